// bqsr_apply.cu -- recalibrated QUAL bytes (replaces the per-read closure of ApplyBQSR, filters/bqsr.go:947-1003).
//
// One warp per read in output order.  Every base with QUAL >= 6 is replaced by
// LUT[read-group covariate][QUAL][cycle][context] -- the byte table bqsr_finalize.cu builds from the hierarchical
// Bayesian estimate (the reference memoises the same function per worker, :973-1000).  The result is written as a
// contiguous QUAL stream in output order (what elp_fetch copies back), so the original QUAL arena stays untouched and
// the read side is a gather through the sorted offsets while the write side streams.
// With lut == nullptr the kernel only materialises the output-order QUAL stream (no BQSR requested).
#include "ctx.h"

namespace {

constexpr int WARPS_PER_BLOCK = 8;

struct ApplyArgs {
    uint64_t n;
    const uint16_t* flag; const int32_t *rg, *lseq; const uint64_t *qual_off, *seq_off, *out_off;
    const uint8_t *seq, *qual; uint8_t* out;
    const int32_t* rg_cov; int n_rg; const uint8_t* cov_exists;
    const uint8_t* lut; int lut_maxcyc, max_cycle;
    uint32_t* err;
};


__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) bqsr_apply_kernel(ApplyArgs A) {
    const unsigned lane = lane_id();
    const uint64_t k = (uint64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (k >= A.n) return;
    const int L = A.lseq[k];
    const uint64_t qoff = A.qual_off[k], ooff = A.out_off[k];
    bool recal = A.lut != nullptr;
    int cov = 0;
    if (recal) {
        const int g = A.rg[k];
        if (g < 0 || g >= A.n_rg) { if (lane == 0) atomicOr(A.err, DERR_NORG); recal = false; }   // readGroupCovariate panics, bqsr.go:38
        else { cov = A.rg_cov[g]; if (!A.cov_exists[cov]) recal = false; }                          // no recalibration, bqsr table empty (:950-953)
    }
    if (!recal) { for (int i = lane; i < L; i += 32) A.out[ooff + i] = A.qual[qoff + i]; return; }
    const uint16_t f = A.flag[k];
    const uint8_t* seqp = A.seq + A.seq_off[k]; const uint8_t* qualp = A.qual + qoff; uint8_t* outp = A.out + ooff;
    // low-quality tails on the FULL read (computeStrandedClippedSeq, bqsr.go:312-331), via ballots
    const int nit = (L + 31) >> 5;
    int leftPos = L, rightPos = -1;
    for (int it = 0; it < nit; it++) {
        const int i = lane + it * 32;
        const unsigned b = __ballot_sync(FULL_MASK, i < L && qualp[i] > 2);
        if (b) { if (leftPos == L) leftPos = it * 32 + __ffs(b) - 1; rightPos = it * 32 + 31 - __clz(b); }
    }
    const int reversed = (f & F_REVERSED) ? 1 : 0, last = (f & F_LAST) ? 1 : 0;
    const int rof = 1 - 2 * last, cf = rof + reversed * (L - 1) * rof, inc = (1 - 2 * reversed) * rof;   // bqsr.go:376-383
    const uint32_t ncyc = 2u * (uint32_t)A.lut_maxcyc + 1u;
    const uint8_t* lut_cov = A.lut + (size_t)cov * 94u * ncyc * 17u;
    uint32_t errbits = 0;
    // branch-free inner loop: base index via popc/ffs on the BAM nibble, the neighbouring base through a warp shuffle
    // (one extra byte load only on the lane at the 32-base boundary), LUT address in 32-bit arithmetic
    const int dirn = reversed ? 1 : -1;                    // context neighbour: previous base in sequencing direction
    for (int it = 0; it < nit; it++) {
        const int i = lane + it * 32;
        const bool in = i < L;
        const int ic = in ? i : L - 1;
        uint32_t q = qualp[ic];
        const uint32_t sb = seqp[ic >> 1];
        const uint32_t nib = (ic & 1) ? (sb & 15u) : (sb >> 4);
        const int bi = (__popc(nib) == 1) ? (__ffs(nib) - 1) : -1;
        // neighbour base index: lane+dirn in this iteration, or the boundary base of the adjacent iteration
        int nbi = __shfl_sync(FULL_MASK, bi, (lane + dirn) & 31);
        const int ni = ic + dirn;
        if (((int)lane + dirn) < 0 || ((int)lane + dirn) > 31 || ni >= L) {
            if (ni >= 0 && ni < L) { const uint32_t nb = seqp[ni >> 1]; const uint32_t nn = (ni & 1) ? (nb & 15u) : (nb >> 4); nbi = (__popc(nn) == 1) ? (__ffs(nn) - 1) : -1; }
            else nbi = -1;
        }
        // both bases inside [leftPos, rightPos] (tails with QUAL <= 2 read as N) and both ACGT
        const int lo_i = reversed ? ic : ni, hi_i = reversed ? ni : ic;
        const bool okc = (bi >= 0) & (nbi >= 0) & (lo_i >= leftPos) & (hi_i <= rightPos) & (ni >= 0) & (ni < L);
        const uint32_t ctx = okc ? (reversed ? (uint32_t)((3 - nbi) | ((3 - bi) << 2)) : (uint32_t)(nbi | (bi << 2))) : 16u;
        const int cyc = cf + ic * inc;
        const bool recal_b = q >= 6;
        const bool badq = q > 93, badc = (cyc > A.max_cycle) | (cyc < -A.max_cycle);
        if (recal_b & in & (badq | badc)) errbits |= badq ? DERR_QUAL_RANGE : DERR_CYCLE;
        const uint32_t qi = badq ? 93u : q;
        const int cyi = badc ? 0 : cyc;
        const uint32_t nq = lut_cov[(qi * ncyc + (uint32_t)(cyi + A.lut_maxcyc)) * 17u + ctx];
        if (recal_b & !badq & !badc) q = nq;
        if (in) outp[i] = (uint8_t)q;
    }
    for (int o = 16; o; o >>= 1) errbits |= __shfl_xor_sync(FULL_MASK, errbits, o);
    if (errbits && lane == 0) atomicOr(A.err, errbits);
}

}  // namespace

int run_apply_kernel(elp_ctx* c, bool with_lut) {
    const uint64_t n = c->n;
    uint64_t total = 0;
    if (n) {
        CUDA_TRY(c, cudaMemcpyAsync(&total, c->s_out_off.p + n, 8, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    }
    CUDA_TRY(c, c->qual_out.reserve(total + 64, c->stream));
    if (n) {
        ApplyArgs A{};
        A.n = n; A.flag = c->s_flag.p; A.rg = c->s_rg.p; A.lseq = c->s_lseq.p; A.qual_off = c->s_qual_off.p; A.seq_off = c->s_seq_off.p; A.out_off = c->s_out_off.p;
        A.seq = c->seq.p; A.qual = c->qual.p; A.out = c->qual_out.p; A.rg_cov = c->d_rg_cov; A.n_rg = c->n_rg; A.cov_exists = c->d_cov_exists;
        A.lut = with_lut ? c->d_lut : nullptr; A.lut_maxcyc = c->lut_maxcyc; A.max_cycle = c->max_cycle; A.err = c->d_err;
        const double bytes = (double)n * (2 + 4 + 4 + 8 + 8 + 8) + (double)c->n_seq + 2.0 * (double)c->n_qual;
        c->begin(with_lut ? "bqsr_apply" : "qual_materialize", bytes);
        bqsr_apply_kernel<<<(unsigned)((n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK), WARPS_PER_BLOCK * 32, 0, c->stream>>>(A);
        c->end(); LAUNCH_CHECK(c);
    }
    int rc = check_device_errors(c);
    if (rc) return rc;
    c->qual_out_valid = true;
    return E_OK;
}

int phase_bqsr_apply(elp_ctx* c) {
    if (!c->sorted) return c->fail(E_STATE, "elp_bqsr_apply called before elp_sort_markdup");
    if (!c->finalized) return c->fail(E_STATE, "elp_bqsr_apply called before elp_bqsr_finalize");
    return run_apply_kernel(c, true);
}
