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

__device__ __forceinline__ int nib_at(const uint8_t* seq, uint64_t soff, int i) { const uint8_t b = seq[soff + (uint64_t)(i >> 1)]; return (i & 1) ? (b & 15) : (b >> 4); }
__device__ __forceinline__ int nib_index(int nib) { return nib == 1 ? 0 : nib == 2 ? 1 : nib == 4 ? 2 : nib == 8 ? 3 : -1; }

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
    const uint64_t soff = A.seq_off[k];
    // low-quality tails on the FULL read (computeStrandedClippedSeq, bqsr.go:312-331)
    int leftPos = L, rightPos = -1;
    for (int i = lane; i < L; i += 32) if (A.qual[qoff + i] > 2) { leftPos = min(leftPos, i); rightPos = max(rightPos, i); }
    for (int o = 16; o; o >>= 1) { leftPos = min(leftPos, __shfl_xor_sync(FULL_MASK, leftPos, o)); rightPos = max(rightPos, __shfl_xor_sync(FULL_MASK, rightPos, o)); }
    const int reversed = (f & F_REVERSED) ? 1 : 0, last = (f & F_LAST) ? 1 : 0;
    const int rof = 1 - 2 * last, cf = rof + reversed * (L - 1) * rof, inc = (1 - 2 * reversed) * rof;   // bqsr.go:376-383
    const int ncyc = 2 * A.lut_maxcyc + 1;
    uint32_t errbits = 0;
    for (int i = lane; i < L; i += 32) {
        uint8_t q = A.qual[qoff + i];
        if (q >= 6) {                                              // minInterestingQual
            if (q > 93) errbits |= DERR_QUAL_RANGE;
            else {
                const int cyc = cf + i * inc;
                if (cyc > A.max_cycle || cyc < -A.max_cycle) errbits |= DERR_CYCLE;   // checkCycleCovariate :364-369
                else {
                    int ctx = 16;                                  // 16 = no context (key -1)
                    const int bi = nib_index(nib_at(A.seq, soff, i));
                    if (bi >= 0) {
                        if (!reversed) { if (i >= 1 && i - 1 >= leftPos && i <= rightPos) { const int pb = nib_index(nib_at(A.seq, soff, i - 1)); if (pb >= 0) ctx = pb | (bi << 2); } }
                        else { if (i + 1 <= L - 1 && i >= leftPos && i + 1 <= rightPos) { const int nb = nib_index(nib_at(A.seq, soff, i + 1)); if (nb >= 0) ctx = (3 - nb) | ((3 - bi) << 2); } }
                    }
                    q = A.lut[(((size_t)cov * 94 + q) * ncyc + (size_t)(cyc + A.lut_maxcyc)) * 17 + ctx];
                }
            }
        }
        A.out[ooff + i] = q;
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
