// bqsr_apply.cu -- recalibrated QUAL bytes (replaces the per-read closure of ApplyBQSR, filters/bqsr.go:947-1003).
//
// Every base with QUAL >= 6 is replaced by LUT[read-group covariate][QUAL][cycle][context] -- the byte table
// bqsr_finalize.cu builds from the hierarchical Bayesian estimate (the reference memoises the same function per worker,
// :973-1000).  The result is written as a contiguous QUAL stream in output order (what elp_fetch copies back); the
// original QUAL arena stays untouched.
//
// Kernel shape (the kernel was instruction-issue bound with one base per lane): 8 lanes per read, 4 reads per warp.
//   1. the group's 8 lanes copy the read's QUAL and SEQ strips into shared memory with aligned 16-byte loads
//      (the strips sit at arbitrary byte offsets of the arenas, so the aligned window around them is staged) and find the
//      low-quality tails (computeStrandedClippedSeq, bqsr.go:312-331) on the fly with SIMD byte compares;
//   2. every lane walks ~L/8 CONSECUTIVE bases out of shared memory: the previous base of the 2-mer context is simply
//      the last one it saw, the cycle advances by +-1, the LUT address by +-17;
//   3. the strip is written back as 16-byte stores aligned on the OUTPUT stream (funnel-shifted out of shared memory);
//      only the first/last partial chunk of a read uses byte stores.
// With lut == nullptr the kernel only materialises the output-order QUAL stream (no BQSR requested).
#include "ctx.h"

namespace {

constexpr int G = 8;                    // lanes per read
constexpr int RPW = 32 / G;             // reads per warp
constexpr int WARPS = 8;
constexpr int MAXL = 512;               // longest read handled (cycles beyond --max-cycle 500 are an error anyway)
constexpr int QSTRIP = MAXL + 48;       // staged QUAL window: <= 15 bytes of lead-in + L + padding, multiple of 16
constexpr int SSTRIP = MAXL / 2 + 48;

struct ApplyArgs {
    uint64_t n;
    const uint16_t* flag; const int32_t *rg, *lseq; const uint64_t *qual_off, *seq_off, *out_off;
    const uint8_t *seq, *qual; uint8_t* out;
    const int32_t* rg_cov; int n_rg; const uint8_t* cov_exists;
    const uint8_t* lut; int lut_maxcyc, max_cycle;
    uint32_t* err;
};

__device__ __forceinline__ uint32_t lds_unaligned32(const uint8_t* p) {   // 4 bytes at any shared-memory address
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    const uint32_t base = a & ~3u, sh = (a & 3u) * 8u;
    uint32_t lo, hi;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(lo) : "r"(base));
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(hi) : "r"(base + 4));
    return __funnelshift_r(lo, hi, sh);
}

__global__ void __launch_bounds__(WARPS * 32) bqsr_apply_kernel(ApplyArgs A) {
    __shared__ __align__(16) uint8_t sm_q[WARPS][RPW][QSTRIP];
    __shared__ __align__(16) uint8_t sm_s[WARPS][RPW][SSTRIP];
    const unsigned lane = lane_id(), w = threadIdx.x >> 5, sub = lane & (G - 1), grp = lane / G;
    const uint64_t k = ((uint64_t)blockIdx.x * WARPS + w) * RPW + grp;
    const bool valid = k < A.n;
    int L = valid ? A.lseq[k] : 0;
    uint32_t errbits = 0;
    if (L > MAXL) { errbits |= DERR_READLEN_LIMIT; L = 0; }
    const uint64_t qoff = valid ? A.qual_off[k] : 0, ooff = valid ? A.out_off[k] : 0;
    bool recal = valid && A.lut != nullptr && L > 0;
    int cov = 0;
    if (recal) {
        const int g = A.rg[k];
        if (g < 0 || g >= A.n_rg) { errbits |= DERR_NORG; recal = false; }                 // readGroupCovariate panics, bqsr.go:38
        else { cov = A.rg_cov[g]; if (!A.cov_exists[cov]) recal = false; }                  // no recalibration, bqsr table empty (:950-953)
    }
    uint8_t* sq = sm_q[w][grp];
    uint8_t* ss = sm_s[w][grp];
    // ---- 1. stage the strips (aligned 16-byte windows) and find the low-quality tails ----
    const uint32_t qsh = (uint32_t)(qoff & 15);
    int leftPos = L, rightPos = -1;
    if (L > 0) {
        const uint64_t qa = qoff & ~15ull;
        const int nqc = (int)((qsh + (uint32_t)L + 15u) >> 4);
        for (int c = sub; c < nqc; c += G) {
            const uint4 v = ld_stream_u4(A.qual + qa + 16ull * c);
            *reinterpret_cast<uint4*>(sq + 16 * c) = v;
            if (recal) {
                const uint32_t wd[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    uint32_t m = __vcmpgtu4(wd[t], 0x02020202u);                           // 0xff per byte with QUAL > 2
                    const int r0 = 16 * c + 4 * t - (int)qsh;                               // read coordinate of byte 0 of this word
                    if (r0 < 0) m &= (r0 <= -4) ? 0u : (0xffffffffu << (8 * (-r0)));
                    if (r0 + 4 > L) m &= (r0 >= L) ? 0u : (0xffffffffu >> (8 * (r0 + 4 - L)));
                    if (m) { leftPos = min(leftPos, r0 + ((__ffs(m) - 1) >> 3)); rightPos = max(rightPos, r0 + ((31 - __clz(m)) >> 3)); }
                }
            }
        }
    }
    uint32_t ssh = 0;
    if (recal) {
        const uint64_t soff = A.seq_off[k];
        ssh = (uint32_t)(soff & 15);
        const uint64_t sa = soff & ~15ull;
        const int nsc = (int)((ssh + (uint32_t)((L + 1) >> 1) + 15u) >> 4);
        for (int c = sub; c < nsc; c += G) *reinterpret_cast<uint4*>(ss + 16 * c) = ld_stream_u4(A.seq + sa + 16ull * c);
    }
#pragma unroll
    for (int o = 1; o < G; o <<= 1) { leftPos = min(leftPos, __shfl_xor_sync(FULL_MASK, leftPos, o)); rightPos = max(rightPos, __shfl_xor_sync(FULL_MASK, rightPos, o)); }
    __syncwarp();
    // ---- 2. every lane recalibrates a run of consecutive bases in place ----
    if (recal) {
        const uint16_t f = A.flag[k];
        const int reversed = (f & F_REVERSED) ? 1 : 0, last = (f & F_LAST) ? 1 : 0;
        const int rof = 1 - 2 * last, cf = rof + reversed * (L - 1) * rof, inc = (1 - 2 * reversed) * rof;   // prepareCycleCovariates, bqsr.go:376-383
        const uint32_t ncyc17 = (2u * (uint32_t)A.lut_maxcyc + 1u) * 17u;
        const uint8_t* lut_cov = A.lut + (size_t)cov * 94u * ncyc17;
        const int C = (L + G - 1) / G, c0 = sub * C, c1 = min(L, c0 + C);
        const uint8_t* sb = ss + ssh;
        uint8_t* qb = sq + qsh;
        auto base_idx = [&](int i) -> int { const uint32_t b = sb[i >> 1]; const uint32_t nb = (i & 1) ? (b & 15u) : (b >> 4); return (__popc(nb) == 1) ? (__ffs(nb) - 1) : -1; };
        // neighbour in sequencing direction: previous read index for forward reads, next for reverse reads; bases outside
        // [leftPos, rightPos] read as N (low-quality tails)
        int cur = (c0 < c1) ? base_idx(c0) : -1;
        int nbr = -1;
        if (!reversed && c0 >= 1 && c0 < c1) nbr = base_idx(c0 - 1);
        for (int i = c0; i < c1; i++) {
            int nxt = -1;
            if (i + 1 < L) nxt = base_idx(i + 1);
            const uint32_t q = qb[i];
            if (q >= 6) {                                                                   // minInterestingQual
                const int cyc = cf + i * inc;
                if (q > 93) errbits |= DERR_QUAL_RANGE;
                else if (cyc > A.max_cycle || cyc < -A.max_cycle) errbits |= DERR_CYCLE;   // checkCycleCovariate :364-369
                else {
                    uint32_t ctx = 16;                                                      // 16 = no context (key -1)
                    if (!reversed) { if (cur >= 0 && nbr >= 0 && i - 1 >= leftPos && i <= rightPos) ctx = (uint32_t)(nbr | (cur << 2)); }
                    else { if (cur >= 0 && nxt >= 0 && i >= leftPos && i + 1 <= rightPos) ctx = (uint32_t)((3 - nxt) | ((3 - cur) << 2)); }
                    qb[i] = lut_cov[q * ncyc17 + (uint32_t)(cyc + A.lut_maxcyc) * 17u + ctx];
                }
            }
            nbr = cur; cur = nxt;
        }
    }
    __syncwarp();
    // ---- 3. write the strip in 16-byte chunks aligned on the output stream ----
    if (L > 0) {
        const uint32_t osh = (uint32_t)(ooff & 15);
        const uint64_t oa = ooff & ~15ull;
        const int noc = (int)((osh + (uint32_t)L + 15u) >> 4);
        const uint8_t* src = sq + qsh;
        for (int c = sub; c < noc; c += G) {
            const int r0 = 16 * c - (int)osh;                                               // read coordinate of the chunk's first byte
            if (r0 >= 0 && r0 + 16 <= L) {
                uint4 v;
                v.x = lds_unaligned32(src + r0); v.y = lds_unaligned32(src + r0 + 4); v.z = lds_unaligned32(src + r0 + 8); v.w = lds_unaligned32(src + r0 + 12);
                *reinterpret_cast<uint4*>(A.out + oa + 16ull * c) = v;
            } else {
                const int lo = max(r0, 0), hi = min(r0 + 16, L);                            // partial chunk shared with the neighbouring read
                for (int r = lo; r < hi; r++) A.out[ooff + r] = src[r];
            }
        }
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
        const uint64_t reads_per_block = (uint64_t)WARPS * RPW;
        c->begin(with_lut ? "bqsr_apply" : "qual_materialize", bytes);
        bqsr_apply_kernel<<<(unsigned)((n + reads_per_block - 1) / reads_per_block), WARPS * 32, 0, c->stream>>>(A);
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
