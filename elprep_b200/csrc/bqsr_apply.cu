// bqsr_apply.cu -- recalibrated QUAL bytes (replaces the per-read closure of ApplyBQSR, filters/bqsr.go:947-1003).
//
// Every base with QUAL >= 6 is replaced by LUT[read-group covariate][QUAL][cycle][context] -- the byte table
// bqsr_finalize.cu builds from the hierarchical Bayesian estimate (the reference memoises the same function per worker,
// :973-1000).  The result is written as a contiguous QUAL stream in output order (what elp_fetch copies back); the
// original QUAL arena stays untouched.
//
// Kernel shape (same decomposition as the gather's chunk kernel, bqsr_gather.cu): a lane owns 16 consecutive bases of a
// read, a warp takes 32 / lanes_per_read reads.
//   1. 16 QUAL bytes and 16 SEQ nibbles come out of aligned 128-bit loads (funnel-shifted to the read's byte offset);
//      base codes, the 2-mer context of every base and the low-quality tails (computeStrandedClippedSeq, bqsr.go:312-331)
//      are computed word-parallel; the tails are reduced across the lanes of the read with member-mask reductions;
//   2. per base: one byte load from the LUT (L1/L2 resident) and a byte insert; cycle advances by +-1, LUT address by +-17;
//   3. the lanes park their 16 result bytes in shared memory and the read's strip is written as 16-byte stores aligned
//      on the OUTPUT stream (funnel-shifted out of shared memory); only the partial chunks at the ends use byte stores.
// With lut == nullptr the kernel only materialises the output-order QUAL stream (no BQSR requested).
#include "ctx.h"
#include "bqsr_simd.cuh"
#include <algorithm>

namespace {

constexpr int WARPS = 8;

// nibble flags for chunk index >= k and <= k - 1 (filled by run_apply_kernel)
__constant__ unsigned long long c_ge[CHUNK + 1], c_le[CHUNK + 1];
__device__ __forceinline__ unsigned long long range16(int lo, int hi) { return c_ge[min(max(lo, 0), CHUNK)] & c_le[min(max(hi + 1, 0), CHUNK)]; }

struct ApplyArgs {
    uint64_t n;
    const uint16_t* flag; const int32_t *rg, *lseq; const uint64_t *qual_off, *seq_off, *out_off;
    const uint8_t *seq, *qual; uint8_t* out;
    const int32_t* rg_cov; int n_rg; const uint8_t* cov_exists;
    const uint8_t* lut; int lut_maxcyc, max_cycle;
    int lanes_per_read;
    uint32_t* err;
};

#ifndef APPLY_MINB
#define APPLY_MINB 8
#endif
__global__ void __launch_bounds__(WARPS * 32, APPLY_MINB) bqsr_apply_kernel(ApplyArgs A) {
    // image of the block's slice of the output stream, placed with the slice's 16-byte phase: the reads of a block are
    // consecutive in output order, so the whole slice leaves as aligned 128-bit stores (two partial chunks per BLOCK)
    __shared__ __align__(16) uint8_t sm_img[WARPS * 32 * CHUNK + 32];
    const unsigned lane = lane_id(), w = threadIdx.x >> 5;
    const int lpr = A.lanes_per_read, rpw = 32 / lpr;
    const int r = (int)lane / lpr, c = (int)lane - r * lpr;
    const bool lane_used = r < rpw;
    const uint64_t k = ((uint64_t)blockIdx.x * WARPS + w) * (uint64_t)rpw + (uint64_t)r;
    const bool valid = lane_used && k < A.n;
    int L = valid ? A.lseq[k] : 0;
    uint32_t errbits = 0;
    if (L > CHUNK * lpr) { errbits |= DERR_READLEN_LIMIT; L = 0; }
    const uint64_t qoff = valid ? A.qual_off[k] : 0, ooff = valid ? A.out_off[k] : 0;
    const uint64_t kb0 = (uint64_t)blockIdx.x * WARPS * (uint64_t)rpw, kb1 = min(A.n, kb0 + (uint64_t)(WARPS * rpw));
    const uint64_t o0 = A.out_off[kb0], o1 = A.out_off[kb1];               // the block's slice [o0, o1) of the output stream
    const uint32_t phase = (uint32_t)(o0 & 15);
    bool recal = valid && A.lut != nullptr && L > 0;
    int cov = 0;
    if (recal) {
        const int g = A.rg[k];
        if (g < 0 || g >= A.n_rg) { errbits |= DERR_NORG; recal = false; }                 // readGroupCovariate panics, bqsr.go:38
        else { cov = A.rg_cov[g]; if (!A.cov_exists[cov]) recal = false; }                  // no recalibration, bqsr table empty (:950-953)
    }
    const int i0 = c * CHUNK, nb = min(max(L - i0, 0), CHUNK);
    // ---- 1. loads and word-parallel covariates ----
    uint32_t Q[4] = {0, 0, 0, 0};
    unsigned long long C = 0;
    if (nb > 0) {
        load16_unaligned(A.qual + qoff + (uint64_t)i0, Q);
        if (A.lut != nullptr) {   // (not `recal`: the read-group look-ups above overlap with these loads)
            const unsigned long long nibs = load16_nibbles_bam(A.seq, A.seq_off[k] * 2 + (uint64_t)i0);
            C = (unsigned long long)codes_of((uint32_t)nibs) | ((unsigned long long)codes_of((uint32_t)(nibs >> 32)) << 32);
        }
    }
    if (nb < CHUNK) { const unsigned long long inlen = range16(0, nb - 1); C = (C & (inlen * 15ull)) | ((ONES & ~inlen) << 3); }   // codes past the read end: 8
    int first, last;
    qual_gt2_span(Q, nb, i0, first, last);
    const unsigned gmask = lane_used ? ((lpr == 32 ? 0xffffffffu : ((1u << lpr) - 1u)) << (r * lpr)) : (1u << lane);
    const int leftPos = __reduce_min_sync(gmask, first), rightPos = __reduce_max_sync(gmask, last);
    const uint32_t c_hi = (uint32_t)(C >> 32), c_lo = (uint32_t)C;
    uint32_t edge_prev = __shfl_up_sync(FULL_MASK, c_hi, 1) >> 28, edge_next = __shfl_down_sync(FULL_MASK, c_lo, 1) & 15u;
    if (c == 0) edge_prev = 8;
    if (c == lpr - 1 || lane == 31) edge_next = 8;
    // ---- 2. LUT lookups ----
    if (recal && nb > 0) {
        const uint16_t f = A.flag[k];
        const bool rev = f & F_REVERSED;
        const int lastf = (f & F_LAST) ? 1 : 0;
        const int rof = 1 - 2 * lastf, inc = rev ? -rof : rof, cf = rof + (rev ? (L - 1) * rof : 0);   // prepareCycleCovariates, bqsr.go:376-383 (full read length)
        const unsigned long long Pn = rev ? ((C >> 4) | ((unsigned long long)edge_next << 60)) : ((C << 4) | edge_prev);
        const unsigned long long M3 = 0x3333333333333333ull, xr = rev ? M3 : 0ull;
        const unsigned long long ctxw = ((Pn ^ xr) & M3) | (((C ^ xr) & M3) << 2);          // key>>4 = prev | cur<<2, complemented for reverse reads
        const int wlo = rev ? leftPos : leftPos + 1, whi = rev ? rightPos - 1 : rightPos;      // low-quality tails read as N
        const unsigned long long okc = ~((Pn | C) >> 3) & ONES & range16(wlo - i0, whi - i0);
        const uint32_t ncyc17 = (2u * (uint32_t)A.lut_maxcyc + 1u) * 17u;
        const uint8_t* lut_cov = A.lut + (size_t)cov * 94u * ncyc17;
        const uint32_t okc_w[2] = {(uint32_t)okc, (uint32_t)(okc >> 32)}, ctx_w[2] = {(uint32_t)ctxw, (uint32_t)(ctxw >> 32)};
        const int cyc_a = cf + i0 * inc, cyc_b = cyc_a + (nb - 1) * inc;                    // cycles of the chunk's first / last base
        // bytes past the read end must not look like bases (they belong to the next read of the arena and are never written back)
#pragma unroll
        for (int wq = 0; wq < 4; wq++) { const int keep = nb - 4 * wq; if (keep < 4) Q[wq] = keep <= 0 ? 0u : (Q[wq] & (0xffffffffu >> (8 * (4 - keep)))); }
        if (max(cyc_a, cyc_b) > A.max_cycle || min(cyc_a, cyc_b) < -A.max_cycle) {
            // checkCycleCovariate (:364-369) fails somewhere in this chunk: an error if one of those bases is recalibrated
            int cyc = cyc_a;
#pragma unroll
            for (int j = 0; j < CHUNK; j++) {
                const uint32_t q = (Q[j >> 2] >> (8 * (j & 3))) & 0xffu;
                if (q >= 6) errbits |= q > 93 ? DERR_QUAL_RANGE : ((cyc > A.max_cycle || cyc < -A.max_cycle) ? DERR_CYCLE : 0u);
                cyc += inc;
            }
        } else {
            uint32_t idx = (uint32_t)(cyc_a + A.lut_maxcyc) * 17u;
            const uint32_t step = (uint32_t)(inc * 17);
            uint32_t over = 0;
#pragma unroll
            for (int j = 0; j < CHUNK; j++) {
                const uint32_t q = (Q[j >> 2] >> (8 * (j & 3))) & 0xffu;
                const uint32_t nib = (ctx_w[j >> 3] >> (4 * (j & 7))) & 15u;
                const uint32_t ctx = ((okc_w[j >> 3] >> (4 * (j & 7))) & 1u) ? nib : 16u;   // 16 = no context (key -1)
                over |= q;
                if (q - 6u <= 87u) {                                                        // minInterestingQual <= q <= 93
                    const uint32_t v = __ldg(lut_cov + q * ncyc17 + idx + ctx);
                    Q[j >> 2] = __byte_perm(Q[j >> 2], v, (j & 3) == 0 ? 0x3214 : ((j & 3) == 1 ? 0x3240 : ((j & 3) == 2 ? 0x3410 : 0x4210)));
                }
                idx += step;
            }
            // a QUAL above 93 (any byte with bit 7, or 94..127) is an error for a recalibrated base
            if (over & 0x80u) errbits |= DERR_QUAL_RANGE;
            else if (over >= 94u) {
#pragma unroll
                for (int wq = 0; wq < 4; wq++) { const uint32_t v = Q[wq]; if ((((v & 0x7f7f7f7fu) + 0x22222222u) | v) & 0x80808080u) errbits |= DERR_QUAL_RANGE; }
            }
        }
    }
    // ---- 3. park the chunk in the block image, then write the image in 16-byte chunks aligned on the output stream ----
    if (nb > 0) {
        uint8_t* dst = sm_img + phase + (uint32_t)(ooff - o0) + (uint32_t)i0;
#pragma unroll
        for (int j = 0; j < CHUNK; j++) if (j < nb) dst[j] = (uint8_t)(Q[j >> 2] >> (8 * (j & 3)));
    }
    __syncthreads();
    {
        const uint32_t total = (uint32_t)(o1 - o0), end = phase + total;           // image bytes [phase, end)
        uint8_t* gbase = A.out + (o0 - phase);                                      // 16-byte aligned
        for (uint32_t m = threadIdx.x; m * 16u < end; m += blockDim.x) {
            const uint32_t b0 = m * 16u;
            if (b0 >= phase && b0 + 16u <= end) *reinterpret_cast<uint4*>(gbase + b0) = *reinterpret_cast<const uint4*>(sm_img + b0);
            else { const uint32_t lo = max(b0, phase), hi = min(b0 + 16u, end); for (uint32_t t = lo; t < hi; t++) gbase[t] = sm_img[t]; }   // shared with a neighbouring block
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
    c->qual_out_total = total;
    if (n) {
        static bool tables_set[64] = {false};
        if (!tables_set[c->device & 63]) {
            unsigned long long ge[CHUNK + 1], le[CHUNK + 1];
            for (int b = 0; b <= CHUNK; b++) { ge[b] = b == CHUNK ? 0ull : (ONES << (4 * b)); le[b] = b == 0 ? 0ull : (ONES >> (4 * (CHUNK - b))); }
            CUDA_TRY(c, cudaMemcpyToSymbol(c_ge, ge, sizeof ge)); CUDA_TRY(c, cudaMemcpyToSymbol(c_le, le, sizeof le));
            tables_set[c->device & 63] = true;
        }
        ApplyArgs A{};
        A.n = n; A.flag = c->s_flag.p; A.rg = c->s_rg.p; A.lseq = c->s_lseq.p; A.qual_off = c->s_qual_off.p; A.seq_off = c->s_seq_off.p; A.out_off = c->s_out_off.p;
        A.seq = c->seq.p; A.qual = c->qual.p; A.out = c->qual_out.p; A.rg_cov = c->d_rg_cov; A.n_rg = c->n_rg; A.cov_exists = c->d_cov_exists;
        A.lut = with_lut ? c->d_lut : nullptr; A.lut_maxcyc = c->lut_maxcyc; A.max_cycle = c->max_cycle; A.err = c->d_err;
        const double bytes = (double)n * (2 + 4 + 4 + 8 + 8 + 8) + (double)c->n_seq + 2.0 * (double)c->n_qual;
        A.lanes_per_read = std::min(32, std::max(1, (c->h_ranges.lseq_max + CHUNK - 1) / CHUNK));
        const uint64_t reads_per_block = (uint64_t)WARPS * (32 / A.lanes_per_read);
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
    // the look-up table must cover every cycle of the reads now loaded (it was sized by elp_bqsr_finalize, possibly before they arrived)
    int rc = phase_adapt(c);
    if (rc) return rc;
    const int need = std::max(1, std::min(c->max_cycle, std::max(c->h_ranges.lseq_max, 1)));
    if (c->lut_maxcyc < need) { rc = build_apply_lut(c, need); if (rc) return rc; }
    return run_apply_kernel(c, true);
}
