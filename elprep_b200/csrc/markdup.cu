// markdup.cu -- duplicate marking on the device (replaces filters/mark-duplicates.go:36-445).
//
// The reference classifies reads one by one into three sharded concurrent maps with CAS "best" handles.  Here the
// same equivalence classes are formed by sorting exact packed keys (no hashing of group keys, so no collisions):
//   adapt_kernel        adaptAlignment (:153-156): unclipped 5' position (:79-110) + clamped phred sum (:36-68),
//                       plus a 39-bit (library, QNAME) hash for the mate join and the value ranges that size the keys
//   fragment groups     key (lib, refid, unclipped pos, strand | pair-read-first, score desc) -> radix sort ->
//                       frag_mark_kernel: one thread per group head walks its run (classifyFragment :210-254)
//   mate join           sort by the (lib,QNAME) hash, verify on bytes, pair up in arrival order
//                       (DeleteOrStore on pairFragment :336)
//   pair groups         128-bit key (lib, refid1, refid2, upos1, upos2, rev1, rev2 | score desc) -> radix sort ->
//                       pair_mark_kernel (classifyPair :329-396): both mates of every loser get 0x400
// Results are deterministic; where the reference depends on goroutine scheduling (equal score AND equal QNAME) the
// outcome equals a single goroutine processing reads in arrival order (the later read/pair survives, :231-238,380-386).
#include "ctx.h"
#include <climits>

namespace {

constexpr uint32_t NONE = 0xffffffffu;

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

// ---------------------------------------------------------------- K1 adapt
// 8 lanes per read (4 reads per warp); QUAL strips are read as aligned 16-byte chunks with byte masks.
#ifndef ADAPT_MINB
#define ADAPT_MINB 5
#endif
__global__ void __launch_bounds__(256, ADAPT_MINB) adapt_kernel(uint64_t n, const uint16_t* __restrict__ flag, const int32_t* __restrict__ pos,
                                                     const int32_t* __restrict__ rg, const int32_t* __restrict__ rg_lib, int n_rg,
                                                     const uint64_t* __restrict__ cigar_off, const uint32_t* __restrict__ cigar,
                                                     const uint64_t* __restrict__ qual_off, const uint8_t* __restrict__ qual,
                                                     const uint64_t* __restrict__ qname_off, const uint8_t* __restrict__ qname,
                                                     int32_t* __restrict__ upos_out, int32_t* __restrict__ score_out, uint64_t* __restrict__ qhash_out,
                                                     DeviceRanges* __restrict__ ranges, uint32_t* __restrict__ err) {
    const unsigned lane = lane_id(), sub = lane & 7, grp = lane >> 3;
    const uint64_t gw = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    int32_t pos_max = 0, upos_min = INT_MAX, upos_max = INT_MIN, score_max = 0, lseq_max = 0, pos_min = 0, qname_max = 0;
    uint32_t n_enter = 0, n_pairs = 0, errbits = 0;
    for (uint64_t i0 = gw * 4; i0 < n; i0 += nw * 4) {
        const uint64_t i = i0 + grp;
        const bool valid = i < n;
        uint16_t f = 0; int32_t p = 0; uint64_t q0 = 0, q1 = 0;
        if (valid) { f = flag[i]; p = pos[i]; q0 = qual_off[i]; q1 = qual_off[i + 1]; qname_max = max(qname_max, (int32_t)(qname_off[i + 1] - qname_off[i])); }
        const int32_t len = (int32_t)(q1 - q0);
        pos_max = max(pos_max, p); pos_min = min(pos_min, p); lseq_max = max(lseq_max, len);
        const bool entering = valid && (f & (F_UNMAPPED | F_SECONDARY | F_SUPPLEMENTARY)) == 0;   // mark-duplicates.go:436
        int32_t up = 0, sc = 0; uint64_t qh = 0;
        uint32_t s = 0, bad = 0; uint64_t h = 0;
        const bool true_pair = entering && (f & (F_MULTIPLE | F_NEXTUNMAPPED)) == F_MULTIPLE;   // :182-184
        if (entering) {
            // computeUnclippedPosition (:79-110), serial over the (short) CIGAR on the group's first lane
            if (sub == 0) {
                const uint64_t c0 = cigar_off[i], c1 = cigar_off[i + 1];
                up = p;
                if (c1 > c0) {
                    if (f & F_REVERSED) {
                        int32_t clipped = 1; up--;
                        for (uint64_t k = c1; k-- > c0;) {
                            const uint32_t op = cigar[k]; const uint32_t o = op & 15; const int32_t l = (int32_t)(op >> 4);
                            const int32_t cl = (o == 4 || o == 5), r = (o == 0 || o == 2 || o == 3 || o == 7 || o == 8);
                            clipped *= cl;
                            up += (r | clipped) * l;
                        }
                    } else {
                        for (uint64_t k = c0; k < c1; k++) { const uint32_t op = cigar[k]; const uint32_t o = op & 15; if (!(o == 4 || o == 5)) break; up -= (int32_t)(op >> 4); }
                    }
                }
            }
            // computePhredScore (:57-68): sum of q>=15 over (q & 0x7f) (the table index byte(char<<1) wraps), error if >93
            const uint64_t abase = q0 & ~15ull;
            const uint32_t nch = (uint32_t)((((q1 + 15) & ~15ull) - abase) >> 4);
            for (uint32_t c = sub; c < nch; c += 8) {
                const uint64_t a = abase + 16ull * c;
                const uint4 v = ld_stream_u4(qual + a);
                const int lo = (int)(q0 > a ? q0 - a : 0), hi = (int)((q1 < a + 16 ? q1 : a + 16) - a);
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
                if (lo == 0 && hi == 16) {          // interior chunk: no byte masks
#pragma unroll
                    for (int t = 0; t < 4; t++) {
                        const uint32_t x = w[t] & 0x7f7f7f7fu;
                        bad |= (x + 0x22222222u) & 0x80808080u;
                        const uint32_t ge15 = ((x + 0x71717171u) & 0x80808080u) >> 7;
                        s = __dp4a(x & (ge15 * 0xffu), 0x01010101u, s);
                    }
                    continue;
                }
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    int wlo = lo - 4 * t, whi = hi - 4 * t;
                    wlo = wlo < 0 ? 0 : (wlo > 4 ? 4 : wlo); whi = whi < 0 ? 0 : (whi > 4 ? 4 : whi);
                    if (whi <= wlo) continue;
                    const uint32_t mhi = whi == 4 ? 0xffffffffu : ((1u << (8 * whi)) - 1), mlo = wlo == 0 ? 0u : ((1u << (8 * wlo)) - 1);
                    const uint32_t x = w[t] & 0x7f7f7f7fu & mhi & ~mlo;                 // bytes < 128: the adds below cannot carry across bytes
                    bad |= (x + 0x22222222u) & 0x80808080u;                               // some byte > 93
                    const uint32_t ge15 = ((x + 0x71717171u) & 0x80808080u) >> 7;         // 1 per byte >= 15
                    s = __dp4a(x & (ge15 * 0xffu), 0x01010101u, s);
                }
            }
            // (lib, QNAME) hash for the mate join -- only needs to be a function of the bytes; equality is verified on bytes
            if (true_pair) {
                const uint64_t n0 = qname_off[i], n1 = qname_off[i + 1];
                // four bytes per step and lane; position-salted 32-bit mixes summed in 64 bits (order independent across lanes)
                for (uint64_t k = n0 + 4 * sub; k < n1; k += 32) {
                    uint32_t wv = 0;
#pragma unroll
                    for (int t = 0; t < 4; t++) if (k + t < n1) wv |= (uint32_t)qname[k + t] << (8 * t);
                    uint32_t m = (wv ^ ((uint32_t)(k - n0) * 0x9E3779B1u)) * 0x85EBCA6Bu;
                    m ^= m >> 15; m *= 0xC2B2AE35u; m ^= m >> 13;
                    h += (uint64_t)m * 0x9E3779B97F4A7C15ull;
                }
            }
        }
        // reduce over the 8 lanes of the group (all lanes of the warp take part)
#pragma unroll
        for (int o = 4; o; o >>= 1) { s += __shfl_xor_sync(FULL_MASK, s, o); bad |= __shfl_xor_sync(FULL_MASK, bad, o); h += __shfl_xor_sync(FULL_MASK, h, o); }
        if (entering) {
            sc = (int32_t)s;
            if (bad) errbits |= DERR_QUAL;
            if (sub == 0) {
                const int32_t g = rg[i];
                const int32_t lib = (g >= 0 && g < n_rg) ? rg_lib[g] : -1;
                if (true_pair) { qh = mix64(h + (uint64_t)(uint32_t)(lib + 1) * 0x9E3779B97F4A7C15ull + (qname_off[i + 1] - qname_off[i])); n_pairs++; }
                upos_min = min(upos_min, up); upos_max = max(upos_max, up); score_max = max(score_max, sc); n_enter++;
            }
        }
        if (valid && sub == 0) { upos_out[i] = up; score_out[i] = sc; qhash_out[i] = qh; }
    }
    // block reduction of the ranges, then one atomic per block
    __shared__ int32_t sh_i[8][7];
    __shared__ uint32_t sh_u[8][3];
    for (int o = 16; o; o >>= 1) {
        pos_max = max(pos_max, __shfl_xor_sync(FULL_MASK, pos_max, o)); pos_min = min(pos_min, __shfl_xor_sync(FULL_MASK, pos_min, o));
        upos_min = min(upos_min, __shfl_xor_sync(FULL_MASK, upos_min, o)); upos_max = max(upos_max, __shfl_xor_sync(FULL_MASK, upos_max, o));
        score_max = max(score_max, __shfl_xor_sync(FULL_MASK, score_max, o)); lseq_max = max(lseq_max, __shfl_xor_sync(FULL_MASK, lseq_max, o)); qname_max = max(qname_max, __shfl_xor_sync(FULL_MASK, qname_max, o));
        n_enter += __shfl_xor_sync(FULL_MASK, n_enter, o); n_pairs += __shfl_xor_sync(FULL_MASK, n_pairs, o); errbits |= __shfl_xor_sync(FULL_MASK, errbits, o);
    }
    const unsigned w = threadIdx.x >> 5;
    if (lane == 0) { sh_i[w][0] = pos_max; sh_i[w][1] = upos_min; sh_i[w][2] = upos_max; sh_i[w][3] = score_max; sh_i[w][4] = lseq_max; sh_i[w][5] = pos_min; sh_i[w][6] = qname_max; sh_u[w][0] = n_enter; sh_u[w][1] = n_pairs; sh_u[w][2] = errbits; }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nwb = blockDim.x >> 5;
        for (int k = 1; k < nwb; k++) {
            sh_i[0][0] = max(sh_i[0][0], sh_i[k][0]); sh_i[0][1] = min(sh_i[0][1], sh_i[k][1]); sh_i[0][2] = max(sh_i[0][2], sh_i[k][2]);
            sh_i[0][3] = max(sh_i[0][3], sh_i[k][3]); sh_i[0][4] = max(sh_i[0][4], sh_i[k][4]); sh_i[0][5] = min(sh_i[0][5], sh_i[k][5]); sh_i[0][6] = max(sh_i[0][6], sh_i[k][6]);
            sh_u[0][0] += sh_u[k][0]; sh_u[0][1] += sh_u[k][1]; sh_u[0][2] |= sh_u[k][2];
        }
        atomicMax(&ranges->pos_max, sh_i[0][0]); atomicMin(&ranges->upos_min, sh_i[0][1]); atomicMax(&ranges->upos_max, sh_i[0][2]);
        atomicMax(&ranges->score_max, sh_i[0][3]); atomicMax(&ranges->lseq_max, sh_i[0][4]); atomicMax(&ranges->qname_max, sh_i[0][6]);
        atomicAdd(&ranges->n_entering, sh_u[0][0]); atomicAdd(&ranges->n_true_pairs, sh_u[0][1]);
        if (sh_i[0][5] < 0) sh_u[0][2] |= DERR_QUAL_RANGE;   // negative POS: not representable in the compact sort key
        if (sh_u[0][2]) atomicOr(err, sh_u[0][2]);
    }
}

// ---------------------------------------------------------------- fragment groups
struct FragLayout { int bS, bU, bR, bL; int32_t upos_min, score_max; int key_bits; };

__global__ void __launch_bounds__(256) frag_keys_kernel(uint64_t n, const uint16_t* __restrict__ flag, const int32_t* __restrict__ refid, const int32_t* __restrict__ rg,
                                                         const int32_t* __restrict__ rg_lib, int n_rg, const int32_t* __restrict__ upos, const int32_t* __restrict__ score,
                                                         FragLayout L, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint16_t f = flag[i];
    uint64_t key;
    if ((f & (F_UNMAPPED | F_SECONDARY | F_SUPPLEMENTARY)) == 0) {
        const int32_t g = rg[i];
        const uint64_t lib = (uint64_t)(((g >= 0 && g < n_rg) ? rg_lib[g] : -1) + 1);
        const uint64_t is_frag = ((f & (F_MULTIPLE | F_NEXTUNMAPPED)) != F_MULTIPLE) ? 1 : 0;   // isTrueFragment :177-179
        key = (uint64_t)(uint32_t)(L.score_max - score[i]);
        int sh = L.bS;
        key |= is_frag << sh; sh += 1;
        key |= (uint64_t)((f & F_REVERSED) ? 1 : 0) << sh; sh += 1;
        key |= (uint64_t)(uint32_t)(upos[i] - L.upos_min) << sh; sh += L.bU;
        key |= (uint64_t)(uint32_t)(refid[i] + 1) << sh; sh += L.bR;
        key |= lib << sh;
    } else {
        key = L.key_bits >= 64 ? ~0ull : ((1ull << L.key_bits) - 1);   // sorts after every real key (library field holds an unused value)
    }
    keys[i] = key; vals[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(256) frag_mark_kernel(uint64_t m, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, int bS,
                                                         const uint64_t* __restrict__ qname_off, const uint8_t* __restrict__ qname, uint16_t* __restrict__ flag) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const uint64_t k0 = keys[j], g = k0 >> (bS + 1);
    if (j > 0 && (keys[j - 1] >> (bS + 1)) == g) return;   // not a group head
    const bool head_is_pair = ((k0 >> bS) & 1) == 0;
    if (head_is_pair) {
        // a true-pair read in the group: every true fragment is a duplicate, pair reads are untouched (:225-227,245-252)
        for (uint64_t t = j + 1; t < m; t++) { const uint64_t k = keys[t]; if ((k >> (bS + 1)) != g) break; if ((k >> bS) & 1) atomic_or_u16(flag, vals[t], F_DUPLICATE); }
        return;
    }
    // only fragments: best score first. Winner = max score, then smallest QNAME (:228-243); full ties: the later arrival survives
    uint64_t t1 = j + 1;
    uint64_t win = j;
    while (t1 < m && keys[t1] == k0) {
        const uint32_t a = vals[t1], b = vals[win];
        if (qname_compare(qname, qname_off[a], qname_off[a + 1], qname_off[b], qname_off[b + 1]) <= 0) win = t1;
        t1++;
    }
    for (uint64_t t = j; t < m; t++) {
        if (t >= t1 && (keys[t] >> (bS + 1)) != g) break;
        if (t != win) atomic_or_u16(flag, vals[t], F_DUPLICATE);
    }
}

// ---------------------------------------------------------------- mate join
__global__ void __launch_bounds__(256) join_keys_kernel(uint64_t n, const uint16_t* __restrict__ flag, const uint64_t* __restrict__ qhash,
                                                         uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, uint32_t* __restrict__ mate) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint16_t f = flag[i];
    const bool in = (f & (F_UNMAPPED | F_SECONDARY | F_SUPPLEMENTARY)) == 0 && (f & (F_MULTIPLE | F_NEXTUNMAPPED)) == F_MULTIPLE;
    keys[i] = in ? (qhash[i] & ((1ull << 39) - 1)) : (1ull << 39);
    vals[i] = (uint32_t)i;
    mate[i] = NONE;
}

__global__ void __launch_bounds__(256) join_kernel(uint64_t m, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                    const int32_t* __restrict__ rg, const int32_t* __restrict__ rg_lib, int n_rg,
                                                    const uint64_t* __restrict__ qname_off, const uint8_t* __restrict__ qname, uint32_t* __restrict__ mate) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const uint64_t k0 = keys[j];
    if (j > 0 && keys[j - 1] == k0) return;
    uint64_t e = j + 1;
    while (e < m && keys[e] == k0) e++;
    // arrival order inside the run (stable sort): first unmatched same-(lib,QNAME) read stores, the next one deletes and pairs (:336)
    for (uint64_t a = j; a < e; a++) {
        const uint32_t va = vals[a];
        if (mate[va] != NONE) continue;
        const int32_t ga = rg[va]; const int32_t la = (ga >= 0 && ga < n_rg) ? rg_lib[ga] : -1;
        for (uint64_t b = a + 1; b < e; b++) {
            const uint32_t vb = vals[b];
            if (mate[vb] != NONE) continue;
            const int32_t gb = rg[vb]; const int32_t lb = (gb >= 0 && gb < n_rg) ? rg_lib[gb] : -1;
            if (la != lb) continue;
            if (qname_compare(qname, qname_off[va], qname_off[va + 1], qname_off[vb], qname_off[vb + 1]) != 0) continue;
            mate[va] = vb; mate[vb] = va;
            break;
        }
    }
}

__global__ void __launch_bounds__(256) pair_flag_kernel(uint64_t n, const uint32_t* __restrict__ mate, uint32_t* __restrict__ flags) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t m = mate[i];
    flags[i] = (m != NONE && m < i) ? 1u : 0u;   // the later mate triggers classifyPair
}

struct PairLayout { int bS, bU, bR, bL; int32_t upos_min, score_max; int key_bits; };

__device__ __forceinline__ void put128(uint64_t& lo, uint64_t& hi, int& sh, uint64_t v, int bits) {
    if (bits == 0) return;
    if (sh < 64) { lo |= v << sh; if (sh + bits > 64) hi |= v >> (64 - sh); }
    else hi |= v << (sh - 64);
    sh += bits;
}

__global__ void __launch_bounds__(256) pair_keys_kernel(uint64_t n, const uint32_t* __restrict__ mate, const uint64_t* __restrict__ slot,
                                                         const uint16_t* __restrict__ flag, const int32_t* __restrict__ refid, const int32_t* __restrict__ rg,
                                                         const int32_t* __restrict__ rg_lib, int n_rg, const int32_t* __restrict__ upos, const int32_t* __restrict__ score,
                                                         PairLayout L, uint64_t* __restrict__ keys /*lo,hi interleaved*/, uint32_t* __restrict__ vals,
                                                         uint32_t* __restrict__ pair_a, uint32_t* __restrict__ pair_b) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t m = mate[i];
    if (!(m != NONE && m < i)) return;
    const uint64_t p = slot[i];
    // classifyPair :342-361: aln1 = arriving read (i), aln2 = stored mate (m); swap into (refid, upos, fwd<rev) order
    uint32_t a1 = (uint32_t)i, a2 = m;
    int32_t r1 = refid[a1], r2 = refid[a2], p1 = upos[a1], p2 = upos[a2];
    uint32_t v1 = (flag[a1] & F_REVERSED) ? 1 : 0, v2 = (flag[a2] & F_REVERSED) ? 1 : 0;
    if (r1 > r2 || (r1 == r2 && (p1 > p2 || (p1 == p2 && v1 && !v2)))) {
        uint32_t t = a1; a1 = a2; a2 = t; int32_t ti = r1; r1 = r2; r2 = ti; ti = p1; p1 = p2; p2 = ti; t = v1; v1 = v2; v2 = t;
    }
    const int32_t g = rg[a1];
    const uint64_t lib = (uint64_t)(((g >= 0 && g < n_rg) ? rg_lib[g] : -1) + 1);
    uint64_t lo = 0, hi = 0; int sh = 0;
    put128(lo, hi, sh, (uint64_t)(uint32_t)(L.score_max - (score[a1] + score[a2])), L.bS);
    put128(lo, hi, sh, (uint64_t)(uint32_t)(p2 - L.upos_min), L.bU);
    put128(lo, hi, sh, (uint64_t)(uint32_t)(p1 - L.upos_min), L.bU);
    put128(lo, hi, sh, v2, 1); put128(lo, hi, sh, v1, 1);
    put128(lo, hi, sh, (uint64_t)(uint32_t)(r2 + 1), L.bR); put128(lo, hi, sh, (uint64_t)(uint32_t)(r1 + 1), L.bR);
    put128(lo, hi, sh, lib, L.bL);
    keys[2 * p] = lo; keys[2 * p + 1] = hi; vals[p] = (uint32_t)p;
    pair_a[p] = a1; pair_b[p] = a2;
}

__device__ __forceinline__ void shr128(uint64_t lo, uint64_t hi, int s, uint64_t& olo, uint64_t& ohi) {
    if (s == 0) { olo = lo; ohi = hi; }
    else if (s < 64) { olo = (lo >> s) | (hi << (64 - s)); ohi = hi >> s; }
    else { olo = hi >> (s - 64); ohi = 0; }
}

__global__ void __launch_bounds__(256) pair_mark_kernel(uint64_t m, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, int bS,
                                                         const uint32_t* __restrict__ pair_a, const uint32_t* __restrict__ pair_b,
                                                         const uint64_t* __restrict__ qname_off, const uint8_t* __restrict__ qname, uint16_t* __restrict__ flag) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const uint64_t klo = keys[2 * j], khi = keys[2 * j + 1];
    uint64_t glo, ghi; shr128(klo, khi, bS, glo, ghi);
    if (j > 0) { uint64_t a, b; shr128(keys[2 * j - 2], keys[2 * j - 1], bS, a, b); if (a == glo && b == ghi) return; }
    // group head = best score. Winner = max score, then smallest aln1.QNAME (:375-395); full ties: the later pair survives
    uint64_t t1 = j + 1, win = j;
    while (t1 < m && keys[2 * t1] == klo && keys[2 * t1 + 1] == khi) {
        const uint32_t a = pair_a[vals[t1]], b = pair_a[vals[win]];
        if (qname_compare(qname, qname_off[a], qname_off[a + 1], qname_off[b], qname_off[b + 1]) <= 0) win = t1;
        t1++;
    }
    for (uint64_t t = j; t < m; t++) {
        if (t >= t1) { uint64_t a, b; shr128(keys[2 * t], keys[2 * t + 1], bS, a, b); if (a != glo || b != ghi) break; }
        if (t != win) { const uint32_t p = vals[t]; atomic_or_u16(flag, pair_a[p], F_DUPLICATE); atomic_or_u16(flag, pair_b[p], F_DUPLICATE); }
    }
}

inline unsigned nblk(uint64_t n, int t) { return (unsigned)((n + t - 1) / t); }

}  // namespace

int check_device_errors(elp_ctx* c) {
    uint32_t e = 0;
    CUDA_TRY(c, cudaMemcpyAsync(&e, c->d_err, 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (!e) return E_OK;
    CUDA_TRY(c, cudaMemsetAsync(c->d_err, 0, 4, c->stream));
    if (e & DERR_QUAL) return c->fail(E_QUAL, "Invalid QUAL character in a read entering duplicate marking");
    if (e & DERR_NORG) return c->fail(E_NORG, "Error: BQSR requires input with read groups. An alignment has no read group. Please fix input, e.g. rerun with the --replace-read-group option.");
    if (e & DERR_CYCLE) return c->fail(E_CYCLE, "cycle value exceeds maximum cycle value");
    if (e & DERR_CLIP) return c->fail(E_CLIP, "reference coordinate matches a non-existing base in read");
    if (e & DERR_REFEND) return c->fail(E_REFEND, "a recalibrated read extends past the end of its reference sequence");
    if (e & DERR_CIGAR_LIMIT) return c->fail(E_LIMIT, "BQSR: CIGAR with more operations than the device kernel supports");
    if (e & DERR_QUAL_RANGE) return c->fail(E_LIMIT, "value outside the supported range (negative POS, or QUAL > 93 in a recalibrated read)");
    if (e & DERR_TILE) return c->fail(E_TILE, "strconv.ParseInt: parsing a tile/x/y field of a QNAME: invalid syntax or value out of range");
    if (e & DERR_TILE_RANGE) return c->fail(E_LIMIT, "optical duplicates: tile/x/y value outside int32");
    if (e & DERR_BAM_CG) return c->fail(E_LIMIT, "BAM record uses the CG:B long-CIGAR convention, which the device parser does not handle");
    if (e & DERR_BAM_RG) return c->fail(E_BAM, "BAM record with an RG:Z value that is not an @RG ID of the header");
    if (e & DERR_BAM) return c->fail(E_BAM, "malformed BAM alignment record (field lengths and block_size do not add up)");
    if (e & DERR_READLEN_LIMIT) return c->fail(E_LIMIT, "BQSR: read longer than the device kernel supports");
    return c->fail(E_CUDA, "unknown device error word 0x%x", e);
}

int phase_adapt(elp_ctx* c) {
    if (c->adapted) return E_OK;
    const uint64_t n = c->n;
    CUDA_TRY(c, c->upos.reserve(n + 1, c->stream));
    CUDA_TRY(c, c->score.reserve(n + 1, c->stream));
    CUDA_TRY(c, c->qhash.reserve(n + 1, c->stream));
    DeviceRanges init{}; init.pos_max = 0; init.upos_min = INT_MAX; init.upos_max = INT_MIN; init.score_max = 0; init.lseq_max = 0;
    CUDA_TRY(c, cudaMemcpyAsync(c->d_ranges, &init, sizeof init, cudaMemcpyHostToDevice, c->stream));
    if (n) {
        int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
        uint64_t want = (n / 4 * 32 + 255) / 256;
        unsigned grid = (unsigned)std::min<uint64_t>(std::max<uint64_t>(want, 1), (uint64_t)sms * 8);
        double bytes = (double)n * (2 + 4 + 4 + 16 + 16 + 8 + 4 + 4 + 8) + (double)c->n_qual + (double)c->n_cigar * 4 + (double)c->n_qname;
        c->begin("adapt", bytes);
        adapt_kernel<<<grid, 256, 0, c->stream>>>(n, c->flag.p, c->pos.p, c->rg.p, c->d_rg_lib, c->n_rg, c->cigar_off.p, c->cigar.p, c->qual_off.p, c->qual.p,
                                                  c->qname_off.p, c->qname.p, c->upos.p, c->score.p, c->qhash.p, c->d_ranges, c->d_err);
        c->end(); LAUNCH_CHECK(c);
    }
    CUDA_TRY(c, cudaMemcpyAsync(&c->h_ranges, c->d_ranges, sizeof(DeviceRanges), cudaMemcpyDeviceToHost, c->stream));
    int rc = check_device_errors(c);   // synchronizes
    if (rc) return rc;
    if (c->h_ranges.n_entering == 0) { c->h_ranges.upos_min = 0; c->h_ranges.upos_max = 0; }
    c->adapted = true;
    return E_OK;
}

int phase_markdup(elp_ctx* c, bool optical) {
    int rc = phase_adapt(c);
    if (rc) return rc;
    const uint64_t n = c->n;
    if (n == 0 || c->h_ranges.n_entering == 0) return optical ? phase_optical(c, 0, nullptr, nullptr, 0) : E_OK;
    const DeviceRanges& R = c->h_ranges;
    CUDA_TRY(c, c->keys_a.reserve(2 * n + 4, c->stream)); CUDA_TRY(c, c->keys_b.reserve(2 * n + 4, c->stream));
    CUDA_TRY(c, c->vals_a.reserve(n + 4, c->stream)); CUDA_TRY(c, c->vals_b.reserve(n + 4, c->stream));
    const int bR = bits_for((uint64_t)c->n_contigs), bL = bits_for((uint64_t)c->n_lib + 1);
    const int bU = bits_for((uint64_t)((int64_t)R.upos_max - (int64_t)R.upos_min));
    // ---- fragments (classifyFragment) ----
    {
        FragLayout L{}; L.bS = bits_for((uint64_t)R.score_max); L.bU = bU; L.bR = bR; L.bL = bL; L.upos_min = R.upos_min; L.score_max = R.score_max;
        L.key_bits = L.bS + 2 + L.bU + L.bR + L.bL;
        if (L.key_bits > 64) return c->fail(E_LIMIT, "fragment signature needs %d bits (>64): too many contigs/libraries for the packed key", L.key_bits);
        c->begin("frag_keys", (double)n * (2 + 4 + 4 + 4 + 4 + 8 + 4));
        frag_keys_kernel<<<nblk(n, 256), 256, 0, c->stream>>>(n, c->flag.p, c->refid.p, c->rg.p, c->d_rg_lib, c->n_rg, c->upos.p, c->score.p, L, c->keys_a.p, c->vals_a.p);
        c->end(); LAUNCH_CHECK(c);
        bool in_b = false;
        rc = radix_sort_u64(c, c->keys_a.p, c->keys_b.p, c->vals_a.p, c->vals_b.p, n, L.key_bits, &in_b, "u64");
        if (rc) return rc;
        const uint64_t m = R.n_entering;
        c->begin("frag_mark", (double)m * 12);
        frag_mark_kernel<<<nblk(m, 256), 256, 0, c->stream>>>(m, in_b ? c->keys_b.p : c->keys_a.p, in_b ? c->vals_b.p : c->vals_a.p, L.bS, c->qname_off.p, c->qname.p, c->flag.p);
        c->end(); LAUNCH_CHECK(c);
    }
    // ---- pairs (classifyPair) ----
    if (R.n_true_pairs >= 2) {
        CUDA_TRY(c, c->mate.reserve(n + 4, c->stream));
        c->begin("join_keys", (double)n * (2 + 8 + 8 + 4 + 4));
        join_keys_kernel<<<nblk(n, 256), 256, 0, c->stream>>>(n, c->flag.p, c->qhash.p, c->keys_a.p, c->vals_a.p, c->mate.p);
        c->end(); LAUNCH_CHECK(c);
        bool in_b = false;
        rc = radix_sort_u64(c, c->keys_a.p, c->keys_b.p, c->vals_a.p, c->vals_b.p, n, 40, &in_b, "u64");
        if (rc) return rc;
        const uint64_t m = R.n_true_pairs;
        c->begin("join", (double)m * 12);
        join_kernel<<<nblk(m, 256), 256, 0, c->stream>>>(m, in_b ? c->keys_b.p : c->keys_a.p, in_b ? c->vals_b.p : c->vals_a.p, c->rg.p, c->d_rg_lib, c->n_rg, c->qname_off.p, c->qname.p, c->mate.p);
        c->end(); LAUNCH_CHECK(c);
        // deterministic pair list, ordered by the arrival of the later mate
        CUDA_TRY(c, c->scan_tmp.reserve(n + 4, c->stream));
        uint64_t* slot = c->keys_b.p;   // n+1 u64, free at this point
        c->begin("pair_flag", (double)n * 8);
        pair_flag_kernel<<<nblk(n, 256), 256, 0, c->stream>>>(n, c->mate.p, c->scan_tmp.p);
        c->end(); LAUNCH_CHECK(c);
        rc = exclusive_scan_u32_to_u64(c, c->scan_tmp.p, slot, n);
        if (rc) return rc;
        uint64_t npairs = 0;
        CUDA_TRY(c, cudaMemcpyAsync(&npairs, slot + n, 8, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        if (npairs >= (optical ? 1u : 2u)) {
            CUDA_TRY(c, c->pair_a.reserve(npairs + 4, c->stream)); CUDA_TRY(c, c->pair_b.reserve(npairs + 4, c->stream));
            PairLayout L{}; L.bS = bits_for((uint64_t)R.score_max * 2); L.bU = bU; L.bR = bR; L.bL = bL; L.upos_min = R.upos_min; L.score_max = R.score_max * 2;
            L.key_bits = L.bS + 2 * L.bU + 2 + 2 * L.bR + L.bL;
            if (L.key_bits > 128) return c->fail(E_LIMIT, "pair signature needs %d bits (>128)", L.key_bits);
            // 128-bit keys live in keys_a as (lo,hi) pairs; slot[] occupies keys_b, so sort into a separate buffer
            CUDA_TRY(c, c->bytes_tmp.reserve((size_t)npairs * 16 + 64, c->stream));
            uint64_t* kb2 = reinterpret_cast<uint64_t*>(c->bytes_tmp.p);
            c->begin("pair_keys", (double)n * 12 + (double)npairs * (2 * 18 + 16 + 12));
            pair_keys_kernel<<<nblk(n, 256), 256, 0, c->stream>>>(n, c->mate.p, slot, c->flag.p, c->refid.p, c->rg.p, c->d_rg_lib, c->n_rg, c->upos.p, c->score.p, L,
                                                                  c->keys_a.p, c->vals_a.p, c->pair_a.p, c->pair_b.p);
            c->end(); LAUNCH_CHECK(c);
            rc = radix_sort_u128(c, c->keys_a.p, kb2, c->vals_a.p, c->vals_b.p, npairs, L.key_bits, &in_b, "u128");
            if (rc) return rc;
            c->begin("pair_mark", (double)npairs * 20);
            pair_mark_kernel<<<nblk(npairs, 256), 256, 0, c->stream>>>(npairs, in_b ? kb2 : c->keys_a.p, in_b ? c->vals_b.p : c->vals_a.p, L.bS, c->pair_a.p, c->pair_b.p,
                                                                      c->qname_off.p, c->qname.p, c->flag.p);
            c->end(); LAUNCH_CHECK(c);
            if (optical) return phase_optical(c, npairs, in_b ? kb2 : c->keys_a.p, in_b ? c->vals_b.p : c->vals_a.p, L.bS);
        }
    }
    return optical ? phase_optical(c, 0, nullptr, nullptr, 0) : E_OK;
}
