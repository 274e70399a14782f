// bqsr_simd.cuh -- word-parallel helpers shared by the BQSR chunk kernels (gather and apply): a lane owns 16 consecutive
// bases of a read and keeps them in registers as 4 QUAL words and 64-bit words of 4-bit fields (one nibble per base).
#pragma once
#include <cstdint>
#include "common.cuh"

namespace {

constexpr int CHUNK = 16;       // bases per lane
constexpr unsigned long long ONES = 0x1111111111111111ull;

__device__ __forceinline__ void load16_unaligned(const uint8_t* p, uint32_t (&o)[4]) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint4* base = reinterpret_cast<const uint4*>(a & ~(uintptr_t)15);
    const uint4 A0 = __ldg(base), A1 = __ldg(base + 1);
    const uint32_t off = (uint32_t)a & 15u;
    uint32_t w0 = A0.x, w1 = A0.y, w2 = A0.z, w3 = A0.w, w4 = A1.x, w5 = A1.y;
    if (off & 8) { w0 = w2; w1 = w3; w2 = w4; w3 = w5; w4 = A1.z; w5 = A1.w; }
    if (off & 4) { w0 = w1; w1 = w2; w2 = w3; w3 = w4; w4 = w5; }
    const uint32_t sh = (off & 3) * 8;
    o[0] = __funnelshift_r(w0, w1, sh); o[1] = __funnelshift_r(w1, w2, sh); o[2] = __funnelshift_r(w2, w3, sh); o[3] = __funnelshift_r(w3, w4, sh);
}
// 16 nibbles starting at nibble index `nidx` of a stream whose nibble 2b is the LOW nibble of byte b
__device__ __forceinline__ unsigned long long load16_nibbles_le(const uint8_t* stream, uint64_t nidx) {
    uint32_t o[4]; load16_unaligned(stream + (nidx >> 1), o);
    const uint32_t sh = (uint32_t)(nidx & 1) * 4;
    return (unsigned long long)__funnelshift_r(o[0], o[1], sh) | ((unsigned long long)__funnelshift_r(o[1], o[2], sh) << 32);
}
// the same for BAM SEQ (nibble 2b is the HIGH nibble of byte b)
__device__ __forceinline__ unsigned long long load16_nibbles_bam(const uint8_t* stream, uint64_t nidx) {
    uint32_t o[4]; load16_unaligned(stream + (nidx >> 1), o);
#pragma unroll
    for (int k = 0; k < 3; k++) o[k] = ((o[k] & 0x0f0f0f0fu) << 4) | ((o[k] >> 4) & 0x0f0f0f0fu);
    const uint32_t sh = (uint32_t)(nidx & 1) * 4;
    return (unsigned long long)__funnelshift_r(o[0], o[1], sh) | ((unsigned long long)__funnelshift_r(o[1], o[2], sh) << 32);
}
// 8 BAM nibbles -> 8 base codes (A=1 C=2 G=4 T=8 -> 0..3, anything else -> 8)
__device__ __forceinline__ uint32_t codes_of(uint32_t v) {
    const uint32_t code = (((v >> 1) & 0x77777777u) - ((v >> 3) & 0x11111111u)) & 0x33333333u;
    uint32_t pc = v - ((v >> 1) & 0x55555555u); pc = (pc & 0x33333333u) + ((pc >> 2) & 0x33333333u);   // bits set per nibble
    const uint32_t t = pc ^ 0x11111111u;                         // zero iff exactly one bit
    const uint32_t bad = (t | (t >> 1) | (t >> 2)) & 0x11111111u;
    return (code & ~(bad * 7u)) | (bad << 3);
}
// one flag per nibble for the bases lo..hi (clamped to the 16 of a chunk)
__device__ __forceinline__ unsigned long long range_flags(int lo, int hi) {
    lo = max(lo, 0); hi = min(hi, CHUNK - 1);
    if (lo > hi) return 0ull;
    return (ONES << (4 * lo)) & (ONES >> (4 * (CHUNK - 1 - hi)));
}


// QUAL > 2 flags (bit 7 of each byte) of a chunk's 16 QUAL bytes, restricted to its first nb bytes; first / last set byte
__device__ __forceinline__ void qual_gt2_span(const uint32_t (&Q)[4], int nb, int i0, int& first, int& last) {
    unsigned long long g01 = 0, g23 = 0;
#pragma unroll
    for (int wq = 0; wq < 4; wq++) {
        const uint32_t v = Q[wq], f = (((v & 0x7f7f7f7fu) + 0x7d7d7d7du) | v) & 0x80808080u;
        if (wq < 2) g01 |= (unsigned long long)f << (32 * wq); else g23 |= (unsigned long long)f << (32 * (wq - 2));
    }
    if (nb < 8) { g23 = 0; g01 &= nb > 0 ? (~0ull >> (8 * (8 - nb))) : 0ull; } else if (nb < 16) g23 &= (nb > 8) ? (~0ull >> (8 * (16 - nb))) : 0ull;
    first = 0x7fffffff; last = -1;
    if (g01) first = i0 + ((__ffsll((long long)g01) - 1) >> 3); else if (g23) first = i0 + 8 + ((__ffsll((long long)g23) - 1) >> 3);
    if (g23) last = i0 + 8 + ((63 - __clzll((long long)g23)) >> 3); else if (g01) last = i0 + ((63 - __clzll((long long)g01)) >> 3);
}

}  // namespace
