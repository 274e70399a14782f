// bqsr_simd.cuh -- word-parallel helpers shared by the BQSR chunk kernels (gather and apply): a lane owns 16 consecutive
// bases of a read and keeps them in registers as 4 QUAL words and 64-bit words of 4-bit fields (one nibble per base).
#pragma once
#include <cstdint>
#include "common.cuh"

namespace {

constexpr int CHUNK = 16;       // bases per lane
constexpr unsigned long long ONES = 0x1111111111111111ull;

// 16 bytes at any byte address: five aligned 32-bit loads, funnel-shifted (reads up to 3 bytes past p + 16: arenas are padded)
__device__ __forceinline__ void load16_unaligned(const uint8_t* p, uint32_t (&o)[4]) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* b = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
    const uint32_t sh = ((uint32_t)a & 3u) * 8u;
    const uint32_t w0 = __ldg(b), w1 = __ldg(b + 1), w2 = __ldg(b + 2), w3 = __ldg(b + 3), w4 = __ldg(b + 4);
    o[0] = __funnelshift_r(w0, w1, sh); o[1] = __funnelshift_r(w1, w2, sh); o[2] = __funnelshift_r(w2, w3, sh); o[3] = __funnelshift_r(w3, w4, sh);
}
// 16 nibbles starting at nibble index `nidx` of a stream whose nibble 2b is the LOW nibble of byte b: the 9 bytes that hold
// them lie inside three aligned words, and byte offset and nibble parity fold into ONE funnel shift (< 32 bits)
__device__ __forceinline__ unsigned long long load16_nibbles_le(const uint8_t* stream, uint64_t nidx) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(stream + (nidx >> 1));
    const uint32_t* b = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
    const uint32_t sh = ((uint32_t)a & 3u) * 8u + (uint32_t)(nidx & 1) * 4u;
    const uint32_t w0 = __ldg(b), w1 = __ldg(b + 1), w2 = __ldg(b + 2);
    return (unsigned long long)__funnelshift_r(w0, w1, sh) | ((unsigned long long)__funnelshift_r(w1, w2, sh) << 32);
}
// the same for BAM SEQ (nibble 2b is the HIGH nibble of byte b): swap the nibbles of every byte first
__device__ __forceinline__ unsigned long long load16_nibbles_bam(const uint8_t* stream, uint64_t nidx) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(stream + (nidx >> 1));
    const uint32_t* b = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
    const uint32_t sh = ((uint32_t)a & 3u) * 8u + (uint32_t)(nidx & 1) * 4u;
    uint32_t w[3] = {__ldg(b), __ldg(b + 1), __ldg(b + 2)};
#pragma unroll
    for (int k = 0; k < 3; k++) w[k] = ((w[k] & 0x0f0f0f0fu) << 4) | ((w[k] >> 4) & 0x0f0f0f0fu);
    return (unsigned long long)__funnelshift_r(w[0], w[1], sh) | ((unsigned long long)__funnelshift_r(w[1], w[2], sh) << 32);
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
    uint32_t f[4];
#pragma unroll
    for (int wq = 0; wq < 4; wq++) { const uint32_t v = Q[wq]; f[wq] = (((v & 0x7f7f7f7fu) + 0x7d7d7d7du) | v) & 0x80808080u; }
    if (nb == CHUNK && (f[0] & f[1] & f[2] & f[3]) == 0x80808080u) { first = i0; last = i0 + CHUNK - 1; return; }   // the usual case
    unsigned long long g01 = (unsigned long long)f[0] | ((unsigned long long)f[1] << 32), g23 = (unsigned long long)f[2] | ((unsigned long long)f[3] << 32);
    if (nb < 8) { g23 = 0; g01 &= nb > 0 ? (~0ull >> (8 * (8 - nb))) : 0ull; } else if (nb < 16) g23 &= (nb > 8) ? (~0ull >> (8 * (16 - nb))) : 0ull;
    first = 0x7fffffff; last = -1;
    if (g01) first = i0 + ((__ffsll((long long)g01) - 1) >> 3); else if (g23) first = i0 + 8 + ((__ffsll((long long)g23) - 1) >> 3);
    if (g23) last = i0 + 8 + ((63 - __clzll((long long)g23)) >> 3); else if (g01) last = i0 + ((63 - __clzll((long long)g01)) >> 3);
}

}  // namespace
