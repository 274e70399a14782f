// bqsr_lane.cuh -- per-lane word-parallel arithmetic of the BQSR count kernel (bqsr_count.cu): a lane owns 32 consecutive bases of
// a read and never touches them one by one.  Everything here is a pure function of registers, written so that the same source
// also compiles as host C++ (tests/c/lane_check.cpp runs it against a base-by-base restatement on the CPU).
//
// Representations of a lane's 32 bases (local index j = 0..31 in the read's stored order):
//   byte words   Q[8]   : byte (j & 3) of word (j >> 2)                       -- QUAL as loaded
//   nibble words N[4]   : nibble k of word w holds base j = 8 w + k           -- SEQ / reference / slot codes, 4 bits per base
//   plane        P      : bit (4 k + w) holds base j = 8 w + k                -- one flag per base; the bit order is a fixed
//                         permutation chosen so that four nibble-flag words combine with three shifted adds, while
//                         brev(P) is exactly the reversal j -> 31 - j (reverse-strand reads are processed in sequencing order)
//   monotonic    G      : bit j holds base j                                  -- only for first / last set searches
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define LANE_FN __host__ __device__ __forceinline__
#else
#define LANE_FN inline
#endif
#if defined(__CUDA_ARCH__)
#define LANE_UNROLL _Pragma("unroll")
#else
#define LANE_UNROLL
#endif

namespace lanes {

#if defined(__CUDA_ARCH__)
LANE_FN uint32_t fsr(uint32_t lo, uint32_t hi, uint32_t sh) { return __funnelshift_r(lo, hi, sh); }
LANE_FN uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) { return __byte_perm(a, b, s); }
LANE_FN uint32_t brev32(uint32_t x) { return __brev(x); }
LANE_FN int popc32(uint32_t x) { return __popc(x); }
LANE_FN int first_set(uint32_t x) { return __ffs((int)x) - 1; }      // x != 0
LANE_FN int last_set(uint32_t x) { return 31 - __clz((int)x); }
#else
LANE_FN uint32_t fsr(uint32_t lo, uint32_t hi, uint32_t sh) { sh &= 31; return sh ? (lo >> sh) | (hi << (32 - sh)) : lo; }
LANE_FN uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) {
    const uint64_t pool = (uint64_t)a | ((uint64_t)b << 32);
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) { const uint32_t sel = (s >> (4 * i)) & 15u; uint32_t byte = (uint32_t)(pool >> (8 * (sel & 7u))) & 255u; if (sel & 8u) byte = (byte & 128u) ? 255u : 0u; r |= byte << (8 * i); }
    return r;
}
LANE_FN uint32_t brev32(uint32_t x) { uint32_t r = 0; for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i); return r; }
LANE_FN int popc32(uint32_t x) { int c = 0; while (x) { x &= x - 1; c++; } return c; }
LANE_FN int first_set(uint32_t x) { int i = 0; while (!((x >> i) & 1u)) i++; return i; }
LANE_FN int last_set(uint32_t x) { int i = 31; while (!((x >> i) & 1u)) i--; return i; }
#endif

// 32 bytes starting at byte k (0..15) of a 48-byte window W[12] (three aligned 16-byte chunks)
LANE_FN void align_bytes32(const uint32_t (&W)[12], uint32_t k, uint32_t (&Q)[8]) {
    const bool s2 = (k & 8u) != 0, s1 = (k & 4u) != 0;
    const uint32_t sh = (k & 3u) * 8u;
    uint32_t V[10], U[9];
LANE_UNROLL
    for (int i = 0; i < 10; i++) V[i] = s2 ? W[i + 2] : W[i];
LANE_UNROLL
    for (int i = 0; i < 9; i++) U[i] = s1 ? V[i + 1] : V[i];
LANE_UNROLL
    for (int i = 0; i < 8; i++) Q[i] = fsr(U[i], U[i + 1], sh);
}

// 32 nibbles starting at nibble (2 kb + par) of a 32-byte window W[8]; bam: the stream keeps the FIRST base of a byte in the HIGH nibble
// (BAM SEQ, sam/bam-files.go) -- bytes are nibble-swapped first, so that the result has base 8 w + k in nibble k of word w
template <bool BAM> LANE_FN void align_nibbles32(const uint32_t (&W)[8], uint32_t kb, uint32_t par, uint32_t (&N)[4]) {
    const bool s2 = (kb & 8u) != 0, s1 = (kb & 4u) != 0;
    const uint32_t sh = (kb & 3u) * 8u + par * 4u;
    uint32_t V[6], U[5];
LANE_UNROLL
    for (int i = 0; i < 6; i++) V[i] = s2 ? W[i + 2] : W[i];
LANE_UNROLL
    for (int i = 0; i < 5; i++) {
        U[i] = s1 ? V[i + 1] : V[i];
        if (BAM) U[i] = ((U[i] & 0x0f0f0f0fu) << 4) | ((U[i] >> 4) & 0x0f0f0f0fu);
    }
LANE_UNROLL
    for (int i = 0; i < 4; i++) N[i] = fsr(U[i], U[i + 1], sh);
}

// plane of bit `beta` of every nibble
LANE_FN uint32_t plane_of(const uint32_t (&N)[4], int beta) {
    const uint32_t M = 0x11111111u;
    return ((N[0] >> beta) & M) + (((N[1] >> beta) & M) << 1) + (((N[2] >> beta) & M) << 2) + (((N[3] >> beta) & M) << 3);
}
// plane of "nibble != 0"
LANE_FN uint32_t plane_any(const uint32_t (&N)[4]) {
    const uint32_t M = 0x11111111u;
    uint32_t f[4];
LANE_UNROLL
    for (int w = 0; w < 4; w++) { const uint32_t x = N[w]; f[w] = (x | (x >> 1) | (x >> 2) | (x >> 3)) & M; }
    return f[0] + (f[1] << 1) + (f[2] << 2) + (f[3] << 3);
}

// QUAL classification through an 8-entry byte table (two registers, looked up with PRMT): index = (q >> sh) & 7.
// Table byte: bits 0..3 one-hot shared-counter slot of the QUAL value (0: not counted -- QUAL < 6, bqsr.go:513), bit 7: QUAL > 2
// (computeStrandedClippedSeq, bqsr.go:312-331).  The host only selects this path when (q >> sh) & 7 separates every QUAL value
// that occurs in the input.  Out: slot nibble words S[4] (layout of N) and the monotonic mask G of QUAL > 2.
LANE_FN void classify_qual(const uint32_t (&Q)[8], uint32_t sh, uint32_t lut_lo, uint32_t lut_hi, uint32_t (&S)[4], uint32_t& G) {
    uint32_t L[8];
LANE_UNROLL
    for (int i = 0; i < 8; i++) {
        const uint32_t f = (Q[i] >> sh) & 0x07070707u;
        const uint32_t y = f | (f >> 4);                      // byte 0 = i0 | i1 << 4, byte 2 = i2 | i3 << 4
        const uint32_t sel = prmt(y, y, 0x3220u);             // selector nibbles (i0, i1, i2, i3) in the low 16 bits
        L[i] = prmt(lut_lo, lut_hi, sel);
    }
    G = 0;
LANE_UNROLL
    for (int i = 0; i < 8; i++) G |= (((L[i] & 0x80808080u) * 0x00204081u) >> 28) << (4 * i);
LANE_UNROLL
    for (int w = 0; w < 4; w++) {
        const uint32_t za = L[2 * w] & 0x0f0f0f0fu, zb = L[2 * w + 1] & 0x0f0f0f0fu;
        const uint32_t ta = za | (za >> 4), tb = zb | (zb >> 4);
        S[w] = prmt(ta, tb, 0x6420u);
    }
}

// previous base (sequencing direction = increasing j after the strand flip): prev(P) bit of base j = P bit of base j - 1;
// edge = flag of the base just before this lane's first one (0 / 1)
LANE_FN uint32_t shift_prev(uint32_t P, uint32_t edge) { return (P << 4) | ((P >> 27) & 0xeu) | edge; }
LANE_FN uint32_t last_flag(uint32_t P) { return P >> 31; }

// plane with the bases j < n set (n = 0..32)
LANE_FN uint32_t range_plane(int n) {
    uint32_t r = 0;
    for (int w = 0; w < 4; w++) {
        int k = n - 8 * w; k = k < 0 ? 0 : (k > 8 ? 8 : k);
        const uint32_t nib = k >= 8 ? 0x11111111u : (0x11111111u & ((1u << (4 * k)) - 1u));
        r |= nib << w;
    }
    return r;
}
// bit position of base j in a plane, and back
LANE_FN int plane_bit(int j) { return 4 * (j & 7) + (j >> 3); }
LANE_FN int plane_base(int bit) { return 8 * (bit & 3) + (bit >> 2); }

// exactly one bit set among the four base planes of a BAM nibble (A = 1, C = 2, G = 4, T = 8; everything else is not ACGT)
LANE_FN uint32_t onehot4(uint32_t a, uint32_t c, uint32_t g, uint32_t t) {
    const uint32_t ac = a ^ c, acg = ac ^ g;
    return (acg ^ t) & ~((a & c) | (ac & g) | (acg & t));
}

// ---- closed-form read clipping (what bqsr_prep2_kernel does per read; host-checkable against the oracle's step-by-step clipping) ----
// hardClipAdaptorSequence + hardClipSoftClippedBases (filters/utils.go:148-222, 506-534) have a closed form for two CIGAR shapes:
//   [H..][S a] M m [S b][H..]                    the adaptor boundary maps to read coordinate boundary - (POS - a): no D/N/I to fall into,
//                                                so computeReadCoordinateForReferenceCoordinate (:267-326) is linear and never fails
//   [H..][S a] M m1 (I|D) d M m2 [S b][H..]     only when no adaptor clipping applies (the mapping through an indel has quirks)
// kind: -1 the read is not recalibrated (SEQ length != CIGAR read length, or clipped away), 0 / 1 the two shapes, 2 anything else
// (general path).  Kept bases [lo, hi) of the stored read, cpos = POS of the first kept base; bp = kept-read index of the indel.
struct ClipShape { int kind, lo, hi, cpos, bp, ins, del; };
template <class CigarAt>
LANE_FN ClipShape closed_form_clip(uint32_t f, int pos0, int pnext, int tlen, int nref, int L0, int nc0, CigarAt cigar_at) {
    ClipShape R; R.kind = 2; R.lo = R.hi = 0; R.cpos = pos0; R.bp = R.ins = R.del = 0;
    if (nc0 < 1 || nc0 > 9) return R;
    int i = 0, a = 0, b = 0, m1 = 0, m2 = 0, d = 0, dop = -1;
    uint32_t op = 0;
    auto nxt = [&]() { op = (i < nc0) ? cigar_at(i) : 0xfu; i++; };      // 0xf: end marker (operation code 15 does not exist)
    auto opc = [&]() { return (int)(op & 15u); };
    auto opl = [&]() { return (int)(op >> 4); };
    nxt();
    while (opc() == 5 && i <= nc0) nxt();
    if (opc() == 4) { a = opl(); nxt(); }
    if ((opc() == 0 || opc() == 7 || opc() == 8) && opl() > 0) { m1 = opl(); nxt(); } else return R;
    if ((opc() == 1 || opc() == 2) && opl() > 0) {
        dop = opc(); d = opl(); nxt();
        if ((opc() == 0 || opc() == 7 || opc() == 8) && opl() > 0) { m2 = opl(); nxt(); } else return R;
    }
    if (opc() == 4 && i <= nc0) { b = opl(); nxt(); }
    while (opc() == 5 && i <= nc0) nxt();
    if (i != nc0 + 1) return R;                                             // something else follows
    const int ins = dop == 1 ? d : 0, del = dop == 2 ? d : 0;
    if (a + m1 + ins + m2 + b != L0) { R.kind = -1; return R; }             // SEQ length != read length of the CIGAR (bqsr.go:236-238)
    int lo = a, hi = a + m1 + ins + m2, cpos = pos0;
    const int alnEnd = pos0 + m1 + m2 + del - 1;
    // hardClipAdaptorSequence (utils.go:148-222)
    const bool paired = (f & 0x1u) != 0, rev = (f & 0x10u) != 0, nrev = (f & 0x20u) != 0;
    const bool next_unmapped = (f & 0x8u) || nref < 0 || pnext == 0;          // isStrictNextUnmapped (utils.go:144)
    bool well = false;
    if (tlen != 0 && paired && !next_unmapped && rev != nrev) well = rev ? (alnEnd > pnext) : (pos0 <= pnext + tlen);
    if (well) {
        const int boundary = rev ? pnext - 1 : pos0 + (tlen < 0 ? -tlen : tlen);
        if (boundary >= pos0 && boundary <= alnEnd) {
            if (dop >= 0) return R;                                         // through an indel: general path
            const int rc = boundary - (pos0 - a);                          // read coordinate of the boundary
            if (rev) { lo = rc + 1; cpos = boundary + 1; } else hi = rc;
        }
    }
    if (hi - lo <= 0) { R.kind = -1; return R; }                            // clipped away: dropped (bqsr.go:483-490)
    R.kind = dop >= 0 ? 1 : 0; R.lo = lo; R.hi = hi; R.cpos = cpos; R.bp = m1; R.ins = ins; R.del = del;
    return R;
}

// ---- one pass of one lane, in three stages separated by the two cross-lane steps of the count kernel --------------------------------
// (RT is a callable n -> range_plane(clamp(n, 0, 32)): a shared-memory table on the device, the function itself on the host)
struct LaneRec {            // the fields of a 32-byte work record that the lane arithmetic needs
    int Lk; bool rev; uint32_t skip0, skip1; int bp, insl;
};
struct LaneWin {            // the staged windows and their sub-chunk offsets
    uint32_t QW[12], SW[8], RW[8], RW2[8];
    uint32_t kq, kb, spar, kr, rpar, kr2, rpar2;
};
template <int S> struct LaneS1 { uint32_t pA, pC, pG, pT, Mm, Sp[S]; int lf, ll; };

// stage 1 (stored order): slot planes, base planes, mismatch plane; lf / ll = kept-read coordinate of the lane's first / last base with QUAL > 2
template <int S, bool INDEL, class RTF>
LANE_FN void lane_stage1(const LaneWin& w, const LaneRec& rec, int c, uint32_t sh, uint32_t lut_lo, uint32_t lut_hi, RTF RT, LaneS1<S>& o) {
    const int nb = rec.Lk - 32 * c < 0 ? 0 : (rec.Lk - 32 * c > 32 ? 32 : rec.Lk - 32 * c);
    const int ow = rec.rev ? rec.Lk - 32 * c - 32 : 32 * c;
    uint32_t Q[8], Sn[4], Gm, N[4], R[4];
    align_bytes32(w.QW, w.kq, Q);
    classify_qual(Q, sh, lut_lo, lut_hi, Sn, Gm);
    align_nibbles32<true>(w.SW, w.kb, w.spar, N);
    align_nibbles32<false>(w.RW, w.kr, w.rpar, R);
LANE_UNROLL
    for (int i = 0; i < 4; i++) R[i] ^= N[i];
    o.Mm = plane_any(R);
    if (INDEL) {
        align_nibbles32<false>(w.RW2, w.kr2, w.rpar2, R);
LANE_UNROLL
        for (int i = 0; i < 4; i++) R[i] ^= N[i];
        // before the indel the first window, after it (and after inserted bases, which have no reference base) the second
        o.Mm = (o.Mm & RT(rec.bp - ow)) | (plane_any(R) & ~RT(rec.bp + rec.insl - ow));
    }
    o.pA = plane_of(N, 0); o.pC = plane_of(N, 1); o.pG = plane_of(N, 2); o.pT = plane_of(N, 3);
LANE_UNROLL
    for (int s = 0; s < S; s++) o.Sp[s] = plane_of(Sn, s);
    const uint32_t bytes_in = nb >= 32 ? 0xffffffffu : (nb <= 0 ? 0u : (rec.rev ? ~(0xffffffffu >> nb) : ((1u << nb) - 1u)));
    const uint32_t gv = Gm & bytes_in;
    o.lf = gv ? ow + first_set(gv) : 0x7fffffff; o.ll = gv ? ow + last_set(gv) : -1;
}
template <int S> struct LaneS2 { uint32_t pA, pC, pG, pT, Mm, Sp[S], counted, Vc, epack; };
// stage 2 (sequencing order from here on): strand flip, counted bases, bases that are not N for the context covariate
template <int S, class RTF>
LANE_FN void lane_stage2(const LaneS1<S>& a, const LaneRec& rec, int c, int leftPos, int rightPos, RTF RT, LaneS2<S>& o) {
    const int nb = rec.Lk - 32 * c < 0 ? 0 : (rec.Lk - 32 * c > 32 ? 32 : rec.Lk - 32 * c);
    const int i0 = 32 * c, Lk = rec.Lk;
    o.pA = a.pA; o.pC = a.pC; o.pG = a.pG; o.pT = a.pT; o.Mm = a.Mm;
LANE_UNROLL
    for (int s = 0; s < S; s++) o.Sp[s] = a.Sp[s];
    if (rec.rev) {           // reverse complement (bqsr.go:140-146)
        o.pA = brev32(a.pT); o.pT = brev32(a.pA); o.pC = brev32(a.pG); o.pG = brev32(a.pC); o.Mm = brev32(a.Mm);
LANE_UNROLL
        for (int s = 0; s < S; s++) o.Sp[s] = brev32(a.Sp[s]);
    }
    const uint32_t valid = onehot4(o.pA, o.pC, o.pG, o.pT) & RT(nb);
    uint32_t skip = 0;           // known sites (calculateSkipSlice, bqsr.go:389-414): inclusive ranges of kept-read coordinates
    if ((rec.skip0 & 0xffffu) != 0xffffu) {
        const int fs = (int)(rec.skip0 & 0xffffu), fe = (int)(rec.skip0 >> 16);
        skip = rec.rev ? (RT(Lk - fs - i0) & ~RT(Lk - 1 - fe - i0)) : (RT(fe + 1 - i0) & ~RT(fs - i0));
        if ((rec.skip1 & 0xffffu) != 0xffffu) {
            const int fs2 = (int)(rec.skip1 & 0xffffu), fe2 = (int)(rec.skip1 >> 16);
            skip |= rec.rev ? (RT(Lk - fs2 - i0) & ~RT(Lk - 1 - fe2 - i0)) : (RT(fe2 + 1 - i0) & ~RT(fs2 - i0));
        }
    }
    o.counted = valid & ~skip;           // ACGT, inside the read, not a known site; QUAL >= 6 comes with the slot planes
    const int wl = rec.rev ? Lk - 1 - rightPos : leftPos, wh = rec.rev ? Lk - 1 - leftPos : rightPos;   // low-quality tails read as N (bqsr.go:312-331)
    o.Vc = valid & RT(wh + 1 - i0) & ~RT(wl - i0);
    o.epack = (o.pA >> 31) | ((o.pC >> 31) << 1) | ((o.pG >> 31) << 2) | ((o.pT >> 31) << 3) | ((o.Vc >> 31) << 4);
}
struct LaneS3 { uint32_t qA, qC, qG, qT, okc; };
// stage 3: previous base in sequencing direction (ein: the packed last flags of the lane below, 0 for a read's first lane)
template <int S> LANE_FN void lane_stage3(const LaneS2<S>& a, uint32_t ein, LaneS3& o) {
    o.qA = shift_prev(a.pA, ein & 1u); o.qC = shift_prev(a.pC, (ein >> 1) & 1u); o.qG = shift_prev(a.pG, (ein >> 2) & 1u); o.qT = shift_prev(a.pT, (ein >> 3) & 1u);
    o.okc = a.counted & a.Vc & shift_prev(a.Vc, (ein >> 4) & 1u);
}

}  // namespace lanes
