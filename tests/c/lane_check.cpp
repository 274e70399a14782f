// lane_check.cpp -- CPU check of the per-lane arithmetic of the BQSR count kernel (elprep_b200/csrc/bqsr_lane.cuh, the very source the
// CUDA kernel compiles): random reads are laid out in byte arenas exactly as on the device (QUAL bytes, BAM SEQ nibbles, one-hot
// reference nibbles, arbitrary byte / nibble offsets), the windows are cut out as aligned 16-byte chunks the way the kernel's cp.async
// stage does, the three lane stages run for every lane of every read, and the resulting planes are compared base by base with a direct
// restatement of (*BaseRecalibrator).Recalibrate's per-base rules (filters/bqsr.go:467-551).  The cycle / context counters are then
// accumulated with the kernel's own vertical-counter and popcount scheme and compared with counters incremented base by base.
//   g++ -O2 -std=c++17 -I elprep_b200/csrc -o tests/c/_build/lane_check tests/c/lane_check.cpp && tests/c/_build/lane_check
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>
#include "bqsr_lane.cuh"

static uint64_t rng_state = 88172645463325252ull;
static uint64_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }
static int below(int n) { return (int)(rnd() % (uint64_t)n); }

static const int QV[4] = {2, 12, 23, 37};
template <int S, bool INDEL> static int run(int n_reads, int maxL, const int* qvals, int nq) {
    using namespace lanes;
    const int lpr = (maxL + 31) / 32;
    // classifier as the host plans it (plan_fast in bqsr_gather.cu)
    uint32_t sh = 0, lut_lo = 0, lut_hi = 0; int slot_of[128]; for (int i = 0; i < 128; i++) slot_of[i] = -1;
    { int ns = 0; for (int i = 0; i < nq; i++) if (qvals[i] >= 6) slot_of[qvals[i]] = ns++; if (ns != S) { printf("bad S\n"); return 1; } }
    bool found = false;
    for (sh = 0; sh <= 4 && !found; sh++) {
        uint32_t seen = 0; bool good = true;
        for (int i = 0; i < nq; i++) { uint32_t ix = ((uint32_t)qvals[i] >> sh) & 7u; if (seen & (1u << ix)) good = false; seen |= 1u << ix; }
        if (good) { found = true; break; }
    }
    if (!found) { printf("no classifier\n"); return 1; }
    uint8_t lut[8] = {0};
    for (int i = 0; i < nq; i++) { int q = qvals[i]; uint8_t b = q > 2 ? 0x80 : 0; if (q >= 6) b |= (uint8_t)(1u << slot_of[q]); lut[((uint32_t)q >> sh) & 7u] = b; }
    memcpy(&lut_lo, lut, 4); memcpy(&lut_hi, lut + 4, 4);
    auto RT = [](int n) { return range_plane(n < 0 ? 0 : (n > 32 ? 32 : n)); };

    const size_t QN = 1 << 16, REFN = 1 << 16;
    std::vector<uint8_t> qual(QN + 256), seq(QN / 2 + 256), refhot(REFN / 2 + 1024);
    const int REFPAD = 512;
    std::map<long, long> cyc_naive, ctx_naive, cyc_mis_naive, ctx_mis_naive, cyc_k, ctx_k, cyc_mis_k, ctx_mis_k;
    int errors = 0;
    // kernel-style accumulators for one "segment" of reads of one class: lanes (r = 0, c)
    std::vector<uint32_t> pl(lpr * S * 8, 0); std::vector<uint32_t> cx(lpr * 16 * S, 0);
    for (int rd = 0; rd < n_reads; rd++) {
        for (auto& b : qual) b = (uint8_t)qvals[below(nq)];
        for (auto& b : seq) { static const uint8_t nb[6] = {1, 2, 4, 8, 15, 3}; int a = below(100) < 96 ? below(4) : 4 + below(2), c2 = below(100) < 96 ? below(4) : 4 + below(2); b = (uint8_t)((nb[a] << 4) | nb[c2]); }
        for (auto& b : refhot) { static const uint8_t nb[5] = {1, 2, 4, 8, 0}; b = (uint8_t)(nb[below(100) < 98 ? below(4) : 4] | (nb[below(100) < 98 ? below(4) : 4] << 4)); }
        if (rd % 3 == 0) {   // make most read bases equal to the reference so that mismatches are sparse as in real data: done below per read
        }
        LaneRec rec;
        rec.Lk = 1 + below(maxL); rec.rev = below(2);
        if (below(4) == 0) rec.Lk = maxL;
        rec.skip0 = 0xffffu; rec.skip1 = 0xffffu;
        if (below(3) == 0) { int fs = below(rec.Lk), fe = fs + below(12); if (fe > rec.Lk - 1) fe = rec.Lk - 1; rec.skip0 = (uint32_t)fs | ((uint32_t)fe << 16);
            if (below(2) == 0) { int fs2 = below(rec.Lk), fe2 = fs2 + below(3); if (fe2 > rec.Lk - 1) fe2 = rec.Lk - 1; rec.skip1 = (uint32_t)fs2 | ((uint32_t)fe2 << 16); } }
        rec.bp = 0; rec.insl = 0; int delta = 0;
        if (INDEL) { rec.bp = 1 + below(rec.Lk > 2 ? rec.Lk - 1 : 1); if (below(2)) { rec.insl = 1 + below(6); if (rec.bp + rec.insl > rec.Lk) rec.insl = rec.Lk - rec.bp; delta = -rec.insl; } else delta = 1 + below(40); }
        const long qbyte0 = 64 + below(4000), snib0 = 128 + below(8000), cpos = 1 + below(20000) + (below(10) == 0 ? 0 : 300);
        // ~97 % of the read's bases copy the reference
        auto ref_at = [&](long idx) -> int { const uint8_t b = refhot[REFPAD + (idx >> 1)]; return (idx & 1) ? (b >> 4) : (b & 15); };
        auto set_seq = [&](long nidx, int nib) { uint8_t& b = seq[nidx >> 1]; if (nidx & 1) b = (uint8_t)((b & 0xf0) | nib); else b = (uint8_t)((b & 0x0f) | (nib << 4)); };   // BAM: even index in the high nibble
        auto seq_at = [&](long nidx) -> int { const uint8_t b = seq[nidx >> 1]; return (nidx & 1) ? (b & 15) : (b >> 4); };
        auto ref_of = [&](int o, bool& has) -> int { has = true; if (!INDEL) return ref_at(cpos - 1 + o); if (o < rec.bp) return ref_at(cpos - 1 + o); if (o >= rec.bp + rec.insl) return ref_at(cpos - 1 + o + delta); has = false; return 0; };
        for (int o = 0; o < rec.Lk; o++) { bool has; int rn = ref_of(o, has); if (has && rn && below(100) < 97) set_seq(snib0 + o, rn); }
        // ---- direct restatement ----
        int leftPos = 0x7fffffff, rightPos = -1;
        for (int o = 0; o < rec.Lk; o++) if (qual[qbyte0 + o] > 2) { if (leftPos == 0x7fffffff) leftPos = o; rightPos = o; }
        auto in_skip = [&](int o) { for (uint32_t sk : {rec.skip0, rec.skip1}) if ((sk & 0xffffu) != 0xffffu && o >= (int)(sk & 0xffffu) && o <= (int)(sk >> 16)) return true; return false; };
        std::vector<int> e_counted(rec.Lk), e_slot(rec.Lk), e_okc(rec.Lk), e_ctx(rec.Lk), e_mis(rec.Lk);
        auto code_of = [](int nib) { return nib == 1 ? 0 : nib == 2 ? 1 : nib == 4 ? 2 : nib == 8 ? 3 : -1; };
        for (int s = 0; s < rec.Lk; s++) {
            const int o = rec.rev ? rec.Lk - 1 - s : s;
            const int nib = seq_at(snib0 + o), cd = code_of(nib), q = qual[qbyte0 + o];
            const bool nonN = cd >= 0 && o >= leftPos && o <= rightPos;
            e_slot[s] = q >= 6 ? slot_of[q] : -1;
            e_counted[s] = cd >= 0 && !in_skip(o) && e_slot[s] >= 0;
            bool prev_nonN = false; int pcd = -1;
            if (s > 0) { const int op = rec.rev ? rec.Lk - s : s - 1; pcd = code_of(seq_at(snib0 + op)); prev_nonN = pcd >= 0 && op >= leftPos && op <= rightPos; }
            e_okc[s] = e_counted[s] && nonN && prev_nonN;
            const int sc = rec.rev ? 3 - cd : cd, sp = rec.rev ? 3 - pcd : pcd;
            e_ctx[s] = e_okc[s] ? (sp | (sc << 2)) : -1;
            bool has; const int rn = ref_of(o, has);
            e_mis[s] = e_counted[s] && has && rn != nib;
            if (e_counted[s]) { cyc_naive[(long)e_slot[s] * 100000 + s]++; if (e_mis[s]) cyc_mis_naive[(long)e_slot[s] * 100000 + s]++; }
            if (e_okc[s]) { ctx_naive[e_slot[s] * 16 + e_ctx[s]]++; if (e_mis[s]) ctx_mis_naive[e_slot[s] * 16 + e_ctx[s]]++; }
        }
        // ---- the lanes ----
        std::vector<LaneS1<S>> s1(lpr); std::vector<LaneS2<S>> s2(lpr); std::vector<LaneS3> s3(lpr);
        for (int c = 0; c < lpr; c++) {
            const int ow = rec.rev ? rec.Lk - 32 * c - 32 : 32 * c;
            LaneWin w; memset(&w, 0, sizeof w);
            const long qa = qbyte0 + ow, q16 = qa & ~15L; w.kq = (uint32_t)(qa & 15); memcpy(w.QW, &qual[q16], 48);
            const long ni = snib0 + ow, sb = ni >> 1, s16 = sb & ~15L; w.kb = (uint32_t)(sb & 15); w.spar = (uint32_t)(ni & 1); memcpy(w.SW, &seq[s16], 32);
            const long ri = cpos - 1 + ow, rbyte = REFPAD + (ri >> 1), r16 = rbyte & ~15L; w.kr = (uint32_t)(rbyte & 15); w.rpar = (uint32_t)(ri & 1); memcpy(w.RW, &refhot[r16], 32);
            if (INDEL) { const long ri2 = ri + delta, rb2 = REFPAD + (ri2 >> 1), r216 = rb2 & ~15L; w.kr2 = (uint32_t)(rb2 & 15); w.rpar2 = (uint32_t)(ri2 & 1); memcpy(w.RW2, &refhot[r216], 32); }
            lane_stage1<S, INDEL>(w, rec, c, sh, lut_lo, lut_hi, RT, s1[c]);
        }
        int lp = 0x7fffffff, rp = -1;
        for (int c = 0; c < lpr; c++) { if (s1[c].lf < lp) lp = s1[c].lf; if (s1[c].ll > rp) rp = s1[c].ll; }
        if (lp != leftPos || rp != rightPos) { if (errors++ < 10) printf("read %d: tails %d %d expected %d %d\n", rd, lp, rp, leftPos, rightPos); }
        for (int c = 0; c < lpr; c++) lane_stage2<S>(s1[c], rec, c, lp, rp, RT, s2[c]);
        for (int c = 0; c < lpr; c++) lane_stage3<S>(s2[c], c ? s2[c - 1].epack : 0u, s3[c]);
        for (int c = 0; c < lpr; c++) {
            const uint32_t c0 = s2[c].pC | s2[c].pT, c1 = s2[c].pG | s2[c].pT, p0 = s3[c].qC | s3[c].qT, p1 = s3[c].qG | s3[c].qT;
            for (int t = 0; t < 32; t++) {
                const int s = 32 * c + t, b = plane_bit(t);
                int slot = -1; for (int k = 0; k < S; k++) if ((s2[c].Sp[k] >> b) & 1u) slot = slot < 0 ? k : 99;
                const bool counted = ((s2[c].counted >> b) & 1u) && slot >= 0, okc = ((s3[c].okc >> b) & 1u) && slot >= 0, mis = ((s2[c].Mm >> b) & 1u) && counted;
                const int ctx = (int)(((p0 >> b) & 1u) | (((p1 >> b) & 1u) << 1) | (((c0 >> b) & 1u) << 2) | (((c1 >> b) & 1u) << 3));
                const bool in = s < rec.Lk;
                const bool e_c = in && e_counted[s], e_o = in && e_okc[s], e_m = in && e_mis[s];
                if (counted != e_c || okc != e_o || mis != e_m || (counted && slot != e_slot[s]) || (okc && ctx != e_ctx[s])) {
                    if (errors++ < 20) printf("read %d (Lk %d rev %d) lane %d t %d: counted %d/%d okc %d/%d mis %d/%d slot %d/%d ctx %d/%d\n", rd, rec.Lk, rec.rev, c, t, counted, e_c, okc, e_o, mis, e_m, slot,
                                              in ? e_slot[s] : -9, ctx, in ? e_ctx[s] : -9);
                }
            }
            // kernel-style accumulation
            for (int k = 0; k < S; k++) {
                uint32_t carry = s2[c].Sp[k] & s2[c].counted;
                for (int i = 0; i < 8; i++) { uint32_t& P = pl[(c * S + k) * 8 + i]; const uint32_t t = P & carry; P ^= carry; carry = t; }
            }
            const uint32_t cur[4] = {s2[c].pA, s2[c].pC, s2[c].pG, s2[c].pT}, prv[4] = {s3[c].qA, s3[c].qC, s3[c].qG, s3[c].qT};
            for (int cc = 0; cc < 4; cc++) for (int pp = 0; pp < 4; pp++) for (int k = 0; k < S; k++) cx[c * 16 * S + (pp + 4 * cc) * S + k] += (uint32_t)popc32(prv[pp] & cur[cc] & s2[c].Sp[k] & s3[c].okc);
            uint32_t mm = s2[c].Mm & s2[c].counted;
            for (int k = 0; k < S; k++) { uint32_t ms = mm & s2[c].Sp[k]; while (ms) { const int b = first_set(ms); ms &= ms - 1; cyc_mis_k[(long)k * 100000 + 32 * c + plane_base(b)]++;
                    if ((s3[c].okc >> b) & 1u) ctx_mis_k[k * 16 + (int)(((p0 >> b) & 1u) | (((p1 >> b) & 1u) << 1) | (((c0 >> b) & 1u) << 2) | (((c1 >> b) & 1u) << 3))]++; } }
        }
        if ((rd + 1) % 255 == 0 || rd == n_reads - 1) {   // segment flush
            for (int c = 0; c < lpr; c++) for (int k = 0; k < S; k++) for (int b = 0; b < 32; b++) { uint32_t cnt = 0; for (int i = 0; i < 8; i++) cnt |= ((pl[(c * S + k) * 8 + i] >> b) & 1u) << i; if (cnt) cyc_k[(long)k * 100000 + 32 * c + plane_base(b)] += cnt; }
            for (int c = 0; c < lpr; c++) for (int cell = 0; cell < 16 * S; cell++) if (cx[c * 16 * S + cell]) ctx_k[(cell % S) * 16 + cell / S] += cx[c * 16 * S + cell];
            std::fill(pl.begin(), pl.end(), 0); std::fill(cx.begin(), cx.end(), 0);
        }
    }
    if (cyc_k != cyc_naive) { printf("cycle counters differ\n"); errors++; }
    if (ctx_k != ctx_naive) { printf("context counters differ\n"); errors++; }
    if (cyc_mis_k != cyc_mis_naive) { printf("cycle mismatch counters differ\n"); errors++; }
    if (ctx_mis_k != ctx_mis_naive) { printf("context mismatch counters differ\n"); errors++; }
    long tot = 0; for (auto& kv : cyc_naive) tot += kv.second; long tm = 0; for (auto& kv : cyc_mis_naive) tm += kv.second; long tc = 0; for (auto& kv : ctx_naive) tc += kv.second;
    printf("S=%d indel=%d maxL=%d reads=%d: %ld counted bases, %ld with context, %ld mismatches, %d errors\n", S, (int)INDEL, maxL, n_reads, tot, tc, tm, errors);
    return errors;
}

int main() {
    int e = 0;
    e += run<3, false>(3000, 150, QV, 4);
    e += run<3, true>(3000, 150, QV, 4);
    { const int q2[2] = {2, 30}; e += run<1, false>(1000, 100, q2, 2); }
    { const int q5[5] = {2, 11, 25, 37, 40}; e += run<4, false>(1500, 250, q5, 5); }
    { const int q3[3] = {3, 9, 41}; e += run<2, true>(1000, 36, q3, 3); }
    { const int q8[8] = {0, 1, 2, 3, 4, 13, 22, 31}; e += run<3, false>(600, 1024, q8, 8); }   // eight values, separated by q & 7
    printf(e ? "FAILED\n" : "OK\n");
    return e ? 1 : 0;
}
