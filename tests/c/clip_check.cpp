// clip_check.cpp -- the closed-form read clipping of bqsr_prep2_kernel (lanes::closed_form_clip, elprep_b200/csrc/bqsr_lane.cuh) against the
// oracle's step-by-step restatement of hardClipAdaptorSequence + hardClipSoftClippedBases (oracle/oracle.c: orc_probe_clip, following
// filters/utils.go:148-534) on random reads of the shapes the closed form claims, plus the known-site read coordinates against
// orc_probe_readcoord on the clipped CIGAR.  Linked against oracle/_build/liboracle.so -- test infrastructure only.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "bqsr_lane.cuh"
extern "C" {
int orc_probe_clip(int32_t pos, uint16_t flag, int32_t pnext, int32_t tlen, int32_t refid, int32_t nref, const uint32_t* cigar, int32_t ncigar, int32_t lseq,
                   int32_t* lo, int32_t* hi, int32_t* newpos, uint32_t* newcigar, int cap);
int orc_probe_readcoord(const uint32_t* cigar, int32_t ncigar, int soft_start, int ref_index, int tail_right, int* ok);
}
static uint64_t st = 0x9E3779B97F4A7C15ull;
static uint64_t rnd() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; }
static int below(int n) { return (int)(rnd() % (uint64_t)n); }
static uint32_t mk(int len, int op) { return ((uint32_t)len << 4) | (uint32_t)op; }

int main() {
    int errors = 0; long n_simple = 0, n_indel = 0, n_drop = 0, n_general = 0, n_adaptor = 0, n_sites = 0;
    for (int it = 0; it < 400000; it++) {
        std::vector<uint32_t> cg;
        const int L = 30 + below(221);
        int rest = L;
        const int shape = below(10);
        if (below(6) == 0) cg.push_back(mk(1 + below(20), 5));
        int a = 0, b = 0;
        if (below(3) == 0) { a = 1 + below(rest / 4); rest -= a; }
        if (below(3) == 0) { b = 1 + below(rest / 4); rest -= b; }
        if (a) cg.push_back(mk(a, 4));
        const int mops[3] = {0, 7, 8};
        if (shape < 6) cg.push_back(mk(rest, mops[below(10) ? 0 : 1 + below(2)]));
        else if (shape < 9) {               // one insertion / deletion
            const int m1 = 1 + below(rest - 2), isins = below(2);
            int d = 1 + below(isins ? 8 : 30);
            if (isins && m1 + d >= rest) d = rest - m1 - 1;
            if (d < 1) { cg.push_back(mk(rest, 0)); }
            else { cg.push_back(mk(m1, 0)); cg.push_back(mk(d, isins ? 1 : 2)); cg.push_back(mk(rest - m1 - (isins ? d : 0), 0)); }
        } else {                            // something the closed form must hand to the general path
            const int m1 = 1 + below(rest - 6);
            cg.push_back(mk(m1, 0)); cg.push_back(mk(2, 1)); cg.push_back(mk(2, 0)); cg.push_back(mk(3, 2)); cg.push_back(mk(rest - m1 - 4, 0));
        }
        if (b) cg.push_back(mk(b, 4));
        if (below(6) == 0) cg.push_back(mk(1 + below(20), 5));
        int reflen = 0, rl = 0;
        for (uint32_t op : cg) { const int o = op & 15, l = op >> 4; if (o == 0 || o == 2 || o == 3 || o == 7 || o == 8) reflen += l; if (o == 0 || o == 1 || o == 4 || o == 7 || o == 8) rl += l; }
        if (rl != L) { printf("generator bug\n"); return 2; }
        const int pos = 500 + below(100000);
        uint16_t flag = (uint16_t)(0x1 | (below(2) ? 0x10 : 0) | (below(2) ? 0x20 : 0) | (below(2) ? 0x40 : 0x80) | (below(12) == 0 ? 0x8 : 0));
        if (below(10) == 0) flag &= ~0x1;
        const int rev = (flag & 0x10) != 0;
        int pnext = pos + below(2 * L) - L, tlen = below(3) == 0 ? 0 : (below(2 * L + 40) - L - 20);
        if (below(3) == 0) { if (rev) { pnext = pos + below(reflen + 10) - 5; } else { tlen = below(reflen + 20) - 5; pnext = pos + below(40) - 20; } }
        if (below(15) == 0) pnext = 0;
        const int nref = below(20) == 0 ? -1 : 3;
        const lanes::ClipShape cs = lanes::closed_form_clip((uint32_t)flag, pos, pnext, tlen, nref, L, (int)cg.size(), [&](int i) { return cg[i]; });
        int32_t lo = 0, hi = 0, np = 0; uint32_t nc[64];
        const int rc = orc_probe_clip(pos, flag, pnext, tlen, 3, nref, cg.data(), (int)cg.size(), L, &lo, &hi, &np, nc, 64);
        if (cs.kind == 2) { n_general++; if (shape < 6 ) { if (errors++ < 10) printf("it %d: single-M read went to the general path\n", it); } continue; }
        if (cs.kind < 0) { n_drop++; if (rc != -1) { if (errors++ < 10) printf("it %d: closed form drops the read, oracle rc %d lo %d hi %d\n", it, rc, lo, hi); } continue; }
        if (rc < 0) { if (errors++ < 10) printf("it %d: oracle drops / fails (%d), closed form keeps [%d,%d)\n", it, rc, cs.lo, cs.hi); continue; }
        (cs.kind ? n_indel : n_simple)++;
        if (cs.lo != lo || cs.hi != hi || cs.cpos != np) { if (errors++ < 10) printf("it %d: kept [%d,%d) pos %d, oracle [%d,%d) pos %d (flag %x pnext %d tlen %d)\n", it, cs.lo, cs.hi, cs.cpos, lo, hi, np, flag, pnext, tlen); continue; }
        if (hi - lo != L - (a + b) || np != pos) n_adaptor++;
        // the clipped CIGAR must be [H] M [H] (or M (I|D) M between the hard clips) of the claimed lengths
        std::vector<uint32_t> core; for (int k = 0; k < rc; k++) if ((nc[k] & 15) != 5) core.push_back(nc[k]);
        const int Lk = cs.hi - cs.lo;
        bool ok = cs.kind == 0 ? (core.size() == 1 && (int)(core[0] >> 4) == Lk) :
                                 (core.size() == 3 && (int)(core[0] >> 4) == cs.bp && (int)(core[1] >> 4) == (cs.ins ? cs.ins : cs.del) && (int)(core[1] & 15) == (cs.ins ? 1 : 2));
        if (!ok) { if (errors++ < 10) printf("it %d: clipped CIGAR has %zu non-H operations\n", it, core.size()); continue; }
        // known sites on reads without an indel: [max(0, s - cpos), min(Lk - 1, e - cpos)] vs getReadCoordinateForReferenceCoordinate on the clipped CIGAR
        if (cs.kind == 0) for (int t = 0; t < 3; t++) {
            const int s = cs.cpos - 3 + below(Lk + 6), e = s + below(12);
            if (e < cs.cpos || s > cs.cpos + Lk - 1) continue;      // intervals.Intersect only returns overlapping sites
            int ok1, ok2;
            int fs = orc_probe_readcoord(nc, rc, cs.cpos, s, 0, &ok1); if (!ok1 || fs < 0) fs = 0;
            int fe = orc_probe_readcoord(nc, rc, cs.cpos, e, 0, &ok2); if (!ok2 || fe > Lk - 1) fe = Lk - 1;
            int cfs = s - cs.cpos, cfe = e - cs.cpos; if (cfs < 0) cfs = 0; if (cfe > Lk - 1) cfe = Lk - 1;
            n_sites++;
            if (fs != cfs || fe != cfe) { if (errors++ < 10) printf("it %d: site [%d,%d] -> read [%d,%d], oracle [%d,%d]\n", it, s, e, cfs, cfe, fs, fe); }
        }
    }
    printf("simple %ld (adaptor-clipped or soft-clipped away from the ends: %ld) indel %ld dropped %ld general %ld sites %ld: %d errors\n%s\n", n_simple, n_adaptor, n_indel, n_drop, n_general, n_sites, errors, errors ? "FAILED" : "OK");
    return errors ? 1 : 0;
}
