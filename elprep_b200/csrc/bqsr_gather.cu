// bqsr_gather.cu -- BQSR covariate gather on the device (replaces (*BaseRecalibrator).Recalibrate,
// filters/bqsr.go:467-551, with the read clipping of filters/utils.go:130-534).
//
// One warp per read, in output (coordinate) order so that reference bases are fetched from HBM once and re-used
// through L2.  Lane 0 performs the (short, serial) CIGAR surgery of hardClipAdaptorSequence /
// hardClipSoftClippedBases on a shared-memory copy of the CIGAR; all 32 lanes then walk the kept bases:
// mismatch vs reference (computeSnpEvents :254-285), known-sites mask (calculateSkipSlice :389-414), cycle
// (:376-387) and 2-mer context (:64-146,312-362) covariates, and the three integer tables.
// Table layout: dense int64 [n_cov][94][1 + (2*max_cycle+1) + 16][2] = (observations, mismatches); the
// QualityScores column is derived as the row sum of the Cycles columns (every counted base updates both).
#include <algorithm>
#include <vector>
#include "ctx.h"

namespace {

constexpr int MAXC = 64;        // CIGAR operations per read handled by the kernel
constexpr int MAXIT = 16;       // 32*MAXIT = 512 bases per clipped read (cycles beyond max_cycle=500 are an error anyway)
constexpr int WARPS_PER_BLOCK = 8;

__device__ __forceinline__ int op_of(uint32_t c) { return (int)(c & 15); }
__device__ __forceinline__ int len_of(uint32_t c) { return (int)(c >> 4); }
__device__ __forceinline__ uint32_t mk(int len, int op) { return ((uint32_t)len << 4) | (uint32_t)op; }
// BAM op codes: M0 I1 D2 N3 S4 H5 P6 =7 X8
__device__ __forceinline__ int cons_read(int o) { return o == 0 || o == 1 || o == 4 || o == 7 || o == 8; }
__device__ __forceinline__ int cons_ref(int o) { return o == 0 || o == 2 || o == 3 || o == 7 || o == 8; }

struct Clip {          // working copy of one alignment (lane 0 only)
    int32_t pos; int nc; int s0, slen; int err;
    uint32_t* cg;      // shared memory, MAXC+4 entries
    uint32_t* tmp;     // shared memory, MAXC+4 entries
};

__device__ int32_t aln_end(const Clip& a) { int32_t l = 0; for (int i = 0; i < a.nc; i++) l += cons_ref(op_of(a.cg[i])) * len_of(a.cg[i]); return a.pos + l - 1; }
__device__ int soft_start(const Clip& a) { int32_t s = a.pos; for (int i = 0; i < a.nc; i++) { int o = op_of(a.cg[i]); if (o == 4) s -= len_of(a.cg[i]); else if (o != 5) break; } return s; }
__device__ int soft_end(const Clip& a) {
    int32_t end = aln_end(a), se = end;
    for (int i = a.nc - 1; i >= 0; i--) { int o = op_of(a.cg[i]); if (o == 4) se += len_of(a.cg[i]); else if (o != 5) return se; }
    return end;
}
__device__ int read_len(const uint32_t* cg, int nc) { int l = 0; for (int i = 0; i < nc; i++) l += cons_read(op_of(cg[i])) * len_of(cg[i]); return l; }

// computeReadCoordinateForReferenceCoordinate, filters/utils.go:267-326
__device__ int compute_read_coord(const uint32_t* cv, int nc, int softStart, int refIndex, int* falls) {
    const int goal = refIndex - softStart;
    *falls = 0;
    if (goal < 0) return -1;
    int readBases = 0, refBases = 0, fallsInside = 0, endsJustBefore = 0, fob = 0, index = 0;
    while (refBases != goal && index < nc) {
        const uint32_t el = cv[index]; index++;
        const int eo = op_of(el), elen = len_of(el);
        int shift = 0;
        if (cons_ref(eo) || eo == 4) { shift = (refBases + elen < goal) ? elen : goal - refBases; refBases += shift; }
        if (refBases != goal) readBases += cons_read(eo) * elen;
        else {
            if (shift >= elen && index == nc) return -1;
            int no = -1;
            if (shift < elen) fallsInside = (eo == 2 || eo == 3);
            else {
                uint32_t nx = cv[index]; index++;
                if (op_of(nx) == 1) { readBases += len_of(nx); if (index == nc) return -1; nx = cv[index]; index++; }
                no = op_of(nx);
                endsJustBefore = (no == 2 || no == 3);
            }
            fob = endsJustBefore || fallsInside;
            if (!fob) readBases += cons_read(eo) * shift;
            else if (endsJustBefore) readBases += cons_read(eo) * (shift - 1);
            else if (fallsInside || (endsJustBefore && (no == 2 || no == 3))) readBases--;
        }
    }
    if (refBases != goal) return -1;
    *falls = fob;
    return readBases;
}
// getReadCoordinateForReferenceCoordinate, filters/utils.go:335-349 (+ readStartsWithInsertion, bqsr.go:287-299)
__device__ int get_read_coord(const uint32_t* cv, int nc, int softStart, int refIndex, bool tail_right, bool* ok) {
    int falls; int rb = compute_read_coord(cv, nc, softStart, refIndex, &falls);
    if (rb == -1) { *ok = false; return -1; }
    if (tail_right && falls) rb++;
    if (!tail_right && rb == 0) {
        for (int i = 0; i < nc; i++) {
            const int o = op_of(cv[i]);
            if (o == 1) { const int fl = len_of(cv[i]), m = read_len(cv, nc) - 1; rb = fl < m ? fl : m; break; }
            if (o == 5 || o == 4) continue;
            break;
        }
    }
    *ok = true; return rb;
}
__device__ int hard_soft_offset(const uint32_t* c, int nc) {   // utils.go:351-371
    int size = 0, i = 0;
    for (; i < nc; i++) { if (op_of(c[i]) == 5) size += len_of(c[i]); else break; }
    for (; i < nc; i++) { if (op_of(c[i]) == 4) size += len_of(c[i]); else break; }
    return size;
}
__device__ __forceinline__ int clip_shift(uint32_t op, int cigarLength) {   // utils.go:377-386
    const int o = op_of(op);
    if (o == 1) return -cigarLength;
    if (o == 2 || o == 3) return len_of(op);
    return 0;
}
__device__ int clean_hard_clipped(uint32_t* c, int nc) {   // utils.go:473-504
    int total = 0, index = 0;
    for (; index < nc; index++) { const int o = op_of(c[index]); if (o == 5 || o == 2 || o == 3) total += len_of(c[index]); else break; }
    if (index > 0) { c[0] = mk(total, 5); for (int k = index; k < nc; k++) c[1 + k - index] = c[k]; nc = 1 + nc - index; }
    total = 0; index = nc - 1;
    for (; index >= 0; index--) { const int o = op_of(c[index]); if (o == 5 || o == 2 || o == 3) total += len_of(c[index]); else break; }
    if (index < nc - 1) { c[index + 1] = mk(total, 5); nc = index + 2; }
    return nc;
}
// hardClipCigar, utils.go:407-471: writes into a.tmp, returns the new op count
__device__ int hard_clip_cigar(const Clip& a, int start, int stop) {
    const uint32_t* cv = a.cg; const int nc = a.nc; uint32_t* out = a.tmp;
    int index = 0, total = stop - start + 1, ashift = 0, no = 0;
    if (start == 0) {
        int ci = 0;
        for (int k = 0; k < nc; k++) { ci = k; if (op_of(cv[k]) != 5) break; total += len_of(cv[k]); }
        for (; index <= stop && ci < nc; ci++) {
            const uint32_t op = cv[ci]; const int L = len_of(op), shift = cons_read(op_of(op)) * L;
            if (index + shift == stop + 1) { ashift += clip_shift(op, L); out[no++] = mk(total + ashift, 5); }
            else if (index + shift > stop + 1) {
                const int after = L - (stop - index + 1);
                ashift += clip_shift(op, stop - index + 1);
                out[no++] = mk(total + ashift, 5); out[no++] = mk(after, op_of(op));
            }
            index += shift;
            ashift += clip_shift(op, shift);
        }
        for (; ci < nc; ci++) out[no++] = cv[ci];
    } else {
        int ci = 0;
        for (; index < start && ci < nc; ci++) {
            const uint32_t op = cv[ci]; const int L = len_of(op), shift = cons_read(op_of(op)) * L;
            if (index + shift < start) out[no++] = op;
            else {
                const int after = start - index;
                ashift += clip_shift(op, L - (start - index));
                if (op_of(op) == 5) total += after; else out[no++] = mk(after, op_of(op));
            }
            index += shift;
        }
        for (; ci < nc; ci++) { const uint32_t op = cv[ci]; ashift += clip_shift(op, len_of(op)); if (op_of(op) == 5) total += len_of(op); }
        out[no++] = mk(total + ashift, 5);
    }
    return clean_hard_clipped(out, no);
}
// hardClip, utils.go:388-405 (the read is mapped here, so POS always shifts for left clips)
__device__ void hard_clip(Clip& a, int start, int stop) {
    const int ncl = hard_clip_cigar(a, start, stop);
    const int readLength = a.slen, newLength = readLength - (stop - start + 1);
    const int copyStart = (start == 0) ? stop + 1 : 0;
    if (newLength < 0 || copyStart + newLength > readLength) { a.err = 1; a.slen = 0; return; }
    const int shift = hard_soft_offset(a.tmp, ncl) - hard_soft_offset(a.cg, a.nc);
    a.s0 += copyStart; a.slen = newLength;
    for (int k = 0; k < ncl; k++) a.cg[k] = a.tmp[k];
    a.nc = ncl;
    if (start == 0) a.pos += shift;
}

struct GatherArgs {
    uint64_t n;
    const int32_t *refid, *pos, *nref, *pnext, *tlen, *rg, *lseq; const uint16_t* flag; const uint8_t* mapq;
    const uint64_t *qual_off, *seq_off, *cigar_off; const uint32_t* ncigar;
    const uint32_t* cigar; const uint8_t *seq, *qual;
    const int32_t* rg_cov; int n_rg;
    const int32_t* contig_len; int n_contigs;
    const uint8_t* const* ref; const uint64_t* ref_len;
    const int32_t* const* sites; const uint64_t* n_sites;
    TableGeom geom; unsigned long long* tables; uint32_t* err;
    // shared-memory privatisation: observation counters of the frequent QUAL values live in shared memory
    int8_t qslot[94]; uint8_t slot_q[94]; int n_slots, Lc, ncols_s;
};

__device__ __forceinline__ int nib_at(const uint8_t* seq, uint64_t soff, int i) { const uint8_t b = seq[soff + (uint64_t)(i >> 1)]; return (i & 1) ? (b & 15) : (b >> 4); }
__device__ __forceinline__ int nib_index(int nib) { return nib == 1 ? 0 : nib == 2 ? 1 : nib == 4 ? 2 : nib == 8 ? 3 : -1; }   // A C G T, else -1
__device__ __forceinline__ int ref_class(uint8_t b) {   // baseToIntMap, bqsr.go:247-252 (0 for N and IUPAC codes)
    switch (b) { case 'a': case 'A': case '*': return 1; case 'c': case 'C': return 2; case 'g': case 'G': return 3; case 't': case 'T': return 4; }
    return 0;
}

// observation / mismatch update of one table cell: hot QUAL values go to the CTA's shared-memory table (flushed once at
// the end), everything else and all (rare) mismatch counts go straight to the global int64 table
__device__ __forceinline__ void count_cell(const GatherArgs& A, uint32_t* sm_tab, int cov, int q, int col_s, int col_g, int snp) {
    const int slot = A.qslot[q];
    if (slot >= 0) atomicAdd(&sm_tab[((size_t)cov * A.n_slots + slot) * A.ncols_s + col_s], 1u);
    else atomicAdd(A.tables + 2 * A.geom.idx(cov, q, col_g), 1ull);
    if (snp) atomicAdd(A.tables + 2 * A.geom.idx(cov, q, col_g) + 1, 1ull);
}

__device__ __forceinline__ void gather_read(const GatherArgs& A, const uint64_t k, const unsigned lane, const unsigned w, uint32_t* cg, uint32_t* tmp_cg,
                                            int* sfs, int* sfe, uint32_t* sm_tab) {
    // ---- recalibrateAln, bqsr.go:225-244 (all lanes evaluate the same scalars) ----
    const uint16_t f = A.flag[k];
    const uint8_t mq = A.mapq[k];
    const int32_t refid = A.refid[k], pos0 = A.pos[k], g = A.rg[k], L0 = A.lseq[k];
    const int nc0 = (int)A.ncigar[k];
    if (!(mq > 0 && mq < 255)) return;
    if (f & (F_SECONDARY | F_DUPLICATE | F_QCFAILED)) return;
    if ((f & F_UNMAPPED) || refid < 0 || pos0 == 0) return;          // isStrictUnmapped, utils.go:140
    if (pos0 <= 0 || L0 <= 0) return;
    if (g < 0 || g >= A.n_rg) return;                               // aln.RG() != nil
    if (refid >= A.n_contigs || pos0 > A.contig_len[refid]) return;  // alignmentAgreesWithHeader, utils.go:130-138
    int c_pos = pos0, c_nc = nc0, c_s0 = 0, c_len = L0, c_err = 0;
    const uint64_t coff = A.cigar_off[k];
    const uint32_t op0 = nc0 > 0 ? A.cigar[coff] : 0u;
    bool fast = false;
    if (nc0 == 1 && op_of(op0) == 0) {
        // ---- fast path: a single M operation (the common case). No soft clips; the adaptor boundary test of
        // hardClipAdaptorSequence (utils.go:148-222) and the resulting clip have closed forms here ----
        if (len_of(op0) != L0) return;                                   // SEQ length != read length from the CIGAR
        const int32_t pnext = A.pnext[k], tlen = A.tlen[k], nref = A.nref[k];
        const bool next_unmapped = (f & F_NEXTUNMAPPED) || nref < 0 || pnext == 0;
        if (tlen != 0 && (f & F_MULTIPLE) && !next_unmapped && (((f & F_REVERSED) != 0) != ((f & F_NEXTREVERSED) != 0))) {
            const int alnEnd = pos0 + L0 - 1;
            const bool well = (f & F_REVERSED) ? (alnEnd > pnext) : (pos0 <= pnext + tlen);
            const int boundary = (f & F_REVERSED) ? (int)pnext - 1 : (int)pos0 + (tlen < 0 ? -tlen : tlen);
            if (well && boundary >= pos0 && boundary <= alnEnd) {
                const int goal = boundary - pos0;                          // read coordinate of the boundary (utils.go:267-349 for one M op)
                if (f & F_REVERSED) { c_s0 = goal + 1; c_len = L0 - goal - 1; c_pos = pos0 + goal + 1; }   // hardClip(0, goal)
                else { c_len = goal; }                                     // hardClip(goal, L-1); goal == 0 clips everything
            }
        }
        if (c_len <= 0) return;
        if (lane == 0) cg[0] = mk(c_len, 0);
        c_nc = 1;
        fast = true;
        __syncwarp();
    }
    if (!fast) {
        if (nc0 > MAXC) { if (lane == 0) atomicOr(A.err, DERR_CIGAR_LIMIT); return; }
        for (int i = lane; i < nc0; i += 32) cg[i] = A.cigar[coff + i];
        __syncwarp();
        // no N operation, SEQ length == read length from the CIGAR
        int bad = 0, rl = 0;
        for (int i = lane; i < nc0; i += 32) { const int o = op_of(cg[i]); bad |= (o == 3); rl += cons_read(o) * len_of(cg[i]); }
        for (int o = 16; o; o >>= 1) { bad |= __shfl_xor_sync(FULL_MASK, bad, o); rl += __shfl_xor_sync(FULL_MASK, rl, o); }
        if (bad || rl != L0) return;

        // ---- clipping on lane 0 ----
        if (lane == 0) {
            Clip a; a.pos = pos0; a.nc = nc0; a.s0 = 0; a.slen = L0; a.err = 0; a.cg = cg; a.tmp = tmp_cg;
            const int32_t pnext = A.pnext[k], tlen = A.tlen[k], nref = A.nref[k];
            // hardClipAdaptorSequence, utils.go:148-222
            bool well = false; int alnEnd = -1;
            const bool next_unmapped = (f & F_NEXTUNMAPPED) || nref < 0 || pnext == 0;   // isStrictNextUnmapped, utils.go:144
            if (tlen != 0 && (f & F_MULTIPLE) && !next_unmapped && (((f & F_REVERSED) != 0) != ((f & F_NEXTREVERSED) != 0))) {
                if (f & F_REVERSED) { alnEnd = aln_end(a); well = alnEnd > pnext; }
                else well = pos0 <= pnext + tlen;
            }
            if (well) {
                const int boundary = (f & F_REVERSED) ? (int)pnext - 1 : (int)pos0 + (tlen < 0 ? -tlen : tlen);
                if (boundary >= pos0) {
                    if (alnEnd < 0) alnEnd = aln_end(a);
                    if (boundary <= alnEnd) {
                        bool ok;
                        if (f & F_REVERSED) { const int stop = get_read_coord(a.cg, a.nc, soft_start(a), boundary, false, &ok); if (!ok) a.err = 2; else hard_clip(a, 0, stop); }
                        else { const int start = get_read_coord(a.cg, a.nc, soft_start(a), boundary, true, &ok); if (!ok) a.err = 2; else hard_clip(a, start, a.slen - 1); }
                    }
                }
            }
            // hardClipSoftClippedBases, utils.go:506-534
            if (!a.err && a.slen > 0) {
                int readIndex = 0, cutLeft = -1, cutRight = -1; bool rightTail = false;
                for (int i = 0; i < a.nc; i++) {
                    const int o = op_of(a.cg[i]), ln = len_of(a.cg[i]);
                    if (o == 4) { if (rightTail) cutRight = readIndex; else cutLeft = readIndex + ln - 1; }
                    else if (o != 5) rightTail = true;
                    readIndex += cons_read(o) * ln;
                }
                if (cutRight >= 0) hard_clip(a, cutRight, a.slen - 1);
                if (!a.err && a.slen > 0 && cutLeft >= 0) hard_clip(a, 0, cutLeft);
            }
            c_pos = a.pos; c_nc = a.nc; c_s0 = a.s0; c_len = a.slen; c_err = a.err;
        }
    c_pos = __shfl_sync(FULL_MASK, c_pos, 0); c_nc = __shfl_sync(FULL_MASK, c_nc, 0); c_s0 = __shfl_sync(FULL_MASK, c_s0, 0);
        c_len = __shfl_sync(FULL_MASK, c_len, 0); c_err = __shfl_sync(FULL_MASK, c_err, 0);
        __syncwarp();
    }
    if (c_err) { if (lane == 0) atomicOr(A.err, DERR_CLIP); return; }
    if (c_len == 0) return;
    if (c_len > 32 * MAXIT) { if (lane == 0) atomicOr(A.err, DERR_READLEN_LIMIT); return; }
    const int L = c_len;
    const uint64_t qoff = A.qual_off[k] + (uint64_t)c_s0, soff = A.seq_off[k];
    const int cov = A.rg_cov[g];

    // ---- low-quality tails (computeStrandedClippedSeq, bqsr.go:312-331): first / last base with QUAL > 2, via ballots ----
    const int nit = (L + 31) >> 5;
    int leftPos = L, rightPos = -1;
    for (int it = 0; it < nit; it++) {
        const int i = lane + it * 32;
        const unsigned b = __ballot_sync(FULL_MASK, i < L && A.qual[qoff + i] > 2);
        if (b) { if (leftPos == L) leftPos = it * 32 + __ffs(b) - 1; rightPos = it * 32 + 31 - __clz(b); }
    }
    const bool have_ctx = leftPos <= rightPos;

    // ---- known sites (calculateSkipSlice, bqsr.go:389-414): clipped read has no S, so softStart/softEnd = POS / End ----
    uint32_t skipmask = 0;   // bit it: base lane+32*it is masked
    const uint32_t ns = (uint32_t)A.n_sites[refid];
    const bool single_m = (c_nc == 1 && (op_of(cg[0]) == 0 || op_of(cg[0]) == 7 || op_of(cg[0]) == 8)) ||
                          (c_nc == 2 && ((op_of(cg[0]) == 5 && op_of(cg[1]) == 0) || (op_of(cg[0]) == 0 && op_of(cg[1]) == 5))) ||
                          (c_nc == 3 && op_of(cg[0]) == 5 && op_of(cg[1]) == 0 && op_of(cg[2]) == 5);
    if (ns) {
        const int32_t* sv = A.sites[refid];
        int refl;
        if (single_m) refl = L;
        else {
            refl = 0;
            for (int i = lane; i < c_nc; i += 32) refl += cons_ref(op_of(cg[i])) * len_of(cg[i]);
            for (int o = 16; o; o >>= 1) refl += __shfl_xor_sync(FULL_MASK, refl, o);
        }
        const int ss = c_pos, se = c_pos + refl - 1;
        // intervals.Intersect (intervals/intervals.go:166-173): s0 = first interval with End >= ss (32-ary search), then the
        // run of intervals with Start <= se
        uint32_t lo = 0, hi = ns;   // invariant: End[i] < ss for i < lo, End[i] >= ss for i >= hi
        while (lo < hi) {
            const uint32_t span = hi - lo, step = (span + 31) >> 5;
            const uint32_t probe = lo + lane * step;
            const bool inr = probe < hi;
            const unsigned bm = __ballot_sync(FULL_MASK, inr && sv[2 * probe + 1] >= ss);
            const unsigned vm = __ballot_sync(FULL_MASK, inr);
            if (bm) { const uint32_t f1 = __ffs(bm) - 1; hi = lo + f1 * step; lo = f1 ? lo + (f1 - 1) * step + 1 : lo; }
            else { lo = lo + (uint32_t)(__popc(vm) - 1) * step + 1; }
            if (lo > hi) lo = hi;
        }
        const uint32_t s0 = lo;
        uint32_t s1 = s0;
        while (s1 < ns && sv[2 * s1] <= se) s1++;
        for (uint32_t sb = s0; sb < s1; sb += 32) {
            const uint32_t sidx = sb + lane;
            if (sidx < s1) {
                bool ok; int fs = get_read_coord(cg, c_nc, ss, sv[2 * sidx], false, &ok);
                if (!ok || fs < 0) fs = 0;
                int fe = get_read_coord(cg, c_nc, ss, sv[2 * sidx + 1], false, &ok);
                if (!ok || fe > L - 1) fe = L - 1;
                sfs[lane] = fs; sfe[lane] = fe;
            }
            __syncwarp();
            const int cnt = (s1 - sb) < 32 ? (int)(s1 - sb) : 32;
            for (int it = 0; it < nit; it++) {
                const int i = lane + it * 32;
                for (int q = 0; q < cnt; q++) if (i >= sfs[q] && i <= sfe[q]) skipmask |= 1u << it;
            }
            __syncwarp();
        }
    }

    // ---- per base ----
    const uint8_t* ref = A.ref[refid]; const int64_t reflen = (int64_t)A.ref_len[refid];
    const int reversed = (f & F_REVERSED) ? 1 : 0, last = (f & F_LAST) ? 1 : 0;
    const int rof = 1 - 2 * last, cf = rof + reversed * (L - 1) * rof, inc = (1 - 2 * reversed) * rof;   // prepareCycleCovariates, bqsr.go:376-383
    const int lead_h = (single_m && op_of(cg[0]) == 5) ? 1 : 0;   // index of the M op in the single-M fast path
    (void)lead_h;
    const uint32_t sm_base = (uint32_t)cov * (uint32_t)A.n_slots * (uint32_t)A.ncols_s;
    const uint8_t* seqp = A.seq + soff; const uint8_t* qualp = A.qual + qoff;
    uint32_t errbits = 0;
    const int dirn = reversed ? 1 : -1;                    // context neighbour: previous base in sequencing direction
    const int64_t j0 = (int64_t)c_pos - 1;
    for (int it = 0; it < nit; it++) {
        const int i = lane + it * 32;
        const bool in = i < L;
        const int ic = in ? i : L - 1;
        const int oi = c_s0 + ic;
        const uint32_t sb = seqp[oi >> 1];
        const uint32_t nib = (oi & 1) ? (sb & 15u) : (sb >> 4);
        const int bi = (__popc(nib) == 1) ? (__ffs(nib) - 1) : -1;      // A C G T -> 0..3, everything else -1 (bqsr.go:509)
        const int q = qualp[ic];
        // neighbour base (sequencing direction) through a shuffle; one extra load on the lane at a 32-base boundary
        int nbi = __shfl_sync(FULL_MASK, bi, (lane + dirn) & 31);
        const int ni = ic + dirn;
        if (((int)lane + dirn) < 0 || ((int)lane + dirn) > 31 || ni >= L) {
            if (ni >= 0 && ni < L) { const int on = c_s0 + ni; const uint32_t nb = seqp[on >> 1]; const uint32_t nn = (on & 1) ? (nb & 15u) : (nb >> 4); nbi = (__popc(nn) == 1) ? (__ffs(nn) - 1) : -1; }
            else nbi = -1;
        }
        const bool counted = in & !((skipmask >> it) & 1) & (bi >= 0) & (q >= 6);   // bqsr.go:506-515
        if (!counted) continue;
        if (q > 93) { errbits |= DERR_QUAL_RANGE; continue; }
        // reference position of base i (computeSnpEvents, bqsr.go:254-285)
        int64_t jj = -1;
        if (single_m) jj = j0 + ic;
        else {
            int ri = 0; int64_t j = j0;
            for (int c = 0; c < c_nc; c++) {
                const int o = op_of(cg[c]), ln = len_of(cg[c]);
                if (o == 0 || o == 7 || o == 8) { if (ic < ri + ln) { jj = j + (ic - ri); break; } ri += ln; j += ln; }
                else if (o == 2 || o == 3) j += ln;
                else if (o == 1 || o == 4) { if (ic < ri + ln) break; ri += ln; }
            }
        }
        int snp = 0;
        if (jj >= 0) {
            if (jj >= reflen) { errbits |= DERR_REFEND; continue; }
            // baseToIntMap (bqsr.go:247-252): A/a/* C/c G/g T/t are classes, everything else is class 0 (never equals an ACGT read base)
            uint32_t u = ref[jj]; if (u == '*') u = 'A';
            u &= 0xDFu;
            const uint32_t x = u - 'A';
            const bool rvalid = x < 32 && ((0x00080045u >> x) & 1);      // A, C, G, T
            uint32_t ridx = (u >> 1) & 3; ridx ^= ridx >> 1;             // A0 C1 G2 T3
            snp = !(rvalid && (int)ridx == bi);
        }
        const int cyc = cf + ic * inc;
        if (cyc > A.geom.max_cycle || cyc < -A.geom.max_cycle) { errbits |= DERR_CYCLE; continue; }
        // context: 2-mer in sequencing direction; key>>4 = prev | cur<<2 (keyFromContext, bqsr.go:64-76); tails with QUAL<=2 read as N
        const int lo_i = reversed ? ic : ni, hi_i = reversed ? ni : ic;
        const bool okc = have_ctx & (nbi >= 0) & (lo_i >= leftPos) & (hi_i <= rightPos) & (ni >= 0) & (ni < L);
        const int ctx = okc ? (reversed ? ((3 - nbi) | ((3 - bi) << 2)) : (nbi | (bi << 2))) : -1;
        const int slot = A.qslot[q];
        if (slot >= 0) {
            const uint32_t row = sm_base + (uint32_t)slot * (uint32_t)A.ncols_s;
            atomicAdd(&sm_tab[row + (uint32_t)(cyc + A.Lc)], 1u);
            if (ctx >= 0) atomicAdd(&sm_tab[row + (uint32_t)(2 * A.Lc + 1 + ctx)], 1u);
        } else {
            atomicAdd(A.tables + 2 * A.geom.idx(cov, q, A.geom.col_cycle(cyc)), 1ull);
            if (ctx >= 0) atomicAdd(A.tables + 2 * A.geom.idx(cov, q, A.geom.col_ctx(ctx)), 1ull);
        }
        if (snp) {   // mismatches are rare: straight to the global table
            atomicAdd(A.tables + 2 * A.geom.idx(cov, q, A.geom.col_cycle(cyc)) + 1, 1ull);
            if (ctx >= 0) atomicAdd(A.tables + 2 * A.geom.idx(cov, q, A.geom.col_ctx(ctx)) + 1, 1ull);
        }
    }
    for (int o = 16; o; o >>= 1) errbits |= __shfl_xor_sync(FULL_MASK, errbits, o);
    if (errbits && lane == 0) atomicOr(A.err, errbits);
}

// persistent kernel: every warp walks reads k = warp, warp + W, ... in output order (W consecutive reads are in flight chip-wide,
// so reference bases stay L2-resident); the CTA's shared-memory table is flushed once with 64-bit global atomics
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) bqsr_gather_kernel(GatherArgs A) {
    extern __shared__ uint32_t sm_tab[];
    __shared__ uint32_t sh_cg[WARPS_PER_BLOCK][MAXC + 4];
    __shared__ uint32_t sh_tmp[WARPS_PER_BLOCK][MAXC + 4];
    __shared__ int sh_fs[WARPS_PER_BLOCK][32], sh_fe[WARPS_PER_BLOCK][32];
    const unsigned lane = lane_id(), w = threadIdx.x >> 5;
    const int cells = A.geom.n_cov * A.n_slots * A.ncols_s;
    for (int i = threadIdx.x; i < cells; i += blockDim.x) sm_tab[i] = 0;
    __syncthreads();
    for (uint64_t k = (uint64_t)blockIdx.x * WARPS_PER_BLOCK + w; k < A.n; k += (uint64_t)gridDim.x * WARPS_PER_BLOCK) {
        gather_read(A, k, lane, w, sh_cg[w], sh_tmp[w], sh_fs[w], sh_fe[w], sm_tab);
        __syncwarp();
    }
    __syncthreads();
    for (int i = threadIdx.x; i < cells; i += blockDim.x) {
        const uint32_t v = sm_tab[i];
        if (!v) continue;
        const int col_s = i % A.ncols_s, cs = i / A.ncols_s, slot = cs % A.n_slots, cov = cs / A.n_slots;
        const int col_g = col_s < 2 * A.Lc + 1 ? A.geom.col_cycle(col_s - A.Lc) : A.geom.col_ctx(col_s - (2 * A.Lc + 1));
        atomicAdd(A.tables + 2 * A.geom.idx(cov, A.slot_q[slot], col_g), (unsigned long long)v);
    }
}

// QUAL value histogram of a prefix of the QUAL arena: picks which values get shared-memory slots
__global__ void __launch_bounds__(256) qual_sample_kernel(const uint8_t* __restrict__ qual, uint64_t n, uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) atomicAdd(&sh[qual[i]], 1u);
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}

// QualityScores[(rg,q)] = sum over cycles of Cycles[(rg,q,cycle)] (each counted base updates both, bqsr.go:518-529)
__global__ void derive_q_kernel(TableGeom geom, long long* tables) {
    const int row = blockIdx.x;   // cov*94 + q
    long long obs = 0, mis = 0;
    const size_t base = (size_t)row * geom.ncols();
    for (int c = 1 + threadIdx.x; c < 1 + 2 * geom.max_cycle + 1; c += blockDim.x) { obs += tables[2 * (base + c)]; mis += tables[2 * (base + c) + 1]; }
    __shared__ long long so[8], sm[8];
    for (int o = 16; o; o >>= 1) { obs += __shfl_xor_sync(FULL_MASK, obs, o); mis += __shfl_xor_sync(FULL_MASK, mis, o); }
    if ((threadIdx.x & 31) == 0) { so[threadIdx.x >> 5] = obs; sm[threadIdx.x >> 5] = mis; }
    __syncthreads();
    if (threadIdx.x == 0) { long long a = 0, b = 0; for (int i = 0; i < (int)(blockDim.x >> 5); i++) { a += so[i]; b += sm[i]; } tables[2 * base] = a; tables[2 * base + 1] = b; }
}

}  // namespace

int phase_bqsr_gather(elp_ctx* c) {
    if (!c->sorted) return c->fail(E_STATE, "elp_bqsr_gather called before elp_sort_markdup");
    int rc = upload_side_inputs(c);
    if (rc) return rc;
    for (int i = 0; i < c->n_contigs; i++) if (!c->d_ref[i]) {
        // a contig without reference bases is only an error if a read maps to it; keep it simple and require all of them
    }
    const size_t cells = c->geom.cells();
    CUDA_TRY(c, cudaMemsetAsync(c->d_tables, 0, cells * 2 * sizeof(int64_t), c->stream));
    const uint64_t n = c->n;
    if (n) {
        GatherArgs A{};
        A.n = n; A.refid = c->s_refid.p; A.pos = c->s_pos.p; A.nref = c->s_nref.p; A.pnext = c->s_pnext.p; A.tlen = c->s_tlen.p; A.rg = c->s_rg.p; A.lseq = c->s_lseq.p;
        A.flag = c->s_flag.p; A.mapq = c->s_mapq.p; A.qual_off = c->s_qual_off.p; A.seq_off = c->s_seq_off.p; A.cigar_off = c->s_cigar_off.p; A.ncigar = c->s_ncigar.p;
        A.cigar = c->cigar.p; A.seq = c->seq.p; A.qual = c->qual.p; A.rg_cov = c->d_rg_cov; A.n_rg = c->n_rg; A.contig_len = c->d_contig_len; A.n_contigs = c->n_contigs;
        A.ref = c->d_ref_ptrs; A.ref_len = c->d_ref_len; A.sites = c->d_site_ptrs; A.n_sites = c->d_n_sites;
        A.geom = c->geom; A.tables = reinterpret_cast<unsigned long long*>(c->d_tables); A.err = c->d_err;
        // the gather needs up-to-date duplicate flags in output order: s_flag was gathered after duplicate marking
        uint64_t ref_bytes = 0; for (auto l : c->ref_len) ref_bytes += l;
        const double bytes = (double)n * (19 + 4 + 8 + 8) + (double)c->n_cigar * 4 + (double)c->n_seq + (double)c->n_qual + (double)ref_bytes;
        // slot map: the most frequent QUAL values >= 6 of a sample get shared-memory counters (<= 48 KB per CTA)
        const int Lc = std::max(1, std::min(c->max_cycle, c->h_ranges.lseq_max));
        A.Lc = Lc; A.ncols_s = 2 * Lc + 1 + 16;
        for (int q = 0; q < 94; q++) { A.qslot[q] = -1; A.slot_q[q] = 0; }
        {
            CUDA_TRY(c, c->scan_tmp.reserve(256 + 4, c->stream));
            CUDA_TRY(c, cudaMemsetAsync(c->scan_tmp.p, 0, 256 * 4, c->stream));
            const uint64_t ns = std::min<uint64_t>(c->n_qual, 8u << 20);
            c->begin("qual_sample", (double)ns);
            qual_sample_kernel<<<64, 256, 0, c->stream>>>(c->qual.p, ns, c->scan_tmp.p);
            c->end(); LAUNCH_CHECK(c);
            uint32_t h[256];
            CUDA_TRY(c, cudaMemcpyAsync(h, c->scan_tmp.p, sizeof h, cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(c, cudaStreamSynchronize(c->stream));
            std::vector<int> qs;
            for (int q = 6; q < 94; q++) if (h[q]) qs.push_back(q);
            std::sort(qs.begin(), qs.end(), [&](int a, int b) { return h[a] != h[b] ? h[a] > h[b] : a < b; });
            const size_t per_slot = (size_t)std::max(1, c->geom.n_cov) * A.ncols_s * 4;
            const int max_slots = (int)std::min<size_t>(94, (48 * 1024) / per_slot);
            A.n_slots = std::min<int>((int)qs.size(), max_slots);
            for (int s = 0; s < A.n_slots; s++) { A.qslot[qs[s]] = (int8_t)s; A.slot_q[s] = (uint8_t)qs[s]; }
        }
        const size_t smem = (size_t)c->geom.n_cov * A.n_slots * A.ncols_s * 4;
        int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
        uint64_t grid = std::min<uint64_t>((n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK, (uint64_t)sms * 4);
        grid = std::max<uint64_t>(grid, (n + (4u << 20) - 1) / (4u << 20));   // <= 4 M reads per CTA keeps the 32-bit shared counters far from overflow
        CUDA_TRY(c, cudaFuncSetAttribute(bqsr_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));
        c->begin("bqsr_gather", bytes);
        bqsr_gather_kernel<<<(unsigned)grid, WARPS_PER_BLOCK * 32, smem, c->stream>>>(A);
        c->end(); LAUNCH_CHECK(c);
    }
    c->begin("bqsr_derive_q", 0);
    derive_q_kernel<<<c->geom.n_cov * 94, 256, 0, c->stream>>>(c->geom, reinterpret_cast<long long*>(c->d_tables));
    c->end(); LAUNCH_CHECK(c);
    rc = check_device_errors(c);
    if (rc) return rc;
    c->gathered = true; c->finalized = false;
    return E_OK;
}
