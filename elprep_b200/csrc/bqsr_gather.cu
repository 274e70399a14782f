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
};

__device__ __forceinline__ int nib_at(const uint8_t* seq, uint64_t soff, int i) { const uint8_t b = seq[soff + (uint64_t)(i >> 1)]; return (i & 1) ? (b & 15) : (b >> 4); }
__device__ __forceinline__ int nib_index(int nib) { return nib == 1 ? 0 : nib == 2 ? 1 : nib == 4 ? 2 : nib == 8 ? 3 : -1; }   // A C G T, else -1
__device__ __forceinline__ int ref_class(uint8_t b) {   // baseToIntMap, bqsr.go:247-252 (0 for N and IUPAC codes)
    switch (b) { case 'a': case 'A': case '*': return 1; case 'c': case 'C': return 2; case 'g': case 'G': return 3; case 't': case 'T': return 4; }
    return 0;
}

__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) bqsr_gather_kernel(GatherArgs A) {
    __shared__ uint32_t sh_cg[WARPS_PER_BLOCK][MAXC + 4];
    __shared__ uint32_t sh_tmp[WARPS_PER_BLOCK][MAXC + 4];
    __shared__ int sh_fs[WARPS_PER_BLOCK][32], sh_fe[WARPS_PER_BLOCK][32];
    const unsigned lane = lane_id(), w = threadIdx.x >> 5;
    const uint64_t k = (uint64_t)blockIdx.x * WARPS_PER_BLOCK + w;
    if (k >= A.n) return;
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
    if (nc0 > MAXC) { if (lane == 0) atomicOr(A.err, DERR_CIGAR_LIMIT); return; }
    uint32_t* cg = sh_cg[w];
    const uint64_t coff = A.cigar_off[k];
    for (int i = lane; i < nc0; i += 32) cg[i] = A.cigar[coff + i];
    __syncwarp();
    // no N operation, SEQ length == read length from the CIGAR
    int bad = 0, rl = 0;
    for (int i = lane; i < nc0; i += 32) { const int o = op_of(cg[i]); bad |= (o == 3); rl += cons_read(o) * len_of(cg[i]); }
    for (int o = 16; o; o >>= 1) { bad |= __shfl_xor_sync(FULL_MASK, bad, o); rl += __shfl_xor_sync(FULL_MASK, rl, o); }
    if (bad || rl != L0) return;

    // ---- clipping on lane 0 ----
    int c_pos = pos0, c_nc = nc0, c_s0 = 0, c_len = L0, c_err = 0;
    if (lane == 0) {
        Clip a; a.pos = pos0; a.nc = nc0; a.s0 = 0; a.slen = L0; a.err = 0; a.cg = cg; a.tmp = sh_tmp[w];
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
    if (c_err) { if (lane == 0) atomicOr(A.err, DERR_CLIP); return; }
    if (c_len == 0) return;
    if (c_len > 32 * MAXIT) { if (lane == 0) atomicOr(A.err, DERR_READLEN_LIMIT); return; }
    const int L = c_len;
    const uint64_t qoff = A.qual_off[k] + (uint64_t)c_s0, soff = A.seq_off[k];
    const int cov = A.rg_cov[g];

    // ---- low-quality tails (computeStrandedClippedSeq, bqsr.go:312-331) ----
    int leftPos = L, rightPos = -1;
    for (int i = lane; i < L; i += 32) if (A.qual[qoff + i] > 2) { leftPos = min(leftPos, i); rightPos = max(rightPos, i); }
    for (int o = 16; o; o >>= 1) { leftPos = min(leftPos, __shfl_xor_sync(FULL_MASK, leftPos, o)); rightPos = max(rightPos, __shfl_xor_sync(FULL_MASK, rightPos, o)); }
    const bool have_ctx = leftPos <= rightPos;

    // ---- known sites (calculateSkipSlice, bqsr.go:389-414): clipped read has no S, so softStart/softEnd = POS / End ----
    uint32_t skipmask = 0;   // bit it: base lane+32*it is masked
    {
        const int32_t* sv = A.sites[refid]; const uint64_t ns = A.n_sites[refid];
        if (ns) {
            int refl = 0;
            for (int i = lane; i < c_nc; i += 32) refl += cons_ref(op_of(cg[i])) * len_of(cg[i]);
            for (int o = 16; o; o >>= 1) refl += __shfl_xor_sync(FULL_MASK, refl, o);
            const int ss = c_pos, se = c_pos + refl - 1;
            // intervals.Intersect, intervals/intervals.go:166-173
            uint64_t a = 0, b = ns;
            while (a < b) { const uint64_t m = (a + b) >> 1; if (!(sv[2 * m + 1] >= ss)) a = m + 1; else b = m; }
            const uint64_t s0 = a;
            a = 0; b = ns;
            while (a < b) { const uint64_t m = (a + b) >> 1; if (!(sv[2 * m] > se)) a = m + 1; else b = m; }
            const uint64_t s1 = a;
            for (uint64_t sb = s0; sb < s1; sb += 32) {
                const uint64_t s = sb + lane;
                if (s < s1) {
                    bool ok; int fs = get_read_coord(cg, c_nc, ss, sv[2 * s], false, &ok);
                    if (!ok || fs < 0) fs = 0;
                    int fe = get_read_coord(cg, c_nc, ss, sv[2 * s + 1], false, &ok);
                    if (!ok || fe > L - 1) fe = L - 1;
                    sh_fs[w][lane] = fs; sh_fe[w][lane] = fe;
                }
                __syncwarp();
                const int cnt = (s1 - sb) < 32 ? (int)(s1 - sb) : 32;
                for (int it = 0; it * 32 < L; it++) {
                    const int i = lane + it * 32;
                    for (int q = 0; q < cnt; q++) if (i >= sh_fs[w][q] && i <= sh_fe[w][q]) skipmask |= 1u << it;
                }
                __syncwarp();
            }
        }
    }

    // ---- per base ----
    const uint8_t* ref = A.ref[refid]; const uint64_t reflen = A.ref_len[refid];
    const int reversed = (f & F_REVERSED) ? 1 : 0, last = (f & F_LAST) ? 1 : 0;
    const int rof = 1 - 2 * last, cf = rof + reversed * (L - 1) * rof, inc = (1 - 2 * reversed) * rof;   // prepareCycleCovariates, bqsr.go:376-383
    uint32_t errbits = 0;
    for (int it = 0; it * 32 < L; it++) {
        const int i = lane + it * 32;
        if (i >= L) break;
        const int oi = c_s0 + i;
        const int nib = nib_at(A.seq, soff, oi);
        const int bi = nib_index(nib);
        const uint8_t q = A.qual[qoff + i];
        // reference position of base i (computeSnpEvents, bqsr.go:254-285)
        int snp = 0;
        {
            int ri = 0; int64_t j = (int64_t)c_pos - 1;
            for (int c = 0; c < c_nc; c++) {
                const int o = op_of(cg[c]), ln = len_of(cg[c]);
                if (o == 0 || o == 7 || o == 8) {
                    if (i < ri + ln) {
                        if (i >= ri) {
                            const int64_t jj = j + (i - ri);
                            if (jj >= (int64_t)reflen) errbits |= DERR_REFEND;
                            else { const int rc = ref_class(ref[jj]); const int bc = bi < 0 ? 0 : bi + 1; snp = (bc != rc); }
                        }
                        break;
                    }
                    ri += ln; j += ln;
                } else if (o == 2 || o == 3) j += ln;
                else if (o == 1 || o == 4) { if (i < ri + ln) break; ri += ln; }
            }
        }
        if ((skipmask >> it) & 1) continue;
        if (bi < 0) continue;                 // bqsr.go:509
        if (q < 6) continue;                  // minInterestingQual, bqsr.go:513
        if (q > 93) { errbits |= DERR_QUAL_RANGE; continue; }
        const int cyc = cf + i * inc;
        if (cyc > A.geom.max_cycle || cyc < -A.geom.max_cycle) { errbits |= DERR_CYCLE; continue; }
        unsigned long long* t = A.tables + 2 * A.geom.idx(cov, q, A.geom.col_cycle(cyc));
        atomicAdd(t, 1ull);
        if (snp) atomicAdd(t + 1, 1ull);
        // context: 2-mer in sequencing direction; key>>4 = prev | cur<<2 (keyFromContext, bqsr.go:64-76)
        if (have_ctx) {
            int ctx = -1;
            if (!reversed) {
                if (i >= 1 && i - 1 >= leftPos && i <= rightPos) { const int pb = nib_index(nib_at(A.seq, soff, oi - 1)); if (pb >= 0) ctx = pb | (bi << 2); }
            } else {
                if (i + 1 <= L - 1 && i >= leftPos && i + 1 <= rightPos) { const int nb = nib_index(nib_at(A.seq, soff, oi + 1)); if (nb >= 0) ctx = (3 - nb) | ((3 - bi) << 2); }
            }
            if (ctx >= 0) {
                unsigned long long* tx = A.tables + 2 * A.geom.idx(cov, q, A.geom.col_ctx(ctx));
                atomicAdd(tx, 1ull);
                if (snp) atomicAdd(tx + 1, 1ull);
            }
        }
    }
    for (int o = 16; o; o >>= 1) errbits |= __shfl_xor_sync(FULL_MASK, errbits, o);
    if (errbits && lane == 0) atomicOr(A.err, errbits);
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
        c->begin("bqsr_gather", bytes);
        bqsr_gather_kernel<<<(unsigned)((n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK), WARPS_PER_BLOCK * 32, 0, c->stream>>>(A);
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
