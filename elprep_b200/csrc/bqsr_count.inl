// bqsr_count.inl -- the covariate histogram of BQSR as register-resident counters (included by bqsr_gather.cu, inside its
// anonymous namespace).  Replaces the shared-memory-atomic chunk kernel for inputs whose QUAL alphabet is small (every
// current Illumina instrument bins QUAL to <= 8 values, <= 4 of them >= 6): (*BaseRecalibrator).Recalibrate, filters/bqsr.go:467-551.
//
// Why: the per-base shared-memory atomics of the old kernel cost ~7 wavefronts of the SM's single LSU pipe per base step and
// ~50 warp instructions per base; both are far above what the 253 B/read of HBM traffic allow.  Here a lane owns 32 consecutive
// bases of a read IN SEQUENCING ORDER and keeps everything as one-bit-per-base planes (bqsr_lane.cuh):
//   * Cycles table: the cycle of a lane's bit position is fixed (cycle = +-(32 c + t + 1)), so per QUAL slot the lane adds its
//     32-bit "counted" plane into an 8-plane bit-sliced (vertical) counter -- 16 LOP3 per slot and pass, no memory traffic; the
//     planes are flushed to the global table with 64-bit atomics every <= 255 passes.
//   * Contexts table: position independent, so 16 contexts x S slots are popcounts of three-input ANDs of the planes, summed in
//     packed 16-bit accumulators and flushed with one warp reduction per segment.
//   * mismatches (sparse, but piled on few cells) go to CTA-private shared-memory tables, flushed once per CTA.
// Both only work if all reads a warp sees share (read-group covariate, first/second of pair): bqsr_prep2_kernel therefore sorts
// the eligible reads into per-class lists of 32-byte work records (closed-form clipping for reads whose CIGAR is
// [H][S]M[S][H] or that plus one insertion/deletion; everything else goes to the old, general kernels through a list).
// QUAL, SEQ and reference windows are staged through shared memory with cp.async, CNT_STAGES passes deep.

#ifndef CNT_MINB
#define CNT_MINB 2
#endif
#ifndef CNT_STAGES_N
#define CNT_STAGES_N 2
#endif
// software pipeline of a warp: records are fetched CNT_D2 iterations, windows CNT_D2 - CNT_D1 iterations ahead of their use; CNT_WN copy groups stay in flight
constexpr int CNT_WARPS = 8, CNT_STAGES = CNT_STAGES_N, CNT_WN = CNT_STAGES - 1, CNT_D1 = CNT_WN + 1, CNT_D2 = 2 * CNT_WN + 1, CNT_RECRING = CNT_D2 + 1, SEG_PASSES = 255, MAX_CLS = 64;
constexpr uint32_t KEY_NONE = 0xffffffffu;

struct Prep2Args {
    int n_cls;                       // 2 * n_cov
    const uint32_t* region_base;     // [n_cls + 1] record index where the class' region starts
    uint32_t* key_count;             // [2 * n_cls + 1]: per (class, variant) records written; last: complex reads
    uint4* recs;                     // 2 x uint4 per record
    uint32_t* cx_list;               // reads for the general path
    int lpr, max_cycle;
};

// reads per (covariate, mate) -- sizes the regions of the record lists
__global__ void __launch_bounds__(256) class_hist_kernel(uint64_t n, const int32_t* __restrict__ rg, const uint16_t* __restrict__ flag, const int32_t* __restrict__ rg_cov, int n_rg,
                                                          int n_cls, uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[MAX_CLS];
    if (threadIdx.x < MAX_CLS) sh[threadIdx.x] = 0;
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const int32_t g = rg[i];
        if (g >= 0 && g < n_rg) atomicAdd(&sh[rg_cov[g] * 2 + ((flag[i] & F_LAST) ? 1 : 0)], 1u);
    }
    __syncthreads();
    if (threadIdx.x < n_cls && sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}
__global__ void class_scan_kernel(int n_cls, const uint32_t* __restrict__ hist, uint32_t* __restrict__ region_base) {
    if (threadIdx.x == 0) { uint32_t run = 0; for (int i = 0; i < n_cls; i++) { region_base[i] = run; run += hist[i]; } region_base[n_cls] = run; }
}

// one thread per read (output order): recalibrateAln eligibility (bqsr.go:225-244), then the clipping of filters/utils.go:148-534 in
// closed form for the CIGAR shapes where it IS closed form:
//   [H..][S a] M m [S b][H..]                       adaptor boundary -> read coordinate is linear (no D/N/I to fall into)
//   [H..][S a] M m1 (I|D) d M m2 [S b][H..]        without adaptor clipping and without known sites on the read
// (derivation in DESIGN.md section "gather"); every other eligible read is appended to the list of the general kernels.
__global__ void __launch_bounds__(256) bqsr_prep2_kernel(GatherArgs A, Prep2Args P) {
    __shared__ uint32_t s_cnt[2 * MAX_CLS + 1], s_base[2 * MAX_CLS + 1];
    const int nkeys = 2 * P.n_cls + 1;
    for (int i = threadIdx.x; i < nkeys; i += blockDim.x) s_cnt[i] = 0;
    __syncthreads();
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t key = KEY_NONE;
    uint64_t ra = 0, rb = 0; uint32_t r_cpos = 0, r_skip0 = 0xffffu, r_skip1 = 0xffffu, r_indel = 0;
    if (k < A.n) {
        const uint16_t f = A.flag[k];
        const uint8_t mq = A.mapq[k];
        const int32_t refid = A.refid[k], pos0 = A.pos[k], g = A.rg[k], L0 = A.lseq[k];
        const int nc0 = (int)A.ncigar[k];
        bool elig = !(A.optf[k] & 1u) && (mq > 0 && mq < 255) && !(f & (F_SECONDARY | F_DUPLICATE | F_QCFAILED)) && !((f & F_UNMAPPED) || refid < 0 || pos0 == 0) && pos0 > 0 && L0 > 0 &&
                    g >= 0 && g < A.n_rg && refid < A.n_contigs;
        if (elig && pos0 > A.contig_len[refid]) elig = false;
        if (elig) {
            key = 2 * P.n_cls;   // general path unless the shape below matches
            const uint64_t coff = A.cigar_off[k];
            const lanes::ClipShape cs = lanes::closed_form_clip((uint32_t)f, pos0, A.pnext[k], A.tlen[k], A.nref[k], L0, nc0, [&](int i) { return __ldg(A.cigar + coff + i); });
            bool shape = cs.kind == 0 || cs.kind == 1;
            if (cs.kind < 0) key = KEY_NONE;
            const int dop = cs.kind == 1 ? (cs.ins ? 1 : 2) : -1, ins = cs.ins, del = cs.del, lo = cs.lo, m1 = cs.bp;
            const int32_t cpos = cs.cpos;
            const int Lk = cs.hi - cs.lo;
            if (shape) {
                const int kept_ref = Lk - ins + del;
                if (Lk > P.max_cycle || Lk > 32 * P.lpr || Lk > 2047 || ins > 255 || del > 4095 || refid >= (1 << 23) ||
                    (uint64_t)(cpos - 1) + (uint64_t)kept_ref > A.ref_len[refid]) shape = false;
            }
            {
                // ---- known sites (calculateSkipSlice, bqsr.go:389-414): read coordinates are linear without an indel ----
                if (shape) {
                    const uint64_t ns = A.n_sites[refid];
                    if (ns) {
                        const int32_t* sv = A.sites[refid];
                        const int ss = cpos, se = cpos + (Lk - ins + del) - 1;
                        uint64_t l = 0, h = ns;
                        while (l < h) { const uint64_t m = (l + h) >> 1; if (!(__ldg(sv + 2 * m + 1) >= ss)) l = m + 1; else h = m; }
                        uint64_t s1 = l; int nsk = 0;
                        while (s1 < ns && __ldg(sv + 2 * s1) <= se) {
                            if (dop >= 0 || nsk == 2) { shape = false; break; }
                            int fs = __ldg(sv + 2 * s1) - cpos, fe = __ldg(sv + 2 * s1 + 1) - cpos;
                            if (fs < 0) fs = 0;
                            if (fe > Lk - 1) fe = Lk - 1;
                            if (fs <= fe) { const uint32_t w = (uint32_t)fs | ((uint32_t)fe << 16); if (nsk == 0) r_skip0 = w; else r_skip1 = w; nsk++; }
                            s1++;
                        }
                    }
                }
                if (shape) {
                    const int cov = A.rg_cov[g], mate = (f & F_LAST) ? 1 : 0;
                    const int variant = dop >= 0 ? 1 : 0;
                    key = (uint32_t)((cov * 2 + mate) * 2 + variant);
                    ra = (A.qual_off[k] + (uint64_t)lo) | ((uint64_t)Lk << 40) | ((uint64_t)((f & F_REVERSED) ? 1 : 0) << 51);
                    rb = (A.seq_off[k] * 2 + (uint64_t)lo) | ((uint64_t)(uint32_t)refid << 41);
                    r_cpos = (uint32_t)cpos;
                    if (variant) r_indel = (uint32_t)(m1) | ((uint32_t)ins << 11) | ((uint32_t)((dop == 1 ? -ins : del) + 4096) << 19);   // bp is relative to the kept read (lo == a here)
                } else if (key != KEY_NONE) key = 2 * P.n_cls;
            }
        }
    }
    uint32_t rank = 0;
    if (key != KEY_NONE) rank = atomicAdd(&s_cnt[key], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < nkeys; i += blockDim.x) if (s_cnt[i]) s_base[i] = atomicAdd(P.key_count + i, s_cnt[i]);
    __syncthreads();
    if (key == KEY_NONE) return;
    const uint32_t at = s_base[key] + rank;
    if (key == (uint32_t)(2 * P.n_cls)) { P.cx_list[at] = (uint32_t)k; return; }
    const uint32_t cls = key >> 1;
    // simple records grow from the front of the class' region, indel records from its back
    const uint32_t slot = (key & 1) ? (P.region_base[cls + 1] - 1 - at) : (P.region_base[cls] + at);
    P.recs[2 * (uint64_t)slot] = make_uint4((uint32_t)ra, (uint32_t)(ra >> 32), (uint32_t)rb, (uint32_t)(rb >> 32));
    P.recs[2 * (uint64_t)slot + 1] = make_uint4(r_cpos, r_skip0, r_skip1, r_indel);
}

// segments of <= SEG_PASSES passes of one (class, variant) list: the unit a warp of the count kernel takes from the queue
__global__ void seg_build_kernel(int n_cls, int rpw, const uint32_t* __restrict__ region_base, const uint32_t* __restrict__ key_count,
                                 uint4* __restrict__ segs0, uint4* __restrict__ segs1, uint32_t* __restrict__ n_seg /*[2]*/) {
    __shared__ uint32_t first[2 * MAX_CLS];
    const uint32_t per = (uint32_t)SEG_PASSES * (uint32_t)rpw;
    if (threadIdx.x == 0) {
        uint32_t run[2] = {0, 0};
        for (int key = 0; key < 2 * n_cls; key++) { first[key] = run[key & 1]; run[key & 1] += (key_count[key] + per - 1) / per; }
        n_seg[0] = run[0]; n_seg[1] = run[1];
    }
    __syncthreads();
    for (int key = threadIdx.x; key < 2 * n_cls; key += blockDim.x) {
        const uint32_t cnt = key_count[key], cls = (uint32_t)key >> 1;
        const uint32_t rec0 = (key & 1) ? (region_base[cls + 1] - cnt) : region_base[cls];
        uint4* out = (key & 1) ? segs1 : segs0;
        for (uint32_t s = 0, done = 0; done < cnt; s++, done += per) out[first[key] + s] = make_uint4(rec0 + done, min(per, cnt - done), cls, 0u);
    }
}

struct CountArgs {
    const uint8_t* qual; const uint8_t* seq; const uint8_t* const* refhot;
    const uint4* recs; const uint4* segs; const uint32_t* n_seg; uint32_t* seg_next;
    unsigned long long* tables; TableGeom geom;
    int lpr, rpw; uint32_t sh, lut_lo, lut_hi; uint8_t slot_q[4];
    uint32_t rec_bytes;             // ring slot of one pass' records: 32 * rpw
    int n_cls; uint32_t mm_cells;   // CTA-private mismatch tables in shared memory (0: straight to the global table)
};

__device__ __forceinline__ void cp_async16(uint32_t dst_shared, const void* src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_shared), "l"(src) : "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t a) { uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ void red_add_u64(unsigned long long* p, unsigned long long v) { asm volatile("red.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }

template <int S, bool INDEL>
__global__ void __launch_bounds__(CNT_WARPS * 32, CNT_MINB) bqsr_count_kernel(CountArgs A) {
    constexpr int NCH = INDEL ? 9 : 7;                                  // 16-byte chunks per lane and pass: QUAL 3, SEQ 2, REF 2 (+2)
    constexpr int STAGE_BYTES = NCH * 512;                              // per warp
    const uint32_t REC_BYTES = A.rec_bytes, WARP_BYTES = CNT_STAGES * STAGE_BYTES + CNT_RECRING * REC_BYTES;
    extern __shared__ __align__(16) unsigned char cnt_smem[];
    __shared__ uint32_t s_rt[33];
    if (threadIdx.x < 33) s_rt[threadIdx.x] = lanes::range_plane((int)threadIdx.x);
    __syncthreads();
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    // mismatches are sparse but pile up on few cells (a low QUAL value's 16 contexts take most of them): privatised per CTA in shared memory
    // -- global atomics on ~100 hot addresses serialise in L2 and were what bounded the first version of this kernel
    uint32_t* mm_cyc = reinterpret_cast<uint32_t*>(cnt_smem + (size_t)CNT_WARPS * WARP_BYTES);       // [n_cls][S][32 * lpr]
    const int Lpad = 32 * A.lpr;
    uint32_t* mm_ctx = mm_cyc + (size_t)A.n_cls * S * Lpad;                                          // [n_cls / 2][S][16]
    const bool mm_sh = A.mm_cells != 0;
    if (mm_sh) for (uint32_t i = threadIdx.x; i < A.mm_cells; i += blockDim.x) mm_cyc[i] = 0;
    __syncthreads();
    const uint32_t mm_cyc_s = (uint32_t)__cvta_generic_to_shared(mm_cyc), mm_ctx_s = (uint32_t)__cvta_generic_to_shared(mm_ctx);
    const uint32_t wbase = (uint32_t)__cvta_generic_to_shared(cnt_smem) + warp * WARP_BYTES;
    const uint32_t rbase = wbase + CNT_STAGES * STAGE_BYTES;
    const uint32_t rt_addr = (uint32_t)__cvta_generic_to_shared(s_rt);
    auto RT = [&](int n) -> uint32_t { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(rt_addr + 4u * (uint32_t)min(max(n, 0), 32))); return v; };
    const int lpr = A.lpr, rpw = A.rpw;
    const int r = (int)lane / lpr, c = (int)lane - r * lpr;
    const bool lane_used = r < rpw;
    const unsigned gmask = lane_used ? ((lpr == 32 ? 0xffffffffu : ((1u << lpr) - 1u)) << (r * lpr)) : (1u << lane);

    for (;;) {
        uint32_t seg = 0;
        if (lane == 0) seg = atomicAdd(A.seg_next, 1u);
        seg = __shfl_sync(FULL_MASK, seg, 0);
        if (seg >= __ldg(A.n_seg)) break;
        const uint4 sd = __ldg(A.segs + seg);
        const uint32_t rec_first = sd.x, n_rec = sd.y, cls = sd.z;
        const int n_pass = (int)((n_rec + (uint32_t)rpw - 1) / (uint32_t)rpw);
        const int cov = (int)(cls >> 1), sign = (cls & 1) ? -1 : 1;
        uint32_t pl[S][8], cx[8 * S];
#pragma unroll
        for (int s = 0; s < S; s++)
#pragma unroll
            for (int i = 0; i < 8; i++) pl[s][i] = 0;
#pragma unroll
        for (int i = 0; i < 8 * S; i++) cx[i] = 0;

        for (int it = 0; it < n_pass + CNT_D2; it++) {
            // ---- (a) records of pass `it` -> ring slot it % CNT_RECRING ----
            if (it < n_pass) {
                const uint32_t nrec_here = min((uint32_t)rpw, n_rec - (uint32_t)it * (uint32_t)rpw);
                for (uint32_t x = lane; x < 2 * nrec_here; x += 32) cp_async16(rbase + (uint32_t)(it % CNT_RECRING) * REC_BYTES + x * 16u, A.recs + 2 * ((uint64_t)rec_first + (uint64_t)it * rpw) + x);
            }
            // ---- (b) QUAL / SEQ / reference windows of pass it - CNT_D1 -> its stage ----
            const int pd = it - CNT_D1;
            if (pd >= 0 && pd < n_pass && lane_used && (uint32_t)(pd * rpw + r) < n_rec) {
                const uint32_t ra = rbase + (uint32_t)(pd % CNT_RECRING) * REC_BYTES + (uint32_t)r * 32u;
                const uint4 r0 = lds128(ra), r1 = lds128(ra + 16);
                const uint64_t wa = ((uint64_t)r0.y << 32) | r0.x, wb = ((uint64_t)r0.w << 32) | r0.z;
                const int Lk = (int)((wa >> 40) & 0x7ff); const bool rev = (wa >> 51) & 1;
                const int ow = rev ? Lk - 32 * c - 32 : 32 * c;
                const uint32_t st = wbase + (uint32_t)(pd % CNT_STAGES) * STAGE_BYTES + lane * 16u;
                const uint8_t* qp = A.qual + (int64_t)(wa & ((1ull << 40) - 1)) + ow;
                const uint8_t* q16 = reinterpret_cast<const uint8_t*>(reinterpret_cast<uintptr_t>(qp) & ~(uintptr_t)15);
                cp_async16(st, q16); cp_async16(st + 512, q16 + 16); cp_async16(st + 1024, q16 + 32);
                const int64_t ni = (int64_t)(wb & ((1ull << 41) - 1)) + ow;
                const uint8_t* sp = A.seq + (ni >> 1);
                const uint8_t* s16 = reinterpret_cast<const uint8_t*>(reinterpret_cast<uintptr_t>(sp) & ~(uintptr_t)15);
                cp_async16(st + 1536, s16); cp_async16(st + 2048, s16 + 16);
                const uint8_t* refp = A.refhot[(uint32_t)(wb >> 41)];
                const int64_t ri = (int64_t)(int32_t)r1.x - 1 + ow;
                const uint8_t* rp = refp + (ri >> 1);                                   // (arithmetic shift: ri may be slightly negative, the array is padded in front)
                const uint8_t* r16 = reinterpret_cast<const uint8_t*>(reinterpret_cast<uintptr_t>(rp) & ~(uintptr_t)15);
                cp_async16(st + 2560, r16); cp_async16(st + 3072, r16 + 16);
                if (INDEL) {
                    const int64_t ri2 = ri + ((int)(r1.w >> 19) - 4096);
                    const uint8_t* rp2 = refp + (ri2 >> 1);
                    const uint8_t* r216 = reinterpret_cast<const uint8_t*>(reinterpret_cast<uintptr_t>(rp2) & ~(uintptr_t)15);
                    cp_async16(st + 3584, r216); cp_async16(st + 4096, r216 + 16);
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group %0;" ::"n"(CNT_WN) : "memory");
            __syncwarp();
            // ---- (c) compute pass it - CNT_D2 ----
            const int pc = it - CNT_D2;
            if (pc >= 0) {
                const bool have = lane_used && (uint32_t)(pc * rpw + r) < n_rec;
                const uint32_t ra = rbase + (uint32_t)(pc % CNT_RECRING) * REC_BYTES + (uint32_t)r * 32u;
                uint4 r0 = make_uint4(0, 0, 0, 0), r1 = make_uint4(0, 0xffffu, 0xffffu, 0);
                if (have) { r0 = lds128(ra); r1 = lds128(ra + 16); }
                const uint64_t wa = ((uint64_t)r0.y << 32) | r0.x, wb = ((uint64_t)r0.w << 32) | r0.z;
                const int Lk = have ? (int)((wa >> 40) & 0x7ff) : 0; const bool rev = (wa >> 51) & 1;
                const int ow = rev ? Lk - 32 * c - 32 : 32 * c;
                const uint32_t st = wbase + (uint32_t)(pc % CNT_STAGES) * STAGE_BYTES + lane * 16u;
                lanes::LaneWin lw;
                {
                    const uint4 a0 = lds128(st), a1 = lds128(st + 512), a2 = lds128(st + 1024);
                    lw.QW[0] = a0.x; lw.QW[1] = a0.y; lw.QW[2] = a0.z; lw.QW[3] = a0.w; lw.QW[4] = a1.x; lw.QW[5] = a1.y; lw.QW[6] = a1.z; lw.QW[7] = a1.w;
                    lw.QW[8] = a2.x; lw.QW[9] = a2.y; lw.QW[10] = a2.z; lw.QW[11] = a2.w;
                    const uint4 s0 = lds128(st + 1536), s1 = lds128(st + 2048);
                    lw.SW[0] = s0.x; lw.SW[1] = s0.y; lw.SW[2] = s0.z; lw.SW[3] = s0.w; lw.SW[4] = s1.x; lw.SW[5] = s1.y; lw.SW[6] = s1.z; lw.SW[7] = s1.w;
                    const uint4 f0 = lds128(st + 2560), f1 = lds128(st + 3072);
                    lw.RW[0] = f0.x; lw.RW[1] = f0.y; lw.RW[2] = f0.z; lw.RW[3] = f0.w; lw.RW[4] = f1.x; lw.RW[5] = f1.y; lw.RW[6] = f1.z; lw.RW[7] = f1.w;
                    lw.kq = (uint32_t)((wa & ((1ull << 40) - 1)) + (uint64_t)(int64_t)ow + (uint64_t)reinterpret_cast<uintptr_t>(A.qual)) & 15u;
                    const int64_t ni = (int64_t)(wb & ((1ull << 41) - 1)) + ow;
                    lw.kb = (uint32_t)((uint64_t)(ni >> 1) + (uint64_t)reinterpret_cast<uintptr_t>(A.seq)) & 15u; lw.spar = (uint32_t)(ni & 1);
                    const uint8_t* refp = have ? A.refhot[(uint32_t)(wb >> 41)] : nullptr;
                    const int64_t ri = (int64_t)(int32_t)r1.x - 1 + ow;
                    lw.kr = (uint32_t)((uint64_t)(ri >> 1) + (uint64_t)reinterpret_cast<uintptr_t>(refp)) & 15u; lw.rpar = (uint32_t)(ri & 1);
                    lw.kr2 = 0; lw.rpar2 = 0;
                    if (INDEL) {
                        const uint4 b0 = lds128(st + 3584), b1 = lds128(st + 4096);
                        lw.RW2[0] = b0.x; lw.RW2[1] = b0.y; lw.RW2[2] = b0.z; lw.RW2[3] = b0.w; lw.RW2[4] = b1.x; lw.RW2[5] = b1.y; lw.RW2[6] = b1.z; lw.RW2[7] = b1.w;
                        const int64_t ri2 = ri + ((int)(r1.w >> 19) - 4096);
                        lw.kr2 = (uint32_t)((uint64_t)(ri2 >> 1) + (uint64_t)reinterpret_cast<uintptr_t>(refp)) & 15u; lw.rpar2 = (uint32_t)(ri2 & 1);
                    }
                }
                lanes::LaneRec lr;
                lr.Lk = Lk; lr.rev = rev; lr.skip0 = r1.y; lr.skip1 = r1.z; lr.bp = (int)(r1.w & 0x7ff); lr.insl = (int)((r1.w >> 11) & 0xff);
                lanes::LaneS1<S> s1;
                lanes::lane_stage1<S, INDEL>(lw, lr, c, A.sh, A.lut_lo, A.lut_hi, RT, s1);
                // low-quality tails: first / last base of the READ with QUAL > 2, over the lanes of the read
                // (a segmented REDUX compiles to a loop over the distinct member masks: one vote and two shuffles instead.  Kept-read coordinates
                // grow with the lane for forward reads and fall for reverse reads, so the extreme positions sit in the first / last lane that has any)
                const unsigned gb = __ballot_sync(FULL_MASK, s1.ll >= 0) & gmask;
                const int lo_lane = gb ? __ffs((int)gb) - 1 : (int)lane, hi_lane = gb ? 31 - __clz((int)gb) : (int)lane;
                const int lf_x = __shfl_sync(FULL_MASK, s1.lf, rev ? hi_lane : lo_lane), ll_x = __shfl_sync(FULL_MASK, s1.ll, rev ? lo_lane : hi_lane);
                const int leftPos = gb ? lf_x : 0x7fffffff, rightPos = gb ? ll_x : -1;
                lanes::LaneS2<S> s2;
                lanes::lane_stage2<S>(s1, lr, c, leftPos, rightPos, RT, s2);
                uint32_t ein = __shfl_up_sync(FULL_MASK, s2.epack, 1);     // the base before this lane's first one belongs to the lane below
                if (c == 0) ein = 0;
                lanes::LaneS3 s3;
                lanes::lane_stage3<S>(s2, ein, s3);
                const uint32_t pA = s2.pA, pC = s2.pC, pG = s2.pG, pT = s2.pT, qA = s3.qA, qC = s3.qC, qG = s3.qG, qT = s3.qT;
                const uint32_t counted = s2.counted, okc = s3.okc, Mm = s2.Mm;
                const int i0 = 32 * c;
                uint32_t Sp[S];
#pragma unroll
                for (int s = 0; s < S; s++) Sp[s] = s2.Sp[s];
                // ---- Cycles: vertical counters ----
#pragma unroll
                for (int s = 0; s < S; s++) {
                    uint32_t carry = Sp[s] & counted;
#pragma unroll
                    for (int i = 0; i < 8; i++) { const uint32_t t = pl[s][i] & carry; pl[s][i] ^= carry; carry = t; }
                }
                // ---- Contexts: key >> 4 = prev | cur << 2 (bqsr.go:64-76) ----
                {
                    const uint32_t cur[4] = {pA, pC, pG, pT}, prv[4] = {qA, qC, qG, qT};
                    uint32_t So[S];
#pragma unroll
                    for (int s = 0; s < S; s++) So[s] = Sp[s] & okc;
#pragma unroll
                    for (int cc = 0; cc < 4; cc++)
#pragma unroll
                        for (int pp = 0; pp < 4; pp++)
#pragma unroll
                            for (int s = 0; s < S; s++) {
                                const int cell = (pp + 4 * cc) * S + s;
                                cx[cell >> 1] += (uint32_t)__popc(prv[pp] & cur[cc] & So[s]) << (16 * (cell & 1));
                            }
                }
                // ---- mismatches (sparse): straight to the global table ----
                uint32_t mm = Mm & counted;
                if (mm) {
                    const uint32_t c0 = pC | pT, c1 = pG | pT, p0 = qC | qT, p1 = qG | qT;
#pragma unroll
                    for (int s = 0; s < S; s++) {
                        uint32_t ms = mm & Sp[s];
                        while (ms) {
                            const int b = __ffs((int)ms) - 1; ms &= ms - 1;
                            const int t = lanes::plane_base(b);
                            const int q = A.slot_q[s];
                            const int ctx = (int)(((p0 >> b) & 1u) | (((p1 >> b) & 1u) << 1) | (((c0 >> b) & 1u) << 2) | (((c1 >> b) & 1u) << 3));
                            if (mm_sh) {
                                asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(mm_cyc_s + 4u * (uint32_t)(((int)cls * S + s) * Lpad + i0 + t)) : "memory");
                                if ((okc >> b) & 1u) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(mm_ctx_s + 4u * (uint32_t)((cov * S + s) * 16 + ctx)) : "memory");
                            } else {
                                red_add_u64(A.tables + 2 * A.geom.idx(cov, q, A.geom.col_cycle(sign * (i0 + t + 1))) + 1, 1ull);
                                if ((okc >> b) & 1u) red_add_u64(A.tables + 2 * A.geom.idx(cov, q, A.geom.col_ctx(ctx)) + 1, 1ull);
                            }
                        }
                    }
                }
            }
            __syncwarp();
        }
        // ---- flush the segment ----
#pragma unroll
        for (int s = 0; s < S; s++) {
            const int q = A.slot_q[s];
#pragma unroll 1
            for (int b = 0; b < 32; b++) {
                uint32_t cnt = 0;
#pragma unroll
                for (int i = 0; i < 8; i++) cnt |= ((pl[s][i] >> b) & 1u) << i;
                if (cnt) red_add_u64(A.tables + 2 * A.geom.idx(cov, q, A.geom.col_cycle(sign * (32 * c + lanes::plane_base(b) + 1))), (unsigned long long)cnt);
            }
        }
#pragma unroll
        for (int cell = 0; cell < 16 * S; cell++) {
            uint32_t v = (cx[cell >> 1] >> (16 * (cell & 1))) & 0xffffu;
            v = __reduce_add_sync(FULL_MASK, v);
            if (lane == 0 && v) red_add_u64(A.tables + 2 * A.geom.idx(cov, A.slot_q[cell % S], A.geom.col_ctx(cell / S)), (unsigned long long)v);
        }
    }
    if (mm_sh) {
        __syncthreads();
        const uint32_t n_cyc = (uint32_t)(A.n_cls * S * Lpad);
        for (uint32_t i = threadIdx.x; i < A.mm_cells; i += blockDim.x) {
            const uint32_t v = mm_cyc[i];
            if (!v) continue;
            if (i < n_cyc) {
                const int t = (int)(i % (uint32_t)Lpad), s = (int)((i / (uint32_t)Lpad) % S), cl = (int)(i / ((uint32_t)Lpad * S));
                red_add_u64(A.tables + 2 * A.geom.idx(cl >> 1, A.slot_q[s], A.geom.col_cycle(((cl & 1) ? -1 : 1) * (t + 1))) + 1, (unsigned long long)v);
            } else {
                const uint32_t k = i - n_cyc; const int ctx = (int)(k & 15u), s = (int)((k >> 4) % S), cv = (int)((k >> 4) / S);
                red_add_u64(A.tables + 2 * A.geom.idx(cv, A.slot_q[s], A.geom.col_ctx(ctx)) + 1, (unsigned long long)v);
            }
        }
    }
}

// reference bases -> one-hot nibbles (A/a/* 1, C/c 2, G/g 4, T/t 8, everything else 0: never equal to a read base), low nibble first
__global__ void __launch_bounds__(256) ref_pack_hot_kernel(const uint8_t* __restrict__ ref, uint64_t n, uint8_t* __restrict__ out, uint64_t n_out) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_out) return;
    uint32_t v = 0;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const uint64_t j = 2 * t + h; uint32_t cd = 0;
        if (j < n) { const uint8_t b = ref[j]; if (b == 'A' || b == 'a' || b == '*') cd = 1; else if (b == 'C' || b == 'c') cd = 2; else if (b == 'G' || b == 'g') cd = 4; else if (b == 'T' || b == 't') cd = 8; }
        v |= cd << (4 * h);
    }
    out[t] = (uint8_t)v;
}
