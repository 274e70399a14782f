// Synthetic 150-bp paired-read workload generator -- bench/test input only.
// Host C++ (no CUDA): produces the columnar elp_batch payload, the reference genome (1 B/base, as
// fasta.MappedFasta.Seq returns it, fasta/fasta-files.go:355) and flattened known-sites intervals
// (intervals/intervals.go:103). Counter-based RNG: every value is a pure function of (seed, pair, stream).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

namespace {

inline uint64_t mix(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
struct Rng {  // splitmix64 stream keyed by (seed, id, stream)
    uint64_t s;
    Rng(uint64_t seed, uint64_t id, uint64_t stream) : s(mix(seed ^ mix(id * 0xD1342543DE82EF95ULL + stream))) {}
    inline uint64_t next() { s += 0x9E3779B97F4A7C15ULL; uint64_t z = s; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; return z ^ (z >> 31); }
    inline double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    inline uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }
};

struct Params {
    uint64_t seed; int64_t n_pairs; int32_t n_contigs; const int32_t* contig_len; int32_t L;
    double dup_frac, optical_frac, unmapped_frac, mate_unmapped_frac, secondary_frac, supplementary_frac, cross_contig_frac;
    int32_t n_rg; int32_t wide_quals; int32_t exome; int32_t first_refid;  // first_refid: refid offset (always 0 here)
    int32_t threads;
    const uint8_t* home;        // per contig: 1 = fragments start here (nullptr: everywhere); mates of cross-contig pairs may land on any contig
    uint64_t pair_id_base;      // added to the pair index in QNAMEs (several generators -> one name space)
    uint64_t genome_seed;       // seed of the reference genome (shared by the generators of one genome)
};

inline uint8_t genome_base(uint64_t seed, int32_t contig, int64_t pos0) {  // 0-based
    uint64_t h = mix(seed ^ 0xABCDEF12345ULL ^ ((uint64_t)contig << 40) ^ (uint64_t)(pos0 >> 4));
    return "ACGT"[(h >> ((pos0 & 15) * 2)) & 3];
}

struct Frag { int32_t contig; int32_t start; int32_t insert; bool first_fwd; };
struct PairClass { bool dup, optical, unmapped, mate_unmapped, secondary, supplementary, cross; int64_t src; };

struct Gen {
    Params P; std::vector<double> cum; double total_len;
    explicit Gen(const Params& p) : P(p) {
        cum.resize(p.n_contigs + 1); cum[0] = 0;
        for (int i = 0; i < p.n_contigs; i++) cum[i + 1] = cum[i] + ((!p.home || p.home[i]) ? (double)p.contig_len[i] : 0.0);
        total_len = cum[p.n_contigs];
    }
    PairClass classify(int64_t p) const {
        Rng r(P.seed, (uint64_t)p, 1);
        PairClass c{};
        double u = r.uni();
        c.dup = p > 0 && u < P.dup_frac;
        c.src = c.dup ? (int64_t)(r.uni() * (double)p) : -1;
        c.optical = c.dup && r.uni() < P.optical_frac;
        double v = r.uni();
        double a = P.unmapped_frac, b = a + P.mate_unmapped_frac, d = b + P.cross_contig_frac;
        if (!c.dup) { c.unmapped = v < a; c.mate_unmapped = v >= a && v < b; c.cross = v >= b && v < d && P.n_contigs > 1; }
        double w = r.uni();
        c.secondary = !c.unmapped && w < P.secondary_frac;
        c.supplementary = !c.unmapped && w >= P.secondary_frac && w < P.secondary_frac + P.supplementary_frac;
        return c;
    }
    int64_t root(int64_t p) const { for (;;) { PairClass c = classify(p); if (!c.dup) return p; p = c.src; } }
    Frag fragment(int64_t p) const {  // p must be a root (non-duplicate) pair
        Rng r(P.seed, (uint64_t)p, 2);
        Frag f;
        double g = std::sqrt(-2.0 * std::log(1.0 - r.uni())) * std::cos(6.283185307179586 * r.uni());
        int ins = (int)std::lround(400.0 + 60.0 * g);
        ins = std::max(60, std::min(1000, ins));
        if (r.uni() < 0.02) ins = 60 + (int)r.below(90);   // short inserts: reads run into the adaptor (exercises hardClipAdaptorSequence)
        double u = r.uni() * total_len;
        int c = (int)(std::upper_bound(cum.begin(), cum.end(), u) - cum.begin()) - 1;
        c = std::max(0, std::min(P.n_contigs - 1, c));
        int32_t ln = P.contig_len[c];
        int32_t span = std::max(ins, P.L) + P.L + 16;  // room for adaptor read-through and indels
        int32_t maxstart = ln - span - 200;
        if (maxstart < 200) maxstart = 200;
        int32_t start;
        if (P.exome) {  // reads pile up on ~2 % of the genome: exons of 300 bp every 15 kbp
            int32_t n_ex = std::max(1, ln / 15000);
            int32_t ex = (int32_t)r.below((uint32_t)n_ex);
            int32_t ex_start = 1000 + ex * 15000 + (int32_t)(mix(P.seed ^ (uint64_t)c << 32 ^ (uint64_t)ex) % 5000);
            start = ex_start - 250 + (int32_t)r.below(550);
        } else {
            start = 200 + (int32_t)(r.uni() * (double)(maxstart - 200));
        }
        start = std::max(200, std::min(maxstart, start));
        f.contig = c; f.start = start; f.insert = ins; f.first_fwd = (r.next() & 1) != 0;
        return f;
    }
};

// one read's alignment shape, in reference orientation
struct ReadShape { int32_t pos; uint32_t cig[4]; int ncig; };

// kind: 0 = all M, 1 = soft clip (left or right), 2 = insertion, 3 = deletion, 4 = soft clip + indel
void make_shape(Rng& r, int L, int32_t anchor, bool anchor_is_end, ReadShape& s) {
    double u = r.uni();
    // clip / indel geometry; the L >= 100 formulas are the original ones (same random stream), shorter reads scale them down
    const int kmax = L >= 100 ? 40 : std::max(1, L / 4);
    int k = 1 + (int)r.below((uint32_t)kmax), il = 1 + (int)r.below(5), ip;
    if (L - 80 > 0 && L >= 100) ip = 20 + (int)r.below((uint32_t)(L - 80));
    else { const int lo = std::max(2, L / 5); ip = lo + (int)r.below((uint32_t)std::max(1, L - 2 * lo - 8 - kmax)); }
    bool left = (r.next() & 1) != 0;
    auto M = [](int n) { return ((uint32_t)n << 4) | 0u; };
    auto I = [](int n) { return ((uint32_t)n << 4) | 1u; };
    auto D = [](int n) { return ((uint32_t)n << 4) | 2u; };
    auto S = [](int n) { return ((uint32_t)n << 4) | 4u; };
    int reflen = L; s.ncig = 0;
    if (u < 0.84) { s.cig[s.ncig++] = M(L); }
    else if (u < 0.92) { if (left) { s.cig[s.ncig++] = S(k); s.cig[s.ncig++] = M(L - k); } else { s.cig[s.ncig++] = M(L - k); s.cig[s.ncig++] = S(k); } reflen = L - k; }
    else if (u < 0.96) { s.cig[s.ncig++] = M(ip); s.cig[s.ncig++] = I(il); s.cig[s.ncig++] = M(L - ip - il); reflen = L - il; }
    else if (u < 0.99) { s.cig[s.ncig++] = M(ip); s.cig[s.ncig++] = D(il); s.cig[s.ncig++] = M(L - ip); reflen = L + il; }
    else {
        if (left) { s.cig[s.ncig++] = S(k); s.cig[s.ncig++] = M(ip); s.cig[s.ncig++] = D(il); s.cig[s.ncig++] = M(L - k - ip); }
        else { s.cig[s.ncig++] = M(ip); s.cig[s.ncig++] = I(il); s.cig[s.ncig++] = M(L - ip - il - k); s.cig[s.ncig++] = S(k); }
        reflen = left ? (L - k + il) : (L - il - k);
    }
    // anchor: forward read keeps its first aligned base at the fragment start; reverse read keeps its last aligned base at the fragment end
    s.pos = anchor_is_end ? (anchor - reflen + 1) : anchor;
}

static const uint8_t Q4[4] = {2, 12, 23, 37};

struct Out {
    int32_t *refid, *pos; uint16_t* flag; uint8_t* mapq; int32_t *nref, *pnext, *tlen, *rg;
    uint64_t* qname_off; uint8_t* qname; uint64_t* cigar_off; uint32_t* cigar; int32_t* lseq; uint8_t *seq, *qual;
};

// per pair: number of records, cigar ops, qname bytes
struct PairSize { int32_t recs, cig, qn; };

struct PairGen {
    const Gen& G;
    explicit PairGen(const Gen& g) : G(g) {}

    int qname(int64_t p, const PairClass& c, int64_t root, char* buf) const {
        // SYN:<run>:FC1:<lane>:<tile>:<x>:<y>  (7 fields: tile,x,y = fields 4,5,6; mark-optical-duplicates.go:50-71)
        uint64_t key = (uint64_t)(c.optical ? root : p) + G.P.pair_id_base;
        uint64_t v = (key * 0x9E3779B97F4A7C15ULL) & ((1ULL << 40) - 1);
        int lane = 1 + (int)(v & 3), tile = 1101 + (int)((v >> 2) & 1023), x = 1 + (int)((v >> 12) & 0x3fff), y = 1 + (int)((v >> 26) & 0x3fff);
        if (c.optical) {  // same tile, within 50 px of the origin; the run field keeps the name unique
            Rng r(G.P.seed, (uint64_t)p, 7);
            x += (int)r.below(50); y += (int)r.below(50);
            return std::snprintf(buf, 64, "SYN:%lld:FC1:%d:%d:%d:%d", (long long)(p + 2 + (int64_t)G.P.pair_id_base), lane, tile, x, y);
        }
        return std::snprintf(buf, 64, "SYN:1:FC1:%d:%d:%d:%d", lane, tile, x, y);
    }

    // emit one read's bases+quals in reference orientation
    void fill_read(Rng& r, int32_t contig, const ReadShape& s, int L, int32_t frag_lo, int32_t frag_hi, bool second, uint8_t* seqdst, uint8_t* qualdst) const {
        const Params& P = G.P;
        uint8_t bases[8192];
        int ri = 0; int64_t ref0 = (int64_t)s.pos - 1;
        for (int k = 0; k < s.ncig; k++) {
            int ln = (int)(s.cig[k] >> 4); int op = (int)(s.cig[k] & 15);
            if (op == 0) { for (int q = 0; q < ln; q++, ri++, ref0++) { bool in_frag = (ref0 + 1) >= frag_lo && (ref0 + 1) <= frag_hi; bases[ri] = in_frag ? genome_base(P.genome_seed, contig, ref0) : (uint8_t)"ACGT"[r.next() & 3]; } }
            else if (op == 2) ref0 += ln;
            else { for (int q = 0; q < ln; q++, ri++) bases[ri] = (uint8_t)"ACGT"[r.next() & 3]; }
        }
        for (int i = 0; i < L; i++) {
            // position-dependent quality: tails decay
            double edge = (double)std::min(i, L - 1 - i) / (double)L;
            double u = r.uni();
            uint8_t q;
            if (P.wide_quals) { int base = (int)(8 + 32 * std::min(1.0, edge * 6 + 0.35)); int qq = base - (int)r.below(12) + (int)r.below(6); q = (uint8_t)std::max(2, std::min(42, qq)); if (u < 0.02) q = 2; }
            else { double p37 = 0.55 + 0.35 * std::min(1.0, edge * 8), p23 = p37 + 0.6 * (1 - p37), p12 = p23 + 0.7 * (1 - p23); q = u < p37 ? Q4[3] : (u < p23 ? Q4[2] : (u < p12 ? Q4[1] : Q4[0])); }
            double qtrue = (double)q - (second ? 3.0 : 0.0);
            if (r.uni() < std::pow(10.0, -qtrue / 10.0)) { uint8_t b = bases[i]; uint8_t nb; do { nb = (uint8_t)"ACGT"[r.next() & 3]; } while (nb == b); bases[i] = nb; }
            if (r.uni() < 0.001) { bases[i] = 'N'; q = 2; }
            qualdst[i] = q;
        }
        for (int i = 0; i < L; i += 2) {
            auto nib = [](uint8_t b) -> uint8_t { return b == 'A' ? 1 : b == 'C' ? 2 : b == 'G' ? 4 : b == 'T' ? 8 : 15; };
            uint8_t hi = nib(bases[i]), lo = (i + 1 < L) ? nib(bases[i + 1]) : 0;
            seqdst[i >> 1] = (uint8_t)((hi << 4) | lo);
        }
    }

    // Generates pair p. If out == nullptr only sizes are computed.
    PairSize gen(int64_t p, const Out* out, int64_t rec0, uint64_t cig0, uint64_t qn0) const {
        const Params& P = G.P; const int L = P.L;
        PairClass c = G.classify(p);
        int64_t root = c.dup ? G.root(p) : p;
        Frag f = G.fragment(root);
        PairClass rc = c.dup ? G.classify(root) : c;
        bool unmapped = rc.unmapped, mate_unmapped = rc.mate_unmapped, cross = rc.cross;
        char qn[64]; int qnl = qname(p, c, root, qn);
        // alignment shapes come from the ROOT pair's stream (duplicates share coordinates, CIGAR and strand)
        Rng rs(P.seed, (uint64_t)root, 3);
        ReadShape sf, sr;  // forward (leftmost) and reverse (rightmost) reads
        make_shape(rs, L, f.start, false, sf);
        make_shape(rs, L, f.start + f.insert - 1, true, sr);
        int32_t ccontig = f.contig, cpos = 0;
        if (cross) { Rng rx(P.seed, (uint64_t)root, 4); ccontig = (f.contig + 1 + (int)rx.below((uint32_t)(P.n_contigs - 1))) % P.n_contigs; int32_t ln = P.contig_len[ccontig]; cpos = 200 + (int32_t)rx.below((uint32_t)std::max(1, ln - 800)); sr.pos = cpos; sr.ncig = 1; sr.cig[0] = ((uint32_t)L << 4); }
        Rng rm(P.seed, (uint64_t)root, 5);
        double um = rm.uni();
        uint8_t mapq = um < 0.88 ? 60 : (um < 0.92 ? 0 : (uint8_t)(1 + rm.below(59)));
        int rg = P.n_rg > 0 ? (int)(mix(P.seed ^ (uint64_t)root * 77) % (uint64_t)P.n_rg) : -1;
        int nrec = 2 + (c.secondary ? 1 : 0) + (c.supplementary ? 1 : 0);
        PairSize sz{nrec, 0, nrec * qnl};
        // records: [first-in-pair, second-in-pair, (secondary|supplementary copy of the first)]
        bool first_is_fwd = f.first_fwd;
        for (int k = 0; k < nrec; k++) {
            bool is_extra = k >= 2;
            bool this_first = is_extra ? true : (k == 0);
            bool this_fwd = this_first ? first_is_fwd : !first_is_fwd;
            const ReadShape& me = this_fwd ? sf : sr; const ReadShape& mate = this_fwd ? sr : sf;
            int32_t my_contig = (cross && !this_fwd) ? ccontig : f.contig, mate_contig = (cross && this_fwd) ? ccontig : f.contig;
            uint16_t flag = 0x1; int32_t refid, pos, nref, pnext, tlen; int ncig; const uint32_t* cg; uint8_t mq = mapq;
            bool me_unmapped = unmapped || (mate_unmapped && !this_first);
            bool mt_unmapped = unmapped || (mate_unmapped && this_first);
            flag |= this_first ? 0x40 : 0x80;
            if (me_unmapped) flag |= 0x4;
            if (mt_unmapped) flag |= 0x8;
            if (!me_unmapped && !this_fwd) flag |= 0x10;
            if (!mt_unmapped && this_fwd) flag |= 0x20;
            if (!me_unmapped && !mt_unmapped && !cross) flag |= 0x2;
            if (unmapped) { refid = -1; pos = 0; nref = -1; pnext = 0; tlen = 0; ncig = 0; cg = nullptr; mq = 0; }
            else if (mate_unmapped) {  // SAM convention: the unmapped mate sits at the mapped mate's coordinates
                const ReadShape& mapped = first_is_fwd ? sf : sr;
                refid = f.contig; pos = mapped.pos; nref = f.contig; pnext = mapped.pos; tlen = 0;
                if (me_unmapped) { ncig = 0; cg = nullptr; mq = 0; } else { ncig = mapped.ncig; cg = mapped.cig; }
            } else {
                refid = my_contig; pos = me.pos; nref = mate_contig; pnext = mate.pos; ncig = me.ncig; cg = me.cig;
                tlen = cross ? 0 : (this_fwd ? f.insert : -f.insert);
            }
            if (is_extra) { flag |= (c.secondary ? 0x100 : 0x800); if (c.supplementary) { Rng rx(P.seed, (uint64_t)p, 6); pos = std::max(1, pos + 300 + (int32_t)rx.below(2000)); pos = std::min(pos, P.contig_len[refid < 0 ? 0 : refid] - L - 10); } }
            sz.cig += ncig;
            if (out) {
                int64_t i = rec0 + k;
                out->refid[i] = refid; out->pos[i] = pos; out->flag[i] = flag; out->mapq[i] = mq; out->nref[i] = nref; out->pnext[i] = pnext; out->tlen[i] = tlen; out->rg[i] = rg;
                out->qname_off[i] = qn0; std::memcpy(out->qname + qn0, qn, (size_t)qnl); qn0 += (uint64_t)qnl;
                out->cigar_off[i] = cig0; for (int q = 0; q < ncig; q++) out->cigar[cig0 + (uint64_t)q] = cg[q]; cig0 += (uint64_t)ncig;
                out->lseq[i] = L;
                // fresh bases/quals per record (duplicates differ in sequencing errors and quality)
                Rng rb(P.seed, (uint64_t)p * 4 + (uint64_t)k, 8);
                ReadShape shp; shp.pos = pos; shp.ncig = ncig; for (int q = 0; q < ncig; q++) shp.cig[q] = cg[q];
                if (ncig == 0) { shp.ncig = 1; shp.cig[0] = ((uint32_t)L << 4) | 4u; }  // unmapped: random bases
                int32_t frag_lo = (cross || is_extra) ? -1 : f.start, frag_hi = (cross || is_extra) ? 0x7fffffff : f.start + f.insert - 1;
                if (cross || is_extra) { frag_lo = 1; }
                fill_read(rb, refid < 0 ? 0 : refid, shp, L, frag_lo, frag_hi, !this_first, out->seq + (uint64_t)i * (uint64_t)((L + 1) / 2), out->qual + (uint64_t)i * (uint64_t)L);
            }
        }
        return sz;
    }
};

template <class F> void par_for(int64_t n, int threads, F f) {
    threads = std::max(1, threads);
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++) th.emplace_back([=]() { int64_t lo = n * t / threads, hi = n * (t + 1) / threads; f(lo, hi, t); });
    for (auto& x : th) x.join();
}

}  // namespace

extern "C" {

struct synth_params {
    uint64_t seed; int64_t n_pairs; int32_t n_contigs; const int32_t* contig_len; int32_t L;
    double dup_frac, optical_frac, unmapped_frac, mate_unmapped_frac, secondary_frac, supplementary_frac, cross_contig_frac;
    int32_t n_rg, wide_quals, exome, threads;
    const uint8_t* home; uint64_t pair_id_base, genome_seed;
};

static Params to_params(const synth_params* sp) {
    Params P{}; P.seed = sp->seed; P.n_pairs = sp->n_pairs; P.n_contigs = sp->n_contigs; P.contig_len = sp->contig_len; P.L = sp->L;
    P.dup_frac = sp->dup_frac; P.optical_frac = sp->optical_frac; P.unmapped_frac = sp->unmapped_frac; P.mate_unmapped_frac = sp->mate_unmapped_frac;
    P.secondary_frac = sp->secondary_frac; P.supplementary_frac = sp->supplementary_frac; P.cross_contig_frac = sp->cross_contig_frac;
    P.n_rg = sp->n_rg; P.wide_quals = sp->wide_quals; P.exome = sp->exome; P.threads = sp->threads; P.home = sp->home; P.pair_id_base = sp->pair_id_base; P.genome_seed = sp->genome_seed ? sp->genome_seed : sp->seed; return P;
}

// pass 1: per-pair sizes -> totals; pair_rec0/pair_cig0/pair_qn0 are exclusive prefix sums (length n_pairs+1)
int synth_sizes(const synth_params* sp, int64_t* pair_rec0, uint64_t* pair_cig0, uint64_t* pair_qn0) {
    Params P = to_params(sp); Gen G(P); PairGen PG(G);
    int64_t n = P.n_pairs;
    par_for(n, P.threads, [&](int64_t lo, int64_t hi, int) { for (int64_t p = lo; p < hi; p++) { PairSize s = PG.gen(p, nullptr, 0, 0, 0); pair_rec0[p + 1] = s.recs; pair_cig0[p + 1] = (uint64_t)s.cig; pair_qn0[p + 1] = (uint64_t)s.qn; } });
    pair_rec0[0] = 0; pair_cig0[0] = 0; pair_qn0[0] = 0;
    for (int64_t p = 0; p < n; p++) { pair_rec0[p + 1] += pair_rec0[p]; pair_cig0[p + 1] += pair_cig0[p]; pair_qn0[p + 1] += pair_qn0[p]; }
    return 0;
}

int synth_fill(const synth_params* sp, const int64_t* pair_rec0, const uint64_t* pair_cig0, const uint64_t* pair_qn0,
               int32_t* refid, int32_t* pos, uint16_t* flag, uint8_t* mapq, int32_t* nref, int32_t* pnext, int32_t* tlen, int32_t* rg,
               uint64_t* qname_off, uint8_t* qname, uint64_t* cigar_off, uint32_t* cigar, int32_t* lseq, uint8_t* seq, uint8_t* qual) {
    Params P = to_params(sp); Gen G(P); PairGen PG(G);
    Out o{refid, pos, flag, mapq, nref, pnext, tlen, rg, qname_off, qname, cigar_off, cigar, lseq, seq, qual};
    int64_t n = P.n_pairs;
    par_for(n, P.threads, [&](int64_t lo, int64_t hi, int) { for (int64_t p = lo; p < hi; p++) PG.gen(p, &o, pair_rec0[p], pair_cig0[p], pair_qn0[p]); });
    qname_off[pair_rec0[n]] = pair_qn0[n]; cigar_off[pair_rec0[n]] = pair_cig0[n];
    return 0;
}

// reference genome bases for one contig (1 B/base), 0-based [lo,hi)
int synth_genome(uint64_t seed, int32_t contig, int64_t lo, int64_t hi, uint8_t* dst, int32_t threads) {
    par_for(hi - lo, threads, [&](int64_t a, int64_t b, int) { for (int64_t i = a; i < b; i++) dst[i] = genome_base(seed, contig, lo + i); });
    return 0;
}

// known sites for one contig: ~1 per 1000 bp (1-bp) plus 0.1 % 10-bp intervals; sorted, non-overlapping. returns count (pairs start,end 1-based inclusive)
int64_t synth_known_sites(uint64_t seed, int32_t contig, int32_t contig_len, int32_t* se, int64_t cap) {
    int64_t n = 0;
    for (int32_t blk = 0; (int64_t)blk * 1000 + 1000 < contig_len; blk++) {
        uint64_t h = mix(seed ^ 0x51735ULL ^ ((uint64_t)contig << 36) ^ (uint64_t)blk);
        int32_t p = blk * 1000 + 1 + (int32_t)(h % 980);
        int32_t len = ((h >> 40) % 1000 == 0) ? 10 : 1;
        if (n < cap) { se[2 * n] = p; se[2 * n + 1] = p + len - 1; }
        n++;
    }
    return n;
}

// columns -> BAM alignment records (sam/bam-files.go:300-400 layout; each with its block_size, RG:Z and an NM:C tag).
// pass 1 (out == nullptr): fills rec_off[n+1]; pass 2: writes the bytes.  rg_ids: concatenated @RG ID strings, rg_id_off[n_rg+1].
int synth_encode_bam(int64_t n, const int32_t* refid, const int32_t* pos, const uint16_t* flag, const uint8_t* mapq, const int32_t* nref, const int32_t* pnext,
                     const int32_t* tlen, const int32_t* rg, const uint64_t* qname_off, const uint8_t* qname, const uint64_t* cigar_off, const uint32_t* cigar,
                     const int32_t* lseq, const uint64_t* seq_off, const uint8_t* seq, const uint64_t* qual_off, const uint8_t* qual,
                     const uint8_t* rg_ids, const uint32_t* rg_id_off, uint64_t* rec_off, uint8_t* out, int32_t threads) {
    auto rec_len = [&](int64_t i) -> uint64_t {
        const uint64_t qn = qname_off[i + 1] - qname_off[i], nc = cigar_off[i + 1] - cigar_off[i], L = (uint64_t)lseq[i];
        uint64_t aux = 4;   // NM:C
        if (rg[i] >= 0) aux += 3 + (rg_id_off[rg[i] + 1] - rg_id_off[rg[i]]) + 1;
        return 36 + qn + 1 + 4 * nc + (L + 1) / 2 + L + aux;
    };
    if (!out) { rec_off[0] = 0; for (int64_t i = 0; i < n; i++) rec_off[i + 1] = rec_off[i] + rec_len(i); return 0; }
    par_for(n, threads, [&](int64_t lo, int64_t hi, int) {
        for (int64_t i = lo; i < hi; i++) {
            uint8_t* r = out + rec_off[i];
            const uint64_t len = rec_off[i + 1] - rec_off[i];
            const uint32_t qn = (uint32_t)(qname_off[i + 1] - qname_off[i]), nc = (uint32_t)(cigar_off[i + 1] - cigar_off[i]); const int32_t L = lseq[i];
            auto w32 = [&](int o, uint32_t v) { r[o] = v & 255; r[o + 1] = (v >> 8) & 255; r[o + 2] = (v >> 16) & 255; r[o + 3] = v >> 24; };
            w32(0, (uint32_t)(len - 4)); w32(4, (uint32_t)refid[i]); w32(8, (uint32_t)(pos[i] - 1));
            r[12] = (uint8_t)(qn + 1); r[13] = mapq[i]; r[14] = 0x48; r[15] = 0x12; r[16] = nc & 255; r[17] = (nc >> 8) & 255; r[18] = flag[i] & 255; r[19] = flag[i] >> 8;
            w32(20, (uint32_t)L); w32(24, (uint32_t)nref[i]); w32(28, (uint32_t)(pnext[i] - 1)); w32(32, (uint32_t)tlen[i]);
            uint8_t* p = r + 36;
            std::memcpy(p, qname + qname_off[i], qn); p += qn; *p++ = 0;
            std::memcpy(p, cigar + cigar_off[i], 4ull * nc); p += 4ull * nc;
            std::memcpy(p, seq + seq_off[i], (size_t)(L + 1) / 2); p += (L + 1) / 2;
            std::memcpy(p, qual + qual_off[i], (size_t)L); p += L;
            *p++ = 'N'; *p++ = 'M'; *p++ = 'C'; *p++ = (uint8_t)(i & 7);
            if (rg[i] >= 0) { *p++ = 'R'; *p++ = 'G'; *p++ = 'Z'; const uint32_t a = rg_id_off[rg[i]], b = rg_id_off[rg[i] + 1]; std::memcpy(p, rg_ids + a, b - a); p += b - a; *p++ = 0; }
        }
    });
    return 0;
}
// out[out_off[k] .. out_off[k+1]) = data[off[idx[k]] .. off[idx[k]+1])  (elem bytes per element): the ragged gather behind
// AlignmentBatch.take at bench scale (30 M reads), where numpy index arrays would not fit in host memory
int synth_take_ragged(const int64_t* idx, int64_t n, const uint64_t* off, const uint8_t* data, int32_t elem, const uint64_t* out_off, uint8_t* out, int32_t threads) {
    par_for(n, threads, [&](int64_t lo, int64_t hi, int) {
        for (int64_t k = lo; k < hi; k++) {
            const uint64_t a = off[idx[k]], b = off[idx[k] + 1];
            std::memcpy(out + out_off[k] * (uint64_t)elem, data + a * (uint64_t)elem, (size_t)((b - a) * (uint64_t)elem));
        }
    });
    return 0;
}
}
