// A C++ client of include/elprep_b200.hpp (the host layer above the C ABI, reference names).  Compiled by tests/test_zz_cpp_host.py.
//   cpp_host_client marshal   -> prints the columns DeviceSam::marshal builds for a fixed set of alignments (no device needed)
//   cpp_host_client           -> the four-read scenario of tests/c/cabi_client.c through DeviceSam; without a GPU the constructor must
//                                throw elprep::Panic with ELP_ENODEVICE (prints "nodevice: ...")
#include <cstdio>
#include <string>
#include "elprep_b200.hpp"

using namespace elprep;

template <class V> static void line(const char* name, const V& v) { std::printf("%s", name); for (auto x : v) std::printf(" %llu", (unsigned long long)(uint64_t)(int64_t)x); std::printf("\n"); }

static Alignment aln(const char* qn, uint16_t flag, const char* rname, int32_t pos, uint8_t mapq, std::vector<CigarOperation> cig, const char* rnext, int32_t pnext, int32_t tlen,
                     const char* seq, std::vector<uint8_t> qual, const char* rg, bool sr = false) {
    Alignment a; a.QNAME = qn; a.FLAG = flag; a.RNAME = rname; a.POS = pos; a.MAPQ = mapq; a.CIGAR = std::move(cig); a.RNEXT = rnext; a.PNEXT = pnext; a.TLEN = tlen;
    a.SEQ = seq; a.QUAL = std::move(qual); if (rg) { a.RG = rg; a.hasRG = true; } a.sr = sr; return a;
}

int main(int argc, char** argv) {
    Header h; h.sq = {{"chr1", 100000}, {"chr2", 50000}};
    Header::ReadGroup g1; g1.ID = "rg1"; g1.LB = "libA"; g1.hasLB = true; Header::ReadGroup g2; g2.ID = "rg2"; h.rg = {g1, g2};
    if (argc > 1 && std::string(argv[1]) == "marshal") {
        std::vector<Alignment> b = {
            aln("readA", 99, "chr1", 100, 60, {{3, 'S'}, {2, 'M'}, {3, 'M'}, {1, 'I'}, {4, 'M'}}, "=", 250, 160, "ACGTNacgtRYKM", {30, 31, 32, 33, 2, 2, 20, 21, 22, 23, 24, 25, 26}, "rg2"),
            aln("b", 147, "chr2", 7, 0, {{5, 'M'}}, "chr1", 9, -3, "TTTTT", {40, 40, 40, 40, 40}, "rg1", true),
            aln("", 4, "*", 0, 0, {}, "*", 0, 0, "", {}, nullptr),
            aln("weird", 0, "chrUn", 5, 3, {{2, 'H'}, {1, '='}, {1, 'X'}, {2, 'D'}, {1, 'N'}, {1, 'P'}}, "chrUn", 1, 0, "G*", {1, 93}, "rg1")};
        Columns c = DeviceSam::marshal(h, b);
        line("refid", c.refid); line("nref", c.nref); line("pos", c.pos); line("pnext", c.pnext); line("tlen", c.tlen); line("rg", c.rg); line("lseq", c.lseq);
        line("flag", c.flag); line("mapq", c.mapq); line("opt", c.opt); line("qname", c.qname); line("qname_off", c.qname_off); line("cigar", c.cigar); line("cigar_off", c.cigar_off);
        line("seq", c.seq); line("qual", c.qual);
        return 0;
    }
    try {
        Options o; o.MarkDuplicates = true; o.AlsoOpticals = true;
        Header h1; h1.sq = {{"chr1", 100000}}; h1.rg = {g1};
        DeviceSam d(h1, o);
        std::vector<Alignment> b = {aln("r1", 0, "chr1", 50, 60, {{4, 'M'}}, "*", 0, 0, "ACGT", {30, 30, 30, 30}, "rg1"), aln("r2", 0, "chr1", 50, 60, {{4, 'M'}}, "*", 0, 0, "ACGT", {20, 20, 20, 20}, "rg1"),
                                    aln("r3", 16, "chr1", 20, 60, {{4, 'M'}}, "*", 0, 0, "ACGT", {25, 25, 25, 25}, "rg1"), aln("r4", 4, "*", 0, 0, {}, "*", 0, 0, "ACGT", {2, 2, 2, 2}, "rg1")};
        d.AddNodes(b);
        d.Finalize(SortingOrder::Coordinate);
        std::printf("order");
        std::vector<uint16_t> flags;
        d.RunPipeline([&](const Alignment& a) { std::printf(" %s", a.QNAME.c_str()); flags.push_back(a.FLAG); });
        std::printf(" flags"); for (auto f : flags) std::printf(" %u", f); std::printf("\n");
        auto m = d.MarkOpticalDuplicates();
        std::printf("libA unpaired %lld dups %lld unmapped %lld\n", (long long)m["libA"].UnpairedReadsExamined, (long long)m["libA"].UnpairedReadDuplicates, (long long)m["libA"].UnmappedReads);
        try { d.ApplyBQSR(); std::printf("apply-before-finalize did not throw\n"); } catch (const Panic& p) { std::printf("apply-before-finalize rc %d\n", p.code); }
    } catch (const Panic& p) {
        if (p.code == ELP_ENODEVICE) { std::printf("nodevice: %s\n", p.what()); return 0; }
        std::printf("panic %d: %s\n", p.code, p.what()); return 1;
    }
    return 0;
}
