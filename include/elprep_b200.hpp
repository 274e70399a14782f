// elprep_b200.hpp -- the host side above the C ABI in C++ (header only, C++17): elPrep's own types and operator names for the hot
// path, for callers that are compiled code like the reference (which is Go; this image has no Go toolchain, go/gpu/device_sam.go is
// the same layer as uncompiled cgo source).  Nothing here computes: it marshals sam.Alignment-shaped records into the columnar
// elp_batch, calls the library, and turns a failing return code into an exception carrying the text the reference panics with.
//
//   reference                                                       here
//   sam.Alignment (sam/sam-types.go:289-331)                        elprep::Alignment
//   sam.Header SQ / RG tables (sam/sam-types.go:60-118)             elprep::Header
//   (*sam.Sam).AddNodes receive + Finalize (filter-pipeline.go:108-128)   DeviceSam::AddNodes(batch) ... DeviceSam::Finalize(order)
//   filters.MarkDuplicates / MarkOpticalDuplicates (mark-duplicates.go:406, mark-optical-duplicates.go:468)   Options{MarkDuplicates, AlsoOpticals}, DeviceSam::MarkOpticalDuplicates()
//   (*BaseRecalibrator).Recalibrate (bqsr.go:467)                   DeviceSam::Recalibrate()
//   FinalizeBQSRTables + PrintBQSRTables (bqsr.go:677, print-bqsr.go:269)   DeviceSam::FinalizeBQSRTables(recalFile)
//   ApplyBQSR (bqsr.go:936)                                         DeviceSam::ApplyBQSR()
//   (*sam.Sam).RunPipeline as PipelineInput (filter-pipeline.go:242-279)   DeviceSam::RunPipeline(sink)
//   log.Panic(msg)                                                  throw elprep::Panic(msg)
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "elprep_b200.h"

namespace elprep {

struct Panic : std::runtime_error { int code; Panic(int rc, const std::string& msg) : std::runtime_error(msg), code(rc) {} };

struct CigarOperation { int32_t Length; char Operation; };           // sam/sam-types.go:675-678

// sam.Alignment with the optional fields the path reads (RG:Z, the `sr` tag of `elprep split`)
struct Alignment {
    std::string QNAME; uint16_t FLAG = 0; std::string RNAME = "*"; int32_t POS = 0; uint8_t MAPQ = 0; std::vector<CigarOperation> CIGAR;
    std::string RNEXT = "*"; int32_t PNEXT = 0, TLEN = 0;
    std::string SEQ;                   // bases as text; packed to BAM nibbles on the way in (sam-types.go:270-277: unknown characters become N)
    std::vector<uint8_t> QUAL;         // phred values without +33, one per base
    std::string RG; bool hasRG = false; bool sr = false;
};

struct Header {
    struct SQ { std::string SN; int32_t LN; };
    struct ReadGroup { std::string ID, LB, PU; bool hasLB = false, hasPU = false; };
    std::vector<SQ> sq; std::vector<ReadGroup> rg;
};

enum class SortingOrder { Keep = ELP_SO_KEEP, Unknown = ELP_SO_UNKNOWN, Unsorted = ELP_SO_UNSORTED, Queryname = ELP_SO_QUERYNAME, Coordinate = ELP_SO_COORDINATE };

struct Options {                       // the `elprep filter` flags of the path (cmd/filter.go:435-481)
    int Device = 0, MaxCycle = 500, QuantizeLevels = 0; std::vector<uint8_t> SQQ; std::string TablenamePrefix = "GATK";
    int OpticalDuplicatesPixelDistance = 100; bool MarkDuplicates = false, AlsoOpticals = false;
};

struct DuplicatesCtr {                 // filters/mark-optical-duplicates.go:96-110 (the exported counters + the derived numbers)
    int64_t UnpairedReadsExamined = 0, ReadPairsExamined = 0, SecondaryOrSupplementaryReads = 0, UnmappedReads = 0, UnpairedReadDuplicates = 0,
            ReadPairDuplicates = 0, ReadPairOpticalDuplicates = 0, EstimatedLibrarySize = 0;
    double PercentDuplication = 0;
};

// the columns of one batch, exactly elp_batch's layout
struct Columns {
    std::vector<int32_t> refid, pos, nref, pnext, tlen, rg, lseq; std::vector<uint16_t> flag; std::vector<uint8_t> mapq, opt, qname, seq, qual;
    std::vector<uint64_t> qname_off{0}, cigar_off{0}; std::vector<uint32_t> cigar;
};

class DeviceSam {
  public:
    DeviceSam(const Header& header, const Options& opts) : header_(header), opts_(opts) {
        for (size_t i = 0; i < header.sq.size(); i++) refid_[header.sq[i].SN] = (int32_t)i;
        for (size_t i = 0; i < header.rg.size(); i++) rgidx_[header.rg[i].ID] = (int32_t)i;
        std::vector<const char*> names, ids, lbs, pus; std::vector<int32_t> lens;
        for (auto& s : header.sq) { names.push_back(s.SN.c_str()); lens.push_back(s.LN); }
        for (auto& g : header.rg) { ids.push_back(g.ID.c_str()); lbs.push_back(g.hasLB ? g.LB.c_str() : nullptr); pus.push_back(g.hasPU ? g.PU.c_str() : nullptr); }
        elp_config cfg; std::memset(&cfg, 0, sizeof cfg);
        cfg.device = opts.Device; cfg.n_contigs = (int32_t)header.sq.size(); cfg.contig_names = names.data(); cfg.contig_lengths = lens.data();
        cfg.n_read_groups = (int32_t)header.rg.size(); cfg.rg_id = ids.data(); cfg.rg_lb = lbs.data(); cfg.rg_pu = pus.data();
        cfg.max_cycle = opts.MaxCycle; cfg.quantize_levels = opts.QuantizeLevels; cfg.sqq = opts.SQQ.empty() ? nullptr : opts.SQQ.data(); cfg.n_sqq = (int32_t)opts.SQQ.size();
        cfg.tablename_prefix = opts.TablenamePrefix.c_str(); cfg.optical_pixel_distance = opts.OpticalDuplicatesPixelDistance;
        const int rc = elp_create(&cfg, &ctx_);
        if (rc != ELP_OK) throw Panic(rc, elp_last_error(nullptr));      // ELP_ENODEVICE included: there is no CPU path to fall back to
    }
    ~DeviceSam() { if (ctx_) elp_destroy(ctx_); }
    DeviceSam(const DeviceSam&) = delete; DeviceSam& operator=(const DeviceSam&) = delete;

    // filters.AddREFID (filters/simple-filters.go:208-231): "*" and names that are no @SQ give -1; RNEXT "=" is RNAME
    static int32_t lookup(const std::unordered_map<std::string, int32_t>& t, const std::string& name) { auto it = t.find(name); return it == t.end() ? -1 : it->second; }

    // []*sam.Alignment -> columns (what a pargo stage does per batch, sam/filter-pipeline.go:92-104).  Static: usable without a device.
    static Columns marshal(const Header& header, const std::vector<Alignment>& batch) {
        std::unordered_map<std::string, int32_t> ref, rgi;
        for (size_t i = 0; i < header.sq.size(); i++) ref[header.sq[i].SN] = (int32_t)i;
        for (size_t i = 0; i < header.rg.size(); i++) rgi[header.rg[i].ID] = (int32_t)i;
        static const char ops[] = "MIDNSHP=X", bases[] = "=ACMGRSVTWYHKDBN";
        Columns c;
        for (const Alignment& a : batch) {
            const int32_t r = lookup(ref, a.RNAME);
            c.refid.push_back(r); c.nref.push_back(a.RNEXT == "=" ? r : lookup(ref, a.RNEXT));
            c.pos.push_back(a.POS); c.pnext.push_back(a.PNEXT); c.tlen.push_back(a.TLEN); c.flag.push_back(a.FLAG); c.mapq.push_back(a.MAPQ);
            int32_t g = -1;
            if (a.hasRG) { g = lookup(rgi, a.RG); if (g < 0) throw Panic(ELP_EINVAL, "RG:Z value " + a.RG + " is not an @RG ID"); }
            c.rg.push_back(g); c.opt.push_back(a.sr ? (uint8_t)ELP_OPT_SR : (uint8_t)0);
            c.qname.insert(c.qname.end(), a.QNAME.begin(), a.QNAME.end()); c.qname_off.push_back(c.qname.size());
            for (const CigarOperation& op : a.CIGAR) {            // adjacent identical operations merge, as slowScanCigarString does (sam-types.go:700-724)
                const char* p = std::strchr(ops, op.Operation);
                if (!p || !op.Operation) throw Panic(ELP_EINVAL, std::string("unknown CIGAR operation ") + op.Operation);
                const uint32_t code = (uint32_t)(p - ops);
                if (c.cigar.size() > c.cigar_off.back() && (c.cigar.back() & 15u) == code) c.cigar.back() = (((c.cigar.back() >> 4) + (uint32_t)op.Length) << 4) | code;
                else c.cigar.push_back(((uint32_t)op.Length << 4) | code);
            }
            c.cigar_off.push_back(c.cigar.size());
            if (a.QUAL.size() != a.SEQ.size()) throw Panic(ELP_EINVAL, "QUAL must have one value per base");
            c.lseq.push_back((int32_t)a.SEQ.size());
            for (size_t i = 0; i < a.SEQ.size(); i += 2) {
                auto nib = [&](char ch) -> uint8_t { const char* p = ch ? std::strchr(bases, ch) : nullptr; return p ? (uint8_t)(p - bases) : (uint8_t)15; };
                c.seq.push_back((uint8_t)((nib(a.SEQ[i]) << 4) | (i + 1 < a.SEQ.size() ? nib(a.SEQ[i + 1]) : 0)));
            }
            c.qual.insert(c.qual.end(), a.QUAL.begin(), a.QUAL.end());
        }
        return c;
    }

    // the receive side of AddNodes: one batch of alignments, kept so that RunPipeline can hand them back in output order
    void AddNodes(const std::vector<Alignment>& batch) {
        if (batch.empty()) return;
        Columns c = marshal(header_, batch);
        static const uint8_t zero8[1] = {0}; static const uint32_t zero32[1] = {0};
        elp_batch b; std::memset(&b, 0, sizeof b);
        b.n = batch.size(); b.refid = c.refid.data(); b.pos = c.pos.data(); b.flag = c.flag.data(); b.mapq = c.mapq.data(); b.nref = c.nref.data(); b.pnext = c.pnext.data();
        b.tlen = c.tlen.data(); b.rg = c.rg.data(); b.qname_off = c.qname_off.data(); b.qname = c.qname.empty() ? zero8 : c.qname.data(); b.cigar_off = c.cigar_off.data();
        b.cigar = c.cigar.empty() ? zero32 : c.cigar.data(); b.l_seq = c.lseq.data(); b.seq = c.seq.empty() ? zero8 : c.seq.data(); b.qual = c.qual.empty() ? zero8 : c.qual.data();
        b.opt_flags = c.opt.data();
        check(elp_append_batch(ctx_, &b));
        alns_.insert(alns_.end(), batch.begin(), batch.end());
    }
    // the Finalize of AddNodes: By(CoordinateLess / QNAMELess).ParallelStableSort + filters.MarkDuplicates (+ MarkOpticalDuplicates)
    void Finalize(SortingOrder order) { check(elp_sort_markdup(ctx_, (int)order, opts_.MarkDuplicates ? (opts_.AlsoOpticals ? ELP_MARKDUP_OPTICAL : ELP_MARKDUP) : 0)); }

    void SetReference(int contig, const std::string& bases) { check(elp_set_reference(ctx_, contig, reinterpret_cast<const uint8_t*>(bases.data()), bases.size())); }
    void SetKnownSites(int contig, const std::vector<int32_t>& startEnd) { check(elp_set_known_sites(ctx_, contig, startEnd.empty() ? nullptr : startEnd.data(), startEnd.size() / 2, 0)); }
    void Recalibrate() { check(elp_bqsr_gather(ctx_)); }
    void FinalizeBQSRTables(const std::string& recalFile = "") { check(elp_bqsr_finalize(ctx_, recalFile.empty() ? nullptr : recalFile.c_str())); }
    void ApplyBQSR() { check(elp_bqsr_apply(ctx_)); }
    void PrintBQSRTablesToIntermediateFile(const std::string& name) { check(elp_bqsr_tables_write_elrecal(ctx_, name.c_str())); }
    void LoadAndCombineBQSRTables(const std::vector<std::string>& files) { check(elp_bqsr_tables_clear(ctx_)); for (auto& f : files) check(elp_bqsr_tables_add_elrecal(ctx_, f.c_str())); }

    std::map<std::string, DuplicatesCtr> MarkOpticalDuplicates() {
        std::map<std::string, DuplicatesCtr> res;
        for (int32_t slot = 0; slot < elp_optical_n_libraries(ctx_); slot++) {
            elp_dup_metrics m; check(elp_optical_metrics(ctx_, slot, &m));
            DuplicatesCtr d; d.UnpairedReadsExamined = m.unpaired_reads_examined; d.ReadPairsExamined = m.read_pairs_examined;
            d.SecondaryOrSupplementaryReads = m.secondary_or_supplementary_reads; d.UnmappedReads = m.unmapped_reads; d.UnpairedReadDuplicates = m.unpaired_read_duplicates;
            d.ReadPairDuplicates = m.read_pair_duplicates; d.ReadPairOpticalDuplicates = m.read_pair_optical_duplicates; d.EstimatedLibrarySize = m.estimated_library_size;
            d.PercentDuplication = m.percent_duplication;
            res[elp_optical_library_name(ctx_, slot)] = d;
        }
        return res;
    }
    void PrintDuplicatesMetrics(const std::string& path, const std::string& commandLine, const std::string& startedOn) {
        check(elp_print_duplicates_metrics(ctx_, path.c_str(), commandLine.c_str(), startedOn.c_str()));
    }

    // the source side of the write phase: the alignments in output order with the FLAG and QUAL the device computed
    void RunPipeline(const std::function<void(const Alignment&)>& sink, uint64_t chunk = 1u << 18) {
        const uint64_t n = elp_n_reads(ctx_);
        std::vector<uint64_t> idx(chunk), off(chunk + 1); std::vector<uint16_t> flag(chunk); std::vector<uint8_t> qual;
        for (uint64_t first = 0; first < n; first += chunk) {
            const uint64_t m = std::min(chunk, n - first);
            qual.resize(elp_fetch_qual_bytes(ctx_, first, m) + 1);
            check(elp_fetch(ctx_, first, m, idx.data(), flag.data(), off.data(), qual.data(), qual.size()));
            for (uint64_t k = 0; k < m; k++) {
                Alignment& a = alns_[idx[k]];
                a.FLAG = flag[k]; a.QUAL.assign(qual.begin() + off[k], qual.begin() + off[k + 1]);
                sink(a);
            }
        }
    }
    uint64_t Len() const { return elp_n_reads(ctx_); }
    elp_ctx* Context() { return ctx_; }

  private:
    void check(int rc) { if (rc != ELP_OK) throw Panic(rc, elp_last_error(ctx_)); }
    elp_ctx* ctx_ = nullptr; Header header_; Options opts_; std::vector<Alignment> alns_;
    std::unordered_map<std::string, int32_t> refid_, rgidx_;
};

}  // namespace elprep
