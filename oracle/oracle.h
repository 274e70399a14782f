/*
 * oracle/oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the elPrep 5.1.3 hot path: coordinate sort ->
 * mark duplicates -> BQSR gather -> finalize -> apply.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may load this library; the product (libelprep_b200.so) never does.
 *
 * PARITY UNPINNED: the reference ships no test, fixture or golden vector for
 * sort / mark-duplicates / BQSR (its single test file pins only
 * intervals.Flatten/Overlap/Intersect, reused in tests/test_oracle_kat.py),
 * and no Go toolchain exists in this image, so the Go source is the only
 * specification.  Every function cites the reference lines it restates.
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Columnar view of a set of sam.Alignment records (sam/sam-types.go:289-331).
 * cigar: BAM encoding (len<<4 | op), op index into "MIDNSHP=X".
 * seq: 4-bit BAM nibbles, high nibble first, each read byte-aligned at seq_off[i].
 * qual: phred bytes without +33, read i at qual_off[i], length lseq[i]. */
typedef struct {
    int64_t n;
    const int32_t *refid;   /* REFID temp (filters/simple-filters.go:208-231) */
    const int32_t *pos;
    uint16_t *flag;         /* mutated by mark duplicates */
    const uint8_t *mapq;
    const int32_t *nref;    /* NextREFID temp */
    const int32_t *pnext;
    const int32_t *tlen;
    const int32_t *rg;      /* index into the @RG table, -1 = no RG tag */
    const uint64_t *qname_off; const uint8_t *qname;
    const uint64_t *cigar_off; const uint32_t *cigar;
    const int32_t *lseq;
    const uint64_t *seq_off;  const uint8_t *seq;
    const uint64_t *qual_off; uint8_t *qual; /* mutated by apply */
    const uint8_t *opt_flags; /* may be NULL; bit 0: the read carries the sr tag (sam/split-merge.go:286-293) */
} orc_reads;

typedef struct {
    int32_t n_contigs; const int32_t *contig_len;     /* @SQ LN */
    int32_t n_rg;
    const int32_t *rg_lib;  /* library id per @RG (equal LB strings share an id), -1 = no LB */
    const int32_t *rg_cov;  /* read-group covariate id per @RG (PU if present else ID; filters/bqsr.go:35-51) */
    int32_t n_cov;
} orc_header;

/* dense BQSR tables; an entry "exists" in the reference's map iff obs > 0 */
#define ORC_NQ 256
#define ORC_NCTX 16
typedef struct {
    int32_t n_cov, max_cycle;
    int64_t *q_obs, *q_mis;   /* [n_cov][256] */
    int64_t *c_obs, *c_mis;   /* [n_cov][256][2*max_cycle+1], cycle index = cycle + max_cycle */
    int64_t *x_obs, *x_mis;   /* [n_cov][256][16], context index = key>>4 */
    uint8_t *q_emp, *c_emp, *x_emp; /* EmpiricalQuality after finalize, same shapes */
} orc_tables;

/* filters/mark-duplicates.go:57-110 */
int32_t orc_phred_score(const uint8_t *qual, int32_t n, int *invalid);
int32_t orc_unclipped_position(int32_t pos, int reversed, const uint32_t *cigar, int32_t ncigar);
/* sam/sam-types.go:408-473 */
uint16_t orc_mod_flag(uint16_t flag);
int orc_coordinate_less(const orc_reads *r, int64_t a, int64_t b);
/* sam/sam-types.go:599-641 + sam/filter-pipeline.go:113-117: perm[k] = index of k-th record in sorted order */
int orc_coordinate_sort(const orc_reads *r, int64_t *perm, int n_threads);
/* sam/sam-types.go:479-481: By(QNAMELess) stable sort */
int orc_queryname_sort(const orc_reads *r, int64_t *perm);
/* filters/mark-duplicates.go:406-445 (+ adapt/classifyFragment/classifyPair). Sets 0x400 in r->flag.
 * upos_out/score_out optional (may be NULL). returns 0, or -1 on "Invalid QUAL character". */
int orc_mark_duplicates(const orc_reads *r, const orc_header *h, int n_threads, int32_t *upos_out, int32_t *score_out);

/* filters/mark-optical-duplicates.go:95-110 (DuplicatesCtr) + the derived metrics (:519-581) */
typedef struct {
    int64_t unpaired_reads_examined, read_pairs_examined, secondary_or_supplementary, unmapped_reads,
            unpaired_read_duplicates, read_pair_duplicates, read_pair_optical_duplicates;
    int64_t estimated_library_size; double percent_duplication; double roi[100]; int32_t has_roi;
} orc_dup_metrics;
typedef struct orc_optical_result orc_optical_result;
/* MarkDuplicates(alsoOpticals = true) + MarkOpticalDuplicates (mark-optical-duplicates.go:468-517) with the reads visited
 * in `order` (NULL = as given). Slot 0 = "Unknown Library", slot l+1 = library id l. error(): 0, -1 invalid QUAL,
 * -2 "origin for duplicate read pair unknown", -3 a QNAME tile/x/y field that strconv.ParseInt rejects. */
orc_optical_result *orc_markdup_optical(const orc_reads *r, const orc_header *h, int n_threads, const int64_t *order, int pixel_distance);
int orc_optical_error(const orc_optical_result *res);
int orc_optical_slots(const orc_optical_result *res);
void orc_optical_get(const orc_optical_result *res, int slot, orc_dup_metrics *out);
/* which: 0 duplicatesCountHistogram, 1 nonOptical..., 2 optical...; returns number of non-zero entries (ascending key) */
int64_t orc_optical_hist(const orc_optical_result *res, int slot, int which, int64_t *keys, int64_t *counts, int64_t cap);
void orc_optical_free(orc_optical_result *res);
void orc_derive_dup_metrics(orc_dup_metrics *m);
/* PrintDuplicatesMetrics (:601-699); libraries printed in ascending name order (the reference iterates a Go map) */
int orc_optical_print(const orc_optical_result *res, const char *const *lib_names, const char *path, const char *command_line, const char *started_on);

/* intervals/intervals.go:103-173 (KATs from intervals/intervals_test.go) */
int64_t orc_flatten(int32_t *se, int64_t n);            /* in place, returns new count */
int orc_overlap(const int32_t *se, int64_t n, int32_t start, int32_t end);
void orc_intersect(const int32_t *se, int64_t n, int32_t start, int32_t end, int64_t *lo, int64_t *hi);

/* filters/bqsr.go:467-551 (+ filters/utils.go:130-534). ref: concatenated contig bases (1 B/base) with ref_off[n_contigs+1];
 * sites: flattened sorted (start,end) pairs concatenated with site_off[n_contigs+1] (in intervals). returns 0 or <0 error. */
int orc_bqsr_gather(const orc_reads *r, const orc_header *h, const uint8_t *ref, const uint64_t *ref_off,
                    const int32_t *sites, const uint64_t *site_off, orc_tables *t, int n_threads);
/* filters/bqsr.go:677-694 */
void orc_bqsr_finalize(orc_tables *t);
/* filters/bqsr.go:936-1006. quantize_levels 0 = identity. sqq may be NULL. returns 0 or <0 */
int orc_bqsr_apply(const orc_reads *r, const orc_header *h, const orc_tables *t, int quantize_levels,
                   const uint8_t *sqq, int n_sqq, int n_threads);
/* filters/print-bqsr.go:269-298; cov_names[n_cov] */
int orc_bqsr_report(const orc_tables *t, const char *const *cov_names, const char *prefix, const char *path);
/* filters/bqsr.go:655-674: per-covariate combined entry (ascending-qual order) */
void orc_combined(const orc_tables *t, int cov, double *reported_q, int64_t *obs, int64_t *mis, uint8_t *emp, int *exists);

/* single-read probes used by the known-answer tests (Appendix C of SURVEY.md) */
/* clip one read as bqsr gather does; out: kept [lo,hi) in original read coordinates, new POS, new cigar (cap ops). returns new ncigar, -1 if read dropped */
int orc_probe_clip(int32_t pos, uint16_t flag, int32_t pnext, int32_t tlen, int32_t refid, int32_t nref,
                   const uint32_t *cigar, int32_t ncigar, int32_t lseq,
                   int32_t *lo, int32_t *hi, int32_t *newpos, uint32_t *newcigar, int cap);
/* filters/utils.go:267-349 */
int orc_probe_readcoord(const uint32_t *cigar, int32_t ncigar, int soft_start, int ref_index, int tail_right, int *ok);
/* context keys (filters/bqsr.go:64-146,312-362) for one read; keys[lseq]; returns count (0 if all quals <= 2) */
int orc_probe_context(const uint8_t *seq_nibbles, const uint8_t *qual, int32_t lseq, int reversed, int32_t *keys);
void orc_probe_cycle(uint16_t flag, int32_t lseq, int32_t *cycles);
/* filters/bqsr.go:623-649 */
uint8_t orc_empirical_quality(int64_t obs, int64_t mis, double prior);
double orc_prior_cache(int d);
void orc_static_quantized(const uint8_t *sqq, int n, uint8_t *out254);
void orc_quantized(const orc_tables *t, int levels, int64_t *obs94, uint8_t *scores94);

#ifdef __cplusplus
}
#endif
#endif
