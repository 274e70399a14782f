/* A plain-C client of include/elprep_b200.h: what a cgo/JNI/FFI layer does, without any Python in between.
 * Compiled by tests/test_cabi.py with gcc (this also proves the header is valid C).
 * usage: cabi_client            -> exercises the path on device 0 and prints "order ... flags ..."
 * Without a usable GPU elp_create must fail with ELP_ENODEVICE (the library has no CPU fallback): prints "nodevice". */
#include <stdio.h>
#include <string.h>
#include "elprep_b200.h"

int main(void) {
    const char *names[1] = {"chr1"};
    const int32_t lens[1] = {100000};
    const char *rg_id[1] = {"rg1"}, *rg_lb[1] = {"libA"}, *rg_pu[1] = {NULL};
    elp_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.device = 0; cfg.n_contigs = 1; cfg.contig_names = names; cfg.contig_lengths = lens;
    cfg.n_read_groups = 1; cfg.rg_id = rg_id; cfg.rg_lb = rg_lb; cfg.rg_pu = rg_pu;
    cfg.max_cycle = 500; cfg.tablename_prefix = "GATK"; cfg.optical_pixel_distance = 100;
    elp_ctx *ctx = NULL;
    int rc = elp_create(&cfg, &ctx);
    if (rc == ELP_ENODEVICE) { printf("nodevice: %s\n", elp_last_error(NULL)); return 0; }
    if (rc != ELP_OK) { printf("create failed %d: %s\n", rc, elp_last_error(NULL)); return 1; }

    /* four single-end reads, 4 bases each: two at POS 50 (the lower score becomes a duplicate), one at 20, one unmapped */
    enum { N = 4 };
    int32_t refid[N] = {0, 0, 0, -1}, pos[N] = {50, 50, 20, 0}, nref[N] = {-1, -1, -1, -1}, pnext[N] = {0, 0, 0, 0}, tlen[N] = {0, 0, 0, 0};
    int32_t rg[N] = {0, 0, 0, 0}, lseq[N] = {4, 4, 4, 4};
    uint16_t flag[N] = {0, 0, 16, 4};
    uint8_t mapq[N] = {60, 60, 60, 0};
    uint64_t qname_off[N + 1] = {0, 2, 4, 6, 8}; const uint8_t qname[] = "r1r2r3r4";
    uint64_t cigar_off[N + 1] = {0, 1, 2, 3, 3}; uint32_t cigar[3] = {4u << 4, 4u << 4, 4u << 4};      /* 4M */
    uint8_t seq[N * 2] = {0x12, 0x48, 0x12, 0x48, 0x12, 0x48, 0x12, 0x48};                             /* ACGT */
    uint8_t qual[N * 4] = {30, 30, 30, 30, 20, 20, 20, 20, 25, 25, 25, 25, 2, 2, 2, 2};
    elp_batch b;
    memset(&b, 0, sizeof b);
    b.n = N; b.refid = refid; b.pos = pos; b.flag = flag; b.mapq = mapq; b.nref = nref; b.pnext = pnext; b.tlen = tlen; b.rg = rg;
    b.qname_off = qname_off; b.qname = qname; b.cigar_off = cigar_off; b.cigar = cigar; b.l_seq = lseq; b.seq = seq; b.qual = qual;
    if ((rc = elp_append_batch(ctx, &b)) != ELP_OK) { printf("append failed %d: %s\n", rc, elp_last_error(ctx)); return 1; }
    if ((rc = elp_sort_markdup(ctx, ELP_SO_COORDINATE, ELP_MARKDUP_OPTICAL)) != ELP_OK) { printf("sort failed %d: %s\n", rc, elp_last_error(ctx)); return 1; }
    uint64_t idx[N], qoff[N + 1]; uint16_t oflag[N]; uint8_t oqual[N * 4];
    if ((rc = elp_fetch(ctx, 0, N, idx, oflag, qoff, oqual, sizeof oqual)) != ELP_OK) { printf("fetch failed %d: %s\n", rc, elp_last_error(ctx)); return 1; }
    printf("order %llu %llu %llu %llu flags %u %u %u %u\n", (unsigned long long)idx[0], (unsigned long long)idx[1], (unsigned long long)idx[2],
           (unsigned long long)idx[3], oflag[0], oflag[1], oflag[2], oflag[3]);
    elp_dup_metrics m;
    if ((rc = elp_optical_metrics(ctx, 1, &m)) != ELP_OK) { printf("metrics failed %d: %s\n", rc, elp_last_error(ctx)); return 1; }
    printf("libA unpaired %lld dups %lld unmapped %lld\n", (long long)m.unpaired_reads_examined, (long long)m.unpaired_read_duplicates, (long long)m.unmapped_reads);
    /* wrong phase order is an error return, never an abort */
    rc = elp_bqsr_apply(ctx);
    printf("apply-before-finalize rc %d\n", rc);
    elp_destroy(ctx);
    return 0;
}
