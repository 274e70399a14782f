"""CPU tests of the host-side mirror of the reference's operator API (composition, sorting-order logic, multi-rank partition)."""
import numpy as np

from elprep_b200 import filters, sam


def test_compose_filters_skips_nil():
    h = sam.Header(sq=[{"SN": "c", "LN": 10}])
    calls = []
    f1 = lambda hdr: None                                   # a Filter may return nil (sam/filter-pipeline.go:39-40,166-170)
    f2 = lambda hdr: (lambda b: calls.append("f2") or b)
    out = filters.compose_filters(h, [f1, None, f2])
    assert len(out) == 1 and out[0]("x") == "x" and calls == ["f2"]


def test_effective_sorting_order():
    h = sam.Header(sq=[{"SN": "c", "LN": 10}], so=sam.Coordinate)
    assert filters.effective_sorting_order(sam.Coordinate, h, sam.Coordinate) == sam.Keep          # already sorted: skip (:214-217)
    h = sam.Header(sq=[{"SN": "c", "LN": 10}], so=sam.Unsorted)
    assert filters.effective_sorting_order(sam.Coordinate, h, sam.Unsorted) == sam.Coordinate and h.HDSO() == sam.Coordinate
    h = sam.Header(sq=[{"SN": "c", "LN": 10}], so=sam.Queryname)
    assert filters.effective_sorting_order(sam.Keep, h, sam.Queryname) == sam.Keep                 # Keep -> original order, unchanged


def test_markduplicates_needs_rg_id():
    h = sam.Header(sq=[{"SN": "c", "LN": 10}], rg=[{"LB": "x"}])
    md, _, _ = filters.MarkDuplicates(False)
    try:
        md(h)
        assert False
    except ValueError as e:
        assert "Missing mandatory ID entry" in str(e)        # filters/mark-duplicates.go:419


def test_header_tables():
    h = sam.Header(sq=[{"SN": "a", "LN": 5}, {"SN": "b", "LN": 7}], rg=[{"ID": "r1", "LB": "L", "PU": "p"}, {"ID": "r2", "LB": "L"}, {"ID": "r3"}, {"ID": "r4", "PU": "p"}])
    lib, names = h.rg_lib_ids()
    assert lib.tolist() == [0, 0, -1, -1] and names == ["L"]
    cov, cn = h.rg_cov_ids()
    assert cov.tolist() == [0, 1, 2, 0] and cn == ["p", "r2", "r3"]                                # PU if present else ID (bqsr.go:35-51)
    assert h.refid_table() == {"*": -1, "a": 0, "b": 1}
    b = sam.AlignmentBatch.from_records(h, [dict(QNAME="q", RNAME="b", RNEXT="=", POS=3, CIGAR="2M1M", SEQ="ACG", QUAL=[1, 2, 3], RG="r2")])
    assert b.refid[0] == 1 and b.nref[0] == 1 and sam.decode_cigar(b.cigar) == "3M"               # adjacent identical ops merge (sam-types.go:708-710)


def test_simple_filters_as_column_predicates():
    """filters/simple-filters.go:71-103,131-133,332-347 over an AlignmentBatch"""
    from elprep_b200 import filters
    h = sam.Header(sq=[{"SN": "chr1", "LN": 1000}], rg=[{"ID": "rg1"}])
    recs = [dict(QNAME="a", FLAG=0, RNAME="chr1", POS=5, MAPQ=60, CIGAR="4M", SEQ="ACGT", QUAL=[30] * 4),
            dict(QNAME="b", FLAG=4, RNAME="*", POS=0, MAPQ=0, CIGAR="*", SEQ="ACGT", QUAL=[30] * 4),
            dict(QNAME="c", FLAG=0, RNAME="chr1", POS=0, MAPQ=10, CIGAR="2S2M", SEQ="ACGT", QUAL=[30] * 4),      # mapped by FLAG, POS 0
            dict(QNAME="d", FLAG=0x400, RNAME="chr1", POS=9, MAPQ=29, CIGAR="2M1I1M", SEQ="ACGT", QUAL=[30] * 4),
            dict(QNAME="e", FLAG=16, RNAME="chr1", POS=9, MAPQ=30, CIGAR="1H4M", SEQ="ACGT", QUAL=[30] * 4)]
    b = sam.AlignmentBatch.from_records(h, recs)
    names = lambda x: [x.qname_str(i) for i in range(x.n)]
    assert names(filters.RemoveUnmappedReads(h)(b)) == ["a", "c", "d", "e"]
    assert names(filters.RemoveUnmappedReadsStrict(h)(b)) == ["a", "d", "e"]
    assert names(filters.RemoveNonExactMappingReads(h)(b)) == ["a", "b", "c"]
    assert names(filters.RemoveMappingQualityLessThan(30)(h)(b)) == ["a", "e"] and filters.RemoveMappingQualityLessThan(0) is None
    assert names(filters.RemoveDuplicateReads(h)(b)) == ["a", "b", "c", "e"]
    assert filters.RemoveUnmappedReads(h)(b.take([0, 3])) .n == 2


def test_spread_exchange_degenerate_cases():
    """elprep_b200.multi with one worker (nothing to exchange) and with a batch that has no cross-group pairs"""
    from elprep_b200 import multi, synth
    contigs = [("c1", 200_000), ("c2", 100_000)]
    w = synth.make_workload(500, contigs, seed=5, cross_contig_frac=0.0, want_reference=False)
    calls = []

    def md(batch, header):
        calls.append(batch.n)
        return batch.flag.copy(), None
    owner1 = multi.owner_table(w.header, multi.contig_groups(contigs, 1))
    before = w.batch.flag.copy()
    assert multi.exchange_spread_duplicates(w.batch, w.header, owner1, 0, 1, md, lambda o: [o]) is None
    assert np.array_equal(before, w.batch.flag) and calls == []
    owner2 = multi.owner_table(w.header, multi.contig_groups(contigs, 2))
    assert sorted(set(owner2.tolist())) == [0, 1]
    idx, powner = multi.spread_reads(w.batch, owner2, 0)
    assert idx.size == 0 and powner.size == 0
    own = [multi.partition(w.batch, owner2, r, 2) for r in range(2)]
    assert sorted(np.concatenate(own).tolist()) == list(range(w.batch.n))          # every read has exactly one owner
    assert bool((w.batch.refid[own[1]] < 0).sum() == (w.batch.refid < 0).sum())       # unmapped reads go to the last rank
