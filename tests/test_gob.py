"""encoding/gob intermediate files: .elrecal and duplication metrics.  UNVERIFIED against a Go binary (none in this
image): the specification's own byte example pins the primitives, and the library's writer / reader (elprep_b200/csrc/gob.cpp) are checked
against the independent restatement in tests/gobref.py."""
import numpy as np
import pytest

import gobref


def test_spec_example_bytes():
    """the byte example of the encoding/gob documentation: type Point struct{X, Y int}; Point{22, 33}"""
    assert gobref.encode_point(22, 33) == gobref.POINT_EXAMPLE
    assert gobref.decode(gobref.POINT_EXAMPLE) == {"X": 22, "Y": 33}
    assert gobref.enc_int(-129) == bytes([0xfe, 0x01, 0x01]) and gobref.enc_uint(256) == bytes([0xfe, 0x01, 0x00])   # "(-129) is FE 01 01", "256 is FE 01 00"


def test_restatement_round_trip():
    t = [{(30, 0, "rgA"): (0, 1000, 7), (2, 0, "rgB"): (0, 5, 0)}, {(30, -150, "rgA"): (0, 3, 1), (30, 7, "rgA"): (29, 1 << 40, 12)}, {(30, 66, "rgA"): (0, 9, 0)}]
    d = gobref.decode(gobref.encode_elrecal(t))
    got = [{(k.get("Qual", 0), k.get("Covariate", 0), k.get("ReadGroup", "")): (v.get("EmpiricalQuality", 0), v.get("Observations", 0), v.get("Mismatches", 0)) for k, v in d[name]}
           for name in ("QualityScores", "Cycles", "Contexts")]
    assert got == t


@pytest.mark.gpu
def test_elrecal_written_by_the_library_and_modes(tmp_path):
    """--bqsr-tables-only on two halves of the reads, then --bqsr-apply over both files == one --bqsr run (tables, report, QUAL)"""
    from elprep_b200 import device, synth
    from util import gpu_pipeline
    w = synth.make_workload(6000, [("chr20", 300_000), ("chr21", 200_000)], seed=77)
    whole = gpu_pipeline(w, keep_ctx=True)
    n = w.batch.n
    files = []
    for part, idx in enumerate((np.arange(0, n // 2), np.arange(n // 2, n))):
        ctx = device.Context(w.header)
        for ci in range(2):
            ctx.set_reference(ci, w.contig_bases[ci]); ctx.set_known_sites(ci, w.sites[ci])
        ctx.append(w.batch.take(idx))
        ctx.sort_markdup(device.SO_COORDINATE, False)
        ctx.bqsr_gather()
        f = str(tmp_path / f"part{part}.elrecal")
        ctx.write_elrecal(f); files.append(f)
        dense = ctx.tables_get()
        d = gobref.decode(open(f, "rb").read())
        names = ctx.cov_names()
        for name, cols in (("QualityScores", lambda cov: 0), ("Cycles", lambda cov: 1 + cov + 500), ("Contexts", lambda cov: 1 + 1001 + (cov >> 4))):
            seen = 0
            for k, v in d[name]:
                cv, q = names.index(k["ReadGroup"]), k.get("Qual", 0)
                assert dense[cv, q, cols(k.get("Covariate", 0)), 0] == v["Observations"] and dense[cv, q, cols(k.get("Covariate", 0)), 1] == v.get("Mismatches", 0)
                seen += 1
            lo, hi = {"QualityScores": (0, 1), "Cycles": (1, 1002), "Contexts": (1002, 1018)}[name]
            assert seen == int((dense[:, :, lo:hi, 0] > 0).sum())
        ctx.close()
    # the duplicate-marked whole run, without --mark-duplicates in the parts above the tables differ; compare against a whole run without it
    ref = gpu_pipeline(w, markdup=False, keep_ctx=True)
    a = ref["ctx"]
    b = device.Context(w.header)
    b.append(w.batch); b.sort_markdup(device.SO_COORDINATE, False)
    b.tables_clear()
    for f in files:
        b.add_elrecal(f)
    assert np.array_equal(b.tables_get(), ref["tables"])
    # a file written by the restatement is read the same way
    g = str(tmp_path / "ref.elrecal")
    t = ref["tables"]; names = a.cov_names()
    tabs = [{}, {}, {}]
    for cv, q, col in zip(*np.nonzero(t[:, :, :, 0])):
        o, m = int(t[cv, q, col, 0]), int(t[cv, q, col, 1])
        if col == 0: tabs[0][(int(q), 0, names[cv])] = (0, o, m)
        elif col < 1002: tabs[1][(int(q), int(col) - 1 - 500, names[cv])] = (0, o, m)
        else: tabs[2][(int(q), 2 | ((int(col) - 1002) << 4), names[cv])] = (0, o, m)
    open(g, "wb").write(gobref.encode_elrecal(tabs))
    b.tables_clear(); b.add_elrecal(g)
    assert np.array_equal(b.tables_get(), t)
    rep = str(tmp_path / "b.recal")
    b.bqsr_finalize(rep); b.bqsr_apply()
    assert open(rep).read() == ref["report"]
    assert np.array_equal(b.fetch()[3][:ref["qual"].size], ref["qual"])
    a.close(); b.close(); whole["ctx"].close()


@pytest.mark.gpu
def test_duplicates_metrics_gob(tmp_path):
    from elprep_b200 import device, synth, _lib
    w = synth.make_workload(5000, [("chr20", 300_000)], seed=5, optical_frac=0.5, dup_frac=0.3)
    ctx = device.Context(w.header)
    ctx.append(w.batch); ctx.sort_markdup(device.SO_COORDINATE, _lib.MARKDUP_OPTICAL)
    m = ctx.optical_metrics(); libs = ctx.optical_libraries()
    f = str(tmp_path / "m.gob"); ctx.optical_write_gob(f)
    d = dict(gobref.decode(open(f, "rb").read()))
    keys = ("unpaired_reads_examined", "read_pairs_examined", "secondary_or_supplementary_reads", "unmapped_reads", "unpaired_read_duplicates", "read_pair_duplicates", "read_pair_optical_duplicates")
    for lib, mm in zip(libs, m):
        if any(mm[k] for k in keys):
            assert [d[lib].get(n, 0) for n in gobref.COUNTERS] == [mm[k] for k in keys]
    # two workers' files sum up (LoadAndCombineDuplicateMetrics)
    other = device.Context(w.header)
    other.optical_add_gob(f); other.optical_add_gob(f)
    m2 = other.optical_metrics()
    for a, b in zip(m, m2):
        for k in keys:
            assert b[k] == 2 * a[k]
    ctx.close(); other.close()
