"""GPU parity: the CUDA path, called through the C ABI, against the oracle on the same seeded inputs.
Bit-exact for FLAG (0x400 bits), output order, table counters, EmpiricalQuality, report text and QUAL bytes."""
import numpy as np
import pytest

from elprep_b200 import sam, synth
from util import gpu_pipeline, oracle_pipeline, oracle_tables_dense

pytestmark = pytest.mark.gpu

SMALL = [("chr20", 600_000), ("chr21", 300_000), ("chrM", 16_569)]


def _ctx():
    from elprep_b200 import device
    return device.Context(sam.Header(sq=[{"SN": "c", "LN": 1000}]))


@pytest.mark.parametrize("n,bits", [(0, 8), (1, 8), (5, 3), (1000, 17), (6144, 34), (6145, 34), (100_000, 52), (1_000_003, 40), (300_000, 64), (200_000, 1)])
def test_radix_sort_u64(n, bits):
    rng = np.random.default_rng(n + bits)
    keys = rng.integers(0, 2 ** 63, size=n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=n, dtype=np.uint64)
    if bits < 64:
        keys &= np.uint64((1 << bits) - 1)
    if n > 100:
        keys[rng.integers(0, n, size=n // 3)] = keys[0]      # many duplicates: stability matters
    vals = np.arange(n, dtype=np.uint32)
    ctx = _ctx()
    k2, v2 = ctx.debug_sort_u64(keys, vals, bits)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k2, keys[order]) and np.array_equal(v2, vals[order])
    ctx.close()


@pytest.mark.parametrize("n,bits", [(0, 70), (3, 70), (4096, 87), (4097, 87), (250_000, 128), (100_000, 65)])
def test_radix_sort_u128(n, bits):
    rng = np.random.default_rng(n + bits)
    lo = rng.integers(0, 2 ** 63, size=n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=n, dtype=np.uint64)
    hi = rng.integers(0, 2 ** 63, size=n, dtype=np.uint64)
    hb = bits - 64
    hi &= np.uint64((1 << hb) - 1) if hb < 64 else np.uint64(2 ** 64 - 1)
    if n > 100:
        dup = rng.integers(0, n, size=n // 3); hi[dup] = hi[0]; lo[dup] = lo[0]
    vals = np.arange(n, dtype=np.uint32)
    ctx = _ctx()
    h2, l2, v2 = ctx.debug_sort_u128(hi, lo, vals, bits)
    order = np.lexsort((lo, hi))      # stable, hi most significant
    assert np.array_equal(h2, hi[order]) and np.array_equal(l2, lo[order]) and np.array_equal(v2, vals[order])
    ctx.close()


def _compare(w, g, o, bqsr=True):
    assert np.array_equal(g["perm"], o["perm"]), "output order differs"
    assert np.array_equal(g["flag"], o["flag"]), "FLAG differs"
    assert np.array_equal(g["qual_off"], o["qual_off"])
    if bqsr:
        d, e = oracle_tables_dense(o["tables"], 500)
        assert np.array_equal(g["tables"], d), "BQSR table counters differ"
        assert np.array_equal(g["emp"], e), "EmpiricalQuality differs"
        assert g["report"] == o["report"], "recalibration report text differs"
    assert np.array_equal(g["qual"], o["qual"]), "QUAL bytes differ"


def test_adapt_matches_oracle(orc):
    w = synth.make_workload(20_000, SMALL, seed=11, want_reference=False)
    from elprep_b200 import device
    ctx = device.Context(w.header)
    ctx.append(w.batch)
    up, sc = ctx.debug_adapt()
    b = w.batch.copy()
    ou, os_ = orc.mark_duplicates(b, w.header, want_adapt=True)
    enter = (w.batch.flag & 0x904) == 0
    assert np.array_equal(up[enter], ou[enter]) and np.array_equal(sc[enter], os_[enter])
    ctx.close()


@pytest.mark.parametrize("seed,kw", [(1, {}), (2, dict(exome=True)), (3, dict(wide_quals=True)), (4, dict(dup_frac=0.5, optical_frac=0.5)),
                                      (5, dict(unmapped_frac=0.3)), (6, dict(n_rg=1)), (7, dict(n_rg=0))])
def test_full_path_small(seed, kw):
    w = synth.make_workload(15_000, SMALL, seed=seed, **kw)
    bqsr = kw.get("n_rg", 4) != 0     # without read groups BQSR panics in the reference; sort+markdup still run
    g = gpu_pipeline(w, bqsr=bqsr, n_batches=3)
    o = oracle_pipeline(w, bqsr=bqsr)
    _compare(w, g, o, bqsr)
    assert int(((g["flag"] & 0x400) != 0).sum()) > 0 or kw.get("unmapped_frac", 0) > 0.2


def test_full_path_c1_shape():
    """config[0] shape (single contig), 200k reads: sort + markdup + BQSR"""
    w = synth.make_workload(100_000, [("chr20", 6_444_416)], seed=20260924)
    g = gpu_pipeline(w, n_batches=4)
    o = oracle_pipeline(w, threads=8)
    _compare(w, g, o)


def test_no_sort_keep_order():
    w = synth.make_workload(5_000, SMALL, seed=9)
    g = gpu_pipeline(w, sort=False)
    o = oracle_pipeline(w, sort=False)
    _compare(w, g, o)
    assert np.array_equal(g["perm"], np.arange(w.batch.n, dtype=np.uint64))


def test_sqq_and_quantize():
    w = synth.make_workload(8_000, SMALL, seed=21)
    g = gpu_pipeline(w, quantize_levels=4, sqq=[10, 20, 30])
    o = oracle_pipeline(w, quantize_levels=4, sqq=[10, 20, 30])
    _compare(w, g, o)


def test_long_tie_runs_and_unmapped_block():
    """many reads at identical (refid,pos,strand) and a large unmapped block: exercises the long-run tie-break path"""
    w = synth.make_workload(6_000, [("chr20", 3_000)], seed=31, unmapped_frac=0.4, dup_frac=0.6)
    g = gpu_pipeline(w, bqsr=False)
    o = oracle_pipeline(w, bqsr=False)
    _compare(w, g, o, bqsr=False)


def test_empty_and_single():
    h = synth.make_header(SMALL)
    from elprep_b200 import device
    ctx = device.Context(h)
    ctx.sort_markdup()
    idx, flag, qoff, qual = ctx.fetch()
    assert idx.size == 0 and flag.size == 0
    ctx.close()
    w = synth.make_workload(1, SMALL, seed=3)
    _compare(w, gpu_pipeline(w), oracle_pipeline(w))


def _adversarial_records():
    """hand-built reads covering clip/indel/adaptor/N/low-qual-tail corner cases"""
    rng = np.random.default_rng(77)
    recs, L = [], 60
    cigars = ["60M", "5S55M", "55M5S", "3H5S50M5S", "20M3I37M", "20M4D40M", "10S20M2I10M3D18M", "5S20M5D30M5S", "1M1I58M", "58M1I1M",
              "2S10M2D10M2I10M2D10M2I14M", "30M30S", "30S30M", "10M10N40M"]
    for t in range(400):
        c = cigars[t % len(cigars)]
        pos = 1000 + int(rng.integers(0, 400))
        rev = bool(rng.integers(0, 2))
        paired = rng.random() < 0.8
        flag = (0x1 | (0x40 if rng.random() < .5 else 0x80)) if paired else 0
        if rev:
            flag |= 0x10
        elif paired:
            flag |= 0x20
        ins = int(rng.integers(20, 160))
        pnext = pos - ins + 50 if rev else pos + ins - 50
        tlen = (-ins if rev else ins) if paired else 0
        q = rng.choice([2, 2, 5, 6, 12, 23, 37, 40], size=L).astype(int)
        if t % 7 == 0:
            q[:8] = 2
        if t % 11 == 0:
            q[-9:] = 1
        seq = "".join(rng.choice(list("ACGTN"), p=[.24, .24, .24, .24, .04], size=L))
        recs.append(dict(QNAME=f"adv{t:04d}" if paired else f"frag{t:04d}", FLAG=flag, RNAME="chr20", POS=pos, MAPQ=int(rng.choice([0, 30, 60, 255])),
                         CIGAR=c, RNEXT="=" if paired else "*", PNEXT=max(1, pnext) if paired else 0, TLEN=tlen, SEQ=seq, QUAL=[int(x) for x in q],
                         RG=["rg1", "rg2", "rg3", "rg4"][t % 4]))
    return recs


def test_adversarial_clipping_cases():
    contigs = [("chr20", 4_000)]
    base = synth.make_workload(10, contigs, seed=5)       # header + reference + sites
    b = sam.AlignmentBatch.from_records(base.header, _adversarial_records())
    sites = [np.array([[1005, 1005], [1100, 1109], [1200, 1200], [1250, 1300]], dtype=np.int32)]
    w = synth.Workload(base.header, b, base.contig_bases, sites, {})
    _compare(w, gpu_pipeline(w), oracle_pipeline(w))


def test_invalid_qual_is_an_error():
    from elprep_b200 import device
    w = synth.make_workload(200, SMALL, seed=8, want_reference=False)
    w.batch.qual[5] = 100
    ctx = device.Context(w.header)
    ctx.append(w.batch)
    with pytest.raises(device.ElprepError) as ei:
        ctx.sort_markdup()
    assert ei.value.code == -10 and "Invalid QUAL character" in str(ei.value)
    ctx.close()


def test_reference_style_api():
    """the phase order of runBestPracticesPipelineIntermediateSam (cmd/filter.go:142-211) through the mirrored operator API"""
    import os
    import tempfile
    from elprep_b200 import filters
    w = synth.make_workload(6_000, SMALL, seed=41)
    n = w.batch.n
    batches = [w.batch.take(np.arange(a, b)) for a, b in ((0, n // 2), (n // 2, n))]
    reads = filters.DeviceSam()
    md, fragments, pairs = filters.MarkDuplicates(False)
    filters.InputBatches(w.header, batches).RunPipeline(reads, [filters.AddREFID, md], sam.Coordinate)      # phase 1
    assert w.header.HDSO() == sam.Coordinate
    recal = filters.NewBaseRecalibrator(w.sites, w.contig_bases)
    tables = recal.Recalibrate(reads, 500)                                                                  # phase 3
    with tempfile.TemporaryDirectory() as d:
        tables.FinalizeBQSRTables()
        tables.PrintBQSRTables(os.path.join(d, "x.recal"))                                                  # phase 4
        report = open(os.path.join(d, "x.recal")).read()
    reads.RunPipeline(reads, [tables.ApplyBQSR(0, [], 500)], sam.Keep)                                      # phase 5
    out = filters.HostResult()
    reads.RunPipeline(out, [], sam.Keep)                                                                    # phase 6
    o = oracle_pipeline(w)
    assert np.array_equal(out.record_index, o["perm"]) and np.array_equal(out.flag, o["flag"])
    assert np.array_equal(out.qual[:int(out.qual_off[-1])], o["qual"]) and report == o["report"]
