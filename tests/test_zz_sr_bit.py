"""The `sr` tag of `elprep split` (sam/split-merge.go:286-293): a read that carries it is never recalibrated -- recalibrateAln returns
false before anything else (filters/bqsr.go:225-229), so it adds nothing to the tables -- and ApplyBQSR, which does not look at the tag,
still rewrites its QUAL.  The C ABI carries the tag as bit ELP_OPT_SR of the opt_flags column.
CPU: the oracle honours the clause.  GPU: tables, QUAL and the opt_flags read-back against the oracle."""
import numpy as np
import pytest

from elprep_b200 import synth
from util import gpu_pipeline, oracle_pipeline, oracle_tables_dense

SMALL = [("chr20", 600_000), ("chr21", 300_000)]


def _workload(seed=41):
    w = synth.make_workload(5000, SMALL, seed=seed)
    rng = np.random.default_rng(seed)
    w.batch.opt_flags = (rng.random(w.batch.n) < 0.15).astype(np.uint8)      # bit 0 = ELP_OPT_SR
    return w


def test_oracle_skips_sr_reads():
    w = _workload()
    with_sr = oracle_pipeline(w)
    w0 = _workload()
    w0.batch.opt_flags[:] = 0
    without = oracle_pipeline(w0)
    keep = w.batch.opt_flags == 0
    sub = synth.make_workload(5000, SMALL, seed=41)
    sub = synth.Workload(sub.header, sub.batch.take(np.nonzero(keep)[0]), sub.contig_bases, sub.sites, sub.params)
    only_untagged = oracle_pipeline(sub, markdup=False)
    # the tables of the tagged run are the tables of the run that never saw the tagged reads (duplicate marking off in both: it sees every read)
    w1 = _workload(); tagged_nomd = oracle_pipeline(w1, markdup=False)
    assert np.array_equal(oracle_tables_dense(tagged_nomd["tables"], 500)[0], oracle_tables_dense(only_untagged["tables"], 500)[0])
    assert not np.array_equal(oracle_tables_dense(with_sr["tables"], 500)[0], oracle_tables_dense(without["tables"], 500)[0])


@pytest.mark.gpu
def test_sr_reads_are_not_recalibrated():
    w = _workload()
    g = gpu_pipeline(w, n_batches=3, keep_ctx=True)
    o = oracle_pipeline(w)
    d, e = oracle_tables_dense(o["tables"], 500)
    assert np.array_equal(g["perm"], o["perm"]) and np.array_equal(g["flag"], o["flag"])
    assert np.array_equal(g["tables"], d), "sr reads must not reach the tables"
    assert np.array_equal(g["emp"], e) and np.array_equal(g["qual"], o["qual"])
    ctx = g["ctx"]
    assert np.array_equal(ctx.fetch_opt_flags(), w.batch.opt_flags[o["perm"].astype(np.int64)])      # what RemoveOptionalReads looks at, in output order
    ctx.close()
