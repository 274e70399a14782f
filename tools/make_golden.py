"""Regenerates tests/golden/oracle_regression.json: digests of the ORACLE's outputs on small seeded workloads.
These are regression fixtures of the restatement itself (they do not come from the reference, which cannot be built here --
see oracle/oracle.h "PARITY UNPINNED"); they catch silent drift of oracle/, of the synthetic generator and of the text formats.
usage: python tools/make_golden.py"""
import hashlib, json, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle
from elprep_b200 import synth
from util import oracle_pipeline

CASES = {
    "wgs_small": dict(n_pairs=3000, contigs=[("chr20", 300_000), ("chr21", 150_000)], kw=dict(seed=11)),
    "dups_optical": dict(n_pairs=2500, contigs=[("chr20", 200_000)], kw=dict(seed=12, dup_frac=0.5, optical_frac=0.5, n_rg=1)),
    "short_reads": dict(n_pairs=2000, contigs=[("chr20", 200_000), ("chrM", 16_569)], kw=dict(seed=13, L=50, unmapped_frac=0.1)),
}


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:24]


def run_case(c):
    w = synth.make_workload(c["n_pairs"], c["contigs"], **c["kw"])
    o = oracle_pipeline(w)
    b = w.batch.copy()
    with tempfile.TemporaryDirectory() as d:
        m = oracle.markdup_optical(b, w.header, metrics_path=os.path.join(d, "m.txt"), command_line="elprep filter in out", started_on="T")
        metrics_txt = open(os.path.join(d, "m.txt")).read()
    return dict(n=int(w.batch.n), input=digest(np.concatenate([w.batch.pos.view(np.uint8), w.batch.qual, w.batch.seq, w.batch.qname])),
                perm=digest(o["perm"]), flag=digest(o["flag"]), qual=digest(o["qual"]), report=hashlib.sha256(o["report"].encode()).hexdigest()[:24],
                duplicates=int(((o["flag"] & 0x400) != 0).sum()), metrics=hashlib.sha256(metrics_txt.encode()).hexdigest()[:24],
                counters=[list(c.values()) for c in m.counters])


if __name__ == "__main__":
    out = {k: run_case(c) for k, c in CASES.items()}
    p = os.path.join(ROOT, "tests", "golden", "oracle_regression.json")
    json.dump(out, open(p, "w"), indent=1, sort_keys=True)
    print("wrote", p)
