"""The oracle's outputs on small seeded workloads against committed digests (tests/golden/oracle_regression.json, written by
tools/make_golden.py).  Regression fixtures of the restatement -- not reference outputs (the reference cannot be built here)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_oracle_outputs_match_committed_digests():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_golden
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_regression.json")))
    assert set(gold) == set(make_golden.CASES)
    for name, case in make_golden.CASES.items():
        got = make_golden.run_case(case)
        assert got == gold[name], name
