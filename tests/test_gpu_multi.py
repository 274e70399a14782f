"""C4 under pytest: one genome partitioned over two GPUs through the NCCL path inside the library must give the
single-GPU results (tools/multigpu_check.py does the work, one rank per GPU).  Skipped on boxes with one GPU."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpus_equal_one_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, CHECK_PAIRS="20000", NCCL_DEBUG_FILE="/dev/stderr")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29531",
                        os.path.join(ROOT, "tools", "multigpu_check.py")], capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("{")][-1]
    res = json.loads(line)
    assert res["ok"] and res["world"] == 2
    for r in res["ranks"]:
        assert r["flags_equal"] and r["tables_equal"] and r["qual_equal_sampled"] and r["metrics_equal"] and r["cross_pairs_reads"] > 0
