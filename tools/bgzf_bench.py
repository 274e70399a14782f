"""Host BGZF throughput (no GPU): BAM-like bytes (synthetic alignment records) through elp_bgzf_deflate / elp_bgzf_inflate on n host threads.
usage: python tools/bgzf_bench.py [reads] [threads]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from elprep_b200 import bgzf, synth  # noqa: E402
from util import encode_bam  # noqa: E402


def main():
    reads = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
    threads = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
    w = synth.make_workload(reads // 2, [("chr20", 3_000_000)], seed=3, want_reference=False)
    rec, _ = encode_bam(w.batch, w.header)
    raw = np.ascontiguousarray(rec, dtype=np.uint8)
    res = {"bytes": int(raw.size), "threads": threads}
    for level in (1, 6):
        t0 = time.perf_counter(); comp = bgzf.deflate(raw, level=level, n_threads=threads); t1 = time.perf_counter()
        back = bgzf.inflate(comp, n_threads=threads); t2 = time.perf_counter()
        assert np.array_equal(back, raw)
        res[f"level{level}"] = {"ratio": round(raw.size / comp.size, 2), "deflate_MBps": round(raw.size / (t1 - t0) / 1e6), "inflate_MBps": round(raw.size / (t2 - t1) / 1e6)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
