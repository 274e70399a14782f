"""PCIe probe through the library's own copy path: upload (elp_append_batch_async) and download (elp_fetch_async) alone, together, and with the
device phases of a third context running beside them.  Prints the milliseconds of each.  ELPREP_B200_COPY_CHUNK_MB selects the copy granularity."""
import json, os, sys, time
import numpy as np
import torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from elprep_b200 import device, synth  # noqa: E402
import bench  # noqa: E402


def main():
    reads = int(os.environ.get("PROBE_READS", 30_000_000))
    contigs = synth.scaled_hg38(20.0)
    w = synth.make_workload(reads // 2, contigs, seed=20260924, threads=32)
    n = w.batch.n
    hb = bench.pinned(w.batch)

    def mk():
        c = device.Context(w.header, profile=False)
        for ci in range(len(contigs)):
            c.set_reference(ci, w.contig_bases[ci]); c.set_known_sites(ci, w.sites[ci], already_flat=True)
        return c

    def phases(c):
        c.sort_markdup(device.SO_COORDINATE, True); c.bqsr_gather(); c.bqsr_finalize(None); c.bqsr_apply()
    A, B, Cx = mk(), mk(), mk()
    for c in (A, B, Cx):
        c.append(hb); phases(c)
    o = tuple(torch.empty(s, dtype=dt, pin_memory=True) for s, dt in ((n, torch.int32), (n, torch.int16), (n + 1, torch.int64), (int(hb.qual.size), torch.uint8)))
    out = (o[0].numpy().view(np.uint32), o[1].numpy().view(np.uint16), o[2].numpy().view(np.uint64), o[3].numpy())
    res = {"reads": n, "d2h_gb": sum(a.nbytes for a in out) / 1e9, "chunk_mb": os.environ.get("ELPREP_B200_COPY_CHUNK_MB", "32")}

    def t(fn):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); return round(1e3 * (time.perf_counter() - t0), 1)
    for rep in range(2):
        res[f"h2d_alone_{rep}"] = t(lambda: (A.reset(), A.append_async(hb), A.append_wait()))
        res[f"d2h_alone_{rep}"] = t(lambda: (B.fetch_async(out), B.fetch_wait()))

        def both(with_phases):
            A.reset(); A.append_async(hb); B.fetch_async(out)
            t0 = time.perf_counter()
            if with_phases:
                phases(Cx)
            t1 = time.perf_counter(); B.fetch_wait(); t2 = time.perf_counter(); A.append_wait(); t3 = time.perf_counter()
            return [round(1e3 * (x - t0), 1) for x in (t1, t2, t3)]
        res[f"both_{rep}"] = both(False)
        Cx.reset(); Cx.append(hb); torch.cuda.synchronize()
        res[f"both_with_phases_{rep}"] = both(True)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
