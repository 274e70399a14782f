"""Debug aid: run the apply phase with both kernels on the same context and report where their QUAL output differs."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from elprep_b200 import device, synth


def run(n=6000, seed=77, reps=3):
    w = synth.make_workload(n, [("chr20", 300_000), ("chr21", 200_000)], seed=seed)
    ctx = device.Context(w.header)
    for ci in range(2):
        ctx.set_reference(ci, w.contig_bases[ci]); ctx.set_known_sites(ci, w.sites[ci])
    ctx.append(w.batch); ctx.sort_markdup(device.SO_COORDINATE, False); ctx.bqsr_gather(); ctx.bqsr_finalize(None)
    outs = {}
    for mode in ["v1"] + ["v2"] * reps:
        os.environ["ELPREP_B200_APPLY"] = mode
        ctx.bqsr_apply()
        idx, flag, qoff, qual = ctx.fetch()
        outs.setdefault(mode, []).append(qual[:int(qoff[-1])].copy())
    ref = outs["v1"][0]
    for k, o in enumerate(outs["v2"]):
        bad = np.nonzero(o != ref)[0]
        print(f"n={n} seed={seed} v2 run {k}: {bad.size} of {ref.size} bytes differ")
        if bad.size:
            rd = np.searchsorted(qoff, bad, side="right") - 1
            qin = w.batch.qual
            for b_, r_ in list(zip(bad, rd))[:12]:
                src = int(idx[r_]); L = int(w.batch.lseq[src]); j = int(b_ - qoff[r_])
                qi = int(qin[int(w.batch.qual_off[src]) + j])
                print(f"   sorted read {r_} (input {src}) L={L} flag={int(flag[r_]):#x} base {j}: in {qi} v1 {int(ref[b_])} v2 {int(o[b_])}")
            print("   reads hit:", np.unique(rd).size, " positions-in-read histogram (first 10):", np.bincount((bad - qoff[rd]).astype(np.int64))[:10])
    ctx.close()


if __name__ == "__main__":
    run()
    run(n=40000, seed=5)
