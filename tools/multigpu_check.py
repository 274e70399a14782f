"""C4 check: the same genome partitioned over N GPUs must give the results of one GPU.
Run under torchrun (one rank per GPU):  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py
Every rank builds the WHOLE input (all groups' generators, small sizes), runs it through one context on its own GPU (the single-GPU answer),
then runs only its contig group's reads through the collective path (NCCL behind the C ABI: spread-pair exchange, table allreduce, metrics
allreduce) and compares: FLAG and QUAL of its reads, the summed BQSR tables, the summed duplication metrics."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from elprep_b200 import device, multi, sam, synth, _lib  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    pairs = int(os.environ.get("CHECK_PAIRS", 40000))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    contigs = synth.scaled_hg38(400.0)
    groups = multi.contig_groups(contigs, world)
    header = synth.make_header(contigs)
    owner = multi.owner_table(header, groups)
    names = [n for n, _ in contigs]
    parts, ref_w = [], None
    for r in range(world):
        home = np.array([1 if owner[i] == r else 0 for i in range(len(contigs))], np.uint8)
        w = synth.make_workload(pairs, contigs, seed=500 + r, home=home, pair_id_base=r * 10**9, genome_seed=77, cross_contig_frac=0.05, optical_frac=0.4, dup_frac=0.2,
                                want_reference=(r == 0))
        if r == 0:
            ref_w = w
        parts.append(w.batch)
    whole = sam.AlignmentBatch.concat(parts)

    def setup(ctx):
        for ci in range(len(contigs)):
            ctx.set_reference(ci, ref_w.contig_bases[ci]); ctx.set_known_sites(ci, ref_w.sites[ci])
    # ---- one GPU, whole genome ----
    a = device.Context(header, device=local)
    setup(a)
    a.append(whole); a.sort_markdup(device.SO_COORDINATE, _lib.MARKDUP_OPTICAL)
    a.bqsr_gather(); t_whole = a.tables_get(); a.bqsr_finalize(None); a.bqsr_apply()
    idx, flag, qoff, qual = a.fetch()
    m_whole = a.optical_metrics()
    flag_by_read = np.empty(whole.n, np.uint16); flag_by_read[idx.astype(np.int64)] = flag
    qual_by_read = {}
    a.close()
    # ---- N GPUs ----
    mine = np.nonzero(np.where(whole.refid >= 0, owner[np.maximum(whole.refid, 0)], world - 1) == rank)[0]
    b = device.Context(header, device=local)
    setup(b)
    uid = [device.Context.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    b.comm_init(uid[0], rank, world); b.comm_set_partition(owner)
    b.append(whole.take(mine))
    b.sort_markdup(device.SO_COORDINATE, _lib.MARKDUP_OPTICAL)
    b.bqsr_gather(); b.tables_allreduce(); t_multi = b.tables_get()
    b.bqsr_finalize(None); b.bqsr_apply()
    b.optical_allreduce(); m_multi = b.optical_metrics()
    idx2, flag2, qoff2, qual2 = b.fetch()
    # compare per read: output record k of the rank is read mine[idx2[k]] of the whole input
    gidx = mine[idx2.astype(np.int64)]
    res = {"rank": rank, "reads": int(mine.size), "cross_pairs_reads": int(((owner[np.maximum(whole.refid[mine], 0)] != owner[np.maximum(whole.nref[mine], 0)]) & (whole.refid[mine] >= 0) & (whole.nref[mine] >= 0)).sum()),
           "flags_equal": bool(np.array_equal(flag2, flag_by_read[gidx])), "tables_equal": bool(np.array_equal(t_whole, t_multi)),
           "duplicates": int(((flag2 & 0x400) != 0).sum())}
    # QUAL: order of the whole run restricted to this rank's reads = this rank's order (concatenation of contig groups), so compare per read through offsets
    pos_in_whole = np.empty(whole.n, np.int64); pos_in_whole[idx.astype(np.int64)] = np.arange(whole.n)
    ok_q = True
    for k in np.linspace(0, mine.size - 1, num=min(3000, mine.size), dtype=np.int64):
        w_k = pos_in_whole[gidx[k]]
        if not np.array_equal(qual2[int(qoff2[k]):int(qoff2[k + 1])], qual[int(qoff[w_k]):int(qoff[w_k + 1])]):
            ok_q = False; break
    res["qual_equal_sampled"] = ok_q
    keys = ("unpaired_reads_examined", "read_pairs_examined", "secondary_or_supplementary_reads", "unmapped_reads", "unpaired_read_duplicates", "read_pair_duplicates", "read_pair_optical_duplicates", "estimated_library_size")
    res["metrics_equal"] = all(x[k] == y[k] for x, y in zip(m_whole, m_multi) for k in keys) and all(x["hist"] == y["hist"] for x, y in zip(m_whole, m_multi))
    b.close()
    allres = [None] * world
    dist.all_gather_object(allres, res)
    if rank == 0:
        ok = all(r["flags_equal"] and r["tables_equal"] and r["qual_equal_sampled"] and r["metrics_equal"] for r in allres)
        print(json.dumps({"world": world, "ok": ok, "ranks": allres}))
        sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
