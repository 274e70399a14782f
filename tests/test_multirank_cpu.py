"""world_size-2 (gloo, CPU) test of the multi-GPU partitioning logic: contig groups (sfm-style, sam/split-merge.go:178-213),
local sort + markdup + gather per rank, ONE all_reduce(sum) of the integer BQSR tables (the summation of
LoadAndCombineBQSRTables, filters/print-bqsr.go:310-329), and concatenation of the groups as the global order."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import bench
    import oracle
    from elprep_b200 import synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    contigs = [("c1", 300_000), ("c2", 250_000), ("c3", 120_000), ("c4", 80_000)]
    w = synth.make_workload(6000, contigs, seed=77, cross_contig_frac=0.0, threads=2)
    groups = bench.contig_groups(contigs, world)
    mine = {n for n, _ in groups[rank]}
    names = [n for n, _ in contigs]
    own = np.array([names[r] in mine if r >= 0 else (rank == world - 1) for r in w.batch.refid])     # unmapped reads go to the last rank
    sub = w.batch.take(np.nonzero(own)[0])
    oracle.mark_duplicates(sub, w.header, n_threads=1)
    perm = oracle.coordinate_sort(sub, n_threads=1)
    srt = sub.take(perm)
    ref = oracle.Reference(w.header, w.contig_bases, w.sites)
    t = oracle.bqsr_gather(srt, w.header, ref)
    tabs = [torch.from_numpy(a) for a in (t.q_obs, t.q_mis, t.c_obs, t.c_mis, t.x_obs, t.x_mis)]
    for x in tabs:
        dist.all_reduce(x)                                   # the single collective of the path
    np.save(os.path.join(out_dir, f"flags_{rank}.npy"), np.stack([np.nonzero(own)[0][perm], srt.flag.astype(np.int64)]))
    if rank == 0:
        np.savez(os.path.join(out_dir, "tables.npz"), *[x.numpy() for x in tabs])
    dist.barrier(); dist.destroy_process_group()


def test_contig_groups_balance():
    sys.path.insert(0, ROOT)
    import bench
    from elprep_b200 import synth
    for n in (1, 2, 4, 8):
        g = bench.contig_groups(synth.HG38, n)
        loads = [sum(l for _, l in x) for x in g]
        assert sorted(c for x in g for c, _ in x) == sorted(c for c, _ in synth.HG38)
        assert max(loads) <= 1.15 * (sum(loads) / n) or n == 8 and max(loads) <= 1.3 * (sum(loads) / n)


def test_two_rank_partition_equals_whole(tmp_path):
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    import oracle
    from elprep_b200 import synth
    port = 29600 + (os.getpid() % 300)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    contigs = [("c1", 300_000), ("c2", 250_000), ("c3", 120_000), ("c4", 80_000)]
    w = synth.make_workload(6000, contigs, seed=77, cross_contig_frac=0.0, threads=2)
    b = w.batch.copy()
    oracle.mark_duplicates(b, w.header, n_threads=1)
    perm = oracle.coordinate_sort(b, n_threads=1)
    srt = b.take(perm)
    t = oracle.bqsr_gather(srt, w.header, oracle.Reference(w.header, w.contig_bases, w.sites))
    z = np.load(os.path.join(tmp_path, "tables.npz"))
    for got, exp in zip([z[k] for k in z.files], (t.q_obs, t.q_mis, t.c_obs, t.c_mis, t.x_obs, t.x_mis)):
        assert np.array_equal(got, exp), "summed per-rank tables differ from the whole-genome tables"
    # duplicate flags: every read gets the same FLAG as in the whole run; per-rank orders are sub-sequences of the global order
    whole = dict(zip(perm.tolist(), srt.flag.tolist()))
    for r in range(2):
        idx, fl = np.load(os.path.join(tmp_path, f"flags_{r}.npy"))
        assert all(whole[int(i)] == int(f) for i, f in zip(idx, fl))
        pos_in_whole = {int(v): k for k, v in enumerate(perm)}
        ranks = [pos_in_whole[int(i)] for i in idx]
        assert ranks == sorted(ranks)


# ---- cross-group mate pairs: the reference's "spread" reads (sam/split-merge.go:286-293) as an exchange between the ranks ----
SPREAD_CONTIGS = [("c1", 300_000), ("c2", 250_000), ("c3", 120_000), ("c4", 80_000)]
SPREAD_KW = dict(seed=91, cross_contig_frac=0.3, dup_frac=0.4, optical_frac=0.3, want_reference=False, threads=2)


def _spread_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import pickle
    import torch.distributed as dist
    import oracle
    from elprep_b200 import synth, multi
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = synth.make_workload(5000, SPREAD_CONTIGS, **SPREAD_KW)
    groups = multi.contig_groups(SPREAD_CONTIGS, world)
    owner = multi.owner_table(w.header, groups)
    own = multi.partition(w.batch, owner, rank, world)
    sub = w.batch.take(own)

    def orc_markdup(batch, header):                    # same contract as multi.device_markdup, on the CPU restatement
        b = batch.copy()
        m = oracle.markdup_optical(b, header)
        return b.flag, m

    n_spread = multi.spread_reads(sub, owner, rank)[0].size
    sm = multi.exchange_spread_duplicates(sub, w.header, owner, rank, world, orc_markdup, multi.torch_gather_objects())
    mm = oracle.markdup_optical(sub, w.header)         # the rank's main pass (spread reads arrive with their 0x400 bit preset)
    with open(os.path.join(out_dir, f"spread_{rank}.pkl"), "wb") as f:
        pickle.dump(dict(own=own, flag=sub.flag.copy(), n_spread=n_spread, main=(mm.counters, mm.hist), spread=(sm.counters, sm.hist) if sm else None), f)
    dist.barrier(); dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_two_rank_spread_pairs(tmp_path, world):
    import pickle
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    import oracle
    from elprep_b200 import synth
    port = 29900 + (os.getpid() % 90) + 100 * world
    mp.spawn(_spread_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    w = synth.make_workload(5000, SPREAD_CONTIGS, **SPREAD_KW)
    whole = w.batch.copy()
    wm = oracle.markdup_optical(whole, w.header)
    res = [pickle.load(open(os.path.join(tmp_path, f"spread_{r}.pkl"), "rb")) for r in range(world)]
    assert sum(r["n_spread"] for r in res) > 200
    seen = np.zeros(whole.n, bool)
    spread_dups = 0
    for r in res:
        assert np.array_equal(r["flag"], whole.flag[r["own"]]), "FLAG of a partitioned run differs from the whole-file run"
        seen[r["own"]] = True
    assert seen.all()
    # duplication metrics: per-read counters from the main passes, pair-level numbers from main + spread passes
    n_slots = len(wm.counters)
    for slot in range(n_slots):
        tot = {k: 0 for k in oracle.COUNTERS}
        hist = [dict(), dict(), dict()]
        for r in res:
            for k in oracle.COUNTERS:
                tot[k] += r["main"][0][slot][k]
            parts = [r["main"][1][slot]] + ([r["spread"][1][slot]] if r["spread"] else [])
            if r["spread"]:
                for k in ("read_pair_duplicates", "read_pair_optical_duplicates"):
                    tot[k] += r["spread"][0][slot][k]
                spread_dups += r["spread"][0][slot]["read_pair_duplicates"]
            for p in parts:
                for wh in range(3):
                    for key, v in p[wh].items():
                        hist[wh][key] = hist[wh].get(key, 0) + v
        exp = wm.counters[slot]
        for k in oracle.COUNTERS:
            if k == "read_pairs_examined":          # halved per worker (:503-505): off by at most one per worker
                assert 0 <= exp[k] - tot[k] <= world
            else:
                assert tot[k] == exp[k], (slot, k)
        assert hist == wm.hist[slot], slot
    assert spread_dups > 10


def _split_worker(rank, world, port, out_dir):
    """the input side of `bench.py --gpus 2`: one genome, each rank generates the pairs that start in its contig group, then the split step"""
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import bench
    from elprep_b200 import multi, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    contigs = [("c1", 300_000), ("c2", 250_000), ("c3", 120_000), ("c4", 80_000)]
    groups = bench.contig_groups(contigs, world)
    header = synth.make_header(contigs)
    owner = multi.owner_table(header, groups)
    home = np.array([1 if owner[i] == rank else 0 for i in range(len(contigs))], np.uint8)
    w = synth.make_workload(5000, contigs, seed=900 + rank, home=home, pair_id_base=rank * 10**9, genome_seed=5, cross_contig_frac=0.05, threads=2, want_reference=False)
    a = multi.redistribute(w.batch, owner, rank, world, multi.torch_gather_objects())
    b = multi.redistribute(w.batch, owner, rank, world, multi.torch_gather_objects(), take=lambda x, idx: synth.take(x, idx, threads=2))
    same = all(np.array_equal(getattr(a, f), getattr(b, f)) for f in a.FIELDS)
    at_home = bool(np.all(np.where(a.refid >= 0, owner[np.maximum(a.refid, 0)], rank) == rank))
    cross = int(((a.refid >= 0) & (a.nref >= 0) & (owner[np.maximum(a.refid, 0)] != owner[np.maximum(a.nref, 0)])).sum())
    np.save(os.path.join(out_dir, f"split_{rank}.npy"), np.array([int(same), int(at_home), cross, a.n, w.batch.n]))
    dist.barrier(); dist.destroy_process_group()


def test_two_rank_split_of_one_genome(tmp_path):
    """after the split step every read sits on the rank that owns its contig, no read is lost, pairs still span the ranks, and the threaded
    gather bench.py uses gives the same batch as the numpy one"""
    import torch.multiprocessing as mp
    port = 29300 + (os.getpid() % 300)
    mp.spawn(_split_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r = [np.load(os.path.join(str(tmp_path), f"split_{k}.npy")) for k in range(2)]
    assert all(x[0] == 1 and x[1] == 1 and x[2] > 0 for x in r)
    assert r[0][3] + r[1][3] == r[0][4] + r[1][4]
