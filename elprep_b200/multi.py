"""Host-side plumbing for several GPUs: contig-group partitioning and the one exchange duplicate marking needs.

The reference scales out with ``elprep sfm``: the input is split per contig group (``computeContigGroups``,
``sam/split-merge.go:178-213``), reads whose mate maps to another group are copied into a separate *spread* file
(``SplitFilePerChromosome``, ``sam/split-merge.go:286-293``), every file is filtered on its own and the integer BQSR
tables are summed (``LoadAndCombineBQSRTables``, ``filters/print-bqsr.go:310-329``).  Here one process owns one GPU and one
contig group; the spread file becomes an exchange between the processes:

  1. a pair whose mates lie in different groups is classified by the rank that owns the mate with the smaller REFID
     (two pairs with equal duplicate signature have equal (refid1, refid2), so every candidate of a signature meets there);
  2. the other mate's record is sent to that rank (``torch.distributed``, object all-gather: the volume is the
     cross-group fraction of the reads);
  3. the owner marks duplicates among its spread pairs in a small separate context (the analogue of filtering the spread
     file) and returns the 0x400 bits of the visiting mates;
  4. every rank ORs the bits into the FLAG column of its own reads *before* they enter the main context -- the reference
     only ever sets the bit (``aln.FLAG |= sam.Duplicate``), so a preset bit survives duplicate marking unchanged.
In the main context a spread read is a paired read without its mate: it never joins a pair there, but it still shadows
the fragments at its position (``classifyFragment``, ``filters/mark-duplicates.go:225-252``), exactly as the tagged copy the
reference leaves in the group file does.  Pair-level duplication metrics of the spread pairs come from the small context,
per-read counters from the main one (``merge_spread_metrics``).

Everything here is numpy + torch.distributed; the duplicate marking itself is passed in as a callable (the device
context in production, any other implementation of the same contract in tests).
"""
from __future__ import annotations

import numpy as np

from . import sam

F_MULTIPLE, F_UNMAPPED, F_NEXTUNMAPPED, F_SECONDARY, F_DUPLICATE, F_SUPPLEMENTARY = 0x1, 0x4, 0x8, 0x100, 0x400, 0x800


def contig_groups(contigs, n):
    """sfm-style contig groups (sam/split-merge.go:178-213 balances by contig length): greedy longest-first bin packing.
    ``contigs``: [(name, length)] -> list of n lists of (name, length)."""
    groups = [[] for _ in range(n)]
    load = [0] * n
    for name, ln in sorted(contigs, key=lambda x: -x[1]):
        k = int(np.argmin(load))
        groups[k].append((name, ln))
        load[k] += ln
    return groups


def owner_table(header, groups):
    """refid -> rank owning the contig (int32 array over @SQ)"""
    names = header.contig_names()
    where = {name: r for r, g in enumerate(groups) for name, _ in g}
    return np.array([where[n] for n in names], dtype=np.int32)


def partition(batch, owner, rank, world):
    """indices of the reads this rank owns: by contig group; unmapped reads (REFID -1) go to the last rank, where they
    sort last (sam/sam-types.go:428-432)"""
    refid = batch.refid
    own = np.where(refid >= 0, owner[np.maximum(refid, 0)], world - 1)
    return np.nonzero(own == rank)[0]


def spread_reads(batch, owner, rank):
    """-> (idx, pair_owner): local reads that are one mate of a cross-group true pair, and the rank that classifies the pair.
    True pair: FLAG & (0x1|0x8) == 0x1 (isTruePair, mark-duplicates.go:182-184); only reads entering duplicate marking
    (not unmapped / secondary / supplementary, :436)."""
    f = batch.flag
    cand = ((f & (F_MULTIPLE | F_NEXTUNMAPPED)) == F_MULTIPLE) & ((f & (F_UNMAPPED | F_SECONDARY | F_SUPPLEMENTARY)) == 0) & (batch.refid >= 0) & (batch.nref >= 0)
    idx = np.nonzero(cand)[0]
    if idx.size == 0:
        return idx, np.zeros(0, np.int32)
    o1, o2 = owner[batch.refid[idx]], owner[batch.nref[idx]]
    cross = o1 != o2
    idx = idx[cross]
    pair_owner = owner[np.minimum(batch.refid[idx], batch.nref[idx])]
    return idx, pair_owner.astype(np.int32)


def _pack(batch):
    return {f: getattr(batch, f) for f in sam.AlignmentBatch.FIELDS}


def _unpack(d):
    return sam.AlignmentBatch(**d)


def exchange_spread_duplicates(batch, header, owner, rank, world, markdup, gather_objects):
    """Sets the 0x400 bits of cross-group pairs in ``batch.flag`` (this rank's reads, arrival order) in place.

    markdup(batch, header) -> (flags uint16[n], metrics or None): MarkDuplicates over a stand-alone batch.
    gather_objects(obj) -> list of every rank's obj (torch.distributed.all_gather_object in production).
    Returns the metrics object ``markdup`` produced for the spread pairs this rank classified (or None)."""
    idx, powner = spread_reads(batch, owner, rank)
    mine = idx[powner == rank]
    # 1. ship the visiting mates to the rank that classifies their pair
    out = {}
    for dst in range(world):
        if dst == rank:
            continue
        sel = idx[powner == dst]
        if sel.size:
            out[dst] = (_pack(batch.take(sel)), sel)
    inbox = gather_objects({dst: payload[0] for dst, payload in out.items()})
    parts, origin = [batch.take(mine)], [(rank, mine.size)]
    for src in range(world):
        if src != rank and rank in inbox[src]:
            b = _unpack(inbox[src][rank])
            parts.append(b)
            origin.append((src, b.n))
    spread = sam.AlignmentBatch.concat(parts) if len(parts) > 1 else parts[0]
    # 2. duplicate marking among the spread pairs (the analogue of filtering the reference's spread file)
    metrics = None
    if spread.n:
        flags, metrics = markdup(spread, header)
    else:
        flags = np.zeros(0, np.uint16)
    dup = (flags & F_DUPLICATE) != 0
    # 3. return the bits of the visitors, take ours
    back, pos = {}, 0
    for src, cnt in origin:
        if src == rank:
            batch.flag[mine[dup[pos:pos + cnt]]] |= F_DUPLICATE
        else:
            back[src] = dup[pos:pos + cnt]
        pos += cnt
    replies = gather_objects(back)
    for dst, (_, sel) in out.items():
        bits = replies[dst][rank]
        batch.flag[sel[bits]] |= F_DUPLICATE
    return metrics


def torch_gather_objects():
    """all_gather_object over the default process group"""
    import torch.distributed as dist

    def gather(obj):
        res = [None] * dist.get_world_size()
        dist.all_gather_object(res, obj)
        return res
    return gather


def device_markdup(device_ordinal=0, optical=False, optical_pixel_distance=100):
    """MarkDuplicates of a stand-alone batch on the GPU (a small second context beside the rank's main one)."""
    from . import device, _lib

    def run(batch, header):
        ctx = device.Context(header, device=device_ordinal, optical_pixel_distance=optical_pixel_distance)
        try:
            ctx.append(batch)
            ctx.sort_markdup(device.SO_KEEP, _lib.MARKDUP_OPTICAL if optical else _lib.MARKDUP)
            idx, flag, _, _ = ctx.fetch(want_qual=False)
            out = np.empty(batch.n, np.uint16)
            out[idx.astype(np.int64)] = flag
            return out, (ctx.optical_metrics() if optical else None)
        finally:
            ctx.close()
    return run


PAIR_LEVEL = ("read_pair_duplicates", "read_pair_optical_duplicates")


def merge_spread_metrics(ctx, spread_metrics):
    """adds the pair-level numbers of the spread pairs (ReadPairDuplicates, ReadPairOpticalDuplicates, the three
    histograms) to the main context's metrics; the per-read counters of those reads are already there"""
    from . import _lib
    for slot, m in enumerate(spread_metrics or []):
        counters = [m[k] if k in PAIR_LEVEL else 0 for k in _lib.ElpDupMetrics.COUNTERS]   # (per-read counters stay 0)
        ctx.optical_merge(slot, counters, m["hist"])


def redistribute(batch, owner, rank, world, gather_objects, take=None):
    """The `split` step of an sfm run for one rank's share of the input (sam/split-merge.go:178-311): reads whose contig belongs to another
    rank are handed to that rank, reads arriving from the others are appended.  Unmapped reads (REFID -1) stay where they are.  Host-side
    setup (numpy + one object all-gather), not part of any timed region: the device library starts from reads that are already home."""
    take = take or (lambda b, idx: b.take(idx))      # (bench.py passes synth.take: host threads instead of element-wise numpy index arrays)
    refid = batch.refid
    dest = np.where(refid >= 0, owner[np.maximum(refid, 0)], rank)
    out = {}
    for dst in range(world):
        if dst != rank:
            sel = np.nonzero(dest == dst)[0]
            if sel.size:
                out[dst] = _pack(take(batch, sel))
    inbox = gather_objects(out)
    parts = [take(batch, np.nonzero(dest == rank)[0])]
    for src in range(world):
        if src != rank and rank in inbox[src]:
            parts.append(_unpack(inbox[src][rank]))
    return sam.AlignmentBatch.concat(parts) if len(parts) > 1 else parts[0]
