"""Shared helpers for the parity tests: run the same workload through the oracle (CPU restatement of the
reference) and through the CUDA library (C ABI via ctypes), and return comparable results."""
import os
import tempfile

import numpy as np


def oracle_pipeline(w, bqsr=True, threads=4, max_cycle=500, quantize_levels=0, sqq=None, sort=True, markdup=True):
    import oracle
    b = w.batch.copy()
    if markdup:
        oracle.mark_duplicates(b, w.header, n_threads=1)
    perm = oracle.coordinate_sort(b, n_threads=threads) if sort else np.arange(b.n, dtype=np.int64)
    srt = b.take(perm)
    res = dict(perm=perm.astype(np.uint64), flag=srt.flag.copy(), qual=srt.qual.copy(), qual_off=srt.qual_off.copy())
    if bqsr:
        ref = oracle.Reference(w.header, w.contig_bases, w.sites)
        t = oracle.bqsr_gather(srt, w.header, ref, max_cycle=max_cycle, n_threads=threads)
        oracle.bqsr_finalize(t)
        with tempfile.TemporaryDirectory() as d:
            p = os.path.join(d, "o.recal")
            oracle.bqsr_report(t, oracle.OracleHeader(w.header).cov_names, p)
            res["report"] = open(p).read()
        oracle.bqsr_apply(srt, w.header, t, quantize_levels=quantize_levels, sqq=sqq, n_threads=threads)
        res["qual"] = srt.qual.copy()
        res["tables"] = t
    return res


def oracle_tables_dense(t, max_cycle=500):
    """oracle tables -> the C ABI's dense layout [n_cov][94][1+(2mc+1)+16][2] (+ empirical [..][..][..])"""
    n_cov = t.n_cov
    ncol = 1 + (2 * max_cycle + 1) + 16
    d = np.zeros((n_cov, 94, ncol, 2), dtype=np.int64)
    e = np.zeros((n_cov, 94, ncol), dtype=np.uint8)
    d[:, :, 0, 0] = t.q_obs[:, :94]; d[:, :, 0, 1] = t.q_mis[:, :94]; e[:, :, 0] = t.q_emp[:, :94]
    d[:, :, 1:1 + 2 * max_cycle + 1, 0] = t.c_obs[:, :94]; d[:, :, 1:1 + 2 * max_cycle + 1, 1] = t.c_mis[:, :94]; e[:, :, 1:1 + 2 * max_cycle + 1] = t.c_emp[:, :94]
    d[:, :, 1 + 2 * max_cycle + 1:, 0] = t.x_obs[:, :94]; d[:, :, 1 + 2 * max_cycle + 1:, 1] = t.x_mis[:, :94]; e[:, :, 1 + 2 * max_cycle + 1:] = t.x_emp[:, :94]
    assert t.q_obs[:, 94:].sum() == 0 and t.c_obs[:, 94:].sum() == 0
    return d, e


def gpu_pipeline(w, bqsr=True, n_batches=1, max_cycle=500, quantize_levels=0, sqq=None, sort=True, markdup=True, profile=False, keep_ctx=False):
    from elprep_b200 import device
    ctx = device.Context(w.header, max_cycle=max_cycle, quantize_levels=quantize_levels, sqq=sqq, profile=profile)
    try:
        if bqsr:
            for ci in range(len(w.header.SQ)):
                ctx.set_reference(ci, w.contig_bases[ci])
                ctx.set_known_sites(ci, w.sites[ci], already_flat=True)
        n = w.batch.n
        if n_batches <= 1 or n < n_batches:
            ctx.append(w.batch)
        else:
            bounds = [n * i // n_batches for i in range(n_batches + 1)]
            for a, b in zip(bounds[:-1], bounds[1:]):
                ctx.append(w.batch.take(np.arange(a, b)))
        ctx.sort_markdup(device.SO_COORDINATE if sort else device.SO_KEEP, markdup)
        res = {}
        if bqsr:
            ctx.bqsr_gather()
            res["tables"] = ctx.tables_get()
            with tempfile.TemporaryDirectory() as d:
                p = os.path.join(d, "g.recal")
                ctx.bqsr_finalize(p)
                res["report"] = open(p).read()
            res["emp"] = ctx.empirical_get()
            ctx.bqsr_apply()
        idx, flag, qoff, qual = ctx.fetch()
        res.update(perm=idx, flag=flag, qual=qual[:int(qoff[-1])] if n else qual[:0], qual_off=qoff)
        if profile:
            res["stats"] = ctx.kernel_stats()
        res["launches"] = ctx.launch_count()
        if keep_ctx:
            res["ctx"] = ctx
        return res
    finally:
        if not keep_ctx:
            ctx.close()
