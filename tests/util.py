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


# ---- BAM alignment records (sam/bam-files.go:300-400) for the device ingest tests ----
def encode_bam(batch, header, rng=None, with_aux=True, extra_tags=None):
    """AlignmentBatch -> (uint8 record bytes, uint64 record offsets [n+1]).  Each record carries its block_size, the fixed
    fields of parseBamAlignment, NUL-terminated name, CIGAR words, SEQ nibbles, QUAL bytes and typed optional fields
    (RG:Z plus a mix of the other value types, so that the tag walk is exercised)."""
    import struct
    rng = rng or np.random.default_rng(0)
    ids = [r["ID"] for r in header.RG]
    out, offs = bytearray(), [0]
    qo, co = batch.qname_off.astype(np.int64), batch.cigar_off.astype(np.int64)
    so, uo = batch.seq_off.astype(np.int64), batch.qual_off.astype(np.int64)
    for i in range(batch.n):
        name = bytes(batch.qname[qo[i]:qo[i + 1]]) + b"\0"
        cig = batch.cigar[co[i]:co[i + 1]].astype("<u4").tobytes()
        L = int(batch.lseq[i])
        seq = bytes(batch.seq[so[i]:so[i] + (L + 1) // 2]); qual = bytes(batch.qual[uo[i]:uo[i] + L])
        aux = b""
        if with_aux:
            k = int(rng.integers(0, 4))
            if k >= 1: aux += b"NMC" + struct.pack("<B", int(rng.integers(0, 9)))
            if k >= 2: aux += b"MDZ" + b"75A74\0"
            if int(batch.rg[i]) >= 0: aux += b"RGZ" + ids[int(batch.rg[i])].encode() + b"\0"
            if k >= 3: aux += b"ASi" + struct.pack("<i", -5) + b"XSf" + struct.pack("<f", 1.5) + b"ZBBs" + struct.pack("<I", 3) + struct.pack("<3h", 1, -2, 3) + b"XAA" + b"q"
        elif int(batch.rg[i]) >= 0:
            aux += b"RGZ" + ids[int(batch.rg[i])].encode() + b"\0"
        if extra_tags is not None:
            aux += extra_tags[i]
        body = struct.pack("<iiBBHHHiiii", int(batch.refid[i]), int(batch.pos[i]) - 1, len(name), int(batch.mapq[i]), 4680, (co[i + 1] - co[i]) & 0xffff,
                           int(batch.flag[i]), L, int(batch.nref[i]), int(batch.pnext[i]) - 1, int(batch.tlen[i])) + name + cig + seq + qual + aux
        out += struct.pack("<I", len(body)) + body
        offs.append(len(out))
    return np.frombuffer(bytes(out), dtype=np.uint8).copy(), np.array(offs, dtype=np.uint64)


def decode_bam(raw, offs, header):
    """restatement of parseBamAlignment (sam/bam-files.go:314-400) for the fields of the path -> AlignmentBatch"""
    import struct
    from elprep_b200 import sam
    ids = {r["ID"]: k for k, r in enumerate(header.RG)}
    cols = {k: [] for k in ("refid", "pos", "flag", "mapq", "nref", "pnext", "tlen", "rg", "lseq")}
    qn, cg, sq, ql, qoff, coff = bytearray(), [], bytearray(), bytearray(), [0], [0]
    b = raw.tobytes()
    sizes = {"A": 1, "c": 1, "C": 1, "s": 2, "S": 2, "i": 4, "I": 4, "f": 4}
    for i in range(len(offs) - 1):
        r = b[int(offs[i]):int(offs[i + 1])]
        bs, refid, pos, lname, mapq, _bin, ncig, flag, lseq, nref, pnext, tlen = struct.unpack_from("<IiiBBHHHiiii", r, 0)
        assert bs + 4 == len(r)
        x = 36
        qn += r[x:x + lname - 1]; qoff.append(len(qn)); x += lname
        cg += list(struct.unpack_from("<%dI" % ncig, r, x)); coff.append(len(cg)); x += 4 * ncig
        sq += r[x:x + (lseq + 1) // 2]; x += (lseq + 1) // 2
        ql += r[x:x + lseq]; x += lseq
        rg = -1
        while x < len(r):
            tag, ty = r[x:x + 2], chr(r[x + 2]); x += 3
            if ty in sizes: x += sizes[ty]
            elif ty in "ZH":
                e = r.index(b"\0", x)
                if tag == b"RG" and ty == "Z": rg = ids[r[x:e].decode()]
                x = e + 1
            elif ty == "B":
                sub, cnt = chr(r[x]), struct.unpack_from("<I", r, x + 1)[0]; x += 5 + cnt * sizes[sub]
            else: raise ValueError("bad type")
        for k, v in zip(cols, (refid, pos + 1, flag, mapq, nref, pnext + 1, tlen, rg, lseq)):
            cols[k].append(v)
    return sam.AlignmentBatch(qname_off=np.array(qoff, np.uint64), qname=np.frombuffer(bytes(qn), np.uint8), cigar_off=np.array(coff, np.uint64),
                              cigar=np.array(cg, np.uint32), seq=np.frombuffer(bytes(sq), np.uint8), qual=np.frombuffer(bytes(ql), np.uint8).copy(),
                              **{k: np.array(v) for k, v in cols.items()})
