"""ctypes wrapper around oracle/_build/liboracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl
reference`` leg may import this package.  The product (elprep_b200) never does.
PARITY UNPINNED: see oracle/oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("oracle.c", "gomath.c", "oracle.h", "gomath.h", "Makefile")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs if os.path.exists(s)):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


class _Reads(C.Structure):
    _fields_ = [("n", C.c_int64), ("refid", C.c_void_p), ("pos", C.c_void_p), ("flag", C.c_void_p), ("mapq", C.c_void_p),
                ("nref", C.c_void_p), ("pnext", C.c_void_p), ("tlen", C.c_void_p), ("rg", C.c_void_p),
                ("qname_off", C.c_void_p), ("qname", C.c_void_p), ("cigar_off", C.c_void_p), ("cigar", C.c_void_p),
                ("lseq", C.c_void_p), ("seq_off", C.c_void_p), ("seq", C.c_void_p), ("qual_off", C.c_void_p), ("qual", C.c_void_p), ("opt_flags", C.c_void_p)]


class _Header(C.Structure):
    _fields_ = [("n_contigs", C.c_int32), ("contig_len", C.c_void_p), ("n_rg", C.c_int32), ("rg_lib", C.c_void_p),
                ("rg_cov", C.c_void_p), ("n_cov", C.c_int32)]


class _Tables(C.Structure):
    _fields_ = [("n_cov", C.c_int32), ("max_cycle", C.c_int32)] + [(k, C.c_void_p) for k in
                ("q_obs", "q_mis", "c_obs", "c_mis", "x_obs", "x_mis", "q_emp", "c_emp", "x_emp")]


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_phred_score.restype = C.c_int32
        L.orc_unclipped_position.restype = C.c_int32
        L.orc_mod_flag.restype = C.c_uint16
        L.orc_mod_flag.argtypes = [C.c_uint16]
        L.orc_flatten.restype = C.c_int64
        L.orc_empirical_quality.restype = C.c_uint8
        L.orc_empirical_quality.argtypes = [C.c_int64, C.c_int64, C.c_double]
        L.orc_prior_cache.restype = C.c_double
        for f in ("gm_log", "gm_log2", "gm_log10", "gm_exp", "gm_lgamma", "gm_round"):
            getattr(L, f).restype = C.c_double
            getattr(L, f).argtypes = [C.c_double]
        L.gm_pow.restype = C.c_double
        L.gm_pow.argtypes = [C.c_double, C.c_double]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class OracleReads:
    """Keeps numpy arrays alive and exposes an ``orc_reads`` struct."""

    def __init__(self, batch):
        self.b = batch
        self._keep = [batch.qual_off, batch.seq_off]
        s = _Reads()
        s.n = batch.n
        for k in ("refid", "pos", "flag", "mapq", "nref", "pnext", "tlen", "rg", "qname_off", "qname", "cigar_off", "cigar",
                  "lseq", "seq", "qual"):
            setattr(s, k, _p(getattr(batch, k)))
        s.seq_off = _p(batch.seq_off)
        s.qual_off = _p(batch.qual_off)
        s.opt_flags = _p(batch.opt_flags)
        self.s = s


class OracleHeader:
    def __init__(self, header):
        self.contig_len = header.contig_lengths()
        self.rg_lib, self.lib_names = header.rg_lib_ids()
        self.rg_cov, self.cov_names = header.rg_cov_ids()
        s = _Header()
        s.n_contigs = len(self.contig_len)
        s.contig_len = _p(self.contig_len)
        s.n_rg = len(self.rg_lib)
        s.rg_lib = _p(self.rg_lib)
        s.rg_cov = _p(self.rg_cov)
        s.n_cov = len(self.cov_names)
        self.s = s


class OracleTables:
    NQ, NCTX = 256, 16

    def __init__(self, n_cov, max_cycle=500):
        self.n_cov, self.max_cycle = n_cov, max_cycle
        nc = 2 * max_cycle + 1
        self.q_obs = np.zeros((n_cov, 256), dtype=np.int64)
        self.q_mis = np.zeros((n_cov, 256), dtype=np.int64)
        self.c_obs = np.zeros((n_cov, 256, nc), dtype=np.int64)
        self.c_mis = np.zeros((n_cov, 256, nc), dtype=np.int64)
        self.x_obs = np.zeros((n_cov, 256, 16), dtype=np.int64)
        self.x_mis = np.zeros((n_cov, 256, 16), dtype=np.int64)
        self.q_emp = np.zeros((n_cov, 256), dtype=np.uint8)
        self.c_emp = np.zeros((n_cov, 256, nc), dtype=np.uint8)
        self.x_emp = np.zeros((n_cov, 256, 16), dtype=np.uint8)
        s = _Tables()
        s.n_cov, s.max_cycle = n_cov, max_cycle
        for k in ("q_obs", "q_mis", "c_obs", "c_mis", "x_obs", "x_mis", "q_emp", "c_emp", "x_emp"):
            setattr(s, k, _p(getattr(self, k)))
        self.s = s


def coordinate_sort(batch, n_threads=1):
    r = OracleReads(batch)
    perm = np.zeros(batch.n, dtype=np.int64)
    lib().orc_coordinate_sort(C.byref(r.s), _p(perm), C.c_int(n_threads))
    return perm


def queryname_sort(batch):
    r = OracleReads(batch)
    perm = np.zeros(batch.n, dtype=np.int64)
    lib().orc_queryname_sort(C.byref(r.s), _p(perm))
    return perm


def coordinate_less(batch, a, b):
    r = OracleReads(batch)
    return bool(lib().orc_coordinate_less(C.byref(r.s), C.c_int64(a), C.c_int64(b)))


def mark_duplicates(batch, header, n_threads=1, want_adapt=False):
    """Sets 0x400 in batch.flag in place. Raises on 'Invalid QUAL character'."""
    r, h = OracleReads(batch), OracleHeader(header)
    upos = np.zeros(batch.n, dtype=np.int32)
    score = np.zeros(batch.n, dtype=np.int32)
    rc = lib().orc_mark_duplicates(C.byref(r.s), C.byref(h.s), C.c_int(n_threads), _p(upos), _p(score))
    if rc != 0:
        raise ValueError("Invalid QUAL character")
    return (upos, score) if want_adapt else None


class _DupMetrics(C.Structure):
    _fields_ = [(k, C.c_int64) for k in ("unpaired_reads_examined", "read_pairs_examined", "secondary_or_supplementary", "unmapped_reads",
                                         "unpaired_read_duplicates", "read_pair_duplicates", "read_pair_optical_duplicates",
                                         "estimated_library_size")] + [("percent_duplication", C.c_double), ("roi", C.c_double * 100), ("has_roi", C.c_int32)]


COUNTERS = ("unpaired_reads_examined", "read_pairs_examined", "secondary_or_supplementary", "unmapped_reads",
            "unpaired_read_duplicates", "read_pair_duplicates", "read_pair_optical_duplicates")


class OpticalMetrics:
    """Result of MarkOpticalDuplicates: per slot (0 = "Unknown Library", l+1 = library l) the seven counters, the derived
    metrics and the three count histograms as {key: count} dicts (all, non-optical, optical)."""

    def __init__(self, lib_names):
        self.lib_names = ["Unknown Library"] + list(lib_names)
        self.counters, self.library_size, self.percent_duplication, self.roi, self.hist = [], [], [], [], []


def markdup_optical(batch, header, order=None, pixel_distance=100, n_threads=1, metrics_path=None, command_line="", started_on=""):
    """MarkDuplicates(alsoOpticals=True) (sets 0x400 in batch.flag in place) + MarkOpticalDuplicates over ``order``."""
    L = lib()
    L.orc_markdup_optical.restype = C.c_void_p
    L.orc_optical_hist.restype = C.c_int64
    r, h = OracleReads(batch), OracleHeader(header)
    o = np.ascontiguousarray(order, dtype=np.int64) if order is not None else None
    res = C.c_void_p(L.orc_markdup_optical(C.byref(r.s), C.byref(h.s), C.c_int(n_threads), _p(o), C.c_int(pixel_distance)))
    try:
        err = L.orc_optical_error(res)
        if err == -1:
            raise ValueError("Invalid QUAL character")
        if err == -2:
            raise ValueError("origin for duplicate read pair unknown")
        if err == -3:
            raise ValueError("strconv.ParseInt: parsing a QNAME tile field: invalid syntax")
        out = OpticalMetrics(h.lib_names)
        for slot in range(L.orc_optical_slots(res)):
            m = _DupMetrics()
            L.orc_optical_get(res, C.c_int(slot), C.byref(m))
            out.counters.append({k: int(getattr(m, k)) for k in COUNTERS})
            out.library_size.append(int(m.estimated_library_size))
            out.percent_duplication.append(float(m.percent_duplication))
            out.roi.append(list(m.roi) if m.has_roi else None)
            hs = []
            for which in range(3):
                n = L.orc_optical_hist(res, C.c_int(slot), C.c_int(which), None, None, C.c_int64(0))
                keys, cnt = np.zeros(n, np.int64), np.zeros(n, np.int64)
                L.orc_optical_hist(res, C.c_int(slot), C.c_int(which), _p(keys), _p(cnt), C.c_int64(n))
                hs.append({int(k): int(v) for k, v in zip(keys, cnt)})
            out.hist.append(hs)
        if metrics_path is not None:
            arr = (C.c_char_p * max(1, len(h.lib_names)))(*[x.encode() for x in h.lib_names])
            if L.orc_optical_print(res, arr, metrics_path.encode(), command_line.encode(), started_on.encode()) != 0:
                raise OSError("cannot write " + metrics_path)
        return out
    finally:
        L.orc_optical_free(res)


class Reference:
    """Concatenated contig bases (1 B/base, as fasta.MappedFasta.Seq returns) + known sites."""

    def __init__(self, header, contig_bases, sites=None):
        """contig_bases: list of uint8 arrays per contig; sites: list of (k,2) int32 arrays per contig, sorted+flattened."""
        nc = len(header.SQ)
        self.ref_off = np.zeros(nc + 1, dtype=np.uint64)
        self.ref_off[1:] = np.cumsum([len(b) for b in contig_bases])
        self.ref = np.concatenate([np.asarray(b, dtype=np.uint8) for b in contig_bases]) if nc else np.zeros(0, np.uint8)
        sites = sites or [np.zeros((0, 2), np.int32) for _ in range(nc)]
        self.site_off = np.zeros(nc + 1, dtype=np.uint64)
        self.site_off[1:] = np.cumsum([len(s) for s in sites])
        self.sites = (np.concatenate([np.asarray(s, dtype=np.int32).reshape(-1, 2) for s in sites]).reshape(-1)
                      if nc else np.zeros(0, np.int32))
        self.sites = np.ascontiguousarray(self.sites, dtype=np.int32)
        if self.sites.size == 0:
            self.sites = np.zeros(2, dtype=np.int32)


ERRORS = {-3: "reference coordinate matches a non-existing base in read", -4: "clip out of range", -5: "cigar too long for oracle",
          -6: "read extends past contig end", -7: "context covariate would panic (non-N IUPAC base at read start)",
          -8: "cycle value exceeds maximum cycle value", -9: "BQSR requires input with read groups"}


def bqsr_gather(batch, header, reference, max_cycle=500, n_threads=1):
    r, h = OracleReads(batch), OracleHeader(header)
    t = OracleTables(h.s.n_cov, max_cycle)
    rc = lib().orc_bqsr_gather(C.byref(r.s), C.byref(h.s), _p(reference.ref), _p(reference.ref_off), _p(reference.sites),
                               _p(reference.site_off), C.byref(t.s), C.c_int(n_threads))
    if rc != 0:
        raise ValueError(ERRORS.get(rc, f"oracle error {rc}"))
    return t


def bqsr_finalize(t):
    lib().orc_bqsr_finalize(C.byref(t.s))


def bqsr_apply(batch, header, t, quantize_levels=0, sqq=None, n_threads=1):
    r, h = OracleReads(batch), OracleHeader(header)
    sq = np.ascontiguousarray(sqq if sqq is not None else [], dtype=np.uint8)
    rc = lib().orc_bqsr_apply(C.byref(r.s), C.byref(h.s), C.byref(t.s), C.c_int(quantize_levels), _p(sq) if sq.size else None,
                              C.c_int(sq.size), C.c_int(n_threads))
    if rc != 0:
        raise ValueError(ERRORS.get(rc, f"oracle error {rc}"))


def bqsr_report(t, cov_names, path, prefix="GATK"):
    arr = (C.c_char_p * len(cov_names))(*[s.encode() for s in cov_names])
    rc = lib().orc_bqsr_report(C.byref(t.s), arr, prefix.encode(), path.encode())
    if rc != 0:
        raise OSError("cannot write report")


def combined(t, cov):
    rq, o, m, e, ex = C.c_double(), C.c_int64(), C.c_int64(), C.c_uint8(), C.c_int()
    lib().orc_combined(C.byref(t.s), C.c_int(cov), C.byref(rq), C.byref(o), C.byref(m), C.byref(e), C.byref(ex))
    return dict(reported=rq.value, obs=o.value, mis=m.value, emp=e.value, exists=bool(ex.value))
