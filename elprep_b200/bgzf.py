"""ctypes wrappers of the context-free host utilities (include/elprep_b200.h): BGZF inflate / deflate on a host thread pool
and the BAM header walk.  No GPU involved."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


class BgzfError(ValueError):
    pass


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def inflate(data, n_threads=8):
    """BGZF blocks (uint8 array / bytes) -> uint8 array of the uncompressed bytes"""
    L = _lib.load()
    d = np.frombuffer(data, dtype=np.uint8) if isinstance(data, (bytes, bytearray, memoryview)) else np.ascontiguousarray(data, dtype=np.uint8)
    bound = int(L.elp_bgzf_inflate_bound(_vp(d), d.size))
    if bound < 0:
        raise BgzfError(f"malformed BGZF input (code {bound})")
    out = np.empty(max(bound, 1), dtype=np.uint8)
    n = C.c_uint64()
    rc = L.elp_bgzf_inflate(_vp(d), d.size, _vp(out), out.size, C.byref(n), n_threads)
    if rc != 0:
        raise BgzfError(f"BGZF inflate failed (code {rc})")
    return out[:int(n.value)]


def deflate(data, level=-1, n_threads=8, write_eof=True):
    """uint8 array / bytes -> BGZF blocks (uint8 array), optionally terminated by the EOF marker block"""
    L = _lib.load()
    d = np.frombuffer(data, dtype=np.uint8) if isinstance(data, (bytes, bytearray, memoryview)) else np.ascontiguousarray(data, dtype=np.uint8)
    out = np.empty(int(L.elp_bgzf_deflate_bound(d.size)), dtype=np.uint8)
    n = C.c_uint64()
    rc = L.elp_bgzf_deflate(_vp(d), d.size, _vp(out), out.size, C.byref(n), level, n_threads, int(write_eof))
    if rc != 0:
        raise BgzfError(f"BGZF deflate failed (code {rc})")
    return out[:int(n.value)]


def bam_header_size(bam):
    """-> (bytes before the first alignment record, number of reference sequences) of an inflated BAM file"""
    L = _lib.load()
    d = np.frombuffer(bam, dtype=np.uint8) if isinstance(bam, (bytes, bytearray, memoryview)) else np.ascontiguousarray(bam, dtype=np.uint8)
    nref = C.c_int32()
    sz = int(L.elp_bam_header_size(_vp(d), d.size, C.byref(nref)))
    if sz < 0:
        raise BgzfError("not a BAM header")
    return sz, int(nref.value)
