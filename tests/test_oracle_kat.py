"""Known-answer tests pinning the oracle to the reference source.
Each case cites the reference lines it was hand-derived from.  The interval cases are the
reference's own golden vectors (intervals/intervals_test.go:53-213)."""
import ctypes as C
import json
import math
import os

import numpy as np
import pytest

from elprep_b200 import sam

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def cig(s):
    return np.array(sam.encode_cigar(s), dtype=np.uint32)


def upos(orc, pos, rev, c):
    a = cig(c)
    return orc.lib().orc_unclipped_position(C.c_int32(pos), C.c_int(rev), a.ctypes.data_as(C.c_void_p), C.c_int32(len(a)))


# ---- C1..C6 computeUnclippedPosition (filters/mark-duplicates.go:79-110)
@pytest.mark.parametrize("pos,rev,c,exp", [
    (100, 0, "5S145M", 95), (100, 0, "3H5S142M", 92), (100, 1, "145M5S", 249),
    (100, 1, "5S140M2D5M", 246), (100, 1, "10M2I138M", 247), (100, 0, "*", 100), (100, 1, "*", 100)])
def test_unclipped_position(orc, pos, rev, c, exp):
    assert upos(orc, pos, rev, c) == exp


# ---- C7..C9 computePhredScore (filters/mark-duplicates.go:36-68)
def phred(orc, q):
    a = np.array(q, dtype=np.uint8)
    inv = C.c_int()
    s = orc.lib().orc_phred_score(a.ctypes.data_as(C.c_void_p), C.c_int32(len(a)), C.byref(inv))
    return s, inv.value


def test_phred_score(orc):
    assert phred(orc, [37] * 150) == (5550, 0)
    assert phred(orc, [2, 12, 14, 15, 23]) == (38, 0)
    assert phred(orc, [30, 94])[1] == 1           # "Invalid QUAL character"
    assert phred(orc, [30, 127])[1] == 1
    assert phred(orc, [128 + 30]) == (30, 0)      # byte(char<<1) wraps: 158 aliases 30
    assert phred(orc, []) == (0, 0)


# ---- C19/C20 CoordinateLess + modFlag (sam/sam-types.go:408-473)
def test_mod_flag(orc):
    L = orc.lib()
    assert L.orc_mod_flag(0x20 | 0x8) == 0            # unpaired: 0x8, 0x20 ignored
    assert L.orc_mod_flag(0x4 | 0x10) == 0x4          # unmapped: 0x10 ignored
    assert L.orc_mod_flag(0x1 | 0x8 | 0x20) == 0x9    # mate unmapped: 0x20 ignored
    assert L.orc_mod_flag(99) == 99


def _hdr():
    return sam.Header(sq=[{"SN": f"chr{i}", "LN": 100000} for i in range(1, 5)],
                      rg=[{"ID": "rg1", "LB": "libA", "PU": "pu1"}, {"ID": "rg2", "LB": "libA"}, {"ID": "rg3"}])


def test_coordinate_less(orc):
    h = _hdr()
    recs = [dict(QNAME="a", RNAME="*", POS=0, FLAG=4), dict(QNAME="a", RNAME="chr4", POS=5, FLAG=0),
            dict(QNAME="b", RNAME="chr4", POS=5, FLAG=16), dict(QNAME="c", RNAME="chr4", POS=5, FLAG=0),
            dict(QNAME="c", RNAME="chr4", POS=5, FLAG=0, MAPQ=3),
            dict(QNAME="p", RNAME="chr1", POS=9, FLAG=1, RNEXT="*", PNEXT=7), dict(QNAME="p", RNAME="chr1", POS=9, FLAG=1, RNEXT="chr2", PNEXT=1),
            dict(QNAME="p", RNAME="chr1", POS=9, FLAG=1, RNEXT="chr2", PNEXT=1, TLEN=-5)]
    b = sam.AlignmentBatch.from_records(h, recs)
    less = lambda i, j: orc.coordinate_less(b, i, j)
    assert less(1, 0) and not less(0, 1)          # refid -1 sorts last (:428-432)
    assert less(1, 2) and not less(2, 1)          # forward before reverse (:437-438)
    assert less(1, 3) and not less(3, 1)          # QNAME bytes (:439-446)
    assert less(3, 4) and not less(4, 3)          # MAPQ (:454-457)
    assert less(5, 6) and not less(6, 5)          # NextREFID signed: -1 first (:462-465)
    assert less(7, 6) and not less(6, 7)          # TLEN last (:472)
    perm = orc.coordinate_sort(b)
    assert list(perm) == [5, 7, 6, 1, 3, 4, 2, 0]


# ---- C10, C21..C24 fragment / pair rules (filters/mark-duplicates.go:177-396)
def _dupflags(orc, recs, threads=1):
    h = _hdr()
    b = sam.AlignmentBatch.from_records(h, recs)
    orc.mark_duplicates(b, h, n_threads=threads)
    return [bool(f & 0x400) for f in b.flag]


def R(q, flag, pos, qual, rname="chr1", cigar="4M", rg="rg1", **kw):
    return dict(QNAME=q, FLAG=flag, RNAME=rname, POS=pos, CIGAR=cigar, SEQ="ACGT", QUAL=qual, RG=rg, **kw)


def test_fragment_rule(orc):
    # C21: ties -> smallest QNAME wins (:222-243)
    assert _dupflags(orc, [R("b", 0, 10, [25] * 4), R("a", 0, 10, [25] * 4), R("c", 0, 10, [20] * 4)]) == [True, False, True]
    # C22: a pair read shadows all fragments, itself untouched (:225-227,245-252)
    assert _dupflags(orc, [R("p", 0x1 | 0x40, 10, [15] * 4, RNEXT="=", PNEXT=50), R("f", 0, 10, [40] * 4)]) == [False, True]
    assert _dupflags(orc, [R("f", 0, 10, [40] * 4), R("p", 0x1 | 0x40, 10, [15] * 4, RNEXT="=", PNEXT=50)]) == [True, False]
    # C10: paired with unmapped mate (0x1|0x8) is a true fragment (:177-184)
    assert _dupflags(orc, [R("x", 0x1 | 0x8 | 0x40, 10, [40] * 4), R("y", 0, 10, [30] * 4)]) == [False, True]
    # different library / strand / unclipped pos -> separate groups (:188-216)
    assert _dupflags(orc, [R("x", 0, 10, [40] * 4), R("y", 0, 10, [30] * 4, rg="rg3"), R("z", 16, 10, [30] * 4), R("w", 0, 12, [30] * 4, cigar="2S2M")]) == [False, False, False, True]
    # secondary / supplementary / unmapped never enter (:436)
    assert _dupflags(orc, [R("x", 0, 10, [40] * 4), R("y", 0x100, 10, [30] * 4), R("z", 0x800, 10, [30] * 4), R("u", 4, 10, [30] * 4)]) == [False, False, False, False]


def test_pair_rule(orc):
    def pair(q, p1, p2, qual, flags=(99, 147)):
        return [R(q, flags[0], p1, qual, RNEXT="=", PNEXT=p2), R(q, flags[1], p2, qual, RNEXT="=", PNEXT=p1)]
    # C23: equal score -> smaller QNAME wins; both mates of the loser marked (:375-395)
    assert _dupflags(orc, pair("q2", 100, 300, [30] * 4) + pair("q1", 100, 300, [30] * 4)) == [True, True, False, False]
    assert _dupflags(orc, pair("q1", 100, 300, [30] * 4) + pair("q2", 100, 300, [31] * 4)) == [True, True, False, False]
    # mates given in either order form the same key (C24 orientation :347-353)
    a = pair("q1", 100, 300, [30] * 4)
    b = pair("q2", 100, 300, [20] * 4)
    assert _dupflags(orc, a + b[::-1]) == [False, False, True, True]
    # different second coordinate -> not duplicates of each other (but first mates share a fragment group with pairs only)
    assert _dupflags(orc, pair("q1", 100, 300, [30] * 4) + pair("q2", 100, 301, [20] * 4)) == [False, False, False, False]
    # unmatched mate never pair-marked
    assert _dupflags(orc, pair("q1", 100, 300, [30] * 4) + [R("q2", 99, 100, [20] * 4, RNEXT="=", PNEXT=300)]) == [False, False, False]


def test_markdup_threads_agree(orc):
    rng = np.random.default_rng(5)
    recs = []
    for i in range(3000):
        p1 = int(rng.integers(1, 60)); p2 = p1 + int(rng.integers(1, 6))
        q = [int(x) for x in rng.integers(10, 41, size=4)]
        if rng.random() < 0.7:
            recs += [R(f"r{i:05d}", 99, p1, q, RNEXT="=", PNEXT=p2), R(f"r{i:05d}", 147, p2, q, RNEXT="=", PNEXT=p1)]
        else:
            recs.append(R(f"r{i:05d}", 16 if rng.random() < .5 else 0, p1, q))
    assert _dupflags(orc, recs, 1) == _dupflags(orc, recs, 4)


# ---- C26 intervals: the reference's own golden vectors (intervals/intervals_test.go)
def test_intervals_reference_vectors(orc):
    L = orc.lib()
    with open(os.path.join(GOLD, "intervals_kat.json")) as f:
        kat = json.load(f)
    for inp, exp in kat["flatten"]:
        a = np.array(inp, dtype=np.int32).reshape(-1)
        n = L.orc_flatten(a.ctypes.data_as(C.c_void_p), C.c_int64(len(inp)))
        assert a[:2 * n].reshape(-1, 2).tolist() == exp
    for iv, s, e, exp in kat["overlap"]:
        a = np.array(iv, dtype=np.int32).reshape(-1)
        assert bool(L.orc_overlap(a.ctypes.data_as(C.c_void_p), C.c_int64(len(iv)), C.c_int32(s), C.c_int32(e))) == exp
    for iv, s, e, exp in kat["intersect"]:
        a = np.array(iv, dtype=np.int32).reshape(-1)
        lo, hi = C.c_int64(), C.c_int64()
        L.orc_intersect(a.ctypes.data_as(C.c_void_p), C.c_int64(len(iv)), C.c_int32(s), C.c_int32(e), C.byref(lo), C.byref(hi))
        assert iv[lo.value:hi.value] == exp
    # Flatten 7: property test from the reference (makeLargeIntervalsSlice)
    rng = np.random.default_rng(1)
    n = 0x3000
    iv = np.zeros((n, 2), dtype=np.int32); iv[0] = (0, 3)
    for i in range(1, n):
        iv[i, 0] = iv[i - 1, 1] - 1 if rng.integers(100) < 20 else iv[i - 1, 1] + 1
        iv[i, 1] = iv[i, 0] + 3
    a = iv.reshape(-1).copy()
    m = L.orc_flatten(a.ctypes.data_as(C.c_void_p), C.c_int64(n))
    out = a[:2 * m].reshape(-1, 2)
    assert (out[:, 0] <= out[:, 1]).all() and (out[1:, 0] > out[:-1, 1]).all()


# ---- C11..C14 cycle covariate (filters/bqsr.go:376-387)
def test_cycle(orc):
    cyc = np.zeros(150, dtype=np.int32)
    for flag, first, last in [(99, 1, 150), (83, 150, 1), (163, -1, -150), (147, -150, -1)]:
        orc.lib().orc_probe_cycle(C.c_uint16(flag), C.c_int32(150), cyc.ctypes.data_as(C.c_void_p))
        assert (cyc[0], cyc[149]) == (first, last)


# ---- C15..C18 context covariate (filters/bqsr.go:64-146,312-362)
def ctx(orc, seq, qual, rev):
    s = np.frombuffer(sam.encode_seq(seq), dtype=np.uint8)
    q = np.array(qual, dtype=np.uint8)
    k = np.full(len(seq), -7, dtype=np.int32)
    n = orc.lib().orc_probe_context(s.ctypes.data_as(C.c_void_p), q.ctypes.data_as(C.c_void_p), C.c_int32(len(seq)), C.c_int(rev), k.ctypes.data_as(C.c_void_p))
    return list(k[:max(n, 0)]) if n >= 0 else n


def test_context(orc):
    assert ctx(orc, "ACGT", [30] * 4, 0) == [-1, 66, 146, 226]
    assert ctx(orc, "AACG", [30] * 4, 1) == [242, 226, 146, -1]
    assert ctx(orc, "ACGTA", [2, 2, 30, 30, 2], 0) == [-1, -1, -1, 226, -1]
    assert ctx(orc, "ACGT", [2, 2, 1, 0], 0) == []                       # all quals <= 2 -> nil seq -> no keys (:329-331)
    assert ctx(orc, "ANGT", [30] * 4, 0) == [-1, -1, -1, 226]           # N resets with penalty 2 (:115-128)
    assert ctx(orc, "NCGT", [30] * 4, 0) == [-1, -1, 146, 226]
    assert ctx(orc, "A", [30], 0) == [-1]
    assert ctx(orc, "AMGT", [30] * 4, 0) == -100                        # Go would panic (index -1) on a non-N IUPAC code at the start


# ---- C31..C36 clipping (filters/utils.go:148-534)
def clip(orc, pos, flag, pnext, tlen, c, lseq, refid=0, nref=0):
    a = cig(c)
    lo, hi, np_ = C.c_int32(), C.c_int32(), C.c_int32()
    out = np.zeros(64, dtype=np.uint32)
    n = orc.lib().orc_probe_clip(C.c_int32(pos), C.c_uint16(flag), C.c_int32(pnext), C.c_int32(tlen), C.c_int32(refid), C.c_int32(nref),
                                 a.ctypes.data_as(C.c_void_p), C.c_int32(len(a)), C.c_int32(lseq), C.byref(lo), C.byref(hi), C.byref(np_),
                                 out.ctypes.data_as(C.c_void_p), C.c_int(64))
    if n < 0:
        return n
    return lo.value, hi.value, np_.value, sam.decode_cigar(out[:n])


def test_clipping(orc):
    assert clip(orc, 100, 0, 0, 0, "5S145M", 150) == (5, 150, 100, "5H145M")            # C31
    assert clip(orc, 1000, 99, 970, 120, "150M", 150) == (0, 120, 1000, "120M30H")      # C32
    assert clip(orc, 1000, 83, 1030, -120, "150M", 150) == (30, 150, 1030, "30H120M")   # C33
    assert clip(orc, 1000, 99, 1000, 100, "100M50S", 150) == (0, 100, 1000, "100M50H")  # C36
    assert clip(orc, 100, 0, 0, 0, "150M", 150) == (0, 150, 100, "150M")
    assert clip(orc, 100, 0, 0, 0, "5S140M5S", 150) == (5, 145, 100, "5H140M5H")
    assert clip(orc, 100, 0, 0, 0, "3H5S140M5S2H", 150) == (5, 145, 100, "8H140M7H")
    # deletion adjacent to the clip is absorbed into H and shifts POS (utils.go:473-504, 351-375)
    assert clip(orc, 1000, 83, 1010, -130, "10M2D140M", 150) == (10, 150, 1012, "12H140M")
    # boundary 1010 = first base after the insertion = read index 15 (utils.go:296-314): clip [0,15], H = 16 - 5 inserted
    assert clip(orc, 1000, 83, 1011, -130, "10M5I135M", 150) == (16, 150, 1011, "11H134M")
    # clip point inside an insertion (utils.go:429-437): forward read, boundary 1012 -> read index 17 -> keep [0,17)
    assert clip(orc, 1000, 99, 995, 12, "10M5I135M", 150) == (0, 17, 1000, "10M5I2M133H")


def test_read_coordinate(orc):                                                          # C34 (utils.go:267-349)
    a = cig("10M5D140M")
    ok = C.c_int()
    f = lambda ref, tail: (orc.lib().orc_probe_readcoord(a.ctypes.data_as(C.c_void_p), C.c_int32(len(a)), C.c_int(1000), C.c_int(ref), C.c_int(tail), C.byref(ok)), ok.value)
    assert f(1012, 0) == (9, 1) and f(1012, 1) == (10, 1)
    assert f(1005, 0) == (5, 1) and f(1015, 0) == (10, 1)
    assert f(999, 0)[1] == 0


# ---- C25, C30 + BQSR numerics (filters/bqsr.go:553-706)
def test_prior_cache(orc):
    for d in range(20):
        assert orc.lib().orc_prior_cache(C.c_int(d)) == math.log10(0.9 * math.exp(-d * d / 0.5))
    assert orc.lib().orc_prior_cache(C.c_int(20)) == -1.7976931348623157e308


def _py_empirical(obs, mis, prior):
    """independent restatement of bqsr.go:593-649 with scipy/math (argmax over 61 bins)"""
    from scipy import special
    n, k = obs + 2, mis + 1
    best, bi = -1.7976931348623157e308, 0
    for i in range(61):
        d = min(abs(int(i - prior)), 20)
        p1 = -1.7976931348623157e308 if d == 20 else math.log10(0.9 * math.exp(-d * d / 0.5))
        l = i / -10.0
        if l == 0.0:
            p2 = -1.7976931348623157e308
        else:
            lg = lambda x: float(special.gammaln(float(x))) * math.log10(math.e)
            p2 = (lg(n + 1) - lg(k + 1) - lg(n - k + 1)) + l * k + math.log10(1.0 - 10.0 ** l) * (n - k)
        if best < p1 + p2:
            best, bi = p1 + p2, i
    return min(bi, 93)


def test_empirical_quality(orc):
    eq = orc.lib().orc_empirical_quality
    assert eq(0, 0, 30.0) == 30                     # no data: the prior decides
    assert eq(10**6, 10**3, 30.0) == 30             # 1e-3 error rate -> Q30
    rng = np.random.default_rng(3)
    for _ in range(400):
        obs = int(10 ** rng.uniform(0, 9))
        mis = int(obs * 10 ** rng.uniform(-6, -0.3))
        prior = float(rng.choice([rng.integers(2, 60), rng.uniform(2, 60)]))
        assert eq(obs, mis, prior) == _py_empirical(obs, mis, prior), (obs, mis, prior)


def test_queryname_order(orc):
    # By(QNAMELess).ParallelStableSort (sam-types.go:479-481): bytewise QNAME order, a prefix first, equal names keep arrival order
    h = _hdr()
    recs = [dict(QNAME=q, FLAG=f, RNAME="chr1", POS=p) for q, f, p in
            [("r10", 0, 5), ("r2", 0, 9), ("r1", 16, 7), ("r1", 0, 3), ("r", 0, 1), ("R9", 0, 2), ("r1:x", 0, 4)]]
    b = sam.AlignmentBatch.from_records(h, recs)
    assert list(orc.queryname_sort(b)) == [5, 4, 2, 3, 0, 6, 1]
