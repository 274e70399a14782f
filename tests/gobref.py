"""A small, independent restatement of Go's encoding/gob wire format (TEST INFRASTRUCTURE), enough for the two intermediate files of an
`elprep sfm` run: filters.BaseRecalibratorTables (.elrecal) and map[string]*DuplicatesCtr.  Written from the encoding/gob specification;
``POINT_EXAMPLE`` is the byte example of that specification (type Point struct{X, Y int}; Point{22, 33}) and pins the primitives."""
import io

POINT_EXAMPLE = bytes.fromhex("1fff810301010550 6f696e7401ff8200 0102010158010400 0101590104000000 07ff82012c014200".replace(" ", ""))

T_INT, T_UINT, T_STRING = 2, 3, 6


def enc_uint(v):
    if v < 128:
        return bytes([v])
    raw = v.to_bytes((v.bit_length() + 7) // 8, "big")
    return bytes([256 - len(raw)]) + raw


def enc_int(v):
    return enc_uint(((~v) << 1) | 1 if v < 0 else v << 1)


def enc_str(s):
    b = s.encode()
    return enc_uint(len(b)) + b


def frame(payload):
    return enc_uint(len(payload)) + payload


def _common(name, tid):
    out, prev = b"", -1
    if name:
        out += enc_uint(1) + enc_str(name); prev = 0
    return out + enc_uint(1 - prev) + enc_int(tid) + b"\0"


def def_struct(tid, name, fields):
    body = enc_int(-tid) + enc_uint(3) + enc_uint(1) + _common(name, tid) + enc_uint(1) + enc_uint(len(fields))
    for fn, ft in fields:
        body += enc_uint(1) + enc_str(fn) + enc_uint(1) + enc_int(ft) + b"\0"
    return frame(body + b"\0\0")


def def_map(tid, name, key, elem):
    return frame(enc_int(-tid) + enc_uint(4) + enc_uint(1) + _common(name, tid) + enc_uint(1) + enc_int(key) + enc_uint(1) + enc_int(elem) + b"\0\0")


def enc_struct(values, kinds):
    """values in field order; kinds: 'u', 'i', 's' or a callable; zero values are omitted"""
    out, prev = b"", -1
    for f, (v, k) in enumerate(zip(values, kinds)):
        if not v and not callable(k):
            continue
        out += enc_uint(f - prev) + ({"u": enc_uint, "i": enc_int, "s": enc_str}[k](v) if not callable(k) else k(v)); prev = f
    return out + b"\0"


def encode_point(x, y, tid=65):
    return def_struct(tid, "Point", [("X", T_INT), ("Y", T_INT)]) + frame(enc_int(tid) + enc_struct([x, y], "ii"))


def encode_elrecal(tables):
    """tables: three dicts {(qual, covariate, read_group): (empirical_quality, observations, mismatches)} -> bytes, ids as a fresh Go process assigns them"""
    out = def_struct(64, "BaseRecalibratorTables", [("QualityScores", 67), ("Cycles", 67), ("Contexts", 67)])
    out += def_map(67, "bqsrTable", 65, 66)
    out += def_struct(65, "bqsrTableKey", [("Qual", T_UINT), ("Covariate", T_INT), ("ReadGroup", T_STRING)])
    out += def_struct(66, "bqsrEntry", [("EmpiricalQuality", T_UINT), ("Observations", T_INT), ("Mismatches", T_INT)])

    def enc_map(t):
        b = enc_uint(len(t))
        for (q, cov, rg), (e, o, m) in t.items():
            b += enc_struct([q, cov, rg], "uis") + enc_struct([e, o, m], "uii")
        return b
    body = enc_int(64)
    prev = -1
    for f, t in enumerate(tables):
        body += enc_uint(f - prev) + enc_map(t); prev = f
    return out + frame(body + b"\0")


COUNTERS = ("UnpairedReadsExamined", "ReadPairsExamined", "SecondaryOrSupplementaryReads", "UnmappedReads", "UnpairedReadDuplicates", "ReadPairDuplicates", "ReadPairOpticalDuplicates")


def encode_metrics(ctrs):
    """ctrs: {library: [7 counters]} -> gob of map[string]*DuplicatesCtr"""
    out = def_map(65, "", T_STRING, 64) + def_struct(64, "DuplicatesCtr", [(n, T_INT) for n in COUNTERS])
    body = enc_int(65) + b"\0" + enc_uint(len(ctrs))
    for lib, c in ctrs.items():
        body += enc_str(lib) + enc_struct(list(c), "i" * 7)
    return out + frame(body)


# ---- decoder ----
class _R:
    def __init__(self, b):
        self.f = io.BytesIO(b)

    def u(self):
        c = self.f.read(1)[0]
        if c < 128:
            return c
        return int.from_bytes(self.f.read(256 - c), "big")

    def i(self):
        x = self.u()
        return ~(x >> 1) if x & 1 else x >> 1

    def s(self):
        return self.f.read(self.u()).decode()


def decode(data):
    """-> the first Encode()d value as nested dict / list of pairs / int / str"""
    r, types = _R(data), {}

    def common():
        name, tid, f = "", 0, -1
        while True:
            d = r.u()
            if not d:
                return name, tid
            f += d
            if f == 0:
                name = r.s()
            else:
                tid = r.i()

    def value(tid):
        if tid in (2,):
            return r.i()
        if tid in (1, 3):
            return r.u()
        if tid in (5, 6):
            return r.s()
        t = types[tid]
        if t[0] == "struct":
            out, f = {}, -1
            while True:
                d = r.u()
                if not d:
                    return out
                f += d
                out[t[1][f][0]] = value(t[1][f][1])
        if t[0] == "map":
            return [(value(t[1]), value(t[2])) for _ in range(r.u())]
        raise ValueError(t)
    while True:
        ln = r.u(); end = r.f.tell() + ln
        tid = r.i()
        if tid < 0:
            f = -1
            while True:
                d = r.u()
                if not d:
                    break
                f += d
                if f == 2:
                    g, fields = -1, []
                    while True:
                        d2 = r.u()
                        if not d2:
                            break
                        g += d2
                        if g == 0:
                            common()
                        else:
                            for _ in range(r.u()):
                                fn, ft, h = "", 0, -1
                                while True:
                                    d3 = r.u()
                                    if not d3:
                                        break
                                    h += d3
                                    if h == 0:
                                        fn = r.s()
                                    else:
                                        ft = r.i()
                                fields.append((fn, ft))
                    types[-tid] = ("struct", fields)
                elif f == 3:
                    g, key, elem = -1, 0, 0
                    while True:
                        d2 = r.u()
                        if not d2:
                            break
                        g += d2
                        if g == 0:
                            common()
                        elif g == 1:
                            key = r.i()
                        else:
                            elem = r.i()
                    types[-tid] = ("map", key, elem)
                else:
                    raise ValueError("unsupported wire type")
            assert r.f.tell() == end
            continue
        if types.get(tid, ("",))[0] != "struct":
            r.u()
        return value(tid)
