"""Known-answer tests for the oracle's MarkOpticalDuplicates restatement (filters/mark-optical-duplicates.go).
Hand-derived from the cited lines; an independent brute-force clustering cross-checks the graph path."""
import os

import numpy as np
import pytest

from elprep_b200 import sam


def _hdr(one_lib=False):
    rg = [{"ID": "rg1", "LB": "libA", "PU": "pu1"}, {"ID": "rg2", "LB": "libA"}, {"ID": "rg3"}]
    if not one_lib:
        rg.append({"ID": "rg4", "LB": "libB"})
    return sam.Header(sq=[{"SN": f"chr{i}", "LN": 100000} for i in range(1, 5)], rg=rg)


def R(q, flag, pos, qual, rname="chr1", cigar="4M", rg="rg1", **kw):
    return dict(QNAME=q, FLAG=flag, RNAME=rname, POS=pos, CIGAR=cigar, SEQ="ACGT", QUAL=qual, RG=rg, **kw)


def pair(q, p1, p2, score, flags=(99, 147), rg="rg1"):
    return [R(q, flags[0], p1, [score] * 4, RNEXT="=", PNEXT=p2, rg=rg), R(q, flags[1], p2, [score] * 4, RNEXT="=", PNEXT=p1, rg=rg)]


def run(orc, recs, h=None, **kw):
    h = h or _hdr()
    b = sam.AlignmentBatch.from_records(h, recs)
    return orc.markdup_optical(b, h, **kw), b


def test_counters_and_small_lists(orc):
    # one group of three pairs on tile 1101: q1 (origin, best score) and q2 are 60 px apart, q3 is far away
    recs = pair("M:1:FC:1:1101:1000:2000", 100, 300, 40) + pair("M:1:FC:1:1101:1050:2060", 100, 300, 30) + pair("M:1:FC:1:1101:5000:2000", 100, 300, 20)
    recs += [R("frag", 0, 500, [30] * 4), R("fragdup", 0, 500, [20] * 4), R("un", 4, 0, [30] * 4, rname="*"), R("sec", 0x100, 10, [30] * 4),
             R("nolib", 0, 700, [30] * 4, rg="rg3")]
    m, b = run(orc, recs)
    assert m.lib_names == ["Unknown Library", "libA", "libB"]
    a = m.counters[1]
    # :473-494: examined counts are per read, pairs halved (:503-505); the two losing pairs are duplicates (:189)
    assert a == dict(unpaired_reads_examined=2, read_pairs_examined=3, secondary_or_supplementary=1, unmapped_reads=1,
                     unpaired_read_duplicates=1, read_pair_duplicates=2, read_pair_optical_duplicates=1)
    assert m.counters[0]["unpaired_reads_examined"] == 1 and m.counters[2]["read_pairs_examined"] == 0
    # :310-325: list = origin + 2 duplicates, one optical -> all[3], nonOptical[2], optical[1+1]
    assert m.hist[1] == [{3: 1}, {2: 1}, {2: 1}]
    # percentDuplication (:524) = (1 + 2*2) / (2 + 3*2)
    assert m.percent_duplication[1] == pytest.approx(5 / 8)
    assert np.isnan(m.percent_duplication[2])
    # pixel distance is inclusive (unpedantic.go:32-34)
    m2, _ = run(orc, recs, pixel_distance=59)
    assert m2.counters[1]["read_pair_optical_duplicates"] == 0 and m2.hist[1] == [{3: 1}, {3: 1}, {}]
    m3, _ = run(orc, recs, pixel_distance=60)
    assert m3.counters[1]["read_pair_optical_duplicates"] == 1


def test_strand_lists_and_read_group(orc):
    # the list read is the first-of-pair read (:216-221); lists are split by ITS strand (:287-301)
    near = ["M:1:FC:1:1101:1000:2000", "M:1:FC:1:1101:1001:2001", "M:1:FC:1:1101:1002:2002"]
    recs = pair(near[0], 100, 300, 40) + pair(near[1], 100, 300, 30) + pair(near[2], 100, 300, 20, flags=(163, 83))   # third: first-of-pair is the reverse read
    m, _ = run(orc, recs)
    assert m.counters[1]["read_pair_optical_duplicates"] == 1          # forward list {q0,q1}: 1; reverse list {q2}: 0
    assert m.hist[1] == [{3: 1}, {2: 1}, {2: 1}]
    # different read groups never cluster (:83), even inside one library
    recs = pair(near[0], 100, 300, 40) + pair(near[1], 100, 300, 30, rg="rg2")
    m, _ = run(orc, recs)
    assert m.counters[1]["read_pair_duplicates"] == 1 and m.counters[1]["read_pair_optical_duplicates"] == 0
    # different tiles never cluster (:89)
    recs = pair("M:1:FC:1:1101:1000:2000", 100, 300, 40) + pair("M:1:FC:1:1102:1000:2000", 100, 300, 30)
    assert run(orc, recs)[0].counters[1]["read_pair_optical_duplicates"] == 0


def test_qname_formats(orc):
    # 5 columns: tile,x,y = fields 2,3,4 (:62-65)
    recs = pair("FC:1:7:100:200", 100, 300, 40) + pair("FC:1:7:110:210", 100, 300, 30)
    assert run(orc, recs)[0].counters[1]["read_pair_optical_duplicates"] == 1
    # any other column count: no tile info, never an optical duplicate (:66-71, :85)
    recs = pair("a:7:100:200", 100, 300, 40) + pair("b:7:100:200", 100, 300, 30)
    m, _ = run(orc, recs)
    assert m.counters[1]["read_pair_duplicates"] == 1 and m.counters[1]["read_pair_optical_duplicates"] == 0
    # a field that strconv.ParseInt rejects panics -- but only when the list has at least two reads (:339-341)
    with pytest.raises(ValueError):
        run(orc, pair("FC:1:7:100:2x0", 100, 300, 40) + pair("FC:1:7:110:210", 100, 300, 30))
    m, _ = run(orc, pair("FC:1:7:100:2x0", 100, 300, 40) + pair("FC:1:7:110:210", 500, 700, 30))
    assert m.hist[1][0] == {1: 2}
    # signs are accepted by ParseInt
    recs = pair("FC:1:+7:-5:+20", 100, 300, 40) + pair("FC:1:7:5:20", 100, 300, 30)
    assert run(orc, recs)[0].counters[1]["read_pair_optical_duplicates"] == 1


def _brute(tiles, dist):
    n = len(tiles)
    par = list(range(n))
    def find(x):
        while par[x] != x:
            x = par[x]
        return x
    for i in range(n):
        for j in range(i + 1, n):
            (t1, x1, y1), (t2, x2, y2) = tiles[i], tiles[j]
            if t1 == t2 and abs(x1 - x2) <= dist and abs(y1 - y2) <= dist:
                par[find(i)] = find(j)
    return n - len({find(i) for i in range(n)})


def test_graph_path_against_brute_force(orc):
    # lists of >= 4 reads take countOpticalDuplicatesWithGraph (:244-273): Σ(cluster size - 1), transitive chains included
    rng = np.random.default_rng(11)
    for trial in range(20):
        k = int(rng.integers(4, 40))
        tiles = [(int(rng.integers(1, 3)), int(rng.integers(0, 400)), int(rng.integers(0, 400))) for _ in range(k)]
        recs = []
        for i, (t, x, y) in enumerate(tiles):
            recs += pair(f"M:{i}:FC:1:{t}:{x}:{y}", 100, 300, 40 - (i > 0))
        m, _ = run(orc, recs, h=_hdr(one_lib=True))
        exp = _brute(tiles, 100)
        assert m.counters[1]["read_pair_optical_duplicates"] == exp
        assert m.hist[1][0] == {k: 1}
        assert m.hist[1][1] == ({k - exp: 1} if k - exp > 0 else {})
        assert m.hist[1][2] == ({exp + 1: 1} if exp else {})


def test_library_size_and_report(orc, tmp_path):
    # estimateLibrarySize (:533-562): the root of c/x - 1 + exp(-n/x) in x, by bisection, truncated
    import math
    recs = []
    for i in range(40):
        recs += pair(f"M:{i}:FC:1:1:{i * 1000}:5", 100 + (i % 25), 300, 40 - (i >= 25))    # 25 distinct positions, 15 duplicates
    p = str(tmp_path / "m.txt")
    m, _ = run(orc, recs, h=_hdr(one_lib=True), metrics_path=p, command_line="elprep filter in out", started_on="T")
    c = m.counters[1]
    assert (c["read_pairs_examined"], c["read_pair_duplicates"], c["read_pair_optical_duplicates"]) == (40, 15, 0)
    n, u = 40.0, 25.0
    size = m.library_size[1]
    f = lambda x: u / x - 1 + math.exp(-n / x)
    assert f(size) >= -1e-3 and f(size + 1) <= 1e-3 and size > 0
    assert m.roi[1][0] == pytest.approx(size * (1 - math.exp(-n / size)) / u)
    txt = open(p).read().split("\n")
    assert txt[0] == "## htsjdk.samtools.metrics.StringHeader" and txt[1] == "# elprep filter in out" and txt[3] == "# Started on: T"
    assert txt[7].split("\t")[0] == "Unknown Library" and txt[7].split("\t")[8] == "NaN" and len(txt[7].split("\t")) == 9
    assert txt[8].split("\t") == ["libA", "0", "40", "0", "0", "0", "15", "0", "0.375", str(size)]
    assert txt[10] == "## HISTOGRAM\tjava.lang.Double"
    # BIN 1: 10 origins without duplicates; BIN 2: 15 origins with one (all_sets, optical_sets, non_optical_sets)
    assert txt[12].split("\t")[2:] == ["10", "0", "10"] and txt[13].split("\t")[2:] == ["15", "0", "15"]
    assert txt[12].split("\t")[0] == "1.0"


def test_preset_duplicate_flags(orc):
    # FLAG 0x400 already set on input stays set (FLAG |= ...); the best pair with both mates flagged counts as a
    # ReadPairDuplicate (:189) but is not attached to itself (:215)
    recs = pair("M:1:FC:1:1:10:10", 100, 300, 40, flags=(99 | 0x400, 147 | 0x400)) + pair("M:2:FC:1:1:20:20", 100, 300, 30)
    m, b = run(orc, recs)
    assert m.counters[1]["read_pair_duplicates"] == 2 and m.counters[1]["read_pair_optical_duplicates"] == 1
    assert m.hist[1][0] == {2: 1}
