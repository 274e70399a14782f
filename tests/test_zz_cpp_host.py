"""The C++ host layer above the C ABI (include/elprep_b200.hpp: elPrep's types and operator names for compiled callers).
CPU: it compiles with all warnings on, its marshaller builds the same columns as the Python one (sam.AlignmentBatch.from_records), and
without a GPU the constructor throws with ELP_ENODEVICE (no CPU fallback).  GPU: the four-read scenario of the plain-C client through
DeviceSam (AddNodes / Finalize / RunPipeline / MarkOpticalDuplicates)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "cpp_host_client")
    libdir = os.path.join(ROOT, "elprep_b200", "lib")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "cpp_host_client.cpp"),
                           "-o", exe, "-L", libdir, "-lelprep_b200", "-Wl,-rpath," + libdir])
    return exe


def test_cpp_marshal_equals_python_marshal(tmp_path):
    from elprep_b200 import sam
    exe = _build(tmp_path)
    out = subprocess.run([exe, "marshal"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stdout + out.stderr
    cols = {}
    for ln in out.stdout.strip().split("\n"):
        name, *vals = ln.split(" ")
        cols[name] = np.array([int(v) for v in vals], dtype=np.uint64)
    header = sam.Header(sq=[{"SN": "chr1", "LN": 100000}, {"SN": "chr2", "LN": 50000}], rg=[{"ID": "rg1", "LB": "libA"}, {"ID": "rg2"}])
    recs = [dict(QNAME="readA", FLAG=99, RNAME="chr1", POS=100, MAPQ=60, CIGAR="3S2M3M1I4M", RNEXT="=", PNEXT=250, TLEN=160, SEQ="ACGTNacgtRYKM",
                 QUAL=[30, 31, 32, 33, 2, 2, 20, 21, 22, 23, 24, 25, 26], RG="rg2"),
            dict(QNAME="b", FLAG=147, RNAME="chr2", POS=7, MAPQ=0, CIGAR="5M", RNEXT="chr1", PNEXT=9, TLEN=-3, SEQ="TTTTT", QUAL=[40] * 5, RG="rg1"),
            dict(QNAME="", FLAG=4, RNAME="*", POS=0, MAPQ=0, CIGAR="*", RNEXT="*", PNEXT=0, TLEN=0, SEQ="", QUAL=[], RG=None),
            dict(QNAME="weird", FLAG=0, RNAME="chrUn", POS=5, MAPQ=3, CIGAR="2H1=1X2D1N1P", RNEXT="chrUn", PNEXT=1, TLEN=0, SEQ="G*", QUAL=[1, 93], RG="rg1")]
    b = sam.AlignmentBatch.from_records(header, recs)
    u = lambda a: np.asarray(a).astype(np.int64).astype(np.uint64)       # the client prints signed values as their 64-bit two's complement
    for name, ref in (("refid", b.refid), ("nref", b.nref), ("pos", b.pos), ("pnext", b.pnext), ("tlen", b.tlen), ("rg", b.rg), ("lseq", b.lseq), ("flag", b.flag),
                      ("mapq", b.mapq), ("qname", b.qname), ("qname_off", b.qname_off), ("cigar", b.cigar), ("cigar_off", b.cigar_off), ("seq", b.seq), ("qual", b.qual)):
        assert np.array_equal(cols[name], u(ref)), name
    assert cols["opt"].tolist() == [0, 1, 0, 0]


def test_cpp_host_fails_loudly_without_gpu(tmp_path):
    import torch
    exe = _build(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    if not torch.cuda.is_available():
        assert out.stdout.startswith("nodevice:") and "no CPU fallback" in out.stdout


@pytest.mark.gpu
def test_cpp_host_runs_the_path(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.strip().split("\n")
    assert lines[0] == "order r3 r1 r2 r4 flags 16 0 1024 4", out.stdout     # the same reads and answer as tests/c/cabi_client.c
    assert lines[1] == "libA unpaired 3 dups 1 unmapped 1"
    assert lines[2] == "apply-before-finalize rc -16"
