"""The C-ABI library loads without a GPU and exports every symbol include/elprep_b200.h declares;
without a CUDA device elp_create fails loudly (there is no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_match_header():
    from elprep_b200 import _lib
    L = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "elprep_b200.h")).read()
    declared = set(re.findall(r"\b(elp_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.EXPORTS)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from elprep_b200 import device, sam
    with pytest.raises(device.ElprepError) as ei:
        device.Context(sam.Header(sq=[{"SN": "c", "LN": 10}]))
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)


def test_product_does_not_import_oracle():
    """the oracle is test infrastructure: nothing under elprep_b200/ or include/ may reference it"""
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "elprep_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                s = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"^\s*(import|from)\s+oracle\b", s, re.M) or "liboracle" in s or "oracle/" in s.replace("the oracle/", ""):
                    bad.append(f)
    assert not bad, bad


def _build_c_client(tmp_path):
    import subprocess
    exe = str(tmp_path / "cabi_client")
    libdir = os.path.join(ROOT, "elprep_b200", "lib")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "cabi_client.c"),
                           "-o", exe, "-L", libdir, "-lelprep_b200", "-Wl,-rpath," + libdir])
    return exe


def test_c_client_compiles_and_fails_loudly_without_gpu(tmp_path):
    """the header is valid C99 and a plain C program links against the library; without a GPU elp_create returns ELP_ENODEVICE"""
    import subprocess
    import torch
    exe = _build_c_client(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    if not torch.cuda.is_available():
        assert out.stdout.startswith("nodevice:") and "no CPU fallback" in out.stdout


@pytest.mark.gpu
def test_c_client_runs_the_path(tmp_path):
    import subprocess
    exe = _build_c_client(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.strip().split("\n")
    # coordinate order: POS 20 (read 2), the two reads at POS 50 by QNAME (r1 < r2), the unmapped read last; r2 (score 80 < 120) is the duplicate
    assert lines[0] == "order 2 0 1 3 flags 16 0 1024 4", out.stdout
    assert lines[1] == "libA unpaired 3 dups 1 unmapped 1"
    assert lines[2] == "apply-before-finalize rc -16"


def test_header_is_valid_c_and_cxx(tmp_path):
    """include/elprep_b200.h compiles on its own as C99 and as C++17 (extern "C" guards), with all warnings on"""
    import subprocess
    inc = os.path.join(ROOT, "include")
    (tmp_path / "t.c").write_text('#include "elprep_b200.h"\nint main(void) { elp_config c; elp_batch b; elp_dup_metrics m; (void)c; (void)b; (void)m; return ELP_OK; }\n')
    (tmp_path / "t.cpp").write_text('#include "elprep_b200.h"\nint main() { elp_ctx* x = nullptr; return elp_n_reads(x) == 0 ? 0 : 1; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", inc, "-c", str(tmp_path / "t.c"), "-o", str(tmp_path / "t.o")])
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", inc, "-c", str(tmp_path / "t.cpp"), "-o", str(tmp_path / "t2.o")])
