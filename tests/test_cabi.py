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
