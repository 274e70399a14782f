"""CPU check of the word-parallel lane arithmetic of the BQSR count kernel: tests/c/lane_check.cpp compiles the kernel's own header
(elprep_b200/csrc/bqsr_lane.cuh) as host C++ and compares it base by base with a direct restatement of the per-base rules of
(*BaseRecalibrator).Recalibrate (filters/bqsr.go:467-551) on random reads, windows cut out of byte arenas as on the device."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lane_arithmetic_matches_per_base_rules(tmp_path):
    exe = str(tmp_path / "lane_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "elprep_b200", "csrc"), "-o", exe, os.path.join(ROOT, "tests", "c", "lane_check.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout[-2000:]


def test_closed_form_clipping_matches_oracle(tmp_path, orc):
    """lanes::closed_form_clip (what bqsr_prep2_kernel applies to [H][S]M[S][H] and one-indel reads) against the oracle's step-by-step
    hardClipAdaptorSequence + hardClipSoftClippedBases and getReadCoordinateForReferenceCoordinate on 400 k random reads"""
    import oracle
    so = oracle.build()
    exe = str(tmp_path / "clip_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "elprep_b200", "csrc"), "-o", exe, os.path.join(ROOT, "tests", "c", "clip_check.cpp"),
                           "-L", os.path.dirname(so), "-loracle", "-Wl,-rpath," + os.path.dirname(so)])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout[-2000:]
