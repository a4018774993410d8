"""csrc/env_core.cuh::div_by_const (the height-field index division without a divide) is bit-identical to the IEEE
division for every fp32 numerator in [2^-31, 2^24), both signs -- exhaustive, compiled C with real fmaf."""
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))


def test_div_by_const_is_exact_for_every_fp32_numerator():
    exe = os.path.join(tempfile.mkdtemp(), "div_const_check")
    subprocess.check_call(["gcc", "-O2", "-mfma", "-ffp-contract=off", os.path.join(HERE, "csrc", "div_const_check.c"), "-o", exe, "-lm"])
    out = subprocess.run([exe, "0.1", "0.05", "0.25"], capture_output=True, text=True, check=True).stdout.split("\n")
    rows = [l.split() for l in out if l.strip()]
    assert len(rows) == 3 and all(int(r[1]) == 0 for r in rows), rows
