"""The exact-division forms of the TSDF integrate kernel (gps_slam_b200/csrc/tsdf_kernels.cu: div_255, div_32767, div_int), checked
on the CPU: the constants / reciprocal table in the CUDA source are RN(1/c), and q' = fma(fma(-q, c, x), r, q) equals the IEEE quotient
on a sampled range (the exhaustive runs over all 2^32 inputs are tools/div_proof/, ~2 min; results in its README)."""
import os
import re
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "gps_slam_b200", "csrc", "tsdf_kernels.cu")).read()


def test_reciprocal_constants_are_correctly_rounded():
    table = re.search(r"c_rcpInt\[257\] = \{(.*?)\};", SRC, re.S).group(1)
    vals = [float.fromhex(t.strip().rstrip("f")) for t in table.split(",")]
    assert len(vals) == 257
    for w in range(1, 257):
        assert np.float32(vals[w]) == np.float32(1) / np.float32(w), w
    for const, lit in ((255.0, "0x1.010102p-8f"), (32767.0, "0x1.0002p-15f")):
        assert lit in SRC
        assert np.float32(float.fromhex(lit.rstrip("f"))) == np.float32(1) / np.float32(const)


def test_fma_quotient_equals_ieee_division_on_a_sample(tmp_path):
    exe = str(tmp_path / "proof_const")
    subprocess.check_call(["gcc", "-O2", "-mfma", "-fopenmp", "-ffp-contract=off", os.path.join(ROOT, "tools", "div_proof", "proof_const.c"),
                           "-o", exe, "-lm"])
    for c in ("255", "32767", "3", "7", "101", "255.0"):
        # all floats in [0.5, 4) and their neighbourhood of exponents: 3 x 2^23 values
        out = subprocess.run([exe, c, "0x3f000000", "0x40800000"], capture_output=True, text=True, check=True).stdout
        assert "mismatches: 0 " in out, out
