"""The hard-shadow product in the reference tree's association (csrc/cuda/shadow_assoc.h), checked on the CPU from the same
source the kernels compile: on random binary trees with random blocking leaves the stack reduction over path codes gives
the bits of BVHStrategy::ShadowHelper's recursion (`l * r`, a non-blocking leaf being an exact 1)."""
import json
import subprocess
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent


def test_path_code_reduction_equals_the_tree_recursion(tmp_path):
    exe = tmp_path / "shadow_assoc_check"
    cuda_inc = next((p for p in (Path("/usr/local/cuda/include"), Path("/usr/local/cuda/targets/x86_64-linux/include"))
                     if p.exists()), None)
    inc = [f"-I{cuda_inc}"] if cuda_inc else []
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", *inc,
                    str(REPO / "tests" / "shadow_assoc_check.cpp"), "-o", str(exe)], check=True)
    for seed in (1, 2, 3):
        out = subprocess.run([str(exe), "60000", str(seed)], capture_output=True, text=True)
        assert out.returncode == 0, out.stdout + out.stderr
        r = json.loads(out.stdout.strip().splitlines()[-1])
        assert r["mismatch"] == 0 and r["cases"] > 50000
        assert r["order_matters"] > 0.2 * r["cases"]        # the visit-order product differs often: the check is not vacuous
