"""The hard-shadow product in the reference tree's association (csrc/cuda/shadow_assoc.h), checked on the CPU from the same
source the kernels compile: on random binary trees with random blocking leaves the stack reduction over path codes gives
the bits of BVHStrategy::ShadowHelper's recursion (`l * r`, a non-blocking leaf being an exact 1)."""
import json
import subprocess
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent


def _build_checker(tmp_path):
    exe = tmp_path / "shadow_assoc_check"
    cuda_inc = next((p for p in (Path("/usr/local/cuda/include"), Path("/usr/local/cuda/targets/x86_64-linux/include"))
                     if p.exists()), None)
    inc = [f"-I{cuda_inc}"] if cuda_inc else []
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", *inc,
                    str(REPO / "tests" / "shadow_assoc_check.cpp"), "-o", str(exe)], check=True)
    return exe


def test_path_code_reduction_equals_the_tree_recursion(tmp_path):
    exe = _build_checker(tmp_path)
    for seed in (1, 2, 3):
        out = subprocess.run([str(exe), "60000", str(seed)], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stdout + out.stderr
        r = json.loads(out.stdout.strip().splitlines()[-1])
        assert r["mismatch"] == 0 and r["cases"] > 50000
        assert r["order_matters"] > 0.2 * r["cases"]        # the visit-order product differs often: the check is not vacuous


@pytest.mark.parametrize("name,bunny", [("water_small", True), ("spheres", False)])
def test_scene_tree_path_codes_give_the_recursion_product(tmp_path, workdir, name, bunny):
    """The same on the scenes' own flattened trees (the reference's topology, BVH.hpp:49-125), with the path codes made by
    wrt_make_path_codes — the function wrt_upload_scene stages them with: for clusters of 3..12 blocking primitives the
    sort by primitive index + stack reduction gives the bits of ShadowHelper's recursion over that tree."""
    from conftest import load_golden_scene
    scene, _ = load_golden_scene(workdir, name)
    nodes = scene.nodes()
    assert nodes.shape[0] == 2 * scene.n_prims
    f = tmp_path / "tree.bin"
    nodes.tofile(f)
    exe = _build_checker(tmp_path)
    out = subprocess.run([str(exe), "tree", str(f), "40000", "11"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = [json.loads(x) for x in out.stdout.strip().splitlines()]
    assert lines[0]["prims"] == scene.n_prims and 1 <= lines[0]["depth"] <= 64
    assert lines[-1]["mismatch"] == 0 and lines[-1]["order_matters"] > 0.05 * lines[-1]["cases"]
