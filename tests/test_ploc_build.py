"""The device BVH build (csrc/cuda/ploc_bvh.h + bvh_build.cuh) on the CPU: tests/ploc_check.cpp compiles the SAME per-item
steps the kernels run, emulates the passes in barrier order and checks the resulting tree — every primitive in exactly one
leaf, leaf boxes == the reference's per-primitive boxes, every inner box (and dilated box) the exact union of its
children, sibling pairs adjacent, 2n records, depth, and a surface-area cost close to the host's binned-SAH build
(the tree only has to be valid for parity — DESIGN.md section 4 — but its cost is what the traversal kernels pay).
The GPU tests then run every parity case on the tree the kernels really build."""
import json
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import REPO
from whittedstyle_raytracer_b200 import fixtures

PKG = REPO / "whittedstyle_raytracer_b200"


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    exe = tmp_path_factory.mktemp("ploc") / "ploc_check"
    cuda_inc = next((p for p in (Path("/usr/local/cuda/include"), Path("/usr/local/cuda/targets/x86_64-linux/include"))
                     if (p / "vector_types.h").exists()), None)
    if cuda_inc is None:
        pytest.skip("CUDA headers (vector_types.h) not found")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", f"-I{cuda_inc}",
                    str(REPO / "tests" / "ploc_check.cpp"), "-o", str(exe), f"-L{PKG}", "-lwrt_host",
                    f"-Wl,-rpath,{PKG}", "-pthread"], check=True)
    return exe


def _soup(n, seed):
    rng = np.random.default_rng(seed)
    c = rng.uniform(-8, 8, (n, 3))
    lines = ["imsize 8 8", "eye 0 1 10", "viewdir 0 -0.1 -1", "hfov 60", "updir 0 1 0", "bkgcolor 0.2 0.3 0.5 1.0",
             "light 5 20 10 1 1 1 1", "mtlcolor 0.7 0.6 0.5 1 1 1 0.2 0.7 0.3 20 1 1"]
    tri = c[:, None, :] + rng.uniform(-0.12, 0.12, (n, 3, 3))
    lines += ["v %.4f %.4f %.4f" % tuple(q) for q in tri.reshape(-1, 3)]
    lines += ["f %d %d %d" % (3 * i + 1, 3 * i + 2, 3 * i + 3) for i in range(n)]
    return "\n".join(lines) + "\n"


CASES = {
    # name: (config text, with bunny, max cost ratio vs the host's binned SAH)
    "bunny": (lambda: fixtures.water_bunny_tex_config(8, 8), True, 1.05),
    "spheres_1k": (lambda: fixtures.f4_spheres_config(8, 8), False, 1.30),
    "soup_20k": (lambda: _soup(20000, 3), False, 1.30),
    "identical_boxes": (lambda: "imsize 4 4\neye 0 0 5\nviewdir 0 0 -1\nhfov 60\nupdir 0 1 0\nbkgcolor 0 0 0 1\n"
                                "mtlcolor 1 1 1 1 1 1 0.2 0.7 0.3 20 1 1\n" + "sphere 0 0 0 1\n" * 37, False, 2.0),
    "two": (lambda: "imsize 4 4\neye 0 0 5\nviewdir 0 0 -1\nhfov 60\nupdir 0 1 0\nbkgcolor 0 0 0 1\n"
                    "mtlcolor 1 1 1 1 1 1 0.2 0.7 0.3 20 1 1\nsphere 0 0 0 1\nsphere 3 0 0 1\n", False, 1.01),
    "one": (lambda: "imsize 4 4\neye 0 0 5\nviewdir 0 0 -1\nhfov 60\nupdir 0 1 0\nbkgcolor 0 0 0 1\n"
                    "mtlcolor 1 1 1 1 1 1 0.2 0.7 0.3 20 1 1\nsphere 0 0 0 1\n", False, 1.01),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_ploc_tree_is_a_valid_tree_over_the_reference_boxes(checker, tmp_path, name):
    text, bunny, max_ratio = CASES[name]
    fixtures.ensure_assets(tmp_path)
    fixtures.write_config(tmp_path, name, text())
    p = subprocess.run([str(checker), str(tmp_path / f"{name}.txt"), str(tmp_path / "bunny.obj") if bunny else "", str(tmp_path)],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    r = json.loads(p.stdout.strip().splitlines()[-1])
    assert "error" not in r, r
    assert r["records"] == max(2 * r["n"], 2) == r["ref_nodes"]
    assert r["missing"] == 0 and r["bad_union"] == 0 and r["bad_leaf_box"] == 0
    assert r["bad_dilated_union"] == 0 and r["dilated_leaf_diff"] == 0
    assert r["depth"] == r["depth_seen"] <= 64
    assert r["passes"] <= 4 * max(1, int(np.ceil(np.log2(max(r["n"], 2))))) + 8
    assert r["cost"] <= max_ratio * r["host_sah_cost"] + 1e-6, r
