"""The soft-shadow shaft test (csrc/cuda/shaft_cull.h) compiled for the CPU from the SAME source the
kernel uses, checked by brute force (tests/shaft_cull_check.cpp): an "empty shaft" verdict must mean that
no sample ray, built with the kernel's float operations, hits the own box of any primitive, and every
sample's 1/d must lie inside the shaft's bounds.  The GPU parity tests then pin the kernel end to end
(soft-shadow frames are bit-compared with the oracle, which traces all 50 samples of every request)."""
import json
import os
import subprocess
from pathlib import Path

import pytest

from conftest import REPO
from whittedstyle_raytracer_b200 import fixtures

PKG = REPO / "whittedstyle_raytracer_b200"


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    exe = tmp_path_factory.mktemp("shaft") / "shaft_cull_check"
    cuda_inc = next((p for p in (Path("/usr/local/cuda/include"), Path("/usr/local/cuda/targets/x86_64-linux/include"))
                     if (p / "vector_types.h").exists()), None)
    if cuda_inc is None:
        pytest.skip("CUDA headers (vector_types.h) not found")
    extra = os.environ.get("WRT_TEST_CXXFLAGS", "").split()       # A/B builds of the header's compile-time variants
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", *extra, f"-I{cuda_inc}",
                    str(REPO / "tests" / "shaft_cull_check.cpp"), "-o", str(exe), f"-L{PKG}", "-lwrt_host",
                    f"-Wl,-rpath,{PKG}", "-pthread"], check=True)
    return exe


def _soft(text):
    return text if "shadow soft" in text else text.replace("\nlight ", "\nshadow soft\nlight ", 1)


CASES = {
    # name: (config text, with bunny, requests)
    "water_bunny_tex": (lambda: fixtures.water_bunny_tex_config(64, 48, soft=True), True, 4000),
    "bunny_shadow": (lambda: fixtures.bunny_shadow_config(64, 48, soft=True), True, 2000),
    "spheres": (lambda: _soft(fixtures.spheres_config(64, 48)), False, 6000),
    "smooth": (lambda: _soft(fixtures.smooth_config(64, 48)), False, 6000),
    # a light right above a dense little scene: origins nearly under the light (sign-indefinite shafts),
    # a light very close to the geometry, a light far away on an axis
    "near_light": (lambda: fixtures._CAMERA.format(w=32, h=24) + "\nshadow soft\nlight 0 1.5 -1 1 1 1 1\n"
                   "light 300 0.01 0.01 1 1 1 1\nlight 0.5 0.2 -0.5 1 1 1 1\n" + fixtures._WALL_VERTS +
                   "\nmtlcolor 1 1 1 1 1 1 0.2 0.6 0.2 10 1 1\nf 1 2 3\nf 1 4 2\nsphere 0 0 -1 0.4\nsphere 1 1 -2 0.3\n",
                   True, 3000),
    # lights whose corners have a 0 coordinate: origins that share it make every sample axis-degenerate there
    "axis_degenerate": (lambda: fixtures._CAMERA.format(w=32, h=24) + "\nshadow soft\nlight 0 0 60 1 1 1 1\n"
                        "light 50 0 0 1 1 1 1\nv -1 -2 -4\nv 1 -2 -4\nv 0 -2 -6\n"
                        "mtlcolor 1 1 1 1 1 1 0.2 0.6 0.2 10 1 1\nf 1 2 3\nsphere 0 0 -1 0.4\nsphere 1 0 2 0.3\n"
                        "sphere -2 0.5 1 0.5\nsphere 0 -1 3 0.25\nsphere 4 0 -3 0.6\nsphere -3 2 -2 0.5\n", False, 8000),
}


@pytest.mark.parametrize("name", list(CASES))
def test_empty_shaft_verdicts_hold_by_brute_force(name, checker, workdir):
    text, bunny, n = CASES[name]
    fixtures.write_config(workdir, f"shaft_{name}", text())
    obj = str(workdir / "bunny.obj") if bunny else "-"
    out = subprocess.run([str(checker), str(workdir / f"shaft_{name}.txt"), obj, str(workdir), str(n), "7"],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    r = json.loads(out.stdout.strip().splitlines()[-1])
    assert r["violations"] == 0 and r["bound_violations"] == 0 and r["list_violations"] == 0
    assert r["filter_violations"] == 0                       # no pruned candidate is ever hit by a sample ray
    assert r["wide_mismatch"] == 0                           # the 4-wide walks (wide_bvh.h) reach exactly the same leaves
    if name in ("water_bunny_tex", "bunny_shadow"):
        assert r["filter_removed"] > 0.2 * (r["filter_removed"] + r["filter_kept"]) and r["filter_pairs"] > 100000
    assert r["list_rays"] > 0
    assert r["empty"] + r["nonempty"] + r["gave_up"] > 0
    if name in ("water_bunny_tex", "bunny_shadow"):
        assert r["empty"] > 0 and r["nonempty"] > 0          # the test exercises both verdicts
    if name == "axis_degenerate":
        assert r["degenerate_rays"] > 1000 and r["empty"] > 0
