"""The closest-hit pruning rule (csrc/cuda/prune_rule.h) compiled for the CPU from the SAME source the kernels use and
checked by brute force (tests/prune_rule_check.cpp): for adversarial rays — targets on and just outside triangle edges
and vertices (Triangle.hpp:41 accepts barycentrics down to -1e-5), elevations down to 1e-4 rad over the triangle's
plane, box faces seen edge-on, origins on the plane, the 40-unit wall triangles — the own box of the TRUE closest
primitive (every primitive tested with the reference's arithmetic) is never entered beyond the limit derived from its
own t, so no visiting order of the pruned walk can skip it.  Round 1's margin t*(1+1e-3)+1e-3 fails this check (and
failed on the GPU: ADVICE r1, test_adversarial_rays_pruned_equals_exhaustive)."""
import json
import subprocess

import pytest

from conftest import REPO
from whittedstyle_raytracer_b200 import fixtures

PKG = REPO / "whittedstyle_raytracer_b200"


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    exe = tmp_path_factory.mktemp("prune") / "prune_rule_check"
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", str(REPO / "tests" / "prune_rule_check.cpp"),
                    "-o", str(exe), f"-L{PKG}", "-lwrt_host", f"-Wl,-rpath,{PKG}", "-pthread"], check=True)
    return exe


CASES = {
    "water_bunny_tex": (lambda: fixtures.water_bunny_tex_config(64, 48), True, 60000),
    "bunny_shadow": (lambda: fixtures.bunny_shadow_config(64, 48), True, 40000),
    "bump": (lambda: fixtures.bump_config(64, 48), False, 300000),
    "smooth": (lambda: fixtures.smooth_config(64, 48), False, 300000),
}


@pytest.mark.parametrize("name", list(CASES))
def test_true_closest_hit_is_never_behind_the_prune_limit(name, checker, workdir):
    text, bunny, n = CASES[name]
    fixtures.write_config(workdir, f"prune_{name}", text())
    obj = str(workdir / "bunny.obj") if bunny else "-"
    out = subprocess.run([str(checker), str(workdir / f"prune_{name}.txt"), obj, str(workdir), str(n), "11"],
                         capture_output=True, text=True)
    r = json.loads(out.stdout.strip().splitlines()[-1])
    assert out.returncode == 0 and r["violations"] == 0, out.stdout + out.stderr
    assert r["hits"] > 0.5 * n
    if bunny:
        assert r["entry_after_hit"] > 0          # the set-up does produce hits in front of their own box
        assert r["old_rule_violations"] > 0      # ... and beyond round 1's margin
        assert r["worst_gap_over_margin"] < 0.5
