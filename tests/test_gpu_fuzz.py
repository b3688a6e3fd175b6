"""Randomised scenes (seeded): spheres + triangles, opaque / glass / mirror materials,
point / attenuated / directional lights, textures, normal maps, depth cueing, both
projections and both shadow modes — CUDA image and ray counts against the CPU oracle."""
import numpy as np
import pytest

import oracle_bindings as ob
from conftest import image_diff
from whittedstyle_raytracer_b200 import Renderer, Scene
from whittedstyle_raytracer_b200.renderer import TRAVERSAL_EXHAUSTIVE, TRAVERSAL_PRUNED

pytestmark = pytest.mark.gpu


def fmt(x):
    return f"{x:.4f}".rstrip("0").rstrip(".") if abs(x) >= 1e-4 else "0"


def random_scene(rng: np.random.Generator) -> str:
    w, h = int(rng.integers(40, 120)), int(rng.integers(30, 90))
    lines = [f"imsize {w} {h}", "eye 0 1 8", f"viewdir {fmt(rng.uniform(-0.2, 0.2))} {fmt(rng.uniform(-0.3, 0.1))} -1",
             f"hfov {int(rng.integers(40, 90))}", "updir 0 1 0", f"bkgcolor {fmt(rng.random())} {fmt(rng.random())} {fmt(rng.random())} 1.0"]
    if rng.random() < 0.25:
        lines.append("projection parallel")
    if rng.random() < 0.35:
        lines.append("shadow soft")
    if rng.random() < 0.3:
        lines.append("depthcueing 0.2 0.2 0.3 1.0 0.2 25 4")
    for _ in range(int(rng.integers(1, 3))):
        p = rng.uniform(-8, 8, 3)
        p[1] = abs(p[1]) + 3
        kind = rng.random()
        if kind < 0.6:
            lines.append(f"light {fmt(p[0])} {fmt(p[1])} {fmt(p[2])} 1 {fmt(rng.uniform(.4, 1))} {fmt(rng.uniform(.4, 1))} {fmt(rng.uniform(.4, 1))}")
        elif kind < 0.8:
            lines.append(f"attlight {fmt(p[0])} {fmt(p[1])} {fmt(p[2])} 1 1 1 1 0.5 0.02 0.002")
        else:
            lines.append(f"light {fmt(rng.uniform(-1, 1))} -1 {fmt(rng.uniform(-1, 0))} 0 0.8 0.8 0.7")

    def material():
        kind = rng.random()
        od = [fmt(x) for x in rng.uniform(.1, 1, 3)]
        if kind < 0.45:      # opaque, maybe a little mirror
            ks, alpha, eta = rng.choice([0, 0, 0.3]), 1, 1
        elif kind < 0.8:     # glass
            ks, alpha, eta = rng.uniform(.1, .4), rng.uniform(.1, .5), rng.uniform(1.1, 1.7)
        else:                # mirror
            ks, alpha, eta = rng.uniform(.5, .9), 1, 1
        return (f"mtlcolor {od[0]} {od[1]} {od[2]} 1 1 1 {fmt(rng.uniform(.05, .3))} {fmt(rng.uniform(.3, .8))} "
                f"{fmt(ks)} {int(rng.integers(4, 80))} {fmt(alpha)} {fmt(eta)}")

    nv = 0
    lines += ["vt 0 0", "vt 1 0", "vt 1 1", "vt 0 1"]
    for _ in range(int(rng.integers(2, 9))):
        lines.append(material())
        if rng.random() < 0.35:
            lines.append("texture textures/harbor.ppm")
            if rng.random() < 0.5:
                lines.append("bump textures/bumps.ppm")
        if rng.random() < 0.5:
            c = rng.uniform(-4, 4, 3)
            c[2] -= 2
            lines.append(f"sphere {fmt(c[0])} {fmt(c[1])} {fmt(c[2])} {fmt(rng.uniform(.4, 1.8))}")
        else:
            base = rng.uniform(-5, 5, 3)
            base[2] -= 3
            pts = [base + rng.uniform(-3, 3, 3) for _ in range(3)]
            for p in pts:
                lines.append(f"v {fmt(p[0])} {fmt(p[1])} {fmt(p[2])}")
            a, b, c = nv + 1, nv + 2, nv + 3
            nv += 3
            lines.append(f"f {a}/1 {b}/2 {c}/3" if rng.random() < 0.5 else f"f {a} {b} {c}")
    lines.append("mtlcolor 0.6 0.6 0.6 1 1 1 0.2 0.7 0 10 1 1")
    lines += ["v -14 -2.5 8", "v 14 -2.5 8", "v 14 -2.5 -16", "v -14 -2.5 -16"]
    lines += [f"f {nv + 1} {nv + 2} {nv + 3}", f"f {nv + 1} {nv + 3} {nv + 4}"]
    return "\n".join(lines) + "\n"


@pytest.mark.parametrize("seed", range(24))
def test_random_scene_matches_oracle(workdir, seed):
    rng = np.random.default_rng(1000 + seed)
    text = random_scene(rng)
    scene = Scene(text=text, asset_dir=workdir)
    ref, ost = ob.OracleScene(scene).render()
    r = Renderer(scene)
    for traversal in (TRAVERSAL_PRUNED, TRAVERSAL_EXHAUSTIVE):
        r.ctx.set_options(traversal=traversal)
        img = r.render()
        d = image_diff(img, ref)
        assert d["within1"] >= 0.999 * d["n"] and d["exact"] >= 0.995 * d["n"], (seed, traversal, d, text)
        st = r.last_stats
        assert st["closest_rays"] == ost.closest_rays and st["shadow_rays"] == ost.shadow_rays, (seed, text)
    r.ctx.close()
