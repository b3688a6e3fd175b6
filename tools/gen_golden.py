"""Generates tests/golden/* by running the UNMODIFIED reference (oracle/_ref, built by
`make -C oracle ref` from /root/reference) on the fixture scenes.

Run in the build container only (needs /root/reference for the harness build):
    python tools/gen_golden.py
Outputs (committed, small):
    tests/golden/img_<scene>.npz     reference 8-bit image, config text, ray counts
    tests/golden/rays_<scene>.npz    ray batches + the reference's closest hits, shadow
                                     coefficients and float traceRay colours
    tests/golden/bvh_bunny.npz       left-to-right leaf order of the reference's BVH
    tests/golden/cli_errors.json     stdout / exit status of the reference executable on bad configs
"""
from __future__ import annotations

import json
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))

import oracle_bindings as ob  # noqa: E402
from whittedstyle_raytracer_b200 import Scene, fixtures  # noqa: E402

GOLD = REPO / "tests" / "golden"

# scene name -> (config text, uses bunny)
SCENES = {
    "config_800x600": (fixtures.bunny_shadow_config(800, 600), True),
    "config_small": (fixtures.bunny_shadow_config(200, 150), True),
    "water_small": (fixtures.water_bunny_tex_config(200, 150), True),
    "spheres": (fixtures.spheres_config(), False),
    "parallel": (fixtures.parallel_config(), False),
    "bump": (fixtures.bump_config(), False),
    "directional": (fixtures.directional_config(), False),
    "smooth": (fixtures.smooth_config(), False),
}
SOFT_SCENES = {   # non-deterministic in the reference: statistical golden (mean of several renders)
    "water_soft": (fixtures.water_bunny_tex_config(120, 90, soft=True), True),
    "spheres_soft": (fixtures.spheres_config(160, 120) + "shadow soft\n", False),
}
RAY_SCENES = ["config_small", "water_small", "spheres", "bump", "directional", "smooth"]


def unit(v):
    return v / np.linalg.norm(v, axis=1, keepdims=True)


def ray_batches(scene: Scene, orc: ob.OracleScene, rng: np.random.Generator):
    """Primary, random, surface-leaving and vertex/edge-grazing rays."""
    o0, d0 = orc.primary_rays()
    pick = rng.choice(len(o0), size=min(3000, len(o0)), replace=False)
    origs, dirs = [o0[pick]], [d0[pick]]
    n = 2000
    origs.append(rng.uniform(-6, 6, (n, 3)).astype(np.float32))
    dirs.append(unit(rng.normal(size=(n, 3))).astype(np.float32))
    # rays leaving hit points (secondary-ray-like, some starting inside objects)
    h = orc.trace_closest(o0[pick], d0[pick])
    hp = h["pos"][h["hit"] == 1]
    if len(hp):
        sel = hp[rng.integers(0, len(hp), n)]
        dd = unit(rng.normal(size=(n, 3))).astype(np.float32)
        origs.append((sel + 1e-4 * dd).astype(np.float32))
        dirs.append(dd)
    # rays aimed exactly at triangle vertices / edge midpoints (shared edges -> ties)
    d = scene.desc
    if d.n_prims:
        g = np.ctypeslib.as_array(d.prim_geom, shape=(d.n_prims, 12))
        fl = np.ctypeslib.as_array(d.prim_flags, shape=(d.n_prims,))
        tri = g[(fl & 1) == 0]
        if len(tri):
            k = rng.integers(0, len(tri), n)
            v0, e1, e2 = tri[k, 0:3], tri[k, 4:7], tri[k, 8:11]
            w = rng.integers(0, 4, n)[:, None]
            target = np.where(w == 0, v0, np.where(w == 1, v0 + e1, np.where(w == 2, v0 + 0.5 * e1, v0 + 0.5 * e1 + 0.5 * e2)))
            eye = np.array([[scene.camera.eye[0], scene.camera.eye[1], scene.camera.eye[2]]], np.float32)
            eye = eye + rng.uniform(-0.5, 0.5, (n, 3)).astype(np.float32) * (rng.random((n, 1)) < 0.5)
            origs.append(eye.astype(np.float32))
            dirs.append(unit((target - eye).astype(np.float32)).astype(np.float32))
    # axis-parallel directions (zero components -> inf / NaN slab paths)
    ax = np.eye(3, dtype=np.float32)[rng.integers(0, 3, 300)] * rng.choice([-1.0, 1.0], (300, 1)).astype(np.float32)
    origs.append(rng.uniform(-3, 3, (300, 3)).astype(np.float32))
    dirs.append(ax)
    return np.concatenate(origs).astype(np.float32), np.concatenate(dirs).astype(np.float32)


def main():
    if not ob.have_reference():
        raise SystemExit("oracle/_ref/libwhitted_ref.so missing: run `make -C oracle ref` first")
    GOLD.mkdir(parents=True, exist_ok=True)
    wd = Path(tempfile.mkdtemp(prefix="wrt_gold_"))
    fixtures.ensure_assets(wd)
    rng = np.random.default_rng(20261018)

    for name, (text, bunny) in SCENES.items():
        fixtures.write_config(wd, name, text)
        ref = ob.ReferenceScene(wd, name, bunny=bunny)
        img, secs = ref.render()
        closest, shadow = ref.counters(reset=True)
        assert img.min() >= 0 and img.max() <= 255
        np.savez_compressed(GOLD / f"img_{name}.npz", rgb=img.astype(np.uint8), config=text, bunny=bunny,
                            closest_rays=closest, shadow_rays=shadow)
        print(f"{name}: {img.shape} rays {closest}+{shadow} ref {secs:.2f}s")
        if name in RAY_SCENES:
            scene = Scene.from_workdir(wd, name, bunny=bunny)
            orc = ob.OracleScene(scene)
            o, d = ray_batches(scene, orc, rng)
            hits = ref.trace_closest(o, d)
            m = hits["hit"] == 1
            pos, nd = hits["pos"][m][:4000], hits["ndir"][m][:4000]
            L = scene.desc.lights[0]
            lp = np.tile(np.array([[L.pos[0], L.pos[1], L.pos[2]]], np.float32), (len(pos), 1))
            far = rng.uniform(-30, 30, (len(pos), 3)).astype(np.float32)
            lp = np.where((np.arange(len(pos)) % 2 == 0)[:, None], lp, far).astype(np.float32)
            sh_hard = ref.shadow_hard(pos, nd, lp)
            sh_soft = ref.shadow_soft(pos, nd, lp)
            ldir = np.tile(np.array([[-1, -2, -1.5, 0]], np.float32), (len(pos), 1))
            ldir[1::2, :3] = rng.normal(size=(len(ldir[1::2]), 3)).astype(np.float32)
            sh_dir = ref.shadow_directional(pos, hits["object"][m][:4000], ldir)
            colour = ref.trace_ray(o[:3000], d[:3000])
            np.savez_compressed(GOLD / f"rays_{name}.npz", orig=o, dir=d, hits=hits, sh_pos=pos, sh_ndir=nd,
                                sh_light=lp, sh_hard=sh_hard, sh_soft=sh_soft, sh_ldir=ldir,
                                sh_self=hits["object"][m][:4000], sh_dir=sh_dir, colour=colour)
            print(f"   rays {len(o)}  hits {int(m.sum())}  shadow queries {len(pos)}")
        if name == "config_small":
            order, depth = ref.leaf_order()
            np.savez_compressed(GOLD / "bvh_bunny.npz", leaf_order=order.astype(np.int32), depth=depth)

    for name, (text, bunny) in SOFT_SCENES.items():
        fixtures.write_config(wd, name, text)
        ref = ob.ReferenceScene(wd, name, bunny=bunny)
        acc = None
        k = 6
        for _ in range(k):
            img, _ = ref.render()
            acc = img.astype(np.float64) if acc is None else acc + img
        closest, shadow = ref.counters(reset=True)
        np.savez_compressed(GOLD / f"soft_{name}.npz", mean_rgb=(acc / k).astype(np.float32), one_rgb=img.astype(np.uint8),
                            config=text, bunny=bunny, renders=k, closest_rays=closest // k, shadow_rays=shadow // k)
        print(f"{name}: soft golden from {k} reference renders, rays {closest // k}+{shadow // k}")

    # reference executable on malformed configs: message + exit status
    bad = {
        "unknown_keyword": "imsize 4 4\nfoo 1\n",
        "missing_required": "imsize 4 4\neye 0 0 0\n",
        "bad_float": "imsize 4 4\neye 0 0 x\n",
        "exponent_float": "imsize 4 4\neye 0 0 1e3\n",
        "plus_sign": "imsize 4 4\neye 0 0 +1\n",
        "leading_dot": "imsize 4 4\neye 0 0 .5\n",
        "trailing_dot": "imsize 4 4\neye 0 0 5.\n",
        "truncated_record": "imsize 4 4\neye 0 0\n",
        "same_view_up": "imsize 4 4\neye 0 0 0\nviewdir 0 1 0\nupdir 0 1 0\nhfov 60\nbkgcolor 0 0 0 1\n",
        "bad_face": "imsize 4 4\neye 0 0 0\nviewdir 0 0 -1\nupdir 0 1 0\nhfov 60\nbkgcolor 0 0 0 1\nv 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2/1 3\n",
        "face_index_oob": "imsize 4 4\neye 0 0 0\nviewdir 0 0 -1\nupdir 0 1 0\nhfov 60\nbkgcolor 0 0 0 1\nv 0 0 0\nv 1 0 0\nf 1 2 3\n",
        "missing_texture": "imsize 4 4\neye 0 0 0\nviewdir 0 0 -1\nupdir 0 1 0\nhfov 60\nbkgcolor 0 0 0 1\ntexture nope.ppm\n",
        "negative_imsize": "imsize -4 4\n",
        "last_keyword_no_newline": "imsize 4 4\neye 0 0 0\nviewdir 0 0 -1\nupdir 0 1 0\nhfov 60\nbkgcolor 0 0 0 1\nv -1 -1 -3\nv 1 -1 -3\nv 0 1 -3\nmtlcolor 0.8 0.4 0.2 1 1 1 0.3 0.6 0.2 10 1 1\nf 1 2 3\nlight 2 3 1 1 1 1 1\nshadow",
        "ok_minimal": "imsize 6 4\neye 0 0 0\nviewdir 0 0 -1\nupdir 0 1 0\nhfov 60\nbkgcolor 0.5 0.25 1 1\nv -1 -1 -3\nv 1 -1 -3\nv 0 1 -3\nmtlcolor 0.8 0.4 0.2 1 1 1 0.3 0.6 0.2 10 1 1\nf 1 2 3\nlight 2 3 1 1 1 1 1\n",
        "empty_scene_crashes_reference": "imsize 4 4\neye 0 0 0\nviewdir 0 0 -1\nupdir 0 1 0\nhfov 60\nbkgcolor 0.5 0.25 1 1\n",
    }
    out = {}
    bd = Path(tempfile.mkdtemp(prefix="wrt_bad_"))
    for k, text in bad.items():
        (bd / f"{k}.txt").write_text(text)
        p = subprocess.run([str(ob.REF_EXE), f"{k}.txt"], cwd=bd, capture_output=True, text=True, errors="replace")
        first_err = next((ln for ln in p.stdout.splitlines() if ln.startswith("ERROR")), "")
        ppm = (bd / f"{k}.ppm")
        out[k] = dict(config=text, returncode=p.returncode, error_line=first_err,
                      ppm=ppm.read_text() if ppm.exists() else None)
        print(f"cli {k}: rc={p.returncode} {first_err!r}")
    (GOLD / "cli_errors.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
