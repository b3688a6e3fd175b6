"""Generates tests/golden/assoc_<case>.npz: hard-shadow queries with MANY translucent crossings, answered by the UNMODIFIED
reference (oracle/_ref, `make -C oracle ref`; build container only) — the rays on which BVHStrategy::ShadowHelper's
tree-association product (`l * r`, BVHStrategy.hpp:43-47) differs from a left-to-right product in the last bits.

    python tools/gen_assoc_golden.py

  assoc_glass_row.npz   a row of six glass spheres of different alphas over a floor (no bunny): rays along the row, up to
                        six different factors != 1 (one per sphere)
  assoc_glass_bunny.npz the glass bunny of water_small (alpha 0.2): rays from behind it towards the light, 4, 6, ... equal factors
Each holds the config text, pos / ndir / lightpos of 20 000 queries and the reference's coefficients."""
from __future__ import annotations

import sys
import tempfile
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))

import oracle_bindings as ob  # noqa: E402
from whittedstyle_raytracer_b200 import fixtures  # noqa: E402

GOLD = REPO / "tests" / "golden"
N = 20_000


def glass_row_text() -> str:
    text = fixtures._CAMERA.format(w=160, h=120) + "light 8 0.2 -2 1 1 1 1\n"
    for k, a in enumerate([0.1, 0.25, 0.4, 0.55, 0.7, 0.85]):
        text += f"mtlcolor 0.8 0.8 0.9 1 1 1 0.2 0.6 0.3 20 {a} 1.3\nsphere {-3 + k} 0 -2 0.42\n"
    return text + ("mtlcolor 0.7 0.7 0.7 1 1 1 0.2 0.8 0.0 10 1 1\nv -12 -0.6 6\nv 12 -0.6 6\nv 12 -0.6 -14\nv -12 -0.6 -14\n"
                   "f 1 2 3\nf 1 3 4\n")


def main() -> None:
    assert ob.have_reference(), "oracle/_ref is not built (needs /root/reference)"
    wd = Path(tempfile.mkdtemp(prefix="wrt_assoc_golden_"))
    fixtures.ensure_assets(wd)
    rng = np.random.default_rng(23)
    # ---- the row of glass spheres ----
    text = glass_row_text()
    fixtures.write_config(wd, "glass_row", text)
    pos = np.stack([rng.uniform(-6, 3.5, N), rng.uniform(-0.4, 0.4, N), rng.uniform(-2.4, -1.6, N)], 1).astype(np.float32)
    nd = rng.normal(size=(N, 3)).astype(np.float32)
    nd /= np.linalg.norm(nd, axis=1, keepdims=True)
    lp = np.stack([np.full(N, 8.0), rng.uniform(-0.4, 0.4, N), rng.uniform(-2.4, -1.6, N)], 1).astype(np.float32)
    ref = ob.ReferenceScene(wd, "glass_row", bunny=False).shadow_hard(pos, nd, lp)
    print("glass_row: distinct coefficients", len(np.unique(ref)), " in (0, 0.1):", float(((ref > 0) & (ref < 0.1)).mean()))
    np.savez_compressed(GOLD / "assoc_glass_row.npz", config=text, bunny=False, pos=pos, ndir=nd, light=lp, hard=ref)
    # ---- the glass bunny ----
    text = fixtures.water_bunny_tex_config(200, 150)
    fixtures.write_config(wd, "glass_bunny", text)
    lp = np.tile(np.array([[-20, 70, 20]], np.float32), (N, 1))
    target = np.stack([rng.uniform(-1.2, 1.2, N), rng.uniform(-1.0, 1.2, N), rng.uniform(-3.2, -1.2, N)], 1)
    dirs = target - lp
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    pos = (target + dirs * rng.uniform(2.0, 6.0, (N, 1))).astype(np.float32)
    nd = (-dirs).astype(np.float32)
    ref = ob.ReferenceScene(wd, "glass_bunny", bunny=True).shadow_hard(pos, nd, lp)
    a4 = np.float32(0.8) * np.float32(0.8) * np.float32(0.8) * np.float32(0.8)
    print("glass_bunny: distinct coefficients", len(np.unique(ref)), " four or more crossings:", float(((ref > 0) & (ref <= a4)).mean()))
    np.savez_compressed(GOLD / "assoc_glass_bunny.npz", config=text, bunny=True, pos=pos, ndir=nd, light=lp, hard=ref)


if __name__ == "__main__":
    main()
