#!/usr/bin/env python
"""Golden vectors of BASELINE.json configs[4] (glass bunny + shadow soft @ 7680x4320) from the CPU oracle.

A full 8K soft-shadow frame is 3.04 G rays: ~10 minutes on 8 cores with the OpenMP oracle, far too long for a test,
so it is rendered ONCE here (build container) and what the GPU tests need is committed as a small fixture:
  * the oracle's ray counts (closest, shadow, per depth) — the reference-equivalent counts of the whole frame;
  * per-band channel sums (bands of 32 rows) of the 8-bit image — a full-frame check that localises a defect;
  * every 16th pixel in x and y (480 x 270) — pixel-exact sample;
The oracle itself is pinned to the unmodified reference by tests/test_oracle_vs_reference.py.

    python tools/gen_full_frame_golden.py            # ~10 min; writes tests/golden/full_glass_bunny_soft_8k.npz
"""
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))

NAME = "glass_bunny_soft_8k"
BAND = 32
STRIDE = 16


def main():
    import oracle_bindings as ob
    from whittedstyle_raytracer_b200 import Scene, fixtures
    wd = Path(tempfile.mkdtemp(prefix="gold8k_"))
    fixtures.ensure_assets(wd)
    fixtures.write_config(wd, NAME, fixtures.bench_config_text(NAME))
    scene = Scene.from_workdir(wd, NAME, glass=True)
    if len(sys.argv) > 2 and sys.argv[1] == "--from-npz":      # (image + counts of an earlier run of this script)
        z = np.load(sys.argv[2])
        img, closest, shadow, per_depth = z["img"], int(z["closest"]), int(z["shadow"]), z["per_depth"]
    else:
        t0 = time.time()
        img, st = ob.OracleScene(scene).render()
        closest, shadow, per_depth = int(st.closest_rays), int(st.shadow_rays), np.array([int(x) for x in st.rays_per_depth])
        print(f"oracle: {closest} closest + {shadow} shadow rays in {time.time() - t0:.0f} s")
    h, w = img.shape[:2]
    bands = img.reshape(h // BAND, BAND, w, 3).astype(np.int64).sum(axis=(1, 2))
    out = REPO / "tests" / "golden" / f"full_{NAME}.npz"
    np.savez_compressed(out, closest_rays=closest, shadow_rays=shadow, rays_per_depth=np.asarray(per_depth, np.int64),
                        band_rows=BAND, band_sums=bands, sample_stride=STRIDE, sample=img[::STRIDE, ::STRIDE].copy(),
                        width=w, height=h, config=fixtures.bench_config_text(NAME))
    print("wrote", out, out.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
