"""Small end-to-end case for compute-sanitizer: every kernel family runs once (hard + soft + directional,
textures, bump, spheres, tile sharding, batch queries)."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import numpy as np
from whittedstyle_raytracer_b200 import Scene, Renderer, fixtures
wd = Path("/tmp/wrt_san"); fixtures.ensure_assets(wd)
cases = [("a", fixtures.water_bunny_tex_config(64, 48), True), ("b", fixtures.water_bunny_tex_config(40, 30, soft=True), True),
         ("c", fixtures.directional_config(48, 36), False), ("d", fixtures.bump_config(48, 36), False),
         ("e", fixtures.spheres_config(48, 36) + "shadow soft\n", False)]
for name, text, bunny in cases:
    fixtures.write_config(wd, name, text)
    s = Scene.from_workdir(wd, name, bunny=bunny)
    r = Renderer(s)
    img = r.render()
    r.ctx.set_tiles(8, 4, 1, 3)
    part = np.zeros_like(img); r.render(out=part)
    o = np.random.default_rng(0).uniform(-3, 3, (2000, 3)).astype(np.float32)
    d = np.random.default_rng(1).normal(size=(2000, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
    h = r.interStrategy.UpdateInter(o, d)
    r.interStrategy.getShadowCoeffi(o, d, o + 5); r.interStrategy.getSoftShadowSample(o, d, o + 5)
    r.interStrategy.getDirectionalShadowCoeffi(o, np.zeros(2000, np.int32), np.concatenate([d, np.zeros((2000, 1), np.float32)], axis=1))
    print(name, img.mean(), r.last_stats["rays"], int(h["hit"].sum()))
    r.ctx.close()
print("sanitize case ok")
