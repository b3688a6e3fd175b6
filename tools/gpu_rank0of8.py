"""Development aid: rank 0 of 8's share of the metric frame, 3 frames (for an ncu launch list)."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from whittedstyle_raytracer_b200 import Scene, Renderer, fixtures
wd = Path("/tmp/wrt_perf"); fixtures.ensure_assets(wd)
name = "water_bunny_tex_soft_4k"
fixtures.write_config(wd, name, fixtures.bench_config_text(name))
s = Scene.from_workdir(wd, name); r = Renderer(s)
r.ctx.set_tiles(8, 4, 0, int(sys.argv[1]) if len(sys.argv) > 1 else 8)
for it in range(3):
    r.render()
print(r.last_stats["gpu_ms"], r.last_stats["rays_per_depth"])
