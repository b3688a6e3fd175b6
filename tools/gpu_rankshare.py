"""Development aid: time one rank's share (rank 0 of 8) of the metric frame on a single GPU."""
import sys, os
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from whittedstyle_raytracer_b200 import Scene, Renderer, fixtures
wd = Path("/tmp/wrt_perf"); fixtures.ensure_assets(wd)
name = "water_bunny_tex_soft_4k"
fixtures.write_config(wd, name, fixtures.bench_config_text(name))
s = Scene.from_workdir(wd, name); r = Renderer(s)
for world in (8, 1):
    r.ctx.set_tiles(8, 4, 0, world)
    ts = []
    for it in range(9):
        r.render(); ts.append(r.last_stats["gpu_ms"])
    print("world", world, "rank0 ms", round(min(ts[1:]), 3), "rays", r.last_stats["rays"])
r.ctx.enable_kernel_timing(True)
for world in (8, 1):
    r.ctx.set_tiles(8, 4, 0, world)
    for it in range(3):
        r.render()
    print("world", world, "serialized ms", round(r.last_stats["gpu_ms"], 3), {k: round(v, 3) for k, v in r.ctx.kernel_times().items() if v > 0})
