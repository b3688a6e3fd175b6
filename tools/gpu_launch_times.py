"""Development aid: per-launch times (events around every launch, serialised) of one rank's 1/world share of a workload."""
import sys, os
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from whittedstyle_raytracer_b200 import Scene, Renderer, fixtures
name = sys.argv[1] if len(sys.argv) > 1 else "water_bunny_tex_soft_4k"
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
wd = Path("/tmp/wrt_perf"); fixtures.ensure_assets(wd)
fixtures.write_config(wd, name, fixtures.bench_config_text(name))
s = Scene.from_workdir(wd, name); r = Renderer(s)
r.ctx.set_tiles(8, 4, 0, world)
r.ctx.enable_kernel_timing(True)
for it in range(3):
    r.render()
os.environ["WRT_TIMING_DUMP"] = "1"
sys.stderr.write(f"---- {name} world {world}\n"); sys.stderr.flush()
r.render()
print(name, "world", world, "serialized ms", round(r.last_stats["gpu_ms"], 3), {k: round(v, 3) for k, v in r.ctx.kernel_times().items() if v > 0})
print({k: v for k, v in r.last_stats.items() if "ray" in k or "request" in k})
