"""Development aid: N frames of one bench config on cuda:0 (the command ncu captures wrap)."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from whittedstyle_raytracer_b200 import Scene, Renderer, fixtures

name = sys.argv[1] if len(sys.argv) > 1 else "water_bunny_tex_soft_4k"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 2
wd = Path("/tmp/wrt_perf"); fixtures.ensure_assets(wd)
fixtures.write_config(wd, name, fixtures.bench_config_text(name))
cfg = fixtures.BENCH_CONFIGS[name]
r = Renderer(Scene.from_workdir(wd, name, bunny=cfg.get("bunny", True), glass=bool(cfg.get("glass"))))
for _ in range(frames):
    r.render()
    print(name, r.last_stats["gpu_ms"], r.last_stats["shadow_rays_traced"], flush=True)
r.ctx.close()
