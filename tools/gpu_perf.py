"""Ad-hoc 4K timing of the benchmark configs (run under gpurun)."""
import sys, time, json
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
from whittedstyle_raytracer_b200 import Scene, Renderer, fixtures
from whittedstyle_raytracer_b200.renderer import TRAVERSAL_EXHAUSTIVE, TRAVERSAL_PRUNED

wd = Path("/tmp/wrt_perf"); fixtures.ensure_assets(wd)
names = sys.argv[1:] or ["config", "bunny_shadow_4k", "gla_bunny_tex_4k", "water_bunny_tex_soft_4k"]
for name in names:
    fixtures.write_config(wd, name, fixtures.bench_config_text(name))
    s = Scene.from_workdir(wd, name)
    r = Renderer(s)
    r.ctx.enable_kernel_timing(True)
    for mode in (TRAVERSAL_PRUNED, TRAVERSAL_EXHAUSTIVE):
        r.ctx.set_options(traversal=mode)
        ts = []
        for it in range(4):
            img = r.render(); ts.append(r.last_stats["gpu_ms"])
        st = r.last_stats
        kt = r.ctx.kernel_times()
        print(name, "mode", mode, "gpu_ms", [round(t, 3) for t in ts], "rays", st["rays"], "Mrays/s", round(st["rays"] / min(ts[1:]) / 1e3, 1),
              "per_depth", st["rays_per_depth"], {k: round(v, 3) for k, v in kt.items() if v > 0}, flush=True)
    r.ctx.close()
