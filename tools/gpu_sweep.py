"""Development sweep of the launch/refill knobs (env overrides read by wrt_create)."""
import itertools, os, subprocess, sys, json
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
code = r'''
import sys; sys.path.insert(0, %r)
from pathlib import Path
from whittedstyle_raytracer_b200 import Scene, Renderer, fixtures
wd = Path("/tmp/wrt_perf"); fixtures.ensure_assets(wd)
out = []
for name in %r:
    fixtures.write_config(wd, name, fixtures.bench_config_text(name))
    s = Scene.from_workdir(wd, name); r = Renderer(s)
    ts = []
    for it in range(4):
        r.render(); ts.append(r.last_stats["gpu_ms"])
    out.append(round(min(ts[1:]), 3)); r.ctx.close()
print(out)
'''
names = ["bunny_shadow_4k", "water_bunny_tex_soft_4k"]
grid = sys.argv[1:] or ["WRT_TRACE_BLOCKS=8", "WRT_TRACE_BLOCKS=9", "WRT_TRACE_BLOCKS=12", "WRT_TRACE_BLOCKS=16",
                        "WRT_REFILL=4", "WRT_REFILL=16", "WRT_REFILL=32", "WRT_REFILL_SOFT=8", "WRT_REFILL_SOFT=24", "WRT_REFILL_SOFT=32",
                        "WRT_REFILL0=16", "WRT_REFILL0=8"]
for setting in ["default"] + grid:
    env = dict(os.environ)
    if setting != "default":
        for kv in setting.split(","):
            k, v = kv.split("="); env[k] = v
    p = subprocess.run([sys.executable, "-c", code % (str(REPO), names)], env=env, capture_output=True, text=True)
    print(setting, p.stdout.strip().splitlines()[-1] if p.stdout.strip() else p.stderr[-300:], flush=True)
