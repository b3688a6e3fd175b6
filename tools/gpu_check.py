"""Ad-hoc GPU bring-up check: GPU vs oracle on every fixture config (run under gpurun)."""
import sys, time, json
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
from whittedstyle_raytracer_b200 import Scene, Renderer, fixtures
from whittedstyle_raytracer_b200.renderer import TRAVERSAL_EXHAUSTIVE, TRAVERSAL_PRUNED
import oracle_bindings as ob

wd = Path("/tmp/wrt_check"); fixtures.ensure_assets(wd)
cases = []
fixtures.write_config(wd, "config", fixtures.bunny_shadow_config(320, 240)); cases.append(("config", True))
fixtures.write_config(wd, "water", fixtures.water_bunny_tex_config(320, 240)); cases.append(("water", True))
fixtures.write_config(wd, "water_soft", fixtures.water_bunny_tex_config(160, 120, soft=True)); cases.append(("water_soft", True))
for k, f in fixtures.COVERAGE_CONFIGS.items():
    fixtures.write_config(wd, k, f()); cases.append((k, False))
res = {}
for name, bunny in cases:
    s = Scene.from_workdir(wd, name, bunny=bunny)
    o = ob.OracleScene(s)
    t = time.time(); ref, fl, st = o.render(want_float=True); tc = time.time() - t
    r = Renderer(s)
    out = {}
    for mode in (TRAVERSAL_EXHAUSTIVE, TRAVERSAL_PRUNED):
        r.ctx.set_options(traversal=mode)
        img = r.render()
        d = np.abs(img.astype(int) - ref.astype(int))
        gs = r.last_stats
        out[mode] = dict(ndiff=int((d.max(axis=2) > 0).sum()), maxdiff=int(d.max()), n=int(d.shape[0] * d.shape[1]),
                         closest=(gs["closest_rays"], int(st.closest_rays)), shadow=(gs["shadow_rays"], int(st.shadow_rays)),
                         gpu_ms=gs["gpu_ms"])
    po, pd = o.primary_rays()
    ho = o.trace_closest(po, pd); hg = r.interStrategy.UpdateInter(po, pd)
    eq = {k: bool(np.array_equal(ho[k], hg[k])) for k in ["hit", "object", "t", "pos", "ndir", "uv", "texture", "normalmap", "material", "prim"]}
    res[name] = dict(render=out, closest_equal=eq, oracle_s=tc)
    print(name, json.dumps(res[name]), flush=True)
    r.ctx.close()
Path(REPO / "gpurun_out").mkdir(exist_ok=True)
json.dump(res, open(REPO / "gpurun_out/gpu_check.json", "w"), indent=1)
