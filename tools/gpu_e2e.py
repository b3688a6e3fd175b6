"""Where the end-to-end step spends its time (development aid)."""
import sys, time, ctypes as C
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import torch
from whittedstyle_raytracer_b200 import Scene, Renderer, fixtures, cabi
wd = Path("/tmp/wrt_perf"); fixtures.ensure_assets(wd)
name = "water_bunny_tex_soft_4k"
fixtures.write_config(wd, name, fixtures.bench_config_text(name))
s = Scene.from_workdir(wd, name); r = Renderer(s); lib = cabi.load_cuda(); ctx = r.ctx
img = torch.empty((s.height, s.width, 3), dtype=torch.uint8).pin_memory()
st = cabi.WrtStats()
for it in range(4):
    t0 = time.perf_counter(); ctx._check(lib.wrt_upload_scene(ctx.h, s.desc_ptr)); t1 = time.perf_counter()
    ctx._check(lib.wrt_set_camera(ctx.h, s.camera_ptr)); ctx._check(lib.wrt_render(ctx.h, img.data_ptr(), C.byref(st))); t2 = time.perf_counter()
    print(f"upload {1e3*(t1-t0):.2f} ms, render+D2H {1e3*(t2-t1):.2f} ms (gpu {st.gpu_ms:.2f})")
