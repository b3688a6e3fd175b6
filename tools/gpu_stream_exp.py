"""Development aid: does the launching stream / L2 flush change the frame time?"""
import sys, time
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import torch
from whittedstyle_raytracer_b200 import Scene, Renderer, fixtures
from whittedstyle_raytracer_b200.parallel import DistributedRenderer
wd = Path("/tmp/wrt_perf"); fixtures.ensure_assets(wd)
name = "water_bunny_tex_soft_4k"
fixtures.write_config(wd, name, fixtures.bench_config_text(name))
s = Scene.from_workdir(wd, name)
r = Renderer(s)
for it in range(3):
    r.render()
print("wrt_render (own stream): gpu_ms", round(r.last_stats["gpu_ms"], 3))
n = r.ctx.tile_pixel_count(0, 1)
packed = torch.zeros(n * 3, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
def timed(stream_obj, flush=None, reps=3):
    out = []
    for _ in range(reps):
        if flush is not None:
            flush.fill_(1)
        with torch.cuda.stream(stream_obj):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            r.render_device(packed.data_ptr(), torch.cuda.current_stream().cuda_stream)
            e1.record()
            st = r.finish_device()
        torch.cuda.synchronize()
        out.append((round(e0.elapsed_time(e1), 3), round(st["gpu_ms"], 3)))
    return out
print("render_device default stream:", timed(torch.cuda.default_stream()))
side = torch.cuda.Stream()
print("render_device side stream:   ", timed(side))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
print("default stream + L2 flush:   ", timed(torch.cuda.default_stream(), flush))
print("side stream + L2 flush:      ", timed(side, flush))
big = torch.empty(8 << 30, dtype=torch.uint8, device="cuda")
print("after allocating 8 GiB more: ", timed(side))
