"""Development aid: time the library named by WRT_CUDA_LIB (default: the in-tree build) on some workloads: whole frame and rank 0's
share of an 8-rank frame (min of 8 frames), the serialised per-family times, and the CRC of the whole image (variants must agree).
Usage: gpu_variant_time.py [workload ...]"""
import os
import sys
import zlib
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from whittedstyle_raytracer_b200 import Scene, Renderer, fixtures  # noqa: E402

names = sys.argv[1:] or ["water_bunny_tex_soft_4k", "bunny_shadow_4k"]
wd = Path("/tmp/wrt_perf")
fixtures.ensure_assets(wd)
tag = os.path.basename(os.environ.get("WRT_CUDA_LIB", "default"))
for name in names:
    fixtures.write_config(wd, name, fixtures.bench_config_text(name))
    r = Renderer(Scene.from_workdir(wd, name))
    out = {}
    for world in (1, 8):
        r.ctx.set_tiles(8, 4, 0, world)
        ts = []
        for it in range(9):
            img = r.render()
            ts.append(r.last_stats["gpu_ms"])
        out[world] = min(ts[1:])
        if world == 1:
            crc = zlib.crc32(img.tobytes())
    print(f"[{tag}] {name}: frame {out[1]:.3f} ms  1/8 share {out[8]:.3f} ms  crc {crc:08x}", flush=True)
    r.ctx.enable_kernel_timing(True)
    for world in (1, 8):
        r.ctx.set_tiles(8, 4, 0, world)
        for it in range(3):
            r.render()
        print(f"    world {world} serialised {r.last_stats['gpu_ms']:.3f} ms",
              {k: round(v, 3) for k, v in r.ctx.kernel_times().items() if v > 0}, flush=True)
    r.ctx.close()
