"""In-process multi-GPU frame time (wrt_multi_*: host threads, peer stores into GPU 0's frame) for N = 1, 2, 4, 8 GPUs of
this box, beside the torchrun + NCCL-gather path bench.py measures.  Prints one JSON line per N.
    python tools/gpu_multi_bench.py [workload] [steps]"""
import ctypes as C
import json
import sys
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import numpy as np
import torch

from whittedstyle_raytracer_b200 import MultiRenderer, Scene, fixtures

name = sys.argv[1] if len(sys.argv) > 1 else "water_bunny_tex_soft_4k"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
wd = Path("/tmp/wrt_multi_bench"); fixtures.ensure_assets(wd)
fixtures.write_config(wd, name, fixtures.bench_config_text(name))
cfg = fixtures.BENCH_CONFIGS[name]
scene = Scene.from_workdir(wd, name, bunny=cfg.get("bunny", True), glass=bool(cfg.get("glass")))
host = torch.empty((scene.height, scene.width, 3), dtype=torch.uint8).pin_memory()
ref = None
for n in (1, 2, 4, 8):
    if torch.cuda.device_count() < n:
        break
    m = MultiRenderer(scene, list(range(n)))
    for _ in range(3):
        m.render(out=host.data_ptr())
    gpu_ms, wall = [], []
    for _ in range(steps):
        t0 = time.perf_counter()
        m.render(out=host.data_ptr())
        wall.append((time.perf_counter() - t0) * 1e3)
        gpu_ms.append(m.last_stats["gpu_ms"])
    e2e = []
    for _ in range(steps):
        t0 = time.perf_counter()
        m.upload()
        m.render(out=host.data_ptr())
        e2e.append((time.perf_counter() - t0) * 1e3)
    chk = int(host.numpy().astype(np.uint64).sum())
    ref = chk if ref is None else ref
    print(json.dumps({"workload": name, "n_gpus": n, "peer_stores": m.uses_peer_stores, "slowest_gpu_render_ms": round(float(np.median(gpu_ms)), 3),
                      "frame_wall_ms_incl_d2h": round(float(np.median(wall)), 3), "e2e_upload_render_d2h_ms": round(float(np.median(e2e)), 3),
                      "image_checksum": chk, "same_image_as_1gpu": chk == ref, "rays": m.last_stats["rays"]}), flush=True)
    m.close()
