"""Multi-GPU frame driver: one process per GPU (torch.distributed), scene replicated,
image sharded as interleaved tiles (include/wrt_tiles.h), finished 8-bit tiles gathered to
rank 0 over NVLink with one NCCL gather and de-interleaved there by a small kernel.

The reference has no parallelism at all; the only exchange step this
path has is that single gather, so no other collective exists.  Results are rank-count
invariant: every pixel is independent and the soft-shadow RNG is keyed on the global pixel.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import cabi
from .renderer import Renderer
from .scene import Scene

TILE_W, TILE_H = 8, 4      # one warp-sized pixel block per tile: finest interleave, best balance


def tile_slot_count(width, height, rank, world, tile=(TILE_W, TILE_H)) -> int:
    n = cabi.load_host().wrt_tile_slot_count(width, height, tile[0], tile[1], rank, world)
    if n < 0:
        raise ValueError("bad tile geometry")
    return int(n)


def tile_pixel_map(width, height, rank, world, tile=(TILE_W, TILE_H)) -> np.ndarray:
    """out[slot] = y*width + x of the pixel rank `rank` renders in that slot (-1: padding)."""
    n = tile_slot_count(width, height, rank, world, tile)
    out = np.empty(n, np.int64)
    if cabi.load_host().wrt_tile_pixel_map(width, height, tile[0], tile[1], rank, world, out.ctypes.data, n) != 0:
        raise ValueError(cabi.load_host().wrt_host_last_error().decode())
    return out


def scatter_tiles_host(gathered: np.ndarray, width, height, world, tile=(TILE_W, TILE_H)) -> np.ndarray:
    """CPU twin of the rank-0 scatter kernel: (world, stride) uint8 -> (H, W, 3) uint8."""
    gathered = np.ascontiguousarray(gathered, np.uint8)
    img = np.zeros((height, width, 3), np.uint8)
    rc = cabi.load_host().wrt_scatter_tiles_host(width, height, tile[0], tile[1], world, gathered.ctypes.data,
                                                 gathered.shape[1], img.ctypes.data)
    if rc != 0:
        raise ValueError(cabi.load_host().wrt_host_last_error().decode())
    return img


class TileGather:
    """Gather of the ranks' tile-order RGB buffers to rank 0.  Works on any
    torch.distributed backend (NCCL on GPUs; gloo in the CPU tests of the host logic)."""

    def __init__(self, width, height, rank, world, tile=(TILE_W, TILE_H)):
        self.width, self.height, self.rank, self.world, self.tile = width, height, rank, world, tile
        self.slots = [tile_slot_count(width, height, r, world, tile) for r in range(world)]
        self.stride = max(self.slots) * 3          # rank 0 owns the most tiles; others are padded to it

    def new_buffer(self, device):
        import torch
        return torch.zeros(self.stride, dtype=torch.uint8, device=device)

    def gather(self, packed, gathered=None):
        """packed: this rank's (stride,) uint8 tensor.  Returns the (world, stride) tensor on rank 0."""
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return packed.view(1, -1)
        if self.rank == 0:
            if gathered is None:
                gathered = torch.empty((self.world, self.stride), dtype=torch.uint8, device=packed.device)
            dist.gather(packed, [gathered[r] for r in range(self.world)], dst=0)
            return gathered
        dist.gather(packed, None, dst=0)
        return None


def any_rank_flag(flag: bool, device="cpu") -> bool:
    """True on every rank when `flag` is true on at least one (one int32 all-reduce; any backend)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return bool(int(t.item()))


class DistributedRenderer:
    """Per-rank renderer + the gather.  `frame()` leaves the full image on rank 0's GPU."""

    def __init__(self, scene: Scene, rank: int, world: int, device: int, tile=(TILE_W, TILE_H)):
        import torch
        self.torch = torch
        self.rank, self.world, self.device = rank, world, device
        self.renderer = Renderer(scene, device=device)
        self.renderer.ctx.set_tiles(tile[0], tile[1], rank, world)
        cam = scene.camera
        self.width, self.height = cam.width, cam.height
        self.tg = TileGather(cam.width, cam.height, rank, world, tile)
        dev = torch.device("cuda", device)
        self.packed = self.tg.new_buffer(dev)
        self.gathered = torch.empty((world, self.tg.stride), dtype=torch.uint8, device=dev) if rank == 0 else None
        self.image = torch.zeros((cam.height, cam.width, 3), dtype=torch.uint8, device=dev) if rank == 0 else None
        self.render_done = None                        # optional torch.cuda.Event recorded after the render kernels

    def frame(self):
        """Enqueues render -> gather -> scatter on torch's current stream (asynchronous)."""
        stream = self.torch.cuda.current_stream().cuda_stream
        self.renderer.render_device(self.packed.data_ptr(), stream)
        if self.render_done is not None:
            self.render_done.record()                  # this rank's own tiles are finished here
        self._gather_scatter(stream)
        return self.image

    def _gather_scatter(self, stream):
        g = self.tg.gather(self.packed, self.gathered)
        if self.rank == 0:
            self.renderer.scatter_tiles(g.data_ptr(), self.world, self.tg.stride, self.image.data_ptr(), stream)

    def finish(self) -> dict:
        """Waits for this rank's frame.  A rank whose frame overflowed a ray queue has re-rendered it inside
        finish_device() — after frame() had already gathered its incomplete tiles — so the ranks agree (one flag
        all-reduce, N > 1 only) on whether the gather + scatter must be repeated."""
        stats = self.renderer.finish_device()
        if self.world > 1:
            if any_rank_flag(bool(stats["overflow_retries"]), self.packed.device):
                self._gather_scatter(self.torch.cuda.current_stream().cuda_stream)
                self.torch.cuda.current_stream().synchronize()
        return stats

    def close(self):
        self.renderer.ctx.close()
