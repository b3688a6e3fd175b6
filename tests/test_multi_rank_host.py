"""N > 1 host logic on CPU: tile ownership (include/wrt_tiles.h) and the gather to rank 0,
run as world_size-2 and -3 `gloo` process groups.  The pixels come from the CPU oracle
here — the point is the sharding/gather plumbing that bench.py --gpus N and
DistributedRenderer use over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from whittedstyle_raytracer_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("w,h,world,tile", [(800, 600, 2, (32, 16)), (97, 61, 3, (8, 4)), (3840, 2160, 8, (32, 16)),
                                           (33, 5, 4, (32, 16)), (64, 32, 1, (32, 16))])
def test_tile_maps_partition_the_image(w, h, world, tile):
    seen = np.zeros(w * h, np.int32)
    sizes = []
    for r in range(world):
        m = parallel.tile_pixel_map(w, h, r, world, tile)
        assert len(m) == parallel.tile_slot_count(w, h, r, world, tile)
        valid = m[m >= 0]
        seen[valid] += 1
        sizes.append(len(valid))
        # a warp's 32 consecutive slots form one 8x4 pixel block
        blk = m[:32]
        if len(blk) == 32 and (blk >= 0).all():
            assert (blk % w).max() - (blk % w).min() == 7 and (blk // w).max() - (blk // w).min() == 3
    assert (seen == 1).all()                       # every pixel rendered exactly once
    assert max(sizes) - min(sizes) <= 4 * tile[0] * tile[1]   # clipped border tiles hold fewer pixels


def test_tiles_balance_the_bunny():
    """Interleaving must spread the bunny's screen area (x 1300..2500, y 900..1900 at 4K,
    ~7 % of the pixels, ~half of the rays) evenly over 8 ranks."""
    w, h, world = 3840, 2160, 8
    share = []
    for r in range(world):
        m = parallel.tile_pixel_map(w, h, r, world)
        m = m[m >= 0]
        x, y = m % w, m // w
        share.append(int(((x >= 1300) & (x < 2500) & (y >= 900) & (y < 1900)).sum()))
    assert max(share) / (sum(share) / world) < 1.08, share


def _worker(rank, world, port, w, h, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        yy, xx = np.mgrid[0:h, 0:w]
        full = np.stack([xx % 251, yy % 241, (xx * 7 + yy * 13) % 239], axis=-1).astype(np.uint8)
        tg = parallel.TileGather(w, h, rank, world)
        m = parallel.tile_pixel_map(w, h, rank, world)
        packed = tg.new_buffer("cpu")
        flat = full.reshape(-1, 3)
        buf = packed.numpy().reshape(-1, 3)
        buf[:len(m)][m >= 0] = flat[m[m >= 0]]      # what this rank's resolve kernel would write
        g = tg.gather(packed)
        if rank == 0:
            img = parallel.scatter_tiles_host(g.numpy(), w, h, world)
            q.put(bool(np.array_equal(img, full)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,w,h", [(2, 200, 150), (3, 97, 61)])
def test_gather_to_rank0_over_gloo(world, w, h):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, w, h, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def _overflow_worker(rank, world, port, w, h, q):
    """The repeat-gather protocol of DistributedRenderer.finish(): rank 1's first frame is incomplete (its queue
    overflowed: half of its pixels are missing); after its re-render every rank learns about it through one flag
    all-reduce and gather + scatter run again."""
    import torch.distributed as dist
    from whittedstyle_raytracer_b200 import parallel
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        yy, xx = np.mgrid[0:h, 0:w]
        full = np.stack([xx % 251, yy % 241, (xx * 7 + yy * 13) % 239], axis=-1).astype(np.uint8)
        tg = parallel.TileGather(w, h, rank, world)
        m = parallel.tile_pixel_map(w, h, rank, world)
        packed = tg.new_buffer("cpu")
        buf = packed.numpy().reshape(-1, 3)
        flat = full.reshape(-1, 3)
        ok = m >= 0
        if rank == 1:
            ok = ok & (np.arange(len(m)) < len(m) // 2)          # dropped rays: the second half never arrived
        buf[:len(m)][ok] = flat[m[ok]]
        g = tg.gather(packed)
        first_ok = None
        if rank == 0:
            first_ok = bool(np.array_equal(parallel.scatter_tiles_host(g.numpy(), w, h, world), full))
        overflowed = rank == 1
        if overflowed:                                           # wrt_finish_device re-rendered the frame
            buf[:len(m)][m >= 0] = flat[m[m >= 0]]
        again = parallel.any_rank_flag(overflowed)
        if again:
            g = tg.gather(packed)
        if rank == 0:
            img = parallel.scatter_tiles_host(g.numpy(), w, h, world)
            q.put((first_ok, again, bool(np.array_equal(img, full))))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_overflowed_rank_triggers_a_second_gather():
    world, w, h = 3, 160, 96
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_overflow_worker, args=(r, world, port, w, h, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    first_ok, again, final_ok = q.get(timeout=5)
    assert first_ok is False and again is True and final_ok is True
