"""CPU (gloo, world_size 2) checks of the multi-GPU design of the Gaussian path: block ownership is a partition,
and because the GES blend is an order-independent sum, all-reducing the per-rank partial accumulation images reproduces the
single-process image -- checked with the numpy oracle's rasteriser standing in for the CUDA kernels."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gps_slam_b200 import parallel
from oracle import gs_oracle as go
from tests.helpers_gs import camera, random_splats, scene_images


def test_block_ownership_is_a_partition():
    p = random_splats(5000, seed=1)
    for world in (1, 2, 4, 8):
        own = parallel.owner_of(p["means"], world)
        assert own.min() >= 0 and own.max() < world
        sizes = np.bincount(own, minlength=world)
        assert sizes.sum() == 5000
        if world > 1:
            assert sizes.min() > 5000 / world * 0.7, sizes   # reasonably balanced
        shards = [parallel.shard_params(p, r, world) for r in range(world)]
        assert sum(len(s["means"]) for s in shards) == 5000
    # all points of one 4 cm block go to the same rank
    base = np.array([[1.003, 2.001, 0.485]], np.float32)
    pts = base + np.random.RandomState(0).uniform(0, 0.03, (50, 3)).astype(np.float32)
    assert len(set(parallel.owner_of(pts, 8))) == 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        W, H, N = 96, 64, 400
        p = random_splats(N, seed=5, scale_lo=0.01, scale_hi=0.04)
        c2w, K = camera(W, H, 5)
        ref_depth, base, gt = scene_images(W, H, 5)
        mine = parallel.shard_params(p, rank, world)
        it = go.ges_iteration(mine, c2w, K, W, H, ref_depth, base, gt)
        acc5 = torch.from_numpy(np.concatenate([it["render"].reshape(-1), it["alphas"].reshape(-1)]).astype(np.float32))
        parallel.allreduce_sum_(acc5)                      # the one collective of an iteration
        n = np.array([len(mine["means"])], np.int64)
        t = torch.from_numpy(n)
        parallel.allreduce_sum_(t)
        if rank == 0:
            full = go.ges_iteration(p, c2w, K, W, H, ref_depth, base, gt)
            render = acc5[: W * H * 4].numpy().reshape(H, W, 4)
            alphas = acc5[W * H * 4:].numpy().reshape(H, W)
            rgb, _ = go.composite(render, alphas, ref_depth, base)
            out.put((int(t.item()), float(np.abs(render - full["render"]).max()), float(np.abs(alphas - full["alphas"]).max()),
                     float(np.abs(rgb - full["rgb"]).max())))
    finally:
        dist.destroy_process_group()


def test_partial_images_allreduce_to_the_full_image_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = out.get(timeout=180)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    total, d_render, d_alpha, d_rgb = res
    assert total == 400
    # sums are re-associated across ranks: fp32 tolerance, not bit-exact
    assert d_render < 2e-5 and d_alpha < 2e-5 and d_rgb < 2e-5, res


def test_voxel_block_owner_is_a_balanced_partition():
    """host mirror of the TSDF block-ownership rule (owner = hashIndex(blockPos) mod world, csrc/tsdf.h block_owner): every block has exactly
    one owner in range, the rule is the reference's hash (ITMRepresentationAccess.h:7-11) folded to 20 bits, and a room-sized set of blocks
    spreads evenly over 2 ... 8 ranks"""
    import numpy as np
    from gps_slam_b200 import parallel
    rng = np.random.RandomState(3)
    blocks = rng.randint(-200, 200, size=(50000, 3))
    h = ((blocks[:, 0].astype(np.int64) * 73856093) ^ (blocks[:, 1].astype(np.int64) * 19349669) ^ (blocks[:, 2].astype(np.int64) * 83492791)) & 0xfffff
    for world in (1, 2, 3, 4, 8):
        own = parallel.voxel_block_owner(blocks, world)
        assert own.min() >= 0 and own.max() < world
        assert np.array_equal(own, h % world)
        counts = np.bincount(own, minlength=world)
        assert counts.min() > 0.9 * len(blocks) / world, counts
    # negative coordinates hash like the device's 32-bit unsigned arithmetic
    assert parallel.voxel_block_owner(np.array([[-1, -1, -1]]), 7)[0] == (((-73856093) & 0xffffffff) ^ ((-19349669) & 0xffffffff) ^ ((-83492791) & 0xffffffff)) % (1 << 20) % 7
