#!/usr/bin/env python
"""BASELINE.json config 5: synthetic stress scene at 1920x1080 -- 1 M Gaussians / 10 M voxels -- rasteriser-backward HBM-roofline sweep,
on 1 ... 8 B200s (one process per GPU under torchrun; Gaussians sharded by spatial block with the peer-memory exchange, voxel hash sharded).

Gaussian side, for N in {125k, 250k, 500k, 1M} Gaussians in total: each stage of one training iteration timed with CUDA events on this
rank's shard (L2 flushed between launches), the whole iteration (collective: exchange barriers inside) as the MAX over ranks; algorithmic
bytes 24 P + 48 I + 80 N_vis of the backward over the measured HBM peak, pairs tested / passed by the backward.
TSDF side: a 1920x1080 depth frame of a synthetic scene fused into an empty map until ~10 M voxels (19.5 k blocks of 512) are allocated,
per-stage device times of ProcessFrame on fresh frames + one free-view raycast.

Prints one JSON line (rank 0).  `python bench.py --config 5` prints the same record wrapped in the bench line's keys.
usage (GPU box): python tools/stress_sweep.py            |  torchrun --nproc-per-node 8 tools/stress_sweep.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

W, H = 1920, 1080


def _gaussian_sweep(dev, local, rank, world, peak, comm):
    import numpy as np
    import torch
    import torch.distributed as dist
    from gps_slam_b200 import engine as E, parallel
    from tests.helpers_gs import camera, random_splats
    c2w, K = camera(W, H, 77)
    intr = dict(width=W, height=H, fx=float(K[0, 0]), fy=float(K[1, 1]), cx=float(K[0, 2]), cy=float(K[1, 2]))
    yy, xx = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
    ref_depth = (2.2 + 0.7 * torch.sin(xx / 97.0) * torch.cos(yy / 61.0)).float().contiguous()
    g = torch.Generator(device=dev).manual_seed(1)
    base = torch.rand(H, W, 3, device=dev, generator=g)
    gt = (base + 0.1 * torch.randn(H, W, 3, device=dev, generator=g)).clamp(0, 1)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.Stream(device=dev)

    def vmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def vsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    rows = []
    for N in (125_000, 250_000, 500_000, 1_000_000):
        p = parallel.shard_params(random_splats(N, seed=77, scale_lo=0.003, scale_hi=0.010), rank, world)
        n_own = len(p["means"])
        eng = E.GaussianEngine(W, H, capacity=max(n_own, 1024), device=local)
        eng.set_stream(stream.cuda_stream)
        if comm is not None:
            eng.set_comm(comm)
        eng.set_params(p)
        eng.initOptimizers()

        def t(fn, reps=10):
            ms = []
            with torch.cuda.stream(stream):
                for _ in range(2):
                    fn()
                for _ in range(reps):
                    flush.fill_(1)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    fn()
                    e1.record(stream)
                    e1.synchronize()
                    ms.append(e0.elapsed_time(e1))
            return float(np.mean(ms)) * 1e3
        with torch.cuda.stream(stream):
            eng.train_step(c2w, intr, ref_depth, base, gt)
        row = {"gaussians": N, "gaussians_this_rank": n_own}
        row["project_sh_us"] = vmax(t(lambda: eng.run_stage(0)))
        row["project_sh+bin_us"] = vmax(t(lambda: eng.run_stage(1)))
        row["raster_fwd_us"] = vmax(t(lambda: eng.run_stage(2)))
        bwd_local = t(lambda: eng.run_stage(3))
        row["raster_bwd_us"] = vmax(bwd_local)
        with torch.cuda.stream(stream):
            tested, passed = eng.bwd_pair_stats()
            eng.run_stage(4)
        row["train_step_us"] = vmax(t(lambda: eng.train_step(c2w, intr, ref_depth, base, gt)))
        cnt = eng.counters()
        I, nvis = int(cnt[0]), int(cnt[4])
        # roofline of the backward on THIS rank's shard (the kernel is rank-local): its own algorithmic bytes over its own time
        alg = 24 * W * H + 48 * I + 80 * nvis
        row.update(isects_this_rank=I, visible_this_rank=nvis, isects=int(vsum(I)), visible=int(vsum(nvis)),
                   bwd_algorithmic_bytes_this_rank=alg, bwd_achieved_gbs=alg / bwd_local / 1e3, bwd_hbm_frac=alg / bwd_local / 1e3 / peak,
                   bwd_pairs_tested=int(vsum(tested)), bwd_pairs_passed=int(vsum(passed)), overflow_flags=int(vmax(int(cnt[2]))),
                   train_iterations_per_sec=1e6 / row["train_step_us"])
        rows.append(row)
        eng.close()
    return rows


def _tsdf_stress(dev, local, rank, world):
    """~10 M voxels: frames of the synthetic room at 1920x1080 until 19.5 k blocks are allocated, then the per-stage times of fresh frames"""
    import numpy as np
    import torch
    from gps_slam_b200 import engine as E, parallel, synthetic as syn
    intr = dict(width=W, height=H, fx=960.0, fy=960.0, cx=959.5, cy=539.5)
    n = 14
    poses = syn.trajectory(n)
    frames = [syn.render_frame(poses[i], intr, device=dev) for i in range(n)]
    torch.cuda.synchronize()
    eng = parallel.make_sharded_tsdf(intr, rank, world, local) if world > 1 else E.TsdfEngine(intr, device=local)
    rows, alloc = [], []
    eng.enable_stage_timing(True)
    for i in range(n):
        eng.ProcessFrameDevice(frames[i][0], frames[i][1], syn.c2w_to_colmajor(poses[i]))
        rows.append(eng.stage_times())
        alloc.append(eng.num_blocks - 1 - eng.counter(0))
    eng.enable_stage_timing(False)
    st = torch.cuda.ExternalStream(eng.stream())
    ms = []
    for i in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        eng.runRaycast(syn.c2w_to_colmajor(poses[n - 1 - i]), intr)
        e1.record(st)
        e1.synchronize()
        ms.append(e0.elapsed_time(e1))
    m = np.asarray(rows[4:]).mean(0) * 1e3
    out = {k: float(v) for k, v in zip(("track_us", "allocate_us", "integrate_us", "expected_depth_us", "raycast_us", "icp_maps_us"), m)}
    out.update(frames=n, allocated_blocks=alloc[-1], allocated_voxels=alloc[-1] * 512, visible_blocks_last_frame=eng.counter(2),
               free_view_raycast_us=float(np.median(ms)) * 1e3, shard_error=eng.shard_error() if world > 1 else 0,
               sharded="mode 1 (owner integrates, every rank stores)" if world > 1 else "single GPU")
    eng.close()
    return out


def main(bench_line=False, args=None):
    import torch
    import torch.distributed as dist
    import bench
    from gps_slam_b200 import engine as E, parallel
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    E.load_library()
    peak, src = bench.load_peaks()
    comm = parallel.make_peer_comm(local, rank, world, W, H) if world > 1 else None
    rows = _gaussian_sweep(dev, local, rank, world, peak, comm)
    tsdf = _tsdf_stress(dev, local, rank, world)
    if comm is not None:
        comm.close()
    rec = {"what": "config 5 stress sweep, 1920x1080, %d x B200" % world, "n_gpus": world, "peak_gbs": peak, "peak_source": src, "rows": rows,
           "tsdf_10M_voxels": tsdf}
    if rank == 0:
        if bench_line:
            big = rows[-1]
            print(json.dumps({
                "metric": "gs_train_iterations_per_sec", "value": big["train_iterations_per_sec"], "unit": "iterations/s", "n_gpus": world,
                "steps": 10, "warmup": 2, "ms_per_step": big["train_step_us"] * 1e-3, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "BASELINE config 5: synthetic stress scene, 1,000,000 Gaussians at 1920x1080 (one GES training iteration: "
                                       "projection + SH, binning, rasteriser forward + loss, backward, Adam), 10 M-voxel TSDF frame times beside it",
                           "baseline_config": 5, "l2": "256 MiB fill between timed launches"},
                "roofline": {"kernel": "k_raster_bwd", "bound": "hbm", "achieved": big["bwd_achieved_gbs"], "peak": peak, "unit": "GB/s",
                             "frac": big["bwd_hbm_frac"], "traffic": None, "peak_source": src},
                "sweep": rec}))
        else:
            print(json.dumps(rec))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
