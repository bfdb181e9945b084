#!/usr/bin/env python
"""BASELINE.json config 5: synthetic stress scene at 1920x1080, rasteriser-backward HBM-roofline sweep over the Gaussian count.
For N in {125k, 250k, 500k, 1M}: one training iteration per stage timed with CUDA events (L2 flushed between launches), algorithmic
bytes 24 P + 48 I + 80 N_vis of the backward over the measured HBM peak.  Prints one JSON line.
usage (GPU box): python tools/stress_sweep.py > gpurun_out/stress_sweep.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    import bench
    from gps_slam_b200 import engine as E
    from tests.helpers_gs import camera, random_splats
    E.load_library()
    W, H = 1920, 1080
    dev = torch.device("cuda", 0)
    peak, src = bench.load_peaks()
    c2w, K = camera(W, H, 77)
    intr = dict(width=W, height=H, fx=float(K[0, 0]), fy=float(K[1, 1]), cx=float(K[0, 2]), cy=float(K[1, 2]))
    yy, xx = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
    ref_depth = (2.2 + 0.7 * torch.sin(xx / 97.0) * torch.cos(yy / 61.0)).float().contiguous()
    g = torch.Generator(device=dev).manual_seed(1)
    base = torch.rand(H, W, 3, device=dev, generator=g)
    gt = (base + 0.1 * torch.randn(H, W, 3, device=dev, generator=g)).clamp(0, 1)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.Stream(device=dev)
    out = []
    for N in (125_000, 250_000, 500_000, 1_000_000):
        p = random_splats(N, seed=77, scale_lo=0.003, scale_hi=0.010)
        eng = E.GaussianEngine(W, H, capacity=N)
        eng.set_stream(stream.cuda_stream)
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
        row = {"gaussians": N}
        row["project_sh_us"] = t(lambda: eng.run_stage(0))
        row["project_sh+bin_us"] = t(lambda: eng.run_stage(1))
        row["raster_fwd_us"] = t(lambda: eng.run_stage(2))
        row["raster_bwd_us"] = t(lambda: eng.run_stage(3))
        with torch.cuda.stream(stream):
            eng.run_stage(4)
        row["train_step_us"] = t(lambda: eng.train_step(c2w, intr, ref_depth, base, gt))
        cnt = eng.counters()
        I, nvis = int(cnt[0]), int(cnt[4])
        alg = 24 * W * H + 48 * I + 80 * nvis
        row.update(isects=I, visible=nvis, bwd_algorithmic_bytes=alg, bwd_achieved_gbs=alg / row["raster_bwd_us"] / 1e3,
                   bwd_hbm_frac=alg / row["raster_bwd_us"] / 1e3 / peak, overflow_flags=int(cnt[2]))
        out.append(row)
        eng.close()
    print(json.dumps({"what": "config 5 stress sweep, 1920x1080, 1xB200", "peak_gbs": peak, "peak_source": src, "rows": out}))


if __name__ == "__main__":
    main()
