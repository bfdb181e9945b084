#!/usr/bin/env python
"""Times ONE optimiser iteration of path A (gesForward + L1 + backward + Adam, reference src/raw_gs_model.cpp:188-417, 677-705)
two ways on the same GPU, same Gaussians, same camera:

  reference : the reference's own gsplat CUDA kernels + autograd wrappers (oracle/_ref/libgsplat_ref.so) + torch ops + 6x
              torch.optim.Adam, exactly the launch sequence the reference issues (oracle/gsplat_ref.py);
  ours      : gsb_gs_train_step (7 kernels).

The Gaussians are the steady state of the bench workload (13 cycles of the office0-shaped sequence).  SURVEY.md 8(d) names this
the >=10x comparison point, since the reference binary itself cannot be built offline.  Measurement tool, not part of bench.py;
prints one JSON line.   usage (GPU box): python tools/time_reference_gs.py [--cycles 13] [--iters 20]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cycles", type=int, default=13)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    import numpy as np
    import torch
    import bench
    from gps_slam_b200 import engine as E, slam
    from oracle import gsplat_ref
    E.load_library()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    n_frames = a.cycles * 10
    intr, poses, rgba, depth = bench.make_frames(n_frames, dev)
    stream = torch.cuda.Stream(device=dev)
    pipe = slam.SlamPipeline(intr, mode="train", device=0, stream=stream)
    with torch.cuda.stream(stream):
        for f in range(n_frames):
            pipe.process_frame(f, rgba, depth, poses, True)
    torch.cuda.synchronize()
    cam = [c for c in pipe.opt_cams if c.depth_map is not None][-1]
    g = pipe.gs
    N = g.getGaussianNum()
    params = g.get_params()
    W, H = intr["width"], intr["height"]
    K = np.array([[intr["fx"], 0, intr["cx"]], [0, intr["fy"], intr["cy"]], [0, 0, 1]], np.float32)
    ref_depth = cam.depth_map.cpu().numpy()
    base = cam.color_map.cpu().numpy()
    gt = cam.image.cpu().numpy()

    # ---- ours
    with torch.cuda.stream(stream):
        g.initOptimizers()
        for _ in range(3):
            g.train_step(cam.c2w_slam, intr, cam.depth_map, cam.color_map, cam.image)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(a.iters):
            g.train_step(cam.c2w_slam, intr, cam.depth_map, cam.color_map, cam.image)
        e1.record(stream)
    torch.cuda.synchronize()
    ours_ms = e0.elapsed_time(e1) / a.iters
    ours_loss = g.loss()

    # ---- reference kernels, reference launch sequence (device-resident inputs; timing excludes the numpy conversions)
    lrs = dict(means=1.6e-4, scales=5e-3, quats=1e-3, featuresDc=2.5e-3, featuresRest=5e-4, opacities=5e-2)
    params["featuresRest"] = params["featuresRest"].reshape(N, 15, 3)
    ref = gsplat_ref.RefGaussians(params, lrs=lrs)
    c2w_t = torch.as_tensor(np.asarray(cam.c2w_slam, np.float32), device=dev)
    K_t = torch.as_tensor(K, device=dev)
    rd_t, base_t, gt_t = [torch.as_tensor(x, device=dev) for x in (ref_depth, base, gt)]

    def ref_iter():
        r = ref.forward(c2w_t, K_t, W, H, rd_t, base_t)
        loss = torch.abs(gt_t - r["rgb"]).mean()
        loss.backward()
        for k in ref.KEYS:
            ref.opt[k].step()
        for k in ref.KEYS:
            ref.opt[k].zero_grad()
        return loss

    for _ in range(3):
        ref_iter()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        loss = ref_iter()
    e1.record()
    torch.cuda.synchronize()
    ref_ms = e0.elapsed_time(e1) / a.iters
    print(json.dumps({"what": "one optimiser iteration of path A (forward + L1 + backward + Adam), 1xB200, CUDA events",
                      "gaussians": N, "width": W, "height": H, "iters": a.iters,
                      "reference_kernels_ms": ref_ms, "ours_ms": ours_ms, "speedup": ref_ms / ours_ms,
                      "reference_loss": float(loss.detach()), "ours_loss": ours_loss,
                      "note": "reference = /root/reference/gsplat kernels compiled for sm_100a (glm stand-in) + torch 2.11 ops/autograd/Adam in "
                              "the reference's launch order (includes its host syncs in isect_tiles); both start from the same Gaussians"}))
    pipe.close()


if __name__ == "__main__":
    main()
