#!/usr/bin/env python
"""GPU-vs-GPU denominator for the TSDF side: the REFERENCE's InfiniTAM CUDA engine, compiled for
sm_100a where it lies (oracle/itm_ref_cuda/Makefile -> oracle/_ref/libitm_ref_cuda.so), and this repository's engine through the same
frames of the bench sequence on the same B200: per ProcessFrame (host frame in, H2D inside, as both programs do) and per free-view raycast
(runRaycast at a keyframe pose).  Both are timed by host wall clock around the call with a device synchronise, because the reference's
engine synchronises inside ProcessFrame anyway.  Writes one JSON line.
usage (GPU box): python tools/time_reference_itm_cuda.py [--frames 400] [--timed 100]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=400)
    ap.add_argument("--timed", type=int, default=100)
    a = ap.parse_args()
    import numpy as np
    import torch
    from gps_slam_b200 import engine as E, synthetic as syn
    from oracle.itm_ref import ItmRef, available
    if not available("cuda"):
        print(json.dumps({"unavailable": "oracle/_ref/libitm_ref_cuda.so not built"}))
        return
    os.environ["ITMREF_DEVICE"] = "cuda"
    dev = torch.device("cuda", 0)
    intr = syn.intrinsics("replica")
    n = a.frames
    poses = syn.trajectory(n)
    frames = []
    for i in range(n):
        rgba, d = syn.render_frame(poses[i], intr, device=dev)
        frames.append((rgba.cpu().numpy(), d.cpu().numpy()))
    t0 = n - a.timed
    out = {"frames": n, "timed_frames": a.timed, "width": intr["width"], "height": intr["height"]}
    free_poses = [syn.c2w_to_colmajor(poses[i]) for i in range(t0, n, max(1, a.timed // 9))][:9]

    def run(name, process, raycast, sync, close, visible=None):
        pf, rc = [], []
        for i in range(n):
            if i >= t0:
                sync()
                t = time.perf_counter()
            process(frames[i][0], frames[i][1], syn.c2w_to_colmajor(poses[i]))
            if i >= t0:
                sync()
                pf.append(time.perf_counter() - t)
        for p in free_poses:
            sync()
            t = time.perf_counter()
            raycast(p)
            sync()
            rc.append(time.perf_counter() - t)
        r = {"process_frame_ms": float(np.mean(pf)) * 1e3, "process_frame_ms_p50": float(np.median(pf)) * 1e3,
             "free_view_raycast_ms": float(np.mean(rc)) * 1e3,
             "tsdf_side_ms_per_10_frame_step": (10 * float(np.mean(pf)) + 9 * float(np.mean(rc))) * 1e3}
        if visible:
            r["visible_blocks_last_frame"] = visible()
        close()
        out[name] = r

    ref = ItmRef(intr, tracker=0, threads=1, kind="cuda")
    run("reference_cuda_engine_sm100a", ref.process_frame, lambda p: ref.run_raycast(p, intr), torch.cuda.synchronize, ref.close)
    eng = E.TsdfEngine(intr, tracker=0)
    pin = [torch.empty(frames[0][0].shape, dtype=torch.uint8).pin_memory(), torch.empty(frames[0][1].shape, dtype=torch.int16).pin_memory()]

    def ours(rgba, d, c2w):
        pin[0].numpy()[...] = rgba   # the reference's ProcessFrame also copies the caller's image first (CLIEngine / createTsdfEngine)
        pin[1].numpy()[...] = d
        eng.ProcessFrame(pin[0], pin[1], c2w)
    run("gps_slam_b200", ours, lambda p: eng.runRaycast(p, intr), eng.sync, eng.close, visible=lambda: eng.counter(2))
    a_, b_ = out["reference_cuda_engine_sm100a"], out["gps_slam_b200"]
    out["speedup_process_frame"] = a_["process_frame_ms"] / b_["process_frame_ms"]
    out["speedup_free_view_raycast"] = a_["free_view_raycast_ms"] / b_["free_view_raycast_ms"]
    out["speedup_tsdf_side"] = a_["tsdf_side_ms_per_10_frame_step"] / b_["tsdf_side_ms_per_10_frame_step"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
