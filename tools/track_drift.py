#!/usr/bin/env python
"""Per-frame translation error of the engine's online tracker (extended ICP, use_gt_pose=false) on the synthetic orbit, TSDF-only loop:
where, if anywhere, the tracker loses the map.  Written to compare with the reference's CPU tracker on the same frames
(oracle/_ref/libitm_ref_fast.so, tracker = extended): that one loses track between frames 200 and 250 of this sequence -- the orbit passes
a view the point-to-plane ICP cannot constrain -- so BASELINE config 3's 2000-frame figure says more about the synthetic scene than about
either implementation.
usage (GPU box): python tools/track_drift.py [frames] [scale]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    from gps_slam_b200 import engine as E, synthetic as syn
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 330
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    intr = syn.intrinsics("replica", scale)
    poses = syn.trajectory(n)
    dev = torch.device("cuda", 0)
    eng = E.TsdfEngine(intr, tracker=1)
    eng.set_pose(syn.c2w_to_colmajor(poses[0]))
    errs, iters = [], []
    for i in range(n):
        rgba, d = syn.render_frame(poses[i], intr, device=dev)
        torch.cuda.synchronize()   # the frame is rendered on torch's stream, the engine reads it on its own
        eng.ProcessFrameDevice(rgba, d, None)
        eng.sync()     # the frame tensors are recycled by torch's allocator on the next iteration: the engine's stream must be done with them
        est = eng.pose()[1].reshape(4, 4).T
        errs.append(float(np.linalg.norm(est[:3, 3] - poses[i][:3, 3])))
        iters.append(int(eng.tracker_result()[2]))
    eng.close()
    print(json.dumps({"frames": n, "scale": scale, "translation_error_m_every_10th": [round(e, 5) for e in errs[::10]],
                      "first_frame_above_5cm": next((i for i, e in enumerate(errs) if e > 0.05), None),
                      "icp_evaluations_per_frame": float(np.mean(iters))}))


if __name__ == "__main__":
    main()
