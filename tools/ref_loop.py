#!/usr/bin/env python
"""PSNR of the engine's SLAM loop against the same loop with the REFERENCE's gsplat kernels on the Gaussian side (tests/ref_slam.py):
same frames, same camera-sampling sequence, same spawned pixels.  Writes one JSON line (kept as profiles/r02_psnr_vs_reference.json,
which bench.py quotes as quality.psnr_vs_reference_db).   usage (GPU box): python tools/ref_loop.py [--frames 81] [--scale 1.0]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=81)
    ap.add_argument("--scale", type=float, default=1.0)
    a = ap.parse_args()
    from gps_slam_b200 import engine as E
    E.load_library()
    from tests import ref_slam
    r = ref_slam.run_both(a.frames, a.scale)
    r["what"] = ("PSNR (20 log10(1/sqrt(mse)) against the input frames, every 10th training view) of the engine's loop minus that of the same "
                 "loop on the reference's gsplat CUDA kernels + torch.optim.Adam; TSDF side = this engine (bit-exact to the reference's)")
    print(json.dumps(r))


if __name__ == "__main__":
    main()
