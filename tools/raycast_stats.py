"""diagnostic: where the TSDF raycast spends its steps on the bench scene (gsb_tsdf_raycast_stats)"""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from gps_slam_b200 import synthetic as syn
from gps_slam_b200.engine import TsdfEngine, _check

n = int(sys.argv[1]) if len(sys.argv) > 1 else 120
intr = syn.intrinsics("replica")
poses, frames = syn.sequence(n, intr, device="cuda")
eng = TsdfEngine(intr, tracker=0)
eng.L.gsb_tsdf_raycast_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p]
for i in range(n):
    eng.ProcessFrameDevice(frames[i][0].data_ptr(), frames[i][1].data_ptr(), syn.c2w_to_colmajor(poses[i]))
eng.sync()
print("frames", n, "visible blocks", eng.counter(2))
for k in (n - 1, n // 2, 0):
    tot = np.zeros(8, np.uint64)
    c2w = syn.c2w_to_colmajor(poses[k])
    _check(eng.L.gsb_tsdf_raycast_stats(eng.h_, c2w.ctypes.data, intr["fx"], intr["fy"], intr["cx"], intr["cy"], tot.ctypes.data))
    rays, steps, miss, interp, block, wmax, warps = [int(x) for x in tot[:7]]
    print("pose %d: rays %d  steps/ray %.2f  (unallocated %.2f, block changes %.2f, trilinear reads %.2f)  slowest ray per warp %.2f steps" % (
        k, rays, steps / rays, miss / rays, block / rays, interp / rays, wmax / max(warps, 1)))
eng.close()
