"""profiling target: 100 frames of TSDF fusion + live raycast, then 4 free-view raycasts (ncu -k regex:k_raycast -s <skip>)"""
import sys
sys.path.insert(0, ".")
from gps_slam_b200 import synthetic as syn
from gps_slam_b200.engine import TsdfEngine
n = 100
intr = syn.intrinsics("replica")
poses, frames = syn.sequence(n, intr, device="cuda")
eng = TsdfEngine(intr, tracker=0)
for i in range(n):
    eng.ProcessFrameDevice(frames[i][0].data_ptr(), frames[i][1].data_ptr(), syn.c2w_to_colmajor(poses[i]))
for k in (10, 40, 70, 99):
    eng.runRaycast(syn.c2w_to_colmajor(poses[k]), intr)
eng.sync()
eng.close()
