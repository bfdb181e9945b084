"""Where the time of the sharded raycast goes: two (or N) engines in ONE process, one per GPU, peer-mapped directly; a few full-resolution
frames build a scene, then the live raycast stage (no barriers) is timed per rank with CUDA events -- each rank alone and all ranks at once --
for both shard modes and for the probe variants (results kept local / per-thread peer stores / bulk peer stores).
usage: python tools/shard_probe.py [world]"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gps_slam_b200 import engine as E, synthetic as syn  # noqa: E402


def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else min(2, torch.cuda.device_count())
    same_dev = torch.cuda.device_count() < world
    intr = syn.intrinsics("replica")
    n_frames = 6
    poses, frames = syn.sequence(n_frames, intr)
    out = {"world": world, "one_device": same_dev}
    for mode in (1, 0):
        engs = [E.TsdfEngine(intr, rank=r, world=world, device=0 if same_dev else r) for r in range(world)]
        for e in engs:
            e.attach_local(engs)
            e.set_shard_mode(mode)
        fr = []
        for r in range(world):
            dev = torch.device("cuda", 0 if same_dev else r)
            fr.append([(frames[i][0].to(dev), frames[i][1].to(dev)) for i in range(n_frames)])
        torch.cuda.synchronize()
        for i in range(n_frames):
            for r, e in enumerate(engs):
                e.ProcessFrameDevice(fr[r][i][0], fr[r][i][1], syn.c2w_to_colmajor(poses[i]))
        for e in engs:
            e.sync()
        res = {}
        for probe in (0, 1, 2):
            for e in engs:
                e.L.gsb_tsdf_shard_probe(e.h_, probe)
            for stage, name in ((3, "live_raycast"), (1, "integrate")):
                if stage == 1 and probe:
                    continue
                alone, together = [], []
                for r, e in enumerate(engs):
                    torch.cuda.set_device(0 if same_dev else r)
                    st = torch.cuda.ExternalStream(e.stream())
                    ms = []
                    for _ in range(5):
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record(st)
                        e.run_stage(stage)
                        e1.record(st)
                        e1.synchronize()
                        ms.append(e0.elapsed_time(e1))
                    alone.append(round(float(np.median(ms)) * 1e3, 1))
                evs = []
                for _ in range(5):
                    cur = []
                    for r, e in enumerate(engs):
                        torch.cuda.set_device(0 if same_dev else r)
                        st = torch.cuda.ExternalStream(e.stream())
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record(st)
                        e.run_stage(stage)
                        e1.record(st)
                        cur.append((e0, e1))
                    for e in engs:
                        e.sync()
                    evs.append([a.elapsed_time(b) for a, b in cur])
                together = [round(float(x) * 1e3, 1) for x in np.median(np.array(evs), axis=0)]
                res["%s probe%d" % (name, probe)] = {"alone_us": alone, "together_us": together}
        out["mode%d" % mode] = res
        for e in engs:
            e.close()
    single = E.TsdfEngine(intr, device=0)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    f0 = [(frames[i][0].to(dev), frames[i][1].to(dev)) for i in range(n_frames)]
    torch.cuda.synchronize()
    for i in range(n_frames):
        single.ProcessFrameDevice(f0[i][0], f0[i][1], syn.c2w_to_colmajor(poses[i]))
    single.sync()
    st = torch.cuda.ExternalStream(single.stream())
    for stage, name in ((3, "live_raycast"), (1, "integrate")):
        ms = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            single.run_stage(stage)
            e1.record(st)
            e1.synchronize()
            ms.append(e0.elapsed_time(e1))
        out["single_%s_us" % name] = round(float(np.median(ms)) * 1e3, 1)
    single.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
