"""Host-side SLAM loop over the C-ABI engines -- the Python mirror of SLAMPipeline::SLAMTrainCams
(reference slam/slam_pipeline.cpp:52-173) for the parts that sit on the hot path.  bench.py drives it.

mode "recon": TSDF fusion only per frame (reference work_mode == "recon", slam_pipeline.cpp:96).
"""
import numpy as np
import torch

from . import engine as E
from . import synthetic as syn

DEFAULT_MODE = "recon"


def workload_name(mode):
    if mode == "recon":
        return "Replica-shaped 1200x680 synthetic RGB-D, work_mode=recon (TSDF fusion + raycast per frame, use_gt_pose=true)"
    return "Replica office0-shaped 1200x680 synthetic RGB-D, work_mode=train (gsplat GES + TSDF, use_gt_pose=true)"


class SlamPipeline:
    def __init__(self, intr, mode="recon", device=0, stream=None, rank=0, world=1):
        self.intr, self.mode, self.rank, self.world = intr, mode, rank, world
        self.tsdf = E.TsdfEngine(intr, tracker=0, device=device)
        self.stream = stream
        if stream is not None:
            self.tsdf.set_stream(stream.cuda_stream)
        self.W, self.H = intr["width"], intr["height"]
        self._pose_host = np.zeros(16, np.float32)
        self._vis_sum, self._vis_n = 0, 0

    def reset(self):
        self.tsdf.resetAll()
        self._vis_sum, self._vis_n = 0, 0

    def close(self):
        self.tsdf.close()

    def process_frame(self, idx, rgba_all, depth_all, poses, resident):
        c2w = syn.c2w_to_colmajor(poses[idx])
        if resident:
            self.tsdf.ProcessFrameDevice(rgba_all[idx], depth_all[idx], c2w)
        else:
            self.tsdf.ProcessFrame(rgba_all[idx], depth_all[idx], c2w)

    def end_of_step(self, resident):
        if not resident:
            # the call a user makes after a cycle: read the pose estimate back (est_pose, slam_pipeline.cpp:81-82)
            self.tsdf.sync()
            self._pose_host = self.tsdf.pose()[1]

    def stats(self):
        return {"visible_blocks_last_frame": self.tsdf.counter(2), "allocated_blocks": self.tsdf.num_blocks - 1 - self.tsdf.counter(0)}

    def io_bytes_per_step(self, frames_per_step):
        return frames_per_step * self.W * self.H * 6, 64

    def scaling(self):
        return "weak"

    def time_dominant_kernel(self, stream, peak_gbs, reps=20):
        """integrate kernel: algorithmic bytes V*(4+16+2*4096) + 8*P, CUDA events on the launching stream,
        L2 flushed between launches"""
        V = self.tsdf.counter(2)
        P = self.W * self.H
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        ms = []
        with torch.cuda.stream(stream):
            for _ in range(3):
                self.tsdf.run_stage(1)
            for _ in range(reps):
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                self.tsdf.run_stage(1)
                e1.record(stream)
                e1.synchronize()
                ms.append(e0.elapsed_time(e1))
        t = float(np.mean(ms)) * 1e-3
        alg = V * (4 + 16 + 2 * 4096) + 8 * P
        ach = alg / t / 1e9
        return {"kernel": "k_integrate_tma", "bound": "hbm", "achieved": ach, "peak": peak_gbs, "unit": "GB/s", "frac": ach / peak_gbs,
                "traffic": None, "algorithmic_bytes": alg, "avg_launch_us": t * 1e6, "units": {"visible_blocks": V, "pixels": P}}
