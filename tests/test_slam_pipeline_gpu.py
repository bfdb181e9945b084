"""The host loop (gps_slam_b200/slam.py = SLAMPipeline::SLAMTrainCams, slam/slam_pipeline.cpp:52-173) on a short half-resolution
sequence: the two-stream schedule (TSDF side overlapped with the Gaussian optimiser) must produce what the plain single-stream
order produces -- same map, same number of Gaussians, same render up to the fp32 re-association of the backward's atomics."""
import numpy as np
import pytest
import torch

from gps_slam_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def run(overlap, n_frames=31, scale=0.5):
    from gps_slam_b200 import slam
    intr = syn.intrinsics("replica", scale)
    dev = torch.device("cuda", 0)
    poses = syn.trajectory(n_frames)
    frames = [syn.render_frame(poses[i], intr, device=dev) for i in range(n_frames)]
    rgba = torch.stack([f[0] for f in frames])
    depth = torch.stack([f[1] for f in frames])
    stream = torch.cuda.Stream(device=dev)
    pipe = slam.SlamPipeline(intr, mode="train", device=0, stream=stream, overlap=overlap, gs_capacity=1 << 19)
    try:
        with torch.cuda.stream(stream):
            for f in range(n_frames):
                pipe.process_frame(f, rgba, depth, poses, True)
                if f % 10 == 9:
                    pipe.end_of_step(True)
            pipe.end_of_step(True)
        torch.cuda.synchronize()
        n = pipe.gs.getGaussianNum()
        params = pipe.gs.get_params()
        H, W = intr["height"], intr["width"]
        rgb, d, a = torch.empty((H, W, 3), device=dev), torch.empty((H, W), device=dev), torch.empty((H, W), device=dev)
        with torch.cuda.stream(stream):
            pipe.render_eval(poses[20], rgb, d, a)
        torch.cuda.synchronize()
        vis = pipe.tsdf.visible_ids().copy()
        blocks = pipe.tsdf.counter(0)
        return n, params, rgb.cpu().numpy(), vis, blocks, pipe.cycles
    finally:
        pipe.close()


def test_two_stream_schedule_matches_single_stream(engine_lib):
    n0, p0, img0, vis0, blocks0, cyc0 = run(False)
    n1, p1, img1, vis1, blocks1, cyc1 = run(True)
    assert cyc0 == cyc1 == 3 and n0 > 1000
    assert np.array_equal(vis0, vis1) and blocks0 == blocks1            # the TSDF side is bit-identical
    assert n0 == n1, "Gaussian count %d vs %d" % (n0, n1)
    assert np.array_equal(p0["means"].shape, p1["means"].shape)
    # parameters: identical up to float-atomic summation order in the backward (a few ulp per step, amplified by Adam's sign
    # sensitivity on ~0 gradients for a handful of values)
    for k in p0:
        d = np.abs(p0[k].reshape(n0, -1) - p1[k].reshape(n1, -1))
        assert np.median(d) < 1e-6 and (d > 1e-3).mean() < 1e-3, k
    assert np.abs(img0 - img1).mean() < 1e-4
