"""GPU parity of the glue between the TSDF engine and the Gaussian model and of the Gaussian spawn
against oracle/slam_glue.py, the torch-CPU restatement of the reference's tensor code (runRaycastByCam, ITMU*ImageToTensor,
initNewGaussians, computeNormalMap, addGaussians, RawGaussianParams::init).

  * gsb_gs_frame_to_float: bit-exact (one IEEE division per value);
  * gsb_gs_raycast_maps: colour / confidence bit-exact, depth within 2e-6 m (the reference multiplies a 4x4 into all vertices with a
    GEMM whose summation order is libtorch's; ours is a fixed left-to-right dot product);
  * gsb_gs_spawn: the selected pixels are a subset of the reference's mask of the right size (the reference draws a randperm prefix,
    we keep each masked pixel with probability ratio -- a documented deviation), and GIVEN the selected pixels every parameter of the
    new Gaussians follows the reference's init: means exact, scales (3-NN, clamp, z x 0.1, log) / DC / opacity to fp32 rounding, the rotation through the axis it gives the Gaussian;
  * world = 2: the two ranks' spawns partition the world = 1 spawn and carry the same parameters (KNN over every rank's points).
"""
import numpy as np
import pytest
import torch

from gps_slam_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

CFG = dict(color_error_thres=0.05, depth_vis_min=0.0, depth_vis_max=5.0, alpha_vis_max=5.0, sample_ratio=0.25, max_init_scale=0.01,
           min_init_scale=-1.0, default_opacity=0.5)


def fused_scene(intr, n_frames=6):
    """TSDF engine after n_frames, a free-view raycast at the last pose; returns (engine, poses, frames, host copies of the free images)"""
    from gps_slam_b200.engine import TsdfEngine
    poses, frames = syn.sequence(n_frames, intr)
    eng = TsdfEngine(intr, tracker=0)
    for i in range(n_frames):
        eng.ProcessFrame(frames[i][0].numpy(), frames[i][1].numpy(), syn.c2w_to_colmajor(poses[i]))
    eng.runRaycast(syn.c2w_to_colmajor(poses[-1]), intr)
    eng.sync()
    return eng, poses, frames


def device_maps(gs, tsdf, c2w, H, W, dev):
    depth_map = torch.empty((H, W), device=dev)
    color_map = torch.empty((H, W, 3), device=dev)
    conf_map = torch.empty((H, W), device=dev)
    gs.raycast_maps(tsdf.GetFreeVertex(), tsdf.GetFreeImage(), c2w, tsdf.getVoxelSize(), depth_map, color_map, conf_map)
    gs.sync()
    return depth_map, color_map, conf_map


def test_frame_to_float_and_raycast_maps(engine_lib):
    from gps_slam_b200.engine import GaussianEngine
    from oracle import slam_glue as sg
    intr = syn.intrinsics("replica", 0.5)
    H, W = intr["height"], intr["width"]
    dev = torch.device("cuda", 0)
    tsdf, poses, frames = fused_scene(intr)
    gs = GaussianEngine(W, H, capacity=1024)
    try:
        # f2: Camera::image / depth as float
        rgb_d, dep_d = torch.empty((H, W, 3), device=dev), torch.empty((H, W), device=dev)
        gs.frame_to_float(frames[-1][0].to(dev), frames[-1][1].to(dev), rgb_d, dep_d)
        gs.sync()
        rgb_o, dep_o = sg.frame_to_float(frames[-1][0].numpy(), frames[-1][1].numpy())
        assert np.array_equal(rgb_d.cpu().numpy().view(np.uint32), rgb_o.numpy().view(np.uint32))
        assert np.array_equal(dep_d.cpu().numpy().view(np.uint32), dep_o.numpy().view(np.uint32))
        # f1: runRaycastByCam's tensors, at the last pose and at a pose that differs from the raycast pose (cam.c2w != c2w_slam)
        vert = tsdf.raycast(live=False)
        img = tsdf.free_image()
        assert (vert[..., 3] > 0).mean() > 0.5
        for c2w in (poses[-1], poses[0]):
            depth_map, color_map, conf_map = device_maps(gs, tsdf, c2w, H, W, dev)
            o = sg.raycast_maps(vert, img, c2w, tsdf.getVoxelSize())
            assert np.array_equal(color_map.cpu().numpy().view(np.uint32), o["color_map"].numpy().view(np.uint32)), "color_map"
            assert np.array_equal(conf_map.cpu().numpy().view(np.uint32), o["confidence_map"][..., 0].numpy().view(np.uint32)), "confidence"
            d_o = o["depth_map"][..., 0].numpy()
            d_g = depth_map.cpu().numpy()
            assert np.array_equal(d_g == 0, d_o == 0), "invalid-vertex mask"
            assert np.abs(d_g - d_o).max() < 2e-6, np.abs(d_g - d_o).max()
    finally:
        gs.close()
        tsdf.close()


def check_spawn_against_oracle(gs, n_before, n_new, pix, vertex_map, image, normal_map, mask):
    from oracle import slam_glue as sg
    # the sampled set: inside the reference's mask, raster order, about ratio * |mask|
    M = int(mask.sum())
    assert n_new == len(pix) and np.all(np.diff(pix) > 0)
    assert mask.reshape(-1)[pix].all(), "a pixel outside the reference's sample mask was spawned"
    exp, sd = CFG["sample_ratio"] * M, np.sqrt(M * CFG["sample_ratio"] * (1 - CFG["sample_ratio"]))
    assert abs(n_new - exp) < 6 * sd + 1, "sampled %d of %d masked pixels (expected %.0f +- %.0f)" % (n_new, M, exp, sd)
    sel = torch.from_numpy(pix.astype(np.int64))
    o = sg.init_params(vertex_map.reshape(-1, 3)[sel], image.reshape(-1, 3)[sel], normal_map.reshape(-1, 3)[sel], CFG["default_opacity"],
                       CFG["max_init_scale"], CFG["min_init_scale"])
    got = gs.get_params()
    new = {k: v[n_before:n_before + n_new] for k, v in got.items()}
    assert np.array_equal(new["means"].view(np.uint32), o["means"].view(np.uint32)), "means"
    assert np.abs(new["scales"] - o["scales"]).max() < 2e-5, ("log scales", np.abs(new["scales"] - o["scales"]).max())
    # the KNN clamp must bite for some and not for others, or the test would not see the scale path
    s0 = np.exp(o["scales"][:, 0])
    assert (s0 < CFG["max_init_scale"] * 0.999).any() and (s0 >= CFG["max_init_scale"] * 0.999).any()
    assert np.allclose(np.exp(new["scales"][:, 2]), 0.1 * np.exp(new["scales"][:, 0]), rtol=1e-5)
    # The quaternion turns the z axis onto the surface normal (computeQuat).  The normal comes from Sobel differences of millimetre-spaced
    # vertices that are metres from the origin (relative rounding ~1e-4, summation order of conv2d unspecified), and acos / the axis
    # normalisation amplify that without bound where the normal is (anti)parallel to z -- there even the reference's own result is
    # decided by rounding.  What the new Gaussian depends on is its covariance R diag(s^2, s^2, (0.1 s)^2) R^T, i.e. only R z: compare that,
    # and the quaternions themselves where the construction is well conditioned.
    def rz(q):
        q = q / np.linalg.norm(q, axis=1, keepdims=True)   # (gsplat normalises the quaternion inside the projection)
        w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
        return np.stack([2 * (x * z + w * y), 2 * (y * z - w * x), 1 - 2 * (x * x + y * y)], 1)
    # (a zero normal -- the reference zeroes it where the WORLD z of the vertex is <= 0 -- gives the non-unit (cos 45deg, 0, 0, 0) in both)
    assert np.abs(np.linalg.norm(new["quats"], axis=1) - np.linalg.norm(o["quats"], axis=1)).max() < 1e-5
    cosang = np.clip((rz(new["quats"]) * rz(o["quats"])).sum(1), -1, 1)
    assert np.degrees(np.arccos(cosang)).max() < 0.3, ("rotated z axis", np.degrees(np.arccos(cosang)).max())
    well = np.abs(rz(o["quats"])[:, 2]) < 0.9
    assert well.sum() > 0.2 * len(well)
    dq = np.abs(new["quats"] - o["quats"])[well]
    assert dq.max() < 5e-4, ("quats", dq.max())
    assert np.abs(new["featuresDc"] - o["featuresDc"]).max() < 1e-6, "featuresDc"
    assert not new["featuresRest"].any()
    assert np.abs(new["opacities"] - o["opacities"]).max() < 1e-6, "opacities"


def test_spawn_matches_reference_init(engine_lib):
    """cycle 1: empty model (mask from the TSDF colour error); cycle 2: mask from the render of the cycle-1 Gaussians"""
    from gps_slam_b200.engine import GaussianEngine
    from oracle import slam_glue as sg
    intr = syn.intrinsics("replica", 0.5)
    H, W = intr["height"], intr["width"]
    dev = torch.device("cuda", 0)
    tsdf, poses, frames = fused_scene(intr)
    gs = GaussianEngine(W, H, capacity=1 << 18)
    try:
        c2w = poses[-1]
        image = torch.empty((H, W, 3), device=dev)
        gs.frame_to_float(frames[-1][0].to(dev), None, image, None)
        depth_map, color_map, _ = device_maps(gs, tsdf, c2w, H, W, dev)
        o = sg.raycast_maps(tsdf.raycast(live=False), tsdf.free_image(), c2w, tsdf.getVoxelSize())
        normal = sg.compute_normal_map(o["vertex_map"])
        img_o = image.cpu()
        # ---- empty model
        mask = sg.sample_mask(o, img_o, None, None, CFG["color_error_thres"], CFG["depth_vis_min"], CFG["depth_vis_max"], CFG["alpha_vis_max"]).numpy()
        gs.addGaussians(c2w, intr, tsdf.GetFreeVertex(), tsdf.getVoxelSize(), depth_map, color_map, image, seed=1234, **CFG)
        n1 = gs.getGaussianNum()
        assert n1 > 1000
        check_spawn_against_oracle(gs, 0, n1, gs.spawn_pixels(n1), o["vertex_map"], img_o, normal, mask)
        # ---- with Gaussians: the mask needs the render of the current model (gesForward), taken from the engine itself
        rgb, dep, alpha = torch.empty((H, W, 3), device=dev), torch.empty((H, W), device=dev), torch.empty((H, W), device=dev)
        gs.forward(c2w, intr, depth_map, color_map, rgb, dep, alpha)
        gs.sync()
        mask2 = sg.sample_mask(o, img_o, rgb.cpu(), alpha.cpu(), CFG["color_error_thres"], CFG["depth_vis_min"], CFG["depth_vis_max"],
                               CFG["alpha_vis_max"]).numpy()
        gs.addGaussians(c2w, intr, tsdf.GetFreeVertex(), tsdf.getVoxelSize(), depth_map, color_map, image, seed=99, **CFG)
        n2 = gs.getGaussianNum() - n1
        assert n2 > 100
        check_spawn_against_oracle(gs, n1, n2, gs.spawn_pixels(n2), o["vertex_map"], img_o, normal, mask2)
    finally:
        gs.close()
        tsdf.close()


def test_spawn_sharded_equals_single(engine_lib):
    """world = 2 on one device: the ranks' new Gaussians partition the world = 1 set (ownership by 4 cm block) with identical
    parameters -- in particular the 3-NN scale of a Gaussian whose neighbours belong to the other rank"""
    from gps_slam_b200 import parallel
    from gps_slam_b200.engine import GaussianEngine
    intr = syn.intrinsics("replica", 0.5)
    H, W = intr["height"], intr["width"]
    dev = torch.device("cuda", 0)
    tsdf, poses, frames = fused_scene(intr)
    engines = [GaussianEngine(W, H, capacity=1 << 18) for _ in range(3)]
    try:
        c2w = poses[-1]
        image = torch.empty((H, W, 3), device=dev)
        engines[0].frame_to_float(frames[-1][0].to(dev), None, image, None)
        depth_map, color_map, _ = device_maps(engines[0], tsdf, c2w, H, W, dev)
        # the render of the (empty) model, as the multi-GPU path supplies it: TSDF colour, zero weight
        alpha0 = torch.zeros((H, W), device=dev)
        engines[0].addGaussians(c2w, intr, tsdf.GetFreeVertex(), tsdf.getVoxelSize(), depth_map, color_map, image, seed=7, **CFG)
        n = engines[0].getGaussianNum()
        single = engines[0].get_params()
        pix = engines[0].spawn_pixels(n)
        parts = []
        for r in (0, 1):
            e = engines[1 + r]
            e.addGaussians(c2w, intr, tsdf.GetFreeVertex(), tsdf.getVoxelSize(), depth_map, color_map, image, seed=7, rank=r, world=2,
                           render_rgb=color_map, render_alpha=alpha0, **CFG)
            k = e.getGaussianNum()
            parts.append((e.spawn_pixels(k), e.get_params()))
        assert len(parts[0][0]) + len(parts[1][0]) == n and min(len(parts[0][0]), len(parts[1][0])) > 0.3 * n
        own = parallel.owner_of(single["means"], 2)
        for r in (0, 1):
            assert np.array_equal(parts[r][0], pix[own == r]), "rank %d spawned a different pixel set" % r
            for k in single:
                assert np.array_equal(parts[r][1][k], single[k][own == r]), "rank %d: %s differs from the single-GPU spawn" % (r, k)
    finally:
        for e in engines:
            e.close()
        tsdf.close()
