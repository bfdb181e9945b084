"""shared synthetic inputs for the gsplat-path tests"""
import numpy as np


def random_splats(n, seed=0, spread=1.0, cam_z=3.0, scale_lo=0.003, scale_hi=0.03):
    """Gaussians in front of an identity-rotation camera placed at the origin looking down +z"""
    rng = np.random.RandomState(seed)
    means = np.stack([rng.uniform(-1.6, 1.6, n) * spread, rng.uniform(-0.9, 0.9, n) * spread, rng.uniform(0.8, cam_z, n)], 1).astype(np.float32)
    scales = np.log(np.exp(rng.uniform(np.log(scale_lo), np.log(scale_hi), (n, 3)))).astype(np.float32)
    quats = rng.normal(size=(n, 4)).astype(np.float32)
    dc = rng.normal(scale=0.5, size=(n, 3)).astype(np.float32)
    rest = rng.normal(scale=0.1, size=(n, 15, 3)).astype(np.float32)
    opac = rng.normal(loc=0.0, scale=1.5, size=(n, 1)).astype(np.float32)
    return dict(means=means, scales=scales, quats=quats, featuresDc=dc, featuresRest=rest, opacities=opac)


def camera(W, H, seed=0):
    rng = np.random.RandomState(100 + seed)
    # small rotation + translation so that viewmat is a general rigid transform
    a = rng.uniform(-0.08, 0.08, 3)
    Rx = np.array([[1, 0, 0], [0, np.cos(a[0]), -np.sin(a[0])], [0, np.sin(a[0]), np.cos(a[0])]])
    Ry = np.array([[np.cos(a[1]), 0, np.sin(a[1])], [0, 1, 0], [-np.sin(a[1]), 0, np.cos(a[1])]])
    Rz = np.array([[np.cos(a[2]), -np.sin(a[2]), 0], [np.sin(a[2]), np.cos(a[2]), 0], [0, 0, 1]])
    c2w = np.eye(4)
    c2w[:3, :3] = Rz @ Ry @ Rx
    c2w[:3, 3] = rng.uniform(-0.1, 0.1, 3)
    f = 0.5 * W
    K = np.array([[f, 0, (W - 1) / 2.0], [0, f, (H - 1) / 2.0], [0, 0, 1]], np.float32)
    return c2w.astype(np.float32), K


def scene_images(W, H, seed=0):
    rng = np.random.RandomState(200 + seed)
    yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    ref_depth = (2.0 + 0.6 * np.sin(xx / 37.0) * np.cos(yy / 23.0)).astype(np.float32)
    ref_depth[rng.uniform(size=(H, W)) < 0.03] = 0.0          # holes in the TSDF raycast
    base = rng.uniform(0.1, 0.9, size=(H, W, 3)).astype(np.float32)
    gt = np.clip(base + rng.normal(scale=0.15, size=(H, W, 3)), 0, 1).astype(np.float32)
    return ref_depth, base, gt
