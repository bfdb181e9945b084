"""Staged C-ABI entry points (gsb_gs_projection_fwd/_bwd, gsb_gs_sh_fwd/_bwd, gsb_gs_isect_tiles, gsb_gs_rasterize_ges_fwd/_bwd,
gsb_gs_adam_step) against the reference function each one replaces, called on IDENTICAL device inputs: the reference's own
gsplat kernels / autograd wrappers compiled for sm_100a (oracle/_ref/libgsplat_ref.so), and torch.optim.Adam for the optimiser.
Integer outputs bit-exact, floats within the bars of tests/gs_checks.py."""
import numpy as np
import pytest
import torch

from tests import gs_checks as gc
from tests.helpers_gs import camera, random_splats, scene_images

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def gsref():
    from oracle import gsplat_ref
    if not gsplat_ref.available():
        pytest.skip("oracle/_ref/libgsplat_ref.so not built (oracle/gsplat_ref/Makefile needs /root/reference at build time)")
    return gsplat_ref.ops()


def scene(N, W, H, seed, **kw):
    p = random_splats(N, seed=seed, **kw)
    c2w, K = camera(W, H, seed)
    ref_depth, base, gt = scene_images(W, H, seed)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(DEV)
    from oracle.gsplat_ref import pose_inv
    c2w_t = t(c2w)
    d = dict(means=t(p["means"]), quats=t(p["quats"]), scales=torch.exp(t(p["scales"])), opac=torch.sigmoid(t(p["opacities"])),
             shs=torch.cat([t(p["featuresDc"])[:, None, :], t(p["featuresRest"])], 1).contiguous(), viewmat=pose_inv(c2w_t)[None], K=t(K)[None],
             cam_t=c2w_t[:3, 3], ref_depth=torch.where(t(ref_depth) < 0.01, torch.full_like(t(ref_depth), 1000.0), t(ref_depth)), base=t(base))
    return d


def n(t):
    return t.detach().cpu().numpy()


CASES = [(1500, 320, 192, 7, {}), (200, 400, 300, 5, dict(scale_lo=0.05, scale_hi=0.4)), (30000, 1200, 680, 23, dict(scale_lo=0.004, scale_hi=0.02))]


@pytest.mark.parametrize("N,W,H,seed,kw", CASES)
def test_staged_ops_match_reference_functions(engine_lib, gsref, N, W, H, seed, kw):
    from gps_slam_b200.gsplat_ops import GsplatOps
    s = scene(N, W, H, seed, **kw)
    ops = GsplatOps(W, H, capacity=N)
    g = torch.Generator(device=DEV).manual_seed(seed)
    rnd = lambda *shape: torch.randn(*shape, device=DEV, generator=g)
    try:
        # ---- projection forward
        means, quats, scales = [x.clone().requires_grad_(True) for x in (s["means"], s["quats"], s["scales"])]
        r_radii, r_m2d, r_depths, r_conics = gsref.fully_fused_projection(means, quats, scales, s["viewmat"], s["K"], W, H, 0.3, 0.01, 1e10, 0.0)[:4]
        radii, m2d, depths, conics = ops.fully_fused_projection_fwd(s["means"], s["quats"], s["scales"], s["viewmat"], s["K"])
        odd = np.nonzero(n(radii)[0] != n(r_radii)[0])[0]
        assert len(odd) <= max(1, int(1e-4 * N)) and np.all(np.abs(n(radii)[0][odd] - n(r_radii)[0][odd]) == 1)
        vis = (n(r_radii)[0] > 0) & (n(radii)[0] > 0)
        gc.close_frac("means2d", n(m2d)[0][vis], n(r_m2d)[0][vis], 2e-4, 2e-6)
        gc.close_frac("depths", n(depths)[0][vis], n(r_depths)[0][vis], 2e-6, 2e-6)
        gc.close_frac("conics", n(conics)[0][vis], n(r_conics)[0][vis], 2e-6, 5e-5)
        # ---- projection backward (reference: autograd through FullyFusedProjection::backward)
        v_m2d, v_depths, v_conics = rnd(1, N, 2), rnd(1, N), rnd(1, N, 3)
        rv_means, rv_quats, rv_scales = torch.autograd.grad([r_m2d, r_depths, r_conics], [means, quats, scales], [v_m2d, v_depths, v_conics])
        v_means, v_quats, v_scales = ops.fully_fused_projection_bwd(s["means"], s["quats"], s["scales"], s["viewmat"], s["K"], r_radii,
                                                                     r_conics.detach(), v_m2d, v_depths, v_conics)
        same = np.ones(N, bool)
        same[odd] = False
        for name, a, b in (("v_means", v_means, rv_means), ("v_quats", v_quats, rv_quats), ("v_scales", v_scales, rv_scales)):
            gc.close_scaled(name, n(a)[vis & same], n(b)[vis & same], 3e-3, 1e-4)
        # ---- SH forward / backward
        rad_c = torch.clamp_max(r_radii, 100)
        mask = rad_c > 0
        dirs = (s["means"] - s["cam_t"][None, :])[None]
        shs = s["shs"][None].clone().requires_grad_(True)
        dirs_r = dirs.clone().requires_grad_(True)
        r_col = gsref.spherical_harmonics(3, dirs_r, shs, mask)
        col = ops.compute_sh_fwd(3, dirs, s["shs"][None], mask)
        gc.close_frac("sh colors", n(col)[0][vis], n(r_col)[0][vis], 5e-6, 5e-5)
        v_col = rnd(1, N, 3)
        rv_shs, rv_dirs = torch.autograd.grad([r_col], [shs, dirs_r], [v_col])
        v_shs, v_dirs = ops.compute_sh_bwd(3, dirs, s["shs"][None], mask, v_col)
        gc.close_scaled("v_coeffs", n(v_shs)[0][vis], n(rv_shs)[0][vis], 1e-4)
        gc.close_scaled("v_dirs", n(v_dirs)[0][vis], n(rv_dirs)[0][vis], 2e-3, 1e-4)
        assert float(n(v_shs)[0][~n(mask)[0]].__abs__().max(initial=0.0)) == 0.0
        # ---- binning: the reference's own projection outputs in, integers out -> bit-exact
        tw, th = -(-W // 16), -(-H // 16)
        r_tpg, r_isect, r_flat, _, _ = gsref.isect_tiles_no_depth(r_m2d.detach(), rad_c, r_depths.detach(), 16, tw, th)
        r_off = gsref.isect_offset_encode_no_depth(r_isect, 1, tw, th)
        tpg, isect, flat, off = ops.isect_tiles_no_depth(r_m2d.detach(), rad_c)
        assert np.array_equal(n(tpg), n(r_tpg)) and np.array_equal(n(isect), n(r_isect))
        assert np.array_equal(n(flat), n(r_flat)) and np.array_equal(n(off), n(r_off))
        # ---- rasteriser forward on the reference's bins
        colors4 = torch.cat([torch.clamp_min(r_col.detach() + 0.5, 0.0), r_depths.detach()[..., None]], 2)
        ref_depth = s["ref_depth"].reshape(1, H, W, 1)
        ins = [x.detach().clone().requires_grad_(True) for x in (r_m2d, r_conics, colors4, s["opac"])]
        r_tpg2, r_isect2, r_flat2, r_ggs, r_gst = gsref.isect_tiles_no_depth(r_m2d.detach(), rad_c, r_depths.detach(), 16, tw, th)
        r_render, r_alpha = gsref.rasterize_ges(ins[0], ins[1], ins[2], ins[3], rad_c, ref_depth, s["base"].reshape(1, H, W, 3), W, H, 16, r_off, r_flat,
                                                r_ggs, r_gst, False, 0.1)
        render, alpha = ops.rasterize_to_pixels_fwd_ges(r_m2d.detach(), r_conics.detach(), colors4, s["opac"].reshape(-1), ref_depth, 0.1, r_off, r_flat)
        gc.close_frac("render", n(render), n(r_render), 2e-4, 2e-4, 2e-4)
        gc.close_frac("alphas", n(alpha), n(r_alpha), 2e-4, 2e-4, 2e-4)
        # ---- rasteriser backward
        v_render, v_alpha = rnd(1, H, W, 4) * 1e-3, rnd(1, H, W, 1) * 1e-3
        rv = torch.autograd.grad([r_render, r_alpha], ins, [v_render, v_alpha])
        mine = ops.rasterize_to_pixels_bwd_ges(r_m2d.detach(), r_conics.detach(), colors4, s["opac"].reshape(-1), rad_c, ref_depth, 0.1, v_render, v_alpha)
        for name, a, b in zip(("v_means2d", "v_conics", "v_colors", "v_opacities"), mine, rv):
            gc.close_scaled(name, n(a).reshape(N, -1)[vis], n(b).reshape(N, -1)[vis], 2e-3, 1e-4)
    finally:
        ops.close()


def test_adam_step_matches_torch_optim(engine_lib):
    from gps_slam_b200.gsplat_ops import GsplatOps
    ops = GsplatOps(64, 64, capacity=128)
    try:
        g = torch.Generator(device=DEV).manual_seed(3)
        p0 = torch.randn(50000, device=DEV, generator=g)
        ref = p0.clone().requires_grad_(True)
        opt = torch.optim.Adam([ref], lr=5e-3, eps=1e-15)
        mine, m, v = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
        for step in range(1, 6):
            grad = torch.randn(50000, device=DEV, generator=g) * (10.0 ** float(torch.randint(-8, 0, (1,)).item()))
            ref.grad = grad.clone()
            opt.step()
            ops.adam_step(mine, grad, m, v, 5e-3, step)
            np.testing.assert_allclose(mine.cpu().numpy(), ref.detach().cpu().numpy(), rtol=2e-6, atol=1e-8)
    finally:
        ops.close()
