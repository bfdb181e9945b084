"""render_method "raw" (depth-sorted front-to-back 3DGS compositing) and the fused SSIM map -- SURVEY.md section 8(f) row 3 --
against the reference functions they replace (the reference's own kernels, oracle/_ref/libgsplat_ref.so) on identical inputs:
isectTiles / isectOffsetEncode (bit-exact), RasterizeToPixels forward + backward, FusedSSIMMap forward + backward."""
import numpy as np
import pytest
import torch

from tests import gs_checks as gc
from tests.test_gs_staged_gpu import n, scene

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def gsref():
    from oracle import gsplat_ref
    if not gsplat_ref.available():
        pytest.skip("oracle/_ref/libgsplat_ref.so not built (oracle/gsplat_ref/Makefile needs /root/reference at build time)")
    return gsplat_ref.ops()


@pytest.mark.parametrize("N,W,H,seed,kw,bg", [(1500, 320, 192, 7, {}, False), (3000, 400, 300, 9, dict(scale_lo=0.01, scale_hi=0.08), True),
                                               (30000, 1200, 680, 23, dict(scale_lo=0.004, scale_hi=0.02), False)])
def test_raw_rasteriser_matches_reference(engine_lib, gsref, N, W, H, seed, kw, bg):
    from gps_slam_b200.gsplat_ops import GsplatOps
    s = scene(N, W, H, seed, **kw)
    ops = GsplatOps(W, H, capacity=N)
    g = torch.Generator(device=DEV).manual_seed(seed)
    try:
        radii, m2d, depths, conics = gsref.fully_fused_projection(s["means"], s["quats"], s["scales"], s["viewmat"], s["K"], W, H, 0.3, 0.01, 1e10, 0.0)[:4]
        radii = torch.clamp_max(radii, 100)
        vis = n(radii)[0] > 0
        tw, th = -(-W // 16), -(-H // 16)
        # ---- depth-sorted bins: bit-exact
        r_tpg, r_isect, r_flat = gsref.isect_tiles(m2d, radii, depths, 16, tw, th)
        r_off = gsref.isect_offset_encode(r_isect, 1, tw, th)
        tpg, isect, flat, off = ops.isect_tiles(m2d, radii, depths)
        assert np.array_equal(n(tpg), n(r_tpg)) and np.array_equal(n(isect), n(r_isect))
        assert np.array_equal(n(flat), n(r_flat)) and np.array_equal(n(off), n(r_off))
        # ---- forward
        colors4 = torch.cat([torch.rand(1, N, 3, device=DEV, generator=g), depths[..., None]], 2).contiguous()
        opac = s["opac"].reshape(-1).contiguous()
        ins = [x.detach().clone().requires_grad_(True) for x in (m2d, conics, colors4, opac)]
        bg_t = torch.tensor([[0.3, 0.5, 0.7, 0.0]], device=DEV) if bg else None
        if bg:
            r_render, r_alpha = gsref.rasterize_raw_bg(ins[0], ins[1], ins[2], ins[3], bg_t, W, H, 16, r_off, r_flat, False)
        else:
            r_render, r_alpha = gsref.rasterize_raw(ins[0], ins[1], ins[2], ins[3], W, H, 16, r_off, r_flat, False)
        render, alpha, last = ops.rasterize_to_pixels_fwd(m2d, conics, colors4, opac, bg_t, r_off, r_flat)
        gc.close_frac("render", n(render), n(r_render), 3e-4, 3e-4, 3e-4)
        gc.close_frac("alphas", n(alpha), n(r_alpha), 3e-4, 3e-4, 3e-4)
        # ---- backward
        v_render, v_alpha = torch.randn(1, H, W, 4, device=DEV, generator=g) * 1e-3, torch.randn(1, H, W, 1, device=DEV, generator=g) * 1e-3
        # RasterizeToPixels::backward never hands the background to the kernel (gsplat_wapper.hpp:300-312 there: a fresh empty optional),
        # so the reference's autograd gives the gradients of the background-free composite: compared with ours WITHOUT a background
        rv = torch.autograd.grad([r_render, r_alpha], ins, [v_render, v_alpha])
        mine = ops.rasterize_to_pixels_bwd(m2d, conics, colors4, opac, None, r_off, r_flat, alpha, last, v_render, v_alpha)
        for name, a, b in zip(("v_means2d", "v_conics", "v_colors", "v_opacities"), mine, rv):
            gc.close_scaled(name, n(a).reshape(N, -1)[vis], n(b).reshape(N, -1)[vis], 3e-3, 2e-4)
        if bg:
            # ... and the kernel itself given the background (rasterize_to_pixels_bwd_tensor), against ours WITH it
            r_last = gsref.rasterize_raw_fwd(m2d, conics, colors4, opac, bg_t, W, H, 16, r_off, r_flat)[2]
            rk = gsref.rasterize_raw_bwd(m2d, conics, colors4, opac, bg_t, W, H, 16, r_off, r_flat, r_alpha.detach(), r_last, v_render, v_alpha)
            mine_bg = ops.rasterize_to_pixels_bwd(m2d, conics, colors4, opac, bg_t, r_off, r_flat, alpha, last, v_render, v_alpha)
            for name, a, b in zip(("v_means2d", "v_conics", "v_colors", "v_opacities"), mine_bg, rk):
                gc.close_scaled(name + " (background)", n(a).reshape(N, -1)[vis], n(b).reshape(N, -1)[vis], 3e-3, 2e-4)
            assert np.abs(n(mine_bg[0]) - n(mine[0])).max() > 0      # the term is not a no-op on this scene
    finally:
        ops.close()


@pytest.mark.parametrize("H,W", [(68, 120), (340, 600), (680, 1200)])
def test_fused_ssim_matches_reference(engine_lib, gsref, H, W):
    from gps_slam_b200.gsplat_ops import GsplatOps
    ops = GsplatOps(W, H, capacity=128)
    try:
        g = torch.Generator(device=DEV).manual_seed(H)
        img2 = torch.rand(1, 3, H, W, device=DEV, generator=g)
        img1 = (img2 + 0.1 * torch.randn(1, 3, H, W, device=DEV, generator=g)).clamp(0, 1)
        C1, C2 = 0.01 ** 2, 0.03 ** 2
        a = img1.clone().requires_grad_(True)
        for padding in ("same", "valid"):
            r_map = gsref.fused_ssim_map(C1, C2, a, img2, padding, True)
            m, d1, d2, d3 = ops.fusedssim(C1, C2, img1, img2, True)
            mine = m[:, :, 5:-5, 5:-5] if padding == "valid" else m
            np.testing.assert_allclose(n(mine), n(r_map), rtol=2e-4, atol=2e-5)
            dL = torch.randn(r_map.shape, device=DEV, generator=g)
            (r_grad,) = torch.autograd.grad([r_map], [a], [dL])
            dL_full = dL
            if padding == "valid":
                dL_full = torch.zeros_like(img1)
                dL_full[:, :, 5:-5, 5:-5] = dL
            grad = ops.fusedssim_backward(C1, C2, img1, img2, dL_full, d1, d2, d3)
            gc.close_scaled("dL_dimg1 " + padding, n(grad), n(r_grad), 2e-4)
    finally:
        ops.close()
