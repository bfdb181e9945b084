"""Python mirror of the gsplat function layer the reference's Gaussian model calls (reference gsplat/gsplat_wapper.hpp:14-709 and
the gsplat::*_tensor functions of gsplat/rasterizer/bindings.h underneath), bound to the staged C-ABI entry points of
libgpsslam_b200.so.  Same names, argument order and tensor shapes as the reference functions (camera dimension C = 1 explicit),
so the parity tests read like calls into the reference.  torch is used for device memory only: every function allocates its
outputs as torch CUDA tensors and hands raw pointers to the library; there is no fallback path.

    ops = GsplatOps(width, height)               # one engine = workspace + stream
    radii, means2d, depths, conics = ops.fully_fused_projection_fwd(means, quats, scales, viewmats, Ks)
"""
import ctypes as C

import torch

from . import engine as E


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _f32(t):
    assert t.is_cuda and t.dtype == torch.float32
    return t.contiguous()


class GsplatOps:
    def __init__(self, width, height, capacity=1 << 20, device=0, **overrides):
        self.eng = E.GaussianEngine(width, height, capacity=capacity, device=device, **overrides)
        self.L, self.h = self.eng.L, self.eng.h_
        self.W, self.H = width, height
        self.tile_w, self.tile_h = self.eng.tile_w, self.eng.tile_h
        self.dev = torch.device("cuda", device)
        # run on torch's current stream so that tensor lifetimes and ordering follow the caller's
        # (torch's default stream has handle 0, which the C ABI reads as "use the engine's private stream": pass cudaStreamLegacy = 1)
        self.eng.set_stream(torch.cuda.current_stream(self.dev).cuda_stream or 1)
        L = self.L
        i, f, p = C.c_int, C.c_float, C.c_void_p
        L.gsb_gs_projection_fwd.argtypes = [p, i] + [p] * 5 + [i] + [p] * 4
        L.gsb_gs_projection_bwd.argtypes = [p, i] + [p] * 13
        L.gsb_gs_sh_fwd.argtypes = [p, i, i, p, p, p, p]
        L.gsb_gs_sh_bwd.argtypes = [p, i, i, p, p, p, p, p, p]
        L.gsb_gs_isect_tiles.argtypes = [p, i, p, p, p, C.POINTER(C.c_int)]
        L.gsb_gs_isect_fetch.argtypes = [p, i, p, p, p]
        L.gsb_gs_rasterize_ges_fwd.argtypes = [p, i, p, p, p, p, p, f, p, p, i, p, p]
        L.gsb_gs_rasterize_ges_bwd.argtypes = [p, i, p, p, p, p, p, p, f, p, p, p, p, p, p]
        L.gsb_gs_adam_step.argtypes = [p, C.c_longlong, p, p, p, p, f, f, f, f, i]
        L.gsb_gs_isect_tiles_depth.argtypes = [p, i, p, p, p, p, C.POINTER(C.c_int)]
        L.gsb_gs_isect_fetch_depth.argtypes = [p, i, p, p, p]
        L.gsb_gs_rasterize_fwd.argtypes = [p, i, p, p, p, p, p, p, p, i, p, p, p]
        L.gsb_gs_rasterize_bwd.argtypes = [p, i, p, p, p, p, p, p, p, i, p, p, p, p, p, p, p, p]
        L.gsb_gs_ssim_fwd.argtypes = [p, i, i, i, f, f, p, p, p, p, p, p]
        L.gsb_gs_ssim_bwd.argtypes = [p, i, i, i, p, p, p, p, p, p, p]

    def close(self):
        self.eng.close()

    # gsplat::fully_fused_projection_fwd_tensor (+ clamp_max(radii, max_gs_radii) when clamp_radii > 0)
    def fully_fused_projection_fwd(self, means, quats, scales, viewmats, Ks, clamp_radii=0):
        n = means.shape[0]
        radii = torch.empty((1, n), dtype=torch.int32, device=self.dev)
        means2d = torch.empty((1, n, 2), device=self.dev)
        depths = torch.empty((1, n), device=self.dev)
        conics = torch.empty((1, n, 3), device=self.dev)
        vm, K = viewmats.reshape(4, 4).float().cpu().contiguous(), Ks.reshape(3, 3).float().cpu().contiguous()
        E._check(self.L.gsb_gs_projection_fwd(self.h, n, _p(_f32(means)), _p(_f32(quats)), _p(_f32(scales)), C.c_void_p(vm.data_ptr()),
                                              C.c_void_p(K.data_ptr()), clamp_radii, _p(radii), _p(means2d), _p(depths), _p(conics)))
        return radii, means2d, depths, conics

    # gsplat::fully_fused_projection_bwd_tensor -> v_means, v_quats, v_scales
    def fully_fused_projection_bwd(self, means, quats, scales, viewmats, Ks, radii, conics, v_means2d, v_depths, v_conics):
        n = means.shape[0]
        v_means, v_quats, v_scales = torch.empty((n, 3), device=self.dev), torch.empty((n, 4), device=self.dev), torch.empty((n, 3), device=self.dev)
        vm, K = viewmats.reshape(4, 4).float().cpu().contiguous(), Ks.reshape(3, 3).float().cpu().contiguous()
        E._check(self.L.gsb_gs_projection_bwd(self.h, n, _p(_f32(means)), _p(_f32(quats)), _p(_f32(scales)), C.c_void_p(vm.data_ptr()),
                                              C.c_void_p(K.data_ptr()), _p(radii.contiguous()), _p(_f32(conics)), _p(_f32(v_means2d)),
                                              _p(_f32(v_depths)), _p(_f32(v_conics)), _p(v_means), _p(v_quats), _p(v_scales)))
        return v_means, v_quats, v_scales

    # gsplat::compute_sh_fwd_tensor
    def compute_sh_fwd(self, degrees_to_use, dirs, coeffs, masks=None):
        n = coeffs.shape[-3]
        colors = torch.empty(tuple(dirs.shape[:-1]) + (3,), device=self.dev)
        m = masks.to(torch.uint8).contiguous() if masks is not None else None
        E._check(self.L.gsb_gs_sh_fwd(self.h, n, degrees_to_use, _p(_f32(dirs)), _p(_f32(coeffs)), _p(m), _p(colors)))
        return colors

    # gsplat::compute_sh_bwd_tensor -> v_coeffs, v_dirs
    def compute_sh_bwd(self, degrees_to_use, dirs, coeffs, masks, v_colors, compute_v_dirs=True):
        n = coeffs.shape[-3]
        v_coeffs = torch.empty_like(coeffs)
        v_dirs = torch.empty_like(dirs) if compute_v_dirs else None
        m = masks.to(torch.uint8).contiguous() if masks is not None else None
        E._check(self.L.gsb_gs_sh_bwd(self.h, n, degrees_to_use, _p(_f32(dirs)), _p(_f32(coeffs)), _p(m), _p(_f32(v_colors)), _p(v_coeffs), _p(v_dirs)))
        return v_coeffs, v_dirs

    # isectTilesNoDepth + isectOffsetEncodeNoDepth (gsplat_wapper.cpp:58-93): tiles_per_gauss, isect_ids, flatten_ids, isect_offsets
    def isect_tiles_no_depth(self, means2d, radii):
        n = means2d.shape[-2]
        tpg = torch.empty((1, n), dtype=torch.int32, device=self.dev)
        cnt = C.c_int(0)
        E._check(self.L.gsb_gs_isect_tiles(self.h, n, _p(_f32(means2d)), _p(radii.contiguous()), _p(tpg), C.byref(cnt)))
        ni = cnt.value
        isect_ids = torch.empty((ni,), dtype=torch.int64, device=self.dev)
        flatten_ids = torch.empty((ni,), dtype=torch.int32, device=self.dev)
        offsets = torch.empty((1, self.tile_h, self.tile_w), dtype=torch.int32, device=self.dev)
        E._check(self.L.gsb_gs_isect_fetch(self.h, ni, _p(isect_ids), _p(flatten_ids), _p(offsets)))
        return tpg, isect_ids, flatten_ids, offsets

    # gsplat::rasterize_to_pixels_fwd_ges_tensor -> render_colors [1,H,W,4], render_alphas [1,H,W,1]
    def rasterize_to_pixels_fwd_ges(self, means2d, conics, colors, opacities, ref_depth_map, delta_depth, isect_offsets, flatten_ids):
        n = means2d.shape[-2]
        render = torch.empty((1, self.H, self.W, 4), device=self.dev)
        alphas = torch.empty((1, self.H, self.W, 1), device=self.dev)
        E._check(self.L.gsb_gs_rasterize_ges_fwd(self.h, n, _p(_f32(means2d)), _p(_f32(conics)), _p(_f32(colors)), _p(_f32(opacities)),
                                                 _p(_f32(ref_depth_map)), delta_depth, _p(isect_offsets.contiguous()), _p(flatten_ids.contiguous()),
                                                 flatten_ids.numel(), _p(render), _p(alphas)))
        return render, alphas

    # gsplat::rasterize_to_pixels_bwd_ges_gs_parallel_tensor -> v_means2d, v_conics, v_colors, v_opacities
    def rasterize_to_pixels_bwd_ges(self, means2d, conics, colors, opacities, radiis, ref_depth_map, delta_depth, v_render_colors, v_render_alphas):
        n = means2d.shape[-2]
        v_means2d, v_conics, v_colors = torch.empty_like(means2d), torch.empty_like(conics), torch.empty_like(colors)
        v_opac = torch.empty_like(opacities)
        E._check(self.L.gsb_gs_rasterize_ges_bwd(self.h, n, _p(_f32(means2d)), _p(_f32(conics)), _p(_f32(colors)), _p(_f32(opacities)),
                                                 _p(radiis.contiguous()), _p(_f32(ref_depth_map)), delta_depth, _p(_f32(v_render_colors)),
                                                 _p(_f32(v_render_alphas)), _p(v_means2d), _p(v_conics), _p(v_colors), _p(v_opac)))
        return v_means2d, v_conics, v_colors, v_opac

    # torch::optim::Adam::step on one tensor, in place
    def adam_step(self, param, grad, exp_avg, exp_avg_sq, lr, step, beta1=0.9, beta2=0.999, eps=1e-15):
        E._check(self.L.gsb_gs_adam_step(self.h, param.numel(), _p(param), _p(_f32(grad)), _p(exp_avg), _p(exp_avg_sq), lr, beta1, beta2, eps, step))

    # ---- render_method "raw" + fused SSIM
    # isectTiles + isectOffsetEncode (gsplat_wapper.cpp:14-50): bins ordered by (tile, depth)
    def isect_tiles(self, means2d, radii, depths):
        n = means2d.shape[-2]
        tpg = torch.empty((1, n), dtype=torch.int32, device=self.dev)
        cnt = C.c_int(0)
        E._check(self.L.gsb_gs_isect_tiles_depth(self.h, n, _p(_f32(means2d)), _p(radii.contiguous()), _p(_f32(depths)), _p(tpg), C.byref(cnt)))
        ni = cnt.value
        isect_ids = torch.empty((ni,), dtype=torch.int64, device=self.dev)
        flatten_ids = torch.empty((ni,), dtype=torch.int32, device=self.dev)
        offsets = torch.empty((1, self.tile_h, self.tile_w), dtype=torch.int32, device=self.dev)
        E._check(self.L.gsb_gs_isect_fetch_depth(self.h, ni, _p(isect_ids), _p(flatten_ids), _p(offsets)))
        return tpg, isect_ids, flatten_ids, offsets

    # gsplat::rasterize_to_pixels_fwd_tensor -> render_colors [1,H,W,4], render_alphas [1,H,W,1], last_ids [1,H,W]
    def rasterize_to_pixels_fwd(self, means2d, conics, colors, opacities, backgrounds, isect_offsets, flatten_ids):
        n = means2d.shape[-2]
        render = torch.empty((1, self.H, self.W, 4), device=self.dev)
        alphas = torch.empty((1, self.H, self.W, 1), device=self.dev)
        last_ids = torch.empty((1, self.H, self.W), dtype=torch.int32, device=self.dev)
        bg = _f32(backgrounds) if backgrounds is not None else None
        E._check(self.L.gsb_gs_rasterize_fwd(self.h, n, _p(_f32(means2d)), _p(_f32(conics)), _p(_f32(colors)), _p(_f32(opacities)), _p(bg),
                                             _p(isect_offsets.contiguous()), _p(flatten_ids.contiguous()), flatten_ids.numel(), _p(render),
                                             _p(alphas), _p(last_ids)))
        return render, alphas, last_ids

    # gsplat::rasterize_to_pixels_bwd_tensor -> v_means2d, v_conics, v_colors, v_opacities
    def rasterize_to_pixels_bwd(self, means2d, conics, colors, opacities, backgrounds, isect_offsets, flatten_ids, render_alphas, last_ids,
                                v_render_colors, v_render_alphas):
        n = means2d.shape[-2]
        v_means2d, v_conics, v_colors = torch.empty_like(means2d), torch.empty_like(conics), torch.empty_like(colors)
        v_opac = torch.empty_like(opacities)
        bg = _f32(backgrounds) if backgrounds is not None else None
        E._check(self.L.gsb_gs_rasterize_bwd(self.h, n, _p(_f32(means2d)), _p(_f32(conics)), _p(_f32(colors)), _p(_f32(opacities)), _p(bg),
                                             _p(isect_offsets.contiguous()), _p(flatten_ids.contiguous()), flatten_ids.numel(),
                                             _p(_f32(render_alphas)), _p(last_ids.contiguous()), _p(_f32(v_render_colors)),
                                             _p(_f32(v_render_alphas)), _p(v_means2d), _p(v_conics), _p(v_colors), _p(v_opac)))
        return v_means2d, v_conics, v_colors, v_opac

    # fusedssim (ssim.cu:368-407): img [B,CH,H,W] -> ssim_map, dm_dmu1, dm_dsigma1_sq, dm_dsigma12
    def fusedssim(self, C1, C2, img1, img2, train=True):
        B, CH, H, W = img1.shape
        out = [torch.empty_like(img1) for _ in range(4 if train else 1)]
        E._check(self.L.gsb_gs_ssim_fwd(self.h, B * CH, H, W, C1, C2, _p(_f32(img1)), _p(_f32(img2)), _p(out[0]),
                                        _p(out[1]) if train else None, _p(out[2]) if train else None, _p(out[3]) if train else None))
        return tuple(out)

    # fusedssim_backward (ssim.cu:409-460) -> dL_dimg1
    def fusedssim_backward(self, C1, C2, img1, img2, dL_dmap, dm_dmu1, dm_dsigma1_sq, dm_dsigma12):
        B, CH, H, W = img1.shape
        d = torch.empty_like(img1)
        E._check(self.L.gsb_gs_ssim_bwd(self.h, B * CH, H, W, _p(_f32(img1)), _p(_f32(img2)), _p(_f32(dL_dmap)), _p(_f32(dm_dmu1)),
                                        _p(_f32(dm_dsigma1_sq)), _p(_f32(dm_dsigma12)), _p(d)))
        return d
