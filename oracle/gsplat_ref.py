"""TEST INFRASTRUCTURE ONLY -- the REFERENCE's own gsplat CUDA kernels + autograd wrappers (oracle/_ref/libgsplat_ref.so,
built by oracle/gsplat_ref/Makefile from the sources under /root/reference/gsplat, for sm_100a), driven from Python in the
order RawGaussianModel::gesForward / computeLoss / optimizersStep call them (reference src/raw_gs_model.cpp:188-417,
654-705).  Needs a GPU.  Purpose: pin oracle/gs_oracle.py (the numpy restatement) and the CUDA engine against what the
reference itself computes.  Only tests/ and bench.py's reference-kernel timing may import this; the product never does.

Third-party pieces that are not the reference's bytes: glm 1.0.1 is replaced by oracle/gsplat_ref/glm_compat (plain
2x2/3x3 algebra), libtorch is this image's torch 2.11 (the reference ships an unpinned libtorch in ThirdLibs.zip).
"""
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_ref", "libgsplat_ref.so")
_loaded = False


def available():
    return os.path.exists(LIB)


def ops():
    """torch.ops.gsplat_ref (loads the library on first use)"""
    global _loaded
    import torch
    if not _loaded:
        torch.ops.load_library(LIB)
        _loaded = True
    return torch.ops.gsplat_ref


def pose_inv(c2w):
    """reference src/tensor_math.cpp:56-67 (poseInv): [R|t]^-1 = [R^T | -R^T t]"""
    import torch
    R = c2w[:3, :3]
    t = c2w[:3, 3:4]
    Rt = R.transpose(0, 1)
    top = torch.cat([Rt, -torch.matmul(Rt, t)], 1)
    bottom = torch.tensor([[0.0, 0.0, 0.0, 1.0]], device=c2w.device, dtype=c2w.dtype)
    return torch.cat([top, bottom], 0)


class RefGaussians:
    """the six parameter tensors of RawGaussianParams (reference include/raw_gs_param.h:12-18) as leaf tensors with the
    reference's optimisers (src/raw_gs_model.cpp:654-675: one Adam per tensor, eps 1e-15)."""
    KEYS = ("means", "scales", "quats", "featuresDc", "featuresRest", "opacities")

    def __init__(self, params, device="cuda", lrs=None, ops_ns=None):
        """ops_ns: the torch.ops namespace to drive (default: the reference's own wrappers, torch.ops.gsplat_ref).  The tests pass
        torch.ops.gsplat_b200 (the product's C++ host layer, same op schemas) to run the identical script against it."""
        import torch
        self.ops_ns = ops_ns
        self.dev = torch.device(device)
        self.p = {k: torch.tensor(np.asarray(params[k], np.float32), device=self.dev, requires_grad=True) for k in self.KEYS}
        self.opt = None
        if lrs is not None:
            self.opt = {k: torch.optim.Adam([self.p[k]], lr=lrs[k], eps=1e-15) for k in self.KEYS}

    def forward(self, c2w, K, W, H, ref_depth_raw, base_color, delta_depth=0.1, max_radii=100, tile_size=16, degree=3, keep=False):
        """RawGaussianModel::gesForward (src/raw_gs_model.cpp:188-341); returns rgb [H,W,3], depth [H,W,1], alpha [H,W,1] and, with
        keep=True, every intermediate"""
        import torch
        o = self.ops_ns if self.ops_ns is not None else ops()
        dev = self.dev
        def t(x):
            return x.to(dev, torch.float32) if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x, np.float32), device=dev)
        c2w, Ks = t(c2w), t(K)
        ref_depth = t(ref_depth_raw).reshape(1, H, W, 1)
        base = t(base_color).reshape(1, H, W, 3)
        tile_w, tile_h = -(-W // tile_size), -(-H // tile_size)
        cam_T = c2w[:3, 3:4]
        ref_clamped = torch.where(ref_depth < 0.01, torch.full_like(ref_depth, 1000.0), ref_depth)
        viewmat = pose_inv(c2w)
        means = self.p["means"].contiguous()
        scales = torch.exp(self.p["scales"]).contiguous()
        radiis, means2d, depths, conics = o.fully_fused_projection(means, self.p["quats"], scales, viewmat.unsqueeze(0), Ks.unsqueeze(0),
                                                                   W, H, 0.3, 0.01, 1e10, 0.0)[:4]
        if max_radii > 0:
            radiis = torch.clamp_max(radiis, max_radii)
        shs = torch.cat([self.p["featuresDc"][:, None, :], self.p["featuresRest"]], 1)
        view_dirs = means - cam_T.transpose(0, 1)
        visible = radiis > 0
        sh_raw = o.spherical_harmonics(degree, view_dirs.unsqueeze(0), shs.unsqueeze(0), visible)
        colors = torch.clamp_min(sh_raw + 0.5, 0.0)
        tpg, isect_ids, flatten_ids, group_gs_ids, group_starts = o.isect_tiles_no_depth(means2d, radiis, depths, tile_size, tile_w, tile_h)
        offsets = o.isect_offset_encode_no_depth(isect_ids, 1, tile_w, tile_h)
        colors4 = torch.cat([colors, depths.unsqueeze(-1)], 2)
        opac = torch.sigmoid(self.p["opacities"])
        render, wsum = o.rasterize_ges(means2d, conics, colors4, opac, radiis, ref_clamped, base, W, H, tile_size, offsets, flatten_ids,
                                       group_gs_ids, group_starts, False, delta_depth)
        raw_rgb, raw_depth = render[..., :3], render[..., 3:]
        bw = torch.ones_like(wsum)
        rgb = (raw_rgb + base * bw) / (wsum + bw)
        dw = torch.zeros_like(wsum)
        dw.masked_fill_(ref_depth > 0, 1)
        depth = (raw_depth + ref_depth * dw) / (wsum + dw)
        out = dict(rgb=rgb[0], depth=depth[0], alpha=wsum[0])
        if keep:
            out.update(viewmat=viewmat, radii=radiis, means2d=means2d, depths=depths, conics=conics, sh_raw=sh_raw, colors=colors,
                       opac=opac, tiles_per_gauss=tpg, isect_ids=isect_ids, flatten_ids=flatten_ids, group_gs_ids=group_gs_ids,
                       group_starts=group_starts, tile_offsets=offsets, render=render, colors4=colors4, wsum=wsum)
        return out

    def train_iteration(self, c2w, K, W, H, ref_depth_raw, base_color, gt_rgb, step=True, **kw):
        """forward + computeLoss (l1 = mean |gt - rgb|, src/tensor_math.cpp:41-44) + backward (+ optimizersStep/ZeroGrad);
        returns the dict of intermediates (numpy) with the same keys as oracle.gs_oracle.ges_iteration"""
        import torch
        for k in self.KEYS:
            self.p[k].grad = None
        r = self.forward(c2w, K, W, H, ref_depth_raw, base_color, keep=True, **kw)
        for k in ("means2d", "conics", "colors4", "opac", "render", "wsum"):
            r[k].retain_grad()
        gt = torch.as_tensor(np.asarray(gt_rgb, np.float32), device=self.dev)
        loss = torch.abs(gt - r["rgb"]).mean()
        loss.backward()
        grads = {k: self.p[k].grad.detach().cpu().numpy().copy() for k in self.KEYS}
        if step and self.opt is not None:
            for k in self.KEYS:
                self.opt[k].step()
            for k in self.KEYS:
                self.opt[k].zero_grad()
        n = lambda t: t.detach().cpu().numpy()
        N = self.p["means"].shape[0]
        return dict(
            viewmat=n(r["viewmat"]),
            proj=dict(radii=n(r["radii"])[0], means2d=n(r["means2d"])[0], depths=n(r["depths"])[0], conics=n(r["conics"])[0]),
            colors=n(r["colors"])[0], sh_raw=n(r["sh_raw"])[0], tiles_per_gauss=n(r["tiles_per_gauss"]).reshape(-1),
            isect_ids=n(r["isect_ids"]), flatten_ids=n(r["flatten_ids"]), tile_offsets=n(r["tile_offsets"]).reshape(-1),
            group_gs_ids=n(r["group_gs_ids"]), group_starts=n(r["group_starts"]),
            render=n(r["render"])[0], alphas=n(r["alpha"])[..., 0], rgb=n(r["rgb"]), depth=n(r["depth"])[..., 0], loss=float(loss.detach()),
            v_render=n(r["render"].grad)[0], v_alphas=n(r["wsum"].grad)[0, ..., 0],
            v_means2d=n(r["means2d"].grad)[0], v_conics=n(r["conics"].grad)[0], v_colors=n(r["colors4"].grad)[0],
            v_opacities=n(r["opac"].grad).reshape(N), grads=grads)

    def params(self):
        return {k: self.p[k].detach().cpu().numpy().copy() for k in self.KEYS}


def ges_iteration(params, c2w, K, W, H, ref_depth_raw, base_color, gt_rgb, ops_ns=None, **kw):
    """one gesForward + loss + backward with the reference's kernels; same signature / keys as gs_oracle.ges_iteration"""
    return RefGaussians(params, ops_ns=ops_ns).train_iteration(c2w, K, W, H, ref_depth_raw, base_color, gt_rgb, step=False, **kw)
