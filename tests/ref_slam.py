"""TEST INFRASTRUCTURE -- the SLAM loop with the Gaussian side computed by the REFERENCE's own gsplat kernels
(oracle/_ref/libgsplat_ref.so through oracle/gsplat_ref.RefGaussians: reference kernels + autograd wrappers + torch.optim.Adam in
the launch order of RawGaussianModel::gesForward / computeLoss / optimizersStep) and a torch restatement, on the GPU, of

    initNewGaussians        slam/slam_pipeline.cpp:450-526      (sample mask)
    addGaussians            slam/slam_gs_model.cpp:5-56         (selection, RawGaussianParams::init, add)
    RawGaussianParams::init src/raw_gs_param.cpp:11-74          (distCUDA2 = the reference's simple_knn kernel)
    localOptimize           slam/slam_pipeline.cpp:195-291      (initOptimizers, 20 x forward / L1 / backward / step / zero_grad)
    removeRedundantGs       slam/slam_pipeline.cpp:564-586      (prune mask) + prunePoints

The TSDF side (fusion, raycasts, raycast -> tensor glue) is this repository's engine, which is bit-exact to the reference's
(tests/test_tsdf_parity_gpu.py, tests/test_glue_gpu.py).  Host logic (window / keyframes / camera sampling) is SlamPipeline's, with
the same pinned random sequence, so both loops optimise the same cameras in the same order.  One thing is injected into the
reference side because the reference leaves it to an unseeded RNG: WHICH masked pixels spawn (torch::randperm prefix there); the
engine's deterministic per-pixel hash is used for both, so the two models start every cycle from the same new Gaussians.

Used by tests/test_psnr_vs_reference_gpu.py and tools/ref_loop.py; never by the product.
"""
import numpy as np
import torch

from gps_slam_b200 import slam
from gps_slam_b200 import synthetic as syn
from oracle import gsplat_ref
from oracle import slam_glue as sg

LRS = dict(means=1.6e-4, scales=5e-3, quats=1e-3, featuresDc=2.5e-3, featuresRest=5e-4, opacities=5e-2)   # office0.yaml:25-45
KEYS = gsplat_ref.RefGaussians.KEYS


def hash_u32(x):
    x = x & 0xffffffff
    x = x ^ (x >> 16)
    x = (x * 0x7feb352d) & 0xffffffff
    x = x ^ (x >> 15)
    x = (x * 0x846ca68b) & 0xffffffff
    return x ^ (x >> 16)


def engine_sampling(P, seed, ratio, device):
    """the engine's keep-decision per pixel (gs_spawn.cu k_spawn_select): hash(pixel ^ hash(seed)) < ratio * 2^32"""
    i = torch.arange(P, device=device, dtype=torch.int64)
    hs = int(hash_u32(torch.tensor([seed & 0xffffffff], dtype=torch.int64))[0])
    thr = min(0xffffffff, int(ratio * 4294967296.0))
    return hash_u32(i ^ hs) < thr


class RefSlamPipeline(slam.SlamPipeline):
    def __init__(self, intr, **kw):
        super().__init__(intr, **kw)
        self.K = np.array([[intr["fx"], 0, intr["cx"]], [0, intr["fy"], intr["cy"]], [0, 0, 1]], np.float32)
        self.ref = None          # RefGaussians once the model is non-empty

    def reset(self):
        super().reset()
        self.ref = None

    # ---- everything below replaces a gsb_gs_* call of SlamPipeline by the reference's code path
    def _params_cat(self, new):
        if self.ref is None:
            return new
        cur = {k: self.ref.p[k].detach() for k in KEYS}
        return {k: torch.cat([cur[k], new[k]], 0) for k in KEYS}

    def _set_model(self, params):
        r = gsplat_ref.RefGaussians.__new__(gsplat_ref.RefGaussians)
        r.ops_ns, r.dev, r.opt = None, self.device, None
        r.p = {k: params[k].detach().clone().to(self.device).requires_grad_(True) for k in KEYS}
        self.ref = r
        self.n_gauss = int(r.p["means"].shape[0])

    def _vertex_map(self):
        v = torch.from_numpy(self.tsdf.raycast(live=False)).to(self.device)          # [H,W,4], voxel units + confidence
        return (v[..., :3] * v[..., 3:4].gt(0)).contiguous() * self.tsdf.getVoxelSize()

    def _init_new_gaussians(self):
        cam = self.window[-1]
        c = self.cfg
        H, W = self.H, self.W
        with torch.no_grad(), torch.cuda.stream(self.sG):
            vertex = self._vertex_map()
            maps = dict(depth_map=cam.depth_map.unsqueeze(-1), color_map=cam.color_map, vertex_map=vertex)
            if self.ref is None:
                rgb = alpha = None
            else:
                out = self.ref.forward(cam.c2w_slam, self.K, W, H, cam.depth_map, cam.color_map)
                rgb, alpha = out["rgb"], out["alpha"]
            mask = sg.sample_mask(maps, cam.image, rgb, alpha, c["color_error_thres"], c["depth_vis_min"], c["depth_vis_max"], c["alpha_vis_max"])
            keep = mask.reshape(-1) & engine_sampling(H * W, self.seed * 7919 + cam.id, c["new_gs_sample_ratio"], self.device)
            sel = torch.nonzero(keep).reshape(-1)
            before = self.n_gauss
            if sel.numel() > 0:
                normal = sg.compute_normal_map(vertex)
                xyz = vertex.reshape(-1, 3)[sel].contiguous()
                d2 = gsplat_ref.ops().simple_knn(xyz)
                raw = torch.sqrt(d2).clamp(c["min_init_scale"], c["max_init_scale"]).unsqueeze(1).repeat(1, 3)
                raw[:, 2] = raw[:, 2] * 0.1
                z = torch.zeros_like(raw)
                z[:, 2] = 1
                n = sel.numel()
                new = dict(means=xyz, scales=raw.log(), quats=sg.compute_quat(z, normal.reshape(-1, 3)[sel]),
                           featuresDc=(cam.image.reshape(-1, 3)[sel] - 0.5) / sg.C0, featuresRest=torch.zeros((n, 15, 3), device=self.device),
                           opacities=torch.logit(c["default_opacities"] * torch.ones((n, 1), device=self.device)))
                self._set_model(self._params_cat(new))
            self.spawned_last = self.n_gauss - before

    def _local_optimize(self):
        if self.ref is None:
            return
        with torch.cuda.stream(self.sG):
            self.ref.opt = {k: torch.optim.Adam([self.ref.p[k]], lr=LRS[k], eps=1e-15) for k in KEYS}
            cams = self.opt_cams
            current = list(range(len(cams)))
            for _ in range(self.cfg["local_opt_iters"]):
                if not current:
                    current = list(range(len(cams)))
                i = self.rng.randrange(len(current))
                ci = current[i]
                current[i] = current[-1]
                current.pop()
                cam = cams[ci]
                r = self.ref.forward(cam.c2w_slam, self.K, self.W, self.H, cam.depth_map, cam.color_map)
                loss = torch.abs(cam.image - r["rgb"]).mean()
                loss.backward()
                for k in KEYS:
                    self.ref.opt[k].step()
                for k in KEYS:
                    self.ref.opt[k].zero_grad(set_to_none=True)
            self.last_loss = float(loss.detach())

    def _remove_redundant(self):
        if self.ref is None:
            return
        c = self.cfg
        with torch.no_grad(), torch.cuda.stream(self.sG):
            s = torch.exp(self.ref.p["scales"]).max(-1)[0]
            o = torch.sigmoid(self.ref.p["opacities"]).squeeze(-1)
            remove = (s < c["small_scale_thres"]) | (s > c["large_scale_thres"]) | (o < c["low_opac_thres"])
            if int(remove.sum()) > 0:
                self._set_model({k: self.ref.p[k].detach()[~remove] for k in KEYS})

    def _forward_all(self, cam, rgb, depth, alpha):
        with torch.no_grad(), torch.cuda.stream(self.sG):
            if self.ref is None:
                rgb.copy_(cam.color_map)
                alpha.zero_()
                depth.copy_(cam.depth_map)
                return
            out = self.ref.forward(cam.c2w_slam, self.K, self.W, self.H, cam.depth_map, cam.color_map)
            rgb.copy_(out["rgb"])
            depth.copy_(out["depth"][..., 0])
            alpha.copy_(out["alpha"][..., 0])

    def stats(self):
        return {"gaussians": self.n_gauss, "cycles": self.cycles, "keyframes": len(self.keyframes)}


def psnr(a, b):
    mse = float(((a - b) ** 2).mean())
    return 20.0 * np.log10(1.0 / np.sqrt(mse))


def run_both(n_frames, scale=1.0, eval_every=10, device=0, verbose=False):
    """the same n_frames through the engine's loop and through the reference-kernel loop; returns the comparison record"""
    dev = torch.device("cuda", device)
    intr = syn.intrinsics("replica", scale)
    poses = syn.trajectory(n_frames)
    H, W = intr["height"], intr["width"]
    rgba = torch.empty((n_frames, H, W, 4), dtype=torch.uint8, device=dev)
    depth = torch.empty((n_frames, H, W), dtype=torch.int16, device=dev)
    for i in range(n_frames):
        rgba[i], depth[i] = syn.render_frame(poses[i], intr, device=dev)
    res = {}
    renders = {}
    for name, cls in (("engine", slam.SlamPipeline), ("reference_kernels", RefSlamPipeline)):
        # Everything on the DEFAULT stream, as in the reference's program: its backward kernel is launched on the legacy default stream
        # (rasterize_to_pixels_bwd_ges_new_parallel.cu:264, no stream argument) while the zero-fill of its outputs goes to torch's
        # current stream -- on a side stream the two race and the kernel's atomics are partly wiped (seen as lost opacity gradients).
        pipe = cls(intr, mode="train", device=device, stream=None, overlap=False, gs_capacity=1 << 20)
        counts, spawned = [], []
        if True:
            import time
            torch.cuda.synchronize()
            t_loop = time.perf_counter()
            t_last10 = t_loop
            for f in range(n_frames):
                if f == n_frames - 11:
                    torch.cuda.synchronize()
                    t_last10 = time.perf_counter()
                pipe.process_frame(f, rgba, depth, poses, True)
                if f % 10 == 0 and f > 0:
                    counts.append(pipe.gs.getGaussianNum() if name == "engine" else pipe.n_gauss)
                    spawned.append(pipe.spawned_last)
            torch.cuda.synchronize()
            t_end = time.perf_counter()
            loop_s, last_cycle_s = t_end - t_loop, t_end - t_last10
            rgb, dep, alpha = torch.empty((H, W, 3), device=dev), torch.empty((H, W), device=dev), torch.empty((H, W), device=dev)
            ps, imgs = [], []
            for i in range(0, n_frames, eval_every):
                pipe.render_eval(poses[i], rgb, dep, alpha)
                torch.cuda.synchronize()
                img = rgb.clamp(0, 1).clone()
                ps.append(psnr(img, rgba[i][..., :3].float() / 255.0))
                imgs.append(img)
        res[name] = dict(psnr_db=float(np.mean(ps)), psnr_each=[round(float(p), 3) for p in ps], gaussians_after_each_cycle=counts,
                         spawned_each_cycle=spawned, last_loss=pipe.last_loss if name != "engine" else pipe.gs.loss(),
                         loop_seconds=loop_s, frames_per_sec=n_frames / loop_s, last_cycle_ms=last_cycle_s * 1e3)
        renders[name] = imgs
        pipe.close()
        if verbose:
            print(name, res[name])
    between = [psnr(a, b) for a, b in zip(renders["engine"], renders["reference_kernels"])]
    res["psnr_engine_vs_reference_render_db"] = float(np.mean(between))
    res["psnr_vs_reference_db"] = res["engine"]["psnr_db"] - res["reference_kernels"]["psnr_db"]
    res["frames"], res["cycles"], res["width"], res["height"] = n_frames, (n_frames - 1) // 10, W, H
    return res
