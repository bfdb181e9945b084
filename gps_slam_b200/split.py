"""Functional split of the SLAM loop over the GPUs of one box ("path to 6x"): rank 0 owns the TSDF side -- fusion of every
frame, the key-frame / window raycasts of every cycle, the raycast -> tensor glue -- and stores the camera maps of a cycle into the other
ranks' mailboxes over NVLink (csrc/peer_mbox.cu); ranks 1 .. G-1 hold the Gaussian shards (spatial-block ownership over G-1 ranks, peer-memory
exchange of the partial images among themselves) and never touch a voxel.

Why: measured on 8 B200s the TSDF side of a 10-frame step is ~6.6 ms whatever the number of GPUs (its raycasts are latency chains, sharding
them buys nothing: profiles/r02_shard_modes.md) while the Gaussian side shrinks with the shard; on shared SMs the two add up, on separate GPUs
the step is the longer of the two.

Every rank runs the same host logic (window, key frames, the pinned random sequence), so all agree on which cameras a cycle raycasts and in
which order: slot k of a cycle's mailbox half is its k-th raycast.  Hand-shakes are mailbox counters waited on by the streams:
READY (on a Gaussian rank) = cycles whose maps have landed; CONSUMED + r (on rank 0) = cycles rank r is done with; two mailbox halves
alternate, so the TSDF rank runs at most one cycle ahead.  use_gt_pose only (the tracked pose would have to travel as well).
"""
import numpy as np
import torch
import torch.distributed as dist

from . import engine as E
from . import parallel
from . import slam
from . import synthetic as syn

MAX_CAMS = 10          # raycasts per cycle: localframe_cam_window_length + 1 window cameras + keyframe_select_max key frames
F_READY, F_EVAL_READY = 0, 1
F_CONSUMED, F_EVAL_DONE = 16, 64   # + rank


class DevBuf:
    """a device address the engine wrappers accept in place of a tensor (engine._ptr calls data_ptr())"""
    __slots__ = ("ptr",)

    def __init__(self, ptr):
        self.ptr = ptr

    def data_ptr(self):
        return self.ptr


class _CudaView:
    """float32 device memory as a zero-copy torch tensor (torch.as_tensor reads __cuda_array_interface__)"""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


class SplitSlamPipeline(slam.SlamPipeline):
    def __init__(self, intr, device=0, stream=None, rank=0, world=4, cfg=None, seed=42, gs_capacity=1 << 21):
        assert world >= 3, "the functional split needs one TSDF rank and at least two Gaussian ranks"
        self.intr, self.mode, self.rank, self.world = intr, "train", rank, world
        self.cfg = dict(slam.OFFICE0, **(cfg or {}))
        c = self.cfg
        self.device = torch.device("cuda", device)
        self.use_gt_pose = True
        self.W, self.H = intr["width"], intr["height"]
        self.is_tsdf = rank == 0
        self.gs_rank, self.gs_world = rank - 1, world - 1
        self.tsdf_sharded = False
        self.exchange = "peer"
        self.acc5 = self.sp_rgb = self.sp_depth = self.sp_alpha = None
        P = self.W * self.H
        self.slot_bytes = 16 * P                               # depth f32 [P] + colour f32 [P,3]
        self.half_bytes = (MAX_CAMS + 1) * self.slot_bytes     # + the free-view vertex image (float4 [P])
        self.mbox = E.PeerMailbox(device, rank, world, 2 * self.half_bytes)   # same size on every rank (the TSDF rank only uses its counters)
        handles = [None] * world
        dist.all_gather_object(handles, self.mbox.export_handle())
        self.mbox.attach(handles)
        self.stream = stream if stream is not None else torch.cuda.current_stream(self.device)
        self.sG = self.sT = self.stream
        self.pool = slam.BufferPool(self.device)
        self.own_comm = False
        self.comm = None
        if self.is_tsdf:
            self.tsdf = E.TsdfEngine(intr, voxel_size=c["voxel_size"], mu=c["trunc_dist"], view_frustum_min=c["viewFrustum_min"],
                                     view_frustum_max=c["viewFrustum_max"], tracker=0, device=device)
            self.tsdf.set_stream(self.stream.cuda_stream or 1)
            self.gs = E.GaussianEngine(self.W, self.H, capacity=1024, device=device)    # only its stateless image kernels (glue) are used here
            self.gs.set_stream(self.stream.cuda_stream or 1)
            comm_handle = b"\0" * 64
        else:
            self.tsdf = None
            self.gs = E.GaussianEngine(self.W, self.H, capacity=gs_capacity, device=device)
            self.gs.set_stream(self.stream.cuda_stream or 1)
            self.comm = E.PeerComm(device, self.gs_rank, self.gs_world, self.W, self.H)
            comm_handle = self.comm.export_handle()
            self.stage = torch.empty((self.H, self.W, 4), dtype=torch.uint8, device=self.device)   # host frames of the e2e leg land here
        ch = [None] * world
        dist.all_gather_object(ch, comm_handle)
        if not self.is_tsdf:
            self.comm.attach(ch[1:])
            self.own_comm = True
            self.gs.set_comm(self.comm)
        dist.barrier()
        pre = [self.pool.get((self.H, self.W, 3)) for _ in range(160 if not self.is_tsdf else 24)] + [self.pool.get((self.H, self.W)) for _ in range(24)]
        for t in pre:
            self.pool.put(t)
        self.seed = seed
        self.cycle_seq = 0     # cycles since construction (never reset: the mailbox counters only grow)
        self.eval_seq = 0
        self.breakdown = None
        self.ev_gs = self.ev_spawn = None
        self.voxel_size = float(c["voxel_size"])
        self.reset()

    # ------------------------------------------------------------------------------------------------------------
    def reset(self):
        torch.cuda.synchronize(self.device)
        dist.barrier()
        if self.is_tsdf:
            self.tsdf.resetAll()
        else:
            self.gs.set_params(dict(means=np.zeros((0, 3), np.float32), scales=np.zeros((0, 3), np.float32), quats=np.zeros((0, 4), np.float32),
                                    featuresDc=np.zeros((0, 3), np.float32), featuresRest=np.zeros((0, 45), np.float32),
                                    opacities=np.zeros((0, 1), np.float32)))
        for cam in list(getattr(self, "window", [])) + list(getattr(self, "keyframes", [])):
            self._release(cam)
        import collections
        import random
        self.window = collections.deque()
        self.keyframes = []
        self.opt_cams = []
        self.rng = random.Random(self.seed)
        self.curr = None
        self.n_gauss = 0
        self.cycles = 0
        self.last_loss = None
        self.spawned_last = 0
        self._pose_host = np.zeros(16, np.float32)
        self.track_err, self.track_iters = [], []
        if not self.is_tsdf and getattr(self, "_loss_pending", False):
            self.gs.loss_end()
        self._loss_pending = False

    def close(self):
        torch.cuda.synchronize(self.device)
        if self.tsdf:
            self.tsdf.close()
        if self.gs:
            self.gs.close()
        if self.comm is not None:
            self.comm.close()
            self.comm = None
        self.mbox.close()

    def _release(self, cam):
        if isinstance(cam.depth_map, torch.Tensor):
            self.pool.put(cam.depth_map, None)
            self.pool.put(cam.color_map, None)
        cam.depth_map = cam.color_map = None

    def _st(self):
        return self.stream.cuda_stream or 1

    # ---- mailbox layout (byte offsets inside a Gaussian rank's mailbox)
    def _off(self, half, k):
        return half * self.half_bytes + k * self.slot_bytes

    def _off_vertex(self, half):
        return half * self.half_bytes + MAX_CAMS * self.slot_bytes

    # ------------------------------------------------------------------------------------------------------------
    def process_frame(self, idx, rgba_all, depth_all, poses, resident, frame_offset=0):
        c2w = np.asarray(poses[idx], np.float32)
        gt = syn.c2w_to_colmajor(c2w)
        frame = rgba_all[idx - frame_offset]
        if self.is_tsdf:
            if resident:
                self.tsdf.ProcessFrameDevice(frame, depth_all[idx - frame_offset], gt)
            else:
                self.tsdf.ProcessFrame(frame, depth_all[idx - frame_offset], gt)
        self._frame = (frame, resident)
        est = c2w.copy()     # use_gt_pose: pose_d->GetInvM() is the given pose
        self.curr = (idx, c2w, est)
        self._update_frame_list(idx, c2w, est)
        c = self.cfg
        if idx % c["local_opt_interval"] == 0 and idx > 0:
            self.cycle_seq += 1
            seq, half = self.cycle_seq, self.cycle_seq & 1
            order = self._cycle_cameras()
            draws = self._draw_iterations(len(self.opt_cams))
            if self.is_tsdf:
                if seq > 2:     # the half about to be overwritten belongs to cycle seq - 2: every Gaussian rank must be done with it
                    for r in range(1, self.world):
                        self.mbox.wait(F_CONSUMED + r, seq - 2, self._st())
                for k, cam in enumerate(order):
                    self._raycast_by_cam(cam)
                    for r in range(1, self.world):
                        self.mbox.put(r, self._off(half, k), cam.depth_map, 4 * self.W * self.H, self._st())
                        self.mbox.put(r, self._off(half, k) + 4 * self.W * self.H, cam.color_map, 12 * self.W * self.H, self._st())
                for r in range(1, self.world):
                    self.mbox.put(r, self._off_vertex(half), self.tsdf.GetFreeVertex(), 16 * self.W * self.H, self._st())
                    self.mbox.signal(r, F_READY, seq, self._st())
            else:
                base = self.mbox.local_ptr()
                for k, cam in enumerate(order):
                    cam.depth_map = DevBuf(base + self._off(half, k))
                    cam.color_map = DevBuf(base + self._off(half, k) + 4 * self.W * self.H)
                self.mbox.wait(F_READY, seq, self._st())
                self._vertex = DevBuf(base + self._off_vertex(half))
                self._init_new_gaussians()
                self._local_optimize(draws)
                self._remove_redundant()
                self.mbox.signal(0, F_CONSUMED + self.rank, seq, self._st())
            self.cycles += 1

    def _make_cam(self, idx, c2w, est):
        if self.is_tsdf:
            return slam.Cam(idx, c2w, est, None)     # the TSDF rank never needs the float image
        img = self.pool.get((self.H, self.W, 3), self.stream)
        frame, resident = self._frame
        if not resident:
            with torch.cuda.stream(self.stream):
                self.stage.copy_(frame, non_blocking=True)
            frame = self.stage
        self.gs.frame_to_float(frame.data_ptr(), None, img, None)
        return slam.Cam(idx, c2w, est, img)

    def _update_frame_list(self, idx, c2w, est):
        """updateFrameList (slam_pipeline.cpp:293-360); as SlamPipeline's, with this class's buffer release"""
        if idx == 0:
            return
        c = self.cfg
        cam = None
        if idx % c["localframe_cam_window_interval"] == 0:
            cam = self._make_cam(idx, c2w, est)
            self.window.append(cam)
            if len(self.window) == c["localframe_cam_window_length"] + 1:
                old = self.window.popleft()
                if not any(k is old for k in self.keyframes):
                    self._release(old)
                    self.pool.put(old.image, None)
        if not self.keyframes:
            is_key = True
        else:
            last = self.keyframes[-1]
            theta = slam.rot_compare(last.c2w_slam[:3, :3].astype(np.float64), est[:3, :3].astype(np.float64))
            trans = float(np.linalg.norm(last.c2w_slam[:3, 3] - est[:3, 3]))
            is_key = theta > c["keyframe_theta_thres"] or trans > c["keyframe_trans_thres"]
        if is_key:
            self.keyframes.append(cam if cam is not None else self._make_cam(idx, c2w, est))

    def _cycle_cameras(self):
        """the cameras a cycle raycasts, in raycast order (key frames first, then the window: SlamPipeline._key_frame_raycast /
        _local_frame_raycast); draws the key frames from the pinned random sequence"""
        k = min(self.cfg["keyframe_select_max"], len(self.keyframes))
        pool = list(self.keyframes)
        picked = []
        for _ in range(k):
            i = self.rng.randrange(len(pool))
            cam = pool[i]
            pool[i] = pool[-1]
            pool.pop()
            picked.append(cam)
        self.opt_cams = list(self.window) + picked
        order = picked + list(self.window)
        assert len(order) <= MAX_CAMS
        return order

    def _draw_iterations(self, n_cams):
        """camera index of every optimiser iteration of the cycle (localOptimize's sampling without replacement); drawn by every rank so
        that the random sequences stay in step"""
        out, current = [], list(range(n_cams))
        for _ in range(self.cfg["local_opt_iters"]):
            if not current:
                current = list(range(n_cams))
            i = self.rng.randrange(len(current))
            out.append(current[i])
            current[i] = current[-1]
            current.pop()
        return out

    def _raycast_by_cam(self, cam):
        """TSDF rank: runRaycastByCam into pool buffers (free-view raycast + tensor glue)"""
        self.tsdf.runRaycast(syn.c2w_to_colmajor(cam.c2w_slam), self.intr)
        self._release(cam)
        cam.depth_map = self.pool.get((self.H, self.W), self.stream)
        cam.color_map = self.pool.get((self.H, self.W, 3), self.stream)
        self.gs.raycast_maps(self.tsdf.GetFreeVertex(), self.tsdf.GetFreeImage(), cam.c2w, self.tsdf.getVoxelSize(), cam.depth_map, cam.color_map)

    def _init_new_gaussians(self):
        cam = self.window[-1]
        c = self.cfg
        before = self.n_gauss
        self.gs.addGaussians(cam.c2w_slam, self.intr, self._vertex, self.voxel_size, cam.depth_map, cam.color_map, cam.image,
                             seed=self.seed * 7919 + cam.id, color_error_thres=c["color_error_thres"], depth_vis_min=c["depth_vis_min"],
                             depth_vis_max=c["depth_vis_max"], alpha_vis_max=c["alpha_vis_max"], sample_ratio=c["new_gs_sample_ratio"],
                             max_init_scale=c["max_init_scale"], min_init_scale=c["min_init_scale"], default_opacity=c["default_opacities"],
                             rank=self.gs_rank, world=self.gs_world)
        self.n_gauss = self.gs.getGaussianNum()
        self.spawned_last = self.n_gauss - before

    def _local_optimize(self, draws):
        self.gs.initOptimizers()
        for ci in draws:
            cam = self.opt_cams[ci]
            self.gs.train_step(cam.c2w_slam, self.intr, cam.depth_map, cam.color_map, cam.image)

    # ------------------------------------------------------------------------------------------------------------
    def end_of_step(self, resident):
        if not resident:
            if self.is_tsdf:
                self._pose_host = self.tsdf.pose()[1]
            elif self.cycles:
                if self._loss_pending:
                    self.last_loss = self.gs.loss_end()
                self.gs.loss_begin()
                self._loss_pending = True

    def flush_readback(self):
        if not self.is_tsdf and self._loss_pending:
            self.last_loss = self.gs.loss_end()
            self._loss_pending = False

    def render_eval(self, c2w, rgb, depth, alpha):
        """renderEvalImgs for one camera: the TSDF rank raycasts and ships the maps, the Gaussian ranks render.  rgb / depth / alpha and the
        returned TSDF colour image are valid on the Gaussian ranks (zeros on the TSDF rank)"""
        self.eval_seq += 1
        n = self.eval_seq
        cam = slam.Cam(-1, np.asarray(c2w, np.float32), np.asarray(c2w, np.float32), None)
        P = self.W * self.H
        if self.is_tsdf:
            if n > 1:
                for r in range(1, self.world):
                    self.mbox.wait(F_EVAL_DONE + r, n - 1, self._st())
            self._raycast_by_cam(cam)
            for r in range(1, self.world):
                self.mbox.put(r, self._off(0, 0), cam.depth_map, 4 * P, self._st())
                self.mbox.put(r, self._off(0, 0) + 4 * P, cam.color_map, 12 * P, self._st())
                self.mbox.signal(r, F_EVAL_READY, n, self._st())
            base = cam.color_map.clone()
            self._release(cam)
            rgb.zero_(), depth.zero_(), alpha.zero_()
            torch.cuda.synchronize(self.device)
            return base
        basep = self.mbox.local_ptr()
        self.mbox.wait(F_EVAL_READY, n, self._st())
        dm, cm = DevBuf(basep + self._off(0, 0)), DevBuf(basep + self._off(0, 0) + 4 * P)
        self.gs.forward(cam.c2w_slam, self.intr, dm, cm, rgb, depth, alpha)
        torch.cuda.synchronize(self.device)
        base = torch.as_tensor(_CudaView(cm.ptr, (self.H, self.W, 3)), device=self.device).clone()   # the TSDF colour image, out of the mailbox
        self.mbox.signal(0, F_EVAL_DONE + self.rank, n, self._st())
        torch.cuda.synchronize(self.device)
        return base

    def stats(self):
        s = {}
        if self.is_tsdf:
            s.update(visible_blocks_last_frame=self.tsdf.counter(2), allocated_blocks=self.tsdf.num_blocks - 1 - self.tsdf.counter(0))
            n, cnt = 0, np.zeros(8, np.int64)
        else:
            n, cnt = self.gs.getGaussianNum(), self.gs.counters().astype(np.int64)
        t = torch.tensor([n, int(cnt[2]), int(cnt[0]), int(cnt[4]), self.mbox.error()], device=self.device, dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        s.update(gaussians=int(t[0].item()), gaussians_this_rank=n, keyframes=len(self.keyframes), opt_cameras=len(self.opt_cams),
                 last_isects=int(t[2].item()), last_visible=int(t[3].item()), overflow_flags=int(t[1].item()), cycles=self.cycles,
                 mailbox_errors=int(t[4].item()))
        return s

    def parallelism(self):
        return ("functional split over %d GPUs: rank 0 owns the TSDF side (fusion, raycasts, raycast -> tensor glue) and stores the camera maps of "
                "every cycle into the other ranks' mailboxes over NVLink; ranks 1-%d hold the Gaussians sharded by spatial block (peer-memory "
                "exchange of the tile partial sums per optimiser iteration among themselves); no NCCL call on the data path"
                % (self.world, self.world - 1))

    def tracking_stats(self, poses, total):
        return None

    def time_dominant_kernel(self, stream, peak_gbs, reps=20, fresh_frames=None):
        """per-stage device times: the TSDF stages on rank 0, the Gaussian stages on the Gaussian ranks (rank 1 reports); merged on rank 0"""
        torch.cuda.synchronize(self.device)
        P = self.W * self.H
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.device)
        table, part = {}, {}
        if self.is_tsdf:
            for name, st in (("tsdf_allocate(6 kernels)", 0), ("tsdf_integrate", 1), ("tsdf_expected_depth(2)", 2), ("tsdf_raycast", 3),
                             ("tsdf_icp_maps", 4)):
                table[name] = self._time(stream, lambda st=st: self.tsdf.run_stage(st), reps, flush) * 1e6
            c2w_free = syn.c2w_to_colmajor(self.curr[1])
            table["tsdf_free_view_raycast"] = self._time(stream, lambda: self.tsdf.runRaycast(c2w_free, self.intr), reps, flush) * 1e6
            fresh = self._fresh_frame_stages(stream, fresh_frames) if fresh_frames is not None else None
            part = {"table": table, "fresh": fresh, "V": self.tsdf.counter(2)}
        else:
            g = self.gs
            live = [c for c in self.opt_cams if c.depth_map is not None and c.image is not None]
            cam = live[-1]
            with torch.cuda.stream(stream):
                g.initOptimizers()
                g.train_step(cam.c2w_slam, self.intr, cam.depth_map, cam.color_map, cam.image)
            for name, st in (("gs_project_sh", 0), ("gs_project_sh+bin_tiles(4 kernels)", 1), ("gs_raster_fwd_train", 2), ("gs_raster_bwd", 3)):
                table[name] = self._time(stream, lambda st=st: g.run_stage(st), reps, flush) * 1e6
            with torch.cuda.stream(stream):
                g.run_stage(3)

            def bwd_and_params():
                g.run_stage(1)
                g.run_stage(5)
            t15 = self._time(stream, bwd_and_params, reps, flush) * 1e6
            table["gs_bwd_params+adam_rest(2 kernels)"] = t15 - table["gs_project_sh+bin_tiles(4 kernels)"]
            with torch.cuda.stream(stream):
                g.run_stage(1)
                tested, passed = g.bwd_pair_stats()
                g.run_stage(4)
            table["gs_train_step(7 kernels, no flush)"] = self._time(
                stream, lambda: g.train_step(cam.c2w_slam, self.intr, cam.depth_map, cam.color_map, cam.image), reps, flush) * 1e6
            table["gs_exchange_barrier"] = self._time(stream, lambda: self.comm.barrier(stream.cuda_stream), reps, flush) * 1e6
            cnt = g.counters()
            part = {"table": table, "I": int(cnt[0]), "n_vis": int(cnt[4]), "tested": tested, "passed": passed, "n": g.getGaussianNum()}
        parts = [None] * self.world
        dist.all_gather_object(parts, part)
        t0, g1 = parts[0], parts[1]
        table = dict(t0["table"], **g1["table"])
        fresh = t0["fresh"]
        c = self.cfg
        per_frame = sum(v for k, v in (fresh or {}).items() if k in ("track", "allocate(6 kernels)", "integrate", "expected_depth(2)", "raycast", "icp_maps"))
        self.breakdown = {"tsdf_rank_fuse_ms": c["local_opt_interval"] * per_frame * 1e-3,
                          "tsdf_rank_free_view_raycasts_ms": len(self.opt_cams) * table["tsdf_free_view_raycast"] * 1e-3,
                          "gaussian_ranks_iterations_ms": c["local_opt_iters"] * table["gs_train_step(7 kernels, no flush)"] * 1e-3,
                          "of_which_exchange_barriers_ms": 2 * c["local_opt_iters"] * table["gs_exchange_barrier"] * 1e-3,
                          "free_view_raycasts_per_step": len(self.opt_cams), "frames_per_step": c["local_opt_interval"],
                          "iterations_per_step": c["local_opt_iters"],
                          "note": "the two sides run on different GPUs: a step is the longer of the two, not their sum"}
        t_int, Vf = (fresh["integrate"] * 1e-6, fresh["visible_blocks"]) if fresh else (table["tsdf_integrate"] * 1e-6, t0["V"])
        alg_int = Vf * (4 + 16 + 2 * 4096) + 8 * P
        integrate = {"kernel": "k_integrate_tma", "bound": "hbm", "achieved": alg_int / t_int / 1e9, "peak": peak_gbs, "unit": "GB/s",
                     "frac": alg_int / t_int / 1e9 / peak_gbs, "traffic": None, "algorithmic_bytes": alg_int, "avg_launch_us": t_int * 1e6,
                     "units": {"visible_blocks": Vf, "pixels": P}, "fresh_frame_stages_us": fresh,
                     "note": "rank 0 (the TSDF rank), fresh frames after the timed window"}
        t = g1["table"]["gs_raster_bwd"] * 1e-6
        alg = 24 * P + 48 * g1["I"] + 80 * g1["n_vis"]
        ach = alg / t / 1e9
        return {"kernel": "k_raster_bwd", "bound": "hbm", "achieved": ach, "peak": peak_gbs, "unit": "GB/s", "frac": ach / peak_gbs,
                "traffic": None, "algorithmic_bytes": alg, "avg_launch_us": t * 1e6,
                "units": {"pixels": P, "isects": g1["I"], "visible_gaussians": g1["n_vis"], "gaussians": g1["n"],
                          "bwd_pairs_tested": g1["tested"], "bwd_pairs_passed": g1["passed"], "measured_on": "rank 1's shard"},
                "kernels_us": table, "step_breakdown": self.breakdown, "tsdf_integrate": integrate}
