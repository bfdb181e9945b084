"""Host-side SLAM loop over the C-ABI engines -- the Python mirror of SLAMPipeline (reference slam/slam_pipeline.h:6-94,
slam/slam_pipeline.cpp) for the parts that sit on the hot path: SLAMTrainCams (:52-173), updateFrameList (:293-360),
runRaycastByCam (:362-415), localFrameRaycast (:417-448), keyFrameRaycast (:528-561), initNewGaussians (:450-526),
localOptimize (:195-291), removeRedundantGs (:564-586).  Hyper-parameters are configs/release/replica/office0.yaml.

mode "recon": TSDF fusion only per frame (reference work_mode == "recon", slam_pipeline.cpp:96).
mode "train": TSDF fusion per frame and, every local_opt_interval frames, window/keyframe raycasts, Gaussian spawn,
              local_opt_iters optimiser iterations and the prune.

torch is used for device buffers only; every computation is a kernel of libgpsslam_b200.so.
"""
import collections
import contextlib
import math
import os
import random

import numpy as np
import torch

from . import engine as E
from . import parallel
from . import synthetic as syn

DEFAULT_MODE = "train"

# PIPE / MODEL sections of configs/release/replica/office0.yaml
OFFICE0 = dict(
    new_gs_sample_ratio=0.25, color_error_thres=0.05, localframe_cam_window_length=2, localframe_cam_window_interval=5,
    local_opt_iters=20, local_opt_interval=10, keyframe_theta_thres=30.0, keyframe_trans_thres=0.3, keyframe_select_max=7,
    depth_vis_max=5.0, depth_vis_min=0.0, alpha_vis_max=5.0,
    large_scale_thres=0.1, small_scale_thres=0.003, low_opac_thres=0.005,
    max_init_scale=0.01, min_init_scale=-1.0, default_opacities=0.5,
    voxel_size=0.005, trunc_dist=0.02, viewFrustum_min=0.2, viewFrustum_max=10.0,
)


def workload_name(mode):
    if mode == "recon":
        return "Replica-shaped 1200x680 synthetic RGB-D, work_mode=recon (TSDF fusion + raycast per frame, use_gt_pose=true)"
    return "Replica office0-shaped 1200x680 synthetic RGB-D, work_mode=train (gsplat GES + TSDF, use_gt_pose=true), office0.yaml hyper-parameters"


class Cam:
    """the fields of the reference's Camera that the hot path reads (include/camera.h): id, c2w (GT), c2w_slam, image"""
    __slots__ = ("id", "c2w", "c2w_slam", "image", "depth_map", "color_map")

    def __init__(self, idx, c2w, c2w_slam, image):
        self.id, self.c2w, self.c2w_slam, self.image = idx, c2w, c2w_slam, image
        self.depth_map = self.color_map = None


def rot_compare(Ra, Rb):
    """rotCompare (src/tensor_math.cpp:302-317), degrees"""
    c = (np.trace(Ra.T @ Rb) - 1.0) / 2.0
    return math.degrees(math.acos(min(1.0, max(-1.0, float(c)))))


class BufferPool:
    """device image buffers recycled between cycles (the reference allocates fresh tensors per raycast / camera).  A buffer may be
    released while work that reads it is still queued on another stream: put() takes the event after which it is free, and
    get() makes the stream that is going to overwrite it wait for that event (device-side, the host never blocks)."""

    def __init__(self, device):
        self.device, self.free = device, collections.defaultdict(collections.deque)

    def get(self, shape, writer_stream=None):
        f = self.free[shape]
        if not f:
            return torch.empty(shape, dtype=torch.float32, device=self.device)
        t, ev = f.popleft()
        if ev is not None and writer_stream is not None:
            writer_stream.wait_event(ev)
        return t

    def put(self, t, free_after=None):
        if t is not None:
            self.free[tuple(t.shape)].append((t, free_after))


class SlamPipeline:
    def __init__(self, intr, mode="train", device=0, stream=None, rank=0, world=1, cfg=None, seed=42, gs_capacity=1 << 21, use_gt_pose=True,
                 tracker=1, overlap=True, exchange=None, comm=None, tsdf_shard=None):
        """use_gt_pose=False: online tracking (TSDF.use_gt_pose: false) with the extended (1) or icp (2) tracker.
        world > 1: Gaussians sharded across ranks (parallel.py); torch.distributed must be initialised by the caller."""
        self.intr, self.mode, self.rank, self.world = intr, mode, rank, world
        self.cfg = dict(OFFICE0, **(cfg or {}))
        c = self.cfg
        self.device = torch.device("cuda", device)
        self.use_gt_pose = use_gt_pose
        # world > 1: the voxel hash is sharded by spatial block as well (csrc/tsdf.h: replicated hash table, owned-block integrate, row-slab
        # raycasts over peer memory); GSB_TSDF_SHARD=0 keeps a full replica of the map on every rank (A/B timing)
        self.tsdf_sharded = world > 1 and (tsdf_shard if tsdf_shard is not None else os.environ.get("GSB_TSDF_SHARD", "1") != "0")
        tkw = dict(voxel_size=c["voxel_size"], mu=c["trunc_dist"], view_frustum_min=c["viewFrustum_min"], view_frustum_max=c["viewFrustum_max"],
                   tracker=0 if use_gt_pose else tracker)
        if self.tsdf_sharded:
            self.tsdf = parallel.make_sharded_tsdf(intr, rank, world, device, **tkw)
        else:
            self.tsdf = E.TsdfEngine(intr, device=device, **tkw)
        self.W, self.H = intr["width"], intr["height"]
        # multi-GPU exchange of the partial images: "peer" (stores into peer memory below the C ABI, csrc/gs_comm.h) or "nccl" (all-reduce
        # between two C-ABI calls).  comm: an already attached PeerComm (tests with several engines in one process)
        self.exchange = (exchange or os.environ.get("GSB_EXCHANGE", "peer")) if world > 1 else None
        self.comm, self.own_comm = comm, False
        self.acc5 = torch.empty(self.W * self.H * 5, dtype=torch.float32, device=self.device) if (self.exchange == "nccl" and mode == "train") else None
        self.sp_rgb = self.sp_depth = self.sp_alpha = None
        self.gs = E.GaussianEngine(self.W, self.H, capacity=gs_capacity, device=device) if mode == "train" else None
        if self.gs and self.exchange == "peer":
            if self.comm is None:
                self.comm, self.own_comm = parallel.make_peer_comm(device, rank, world, self.W, self.H), True
            self.gs.set_comm(self.comm)
        # Two streams (train mode): the TSDF side of the loop (fusion, raycasts, raycast -> tensor glue) runs on sT, the Gaussian side
        # (spawn, optimiser iterations, prune) on sG.  Nothing on the TSDF side depends on the Gaussians, so the 20 optimiser
        # iterations of a cycle (issue-bound rasteriser kernels) overlap with the fusion + raycasts of the following frames
        # (latency-bound kernels); events order the hand-overs (maps -> optimiser, spawn -> next free-view raycast, buffer reuse).
        # Results are identical to the single-stream order.  overlap=False keeps everything on one stream.
        self.stream = stream if stream is not None else torch.cuda.current_stream(self.device)
        self.sG = self.stream
        prio = int(os.environ.get("GSB_TSDF_STREAM_PRIORITY", "0"))   # 0 normal, -1 high (experiments; the Gaussian side is the critical path)
        self.sT = torch.cuda.Stream(device=self.device, priority=prio) if (overlap and mode == "train") else self.sG
        self.tsdf.set_stream(self.sT.cuda_stream or 1)
        if self.gs:
            self.gs.set_stream(self.sG.cuda_stream or 1)
        self.ev_spawn = None       # recorded on sG after the spawn (the last reader of the engine's free-view vertex image)
        self.ev_gs = None          # recorded on sG after the last enqueued Gaussian-side work that reads camera buffers / the free vertex image
        self.pool = BufferPool(self.device)
        if mode == "train":
            # image buffers are recycled; allocate the working set up front so that no cudaMalloc (a device-wide sync) lands inside
            # the frame loop: window + keyframes hold an image, a depth map and a colour map each
            # (every keyframe keeps its image for good -- the reference's GPU memory grows the same way -- so size for the keyframes of a
            # few-thousand-frame sequence: 160 x 9.8 MB at 1200x680)
            pre = [self.pool.get((self.H, self.W, 3)) for _ in range(160)] + [self.pool.get((self.H, self.W)) for _ in range(32)]
            for t in pre:
                self.pool.put(t)
        self.seed = seed
        self.reset()

    # ------------------------------------------------------------------------------------------------------------
    def reset(self):
        torch.cuda.synchronize(self.device)
        if self.world > 1 and parallel.dist.is_available() and parallel.dist.is_initialized():
            # sharded engines: no rank may clear its map while another one is still storing rows / visibility marks of the last frame into it
            parallel.dist.barrier()
        self.ev_gs = self.ev_spawn = None
        self.tsdf.resetAll()
        if self.gs:
            self.gs.set_params(dict(means=np.zeros((0, 3), np.float32), scales=np.zeros((0, 3), np.float32), quats=np.zeros((0, 4), np.float32),
                                    featuresDc=np.zeros((0, 3), np.float32), featuresRest=np.zeros((0, 45), np.float32),
                                    opacities=np.zeros((0, 1), np.float32)))
        for cam in list(getattr(self, "window", [])) + list(getattr(self, "keyframes", [])):
            self._release(cam)
        self.window = collections.deque()
        self.keyframes = []
        self.opt_cams = []
        self.rng = random.Random(self.seed)   # RandomSelector's std::random_device, pinned so that runs repeat
        self.curr = None
        self.n_gauss = 0
        self.cycles = 0
        self.last_loss = None
        self.spawned_last = 0
        self._pose_host = np.zeros(16, np.float32)
        self.track_err, self.track_iters = [], []
        if self.gs and getattr(self, "_loss_pending", False):
            self.gs.loss_end()
        self._loss_pending = False

    def close(self):
        self.tsdf.close()
        if self.gs:
            self.gs.close()
        if self.own_comm and self.comm is not None:
            self.comm.close()
            self.comm = None

    def _release(self, cam):
        self.pool.put(cam.depth_map, self.ev_gs)
        self.pool.put(cam.color_map, self.ev_gs)
        cam.depth_map = cam.color_map = None

    @contextlib.contextmanager
    def _glue(self):
        """the Gaussian engine's stateless image kernels (frame_to_float, raycast_maps) issued on the TSDF stream"""
        if self.sT is self.sG:
            yield
            return
        self.gs.set_stream(self.sT.cuda_stream or 1)
        try:
            yield
        finally:
            self.gs.set_stream(self.sG.cuda_stream or 1)

    def _handover(self, src, dst):
        """dst waits (on the device) for everything queued on src so far"""
        if src is dst:
            return None
        ev = torch.cuda.Event()
        ev.record(src)
        dst.wait_event(ev)
        return ev

    @contextlib.contextmanager
    def single_stream(self):
        """everything on the Gaussian stream (per-kernel timing, evaluation renders); synchronises on entry and exit"""
        torch.cuda.synchronize(self.device)
        keep = self.sT
        self.sT = self.sG
        self.tsdf.set_stream(self.sG.cuda_stream or 1)
        try:
            yield
        finally:
            torch.cuda.synchronize(self.device)
            self.sT = keep
            self.tsdf.set_stream(self.sT.cuda_stream or 1)

    # ------------------------------------------------------------------------------------------------------------
    def process_frame(self, idx, rgba_all, depth_all, poses, resident, frame_offset=0):
        """one iteration of the SLAMTrainCams loop body (slam_pipeline.cpp:69-143); frame idx is rgba_all[idx - frame_offset]"""
        c2w = np.asarray(poses[idx], np.float32)
        if not self.use_gt_pose and self.tsdf.frames_processed() == 0:
            self.tsdf.set_pose(syn.c2w_to_colmajor(c2w))   # the reference re-bases the trajectory on frame 0; here frame 0 is given
        gt = syn.c2w_to_colmajor(c2w) if self.use_gt_pose else None
        if resident:
            self.tsdf.ProcessFrameDevice(rgba_all[idx - frame_offset], depth_all[idx - frame_offset], gt)
        else:
            self.tsdf.ProcessFrame(rgba_all[idx - frame_offset], depth_all[idx - frame_offset], gt)
        if self.mode == "recon":
            return
        est = self.tsdf.pose()[1].reshape(4, 4).T.copy()      # est_pose = pose_d->GetInvM() as a row-major tensor (:81-82)
        if not self.use_gt_pose:
            self.track_err.append(float(np.linalg.norm(est[:3, 3] - c2w[:3, 3])))
            self.track_iters.append(self.tsdf.tracker_result()[2])
        self.curr = (idx, c2w, est)
        self._update_frame_list(idx, c2w, est)
        c = self.cfg
        if idx % c["local_opt_interval"] == 0 and idx > 0:
            if self.ev_spawn is not None and self.sT is not self.sG:
                self.sT.wait_event(self.ev_spawn)    # the previous spawn has read the engine's free-view vertex image
            self._key_frame_raycast()
            self._local_frame_raycast()
            self._handover(self.sT, self.sG)          # maps and camera images are ready for the Gaussian side
            self._init_new_gaussians()
            if self.sT is not self.sG:
                self.ev_spawn = torch.cuda.Event()
                self.ev_spawn.record(self.sG)
            self._local_optimize()
            self._remove_redundant()
            if self.sT is not self.sG:
                self.ev_gs = torch.cuda.Event()
                self.ev_gs.record(self.sG)
            self.cycles += 1

    def _make_cam(self, idx, c2w, est):
        """curr_cam.toGPU(): float image of the current frame (slam_pipeline.cpp:84)"""
        img = self.pool.get((self.H, self.W, 3), self.sT)
        with self._glue():
            self.gs.frame_to_float(self.tsdf.current_rgba(), None, img, None)
        return Cam(idx, c2w, est, img)

    def _update_frame_list(self, idx, c2w, est):
        """updateFrameList (slam_pipeline.cpp:293-360)"""
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
                    self.pool.put(old.image, self.ev_gs)
        is_key = False
        if not self.keyframes:
            is_key = True
        else:
            last = self.keyframes[-1]
            theta = rot_compare(last.c2w_slam[:3, :3].astype(np.float64), est[:3, :3].astype(np.float64))
            trans = float(np.linalg.norm(last.c2w_slam[:3, 3] - est[:3, 3]))
            is_key = theta > c["keyframe_theta_thres"] or trans > c["keyframe_trans_thres"]
        if is_key:
            self.keyframes.append(cam if cam is not None else self._make_cam(idx, c2w, est))

    def _raycast_by_cam(self, cam):
        """runRaycastByCam (slam_pipeline.cpp:362-415): free-view raycast at the engine's logged pose of that frame + tensor glue"""
        self.tsdf.runRaycast(syn.c2w_to_colmajor(cam.c2w_slam), self.intr)
        # fresh map buffers every time: the previous cycle's optimiser iterations may still be reading the old ones on the other stream
        self._release(cam)
        cam.depth_map = self.pool.get((self.H, self.W), self.sT)
        cam.color_map = self.pool.get((self.H, self.W, 3), self.sT)
        with self._glue():
            self.gs.raycast_maps(self.tsdf.GetFreeVertex(), self.tsdf.GetFreeImage(), cam.c2w, self.tsdf.getVoxelSize(), cam.depth_map,
                                 cam.color_map)

    def _key_frame_raycast(self):
        """keyFrameRaycast, sample_method == "random" (slam_pipeline.cpp:528-561): up to keyframe_select_max keyframes drawn
        without replacement.  Issued BEFORE the window raycasts (the reference does the window first): a free-view raycast does
        not modify the map, so the order is free, and this way the engine's vertex image still belongs to the newest window
        camera when the spawn reads it."""
        k = min(self.cfg["keyframe_select_max"], len(self.keyframes))
        pool = list(self.keyframes)
        picked = []
        for _ in range(k):
            i = self.rng.randrange(len(pool))
            cam = pool[i]
            pool[i] = pool[-1]
            pool.pop()
            self._raycast_by_cam(cam)
            picked.append(cam)
        self.opt_cams = list(self.window) + picked

    def _local_frame_raycast(self):
        """localFrameRaycast (slam_pipeline.cpp:417-448)"""
        for cam in self.window:
            self._raycast_by_cam(cam)

    def _init_new_gaussians(self):
        """initNewGaussians + addGaussians on the newest window camera (slam_pipeline.cpp:115, :450-526)"""
        cam = self.window[-1]
        c = self.cfg
        before = self.n_gauss
        extra = {}
        if self.exchange == "peer":
            extra = dict(rank=self.rank, world=self.world)       # the engine renders through the communicator itself
        elif self.world > 1:
            # the sample mask needs the render of ALL Gaussians: partial forward -> all-reduce -> composite
            if self.sp_rgb is None:
                self.sp_rgb, self.sp_depth, self.sp_alpha = self.pool.get((self.H, self.W, 3)), self.pool.get((self.H, self.W)), self.pool.get((self.H, self.W))
            self._forward_all(cam, self.sp_rgb, self.sp_depth, self.sp_alpha)
            extra = dict(rank=self.rank, world=self.world, render_rgb=self.sp_rgb, render_alpha=self.sp_alpha)
        self.gs.addGaussians(cam.c2w_slam, self.intr, self.tsdf.GetFreeVertex(), self.tsdf.getVoxelSize(), cam.depth_map, cam.color_map,
                             cam.image, seed=self.seed * 7919 + cam.id, color_error_thres=c["color_error_thres"],
                             depth_vis_min=c["depth_vis_min"], depth_vis_max=c["depth_vis_max"], alpha_vis_max=c["alpha_vis_max"],
                             sample_ratio=c["new_gs_sample_ratio"], max_init_scale=c["max_init_scale"], min_init_scale=c["min_init_scale"],
                             default_opacity=c["default_opacities"], **extra)
        self.n_gauss = self.gs.getGaussianNum()    # the one host round trip of the cycle (sizes the next 20 iterations' launches)
        self.spawned_last = self.n_gauss - before

    def _local_optimize(self):
        """localOptimize (slam_pipeline.cpp:195-291)"""
        self.gs.initOptimizers()
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
            if self.exchange == "nccl":
                self.gs.forward_partial(cam.c2w_slam, self.intr, cam.depth_map, self.acc5, True)
                self._allreduce_acc5()                  # the one collective of an iteration: [H,W,5] fp32 partial image
                self.gs.train_finish(cam.depth_map, cam.color_map, cam.image, self.acc5)
            else:
                self.gs.train_step(cam.c2w_slam, self.intr, cam.depth_map, cam.color_map, cam.image)

    def _allreduce_acc5(self):
        """sum of the per-rank partial images, enqueued on the Gaussian stream sG -- the stream forward_partial wrote acc5 on and
        train_finish / render_finish read it on -- whatever torch's current stream is in the caller"""
        with torch.cuda.stream(self.sG):
            parallel.allreduce_sum_(self.acc5)

    def _forward_all(self, cam, rgb, depth, alpha):
        """gesForward over the Gaussians of every rank"""
        if self.exchange == "nccl":
            self.gs.forward_partial(cam.c2w_slam, self.intr, cam.depth_map, self.acc5, False)
            self._allreduce_acc5()
            self.gs.render_finish(cam.depth_map, cam.color_map, self.acc5, rgb, depth, alpha)
        else:
            self.gs.forward(cam.c2w_slam, self.intr, cam.depth_map, cam.color_map, rgb, depth, alpha)

    def _remove_redundant(self):
        c = self.cfg
        # the next cycle re-creates the optimisers (localOptimize -> initOptimizers) before anything reads their state, so drop the
        # state now: the prune then has no Adam moments to carry along (gsb_gs_prune keeps them otherwise, like removeFromOptimizer)
        self.gs.initOptimizers()
        self.gs.prunePoints(c["low_opac_thres"], c["small_scale_thres"], c["large_scale_thres"])

    # ------------------------------------------------------------------------------------------------------------
    def end_of_step(self, resident):
        # step boundary: each stream waits for the other (device-side), so that events recorded on the main stream bracket all the
        # work of the step
        self._handover(self.sT, self.sG)
        self._handover(self.sG, self.sT)
        if not resident:
            # what a caller reads back after a cycle: the pose estimate (host-side state of the engine: 64 B) and the last loss
            # (slam_pipeline.cpp:81-82, progress bar :283).  The loss travels through pinned memory one cycle behind: the copy of cycle
            # k is enqueued here, the value of cycle k-1 -- long finished -- is collected, so the host keeps enqueueing ahead of the GPU
            # instead of draining it every cycle.  flush_readback() collects the last one.
            self._pose_host = self.tsdf.pose()[1]
            if self.gs and self.cycles:
                if self._loss_pending:
                    self.last_loss = self.gs.loss_end()
                self.gs.loss_begin()
                self._loss_pending = True

    def flush_readback(self):
        if self.gs and self._loss_pending:
            self.last_loss = self.gs.loss_end()
            self._loss_pending = False

    def render_eval(self, c2w, rgb, depth, alpha):
        """renderEvalImgs body for one camera (slam_pipeline.cpp:588-660): free-view raycast + gesForward"""
        cam = Cam(-1, np.asarray(c2w, np.float32), np.asarray(c2w, np.float32), None)
        with self.single_stream():
            self._raycast_by_cam(cam)
            self._forward_all(cam, rgb, depth, alpha)
            base = cam.color_map.clone()
        self._release(cam)
        return base

    def stats(self):
        s = {"visible_blocks_last_frame": self.tsdf.counter(2), "allocated_blocks": self.tsdf.num_blocks - 1 - self.tsdf.counter(0)}
        if self.gs:
            cnt = self.gs.counters()
            n = self.gs.getGaussianNum()
            if self.world > 1:
                t = torch.tensor([n], device=self.device, dtype=torch.int64)
                parallel.allreduce_sum_(t)
                s["gaussians_this_rank"] = n
                n = int(t.item())
            s.update(gaussians=n, keyframes=len(self.keyframes), opt_cameras=len(self.opt_cams),
                     last_isects=int(cnt[0]), last_visible=int(cnt[4]), overflow_flags=int(cnt[2]), cycles=self.cycles)
        return s

    def parallelism(self):
        if self.world == 1:
            return "single GPU"
        tsdf = ("voxel hash sharded by spatial block (owner = hashIndex(blockPos) mod %d): hash table / allocation replicated, each rank integrates "
                "the blocks it owns, raycasts split by image rows with peer-memory voxel reads and row stores into every rank, 2 flag barriers "
                "per frame" % self.world) if self.tsdf_sharded else "TSDF replicated"
        if self.exchange == "peer":
            return ("Gaussians sharded by spatial block over %d GPUs; per optimiser iteration the rasteriser stores tile partial sums into the owner "
                    "rank's memory over NVLink (reduce-scatter), the owner composites and stores dL/d(render) into every rank (all-gather), 2 flag "
                    "barriers through peer memory, no NCCL call; %s" % (self.world, tsdf))
        return ("Gaussians sharded by spatial block over %d GPUs, one NCCL all-reduce of the [H,W,5] partial image per optimiser iteration; %s"
                % (self.world, tsdf))

    def tracking_stats(self, poses, total):
        """online-tracking summary (BASELINE.json config 3): translation error of the tracked poses against the generator's, LM
        iterations (= ICP evaluations: one 29-accumulator reduction each) per frame"""
        if self.use_gt_pose or not self.track_err:
            return None
        e = np.asarray(self.track_err)
        it = np.asarray(self.track_iters, np.float64)
        return {"frames": int(len(e)), "ate_rmse_m": float(np.sqrt((e ** 2).mean())), "max_translation_error_m": float(e.max()),
                "final_translation_error_m": float(e[-1]), "icp_evaluations_per_frame": float(it.mean())}

    def io_bytes_per_step(self, frames_per_step):
        return frames_per_step * self.W * self.H * 6, 64 + (8 if self.gs else 0)

    def scaling(self):
        # the frame sequence is fixed; with N GPUs the same Gaussians and the same frames are split N ways
        return "strong"

    def _time(self, stream, fn, reps, flush):
        ms = []
        with torch.cuda.stream(stream):
            for _ in range(2):
                fn()
            for _ in range(reps):
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                fn()
                e1.record(stream)
                e1.synchronize()
                ms.append(e0.elapsed_time(e1))
        return float(np.mean(ms)) * 1e-3

    def time_dominant_kernel(self, stream, peak_gbs, reps=20, fresh_frames=None):
        with self.single_stream():
            return self._time_dominant_kernel(stream, peak_gbs, reps, fresh_frames)

    def _fresh_frame_stages(self, stream, fresh_frames):
        """device time of every ProcessFrame stage on FRESH frames (the frames that follow the timed window), CUDA events between
        the stages inside the engine: what the TSDF kernels cost inside the loop, where every frame changes every visible block"""
        rgba, depth, poses, f0, f1 = fresh_frames
        if f1 <= f0:
            return None
        self.tsdf.enable_stage_timing(True)
        rows, vis = [], []
        with torch.cuda.stream(stream):
            for f in range(f0, f1):
                c2w = np.asarray(poses[f], np.float32)
                self.tsdf.ProcessFrameDevice(rgba[f], depth[f], syn.c2w_to_colmajor(c2w) if self.use_gt_pose else None)
                rows.append(self.tsdf.stage_times())
                vis.append(self.tsdf.counter(2))
        self.tsdf.enable_stage_timing(False)
        m = np.asarray(rows).mean(0) * 1e3
        out = {k: float(v) for k, v in zip(("track", "allocate(6 kernels)", "integrate", "expected_depth(2)", "raycast", "icp_maps"), m)}
        out.update(frames=len(rows), visible_blocks=float(np.mean(vis)))
        if not self.use_gt_pose:
            out["icp_evaluations_per_frame"] = float(np.mean(self.track_iters[-len(rows):])) if self.track_iters else None
        return out

    def _time_dominant_kernel(self, stream, peak_gbs, reps=20, fresh_frames=None):
        """Per-kernel device times (CUDA events on the launching stream, L2 flushed by a 256 MiB fill between launches) and the
        roofline entry of the kernel BASELINE.json names: the rasteriser backward in train mode (algorithmic bytes
        24*P + 48*I + 80*N_vis), the TSDF integrate kernel in recon mode (V*(4+16+2*4096) + 8*P)."""
        P = self.W * self.H
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.device)
        V = self.tsdf.counter(2)
        table = {}
        for name, st in (("tsdf_allocate(6 kernels)", 0), ("tsdf_integrate", 1), ("tsdf_expected_depth(2)", 2), ("tsdf_raycast", 3),
                         ("tsdf_icp_maps", 4)):
            table[name] = self._time(stream, lambda st=st: self.tsdf.run_stage(st), reps, flush) * 1e6
        cams = [c for c in self.opt_cams if c.depth_map is not None]
        if cams:
            # free-view raycast as runRaycastByCam issues it (expected depths over the whole table + raycast + colour; sharded: plus its two
            # barriers and the row stores into every rank)
            c2w_free = syn.c2w_to_colmajor(cams[-1].c2w_slam)
            table["tsdf_free_view_raycast"] = self._time(stream, lambda: self.tsdf.runRaycast(c2w_free, self.intr), reps, flush) * 1e6
        fresh = self._fresh_frame_stages(stream, fresh_frames) if fresh_frames is not None else None
        if fresh:
            # the in-loop figure: fresh frames, events inside ProcessFrame (no L2 flush needed: every frame is new data and V x 8 KB
            # of voxel blocks stream through)
            t_int, Vf = fresh["integrate"] * 1e-6, fresh["visible_blocks"]
            note = "mean over %d FRESH frames processed after the timed window (CUDA events between the stages inside gsb_tsdf_process_frame_device)" % fresh["frames"]
        else:
            t_int, Vf = table["tsdf_integrate"] * 1e-6, V
            note = "timed by re-integrating the current frame (run_stage): a saturated map, few blocks change -- NOT the in-loop cost"
        alg_int = Vf * (4 + 16 + 2 * 4096) + 8 * P
        integrate = {"kernel": "k_integrate_tma", "bound": "hbm", "achieved": alg_int / t_int / 1e9, "peak": peak_gbs, "unit": "GB/s",
                     "frac": alg_int / t_int / 1e9 / peak_gbs, "traffic": None, "algorithmic_bytes": alg_int, "avg_launch_us": t_int * 1e6,
                     "units": {"visible_blocks": Vf, "pixels": P}, "note": note, "fresh_frame_stages_us": fresh}
        live = [c for c in self.opt_cams if c.depth_map is not None and c.image is not None]
        if self.mode != "train" or not live or self.n_gauss == 0:
            integrate["kernels_us"] = table
            return integrate
        g = self.gs
        cam = live[-1]
        with torch.cuda.stream(stream):
            g.initOptimizers()
            g.train_step(cam.c2w_slam, self.intr, cam.depth_map, cam.color_map, cam.image)
        for name, st in (("gs_project_sh", 0), ("gs_project_sh+bin_tiles(4 kernels)", 1), ("gs_raster_fwd_train", 2), ("gs_raster_bwd", 3)):
            table[name] = self._time(stream, lambda st=st: g.run_stage(st), reps, flush) * 1e6
        with torch.cuda.stream(stream):
            g.run_stage(3)   # stage 5 clears the work list, so re-arm it each time: time 3+5 and subtract
        def bwd_and_params():
            g.run_stage(1)
            g.run_stage(5)
        t15 = self._time(stream, bwd_and_params, reps, flush) * 1e6
        table["gs_bwd_params+adam_rest(2 kernels)"] = t15 - table["gs_project_sh+bin_tiles(4 kernels)"]
        with torch.cuda.stream(stream):
            g.run_stage(4)
        with torch.cuda.stream(stream):
            g.run_stage(1)          # a fresh work list for the pair statistics of the backward
            pairs_tested, pairs_passed = g.bwd_pair_stats()
            g.run_stage(4)
        table["gs_train_step(7 kernels, no flush)"] = self._time(
            stream, lambda: g.train_step(cam.c2w_slam, self.intr, cam.depth_map, cam.color_map, cam.image), reps, flush) * 1e6
        if self.comm is not None and self.world > 1:
            table["gs_exchange_barrier"] = self._time(stream, lambda: self.comm.barrier(stream.cuda_stream), reps, flush) * 1e6
        # where a 10-frame step goes, from the per-stage device times above (stages of the two streams overlap in the real loop, so the
        # parts add up to more than ms_per_step when the overlap works and to about ms_per_step when the SMs are the limit)
        c = self.cfg
        per_frame = sum(v for k, v in (fresh or {}).items() if k in ("track", "allocate(6 kernels)", "integrate", "expected_depth(2)", "raycast", "icp_maps"))
        n_free = len(self.opt_cams)
        self.breakdown = {"tsdf_fuse_ms": c["local_opt_interval"] * per_frame * 1e-3,
                          "tsdf_free_view_raycasts_ms": n_free * table.get("tsdf_free_view_raycast", 0.0) * 1e-3,
                          "gaussian_iterations_ms": c["local_opt_iters"] * table["gs_train_step(7 kernels, no flush)"] * 1e-3,
                          "of_which_exchange_barriers_ms": 2 * c["local_opt_iters"] * table.get("gs_exchange_barrier", 0.0) * 1e-3,
                          "free_view_raycasts_per_step": n_free, "frames_per_step": c["local_opt_interval"], "iterations_per_step": c["local_opt_iters"]}
        cnt = g.counters()
        I, n_vis = int(cnt[0]), int(cnt[4])
        t = table["gs_raster_bwd"] * 1e-6
        alg = 24 * P + 48 * I + 80 * n_vis
        ach = alg / t / 1e9
        return {"kernel": "k_raster_bwd", "bound": "hbm", "achieved": ach, "peak": peak_gbs, "unit": "GB/s", "frac": ach / peak_gbs,
                "traffic": None, "algorithmic_bytes": alg, "avg_launch_us": t * 1e6,
                "units": {"pixels": P, "isects": I, "visible_gaussians": n_vis, "gaussians": self.n_gauss,
                          "bwd_pairs_tested": pairs_tested, "bwd_pairs_passed": pairs_passed},
                "kernels_us": table, "step_breakdown": self.breakdown, "tsdf_integrate": integrate}
