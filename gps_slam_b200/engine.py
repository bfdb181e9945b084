"""ctypes binding of the C-ABI engine (include/gpsslam_b200.h) -- the Python-side mirror of the reference interfaces.

`TsdfEngine` follows ITMLib::ITMBasicEngine<ITMVoxel_s_rgb, ITMVoxelBlockHash> as the SLAM pipeline uses it
(reference InfiniTAM/ITMLib/Core/ITMBasicEngine.h:52-110): ProcessFrame, runRaycast, GetFreeImage/GetFreeVertex,
GetTrackingState()->pose_d, getVoxelSize, turnOffTracking (= tracker 0).

There is deliberately no fallback: if the shared library is missing or no CUDA device is present, construction raises.
torch is used only for device memory / streams by the callers; this module needs numpy + ctypes.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgpsslam_b200.so")

HASH_ENTRY = np.dtype([("pos", "<i2", 3), ("pad", "<i2"), ("offset", "<i4"), ("ptr", "<i4")])
VOXEL = np.dtype([("sdf", "<i2"), ("w_depth", "u1"), ("clr", "u1", 3), ("w_color", "u1"), ("pad", "u1")])

(HASH_TABLE, VOXELS, VISIBLE_IDS, VISIBLE_TYPES, DEPTH_F, MINMAX_LIVE, MINMAX_FREE, RAYCAST_LIVE, RAYCAST_FREE,
 POINTS_MAP, NORMALS_MAP, IMAGE_FREE) = range(12)


class TsdfConfig(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("voxel_size", C.c_float), ("mu", C.c_float),
                ("view_frustum_min", C.c_float), ("view_frustum_max", C.c_float),
                ("max_w", C.c_int), ("num_blocks", C.c_int), ("tracker", C.c_int), ("device", C.c_int),
                ("integrate_variant", C.c_int)]


class GsConfig(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("capacity", C.c_int), ("isect_capacity", C.c_int), ("item_capacity", C.c_int),
                ("max_gs_radii", C.c_int), ("delta_depth", C.c_float),
                ("eps2d", C.c_float), ("near_plane", C.c_float), ("far_plane", C.c_float), ("radius_clip", C.c_float),
                ("lr_means", C.c_float), ("lr_scales", C.c_float), ("lr_quats", C.c_float), ("lr_dc", C.c_float), ("lr_rest", C.c_float),
                ("lr_opac", C.c_float), ("scene_scale", C.c_float), ("device", C.c_int)]


class SpawnConfig(C.Structure):
    _fields_ = [("color_error_thres", C.c_float), ("depth_vis_min", C.c_float), ("depth_vis_max", C.c_float), ("alpha_vis_max", C.c_float),
                ("sample_ratio", C.c_float), ("max_init_scale", C.c_float), ("min_init_scale", C.c_float), ("default_opacity", C.c_float),
                ("seed", C.c_uint), ("rank", C.c_int), ("world", C.c_int), ("render_rgb_dev", C.c_void_p), ("render_alpha_dev", C.c_void_p)]


(GS_SPLAT_RECORDS, GS_SPLAT_GRADS, GS_TILE_OFFSETS, GS_FLATTEN_IDS, GS_V_OUT, GS_COUNTERS, GS_GRAD_MEANS, GS_GRAD_SCALES, GS_GRAD_QUATS,
 GS_GRAD_DC, GS_GRAD_REST, GS_GRAD_OPAC, GS_SPAWN_PIXELS) = range(13)


class EngineError(RuntimeError):
    pass


_lib = None


def load_library():
    """dlopen the engine; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.gsb_last_error.restype = C.c_char_p
    L.gsb_version.restype = C.c_char_p
    L.gsb_launch_count.restype = C.c_longlong
    L.gsb_tsdf_default_config.argtypes = [C.POINTER(TsdfConfig)]
    L.gsb_tsdf_create.argtypes = [C.POINTER(TsdfConfig), C.POINTER(C.c_void_p)]
    L.gsb_tsdf_destroy.argtypes = [C.c_void_p]
    L.gsb_tsdf_destroy.restype = None
    L.gsb_tsdf_create_sharded.argtypes = [C.POINTER(TsdfConfig), C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.gsb_tsdf_shard_export.argtypes = [C.c_void_p, C.c_void_p]
    L.gsb_tsdf_shard_attach.argtypes = [C.c_void_p, C.c_void_p]
    L.gsb_tsdf_shard_attach_local.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    L.gsb_tsdf_shard_error.argtypes = [C.c_void_p]
    L.gsb_tsdf_shard_set_mode.argtypes = [C.c_void_p, C.c_int]
    L.gsb_tsdf_shard_probe.argtypes = [C.c_void_p, C.c_int]
    L.gsb_tsdf_mesh.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.POINTER(C.c_longlong)]
    L.gsb_tsdf_shard_info.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 4
    L.gsb_tsdf_reset.argtypes = [C.c_void_p]
    L.gsb_tsdf_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.gsb_tsdf_get_stream.argtypes = [C.c_void_p]
    L.gsb_tsdf_get_stream.restype = C.c_void_p
    L.gsb_tsdf_sync.argtypes = [C.c_void_p]
    L.gsb_tsdf_process_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.gsb_tsdf_process_frame_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.gsb_tsdf_run_raycast.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float]
    for n in ("free_image_dev", "free_vertex_dev", "live_vertex_dev", "points_map_dev", "normals_map_dev"):
        f = getattr(L, "gsb_tsdf_" + n)
        f.argtypes = [C.c_void_p]
        f.restype = C.c_void_p
    L.gsb_tsdf_get_pose.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.gsb_tsdf_set_pose.argtypes = [C.c_void_p, C.c_void_p]
    L.gsb_tsdf_voxel_size.argtypes = [C.c_void_p]
    L.gsb_tsdf_voxel_size.restype = C.c_float
    L.gsb_tsdf_frames_processed.argtypes = [C.c_void_p]
    L.gsb_tsdf_frames_processed.restype = C.c_int
    L.gsb_tsdf_read.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
    L.gsb_tsdf_counter.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    L.gsb_tsdf_run_stage.argtypes = [C.c_void_p, C.c_int]
    L.gsb_tsdf_enable_stage_timing.argtypes = [C.c_void_p, C.c_int]
    L.gsb_tsdf_stage_times.argtypes = [C.c_void_p, C.c_void_p]
    L.gsb_tsdf_icp_eval.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float), C.c_void_p, C.c_void_p]
    L.gsb_tsdf_set_tracking_frames.argtypes = [C.c_void_p, C.c_int]
    L.gsb_tsdf_tracker_result.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_int)]
    L.gsb_tsdf_depth_level.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    # ---- section A: Gaussian model
    vp, fl = C.c_void_p, C.c_float
    L.gsb_gs_default_config.argtypes = [C.POINTER(GsConfig)]
    L.gsb_gs_create.argtypes = [C.POINTER(GsConfig), C.POINTER(vp)]
    L.gsb_gs_destroy.argtypes = [vp]
    L.gsb_gs_destroy.restype = None
    L.gsb_gs_set_stream.argtypes = [vp, vp]
    L.gsb_gs_sync.argtypes = [vp]
    L.gsb_gs_set_params.argtypes = [vp, C.c_int] + [vp] * 6
    L.gsb_gs_append.argtypes = [vp, C.c_int] + [vp] * 6
    L.gsb_gs_get_params.argtypes = [vp, C.c_int] + [vp] * 6
    L.gsb_gs_count.argtypes = [vp, C.POINTER(C.c_int)]
    L.gsb_gs_count_upper.argtypes = [vp]
    L.gsb_gs_init_optimizers.argtypes = [vp]
    L.gsb_gs_set_learning_rates.argtypes = [vp] + [fl] * 6
    L.gsb_gs_render.argtypes = [vp, vp, fl, fl, fl, fl, vp, vp, vp, vp, vp]
    L.gsb_gs_train_step.argtypes = [vp, vp, fl, fl, fl, fl, vp, vp, vp]
    L.gsb_gs_loss.argtypes = [vp, C.POINTER(C.c_double)]
    L.gsb_gs_loss_begin.argtypes = [vp]
    L.gsb_gs_loss_end.argtypes = [vp, C.POINTER(C.c_double)]
    L.gsb_gs_prune.argtypes = [vp, fl, fl, fl]
    L.gsb_gs_read.argtypes = [vp, C.c_int, vp, C.c_size_t]
    L.gsb_gs_enable_grad_dump.argtypes = [vp, C.c_int]
    L.gsb_gs_run_stage.argtypes = [vp, C.c_int]
    L.gsb_gs_spawn.argtypes = [vp, C.POINTER(SpawnConfig), vp, fl, fl, fl, fl, vp, fl, vp, vp, vp]
    L.gsb_gs_forward_partial.argtypes = [vp, vp, fl, fl, fl, fl, vp, vp, C.c_int]
    L.gsb_gs_render_finish.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    L.gsb_gs_train_finish.argtypes = [vp, vp, vp, vp, vp]
    L.gsb_comm_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.gsb_comm_export.argtypes = [vp, vp]
    L.gsb_comm_attach.argtypes = [vp, vp]
    L.gsb_comm_attach_local.argtypes = [vp, C.POINTER(vp)]
    L.gsb_comm_barrier.argtypes = [vp, vp]
    L.gsb_comm_error.argtypes = [vp]
    L.gsb_comm_destroy.argtypes = [vp]
    L.gsb_comm_destroy.restype = None
    L.gsb_gs_set_comm.argtypes = [vp, vp]
    L.gsb_mbox_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_size_t, C.POINTER(vp)]
    L.gsb_mbox_export.argtypes = [vp, vp]
    L.gsb_mbox_attach.argtypes = [vp, vp]
    L.gsb_mbox_attach_local.argtypes = [vp, C.POINTER(vp)]
    L.gsb_mbox_local.argtypes = [vp]
    L.gsb_mbox_local.restype = vp
    L.gsb_mbox_put.argtypes = [vp, C.c_int, C.c_size_t, vp, C.c_size_t, vp]
    L.gsb_mbox_signal.argtypes = [vp, C.c_int, C.c_int, C.c_uint, vp]
    L.gsb_mbox_wait.argtypes = [vp, C.c_int, C.c_uint, vp]
    L.gsb_mbox_error.argtypes = [vp]
    L.gsb_mbox_destroy.argtypes = [vp]
    L.gsb_mbox_destroy.restype = None
    L.gsb_gs_raycast_maps.argtypes = [vp, vp, vp, vp, fl, vp, vp, vp]
    L.gsb_gs_frame_to_float.argtypes = [vp, vp, vp, vp, vp]
    L.gsb_tsdf_current_rgba_dev.argtypes = [vp]
    L.gsb_tsdf_current_rgba_dev.restype = vp
    _lib = L
    return L


def launch_count():
    """kernels launched by the library so far (host-side count)"""
    return int(load_library().gsb_launch_count())


def _check(rc):
    if rc != 0:
        raise EngineError(load_library().gsb_last_error().decode())


def _ptr(x):
    """host numpy array / torch tensor / int -> void*"""
    if x is None:
        return None
    if isinstance(x, int):
        return C.c_void_p(x)
    if isinstance(x, np.ndarray):
        return C.c_void_p(x.ctypes.data)
    return C.c_void_p(x.data_ptr())  # torch tensor (host or device)


class TsdfEngine:
    def __init__(self, intr, voxel_size=0.005, mu=0.02, view_frustum_min=0.2, view_frustum_max=10.0,
                 tracker=0, device=0, num_blocks=0, integrate_variant=0, max_w=100, rank=0, world=1):
        """world > 1: the voxel hash is sharded by spatial block over `world` engines (gsb_tsdf_create_sharded); map the peers with
        export_handle() / attach(handles) between processes or attach_local(engines) inside one process before the first frame."""
        L = load_library()
        self.L = L
        cfg = TsdfConfig()
        L.gsb_tsdf_default_config(C.byref(cfg))
        cfg.width, cfg.height = intr["width"], intr["height"]
        cfg.fx, cfg.fy, cfg.cx, cfg.cy = intr["fx"], intr["fy"], intr["cx"], intr["cy"]
        cfg.voxel_size, cfg.mu = voxel_size, mu
        cfg.view_frustum_min, cfg.view_frustum_max = view_frustum_min, view_frustum_max
        cfg.tracker, cfg.device, cfg.integrate_variant, cfg.max_w = tracker, device, integrate_variant, max_w
        if num_blocks:
            cfg.num_blocks = num_blocks
        self.cfg = cfg
        self.w, self.h = cfg.width, cfg.height
        self.num_blocks = cfg.num_blocks
        self.E = 0x100000 + 0x20000
        self.rank, self.world = rank, world
        h = C.c_void_p()
        _check(L.gsb_tsdf_create_sharded(C.byref(cfg), rank, world, C.byref(h)))
        self.h_ = h

    def close(self):
        if getattr(self, "h_", None):
            self.L.gsb_tsdf_destroy(self.h_)
            self.h_ = None

    def mesh(self, max_tri=None):
        """SaveSceneToMesh up to the file: torch tensor [n, 18] on the engine's device (p0 p1 p2 in metres, c0 c1 c2 in 0..1), in the
        reference CPU mesher's order"""
        import torch
        n = C.c_longlong()
        _check(self.L.gsb_tsdf_mesh(self.h_, None, 0, C.byref(n)))
        cap = (n.value + 1) if max_tri is None else max_tri
        out = torch.empty((max(cap, 1), 18), dtype=torch.float32, device=torch.device("cuda", self.cfg.device))
        _check(self.L.gsb_tsdf_mesh(self.h_, C.c_void_p(out.data_ptr()), cap, C.byref(n)))
        return out[:n.value]

    # ---- sharded scene ----
    def export_handle(self):
        buf = C.create_string_buffer(64)
        _check(self.L.gsb_tsdf_shard_export(self.h_, buf))
        return buf.raw

    def attach(self, handles):
        blob = b"".join(handles)
        assert len(blob) == 64 * self.world
        _check(self.L.gsb_tsdf_shard_attach(self.h_, C.c_char_p(blob)))

    def attach_local(self, engines):
        arr = (C.c_void_p * self.world)(*[e.h_ for e in engines])
        _check(self.L.gsb_tsdf_shard_attach_local(self.h_, arr))

    def set_shard_mode(self, mode):
        """0: storage-sharded (peer reads in the raycast); 1: owner integrates and stores the block into every rank (local reads)"""
        _check(self.L.gsb_tsdf_shard_set_mode(self.h_, mode))

    def shard_error(self):
        return int(self.L.gsb_tsdf_shard_error(self.h_))

    def shard_rows(self):
        """rows [row0, row1) of the ICP maps this rank computes"""
        r0, r1 = C.c_int(), C.c_int()
        _check(self.L.gsb_tsdf_shard_info(self.h_, None, None, C.byref(r0), C.byref(r1)))
        return r0.value, r1.value

    __del__ = close

    # ---- ITMBasicEngine surface ----
    def ProcessFrame(self, rgba_host, depth_mm_host, gt_c2w=None):
        g = None if gt_c2w is None else np.ascontiguousarray(gt_c2w, dtype=np.float32)
        _check(self.L.gsb_tsdf_process_frame(self.h_, _ptr(rgba_host), _ptr(depth_mm_host), _ptr(g)))

    def ProcessFrameDevice(self, rgba_dev, depth_mm_dev, gt_c2w=None):
        g = None if gt_c2w is None else np.ascontiguousarray(gt_c2w, dtype=np.float32)
        _check(self.L.gsb_tsdf_process_frame_device(self.h_, _ptr(rgba_dev), _ptr(depth_mm_dev), _ptr(g)))

    def runRaycast(self, c2w, intr):
        g = np.ascontiguousarray(c2w, dtype=np.float32)
        _check(self.L.gsb_tsdf_run_raycast(self.h_, _ptr(g), intr["fx"], intr["fy"], intr["cx"], intr["cy"]))

    def GetFreeImage(self):
        return self.L.gsb_tsdf_free_image_dev(self.h_)

    def GetFreeVertex(self):
        return self.L.gsb_tsdf_free_vertex_dev(self.h_)

    def current_rgba(self):
        """device pointer of the RGBA frame the engine last consumed (its own upload buffer, or the caller's resident frame)"""
        return self.L.gsb_tsdf_current_rgba_dev(self.h_)

    def GetLiveVertex(self):
        return self.L.gsb_tsdf_live_vertex_dev(self.h_)

    def getVoxelSize(self):
        return self.L.gsb_tsdf_voxel_size(self.h_)

    def frames_processed(self):
        return self.L.gsb_tsdf_frames_processed(self.h_)

    def pose(self):
        M = np.zeros(16, np.float32)
        iM = np.zeros(16, np.float32)
        _check(self.L.gsb_tsdf_get_pose(self.h_, _ptr(M), _ptr(iM)))
        return M, iM

    def icp_eval(self, level, approx_invM16):
        a = np.ascontiguousarray(approx_invM16, dtype=np.float32)
        n, f = C.c_int(0), C.c_float(0)
        g, H = np.zeros(6, np.float32), np.zeros(36, np.float32)
        _check(self.L.gsb_tsdf_icp_eval(self.h_, level, _ptr(a), C.byref(n), C.byref(f), _ptr(g), _ptr(H)))
        return n.value, f.value, g, H

    def set_tracking_frames(self, n):
        _check(self.L.gsb_tsdf_set_tracking_frames(self.h_, n))

    def tracker_result(self):
        r, s, it = C.c_int(0), C.c_float(0), C.c_int(0)
        _check(self.L.gsb_tsdf_tracker_result(self.h_, C.byref(r), C.byref(s), C.byref(it)))
        return r.value, s.value, it.value

    def depth_level(self, level):
        w, h = C.c_int(0), C.c_int(0)
        _check(self.L.gsb_tsdf_depth_level(self.h_, level, None, C.byref(w), C.byref(h)))
        out = np.empty((h.value, w.value), np.float32)
        _check(self.L.gsb_tsdf_depth_level(self.h_, level, _ptr(out), C.byref(w), C.byref(h)))
        return out

    def set_pose(self, invM16):
        a = np.ascontiguousarray(invM16, dtype=np.float32)
        _check(self.L.gsb_tsdf_set_pose(self.h_, _ptr(a)))

    def load_scene(self, hash_entries, voxels, last_free_block, last_free_excess):
        """ITMBasicEngine::LoadFromFile minus the file I/O (see gps_slam_b200/checkpoint.py)"""
        h = np.ascontiguousarray(hash_entries)
        v = np.ascontiguousarray(voxels)
        self.L.gsb_tsdf_load_scene.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_int]
        _check(self.L.gsb_tsdf_load_scene(self.h_, h.ctypes.data, h.size, v.ctypes.data, v.size, last_free_block, last_free_excess))

    def resetAll(self):
        _check(self.L.gsb_tsdf_reset(self.h_))

    # ---- plumbing ----
    def set_stream(self, cuda_stream_ptr):
        _check(self.L.gsb_tsdf_set_stream(self.h_, C.c_void_p(cuda_stream_ptr)))

    def stream(self):
        return self.L.gsb_tsdf_get_stream(self.h_)

    def sync(self):
        _check(self.L.gsb_tsdf_sync(self.h_))

    def run_stage(self, stage):
        _check(self.L.gsb_tsdf_run_stage(self.h_, stage))

    def enable_stage_timing(self, on=True):
        _check(self.L.gsb_tsdf_enable_stage_timing(self.h_, 1 if on else 0))

    def stage_times(self):
        """device ms of track | allocate | integrate | expected depth | raycast | ICP maps of the last ProcessFrame (synchronises)"""
        ms = (C.c_float * 6)()
        _check(self.L.gsb_tsdf_stage_times(self.h_, ms))
        return [float(x) for x in ms]

    def counter(self, which):
        v = C.c_int(0)
        _check(self.L.gsb_tsdf_counter(self.h_, which, C.byref(v)))
        return v.value

    def read(self, what, dtype, shape):
        out = np.empty(shape, dtype=dtype)
        _check(self.L.gsb_tsdf_read(self.h_, what, _ptr(out), out.nbytes))
        return out

    # convenience views used by the parity tests
    def hash_entries(self):
        return self.read(HASH_TABLE, HASH_ENTRY, (self.E,))

    def voxels(self, nblocks=None):
        return self.read(VOXELS, VOXEL, (nblocks or self.num_blocks, 512))

    def visible_ids(self):
        n = self.counter(2)
        return self.read(VISIBLE_IDS, np.int32, (n,))

    def visible_types(self):
        return self.read(VISIBLE_TYPES, np.uint8, (self.E,))

    def depth(self):
        return self.read(DEPTH_F, np.float32, (self.h, self.w))

    def minmax(self, live=True):
        return self.read(MINMAX_LIVE if live else MINMAX_FREE, np.float32, ((self.h + 7) // 8, (self.w + 7) // 8, 2))

    def raycast(self, live=True):
        return self.read(RAYCAST_LIVE if live else RAYCAST_FREE, np.float32, (self.h, self.w, 4))

    def points_map(self):
        return self.read(POINTS_MAP, np.float32, (self.h, self.w, 4))

    def normals_map(self):
        return self.read(NORMALS_MAP, np.float32, (self.h, self.w, 4))

    def free_image(self):
        return self.read(IMAGE_FREE, np.uint8, (self.h, self.w, 4))


class PeerComm:
    """gsb_comm_t: this rank's exchange segment of the multi-GPU Gaussian path and the mapping of every peer's (csrc/gs_comm.h).
    Ranks in different processes exchange the 64-byte IPC handles with export_handle() / attach(handles); engines inside one process
    (tests) use attach_local(list of PeerComm in rank order)."""

    def __init__(self, device, rank, world, width, height):
        self.L = load_library()
        self.rank, self.world = rank, world
        h = C.c_void_p()
        _check(self.L.gsb_comm_create(device, rank, world, width, height, C.byref(h)))
        self.h_ = h

    def export_handle(self):
        buf = C.create_string_buffer(64)
        _check(self.L.gsb_comm_export(self.h_, buf))
        return buf.raw

    def attach(self, handles):
        blob = b"".join(handles)
        assert len(blob) == 64 * self.world
        _check(self.L.gsb_comm_attach(self.h_, C.c_char_p(blob)))

    def attach_local(self, comms):
        arr = (C.c_void_p * self.world)(*[c.h_ for c in comms])
        _check(self.L.gsb_comm_attach_local(self.h_, arr))

    def barrier(self, cuda_stream_ptr):
        _check(self.L.gsb_comm_barrier(self.h_, C.c_void_p(cuda_stream_ptr)))

    def error(self):
        return int(self.L.gsb_comm_error(self.h_))

    def close(self):
        if self.h_:
            self.L.gsb_comm_destroy(self.h_)
            self.h_ = None


class PeerMailbox:
    """gsb_mbox_t: a byte segment per rank that the other ranks write into over NVLink + counters for stream hand-shakes
    (csrc/peer_mbox.cu).  export_handle() / attach(handles) between processes, attach_local(list) inside one process."""

    def __init__(self, device, rank, world, nbytes):
        self.L = load_library()
        self.rank, self.world, self.nbytes = rank, world, nbytes
        h = C.c_void_p()
        _check(self.L.gsb_mbox_create(device, rank, world, nbytes, C.byref(h)))
        self.h_ = h

    def export_handle(self):
        buf = C.create_string_buffer(64)
        _check(self.L.gsb_mbox_export(self.h_, buf))
        return buf.raw

    def attach(self, handles):
        blob = b"".join(handles)
        assert len(blob) == 64 * self.world
        _check(self.L.gsb_mbox_attach(self.h_, C.c_char_p(blob)))

    def attach_local(self, boxes):
        arr = (C.c_void_p * self.world)(*[b.h_ for b in boxes])
        _check(self.L.gsb_mbox_attach_local(self.h_, arr))

    def local_ptr(self):
        return int(self.L.gsb_mbox_local(self.h_))

    def put(self, dst_rank, dst_offset, src, nbytes, cuda_stream_ptr):
        _check(self.L.gsb_mbox_put(self.h_, dst_rank, dst_offset, _ptr(src), nbytes, C.c_void_p(cuda_stream_ptr)))

    def signal(self, dst_rank, flag, value, cuda_stream_ptr):
        _check(self.L.gsb_mbox_signal(self.h_, dst_rank, flag, value & 0xffffffff, C.c_void_p(cuda_stream_ptr)))

    def wait(self, flag, value, cuda_stream_ptr):
        _check(self.L.gsb_mbox_wait(self.h_, flag, value & 0xffffffff, C.c_void_p(cuda_stream_ptr)))

    def error(self):
        return int(self.L.gsb_mbox_error(self.h_))

    def close(self):
        if self.h_:
            self.L.gsb_mbox_destroy(self.h_)
            self.h_ = None


PARAM_NAMES = ("means", "scales", "quats", "featuresDc", "featuresRest", "opacities")
PARAM_WIDTH = dict(means=3, scales=3, quats=4, featuresDc=3, featuresRest=45, opacities=1)


class GaussianEngine:
    """RawGaussianModel / SLAMGaussianModel with render_method "ges" (reference include/raw_gs_model.h:8-298): same parameter
    tensors (means, scales=log, quats=wxyz, featuresDc, featuresRest [N,15,3], opacities=logit [N,1]), forward(cam, ref_depth,
    base_color), one fused optimiser iteration (forward + computeLoss + backward + optimizersStep), initOptimizers, prunePoints,
    add.  Arguments are numpy arrays (copied) or torch CUDA tensors (used in place)."""

    def __init__(self, width, height, capacity=1 << 21, device=0, **overrides):
        L = load_library()
        self.L = L
        cfg = GsConfig()
        L.gsb_gs_default_config(C.byref(cfg))
        cfg.width, cfg.height, cfg.capacity, cfg.device = width, height, capacity, device
        for k, v in overrides.items():
            if not hasattr(cfg, k):
                raise EngineError("unknown Gaussian-engine option %r" % k)
            setattr(cfg, k, v)
        self.cfg = cfg
        self.W, self.H = width, height
        self.tile_w, self.tile_h = (width + 15) // 16, (height + 15) // 16
        h = C.c_void_p()
        _check(L.gsb_gs_create(C.byref(cfg), C.byref(h)))
        self.h_ = h

    def close(self):
        if getattr(self, "h_", None):
            self.L.gsb_gs_destroy(self.h_)
            self.h_ = None

    __del__ = close

    def set_stream(self, cuda_stream_ptr):
        _check(self.L.gsb_gs_set_stream(self.h_, C.c_void_p(cuda_stream_ptr)))

    def sync(self):
        _check(self.L.gsb_gs_sync(self.h_))

    @staticmethod
    def _arr(params):
        out = []
        for k in PARAM_NAMES:
            a = params.get(k)
            if a is None:
                out.append(None)
            elif isinstance(a, np.ndarray):
                out.append(np.ascontiguousarray(a, dtype=np.float32))
            else:
                out.append(a.contiguous())
        return out

    def set_params(self, params):
        a = self._arr(params)
        n = int(a[0].shape[0])
        _check(self.L.gsb_gs_set_params(self.h_, n, *[_ptr(x) for x in a]))

    def add(self, params):
        a = self._arr(params)
        n = int(a[0].shape[0])
        _check(self.L.gsb_gs_append(self.h_, n, *[_ptr(x) for x in a]))

    def getGaussianNum(self):
        v = C.c_int(0)
        _check(self.L.gsb_gs_count(self.h_, C.byref(v)))
        return v.value

    def get_params(self):
        n = self.getGaussianNum()
        out = {k: np.empty((n, PARAM_WIDTH[k]), np.float32) for k in PARAM_NAMES}
        _check(self.L.gsb_gs_get_params(self.h_, n, *[_ptr(out[k]) for k in PARAM_NAMES]))
        out["featuresRest"] = out["featuresRest"].reshape(n, 15, 3)
        return out

    def initOptimizers(self):
        _check(self.L.gsb_gs_init_optimizers(self.h_))

    def set_learning_rates(self, means, scales, quats, dc, rest, opac):
        _check(self.L.gsb_gs_set_learning_rates(self.h_, means, scales, quats, dc, rest, opac))

    @staticmethod
    def _cam(c2w):
        return np.ascontiguousarray(c2w, dtype=np.float32).reshape(16)

    def forward(self, c2w, intr, ref_depth_dev, base_color_dev, rgb_out, depth_out, alpha_out):
        """gesForward without autograd; all image arguments are device tensors / pointers"""
        c = self._cam(c2w)
        _check(self.L.gsb_gs_render(self.h_, _ptr(c), intr["fx"], intr["fy"], intr["cx"], intr["cy"], _ptr(ref_depth_dev), _ptr(base_color_dev),
                                    _ptr(rgb_out), _ptr(depth_out), _ptr(alpha_out)))

    def train_step(self, c2w, intr, ref_depth_dev, base_color_dev, gt_rgb_dev):
        c = self._cam(c2w)
        _check(self.L.gsb_gs_train_step(self.h_, _ptr(c), intr["fx"], intr["fy"], intr["cx"], intr["cy"], _ptr(ref_depth_dev),
                                        _ptr(base_color_dev), _ptr(gt_rgb_dev)))

    # ---- multi-GPU (Gaussians sharded across ranks): partial forward -> all-reduce by the caller -> finish
    def set_comm(self, comm):
        """attach a PeerComm: forward / train_step / addGaussians then run the whole multi-GPU iteration (exchange below the C ABI)"""
        _check(self.L.gsb_gs_set_comm(self.h_, comm.h_ if comm is not None else None))
        self.comm = comm

    def forward_partial(self, c2w, intr, ref_depth_dev, acc5, for_backward):
        c = self._cam(c2w)
        _check(self.L.gsb_gs_forward_partial(self.h_, _ptr(c), intr["fx"], intr["fy"], intr["cx"], intr["cy"], _ptr(ref_depth_dev), _ptr(acc5),
                                             int(for_backward)))

    def render_finish(self, ref_depth_dev, base_color_dev, acc5, rgb_out, depth_out, alpha_out):
        _check(self.L.gsb_gs_render_finish(self.h_, _ptr(ref_depth_dev), _ptr(base_color_dev), _ptr(acc5), _ptr(rgb_out), _ptr(depth_out),
                                           _ptr(alpha_out)))

    def train_finish(self, ref_depth_dev, base_color_dev, gt_rgb_dev, acc5):
        _check(self.L.gsb_gs_train_finish(self.h_, _ptr(ref_depth_dev), _ptr(base_color_dev), _ptr(gt_rgb_dev), _ptr(acc5)))

    def loss(self):
        v = C.c_double(0)
        _check(self.L.gsb_gs_loss(self.h_, C.byref(v)))
        return v.value

    def loss_begin(self):
        """enqueue the read-back of the last step's loss (no host stall); loss_end() returns it"""
        _check(self.L.gsb_gs_loss_begin(self.h_))

    def loss_end(self):
        v = C.c_double(0)
        _check(self.L.gsb_gs_loss_end(self.h_, C.byref(v)))
        return v.value

    def prunePoints(self, min_opac, min_scale, max_scale):
        _check(self.L.gsb_gs_prune(self.h_, min_opac, min_scale, max_scale))

    def raycast_maps(self, free_vertex_ptr, free_image_ptr, c2w, voxel_size, depth_map, color_map, conf_map=None):
        """runRaycastByCam glue: TSDF free-view vertex/colour images -> depth_map [H,W], color_map [H,W,3] (device tensors)"""
        c = self._cam(c2w)
        _check(self.L.gsb_gs_raycast_maps(self.h_, _ptr(free_vertex_ptr), _ptr(free_image_ptr), _ptr(c), voxel_size, _ptr(depth_map),
                                          _ptr(color_map), _ptr(conf_map)))

    def frame_to_float(self, rgba_ptr, depth_mm_ptr, rgb_out, depth_out=None):
        _check(self.L.gsb_gs_frame_to_float(self.h_, _ptr(rgba_ptr), _ptr(depth_mm_ptr), _ptr(rgb_out), _ptr(depth_out)))

    def addGaussians(self, c2w, intr, free_vertex_ptr, voxel_size, depth_map, color_map, gt_rgb, seed, color_error_thres=0.05,
                     depth_vis_min=0.0, depth_vis_max=5.0, alpha_vis_max=5.0, sample_ratio=0.25, max_init_scale=0.01, min_init_scale=-1.0,
                     default_opacity=0.5, rank=0, world=1, render_rgb=None, render_alpha=None):
        """initNewGaussians + addGaussians (defaults: configs/release/replica/office0.yaml).  Multi-GPU: pass the all-reduced render
        of the camera and (rank, world); each rank then spawns only the Gaussians of its own spatial blocks."""
        sc = SpawnConfig(color_error_thres, depth_vis_min, depth_vis_max, alpha_vis_max, sample_ratio, max_init_scale, min_init_scale,
                         default_opacity, seed & 0xffffffff, rank, world,
                         None if render_rgb is None else render_rgb.data_ptr(), None if render_alpha is None else render_alpha.data_ptr())
        c = self._cam(c2w)
        _check(self.L.gsb_gs_spawn(self.h_, C.byref(sc), _ptr(c), intr["fx"], intr["fy"], intr["cx"], intr["cy"], _ptr(free_vertex_ptr),
                                   voxel_size, _ptr(depth_map), _ptr(color_map), _ptr(gt_rgb)))

    def spawn_pixels(self, n):
        """pixel index of each of the n Gaussians the last addGaussians appended (append order = raster order)"""
        return self.read(GS_SPAWN_PIXELS, np.int32, (n,)) if n > 0 else np.zeros(0, np.int32)

    def run_stage(self, stage):
        _check(self.L.gsb_gs_run_stage(self.h_, stage))

    def enable_grad_dump(self, on=True):
        _check(self.L.gsb_gs_enable_grad_dump(self.h_, int(on)))

    def read(self, what, dtype, shape):
        out = np.empty(shape, dtype=dtype)
        _check(self.L.gsb_gs_read(self.h_, what, _ptr(out), out.nbytes))
        return out

    def counters(self):
        return self.read(GS_COUNTERS, np.int32, (8,))

    def bwd_pair_stats(self):
        """(pixel, splat) pairs the rasteriser backward evaluates / that pass its tests, for the work list of the last train step
        (run_stage 6: the backward kernel with its counters compiled in)"""
        self.run_stage(6)
        c = self.read(GS_COUNTERS, np.int32, (16,))
        t = c[8:12].view(np.uint64)
        return int(t[0]), int(t[1])

    def splat_records(self, n):
        r = self.read(GS_SPLAT_RECORDS, np.float32, (n, 12))
        radii = r[:, 3].copy().view(np.int32)
        vis = radii > 0
        z = lambda a: np.where(vis[:, None] if a.ndim == 2 else vis, a, 0)
        return dict(radii=radii, means2d=z(r[:, 0:2]), opacities=z(r[:, 2]), conics=z(r[:, 4:7]), depths=z(r[:, 7]), colors=z(r[:, 8:11]),
                    flags=np.where(vis, r[:, 11].copy().view(np.int32), 0))

    def splat_grads(self, n):
        g = self.read(GS_SPLAT_GRADS, np.float32, (n, 12))
        return dict(v_means2d=g[:, 0:2], v_opacities=g[:, 2], v_depths=g[:, 3], v_conics=g[:, 4:7], v_colors=g[:, 8:11])

    def tile_bins(self):
        off = self.read(GS_TILE_OFFSETS, np.int32, (self.tile_w * self.tile_h + 1,))
        ids = self.read(GS_FLATTEN_IDS, np.int32, (int(off[-1]),)) if off[-1] > 0 else np.zeros(0, np.int32)
        return off, ids

    def v_out(self):
        """dL/d render (rgb, alpha) per pixel; the engine stores 8 floats per pixel (gradient | depth cut, padding)"""
        return np.ascontiguousarray(self.read(GS_V_OUT, np.float32, (self.H, self.W, 8))[..., :4])

    def param_grads(self, n):
        ids = dict(means=GS_GRAD_MEANS, scales=GS_GRAD_SCALES, quats=GS_GRAD_QUATS, featuresDc=GS_GRAD_DC, featuresRest=GS_GRAD_REST,
                   opacities=GS_GRAD_OPAC)
        out = {k: self.read(ids[k], np.float32, (n, PARAM_WIDTH[k])) for k in PARAM_NAMES}
        out["featuresRest"] = out["featuresRest"].reshape(n, 15, 3)
        return out
