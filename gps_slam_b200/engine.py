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
    L.gsb_tsdf_voxel_size.argtypes = [C.c_void_p]
    L.gsb_tsdf_voxel_size.restype = C.c_float
    L.gsb_tsdf_frames_processed.argtypes = [C.c_void_p]
    L.gsb_tsdf_read.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
    L.gsb_tsdf_counter.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    L.gsb_tsdf_run_stage.argtypes = [C.c_void_p, C.c_int]
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
                 tracker=0, device=0, num_blocks=0, integrate_variant=0, max_w=100):
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
        h = C.c_void_p()
        _check(L.gsb_tsdf_create(C.byref(cfg), C.byref(h)))
        self.h_ = h

    def close(self):
        if getattr(self, "h_", None):
            self.L.gsb_tsdf_destroy(self.h_)
            self.h_ = None

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

    def GetLiveVertex(self):
        return self.L.gsb_tsdf_live_vertex_dev(self.h_)

    def getVoxelSize(self):
        return self.L.gsb_tsdf_voxel_size(self.h_)

    def pose(self):
        M = np.zeros(16, np.float32)
        iM = np.zeros(16, np.float32)
        _check(self.L.gsb_tsdf_get_pose(self.h_, _ptr(M), _ptr(iM)))
        return M, iM

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
