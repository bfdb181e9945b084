"""TEST INFRASTRUCTURE ONLY -- ctypes access to oracle/_ref/libitm_ref_{exact,fast}.so, i.e. the
REFERENCE's own InfiniTAM CPU engine (built by oracle/itm_ref/Makefile from /root/reference).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.
"""
import contextlib
import ctypes as C
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
HASH_ENTRY = np.dtype([("pos", "<i2", 3), ("pad", "<i2"), ("offset", "<i4"), ("ptr", "<i4")])
VOXEL = np.dtype([("sdf", "<i2"), ("w_depth", "u1"), ("clr", "u1", 3), ("w_color", "u1"), ("pad", "u1")])
assert HASH_ENTRY.itemsize == 16 and VOXEL.itemsize == 8


@contextlib.contextmanager
def _quiet_stdout():
    """the reference prints its settings with printf; keep them off our stdout (bench.py prints exactly one JSON line)"""
    sys.stdout.flush()
    saved = os.dup(1)
    null = os.open(os.devnull, os.O_WRONLY)
    try:
        os.dup2(null, 1)
        yield
    finally:
        try:
            C.CDLL(None).fflush(None)
        except Exception:
            pass
        os.dup2(saved, 1)
        os.close(null)
        os.close(saved)


def lib_path(kind="exact"):
    return os.path.join(_HERE, "_ref", "libitm_ref_%s.so" % kind)


def available(kind="exact"):
    return os.path.exists(lib_path(kind))


class ItmRef:
    """tracker: 0 = ground-truth poses, 1 = extended tracker (reference default), 2 = icp (ITMDepthTracker)."""

    def __init__(self, intr, voxel=0.005, mu=0.02, vfmin=0.2, vfmax=10.0, tracker=0, threads=1, kind="exact"):
        L = C.CDLL(lib_path(kind))
        self.L = L
        L.itmref_create.restype = C.c_void_p
        L.itmref_create.argtypes = [C.c_int, C.c_int] + [C.c_float] * 8 + [C.c_int, C.c_int]
        for name in ("hash_entries", "voxels", "visible_types", "depth", "points_map", "normals_map"):
            f = getattr(L, "itmref_" + name)
            f.restype = C.c_void_p
            f.argtypes = [C.c_void_p]
        for name in ("minmax", "raycast", "raycast_image"):
            f = getattr(L, "itmref_" + name)
            f.restype = C.c_void_p
            f.argtypes = [C.c_void_p, C.c_int]
        L.itmref_visible_ids.restype = C.c_void_p
        L.itmref_visible_ids.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.itmref_process_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.itmref_pose.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.itmref_free_pose.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.itmref_set_pose_invM.argtypes = [C.c_void_p, C.c_void_p]
        L.itmref_run_raycast.argtypes = [C.c_void_p, C.c_void_p] + [C.c_float] * 4
        L.itmref_icp_eval.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.itmref_set_tracking_frames.argtypes = [C.c_void_p, C.c_int]
        L.itmref_depth_level.restype = C.c_void_p
        L.itmref_depth_level.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        for name in ("destroy", "num_hash_entries", "num_blocks", "last_free_block", "last_free_excess",
                     "tracker_result", "frames_processed"):
            getattr(L, "itmref_" + name).argtypes = [C.c_void_p]
        for name in ("save", "load"):
            f = getattr(L, "itmref_" + name)
            f.argtypes = [C.c_void_p, C.c_char_p]
        self.w, self.h = intr["width"], intr["height"]
        with _quiet_stdout():
            self.h_ = L.itmref_create(self.w, self.h, intr["fx"], intr["fy"], intr["cx"], intr["cy"],
                                      voxel, mu, vfmin, vfmax, tracker, threads)
        self.E = L.itmref_num_hash_entries(self.h_)
        self.nblocks = L.itmref_num_blocks(self.h_)

    def close(self):
        if self.h_:
            self.L.itmref_destroy(self.h_)
            self.h_ = None

    def _arr(self, ptr, dtype, shape):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        buf = (C.c_char * n).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    def process_frame(self, rgba, depth_mm, c2w_colmajor=None):
        rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
        depth_mm = np.ascontiguousarray(depth_mm, dtype=np.int16)
        p = None
        if c2w_colmajor is not None:
            c2w_colmajor = np.ascontiguousarray(c2w_colmajor, dtype=np.float32)
            p = c2w_colmajor.ctypes.data
        return self.L.itmref_process_frame(self.h_, rgba.ctypes.data, depth_mm.ctypes.data, p)

    # ---- state views (no copies; valid until the next call) ----
    def hash_entries(self):
        return self._arr(self.L.itmref_hash_entries(self.h_), HASH_ENTRY, (self.E,))

    def voxels(self):
        return self._arr(self.L.itmref_voxels(self.h_), VOXEL, (self.nblocks, 512))

    def visible_ids(self, live=True):
        n = C.c_int(0)
        p = self.L.itmref_visible_ids(self.h_, int(live), C.byref(n))
        return self._arr(p, np.int32, (n.value,))

    def visible_types(self):
        return self._arr(self.L.itmref_visible_types(self.h_), np.uint8, (self.E,))

    def depth(self):
        return self._arr(self.L.itmref_depth(self.h_), np.float32, (self.h, self.w))

    def minmax(self, live=True):
        return self._arr(self.L.itmref_minmax(self.h_, int(live)), np.float32, (self.h, self.w, 2))

    def raycast(self, live=True):
        return self._arr(self.L.itmref_raycast(self.h_, int(live)), np.float32, (self.h, self.w, 4))

    def raycast_image(self, live=True):
        return self._arr(self.L.itmref_raycast_image(self.h_, int(live)), np.uint8, (self.h, self.w, 4))

    def points_map(self):
        return self._arr(self.L.itmref_points_map(self.h_), np.float32, (self.h, self.w, 4))

    def normals_map(self):
        return self._arr(self.L.itmref_normals_map(self.h_), np.float32, (self.h, self.w, 4))

    def last_free_block(self):
        return self.L.itmref_last_free_block(self.h_)

    def last_free_excess(self):
        return self.L.itmref_last_free_excess(self.h_)

    def pose(self):
        M = np.zeros(16, np.float32)
        iM = np.zeros(16, np.float32)
        self.L.itmref_pose(self.h_, M.ctypes.data, iM.ctypes.data)
        return M, iM

    def set_pose_invM(self, invM16):
        a = np.ascontiguousarray(invM16, dtype=np.float32)
        self.L.itmref_set_pose_invM(self.h_, a.ctypes.data)

    def run_raycast(self, c2w_colmajor, intr):
        a = np.ascontiguousarray(c2w_colmajor, dtype=np.float32)
        self.L.itmref_run_raycast(self.h_, a.ctypes.data, intr["fx"], intr["fy"], intr["cx"], intr["cy"])

    def free_pose(self):
        M = np.zeros(16, np.float32)
        iM = np.zeros(16, np.float32)
        self.L.itmref_free_pose(self.h_, M.ctypes.data, iM.ctypes.data)
        return M, iM

    def icp_eval(self, level, approx_invM16):
        a = np.ascontiguousarray(approx_invM16, dtype=np.float32)
        f = np.zeros(1, np.float32)
        g = np.zeros(6, np.float32)
        H = np.zeros(36, np.float32)
        n = self.L.itmref_icp_eval(self.h_, level, a.ctypes.data, f.ctypes.data, g.ctypes.data, H.ctypes.data)
        return n, float(f[0]), g, H

    def set_tracking_frames(self, n):
        self.L.itmref_set_tracking_frames(self.h_, n)

    def depth_level(self, level):
        w, h = C.c_int(0), C.c_int(0)
        p = self.L.itmref_depth_level(self.h_, level, C.byref(w), C.byref(h))
        return self._arr(p, np.float32, (h.value, w.value))

    def mesh(self, max_tri=4_000_000):
        """triangles of the reference's marching-cubes export [n, 18] = p0 p1 p2 (metres) c0 c1 c2 (0..1), hash-entry order"""
        self.L.itmref_mesh.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        out = np.zeros((max_tri, 18), np.float32)
        n = self.L.itmref_mesh(self.h_, out.ctypes.data_as(C.c_void_p), max_tri)
        return out[:n].copy()

    def tracker_result(self):
        return self.L.itmref_tracker_result(self.h_)

    def save(self, directory):
        """ITMBasicEngine::SaveToFile: writes <directory>/Scene/{hash,excess,voxel,alloc}.dat, last.txt, vba.txt"""
        assert directory.endswith("/")
        with _quiet_stdout():
            assert self.L.itmref_save(self.h_, directory.encode()) == 0

    def load(self, directory):
        """ITMBasicEngine::LoadFromFile"""
        assert directory.endswith("/")
        with _quiet_stdout():
            assert self.L.itmref_load(self.h_, directory.encode()) == 0
