"""Deterministic synthetic RGB-D sequence of Replica shape.

The datasets the reference runs on (Replica, GPS_SLAM Indoor) are not in the container and there is no
network, so every test and bench consumes this generator instead of DatasetReader::read
(reference src/dataset_reader.cpp:269-369).  Frame format is what createTsdfEngine hands to the TSDF
engine (reference slam/InfiniTAM_tools.cpp:33-45): RGBA uint8 [H,W,4], depth uint16 millimetres [H,W]
(stored as int16 like ITMShortImage), and a 4x4 camera-to-world pose.

Scene: an axis-aligned 6.0 x 3.0 x 4.5 m room (office scale) with 6 boxes, analytic ray/box depth,
procedural albedo = seeded sum of 8 sinusoids per channel times a 12 cm checker, so that a 5 mm TSDF
colour volume leaves high-frequency residual for the Gaussians to fit.  Camera: pinhole, looks along +z,
x right, y down (InfiniTAM / gsplat convention), on a smooth Lissajous orbit (~1 cm, ~0.5 deg per frame).

torch is used only as an array library here (CPU in tests, CUDA in bench.py for speed).
"""
import math

import numpy as np
import torch

REPLICA = dict(width=1200, height=680, fx=600.0, fy=600.0, cx=599.5, cy=339.5)
KINECT = dict(width=1280, height=720, fx=605.37, fy=605.25, cx=635.31, cy=366.51)

ROOM_MIN = (0.0, 0.0, 0.0)
ROOM_MAX = (6.0, 3.0, 4.5)
# (min xyz, max xyz) of the 6 boxes; y is "down" so boxes sit on y = 3.0 (the floor)
BOXES = [
    ((0.8, 2.2, 0.6), (1.8, 3.0, 1.4)),
    ((4.2, 1.9, 0.5), (5.4, 3.0, 1.5)),
    ((2.4, 2.5, 3.2), (3.6, 3.0, 4.1)),
    ((0.3, 1.6, 3.3), (1.1, 3.0, 4.2)),
    ((4.8, 2.0, 3.0), (5.7, 3.0, 3.9)),
    ((2.6, 2.3, 1.9), (3.4, 3.0, 2.6)),
]


def intrinsics(name="replica", scale=1.0):
    base = dict(REPLICA if name == "replica" else KINECT)
    if scale != 1.0:
        base = dict(width=int(round(base["width"] * scale)), height=int(round(base["height"] * scale)),
                    fx=base["fx"] * scale, fy=base["fy"] * scale,
                    cx=(base["cx"] + 0.5) * scale - 0.5, cy=(base["cy"] + 0.5) * scale - 0.5)
    return base


def _look_at(eye, target):
    """camera-to-world with +z forward, +x right, +y down; world 'down' is +y."""
    f = target - eye
    f = f / np.linalg.norm(f)
    down = np.array([0.0, 1.0, 0.0])
    r = np.cross(down, f)
    r = r / np.linalg.norm(r)
    d = np.cross(f, r)
    c2w = np.eye(4, dtype=np.float64)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = r, d, f, eye
    return c2w


def trajectory(n_frames, seed=42):
    """[n,4,4] float32 camera-to-world (row-major), smooth Lissajous orbit around the room centre."""
    rng = np.random.RandomState(seed)
    ph = rng.uniform(0, 2 * math.pi, size=4)
    out = np.zeros((n_frames, 4, 4), dtype=np.float32)
    c = np.array([3.0, 1.5, 2.25])
    for i in range(n_frames):
        t = i * 0.004
        eye = c + np.array([1.6 * math.sin(1.0 * t + ph[0]), 0.35 * math.sin(1.7 * t + ph[1]),
                            1.1 * math.sin(1.3 * t + ph[2])])
        ang = 0.9 * t + ph[3]
        tgt = c + np.array([2.6 * math.cos(ang), 0.6 + 0.3 * math.sin(0.7 * t), 1.9 * math.sin(ang)])
        out[i] = _look_at(eye, tgt).astype(np.float32)
    return out


def _albedo(p, seed):
    """procedural colour in [0,1]^3 for world points p [...,3]"""
    g = torch.Generator().manual_seed(seed)
    k = (torch.rand(3, 8, 3, generator=g) * 2 - 1) * 9.0       # spatial frequencies (rad/m)
    ph = torch.rand(3, 8, generator=g) * 2 * math.pi
    amp = torch.rand(3, 8, generator=g) * 0.5 + 0.5
    k, ph, amp = k.to(p), ph.to(p), amp.to(p)
    arg = torch.einsum("...d,ckd->...ck", p, k) + ph
    s = (torch.sin(arg) * amp).sum(-1) / amp.sum(-1)             # [...,3] in [-1,1]
    chk = (torch.floor(p / 0.12).sum(-1) % 2.0)                  # 12 cm checker
    col = 0.5 + 0.32 * s + 0.14 * (chk[..., None] - 0.5)
    return col.clamp(0.02, 0.98)


def render_frame(c2w, intr, seed=42, device="cpu"):
    """-> (rgba uint8 [H,W,4], depth int16 mm [H,W]) for one camera-to-world pose (4x4 array-like)."""
    W, H = intr["width"], intr["height"]
    dt = torch.float64
    c2w = torch.as_tensor(np.asarray(c2w, dtype=np.float64), dtype=dt, device=device)
    ys, xs = torch.meshgrid(torch.arange(H, device=device, dtype=dt), torch.arange(W, device=device, dtype=dt), indexing="ij")
    dcam = torch.stack([(xs - intr["cx"]) / intr["fx"], (ys - intr["cy"]) / intr["fy"], torch.ones_like(xs)], -1)
    d = dcam @ c2w[:3, :3].T
    o = c2w[:3, 3]
    inv = 1.0 / torch.where(d.abs() < 1e-12, torch.full_like(d, 1e-12), d)

    def slab(bmin, bmax):
        t0 = (torch.tensor(bmin, dtype=dt, device=device) - o) * inv
        t1 = (torch.tensor(bmax, dtype=dt, device=device) - o) * inv
        tn = torch.minimum(t0, t1).amax(-1)
        tf = torch.maximum(t0, t1).amin(-1)
        return tn, tf

    _, t_room = slab(ROOM_MIN, ROOM_MAX)          # camera is inside: exit distance
    t = t_room
    for bmin, bmax in BOXES:
        tn, tf = slab(bmin, bmax)
        hit = (tn < tf) & (tn > 1e-4)
        t = torch.where(hit & (tn < t), tn, t)
    p = o + d * t[..., None]
    col = _albedo(p.to(torch.float32), seed)
    rgba = torch.empty(H, W, 4, dtype=torch.uint8, device=device)
    rgba[..., :3] = (col * 255.0 + 0.5).to(torch.uint8)
    rgba[..., 3] = 255
    depth_mm = torch.round(t * 1000.0).clamp(0, 32000).to(torch.int16)   # z == t because dcam.z == 1
    return rgba, depth_mm


def sequence(n_frames, intr=None, seed=42, device="cpu"):
    """-> poses [n,4,4] float32, list of (rgba, depth_mm) tensors on `device`."""
    intr = intr or intrinsics()
    poses = trajectory(n_frames, seed)
    frames = [render_frame(poses[i], intr, seed, device) for i in range(n_frames)]
    return poses, frames


def c2w_to_colmajor(c2w):
    """row-major 4x4 -> the 16 floats of ORUtils::Matrix4 (m[col*4+row]); cf. reference src/tensor_math.cpp:5-25"""
    return np.ascontiguousarray(np.asarray(c2w, dtype=np.float32).T).reshape(16)
