"""The InfiniTAM-facing C++ facade (gps_slam_b200/cxx/InfiniTAM: ITMLib::ITMBasicEngine, ORUtils::Image / Matrix4 / SE3Pose,
ITMLibSettings, ITMRGBDCalib ... with the reference's names) driven by a C++ program that does what the reference's host code does
(tests/cxx/itm_facade_driver.cpp: createTsdfEngine, the CLIEngine frame loop, the pose read-out, runRaycastByCam, SaveToFile /
LoadFromFile), against

  * the ctypes route into the same library (bit-exact: both are thin hosts over one C ABI), and
  * the reference's own CPU engine on the same frames (oracle/_ref/libitm_ref_exact.so, when built): free-view vertex map bit-exact,
    Scene/ directory byte-identical.

Two builds of the driver run: ours (always), and the one whose frame loop is the REFERENCE's slam/TsdfFusion/CLIEngine.cpp compiled
unchanged against the facade headers (oracle/_ref/itm_facade_driver_refcli, when built)."""
import filecmp
import os
import subprocess

import numpy as np
import pytest

from gps_slam_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFCLI = os.path.join(ROOT, "oracle", "_ref", "itm_facade_driver_refcli")
FILES = ("hash.dat", "excess.dat", "last.txt", "voxel.dat", "alloc.dat", "vba.txt")
N_FRAMES = 5


def write_input(path, intr, poses, frames, free_poses, tracker):
    with open(path, "wb") as f:
        f.write(np.array([len(frames), intr["width"], intr["height"], tracker, len(free_poses)], np.int32).tobytes())
        f.write(np.array([intr["fx"], intr["fy"], intr["cx"], intr["cy"], 0.005, 0.02, 0.2, 10.0], np.float32).tobytes())
        for p in poses:
            f.write(syn.c2w_to_colmajor(p).tobytes())
        for p in free_poses:
            f.write(syn.c2w_to_colmajor(p).tobytes())
        for rgba, depth in frames:
            f.write(np.ascontiguousarray(rgba.numpy()).tobytes())
            f.write(np.ascontiguousarray(depth.numpy()).tobytes())


def read_output(path, n, n_free, w, h, with_scene):
    raw = np.fromfile(path, np.uint8)
    pos = 0

    def take(dtype, count, shape):
        nonlocal pos
        nbytes = np.dtype(dtype).itemsize * count
        a = raw[pos:pos + nbytes].view(dtype).reshape(shape)
        pos += nbytes
        return a
    out = dict(poses=take(np.float32, n * 16, (n, 16)), free=[], cams=[], again=[])
    for _ in range(n_free):
        out["free"].append((take(np.uint8, w * h * 4, (h, w, 4)), take(np.float32, w * h * 4, (h, w, 4))))
    for _ in range(2):
        out["cams"].append(take(np.float32, w * h * 4, (h, w, 4)))
    out["voxel"] = float(take(np.float32, 1, (1,))[0])
    if with_scene:
        for _ in range(1):
            out["again"].append((take(np.uint8, w * h * 4, (h, w, 4)), take(np.float32, w * h * 4, (h, w, 4))))
    assert pos == raw.size
    return out


def same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))


@pytest.fixture(scope="module")
def drivers(engine_lib):
    from gps_slam_b200 import build
    d = [("facade", build.build_itm_driver())]       # never skipped: the facade must build
    if os.path.exists(REFCLI):
        d.append(("reference CLIEngine over the facade", REFCLI))
    return d


@pytest.mark.parametrize("tracker", [0, 1, 2])
def test_cxx_facade_matches_ctypes_route(drivers, tmp_path, tracker):
    from gps_slam_b200.engine import TsdfEngine
    intr = syn.intrinsics("replica", 0.25)
    w, h = intr["width"], intr["height"]
    poses, frames = syn.sequence(N_FRAMES + 2, intr)
    free_poses = [poses[N_FRAMES], poses[N_FRAMES + 1]]
    poses, frames = poses[:N_FRAMES], frames[:N_FRAMES]
    inp = str(tmp_path / "frames.bin")
    write_input(inp, intr, poses, frames, free_poses, tracker)

    # ---- ctypes route
    eng = TsdfEngine(intr, tracker=tracker)
    try:
        exp_poses, exp_free, exp_cams = [], [], []
        for i in range(N_FRAMES):
            eng.ProcessFrame(frames[i][0].numpy(), frames[i][1].numpy(), syn.c2w_to_colmajor(poses[i]) if tracker == 0 else None)
            exp_poses.append(eng.pose())
        for p in free_poses:
            eng.runRaycast(syn.c2w_to_colmajor(p), intr)
            exp_free.append((eng.free_image().copy(), eng.raycast(live=False).copy()))
        for k in (0, N_FRAMES - 1):
            eng.runRaycast(exp_poses[k][1], intr)
            exp_cams.append(eng.raycast(live=False).copy())
        voxel = eng.getVoxelSize()
    finally:
        eng.close()
    assert np.abs(exp_free[0][1]).max() > 0

    for name, exe in drivers:
        outp = str(tmp_path / "out.bin")
        r = subprocess.run([exe, inp, outp, str(tmp_path / ("scene_" + os.path.basename(exe)))], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, "%s: rc %d\n%s\n%s" % (name, r.returncode, r.stdout[-2000:], r.stderr[-2000:])
        got = read_output(outp, N_FRAMES, len(free_poses), w, h, True)
        for i in range(N_FRAMES):
            assert same_bits(got["poses"][i], exp_poses[i][1]), "%s: GetInvM() of frame %d" % (name, i)
            if tracker == 0:
                # SetInvM re-derives the pose parameters (SE3Pose::SetInvM -> SetParamsFromModelView), so not bit-equal to the input
                np.testing.assert_allclose(got["poses"][i], syn.c2w_to_colmajor(poses[i]), atol=2e-5)
        for i in range(len(free_poses)):
            assert same_bits(got["free"][i][0], exp_free[i][0]), "%s: free image %d" % (name, i)
            assert same_bits(got["free"][i][1], exp_free[i][1]), "%s: free vertex %d" % (name, i)
        for k in range(2):
            assert same_bits(got["cams"][k], exp_cams[k]), "%s: raycast at camPoses[%d]" % (name, k)
        assert got["voxel"] == voxel == pytest.approx(0.005)
        # SaveToFile -> LoadFromFile into a second engine -> one more frame: the same through gps_slam_b200.checkpoint + ctypes
        from gps_slam_b200 import checkpoint as ck
        eng2 = TsdfEngine(intr, tracker=0)
        try:
            ck.load_scene(str(tmp_path / ("scene_" + os.path.basename(exe))), eng2)
            last = syn.c2w_to_colmajor(poses[N_FRAMES - 1])
            eng2.ProcessFrame(frames[N_FRAMES - 1][0].numpy(), frames[N_FRAMES - 1][1].numpy(), last)
            eng2.runRaycast(last, intr)
            assert same_bits(got["again"][0][0], eng2.free_image()), "%s: image after LoadFromFile" % name
            assert same_bits(got["again"][0][1], eng2.raycast(live=False)), "%s: vertex map after LoadFromFile" % name
        finally:
            eng2.close()
        assert np.abs(got["again"][0][1]).max() > 0


def test_cxx_facade_matches_reference_cpu_engine(drivers, tmp_path):
    """use_gt_pose mode against the reference's InfiniTAM CPU engine on the same frames: free-view raycast bit-exact, and the Scene/
    directory the facade's SaveToFile writes is byte-identical to the reference's"""
    from oracle.itm_ref import ItmRef, available
    if not available("exact"):
        pytest.skip("oracle/_ref/libitm_ref_exact.so not built")
    intr = syn.intrinsics("replica", 0.25)
    w, h = intr["width"], intr["height"]
    poses, frames = syn.sequence(N_FRAMES + 2, intr)
    free_poses = [poses[N_FRAMES], poses[N_FRAMES + 1]]
    poses, frames = poses[:N_FRAMES], frames[:N_FRAMES]
    inp = str(tmp_path / "frames.bin")
    write_input(inp, intr, poses, frames, free_poses, 0)
    ref = ItmRef(intr, tracker=0, threads=1, kind="exact")
    try:
        for i in range(N_FRAMES):
            ref.process_frame(frames[i][0].numpy(), frames[i][1].numpy(), syn.c2w_to_colmajor(poses[i]))
        ref_dir = str(tmp_path / "ref") + "/"
        ref.save(ref_dir)
        exp_free = []
        for p in free_poses:
            ref.run_raycast(syn.c2w_to_colmajor(p), intr)
            exp_free.append(ref.raycast(live=False).copy())
    finally:
        ref.close()
    for name, exe in drivers:
        outp, scene = str(tmp_path / "out.bin"), str(tmp_path / ("scene_" + os.path.basename(exe)))
        r = subprocess.run([exe, inp, outp, scene], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, "%s: rc %d\n%s" % (name, r.returncode, r.stderr[-2000:])
        got = read_output(outp, N_FRAMES, len(free_poses), w, h, True)
        for i in range(len(free_poses)):
            assert same_bits(got["free"][i][1], exp_free[i]), "%s: free vertex map %d differs from the reference CPU engine" % (name, i)
        for f in FILES:
            assert filecmp.cmp(os.path.join(ref_dir, "Scene", f), os.path.join(scene, "Scene", f), shallow=False), "%s: %s" % (name, f)
