"""Checkpoint / wire formats against the reference itself:

  * the Scene/ directory written from our engine is BYTE-IDENTICAL to what the reference CPU engine's SaveToFile writes after the same
    frames, the reference loads ours and we load the reference's, and fusion continues bit-exact on both sides afterwards;
  * the Gaussian point_cloud.ply round-trips the six parameter tensors bit for bit in the reference's property layout."""
import filecmp
import os

import numpy as np
import pytest

from gps_slam_b200 import checkpoint as ck
from gps_slam_b200 import synthetic as syn
from tests.helpers_gs import random_splats
from tests.test_tsdf_parity_gpu import compare_state

pytestmark = pytest.mark.gpu
FILES = ("hash.dat", "excess.dat", "last.txt", "voxel.dat", "alloc.dat", "vba.txt")


def test_scene_directory_is_wire_compatible_with_the_reference(engine_lib, tmp_path):
    from gps_slam_b200.engine import TsdfEngine
    from oracle.itm_ref import ItmRef, available
    if not available("exact"):
        pytest.skip("oracle/_ref/libitm_ref_exact.so not built")
    intr = syn.intrinsics("replica", 0.25)
    poses, frames = syn.sequence(6, intr)
    ref, eng = ItmRef(intr, tracker=0, threads=1, kind="exact"), TsdfEngine(intr, tracker=0)
    ref2, eng2 = ItmRef(intr, tracker=0, threads=1, kind="exact"), TsdfEngine(intr, tracker=0)
    try:
        for i in range(3):
            c2w = syn.c2w_to_colmajor(poses[i])
            ref.process_frame(frames[i][0].numpy(), frames[i][1].numpy(), c2w)
            eng.ProcessFrame(frames[i][0].numpy(), frames[i][1].numpy(), c2w)
        d_ref, d_ours = str(tmp_path / "ref") + "/", str(tmp_path / "ours") + "/"
        ref.save(d_ref)
        ck.save_scene(d_ours, eng)
        for name in FILES:
            a, b = os.path.join(d_ref, "Scene", name), os.path.join(d_ours, "Scene", name)
            assert os.path.getsize(a) == os.path.getsize(b), name
            assert filecmp.cmp(a, b, shallow=False), "%s differs from the reference's file" % name
        # cross-load: the reference continues from OUR files, we continue from the REFERENCE's
        ref2.load(d_ours)
        ck.load_scene(d_ref, eng2)
        for i in range(3, 6):
            c2w = syn.c2w_to_colmajor(poses[i])
            for x in (ref, ref2):
                x.process_frame(frames[i][0].numpy(), frames[i][1].numpy(), c2w)
            for x in (eng, eng2):
                x.ProcessFrame(frames[i][0].numpy(), frames[i][1].numpy(), c2w)
        # a loaded engine rebuilds visibility from the new frames only, exactly like the reference's loaded engine
        compare_state(eng2, ref2, "after load")
        compare_state(eng, ref, "uninterrupted")
    finally:
        for x in (ref, ref2, eng, eng2):
            x.close()


def test_ply_round_trip_and_layout(engine_lib, tmp_path):
    from gps_slam_b200.engine import GaussianEngine
    p = random_splats(777, seed=5)
    eng = GaussianEngine(64, 64, capacity=1024)
    try:
        eng.set_params(p)
        path = str(tmp_path / "point_cloud.ply")
        ck.save_ply(path, eng.get_params())
        q = ck.load_ply(path)
        for k in p:
            assert np.array_equal(np.asarray(p[k], np.float32).reshape(777, -1), q[k].reshape(777, -1)), k
        raw = open(path, "rb").read()
        head = raw[: raw.index(b"end_header\n")].decode().split("\n")
        assert head[2] == "element vertex 777" and head[3:9] == ["property float %s" % c for c in ("x", "y", "z", "nx", "ny", "nz")]
        assert len(raw) - raw.index(b"end_header\n") - 11 == 777 * 62 * 4
        # f_rest is channel-major (featuresRest.transpose(1, 2)): f_rest_0..14 are the red coefficients
        row0 = np.frombuffer(raw, "<f4", 62, raw.index(b"end_header\n") + 11)
        assert np.array_equal(row0[9:24], np.asarray(p["featuresRest"], np.float32)[0, :, 0])
        eng.set_params(q)
        assert eng.getGaussianNum() == 777
    finally:
        eng.close()
