"""Marching-cubes export (ITMBasicEngine::SaveSceneToMesh) against the REFERENCE's CPU mesher
(oracle/_ref/libitm_ref_exact.so: ITMMeshingEngine_CPU::MeshScene for the positions, the reference's buildVertList for the per-vertex colours
that only its CUDA mesher stores), BIT-EXACT and in the same triangle order; the overflow rule (at most max - 1 triangles are kept); the PLY
file in ITMMesh::WritePLY's layout."""
import numpy as np
import pytest

from gps_slam_b200 import synthetic as syn
from tests.test_tsdf_parity_gpu import assert_same

pytestmark = pytest.mark.gpu


def _scene(scale, n_frames):
    from gps_slam_b200.engine import TsdfEngine
    from oracle.itm_ref import ItmRef
    intr = syn.intrinsics("replica", scale)
    poses, frames = syn.sequence(n_frames, intr)
    # 2 cm voxels (8 cm truncation) keep the mesh of the room at a few hundred thousand triangles
    ref = ItmRef(intr, voxel=0.02, mu=0.08, tracker=0, threads=1, kind="exact")
    eng = TsdfEngine(intr, voxel_size=0.02, mu=0.08, tracker=0)
    for i in range(n_frames):
        rgba, d = frames[i][0].numpy(), frames[i][1].numpy()
        c2w = syn.c2w_to_colmajor(poses[i])
        ref.process_frame(rgba, d, c2w)
        eng.ProcessFrame(rgba, d, c2w)
    return eng, ref


def test_mesh_matches_reference_cpu_mesher(engine_lib, tmp_path):
    eng, ref = _scene(0.25, 12)
    try:
        want = ref.mesh(max_tri=3_000_000)
        got = eng.mesh().cpu().numpy()
        assert len(want) > 50_000, len(want)
        assert_same("triangle positions", got[:, :9], want[:, :9])
        assert_same("vertex colours", got[:, 9:], want[:, 9:])
        # geometry sanity: every vertex lies on a voxel-cube edge of the room-scale scene, colours are in [0, 1]
        assert np.isfinite(got).all() and got[:, 9:].min() >= 0 and got[:, 9:].max() <= 1
        assert got[:, :9].min() > -1.0 and got[:, :9].max() < 8.0
        # overflow rule of the CPU mesher: with room for `cap` triangles only the first cap - 1 are kept
        cap = 1000
        few = eng.mesh(max_tri=cap).cpu().numpy()
        assert len(few) == cap - 1
        assert_same("truncated mesh", few, want[:cap - 1])
        few_ref = ref.mesh(max_tri=cap)
        assert len(few_ref) == cap - 1
        # the PLY file
        from gps_slam_b200 import checkpoint
        path = str(tmp_path / "mesh.ply")
        n = checkpoint.save_mesh_ply(path, eng)
        lines = open(path).read().split("\n")
        assert n == len(want) and lines[0] == "ply" and lines[2] == "element vertex %d" % (3 * n) and lines[9] == "element face %d" % n
        assert lines[12] == "%f %f %f %d %d %d" % (*want[0, :3], *(want[0, 9:12] * np.float32(255)).astype(np.int32))
        assert lines[12 + 3 * n] == "3 0 1 2" and len(lines) == 12 + 4 * n + 1
    finally:
        eng.close()
        ref.close()


def test_mesh_of_an_empty_scene(engine_lib):
    from gps_slam_b200.engine import TsdfEngine
    eng = TsdfEngine(syn.intrinsics("replica", 0.25), tracker=0)
    try:
        assert len(eng.mesh()) == 0
    finally:
        eng.close()
