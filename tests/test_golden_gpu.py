"""GPU suite against the committed golden vectors (tests/golden/: outputs of the reference itself, see the make_golden_*.py
scripts there) -- needs neither /root/reference nor oracle/_ref at run time.

  * TSDF / ICP: the engine must reproduce the reference CPU engine's state digests BIT-EXACT (visible list, free-list heads, hash
    table, voxel blocks, raycast, ICP maps, free-view render) and its tracked poses within the tolerances of test_icp_parity_gpu;
  * gsplat GES: the engine against the reference kernels' recorded outputs (integers bit-exact, floats within gs_checks' bars)."""
import os

import numpy as np
import pytest
import torch

from gps_slam_b200 import synthetic as syn
from tests import gs_checks as gc
from tests.golden import make_golden_gs, make_golden_itm

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_tsdf_engine_reproduces_reference_golden(engine_lib):
    from gps_slam_b200.engine import TsdfEngine
    g = np.load(os.path.join(GOLD, "itm_ref_golden.npz"))
    n = int(g["n_frames"])
    intr = syn.intrinsics("replica", float(g["scale"]))
    poses, frames = syn.sequence(n, intr)
    eng = TsdfEngine(intr, tracker=0)
    try:
        for i in range(n):
            eng.ProcessFrame(frames[i][0].numpy(), frames[i][1].numpy(), syn.c2w_to_colmajor(poses[i]))
            assert np.array_equal(eng.visible_ids(), g["f%d_visible_ids" % i]), "frame %d visible list" % i
            assert [eng.counter(0), eng.counter(1)] == list(g["f%d_free_heads" % i]), "frame %d free-list heads" % i
            assert np.array_equal(make_golden_itm.table_digest(eng.hash_entries()), g["f%d_table_sha" % i]), "frame %d hash table" % i
            assert np.array_equal(make_golden_itm.voxel_digest(eng.voxels(), eng.counter(0) + 1), g["f%d_voxel_sha" % i]), "frame %d voxels" % i
            assert np.array_equal(eng.raycast()[::4, ::4].view(np.uint32), g["f%d_raycast_sub4" % i].view(np.uint32)), "frame %d raycast" % i
            assert np.array_equal(make_golden_itm.sha(eng.raycast()), g["f%d_raycast_sha" % i])
            assert np.array_equal(make_golden_itm.sha(eng.points_map()), g["f%d_points_sha" % i])
            assert np.array_equal(make_golden_itm.sha(eng.normals_map()), g["f%d_normals_sha" % i])
        c2w_f, intr_f = make_golden_itm.free_view(syn.trajectory(n + 40), n - 1, intr)
        eng.runRaycast(c2w_f, intr_f)
        assert np.array_equal(make_golden_itm.sha(eng.raycast(live=False)), g["free_vertex_sha"])
        assert np.array_equal(eng.free_image()[::4, ::4], g["free_image_sub4"])
        assert np.array_equal(make_golden_itm.sha(eng.free_image()), g["free_image_sha"])
    finally:
        eng.close()


@pytest.mark.parametrize("tracker,name,tol", [(1, "extended", 2e-4), (2, "icp", 3e-3)])
def test_tracker_reproduces_reference_golden(engine_lib, tracker, name, tol):
    from gps_slam_b200.engine import TsdfEngine
    g = np.load(os.path.join(GOLD, "itm_ref_golden.npz"))
    n = int(g["n_frames"])
    intr = syn.intrinsics("replica", float(g["scale"]))
    poses, frames = syn.sequence(n, intr)
    eng = TsdfEngine(intr, tracker=tracker)
    try:
        eng.set_pose(syn.c2w_to_colmajor(poses[0]))
        for i in range(n):
            eng.ProcessFrame(frames[i][0].numpy(), frames[i][1].numpy(), None)
            M = eng.pose()[0]
            assert np.abs(M - g["track_%s_M" % name][i]).max() < tol, "frame %d: %g" % (i, np.abs(M - g["track_%s_M" % name][i]).max())
    finally:
        eng.close()


@pytest.mark.parametrize("tag", sorted(make_golden_gs.CASES))
def test_gs_engine_reproduces_reference_kernel_golden(engine_lib, tag):
    g = np.load(os.path.join(GOLD, "gs_ref_golden.npz"))
    c = make_golden_gs.CASES[tag]
    N, W, H = c["N"], c["W"], c["H"]

    def from_golden(p, c2w, K, W_, H_, ref_depth, base, gt):
        """the recorded reference outputs in the shape gs_checks.compare_iteration expects; images it has no record of are taken
        from the numpy oracle (itself pinned to the same golden by tests/test_golden_cpu.py)"""
        from oracle import gs_oracle as go
        it = go.ges_iteration(p, c2w, K, W_, H_, ref_depth, base, gt)
        G = lambda k: g[tag + "_" + k]
        it["proj"] = dict(radii=G("radii"), means2d=G("means2d"), depths=G("depths"), conics=G("conics"))
        it["colors"] = G("colors")
        for k in ("isect_ids", "flatten_ids", "tile_offsets", "v_means2d", "v_conics", "v_colors", "v_opacities"):
            it[k] = G(k)
        it["loss"] = float(G("loss"))
        it["grads"] = {k: G("grad_" + k) for k in ("means", "scales", "quats", "featuresDc", "featuresRest", "opacities")}
        if c["image"]:
            it["rgb"], it["alphas"] = G("rgb"), G("alphas")
        return it

    gc.compare_iteration(N, W, H, c["seed"], checker=from_golden)
