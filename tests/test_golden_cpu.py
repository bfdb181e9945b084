"""CPU suite: the oracles against the committed golden vectors (tests/golden/, produced by running the reference itself --
see make_golden_itm.py / make_golden_gs.py there).

  * gs_ref_golden.npz  : outputs of the reference's gsplat CUDA kernels on a B200 -> oracle/gs_oracle.py (numpy) must reproduce the
    integer outputs bit for bit and the floating-point ones within the tolerances of tests/gs_checks.py;
  * itm_ref_golden.npz : outputs of the reference's InfiniTAM CPU engine -> when oracle/_ref is built here, the live reference must
    still reproduce the fixture bit for bit (guards the fixture and the build recipe against drift)."""
import os

import numpy as np
import pytest

from tests import gs_checks as gc
from tests.golden import make_golden_gs, make_golden_itm
from tests.helpers_gs import camera, random_splats, scene_images

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    path = os.path.join(GOLD, name)
    assert os.path.exists(path), "golden fixture %s is missing" % name
    return np.load(path)


@pytest.mark.parametrize("tag", sorted(make_golden_gs.CASES))
def test_numpy_oracle_reproduces_reference_kernel_golden(tag):
    from oracle import gs_oracle as go
    g = _load("gs_ref_golden.npz")
    c = make_golden_gs.CASES[tag]
    N, W, H = c["N"], c["W"], c["H"]
    p = random_splats(N, seed=c["seed"])
    c2w, K = camera(W, H, c["seed"])
    ref_depth, base, gt = scene_images(W, H, c["seed"])
    a = go.ges_iteration(p, c2w, K, W, H, ref_depth, base, gt)
    G = lambda k: g[tag + "_" + k]
    assert np.array_equal(a["proj"]["radii"], G("radii"))
    assert np.array_equal(a["tiles_per_gauss"], G("tiles_per_gauss"))
    assert np.array_equal(a["isect_ids"], G("isect_ids"))
    assert np.array_equal(a["flatten_ids"], G("flatten_ids"))
    assert np.array_equal(a["tile_offsets"], G("tile_offsets"))
    vis = G("radii") > 0
    gc.close_frac("means2d", a["proj"]["means2d"][vis], G("means2d")[vis], 2e-4, 2e-6)
    gc.close_frac("conics", a["proj"]["conics"][vis], G("conics")[vis], 2e-6, 5e-5)
    gc.close_frac("depths", a["proj"]["depths"][vis], G("depths")[vis], 2e-6, 2e-6)
    gc.close_frac("colors", a["colors"][vis], G("colors")[vis], 5e-6, 5e-5)
    assert abs(a["loss"] - float(G("loss"))) <= 1e-5 * max(1.0, abs(float(G("loss"))))
    for k in ("v_means2d", "v_conics", "v_opacities"):
        gc.close_scaled(k, a[k][vis], G(k)[vis], 2e-3)
    gc.close_scaled("v_colors", a["v_colors"][vis, :3], G("v_colors")[vis, :3], 2e-3)
    for k in ("means", "scales", "quats", "featuresDc", "featuresRest", "opacities"):
        gc.close_scaled("grad " + k, a["grads"][k].reshape(N, -1), G("grad_" + k).reshape(N, -1), 3e-3)
    if c["image"]:
        gc.close_frac("render", a["render"], G("render"), 2e-4, 2e-4, 2e-4)
        gc.close_frac("alphas", a["alphas"], G("alphas"), 2e-4, 2e-4, 2e-4)
        gc.close_frac("rgb", a["rgb"], G("rgb"), 2e-4, 2e-4, 2e-4)


def test_reference_cpu_engine_reproduces_itm_golden():
    from oracle import itm_ref
    if not itm_ref.available("exact"):
        pytest.skip("oracle/_ref/libitm_ref_exact.so not built here (needs /root/reference)")
    from gps_slam_b200 import synthetic as syn
    g = _load("itm_ref_golden.npz")
    n = int(g["n_frames"])
    intr = syn.intrinsics("replica", float(g["scale"]))
    poses, frames = syn.sequence(n, intr)
    ref = itm_ref.ItmRef(intr, tracker=0, threads=1, kind="exact")
    try:
        for i in range(n):
            ref.process_frame(frames[i][0].numpy(), frames[i][1].numpy(), syn.c2w_to_colmajor(poses[i]))
            assert np.array_equal(np.array(ref.visible_ids()), g["f%d_visible_ids" % i])
            assert [ref.last_free_block(), ref.last_free_excess()] == list(g["f%d_free_heads" % i])
            assert np.array_equal(make_golden_itm.table_digest(ref.hash_entries()), g["f%d_table_sha" % i])
            assert np.array_equal(make_golden_itm.voxel_digest(ref.voxels(), ref.last_free_block() + 1), g["f%d_voxel_sha" % i])
            assert np.array_equal(make_golden_itm.sha(ref.raycast()), g["f%d_raycast_sha" % i])
            assert np.array_equal(np.array(ref.raycast()[::4, ::4]).view(np.uint32), g["f%d_raycast_sub4" % i].view(np.uint32))
    finally:
        ref.close()


def test_itm_golden_is_sane():
    """fixture-only checks that run everywhere: block lists ascending and growing, poses rigid"""
    g = _load("itm_ref_golden.npz")
    n = int(g["n_frames"])
    prev = 0
    for i in range(n):
        ids = g["f%d_visible_ids" % i]
        assert len(ids) > 100 and np.all(np.diff(ids) > 0)
        heads = g["f%d_free_heads" % i]
        assert prev == 0 or heads[0] <= prev
        prev = heads[0]
        pts = g["f%d_raycast_sub4" % i]
        assert (pts[..., 3] > 0).mean() > 0.5
    for name in ("extended", "icp"):
        M = g["track_%s_M" % name].reshape(n, 4, 4).transpose(0, 2, 1)
        for i in range(n):
            R = M[i, :3, :3].astype(np.float64)
            assert np.abs(R @ R.T - np.eye(3)).max() < 1e-5
