"""Error behaviour and capacity limits of the Gaussian engine's C ABI: every failure is an error code + gsb_last_error() text
(the Python mirror raises EngineError), never a crash, a silent truncation without a flag, or a fallback."""
import numpy as np
import pytest
import torch

from tests import gs_checks as gc
from tests.helpers_gs import camera, random_splats, scene_images

pytestmark = pytest.mark.gpu


def test_capacity_errors(engine_lib):
    from gps_slam_b200.engine import EngineError, GaussianEngine
    eng = GaussianEngine(64, 64, capacity=256)       # rounded up to a multiple of 128
    try:
        eng.set_params(random_splats(256, seed=1))
        with pytest.raises(EngineError, match="capacity"):
            eng.add(random_splats(1, seed=2))
        with pytest.raises(EngineError, match="capacity"):
            eng.set_params(random_splats(257, seed=3))
        assert eng.getGaussianNum() == 256             # the failed calls changed nothing
    finally:
        eng.close()


def test_intersection_and_item_capacity_overflow_is_flagged(engine_lib):
    """more tile intersections / backward work items than the configured capacity: the overflow bits are raised, the extra
    intersections are dropped (clamped offsets), nothing is written out of bounds and the step still finishes"""
    from gps_slam_b200.engine import GaussianEngine
    W, H, N = 320, 192, 2000
    p = random_splats(N, seed=4, scale_lo=0.05, scale_hi=0.3)     # large splats: many tiles each
    c2w, K = camera(W, H, 4)
    intr = gc.intr_of(K, W, H)
    ref_depth, base, gt = scene_images(W, H, 4)
    dev = torch.device("cuda", 0)
    rd, bs, g = [torch.from_numpy(a).to(dev) for a in (ref_depth, base, gt)]
    for kw, bit in ((dict(isect_capacity=4096), 1), (dict(item_capacity=64), 2)):
        eng = GaussianEngine(W, H, capacity=N, **kw)
        try:
            eng.set_params(p)
            eng.initOptimizers()
            eng.train_step(c2w, intr, rd, bs, g)
            eng.sync()
            cnt = eng.counters()
            assert cnt[2] & bit, "overflow bit %d not raised: %s" % (bit, cnt)
            if bit == 1:
                off, ids = eng.tile_bins()
                assert off[-1] == 4096 and len(ids) == 4096 and np.all(off <= 4096)
            assert np.isfinite(eng.loss())
            after = eng.get_params()
            assert all(np.isfinite(after[k]).all() for k in after)
        finally:
            eng.close()


def test_staged_entry_points_reject_bad_arguments(engine_lib):
    from gps_slam_b200.engine import EngineError
    from gps_slam_b200.gsplat_ops import GsplatOps
    ops = GsplatOps(64, 64, capacity=128)
    dev = ops.dev
    try:
        n = 129
        z = lambda *s: torch.zeros(*s, device=dev)
        with pytest.raises(EngineError, match="capacity"):
            ops.fully_fused_projection_fwd(z(n, 3), z(n, 4), z(n, 3), torch.eye(4, device=dev)[None], torch.eye(3, device=dev)[None])
        with pytest.raises(EngineError, match="degree"):
            ops.compute_sh_fwd(2, z(1, 8, 3), z(1, 8, 16, 3))
        with pytest.raises(EngineError, match="step"):
            ops.adam_step(z(8), z(8), z(8), z(8), 1e-3, 0)
    finally:
        ops.close()


def test_non_multiple_of_tile_image_and_border_splats(engine_lib):
    """image sides that are not multiples of 16 and splats hanging over every border (saturating tile casts)"""
    N, W, H = 1200, 333, 211
    it = gc.compare_iteration(N, W, H, seed=17, spread=1.6, scale_lo=0.01, scale_hi=0.08)
    m = it["proj"]["means2d"][it["proj"]["radii"] > 0]
    assert (m[:, 0] < 0).any() and (m[:, 0] > W).any() and (m[:, 1] < 0).any() and (m[:, 1] > H).any()
