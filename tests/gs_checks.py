"""GPU-vs-oracle comparison of the gsplat-GES path, shared by tests/test_gs_parity_gpu.py and __graft_entry__.smoke().

Tolerances (fp32; the oracle is numpy without FMA contraction, the rasteriser uses FMA and __expf like the reference):
  * radii, tile offsets, flatten ids: bit-exact;
  * projection outputs (means2d, conics, depths), colours, opacities: rtol 2e-5 / atol 2e-6 (they only differ through expf);
  * rendered rgb / depth / alpha, dL/d(render): |d| <= 2e-4 + 2e-4 |ref| for all but a 2e-4 fraction of pixels -- a splat
    whose alpha sits within one ulp of the 1/255 cut may legitimately fall on either side;
  * per-splat raster gradients and parameter gradients: max |d| <= 2e-3 of the largest magnitude of that tensor;
  * parameters after the Adam step: compared with the oracle's Adam applied to the GPU's own gradients (rtol 1e-5), so a
    sign flip of a near-zero gradient (which moves a parameter by a full +-lr under Adam) cannot masquerade as an error.
"""
import numpy as np
import torch

from oracle import gs_oracle as go
from tests.helpers_gs import camera, random_splats, scene_images

LR = dict(means=1.6e-4, scales=5e-3, quats=1e-3, featuresDc=2.5e-3, featuresRest=5e-4, opacities=5e-2)


def close_frac(name, a, b, atol, rtol, max_bad_frac=0.0):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, "%s: shape %s vs %s" % (name, a.shape, b.shape)
    bad = np.abs(a - b) > atol + rtol * np.abs(b)
    frac = bad.mean() if bad.size else 0.0
    assert frac <= max_bad_frac, "%s: %.3g of elements differ (max |d| = %.3g at ref %.3g)" % (
        name, frac, np.abs(a - b).max(), np.abs(b).max())


def close_scaled(name, a, b, tol, max_bad_frac=0.0):
    """max |d| <= tol * max |ref|; with max_bad_frac > 0 that share of the elements may exceed it, by at most 10x (a pixel whose
    alpha sits on the 1/255 cut or the 0.999 clamp enters one side's gradient sum and not the other's)"""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, "%s: shape %s vs %s" % (name, a.shape, b.shape)
    scale = np.abs(b).max() + 1e-20
    d = np.abs(a - b) / scale
    err = d.max() if d.size else 0.0
    if max_bad_frac > 0:
        assert (d > tol).mean() <= max_bad_frac and err <= 10 * tol, "%s: %.3g of elements beyond %.3g, max %.3g" % (name, (d > tol).mean(), tol, err)
    else:
        assert err <= tol, "%s: max |d| / max |ref| = %.3g (scale %.3g)" % (name, err, scale)


def intr_of(K, W, H):
    return dict(width=W, height=H, fx=float(K[0, 0]), fy=float(K[1, 1]), cx=float(K[0, 2]), cy=float(K[1, 2]))


def compare_iteration(N, W, H, seed, verbose=False, checker=None, **splat_kw):
    """checker: function with gs_oracle.ges_iteration's signature producing the expected values -- the numpy oracle by
    default, oracle.gsplat_ref.ges_iteration (the reference's own CUDA kernels) in tests/test_gs_reference_gpu.py"""
    from gps_slam_b200.engine import GaussianEngine
    p = random_splats(N, seed=seed, **splat_kw)
    c2w, K = camera(W, H, seed)
    ref_depth, base, gt = scene_images(W, H, seed)
    it = (checker or go.ges_iteration)(p, c2w, K, W, H, ref_depth, base, gt)
    intr = intr_of(K, W, H)
    dev = torch.device("cuda", 0)
    rd_d, base_d, gt_d = [torch.from_numpy(a).to(dev).contiguous() for a in (ref_depth, base, gt)]
    eng = GaussianEngine(W, H, capacity=max(N, 1))
    try:
        eng.set_params(p)
        assert eng.getGaussianNum() == N
        eng.enable_grad_dump(True)
        eng.initOptimizers()
        # ---- gesForward (no grad)
        rgb = torch.empty((H, W, 3), device=dev)
        depth = torch.empty((H, W), device=dev)
        alpha = torch.empty((H, W), device=dev)
        eng.forward(c2w, intr, rd_d, base_d, rgb, depth, alpha)
        eng.sync()
        rec = eng.splat_records(N)
        proj = it["proj"]
        odd = np.nonzero(rec["radii"] != proj["radii"])[0]
        if len(odd) and checker is None:
            raise AssertionError("radii differ at %s" % odd[:5])
        # against the reference's kernels (FMA-contracted, rsqrtf) radius = ceil(3 sqrt(lambda_max)) may fall on the other side of an
        # integer for a splat whose extent is within an ulp of it: allowed for <= 1e-4 of the splats, by exactly one pixel, and
        # those splats are then left out of the bit-exact bin comparison (tests/test_gs_reference_gpu.py pins the binning stage itself
        # by feeding the reference's kernel the engine's own projection)
        assert len(odd) <= max(1, int(1e-4 * N)), "radii differ for %d splats" % len(odd)
        assert np.all(np.abs(rec["radii"][odd] - proj["radii"][odd]) == 1), "radius off by more than a pixel"
        vis = proj["radii"] > 0     # culled Gaussians: the reference leaves torch::empty garbage in these arrays
        close_frac("means2d", rec["means2d"][vis], proj["means2d"][vis], 2e-4, 2e-6)
        close_frac("conics", rec["conics"][vis], proj["conics"][vis], 2e-6, 5e-5)
        close_frac("depths", rec["depths"][vis], proj["depths"][vis], 2e-6, 2e-6)
        close_frac("colors", rec["colors"][vis], it["colors"][vis], 2e-6, 2e-5)
        close_frac("opacities", rec["opacities"][vis], go.real_opacities(p["opacities"]).reshape(-1)[vis], 2e-6, 2e-5)
        off, ids = eng.tile_bins()
        if len(odd) == 0:
            assert off[-1] == len(it["isect_ids"]), "n_isects %d vs %d" % (off[-1], len(it["isect_ids"]))
            assert np.array_equal(off[:-1], it["tile_offsets"]), "tile offsets differ"
            assert np.array_equal(ids, it["flatten_ids"]), "flatten ids differ"
        else:
            tile_a = np.repeat(np.arange(len(off) - 1), np.diff(off))
            off_b = np.append(it["tile_offsets"], len(it["flatten_ids"]))
            tile_b = np.repeat(np.arange(len(off_b) - 1), np.diff(off_b))
            ka, kb = ~np.isin(ids, odd), ~np.isin(it["flatten_ids"], odd)
            assert np.array_equal(ids[ka], it["flatten_ids"][kb]) and np.array_equal(tile_a[ka], tile_b[kb]), "tile bins differ"
        close_frac("rgb", rgb.cpu().numpy(), it["rgb"], 2e-4, 2e-4, 2e-4)
        close_frac("alpha", alpha.cpu().numpy(), it["alphas"], 2e-4, 2e-4, 2e-4)
        ok = np.isfinite(it["depth"])
        close_frac("depth", depth.cpu().numpy()[ok], it["depth"][ok], 2e-4, 2e-4, 2e-4)
        # ---- one optimiser iteration
        eng.train_step(c2w, intr, rd_d, base_d, gt_d)
        loss = eng.loss()
        assert abs(loss - it["loss"]) <= 1e-5 * max(1.0, abs(it["loss"])), "loss %r vs %r" % (loss, it["loss"])
        vo = eng.v_out()
        close_frac("v_render", vo[..., :3], it["v_render"][..., :3], 1e-9, 2e-4, 2e-4)
        close_frac("v_alpha", vo[..., 3], it["v_alphas"], 1e-9, 2e-4, 2e-4)
        sg = eng.splat_grads(N)
        same = np.ones(N, bool)
        same[odd] = False            # a splat whose radius differs by a pixel has a different backward box: not comparable
        bad = 0.0 if checker is None else 1e-4   # vs the reference's kernels: see close_scaled
        for k in ("v_means2d", "v_conics", "v_opacities"):
            close_scaled(k, sg[k][vis & same], it[k][vis & same], 2e-3, bad)
        close_scaled("v_colors", sg["v_colors"][vis & same], it["v_colors"][vis & same, :3], 2e-3, bad)
        pg = eng.param_grads(N)
        for k in ("means", "scales", "quats", "featuresDc", "featuresRest", "opacities"):
            close_scaled("grad " + k, pg[k].reshape(N, -1)[same], it["grads"][k].reshape(N, -1)[same], 3e-3, bad)
        after = eng.get_params()
        for k in ("means", "scales", "quats", "featuresDc", "featuresRest", "opacities"):
            exp = np.array(p[k], np.float32, copy=True).reshape(N, -1)
            g = pg[k].reshape(N, -1)
            m, v = np.zeros_like(exp), np.zeros_like(exp)
            go.adam_step(exp, g, m, v, 1, LR[k])
            # Gaussians outside the frustum never get optimiser state and must be untouched bit for bit
            got = after[k].reshape(N, -1)
            assert np.array_equal(got[~vis], np.asarray(p[k], np.float32).reshape(N, -1)[~vis]), "culled Gaussians moved: " + k
            close_frac("adam " + k, got[vis], exp[vis], 1e-7, 1e-5)
        cnt = eng.counters()
        assert cnt[2] == 0, "capacity overflow flags %d" % cnt[2]
        if verbose:
            print("smoke gs: N=%d visible=%d isects=%d loss=%.6f -- bins bit-exact, render/grad/Adam within tolerance of the oracle" % (
                N, int(vis.sum()), int(off[-1]), loss))
    finally:
        eng.close()
    return it


def smoke_check(verbose=False):
    compare_iteration(1500, 320, 192, seed=7, verbose=verbose)
