"""The C++ host layer (gps_slam_b200/cxx/gsplat/gsplat_wapper.hpp: FullyFusedProjection, SphericalHarmonicsNew, isectTiles[NoDepth],
isectOffsetEncode[NoDepth], RasterizeToPixels, RasterizeToPixelsGes_NewParallel, FusedSSIMMap, simpleKNN -- same names and argument
lists as the reference's gsplat/gsplat_wapper.hpp, over the C ABI) driven through libtorch autograd exactly like
RawGaussianModel::gesForward / computeLoss drive the reference's (src/raw_gs_model.cpp:188-417), and compared with

  * the numpy oracle (always), and
  * the reference's own wrappers + kernels compiled for sm_100a (oracle/_ref/libgsplat_ref.so, when built)

by running ONE script (oracle/gsplat_ref.py: RefGaussians) against both torch.ops namespaces -- gsplat_b200 and gsplat_ref register
the same op schemas.  Bars: integer outputs bit-exact, floating point as in tests/gs_checks.py."""
import numpy as np
import pytest
import torch

from tests import gs_checks as gc
from tests.helpers_gs import camera, random_splats, scene_images

pytestmark = pytest.mark.gpu
DEV = "cuda"
KEYS = ("means", "scales", "quats", "featuresDc", "featuresRest", "opacities")


@pytest.fixture(scope="module")
def b200(engine_lib):
    """torch.ops.gsplat_b200 -- never skipped: a missing or unloadable shim library is a failure"""
    from gps_slam_b200 import build
    torch.ops.load_library(build.build_torch_shim())
    return torch.ops.gsplat_b200


@pytest.fixture(scope="module")
def gsref():
    from oracle import gsplat_ref
    if not gsplat_ref.available():
        pytest.skip("oracle/_ref/libgsplat_ref.so not built (oracle/gsplat_ref/Makefile needs /root/reference at build time)")
    return gsplat_ref.ops()


def n(t):
    return t.detach().cpu().numpy()


def check_iteration(a, b, N, vs_reference_kernels):
    """a: the C++ host layer, b: expected"""
    odd = np.nonzero(a["proj"]["radii"] != b["proj"]["radii"])[0]
    # see tests/gs_checks.py: ceil() of a 3-sigma extent within an ulp of an integer, FMA-contracted reference kernels only
    assert len(odd) <= (max(1, int(1e-4 * N)) if vs_reference_kernels else 0)
    assert np.all(np.abs(a["proj"]["radii"][odd] - b["proj"]["radii"][odd]) == 1)
    if len(odd) == 0:
        for k in ("tiles_per_gauss", "isect_ids", "flatten_ids", "tile_offsets"):
            assert np.array_equal(a[k], b[k]), k
    vis = b["proj"]["radii"] > 0
    same = np.ones(N, bool)
    same[odd] = False
    gc.close_frac("means2d", a["proj"]["means2d"][vis], b["proj"]["means2d"][vis], 2e-4, 2e-6)
    gc.close_frac("conics", a["proj"]["conics"][vis], b["proj"]["conics"][vis], 2e-6, 5e-5)
    gc.close_frac("depths", a["proj"]["depths"][vis], b["proj"]["depths"][vis], 2e-6, 2e-6)
    gc.close_frac("colors", a["colors"][vis], b["colors"][vis], 5e-6, 5e-5)
    gc.close_frac("render", a["render"], b["render"], 2e-4, 2e-4, 2e-4)
    gc.close_frac("alphas", a["alphas"], b["alphas"], 2e-4, 2e-4, 2e-4)
    gc.close_frac("rgb", a["rgb"], b["rgb"], 2e-4, 2e-4, 2e-4)
    ok = np.isfinite(b["depth"]) & np.isfinite(a["depth"])
    gc.close_frac("depth", a["depth"][ok], b["depth"][ok], 2e-4, 2e-4, 2e-4)
    assert abs(a["loss"] - b["loss"]) <= 1e-5 * max(1.0, abs(b["loss"]))
    gc.close_frac("v_render", a["v_render"][..., :3], b["v_render"][..., :3], 1e-9, 2e-4, 2e-4)
    gc.close_frac("v_alphas", a["v_alphas"], b["v_alphas"], 1e-9, 2e-4, 2e-4)
    bad = 1e-4 if vs_reference_kernels else 0.0
    for k in ("v_means2d", "v_conics", "v_opacities"):
        gc.close_scaled(k, a[k][vis & same], b[k][vis & same], 2e-3, bad)
    gc.close_scaled("v_colors", a["v_colors"][vis & same, :3], b["v_colors"][vis & same, :3], 2e-3, bad)
    for k in KEYS:
        gc.close_scaled("grad " + k, a["grads"][k].reshape(N, -1)[same], b["grads"][k].reshape(N, -1)[same], 3e-3, bad)


CASES = [(1500, 320, 192, 7, {}), (300, 96, 64, 3, {}), (200, 400, 300, 5, dict(scale_lo=0.05, scale_hi=0.4)),
         (20000, 1200, 680, 21, dict(scale_lo=0.004, scale_hi=0.02))]


@pytest.mark.parametrize("N,W,H,seed,kw", CASES[:3])
def test_ges_iteration_through_cxx_wrappers_matches_numpy_oracle(b200, N, W, H, seed, kw):
    from oracle import gs_oracle as go
    from oracle import gsplat_ref
    p = random_splats(N, seed=seed, **kw)
    c2w, K = camera(W, H, seed)
    ref_depth, base, gt = scene_images(W, H, seed)
    a = gsplat_ref.ges_iteration(p, c2w, K, W, H, ref_depth, base, gt, ops_ns=b200)
    b = go.ges_iteration(p, c2w, K, W, H, ref_depth, base, gt)
    check_iteration(a, b, N, False)


@pytest.mark.parametrize("N,W,H,seed,kw", CASES)
def test_ges_iteration_through_cxx_wrappers_matches_reference_wrappers(b200, gsref, N, W, H, seed, kw):
    from oracle import gsplat_ref
    p = random_splats(N, seed=seed, **kw)
    c2w, K = camera(W, H, seed)
    ref_depth, base, gt = scene_images(W, H, seed)
    a = gsplat_ref.ges_iteration(p, c2w, K, W, H, ref_depth, base, gt, ops_ns=b200)
    b = gsplat_ref.ges_iteration(p, c2w, K, W, H, ref_depth, base, gt)
    check_iteration(a, b, N, True)


def test_training_trajectory_through_cxx_wrappers(b200, engine_lib):
    """10 iterations of gesForward + L1 + backward through the C++ wrappers with torch.optim.Adam (= the reference's 6 x
    torch::optim::Adam) against the fused engine's gsb_gs_train_step: the two host routes into the same kernels agree."""
    from gps_slam_b200.engine import GaussianEngine
    from oracle import gsplat_ref
    W, H, N, iters = 320, 192, 3000, 10
    p = random_splats(N, seed=13)
    c2w, K = camera(W, H, 13)
    ref_depth, base, gt = scene_images(W, H, 13)
    model = gsplat_ref.RefGaussians(p, lrs=gc.LR, ops_ns=b200)
    losses_cxx = [model.train_iteration(c2w, K, W, H, ref_depth, base, gt)["loss"] for _ in range(iters)]
    dev = torch.device("cuda", 0)
    rd, bs, g = [torch.from_numpy(x).to(dev).contiguous() for x in (ref_depth, base, gt)]
    eng = GaussianEngine(W, H, capacity=N)
    try:
        eng.set_params(p)
        eng.initOptimizers()
        losses = []
        for _ in range(iters):
            eng.train_step(c2w, gc.intr_of(K, W, H), rd, bs, g)
            losses.append(eng.loss())
    finally:
        eng.close()
    assert losses_cxx[-1] < losses_cxx[0]
    np.testing.assert_allclose(losses_cxx, losses, rtol=2e-4)


@pytest.mark.parametrize("N,W,H,seed,kw,bg", [(1500, 320, 192, 7, {}, False), (3000, 400, 300, 9, dict(scale_lo=0.01, scale_hi=0.08), True)])
def test_raw_compositing_through_cxx_wrappers_matches_reference_wrappers(b200, gsref, N, W, H, seed, kw, bg):
    """render_method "raw": isectTiles + isectOffsetEncode + RasterizeToPixels::apply, forward and autograd backward"""
    from tests.test_gs_staged_gpu import scene
    s = scene(N, W, H, seed, **kw)
    g = torch.Generator(device=DEV).manual_seed(seed)
    radii, m2d, depths, conics = gsref.fully_fused_projection(s["means"], s["quats"], s["scales"], s["viewmat"], s["K"], W, H, 0.3, 0.01, 1e10, 0.0)[:4]
    radii = torch.clamp_max(radii, 100)
    vis = n(radii)[0] > 0
    tw, th = -(-W // 16), -(-H // 16)
    colors4 = torch.cat([torch.rand(1, N, 3, device=DEV, generator=g), depths[..., None]], 2).contiguous()
    opac = s["opac"].reshape(-1).contiguous()
    bg_t = torch.tensor([[0.3, 0.5, 0.7, 0.0]], device=DEV)
    v_render, v_alpha = torch.randn(1, H, W, 4, device=DEV, generator=g) * 1e-3, torch.randn(1, H, W, 1, device=DEV, generator=g) * 1e-3
    out = {}
    for name, o in (("ref", gsref), ("b200", b200)):
        tpg, isect, flat = o.isect_tiles(m2d, radii, depths, 16, tw, th)
        off = o.isect_offset_encode(isect, 1, tw, th)
        ins = [x.detach().clone().requires_grad_(True) for x in (m2d, conics, colors4, opac)]
        if bg:
            render, alpha = o.rasterize_raw_bg(ins[0], ins[1], ins[2], ins[3], bg_t, W, H, 16, off, flat, False)
        else:
            render, alpha = o.rasterize_raw(ins[0], ins[1], ins[2], ins[3], W, H, 16, off, flat, False)
        grads = torch.autograd.grad([render, alpha], ins, [v_render, v_alpha])
        out[name] = dict(tpg=n(tpg), isect=n(isect), flat=n(flat), off=n(off), render=n(render), alpha=n(alpha), grads=[n(x) for x in grads])
    a, b = out["b200"], out["ref"]
    for k in ("tpg", "isect", "flat", "off"):
        assert np.array_equal(a[k], b[k]), k
    gc.close_frac("render", a["render"], b["render"], 3e-4, 3e-4, 3e-4)
    gc.close_frac("alphas", a["alpha"], b["alpha"], 3e-4, 3e-4, 3e-4)
    for name, x, y in zip(("v_means2d", "v_conics", "v_colors", "v_opacities"), a["grads"], b["grads"]):
        gc.close_scaled(name, x.reshape(N, -1)[vis], y.reshape(N, -1)[vis], 3e-3, 2e-4)


def test_offset_encode_without_engine_state(b200):
    """isectOffsetEncodeNoDepth on ids that are NOT the tensor the last binning returned takes the stateless route"""
    W, H, N = 320, 192, 1500
    from tests.test_gs_staged_gpu import scene
    s = scene(N, W, H, 7)
    radii, m2d, depths, conics = b200.fully_fused_projection(s["means"], s["quats"], s["scales"], s["viewmat"], s["K"], W, H, 0.3, 0.01, 1e10, 0.0)[:4]
    tw, th = -(-W // 16), -(-H // 16)
    tpg, isect, flat, _, _ = b200.isect_tiles_no_depth(m2d, radii, depths, 16, tw, th)
    off = b200.isect_offset_encode_no_depth(isect, 1, tw, th)
    off2 = b200.isect_offset_encode_no_depth(isect.clone(), 1, tw, th)
    assert off.dtype == off2.dtype == torch.int32 and torch.equal(off, off2)
    assert int(tpg.sum()) == isect.numel() == flat.numel() > 0


@pytest.mark.parametrize("H,W", [(68, 120), (340, 600)])
def test_fused_ssim_through_cxx_wrapper_matches_reference_wrapper(b200, gsref, H, W):
    g = torch.Generator(device=DEV).manual_seed(H)
    img2 = torch.rand(1, 3, H, W, device=DEV, generator=g)
    img1 = (img2 + 0.1 * torch.randn(1, 3, H, W, device=DEV, generator=g)).clamp(0, 1)
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    for padding in ("same", "valid"):
        res = []
        dL = None
        for o in (gsref, b200):
            a = img1.clone().requires_grad_(True)
            m = o.fused_ssim_map(C1, C2, a, img2, padding, True)
            if dL is None:
                dL = torch.randn(m.shape, device=DEV, generator=g)
            (grad,) = torch.autograd.grad([m], [a], [dL])
            res.append((n(m), n(grad)))
        np.testing.assert_allclose(res[1][0], res[0][0], rtol=2e-4, atol=2e-5)
        gc.close_scaled("dL_dimg1 " + padding, res[1][1], res[0][1], 2e-4)


def brute_force_knn(pts):
    """mean of the 3 smallest squared distances to the other points, float64 on the CPU (small P only)"""
    p = pts.astype(np.float64)
    d = ((p[:, None, :] - p[None, :, :]) ** 2).sum(-1)
    np.fill_diagonal(d, np.inf)
    return np.sort(d, 1)[:, :3].mean(1)


def knn_cases():
    rng = np.random.default_rng(5)
    surface = rng.uniform(-2, 2, (4000, 3)).astype(np.float32)
    surface[:, 2] = 0.3 * np.sin(surface[:, 0] * 2) + 0.01 * rng.standard_normal(4000)       # points on a sheet, like a depth map
    clustered = np.concatenate([rng.normal(0, 0.01, (1500, 3)), rng.normal(3, 0.5, (1500, 3)), [[50, 50, 50], [-40, 0, 0]]]).astype(np.float32)
    dup = np.repeat(rng.uniform(0, 1, (500, 3)).astype(np.float32), 3, 0)                      # exact duplicates: zero distances
    line = np.zeros((777, 3), np.float32)
    line[:, 0] = np.linspace(0, 1, 777)                                                       # degenerate extent on two axes
    return dict(surface=surface, clustered=clustered, duplicates=dup, line=line, four=surface[:4].copy(), uniform=rng.uniform(0, 1, (5000, 3)).astype(np.float32))


@pytest.mark.parametrize("name", ["surface", "clustered", "duplicates", "line", "four", "uniform"])
def test_simple_knn_is_exact(b200, name):
    pts = knn_cases()[name]
    got = n(b200.simple_knn(torch.from_numpy(pts).to(DEV)))
    exp = brute_force_knn(pts)
    np.testing.assert_allclose(got, exp, rtol=2e-5, atol=1e-12)


def test_simple_knn_few_points_like_reference(b200):
    """fewer than 4 points: the missing neighbours stay at FLT_MAX (simple_knn.cu:160,187): (d1 + d2 + FLT_MAX) / 3"""
    got = n(b200.simple_knn(torch.tensor([[0.0, 0, 0], [1, 0, 0], [0, 2, 0]], device=DEV)))
    assert np.all(got > 1e38) and np.all(np.isfinite(got))
    assert np.all(np.isinf(n(b200.simple_knn(torch.zeros(1, 3, device=DEV)))))


@pytest.mark.parametrize("P", [50000, 300000])
def test_simple_knn_matches_reference_kernel_bit_for_bit(b200, gsref, P):
    rng = np.random.default_rng(P)
    pts = rng.uniform(-3, 3, (P, 3)).astype(np.float32)
    pts[:, 1] = 0.2 * np.cos(pts[:, 0]) + 0.005 * rng.standard_normal(P).astype(np.float32)
    t = torch.from_numpy(pts).to(DEV)
    exp = n(gsref.simple_knn(t))
    got = n(b200.simple_knn(t))
    assert np.array_equal(got.view(np.uint32), exp.view(np.uint32))
