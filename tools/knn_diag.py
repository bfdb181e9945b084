"""diagnostic: gsplat_b200.simple_knn vs the reference kernel vs float64 brute force"""
import numpy as np, torch, sys
sys.path.insert(0, ".")
from gps_slam_b200 import build
from oracle import gsplat_ref
torch.ops.load_library(build.build_torch_shim())
b200, ref = torch.ops.gsplat_b200, gsplat_ref.ops()
for P in (50000, 300000):
    rng = np.random.default_rng(P)
    pts = rng.uniform(-3, 3, (P, 3)).astype(np.float32)
    pts[:, 1] = 0.2 * np.cos(pts[:, 0]) + 0.005 * rng.standard_normal(P).astype(np.float32)
    t = torch.from_numpy(pts).cuda()
    a, b = b200.simple_knn(t), ref.simple_knn(t)
    bad = (a != b).nonzero().flatten()
    print("P", P, "mismatches", bad.numel(), "max rel", float(((a - b).abs() / b).max()))
    idx = bad[:2000] if bad.numel() else torch.arange(1000, device="cuda")
    q = t[idx].double()
    d = ((q[:, None, :] - t.double()[None, :, :]) ** 2).sum(-1)
    d[torch.arange(idx.numel()), idx] = float("inf")
    exact = d.topk(3, largest=False).values.mean(1)
    print("  on mismatching points: ours rel err %.3g  reference rel err %.3g" % (float(((a[idx].double() - exact).abs() / exact).max()),
                                                                                float(((b[idx].double() - exact).abs() / exact).max())))
    # fp32 with and without fma for the first mismatch
    if bad.numel():
        i = int(bad[0]); print("  first:", i, float(a[i]), float(b[i]), float(exact[0]))
    torch.cuda.synchronize()
    for name, f in (("ours", b200.simple_knn), ("reference", ref.simple_knn)):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        f(t); e0.record()
        for _ in range(5): f(t)
        e1.record(); torch.cuda.synchronize()
        print("  %s: %.3f ms" % (name, e0.elapsed_time(e1) / 5))
