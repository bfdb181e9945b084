#!/usr/bin/env python
"""Generates tests/golden/gs_ref_golden.npz by RUNNING THE REFERENCE's gsplat CUDA kernels on a B200 (needs a GPU and
oracle/_ref/libgsplat_ref.so, see oracle/gsplat_ref/Makefile):

  gpurun -- 'python tests/golden/make_golden_gs.py gpurun_out/gs_ref_golden.npz'   then copy the file into tests/golden/

Two seeded cases of tests/helpers_gs.py (inputs are regenerated from the seed by the tests, only outputs are stored):
every integer output of one gesForward (radii, tiles_per_gauss, isect ids, flatten ids, tile offsets), per-splat projection
outputs and colours, the loss, per-splat raster gradients, the six parameter gradients, and for the small case the rendered
image.  The CPU suite checks oracle/gs_oracle.py against it; the GPU suite checks the engine against it."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

CASES = {"a": dict(N=300, W=96, H=64, seed=3, image=True), "b": dict(N=1500, W=320, H=192, seed=7, image=False)}


def main():
    from oracle import gsplat_ref
    from tests.helpers_gs import camera, random_splats, scene_images
    out = {}
    for tag, c in CASES.items():
        p = random_splats(c["N"], seed=c["seed"])
        c2w, K = camera(c["W"], c["H"], c["seed"])
        ref_depth, base, gt = scene_images(c["W"], c["H"], c["seed"])
        r = gsplat_ref.ges_iteration(p, c2w, K, c["W"], c["H"], ref_depth, base, gt)
        vis = r["proj"]["radii"] > 0
        z = lambda a: np.where(vis.reshape((-1,) + (1,) * (a.ndim - 1)), a, 0).astype(a.dtype)   # culled lanes hold torch::empty garbage
        out[tag + "_radii"] = r["proj"]["radii"].astype(np.int32)
        out[tag + "_means2d"] = z(r["proj"]["means2d"])
        out[tag + "_depths"] = z(r["proj"]["depths"])
        out[tag + "_conics"] = z(r["proj"]["conics"])
        out[tag + "_colors"] = z(r["colors"])
        for k in ("tiles_per_gauss", "isect_ids", "flatten_ids", "tile_offsets"):
            out[tag + "_" + k] = r[k]
        out[tag + "_loss"] = np.float64(r["loss"])
        for k in ("v_means2d", "v_conics", "v_colors", "v_opacities"):
            out[tag + "_" + k] = z(r[k])
        for k, g in r["grads"].items():
            out[tag + "_grad_" + k] = g
        if c["image"]:
            out[tag + "_rgb"] = r["rgb"]
            out[tag + "_alphas"] = r["alphas"]
            out[tag + "_render"] = r["render"]
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "gs_ref_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
