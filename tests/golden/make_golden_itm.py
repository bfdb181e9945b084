#!/usr/bin/env python
"""Generates tests/golden/itm_ref_golden.npz by RUNNING THE REFERENCE: the reference's InfiniTAM CPU engine
(oracle/_ref/libitm_ref_exact.so, built from /root/reference by oracle/itm_ref/Makefile; IEEE fp32, one thread) on the seeded
synthetic sequence of gps_slam_b200/synthetic.py at 1/4 of the Replica resolution (300x170).

  python tests/golden/make_golden_itm.py          (build container; needs oracle/_ref)

Contents, per frame of a 4-frame ground-truth-pose run: visible block list, free-list heads, SHA-256 of the hash table
fields and of the allocated voxel blocks, every 4th pixel of the raycast image, the points/normals maps' SHA-256; one free-view
raycast digest; and for the two tracker flavours (extended, icp) the tracked poses and tracker results of 4 frames.
The fixture lets the parity tests run where /root/reference (hence oracle/_ref) is absent."""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

N_FRAMES = 4
SCALE = 0.25


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def table_digest(h):
    return sha(np.concatenate([np.ascontiguousarray(h["pos"]).reshape(-1).astype(np.int32), np.ascontiguousarray(h["offset"]).reshape(-1),
                               np.ascontiguousarray(h["ptr"]).reshape(-1)]))


def voxel_digest(v, first):
    v = v[first:]
    return sha(np.concatenate([np.ascontiguousarray(v[f]).reshape(-1).astype(np.int32) for f in ("sdf", "w_depth", "clr", "w_color")]))


def free_view(poses_long, i, intr):
    from gps_slam_b200 import synthetic as syn
    return syn.c2w_to_colmajor(poses_long[i + 20]), dict(intr, fx=intr["fx"] * 0.97, fy=intr["fy"] * 0.97)


def main():
    from gps_slam_b200 import synthetic as syn
    from oracle.itm_ref import ItmRef
    intr = syn.intrinsics("replica", SCALE)
    poses, frames = syn.sequence(N_FRAMES, intr)
    out = {"scale": np.float32(SCALE), "n_frames": np.int32(N_FRAMES)}
    ref = ItmRef(intr, tracker=0, threads=1, kind="exact")
    for i in range(N_FRAMES):
        ref.process_frame(frames[i][0].numpy(), frames[i][1].numpy(), syn.c2w_to_colmajor(poses[i]))
        first = ref.last_free_block() + 1
        out["f%d_visible_ids" % i] = np.array(ref.visible_ids())
        out["f%d_free_heads" % i] = np.array([ref.last_free_block(), ref.last_free_excess()], np.int32)
        out["f%d_table_sha" % i] = table_digest(ref.hash_entries())
        out["f%d_voxel_sha" % i] = voxel_digest(ref.voxels(), first)
        out["f%d_raycast_sub4" % i] = np.array(ref.raycast()[::4, ::4])
        out["f%d_raycast_sha" % i] = sha(ref.raycast())
        out["f%d_points_sha" % i] = sha(ref.points_map())
        out["f%d_normals_sha" % i] = sha(ref.normals_map())
    c2w_f, intr_f = free_view(syn.trajectory(N_FRAMES + 40), N_FRAMES - 1, intr)
    ref.run_raycast(c2w_f, intr_f)
    out["free_vertex_sha"] = sha(ref.raycast(live=False))
    out["free_image_sha"] = sha(ref.raycast_image(live=False))
    out["free_image_sub4"] = np.array(ref.raycast_image(live=False)[::4, ::4])
    ref.close()
    for flavour, name in ((1, "extended"), (2, "icp")):
        ref = ItmRef(intr, tracker=flavour, threads=1, kind="exact")
        ref.set_pose_invM(syn.c2w_to_colmajor(poses[0]))
        Ms, res = [], []
        for i in range(N_FRAMES):
            ref.process_frame(frames[i][0].numpy(), frames[i][1].numpy(), None)
            Ms.append(ref.pose()[0].copy())
            res.append(ref.tracker_result())
        out["track_%s_M" % name] = np.stack(Ms)
        out["track_%s_result" % name] = np.array(res, np.int32)
        ref.close()
    path = os.path.join(HERE, "itm_ref_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
