"""Voxel hash sharded by spatial block against the single-GPU engine, BIT-EXACT.

`world` engines live in one process on one device and map each other's segments directly (gsb_tsdf_shard_attach_local) -- the same
kernels, barriers and peer stores as one process per GPU over NVLink, minus the IPC handle exchange (tests/test_parallel_gpu.py covers
that on a multi-GPU box).  Replicated state (hash table, visible list, visible types, free-list heads) must equal the single engine's on
EVERY rank; every voxel block must equal the single engine's on the rank that owns it (hashIndex(blockPos) mod world); the ICP maps must
equal the single engine's rows on the rank that holds the slab; live and free-view vertex / colour images must equal the single engine's on
EVERY rank (each rank marches its strips of rows and stores them into everybody's image)."""
import threading

import numpy as np
import pytest

from gps_slam_b200 import synthetic as syn
from tests.test_tsdf_parity_gpu import assert_same

pytestmark = pytest.mark.gpu


def _engines(intr, world, tracker=0, mode=0):
    from gps_slam_b200.engine import TsdfEngine
    single = TsdfEngine(intr, tracker=tracker)
    shards = [TsdfEngine(intr, tracker=tracker, rank=r, world=world) for r in range(world)]
    for e in shards:
        e.attach_local(shards)
        e.set_shard_mode(mode)
    return single, shards


def _compare(single, shards, world, tag, maps_everywhere=False, mode=0):
    from gps_slam_b200 import parallel
    hs = single.hash_entries()
    first = single.counter(0) + 1
    vs = single.voxels()
    owner = np.full(single.num_blocks, -1, np.int64)
    alloc = hs["ptr"] >= 0
    owner[hs["ptr"][alloc]] = parallel.voxel_block_owner(hs["pos"][alloc][:, :3], world)
    assert (owner[first:] >= 0).all(), "every allocated block has a hash entry"
    H = single.h
    n_own = 0
    for r, e in enumerate(shards):
        assert e.counter(3) == 0 and e.shard_error() == 0, "%s rank %d error flag" % (tag, r)
        he = e.hash_entries()
        for f in ("pos", "offset", "ptr"):
            assert_same("%s rank %d hash.%s" % (tag, r, f), np.ascontiguousarray(he[f]), np.ascontiguousarray(hs[f]))
        assert e.counter(0) == single.counter(0) and e.counter(1) == single.counter(1)
        assert_same("%s rank %d visibleEntryIDs" % (tag, r), e.visible_ids(), single.visible_ids())
        assert_same("%s rank %d entriesVisibleType" % (tag, r), e.visible_types(), single.visible_types())
        mine = np.nonzero(owner == r)[0]
        n_own += len(mine)
        ve = e.voxels()
        if mode == 1:
            mine = np.nonzero(owner >= 0)[0]      # the owners have stored their blocks into every rank
        for f in ("sdf", "w_depth", "clr", "w_color"):
            assert_same("%s rank %d voxel.%s" % (tag, r, f), np.ascontiguousarray(ve[f][mine]), np.ascontiguousarray(vs[f][mine]))
        if mode == 0:
            # a block this rank does not own was never integrated here
            others = np.nonzero((owner >= 0) & (owner != r))[0]
            assert (ve["w_depth"][others] == 0).all(), "%s rank %d holds voxels of blocks it does not own" % (tag, r)
        y0, y1 = (0, H) if maps_everywhere else e.shard_rows()
        assert_same("%s rank %d minmax" % (tag, r), e.minmax(), single.minmax())
        assert_same("%s rank %d raycast" % (tag, r), e.raycast(), single.raycast())   # every rank receives every row
        assert_same("%s rank %d pointsMap rows" % (tag, r), e.points_map()[y0:y1], single.points_map()[y0:y1])
        assert_same("%s rank %d normalsMap rows" % (tag, r), e.normals_map()[y0:y1], single.normals_map()[y0:y1])
    assert n_own == single.num_blocks - first
    rows = [e.shard_rows() for e in shards]
    assert rows[0][0] == 0 and rows[-1][1] == H and all(rows[i][1] == rows[i + 1][0] for i in range(world - 1)), rows


@pytest.mark.parametrize("world,scale,n_frames,mode", [(2, 0.5, 6, 0), (3, 0.25, 24, 0), (8, 0.25, 6, 0), (2, 0.5, 6, 1), (3, 0.25, 24, 1)])
def test_sharded_scene_is_bit_exact(engine_lib, world, scale, n_frames, mode):
    """mode 0: storage-sharded, raycasts read the owners' voxels; mode 1: the owner integrates and stores the block into every rank"""
    import torch
    intr = syn.intrinsics("replica", scale)
    poses, frames = syn.sequence(n_frames, intr)
    dev = torch.device("cuda", 0)
    single, shards = _engines(intr, world, mode=mode)
    # resident frames, kept alive for the whole test: the engines read them asynchronously on their own streams
    dev_frames = [(frames[i][0].to(dev), frames[i][1].to(dev)) for i in range(n_frames)]
    torch.cuda.synchronize()
    try:
        for i in range(n_frames):
            rgba, d = dev_frames[i]
            c2w = syn.c2w_to_colmajor(poses[i])
            single.ProcessFrameDevice(rgba, d, c2w)
            for e in shards:               # the calls only enqueue: one host thread drives every rank
                e.ProcessFrameDevice(rgba, d, c2w)
            if i % 6 == 5 or i == n_frames - 1:
                for e in shards:
                    e.sync()
                _compare(single, shards, world, "frame %d" % i, mode=mode)
                # every rank integrates a real share of the visible blocks
                own = [e.counter(6) for e in shards]
                assert sum(own) == single.counter(2) and min(own) > 0.5 * single.counter(2) / world, own
                c2w_f = syn.c2w_to_colmajor(syn.trajectory(n_frames + 40)[i + 20])
                intr_f = dict(intr, fx=intr["fx"] * 0.97, fy=intr["fy"] * 0.97)
                single.runRaycast(c2w_f, intr_f)
                for e in shards:
                    e.runRaycast(c2w_f, intr_f)
                for r, e in enumerate(shards):
                    e.sync()
                    assert_same("free vertex rank %d" % r, e.raycast(live=False), single.raycast(live=False))
                    assert_same("free image rank %d" % r, e.free_image(), single.free_image())
                    assert e.shard_error() == 0
    finally:
        single.close()
        for e in shards:
            e.close()


@pytest.mark.parametrize("tracker", [1, 2])
def test_sharded_scene_with_tracking(engine_lib, tracker):
    """online tracking over a sharded scene: every rank must track the SAME pose, bit for bit (the replicated allocation depends on
    it), and that pose must match the single engine's.  Tracking reads the pose back once per frame, so each rank gets its own host
    thread here (one process per GPU in production)."""
    import os
    import torch
    world, n_frames = 2, 8
    os.environ["GSB_ICP_CTAS_PER_SM"] = "1"    # both ranks' persistent trackers must fit on the one device at the same time
    intr = syn.intrinsics("replica", 0.5)
    poses, frames = syn.sequence(n_frames, intr)
    dev = torch.device("cuda", 0)
    single, shards = _engines(intr, world, tracker=tracker)
    dev_frames = [(frames[i][0].to(dev), frames[i][1].to(dev)) for i in range(n_frames)]
    torch.cuda.synchronize()
    first = syn.c2w_to_colmajor(poses[0])
    got = [[] for _ in range(world)]
    errs = []

    def drive(eng, out):
        try:
            torch.cuda.set_device(0)
            eng.set_pose(first)
            for rgba, d in dev_frames:
                eng.ProcessFrameDevice(rgba, d, None)
                out.append(eng.pose()[1].copy())
            eng.sync()
        except Exception as ex:  # noqa: BLE001
            errs.append(ex)

    try:
        ref = []
        drive(single, ref)
        threads = [threading.Thread(target=drive, args=(shards[r], got[r])) for r in range(world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join(300)
        assert not errs, errs
        for r in range(world):
            assert shards[r].shard_error() == 0
            assert len(got[r]) == n_frames
            for i in range(n_frames):
                assert_same("frame %d pose rank %d vs rank 0" % (i, r), got[r][i], got[0][i])
        # the 29 sums are added in a different order (per-rank partial sums, then ranks): same pose up to fp32 re-association
        err = max(np.abs(got[0][i] - ref[i]).max() for i in range(n_frames))
        assert err < 2e-4, err
        # every rank holds the complete ICP maps while tracking is on, identical on every rank
        assert_same("pointsMap rank 1 vs rank 0", shards[1].points_map(), shards[0].points_map())
        assert_same("normalsMap rank 1 vs rank 0", shards[1].normals_map(), shards[0].normals_map())
        assert shards[0].tracker_result()[2] == shards[1].tracker_result()[2] > 0
    finally:
        os.environ.pop("GSB_ICP_CTAS_PER_SM", None)
        single.close()
        for e in shards:
            e.close()
