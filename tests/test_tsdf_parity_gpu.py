"""GPU parity of the TSDF path against the REFERENCE's own CPU engine
(oracle/_ref/libitm_ref_exact.so, built from /root/reference by oracle/itm_ref/Makefile; single-threaded so that its
racy marking loop is deterministic).  Everything is compared BIT-EXACT: hash table, free-list heads, visible list,
visible types, voxel block array, float depth, expected-depth ranges, raycast points, ICP maps, free-view render."""
import numpy as np
import pytest

from gps_slam_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint8).reshape(-1)


def assert_same(name, ours, ref):
    assert ours.shape == ref.shape, "%s: shape %s vs %s" % (name, ours.shape, ref.shape)
    b0, b1 = bits(ours), bits(ref)
    if not np.array_equal(b0, b1):
        bad = np.nonzero(b0 != b1)[0]
        item = ours.dtype.itemsize
        first = np.unique(bad // item)[:5]
        raise AssertionError("%s: %d differing elements of %d; first at %s: ours=%s ref=%s" % (
            name, len(np.unique(bad // item)), ours.size, first, ours.reshape(-1)[first], ref.reshape(-1)[first]))


def compare_state(eng, ref, tag):
    assert_same(tag + " depth_f", eng.depth(), ref.depth())
    he, hr = eng.hash_entries(), ref.hash_entries()
    for f in ("pos", "offset", "ptr"):
        assert_same(tag + " hash." + f, np.ascontiguousarray(he[f]), np.ascontiguousarray(hr[f]))
    assert eng.counter(0) == ref.last_free_block(), tag + " lastFreeBlockId"
    assert eng.counter(1) == ref.last_free_excess(), tag + " lastFreeExcessListId"
    assert eng.counter(3) == 0, tag + " engine error flag"
    assert_same(tag + " visibleEntryIDs", eng.visible_ids(), np.array(ref.visible_ids()))
    assert_same(tag + " entriesVisibleType", eng.visible_types(), np.array(ref.visible_types()))
    first = ref.last_free_block() + 1
    ve = eng.voxels()[first:]
    vr = ref.voxels()[first:]
    for f in ("sdf", "w_depth", "clr", "w_color"):
        assert_same(tag + " voxel." + f, np.ascontiguousarray(ve[f]), np.ascontiguousarray(vr[f]))
    mm = eng.minmax()
    assert_same(tag + " minmax", mm, np.ascontiguousarray(ref.minmax()[:mm.shape[0], :mm.shape[1]]))
    assert_same(tag + " raycast", eng.raycast(), ref.raycast())
    assert_same(tag + " pointsMap", eng.points_map(), ref.points_map())
    assert_same(tag + " normalsMap", eng.normals_map(), ref.normals_map())


def run_sequence(intr, n_frames, variant, check_every=1, free_view_at=()):
    from gps_slam_b200.engine import TsdfEngine
    from oracle.itm_ref import ItmRef
    poses, frames = syn.sequence(n_frames, intr)
    ref = ItmRef(intr, tracker=0, threads=1, kind="exact")
    eng = TsdfEngine(intr, tracker=0, integrate_variant=variant)
    try:
        for i in range(n_frames):
            rgba, d = frames[i][0].numpy(), frames[i][1].numpy()
            c2w = syn.c2w_to_colmajor(poses[i])
            ref.process_frame(rgba, d, c2w)
            eng.ProcessFrame(rgba, d, c2w)
            Me, iMe = eng.pose()
            Mr, iMr = ref.pose()
            assert_same("frame %d pose M" % i, Me, Mr)
            assert_same("frame %d pose invM" % i, iMe, iMr)
            if i % check_every == 0 or i == n_frames - 1:
                compare_state(eng, ref, "frame %d" % i)
            if i in free_view_at:
                # free-view raycast from a pose between two frames, slightly different intrinsics
                c2w_f = syn.c2w_to_colmajor(syn.trajectory(n_frames + 40)[i + 20])
                intr_f = dict(intr, fx=intr["fx"] * 0.97, fy=intr["fy"] * 0.97)
                ref.run_raycast(c2w_f, intr_f)
                eng.runRaycast(c2w_f, intr_f)
                mm = eng.minmax(live=False)
                assert_same("free minmax", mm, np.ascontiguousarray(ref.minmax(live=False)[:mm.shape[0], :mm.shape[1]]))
                assert_same("free vertex", eng.raycast(live=False), ref.raycast(live=False))
                assert_same("free image", eng.free_image(), ref.raycast_image(live=False))
    finally:
        eng.close()
        ref.close()


def test_replica_shape_bit_exact(engine_lib):
    """full 1200x680 Replica-shaped frames"""
    run_sequence(syn.intrinsics("replica"), 4, 0, free_view_at=(3,))


def test_quarter_res_long_sequence(engine_lib):
    """300x170, 40 frames: exercises excess-list chaining, visible-list ageing and repeated integration"""
    run_sequence(syn.intrinsics("replica", 0.25), 40, 0, check_every=8, free_view_at=(20, 39))


def test_invalid_and_empty_depth(engine_lib):
    """all-invalid depth (zeros / negatives) must allocate nothing and leave the raycast empty, like the reference"""
    from gps_slam_b200.engine import TsdfEngine
    from oracle.itm_ref import ItmRef
    intr = syn.intrinsics("replica", 0.25)
    poses, frames = syn.sequence(2, intr)
    ref = ItmRef(intr, tracker=0, threads=1, kind="exact")
    eng = TsdfEngine(intr, tracker=0)
    try:
        rgba = frames[0][0].numpy()
        d = np.zeros_like(frames[0][1].numpy())
        d[::3, ::5] = -7
        c2w = syn.c2w_to_colmajor(poses[0])
        ref.process_frame(rgba, d, c2w)
        eng.ProcessFrame(rgba, d, c2w)
        compare_state(eng, ref, "empty frame")
        assert eng.counter(2) == 0 or len(ref.visible_ids()) == eng.counter(2)
        # ragged: half the image invalid, then a normal frame
        d2 = frames[1][1].numpy().copy()
        d2[:, : d2.shape[1] // 2] = 0
        c2w = syn.c2w_to_colmajor(poses[1])
        ref.process_frame(frames[1][0].numpy(), d2, c2w)
        eng.ProcessFrame(frames[1][0].numpy(), d2, c2w)
        compare_state(eng, ref, "half frame")
    finally:
        eng.close()
        ref.close()
