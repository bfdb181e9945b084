"""Host-side pieces of bench.py that need no GPU: where the timed window sits in the sequence, and how the committed ncu captures are picked
(by round / version number, not by file-name order: round 1's bug)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def test_timed_window_sits_at_the_end_of_the_sequence():
    import bench
    # driver shape: 20 steps of 10 frames after 5 warm-up steps, 50 frames left after the window
    w0, t0, t1 = bench.window_of(2000, 20, 5)
    assert (w0, t0, t1) == (1700, 1750, 1950)
    for steps, warmup in ((10, 3), (1, 1), (50, 5), (100, 10)):
        w0, t0, t1 = bench.window_of(2000, steps, warmup)
        assert t1 - t0 == steps * bench.FRAMES_PER_STEP and t0 - w0 == warmup * bench.FRAMES_PER_STEP
        assert w0 >= 0 and t1 <= 2000 and w0 % bench.FRAMES_PER_STEP == 0
    # a request longer than the sequence starts at frame 0 and extends it
    w0, t0, t1 = bench.window_of(2000, 200, 10)
    assert w0 == 0 and t1 - t0 == 2000 and t1 > 2000


def test_newest_ncu_capture_is_picked_by_round_and_version(tmp_path, monkeypatch):
    import bench
    prof = tmp_path / "profiles"
    prof.mkdir()
    for name, bytes_ in (("r01_ncu_full_v8_avg.json", 1.0), ("r01_ncu_full_v26_avg.json", 2.0), ("r02_ncu_full_v3_avg.json", 3.0)):
        (prof / name).write_text(json.dumps({"kernels": {"gs::k_raster_bwd<0>": {"dram_bytes_per_launch": bytes_, "issue_slots_busy_pct": 70.0}}}))
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    traffic, src = bench.ncu_traffic()
    assert src == "r02_ncu_full_v3_avg.json" and traffic["k_raster_bwd"] == 3.0 and traffic["issue:k_raster_bwd"] == 70.0
    (prof / "r02_ncu_full_v3_avg.json").unlink()
    traffic, src = bench.ncu_traffic()
    assert src == "r01_ncu_full_v26_avg.json" and traffic["k_raster_bwd"] == 2.0     # v26 after v8, not "v8" > "v26"


def test_committed_profiles_parse():
    """what bench.py quotes from profiles/ exists and has the keys it reads"""
    import bench
    traffic, src = bench.ncu_traffic()
    assert src and "k_raster_bwd" in traffic and "k_integrate_tma" in traffic
    ref = bench.psnr_vs_reference()
    assert ref and abs(ref["psnr_vs_reference_db"]) <= 0.1
