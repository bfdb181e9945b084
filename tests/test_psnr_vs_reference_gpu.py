"""north_star: "PSNR within 0.1 dB" of the reference's own path.  Five local-optimisation cycles at the full 1200x680 through the
engine's loop and through the reference-kernel loop (tests/ref_slam.py): the two models must render the training views at the same
PSNR (|delta| <= 0.1 dB), hold Gaussian counts within 1 % of each other after every cycle, and agree with each other far better than
either agrees with the input frames."""
import pytest

pytestmark = pytest.mark.gpu


def test_five_cycles_full_resolution(engine_lib):
    from oracle import gsplat_ref
    if not gsplat_ref.available():
        pytest.skip("oracle/_ref/libgsplat_ref.so not built (needs /root/reference at build time)")
    from tests import ref_slam
    r = ref_slam.run_both(51, 1.0)
    e, k = r["engine"], r["reference_kernels"]
    assert abs(r["psnr_vs_reference_db"]) <= 0.1, r
    assert len(e["gaussians_after_each_cycle"]) == 5
    for a, b in zip(e["gaussians_after_each_cycle"], k["gaussians_after_each_cycle"]):
        assert abs(a - b) <= 0.01 * max(a, b) + 5, (e["gaussians_after_each_cycle"], k["gaussians_after_each_cycle"])
    # the first cycle spawns from the TSDF colour error alone: identical pixel sets
    assert e["spawned_each_cycle"][0] == k["spawned_each_cycle"][0]
    assert r["psnr_engine_vs_reference_render_db"] > e["psnr_db"] + 10.0, r
