"""Multi-GPU parity ON HARDWARE: the same short sequence through a world-size-2 run (torchrun, NCCL over NVLink, Gaussians sharded by
spatial block, one exchange of the partial image per optimiser iteration) and through a single-GPU run must end at the same loss,
the same Gaussian count and the same PSNR.  Not bit-exact: the summed image is re-associated across ranks (fp32), which can flip a
pixel that sits on the spawn threshold.  Skipped on a box with fewer than 2 GPUs."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_match_one(engine_lib, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from tests import mp_slam_worker
    n_frames, scale = 31, 0.5
    out = str(tmp_path / "w2.json")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29731",
           os.path.join(ROOT, "tests", "mp_slam_worker.py"), out, str(n_frames), str(scale)]
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    two = json.load(open(out))
    one = mp_slam_worker.run(n_frames, scale, 1, 0, 0)
    assert one["overflow"] == 0 and two["overflow"] == 0
    assert 0.3 * two["gaussians"] < two["gaussians_this_rank"] < 0.7 * two["gaussians"], two      # the shard is really a shard
    assert abs(two["gaussians"] - one["gaussians"]) <= 0.005 * one["gaussians"] + 2, (two["gaussians"], one["gaussians"])
    assert abs(two["loss"] - one["loss"]) <= 2e-3 * one["loss"], (two["loss"], one["loss"])
    assert np.abs(np.array(two["psnr"]) - np.array(one["psnr"])).max() < 0.05, (two["psnr"], one["psnr"])
