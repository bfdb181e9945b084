"""model.pt (RawGaussianParams::saveTensor / loadTensor, reference src/raw_gs_param.cpp:220-254) without a GPU: the archive written by the
C++ host layer round-trips bit for bit, is a torch::serialize archive any libtorch program can open (torch.jit.load sees the seven
named tensors), and point_cloud.ply round-trips too."""
import numpy as np
import torch

from gps_slam_b200 import checkpoint as ck
from tests.helpers_gs import random_splats


def test_model_pt_round_trip_and_container(engine_lib, tmp_path):
    p = random_splats(321, seed=4)
    path = str(tmp_path / "model.pt")
    exposure = np.tile(np.eye(3, 4, dtype=np.float32)[None], (2, 1, 1)) * 1.5
    ck.save_model_pt(path, p, exposure)
    got, exp_back = ck.load_model_pt(path)
    for k in ck.MODEL_PT_KEYS:
        assert np.array_equal(got[k].reshape(-1).view(np.uint32), np.asarray(p[k], np.float32).reshape(-1).view(np.uint32)), k
    assert got["featuresRest"].shape == (321, 15, 3) and got["opacities"].shape == (321, 1)
    assert np.array_equal(exp_back, exposure)
    module = torch.jit.load(path)            # the container is the TorchScript archive torch::serialize::OutputArchive writes
    names = {n for n, _ in module.named_parameters()} | {n for n, _ in module.named_buffers()}
    assert {"means", "scales", "quats", "featuresDc", "featuresRest", "opacities", "exposure"} <= names
    assert torch.equal(dict(list(module.named_parameters()) + list(module.named_buffers()))["means"], torch.from_numpy(np.asarray(p["means"], np.float32)))


def test_ply_round_trip_cpu(tmp_path):
    p = random_splats(100, seed=2)
    path = str(tmp_path / "point_cloud.ply")
    ck.save_ply(path, p)
    q = ck.load_ply(path)
    for k in ck.MODEL_PT_KEYS:
        assert np.array_equal(np.asarray(q[k], np.float32).reshape(-1), np.asarray(p[k], np.float32).reshape(-1)), k
