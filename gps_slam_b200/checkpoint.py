"""Checkpoint / wire formats of the reference, on the host:

  * Gaussians: the 3DGS point_cloud.ply written by RawGaussianParams::savePly (reference src/raw_gs_param.cpp:159-218): binary little
    endian, per vertex x y z, nx ny nz (zeros), f_dc_0..2, f_rest_0..44 (featuresRest TRANSPOSED to [3,15] then flattened),
    opacity (logit), scale_0..2 (log), rot_0..3 (w x y z).
  * TSDF scene: the Scene/ directory of ITMBasicEngine::SaveToFile (reference InfiniTAM/ITMLib/Core/ITMBasicEngine.tpp:119-135;
    Objects/Scene/ITMVoxelBlockHash.h:129-155, ITMLocalVBA.h:36-66; ORUtils/MemoryBlockPersister.h:188-207): every *.dat file is a
    size_t element count followed by the raw array -- hash.dat (ITMHashEntry[1179648]), excess.dat (int[131072]), voxel.dat
    (ITMVoxel_s_rgb[blocks*512]), alloc.dat (int[blocks]); last.txt holds lastFreeExcessListId, vba.txt "lastFreeBlockId allocatedSize".
    The engine never returns blocks to the free lists (no swapping, as in the SLAM configuration), so both lists stay the identity
    permutation the reference initialises them with, and the files are byte-identical to the reference's (tests/test_checkpoint_gpu.py).

  * Gaussians: the model.pt archive of RawGaussianParams::saveTensor / loadTensor (src/raw_gs_param.cpp:220-254), through libtorch.

Pure numpy except model.pt; the device side is gsb_tsdf_read / gsb_tsdf_counter / gsb_tsdf_load_scene and gsb_gs_get_params / gsb_gs_set_params."""
import os

import numpy as np

from . import engine as E

PLY_PROPS = (["x", "y", "z", "nx", "ny", "nz"] + ["f_dc_%d" % i for i in range(3)] + ["f_rest_%d" % i for i in range(45)] + ["opacity"] +
             ["scale_%d" % i for i in range(3)] + ["rot_%d" % i for i in range(4)])


def save_ply(path, params):
    """params: dict of the six parameter arrays (means [N,3], scales log [N,3], quats wxyz [N,4], featuresDc [N,3], featuresRest [N,15,3],
    opacities logit [N,1]) as GaussianEngine.get_params() returns them"""
    n = params["means"].shape[0]
    rest = np.asarray(params["featuresRest"], np.float32).reshape(n, 15, 3).transpose(0, 2, 1).reshape(n, 45)
    rows = np.concatenate([np.asarray(params["means"], np.float32).reshape(n, 3), np.zeros((n, 3), np.float32),
                           np.asarray(params["featuresDc"], np.float32).reshape(n, 3), rest,
                           np.asarray(params["opacities"], np.float32).reshape(n, 1), np.asarray(params["scales"], np.float32).reshape(n, 3),
                           np.asarray(params["quats"], np.float32).reshape(n, 4)], 1).astype("<f4")
    with open(path, "wb") as f:
        f.write(("ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % n).encode())
        for p in PLY_PROPS:
            f.write(("property float %s\n" % p).encode())
        f.write(b"end_header\n")
        f.write(rows.tobytes())


def load_ply(path):
    with open(path, "rb") as f:
        data = f.read()
    end = data.index(b"end_header\n") + len(b"end_header\n")
    header = data[:end].decode().split("\n")
    if header[0] != "ply" or "binary_little_endian" not in header[1]:
        raise ValueError("%s is not a binary little-endian ply" % path)
    n = int([h for h in header if h.startswith("element vertex")][0].split()[2])
    props = [h.split()[2] for h in header if h.startswith("property float")]
    if props != PLY_PROPS:
        raise ValueError("%s does not have the reference's 3DGS property list" % path)
    rows = np.frombuffer(data, "<f4", n * len(props), end).reshape(n, len(props))
    return dict(means=rows[:, 0:3].copy(), featuresDc=rows[:, 6:9].copy(),
                featuresRest=rows[:, 9:54].reshape(n, 3, 15).transpose(0, 2, 1).copy(), opacities=rows[:, 54:55].copy(),
                scales=rows[:, 55:58].copy(), quats=rows[:, 58:62].copy())


MODEL_PT_KEYS = ("means", "scales", "quats", "featuresDc", "featuresRest", "opacities")


def _torch_shim():
    import torch
    from . import build
    torch.ops.load_library(build.build_torch_shim())
    return torch.ops.gsplat_b200


def save_model_pt(path, params, exposure=None, device=None):
    """RawGaussianParams::saveTensor (reference src/raw_gs_param.cpp:220-238): the model.pt archive (torch::serialize::OutputArchive, keys
    means / scales / quats / featuresDc / featuresRest / opacities / exposure), written through the same libtorch API by the C++ host
    layer (cxx/gsplat_b200_ops.cpp).  params: dict of arrays or tensors in the reference's shapes; exposure defaults to the reference's
    initial value, one 3x4 identity (src/raw_gs_param.cpp:61-65)."""
    import torch
    def t(a, shape):
        x = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a, np.float32))
        x = x.detach().to(torch.float32).reshape(shape).contiguous()
        return x.to(device) if device is not None else x
    n = int(np.prod(np.shape(params["means"]))) // 3
    shapes = dict(means=(n, 3), scales=(n, 3), quats=(n, 4), featuresDc=(n, 3), featuresRest=(n, 15, 3), opacities=(n, 1))
    tensors = [t(params[k], shapes[k]) for k in MODEL_PT_KEYS]
    tensors.append(t(exposure, (-1, 3, 4)) if exposure is not None else t(torch.eye(3, 4)[None], (1, 3, 4)))
    _torch_shim().save_model_pt(path, tensors)


def load_model_pt(path):
    """RawGaussianParams::loadTensor (reference src/raw_gs_param.cpp:240-254) -> (dict of numpy arrays, exposure)"""
    tensors = _torch_shim().load_model_pt(path)
    out = {k: tensors[i].detach().cpu().numpy() for i, k in enumerate(MODEL_PT_KEYS)}
    return out, tensors[6].detach().cpu().numpy()


def _write_block(path, arr):
    with open(path, "wb") as f:
        f.write(np.uint64(arr.size).tobytes())
        f.write(np.ascontiguousarray(arr).tobytes())


def _read_block(path, dtype):
    with open(path, "rb") as f:
        n = int(np.frombuffer(f.read(8), np.uint64)[0])
        a = np.frombuffer(f.read(), dtype)
    if a.size != n:
        raise ValueError("%s: header says %d elements, file holds %d" % (path, n, a.size))
    return a


def save_scene(directory, tsdf):
    """ITMBasicEngine::SaveToFile for a TsdfEngine: writes <directory>/Scene/*"""
    scene = os.path.join(directory, "Scene")
    os.makedirs(scene, exist_ok=True)
    os.makedirs(os.path.join(directory, "Relocaliser"), exist_ok=True)
    last_block, last_excess = tsdf.counter(0), tsdf.counter(1)
    _write_block(os.path.join(scene, "voxel.dat"), tsdf.voxels().reshape(-1))
    _write_block(os.path.join(scene, "alloc.dat"), np.arange(tsdf.num_blocks, dtype=np.int32))
    with open(os.path.join(scene, "vba.txt"), "w") as f:
        f.write("%d %d" % (last_block, tsdf.num_blocks * 512))
    with open(os.path.join(scene, "last.txt"), "w") as f:
        f.write("%d" % last_excess)
    _write_block(os.path.join(scene, "hash.dat"), tsdf.hash_entries())
    _write_block(os.path.join(scene, "excess.dat"), np.arange(0x20000, dtype=np.int32))


def load_scene(directory, tsdf):
    """ITMBasicEngine::LoadFromFile for a TsdfEngine"""
    scene = os.path.join(directory, "Scene")
    hash_entries = _read_block(os.path.join(scene, "hash.dat"), E.HASH_ENTRY)
    voxels = _read_block(os.path.join(scene, "voxel.dat"), E.VOXEL)
    alloc = _read_block(os.path.join(scene, "alloc.dat"), np.int32)
    excess = _read_block(os.path.join(scene, "excess.dat"), np.int32)
    if not (np.array_equal(alloc, np.arange(alloc.size)) and np.array_equal(excess, np.arange(excess.size))):
        raise E.EngineError("free lists are not the identity permutation (scene saved by an engine with swapping): not supported")
    with open(os.path.join(scene, "vba.txt")) as f:
        last_block = int(f.read().split()[0])
    with open(os.path.join(scene, "last.txt")) as f:
        last_excess = int(f.read().split()[0])
    tsdf.load_scene(hash_entries, voxels, last_block, last_excess)


def save_mesh_ply(path, tsdf):
    """ITMBasicEngine::SaveSceneToMesh (reference InfiniTAM/ITMLib/Core/ITMBasicEngine.tpp:105-117): marching cubes on the device
    (gsb_tsdf_mesh) + ITMMesh::WritePLY's ASCII layout (Objects/Meshing/ITMMesh.h:39-104): three vertices per triangle
    "%f %f %f r g b" with colours truncated to uchar, then one "3 i j k" face per triangle.  Returns the triangle count."""
    tri = tsdf.mesh().cpu().numpy()
    n = len(tri)
    pos = tri[:, :9].reshape(n * 3, 3)
    col = (tri[:, 9:].reshape(n * 3, 3) * np.float32(255.0)).astype(np.int32).astype(np.uint8)   # static_cast<unsigned char>(c * 255)
    with open(path, "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex %d\n" % (n * 3))
        f.write("property float x\nproperty float y\nproperty float z\nproperty uchar red\nproperty uchar green\nproperty uchar blue\n")
        f.write("element face %d\nproperty list uchar int vertex_indices\nend_header\n" % n)
        f.write("".join("%f %f %f %d %d %d\n" % (p[0], p[1], p[2], c[0], c[1], c[2]) for p, c in zip(pos, col)))
        f.write("".join("3 %d %d %d\n" % (3 * i, 3 * i + 1, 3 * i + 2) for i in range(n)))
    return n
