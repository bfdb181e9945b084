"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink 5 / NVSwitch; gloo on CPU for the tests).

The reference is single-GPU (no collective call site anywhere); this is new design:
  * the Gaussian set is SHARDED by spatial block: owner = hash(floor(mean / 4 cm)) mod world, decided once at spawn time
    (gs_spawn.cu k_spawn_select) -- the same 20-bit multiplicative hash the TSDF uses for its voxel blocks, re-mixed;
  * because the GES blend is an order-independent sum every rank rasterises only its own Gaussians over the whole
    image; ONE all-reduce(sum) of the partial accumulation image [H, W, 5] fp32 per optimiser iteration (16.3 MB at 1200x680)
    gives every rank the exact full image; composite / loss / dL/d(render) are recomputed redundantly, backward and Adam are
    rank-local (parameters are sharded, never replicated, so there is no parameter-gradient all-reduce at all);
  * the TSDF map is replicated: every rank fuses every frame (deterministic kernels -> identical maps, no exchange).
Two exchange paths exist: "peer" (default) -- csrc/gs_comm.h: the rasteriser's epilogue stores every tile's partial sums straight into
the owner rank's memory over NVLink, the owner composites and stores dL/d(render) into every rank's memory, flag barriers through peer
memory, all below the C ABI -- and "nccl", the plain all-reduce of the [H,W,5] image between two C-ABI calls (kept for A/B timing).
"""
import numpy as np
import torch
import torch.distributed as dist

BLOCK_M = 0.04  # ownership granularity: 8 voxels x 5 mm, the TSDF voxel-block size


def _hash_u32(x):
    x = np.asarray(x, dtype=np.uint64) & 0xffffffff
    x ^= x >> 16
    x = (x * 0x7feb352d) & 0xffffffff
    x ^= x >> 15
    x = (x * 0x846ca68b) & 0xffffffff
    x ^= x >> 16
    return x


def owner_of(points, world):
    """rank that owns each 3-D point [N,3] (host mirror of the device rule in gs_spawn.cu)"""
    p = np.asarray(points, dtype=np.float32)
    b = np.floor(p * np.float32(1.0 / BLOCK_M)).astype(np.int64)
    h = ((b[:, 0] * 73856093) & 0xffffffff) ^ ((b[:, 1] * 19349669) & 0xffffffff) ^ ((b[:, 2] * 83492791) & 0xffffffff)
    return (_hash_u32(h) % np.uint64(max(world, 1))).astype(np.int64)


def voxel_block_owner(block_pos, world):
    """rank that owns the voxel data of each TSDF block [N,3] (integer block coordinates): hashIndex(blockPos) mod world, the
    reference's own hash (ITMRepresentationAccess.h:7-11) -- host mirror of tsdf::block_owner (csrc/tsdf.h)"""
    b = np.asarray(block_pos).astype(np.int64)
    h = ((b[:, 0] * 73856093) & 0xffffffff) ^ ((b[:, 1] * 19349669) & 0xffffffff) ^ ((b[:, 2] * 83492791) & 0xffffffff)
    return ((h & 0xfffff) % max(world, 1)).astype(np.int64)


def make_sharded_tsdf(intr, rank, world, device, **kw):
    """TsdfEngine whose voxel hash is sharded over `world` ranks (one process per GPU): the 64-byte CUDA IPC handles of the shared
    segments travel once through torch.distributed; afterwards the TSDF path uses peer memory only (no collective library)"""
    from . import engine as E
    eng = E.TsdfEngine(intr, device=device, rank=rank, world=world, **kw)
    if world > 1:
        handles = [None] * world
        dist.all_gather_object(handles, eng.export_handle())
        eng.attach(handles)
        dist.barrier()
    return eng


def shard_params(params, rank, world):
    """keep the rows of a parameter dict (means, scales, ...) owned by `rank`"""
    if world <= 1:
        return params
    keep = owner_of(params["means"], world) == rank
    return {k: np.asarray(v)[keep] for k, v in params.items()}


def make_peer_comm(device, rank, world, width, height):
    """PeerComm of this rank with every peer's exchange segment mapped: the 64-byte CUDA IPC handles travel once through
    torch.distributed (all_gather_object); afterwards no collective library is involved in the Gaussian path."""
    from . import engine as E
    comm = E.PeerComm(device, rank, world, width, height)
    if world > 1:
        handles = [None] * world
        dist.all_gather_object(handles, comm.export_handle())
        comm.attach(handles)
        dist.barrier()
    return comm


def allreduce_sum_(t):
    """in-place sum over all ranks (no-op for a single process); enqueued on the current CUDA stream for NCCL"""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t
