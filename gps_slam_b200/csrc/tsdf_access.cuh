// Voxel addressing of the hashed TSDF scene, shared by the raycast (tsdf_kernels.cu) and the mesh export (tsdf_mesh.cu).
#pragma once
#include "common.cuh"
#include "tsdf.h"

namespace tsdf
{

// Voxel addressing.  A voxel is named by a 32-bit handle = index into the voxel block array (28 bits: up to 2^19 blocks) | owner rank << 28;
// NO_VOXEL = not allocated.  VbaLocal: one array (single GPU, owner bits always 0).  VbaSharded: one array per rank, mapped over NVLink
// (the hash table is replicated, a block's voxels live on rank hashIndex(blockPos) mod world only).
constexpr unsigned NO_VOXEL = 0xffffffffu;
struct VbaLocal
{
    const Voxel *vba;
    __device__ __forceinline__ unsigned block_handle(int, int, int, int ptr) const { return (unsigned)ptr * SDF_BLOCK_SIZE3; }
    __device__ __forceinline__ const Voxel *at(unsigned h) const { return vba + h; }
};
struct VbaSharded
{
    const ShardView *v;
    __device__ __forceinline__ unsigned block_handle(int bx, int by, int bz, int ptr) const
    {
        return (unsigned)ptr * SDF_BLOCK_SIZE3 | ((unsigned)block_owner(bx, by, bz, v->world) << 28);
    }
    __device__ __forceinline__ const Voxel *at(unsigned h) const { return v->vba[h >> 28] + (h & 0x0fffffffu); }
};

struct VoxelCache
{
    int bx, by, bz;
    unsigned block; // handle of the cached block's first voxel
};

// returns the voxel's handle; vm = 0 not found, 1 cache hit, slot+1 hash hit
template <class A>
__device__ __forceinline__ unsigned find_voxel(const A &vba, const HashEntry *__restrict__ table, int px, int py, int pz, int &vm, VoxelCache &c)
{
    int bx = ((px < 0) ? px - SDF_BLOCK_SIZE + 1 : px) / SDF_BLOCK_SIZE;
    int by = ((py < 0) ? py - SDF_BLOCK_SIZE + 1 : py) / SDF_BLOCK_SIZE;
    int bz = ((pz < 0) ? pz - SDF_BLOCK_SIZE + 1 : pz) / SDF_BLOCK_SIZE;
    int lin = px + (py - bx) * SDF_BLOCK_SIZE + (pz - by) * SDF_BLOCK_SIZE * SDF_BLOCK_SIZE - bz * SDF_BLOCK_SIZE3;
    if (bx == c.bx && by == c.by && bz == c.bz)
    {
        vm = 1;
        return c.block + (unsigned)lin;
    }
    int idx = hash_index(bx, by, bz);
    while (true)
    {
        HashEntry e = load_entry(table, idx);
        if (e.px == bx && e.py == by && e.pz == bz && e.ptr >= 0)
        {
            c.bx = bx, c.by = by, c.bz = bz;
            c.block = vba.block_handle(bx, by, bz, e.ptr);
            vm = idx + 1;
            return c.block + (unsigned)lin;
        }
        if (e.offset < 1)
            break;
        idx = SDF_BUCKET_NUM + e.offset - 1;
    }
    vm = 0;
    return NO_VOXEL;
}

} // namespace tsdf
