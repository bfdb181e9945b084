// Internal C++ interface between the TSDF kernels (tsdf_kernels.cu), the ICP kernels (icp_kernels.cu) and the
// engine / C-ABI layer (engine.cu).  Not installed; the public boundary is include/gpsslam_b200.h.
#pragma once
#include "common.cuh"

namespace tsdf
{

// device-resident scene = ITMScene<ITMVoxel_s_rgb, ITMVoxelBlockHash> + ITMRenderState_VH bookkeeping
struct Scene
{
    HashEntry *table;        // [E]
    Voxel *vba;              // [numBlocks * 512]
    unsigned *allocKey;      // [E]   per-slot allocation request (0 = none), replaces entriesAllocType + blockCoords
    unsigned char *visType;  // [E]   entriesVisibleType
    int *visIds;             // [numBlocks]  visibleEntryIDs (ascending)
    int2 *chunkCounts;       // [ceil(E/1024)]
    int *state;              // [8]: 0 lastFreeBlockId, 1 lastFreeExcessListId, 2 noVisibleEntries, 3 error flag, 4-5 staging
    int E, numBlocks;
    float voxelSize, mu, vfmin, vfmax;
    int maxW;
    // voxel-hash sharding: hash table / visibility / allocation are replicated (deterministic, every rank computes
    // the same), the voxel DATA of a block lives on one rank only: owner = hashIndex(blockPos) mod world.
    int rank, world;
    int *visIdsOwn;          // [numBlocks]  the visible entries this rank owns (ascending); == visIds for world == 1
    // shard mode 1 ("owner computes, everybody stores"): the integrate kernel writes every block it has updated into the other ranks'
    // voxel block arrays as well (TMA bulk stores over NVLink), so every rank keeps a complete copy and raycasts read local memory
    int nPush;
    Voxel *pushVba[15];
};

constexpr int SHARD_MAX_WORLD = 16;
struct PeerVbas
{
    int n;
    Voxel *p[SHARD_MAX_WORLD - 1];
};

// what the sharded kernels see of the other ranks (peer-mapped device pointers, index = rank; own entry = local buffer)
struct ShardView
{
    int rank, world;
    int replicated;          // shard mode 1: voxel reads are local (vba[rank] is complete)
    int probe;               // measurement aid (gsb_tsdf_shard_probe): 1 = results stay local (no peer stores), 2 = per-thread peer stores
    const Voxel *vba[SHARD_MAX_WORLD];
    unsigned char *visType[SHARD_MAX_WORLD];
    float4 *rayLive[SHARD_MAX_WORLD];
    float4 *rayFree[SHARD_MAX_WORLD];
    uchar4 *imageFree[SHARD_MAX_WORLD];
    float4 *pointsMap[SHARD_MAX_WORLD];
    float4 *normalsMap[SHARD_MAX_WORLD];
    unsigned *flags[SHARD_MAX_WORLD];   // [SHARD_MAX_WORLD] barrier arrival epochs, one per peer
    float *icpXchg[SHARD_MAX_WORLD];    // [2][SHARD_MAX_WORLD][32] ICP partial sums (e3), double-buffered by sequence parity
};

__host__ __device__ inline int block_owner(int bx, int by, int bz, int world)
{
    return (int)((((unsigned)bx * 73856093u) ^ ((unsigned)by * 19349669u) ^ ((unsigned)bz * 83492791u)) & (unsigned)SDF_HASH_MASK) % world;
}
// rows [y0, y1) of the ICP maps a rank computes: contiguous slabs of whole 8-row tiles (the raycast itself is dealt in interleaved
// 8-row strips, strip t to rank t mod world)
__host__ __device__ inline void slab_rows(int H, int rank, int world, int &y0, int &y1)
{
    const int tiles = (H + 7) / 8;
    y0 = (int)((long long)tiles * rank / world) * 8;
    y1 = (int)((long long)tiles * (rank + 1) / world) * 8;
    if (y1 > H)
        y1 = H;
}

struct Frame
{
    const short *depth_mm;   // [H*W] raw depth (mm)
    const uchar4 *rgba;      // [H*W]
    float *depth_f;          // [H*W] metres, <= 0 -> -1   (written by allocate())
    int W, H;
};

struct Camera
{
    Mat4 M, invM;            // world->camera, camera->world (ORUtils column-major)
    float fx, fy, cx, cy;
};

void reset_scene(const Scene &s, cudaStream_t st);
void allocate(const Scene &s, const Frame &f, const Camera &cam, cudaStream_t st);
void integrate(const Scene &s, const Frame &f, const Camera &cam, int variant, cudaStream_t st);
void expected_depth_live(const Scene &s, const Camera &cam, int W, int H, float2 *minmax, cudaStream_t st);
void expected_depth_free(const Scene &s, const Camera &cam, int W, int H, float2 *minmax, cudaStream_t st);
void raycast_stats(const Scene &s, const Camera &cam, int W, int H, const float2 *minmax, unsigned long long *totals8, cudaStream_t st);
void raycast(const Scene &s, const Camera &cam, int W, int H, const float2 *minmax, float4 *pointsRay, uchar4 *colour, bool modifyVisible,
             cudaStream_t st);
void icp_maps(const Scene &s, const Camera &cam, int W, int H, const float4 *pointsRay, float4 *pointsMap, float4 *normalsMap, cudaStream_t st);
// sharded forms (world > 1): this rank's share of the rows, voxels read from their owner's memory over NVLink (mode 0) or from the local
// copy (mode 1); results are stored into every rank's images; the visibility marks of the live raycast go to every rank
void raycast_sharded(const Scene &s, const ShardView &v, const Camera &cam, int W, int H, const float2 *minmax, bool live, cudaStream_t st);
void icp_maps_sharded(const Scene &s, const ShardView &v, const Camera &cam, int W, int H, bool pushAll, cudaStream_t st);
void shard_barrier(const ShardView &v, unsigned epoch, int *errFlag, cudaStream_t st);
// marching-cubes export (tsdf_mesh.cu)
size_t mesh_scratch_bytes(const Scene &s);
int mesh_scene(const Scene &s, const ShardView *view, void *scratch, float *outDev, long long maxTri, long long *nTriHost, cudaStream_t st);

} // namespace tsdf
