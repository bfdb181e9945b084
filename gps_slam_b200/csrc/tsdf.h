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
};

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

} // namespace tsdf
