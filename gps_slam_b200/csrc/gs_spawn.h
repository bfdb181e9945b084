// Internal interface of the spawn / glue kernels (gs_spawn.cu).
#pragma once
#include "gs.h"

namespace gs
{
struct SpawnParams
{
    int W, H, P;
    float voxelSize;
    float colorErrorThres;   // PIPE.color_error_thres
    float depthMin, depthMax; // vis_configs.depth_vis_min / depth_vis_max
    float alphaMax;          // vis_configs.alpha_vis_max
    unsigned ratioThreshold; // new_gs_sample_ratio * 2^32
    unsigned seed;
    float maxScale, minScale; // MODEL.max_init_scale / min_init_scale
    float defaultOpacity;    // MODEL.default_opacities
    int forceRender;         // 1: renderRgb / renderAlpha are always valid (caller-supplied render)
    int rank, world;         // multi-GPU: spawn only the Gaussians whose 4 cm block hashes to this rank (world <= 1: all)
};

struct SpawnBuffers
{
    unsigned char *flags; // [P]
    int *chunkCnt;        // [ceil(P/1024)]
    int *pixOf;           // [P] pixel of each new Gaussian
    int *pixAll;          // [P] multi-GPU: pixel of each new Gaussian of ANY rank (the KNN point set)
    unsigned long long *keys; // [tableMask+1]
    int *heads;           // [tableMask+1]
    int *next;            // [P]
    unsigned tableMask;
};

void raycast_maps(int P, const float4 *vertex4, const uchar4 *colour4, const float *w2cRowMajor, float voxelSize, float *depthMap, float *colorMap,
                  float *confMap, cudaStream_t st);
void frame_to_float(int P, const uchar4 *rgba, const short *depth_mm, float *rgb, float *depth, cudaStream_t st);
void spawn(const SpawnParams &sp, const SpawnBuffers &b, const float4 *vertex4, const float *depthMap, const float *colorMap, const float *gt,
           const float *renderRgb, const float *renderAlpha, const ParamPtrs &p, int *nDev, int cap, unsigned char *touched, int *counters,
           cudaStream_t st);
} // namespace gs
