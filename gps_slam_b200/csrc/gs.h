// Internal interface of the gsplat-GES path: kernels in gs_project.cu (built with
// -fmad=false: projection / SH / binning / parameter backward + Adam / prune / spawn) and gs_raster.cu (rasteriser
// forward / backward), model object and C ABI in gs_engine.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gs_comm.h"
#include "gs_math.cuh"

namespace gs
{

// Packed per-splat record produced by the projection pass and consumed by binning and both rasteriser passes.
// 48 B = 3 x float4, so one splat is three 16-byte transactions wherever it is gathered.
//   q0 = (mean2d.x, mean2d.y, opacity, int_as_float(radius))      radius = 0 -> culled (rest of the record undefined)
//   q1 = (conic.a,  conic.b,  conic.c, camera depth)
//   q2 = (r, g, b, int_as_float(bit c set <=> raw SH colour c + 0.5 >= 0  |  bit 8 set <=> more than one backward item))
struct __align__(16) SplatRec
{
    float4 q0, q1, q2;
};

// Raster-side gradient per splat
//   g0 = (v_mean2d.x, v_mean2d.y, v_opacity, v_depth)   g1 = (v_conic.a, v_conic.b, v_conic.c, 0)   g2 = (v_r, v_g, v_b, 0)
struct __align__(16) SplatGrad
{
    float4 g0, g1, g2;
};

struct ParamPtrs
{
    float *means;  // [N,3]
    float *scales; // [N,3] log
    float *quats;  // [N,4] w x y z
    float *dc;     // [N,3]
    float *rest;   // [N,15,3]
    float *opac;   // [N] logit
};

struct AdamStep
{
    AdamScalars a;
    float step_size[6]; // lr / bias_correction1 for means, scales, quats, dc, rest, opac
};

// counters[] layout
enum
{
    CNT_ISECTS = 0,   // n_isects of the current binning (clamped to the capacity)
    CNT_ITEMS = 1,    // number of backward work items of the current projection
    CNT_OVERFLOW = 2, // bit 0: intersections exceeded isectCap, bit 1: backward items exceeded itemCap, bit 2: Gaussian capacity
    CNT_VISIBLE = 3,  // running count of Gaussians with radius > 0 (moved to CNT_VISIBLE_LAST by the binning pass)
    CNT_VISIBLE_LAST = 4,
    CNT_SCRATCH = 5,  // prune / spawn scratch (2 ints)
    CNT_BWD_CURSOR = 7, // work-item cursor of the rasteriser backward
    CNT_PAIRS_TESTED = 8,  // 64-bit (2 ints): (pixel, splat) pairs the statistics build of the rasteriser backward evaluated ...
    CNT_PAIRS_PASSED = 10, // 64-bit: ... and how many of them passed the alpha / depth tests (gsb_gs_run_stage 6)
    CNT_TOTAL = 16
};

struct Bins
{
    int *segCount;      // [T*BIN_CHUNKS] intersections per (tile, id-range chunk); doubles as the scatter cursor; zero between iterations
    int *segOff;        // [T*BIN_CHUNKS] offset of the segment inside its tile's list
    int *tileCount;     // [T+1] per-tile totals
    int *tileOffsets;   // [T+1] exclusive scan; [T] = n_isects
    int *flatten;       // [isectCap] scatter order (arbitrary within a tile)
    int *flattenSorted; // [isectCap] ascending Gaussian id within each tile  == reference flatten_ids
    int isectCap;
    int4 *items;        // [itemCap] (gaussian id, first rect-linear pixel, x0 | y0 << 16, w | h << 16) of <= BWD_PIXELS_PER_ITEM pixels
    int itemCap;
    int *counters;      // [CNT_TOTAL]
};

enum RasterMode
{
    RASTER_RAW = 0,    // render_colors [H,W,4] + alphas [H,W]                 (gsplat::rasterize_to_pixels_fwd_ges_tensor)
    RASTER_RENDER = 1, // composited rgb [H,W,3], depth [H,W], alpha [H,W]      (RawGaussianModel::gesForward outputs)
    RASTER_TRAIN = 2,  // composite + L1 loss + dL/d(render) packed as float4   (gesForward + computeLoss + backward head)
    RASTER_PUSH = 3    // multi-GPU: partial sums stored straight into the gather slots of peer ranks (gs_comm.h)
};

struct RasterIO
{
    const float *refDepth;  // [H,W] TSDF raycast depth (raw; < 0.01 means "no hit")
    const float *baseColor; // [H,W,3] TSDF raycast colour
    const float *gt;        // [H,W,3] TRAIN
    float deltaDepth;
    int clampRef;           // 1: refDepth < 0.01 -> 1000 inside the kernel (src/raw_gs_model.cpp:207); 0: use as given
    float *render4;         // RAW
    float *alphas;          // RAW / RENDER
    float *rgb;             // RENDER
    float *depth;           // RENDER
    float4 *v_out;          // [2 P]: per pixel (v_render_r, v_render_g, v_render_b, v_render_alpha | depth cut, -, -, -); the first half is
                            // written in TRAIN mode, the depth-test threshold (clamped refDepth + deltaDepth) in RAW / TRAIN mode
    float *lossTile;        // TRAIN: per-tile sum |rgb - gt|
};

#ifdef __CUDACC__
// Conservative pixel-space half extents of the region where a splat can pass the alpha test:
// alpha = opacity * exp(-sigma) >= 1/255  <=>  sigma <= tau = ln(255 * opacity); {sigma <= tau} is an ellipse whose axis-aligned
// half extents are sqrt(2 tau c / det) and sqrt(2 tau a / det).  Inflated for rounding; returns false when nothing can pass.
__device__ __forceinline__ bool alpha_extent(float a, float b, float c, float opac, float &ex, float &ey)
{
    float tau = logf(255.0f * opac) * 1.0005f + 1e-3f;
    if (!(tau > 0.f))
        return false;
    float det = a * c - b * b;
    if (!(det > 0.f) || !(a > 0.f) || !(c > 0.f))
    {
        ex = ey = 1e30f; // degenerate conic: no bound
        return true;
    }
    ex = sqrtf(2.0f * tau * c / det) * 1.001f + 0.02f;
    ey = sqrtf(2.0f * tau * a / det) * 1.001f + 0.02f;
    return true;
}

// Pixel rectangle the rasteriser backward visits for one splat: the reference's box j in [int(x)-r+1, int(x)+r],
// i in [int(y)-r+1, int(y)+r] (rasterize_to_pixels_bwd_ges_new_parallel.cu:76-96) clipped to the image and to the alpha extent --
// exactly the pixels that can pass the reference's validity tests.  Returns the pixel count (0: nothing to do).
__device__ __forceinline__ int bwd_rect(float mx, float my, int radius, float a, float b, float c, float opac, int W, int H, int &x0, int &y0,
                                        int &w, int &h)
{
    float ex, ey;
    x0 = y0 = w = h = 0;
    if (!alpha_extent(a, b, c, opac, ex, ey))
        return 0;
    int xl = max((int)mx - radius + 1, 0), xh = min((int)mx + radius, W - 1);
    int yl = max((int)my - radius + 1, 0), yh = min((int)my + radius, H - 1);
    // pixel centre j + 0.5 must lie within mx +- ex
    float fxl = ceilf(mx - ex - 0.5f), fxh = floorf(mx + ex - 0.5f), fyl = ceilf(my - ey - 0.5f), fyh = floorf(my + ey - 0.5f);
    if (fxl > (float)xl)
        xl = (int)fminf(fxl, 1e9f);
    if (fxh < (float)xh)
        xh = (int)fmaxf(fxh, -1e9f);
    if (fyl > (float)yl)
        yl = (int)fminf(fyl, 1e9f);
    if (fyh < (float)yh)
        yh = (int)fmaxf(fyh, -1e9f);
    if (xl > xh || yl > yh)
        return 0;
    x0 = xl, y0 = yl, w = xh - xl + 1, h = yh - yl + 1;
    return w * h;
}
#endif

constexpr int TILE = 16;
constexpr int BIN_CHUNKS = 64; // id-range chunks of the counting sort by (tile, chunk)
inline int bin_chunk_size(int nUpper) { return nUpper > BIN_CHUNKS ? (nUpper + BIN_CHUNKS - 1) / BIN_CHUNKS : 1; }
constexpr int BWD_PIXELS_PER_ITEM = 2048;
constexpr int PARAMS_PER_GAUSSIAN = 59;

// ---- gs_project.cu
void project_sh_fwd(const ParamPtrs &p, const int *nDev, int nUpper, const CamParams &cam, SplatRec *recs, SplatGrad *grads, const Bins &bins,
                    int tileW, int tileH, bool forBackward, cudaStream_t st);
void bin_tiles(const SplatRec *recs, const int *nDev, int nUpper, const Bins &bins, int tileW, int tileH, cudaStream_t st);
// dbg: optional dump of the parameter gradients (same layout as the parameters); nullptr in production
void bwd_params_adam(const ParamPtrs &p, const ParamPtrs &m, const ParamPtrs &v, unsigned char *touched, const AdamStep &step, const int *nDev,
                     int nUpper, const CamParams &cam, const SplatRec *recs, const SplatGrad *grads, float4 *aux /* [N*5] */,
                     const ParamPtrs *dbg, int *counters, cudaStream_t st);
void reduce_loss(const float *lossTile, int T, double scale, double *out, cudaStream_t st);
// removeRedundantGs + prunePoints + removeFromOptimizer: stable compaction of the parameters and, for Gaussians that carry optimiser
// state, of their Adam moments and state flag, into a second set of buffers (the caller swaps the sets afterwards)
struct PruneBuffers
{
    ParamPtrs p, m, v, pOut, mOut, vOut;
    const unsigned char *touched;
    unsigned char *touchedOut;
};
void prune(const PruneBuffers &b, int *nDev, int nUpper, float minOpac, float minScale, float maxScale, int *scanTmp, int *counters, cudaStream_t st);

// staged pieces with the argument layout of the reference's gsplat::*_tensor functions (C = 1)
void staged_project_fwd(int N, const float *means, const float *quats, const float *scales, const CamParams &cam, int *radii, float *means2d,
                        float *depths, float *conics, cudaStream_t st);
void staged_project_bwd(int N, const float *means, const float *quats, const float *scales, const CamParams &cam, const int *radii,
                        const float *conics, const float *v_means2d, const float *v_depths, const float *v_conics, float *v_means, float *v_quats,
                        float *v_scales, cudaStream_t st);
void staged_sh_fwd(int N, const float *dirs, const float *coeffs, const unsigned char *mask, float *colors, cudaStream_t st);
void staged_sh_bwd(int N, const float *dirs, const float *coeffs, const unsigned char *mask, const float *v_colors, float *v_coeffs, float *v_dirs,
                   cudaStream_t st);
void staged_pack(int N, const float *means2d, const float *conics, const float *colors4, const float *depths, const float *opac, const int *radii,
                 SplatRec *recs, SplatGrad *grads, const Bins &bins, int tileW, int tileH, int W, int H, int *tilesPerGauss, bool countTiles,
                 bool forBackward, cudaStream_t st);
void staged_cut(int P, const float *refDepth, float delta, float4 *v_out, cudaStream_t st);
void staged_isect_ids(const Bins &bins, int T, long long *isectIds, cudaStream_t st);
void staged_unpack_grads(int N, const SplatRec *recs, const SplatGrad *grads, float *v_means2d, float *v_conics, float *v_colors4, float *v_opac,
                         cudaStream_t st);
void staged_adam(int n, float *p, const float *g, float *m, float *v, const AdamScalars &a, float step_size, cudaStream_t st);

// ---- gs_raster.cu
void raster_fwd(int mode, const SplatRec *recs, const Bins &bins, int W, int H, int tileW, int tileH, const RasterIO &io, cudaStream_t st);
void raster_bwd(const SplatRec *recs, const Bins &bins, int W, int H, const RasterIO &io, const float *v_depth /* nullable [H,W] */,
                SplatGrad *grads, cudaStream_t st);
void raster_bwd_stats(const SplatRec *recs, const Bins &bins, int W, int H, const RasterIO &io, SplatGrad *grads, cudaStream_t st);
void composite(int mode, const float *acc5 /* [P*4] render then [P] alphas */, int W, int H, int tileW, int tileH, const RasterIO &io, cudaStream_t st);
void raster_fwd_push(const SplatRec *recs, const Bins &bins, int W, int H, int tileW, int tileH, const RasterIO &io, const CommView *cvDev, bool pushAll,
                     cudaStream_t st);
void composite_exchange(int mode, const CommView &cv, const CommView *cvDev, int W, int H, int tileW, int tileH, const RasterIO &io, cudaStream_t st);
void pack_v_out(int P, const float *v_render4, const float *v_alphas, float4 *v_out, float *v_depth, cudaStream_t st);

// ---- gs_raw.cu: depth-sorted front-to-back compositing (render_method "raw")
void sort_tiles_depth(const SplatRec *recs, const Bins &bins, int T, cudaStream_t st);
void isect_ids_depth(const SplatRec *recs, const Bins &bins, int T, long long *isectIds, cudaStream_t st);
void raw_fwd(const SplatRec *recs, const Bins &bins, int W, int H, int tileW, int tileH, const float *background, float *render4, float *alphas,
             int *lastIds, cudaStream_t st);
void raw_bwd(int N, const SplatRec *recs, const Bins &bins, int W, int H, int tileW, int tileH, const float *background, const float *alphas,
             const int *lastIds, const float *v_render4, const float *v_alphas, SplatGrad *grads, cudaStream_t st);
// ---- gs_ssim.cu
void ssim_fwd(int planes, int H, int W, float C1, float C2, const float *img1, const float *img2, float *ssimMap, float *dm_dmu1,
              float *dm_dsigma1_sq, float *dm_dsigma12, cudaStream_t st);
void ssim_bwd(int planes, int H, int W, const float *img1, const float *img2, const float *dL_dmap, const float *dm_dmu1,
              const float *dm_dsigma1_sq, const float *dm_dsigma12, float *dL_dimg1, cudaStream_t st);


// ---- gs_knn.cu: distCUDA2 (mean squared distance to the 3 nearest other points), exact, grid based
constexpr int KNN_MAX_CELLS = 1 << 21;
size_t knn_workspace_bytes(int maxPoints);
void knn_mean_dist3(int P, const float *points, float *meanDist, void *workspace, cudaStream_t st);

} // namespace gs
