// Internal interface of the gsplat-GES path: kernels in gs_project.cu (built with
// -fmad=false: projection / SH / binning / parameter backward + Adam) and gs_raster.cu (rasteriser forward / backward),
// model object and C ABI in gs_engine.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gs_math.cuh"

namespace gs
{

// Packed per-splat record produced by the projection pass and consumed by binning and both rasteriser passes.
// 48 B = 3 x float4, so one splat is three 16-byte transactions wherever it is gathered.
//   q0 = (mean2d.x, mean2d.y, opacity, int_as_float(radius))      radius = 0 -> culled
//   q1 = (conic.a,  conic.b,  conic.c, camera depth)
//   q2 = (r, g, b, int_as_float(bit c set <=> raw SH colour c + 0.5 >= 0))
struct __align__(16) SplatRec
{
    float4 q0, q1, q2;
};

// Raster-side gradient per splat (zeroed by the projection pass, accumulated by the raster backward)
//   g0 = (v_mean2d.x, v_mean2d.y, v_opacity, v_depth)   g1 = (v_conic.a, v_conic.b, v_conic.c, 0)   g2 = (v_r, v_g, v_b, 0)
struct __align__(16) SplatGrad
{
    float4 g0, g1, g2;
};

struct ParamPtrs
{
    float *means;     // [N,3]
    float *scales;    // [N,3] log
    float *quats;     // [N,4] w x y z
    float *dc;        // [N,3]
    float *rest;      // [N,15,3]
    float *opac;      // [N,1] logit
};

struct AdamState
{
    ParamPtrs m, v;
    unsigned char *touched; // [N] 1 once a Gaussian received a gradient in this optimiser cycle
};

struct AdamStep
{
    AdamScalars a;
    float step_size[6]; // lr / bias_correction1 for means, scales, quats, dc, rest, opac
};

struct BinBuffers
{
    int *tileCount;    // [T+1]
    int *tileOffsets;  // [T+1] exclusive scan; [T] = n_isects
    int *tileCursor;   // [T]
    int *flattenIds;   // [isectCap]
    int isectCap;
    int *counters;     // [8]: 0 n_isects, 1 n_bwd_items, 2 overflow flag, 3 n_visible
};

struct BwdItems
{
    int2 *items;       // (gaussian id, chunk index)
    int cap;
};

enum RasterMode
{
    RASTER_RAW = 0,    // write render_colors [H,W,4] + alphas [H,W]           (gsplat::rasterize_to_pixels_fwd_ges_tensor)
    RASTER_RENDER = 1, // write composited rgb [H,W,3], depth [H,W], alpha     (RawGaussianModel::gesForward outputs)
    RASTER_TRAIN = 2   // composite + L1 loss + dL/d(render) packed as float4   (gesForward + computeLoss + backward head)
};

struct RasterOut
{
    float *render4;   // RAW
    float *alphas;    // RAW / RENDER
    float *rgb;       // RENDER
    float *depth;     // RENDER
    float4 *v_out;    // TRAIN: (v_r, v_g, v_b, v_alpha)
    double *loss;     // TRAIN: sum |rgb - gt| (divide by 3P)
};

constexpr int TILE = 16;
constexpr int BWD_GROUPS_PER_ITEM = 64;

// ---- gs_project.cu
void project_sh_fwd(const ParamPtrs &p, int N, const CamParams &cam, SplatRec *recs, SplatGrad *grads, const BinBuffers &bins, int tileW, int tileH,
                    const BwdItems &items, bool forBackward, cudaStream_t st);
void bin_tiles(const SplatRec *recs, int N, const BinBuffers &bins, int tileW, int tileH, cudaStream_t st);
void bwd_params_adam(const ParamPtrs &p, const AdamState &s, const AdamStep &step, int N, const CamParams &cam, const SplatRec *recs,
                     const SplatGrad *grads, cudaStream_t st);
// staged pieces (gsplat::*_tensor shaped)
void staged_project_fwd(int N, const float *means, const float *quats, const float *scales, const CamParams &cam, int *radii, float *means2d,
                        float *depths, float *conics, cudaStream_t st);
void staged_sh_fwd(int N, const float *dirs, const float *coeffs, const unsigned char *mask, float *colors, cudaStream_t st);
void staged_pack(int N, const float *means2d, const float *conics, const float *colors4, const float *opac, const int *radii, SplatRec *recs,
                 cudaStream_t st);
void staged_count_tiles(const SplatRec *recs, int N, const BinBuffers &bins, int tileW, int tileH, int *tilesPerGauss, cudaStream_t st);
void staged_fill_isect_ids(const BinBuffers &bins, int T, long long *isectIds, cudaStream_t st);
void staged_offset_encode(const long long *isectIds, int nIsects, int T, int *offsets, cudaStream_t st);
void staged_sh_bwd(int N, const float *dirs, const float *coeffs, const unsigned char *mask, const float *v_colors, float *v_coeffs, float *v_dirs,
                   cudaStream_t st);
void staged_project_bwd(int N, const float *means, const float *quats, const float *scales, const CamParams &cam, const int *radii,
                        const float *conics, const float *v_means2d, const float *v_depths, const float *v_conics, float *v_means, float *v_quats,
                        float *v_scales, cudaStream_t st);
void staged_adam(int n, float *p, const float *g, float *m, float *v, const AdamScalars &a, float step_size, cudaStream_t st);
void staged_unpack_grads(int N, const SplatGrad *grads, float *v_means2d, float *v_conics, float *v_colors4, float *v_opac, cudaStream_t st);
void build_bwd_items(const SplatRec *recs, int N, const BinBuffers &bins, const BwdItems &items, SplatGrad *grads, cudaStream_t st);

// ---- gs_raster.cu
void raster_fwd(int mode, const SplatRec *recs, const BinBuffers &bins, int W, int H, int tileW, int tileH, const float *refDepth, bool clampRef,
                const float *baseColor, const float *gt, float deltaDepth, const RasterOut &out, cudaStream_t st);
void raster_bwd(const SplatRec *recs, const BinBuffers &bins, const BwdItems &items, int W, int H, const float *refDepth, bool clampRef,
                float deltaDepth, const float4 *v_out, const float *v_depth /* nullable */, SplatGrad *grads, cudaStream_t st);
void pack_v_out(int P, const float *v_render4, const float *v_alphas, float4 *v_out, float *v_depth, cudaStream_t st);

} // namespace gs
