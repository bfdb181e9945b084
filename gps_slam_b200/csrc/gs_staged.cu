// Staged entry points of the gsplat-GES path with the argument layout of the reference's gsplat::*_tensor functions (C = 1), so
// that each autograd wrapper of gsplat/gsplat_wapper.hpp can be re-pointed at this library one function at a time
// (INTEGRATION.md section 2).  They reuse the per-Gaussian device math (gs_math.cuh) and the rasteriser / binning kernels of the
// fused path; what is added here is only the glue between the reference's separate arrays and the packed records.
//
//   staged_project_fwd / _bwd   fully_fused_projection_{fwd,bwd}_tensor      gsplat/rasterizer/fully_fused_projection_{fwd,bwd}.cu
//   staged_sh_fwd / _bwd        compute_sh_{fwd,bwd}_tensor                  gsplat/rasterizer/compute_sh_{fwd,bwd}.cu
//   staged_pack + bin_tiles     isect_tiles_tensor_no_depth + isect_offset_encode_tensor_no_depth   isect_tiles_no_depth.cu:132-461
//   staged_pack + raster_fwd    rasterize_to_pixels_fwd_ges_tensor           rasterize_to_pixels_fwd_ges.cu:338-407
//   staged_pack + raster_bwd    rasterize_to_pixels_bwd_ges_gs_parallel_tensor   rasterize_to_pixels_bwd_ges_new_parallel.cu:304-385
//   staged_adam                 torch::optim::Adam::step on one tensor       (libtorch; hyper-parameters src/raw_gs_model.cpp:661-672)
#include "common.cuh"
#include "gs.h"

namespace gs
{

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// tile rectangle of a splat: isect_tiles_no_depth.cu:70-80 (the float -> uint32 casts saturate on the GPU)
__device__ __forceinline__ void tile_rect_s(float m2x, float m2y, int radius, int tileW, int tileH, int &x0, int &y0, int &x1, int &y1)
{
    float ts = (float)TILE;
    float tr = (float)radius / ts;
    float tx = m2x / ts, ty = m2y / ts;
    x0 = (int)min((unsigned)floorf(tx - tr), (unsigned)tileW);
    y0 = (int)min((unsigned)floorf(ty - tr), (unsigned)tileH);
    x1 = (int)min((unsigned)ceilf(tx + tr), (unsigned)tileW);
    y1 = (int)min((unsigned)ceilf(ty + tr), (unsigned)tileH);
}

__global__ void __launch_bounds__(256) k_staged_project_fwd(int N, const float *__restrict__ means, const float *__restrict__ quats,
                                                             const float *__restrict__ scales, CamParams cam, int *radii, float *means2d,
                                                             float *depths, float *conics)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= N)
        return;
    float mean[3] = {means[g * 3 + 0], means[g * 3 + 1], means[g * 3 + 2]};
    float quat[4] = {quats[g * 4 + 0], quats[g * 4 + 1], quats[g * 4 + 2], quats[g * 4 + 3]};
    float scale[3] = {scales[g * 3 + 0], scales[g * 3 + 1], scales[g * 3 + 2]};
    Proj o = project_one(mean, quat, scale, cam, nullptr);
    radii[g] = o.radius;
    // the reference leaves torch::empty garbage in the other outputs of a culled Gaussian; zeros here
    means2d[g * 2 + 0] = o.m2x, means2d[g * 2 + 1] = o.m2y;
    depths[g] = o.depth;
    conics[g * 3 + 0] = o.ca, conics[g * 3 + 1] = o.cb, conics[g * 3 + 2] = o.cc;
}

__global__ void __launch_bounds__(128) k_staged_project_bwd(int N, const float *__restrict__ means, const float *__restrict__ quats,
                                                             const float *__restrict__ scales, CamParams cam, const int *__restrict__ radii,
                                                             const float *__restrict__ conics, const float *__restrict__ v_means2d,
                                                             const float *__restrict__ v_depths, const float *__restrict__ v_conics,
                                                             float *v_means, float *v_quats, float *v_scales)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= N)
        return;
    float vm[3] = {0.f, 0.f, 0.f}, vq[4] = {0.f, 0.f, 0.f, 0.f}, vs[3] = {0.f, 0.f, 0.f};
    if (radii[g] > 0)
    {
        float mean[3] = {means[g * 3 + 0], means[g * 3 + 1], means[g * 3 + 2]};
        float quat[4] = {quats[g * 4 + 0], quats[g * 4 + 1], quats[g * 4 + 2], quats[g * 4 + 3]};
        float scale[3] = {scales[g * 3 + 0], scales[g * 3 + 1], scales[g * 3 + 2]};
        ProjIntermediates keep;
        project_one(mean, quat, scale, cam, &keep);
        project_vjp(scale, cam, keep, conics[g * 3 + 0], conics[g * 3 + 1], conics[g * 3 + 2], v_means2d[g * 2 + 0], v_means2d[g * 2 + 1],
                    v_depths[g], v_conics[g * 3 + 0], v_conics[g * 3 + 1], v_conics[g * 3 + 2], vm, vq, vs);
    }
#pragma unroll
    for (int i = 0; i < 3; i++)
        v_means[g * 3 + i] = vm[i], v_scales[g * 3 + i] = vs[i];
#pragma unroll
    for (int i = 0; i < 4; i++)
        v_quats[g * 4 + i] = vq[i];
}

__global__ void __launch_bounds__(128) k_staged_sh_fwd(int N, const float *__restrict__ dirs, const float *__restrict__ coeffs,
                                                        const unsigned char *__restrict__ mask, float *colors)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= N)
        return;
    float c[3] = {0.f, 0.f, 0.f};
    if (!mask || mask[g])
    {
        float dir[3] = {dirs[g * 3 + 0], dirs[g * 3 + 1], dirs[g * 3 + 2]};
        ShBasis sb;
        sh_basis(dir, sb);
        const float *cf = coeffs + (size_t)g * 48;
#pragma unroll
        for (int ch = 0; ch < 3; ch++)
            c[ch] = sh_eval_channel(sb, cf + ch, 3);
    }
    colors[g * 3 + 0] = c[0], colors[g * 3 + 1] = c[1], colors[g * 3 + 2] = c[2];
}

__global__ void __launch_bounds__(128) k_staged_sh_bwd(int N, const float *__restrict__ dirs, const float *__restrict__ coeffs,
                                                        const unsigned char *__restrict__ mask, const float *__restrict__ v_colors,
                                                        float *v_coeffs, float *v_dirs)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= N)
        return;
    float basis[16], vdir[3] = {0.f, 0.f, 0.f};
    float vc[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 16; k++)
        basis[k] = 0.f;
    if (!mask || mask[g])
    {
        vc[0] = v_colors[g * 3 + 0], vc[1] = v_colors[g * 3 + 1], vc[2] = v_colors[g * 3 + 2];
        float dir[3] = {dirs[g * 3 + 0], dirs[g * 3 + 1], dirs[g * 3 + 2]};
        ShBasis sb;
        sh_basis(dir, sb);
        sh_vjp(sb, coeffs + (size_t)g * 48, 3, vc, basis, vdir);
    }
    float *o = v_coeffs + (size_t)g * 48;
#pragma unroll
    for (int k = 0; k < 16; k++)
    {
        o[k * 3 + 0] = basis[k] * vc[0];
        o[k * 3 + 1] = basis[k] * vc[1];
        o[k * 3 + 2] = basis[k] * vc[2];
    }
    if (v_dirs)
        v_dirs[g * 3 + 0] = vdir[0], v_dirs[g * 3 + 1] = vdir[1], v_dirs[g * 3 + 2] = vdir[2];
}

// separate arrays -> packed 48-byte records (+ per-(tile, chunk) counts, tiles_per_gauss, backward work items)
__global__ void __launch_bounds__(256) k_staged_pack(int N, const float *__restrict__ means2d, const float *__restrict__ conics,
                                                      const float *__restrict__ colors4, const float *__restrict__ depths,
                                                      const float *__restrict__ opac, const int *__restrict__ radii, SplatRec *recs,
                                                      SplatGrad *grads, int *segCount,
                                                      int chunkSize, int tileW, int tileH, int W, int H, int *tilesPerGauss, int4 *items,
                                                      int itemCap, int *counters, int countTiles, int forBackward)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= N)
        return;
    const int radius = radii ? radii[g] : 1; // the reference's forward has no radii input: every splat is live
    if (radius <= 0)
    {
        recs[g].q0 = make_float4(0.f, 0.f, 0.f, __int_as_float(0));
        if (tilesPerGauss)
            tilesPerGauss[g] = 0;
        if (forBackward)
        {
            float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            grads[g].g0 = z, grads[g].g1 = z, grads[g].g2 = z;
        }
        return;
    }
    const float mx = means2d[g * 2 + 0], my = means2d[g * 2 + 1];
    const float o = opac ? opac[g] : 0.f;
    float ca = 0.f, cb = 0.f, cc = 0.f, r = 0.f, gr = 0.f, b = 0.f, d = 0.f;
    if (conics)
        ca = conics[g * 3 + 0], cb = conics[g * 3 + 1], cc = conics[g * 3 + 2];
    if (colors4)
        r = colors4[g * 4 + 0], gr = colors4[g * 4 + 1], b = colors4[g * 4 + 2], d = colors4[g * 4 + 3];
    if (depths)
        d = depths[g];
    int bits = 7; // colours arrive already clamped; their gradient is returned unmasked (the caller's autograd applies clamp_min)
    if (countTiles)
    {
        int x0, y0, x1, y1;
        tile_rect_s(mx, my, radius, tileW, tileH, x0, y0, x1, y1);
        const int chunk = g / chunkSize;
        for (int ty = y0; ty < y1; ty++)
            for (int tx = x0; tx < x1; tx++)
                atomicAdd(&segCount[(ty * tileW + tx) * BIN_CHUNKS + chunk], 1);
        if (tilesPerGauss)
            tilesPerGauss[g] = (y1 - y0) * (x1 - x0);
        atomicAdd(&counters[CNT_VISIBLE], 1);
    }
    if (forBackward)
    {
        int rx, ry, rw, rh;
        const int npix = bwd_rect(mx, my, radius, ca, cb, cc, o, W, H, rx, ry, rw, rh);
        const int nItems = (npix + BWD_PIXELS_PER_ITEM - 1) / BWD_PIXELS_PER_ITEM;
        if (nItems != 1)
        {
            if (nItems > 1)
                bits |= 256;
            float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            grads[g].g0 = z, grads[g].g1 = z, grads[g].g2 = z;
        }
        if (nItems > 0)
        {
            const int base = atomicAdd(&counters[CNT_ITEMS], nItems);
            if (base + nItems <= itemCap)
                for (int i = 0; i < nItems; i++)
                    items[base + i] = make_int4(g, i * BWD_PIXELS_PER_ITEM, rx | (ry << 16), rw | (rh << 16));
            else
            {
                atomicOr(&counters[CNT_OVERFLOW], 2);
                float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                grads[g].g0 = z, grads[g].g1 = z, grads[g].g2 = z;
            }
        }
    }
    SplatRec rec;
    rec.q0 = make_float4(mx, my, o, __int_as_float(radius));
    rec.q1 = make_float4(ca, cb, cc, d);
    rec.q2 = make_float4(r, gr, b, __int_as_float(bits));
    recs[g] = rec;
}

// isect_ids of the sorted list: with one camera the id is the tile index (isect_tiles_no_depth.cu:104-117)
__global__ void __launch_bounds__(256) k_staged_isect_ids(const int *__restrict__ tileOffsets, int T, long long *isectIds)
{
    const int t = blockIdx.x;
    const int s = tileOffsets[t], e = tileOffsets[t + 1];
    for (int i = s + threadIdx.x; i < e; i += 256)
        isectIds[i] = (long long)t;
}

__global__ void __launch_bounds__(256) k_staged_unpack_grads(int N, const SplatRec *__restrict__ recs, const SplatGrad *__restrict__ grads,
                                                              float *v_means2d, float *v_conics, float *v_colors4, float *v_opac)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= N)
        return;
    float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = g0, g2 = g0;
    if (__float_as_int(recs[g].q0.w) > 0)
        g0 = grads[g].g0, g1 = grads[g].g1, g2 = grads[g].g2;
    v_means2d[g * 2 + 0] = g0.x, v_means2d[g * 2 + 1] = g0.y;
    v_opac[g] = g0.z;
    v_conics[g * 3 + 0] = g1.x, v_conics[g * 3 + 1] = g1.y, v_conics[g * 3 + 2] = g1.z;
    v_colors4[g * 4 + 0] = g2.x, v_colors4[g * 4 + 1] = g2.y, v_colors4[g * 4 + 2] = g2.z, v_colors4[g * 4 + 3] = g0.w;
}

__global__ void __launch_bounds__(256) k_staged_cut(int P, const float *__restrict__ refDepth, float delta, float4 *v_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P)
        v_out[2 * i + 1] = make_float4(refDepth[i] + delta, 0.f, 0.f, 0.f);
}

__global__ void __launch_bounds__(256) k_staged_adam(int n, float *p, const float *__restrict__ g, float *m, float *v, AdamScalars a, float stepSize)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    float mo = m[i], vo = v[i];
    p[i] = adam_update(p[i], g[i], mo, vo, a, stepSize);
    m[i] = mo, v[i] = vo;
}

// ------------------------------------------------------------------------------------------------------------
void staged_project_fwd(int N, const float *means, const float *quats, const float *scales, const CamParams &cam, int *radii, float *means2d,
                        float *depths, float *conics, cudaStream_t st)
{
    if (N <= 0)
        return;
    GS_COUNT_LAUNCHES(1);
    k_staged_project_fwd<<<cdiv(N, 256), 256, 0, st>>>(N, means, quats, scales, cam, radii, means2d, depths, conics);
}

void staged_project_bwd(int N, const float *means, const float *quats, const float *scales, const CamParams &cam, const int *radii,
                        const float *conics, const float *v_means2d, const float *v_depths, const float *v_conics, float *v_means, float *v_quats,
                        float *v_scales, cudaStream_t st)
{
    if (N <= 0)
        return;
    GS_COUNT_LAUNCHES(1);
    k_staged_project_bwd<<<cdiv(N, 128), 128, 0, st>>>(N, means, quats, scales, cam, radii, conics, v_means2d, v_depths, v_conics, v_means,
                                                       v_quats, v_scales);
}

void staged_sh_fwd(int N, const float *dirs, const float *coeffs, const unsigned char *mask, float *colors, cudaStream_t st)
{
    if (N <= 0)
        return;
    GS_COUNT_LAUNCHES(1);
    k_staged_sh_fwd<<<cdiv(N, 128), 128, 0, st>>>(N, dirs, coeffs, mask, colors);
}

void staged_sh_bwd(int N, const float *dirs, const float *coeffs, const unsigned char *mask, const float *v_colors, float *v_coeffs, float *v_dirs,
                   cudaStream_t st)
{
    if (N <= 0)
        return;
    GS_COUNT_LAUNCHES(1);
    k_staged_sh_bwd<<<cdiv(N, 128), 128, 0, st>>>(N, dirs, coeffs, mask, v_colors, v_coeffs, v_dirs);
}

void staged_pack(int N, const float *means2d, const float *conics, const float *colors4, const float *depths, const float *opac, const int *radii,
                 SplatRec *recs, SplatGrad *grads, const Bins &bins, int tileW, int tileH, int W, int H, int *tilesPerGauss, bool countTiles,
                 bool forBackward, cudaStream_t st)
{
    if (N <= 0)
        return;
    GS_COUNT_LAUNCHES(1);
    k_staged_pack<<<cdiv(N, 256), 256, 0, st>>>(N, means2d, conics, colors4, depths, opac, radii, recs, grads, bins.segCount, bin_chunk_size(N), tileW,
                                                tileH, W, H, tilesPerGauss, bins.items, bins.itemCap, bins.counters, countTiles ? 1 : 0,
                                                forBackward ? 1 : 0);
}

void staged_isect_ids(const Bins &bins, int T, long long *isectIds, cudaStream_t st)
{
    GS_COUNT_LAUNCHES(1);
    k_staged_isect_ids<<<T, 256, 0, st>>>(bins.tileOffsets, T, isectIds);
}

void staged_unpack_grads(int N, const SplatRec *recs, const SplatGrad *grads, float *v_means2d, float *v_conics, float *v_colors4, float *v_opac,
                         cudaStream_t st)
{
    if (N <= 0)
        return;
    GS_COUNT_LAUNCHES(1);
    k_staged_unpack_grads<<<cdiv(N, 256), 256, 0, st>>>(N, recs, grads, v_means2d, v_conics, v_colors4, v_opac);
}

void staged_cut(int P, const float *refDepth, float delta, float4 *v_out, cudaStream_t st)
{
    GS_COUNT_LAUNCHES(1);
    k_staged_cut<<<cdiv(P, 256), 256, 0, st>>>(P, refDepth, delta, v_out);
}

void staged_adam(int n, float *p, const float *g, float *m, float *v, const AdamScalars &a, float step_size, cudaStream_t st)
{
    if (n <= 0)
        return;
    GS_COUNT_LAUNCHES(1);
    k_staged_adam<<<cdiv(n, 256), 256, 0, st>>>(n, p, g, m, v, a, step_size);
}

} // namespace gs
