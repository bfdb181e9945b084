// The classic (depth-sorted, front-to-back alpha compositing) 3DGS rasteriser of the reference's render_method "raw"
// (RawGaussianModel::rawForward, src/raw_gs_model.cpp:37-186) -- SURVEY.md section 8(f) row 3 -- hand-written for sm_100a:
//
//   k_sort_tiles_depth   orders every tile list by (camera depth, Gaussian id) = the order of the reference's stable radix sort of
//                        isect_ids = tile << 32 | depth bits (gsplat/rasterizer/isect_tiles.cu:57-301)
//   k_raw_fwd            rasterize_to_pixels_fwd_kernel (rasterize_to_pixels_fwd.cu:17-196): per pixel, front to back,
//                        alpha = min(0.999, o * exp(-sigma)), T *= 1 - alpha, stop when T would drop to <= 1e-4; splat records
//                        (log2-domain conic, colours included) staged per 256-splat batch in shared memory
//   k_raw_bwd            rasterize_to_pixels_bwd_kernel (rasterize_to_pixels_bwd.cu:17-287): per pixel, back to front with the
//                        transmittance recovered by division; the 10 per-splat gradient values of the 32 pixels of a warp are
//                        summed with a 16-shuffle multi-value warp reduction and written by 10 lanes in parallel (the reference:
//                        10 x 5 shuffles, then 10 atomics issued one after the other by lane 0)
//
// COLOR_DIM is 4 (rgb + camera depth, src/raw_gs_model.cpp:132) as on the GES path.
#include "common.cuh"
#include "gs.h"

namespace gs
{

constexpr int RAW_B = 256; // splats per staged batch = threads per tile CTA
constexpr int SORT_DEPTH_SMEM = 2048;

__device__ __forceinline__ float raw_ex2(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sort_tiles_depth(const int *__restrict__ tileOffsets, const SplatRec *__restrict__ recs, int *flattenSorted,
                                                           int *scratch)
{
    __shared__ unsigned long long s[SORT_DEPTH_SMEM];
    const int t = blockIdx.x;
    const int start = tileOffsets[t], L = tileOffsets[t + 1] - start;
    const int tid = threadIdx.x;
    if (L <= 1)
        return;
    if (L <= SORT_DEPTH_SMEM)
    {
        int n = 2;
        while (n < L)
            n <<= 1;
        for (int i = tid; i < n; i += 256)
        {
            unsigned long long key = ~0ull;
            if (i < L)
            {
                const int g = flattenSorted[start + i];
                key = ((unsigned long long)__float_as_uint(__ldg(&recs[g].q1.w)) << 32) | (unsigned)g; // camera depth > 0: bits are monotonic
            }
            s[i] = key;
        }
        __syncthreads();
        for (int k = 2; k <= n; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1)
            {
                for (int i = tid; i < n; i += 256)
                {
                    const int ixj = i ^ j;
                    if (ixj > i)
                    {
                        const unsigned long long a = s[i], b = s[ixj];
                        const bool asc = (i & k) == 0;
                        if ((a > b) == asc)
                            s[i] = b, s[ixj] = a;
                    }
                }
                __syncthreads();
            }
        for (int i = tid; i < L; i += 256)
            flattenSorted[start + i] = (int)(s[i] & 0xffffffffu);
    }
    else
    {
        // very long list (> SORT_DEPTH_SMEM splats on one tile): rank sort from global memory into the scatter buffer (free at this
        // point), then copy back; keys are unique
        for (int i = tid; i < L; i += 256)
        {
            const int g = flattenSorted[start + i];
            const unsigned long long key = ((unsigned long long)__float_as_uint(__ldg(&recs[g].q1.w)) << 32) | (unsigned)g;
            int rank = 0;
            for (int j = 0; j < L; j++)
            {
                const int gj = flattenSorted[start + j];
                const unsigned long long kj = ((unsigned long long)__float_as_uint(__ldg(&recs[gj].q1.w)) << 32) | (unsigned)gj;
                rank += (kj < key);
            }
            scratch[start + rank] = g;
        }
        __syncthreads();
        for (int i = tid; i < L; i += 256)
            flattenSorted[start + i] = scratch[start + i];
    }
}

// isect_ids of the depth-sorted list: tile << 32 | depth bits (isect_tiles.cu:104-117), one camera
__global__ void __launch_bounds__(256) k_isect_ids_depth(const int *__restrict__ tileOffsets, const int *__restrict__ flattenSorted,
                                                          const SplatRec *__restrict__ recs, long long *isectIds)
{
    const int t = blockIdx.x;
    const int s = tileOffsets[t], e = tileOffsets[t + 1];
    for (int i = s + threadIdx.x; i < e; i += 256)
        isectIds[i] = ((long long)t << 32) | (long long)__float_as_uint(__ldg(&recs[flattenSorted[i]].q1.w));
}

// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_raw_fwd(const SplatRec *__restrict__ recs, const int *__restrict__ tileOffsets,
                                                  const int *__restrict__ flattenSorted, int W, int H, int tileW,
                                                  const float *__restrict__ background /* [4] or null */, float4 *__restrict__ render4,
                                                  float *__restrict__ alphas, int *__restrict__ lastIds)
{
    __shared__ float4 sAll[3 * RAW_B]; // mean/opacity | log2 conic, depth | rgb
    const int tile = blockIdx.x;
    const int tyi = tile / tileW, txi = tile - tyi * tileW;
    const int tid = threadIdx.x;
    const int i = tyi * TILE + (tid >> 4), j = txi * TILE + (tid & 15);
    const bool inside = (i < H) && (j < W);
    const int pix = i * W + j;
    const float px = (float)j + 0.5f, py = (float)i + 0.5f;
    const int start = tileOffsets[tile], end = tileOffsets[tile + 1];
    bool done = !inside;
    float T = 1.0f;
    int curIdx = 0;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int b = start; b < end; b += RAW_B)
    {
        // end early if every pixel of the tile is done (also the barrier that protects the staging buffers)
        if (__syncthreads_count(done) >= RAW_B)
            break;
        const int idx = b + tid;
        if (idx < end)
        {
            const SplatRec *r = recs + __ldg(&flattenSorted[idx]);
            const float4 q0 = __ldg(&r->q0), q1 = __ldg(&r->q1);
            sAll[tid] = q0;
            sAll[RAW_B + tid] = make_float4((0.5f * 1.4426950408889634f) * q1.x, 1.4426950408889634f * q1.y, (0.5f * 1.4426950408889634f) * q1.z, q1.w);
            sAll[2 * RAW_B + tid] = __ldg(&r->q2);
        }
        __syncthreads();
        const int n = min(RAW_B, end - b);
        for (int t = 0; t < n && !done; t++)
        {
            const float4 c = sAll[RAW_B + t];
            const float4 xyo = sAll[t];
            const float dx = xyo.x - px, dy = xyo.y - py;
            const float sigma = fmaf(dx, fmaf(c.y, dy, c.x * dx), (c.z * dy) * dy);
            const float alpha = fminf(0.999f, xyo.z * raw_ex2(-sigma));
            if (sigma < 0.f || alpha < 1.f / 255.f)
                continue;
            const float nextT = T * (1.0f - alpha);
            if (nextT <= 1e-4f)
            {
                done = true; // this pixel is done: exclusive
                break;
            }
            const float vis = alpha * T;
            const float4 col = sAll[2 * RAW_B + t];
            a0 += col.x * vis, a1 += col.y * vis, a2 += col.z * vis, a3 += c.w * vis;
            curIdx = b + t;
            T = nextT;
        }
    }
    if (inside)
    {
        alphas[pix] = 1.0f - T;
        if (background)
            a0 += T * background[0], a1 += T * background[1], a2 += T * background[2], a3 += T * background[3];
        render4[pix] = make_float4(a0, a1, a2, a3);
        lastIds[pix] = curIdx;
    }
}

// 16 values per lane -> warp totals with 16 shuffles; lane l ends up with the total of slot l >> 1
__device__ __forceinline__ float raw_warp_reduce16(float (&v)[16], int lane)
{
    const unsigned full = 0xffffffffu;
    float w8[8], w4[4], w2[2];
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
#pragma unroll
    for (int i = 0; i < 8; i++)
    {
        float send = h16 ? v[i] : v[i + 8], keep = h16 ? v[i + 8] : v[i];
        w8[i] = keep + __shfl_xor_sync(full, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
    {
        float send = h8 ? w8[i] : w8[i + 4], keep = h8 ? w8[i + 4] : w8[i];
        w4[i] = keep + __shfl_xor_sync(full, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 2; i++)
    {
        float send = h4 ? w4[i] : w4[i + 2], keep = h4 ? w4[i + 2] : w4[i];
        w2[i] = keep + __shfl_xor_sync(full, send, 4);
    }
    float send = h2 ? w2[0] : w2[1], keep = h2 ? w2[1] : w2[0];
    float w1 = keep + __shfl_xor_sync(full, send, 2);
    return w1 + __shfl_xor_sync(full, w1, 1);
}

__global__ void __launch_bounds__(256) k_raw_bwd(const SplatRec *__restrict__ recs, const int *__restrict__ tileOffsets,
                                                  const int *__restrict__ flattenSorted, int W, int H, int tileW,
                                                  const float *__restrict__ background, const float *__restrict__ alphas,
                                                  const int *__restrict__ lastIds, const float4 *__restrict__ v_render4,
                                                  const float *__restrict__ v_alphas, SplatGrad *__restrict__ grads)
{
    __shared__ float4 sAll[3 * RAW_B];
    __shared__ int sId[RAW_B];
    const int tile = blockIdx.x;
    const int tyi = tile / tileW, txi = tile - tyi * tileW;
    const int tid = threadIdx.x, lane = tid & 31;
    // a warp covers an 8 x 4 pixel patch of the tile: its 32 pixels stop contributing at nearly the same splat
    const int wid = tid >> 5;
    const int i = tyi * TILE + (wid >> 1) * 4 + (lane >> 3), j = txi * TILE + (wid & 1) * 8 + (lane & 7);
    const bool inside = (i < H) && (j < W);
    const int pix = min(i * W + j, W * H - 1);
    const float px = (float)j + 0.5f, py = (float)i + 0.5f;
    const int start = tileOffsets[tile], end = tileOffsets[tile + 1];
    const float Tfinal = 1.0f - alphas[pix];
    float T = Tfinal;
    float b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f; // contribution of the splats behind the current one
    const int binFinal = inside ? lastIds[pix] : 0;
    const float4 vr = v_render4[pix];
    const float va = v_alphas[pix];
    float vbg = 0.f;
    if (background)
        vbg = background[0] * vr.x + background[1] * vr.y + background[2] * vr.z + background[3] * vr.w;
    int warpBinFinal = binFinal;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
        warpBinFinal = max(warpBinFinal, __shfl_xor_sync(0xffffffffu, warpBinFinal, d));
    for (int batchEnd = end - 1; batchEnd >= start; batchEnd -= RAW_B)
    {
        __syncthreads();
        const int n = min(RAW_B, batchEnd + 1 - start);
        const int idx = batchEnd - tid; // slot 0 = furthest back
        if (idx >= start)
        {
            const int g = __ldg(&flattenSorted[idx]);
            const SplatRec *r = recs + g;
            sId[tid] = g;
            sAll[tid] = __ldg(&r->q0);
            sAll[RAW_B + tid] = __ldg(&r->q1);
            sAll[2 * RAW_B + tid] = __ldg(&r->q2);
        }
        __syncthreads();
        for (int t = max(0, batchEnd - warpBinFinal); t < n; t++)
        {
            bool valid = inside && (batchEnd - t <= binFinal);
            float alpha = 0.f, opac = 0.f, dx = 0.f, dy = 0.f, vis = 0.f;
            float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid)
            {
                c = sAll[RAW_B + t];
                const float4 xyo = sAll[t];
                opac = xyo.z;
                dx = xyo.x - px, dy = xyo.y - py;
                // exactly the forward's expression, so that both passes agree on which splats a pixel saw
                const float A = (0.5f * 1.4426950408889634f) * c.x, B = 1.4426950408889634f * c.y, Cc = (0.5f * 1.4426950408889634f) * c.z;
                const float sigma = fmaf(dx, fmaf(B, dy, A * dx), (Cc * dy) * dy);
                vis = raw_ex2(-sigma);
                alpha = fminf(0.999f, opac * vis);
                if (sigma < 0.f || alpha < 1.f / 255.f)
                    valid = false;
            }
            if (!__any_sync(0xffffffffu, valid))
                continue;
            float v16[16];
#pragma unroll
            for (int k = 0; k < 16; k++)
                v16[k] = 0.f;
            if (valid)
            {
                const float ra = 1.0f / (1.0f - alpha);
                T *= ra;
                const float fac = alpha * T;
                const float4 col = sAll[2 * RAW_B + t];
                // slots follow SplatGrad: (vx, vy, vo, vd | vca, vcb, vcc, - | vr, vg, vb, -)
                v16[8] = fac * vr.x, v16[9] = fac * vr.y, v16[10] = fac * vr.z, v16[3] = fac * vr.w;
                float v_alpha = (col.x * T - b0 * ra) * vr.x + (col.y * T - b1 * ra) * vr.y + (col.z * T - b2 * ra) * vr.z + (c.w * T - b3 * ra) * vr.w;
                v_alpha += Tfinal * ra * va;
                if (background)
                    v_alpha += -Tfinal * ra * vbg;
                if (opac * vis <= 0.999f)
                {
                    const float v_sigma = -opac * vis * v_alpha;
                    v16[4] = 0.5f * v_sigma * dx * dx, v16[5] = v_sigma * dx * dy, v16[6] = 0.5f * v_sigma * dy * dy;
                    v16[0] = v_sigma * (c.x * dx + c.y * dy), v16[1] = v_sigma * (c.y * dx + c.z * dy);
                    v16[2] = vis * v_alpha;
                }
                b0 += col.x * fac, b1 += col.y * fac, b2 += col.z * fac, b3 += c.w * fac;
            }
            const float total = raw_warp_reduce16(v16, lane);
            const int slot = lane >> 1;
            if ((lane & 1) == 0 && slot < 11 && slot != 7)
                atomicAdd(reinterpret_cast<float *>(grads + sId[t]) + slot, total);
        }
    }
}

__global__ void __launch_bounds__(256) k_zero_grads(int N, SplatGrad *grads)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < N)
    {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        grads[g].g0 = z, grads[g].g1 = z, grads[g].g2 = z;
    }
}

// ------------------------------------------------------------------------------------------------------------
void sort_tiles_depth(const SplatRec *recs, const Bins &bins, int T, cudaStream_t st)
{
    GS_COUNT_LAUNCHES(1);
    k_sort_tiles_depth<<<T, 256, 0, st>>>(bins.tileOffsets, recs, bins.flattenSorted, bins.flatten);
}

void isect_ids_depth(const SplatRec *recs, const Bins &bins, int T, long long *isectIds, cudaStream_t st)
{
    GS_COUNT_LAUNCHES(1);
    k_isect_ids_depth<<<T, 256, 0, st>>>(bins.tileOffsets, bins.flattenSorted, recs, isectIds);
}

void raw_fwd(const SplatRec *recs, const Bins &bins, int W, int H, int tileW, int tileH, const float *background, float *render4, float *alphas,
             int *lastIds, cudaStream_t st)
{
    GS_COUNT_LAUNCHES(1);
    k_raw_fwd<<<tileW * tileH, 256, 0, st>>>(recs, bins.tileOffsets, bins.flattenSorted, W, H, tileW, background,
                                             reinterpret_cast<float4 *>(render4), alphas, lastIds);
}

void raw_bwd(int N, const SplatRec *recs, const Bins &bins, int W, int H, int tileW, int tileH, const float *background, const float *alphas,
             const int *lastIds, const float *v_render4, const float *v_alphas, SplatGrad *grads, cudaStream_t st)
{
    GS_COUNT_LAUNCHES(2);
    k_zero_grads<<<(N + 255) / 256, 256, 0, st>>>(N, grads);
    k_raw_bwd<<<tileW * tileH, 256, 0, st>>>(recs, bins.tileOffsets, bins.flattenSorted, W, H, tileW, background, alphas, lastIds,
                                             reinterpret_cast<const float4 *>(v_render4), v_alphas, grads);
}

} // namespace gs
