// Fused SSIM map, forward and backward: reference gsplat/rasterizer/ssim.cu:214-460 (fusedssimCUDA,
// fusedssim_backwardCUDA; 11-tap separable Gaussian window, zero padding, NCHW fp32), used by RawGaussianModel::computeLoss when
// ssim_weight > 0 (src/raw_gs_model.cpp:383-395).
//
// One 32x32 output tile per CTA.  The reference runs five (forward) / three (backward) separate load -> convolve-x -> convolve-y
// rounds through shared memory with ~25 block barriers per channel; here the 42x42 halo tiles of both images are loaded once, the
// horizontal pass produces all five row-filtered planes (x, y, x^2, y^2, xy) in one sweep and the vertical pass finishes them:
// 3 barriers per channel.  Taps are accumulated in the reference's order (k = 0..10).
#include "common.cuh"
#include "gs.h"

namespace gs
{

__constant__ float c_gauss11[11] = {0.001028380123898387f, 0.0075987582094967365f, 0.036000773310661316f, 0.10936068743467331f,
                                    0.21300552785396576f, 0.26601171493530273f,   0.21300552785396576f,  0.10936068743467331f,
                                    0.036000773310661316f, 0.0075987582094967365f, 0.001028380123898387f};

constexpr int SB = 32;        // tile edge
constexpr int SH_ = SB + 10;  // halo tile edge

__device__ __forceinline__ float pix_or_zero(const float *img, int y, int x, int H, int W)
{
    return (x >= 0 && y >= 0 && x < W && y < H) ? __ldg(&img[(size_t)y * W + x]) : 0.0f;
}

// grid (ceil(W/32), ceil(H/32), B*CH); block 32 x 8 (each thread finishes 4 output rows)
__global__ void __launch_bounds__(256) k_ssim_fwd(int H, int W, float C1, float C2, const float *__restrict__ img1, const float *__restrict__ img2,
                                                   float *__restrict__ ssimMap, float *__restrict__ dm_dmu1, float *__restrict__ dm_dsigma1_sq,
                                                   float *__restrict__ dm_dsigma12)
{
    __shared__ float sA[SH_][SH_ + 1], sB[SH_][SH_ + 1];
    __shared__ float sR[5][SH_][SB + 1]; // row-filtered x, y, x^2, y^2, xy
    const size_t plane = (size_t)blockIdx.z * H * W;
    const float *a = img1 + plane, *b = img2 + plane;
    const int x0 = blockIdx.x * SB, y0 = blockIdx.y * SB;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    for (int k = tid; k < SH_ * SH_; k += 256)
    {
        const int ly = k / SH_, lx = k - ly * SH_;
        sA[ly][lx] = pix_or_zero(a, y0 + ly - 5, x0 + lx - 5, H, W);
        sB[ly][lx] = pix_or_zero(b, y0 + ly - 5, x0 + lx - 5, H, W);
    }
    __syncthreads();
    for (int k = tid; k < SH_ * SB; k += 256)
    {
        const int ly = k / SB, lx = k - ly * SB;
        float r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f, r4 = 0.f;
#pragma unroll
        for (int t = 0; t < 11; t++)
        {
            const float g = c_gauss11[t], p = sA[ly][lx + t], q = sB[ly][lx + t];
            r0 += g * p, r1 += g * q, r2 += g * (p * p), r3 += g * (q * q), r4 += g * (p * q);
        }
        sR[0][ly][lx] = r0, sR[1][ly][lx] = r1, sR[2][ly][lx] = r2, sR[3][ly][lx] = r3, sR[4][ly][lx] = r4;
    }
    __syncthreads();
    const int lx = threadIdx.x;
#pragma unroll
    for (int r = 0; r < 4; r++)
    {
        const int ly = threadIdx.y * 4 + r;
        float mu1 = 0.f, mu2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
        for (int t = 0; t < 11; t++)
        {
            const float g = c_gauss11[t];
            mu1 += g * sR[0][ly + t][lx], mu2 += g * sR[1][ly + t][lx], e11 += g * sR[2][ly + t][lx], e22 += g * sR[3][ly + t][lx],
                e12 += g * sR[4][ly + t][lx];
        }
        const int x = x0 + lx, y = y0 + ly;
        if (x < W && y < H)
        {
            const float sigma1_sq = e11 - mu1 * mu1, sigma2_sq = e22 - mu2 * mu2, sigma12 = e12 - mu1 * mu2;
            const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu1_mu2 = mu1 * mu2;
            const float C = 2.0f * mu1_mu2 + C1, D = 2.0f * sigma12 + C2, A = mu1_sq + mu2_sq + C1, B = sigma1_sq + sigma2_sq + C2;
            const size_t o = plane + (size_t)y * W + x;
            ssimMap[o] = (C * D) / (A * B);
            if (dm_dmu1)
            {
                dm_dmu1[o] = (mu2 * 2.0f * D) / (A * B) - (mu2 * 2.0f * C) / (A * B) - (mu1 * 2.0f * C * D) / (A * A * B) + (mu1 * 2.0f * C * D) / (A * B * B);
                dm_dsigma1_sq[o] = (-C * D) / (A * B * B);
                dm_dsigma12[o] = (2 * C) / (A * B);
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_ssim_bwd(int H, int W, const float *__restrict__ img1, const float *__restrict__ img2,
                                                   const float *__restrict__ dL_dmap, const float *__restrict__ dm_dmu1,
                                                   const float *__restrict__ dm_dsigma1_sq, const float *__restrict__ dm_dsigma12,
                                                   float *__restrict__ dL_dimg1)
{
    __shared__ float sP[3][SH_][SH_ + 1]; // dL * dm_dmu1, dL * dm_dsigma1_sq, dL * dm_dsigma12 with halo
    __shared__ float sR[3][SH_][SB + 1];
    const size_t plane = (size_t)blockIdx.z * H * W;
    const int x0 = blockIdx.x * SB, y0 = blockIdx.y * SB;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    for (int k = tid; k < SH_ * SH_; k += 256)
    {
        const int ly = k / SH_, lx = k - ly * SH_;
        const int y = y0 + ly - 5, x = x0 + lx - 5;
        const float dl = pix_or_zero(dL_dmap + plane, y, x, H, W);
        sP[0][ly][lx] = pix_or_zero(dm_dmu1 + plane, y, x, H, W) * dl;
        sP[1][ly][lx] = pix_or_zero(dm_dsigma1_sq + plane, y, x, H, W) * dl;
        sP[2][ly][lx] = pix_or_zero(dm_dsigma12 + plane, y, x, H, W) * dl;
    }
    __syncthreads();
    for (int k = tid; k < SH_ * SB; k += 256)
    {
        const int ly = k / SB, lx = k - ly * SB;
        float r0 = 0.f, r1 = 0.f, r2 = 0.f;
#pragma unroll
        for (int t = 0; t < 11; t++)
        {
            const float g = c_gauss11[t];
            r0 += g * sP[0][ly][lx + t], r1 += g * sP[1][ly][lx + t], r2 += g * sP[2][ly][lx + t];
        }
        sR[0][ly][lx] = r0, sR[1][ly][lx] = r1, sR[2][ly][lx] = r2;
    }
    __syncthreads();
    const int lx = threadIdx.x;
#pragma unroll
    for (int r = 0; r < 4; r++)
    {
        const int ly = threadIdx.y * 4 + r;
        float c0 = 0.f, c1 = 0.f, c2 = 0.f;
#pragma unroll
        for (int t = 0; t < 11; t++)
        {
            const float g = c_gauss11[t];
            c0 += g * sR[0][ly + t][lx], c1 += g * sR[1][ly + t][lx], c2 += g * sR[2][ly + t][lx];
        }
        const int x = x0 + lx, y = y0 + ly;
        if (x < W && y < H)
        {
            const size_t o = plane + (size_t)y * W + x;
            const float p1 = __ldg(&img1[o]), p2 = __ldg(&img2[o]);
            float d = 0.0f;
            d += c0;
            d += p1 * 2.0f * c1;
            d += p2 * c2;
            dL_dimg1[o] = d;
        }
    }
}

void ssim_fwd(int planes, int H, int W, float C1, float C2, const float *img1, const float *img2, float *ssimMap, float *dm_dmu1,
              float *dm_dsigma1_sq, float *dm_dsigma12, cudaStream_t st)
{
    GS_COUNT_LAUNCHES(1);
    dim3 grid((W + SB - 1) / SB, (H + SB - 1) / SB, planes), block(32, 8, 1);
    k_ssim_fwd<<<grid, block, 0, st>>>(H, W, C1, C2, img1, img2, ssimMap, dm_dmu1, dm_dsigma1_sq, dm_dsigma12);
}

void ssim_bwd(int planes, int H, int W, const float *img1, const float *img2, const float *dL_dmap, const float *dm_dmu1,
              const float *dm_dsigma1_sq, const float *dm_dsigma12, float *dL_dimg1, cudaStream_t st)
{
    GS_COUNT_LAUNCHES(1);
    dim3 grid((W + SB - 1) / SB, (H + SB - 1) / SB, planes), block(32, 8, 1);
    k_ssim_bwd<<<grid, block, 0, st>>>(H, W, img1, img2, dL_dmap, dm_dmu1, dm_dsigma1_sq, dm_dsigma12, dL_dimg1);
}

} // namespace gs
