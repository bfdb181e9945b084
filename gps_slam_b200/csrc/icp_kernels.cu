// Point-to-plane ICP tracker, hand-written for sm_100a.  Built with -fmad=false.
//
// reference: ITMExtendedTracker (the compiled-in default, InfiniTAM/ITMLib/Trackers/Interface/ITMExtendedTracker.cpp:218-665,
// Shared/ITMExtendedTracker_Shared.h:51-143, 298-328, CUDA/ITMExtendedTracker_CUDA.cu:106-197, 297-454) and ITMDepthTracker
// (Interface/ITMDepthTracker.cpp:93-298, Shared/ITMDepthTracker_Shared.h:7-101, CUDA/ITMDepthTracker_CUDA.cu:60-272).
//
// What is different from the reference's CUDA tracker, by design:
//  * ONE persistent cooperative kernel tracks a whole frame: all pyramid levels, all LM iterations.  Per iteration the grid
//    reduces the 29 accumulators (count, f, g[6], lower-triangular H[21]) with warp shuffles -> one partial per CTA -> a
//    fixed-order final sum by CTA 0, whose first thread then runs the reference's host-side LM step on the device (normalise,
//    accept / reject, damp, 6x6 or 3x3 Cholesky solve, small-angle update, SE3 re-orthonormalisation through the exp/log maps
//    of se3.h).  The reference does a cudaMemset + kernel + blocking 116-byte cudaMemcpy + host solve per iteration, i.e. up to
//    140 host round trips per frame; here there is none inside a frame.
//  * the reduction is deterministic (fixed pixel->thread assignment, fixed summation order); the reference accumulates with
//    float atomicAdd in arbitrary order.
//  * the depth pyramid (C1) is built by one kernel per level on the same stream; the points / normals pyramids of the Extended
//    tracker are not built because only level 0 of the scene hierarchy is ever read (ITMExtendedTracker.cpp:297).
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>

#include "icp.h"
#include "spd_solve.h"

namespace icp
{

constexpr int MAX_LEVELS = 5;
constexpr int CTA = 256;
constexpr int NACC = 29; // count, f, g[6], H[21]

enum IterType
{
    IT_NONE = 0,
    IT_ROTATION = 1,
    IT_TRANSLATION = 2,
    IT_BOTH = 3
};

struct Level
{
    float *depth;
    int w, h;
    float4 intr; // fx fy cx cy of this level
    int type, nIter;
    float thresh; // spaceThresh (extended) / distThresh (icp)
};

struct Params
{
    Level lv[MAX_LEVELS];
    int nLevels;
    int kind; // 1 extended, 2 icp
    const float4 *pointsMap, *normalsMap;
    int sceneW, sceneH;
    float4 sceneIntr;
    Mat4 scenePose;
    float vfmin, vfmax, tukeyCutOff;
    int framesToSkip, framesToWeight, useWeights;
    float terminationThreshold;
    // ICP sharded over the GPUs of a box: this rank accumulates rows [h * rank / world, h * (rank + 1) / world) of every
    // level; the 29 sums are exchanged through peer memory inside the persistent kernel, every rank adds the world partial sums in rank
    // order and takes the identical LM step.  xchg[q] = rank q's exchange block [2][16][32] floats (slot 31 of a row = sequence stamp).
    int rank, world;
    float *xchg[16];
    unsigned xseqBase;   // exchanges completed by earlier launches (identical on every rank)
    int *err;            // host-mapped flag: a peer did not deliver within the spin bound
};

// device-resident tracker state; the host reads it back once per frame
struct State
{
    se3::Pose pose_d, lastGood;
    Mat4 approxInvPose;
    float f_old, lambda;
    float hessian_good[36], nabla_good[6];
    // cached for UpdatePoseQuality
    float hessian_q[36];
    float f_q;
    int n_q;
    int lastType;
    int validDepthMax;
    int converged;
    int iterationsRun;
    // last raw evaluation (icp_eval)
    float eval[NACC];
};

struct Tracker
{
    int kind, W, H;
    float vfmin, vfmax;
    int nLevels;
    int types[MAX_LEVELS], nIter[MAX_LEVELS];
    float thresh[MAX_LEVELS];
    float *pyr[MAX_LEVELS]; // level >= 1 owned
    float *partials;        // [grid][NACC]
    unsigned *barrier;
    State *state;
    State *hostState; // pinned
    int grid;
    // quality classifier (host)
    float svm_w[20], svm_b, q_mu[4], q_sigma[4];
    float *hk_table;
    int lastResult;
    float lastScore;
    // sharded ICP (set_shard): peers' exchange blocks, exchanges done so far, error flag
    int rank, world;
    float *xchg[16];
    unsigned xseq;
    int *err;
};

// ------------------------------------------------------------------------------------------------------------
__global__ void k_convert_depth(const short *__restrict__ in, float *__restrict__ out, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
    {
        short d = in[i];
        out[i] = d <= 0 ? -1.0f : (float)d * (1.0f / 1000.0f) + 0.0f;
    }
}

// C1: filterSubsampleWithHoles (ITMLowLevelEngine_Shared.h:48-69)
__global__ void k_subsample_holes(const float *__restrict__ in, int ow, float *__restrict__ out, int nw, int nh)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= nw || y >= nh)
        return;
    int sx = x * 2, sy = y * 2;
    float acc = 0.0f, good = 0.0f, p;
    p = in[(sx + 0) + (sy + 0) * ow];
    if (p > 0.0f)
        acc += p, good++;
    p = in[(sx + 1) + (sy + 0) * ow];
    if (p > 0.0f)
        acc += p, good++;
    p = in[(sx + 0) + (sy + 1) * ow];
    if (p > 0.0f)
        acc += p, good++;
    p = in[(sx + 1) + (sy + 1) * ow];
    if (p > 0.0f)
        acc += p, good++;
    if (good > 0)
        acc /= good;
    out[x + y * nw] = acc;
}

// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 mul4(const Mat4 &M, float x, float y, float z, float w)
{
    float4 r;
    r.x = M.m[0] * x + M.m[4] * y + M.m[8] * z + M.m[12] * w;
    r.y = M.m[1] * x + M.m[5] * y + M.m[9] * z + M.m[13] * w;
    r.z = M.m[2] * x + M.m[6] * y + M.m[10] * z + M.m[14] * w;
    r.w = M.m[3] * x + M.m[7] * y + M.m[11] * z + M.m[15] * w;
    return r;
}

// interpolateBilinear_withHoles (ITMPixelUtils.h:78-106)
__device__ __forceinline__ float4 bilinear_holes(const float4 *__restrict__ src, float px, float py, int W)
{
    const short ix = (short)floorf(px), iy = (short)floorf(py);
    const float dx = px - (float)ix, dy = py - (float)iy;
    const float4 a = __ldg(&src[ix + iy * W]), b = __ldg(&src[(ix + 1) + iy * W]);
    const float4 c = __ldg(&src[ix + (iy + 1) * W]), d = __ldg(&src[(ix + 1) + (iy + 1) * W]);
    float4 r;
    if (a.w < 0 || b.w < 0 || c.w < 0 || d.w < 0)
    {
        r.x = r.y = r.z = 0.f, r.w = -1.0f;
        return r;
    }
    r.x = (a.x * (1.0f - dx) * (1.0f - dy) + b.x * dx * (1.0f - dy) + c.x * (1.0f - dx) * dy + d.x * dx * dy);
    r.y = (a.y * (1.0f - dx) * (1.0f - dy) + b.y * dx * (1.0f - dy) + c.y * (1.0f - dx) * dy + d.y * dx * dy);
    r.z = (a.z * (1.0f - dx) * (1.0f - dy) + b.z * dx * (1.0f - dy) + c.z * (1.0f - dx) * dy + d.z * dx * dy);
    r.w = (a.w * (1.0f - dx) * (1.0f - dy) + b.w * dx * (1.0f - dy) + c.w * (1.0f - dx) * dy + d.w * dx * dy);
    return r;
}

// C2: computePerPointGH_exDepth_Ab / computePerPointGH_Depth_Ab.  A is always filled as 6 entries (rotation part, translation part).
template <bool EXT>
__device__ __forceinline__ bool per_point(const Params &P, const Level &L, const Mat4 &approxInvPose, int x, int y, float depth, float *A, float &b,
                                          float &weight)
{
    weight = 0.f;
    if (depth <= 1e-8f)
        return false;
    float tx = depth * (((float)x - L.intr.z) / L.intr.x);
    float ty = depth * (((float)y - L.intr.w) / L.intr.y);
    float4 pt = mul4(approxInvPose, tx, ty, depth, 1.0f);
    float4 rp = mul4(P.scenePose, pt.x, pt.y, pt.z, 1.0f);
    if (rp.z <= 0.0f)
        return false;
    float u = P.sceneIntr.x * rp.x / rp.z + P.sceneIntr.z;
    float v = P.sceneIntr.y * rp.y / rp.z + P.sceneIntr.w;
    if (!((u >= 0.0f) && (u <= (float)(P.sceneW - 2)) && (v >= 0.0f) && (v <= (float)(P.sceneH - 2))))
        return false;
    float4 cp = bilinear_holes(P.pointsMap, u, v, P.sceneW);
    if (cp.w < 0.0f)
        return false;
    float dx = cp.x - pt.x, dy = cp.y - pt.y, dz = cp.z - pt.z;
    float dist = dx * dx + dy * dy + dz * dz;
    if (EXT)
    {
        if (dist > P.tukeyCutOff * L.thresh)
            return false;
    }
    else
    {
        if (dist > L.thresh)
            return false;
    }
    float4 n = bilinear_holes(P.normalsMap, u, v, P.sceneW);
    if (EXT)
    {
        weight = fmaxf(0.0f, 1.0f - (depth - P.vfmin) / (P.vfmax - P.vfmin));
        weight *= weight;
        if (P.useWeights)
        {
            if (cp.w < (float)P.framesToSkip)
                return false;
            weight *= (cp.w - (float)P.framesToSkip) / (float)P.framesToWeight;
        }
    }
    b = n.x * dx + n.y * dy + n.z * dz;
    A[0] = +pt.z * n.y - pt.y * n.z;
    A[1] = -pt.z * n.x + pt.x * n.z;
    A[2] = +pt.y * n.x - pt.x * n.y;
    A[3] = n.x, A[4] = n.y, A[5] = n.z;
    return true;
}

// accumulate one level for one pose: acc[0] = count, acc[1] = f, acc[2..7] = g, acc[8..28] = lower-triangular H (row-major)
template <bool EXT>
__device__ __forceinline__ void accumulate(const Params &P, const Level &L, const Mat4 &approxInvPose, float *acc)
{
#pragma unroll
    for (int i = 0; i < NACC; i++)
        acc[i] = 0.f;
    const int type = L.type;
    const int off = (type == IT_TRANSLATION) ? 3 : 0;   // which half of A is active for short iterations
    const int noPara = (type == IT_BOTH) ? 6 : 3;
    const int iBegin = (int)((long long)L.h * P.rank / P.world) * L.w, iEnd = (int)((long long)L.h * (P.rank + 1) / P.world) * L.w;
    for (int i = iBegin + blockIdx.x * blockDim.x + threadIdx.x; i < iEnd; i += gridDim.x * blockDim.x)
    {
        int y = i / L.w, x = i - y * L.w;
        float A6[6], b, wgt;
        if (!per_point<EXT>(P, L, approxInvPose, x, y, __ldg(&L.depth[i]), A6, b, wgt))
            continue;
        float A[6];
#pragma unroll
        for (int k = 0; k < 6; k++)
            A[k] = (noPara == 6) ? A6[k] : (k < 3 ? A6[k + off] : 0.f);
        float lf, gd, hd;
        if (EXT)
        {
            // rho / rho_deriv / rho_deriv2 (ITMExtendedTracker_Shared.h:51-65)
            float hb = L.thresh;
            float tmp = fmaxf(fabsf(b) - hb, 0.0f);
            lf = (b * b - tmp * tmp) * wgt;
            gd = 2.0f * fminf(fmaxf(b, -hb), hb) * wgt;
            hd = (fabsf(b) < hb ? 2.0f : 0.0f) * wgt;
        }
        else
        {
            lf = b * b, gd = b, hd = 1.0f;
        }
        acc[0] += 1.0f;
        acc[1] += lf;
        int c = 8;
#pragma unroll
        for (int r = 0; r < 6; r++)
        {
            if (r < noPara)
            {
                acc[2 + r] += gd * A[r];
#pragma unroll
                for (int cc = 0; cc <= r; cc++)
                    acc[c + cc] += hd * A[r] * A[cc];
            }
            c += r + 1;
        }
    }
}

__device__ __forceinline__ void block_reduce_to(float *acc, float *smem /* [8][NACC] */, float *out)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NACC; i++)
    {
        float v = acc[i];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1)
            v += __shfl_xor_sync(0xffffffffu, v, d);
        if (lane == 0)
            smem[wid * NACC + i] = v;
    }
    __syncthreads();
    if (threadIdx.x < NACC)
    {
        float s = 0.f;
        for (int w = 0; w < CTA / 32; w++)
            s += smem[w * NACC + threadIdx.x];
        out[threadIdx.x] = s;
    }
    __syncthreads();
}

__device__ __forceinline__ void grid_sync(unsigned *bar, unsigned nblocks)
{
    __syncthreads();
    if (threadIdx.x == 0)
    {
        __threadfence();
        unsigned ticket = atomicAdd(bar, 1u);
        unsigned target = (ticket / nblocks + 1u) * nblocks;
        while (*(volatile unsigned *)bar < target)
            ;
        __threadfence();
    }
    __syncthreads();
}

// expand the 21 lower-triangular sums into the reference's 6x6 layout hessian[r + c*6] (ITMExtendedTracker_CUDA.cu:186-190)
__device__ __host__ inline void expand_hessian(const float *acc, int noPara, float *H, float *g)
{
    for (int i = 0; i < 36; i++)
        H[i] = 0.f;
    for (int i = 0; i < 6; i++)
        g[i] = 0.f;
    int counter = 8;
    for (int r = 0; r < 6; r++)
        for (int c = 0; c <= r; c++, counter++)
            if (r < noPara)
                H[r + c * 6] = acc[counter];
    for (int r = 0; r < noPara; ++r)
        for (int c = r + 1; c < noPara; c++)
            H[r + c * 6] = H[c + r * 6];
    for (int r = 0; r < noPara; r++)
        g[r] = acc[2 + r];
}

// C4: one LM step of TrackCamera on the device (ITMExtendedTracker.cpp:512-661 / ITMDepthTracker.cpp:257-292)
template <bool EXT>
__device__ void lm_step(State &S, const float *sum, int type, float terminationThreshold)
{
    const int noPara = (type == IT_BOTH) ? 6 : 3;
    float H[36], g[6];
    expand_hessian(sum, noPara, H, g);
    int n = (int)sum[0];
    float f = sum[1];
    bool reject;
    if (EXT)
    {
        if (n > 100)
        {
            for (int i = 0; i < 36; i++)
                H[i] /= (float)n;
            for (int i = 0; i < 6; i++)
                g[i] /= (float)n;
            f /= (float)n;
        }
        else
            f = FLT_MAX;
        reject = (n <= 0) || (f >= S.f_old);
    }
    else
    {
        f = (n > 100) ? f / (float)n : 1e5f;
        reject = (n <= 0) || (f > S.f_old);
    }
    if (reject)
    {
        S.pose_d = S.lastGood;
        S.approxInvPose = S.pose_d.get_invM();
        S.lambda *= 10.0f;
    }
    else
    {
        S.lastGood = S.pose_d;
        S.f_old = f;
        if (EXT)
        {
            for (int i = 0; i < 36; i++)
                S.hessian_good[i] = H[i];
            for (int i = 0; i < 6; i++)
                S.nabla_good[i] = g[i];
        }
        else
        {
            for (int i = 0; i < 36; i++)
                S.hessian_good[i] = H[i] / (float)n;
            for (int i = 0; i < 6; i++)
                S.nabla_good[i] = g[i] / (float)n;
        }
        S.lambda /= 10.0f;
        S.n_q = n, S.f_q = f;
        for (int i = 0; i < 36; i++)
            S.hessian_q[i] = S.hessian_good[i];
    }
    float A[36];
    for (int i = 0; i < 36; i++)
        A[i] = S.hessian_good[i];
    for (int i = 0; i < 6; i++)
        A[i + i * 6] *= 1.0f + S.lambda;
    // ComputeDelta
    float step[6] = {0, 0, 0, 0, 0, 0};
    if (noPara == 3)
    {
        float small[9];
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++)
                small[r + c * 3] = A[r + c * 6];
        SpdFactor<3>(small).solve(S.nabla_good, step);
    }
    else
        SpdFactor<6>(A).solve(S.nabla_good, step);
    // ApplyDelta
    float st[6];
    if (type == IT_ROTATION)
        st[0] = step[0], st[1] = step[1], st[2] = step[2], st[3] = st[4] = st[5] = 0.f;
    else if (type == IT_TRANSLATION)
        st[0] = st[1] = st[2] = 0.f, st[3] = step[0], st[4] = step[1], st[5] = step[2];
    else
        for (int i = 0; i < 6; i++)
            st[i] = step[i];
    Mat4 T; // column-major m[col*4+row]; the reference names members mCR
    T.m[0] = 1.0f, T.m[4] = st[2], T.m[8] = -st[1], T.m[12] = st[3];
    T.m[1] = -st[2], T.m[5] = 1.0f, T.m[9] = st[0], T.m[13] = st[4];
    T.m[2] = st[1], T.m[6] = -st[0], T.m[10] = 1.0f, T.m[14] = st[5];
    T.m[3] = 0.0f, T.m[7] = 0.0f, T.m[11] = 0.0f, T.m[15] = 1.0f;
    Mat4 next = se3::mul(T, S.approxInvPose);
    S.pose_d.set_invM(next);
    S.pose_d.coerce();
    S.approxInvPose = S.pose_d.get_invM();
    // HasConverged
    bool conv;
    if (EXT)
    {
        conv = true;
        for (int i = 0; i < 6; i++)
            if (fabsf(step[i]) > terminationThreshold)
                conv = false;
    }
    else
    {
        float len = 0.f;
        for (int i = 0; i < 6; i++)
            len += step[i] * step[i];
        conv = sqrtf(len) / 6 < terminationThreshold;
    }
    S.converged = conv ? 1 : 0;
    S.iterationsRun++;
    S.lastType = type;
}

template <bool EXT>
__global__ void __launch_bounds__(CTA) k_icp_track(Params P, State *state, float *partials, unsigned *barrier)
{
    __shared__ float sred[(CTA / 32) * NACC];
    __shared__ float sfinal[NACC];
    __shared__ Mat4 sPose;
    const unsigned nb = gridDim.x;
    // C5 input: CountValidDepths on the full-resolution depth
    {
        const Level &L0 = P.lv[0];
        int cnt = 0;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L0.w * L0.h; i += gridDim.x * blockDim.x)
            cnt += (__ldg(&L0.depth[i]) > 0.0f) ? 1 : 0;
        for (int d = 16; d > 0; d >>= 1)
            cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
        if ((threadIdx.x & 31) == 0 && cnt)
            atomicAdd(&state->validDepthMax, cnt);
    }
    for (int level = P.nLevels - 1; level >= 0; level--)
    {
        const Level &L = P.lv[level];
        if (L.type == IT_NONE)
            continue;
        if (blockIdx.x == 0 && threadIdx.x == 0)
        {
            State &S = *state;
            S.approxInvPose = S.pose_d.get_invM();
            S.lastGood = S.pose_d;
            S.f_old = EXT ? FLT_MAX : 1e20f;
            S.lambda = 1.0f;
            S.converged = 0;
            if (!EXT)
                S.n_q = 0;
        }
        grid_sync(barrier, nb);
        for (int iter = 0; iter < L.nIter; iter++)
        {
            if (threadIdx.x < 16)
                sPose.m[threadIdx.x] = __ldcg(&state->approxInvPose.m[threadIdx.x]);
            __syncthreads();
            float acc[NACC];
            accumulate<EXT>(P, L, sPose, acc);
            block_reduce_to(acc, sred, partials + (size_t)blockIdx.x * NACC);
            grid_sync(barrier, nb);
            if (blockIdx.x == 0)
            {
                float s = 0.f;
                if (threadIdx.x < NACC)
                {
                    for (unsigned b = 0; b < nb; b++)
                        s += __ldcg(&partials[(size_t)b * NACC + threadIdx.x]);
                    sfinal[threadIdx.x] = s;
                }
                if (P.world > 1)
                {
                    // one-shot all-reduce of the 29 sums through peer memory: store my partial sums into row `rank` of every rank's
                    // exchange block (double-buffered by sequence parity: a fast rank is at most one exchange ahead), stamp the rows,
                    // wait for every rank's stamp in my own block, add the rows in rank order
                    const unsigned seq = P.xseqBase + (unsigned)__ldcg(&state->iterationsRun) + 1u;
                    const int buf = (int)(seq & 1u);
                    if (threadIdx.x < NACC)
                        for (int q = 0; q < P.world; q++)
                            P.xchg[q][(buf * 16 + P.rank) * 32 + threadIdx.x] = s;
                    __threadfence_system();
                    __syncthreads();
                    if (threadIdx.x < P.world)
                    {
                        const int q = threadIdx.x;
                        unsigned *stamp = reinterpret_cast<unsigned *>(P.xchg[q] + (buf * 16 + P.rank) * 32 + 31);
                        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(stamp), "r"(seq) : "memory");
                        const unsigned *mine = reinterpret_cast<const unsigned *>(P.xchg[P.rank] + (buf * 16 + q) * 32 + 31);
                        const long long t0 = clock64();
                        for (;;)
                        {
                            unsigned got;
                            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(got) : "l"(mine) : "memory");
                            if (got == seq)
                                break;
                            if (clock64() - t0 > 40000000000LL)
                            {
                                *P.err = 1;
                                break;
                            }
                        }
                    }
                    __syncthreads();
                    if (threadIdx.x < NACC)
                    {
                        float tot = 0.f;
                        for (int q = 0; q < P.world; q++)
                        {
                            float part;
                            asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(part) : "l"(P.xchg[P.rank] + (buf * 16 + q) * 32 + threadIdx.x) : "memory");
                            tot += part;
                        }
                        sfinal[threadIdx.x] = tot;
                    }
                }
                __syncthreads();
                if (threadIdx.x == 0)
                    lm_step<EXT>(*state, sfinal, L.type, P.terminationThreshold);
            }
            grid_sync(barrier, nb);
            if (__ldcg(&state->converged))
                break;
        }
    }
}

// single evaluation at one level for a given pose (parity with ComputeGandH_Depth / ComputeGandH): raw sums in state->eval
template <bool EXT>
__global__ void __launch_bounds__(CTA) k_icp_eval(Params P, int level, Mat4 approxInvPose, State *state, float *partials, unsigned *barrier)
{
    __shared__ float sred[(CTA / 32) * NACC];
    float acc[NACC];
    accumulate<EXT>(P, P.lv[level], approxInvPose, acc);
    block_reduce_to(acc, sred, partials + (size_t)blockIdx.x * NACC);
    grid_sync(barrier, gridDim.x);
    if (blockIdx.x == 0 && threadIdx.x < NACC)
    {
        float s = 0.f;
        for (unsigned b = 0; b < gridDim.x; b++)
            s += __ldcg(&partials[(size_t)b * NACC + threadIdx.x]);
        state->eval[threadIdx.x] = s;
    }
}

// ------------------------------------------------------------------------------------------------------------
// ORUtils::HomkerMap(2) (ORUtils/HomkerMap.h:55-171, after VLFeat homkermap.c) for the tracking-quality SVM
struct Homker
{
    static constexpr int order = 2, numSub = 8 + 8 * order, minExp = -20, maxExp = 8, dim = 2 * order + 1;
    static float spectrum(float omega) { return 2.0f / (expf((float)M_PI * omega) + expf(-(float)M_PI * omega)); }
    static float sinc(float x) { return x == 0.0f ? 1.0f : sinf(x) / x; }
    static float smooth(float period, float omega)
    {
        float kappa_hat = 0;
        const float epsilon = 1e-2f;
        const float omegaRange = 2.0f / (period * epsilon);
        const float domega = 2.0f * omegaRange / (2.0f * 1024.0f + 1.0f);
        for (float omegap = -omegaRange; omegap <= omegaRange; omegap += domega)
        {
            float win = sinc((period / 2.0f) * omegap);
            win *= (period / (2.0f * (float)M_PI));
            kappa_hat += win * spectrum(omegap + omega);
        }
        kappa_hat *= domega;
        return fmaxf(kappa_hat, 0.0f);
    }
    static float *build()
    {
        float period = fmaxf(8.80f * sqrtf(order + 4.44f) - 12.6f, 1.0f);
        const int tableW = numSub * (maxExp - minExp + 1);
        float *table = new float[dim * tableW + 2 * (1 + order)]();
        float *tp = table, *kappa = table + dim * tableW, *freq = kappa + (1 + order);
        float L = 2.0f * (float)M_PI / period;
        int i = 0, j = 0;
        while (i <= order)
        {
            freq[i] = (float)j;
            kappa[i] = smooth(period, j * L);
            ++j;
            if (kappa[i] > 0 || j >= 3 * i)
                ++i;
        }
        for (int e = minExp; e <= maxExp; ++e)
        {
            float mantissa = 1.0f;
            for (int s = 0; s < numSub; ++s, mantissa += 1.0f / numSub)
            {
                float x = ldexpf(mantissa, e);
                float Lx = L * x, Llogx = L * logf(x);
                *tp++ = sqrtf(Lx * kappa[0]);
                for (int k = 1; k <= order; ++k)
                {
                    float q = sqrtf(2.0f * Lx * kappa[k]);
                    *tp++ = q * cosf(freq[k] * Llogx);
                    *tp++ = q * sinf(freq[k] * Llogx);
                }
            }
        }
        return table;
    }
    static void eval(const float *table, float *dst, float x)
    {
        int exponent;
        float mantissa = frexpf(x, &exponent);
        float sign = (mantissa >= 0.0f) ? +1.0f : -1.0f;
        mantissa *= 2.0f * sign;
        exponent--;
        if (mantissa == 0 || exponent <= minExp || exponent >= maxExp)
        {
            for (int j = 0; j < dim; ++j)
                dst[j] = 0.0f;
            return;
        }
        const float sub = 1.0f / numSub;
        const float *v1 = table + (exponent - minExp) * numSub * dim;
        mantissa -= 1.0f;
        while (mantissa >= sub)
        {
            mantissa -= sub;
            v1 += dim;
        }
        const float *v2 = v1 + dim;
        for (int j = 0; j < dim; ++j)
            dst[j] = sign * ((v2[j] - v1[j]) * (numSub * mantissa) + v1[j]);
    }
};

// ------------------------------------------------------------------------------------------------------------
Tracker *create_tracker(int kind, int W, int H, float vfmin, float vfmax)
{
    Tracker *t = new (std::nothrow) Tracker();
    if (!t)
        return nullptr;
    memset(t, 0, sizeof *t);
    t->kind = kind, t->W = W, t->H = H, t->vfmin = vfmin, t->vfmax = vfmax;
    float failureDec;
    if (kind == 1)
    {
        // "type=extended,levels=rrbb,useDepth=1,minstep=1e-4,outlierSpaceC=0.1,outlierSpaceF=0.004,numiterC=20,numiterF=50,
        //  tukeyCutOff=8,framesToSkip=20,framesToWeight=50,failureDec=20.0" (Utils/ITMLibSettings.cpp:54-57)
        t->nLevels = 4;
        const int types[4] = {IT_BOTH, IT_BOTH, IT_ROTATION, IT_ROTATION};
        float stepI = (float)(20 - 50) / 3.0f, valI = 20.0f, stepT = (0.1f - 0.004f) / 3.0f, valT = 0.1f;
        for (int l = 3; l >= 0; l--)
        {
            t->types[l] = types[l];
            t->nIter[l] = (int)roundf(valI);
            t->thresh[l] = valT;
            valI -= stepI, valT -= stepT;
        }
        failureDec = 20.0f;
    }
    else
    {
        // "type=icp,levels=rrrbb,minstep=1e-3,outlierC=0.01,outlierF=0.002,numiterC=10,numiterF=2,failureDec=5.0"
        t->nLevels = 5;
        const int types[5] = {IT_BOTH, IT_BOTH, IT_ROTATION, IT_ROTATION, IT_ROTATION};
        float stepI = (float)(10 - 2) / 4.0f, valI = 10.0f, stepT = (0.01f - 0.002f) / 4.0f, valT = 0.01f;
        for (int l = 4; l >= 0; l--)
        {
            t->types[l] = types[l];
            t->nIter[l] = (int)roundf(valI);
            t->thresh[l] = valT;
            valI -= stepI, valT -= stepT;
        }
        failureDec = 5.0f;
    }
    bool ok = true;
    int w = W, h = H;
    for (int l = 1; l < t->nLevels; l++)
    {
        w /= 2, h /= 2;
        ok = ok && cudaMalloc((void **)&t->pyr[l], sizeof(float) * (size_t)(w > 0 ? w : 1) * (h > 0 ? h : 1)) == cudaSuccess;
    }
    int dev = 0, sms = 148, perSm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (kind == 1)
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_icp_track<true>, CTA, 0);
    else
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_icp_track<false>, CTA, 0);
    if (perSm < 1)
        perSm = 1;
    if (perSm > 2)
        perSm = 2;
    // tests that run several sharded engines on ONE device: their persistent kernels wait for each other, so all of them must be
    // co-resident (one process per GPU -- the product layout -- has the device to itself)
    if (const char *v = getenv("GSB_ICP_CTAS_PER_SM"))
        perSm = atoi(v) >= 1 ? (atoi(v) < perSm ? atoi(v) : perSm) : perSm;
    t->grid = sms * perSm;
    ok = ok && cudaMalloc((void **)&t->partials, sizeof(float) * NACC * (size_t)t->grid) == cudaSuccess;
    ok = ok && cudaMalloc((void **)&t->barrier, sizeof(unsigned)) == cudaSuccess;
    ok = ok && cudaMalloc((void **)&t->state, sizeof(State)) == cudaSuccess;
    ok = ok && cudaMallocHost((void **)&t->hostState, sizeof(State)) == cudaSuccess;
    if (!ok)
    {
        destroy_tracker(t);
        return nullptr;
    }
    cudaMemset(t->barrier, 0, sizeof(unsigned));
    cudaMemset(t->state, 0, sizeof(State));
    // tracking-quality SVM (ITMExtendedTracker.cpp:91-123; ITMDepthTracker uses the same vectors)
    const float w20[20] = {-3.15813f, -2.38038f, 1.93359f, 1.56642f, 1.76306f, -0.747641f, 4.41852f, 1.72048f, -0.482545f, -5.07793f,
                           1.98676f,  -0.45688f, 2.53969f, -3.50527f, -1.68725f, 2.31608f,  5.14778f, 2.31334f, -14.128f,   6.76423f};
    memcpy(t->svm_w, w20, sizeof w20);
    t->svm_b = 9.334260e-01f + failureDec;
    const float mu[4] = {-34.9470512137603f, -33.1379108518478f, 0.195948598235857f, 0.611027292662361f};
    const float sg[4] = {68.1654461020426f, 60.6607826748643f, 0.00343068557187040f, 0.0402595570918749f};
    memcpy(t->q_mu, mu, sizeof mu), memcpy(t->q_sigma, sg, sizeof sg);
    t->hk_table = Homker::build();
    t->lastResult = 2;
    return t;
}

void destroy_tracker(Tracker *t)
{
    if (!t)
        return;
    for (int l = 1; l < MAX_LEVELS; l++)
        if (t->pyr[l])
            cudaFree(t->pyr[l]);
    if (t->partials)
        cudaFree(t->partials);
    if (t->barrier)
        cudaFree(t->barrier);
    if (t->state)
        cudaFree(t->state);
    if (t->hostState)
        cudaFreeHost(t->hostState);
    delete[] t->hk_table;
    delete t;
}

void convert_depth(const short *depth_mm, float *depth_f, int W, int H, cudaStream_t st)
{
    int n = W * H;
    GS_COUNT_LAUNCHES(1);
    k_convert_depth<<<(n + 255) / 256, 256, 0, st>>>(depth_mm, depth_f, n);
}

static void fill_params(Tracker *t, Params &P, const float *depth_f, const float4 *pointsMap, const float4 *normalsMap, float fx, float fy, float cx,
                        float cy, const Mat4 &scenePose, int trackingFrames, cudaStream_t st)
{
    memset(&P, 0, sizeof P);
    P.nLevels = t->nLevels, P.kind = t->kind;
    P.pointsMap = pointsMap, P.normalsMap = normalsMap;
    P.sceneW = t->W, P.sceneH = t->H;
    P.sceneIntr = make_float4(fx, fy, cx, cy);
    P.scenePose = scenePose;
    P.vfmin = t->vfmin, P.vfmax = t->vfmax, P.tukeyCutOff = 8.0f;
    P.framesToSkip = 20, P.framesToWeight = 50;
    P.useWeights = (t->kind == 1 && trackingFrames >= 100) ? 1 : 0; // ITMExtendedTracker_CUDA.cu:141
    P.terminationThreshold = t->kind == 1 ? 1e-4f : 1e-3f;
    P.rank = t->rank, P.world = t->world > 0 ? t->world : 1;
    for (int q = 0; q < 16; q++)
        P.xchg[q] = t->xchg[q];
    P.xseqBase = t->xseq;
    P.err = t->err;
    int w = t->W, h = t->H;
    float4 intr = make_float4(fx, fy, cx, cy);
    const float *prev = depth_f;
    for (int l = 0; l < t->nLevels; l++)
    {
        Level &L = P.lv[l];
        if (l == 0)
            L.depth = const_cast<float *>(depth_f);
        else
        {
            int ow = w;
            w /= 2, h /= 2;
            intr = make_float4(intr.x * 0.5f, intr.y * 0.5f, intr.z * 0.5f, intr.w * 0.5f);
            dim3 b(16, 16), g((w + 15) / 16, (h + 15) / 16);
            GS_COUNT_LAUNCHES(1);
            k_subsample_holes<<<g, b, 0, st>>>(prev, ow, t->pyr[l], w, h);
            L.depth = t->pyr[l];
        }
        L.w = w, L.h = h, L.intr = intr;
        L.type = t->types[l], L.nIter = t->nIter[l], L.thresh = t->thresh[l];
        prev = L.depth;
    }
}

// C5: UpdatePoseQuality (ITMExtendedTracker.cpp:398-468 / ITMDepthTracker.cpp:183-232) on the host, from the device state
static void update_pose_quality(Tracker *t, const State &S, float thresh0)
{
    const size_t total = (size_t)t->W * t->H;
    const int validMax = S.validDepthMax;
    const int nOld = S.n_q;
    const float nf1 = (float)nOld / (float)total, nf2 = (float)nOld / (float)validMax;
    float d1 = 0.f, d2 = 0.f;
    if (S.lastType == IT_BOTH)
    {
        float h[36];
        for (int i = 0; i < 36; i++)
            h[i] = S.hessian_q[i] * nf1;
        d1 = SpdFactor<6>(h).det_squared();
        if (std::isnan(d1))
            d1 = 0.f;
        for (int i = 0; i < 36; i++)
            h[i] = S.hessian_q[i] * nf2;
        d2 = SpdFactor<6>(h).det_squared();
        if (std::isnan(d2))
            d2 = 0.f;
    }
    float residual = sqrtf(((float)nOld * S.f_q + (float)(validMax - nOld) * thresh0) / (float)validMax);
    float inliers = (float)nOld / (float)validMax;
    t->lastResult = 0; // TRACKING_FAILED
    t->lastScore = residual;
    if (validMax != 0 && total != 0 && d1 > 0 && d2 > 0)
    {
        float in[4] = {logf(d1), logf(d2), residual, inliers}, mapped[20];
        for (int j = 0; j < 4; j++)
            Homker::eval(t->hk_table, mapped + 5 * j, (in[j] - t->q_mu[j]) / t->q_sigma[j]);
        float score = t->svm_b;
        for (int i = 0; i < 20; i++)
            score += t->svm_w[i] * mapped[i];
        if (score > 0)
            t->lastResult = 2; // TRACKING_GOOD
        else if (score > -10.0f)
            t->lastResult = 1; // TRACKING_POOR
    }
}

int track_camera(Tracker *t, const float *depth_f, const float4 *pointsMap, const float4 *normalsMap, float fx, float fy, float cx, float cy,
                 const Mat4 &scenePose, int trackingFrames, se3::Pose *pose_d, cudaStream_t st)
{
    Params P;
    fill_params(t, P, depth_f, pointsMap, normalsMap, fx, fy, cx, cy, scenePose, trackingFrames, st);
    State &hs = *t->hostState;
    memset(&hs, 0, sizeof hs);
    hs.pose_d = *pose_d;
    hs.lastGood = *pose_d;
    GS_CUDA_OK(cudaMemcpyAsync(t->state, &hs, sizeof(State), cudaMemcpyHostToDevice, st));
    void *args[] = {(void *)&P, (void *)&t->state, (void *)&t->partials, (void *)&t->barrier};
    GS_COUNT_LAUNCHES(1);
    if (t->kind == 1)
        GS_CUDA_OK(cudaLaunchCooperativeKernel((const void *)k_icp_track<true>, dim3(t->grid), dim3(CTA), args, 0, st));
    else
        GS_CUDA_OK(cudaLaunchCooperativeKernel((const void *)k_icp_track<false>, dim3(t->grid), dim3(CTA), args, 0, st));
    // the one host round trip of the frame: the tracked pose (the engine keeps pose_d on the host, like ITMTrackingState)
    GS_CUDA_OK(cudaMemcpyAsync(&hs, t->state, sizeof(State), cudaMemcpyDeviceToHost, st));
    GS_CUDA_OK(cudaStreamSynchronize(st));
    *pose_d = hs.pose_d;
    t->xseq += (unsigned)hs.iterationsRun;   // one exchange per LM iteration; the same count on every rank
    update_pose_quality(t, hs, t->thresh[0]);
    return 0;
}

void set_shard(Tracker *t, int rank, int world, float *const *xchg, int *err)
{
    t->rank = rank, t->world = world, t->err = err, t->xseq = 0;
    // The persistent kernel now waits for other GPUs in the middle of its run, and those GPUs may be waiting for THIS GPU's Gaussian stream
    // (exchange barriers of the optimiser iterations): the tracker must not occupy the device exclusively, or the two waits close a cycle.
    // One CTA per SM (half the register file) leaves room for the other stream's kernels to run beside it; each rank only accumulates
    // 1 / world of the pixels, so the smaller grid costs nothing.
    if (world > 1)
    {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (t->grid > sms)
            t->grid = sms;
    }
    for (int q = 0; q < 16; q++)
        t->xchg[q] = q < world ? xchg[q] : nullptr;
}

int icp_eval(Tracker *t, const float *depth_f, const float4 *pointsMap, const float4 *normalsMap, float fx, float fy, float cx, float cy,
             const Mat4 &scenePose, int trackingFrames, int level, const Mat4 &approxInvPose, int *nValid, float *f, float *nabla6, float *hessian36,
             cudaStream_t st)
{
    if (level < 0 || level >= t->nLevels)
        return gs_set_error(__FILE__, __LINE__, "bad pyramid level");
    Params P;
    fill_params(t, P, depth_f, pointsMap, normalsMap, fx, fy, cx, cy, scenePose, trackingFrames, st);
    P.rank = 0, P.world = 1;   // a local evaluation over the whole image (the maps are complete on every rank while tracking is on)
    Mat4 pose = approxInvPose;
    void *args[] = {(void *)&P, (void *)&level, (void *)&pose, (void *)&t->state, (void *)&t->partials, (void *)&t->barrier};
    GS_COUNT_LAUNCHES(1);
    if (t->kind == 1)
        GS_CUDA_OK(cudaLaunchCooperativeKernel((const void *)k_icp_eval<true>, dim3(t->grid), dim3(CTA), args, 0, st));
    else
        GS_CUDA_OK(cudaLaunchCooperativeKernel((const void *)k_icp_eval<false>, dim3(t->grid), dim3(CTA), args, 0, st));
    State &hs = *t->hostState;
    GS_CUDA_OK(cudaMemcpyAsync(&hs, t->state, sizeof(State), cudaMemcpyDeviceToHost, st));
    GS_CUDA_OK(cudaStreamSynchronize(st));
    const int type = t->types[level];
    const int noPara = type == IT_BOTH ? 6 : 3;
    expand_hessian(hs.eval, noPara, hessian36, nabla6);
    *nValid = (int)hs.eval[0];
    // Extended returns the raw sum, Depth returns f / n or 1e5 (ITMExtendedTracker_CUDA.cu:194, ITMDepthTracker_CUDA.cu:108)
    *f = t->kind == 1 ? hs.eval[1] : (*nValid > 100 ? hs.eval[1] / (float)*nValid : 1e5f);
    return 0;
}

const float *level_depth(Tracker *t, int level, int *w, int *h)
{
    int ww = t->W, hh = t->H;
    for (int l = 0; l < level; l++)
        ww /= 2, hh /= 2;
    *w = ww, *h = hh;
    return level == 0 ? nullptr : t->pyr[level];
}

void tracker_result(Tracker *t, int *result, float *score, int *iterations)
{
    *result = t->lastResult;
    *score = t->lastScore;
    *iterations = t->hostState->iterationsRun;
}

} // namespace icp
