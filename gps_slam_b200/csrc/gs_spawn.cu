// Gaussian spawn + raycast/frame glue (SURVEY.md section 8 row A13 and "next" rows f1/f2), hand-written for sm_100a.
//
//   k_raycast_maps   runRaycastByCam's tensor glue (slam/slam_pipeline.cpp:386-403, src/cv_utils.cpp:322-341,
//                    src/tensor_math.cpp:69-81): uchar4 colour / 255, vertex * [conf > 0] * voxel_size, depth = z of w2c * vertex,
//                    zero where the vertex is (0,0,0) -- one kernel instead of ~15 torch kernels per raycast.
//   k_frame_to_float Camera::image / depth as float (src/dataset_reader.cpp:269-369 + Camera::toGPU): rgb / 255, mm / 1000.
//   k_spawn_select   initNewGaussians masks (slam/slam_pipeline.cpp:450-526): colour error of the current render (or of the
//                    TSDF colour when there is no Gaussian yet) > color_error_thres, valid raycast depth, weight sum < alpha_vis_max,
//                    then the sampling of addGaussians (slam/slam_gs_model.cpp:22-32).
//   k_knn_*          distCUDA2 (gsplat/rasterizer/simple_knn.cu:151-239): mean squared distance to the 3 nearest new points.  Only
//                    min(max_init_scale, sqrt(.)) is consumed (src/raw_gs_param.cpp:28), so neighbours further than
//                    sqrt(3) * max_init_scale cannot change the result: a uniform hash grid of that cell size is exact.
//   k_spawn_write    RawGaussianParams::init (src/raw_gs_param.cpp:11-74): scales, z axis shrunk by 0.1 and aligned with the
//                    normal (computeNormalMap, src/tensor_math.cpp:278-300: Sobel of the vertex map; computeQuat :184-201),
//                    DC colour from the frame (rgb2sh, gsplat/gsplat_wapper.cpp:127-133), zero higher SH, logit(default opacity).
#include <cfloat>

#include "common.cuh"
#include "gs.h"
#include "gs_spawn.h"

namespace gs
{

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

__global__ void __launch_bounds__(256) k_raycast_maps(int P, const float4 *__restrict__ vertex4, const uchar4 *__restrict__ colour4, Mat4 w2cRow,
                                                       float voxelSize, float *__restrict__ depthMap, float *__restrict__ colorMap,
                                                       float *__restrict__ confMap)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P)
        return;
    float4 v = vertex4[i];
    float ok = v.w > 0.f ? 1.0f : 0.0f;
    float x = v.x * ok * voxelSize, y = v.y * ok * voxelSize, z = v.z * ok * voxelSize;
    const float *m = w2cRow.m; // row-major 4x4
    float zc = m[8] * x + m[9] * y + m[10] * z + m[11];
    float wc = m[12] * x + m[13] * y + m[14] * z + m[15];
    float d = zc / wc;
    if (x + y + z == 0.f)
        d = 0.f;
    depthMap[i] = d;
    uchar4 c = colour4[i];
    colorMap[i * 3 + 0] = (float)c.x / 255.0f;
    colorMap[i * 3 + 1] = (float)c.y / 255.0f;
    colorMap[i * 3 + 2] = (float)c.z / 255.0f;
    if (confMap)
        confMap[i] = v.w;
}

__global__ void __launch_bounds__(256) k_frame_to_float(int P, const uchar4 *__restrict__ rgba, const short *__restrict__ depth_mm,
                                                         float *__restrict__ rgb, float *__restrict__ depth)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P)
        return;
    uchar4 c = rgba[i];
    rgb[i * 3 + 0] = (float)c.x / 255.0f;
    rgb[i * 3 + 1] = (float)c.y / 255.0f;
    rgb[i * 3 + 2] = (float)c.z / 255.0f;
    if (depth)
        depth[i] = (float)depth_mm[i] / 1000.0f;
}

__device__ __forceinline__ unsigned hash_u32(unsigned x)
{
    x ^= x >> 16;
    x *= 0x7feb352dU;
    x ^= x >> 15;
    x *= 0x846ca68bU;
    x ^= x >> 16;
    return x;
}

__device__ __forceinline__ float3 world_vertex(const float4 *vertex4, int i, float voxelSize)
{
    float4 v = __ldg(&vertex4[i]);
    float ok = v.w > 0.f ? 1.0f : 0.0f;
    return make_float3(v.x * ok * voxelSize, v.y * ok * voxelSize, v.z * ok * voxelSize);
}

// flags[i] = 1 when pixel i spawns a Gaussian
__global__ void __launch_bounds__(256) k_spawn_select(SpawnParams sp, const float4 *__restrict__ vertex4, const float *__restrict__ depthMap,
                                                       const float *__restrict__ colorMap, const float *__restrict__ gt,
                                                       const float *__restrict__ renderRgb, const float *__restrict__ renderAlpha,
                                                       const int *__restrict__ nDev, unsigned char *flags)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= sp.P)
        return;
    const bool haveGs = sp.forceRender || *nDev > 0; // with no Gaussian the render equals the TSDF colour and alpha is 0
    float d = depthMap[i];
    float3 v = world_vertex(vertex4, i, sp.voxelSize);
    bool valid = (d > sp.depthMin) && (d < sp.depthMax) && !(v.x + v.y + v.z == 0.f);
    const float *src = haveGs ? renderRgb : colorMap;
    float err = (fabsf(src[i * 3 + 0] - gt[i * 3 + 0]) + fabsf(src[i * 3 + 1] - gt[i * 3 + 1]) + fabsf(src[i * 3 + 2] - gt[i * 3 + 2])) / 3.0f;
    bool m = valid && (err > sp.colorErrorThres);
    if (haveGs)
        m = m && (renderAlpha[i] < sp.alphaMax);
    // multi-GPU: the Gaussian set is sharded by spatial block (4 cm cells, the TSDF block size); each rank spawns its own.  Pixels
    // that another rank spawns are kept as flag 2: they are neighbours for the distCUDA2 scale of this rank's new Gaussians (the
    // mask, the sampling and the ownership are pure functions of the pixel, so every rank sees the same full set).
    bool mine = true;
    if (m && sp.world > 1)
    {
        int bx = (int)floorf(v.x * 25.0f), by = (int)floorf(v.y * 25.0f), bz = (int)floorf(v.z * 25.0f);
        unsigned hsh = ((unsigned)bx * 73856093u) ^ ((unsigned)by * 19349669u) ^ ((unsigned)bz * 83492791u);
        mine = (int)(hash_u32(hsh) % (unsigned)sp.world) == sp.rank;
    }
    // addGaussians keeps a uniformly random subset of the masked pixels (randperm prefix of length ratio * M); here every masked
    // pixel is kept independently with probability ratio (counter-based hash of pixel and seed): same inclusion probability,
    // deterministic, no sort.
    if (m)
    {
        unsigned h = hash_u32((unsigned)i ^ hash_u32(sp.seed));
        m = h < sp.ratioThreshold;
    }
    flags[i] = m ? (mine ? 1 : 2) : 0;
}

// stable compaction of the flagged pixels (two-level scan, 1024 pixels per chunk)
__device__ __forceinline__ int block_excl_scan_1024s(int v, int *ws, int &total)
{
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        int n = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d)
            incl += n;
    }
    if (lane == 31)
        ws[wid] = incl;
    __syncthreads();
    if (wid == 0)
    {
        int w = ws[lane], wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            int n = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d)
                wi += n;
        }
        ws[lane] = wi - w;
        if (lane == 31)
            ws[32] = wi;
    }
    __syncthreads();
    total = ws[32];
    int r = ws[wid] + incl - v;
    __syncthreads();
    return r;
}

// anyOwner = 0: the pixels this rank spawns (flag 1); 1: the pixels any rank spawns (flag 1 or 2)
__device__ __forceinline__ int flag_selected(unsigned char f, int anyOwner) { return anyOwner ? (f != 0) : (f == 1); }

__global__ void __launch_bounds__(1024) k_spawn_count(int P, const unsigned char *__restrict__ flags, int anyOwner, int *chunkCnt)
{
    __shared__ int ws[33];
    int i = blockIdx.x * 1024 + threadIdx.x;
    int f = i < P ? flag_selected(flags[i], anyOwner) : 0;
    int total;
    block_excl_scan_1024s(f, ws, total);
    if (threadIdx.x == 0)
        chunkCnt[blockIdx.x] = total;
}

// slot: counters[CNT_SCRATCH] (number of new Gaussians, limited by the room left) or counters[CNT_SCRATCH + 1] (size of the KNN point set)
__global__ void __launch_bounds__(1024) k_spawn_scan(int *chunkCnt, int nChunks, const int *nDev, int cap, int *counters, int slot)
{
    __shared__ int ws[33];
    __shared__ int carry;
    if (threadIdx.x == 0)
        carry = 0;
    __syncthreads();
    for (int base = 0; base < nChunks; base += 1024)
    {
        int i = base + threadIdx.x;
        int v = i < nChunks ? chunkCnt[i] : 0;
        int total;
        int ex = block_excl_scan_1024s(v, ws, total);
        if (i < nChunks)
            chunkCnt[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0)
            carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0)
    {
        int m = carry;
        if (slot == CNT_SCRATCH)
        {
            int room = cap - *nDev;
            if (m > room)
            {
                m = room < 0 ? 0 : room;
                atomicOr(&counters[CNT_OVERFLOW], 4);
            }
        }
        counters[slot] = m;
    }
}

__global__ void __launch_bounds__(1024) k_spawn_compact(int P, const unsigned char *__restrict__ flags, int anyOwner, const int *__restrict__ chunkOff,
                                                         const int *__restrict__ counters, int slot, int *pixOf)
{
    __shared__ int ws[33];
    int i = blockIdx.x * 1024 + threadIdx.x;
    int f = i < P ? flag_selected(flags[i], anyOwner) : 0;
    int total;
    int ex = block_excl_scan_1024s(f, ws, total);
    if (f)
    {
        int d = chunkOff[blockIdx.x] + ex;
        if (d < counters[slot])
            pixOf[d] = i;
    }
}

// ---- KNN over the new points: uniform hash grid (cell = sqrt(3) * max_init_scale), chained cells
__device__ __forceinline__ unsigned long long cell_key(int cx, int cy, int cz)
{
    return ((unsigned long long)(cx & 0x1fffff) << 42) | ((unsigned long long)(cy & 0x1fffff) << 21) | (unsigned long long)(cz & 0x1fffff);
}
__device__ __forceinline__ unsigned cell_hash(unsigned long long k)
{
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return (unsigned)k;
}
constexpr unsigned long long KEY_EMPTY = 0xffffffffffffffffULL;

__global__ void __launch_bounds__(256) k_knn_build(const int *__restrict__ counters, int countSlot, const int *__restrict__ pixOf,
                                                    const float4 *__restrict__ vertex4, float voxelSize, float invCell, unsigned long long *keys,
                                                    int *heads, int *next, unsigned mask)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= counters[countSlot])
        return;
    float3 v = world_vertex(vertex4, pixOf[i], voxelSize);
    unsigned long long key = cell_key((int)floorf(v.x * invCell), (int)floorf(v.y * invCell), (int)floorf(v.z * invCell));
    unsigned slot = cell_hash(key) & mask;
    for (unsigned probe = 0; probe <= mask; probe++)
    {
        unsigned long long old = atomicCAS(&keys[slot], KEY_EMPTY, key);
        if (old == KEY_EMPTY || old == key)
        {
            next[i] = atomicExch(&heads[slot], i);
            return;
        }
        slot = (slot + 1) & mask;
    }
}

// pixOf: the pixels this rank spawns; pixKnn: the point set of the KNN grid (the same list on one GPU, every rank's pixels otherwise)
__global__ void __launch_bounds__(128) k_spawn_write(SpawnParams sp, const int *__restrict__ counters, const int *__restrict__ pixOf,
                                                      const int *__restrict__ pixKnn, const float4 *__restrict__ vertex4,
                                                      const float *__restrict__ gt, float invCell,
                                                      const unsigned long long *__restrict__ keys, const int *__restrict__ heads,
                                                      const int *__restrict__ next, unsigned mask, ParamPtrs p, const int *__restrict__ nDev,
                                                      unsigned char *touched)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= counters[CNT_SCRATCH])
        return;
    const int pix = pixOf[i];
    const int W = sp.W, H = sp.H;
    const int y = pix / W, x = pix - y * W;
    float3 v = world_vertex(vertex4, pix, sp.voxelSize);
    // --- distCUDA2: mean of the 3 smallest squared distances to the other new points
    float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    int cx = (int)floorf(v.x * invCell), cy = (int)floorf(v.y * invCell), cz = (int)floorf(v.z * invCell);
    for (int dz = -1; dz <= 1; dz++)
        for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++)
            {
                unsigned long long key = cell_key(cx + dx, cy + dy, cz + dz);
                unsigned slot = cell_hash(key) & mask;
                int head = -1;
                for (unsigned probe = 0; probe <= mask; probe++)
                {
                    unsigned long long k = keys[slot];
                    if (k == key)
                    {
                        head = heads[slot];
                        break;
                    }
                    if (k == KEY_EMPTY)
                        break;
                    slot = (slot + 1) & mask;
                }
                for (int j = head; j >= 0; j = next[j])
                {
                    const int pj = pixKnn[j];
                    if (pj == pix)
                        continue;
                    float3 u = world_vertex(vertex4, pj, sp.voxelSize);
                    float ex = u.x - v.x, ey = u.y - v.y, ez = u.z - v.z;
                    float dist = ex * ex + ey * ey + ez * ez;
#pragma unroll
                    for (int k = 0; k < 3; k++)
                        if (best[k] > dist)
                        {
                            float t = best[k];
                            best[k] = dist;
                            dist = t;
                        }
                }
            }
    float meanD2 = (best[0] + best[1] + best[2]) / 3.0f;
    float s = fminf(fmaxf(sqrtf(meanD2), sp.minScale), sp.maxScale); // torch.clamp(min, max)
    // --- normal: Sobel of the vertex map with replicate padding, cross(dy, dx), / (|.| + 1e-8), zero where vertex.z <= 0
    float3 nb[3][3];
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++)
        {
            int yy = min(max(y + a - 1, 0), H - 1), xx = min(max(x + b - 1, 0), W - 1);
            nb[a][b] = world_vertex(vertex4, yy * W + xx, sp.voxelSize);
        }
    float gx[3], gy[3];
#define COMP(q, c) ((c) == 0 ? (q).x : ((c) == 1 ? (q).y : (q).z))
#pragma unroll
    for (int c = 0; c < 3; c++)
    {
        gx[c] = (COMP(nb[0][2], c) - COMP(nb[0][0], c)) + 2.f * (COMP(nb[1][2], c) - COMP(nb[1][0], c)) + (COMP(nb[2][2], c) - COMP(nb[2][0], c));
        gy[c] = (COMP(nb[2][0], c) - COMP(nb[0][0], c)) + 2.f * (COMP(nb[2][1], c) - COMP(nb[0][1], c)) + (COMP(nb[2][2], c) - COMP(nb[0][2], c));
    }
#undef COMP
    float nx = gy[1] * gx[2] - gy[2] * gx[1];
    float ny = gy[2] * gx[0] - gy[0] * gx[2];
    float nz = gy[0] * gx[1] - gy[1] * gx[0];
    float mag = sqrtf(nx * nx + ny * ny + nz * nz) + 1e-8f;
    nx /= mag, ny /= mag, nz /= mag;
    if (v.z <= 0.f)
        nx = ny = nz = 0.f;
    // --- computeQuat((0,0,1), n): axis = z x n, angle = acos(z . n)
    float ax = -ny, ay = nx, az = 0.f;
    float an = sqrtf(ax * ax + ay * ay + az * az) + 1e-8f;
    ax /= an, ay /= an, az /= an;
    float angle = acosf(nz);
    float an2 = sqrtf(ax * ax + ay * ay + az * az) + 1e-8f;
    float half = angle / 2.f;
    float sh = sinf(half);
    const int g = *nDev + i;
    p.means[g * 3 + 0] = v.x, p.means[g * 3 + 1] = v.y, p.means[g * 3 + 2] = v.z;
    p.scales[g * 3 + 0] = logf(s), p.scales[g * 3 + 1] = logf(s), p.scales[g * 3 + 2] = logf(s * 0.1f);
    reinterpret_cast<float4 *>(p.quats)[g] = make_float4(cosf(half), ax / an2 * sh, ay / an2 * sh, az / an2 * sh);
    const float C0 = 0.28209479177387814f;
#pragma unroll
    for (int c = 0; c < 3; c++)
        p.dc[g * 3 + c] = (gt[pix * 3 + c] - 0.5f) / C0;
    for (int e = 0; e < 45; e++)
        p.rest[(size_t)g * 45 + e] = 0.f;
    p.opac[g] = logf(sp.defaultOpacity / (1.0f - sp.defaultOpacity));
    touched[g] = 0;
}

__global__ void k_spawn_commit(int *nDev, const int *counters)
{
    *nDev += counters[CNT_SCRATCH];
}

// ------------------------------------------------------------------------------------------------------------
void raycast_maps(int P, const float4 *vertex4, const uchar4 *colour4, const float *w2cRowMajor, float voxelSize, float *depthMap, float *colorMap,
                  float *confMap, cudaStream_t st)
{
    Mat4 m;
    for (int i = 0; i < 16; i++)
        m.m[i] = w2cRowMajor[i];
    GS_COUNT_LAUNCHES(1);
    k_raycast_maps<<<cdiv(P, 256), 256, 0, st>>>(P, vertex4, colour4, m, voxelSize, depthMap, colorMap, confMap);
}

void frame_to_float(int P, const uchar4 *rgba, const short *depth_mm, float *rgb, float *depth, cudaStream_t st)
{
    GS_COUNT_LAUNCHES(1);
    k_frame_to_float<<<cdiv(P, 256), 256, 0, st>>>(P, rgba, depth_mm, rgb, depth);
}

void spawn(const SpawnParams &sp, const SpawnBuffers &b, const float4 *vertex4, const float *depthMap, const float *colorMap, const float *gt,
           const float *renderRgb, const float *renderAlpha, const ParamPtrs &p, int *nDev, int cap, unsigned char *touched, int *counters,
           cudaStream_t st)
{
    const int P = sp.P;
    const int nChunks = cdiv(P, 1024);
    const float cell = sqrtf(3.0f) * sp.maxScale * 1.001f;
    const float invCell = 1.0f / cell;
    GS_COUNT_LAUNCHES(7);
    k_spawn_select<<<cdiv(P, 256), 256, 0, st>>>(sp, vertex4, depthMap, colorMap, gt, renderRgb, renderAlpha, nDev, b.flags);
    k_spawn_count<<<nChunks, 1024, 0, st>>>(P, b.flags, 0, b.chunkCnt);
    k_spawn_scan<<<1, 1024, 0, st>>>(b.chunkCnt, nChunks, nDev, cap, counters, CNT_SCRATCH);
    k_spawn_compact<<<nChunks, 1024, 0, st>>>(P, b.flags, 0, b.chunkCnt, counters, CNT_SCRATCH, b.pixOf);
    // the KNN point set: this rank's new points on one GPU, every rank's otherwise (so that a new Gaussian next to a block border
    // gets the scale it would get on one GPU)
    const int *pixKnn = b.pixOf;
    int knnSlot = CNT_SCRATCH;
    if (sp.world > 1)
    {
        GS_COUNT_LAUNCHES(3);
        k_spawn_count<<<nChunks, 1024, 0, st>>>(P, b.flags, 1, b.chunkCnt);
        k_spawn_scan<<<1, 1024, 0, st>>>(b.chunkCnt, nChunks, nDev, cap, counters, CNT_SCRATCH + 1);
        k_spawn_compact<<<nChunks, 1024, 0, st>>>(P, b.flags, 1, b.chunkCnt, counters, CNT_SCRATCH + 1, b.pixAll);
        pixKnn = b.pixAll, knnSlot = CNT_SCRATCH + 1;
    }
    cudaMemsetAsync(b.keys, 0xff, sizeof(unsigned long long) * ((size_t)b.tableMask + 1), st);
    cudaMemsetAsync(b.heads, 0xff, sizeof(int) * ((size_t)b.tableMask + 1), st);
    k_knn_build<<<cdiv(P, 256), 256, 0, st>>>(counters, knnSlot, pixKnn, vertex4, sp.voxelSize, invCell, b.keys, b.heads, b.next, b.tableMask);
    k_spawn_write<<<cdiv(P, 128), 128, 0, st>>>(sp, counters, b.pixOf, pixKnn, vertex4, gt, invCell, b.keys, b.heads, b.next, b.tableMask, p, nDev,
                                                touched);
    k_spawn_commit<<<1, 1, 0, st>>>(nDev, counters);
}

} // namespace gs
