// distCUDA2 (reference gsplat/rasterizer/simple_knn.cu:151-239; called at src/raw_gs_param.cpp:28 to initialise the Gaussian scales):
// for every point the mean of the squared distances to its 3 nearest OTHER points.  The result is a property of the point set, not
// of the search structure, so the reference's Morton sort + 1024-point boxes (two blocking cudaMemcpy, a cudaMalloc and four thrust
// vectors per call) is replaced by a uniform grid built with one counting sort and an expanding-ring query:
//
//   k_knn_bounds   bounding box (ordered-uint atomics)                     k_knn_grid     grid geometry from the box, on the device
//   k_knn_count    points per cell (+ the cell of each point)              k_scan_*       exclusive scan over the cells
//   k_knn_scatter  points copied into cell order as float4 (xyz, id)       k_knn_query    ring r = 1, 2, ...: stop once the 3rd best
//                                                                                         distance is inside the searched cube
//
// No host round trip, no allocation: everything lives in a workspace the engine allocates once.  The squared distance is written
// with the FMA contraction the reference's updateKBest (simple_knn.cu:137-138) compiles to, so equal point pairs give bit-equal
// distances; the 3 smallest are kept ascending and averaged as (b0 + b1 + b2) / 3.0f (:187).
// With fewer than 4 points the missing neighbours stay at FLT_MAX, as in the reference.
#include <float.h>

#include "common.cuh"
#include "gs.h"

namespace gs
{

namespace
{
constexpr int KNN_MAX_AXIS = 127;            // cells per axis (+1 for the far boundary) -> at most 128^3 cells
constexpr int KNN_SCAN_BLOCK = 1024;         // elements per scan block
constexpr int KNN_SCAN_BLOCKS = KNN_MAX_CELLS / KNN_SCAN_BLOCK;   // 2048

struct KnnGrid
{
    float minx, miny, minz, cell, invCell;
    int gx, gy, gz;
};

// workspace carve-up (all 16-byte aligned)
struct KnnWs
{
    unsigned *bounds;   // [8] ordered-uint min xyz, max xyz
    KnnGrid *grid;      // [1]
    int *cellCount;     // [KNN_MAX_CELLS]   counts, then scatter cursors
    int *cellStart;     // [KNN_MAX_CELLS+1]
    int *blockSums;     // [KNN_SCAN_BLOCKS]
    int *cellOf;        // [P]
    float4 *sorted;     // [P]
};

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

KnnWs carve(void *base, int P)
{
    char *p = (char *)base;
    KnnWs w;
    w.bounds = (unsigned *)p, p += 64;
    w.grid = (KnnGrid *)p, p += 64;
    w.cellCount = (int *)p, p += (size_t)KNN_MAX_CELLS * 4;
    w.cellStart = (int *)p, p += align16((size_t)(KNN_MAX_CELLS + 1) * 4);
    w.blockSums = (int *)p, p += (size_t)KNN_SCAN_BLOCKS * 4;
    w.cellOf = (int *)p, p += align16((size_t)P * 4);
    w.sorted = (float4 *)p;
    return w;
}

__device__ __forceinline__ unsigned ordered(float f)
{
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unordered(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

__global__ void k_knn_init(unsigned *bounds)
{
    if (threadIdx.x < 3)
        bounds[threadIdx.x] = 0xffffffffu;
    else if (threadIdx.x < 6)
        bounds[threadIdx.x] = 0u;
}

__global__ void __launch_bounds__(256) k_knn_bounds(int P, const float *__restrict__ pts, unsigned *bounds)
{
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x)
#pragma unroll
        for (int a = 0; a < 3; a++)
        {
            float v = pts[(size_t)i * 3 + a];
            lo[a] = fminf(lo[a], v), hi[a] = fmaxf(hi[a], v);
        }
#pragma unroll
    for (int a = 0; a < 3; a++)
    {
#pragma unroll
        for (int o = 16; o; o >>= 1)
        {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
        if ((threadIdx.x & 31) == 0)
        {
            atomicMin(&bounds[a], ordered(lo[a]));
            atomicMax(&bounds[3 + a], ordered(hi[a]));
        }
    }
}

// cubic cells sized so that a surface-like point set leaves a handful of points per occupied cell
__global__ void k_knn_grid(int P, const unsigned *bounds, KnnGrid *grid)
{
    float lo[3], ext[3], big = 0.f;
    for (int a = 0; a < 3; a++)
    {
        lo[a] = unordered(bounds[a]);
        ext[a] = unordered(bounds[3 + a]) - lo[a];
        big = fmaxf(big, ext[a]);
    }
    int target = (int)ceilf(cbrtf((float)P) * 1.5f);
    target = max(1, min(KNN_MAX_AXIS, target));
    float cell = big / (float)target;
    if (!(cell > 0.f))
        cell = 1.f;
    KnnGrid g;
    g.minx = lo[0], g.miny = lo[1], g.minz = lo[2], g.cell = cell, g.invCell = 1.f / cell;
    g.gx = min(KNN_MAX_AXIS + 1, (int)(ext[0] * g.invCell) + 1);
    g.gy = min(KNN_MAX_AXIS + 1, (int)(ext[1] * g.invCell) + 1);
    g.gz = min(KNN_MAX_AXIS + 1, (int)(ext[2] * g.invCell) + 1);
    *grid = g;
}

__device__ __forceinline__ void cell_of(const KnnGrid &g, float x, float y, float z, int &cx, int &cy, int &cz)
{
    cx = min(g.gx - 1, max(0, (int)((x - g.minx) * g.invCell)));
    cy = min(g.gy - 1, max(0, (int)((y - g.miny) * g.invCell)));
    cz = min(g.gz - 1, max(0, (int)((z - g.minz) * g.invCell)));
}

__global__ void __launch_bounds__(256) k_knn_count(int P, const float *__restrict__ pts, const KnnGrid *grid, int *cellCount, int *cellOf)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P)
        return;
    const KnnGrid g = *grid;
    int cx, cy, cz;
    cell_of(g, pts[(size_t)i * 3], pts[(size_t)i * 3 + 1], pts[(size_t)i * 3 + 2], cx, cy, cz);
    const int c = (cx * g.gy + cy) * g.gz + cz;
    cellOf[i] = c;
    atomicAdd(&cellCount[c], 1);
}

// ---- exclusive scan over KNN_MAX_CELLS counts: per-block sums, scan of the sums, per-block scan + offset
__global__ void __launch_bounds__(256) k_scan_sums(const int *__restrict__ in, int *blockSums)
{
    __shared__ int warpSum[8];
    const int4 v = reinterpret_cast<const int4 *>(in)[blockIdx.x * 256 + threadIdx.x];
    int s = v.x + v.y + v.z + v.w;
#pragma unroll
    for (int o = 16; o; o >>= 1)
        s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0)
        warpSum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        int t = 0;
        for (int w = 0; w < 8; w++)
            t += warpSum[w];
        blockSums[blockIdx.x] = t;
    }
}

// one block of 1024 threads, two sums each: exclusive scan in place; total -> *total
__global__ void __launch_bounds__(1024) k_scan_block_sums(int *blockSums, int *total)
{
    __shared__ int warpTot[32];
    const int t = threadIdx.x;
    const int a = blockSums[2 * t], b = blockSums[2 * t + 1];
    int incl = a + b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        int n = __shfl_up_sync(0xffffffffu, incl, o);
        if ((t & 31) >= o)
            incl += n;
    }
    if ((t & 31) == 31)
        warpTot[t >> 5] = incl;
    __syncthreads();
    if (t < 32)
    {
        int w = warpTot[t], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            int n = __shfl_up_sync(0xffffffffu, wi, o);
            if (t >= o)
                wi += n;
        }
        warpTot[t] = wi - w;
    }
    __syncthreads();
    const int excl = incl - (a + b) + warpTot[t >> 5];
    blockSums[2 * t] = excl, blockSums[2 * t + 1] = excl + a;
    if (t == 1023)
        *total = excl + a + b;
}

__global__ void __launch_bounds__(256) k_scan_apply(const int *__restrict__ in, const int *__restrict__ blockSums, int *out)
{
    __shared__ int warpTot[8];
    const int t = threadIdx.x;
    const int4 v = reinterpret_cast<const int4 *>(in)[blockIdx.x * 256 + t];
    const int mine = v.x + v.y + v.z + v.w;
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        int n = __shfl_up_sync(0xffffffffu, incl, o);
        if ((t & 31) >= o)
            incl += n;
    }
    if ((t & 31) == 31)
        warpTot[t >> 5] = incl;
    __syncthreads();
    int before = blockSums[blockIdx.x];
    for (int w = 0; w < (t >> 5); w++)
        before += warpTot[w];
    int4 o4;
    o4.x = before + incl - mine, o4.y = o4.x + v.x, o4.z = o4.y + v.y, o4.w = o4.z + v.z;
    reinterpret_cast<int4 *>(out)[blockIdx.x * 256 + t] = o4;
}

__global__ void __launch_bounds__(256) k_knn_scatter(int P, const float *__restrict__ pts, const int *__restrict__ cellOf,
                                                      const int *__restrict__ cellStart, int *cursor, float4 *sorted)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P)
        return;
    const int c = cellOf[i];
    const int pos = cellStart[c] + atomicAdd(&cursor[c], 1);
    sorted[pos] = make_float4(pts[(size_t)i * 3], pts[(size_t)i * 3 + 1], pts[(size_t)i * 3 + 2], __int_as_float(i));
}

__device__ __forceinline__ void visit(const float4 *__restrict__ sorted, int from, int to, float px, float py, float pz, int self, float *best)
{
    for (int j = from; j < to; j++)
    {
        const float4 q = sorted[j];
        if (__float_as_int(q.w) == self)
            continue;
        // the reference's d.x*d.x + d.y*d.y + d.z*d.z with d = other - point (simple_knn.cu:137-138), written with the FMA contraction
        // nvcc gives that expression in the reference's kernel (SASS of boxMeanDist: FMUL y,y; FFMA x,x; FFMA z,z) so that the
        // distances, and with them the output, are bit-equal to the reference's
        const float dx = q.x - px, dy = q.y - py, dz = q.z - pz;
        float dist = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
#pragma unroll
        for (int k = 0; k < 3; k++)
            if (best[k] > dist)
            {
                const float t = best[k];
                best[k] = dist, dist = t;
            }
    }
}

// thread = one point, taken in cell order so that a warp walks the same neighbourhood
__global__ void __launch_bounds__(128) k_knn_query(int P, const float4 *__restrict__ sorted, const int *__restrict__ cellStart,
                                                    const KnnGrid *grid, float *out)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= P)
        return;
    const KnnGrid g = *grid;
    const float4 me = sorted[s];
    const int self = __float_as_int(me.w);
    int cx, cy, cz;
    cell_of(g, me.x, me.y, me.z, cx, cy, cz);
    float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    const int rMax = max(max(max(cx, g.gx - 1 - cx), max(cy, g.gy - 1 - cy)), max(cz, g.gz - 1 - cz));
    for (int r = 0; r <= rMax; r++)
    {
        const int x0 = max(0, cx - r), x1 = min(g.gx - 1, cx + r), y0 = max(0, cy - r), y1 = min(g.gy - 1, cy + r);
        const int z0 = max(0, cz - r), z1 = min(g.gz - 1, cz + r);
        for (int x = x0; x <= x1; x++)
            for (int y = y0; y <= y1; y++)
            {
                const int row = (x * g.gy + y) * g.gz;
                const bool rim = (x == cx - r) || (x == cx + r) || (y == cy - r) || (y == cy + r);
                if (rim || r == 0)
                    visit(sorted, cellStart[row + z0], cellStart[row + z1 + 1], me.x, me.y, me.z, self, best);   // cells z0..z1 are contiguous
                else
                {
                    if (cz - r >= 0)
                        visit(sorted, cellStart[row + cz - r], cellStart[row + cz - r + 1], me.x, me.y, me.z, self, best);
                    if (cz + r <= g.gz - 1)
                        visit(sorted, cellStart[row + cz + r], cellStart[row + cz + r + 1], me.x, me.y, me.z, self, best);
                }
            }
        // every point outside the searched cube is farther than r cells (minus rounding slack) from this one
        const float reach = (float)r * g.cell * 0.999f;
        if (best[2] <= reach * reach)
            break;
    }
    out[self] = (best[0] + best[1] + best[2]) / 3.0f;
}
} // namespace

size_t knn_workspace_bytes(int maxPoints)
{
    return 64 + 64 + (size_t)KNN_MAX_CELLS * 4 + align16((size_t)(KNN_MAX_CELLS + 1) * 4) + (size_t)KNN_SCAN_BLOCKS * 4 +
           align16((size_t)maxPoints * 4) + (size_t)maxPoints * 16;
}

void knn_mean_dist3(int P, const float *points, float *meanDist, void *workspace, cudaStream_t st)
{
    KnnWs w = carve(workspace, P);
    k_knn_init<<<1, 32, 0, st>>>(w.bounds);
    cudaMemsetAsync(w.cellCount, 0, (size_t)KNN_MAX_CELLS * 4, st);
    const int blocks = (P + 255) / 256;
    k_knn_bounds<<<min(blocks, 148 * 8), 256, 0, st>>>(P, points, w.bounds);
    k_knn_grid<<<1, 1, 0, st>>>(P, w.bounds, w.grid);
    k_knn_count<<<blocks, 256, 0, st>>>(P, points, w.grid, w.cellCount, w.cellOf);
    k_scan_sums<<<KNN_SCAN_BLOCKS, 256, 0, st>>>(w.cellCount, w.blockSums);
    k_scan_block_sums<<<1, 1024, 0, st>>>(w.blockSums, w.cellStart + KNN_MAX_CELLS);
    k_scan_apply<<<KNN_SCAN_BLOCKS, 256, 0, st>>>(w.cellCount, w.blockSums, w.cellStart);
    cudaMemsetAsync(w.cellCount, 0, (size_t)KNN_MAX_CELLS * 4, st);
    k_knn_scatter<<<blocks, 256, 0, st>>>(P, points, w.cellOf, w.cellStart, w.cellCount, w.sorted);
    k_knn_query<<<(P + 127) / 128, 128, 0, st>>>(P, w.sorted, w.cellStart, w.grid, meanDist);
    GS_COUNT_LAUNCHES(9);
}

} // namespace gs
