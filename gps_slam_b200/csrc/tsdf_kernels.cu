// Hashed-voxel TSDF path (SURVEY.md section 8 rows B1-B9), hand-written for sm_100a.
//
// This translation unit is compiled with -fmad=false: every float expression below is written in the
// same operation order as the reference's device-agnostic "Shared" math so that results are bit-identical
// to the reference CPU engine built with -ffp-contract=off (oracle/_ref/libitm_ref_exact.so).
// The kernels are HBM/L2/latency bound, so giving up FMA contraction costs nothing measurable.
//
// What is different from the reference's CUDA drivers (by design, not by accident):
//  * no host round trip anywhere: visible-entry count, free-list heads and rendering ranges stay on the device,
//    all launches are fixed-shape (persistent / grid-stride over device-side counters) -> graph-capturable;
//  * allocation is DETERMINISTIC (ascending hash-slot order via a two-level prefix sum), which reproduces the
//    reference CPU engine's pointers exactly (its CUDA engine pops the free list with atomicSub in arbitrary order);
//  * the winning block of a same-frame slot collision is the last writer in raster order (= the single-threaded
//    CPU loop), arbitrated with atomicMax on a (pixel, step) key instead of a racy byte store;
//  * expected-depth ranges are splatted directly from the block list (min/max are order independent) instead
//    of materialising up to 262,144 RenderingBlocks and a host-synchronised count;
//  * the depth short->float conversion (B1) is fused into the allocation pass;
//  * voxel blocks are streamed through shared memory with TMA bulk copies in the integrate kernel.
#include <stdlib.h>
#include <type_traits>

#include "common.cuh"
#include "tsdf.h"
#include "tsdf_access.cuh"

namespace tsdf
{

// ------------------------------------------------------------------------------------------------------------
// B4 (first half): entries visible in the previous frame become type 3 ("re-test against the frustum")
// reference: setToType3, ITMSceneReconstructionEngine_CUDA.tcu:398-404 / CPU.tpp:170-171
__global__ void k_set_type3(const int *__restrict__ visIds, const int *__restrict__ nVis, unsigned char *visType)
{
    int n = *nVis;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        visType[visIds[i]] = 3;
}

// ------------------------------------------------------------------------------------------------------------
// B1 + B2: depth short(mm) -> float(m) and per-pixel block marking.
// reference: convertDepthAffineToFloat (ITMViewBuilder_Shared.h:27-36), buildHashAllocAndVisibleTypePP
// (ITMSceneReconstructionEngine_Shared.h:207-323)
// allocKey[slot] = max over requesting (pixel, step) of 1 + pixel*16 + step  (0 = no request)
__global__ void __launch_bounds__(256) k_alloc_flags(const short *__restrict__ depth_mm, float *__restrict__ depth_f, int W, int H,
                                                      Mat4 invM, float4 invProj /* 1/fx, 1/fy, cx, cy */, float2 calib, float mu,
                                                      float oneOverBlock, float vfmin, float vfmax,
                                                      const HashEntry *__restrict__ table, unsigned char *visType, unsigned *allocKey,
                                                      int *errFlag)
{
    int loc = blockIdx.x * blockDim.x + threadIdx.x;
    if (loc >= W * H)
        return;
    int y = loc / W, x = loc - y * W;
    short dmm = depth_mm[loc];
    float d = dmm <= 0 ? -1.0f : (float)dmm * calib.x + calib.y;
    depth_f[loc] = d;
    if (d <= 0 || (d - mu) < 0 || (d - mu) < vfmin || (d + mu) > vfmax)
        return;

    float pz = d;
    float px = pz * ((float(x) - invProj.z) * invProj.x);
    float py = pz * ((float(y) - invProj.w) * invProj.y);
    float norm = sqrtf(px * px + py * py + pz * pz);

    float s0 = 1.0f - mu / norm;
    float3 p = mat4_mul_point(invM, px * s0, py * s0, pz * s0, 1.0f);
    p.x *= oneOverBlock, p.y *= oneOverBlock, p.z *= oneOverBlock;
    float s1 = 1.0f + mu / norm;
    float3 pe = mat4_mul_point(invM, px * s1, py * s1, pz * s1, 1.0f);
    pe.x *= oneOverBlock, pe.y *= oneOverBlock, pe.z *= oneOverBlock;

    float3 dir = make_float3(pe.x - p.x, pe.y - p.y, pe.z - p.z);
    norm = sqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
    int noSteps = (int)ceilf(2.0f * norm);
    float den = (float)(noSteps - 1);
    dir.x /= den, dir.y /= den, dir.z /= den;
    if (noSteps > 16)
    {
        atomicExch(errFlag, 1); // key encoding holds 16 steps per pixel (mu <= 8 blocks)
        noSteps = 16;
    }

    for (int i = 0; i < noSteps; i++)
    {
        short bx = (short)floorf(p.x), by = (short)floorf(p.y), bz = (short)floorf(p.z);
        int idx = hash_index(bx, by, bz);
        HashEntry e = load_entry(table, idx);
        bool found = false;
        if (e.px == bx && e.py == by && e.pz == bz && e.ptr >= -1)
        {
            visType[idx] = (e.ptr == -1) ? 2 : 1;
            found = true;
        }
        if (!found)
        {
            if (e.ptr >= -1)
            {
                while (e.offset >= 1)
                {
                    idx = SDF_BUCKET_NUM + e.offset - 1;
                    e = load_entry(table, idx);
                    if (e.px == bx && e.py == by && e.pz == bz && e.ptr >= -1)
                    {
                        visType[idx] = (e.ptr == -1) ? 2 : 1;
                        found = true;
                        break;
                    }
                }
            }
            if (!found)
                atomicMax(&allocKey[idx], 1u + (unsigned)loc * 16u + (unsigned)i);
        }
        p.x += dir.x, p.y += dir.y, p.z += dir.z;
    }
}

// recompute the block a (pixel, step) key refers to -- same arithmetic as k_alloc_flags
__device__ __forceinline__ void block_of_key(unsigned key, const float *__restrict__ depth_f, int W, const Mat4 &invM, float4 invProj,
                                             float mu, float oneOverBlock, short &bx, short &by, short &bz)
{
    unsigned k = key - 1u;
    int loc = (int)(k >> 4), step = (int)(k & 15u);
    int y = loc / W, x = loc - y * W;
    float d = depth_f[loc];
    float pz = d;
    float px = pz * ((float(x) - invProj.z) * invProj.x);
    float py = pz * ((float(y) - invProj.w) * invProj.y);
    float norm = sqrtf(px * px + py * py + pz * pz);
    float s0 = 1.0f - mu / norm;
    float3 p = mat4_mul_point(invM, px * s0, py * s0, pz * s0, 1.0f);
    p.x *= oneOverBlock, p.y *= oneOverBlock, p.z *= oneOverBlock;
    float s1 = 1.0f + mu / norm;
    float3 pe = mat4_mul_point(invM, px * s1, py * s1, pz * s1, 1.0f);
    pe.x *= oneOverBlock, pe.y *= oneOverBlock, pe.z *= oneOverBlock;
    float3 dir = make_float3(pe.x - p.x, pe.y - p.y, pe.z - p.z);
    norm = sqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
    int noSteps = (int)ceilf(2.0f * norm);
    float den = (float)(noSteps - 1);
    dir.x /= den, dir.y /= den, dir.z /= den;
    for (int i = 0; i < step; i++)
        p.x += dir.x, p.y += dir.y, p.z += dir.z;
    bx = (short)floorf(p.x), by = (short)floorf(p.y), bz = (short)floorf(p.z);
}

// ------------------------------------------------------------------------------------------------------------
// B3: deterministic allocation.  Pass 1 counts requests per 1024-slot chunk, pass 2 ranks them.
// reference semantics: CPU.tpp:196-265 (ascending targetIdx; vba index = lastFree--, excess index = lastFreeExcess--)
constexpr int SCAN_CTA = 1024;

__device__ __forceinline__ int2 block_excl_scan2(int a, int b, int2 &total)
{
    // exclusive scan of two int streams across a 1024-thread CTA
    __shared__ int2 warpSums[32];
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int ia = a, ib = b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        int ta = __shfl_up_sync(0xffffffffu, ia, o), tb = __shfl_up_sync(0xffffffffu, ib, o);
        if (lane >= o)
            ia += ta, ib += tb;
    }
    if (lane == 31)
        warpSums[wid] = make_int2(ia, ib);
    __syncthreads();
    if (wid == 0)
    {
        int2 w = warpSums[lane];
        int wa = w.x, wb = w.y;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            int ta = __shfl_up_sync(0xffffffffu, wa, o), tb = __shfl_up_sync(0xffffffffu, wb, o);
            if (lane >= o)
                wa += ta, wb += tb;
        }
        warpSums[lane] = make_int2(wa, wb);
    }
    __syncthreads();
    int2 base = wid > 0 ? warpSums[wid - 1] : make_int2(0, 0);
    total = warpSums[31];
    __syncthreads();
    return make_int2(base.x + ia - a, base.y + ib - b);
}

// sum of chunkCounts[0 .. cta) computed cooperatively by the CTA (<= 1152 chunks)
__device__ __forceinline__ int2 chunk_prefix(const int2 *__restrict__ chunkCounts, int cta)
{
    __shared__ int2 red[32];
    int a = 0, b = 0;
    for (int i = threadIdx.x; i < cta; i += blockDim.x)
    {
        int2 c = chunkCounts[i];
        a += c.x, b += c.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        a += __shfl_xor_sync(0xffffffffu, a, o), b += __shfl_xor_sync(0xffffffffu, b, o);
    if ((threadIdx.x & 31) == 0)
        red[threadIdx.x >> 5] = make_int2(a, b);
    __syncthreads();
    if (threadIdx.x < 32)
    {
        int2 r = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : make_int2(0, 0);
        a = r.x, b = r.y;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            a += __shfl_xor_sync(0xffffffffu, a, o), b += __shfl_xor_sync(0xffffffffu, b, o);
        if (threadIdx.x == 0)
            red[0] = make_int2(a, b);
    }
    __syncthreads();
    int2 r = red[0];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_CTA) k_alloc_count(const unsigned *__restrict__ allocKey, const HashEntry *__restrict__ table,
                                                           int E, int2 *chunkCounts)
{
    int slot = blockIdx.x * SCAN_CTA + threadIdx.x;
    int any = 0, exc = 0;
    if (slot < E && allocKey[slot] != 0u)
    {
        any = 1;
        bool ordered = slot < SDF_BUCKET_NUM && load_entry(table, slot).ptr < -1;
        exc = ordered ? 0 : 1;
    }
    int a = __syncthreads_count(any);
    int b = __syncthreads_count(exc);
    if (threadIdx.x == 0)
        chunkCounts[blockIdx.x] = make_int2(a, b);
}

// state[0] = lastFreeBlockId, state[1] = lastFreeExcessListId (device resident); stateOut = values after this frame.
__global__ void __launch_bounds__(SCAN_CTA) k_alloc_apply(unsigned *allocKey, HashEntry *table, unsigned char *visType, int E, int nChunks,
                                                           const int2 *__restrict__ chunkCounts, const int *__restrict__ state, int *stateOut,
                                                           const float *__restrict__ depth_f, int W, Mat4 invM, float4 invProj, float mu,
                                                           float oneOverBlock)
{
    int slot = blockIdx.x * SCAN_CTA + threadIdx.x;
    unsigned key = slot < E ? allocKey[slot] : 0u;
    int any = key != 0u, exc = 0;
    if (any)
    {
        bool ordered = slot < SDF_BUCKET_NUM && table[slot].ptr < -1;
        exc = ordered ? 0 : 1;
    }
    int2 pre = chunk_prefix(chunkCounts, blockIdx.x);
    int2 tot;
    int2 rank = block_excl_scan2(any, exc, tot);
    const int lastFreeBlock = state[0], lastFreeExcess = state[1];
    if (any)
    {
        allocKey[slot] = 0u;
        // Exhaustion: the reference restores its counters when a pop fails; ranks are computed as if every earlier
        // request succeeded, so behaviour matches whenever the pools do not run dry (and fails closed otherwise).
        int vbaIdx = lastFreeBlock - (pre.x + rank.x);
        short bx, by, bz;
        block_of_key(key, depth_f, W, invM, invProj, mu, oneOverBlock, bx, by, bz);
        HashEntry ne;
        ne.px = bx, ne.py = by, ne.pz = bz, ne.pad = 0;
        ne.offset = 0;
        ne.ptr = vbaIdx; // voxelAllocationList is the identity: blocks are never freed (swapping is out of scope)
        if (!exc)
        {
            if (vbaIdx >= 0)
                table[slot] = ne;
            else
                visType[slot] = 0;
            if (vbaIdx >= 0)
                visType[slot] = 1; // "new entry is visible" (set by the marking pass in the reference)
        }
        else
        {
            int exlIdx = lastFreeExcess - (pre.y + rank.y);
            if (vbaIdx >= 0 && exlIdx >= 0)
            {
                int exlOffset = exlIdx; // excessAllocationList is the identity as well
                table[slot].offset = exlOffset + 1;
                table[SDF_BUCKET_NUM + exlOffset] = ne;
                visType[SDF_BUCKET_NUM + exlOffset] = 1;
            }
        }
    }
    if (blockIdx.x == nChunks - 1 && threadIdx.x == 0)
    {
        int usedA = pre.x + tot.x, usedB = pre.y + tot.y;
        int nb = lastFreeBlock - usedA, ne2 = lastFreeExcess - usedB;
        stateOut[4] = nb < -1 ? -1 : nb; // committed to state[0..1] by k_visible_count (other CTAs still read state[0..1])
        stateOut[5] = ne2 < -1 ? -1 : ne2;
    }
}

// ------------------------------------------------------------------------------------------------------------
// B4: visible list (ascending slot order, like the CPU engine)
// reference: checkBlockVisibility<false> (Reconstruction_Shared.h:325-422), CPU.tpp:268-310
__device__ __forceinline__ bool point_visible(const Mat4 &M, float4 proj, int W, int H, float x, float y, float z)
{
    float3 p = mat4_mul_point(M, x, y, z, 1.0f);
    if (p.z < 1e-10f)
        return false;
    float u = proj.x * p.x / p.z + proj.z;
    float v = proj.y * p.y / p.z + proj.w;
    return u >= 0 && u < W && v >= 0 && v < H;
}

__device__ __forceinline__ bool block_visible(short bx, short by, short bz, const Mat4 &M, float4 proj, float voxelSize, int W, int H)
{
    float factor = (float)SDF_BLOCK_SIZE * voxelSize;
    float x = (float)bx * factor, y = (float)by * factor, z = (float)bz * factor;
    if (point_visible(M, proj, W, H, x, y, z)) return true;   // 0 0 0
    z += factor;
    if (point_visible(M, proj, W, H, x, y, z)) return true;   // 0 0 1
    y += factor;
    if (point_visible(M, proj, W, H, x, y, z)) return true;   // 0 1 1
    x += factor;
    if (point_visible(M, proj, W, H, x, y, z)) return true;   // 1 1 1
    z -= factor;
    if (point_visible(M, proj, W, H, x, y, z)) return true;   // 1 1 0
    y -= factor;
    if (point_visible(M, proj, W, H, x, y, z)) return true;   // 1 0 0
    x -= factor;
    y += factor;
    if (point_visible(M, proj, W, H, x, y, z)) return true;   // 0 1 0
    x += factor;
    y -= factor;
    z += factor;
    return point_visible(M, proj, W, H, x, y, z);             // 1 0 1
}

// world > 1: the second scanned stream counts / ranks the visible entries whose voxel block this rank owns (visIdsOwn, state[6])
__global__ void __launch_bounds__(SCAN_CTA) k_visible_count(unsigned char *visType, const HashEntry *__restrict__ table, int E, Mat4 M,
                                                             float4 proj, float voxelSize, int W, int H, int2 *chunkCounts, int *state,
                                                             int rank, int world)
{
    int slot = blockIdx.x * SCAN_CTA + threadIdx.x;
    int vis = 0, own = 0;
    if (slot == 0)
    {
        state[0] = state[4];
        state[1] = state[5];
    }
    if (slot < E)
    {
        unsigned char t = visType[slot];
        if (t == 3)
        {
            HashEntry e = load_entry_cg(table, slot);
            if (!block_visible(e.px, e.py, e.pz, M, proj, voxelSize, W, H))
            {
                t = 0;
                visType[slot] = 0;
            }
        }
        vis = t > 0;
        if (vis && world > 1)
        {
            HashEntry e = load_entry_cg(table, slot);
            own = block_owner(e.px, e.py, e.pz, world) == rank;
        }
    }
    int a = __syncthreads_count(vis);
    int b = world > 1 ? __syncthreads_count(own) : 0;
    if (threadIdx.x == 0)
        chunkCounts[blockIdx.x] = make_int2(a, b);
}

__global__ void __launch_bounds__(SCAN_CTA) k_visible_compact(const unsigned char *__restrict__ visType, int E, int nChunks,
                                                               const int2 *__restrict__ chunkCounts, int *visIds, int cap, int *nVis,
                                                               const HashEntry *__restrict__ table, int myRank, int world, int *visIdsOwn,
                                                               int *nVisOwn)
{
    int slot = blockIdx.x * SCAN_CTA + threadIdx.x;
    int vis = slot < E && visType[slot] > 0;
    int own = 0;
    if (vis && world > 1)
    {
        HashEntry e = load_entry(table, slot);
        own = block_owner(e.px, e.py, e.pz, world) == myRank;
    }
    int2 pre = chunk_prefix(chunkCounts, blockIdx.x);
    int2 tot;
    int2 rank = block_excl_scan2(vis, own, tot);
    int pos = pre.x + rank.x;
    if (vis && pos < cap)
        visIds[pos] = slot;
    if (own && pre.y + rank.y < cap)
        visIdsOwn[pre.y + rank.y] = slot;
    if (blockIdx.x == nChunks - 1 && threadIdx.x == 0)
    {
        int n = pre.x + tot.x;
        *nVis = n < cap ? n : cap;
        if (world > 1)
        {
            int m = pre.y + tot.y;
            *nVisOwn = m < cap ? m : cap;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// B5: integrate.  One voxel block (512 voxels, 4 KB) per CTA iteration, persistent grid over the device-side
// visible list.  Blocks move HBM -> smem -> HBM with TMA bulk copies (cp.async.bulk, UBLKCP in SASS) through a
// 2-deep ring so the load of block i+1 overlaps the math of block i and the store of block i-1.
// reference: integrateIntoScene_device (Reconstruction_CUDA.tcu:348-383), computeUpdatedVoxelDepthInfo /
// computeUpdatedVoxelColorInfo (Reconstruction_Shared.h:8-54, 105-140)
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile("{\n.reg .pred p;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT_LOOP;\nDONE:\n}" ::"r"(
                     smem_u32(bar)),
                 "r"(parity)
                 : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dstSmem, const void *srcGmem, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dstSmem)),
                 "l"(srcGmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_store_1d(void *dstGmem, const void *srcSmem, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dstGmem), "r"(smem_u32(srcSmem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- exact fp32 division by the small set of divisors integrate_voxel meets, without the ~12-instruction IEEE division sequence.
// For a divisor c with r = RN(1/c):  q = x * r;  q' = fma(fma(-q, c, x), r, q)  is the correctly rounded x / c.  Proven by exhaustion on the
// CPU (IEEE fmaf, every fp32 bit pattern of x): c = 255 and c = 32767 for ALL x except -0 and +-inf (unreachable here: the numerators
// are sums of non-negative products or integer conversions); every integer c in 1..256 for 2^-80 <= |x| < 2^11 (outside that range the
// plain division is used).  The quotients, and with them the voxels, stay bit-identical to the reference's.
__device__ __forceinline__ float div_255(float x)
{
    const float q = __fmul_rn(x, 0x1.010102p-8f);
    return __fmaf_rn(__fmaf_rn(-q, 255.0f, x), 0x1.010102p-8f, q);
}
__device__ __forceinline__ float div_32767(float x)
{
    const float q = __fmul_rn(x, 0x1.0002p-15f);
    return __fmaf_rn(__fmaf_rn(-q, 32767.0f, x), 0x1.0002p-15f, q);
}
// RN(1 / w) for w = 0..256 (entry 0 unused)
__device__ const float c_rcpInt[257] = {0.0f, 0x1.0000000000000p+0f, 0x1.0000000000000p-1f, 0x1.5555560000000p-2f, 0x1.0000000000000p-2f, 0x1.99999a0000000p-3f, 0x1.5555560000000p-3f, 0x1.24924a0000000p-3f, 0x1.0000000000000p-3f, 0x1.c71c720000000p-4f, 0x1.99999a0000000p-4f, 0x1.745d180000000p-4f, 0x1.5555560000000p-4f, 0x1.3b13b20000000p-4f, 0x1.24924a0000000p-4f, 0x1.1111120000000p-4f, 0x1.0000000000000p-4f, 0x1.e1e1e20000000p-5f, 0x1.c71c720000000p-5f, 0x1.af286c0000000p-5f, 0x1.99999a0000000p-5f, 0x1.8618620000000p-5f, 0x1.745d180000000p-5f, 0x1.642c860000000p-5f, 0x1.5555560000000p-5f, 0x1.47ae140000000p-5f, 0x1.3b13b20000000p-5f, 0x1.2f684c0000000p-5f, 0x1.24924a0000000p-5f, 0x1.1a7b960000000p-5f, 0x1.1111120000000p-5f, 0x1.0842100000000p-5f, 0x1.0000000000000p-5f, 0x1.f07c200000000p-6f, 0x1.e1e1e20000000p-6f, 0x1.d41d420000000p-6f, 0x1.c71c720000000p-6f, 0x1.bacf920000000p-6f, 0x1.af286c0000000p-6f, 0x1.a41a420000000p-6f, 0x1.99999a0000000p-6f, 0x1.8f9c180000000p-6f, 0x1.8618620000000p-6f, 0x1.7d05f40000000p-6f, 0x1.745d180000000p-6f, 0x1.6c16c20000000p-6f, 0x1.642c860000000p-6f, 0x1.5c98820000000p-6f, 0x1.5555560000000p-6f, 0x1.4e5e0a0000000p-6f, 0x1.47ae140000000p-6f, 0x1.4141420000000p-6f, 0x1.3b13b20000000p-6f, 0x1.3521d00000000p-6f, 0x1.2f684c0000000p-6f, 0x1.29e4120000000p-6f, 0x1.24924a0000000p-6f, 0x1.1f70480000000p-6f, 0x1.1a7b960000000p-6f, 0x1.15b1e60000000p-6f, 0x1.1111120000000p-6f, 0x1.0c97140000000p-6f, 0x1.0842100000000p-6f, 0x1.0410420000000p-6f, 0x1.0000000000000p-6f, 0x1.f81f820000000p-7f, 0x1.f07c200000000p-7f, 0x1.e9131a0000000p-7f, 0x1.e1e1e20000000p-7f, 0x1.dae6080000000p-7f, 0x1.d41d420000000p-7f, 0x1.cd85680000000p-7f, 0x1.c71c720000000p-7f, 0x1.c0e0700000000p-7f, 0x1.bacf920000000p-7f, 0x1.b4e81c0000000p-7f, 0x1.af286c0000000p-7f, 0x1.a98ef60000000p-7f, 0x1.a41a420000000p-7f, 0x1.9ec8ea0000000p-7f, 0x1.99999a0000000p-7f, 0x1.948b100000000p-7f, 0x1.8f9c180000000p-7f, 0x1.8acb900000000p-7f, 0x1.8618620000000p-7f, 0x1.8181820000000p-7f, 0x1.7d05f40000000p-7f, 0x1.78a4c80000000p-7f, 0x1.745d180000000p-7f, 0x1.702e060000000p-7f, 0x1.6c16c20000000p-7f, 0x1.6816820000000p-7f, 0x1.642c860000000p-7f, 0x1.6058160000000p-7f, 0x1.5c98820000000p-7f, 0x1.58ed240000000p-7f, 0x1.5555560000000p-7f, 0x1.51d07e0000000p-7f, 0x1.4e5e0a0000000p-7f, 0x1.4afd6a0000000p-7f, 0x1.47ae140000000p-7f, 0x1.446f860000000p-7f, 0x1.4141420000000p-7f, 0x1.3e22cc0000000p-7f, 0x1.3b13b20000000p-7f, 0x1.3813820000000p-7f, 0x1.3521d00000000p-7f, 0x1.323e340000000p-7f, 0x1.2f684c0000000p-7f, 0x1.2c9fb40000000p-7f, 0x1.29e4120000000p-7f, 0x1.27350c0000000p-7f, 0x1.24924a0000000p-7f, 0x1.21fb780000000p-7f, 0x1.1f70480000000p-7f, 0x1.1cf06a0000000p-7f, 0x1.1a7b960000000p-7f, 0x1.1811820000000p-7f, 0x1.15b1e60000000p-7f, 0x1.135c820000000p-7f, 0x1.1111120000000p-7f, 0x1.0ecf560000000p-7f, 0x1.0c97140000000p-7f, 0x1.0a68100000000p-7f, 0x1.0842100000000p-7f, 0x1.0624de0000000p-7f, 0x1.0410420000000p-7f, 0x1.0204080000000p-7f, 0x1.0000000000000p-7f, 0x1.fc07f00000000p-8f, 0x1.f81f820000000p-8f, 0x1.f4465a0000000p-8f, 0x1.f07c200000000p-8f, 0x1.ecc07c0000000p-8f, 0x1.e9131a0000000p-8f, 0x1.e573ac0000000p-8f, 0x1.e1e1e20000000p-8f, 0x1.de5d6e0000000p-8f, 0x1.dae6080000000p-8f, 0x1.d77b660000000p-8f, 0x1.d41d420000000p-8f, 0x1.d0cb580000000p-8f, 0x1.cd85680000000p-8f, 0x1.ca4b300000000p-8f, 0x1.c71c720000000p-8f, 0x1.c3f8f00000000p-8f, 0x1.c0e0700000000p-8f, 0x1.bdd2b80000000p-8f, 0x1.bacf920000000p-8f, 0x1.b7d6c40000000p-8f, 0x1.b4e81c0000000p-8f, 0x1.b203640000000p-8f, 0x1.af286c0000000p-8f, 0x1.ac57020000000p-8f, 0x1.a98ef60000000p-8f, 0x1.a6d01a0000000p-8f, 0x1.a41a420000000p-8f, 0x1.a16d400000000p-8f, 0x1.9ec8ea0000000p-8f, 0x1.9c2d140000000p-8f, 0x1.99999a0000000p-8f, 0x1.970e500000000p-8f, 0x1.948b100000000p-8f, 0x1.920fb40000000p-8f, 0x1.8f9c180000000p-8f, 0x1.8d30180000000p-8f, 0x1.8acb900000000p-8f, 0x1.886e600000000p-8f, 0x1.8618620000000p-8f, 0x1.83c9780000000p-8f, 0x1.8181820000000p-8f, 0x1.7f40600000000p-8f, 0x1.7d05f40000000p-8f, 0x1.7ad2200000000p-8f, 0x1.78a4c80000000p-8f, 0x1.767dce0000000p-8f, 0x1.745d180000000p-8f, 0x1.7242880000000p-8f, 0x1.702e060000000p-8f, 0x1.6e1f760000000p-8f, 0x1.6c16c20000000p-8f, 0x1.6a13ce0000000p-8f, 0x1.6816820000000p-8f, 0x1.661ec60000000p-8f, 0x1.642c860000000p-8f, 0x1.623fa80000000p-8f, 0x1.6058160000000p-8f, 0x1.5e75bc0000000p-8f, 0x1.5c98820000000p-8f, 0x1.5ac0560000000p-8f, 0x1.58ed240000000p-8f, 0x1.571ed40000000p-8f, 0x1.5555560000000p-8f, 0x1.5390940000000p-8f, 0x1.51d07e0000000p-8f, 0x1.5015020000000p-8f, 0x1.4e5e0a0000000p-8f, 0x1.4cab880000000p-8f, 0x1.4afd6a0000000p-8f, 0x1.49539e0000000p-8f, 0x1.47ae140000000p-8f, 0x1.460cbc0000000p-8f, 0x1.446f860000000p-8f, 0x1.42d6620000000p-8f, 0x1.4141420000000p-8f, 0x1.3fb0140000000p-8f, 0x1.3e22cc0000000p-8f, 0x1.3c995a0000000p-8f, 0x1.3b13b20000000p-8f, 0x1.3991c20000000p-8f, 0x1.3813820000000p-8f, 0x1.3698e00000000p-8f, 0x1.3521d00000000p-8f, 0x1.33ae460000000p-8f, 0x1.323e340000000p-8f, 0x1.30d1900000000p-8f, 0x1.2f684c0000000p-8f, 0x1.2e025c0000000p-8f, 0x1.2c9fb40000000p-8f, 0x1.2b404a0000000p-8f, 0x1.29e4120000000p-8f, 0x1.288b020000000p-8f, 0x1.27350c0000000p-8f, 0x1.25e2280000000p-8f, 0x1.24924a0000000p-8f, 0x1.2345680000000p-8f, 0x1.21fb780000000p-8f, 0x1.20b4700000000p-8f, 0x1.1f70480000000p-8f, 0x1.1e2ef40000000p-8f, 0x1.1cf06a0000000p-8f, 0x1.1bb4a40000000p-8f, 0x1.1a7b960000000p-8f, 0x1.1945380000000p-8f, 0x1.1811820000000p-8f, 0x1.16e0680000000p-8f, 0x1.15b1e60000000p-8f, 0x1.1485f00000000p-8f, 0x1.135c820000000p-8f, 0x1.12358e0000000p-8f, 0x1.1111120000000p-8f, 0x1.0fef020000000p-8f, 0x1.0ecf560000000p-8f, 0x1.0db20a0000000p-8f, 0x1.0c97140000000p-8f, 0x1.0b7e6e0000000p-8f, 0x1.0a68100000000p-8f, 0x1.0953f40000000p-8f, 0x1.0842100000000p-8f, 0x1.0732600000000p-8f, 0x1.0624de0000000p-8f, 0x1.0519800000000p-8f, 0x1.0410420000000p-8f, 0x1.03091c0000000p-8f, 0x1.0204080000000p-8f, 0x1.0101020000000p-8f, 0x1.0000000000000p-8f};
// x / (float)w for an integer 1 <= w <= 256; rcp = shared-memory copy of c_rcpInt
__device__ __forceinline__ float div_int(float x, int w, const float *rcp)
{
    const float c = (float)w;
    const float ax = fabsf(x);
    if (!(ax >= 0x1p-80f && ax < 0x1p11f))
        return ax == 0.0f ? x : x / c; // +-0 / c = +-0 (black pixels, empty voxels); anything else out of the proven range: plain division
    const float r = rcp[w];
    const float q = __fmul_rn(x, r);
    return __fmaf_rn(__fmaf_rn(-q, c, x), r, q);
}

struct IntegrateParams
{
    Mat4 M;        // world -> camera
    float4 proj;   // fx fy cx cy
    float mu, voxelSize;
    int maxW, W, H;
};

// returns true when the voxel changed
__device__ __forceinline__ bool integrate_voxel(uint2 &raw, int gx, int gy, int gz, const IntegrateParams &P, const float *__restrict__ depth,
                                                const uchar4 *__restrict__ rgb, const float *rcp /* shared copy of c_rcpInt */)
{
    float mx = (float)gx * P.voxelSize, my = (float)gy * P.voxelSize, mz = (float)gz * P.voxelSize;
    float3 pc = mat4_mul_point(P.M, mx, my, mz, 1.0f);
    if (pc.z <= 0)
        return false;
    float u = P.proj.x * pc.x / pc.z + P.proj.z;
    float v = P.proj.y * pc.y / pc.z + P.proj.w;
    if ((u < 1) || (u > P.W - 2) || (v < 1) || (v > P.H - 2))
        return false;
    float dm = __ldg(&depth[(int)(u + 0.5f) + (int)(v + 0.5f) * P.W]);
    if (dm <= 0.0f)
        return false;
    float eta = dm - pc.z;
    if (eta < -P.mu)
        return false;

    short sdf = (short)(raw.x & 0xffffu);
    int oldW = (int)((raw.x >> 16) & 0xffu);
    float oldF = div_32767((float)sdf);
    float newF = fminf(1.0f, eta / P.mu);
    newF = (float)oldW * oldF + 1.0f * newF;
    int newW = oldW + 1;
    newF = div_int(newF, newW, rcp);
    newW = newW < P.maxW ? newW : P.maxW;
    unsigned usdf = (unsigned)(unsigned short)(short)(newF * 32767.0f);
    unsigned cr = (raw.x >> 24) & 0xffu, cg = raw.y & 0xffu, cb = (raw.y >> 8) & 0xffu, cw = (raw.y >> 16) & 0xffu;

    if (!((eta > P.mu) || (fabsf(eta / P.mu) > 0.25f)))
    {
        // colour: bilinear fetch (ITMPixelUtils.h:11-34) + running average
        int ix = (int)floorf(u), iy = (int)floorf(v);
        float dx = u - (float)ix, dy = v - (float)iy;
        uchar4 a = __ldg(&rgb[ix + iy * P.W]);
        uchar4 b = make_uchar4(0, 0, 0, 0), c = b, d4 = b;
        if (dx != 0) b = __ldg(&rgb[(ix + 1) + iy * P.W]);
        if (dy != 0) c = __ldg(&rgb[ix + (iy + 1) * P.W]);
        if (dx != 0 && dy != 0) d4 = __ldg(&rgb[(ix + 1) + (iy + 1) * P.W]);
        float ox = 1.0f - dx, oy = 1.0f - dy;
        float mr = div_255((float)a.x * ox * oy + (float)b.x * dx * oy + (float)c.x * ox * dy + (float)d4.x * dx * dy);
        float mg = div_255((float)a.y * ox * oy + (float)b.y * dx * oy + (float)c.y * ox * dy + (float)d4.y * dx * dy);
        float mb = div_255((float)a.z * ox * oy + (float)b.z * dx * oy + (float)c.z * ox * dy + (float)d4.z * dx * dy);
        float ow = (float)cw;
        float nr = div_255((float)cr) * ow + mr * 1.0f;
        float ng = div_255((float)cg) * ow + mg * 1.0f;
        float nb = div_255((float)cb) * ow + mb * 1.0f;
        float nw = ow + 1.0f;
        const int iw = (int)cw + 1; // nw as an integer, 1..256
        nr = div_int(nr, iw, rcp), ng = div_int(ng, iw, rcp), nb = div_int(nb, iw, rcp);
        float cap = (float)(unsigned char)P.maxW;
        nw = nw < cap ? nw : cap;
        nr *= 255.0f, ng *= 255.0f, nb *= 255.0f;
        int ir = (int)(nr < 0 ? nr - 0.5f : nr + 0.5f), ig = (int)(ng < 0 ? ng - 0.5f : ng + 0.5f), ib = (int)(nb < 0 ? nb - 0.5f : nb + 0.5f);
        cr = (unsigned)min(max(ir, 0), 255), cg = (unsigned)min(max(ig, 0), 255), cb = (unsigned)min(max(ib, 0), 255);
        cw = (unsigned)(unsigned char)nw;
    }
    uint2 out;
    out.x = usdf | ((unsigned)newW << 16) | (cr << 24);
    out.y = cg | (cb << 8) | (cw << 16) | (raw.y & 0xff000000u);
    bool changed = out.x != raw.x || out.y != raw.y;
    raw = out;
    return changed;
}

constexpr int INT_THREADS = 256;
constexpr int INT_STAGES = 4;
constexpr int INT_DESC = 128; // descriptor ring (two halves of 64)

__global__ void __launch_bounds__(INT_THREADS) k_integrate_tma(Voxel *__restrict__ vba, const HashEntry *__restrict__ table,
                                                                const int *__restrict__ visIds, const int *__restrict__ nVis,
                                                                IntegrateParams P, const float *__restrict__ depth,
                                                                const uchar4 *__restrict__ rgb, const __grid_constant__ PeerVbas push)
{
    __shared__ __align__(128) uint2 buf[INT_STAGES][SDF_BLOCK_SIZE3];
    __shared__ __align__(8) unsigned long long full[INT_STAGES];
    __shared__ int sDirty[INT_STAGES];
    // block descriptors (VBA pointer + block position) of this CTA's next INT_DESC work items, fetched by all threads at once:
    // the two dependent global loads (visible id -> hash entry) are then off the single-thread TMA issue path
    __shared__ int dPtr[INT_DESC];
    __shared__ short4 dPos[INT_DESC];
    __shared__ float sRcp[257];

    const int n = *nVis;
    const int tid = threadIdx.x;
    for (int i = tid; i < 257; i += INT_THREADS)
        sRcp[i] = c_rcpInt[i]; // visible to all threads after the first __syncthreads below
    if (tid == 0)
    {
        for (int s = 0; s < INT_STAGES; s++)
            mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    // work items of this CTA: i = blockIdx.x + k * gridDim.x
    const int myCount = (n > (int)blockIdx.x) ? (n - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    auto load_desc = [&](int kBase) {
        // descriptors kBase .. kBase + INT_DESC/2 - 1 go to the half of the ring selected by (kBase / (INT_DESC/2)) & 1
        int k = kBase + tid;
        if (tid < INT_DESC / 2 && k < myCount)
        {
            int slot = __ldg(&visIds[blockIdx.x + k * gridDim.x]);
            HashEntry e = load_entry(table, slot);
            dPtr[k % INT_DESC] = e.ptr;
            dPos[k % INT_DESC] = make_short4(e.px, e.py, e.pz, 0);
        }
    };
    load_desc(0);
    load_desc(INT_DESC / 2);
    __syncthreads();

    auto issue = [&](int k) {
        // executed by thread 0 only
        int s = k % INT_STAGES;
        int ptr = dPtr[k % INT_DESC];
        if (ptr >= 0)
        {
            mbar_expect_tx(&full[s], SDF_BLOCK_SIZE3 * 8);
            tma_load_1d(&buf[s][0], vba + (size_t)ptr * SDF_BLOCK_SIZE3, SDF_BLOCK_SIZE3 * 8, &full[s]);
        }
        else
            mbar_expect_tx(&full[s], 0); // nothing to load: complete the phase so stage parity stays in step
    };

    if (tid == 0)
        for (int k = 0; k < INT_STAGES - 1 && k < myCount; k++)
            issue(k);
    __syncthreads();

    for (int k = 0; k < myCount; k++)
    {
        int s = k % INT_STAGES;
        // refill the half of the descriptor ring that has just been consumed (items k - INT_DESC/2 .. k - 1)
        if (k > 0 && (k % (INT_DESC / 2)) == 0)
            load_desc(k + INT_DESC / 2);
        // prefetch item k + STAGES-1 into the stage freed by item k-1 (its store must have finished READING smem)
        if (tid == 0)
        {
            int kn = k + INT_STAGES - 1;
            if (kn < myCount)
            {
                tma_store_wait_read<0>();
                issue(kn);
            }
            sDirty[s] = 0;
        }
        __syncthreads();
        int ptr = dPtr[k % INT_DESC];
        mbar_wait(&full[s], (unsigned)((k / INT_STAGES) & 1));
        if (ptr >= 0)
        {
            short4 bp = dPos[k % INT_DESC];
            int gx0 = (int)bp.x * SDF_BLOCK_SIZE, gy0 = (int)bp.y * SDF_BLOCK_SIZE, gz0 = (int)bp.z * SDF_BLOCK_SIZE;
            bool any = false;
#pragma unroll
            for (int j = 0; j < SDF_BLOCK_SIZE3 / INT_THREADS; j++)
            {
                int loc = tid + j * INT_THREADS;
                int z = loc >> 6, y = (loc >> 3) & 7, x = loc & 7;
                uint2 raw = buf[s][loc];
                if (integrate_voxel(raw, gx0 + x, gy0 + y, gz0 + z, P, depth, rgb, sRcp))
                {
                    buf[s][loc] = raw;
                    any = true;
                }
            }
            if (any)
                sDirty[s] = 1; // benign same-value race
            fence_proxy_async();
        }
        __syncthreads();
        if (tid == 0 && ptr >= 0 && sDirty[s])
        {
            tma_store_1d(vba + (size_t)ptr * SDF_BLOCK_SIZE3, &buf[s][0], SDF_BLOCK_SIZE3 * 8);
            // sharded scene, mode 1: the updated block also goes into every other rank's array (one 4 KB bulk store each, over NVLink)
            for (int q = 0; q < push.n; q++)
                tma_store_1d(push.p[q] + (size_t)ptr * SDF_BLOCK_SIZE3, &buf[s][0], SDF_BLOCK_SIZE3 * 8);
            tma_store_commit();
        }
    }
    if (tid == 0)
        tma_store_wait_all();
}

// (Two alternates of this kernel were built, measured equal and removed: one CTA per block with direct loads / stores -- the reference's own
// shape -- and a warp-specialised pipeline, a producer warp for descriptors / TMA loads / write-back and eight consumer warps handing stages
// over through mbarriers.  ncu attributes 23 % of this kernel's stall samples to its two CTA barriers per block, but the freed slots were
// already covered by other warps: the kernel is issue-bound either way.)

// ------------------------------------------------------------------------------------------------------------
// B6: expected depth range image at 1/8 resolution
// reference: ProjectSingleBlock / CreateRenderingBlocks (Visualisation_Shared.h:36-118), fillBlocks_device
#define FAR_AWAY 999999.9f
#define VERY_CLOSE 0.05f

__global__ void k_minmax_init(float2 *minmax, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        minmax[i] = make_float2(FAR_AWAY, VERY_CLOSE);
}

__device__ __forceinline__ void splat_block_range(short bx, short by, short bz, const Mat4 &M, float4 proj, int W, int H, float voxelSize,
                                                  float2 *minmax, int mmW, int mmH)
{
    int ulx = W / 8, uly = H / 8, lrx = -1, lry = -1;
    float zmin = FAR_AWAY, zmax = VERY_CLOSE;
#pragma unroll
    for (int corner = 0; corner < 8; ++corner)
    {
        short tx = bx + ((corner & 1) ? 1 : 0), ty = by + ((corner & 2) ? 1 : 0), tz = bz + ((corner & 4) ? 1 : 0);
        float fx = (float)tx * (float)SDF_BLOCK_SIZE * voxelSize, fy = (float)ty * (float)SDF_BLOCK_SIZE * voxelSize,
              fz = (float)tz * (float)SDF_BLOCK_SIZE * voxelSize;
        float3 p = mat4_mul_point(M, fx, fy, fz, 1.0f);
        if (p.z < 1e-6)
            continue;
        float u = (proj.x * p.x / p.z + proj.z) / 8;
        float v = (proj.y * p.y / p.z + proj.w) / 8;
        if (ulx > floorf(u)) ulx = (int)floorf(u);
        if (lrx < ceilf(u)) lrx = (int)ceilf(u);
        if (uly > floorf(v)) uly = (int)floorf(v);
        if (lry < ceilf(v)) lry = (int)ceilf(v);
        if (zmin > p.z) zmin = p.z;
        if (zmax < p.z) zmax = p.z;
    }
    if (ulx < 0) ulx = 0;
    if (uly < 0) uly = 0;
    if (lrx >= W) lrx = W - 1;
    if (lry >= H) lry = H - 1;
    if (ulx > lrx || uly > lry)
        return;
    if (zmin < VERY_CLOSE) zmin = VERY_CLOSE;
    if (zmax < VERY_CLOSE)
        return;
    // the reference image has full-resolution stride; only columns < ceil(W/8) and rows < ceil(H/8) are ever read back
    if (lrx > mmW - 1) lrx = mmW - 1;
    if (lry > mmH - 1) lry = mmH - 1;
    int zminBits = __float_as_int(zmin), zmaxBits = __float_as_int(zmax); // positive floats order like ints
    for (int y = uly; y <= lry; ++y)
        for (int x = ulx; x <= lrx; ++x)
        {
            int *px = reinterpret_cast<int *>(&minmax[x + y * mmW]);
            atomicMin(px, zminBits);
            atomicMax(px + 1, zmaxBits);
        }
}

__global__ void k_project_visible(const HashEntry *__restrict__ table, const int *__restrict__ visIds, const int *__restrict__ nVis, Mat4 M,
                                  float4 proj, int W, int H, float voxelSize, float2 *minmax, int mmW, int mmH)
{
    int n = *nVis;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        HashEntry e = load_entry(table, visIds[i]);
        if (e.ptr >= 0)
            splat_block_range(e.px, e.py, e.pz, M, proj, W, H, voxelSize, minmax, mmW, mmH);
    }
}

// free view (B9): FindVisibleBlocks (Visualisation_CPU.tpp:36-74) + CreateExpectedDepths fused over the whole table
__global__ void k_project_all(const HashEntry *__restrict__ table, int E, Mat4 M, float4 proj, int W, int H, float voxelSize, float2 *minmax,
                              int mmW, int mmH)
{
    int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= E)
        return;
    HashEntry e = load_entry(table, slot);
    if (e.ptr < 0)
        return;
    if (!block_visible(e.px, e.py, e.pz, M, proj, voxelSize, W, H))
        return;
    splat_block_range(e.px, e.py, e.pz, M, proj, W, H, voxelSize, minmax, mmW, mmH);
}

// ------------------------------------------------------------------------------------------------------------
// B7: raycast.  reference: castRay (Visualisation_Shared.h:122-221), readVoxel / readFromSDF_*
// (ITMRepresentationAccess.h:77-232)
template <class A>
__device__ __forceinline__ unsigned read_voxel_lo(const A &vba, const HashEntry *__restrict__ table, int px, int py, int pz, int &vm,
                                                  VoxelCache &c)
{
    const unsigned h = find_voxel(vba, table, px, py, pz, vm, c);
    if (h == NO_VOXEL)
        return 0x00007fffu; // default voxel: sdf = 32767, w_depth = 0, clr = 0
    return __ldg(reinterpret_cast<const unsigned *>(vba.at(h)));
}
__device__ __forceinline__ float lo_sdf(unsigned lo) { return (float)(short)(lo & 0xffffu); }
__device__ __forceinline__ float lo_w(unsigned lo) { return (float)((lo >> 16) & 0xffu); }

template <class A>
__device__ __forceinline__ float sdf_uninterp(const A &vba, const HashEntry *__restrict__ table, float3 p, int &vm, VoxelCache &c)
{
    int ix = (int)(p.x < 0 ? p.x - 0.5f : p.x + 0.5f), iy = (int)(p.y < 0 ? p.y - 0.5f : p.y + 0.5f), iz = (int)(p.z < 0 ? p.z - 0.5f : p.z + 0.5f);
    unsigned lo = read_voxel_lo(vba, table, ix, iy, iz, vm, c);
    return div_32767(lo_sdf(lo)); // == lo_sdf / 32767.0f for every short (exhaustive proof: tools/div_proof)
}

// The 8 corners of a trilinear read are resolved first (same order and therefore same cache / hash-walk sequence as the reference's
// eight readVoxel calls), then all 8 voxels are requested together, then combined with the reference's arithmetic: one memory round
// trip per interpolated read instead of four (sdf) or eight (colour) chained ones.  NO_VOXEL = corner not allocated.
template <class A>
__device__ __forceinline__ void find_corners(const A &vba, const HashEntry *__restrict__ table, int x, int y, int z, int &vm, VoxelCache &c,
                                             unsigned (&off)[8])
{
    off[0] = find_voxel(vba, table, x, y, z, vm, c);
    if (((x & 7) != 7) && ((y & 7) != 7) && ((z & 7) != 7))
    {
        // all 8 corners lie in the block of corner 0 (2 reads out of 3): the seven remaining readVoxel calls would be cache hits on it
        // (or walk the same empty hash chain) without changing any state, so their addresses follow from the first
#pragma unroll
        for (int k = 1; k < 8; k++)
            off[k] = off[0] != NO_VOXEL ? off[0] + (unsigned)((k & 1) + ((k >> 1) & 1) * SDF_BLOCK_SIZE + (k >> 2) * SDF_BLOCK_SIZE * SDF_BLOCK_SIZE)
                                        : NO_VOXEL;
        return;
    }
#pragma unroll
    for (int k = 1; k < 8; k++)
        off[k] = find_voxel(vba, table, x + (k & 1), y + ((k >> 1) & 1), z + (k >> 2), vm, c);
}

template <bool withConf, class A>
__device__ __forceinline__ float sdf_interp(const A &vba, const HashEntry *__restrict__ table, float3 p, int &vm, VoxelCache &c, float &conf)
{
    float fx = floorf(p.x), fy = floorf(p.y), fz = floorf(p.z);
    float cx = p.x - fx, cy = p.y - fy, cz = p.z - fz;
    unsigned off[8];
    find_corners(vba, table, (int)fx, (int)fy, (int)fz, vm, c, off);
    unsigned lo[8];
#pragma unroll
    for (int k = 0; k < 8; k++)
        lo[k] = off[k] != NO_VOXEL ? __ldg(reinterpret_cast<const unsigned *>(vba.at(off[k]))) : 0x00007fffu; // default voxel: sdf = 32767, w_depth = 0
    float res1, res2, r1c = 0, r2c = 0;
    res1 = (1.0f - cx) * lo_sdf(lo[0]) + cx * lo_sdf(lo[1]);
    if (withConf) r1c = (1.0f - cx) * lo_w(lo[0]) + cx * lo_w(lo[1]);
    res1 = (1.0f - cy) * res1 + cy * ((1.0f - cx) * lo_sdf(lo[2]) + cx * lo_sdf(lo[3]));
    if (withConf) r1c = (1.0f - cy) * r1c + cy * ((1.0f - cx) * lo_w(lo[2]) + cx * lo_w(lo[3]));
    res2 = (1.0f - cx) * lo_sdf(lo[4]) + cx * lo_sdf(lo[5]);
    if (withConf) r2c = (1.0f - cx) * lo_w(lo[4]) + cx * lo_w(lo[5]);
    res2 = (1.0f - cy) * res2 + cy * ((1.0f - cx) * lo_sdf(lo[6]) + cx * lo_sdf(lo[7]));
    if (withConf) r2c = (1.0f - cy) * r2c + cy * ((1.0f - cx) * lo_w(lo[6]) + cx * lo_w(lo[7]));
    vm = 1;
    if (withConf) conf = (1.0f - cz) * r1c + cz * r2c;
    return div_32767((1.0f - cz) * res1 + cz * res2); // |x| <= 32767: inside the proven range of the exact form
}

// readFromSDF_color4u_interpolated (ITMRepresentationAccess.h:344-424) + drawPixelColour (Visualisation_Shared.h:386-396)
template <class A>
__device__ __forceinline__ uchar4 colour_interp(const A &vba, const HashEntry *__restrict__ table, float3 p)
{
    VoxelCache c = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0u};
    float fx = floorf(p.x), fy = floorf(p.y), fz = floorf(p.z);
    float cx = p.x - fx, cy = p.y - fy, cz = p.z - fz;
    float rx = 0, ry = 0, rz = 0, wsum = 0;
    int vm;
    unsigned off[8];
    find_corners(vba, table, (int)fx, (int)fy, (int)fz, vm, c, off);
    uint2 raw[8];
#pragma unroll
    for (int k = 0; k < 8; k++)
        raw[k] = off[k] != NO_VOXEL ? __ldg(reinterpret_cast<const uint2 *>(vba.at(off[k]))) : make_uint2(0u, 0u);
#pragma unroll
    for (int k = 0; k < 8; k++)
    {
        int ox = k & 1, oy = (k >> 1) & 1, oz = (k >> 2) & 1;
        unsigned wc = (raw[k].y >> 16) & 0xffu;   // 0 for a missing corner: skipped like the reference's `continue`
        if (off[k] != NO_VOXEL && wc >= 1u)
        {
            float w = (ox ? cx : (1.0f - cx)) * (oy ? cy : (1.0f - cy)) * (oz ? cz : (1.0f - cz));
            rx += w * (float)((raw[k].x >> 24) & 0xffu);
            ry += w * (float)(raw[k].y & 0xffu);
            rz += w * (float)((raw[k].y >> 8) & 0xffu);
            wsum += w;
        }
    }
    rx /= wsum, ry /= wsum, rz /= wsum;
    rx = div_255(rx), ry = div_255(ry), rz = div_255(rz); // == x / 255.0f (x >= +0 or NaN here; exhaustive proof: tools/div_proof)
    uchar4 o;
    // (uchar) of NaN is UB on the host; x86 cvttss2si and the GPU both give 0 after truncation to 8 bits
    o.x = (unsigned char)(int)(rx * 255.0f);
    o.y = (unsigned char)(int)(ry * 255.0f);
    o.z = (unsigned char)(int)(rz * 255.0f);
    o.w = 255;
    return o;
}

// where the visibility marks of the live raycast go (modifyVisible): the local entriesVisibleType, or -- sharded -- every rank's copy.
// The copies are identical when the raycast starts (the allocation pass is replicated), so a mark that does not change the local byte does
// not change any other rank's either and is not sent; the rare ones that do (a block the ray touches that the depth image did not mark)
// are stored into every rank's array (idempotent byte stores over NVLink, ordered by the barrier that follows the raycast).
struct MarkLocal
{
    unsigned char *visType;
    __device__ __forceinline__ void operator()(int slot) const { visType[slot] = 1; }
};
struct MarkAll
{
    const ShardView *v;
    __device__ __forceinline__ void operator()(int slot) const
    {
        if (*(volatile unsigned char *)(v->visType[v->rank] + slot) == 1)
            return;
        for (int q = 0; q < v->world; q++)
            v->visType[q][slot] = 1;
    }
};
struct MarkNone
{
    __device__ __forceinline__ void operator()(int) const {}
};

struct RayCounters
{
    int nSteps, nMiss, nInterp, nBlock;
};

// castRay for pixel (x, y): pt = hit point in voxel units, conf; returns found
template <bool modifyVisible, bool STATS, class A, class Mark>
__device__ __forceinline__ bool cast_ray(const A &vba, const HashEntry *__restrict__ table, const Mark &mark, int x, int y, const Mat4 &invM,
                                         float4 invProj /* 1/fx 1/fy -cx -cy */, float oneOverVoxelSize, float mu, float2 mm, float3 &pt,
                                         float &conf, RayCounters &cnt)
{
    const int lane = threadIdx.x & 31;
    float stepScale = mu * oneOverVoxelSize;

    float cz = mm.x;
    float cxp = cz * ((float(x) + invProj.z) * invProj.x);
    float cyp = cz * ((float(y) + invProj.w) * invProj.y);
    float totalLength = sqrtf(cxp * cxp + cyp * cyp + cz * cz) * oneOverVoxelSize;
    float3 ps = mat4_mul_point(invM, cxp, cyp, cz, 1.0f);
    ps.x *= oneOverVoxelSize, ps.y *= oneOverVoxelSize, ps.z *= oneOverVoxelSize;

    cz = mm.y;
    cxp = cz * ((float(x) + invProj.z) * invProj.x);
    cyp = cz * ((float(y) + invProj.w) * invProj.y);
    float totalLengthMax = sqrtf(cxp * cxp + cyp * cyp + cz * cz) * oneOverVoxelSize;
    float3 pe = mat4_mul_point(invM, cxp, cyp, cz, 1.0f);
    pe.x *= oneOverVoxelSize, pe.y *= oneOverVoxelSize, pe.z *= oneOverVoxelSize;

    float3 rd = make_float3(pe.x - ps.x, pe.y - ps.y, pe.z - ps.z);
    float dn = 1.0f / sqrtf(rd.x * rd.x + rd.y * rd.y + rd.z * rd.z);
    rd.x *= dn, rd.y *= dn, rd.z *= dn;

    pt = ps;
    VoxelCache cache = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0u};
    float sdf = 1.0f, stepLength;
    conf = 0.0f;
    int vm;
    cnt.nSteps = cnt.nMiss = cnt.nInterp = cnt.nBlock = 0;
    bool hitSlot0 = false;
    while (totalLength < totalLengthMax)
    {
        sdf = sdf_uninterp(vba, table, pt, vm, cache);
        if (STATS)
            cnt.nSteps++, cnt.nMiss += (vm == 0), cnt.nBlock += (vm > 1);
        if (modifyVisible)
        {
            // NB: a cache hit reports vm == 1, i.e. slot 0 -- reference quirk kept (Access.h:86-90).  Nearly every step of every ray is
            // such a hit: the write to slot 0 is remembered and issued once per warp after the march instead of ~6 M times to one address
            // (the L2 slice owning that sector serialised them and every load behind them waited: 5x the long-scoreboard stall of the
            // free-view variant)
            if (vm > 1)
                mark(vm - 1);
            else if (vm == 1)
                hitSlot0 = true;
        }
        if (!vm)
            stepLength = SDF_BLOCK_SIZE;
        else
        {
            if (STATS)
                cnt.nInterp += ((sdf <= 0.1f) && (sdf >= -0.5f));
            if ((sdf <= 0.1f) && (sdf >= -0.5f))
                sdf = sdf_interp<false>(vba, table, pt, vm, cache, conf);
            if (sdf <= 0.0f)
                break;
            stepLength = fmaxf(sdf * stepScale, 1.0f);
        }
        pt.x += stepLength * rd.x, pt.y += stepLength * rd.y, pt.z += stepLength * rd.z;
        totalLength += stepLength;
    }
    if (modifyVisible)
    {
        const unsigned act = __activemask();
        if (__any_sync(act, hitSlot0) && lane == __ffs(act) - 1)
            mark(0);
    }
    if (sdf <= 0.0f)
    {
        stepLength = sdf * stepScale;
        pt.x += stepLength * rd.x, pt.y += stepLength * rd.y, pt.z += stepLength * rd.z;
        sdf = sdf_interp<true>(vba, table, pt, vm, cache, conf);
        stepLength = sdf * stepScale;
        pt.x += stepLength * rd.x, pt.y += stepLength * rd.y, pt.z += stepLength * rd.z;
        return true;
    }
    return false;
}

// STATS (diagnostic build of the same loop, gsb_tsdf_raycast_stats): visType is reinterpreted as unsigned long long[8] totals --
// rays, march steps, steps in unallocated space, trilinear reads, steps that changed voxel block, warp-max steps summed over warps, warps
template <bool modifyVisible, bool withColour, bool STATS = false>
__global__ void __launch_bounds__(256) k_raycast(float4 *__restrict__ pointsRay, uchar4 *__restrict__ colourOut, unsigned char *visType,
                                                  const Voxel *__restrict__ vbaPtr, const HashEntry *__restrict__ table, int W, int H, Mat4 invM,
                                                  float4 invProj /* 1/fx 1/fy -cx -cy */, float oneOverVoxelSize, float mu,
                                                  const float2 *__restrict__ minmax, int mmW)
{
    // 32x8 pixel tile per CTA, one 8x4 pixel patch per warp: the 32 rays of a warp share one cell of the 1/8-resolution range
    // image and march through the same few voxel blocks, so they take nearly the same number of steps (a warp runs for as long
    // as its slowest ray) and their hash / voxel reads hit the same lines; each patch row is still one 128-byte store.
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int x = blockIdx.x * 32 + (wid & 3) * 8 + (lane & 7);
    int y = blockIdx.y * 8 + (wid >> 2) * 4 + (lane >> 3);
    if (x >= W || y >= H)
        return;
    float2 mm = __ldg(&minmax[(x >> 3) + (y >> 3) * mmW]);
    const VbaLocal vba = {vbaPtr};
    const MarkLocal mark = {visType};
    float3 pt;
    float conf;
    RayCounters cnt;
    const bool found = cast_ray<modifyVisible, STATS>(vba, table, mark, x, y, invM, invProj, oneOverVoxelSize, mu, mm, pt, conf, cnt);
    int loc = x + y * W;
    if (STATS)
    {
        unsigned long long *tot = reinterpret_cast<unsigned long long *>(visType);
        const unsigned act = __activemask();
        int warpMax = cnt.nSteps;
        for (int o = 16; o; o >>= 1)
            warpMax = max(warpMax, __shfl_xor_sync(act, warpMax, o));
        atomicAdd(tot + 0, 1ull), atomicAdd(tot + 1, (unsigned long long)cnt.nSteps), atomicAdd(tot + 2, (unsigned long long)cnt.nMiss);
        atomicAdd(tot + 3, (unsigned long long)cnt.nInterp), atomicAdd(tot + 4, (unsigned long long)cnt.nBlock);
        if (lane == __ffs(act) - 1)
            atomicAdd(tot + 5, (unsigned long long)warpMax), atomicAdd(tot + 6, 1ull);
        return;
    }
    pointsRay[loc] = make_float4(pt.x, pt.y, pt.z, found ? conf + 1.0f : 0.0f);
    if (withColour)
    {
        // renderColour_device epilogue fused into the ray (B9): processPixelColour(ptRay.w > 0)
        uchar4 col = make_uchar4(0, 0, 0, 0);
        if (found && (conf + 1.0f) > 0)
            col = colour_interp(vba, table, pt);
        colourOut[loc] = col;
    }
}

// Sharded raycast (world > 1): the image is dealt to the ranks in strips of 8 rows, strip t to rank t mod world (interleaved so that near and
// far parts of the view, which cost very different numbers of march steps, are spread evenly); this rank marches the rays of its strips.
// Voxels are read from the rank that owns their block (mode 0: peer loads over NVLink, the hash table is local) or from the local copy
// (mode 1).  Every rank receives every row: the all-gather is the kernel's own output store (128-byte runs per warp).  LIVE: visibility
// marks go to every rank as well.
template <bool LIVE, bool REPL>
__global__ void __launch_bounds__(256) k_raycast_sharded(const __grid_constant__ ShardView v, const HashEntry *__restrict__ table, int W, int H,
                                                          Mat4 invM, float4 invProj, float oneOverVoxelSize, float mu,
                                                          const float2 *__restrict__ minmax, int mmW)
{
    // the CTA's 32 x 8 pixel tile is staged in shared memory and leaves as bulk copies (one 512-byte run per tile row and destination rank):
    // per-thread 16-byte stores into peer memory reach a small fraction of the NVLink rate, TMA bulk stores do not have that problem
    __shared__ __align__(128) float4 sOut[8][32];
    __shared__ __align__(128) uchar4 sCol[8][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int tx = (wid & 3) * 8 + (lane & 7), ty = (wid >> 2) * 4 + (lane >> 3);
    const int x0 = blockIdx.x * 32, y0 = (blockIdx.y * v.world + v.rank) * 8;
    const int x = x0 + tx, y = y0 + ty;
    const bool inside = x < W && y < H;
    float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
    uchar4 col = make_uchar4(0, 0, 0, 0);
    if (inside)
    {
        float2 mm = __ldg(&minmax[(x >> 3) + (y >> 3) * mmW]);
        // mode 1: every rank holds every block (the integrate kernel stored them), only the rows are split
        typename std::conditional<REPL, VbaLocal, VbaSharded>::type vba;
        if constexpr (REPL)
            vba.vba = v.vba[v.rank];
        else
            vba.v = &v;
        float3 pt;
        float conf;
        RayCounters cnt;
        bool found;
        if (LIVE)
        {
            const MarkAll mark = {&v};
            found = cast_ray<true, false>(vba, table, mark, x, y, invM, invProj, oneOverVoxelSize, mu, mm, pt, conf, cnt);
        }
        else
        {
            const MarkNone mark;
            found = cast_ray<false, false>(vba, table, mark, x, y, invM, invProj, oneOverVoxelSize, mu, mm, pt, conf, cnt);
        }
        out = make_float4(pt.x, pt.y, pt.z, found ? conf + 1.0f : 0.0f);
        if (!LIVE && found && (conf + 1.0f) > 0)
            col = colour_interp(vba, table, pt);
    }
    const bool bulk = (W & 3) == 0 && v.probe != 2;   // 16-byte alignment of every tile row of the colour image
    if (bulk)
    {
        sOut[ty][tx] = out;
        if (!LIVE)
            sCol[ty][tx] = col;
        fence_proxy_async();
        __syncthreads();
        // thread r < 8 sends tile row r to every rank (its own included)
        if (threadIdx.x < 8 && y0 + (int)threadIdx.x < H)
        {
            const int r = threadIdx.x;
            const int npx = min(32, W - x0);
            const size_t loc = (size_t)x0 + (size_t)(y0 + r) * W;
            for (int q = 0; q < v.world; q++)
            {
                if (v.probe == 1 && q != v.rank)
                    continue;
                tma_store_1d((LIVE ? v.rayLive[q] : v.rayFree[q]) + loc, &sOut[r][0], (unsigned)npx * 16u);
                if (!LIVE)
                    tma_store_1d(v.imageFree[q] + loc, &sCol[r][0], (unsigned)npx * 4u);
            }
            tma_store_commit();
            tma_store_wait_all();
        }
        return;
    }
    if (!inside)
        return;
    const int loc = x + y * W;
    for (int q = 0; q < v.world; q++)
    {
        if (LIVE)
            v.rayLive[q][loc] = out;
        else
        {
            v.rayFree[q][loc] = out;
            v.imageFree[q][loc] = col;
        }
    }
}

// Cross-GPU barrier of the sharded TSDF path: lane q signals rank q ("I have reached barrier `epoch`") and waits for rank q's signal;
// release / acquire at system scope, bounded spin (a rank that never arrives becomes an error flag, not a hung GPU).
__global__ void k_shard_barrier(const __grid_constant__ ShardView v, unsigned epoch, int *err)
{
    const int q = threadIdx.x;
    if (q >= v.world)
        return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(v.flags[q] + v.rank), "r"(epoch) : "memory");
    const unsigned *mine = v.flags[v.rank] + q;
    const long long t0 = clock64();
    for (;;)
    {
        unsigned got;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(got) : "l"(mine) : "memory");
        if ((int)(got - epoch) >= 0)
            break;
        if (clock64() - t0 > 40000000000LL)
        {
            *err = 1;
            break;
        }
        __nanosleep(200);
    }
    __threadfence_system();
}

// ------------------------------------------------------------------------------------------------------------
// B8: ICP maps.  reference: processPixelICP<true,false> (Visualisation_Shared.h:438-480),
// computeNormalAndAngle<true,false> (:257-338)
__device__ __forceinline__ void icp_map_pixel(const float4 *__restrict__ pointsRay, int x, int y, int W, int H, float voxelSize, float3 light,
                                              float4 &pm, float4 &nm)
{
    int loc = x + y * W;
    float4 point = __ldg(&pointsRay[loc]);
    bool found = point.w > 0.0f;
    float3 n = make_float3(0, 0, 0);
    if (found)
    {
        if (y <= 2 || y >= H - 3 || x <= 2 || x >= W - 3)
            found = false;
    }
    if (found)
    {
        float4 xp = __ldg(&pointsRay[(x + 2) + y * W]), yp = __ldg(&pointsRay[x + (y + 2) * W]);
        float4 xm = __ldg(&pointsRay[(x - 2) + y * W]), ym = __ldg(&pointsRay[x + (y - 2) * W]);
        float4 dx = make_float4(0, 0, 0, 0), dy = dx;
        bool plus1 = false;
        if (xp.w <= 0 || yp.w <= 0 || xm.w <= 0 || ym.w <= 0)
            plus1 = true;
        else
        {
            dx = make_float4(xp.x - xm.x, xp.y - xm.y, xp.z - xm.z, xp.w - xm.w);
            dy = make_float4(yp.x - ym.x, yp.y - ym.y, yp.z - ym.z, yp.w - ym.w);
            float lx = dx.x * dx.x + dx.y * dx.y + dx.z * dx.z, ly = dy.x * dy.x + dy.y * dy.y + dy.z * dy.z;
            float ld = (lx < ly) ? ly : lx;
            if (ld * voxelSize * voxelSize > (0.15f * 0.15f))
                plus1 = true;
        }
        if (plus1)
        {
            xp = __ldg(&pointsRay[(x + 1) + y * W]), yp = __ldg(&pointsRay[x + (y + 1) * W]);
            xm = __ldg(&pointsRay[(x - 1) + y * W]), ym = __ldg(&pointsRay[x + (y - 1) * W]);
            dx = make_float4(xp.x - xm.x, xp.y - xm.y, xp.z - xm.z, xp.w - xm.w);
            dy = make_float4(yp.x - ym.x, yp.y - ym.y, yp.z - ym.z, yp.w - ym.w);
            if (xp.w <= 0 || yp.w <= 0 || xm.w <= 0 || ym.w <= 0)
                found = false;
        }
        if (found)
        {
            n.x = -(dx.y * dy.z - dx.z * dy.y);
            n.y = -(dx.z * dy.x - dx.x * dy.z);
            n.z = -(dx.x * dy.y - dx.y * dy.x);
            float s = 1.0f / sqrtf(n.x * n.x + n.y * n.y + n.z * n.z);
            n.x *= s, n.y *= s, n.z *= s;
            float angle = n.x * light.x + n.y * light.y + n.z * light.z;
            if (!(angle > 0.0))
                found = false;
        }
    }
    if (found)
    {
        pm = make_float4(point.x * voxelSize, point.y * voxelSize, point.z * voxelSize, point.w);
        nm = make_float4(n.x, n.y, n.z, 0.0f);
    }
    else
        pm = nm = make_float4(0, 0, 0, -1.0f);
}

__global__ void __launch_bounds__(256) k_icp_maps(float4 *__restrict__ pointsMap, float4 *__restrict__ normalsMap,
                                                   const float4 *__restrict__ pointsRay, int W, int H, float voxelSize, float3 light)
{
    int x = blockIdx.x * 32 + (threadIdx.x & 31);
    int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H)
        return;
    float4 pm, nm;
    icp_map_pixel(pointsRay, x, y, W, H, voxelSize, light, pm, nm);
    pointsMap[x + y * W] = pm;
    normalsMap[x + y * W] = nm;
}

// sharded: the rows of this rank's slab (the raycast kernel has delivered the two neighbour rows on either side); with PUSH_ALL (tracking on:
// the ICP of the next frame projects into the whole maps) every rank receives every row
template <bool PUSH_ALL>
__global__ void __launch_bounds__(256) k_icp_maps_sharded(const __grid_constant__ ShardView v, int W, int H, int row0, int row1, float voxelSize,
                                                           float3 light)
{
    __shared__ __align__(128) float4 sP[8][32], sN[8][32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int x0 = blockIdx.x * 32, y0 = row0 + blockIdx.y * 8;
    const int x = x0 + tx, y = y0 + ty;
    const bool inside = x < W && y < row1;
    float4 pm = make_float4(0, 0, 0, -1.0f), nm = pm;
    if (inside)
        icp_map_pixel(v.rayLive[v.rank], x, y, W, H, voxelSize, light, pm, nm);
    if (!PUSH_ALL)
    {
        if (inside)
        {
            v.pointsMap[v.rank][x + y * W] = pm;
            v.normalsMap[v.rank][x + y * W] = nm;
        }
        return;
    }
    // every rank receives every row: the tile leaves as bulk copies, one 512-byte run per tile row, map and destination rank
    sP[ty][tx] = pm, sN[ty][tx] = nm;
    fence_proxy_async();
    __syncthreads();
    if (threadIdx.x < 8 && y0 + (int)threadIdx.x < row1)
    {
        const int r = threadIdx.x;
        const unsigned bytes = (unsigned)min(32, W - x0) * 16u;
        const size_t loc = (size_t)x0 + (size_t)(y0 + r) * W;
        for (int q = 0; q < v.world; q++)
        {
            tma_store_1d(v.pointsMap[q] + loc, &sP[r][0], bytes);
            tma_store_1d(v.normalsMap[q] + loc, &sN[r][0], bytes);
        }
        tma_store_commit();
        tma_store_wait_all();
    }
}

// ------------------------------------------------------------------------------------------------------------
// scene reset (ResetScene, CPU.tpp:25-49)
__global__ void k_reset_table(HashEntry *table, int E)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < E)
    {
        HashEntry e;
        e.px = e.py = e.pz = e.pad = 0;
        e.offset = 0;
        e.ptr = -2;
        table[i] = e;
    }
}
__global__ void k_reset_voxels(uint2 *v, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride)
        v[i] = make_uint2(0x00007fffu, 0u);
}

// ------------------------------------------------------------------------------------------------------------
// launch wrappers
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

void reset_scene(const Scene &s, cudaStream_t st)
{
    GS_COUNT_LAUNCHES(2);
    k_reset_table<<<cdiv(s.E, 256), 256, 0, st>>>(s.table, s.E);
    k_reset_voxels<<<148 * 8, 512, 0, st>>>(reinterpret_cast<uint2 *>(s.vba), (size_t)s.numBlocks * SDF_BLOCK_SIZE3);
    cudaMemsetAsync(s.allocKey, 0, sizeof(unsigned) * s.E, st);
    cudaMemsetAsync(s.visType, 0, s.E, st);
    int init[8] = {s.numBlocks - 1, SDF_EXCESS_LIST_SIZE - 1, 0, 0, s.numBlocks - 1, SDF_EXCESS_LIST_SIZE - 1, 0, 0};
    cudaMemcpyAsync(s.state, init, sizeof(init), cudaMemcpyHostToDevice, st);
    cudaStreamSynchronize(st); // `init` lives on this stack frame
}

void allocate(const Scene &s, const Frame &f, const Camera &cam, cudaStream_t st)
{
    const int P = f.W * f.H;
    const int nChunks = cdiv(s.E, SCAN_CTA);
    float4 invProj = make_float4(1.0f / cam.fx, 1.0f / cam.fy, cam.cx, cam.cy);
    float oneOverBlock = 1.0f / (s.voxelSize * SDF_BLOCK_SIZE);
    GS_COUNT_LAUNCHES(6);
    k_set_type3<<<148, 256, 0, st>>>(s.visIds, s.state + 2, s.visType);
    k_alloc_flags<<<cdiv(P, 256), 256, 0, st>>>(f.depth_mm, f.depth_f, f.W, f.H, cam.invM, invProj, make_float2(1.0f / 1000.0f, 0.0f), s.mu,
                                                oneOverBlock, s.vfmin, s.vfmax, s.table, s.visType, s.allocKey, s.state + 3);
    k_alloc_count<<<nChunks, SCAN_CTA, 0, st>>>(s.allocKey, s.table, s.E, s.chunkCounts);
    k_alloc_apply<<<nChunks, SCAN_CTA, 0, st>>>(s.allocKey, s.table, s.visType, s.E, nChunks, s.chunkCounts, s.state, s.state, f.depth_f, f.W,
                                                cam.invM, invProj, s.mu, oneOverBlock);
    float4 proj = make_float4(cam.fx, cam.fy, cam.cx, cam.cy);
    k_visible_count<<<nChunks, SCAN_CTA, 0, st>>>(s.visType, s.table, s.E, cam.M, proj, s.voxelSize, f.W, f.H, s.chunkCounts, s.state, s.rank,
                                                  s.world);
    k_visible_compact<<<nChunks, SCAN_CTA, 0, st>>>(s.visType, s.E, nChunks, s.chunkCounts, s.visIds, s.numBlocks, s.state + 2, s.table, s.rank,
                                                    s.world, s.visIdsOwn, s.state + 6);
}

void integrate(const Scene &s, const Frame &f, const Camera &cam, int variant, cudaStream_t st)
{
    IntegrateParams P;
    P.M = cam.M;
    P.proj = make_float4(cam.fx, cam.fy, cam.cx, cam.cy);
    P.mu = s.mu, P.voxelSize = s.voxelSize, P.maxW = s.maxW, P.W = f.W, P.H = f.H;
    GS_COUNT_LAUNCHES(1);
    // sharded scene: only the visible blocks this rank owns (per-block independent work, no exchange)
    const int *ids = s.world > 1 ? s.visIdsOwn : s.visIds;
    const int *nIds = s.world > 1 ? s.state + 6 : s.state + 2;
    PeerVbas push;
    push.n = s.nPush;
    for (int q = 0; q < s.nPush; q++)
        push.p[q] = s.pushVba[q];
    (void)variant;
    k_integrate_tma<<<148 * 5, INT_THREADS, 0, st>>>(s.vba, s.table, ids, nIds, P, f.depth_f, f.rgba, push);
}

void expected_depth_live(const Scene &s, const Camera &cam, int W, int H, float2 *minmax, cudaStream_t st)
{
    int mmW = cdiv(W, 8), mmH = cdiv(H, 8);
    GS_COUNT_LAUNCHES(2);
    k_minmax_init<<<cdiv(mmW * mmH, 256), 256, 0, st>>>(minmax, mmW * mmH);
    k_project_visible<<<148, 256, 0, st>>>(s.table, s.visIds, s.state + 2, cam.M, make_float4(cam.fx, cam.fy, cam.cx, cam.cy), W, H, s.voxelSize,
                                           minmax, mmW, mmH);
}

void expected_depth_free(const Scene &s, const Camera &cam, int W, int H, float2 *minmax, cudaStream_t st)
{
    int mmW = cdiv(W, 8), mmH = cdiv(H, 8);
    GS_COUNT_LAUNCHES(2);
    k_minmax_init<<<cdiv(mmW * mmH, 256), 256, 0, st>>>(minmax, mmW * mmH);
    k_project_all<<<cdiv(s.E, 256), 256, 0, st>>>(s.table, s.E, cam.M, make_float4(cam.fx, cam.fy, cam.cx, cam.cy), W, H, s.voxelSize, minmax, mmW,
                                                  mmH);
}

void raycast(const Scene &s, const Camera &cam, int W, int H, const float2 *minmax, float4 *pointsRay, uchar4 *colour, bool modifyVisible,
             cudaStream_t st)
{
    dim3 grid(cdiv(W, 32), cdiv(H, 8));
    float4 invProj = make_float4(1.0f / cam.fx, 1.0f / cam.fy, -cam.cx, -cam.cy);
    float oneOverVoxel = 1.0f / s.voxelSize;
    int mmW = cdiv(W, 8);
    GS_COUNT_LAUNCHES(1);
    // (forcing 32 registers / thread for 8 CTAs per SM was measured: no gain, the kernel is issue-bound)
    if (modifyVisible)
        k_raycast<true, false><<<grid, 256, 0, st>>>(pointsRay, nullptr, s.visType, s.vba, s.table, W, H, cam.invM, invProj, oneOverVoxel, s.mu, minmax,
                                                     mmW);
    else if (colour)
        k_raycast<false, true><<<grid, 256, 0, st>>>(pointsRay, colour, nullptr, s.vba, s.table, W, H, cam.invM, invProj, oneOverVoxel, s.mu, minmax,
                                                     mmW);
    else
        k_raycast<false, false><<<grid, 256, 0, st>>>(pointsRay, nullptr, nullptr, s.vba, s.table, W, H, cam.invM, invProj, oneOverVoxel, s.mu, minmax,
                                                      mmW);
}

void raycast_sharded(const Scene &s, const ShardView &v, const Camera &cam, int W, int H, const float2 *minmax, bool live, cudaStream_t st)
{
    const int strips = cdiv(H, 8);
    const int mine = strips > v.rank ? (strips - v.rank + v.world - 1) / v.world : 0;   // strips t = rank, rank + world, ...
    if (mine <= 0)
        return;
    dim3 grid(cdiv(W, 32), mine);
    float4 invProj = make_float4(1.0f / cam.fx, 1.0f / cam.fy, -cam.cx, -cam.cy);
    float oneOverVoxel = 1.0f / s.voxelSize;
    int mmW = cdiv(W, 8);
    GS_COUNT_LAUNCHES(1);
    if (live && v.replicated)
        k_raycast_sharded<true, true><<<grid, 256, 0, st>>>(v, s.table, W, H, cam.invM, invProj, oneOverVoxel, s.mu, minmax, mmW);
    else if (live)
        k_raycast_sharded<true, false><<<grid, 256, 0, st>>>(v, s.table, W, H, cam.invM, invProj, oneOverVoxel, s.mu, minmax, mmW);
    else if (v.replicated)
        k_raycast_sharded<false, true><<<grid, 256, 0, st>>>(v, s.table, W, H, cam.invM, invProj, oneOverVoxel, s.mu, minmax, mmW);
    else
        k_raycast_sharded<false, false><<<grid, 256, 0, st>>>(v, s.table, W, H, cam.invM, invProj, oneOverVoxel, s.mu, minmax, mmW);
}

void icp_maps_sharded(const Scene &s, const ShardView &v, const Camera &cam, int W, int H, bool pushAll, cudaStream_t st)
{
    int y0, y1;
    slab_rows(H, v.rank, v.world, y0, y1);
    if (y1 <= y0)
        return;
    dim3 grid(cdiv(W, 32), cdiv(y1 - y0, 8));
    float3 light = make_float3(-cam.invM.m[8], -cam.invM.m[9], -cam.invM.m[10]);
    GS_COUNT_LAUNCHES(1);
    if (pushAll)
        k_icp_maps_sharded<true><<<grid, 256, 0, st>>>(v, W, H, y0, y1, s.voxelSize, light);
    else
        k_icp_maps_sharded<false><<<grid, 256, 0, st>>>(v, W, H, y0, y1, s.voxelSize, light);
}

void shard_barrier(const ShardView &v, unsigned epoch, int *errFlag, cudaStream_t st)
{
    GS_COUNT_LAUNCHES(1);
    k_shard_barrier<<<1, 32, 0, st>>>(v, epoch, errFlag);
}

// diagnostic: the free-view march with step counters instead of outputs; totals8 = device unsigned long long[8], zeroed by the caller
void raycast_stats(const Scene &s, const Camera &cam, int W, int H, const float2 *minmax, unsigned long long *totals8, cudaStream_t st)
{
    dim3 grid(cdiv(W, 32), cdiv(H, 8));
    float4 invProj = make_float4(1.0f / cam.fx, 1.0f / cam.fy, -cam.cx, -cam.cy);
    GS_COUNT_LAUNCHES(1);
    k_raycast<false, false, true><<<grid, 256, 0, st>>>(nullptr, nullptr, reinterpret_cast<unsigned char *>(totals8), s.vba, s.table, W, H, cam.invM,
                                                          invProj, 1.0f / s.voxelSize, s.mu, minmax, cdiv(W, 8));
}

void icp_maps(const Scene &s, const Camera &cam, int W, int H, const float4 *pointsRay, float4 *pointsMap, float4 *normalsMap, cudaStream_t st)
{
    dim3 grid(cdiv(W, 32), cdiv(H, 8));
    float3 light = make_float3(-cam.invM.m[8], -cam.invM.m[9], -cam.invM.m[10]);
    GS_COUNT_LAUNCHES(1);
    k_icp_maps<<<grid, 256, 0, st>>>(pointsMap, normalsMap, pointsRay, W, H, s.voxelSize, light);
}

} // namespace tsdf
