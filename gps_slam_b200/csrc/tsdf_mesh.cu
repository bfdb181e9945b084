// Marching-cubes export of the TSDF scene (ITMBasicEngine::SaveSceneToMesh).  Built with -fmad=false.
//
// reference: ITMMeshingEngine_CPU::MeshScene (InfiniTAM/ITMLib/Engines/Meshing/CPU/ITMMeshingEngine_CPU.tpp:8-52) for the triangle order and
// the overflow rule, findPointNeighbors / sdfInterp / buildVertList (Engines/Meshing/Shared/ITMMeshingEngine_Shared.h:277-470) for the
// per-voxel arithmetic, meshScene_device (Engines/Meshing/CUDA/ITMMeshingEngine_CUDA.tcu:106-139) for the per-vertex colours (the CPU engine
// leaves them unset).
//
// What is different from the reference's CUDA mesher, by design: the output is DETERMINISTIC and in the CPU engine's order (ascending hash
// entry, then z, y, x inside a block, then the case table's triangle order) -- a counting pass, a prefix sum over the blocks and a writing pass
// instead of one atomicAdd per triangle; only allocated blocks get a CTA (the reference launches one per possible block, 262,144).
#include "common.cuh"
#include "tsdf.h"
#include "tsdf_access.cuh"

namespace tsdf
{

// The marching-cubes case table (P. Bourke, "Polygonising a scalar field", 1994; the table every implementation shares, the reference's
// triangleTable included), packed: case c -> up to five triangles = up to 15 edge indices, 4 bits each, 0xF terminates.  The edge mask of a
// case (the reference's edgeTable) is the set of edge indices that occur in its row.
__device__ const unsigned long long c_mcCases[256] = {
    0xffffffffffffffffull, 0xfffffffffffff380ull, 0xfffffffffffff910ull, 0xffffffffff189381ull,
    0xfffffffffffffa21ull, 0xffffffffffa21380ull, 0xffffffffff920a29ull, 0xfffffff89a8a2382ull,
    0xfffffffffffff2b3ull, 0xffffffffff0b82b0ull, 0xffffffffffb32091ull, 0xfffffffb89b912b1ull,
    0xffffffffff3ab1a3ull, 0xfffffffab8a801a0ull, 0xfffffff9ab9b3093ull, 0xffffffffffb8aa89ull,
    0xfffffffffffff874ull, 0xffffffffff437034ull, 0xffffffffff748910ull, 0xfffffff137174914ull,
    0xffffffffff748a21ull, 0xfffffffa21403743ull, 0xfffffff748209a29ull, 0xffff4973727929a2ull,
    0xffffffffff2b3748ull, 0xfffffff40242b74bull, 0xfffffffb32748109ull, 0xffff1292b9b49b74ull,
    0xfffffff487ab31a3ull, 0xffff4b7401b41ab1ull, 0xffff30bab9b09874ull, 0xfffffffab99b4b74ull,
    0xfffffffffffff459ull, 0xffffffffff380459ull, 0xffffffffff051450ull, 0xfffffff513538458ull,
    0xffffffffff459a21ull, 0xfffffff594a21803ull, 0xfffffff204245a25ull, 0xffff8434535235a2ull,
    0xffffffffffb32459ull, 0xfffffff594b802b0ull, 0xfffffffb32510450ull, 0xffff584b82852512ull,
    0xfffffff45931ab3aull, 0xffffab81a8180594ull, 0xffff30bab5b05045ull, 0xfffffffb8aa85845ull,
    0xffffffffff975879ull, 0xfffffff375359039ull, 0xfffffff751710870ull, 0xffffffffff753351ull,
    0xfffffff21a759879ull, 0xffff37503505921aull, 0xffff25a758528208ull, 0xfffffff7533525a2ull,
    0xfffffff2b3987597ull, 0xffffb72029279759ull, 0xffff751871810b32ull, 0xfffffff51771b12bull,
    0xffffb3a31a758859ull, 0xf0aba010b7905075ull, 0xf07570805a30b0abull, 0xffffffffff5b75abull,
    0xfffffffffffff56aull, 0xffffffffff6a5380ull, 0xffffffffff6a5109ull, 0xfffffff6a5891381ull,
    0xffffffffff162561ull, 0xfffffff803621561ull, 0xfffffff620609569ull, 0xffff823625285895ull,
    0xffffffffff56ab32ull, 0xfffffff56a02b80bull, 0xfffffff6a5b32910ull, 0xffffb892b92916a5ull,
    0xfffffff315356b36ull, 0xffff6b51505b0b80ull, 0xffff9505606306b3ull, 0xfffffff89bb96956ull,
    0xffffffffff8746a5ull, 0xfffffffa56374034ull, 0xfffffff7486a5091ull, 0xffff49737179156aull,
    0xfffffff874156216ull, 0xffff743403625521ull, 0xffff620560509748ull, 0xf962695923497937ull,
    0xfffffff56a4872b3ull, 0xffffb720242746a5ull, 0xffff6a5b32874910ull, 0xf6a54b7b492b9129ull,
    0xffff6b51535b3748ull, 0xfb404b7b016b5b15ull, 0xf74836b630560950ull, 0xffff9b7974b96956ull,
    0xffffffffffa4694aull, 0xfffffff380a946a4ull, 0xfffffff04606a10aull, 0xffffa16468618138ull,
    0xfffffff462421941ull, 0xffff462942921803ull, 0xffffffffff624420ull, 0xfffffff624428238ull,
    0xfffffff32b46a94aull, 0xffff6a4a94b82280ull, 0xffffa164606102b3ull, 0xf1b8b12184a16146ull,
    0xffff36b319639469ull, 0xf14641916b0181b8ull, 0xfffffff4600636b3ull, 0xffffffffff86b846ull,
    0xfffffffa98a876a7ull, 0xffffa76a907a0370ull, 0xffff0818717a176aull, 0xfffffff37117a76aull,
    0xffff768981861621ull, 0xf937390976192962ull, 0xfffffff206607087ull, 0xffffffffff276237ull,
    0xffff76898a86ab32ull, 0xf7a9a76790b72702ull, 0xfb32a767a1871081ull, 0xffff17616a71b12bull,
    0xf63136b619768698ull, 0xffffffffff76b190ull, 0xffff06b0b3607087ull, 0xfffffffffffff6b7ull,
    0xfffffffffffffb67ull, 0xffffffffff67b803ull, 0xffffffffff67b910ull, 0xfffffff67b138918ull,
    0xffffffffff7b621aull, 0xfffffff7b6803a21ull, 0xfffffff7b69a2092ull, 0xffff89a38a3a27b6ull,
    0xffffffffff726327ull, 0xfffffff026067807ull, 0xfffffff910732672ull, 0xffff678891681261ull,
    0xfffffff73171a67aull, 0xffff801781a7167aull, 0xffff7a69a0a70730ull, 0xfffffff9a88a7a67ull,
    0xffffffffff68b486ull, 0xfffffff640603b63ull, 0xfffffff109648b68ull, 0xffff63b139369649ull,
    0xfffffff1a28b6486ull, 0xffff640b60b03a21ull, 0xffff9a2920b648b4ull, 0xf36463b34923a39aull,
    0xfffffff264248328ull, 0xffffffffff264240ull, 0xffff834642432091ull, 0xfffffff642241491ull,
    0xffff1a6648168318ull, 0xfffffff40660a01aull, 0xf39a9303a6834364ull, 0xffffffffff4a649aull,
    0xffffffffffb67594ull, 0xfffffff67b594380ull, 0xfffffffb67045105ull, 0xffff51345343867bull,
    0xfffffffb6721a459ull, 0xffff594380a217b6ull, 0xffff204a24a45b67ull, 0xf67b25a523453843ull,
    0xfffffff945267327ull, 0xffff786260680459ull, 0xffff045051673263ull, 0xf851584812786826ull,
    0xffff73167161a459ull, 0xf459078701671a61ull, 0xfa737a6a305a4a04ull, 0xffffa84a458a7a67ull,
    0xfffffff98b9b6596ull, 0xffff590650360b63ull, 0xffffb65510b508b0ull, 0xfffffff1355363b6ull,
    0xffff65b8b9b59a21ull, 0xfa21965690b603b0ull, 0xf52025a50865b58bull, 0xffff35a3a25363b6ull,
    0xffff283265825985ull, 0xfffffff260069659ull, 0xf826283865081851ull, 0xffffffffff612651ull,
    0xf698965683a61631ull, 0xffff06505960a01aull, 0xffffffffffa65830ull, 0xfffffffffffff65aull,
    0xffffffffffb57a5bull, 0xfffffff03857ba5bull, 0xfffffff091ba57b5ull, 0xffff1381897ba57aull,
    0xfffffff15717b21bull, 0xffffb27571721380ull, 0xffff7b2209729579ull, 0xf289823295b27257ull,
    0xfffffff573532a52ull, 0xffff52a578258028ull, 0xffff2a37353a5109ull, 0xf25752a278129289ull,
    0xffffffffff573531ull, 0xfffffff571170780ull, 0xfffffff735539309ull, 0xffffffffff795789ull,
    0xfffffff8ba8a5485ull, 0xffff03bba50b5405ull, 0xffff54aba8a48910ull, 0xf41314943b54a4baull,
    0xffff8548b2582152ull, 0xfb151b2b543b0b40ull, 0xf58b8545b2950520ull, 0xffffffffff3b2549ull,
    0xffff483543253a52ull, 0xfffffff0244252a5ull, 0xf910854583a532a3ull, 0xffff2492914252a5ull,
    0xfffffff153358548ull, 0xffffffffff501540ull, 0xffff530509358548ull, 0xfffffffffffff549ull,
    0xfffffffba9b947b4ull, 0xffffba97b9794380ull, 0xffffb470414b1ba1ull, 0xf4bab474a1843413ull,
    0xffff219b294b97b4ull, 0xf3801b2b197b9479ull, 0xfffffff04224b47bull, 0xffff42343824b47bull,
    0xffff947732972a92ull, 0xf70207872a4797a9ull, 0xfa040a1a472a3a73ull, 0xffffffffff4782a1ull,
    0xfffffff317714194ull, 0xffff178180714194ull, 0xffffffffff347304ull, 0xfffffffffffff784ull,
    0xffffffffff8ba8a9ull, 0xfffffffa9bb93903ull, 0xfffffffba88a0a10ull, 0xffffffffffa3ba13ull,
    0xfffffff8b99b1b21ull, 0xffff9b2921b93903ull, 0xffffffffffb08b20ull, 0xfffffffffffffb23ull,
    0xfffffff98aa82832ull, 0xffffffffff2902a9ull, 0xffff8a1810a82832ull, 0xfffffffffffff2a1ull,
    0xffffffffff819831ull, 0xfffffffffffff190ull, 0xfffffffffffff830ull, 0xffffffffffffffffull,
};

__device__ __forceinline__ int mc_triangles(unsigned long long row)
{
    // number of leading nibbles != 0xF, divided by 3: rows are 0, 3, 6, ... 15 indices followed by 0xF padding
    const unsigned long long isF = row & (row >> 1) & (row >> 2) & (row >> 3) & 0x1111111111111111ull; // bit 4i set <=> nibble i == 0xF
    const int firstF = isF ? (__ffsll((long long)isF) - 1) >> 2 : 16;
    return firstF / 3;
}

struct Corner
{
    float sdf;
    float3 clr;
};

// sdfInterp (Shared.h:351-361), one component triple at a time
__device__ __forceinline__ float3 mc_interp(float3 p1, float3 p2, float v1, float v2)
{
    if (fabsf(0.0f - v1) < 0.00001f)
        return p1;
    if (fabsf(0.0f - v2) < 0.00001f)
        return p2;
    if (fabsf(v1 - v2) < 0.00001f)
        return p1;
    const float mu = (0.0f - v1) / (v2 - v1);
    return make_float3(p1.x + mu * (p2.x - p1.x), p1.y + mu * (p2.y - p1.y), p1.z + mu * (p2.z - p1.z));
}

// findPointNeighbors + the case index of buildVertList: -1 when a corner is unallocated or carries no measurement (sdf == 1), or when the
// cube is not crossed
template <class A>
__device__ __forceinline__ int mc_case(const A &vba, const HashEntry *__restrict__ table, int gx, int gy, int gz, Corner (&cn)[8])
{
    const int ox[8] = {0, 1, 1, 0, 0, 1, 1, 0}, oy[8] = {0, 0, 1, 1, 0, 0, 1, 1}, oz[8] = {0, 0, 0, 0, 1, 1, 1, 1};
    VoxelCache cache = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0u};
    int cube = 0;
#pragma unroll
    for (int k = 0; k < 8; k++)
    {
        int vm;
        const unsigned h = find_voxel(vba, table, gx + ox[k], gy + oy[k], gz + oz[k], vm, cache);
        if (h == NO_VOXEL)
            return -1;
        const uint2 raw = __ldg(reinterpret_cast<const uint2 *>(vba.at(h)));
        const short s = (short)(raw.x & 0xffffu);
        cn[k].sdf = (float)s / 32767.0f;
        if (cn[k].sdf == 1.0f)
            return -1;
        cn[k].clr = make_float3((float)((raw.x >> 24) & 0xffu) / 255.0f, (float)(raw.y & 0xffu) / 255.0f, (float)((raw.y >> 8) & 0xffu) / 255.0f);
        if (cn[k].sdf < 0)
            cube |= 1 << k;
    }
    return (cube == 0 || cube == 255) ? -1 : cube;
}

constexpr int MESH_CTA = SDF_BLOCK_SIZE3;

// allocated hash entries in ascending slot order: pass 1 counts per 1024-slot chunk, pass 2 ranks (same two-level scheme as the visible list)
__global__ void __launch_bounds__(1024) k_mesh_entries_count(const HashEntry *__restrict__ table, int E, int *chunkCount)
{
    const int slot = blockIdx.x * 1024 + threadIdx.x;
    const int used = slot < E && load_entry(table, slot).ptr >= 0;
    const int n = __syncthreads_count(used);
    if (threadIdx.x == 0)
        chunkCount[blockIdx.x] = n;
}

__device__ __forceinline__ int cta_excl_scan(int v, int *warpSums /* [33] */, int &total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        const int n = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d)
            incl += n;
    }
    if (lane == 31)
        warpSums[wid] = incl;
    __syncthreads();
    if (wid == 0)
    {
        const int w = lane < nw ? warpSums[lane] : 0;
        int wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            const int n = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d)
                wi += n;
        }
        warpSums[lane] = wi - w;
        if (lane == 31)
            warpSums[32] = wi;
    }
    __syncthreads();
    total = warpSums[32];
    const int r = warpSums[wid] + incl - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(1024) k_mesh_entries_compact(const HashEntry *__restrict__ table, int E, const int *__restrict__ chunkCount,
                                                                int nChunks, int *entries, int *nEntries)
{
    __shared__ int ws[33];
    __shared__ int sBase;
    // entries before this chunk
    int part = 0;
    for (int i = threadIdx.x; i < (int)blockIdx.x; i += 1024)
        part += chunkCount[i];
    int tot;
    cta_excl_scan(part, ws, tot);
    if (threadIdx.x == 0)
        sBase = tot;
    __syncthreads();
    const int slot = blockIdx.x * 1024 + threadIdx.x;
    const int used = slot < E && load_entry(table, slot).ptr >= 0;
    const int rank = cta_excl_scan(used, ws, tot);
    if (used)
        entries[sBase + rank] = slot;
    if ((int)blockIdx.x == nChunks - 1 && threadIdx.x == 0)
        *nEntries = sBase + tot;
}

// one CTA per allocated block, one thread per voxel (thread id = x + 8 y + 64 z = the CPU engine's loop order).  WRITE = false: triangles per
// block; WRITE = true: the triangles, at blockOff[block] + (triangles of the lower voxels of the block).
template <bool WRITE, class A>
__global__ void __launch_bounds__(MESH_CTA) k_mesh_cubes(const A vba, const HashEntry *__restrict__ table, const int *__restrict__ entries,
                                                          int *blockTri, const long long *__restrict__ blockOff, float voxelSize, float *out,
                                                          long long maxTri)
{
    __shared__ int ws[33];
    const HashEntry e = load_entry(table, entries[blockIdx.x]);
    const int loc = threadIdx.x;
    const int x = loc & 7, y = (loc >> 3) & 7, z = loc >> 6;
    const int gx = (int)e.px * SDF_BLOCK_SIZE + x, gy = (int)e.py * SDF_BLOCK_SIZE + y, gz = (int)e.pz * SDF_BLOCK_SIZE + z;
    Corner cn[8];
    const int cube = mc_case(vba, table, gx, gy, gz, cn);
    const unsigned long long row = cube >= 0 ? c_mcCases[cube] : ~0ull;
    const int nTri = cube >= 0 ? mc_triangles(row) : 0;
    int total;
    const int before = cta_excl_scan(nTri, ws, total);
    if (!WRITE)
    {
        if (threadIdx.x == 0)
            blockTri[blockIdx.x] = total;
        return;
    }
    if (nTri == 0)
        return;
    // the 12 edge vertices this case uses (buildVertList): edge k joins corners ea[k] -> eb[k]
    const int ea[12] = {0, 1, 2, 3, 4, 5, 6, 7, 0, 1, 2, 3}, eb[12] = {1, 2, 3, 0, 5, 6, 7, 4, 4, 5, 6, 7};
    const float cx[8] = {0, 1, 1, 0, 0, 1, 1, 0}, cy[8] = {0, 0, 1, 1, 0, 0, 1, 1}, cz[8] = {0, 0, 0, 0, 1, 1, 1, 1};
    long long t = blockOff[blockIdx.x] + before;
    for (int i = 0; i < nTri; i++, t++)
    {
        // the CPU engine keeps counting only while noTriangles < noMaxTriangles - 1: triangles beyond that are dropped
        if (t >= maxTri - 1)
            return;
        float *o = out + (size_t)t * 18;
#pragma unroll
        for (int v = 0; v < 3; v++)
        {
            const int edge = (int)((row >> (4 * (3 * i + v))) & 0xFull);
            const int a = ea[edge], b = eb[edge];
            const float3 pa = make_float3((float)gx + cx[a], (float)gy + cy[a], (float)gz + cz[a]);
            const float3 pb = make_float3((float)gx + cx[b], (float)gy + cy[b], (float)gz + cz[b]);
            const float3 p = mc_interp(pa, pb, cn[a].sdf, cn[b].sdf);
            const float3 c = mc_interp(cn[a].clr, cn[b].clr, cn[a].sdf, cn[b].sdf);
            o[v * 3 + 0] = p.x * voxelSize, o[v * 3 + 1] = p.y * voxelSize, o[v * 3 + 2] = p.z * voxelSize;
            o[9 + v * 3 + 0] = c.x, o[9 + v * 3 + 1] = c.y, o[9 + v * 3 + 2] = c.z;
        }
    }
}

// exclusive prefix sum of the per-block triangle counts (one CTA, 64-bit running total)
__global__ void __launch_bounds__(1024) k_mesh_scan(const int *__restrict__ blockTri, int n, long long *blockOff, long long *total)
{
    __shared__ int ws[33];
    __shared__ long long sRun;
    if (threadIdx.x == 0)
        sRun = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024)
    {
        const int i = base + threadIdx.x;
        const int v = i < n ? blockTri[i] : 0;
        int tot;
        const int ex = cta_excl_scan(v, ws, tot);
        if (i < n)
            blockOff[i] = sRun + ex;
        __syncthreads();
        if (threadIdx.x == 0)
            sRun += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0)
        *total = sRun;
}

// host side.  scratch (mesh_scratch_bytes): long long blockOff[numBlocks] | long long total | int entries[numBlocks] |
// int blockTri[numBlocks] | int chunkCount[ceil(E / 1024)] | int nEntries
size_t mesh_scratch_bytes(const Scene &s)
{
    const size_t nChunks = (size_t)(s.E + 1023) / 1024;
    return ((size_t)s.numBlocks + 1) * sizeof(long long) + (2 * (size_t)s.numBlocks + nChunks + 1) * sizeof(int);
}

// maxTri = capacity of outDev in triangles (18 floats each); outDev == nullptr: count only.  *nTriHost = triangles written (resp. present).
// Synchronises the stream.
int mesh_scene(const Scene &s, const ShardView *view, void *scratch, float *outDev, long long maxTri, long long *nTriHost, cudaStream_t st)
{
    *nTriHost = 0;
    if (view && view->world > 1 && !view->replicated)
        return gs_set_error(__FILE__, __LINE__, "mesh export of a storage-sharded scene (shard mode 0) is not built: use shard mode 1 or one GPU");
    const int nChunks = (s.E + 1023) / 1024;
    long long *blockOff = (long long *)scratch;
    long long *total = blockOff + s.numBlocks;
    int *entries = (int *)(total + 1);
    int *blockTri = entries + s.numBlocks;
    int *chunkCount = blockTri + s.numBlocks;
    int *nEntries = chunkCount + nChunks;
    GS_COUNT_LAUNCHES(2);
    k_mesh_entries_count<<<nChunks, 1024, 0, st>>>(s.table, s.E, chunkCount);
    k_mesh_entries_compact<<<nChunks, 1024, 0, st>>>(s.table, s.E, chunkCount, nChunks, entries, nEntries);
    int nB = 0;
    GS_CUDA_OK(cudaMemcpyAsync(&nB, nEntries, sizeof(int), cudaMemcpyDeviceToHost, st));
    GS_CUDA_OK(cudaStreamSynchronize(st));
    if (nB <= 0)
        return 0;
    if (nB > s.numBlocks)
        return gs_set_error(__FILE__, __LINE__, "more allocated hash entries than voxel blocks: corrupt scene");
    const VbaLocal vba = {s.vba};
    GS_COUNT_LAUNCHES(2);
    k_mesh_cubes<false, VbaLocal><<<nB, MESH_CTA, 0, st>>>(vba, s.table, entries, blockTri, nullptr, s.voxelSize, nullptr, 0);
    k_mesh_scan<<<1, 1024, 0, st>>>(blockTri, nB, blockOff, total);
    long long n = 0;
    GS_CUDA_OK(cudaMemcpyAsync(&n, total, sizeof n, cudaMemcpyDeviceToHost, st));
    GS_CUDA_OK(cudaStreamSynchronize(st));
    if (!outDev)
    {
        *nTriHost = n;
        return 0;
    }
    const long long kept = n < maxTri - 1 ? n : (maxTri > 1 ? maxTri - 1 : 0);
    *nTriHost = kept;
    if (kept <= 0)
        return 0;
    GS_COUNT_LAUNCHES(1);
    k_mesh_cubes<true, VbaLocal><<<nB, MESH_CTA, 0, st>>>(vba, s.table, entries, blockTri, blockOff, s.voxelSize, outDev, maxTri);
    GS_CUDA_OK(cudaGetLastError());
    GS_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

} // namespace tsdf
