// gsplat-GES path, per-Gaussian side, hand-written for sm_100a.
// Built with -fmad=false: the forward expressions follow oracle/gs_oracle.py operation by operation so that the integer
// artefacts (radii, tile rectangles, bins) are bit-identical to the restated reference.
//
// What replaces what (reference file:line):
//   k_project_sh      getRealScales/getRealOpacities (include/raw_gs_param.h:80-82) + fully_fused_projection_fwd_kernel
//                     (gsplat/rasterizer/fully_fused_projection_fwd.cu:20-194) + clamp_max(radii) (src/raw_gs_model.cpp:241-242)
//                     + viewDirs / cat(featuresDc, featuresRest) / compute_sh_fwd_kernel / clamp_min(c+0.5, 0)
//                     (src/raw_gs_model.cpp:253-257, gsplat/rasterizer/compute_sh_fwd.cu:12-38) + the first pass of
//                     isect_tiles_no_depth (isect_tiles_no_depth.cu:57-90): one pass over the parameters, no intermediates
//                     in HBM except the 48-byte splat record.
//   k_scan_tiles / k_scatter_tiles / k_sort_tiles
//                     cumsum + second pass of isect_tiles_no_depth + cub::DeviceRadixSort + isect_offset_encode_no_depth
//                     (isect_tiles_no_depth.cu:132-371, 373-461).  Counting sort by tile with a per-tile ordering pass gives the
//                     same (tile, ascending Gaussian id) order as the reference's stable radix sort, with every size kept on
//                     the device (the reference reads n_isects / n_groups back with .item() twice per iteration).
//   k_bwd_params_adam compute_sh_bwd_kernel (compute_sh_bwd.cu:14-54) + fully_fused_projection_bwd_kernel
//                     (fully_fused_projection_bwd.cu:21-265) + exp/sigmoid backward + 6 x torch::optim::Adam::step
//                     (src/raw_gs_model.cpp:654-705) in one pass over parameters and optimiser state.
#include "common.cuh"
#include "gs.h"
#include "tma.cuh"

namespace gs
{

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// tile rectangle of a splat: isect_tiles_no_depth.cu:70-80 (the float -> uint32 casts saturate on the GPU)
__device__ __forceinline__ void tile_rect(float m2x, float m2y, int radius, int tileW, int tileH, int &x0, int &y0, int &x1, int &y1)
{
    float ts = (float)TILE;
    float tr = (float)radius / ts;
    float tx = m2x / ts, ty = m2y / ts;
    x0 = (int)min((unsigned)floorf(tx - tr), (unsigned)tileW);
    y0 = (int)min((unsigned)floorf(ty - tr), (unsigned)tileH);
    x1 = (int)min((unsigned)ceilf(tx + tr), (unsigned)tileW);
    y1 = (int)min((unsigned)ceilf(ty + tr), (unsigned)tileH);
}

__device__ __forceinline__ void load_coeffs(const ParamPtrs &p, int g, float *cl /* [16][3] */)
{
    cl[0] = p.dc[g * 3 + 0], cl[1] = p.dc[g * 3 + 1], cl[2] = p.dc[g * 3 + 2];
    const float *r = p.rest + (size_t)g * 45;
#pragma unroll
    for (int i = 0; i < 45; i++)
        cl[3 + i] = r[i];
}

// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_project_sh(ParamPtrs p, const int *__restrict__ nDev, CamParams cam, SplatRec *__restrict__ recs,
                                                     SplatGrad *__restrict__ grads, int *segCount, int chunkSize, int tileW, int tileH, int4 *items,
                                                     int itemCap, int *counters, int forBackward)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    // The 15x3 higher-order SH coefficients of the warp's 32 consecutive Gaussians are one contiguous 5,760-byte run: one TMA bulk
    // copy per warp brings it into shared memory; lanes then read their own row at stride 45 (odd -> conflict free).  Buffers are
    // padded to a multiple of 128 Gaussians (and the grid never exceeds that), so the full run and every parameter below are readable
    // for any g of the grid: the small parameters are requested before anything depends on them, the Gaussian count included.
    __shared__ __align__(128) float sRest[4][32 * 45];
    __shared__ unsigned long long sBar[4];
    float *rest = sRest[threadIdx.x >> 5];
    unsigned long long *bar = &sBar[threadIdx.x >> 5];
    float mean[3];
    mean[0] = p.means[g * 3 + 0], mean[1] = p.means[g * 3 + 1], mean[2] = p.means[g * 3 + 2];
    const float ls0 = p.scales[g * 3 + 0], ls1 = p.scales[g * 3 + 1], ls2 = p.scales[g * 3 + 2];
    const float4 q4 = reinterpret_cast<const float4 *>(p.quats)[g];
    const float opacLogit = p.opac[g];
    const float dc0 = p.dc[g * 3 + 0], dc1 = p.dc[g * 3 + 1], dc2 = p.dc[g * 3 + 2];
    const int nGauss = *nDev;
    const bool inRange = g < nGauss;
    Proj o;
    o.radius = 0;
    if (inRange)
    {
        float scale[3] = {expf(ls0), expf(ls1), expf(ls2)};
        float quat[4] = {q4.x, q4.y, q4.z, q4.w};
        o = project_one(mean, quat, scale, cam, nullptr);
    }
    const bool vis = o.radius > 0;
    const unsigned full = 0xffffffffu;
    const unsigned vm = __ballot_sync(full, vis);
    // The SH run is 76 % of the bytes this kernel could read, and in a grown map most warps see nothing (Gaussians are appended in raster
    // order: neighbours in id are neighbours in space): it is fetched only by warps with a visible Gaussian, after the projection
    // (the other resident warps cover the copy's latency; an earlier, cheaper test cannot be exact -- the cull uses the unclamped radius)
    if (vm)
    {
        if (lane == 0)
        {
            tma::mbar_init(bar, 1);
            tma::mbar_expect_tx(bar, 32 * 45 * 4);
            tma::load_1d(rest, p.rest + (size_t)(g - lane) * 45, 32 * 45 * 4, bar);
        }
        __syncwarp();
        tma::mbar_wait(bar, 0); // every lane: the CTA must not retire with the copy in flight
    }
    int nItems = 0, rectXY = 0, rectWH = 0;
    int bits = 0;
    float col[3] = {0.f, 0.f, 0.f};
    float opac = 0.f;
    if (vis)
    {
        opac = 1.0f / (1.0f + expf(-opacLogit));
        // SH colour (degree 3), dirs = means - camT, colour = max(SH + 0.5, 0)
        float dir[3] = {mean[0] - cam.cam_pos[0], mean[1] - cam.cam_pos[1], mean[2] - cam.cam_pos[2]};
        ShBasis sb;
        sh_basis(dir, sb);
        float cl[48];
        cl[0] = dc0, cl[1] = dc1, cl[2] = dc2;
#pragma unroll
        for (int i = 0; i < 45; i++)
            cl[3 + i] = rest[lane * 45 + i];
#pragma unroll
        for (int c = 0; c < 3; c++)
        {
            float raw = sh_eval_channel(sb, cl + c, 3) + 0.5f;
            if (raw >= 0.f)
                bits |= 1 << c;
            col[c] = fmaxf(raw, 0.f);
        }
        int x0, y0, x1, y1;
        tile_rect(o.m2x, o.m2y, o.radius, tileW, tileH, x0, y0, x1, y1);
        const int chunk = g / chunkSize; // id-range chunk: bins are counted per (tile, chunk) so that the scatter is ordered across chunks
        for (int ty = y0; ty < y1; ty++)
            for (int tx = x0; tx < x1; tx++)
                atomicAdd(&segCount[(ty * tileW + tx) * BIN_CHUNKS + chunk], 1);
        if (forBackward)
        {
            int rx, ry, rw, rh;
            int npix = bwd_rect(o.m2x, o.m2y, o.radius, o.ca, o.cb, o.cc, opac, cam.W, cam.H, rx, ry, rw, rh);
            nItems = (npix + BWD_PIXELS_PER_ITEM - 1) / BWD_PIXELS_PER_ITEM;
            rectXY = rx | (ry << 16), rectWH = rw | (rh << 16);
            if (nItems != 1)
            {
                // 0 items: nothing will write this splat's gradient; > 1: several warps accumulate with atomics. Start from zero.
                if (nItems > 1)
                    bits |= 256;
                float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                grads[g].g0 = z, grads[g].g1 = z, grads[g].g2 = z;
            }
        }
    }
    // warp-aggregated reservation of backward work items and of the visible count (one atomic per warp)
    if (vm == 0)
    {
        if (inRange)
            recs[g].q0 = make_float4(0.f, 0.f, 0.f, __int_as_float(0));
        return;
    }
    int incl = nItems;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        int n = __shfl_up_sync(full, incl, d);
        if (lane >= d)
            incl += n;
    }
    int total = __shfl_sync(full, incl, 31);
    int base = 0;
    if (lane == 31)
    {
        atomicAdd(&counters[CNT_VISIBLE], __popc(vm));
        if (total > 0)
            base = atomicAdd(&counters[CNT_ITEMS], total);
    }
    base = __shfl_sync(full, base, 31) + incl - nItems;
    if (nItems > 0)
    {
        if (base + nItems <= itemCap)
        {
            for (int i = 0; i < nItems; i++)
                items[base + i] = make_int4(g, i * BWD_PIXELS_PER_ITEM, rectXY, rectWH);
        }
        else
        {
            // dropped work items: nothing will write this splat's raster gradient -> it is zero, not whatever the buffer held
            atomicOr(&counters[CNT_OVERFLOW], 2);
            float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            grads[g].g0 = z, grads[g].g1 = z, grads[g].g2 = z;
        }
    }
    if (!inRange)
        return;
    if (!vis)
    {
        recs[g].q0 = make_float4(0.f, 0.f, 0.f, __int_as_float(0));
        return;
    }
    SplatRec r;
    r.q0 = make_float4(o.m2x, o.m2y, opac, __int_as_float(o.radius));
    r.q1 = make_float4(o.ca, o.cb, o.cc, o.depth);
    r.q2 = make_float4(col[0], col[1], col[2], __int_as_float(bits));
    recs[g] = r;
}

// Binning = counting sort by (tile, id-range chunk).  The Gaussians are cut into BIN_CHUNKS consecutive id ranges; the projection
// pass counts intersections per (tile, chunk) segment, k_seg_scan turns the counts of a tile into segment offsets, k_scan_tiles
// scans the tile totals, k_scatter_tiles drops each id into its segment (order inside a segment is arbitrary, segments of a tile
// are in ascending id-range order), and k_sort_tiles orders the few ids of every segment.  Result: ascending Gaussian id within
// each tile = the order of the reference's stable radix sort by tile id (isect_tiles_no_depth.cu:233-301), with the same-address
// atomic pressure spread over BIN_CHUNKS x more counters and a sort of ~5-element segments instead of whole tile lists.
__global__ void __launch_bounds__(256) k_seg_scan(int *segCount, int *segOff, int *tileCount, int T)
{
    const int lane = threadIdx.x & 31;
    const int t = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (t >= T)
        return;
    static_assert(BIN_CHUNKS == 64, "one int2 per lane");
    int2 *cnt = reinterpret_cast<int2 *>(segCount + (size_t)t * BIN_CHUNKS) + lane;
    const int2 v = *cnt;
    const int sum = v.x + v.y;
    int incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        int n = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d)
            incl += n;
    }
    const int excl = incl - sum;
    reinterpret_cast<int2 *>(segOff + (size_t)t * BIN_CHUNKS)[lane] = make_int2(excl, excl + v.x);
    *cnt = make_int2(0, 0); // the counters become the scatter cursors
    if (lane == 31)
        tileCount[t] = incl;
}

// exclusive scan of the per-tile counts (T + 1 <= a few 10^4 entries) by one CTA; also re-arms the per-iteration counters
__global__ void __launch_bounds__(1024) k_scan_tiles(const int *tileCount, int *tileOffsets, int T, int isectCap, int *counters)
{
    __shared__ int warpSums[32];
    __shared__ int carry;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0)
        carry = 0;
    __syncthreads();
    for (int base = 0; base < T; base += 1024)
    {
        int i = base + tid;
        int v = (i < T) ? tileCount[i] : 0;
        int incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            int n = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d)
                incl += n;
        }
        if (lane == 31)
            warpSums[wid] = incl;
        __syncthreads();
        if (wid == 0)
        {
            int w = warpSums[lane];
            int wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1)
            {
                int n = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= d)
                    wi += n;
            }
            warpSums[lane] = wi - w; // exclusive
        }
        __syncthreads();
        int excl = carry + warpSums[wid] + incl - v;
        if (i < T)
            tileOffsets[i] = min(excl, isectCap);
        __syncthreads();
        if (tid == 1023)
            carry = excl + v;
        __syncthreads();
    }
    if (tid == 0)
    {
        int total = carry;
        if (total > isectCap)
        {
            atomicOr(&counters[CNT_OVERFLOW], 1);
            total = isectCap;
        }
        tileOffsets[T] = total;
        counters[CNT_ISECTS] = total;
        counters[CNT_VISIBLE_LAST] = counters[CNT_VISIBLE];
        counters[CNT_VISIBLE] = 0;
        counters[CNT_BWD_CURSOR] = 0; // re-arm the rasteriser backward's work cursor (saves a memset node per iteration)
    }
}

__global__ void __launch_bounds__(256) k_scatter_tiles(const SplatRec *__restrict__ recs, const int *__restrict__ nDev,
                                                        const int *__restrict__ tileOffsets, const int *__restrict__ segOff, int *segCursor,
                                                        int chunkSize, int *flatten, int isectCap, int tileW, int tileH, int nUpper)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (nDev ? *nDev : nUpper)) // nDev == nullptr: staged call, the count is the launch bound itself
        return;
    float4 q0 = __ldg(&recs[g].q0);
    int radius = __float_as_int(q0.w);
    if (radius <= 0)
        return;
    int x0, y0, x1, y1;
    tile_rect(q0.x, q0.y, radius, tileW, tileH, x0, y0, x1, y1);
    const int chunk = g / chunkSize;
    // The cursor bumps return a value (the slot), so each one costs a full round trip to L2.  They are independent: issue them four
    // at a time and only then consume the results, instead of one round trip per tile.
    constexpr int B = 4;
    const int nx = x1 - x0, total = nx * (y1 - y0);
    for (int base = 0; base < total; base += B)
    {
        int pos[B];
#pragma unroll
        for (int i = 0; i < B; i++)
        {
            const int k = base + i;
            pos[i] = isectCap;
            if (k < total)
            {
                const int row = k / nx;
                const int t = (y0 + row) * tileW + x0 + (k - row * nx);
                const int seg = t * BIN_CHUNKS + chunk;
                pos[i] = __ldg(&tileOffsets[t]) + __ldg(&segOff[seg]) + atomicAdd(&segCursor[seg], 1);
            }
        }
#pragma unroll
        for (int i = 0; i < B; i++)
            if (pos[i] < isectCap)
                flatten[pos[i]] = g;
    }
}

// per tile: order the ids of every (tile, chunk) segment ascending; consumes (zeroes) the segment cursors
constexpr int SORT_SMEM = 8192;
// A segment longer than this is not rank-sorted by one warp (n^2 / 32 steps on a single warp while the other seven idle -- Gaussians are
// appended in raster order, so the ids of one tile cluster in a few id-range chunks and segments of several hundred ids are common in a
// grown map): the whole tile list is sorted cooperatively by a bitonic network instead (n log^2 n / 256 steps per thread).
constexpr int SORT_SEG_MAX = 64;
__global__ void __launch_bounds__(256) k_sort_tiles(const int *__restrict__ tileOffsets, const int *__restrict__ segOff, int *segLen,
                                                     const int *__restrict__ flatten, int *flattenSorted)
{
    __shared__ int s[SORT_SMEM];
    __shared__ int sLen[BIN_CHUNKS], sOff[BIN_CHUNKS];
    __shared__ int sMaxLen;
    const int t = blockIdx.x;
    const int start = tileOffsets[t], L = tileOffsets[t + 1] - start;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0)
        sMaxLen = 0;
    __syncthreads();
    if (tid < BIN_CHUNKS)
    {
        const int off = segOff[(size_t)t * BIN_CHUNKS + tid];
        int len = segLen[(size_t)t * BIN_CHUNKS + tid];
        segLen[(size_t)t * BIN_CHUNKS + tid] = 0;
        len = max(0, min(len, L - off)); // only differs when the intersection capacity overflowed
        sLen[tid] = len, sOff[tid] = off;
        if (len > SORT_SEG_MAX)
            atomicMax(&sMaxLen, len);
    }
    if (L <= 0)
        return;
    const bool inSmem = L <= SORT_SMEM;
    if (inSmem)
        for (int i = tid; i < L; i += 256)
            s[i] = flatten[start + i];
    __syncthreads();
    if (sMaxLen == 0)
    {
        const int *src = inSmem ? s : flatten + start;
        for (int c = wid; c < BIN_CHUNKS; c += 8)
        {
            const int n = sLen[c], o = sOff[c];
            if (n == 0)
                continue;
            if (n <= 32)
            {
                // one element per lane, rank by n broadcasts
                const int a = lane < n ? src[o + lane] : 0x7fffffff;
                int rank = 0;
                for (int j = 0; j < n; j++)
                    rank += (__shfl_sync(0xffffffffu, a, j) < a);
                if (lane < n)
                    flattenSorted[start + o + rank] = a;
            }
            else
            {
                for (int i = lane; i < n; i += 32)
                {
                    const int a = src[o + i];
                    int rank = 0;
                    for (int j = 0; j < n; j++)
                        rank += (src[o + j] < a);
                    flattenSorted[start + o + rank] = a;
                }
            }
        }
        return;
    }
    // a very long segment: bitonic sort of the whole tile list (ascending ids = sorted segments in chunk order)
    if (inSmem)
    {
        int n = 2;
        while (n < L)
            n <<= 1;
        for (int i = L + tid; i < n; i += 256)
            s[i] = 0x7fffffff;
        __syncthreads();
        for (int k = 2; k <= n; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1)
            {
                // one compare-exchange per loop step: pair p -> (i, i | j) with bit j of i clear
                for (int p = tid; p < (n >> 1); p += 256)
                {
                    const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
                    const int a = s[i], b = s[i | j];
                    const bool asc = (i & k) == 0;
                    if ((a > b) == asc)
                        s[i] = b, s[i | j] = a;
                }
                __syncthreads();
            }
        for (int i = tid; i < L; i += 256)
            flattenSorted[start + i] = s[i];
    }
    else
    {
        // > SORT_SMEM splats on one 16x16 tile: rank sort straight from global memory; ids are unique
        for (int i = tid; i < L; i += 256)
        {
            int a = flatten[start + i];
            int rank = 0;
            for (int j = 0; j < L; j++)
                rank += (flatten[start + j] < a);
            flattenSorted[start + rank] = a;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void adam_one(float *p, float *m, float *v, size_t idx, float g, bool hadState, const AdamScalars &a, float stepSize)
{
    float mo = hadState ? m[idx] : 0.f;
    float vo = hadState ? v[idx] : 0.f;
    float pn = adam_update(p[idx], g, mo, vo, a, stepSize);
    p[idx] = pn, m[idx] = mo, v[idx] = vo;
}

// Parameter backward + Adam in two kernels.
//  k_bwd_params: one lane per Gaussian (a warp owns 32 consecutive ones).  The 15x3 higher-order SH coefficients of those 32
//     Gaussians are one contiguous 5,760-byte run: the warp stages it in shared memory with coalesced loads and every lane reads
//     its own 45 coefficients at stride 45 (odd -> conflict free) for the SH VJP.  It applies Adam to the 14 "small" parameters
//     (means, scales, quats, DC colour, opacity) and leaves, per Gaussian, the 15 SH basis values, the 3 colour gradients and a
//     state flag in an 80-byte aux record.
//  k_adam_rest: pure streaming pass over the higher-order SH coefficients and their two moments (76% of all parameter bytes):
//     16-byte accesses, two per thread in flight, gradient = basis[k] * v_colour[c] rebuilt from the aux record.
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
constexpr int ADAM_WARPS = 4;
constexpr int AUX_FLOATS = 20; // basis[1..15], vcol[3], flag, pad
__global__ void __launch_bounds__(ADAM_WARPS * 32, 4) k_bwd_params(ParamPtrs p, ParamPtrs m, ParamPtrs v, unsigned char *touched, AdamStep step,
                                                                   const int *__restrict__ nDev, CamParams cam,
                                                                   const SplatRec *__restrict__ recs, const SplatGrad *__restrict__ grads,
                                                                   float4 *__restrict__ aux, ParamPtrs dbg, int haveDbg, int *counters)
{
    __shared__ __align__(128) float sRest[ADAM_WARPS][32 * 45];
    __shared__ unsigned long long sBar[ADAM_WARPS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int g0 = (blockIdx.x * ADAM_WARPS + wid) * 32;
    const int g = g0 + lane;
    if (blockIdx.x == 0 && threadIdx.x == 0)
        counters[CNT_ITEMS] = 0; // re-arm for the next projection
    // In a grown map four Gaussians out of five are neither visible nor carry optimiser state: only the 16-byte head of the record (the
    // radius) and the state flag are read for everybody; the rest of the record, the raster gradients and the moments follow for the
    // lanes that need them (one more round trip for those, 180 bytes less for all the others).
    const float4 q0r = __ldg(&recs[g].q0);
    const unsigned char tch = touched[g];
    const int N = *nDev;
    if (g0 >= N)
        return;
    const bool inRange = g < N;
    const float4 q0 = inRange ? q0r : make_float4(0.f, 0.f, 0.f, 0.f);
    const int radius = __float_as_int(q0.w);
    const bool vis = inRange && radius > 0;
    const bool had = inRange && tch != 0;
    float4 q1r = make_float4(0.f, 0.f, 0.f, 0.f), q2r = q1r, sg0 = q1r, sg1 = q1r, sg2 = q1r;
    if (vis)
    {
        q1r = __ldg(&recs[g].q1), q2r = __ldg(&recs[g].q2);
        sg0 = __ldg(&grads[g].g0), sg1 = __ldg(&grads[g].g1), sg2 = __ldg(&grads[g].g2);
    }
    if (had)
    {
        prefetch_l2(m.means + g * 3), prefetch_l2(v.means + g * 3), prefetch_l2(m.scales + g * 3), prefetch_l2(v.scales + g * 3);
        prefetch_l2(m.dc + g * 3), prefetch_l2(v.dc + g * 3), prefetch_l2(m.quats + g * 4), prefetch_l2(v.quats + g * 4);
        prefetch_l2(m.opac + g), prefetch_l2(v.opac + g);
    }
    const unsigned full = 0xffffffffu;
    const unsigned visMask = __ballot_sync(full, vis);
    // SH coefficient run of the warp's 32 Gaussians by one TMA bulk copy, overlapped with the projection re-computation below -- only
    // when one of them is visible: the SH VJP is all that reads it, and in a grown map most warps are culled as a whole (Gaussians are
    // appended in raster order, neighbours in id are neighbours in space)
    if (lane == 0 && visMask)
    {
        tma::mbar_init(&sBar[wid], 1);
        tma::mbar_expect_tx(&sBar[wid], 32 * 45 * 4);
        tma::load_1d(sRest[wid], p.rest + (size_t)g0 * 45, 32 * 45 * 4, &sBar[wid]);
    }
    __syncwarp();
    float *rest = sRest[wid];
    if (!vis && visMask)
        tma::mbar_wait(&sBar[wid], 0); // every lane waits (the CTA must not retire with the copy in flight); visible lanes wait below
    float gm[3] = {0.f, 0.f, 0.f}, gsc[3] = {0.f, 0.f, 0.f}, gq[4] = {0.f, 0.f, 0.f, 0.f}, gdc[3] = {0.f, 0.f, 0.f}, gop = 0.f;
    float basis[16], vcol[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 16; k++)
        basis[k] = 0.f;
    if (vis)
    {
        const float4 q1 = q1r, q2 = q2r;
        SplatGrad sg;
        sg.g0 = sg0, sg.g1 = sg1, sg.g2 = sg2;
        int bits = __float_as_int(q2.w);
        vcol[0] = (bits & 1) ? sg.g2.x : 0.f;
        vcol[1] = (bits & 2) ? sg.g2.y : 0.f;
        vcol[2] = (bits & 4) ? sg.g2.z : 0.f;
        float mean[3] = {p.means[g * 3 + 0], p.means[g * 3 + 1], p.means[g * 3 + 2]};
        float scale[3] = {expf(p.scales[g * 3 + 0]), expf(p.scales[g * 3 + 1]), expf(p.scales[g * 3 + 2])};
        float4 q4 = reinterpret_cast<const float4 *>(p.quats)[g];
        float quat[4] = {q4.x, q4.y, q4.z, q4.w};
        ProjIntermediates keep;
        project_one(mean, quat, scale, cam, &keep);
        float dir[3] = {mean[0] - cam.cam_pos[0], mean[1] - cam.cam_pos[1], mean[2] - cam.cam_pos[2]};
        ShBasis sb;
        sh_basis(dir, sb);
        float cl[48];
        cl[0] = p.dc[g * 3 + 0], cl[1] = p.dc[g * 3 + 1], cl[2] = p.dc[g * 3 + 2];
        tma::mbar_wait(&sBar[wid], 0);
#pragma unroll
        for (int i = 0; i < 45; i++)
            cl[3 + i] = rest[lane * 45 + i];
        float vdir[3];
        sh_vjp(sb, cl, 3, vcol, basis, vdir);
        float vmean[3], vquat[4], vscale[3];
        project_vjp(scale, cam, keep, q1.x, q1.y, q1.z, sg.g0.x, sg.g0.y, sg.g0.w, sg.g1.x, sg.g1.y, sg.g1.z, vmean, vquat, vscale);
#pragma unroll
        for (int i = 0; i < 3; i++)
        {
            gm[i] = vmean[i] + vdir[i];
            gsc[i] = vscale[i] * scale[i];
            gdc[i] = basis[0] * vcol[i];
        }
#pragma unroll
        for (int i = 0; i < 4; i++)
            gq[i] = vquat[i];
        float o = q0.z;
        gop = sg.g0.z * o * (1.0f - o);
    }
    if (vis)
    {
        float4 *a = aux + (size_t)g * (AUX_FLOATS / 4);
        a[0] = make_float4(basis[1], basis[2], basis[3], basis[4]);
        a[1] = make_float4(basis[5], basis[6], basis[7], basis[8]);
        a[2] = make_float4(basis[9], basis[10], basis[11], basis[12]);
        a[3] = make_float4(basis[13], basis[14], basis[15], vcol[0]);
        a[4] = make_float4(vcol[1], vcol[2], 0.f, 0.f);
    }
    if (haveDbg && inRange)
    {
#pragma unroll
        for (int i = 0; i < 3; i++)
            dbg.means[g * 3 + i] = gm[i], dbg.scales[g * 3 + i] = gsc[i], dbg.dc[g * 3 + i] = gdc[i];
#pragma unroll
        for (int i = 0; i < 4; i++)
            dbg.quats[g * 4 + i] = gq[i];
        dbg.opac[g] = gop;
        for (int e = 0; e < 45; e++)
            dbg.rest[(size_t)g * 45 + e] = basis[e / 3 + 1] * vcol[e % 3];
    }
    // Adam: a Gaussian that never received a gradient in this optimiser cycle has m = v = 0 and a zero update, so it is
    // skipped exactly; its state is materialised on first touch.
    if (vis || had)
    {
        // small parameter groups: 13 values per Gaussian; all loads are issued before the first store
        float P13[13], M13[13], V13[13], G13[13];
#pragma unroll
        for (int i = 0; i < 3; i++)
        {
            P13[i] = p.means[g * 3 + i], P13[3 + i] = p.scales[g * 3 + i], P13[6 + i] = p.dc[g * 3 + i];
            G13[i] = gm[i], G13[3 + i] = gsc[i], G13[6 + i] = gdc[i];
        }
        {
            float4 q4 = reinterpret_cast<const float4 *>(p.quats)[g];
            P13[9] = q4.x, P13[10] = q4.y, P13[11] = q4.z, P13[12] = q4.w;
            G13[9] = gq[0], G13[10] = gq[1], G13[11] = gq[2], G13[12] = gq[3];
        }
        float po = p.opac[g], mo = 0.f, vo = 0.f;
#pragma unroll
        for (int i = 0; i < 13; i++)
            M13[i] = 0.f, V13[i] = 0.f;
        if (had)
        {
#pragma unroll
            for (int i = 0; i < 3; i++)
            {
                M13[i] = m.means[g * 3 + i], M13[3 + i] = m.scales[g * 3 + i], M13[6 + i] = m.dc[g * 3 + i];
                V13[i] = v.means[g * 3 + i], V13[3 + i] = v.scales[g * 3 + i], V13[6 + i] = v.dc[g * 3 + i];
            }
            float4 mq = reinterpret_cast<const float4 *>(m.quats)[g], vq = reinterpret_cast<const float4 *>(v.quats)[g];
            M13[9] = mq.x, M13[10] = mq.y, M13[11] = mq.z, M13[12] = mq.w;
            V13[9] = vq.x, V13[10] = vq.y, V13[11] = vq.z, V13[12] = vq.w;
            mo = m.opac[g], vo = v.opac[g];
        }
#pragma unroll
        for (int i = 0; i < 13; i++)
        {
            const float ss = i < 3 ? step.step_size[0] : (i < 6 ? step.step_size[1] : (i < 9 ? step.step_size[3] : step.step_size[2]));
            P13[i] = adam_update(P13[i], G13[i], M13[i], V13[i], step.a, ss);
        }
        po = adam_update(po, gop, mo, vo, step.a, step.step_size[5]);
#pragma unroll
        for (int i = 0; i < 3; i++)
        {
            p.means[g * 3 + i] = P13[i], p.scales[g * 3 + i] = P13[3 + i], p.dc[g * 3 + i] = P13[6 + i];
            m.means[g * 3 + i] = M13[i], m.scales[g * 3 + i] = M13[3 + i], m.dc[g * 3 + i] = M13[6 + i];
            v.means[g * 3 + i] = V13[i], v.scales[g * 3 + i] = V13[3 + i], v.dc[g * 3 + i] = V13[6 + i];
        }
        reinterpret_cast<float4 *>(p.quats)[g] = make_float4(P13[9], P13[10], P13[11], P13[12]);
        reinterpret_cast<float4 *>(m.quats)[g] = make_float4(M13[9], M13[10], M13[11], M13[12]);
        reinterpret_cast<float4 *>(v.quats)[g] = make_float4(V13[9], V13[10], V13[11], V13[12]);
        p.opac[g] = po, m.opac[g] = mo, v.opac[g] = vo;
        // state byte: bit 0 = has optimiser state (from now on), bit 1 = the state was created by this step (moments start from zero),
        // bit 2 = received a gradient in this step.  k_adam_rest takes its per-Gaussian decisions from this one byte.
        touched[g] = (unsigned char)(1 | (had ? 0 : 2) | (vis ? 4 : 0));
    }
}

__device__ __forceinline__ void adam_rest4(float4 &P, float4 &M, float4 &V, int i4, int N, const float *__restrict__ aux,
                                           const unsigned char *__restrict__ state, const AdamStep &step)
{
    float pv[4] = {P.x, P.y, P.z, P.w}, mv[4] = {M.x, M.y, M.z, M.w}, vv[4] = {V.x, V.y, V.z, V.w};
#pragma unroll
    for (int c4 = 0; c4 < 4; c4++)
    {
        const int i = i4 * 4 + c4;
        const int gl = i / 45, e = i - gl * 45;
        if (gl >= N)
            continue;
        const float *a = aux + (size_t)gl * AUX_FLOATS;
        const int fl = __ldg(state + gl);
        if (fl == 0)
            continue;
        const int k = e / 3, c = e - k * 3; // basis index k + 1
        const float ge = (fl & 4) ? __ldg(a + k) * __ldg(a + 15 + c) : 0.f;
        float mo = (fl & 2) ? 0.f : mv[c4], vo = (fl & 2) ? 0.f : vv[c4];
        pv[c4] = adam_update(pv[c4], ge, mo, vo, step.a, step.step_size[4]);
        mv[c4] = mo, vv[c4] = vo;
    }
    P = make_float4(pv[0], pv[1], pv[2], pv[3]);
    M = make_float4(mv[0], mv[1], mv[2], mv[3]);
    V = make_float4(vv[0], vv[1], vv[2], vv[3]);
}

__global__ void __launch_bounds__(256) k_adam_rest(float4 *__restrict__ pR, float4 *__restrict__ mR, float4 *__restrict__ vR,
                                                    const float *__restrict__ aux, const unsigned char *__restrict__ state,
                                                    const int *__restrict__ nDev, AdamStep step)
{
    const int N = *nDev;
    const int total4 = (N * 45 + 3) / 4;
    const int i4a = (blockIdx.x * 256 + threadIdx.x) * 2, i4b = i4a + 1;
    if (i4a >= total4)
        return;
    // skip float4s whose (at most two) Gaussians are both inactive
    const int ga0 = (i4a * 4) / 45, ga1 = min((i4a * 4 + 3) / 45, N - 1), gb1 = min((i4b * 4 + 3) / 45, N - 1);
    const int fa = __ldg(state + ga0) | __ldg(state + ga1);
    const int fb = (i4b < total4) ? (__ldg(state + ga1) | __ldg(state + gb1)) : 0;
    float4 Pa, Ma, Va, Pb, Mb, Vb;
    if (fa)
        Pa = pR[i4a], Ma = mR[i4a], Va = vR[i4a];
    if (fb)
        Pb = pR[i4b], Mb = mR[i4b], Vb = vR[i4b];
    if (fa)
    {
        adam_rest4(Pa, Ma, Va, i4a, N, aux, state, step);
        pR[i4a] = Pa, mR[i4a] = Ma, vR[i4a] = Va;
    }
    if (fb)
    {
        adam_rest4(Pb, Mb, Vb, i4b, N, aux, state, step);
        pR[i4b] = Pb, mR[i4b] = Mb, vR[i4b] = Vb;
    }
}

__global__ void __launch_bounds__(256) k_reduce_loss(const float *__restrict__ lossTile, int T, double scale, double *out)
{
    __shared__ double s[256];
    double acc = 0.0;
    for (int i = threadIdx.x; i < T; i += 256)
        acc += (double)lossTile[i];
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int d = 128; d > 0; d >>= 1)
    {
        if (threadIdx.x < d)
            s[threadIdx.x] += s[threadIdx.x + d];
        __syncthreads();
    }
    if (threadIdx.x == 0)
        *out = s[0] * scale;
}

// ------------------------------------------------------------------------------------------------------------
// removeRedundantGs (slam/slam_pipeline.cpp:564-586) + prunePoints: keep = !(max scale < minScale | max scale > maxScale | opacity < minOpac)
__device__ __forceinline__ bool prune_keep(const ParamPtrs &p, int g, float minOpac, float minScale, float maxScale)
{
    float s = fmaxf(fmaxf(expf(p.scales[g * 3 + 0]), expf(p.scales[g * 3 + 1])), expf(p.scales[g * 3 + 2]));
    float o = 1.0f / (1.0f + expf(-p.opac[g]));
    return !((s < minScale) || (s > maxScale) || (o < minOpac));
}

__device__ __forceinline__ int block_excl_scan_1024(int v, int *warpSums, int &total)
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
        warpSums[wid] = incl;
    __syncthreads();
    if (wid == 0)
    {
        int w = warpSums[lane];
        int wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            int n = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d)
                wi += n;
        }
        warpSums[lane] = wi - w;
        if (lane == 31)
            warpSums[32] = wi;
    }
    __syncthreads();
    total = warpSums[32];
    int r = warpSums[wid] + incl - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(1024) k_prune_count(ParamPtrs p, const int *__restrict__ nDev, float minOpac, float minScale, float maxScale,
                                                       int *chunkCnt)
{
    __shared__ int ws[33];
    int g = blockIdx.x * 1024 + threadIdx.x;
    int keep = (g < *nDev) ? (int)prune_keep(p, g, minOpac, minScale, maxScale) : 0;
    int total;
    block_excl_scan_1024(keep, ws, total);
    if (threadIdx.x == 0)
        chunkCnt[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_prune_scan(int *chunkCnt, int nChunks, int *nDev, int *counters)
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
        int ex = block_excl_scan_1024(v, ws, total);
        if (i < nChunks)
            chunkCnt[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0)
            carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0)
    {
        counters[CNT_SCRATCH + 1] = *nDev; // old count, read by the scatter pass
        counters[CNT_SCRATCH] = carry;
        *nDev = carry;
    }
}

// Warp-cooperative copy of the kept rows of one parameter array: the warp's 32 source rows (width WIDTH floats, rows g0 .. g0 + 31) shrink to
// the rows named by `mask`, which land on consecutive destination rows starting at d0 (stable compaction).  Lanes walk the destination run
// float by float -- coalesced stores, and loads that are contiguous over every kept row -- instead of one strided row per thread.
// rowSel (optional, per-lane): copy kept row r only if bit r of rowSel is set (moments travel only with Gaussians that have state).
template <int WIDTH>
__device__ __forceinline__ void warp_compact_rows(float *__restrict__ dst, const float *__restrict__ src, int g0, int d0, unsigned mask,
                                                  int srcRowOfKept /* lane r: source row (0..31) of the r-th kept row */, unsigned rowSel)
{
    const int lane = threadIdx.x & 31;
    const int nKept = __popc(mask);
    const int total = nKept * WIDTH;
    // lane's float index i = lane + 32 k  ->  (row r, element e); 32 < WIDTH is not required: advance by division-free steps
    int r = lane / WIDTH, e = lane - r * WIDTH;
    for (int base = 0; base < total; base += 32)   // warp-uniform trip count: every lane takes part in the shuffle
    {
        const int i = base + lane;
        const int sr = __shfl_sync(0xffffffffu, srcRowOfKept, r & 31);
        if (i < total && ((rowSel >> (r & 31)) & 1u))
            dst[(size_t)d0 * WIDTH + i] = src[(size_t)(g0 + sr) * WIDTH + e];
        e += 32;
        while (e >= WIDTH)
            e -= WIDTH, r++;
    }
}

// Stable compaction of the survivors: parameters always; the Adam moments and the state flag travel with their Gaussian
// (removeFromOptimizer, src/raw_gs_model.cpp:744-765) when it has optimiser state at all -- a Gaussian that has not been touched
// since initOptimizers has m = v = 0 by definition and nothing to move.
__global__ void __launch_bounds__(1024) k_prune_scatter(PruneBuffers b, const int *__restrict__ counters, float minOpac, float minScale,
                                                         float maxScale, const int *__restrict__ chunkOff)
{
    __shared__ int ws[33];
    const int g = blockIdx.x * 1024 + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int nOld = counters[CNT_SCRATCH + 1];
    const int keep = (g < nOld) ? (int)prune_keep(b.p, g, minOpac, minScale, maxScale) : 0;
    int total;
    const int ex = block_excl_scan_1024(keep, ws, total);
    const int d = chunkOff[blockIdx.x] + ex;
    // per warp: kept mask, first destination row, and for lane r the source row of the r-th kept Gaussian
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    if (mask == 0)
        return;
    const int d0 = __shfl_sync(0xffffffffu, d, __ffs(mask) - 1);
    const int srcRow = __fns(mask, 0, lane + 1);   // 0xffffffff beyond the kept count (never selected)
    const int g0 = g - lane;
    unsigned char t = keep ? b.touched[g] : 0;
    if (keep)
        b.touchedOut[d] = t;
    // bit r of `sel` = the r-th kept row has optimiser state
    const unsigned stateOfLane = __ballot_sync(0xffffffffu, t != 0);
    unsigned sel = 0;
    {
        const int has = srcRow >= 0 && srcRow < 32 && ((stateOfLane >> srcRow) & 1u);
        sel = __ballot_sync(0xffffffffu, has);
    }
    const unsigned all = 0xffffffffu;
    warp_compact_rows<3>(b.pOut.means, b.p.means, g0, d0, mask, srcRow, all);
    warp_compact_rows<3>(b.pOut.scales, b.p.scales, g0, d0, mask, srcRow, all);
    warp_compact_rows<4>(b.pOut.quats, b.p.quats, g0, d0, mask, srcRow, all);
    warp_compact_rows<3>(b.pOut.dc, b.p.dc, g0, d0, mask, srcRow, all);
    warp_compact_rows<45>(b.pOut.rest, b.p.rest, g0, d0, mask, srcRow, all);
    warp_compact_rows<1>(b.pOut.opac, b.p.opac, g0, d0, mask, srcRow, all);
    if (sel)
    {
        warp_compact_rows<3>(b.mOut.means, b.m.means, g0, d0, mask, srcRow, sel);
        warp_compact_rows<3>(b.mOut.scales, b.m.scales, g0, d0, mask, srcRow, sel);
        warp_compact_rows<4>(b.mOut.quats, b.m.quats, g0, d0, mask, srcRow, sel);
        warp_compact_rows<3>(b.mOut.dc, b.m.dc, g0, d0, mask, srcRow, sel);
        warp_compact_rows<45>(b.mOut.rest, b.m.rest, g0, d0, mask, srcRow, sel);
        warp_compact_rows<1>(b.mOut.opac, b.m.opac, g0, d0, mask, srcRow, sel);
        warp_compact_rows<3>(b.vOut.means, b.v.means, g0, d0, mask, srcRow, sel);
        warp_compact_rows<3>(b.vOut.scales, b.v.scales, g0, d0, mask, srcRow, sel);
        warp_compact_rows<4>(b.vOut.quats, b.v.quats, g0, d0, mask, srcRow, sel);
        warp_compact_rows<3>(b.vOut.dc, b.v.dc, g0, d0, mask, srcRow, sel);
        warp_compact_rows<45>(b.vOut.rest, b.v.rest, g0, d0, mask, srcRow, sel);
        warp_compact_rows<1>(b.vOut.opac, b.v.opac, g0, d0, mask, srcRow, sel);
    }
}

// ------------------------------------------------------------------------------------------------------------
// launch wrappers
void project_sh_fwd(const ParamPtrs &p, const int *nDev, int nUpper, const CamParams &cam, SplatRec *recs, SplatGrad *grads, const Bins &bins,
                    int tileW, int tileH, bool forBackward, cudaStream_t st)
{
    if (nUpper <= 0)
        return;
    GS_COUNT_LAUNCHES(1);
    k_project_sh<<<cdiv(nUpper, 128), 128, 0, st>>>(p, nDev, cam, recs, grads, bins.segCount, bin_chunk_size(nUpper), tileW, tileH, bins.items,
                                                    bins.itemCap, bins.counters, forBackward ? 1 : 0);
}

void bin_tiles(const SplatRec *recs, const int *nDev, int nUpper, const Bins &bins, int tileW, int tileH, cudaStream_t st)
{
    const int T = tileW * tileH;
    GS_COUNT_LAUNCHES(4);
    k_seg_scan<<<cdiv(T, 8), 256, 0, st>>>(bins.segCount, bins.segOff, bins.tileCount, T);
    k_scan_tiles<<<1, 1024, 0, st>>>(bins.tileCount, bins.tileOffsets, T, bins.isectCap, bins.counters);
    if (nUpper > 0)
        k_scatter_tiles<<<cdiv(nUpper, 256), 256, 0, st>>>(recs, nDev, bins.tileOffsets, bins.segOff, bins.segCount, bin_chunk_size(nUpper),
                                                           bins.flatten, bins.isectCap, tileW, tileH, nUpper);
    k_sort_tiles<<<T, 256, 0, st>>>(bins.tileOffsets, bins.segOff, bins.segCount, bins.flatten, bins.flattenSorted);
}

void bwd_params_adam(const ParamPtrs &p, const ParamPtrs &m, const ParamPtrs &v, unsigned char *touched, const AdamStep &step, const int *nDev,
                     int nUpper, const CamParams &cam, const SplatRec *recs, const SplatGrad *grads, float4 *aux, const ParamPtrs *dbg,
                     int *counters, cudaStream_t st)
{
    if (nUpper <= 0)
        return;
    GS_COUNT_LAUNCHES(2);
    ParamPtrs d = dbg ? *dbg : p;
    k_bwd_params<<<cdiv(nUpper, ADAM_WARPS * 32), ADAM_WARPS * 32, 0, st>>>(p, m, v, touched, step, nDev, cam, recs, grads, aux, d, dbg ? 1 : 0,
                                                                            counters);
    const long long total4 = ((long long)nUpper * 45 + 3) / 4;
    k_adam_rest<<<(int)((total4 + 511) / 512), 256, 0, st>>>(reinterpret_cast<float4 *>(p.rest), reinterpret_cast<float4 *>(m.rest),
                                                              reinterpret_cast<float4 *>(v.rest), reinterpret_cast<const float *>(aux), touched, nDev, step);
}

void reduce_loss(const float *lossTile, int T, double scale, double *out, cudaStream_t st)
{
    GS_COUNT_LAUNCHES(1);
    k_reduce_loss<<<1, 256, 0, st>>>(lossTile, T, scale, out);
}

void prune(const PruneBuffers &b, int *nDev, int nUpper, float minOpac, float minScale, float maxScale, int *scanTmp, int *counters, cudaStream_t st)
{
    if (nUpper <= 0)
        return;
    int nChunks = cdiv(nUpper, 1024);
    GS_COUNT_LAUNCHES(3);
    k_prune_count<<<nChunks, 1024, 0, st>>>(b.p, nDev, minOpac, minScale, maxScale, scanTmp);
    k_prune_scan<<<1, 1024, 0, st>>>(scanTmp, nChunks, nDev, counters);
    k_prune_scatter<<<nChunks, 1024, 0, st>>>(b, counters, minOpac, minScale, maxScale, scanTmp);
}

} // namespace gs
