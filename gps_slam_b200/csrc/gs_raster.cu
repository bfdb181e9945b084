// gsplat-GES rasteriser, forward and Gaussian-parallel backward, hand-written for sm_100a.
//
//   k_raster_fwd   rasterize_to_pixels_fwd_ges_kernel (gsplat/rasterizer/rasterize_to_pixels_fwd_ges.cu:18-221) fused with the
//                  ref-depth clamp (src/raw_gs_model.cpp:207), the weighted-average composite with the TSDF render
//                  (src/raw_gs_model.cpp:317-326), the L1 loss (src/raw_gs_model.cpp:369-417, src/tensor_math.cpp:41-44) and
//                  the head of the backward pass (autograd of the composite and of mean|gt - rgb|): one pass over the image
//                  instead of ~25 elementwise launches.  Splat records (48 B) are staged per batch in shared memory, colours
//                  included (the reference gathers colours from global memory per pixel-splat hit).
//   k_raster_bwd   temp_bwd_kernel (gsplat/rasterizer/rasterize_to_pixels_bwd_ges_new_parallel.cu:18-201): same support (the
//                  2r x 2r pixel box of each splat, NOT the forward's tile footprint), same validity tests, same per-lane
//                  pixel assignment (32 consecutive box-linear ids per step).  A warp owns up to 64 consecutive groups of one
//                  splat and keeps the 10 gradient sums in registers across them, so there is one shuffle reduction and one
//                  store (or 10 atomics for splats wider than 45 px) per work item instead of per 32 pixels, and no idle
//                  warps (the reference launches 32x more warps than it uses).
#include <stdlib.h>

#include "common.cuh"
#include "gs.h"

namespace gs
{

// 2^x by one MUFU.EX2 (flush-to-zero; inputs here are <= 0 or rejected right after)
__device__ __forceinline__ float ex2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

constexpr int RB = 256; // splats per staged batch = threads per tile CTA

template <int MODE>
__global__ void __launch_bounds__(256, 6) k_raster_fwd(const SplatRec *__restrict__ recs, const int *__restrict__ tileOffsets,
                                                     const int *__restrict__ flattenSorted, int W, int H, int tileW, RasterIO io, float invCount,
                                                     const CommView *__restrict__ cv, int pushAll)
{
    // one array, four planes (mean/opacity | log2-conic/depth | colour | alpha-extent box): a single base register addresses all of them
    __shared__ float4 sAll[4 * RB];
    float4 *const s0 = sAll, *const s1 = sAll + RB, *const s2 = sAll + 2 * RB, *const s3 = sAll + 3 * RB;
    __shared__ unsigned char wlist[8][RB]; // per warp: batch slots whose alpha extent touches the warp's rectangle, ascending
    __shared__ float warpLoss[8];
    const int tile = blockIdx.x;
    const int tyi = tile / tileW, txi = tile - tyi * tileW;
    const int tid = threadIdx.x;
    // each warp owns an 8 x 4 pixel rectangle of the tile, so that a splat whose alpha extent misses the rectangle is skipped
    // by the whole warp with one shared-memory read and four compares
    const int wid = tid >> 5, lane = tid & 31;
    const int wx0 = txi * TILE + (wid & 1) * 8, wy0 = tyi * TILE + (wid >> 1) * 4;
    const int i = wy0 + (lane >> 3), j = wx0 + (lane & 7);
    const float rxlo = (float)wx0 + 0.5f, rxhi = (float)wx0 + 7.5f, rylo = (float)wy0 + 0.5f, ryhi = (float)wy0 + 3.5f;
    const bool inside = (i < H) && (j < W);
    const int pix = i * W + j;
    const float px = (float)j + 0.5f, py = (float)i + 0.5f;
    const int start = tileOffsets[tile], end = tileOffsets[tile + 1];

    float rdRaw = inside ? __ldg(&io.refDepth[pix]) : 0.f;
    float rd = (io.clampRef && rdRaw < 0.01f) ? 1000.0f : rdRaw;
    const float cut = rd + io.deltaDepth;
    if (MODE != RASTER_RENDER && inside)
        io.v_out[2 * pix + 1] = make_float4(cut, 0.f, 0.f, 0.f); // depth-test threshold per pixel, read back by the rasteriser backward
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, w = 0.f;

    for (int b = start; b < end; b += RB)
    {
        __syncthreads();
        int idx = b + tid;
        if (idx < end)
        {
            const SplatRec *r = recs + __ldg(&flattenSorted[idx]);
            const float4 q0 = __ldg(&r->q0), q1 = __ldg(&r->q1);
            s0[tid] = q0;
            // conic in log2 units (see k_raster_bwd): alpha = opacity * 2^-(A dx^2 + B dx dy + C dy^2)
            s1[tid] = make_float4((0.5f * 1.4426950408889634f) * q1.x, 1.4426950408889634f * q1.y, (0.5f * 1.4426950408889634f) * q1.z, q1.w);
            s2[tid] = __ldg(&r->q2);
            float ex, ey;
            if (alpha_extent(q1.x, q1.y, q1.z, q0.z, ex, ey))
                s3[tid] = make_float4(q0.x - ex, q0.x + ex, q0.y - ey, q0.y + ey);
            else
                s3[tid] = make_float4(1e30f, -1e30f, 1e30f, -1e30f);
        }
        __syncthreads();
        const int n = min(RB, end - b);
        // the warp tests the batch against its rectangle cooperatively (one splat per lane, 8 rounds) and keeps the survivors
        // in order; the pixel loop below then only visits splats that can touch one of the warp's 32 pixels
        int ns = 0;
#pragma unroll
        for (int k = 0; k < RB / 32; k++)
        {
            const int t = k * 32 + lane;
            bool pass = false;
            if (t < n)
            {
                const float4 bb = s3[t];
                pass = !(bb.y < rxlo || bb.x > rxhi || bb.w < rylo || bb.z > ryhi);
            }
            const unsigned m = __ballot_sync(0xffffffffu, pass);
            if (pass)
                wlist[wid][ns + __popc(m & ((1u << lane) - 1u))] = (unsigned char)t;
            ns += __popc(m);
        }
        __syncwarp();
        if (inside)
        {
            for (int q = 0; q < ns; q++)
            {
                const float4 *rec = sAll + wlist[wid][q];
                const float4 c = rec[RB];
                const float4 xyo = rec[0];
                const float dx = xyo.x - px, dy = xyo.y - py;
                const float sigma = fmaf(dx, fmaf(c.y, dy, c.x * dx), (c.z * dy) * dy);
                const float alpha = fminf(0.999f, xyo.z * ex2_approx(-sigma));
                // (the depth cut rejects well under 1% of the candidates: folded into the one branch)
                if (sigma < 0.f || alpha < 1.f / 255.f || c.w > cut)
                    continue;
                const float4 col = rec[2 * RB];
                a0 += col.x * alpha;
                a1 += col.y * alpha;
                a2 += col.z * alpha;
                a3 += c.w * alpha;
                w += alpha;
            }
        }
    }

    if (MODE == RASTER_RAW)
    {
        if (inside)
        {
            reinterpret_cast<float4 *>(io.render4)[pix] = make_float4(a0, a1, a2, a3);
            io.alphas[pix] = w;
        }
        return;
    }
    if (MODE == RASTER_PUSH)
    {
        // the reduce-scatter IS this store: the tile's partial sums go to slot `rank` of the gather region of the rank that owns the
        // tile (or of every rank, for a render that every rank needs in full), in tile-major order; a warp's 8 x 4 rectangle is four
        // 128-byte runs of float4
        if (inside)
        {
            const size_t t256 = cv->slotFloats / 5;
            const size_t off = (size_t)tile * 256 + (size_t)((i - tyi * TILE) * TILE + (j - txi * TILE));
            const size_t slot = (size_t)cv->rank * cv->slotFloats;
            const float4 acc = make_float4(a0, a1, a2, a3);
            const int q0 = pushAll ? 0 : comm_owner(*cv, tile), q1 = pushAll ? cv->world : q0 + 1;
            for (int q = q0; q < q1; q++)
            {
                float *base = cv->gather[q] + slot;
                reinterpret_cast<float4 *>(base)[off] = acc;
                base[4 * t256 + off] = w;
            }
        }
        return;
    }
    // composite with the TSDF render: rgb = (acc + base) / (w + 1), depth = (acc_d + D*[D>0]) / (w + [D>0])
    float r0 = 0.f, r1 = 0.f, r2 = 0.f;
    float w1 = w + 1.0f;
    if (inside)
    {
        const float *bc = io.baseColor + (size_t)pix * 3;
        r0 = (a0 + __ldg(bc + 0) * 1.0f) / w1;
        r1 = (a1 + __ldg(bc + 1) * 1.0f) / w1;
        r2 = (a2 + __ldg(bc + 2) * 1.0f) / w1;
    }
    if (MODE == RASTER_RENDER)
    {
        if (inside)
        {
            float bw = rdRaw > 0.f ? 1.0f : 0.0f;
            float *o = io.rgb + (size_t)pix * 3;
            o[0] = r0, o[1] = r1, o[2] = r2;
            io.depth[pix] = (a3 + rdRaw * bw) / (w + bw);
            io.alphas[pix] = w;
        }
        return;
    }
    // TRAIN: loss = mean |gt - rgb| over 3P elements; v_rgb = sign(rgb - gt) / (3P);
    //        v_render_c = v_rgb_c / (w + 1); v_render_alpha = -sum_c v_rgb_c * rgb_c / (w + 1)
    float lsum = 0.f;
    if (inside)
    {
        const float *gt = io.gt + (size_t)pix * 3;
        float d0 = r0 - __ldg(gt + 0), d1 = r1 - __ldg(gt + 1), d2 = r2 - __ldg(gt + 2);
        lsum = fabsf(d0) + fabsf(d1) + fabsf(d2);
        float v0 = d0 > 0.f ? invCount : (d0 < 0.f ? -invCount : 0.f);
        float v1 = d1 > 0.f ? invCount : (d1 < 0.f ? -invCount : 0.f);
        float v2 = d2 > 0.f ? invCount : (d2 < 0.f ? -invCount : 0.f);
        float va = -(v0 * r0 + v1 * r1 + v2 * r2) / w1;
        io.v_out[2 * pix] = make_float4(v0 / w1, v1 / w1, v2 / w1, va);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
        lsum += __shfl_xor_sync(0xffffffffu, lsum, d);
    if ((tid & 31) == 0)
        warpLoss[tid >> 5] = lsum;
    __syncthreads();
    if (tid == 0)
    {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; k++)
            s += warpLoss[k];
        io.lossTile[tile] = s;
    }
}

// ------------------------------------------------------------------------------------------------------------
// Sum 16 per-lane values over each HALF warp (lanes 0-15 and 16-31 independently) with 15 shuffles: at each halving step a lane
// keeps one half of its values and hands the other half to its partner.  Returns, in lane l, the half-warp total of slot l & 15.
__device__ __forceinline__ float halfwarp_reduce16(float (&v)[16], int lane)
{
    const unsigned full = 0xffffffffu;
    float w8[8], w4[4], w2[2];
    const bool h8 = lane & 8, h4 = lane & 4, h2 = lane & 2, h1 = lane & 1;
#pragma unroll
    for (int i = 0; i < 8; i++)
    {
        float send = h8 ? v[i] : v[i + 8], keep = h8 ? v[i + 8] : v[i];
        w8[i] = keep + __shfl_xor_sync(full, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
    {
        float send = h4 ? w8[i] : w8[i + 4], keep = h4 ? w8[i + 4] : w8[i];
        w4[i] = keep + __shfl_xor_sync(full, send, 4);
    }
#pragma unroll
    for (int i = 0; i < 2; i++)
    {
        float send = h2 ? w4[i] : w4[i + 2], keep = h2 ? w4[i + 2] : w4[i];
        w2[i] = keep + __shfl_xor_sync(full, send, 2);
    }
    float send = h1 ? w2[0] : w2[1], keep = h1 ? w2[1] : w2[0];
    return keep + __shfl_xor_sync(full, send, 1);
}

// Rasteriser backward.  A HALF warp owns one work item (one splat, <= 2048 pixels of its rectangle): the two halves of a warp walk
// two items side by side, 2 x 16 rect-linear pixels per step each, keep the 10 gradient sums in registers and finish with one
// half-warp reduction and one store (or 10 atomics when the splat has several items).  Compared with a whole warp per item this
// halves the per-item overhead (record loads, rectangle setup, reduction) and wastes fewer lanes on the typical 100-200 pixel
// rectangles.  HAS_VD: a depth-channel gradient image is supplied (depth_weight > 0; off in every release config).
// (A quarter warp per item -- four items side by side, so that the ~450 instructions a warp spends per batch of items outside the pixel
// loop are shared by four items instead of two -- was built and measured: 336 us against 294 us.  The warp runs for as long as its
// longest item, and four neighbours differ more than two.)
template <bool HAS_VD, bool STATS = false>
__global__ void __launch_bounds__(256, 3) k_raster_bwd(const SplatRec *__restrict__ recs, const int4 *__restrict__ items, int *counters, int itemCap,
                                                        int W, const float4 *__restrict__ v_out, const float *__restrict__ v_depthImg,
                                                        SplatGrad *__restrict__ grads)
{
    const unsigned full = 0xffffffffu;
    constexpr int LPI = 16;          // lanes per item
    constexpr int IPW = 32 / LPI;    // items per warp
    constexpr int PXS = 2 * LPI;     // rect-linear pixels of one item per step
    const int lane = threadIdx.x & 31, hl = lane & (LPI - 1), half = lane / LPI;
    const int nItems = min(counters[CNT_ITEMS], itemCap);
    // dynamic distribution of work items over the resident warps: items differ by up to 64x in cost
    int *cursor = counters + CNT_BWD_CURSOR;
    constexpr int GRAB = 8; // consecutive items per cursor bump (same-address atomics are the scarce resource)
    int it = 0, itEnd = 0;
    for (;;)
    {
        if (it >= itEnd)
        {
            // (prefetching the next batch's cursor bump while a batch runs was measured 12 % slower: it coarsens the distribution)
            if (lane == 0)
                it = atomicAdd(cursor, GRAB);
            it = __shfl_sync(full, it, 0);
            if (it >= nItems)
                break;
            itEnd = min(it + GRAB, nItems);
        }
        const int mine = it + half;
        const bool live = mine < itEnd;
        it += IPW;
        const int4 item = __ldg(&items[live ? mine : itEnd - 1]);
        const int g = item.x;
        const float4 q0 = __ldg(&recs[g].q0), q1 = __ldg(&recs[g].q1), q2 = __ldg(&recs[g].q2);
        const float opac = q0.z;
        // sigma in log2 units: alpha = opac * 2^-(A dx^2 + B dx dy + C dy^2)
        const float A = (0.5f * 1.4426950408889634f) * q1.x, B = 1.4426950408889634f * q1.y, C = (0.5f * 1.4426950408889634f) * q1.z;
        // the pixel rectangle was computed once by the projection pass (bwd_rect)
        const int rx = item.z & 0xffff, ry = item.z >> 16, rw = item.w & 0xffff, rh = item.w >> 16;
        const int npix = rw * rh;
        const float inv_rw = 1.0f / (float)rw;
        const int p0 = item.y;
        const int p1 = live ? min(p0 + BWD_PIXELS_PER_ITEM, npix) : p0;
        // Each lane walks two pixel sequences, ids p0 + hl + PXS k and p0 + LPI + hl + PXS k, in rect-linear order.  Pixel centre
        // (px, py) and image index are advanced incrementally: +PXS ids = +q32 rows, +r32 columns with at most one wrap.
        const int q32 = (int)(((float)PXS + 0.5f) * inv_rw), r32 = PXS - q32 * rw; // exact: rw <= 200
        const float r32f = (float)r32, q32f = (float)q32, rwf = (float)rw;
        const float xEnd = (float)(rx + rw);                       // first pixel centre beyond the rect is xEnd + 0.5
        const int dpix = q32 * W + r32, dwrap = W - rw;
        float px[2], py[2];
        int pix[2];
#pragma unroll
        for (int u = 0; u < 2; u++)
        {
            const int id = p0 + u * LPI + hl;
            const int row = (int)(((float)id + 0.5f) * inv_rw); // exact for id < 2^16, rw <= 200
            const int col = id - row * rw;
            px[u] = (float)(rx + col) + 0.5f, py[u] = (float)(ry + row) + 0.5f;
            pix[u] = (ry + row) * W + rx + col;
        }
        // steps of the longest of the warp's items
        int steps = (p1 - p0 + PXS - 1) / PXS;
        steps = max(steps, __shfl_xor_sync(full, steps, 16));
        // gradient sums; the conic / mean gradients are accumulated as moments of t = v_sigma over (dx, dy):
        // v_conic = (Sxx/2, Sxy, Syy/2), v_mean2d = (a Sx + b Sy, b Sx + c Sy)
        float vr = 0.f, vg = 0.f, vb = 0.f, vd = 0.f, sxx = 0.f, sxy = 0.f, syy = 0.f, sx = 0.f, sy = 0.f, vo = 0.f;
        int nTested = 0, nPassed = 0; // STATS build only
        // Per-pixel record of the backward: 32 bytes = (dL/d render rgb, dL/d alpha | depth cut, -, -, -), written by the forward, so one
        // address serves both reads.  Software pipeline, unrolled by two so that the in-flight buffers alternate without register
        // copies: the reads of step k+1 are issued, unconditionally for every pixel of the rectangle, before the arithmetic of step k.
        int id0 = p0 + hl;
        struct PixData
        {
            float4 vo[2];
            float cut[2], vdp[2], dx[2], dy[2];
        };
        auto fetch = [&](PixData &d, int idBase)
        {
#pragma unroll
            for (int u = 0; u < 2; u++)
            {
                d.dx[u] = q0.x - px[u], d.dy[u] = q0.y - py[u];
                d.cut[u] = -1e30f, d.vdp[u] = 0.f, d.vo[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (idBase + u * LPI < p1)
                {
                    const float4 *rec = v_out + 2 * (size_t)pix[u];
                    d.vo[u] = __ldg(rec);
                    d.cut[u] = __ldg(reinterpret_cast<const float *>(rec + 1));
                    if (HAS_VD)
                        d.vdp[u] = __ldg(&v_depthImg[pix[u]]);
                }
                // advance to the pixel PXS ids further
                px[u] += r32f, py[u] += q32f, pix[u] += dpix;
                if (px[u] > xEnd)
                    px[u] -= rwf, py[u] += 1.0f, pix[u] += dwrap;
            }
        };
        auto consume = [&](const PixData &d)
        {
#pragma unroll
            for (int u = 0; u < 2; u++)
            {
                const float dx = d.dx[u], dy = d.dy[u], vdp = d.vdp[u];
                const float s2 = fmaf(dx, fmaf(B, dy, A * dx), (C * dy) * dy);
                const float vis = ex2_approx(-s2);
                const float ar = opac * vis;
                const float alpha = fminf(0.999f, ar);
                // (pixels beyond the item carry cut = -1e30 and fail the depth test)
                if (STATS)
                    nTested += d.cut[u] > -1e29f;
                if (s2 < 0.f || alpha < 1.f / 255.f || q1.w > d.cut[u])
                    continue;
                if (STATS)
                    nPassed++;
                const float4 vo4 = d.vo[u];
                vr += alpha * vo4.x;
                vg += alpha * vo4.y;
                vb += alpha * vo4.z;
                float v_alpha = q2.x * vo4.x + q2.y * vo4.y + q2.z * vo4.z + vo4.w;
                if (HAS_VD)
                {
                    vd += alpha * vdp;
                    v_alpha += q1.w * vdp;
                }
                if (ar <= 0.999f)
                {
                    const float qv = vis * v_alpha;
                    const float t = -opac * qv;
                    const float tx = t * dx, ty = t * dy;
                    sxx += tx * dx;
                    sxy += tx * dy;
                    syy += ty * dy;
                    sx += tx;
                    sy += ty;
                    vo += qv;
                }
            }
        };
        PixData dA, dB;
        fetch(dA, id0);
        for (int k = 0; k < steps; k += 2, id0 += 2 * PXS)
        {
            fetch(dB, id0 + PXS);
            consume(dA);
            fetch(dA, id0 + 2 * PXS);
            consume(dB); // (a step beyond `steps` only sees pixels beyond the item: nothing passes)
        }
        // slots follow the float layout of SplatGrad: (vx, vy, vo, vd | vca, vcb, vcc, - | vr, vg, vb, -)
        float v16[16];
        v16[0] = q1.x * sx + q1.y * sy, v16[1] = q1.y * sx + q1.z * sy, v16[2] = vo, v16[3] = vd;
        v16[4] = 0.5f * sxx, v16[5] = sxy, v16[6] = 0.5f * syy, v16[7] = 0.f;
        v16[8] = vr, v16[9] = vg, v16[10] = vb, v16[11] = 0.f;
        v16[12] = v16[13] = v16[14] = v16[15] = 0.f;
        const float total = halfwarp_reduce16(v16, lane);
        if (STATS)
        {
            for (int o = 16; o; o >>= 1)
                nTested += __shfl_xor_sync(full, nTested, o), nPassed += __shfl_xor_sync(full, nPassed, o);
            if (lane == 0)
            {
                atomicAdd(reinterpret_cast<unsigned long long *>(counters + CNT_PAIRS_TESTED), (unsigned long long)nTested);
                atomicAdd(reinterpret_cast<unsigned long long *>(counters + CNT_PAIRS_PASSED), (unsigned long long)nPassed);
            }
        }
        if (live && hl < 11 && hl != 7)
        {
            float *f = reinterpret_cast<float *>(grads + g) + hl;
            if (__float_as_int(q2.w) & 256)
                atomicAdd(f, total);
            else
                *f = total;
        }
    }
}

// epilogue of the forward pass on an already accumulated image (multi-GPU: the per-rank partial sums were all-reduced):
// acc5 = render_colors [P,4] followed by alphas [P].  Same arithmetic as the tail of k_raster_fwd.
template <int MODE>
__global__ void __launch_bounds__(256) k_composite(const float *__restrict__ acc5, int W, int H, int tileW, RasterIO io, float invCount)
{
    __shared__ float warpLoss[8];
    const int tile = blockIdx.x;
    const int tyi = tile / tileW, txi = tile - tyi * tileW;
    const int tid = threadIdx.x;
    const int i = tyi * TILE + (tid >> 4), j = txi * TILE + (tid & 15);
    const bool inside = (i < H) && (j < W);
    const int pix = i * W + j;
    const int P = W * H;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, w = 0.f, rdRaw = 0.f;
    if (inside)
    {
        float4 a = __ldg(reinterpret_cast<const float4 *>(acc5) + pix);
        a0 = a.x, a1 = a.y, a2 = a.z, a3 = a.w;
        w = __ldg(acc5 + (size_t)4 * P + pix);
        rdRaw = __ldg(&io.refDepth[pix]);
    }
    float r0 = 0.f, r1 = 0.f, r2 = 0.f;
    float w1 = w + 1.0f;
    if (inside)
    {
        const float *bc = io.baseColor + (size_t)pix * 3;
        r0 = (a0 + __ldg(bc + 0) * 1.0f) / w1;
        r1 = (a1 + __ldg(bc + 1) * 1.0f) / w1;
        r2 = (a2 + __ldg(bc + 2) * 1.0f) / w1;
    }
    if (MODE == RASTER_RENDER)
    {
        if (inside)
        {
            float bw = rdRaw > 0.f ? 1.0f : 0.0f;
            float *o = io.rgb + (size_t)pix * 3;
            o[0] = r0, o[1] = r1, o[2] = r2;
            io.depth[pix] = (a3 + rdRaw * bw) / (w + bw);
            io.alphas[pix] = w;
        }
        return;
    }
    float lsum = 0.f;
    if (inside)
    {
        io.v_out[2 * pix + 1] = make_float4(((io.clampRef && rdRaw < 0.01f) ? 1000.0f : rdRaw) + io.deltaDepth, 0.f, 0.f, 0.f);
        const float *gt = io.gt + (size_t)pix * 3;
        float d0 = r0 - __ldg(gt + 0), d1 = r1 - __ldg(gt + 1), d2 = r2 - __ldg(gt + 2);
        lsum = fabsf(d0) + fabsf(d1) + fabsf(d2);
        float v0 = d0 > 0.f ? invCount : (d0 < 0.f ? -invCount : 0.f);
        float v1 = d1 > 0.f ? invCount : (d1 < 0.f ? -invCount : 0.f);
        float v2 = d2 > 0.f ? invCount : (d2 < 0.f ? -invCount : 0.f);
        float va = -(v0 * r0 + v1 * r1 + v2 * r2) / w1;
        io.v_out[2 * pix] = make_float4(v0 / w1, v1 / w1, v2 / w1, va);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
        lsum += __shfl_xor_sync(0xffffffffu, lsum, d);
    if ((tid & 31) == 0)
        warpLoss[tid >> 5] = lsum;
    __syncthreads();
    if (tid == 0)
    {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; k++)
            s += warpLoss[k];
        io.lossTile[tile] = s;
    }
}

// Multi-GPU epilogue (gs_comm.h): sums the G gather slots of a tile in rank order and finishes like k_composite.  TRAIN: launched over
// the tiles this rank owns; the gradient record and the tile loss are stored into EVERY rank's v_out / lossTile.  RENDER: launched over
// all tiles (every rank received every tile), outputs are local.
template <int MODE>
__global__ void __launch_bounds__(256) k_composite_x(const CommView *__restrict__ cv, int tile0, int W, int H, int tileW, RasterIO io, float invCount)
{
    __shared__ float warpLoss[8];
    const int tile = tile0 + blockIdx.x;
    const int tyi = tile / tileW, txi = tile - tyi * tileW;
    const int tid = threadIdx.x;
    const int i = tyi * TILE + (tid >> 4), j = txi * TILE + (tid & 15);
    const bool inside = (i < H) && (j < W);
    const int pix = i * W + j;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, w = 0.f, rdRaw = 0.f;
    if (inside)
    {
        const size_t t256 = cv->slotFloats / 5;
        const size_t off = (size_t)tile * 256 + tid;
        const float *mine = cv->gather[cv->rank];
        for (int r = 0; r < cv->world; r++)
        {
            const float *base = mine + (size_t)r * cv->slotFloats;
            const float4 a = __ldcg(reinterpret_cast<const float4 *>(base) + off);
            a0 += a.x, a1 += a.y, a2 += a.z, a3 += a.w;
            w += __ldcg(base + 4 * t256 + off);
        }
        rdRaw = __ldg(&io.refDepth[pix]);
    }
    float r0 = 0.f, r1 = 0.f, r2 = 0.f;
    float w1 = w + 1.0f;
    if (inside)
    {
        const float *bc = io.baseColor + (size_t)pix * 3;
        r0 = (a0 + __ldg(bc + 0) * 1.0f) / w1;
        r1 = (a1 + __ldg(bc + 1) * 1.0f) / w1;
        r2 = (a2 + __ldg(bc + 2) * 1.0f) / w1;
    }
    if (MODE == RASTER_RENDER)
    {
        if (inside)
        {
            float bw = rdRaw > 0.f ? 1.0f : 0.0f;
            float *o = io.rgb + (size_t)pix * 3;
            o[0] = r0, o[1] = r1, o[2] = r2;
            io.depth[pix] = (a3 + rdRaw * bw) / (w + bw);
            io.alphas[pix] = w;
        }
        return;
    }
    float lsum = 0.f;
    if (inside)
    {
        const float *gt = io.gt + (size_t)pix * 3;
        float d0 = r0 - __ldg(gt + 0), d1 = r1 - __ldg(gt + 1), d2 = r2 - __ldg(gt + 2);
        lsum = fabsf(d0) + fabsf(d1) + fabsf(d2);
        float v0 = d0 > 0.f ? invCount : (d0 < 0.f ? -invCount : 0.f);
        float v1 = d1 > 0.f ? invCount : (d1 < 0.f ? -invCount : 0.f);
        float v2 = d2 > 0.f ? invCount : (d2 < 0.f ? -invCount : 0.f);
        float va = -(v0 * r0 + v1 * r1 + v2 * r2) / w1;
        const float4 rec = make_float4(v0 / w1, v1 / w1, v2 / w1, va);
        for (int q = 0; q < cv->world; q++)
            cv->vout[q][2 * (size_t)pix] = rec; // the all-gather: one 16-byte store per peer (the depth cut next to it is rank-local)
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
        lsum += __shfl_xor_sync(0xffffffffu, lsum, d);
    if ((tid & 31) == 0)
        warpLoss[tid >> 5] = lsum;
    __syncthreads();
    if (tid < cv->world)
    {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; k++)
            s += warpLoss[k];
        cv->lossTile[tid][tile] = s;
    }
}

__global__ void k_pack_v_out(int P, const float *__restrict__ v_render4, const float *__restrict__ v_alphas, float4 *v_out, float *v_depth)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P)
        return;
    float4 r = reinterpret_cast<const float4 *>(v_render4)[i];
    v_out[2 * i] = make_float4(r.x, r.y, r.z, v_alphas[i]);
    v_depth[i] = r.w;
}

// ------------------------------------------------------------------------------------------------------------
void raster_fwd(int mode, const SplatRec *recs, const Bins &bins, int W, int H, int tileW, int tileH, const RasterIO &io, cudaStream_t st)
{
    const int T = tileW * tileH;
    const float invCount = 1.0f / (float)((size_t)3 * W * H);
    GS_COUNT_LAUNCHES(1);
    if (mode == RASTER_RAW)
        k_raster_fwd<RASTER_RAW><<<T, 256, 0, st>>>(recs, bins.tileOffsets, bins.flattenSorted, W, H, tileW, io, invCount, nullptr, 0);
    else if (mode == RASTER_RENDER)
        k_raster_fwd<RASTER_RENDER><<<T, 256, 0, st>>>(recs, bins.tileOffsets, bins.flattenSorted, W, H, tileW, io, invCount, nullptr, 0);
    else
        k_raster_fwd<RASTER_TRAIN><<<T, 256, 0, st>>>(recs, bins.tileOffsets, bins.flattenSorted, W, H, tileW, io, invCount, nullptr, 0);
}

// multi-GPU forward: partial sums into peer gather slots (cvDev: the CommView in device memory)
void raster_fwd_push(const SplatRec *recs, const Bins &bins, int W, int H, int tileW, int tileH, const RasterIO &io, const CommView *cvDev, bool pushAll,
                     cudaStream_t st)
{
    const int T = tileW * tileH;
    GS_COUNT_LAUNCHES(1);
    k_raster_fwd<RASTER_PUSH><<<T, 256, 0, st>>>(recs, bins.tileOffsets, bins.flattenSorted, W, H, tileW, io, 0.f, cvDev, pushAll ? 1 : 0);
}

void composite_exchange(int mode, const CommView &cv, const CommView *cvDev, int W, int H, int tileW, int tileH, const RasterIO &io, cudaStream_t st)
{
    const int T = tileW * tileH;
    const float invCount = 1.0f / (float)((size_t)3 * W * H);
    GS_COUNT_LAUNCHES(1);
    if (mode == RASTER_RENDER)
        k_composite_x<RASTER_RENDER><<<T, 256, 0, st>>>(cvDev, 0, W, H, tileW, io, invCount);
    else
    {
        const int t0 = cv.rank * cv.tilesPerRank, t1 = min(T, cv.rank == cv.world - 1 ? T : t0 + cv.tilesPerRank);
        if (t1 > t0)
            k_composite_x<RASTER_TRAIN><<<t1 - t0, 256, 0, st>>>(cvDev, t0, W, H, tileW, io, invCount);
    }
}

void raster_bwd(const SplatRec *recs, const Bins &bins, int W, int H, const RasterIO &io, const float *v_depth, SplatGrad *grads, cudaStream_t st)
{
    GS_COUNT_LAUNCHES(1);
    // (the work cursor CNT_BWD_CURSOR is zeroed by the binning pass of the same iteration, or by the staged entry point)
    // persistent grid: exactly the resident CTAs (148 SMs x 3 per SM at 80 registers); work is handed out by the device-side cursor
    static int ctasPerSm[2] = {0, 0};
    const int v = v_depth ? 1 : 0;
    if (!ctasPerSm[v])
    {
        int n = 0;
        if (v)
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_raster_bwd<true>, 256, 0);
        else
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_raster_bwd<false>, 256, 0);
        ctasPerSm[v] = n > 0 ? n : 3;
    }
    static int sms = 0; // one process per GPU
    if (!sms)
    {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
            sms = 148;
    }
    const int grid = sms * ctasPerSm[v];
    // (4 CTAs per SM at 64 registers -- 112 bytes of spills -- was measured: 406 us against 294 us at 3 CTAs / 80 registers)
    if (v_depth)
        k_raster_bwd<true><<<grid, 256, 0, st>>>(recs, bins.items, bins.counters, bins.itemCap, W, io.v_out, v_depth, grads);
    else
        k_raster_bwd<false><<<grid, 256, 0, st>>>(recs, bins.items, bins.counters, bins.itemCap, W, io.v_out, nullptr, grads);
}

// the same kernel with its pair counters compiled in (measurement aid: how many (pixel, splat) pairs the backward evaluates and how many
// pass -- counters[CNT_PAIRS_TESTED / CNT_PAIRS_PASSED], zeroed here)
void raster_bwd_stats(const SplatRec *recs, const Bins &bins, int W, int H, const RasterIO &io, SplatGrad *grads, cudaStream_t st)
{
    GS_COUNT_LAUNCHES(1);
    cudaMemsetAsync(bins.counters + CNT_PAIRS_TESTED, 0, 4 * sizeof(int), st);
    k_raster_bwd<false, true><<<148 * 3, 256, 0, st>>>(recs, bins.items, bins.counters, bins.itemCap, W, io.v_out, nullptr, grads);
}

void composite(int mode, const float *acc5, int W, int H, int tileW, int tileH, const RasterIO &io, cudaStream_t st)
{
    const int T = tileW * tileH;
    const float invCount = 1.0f / (float)((size_t)3 * W * H);
    GS_COUNT_LAUNCHES(1);
    if (mode == RASTER_RENDER)
        k_composite<RASTER_RENDER><<<T, 256, 0, st>>>(acc5, W, H, tileW, io, invCount);
    else
        k_composite<RASTER_TRAIN><<<T, 256, 0, st>>>(acc5, W, H, tileW, io, invCount);
}

void pack_v_out(int P, const float *v_render4, const float *v_alphas, float4 *v_out, float *v_depth, cudaStream_t st)
{
    GS_COUNT_LAUNCHES(1);
    k_pack_v_out<<<(P + 255) / 256, 256, 0, st>>>(P, v_render4, v_alphas, v_out, v_depth);
}

} // namespace gs
