// Gaussian model object + its C ABI (include/gpsslam_b200.h, section A).
// Host logic mirrors RawGaussianModel / SLAMGaussianModel (reference include/raw_gs_model.h:8-298, src/raw_gs_model.cpp:188-417,
// 654-705; slam/slam_gs_model.cpp:5-56) for the GES render method: gesForward, computeLoss, backward, optimizersStep,
// initOptimizers, prunePoints, add.  One training iteration is 7 asynchronous launches on one stream with no host round trip
// (the reference: ~40-60 launches and 3 host syncs).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <utility>
#include <vector>

#include "../../include/gpsslam_b200.h"
#include "common.cuh"
#include "gs.h"
#include "gs_comm.h"
#include "gs_spawn.h"

using namespace gs;

struct gsb_gs
{
    gsb_gs_config_t cfg;
    cudaStream_t stream, ownStream;
    int W, H, tileW, tileH, T;
    int cap;         // Gaussian capacity
    int nUpper;      // host-side upper bound of the Gaussian count (exact after gsb_gs_count)
    int *nDev;       // device-side exact count
    ParamPtrs p, tmp, m, v, mTmp, vTmp, dbg; // tmp / mTmp / vTmp: the other half of the prune's double buffers
    unsigned char *touched, *touchedTmp;
    SplatRec *recs;
    SplatGrad *grads;
    float4 *aux;     // [cap*5] per-Gaussian SH basis / colour gradient / state flag between the two backward kernels
    Bins bins;
    float4 *v_out, *vOutOwn;   // vOutOwn / lossTileOwn: the engine's own buffers (v_out / lossTile point into the exchange segment with a communicator)
    float *v_depth;
    float *lossTile, *lossTileOwn;
    double *lossDev;
    int *scanTmp;
    SpawnBuffers sb;
    float *spRgb, *spDepth, *spAlpha; // render of the spawn camera (initNewGaussians)
    int adamStep;
    int *hostInts;   // pinned [8]
    double *hostLoss; // pinned [2]: [0] gsb_gs_loss, [1] gsb_gs_loss_begin / gsb_gs_loss_end
    cudaEvent_t lossEv; // completion of the copy started by gsb_gs_loss_begin
    bool lossPending;
    bool haveDbg;
    CamParams lastCam; // camera / image set of the last train step or stage-0 call (gsb_gs_run_stage)
    RasterIO lastIo;
    bool haveLast;
    gsb_comm *comm;  // multi-GPU exchange (gsb_gs_set_comm); nullptr on one GPU
    void *knnWs;     // gs_knn.cu workspace, allocated on the first gsb_gs_dist_cuda2 call
    int knnCap;
    std::vector<void *> allocs;
};

template <typename T>
static int dev_alloc(gsb_gs *e, T **p, size_t n)
{
    GS_CUDA_OK(cudaMalloc((void **)p, (n ? n : 1) * sizeof(T)));
    e->allocs.push_back((void *)*p);
    // cudaMalloc hands back recycled, uncleared memory: every buffer starts from zero so that nothing (dropped work items after a
    // capacity overflow, padding rows past the Gaussian count) can ever depend on what a previous owner left there
    GS_CUDA_OK(cudaMemset((void *)*p, 0, (n ? n : 1) * sizeof(T)));
    return 0;
}

static int alloc_params(gsb_gs *e, ParamPtrs &q, size_t cap)
{
    int rc = 0;
    rc |= dev_alloc(e, &q.means, cap * 3);
    rc |= dev_alloc(e, &q.scales, cap * 3);
    rc |= dev_alloc(e, &q.quats, cap * 4);
    rc |= dev_alloc(e, &q.dc, cap * 3);
    rc |= dev_alloc(e, &q.rest, cap * 45);
    rc |= dev_alloc(e, &q.opac, cap);
    return rc;
}

extern "C" void gsb_gs_default_config(gsb_gs_config_t *c)
{
    memset(c, 0, sizeof *c);
    c->width = 1200, c->height = 680;
    c->capacity = 1 << 21;
    c->isect_capacity = 1 << 24;
    c->item_capacity = 1 << 23;
    c->max_gs_radii = 100;        // MODEL.max_gs_radii      (configs/release/replica/office0.yaml:83)
    c->delta_depth = 0.1f;        // MODEL.delta_depth
    c->eps2d = 0.3f, c->near_plane = 0.01f, c->far_plane = 1e10f, c->radius_clip = 0.0f; // include/raw_gs_model.h defaults
    c->lr_means = 0.00016f, c->lr_scales = 0.005f, c->lr_quats = 0.001f, c->lr_dc = 0.0025f, c->lr_rest = 0.0005f, c->lr_opac = 0.05f;
    c->scene_scale = 1.0f;
    c->device = 0;
}

extern "C" void gsb_gs_destroy(gsb_gs_t *e)
{
    if (!e)
        return;
    cudaStreamSynchronize(e->stream);
    for (void *q : e->allocs)
        cudaFree(q);
    if (e->hostInts)
        cudaFreeHost(e->hostInts);
    if (e->hostLoss)
        cudaFreeHost(e->hostLoss);
    if (e->lossEv)
        cudaEventDestroy(e->lossEv);
    cudaStreamDestroy(e->ownStream);
    delete e;
}

extern "C" int gsb_gs_create(const gsb_gs_config_t *cfg, gsb_gs_t **out)
{
    if (!cfg || !out)
        return gs_set_error(__FILE__, __LINE__, "null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return gs_set_error(__FILE__, __LINE__, "no CUDA device: gpsslam_b200 has no CPU fallback");
    if (cfg->width <= 0 || cfg->height <= 0 || cfg->capacity <= 0)
        return gs_set_error(__FILE__, __LINE__, "invalid configuration");
    GS_CUDA_OK(cudaSetDevice(cfg->device));
    gsb_gs *e = new (std::nothrow) gsb_gs();
    if (!e)
        return gs_set_error(__FILE__, __LINE__, "out of host memory");
    e->cfg = *cfg;
    e->W = cfg->width, e->H = cfg->height;
    e->tileW = (e->W + TILE - 1) / TILE, e->tileH = (e->H + TILE - 1) / TILE, e->T = e->tileW * e->tileH;
    e->cap = (cfg->capacity + 127) / 128 * 128;
    e->nUpper = 0;
    e->adamStep = 0;
    e->haveDbg = false;
    e->haveLast = false;
    e->comm = nullptr;
    e->hostInts = nullptr, e->hostLoss = nullptr;
    GS_CUDA_OK(cudaStreamCreateWithFlags(&e->ownStream, cudaStreamNonBlocking));
    e->stream = e->ownStream;
    const size_t P = (size_t)e->W * e->H;
    int rc = 0;
    rc |= alloc_params(e, e->p, e->cap);
    rc |= alloc_params(e, e->tmp, e->cap);
    rc |= alloc_params(e, e->m, e->cap);
    rc |= alloc_params(e, e->v, e->cap);
    rc |= alloc_params(e, e->mTmp, e->cap);
    rc |= alloc_params(e, e->vTmp, e->cap);
    memset(&e->dbg, 0, sizeof e->dbg);
    rc |= dev_alloc(e, &e->touched, (size_t)e->cap);
    rc |= dev_alloc(e, &e->touchedTmp, (size_t)e->cap);
    rc |= dev_alloc(e, &e->recs, (size_t)e->cap);
    rc |= dev_alloc(e, &e->grads, (size_t)e->cap);
    rc |= dev_alloc(e, &e->aux, (size_t)e->cap * 5);
    rc |= dev_alloc(e, &e->nDev, 1);
    e->bins.isectCap = cfg->isect_capacity > 0 ? cfg->isect_capacity : (1 << 24);
    e->bins.itemCap = cfg->item_capacity > 0 ? cfg->item_capacity : (1 << 23);
    rc |= dev_alloc(e, &e->bins.tileCount, (size_t)e->T + 1);
    rc |= dev_alloc(e, &e->bins.tileOffsets, (size_t)e->T + 1);
    rc |= dev_alloc(e, &e->bins.segCount, (size_t)e->T * BIN_CHUNKS);
    rc |= dev_alloc(e, &e->bins.segOff, (size_t)e->T * BIN_CHUNKS);
    rc |= dev_alloc(e, &e->bins.flatten, (size_t)e->bins.isectCap);
    rc |= dev_alloc(e, &e->bins.flattenSorted, (size_t)e->bins.isectCap);
    rc |= dev_alloc(e, &e->bins.items, (size_t)e->bins.itemCap);
    rc |= dev_alloc(e, &e->bins.counters, (size_t)CNT_TOTAL);
    rc |= dev_alloc(e, &e->v_out, 2 * P);
    e->vOutOwn = e->v_out;
    rc |= dev_alloc(e, &e->v_depth, P);
    rc |= dev_alloc(e, &e->lossTile, (size_t)e->T);
    e->lossTileOwn = e->lossTile;
    rc |= dev_alloc(e, &e->lossDev, 1);
    rc |= dev_alloc(e, &e->scanTmp, (size_t)e->cap / 1024 + 2);
    {
        unsigned tbl = 1;
        while (tbl < 2 * P)
            tbl <<= 1;
        e->sb.tableMask = tbl - 1;
        rc |= dev_alloc(e, &e->sb.flags, P);
        rc |= dev_alloc(e, &e->sb.chunkCnt, P / 1024 + 2);
        rc |= dev_alloc(e, &e->sb.pixOf, P);
        rc |= dev_alloc(e, &e->sb.pixAll, P);
        rc |= dev_alloc(e, &e->sb.keys, (size_t)tbl);
        rc |= dev_alloc(e, &e->sb.heads, (size_t)tbl);
        rc |= dev_alloc(e, &e->sb.next, P);
        rc |= dev_alloc(e, &e->spRgb, P * 3);
        rc |= dev_alloc(e, &e->spDepth, P);
        rc |= dev_alloc(e, &e->spAlpha, P);
    }
    if (!rc && cudaMallocHost((void **)&e->hostInts, 8 * sizeof(int)) != cudaSuccess)
        rc = gs_set_error(__FILE__, __LINE__, "pinned allocation failed");
    if (!rc && cudaMallocHost((void **)&e->hostLoss, 2 * sizeof(double)) != cudaSuccess)
        rc = gs_set_error(__FILE__, __LINE__, "pinned allocation failed");
    e->lossPending = false;
    if (!rc && cudaEventCreateWithFlags(&e->lossEv, cudaEventDisableTiming) != cudaSuccess)
        rc = gs_set_error(__FILE__, __LINE__, "event creation failed");
    if (rc)
    {
        gsb_gs_destroy(e);
        return 1;
    }
    cudaMemsetAsync(e->nDev, 0, sizeof(int), e->stream);
    cudaMemsetAsync(e->bins.tileCount, 0, sizeof(int) * (e->T + 1), e->stream);
    cudaMemsetAsync(e->bins.segCount, 0, sizeof(int) * (size_t)e->T * BIN_CHUNKS, e->stream);
    cudaMemsetAsync(e->bins.counters, 0, sizeof(int) * CNT_TOTAL, e->stream);
    cudaMemsetAsync(e->touched, 0, (size_t)e->cap, e->stream);
    cudaMemsetAsync(e->lossTile, 0, sizeof(float) * e->T, e->stream);
    GS_CUDA_OK(cudaStreamSynchronize(e->stream));
    *out = e;
    return 0;
}

extern "C" int gsb_gs_set_stream(gsb_gs_t *e, void *st)
{
    e->stream = st ? (cudaStream_t)st : e->ownStream;
    return 0;
}
extern "C" int gsb_gs_sync(gsb_gs_t *e)
{
    GS_CUDA_OK(cudaStreamSynchronize(e->stream));
    GS_CUDA_OK(cudaGetLastError());
    return 0;
}

// exact Gaussian count (one 4-byte D2H + stream sync); also tightens the host-side launch bound
extern "C" int gsb_gs_count(gsb_gs_t *e, int *n)
{
    GS_CUDA_OK(cudaMemcpyAsync(e->hostInts, e->nDev, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    GS_CUDA_OK(cudaStreamSynchronize(e->stream));
    if (e->comm && *e->comm->errHost)
        return gs_set_error(__FILE__, __LINE__, "multi-GPU exchange: a peer rank never reached a barrier (timed out)");
    e->nUpper = e->hostInts[0];
    if (n)
        *n = e->hostInts[0];
    return 0;
}
extern "C" int gsb_gs_count_upper(gsb_gs_t *e) { return e->nUpper; }

static int copy_rows(gsb_gs *e, float *dst, const float *src, size_t n, size_t width, size_t dstRow)
{
    if (!src || n == 0)
        return 0;
    GS_CUDA_OK(cudaMemcpyAsync(dst + dstRow * width, src, n * width * sizeof(float), cudaMemcpyDefault, e->stream));
    return 0;
}

// RawGaussianParams::add (src/raw_gs_param.cpp:123-140): append n Gaussians. Pointers may be host or device (UVA).
// rest may be NULL (zeros, as RawGaussianParams::init produces).  `at` < 0 appends at the current host-side count.
static int put_params(gsb_gs *e, int at, int n, const float *means, const float *scales_log, const float *quats, const float *dc, const float *rest,
                      const float *opac_logit)
{
    if (at + n > e->cap)
        return gs_set_error(__FILE__, __LINE__, "Gaussian capacity exceeded");
    int rc = 0;
    rc |= copy_rows(e, e->p.means, means, n, 3, at);
    rc |= copy_rows(e, e->p.scales, scales_log, n, 3, at);
    rc |= copy_rows(e, e->p.quats, quats, n, 4, at);
    rc |= copy_rows(e, e->p.dc, dc, n, 3, at);
    if (rest)
        rc |= copy_rows(e, e->p.rest, rest, n, 45, at);
    else if (n > 0)
        GS_CUDA_OK(cudaMemsetAsync(e->p.rest + (size_t)at * 45, 0, (size_t)n * 45 * sizeof(float), e->stream));
    rc |= copy_rows(e, e->p.opac, opac_logit, n, 1, at);
    return rc;
}

extern "C" int gsb_gs_set_params(gsb_gs_t *e, int n, const float *means, const float *scales_log, const float *quats, const float *dc,
                                 const float *rest, const float *opac_logit)
{
    if (n < 0 || n > e->cap)
        return gs_set_error(__FILE__, __LINE__, "Gaussian capacity exceeded");
    if (put_params(e, 0, n, means, scales_log, quats, dc, rest, opac_logit))
        return 1;
    e->hostInts[1] = n;
    GS_CUDA_OK(cudaMemcpyAsync(e->nDev, &e->hostInts[1], sizeof(int), cudaMemcpyHostToDevice, e->stream));
    // a new model has no optimiser state
    e->adamStep = 0;
    if (n > 0)
        GS_CUDA_OK(cudaMemsetAsync(e->touched, 0, (size_t)n, e->stream));
    GS_CUDA_OK(cudaStreamSynchronize(e->stream));
    e->nUpper = n;
    return 0;
}

extern "C" int gsb_gs_append(gsb_gs_t *e, int n, const float *means, const float *scales_log, const float *quats, const float *dc,
                             const float *rest, const float *opac_logit)
{
    int cur = 0;
    if (gsb_gs_count(e, &cur))
        return 1;
    if (put_params(e, cur, n, means, scales_log, quats, dc, rest, opac_logit))
        return 1;
    // new Gaussians have no optimiser state
    if (n > 0)
        GS_CUDA_OK(cudaMemsetAsync(e->touched + cur, 0, (size_t)n, e->stream));
    e->hostInts[1] = cur + n;
    GS_CUDA_OK(cudaMemcpyAsync(e->nDev, &e->hostInts[1], sizeof(int), cudaMemcpyHostToDevice, e->stream));
    GS_CUDA_OK(cudaStreamSynchronize(e->stream));
    e->nUpper = cur + n;
    return 0;
}

extern "C" int gsb_gs_get_params(gsb_gs_t *e, int n, float *means, float *scales_log, float *quats, float *dc, float *rest, float *opac_logit)
{
    if (n > e->cap)
        return gs_set_error(__FILE__, __LINE__, "read larger than the capacity");
    struct
    {
        float *dst;
        const float *src;
        size_t w;
    } a[6] = {{means, e->p.means, 3}, {scales_log, e->p.scales, 3}, {quats, e->p.quats, 4}, {dc, e->p.dc, 3}, {rest, e->p.rest, 45}, {opac_logit, e->p.opac, 1}};
    for (auto &x : a)
        if (x.dst && n > 0)
            GS_CUDA_OK(cudaMemcpyAsync(x.dst, x.src, (size_t)n * x.w * sizeof(float), cudaMemcpyDefault, e->stream));
    GS_CUDA_OK(cudaStreamSynchronize(e->stream));
    return 0;
}

// RawGaussianModel::initOptimizers (src/raw_gs_model.cpp:654-675): the 6 Adam optimisers are re-created, i.e. step = 0, m = v = 0.
// State is materialised lazily (see k_bwd_params_adam), so this only clears one byte per Gaussian.
extern "C" int gsb_gs_init_optimizers(gsb_gs_t *e)
{
    e->adamStep = 0;
    if (e->nUpper > 0)
        GS_CUDA_OK(cudaMemsetAsync(e->touched, 0, (size_t)e->nUpper, e->stream));
    return 0;
}

extern "C" int gsb_gs_set_learning_rates(gsb_gs_t *e, float means, float scales, float quats, float dc, float rest, float opac)
{
    e->cfg.lr_means = means, e->cfg.lr_scales = scales, e->cfg.lr_quats = quats, e->cfg.lr_dc = dc, e->cfg.lr_rest = rest, e->cfg.lr_opac = opac;
    return 0;
}

// c2w: row-major 4x4 camera-to-world (cam.c2w_slam). viewmat = poseInv(c2w) (src/tensor_math.cpp:56-67), fp32, same order as the oracle.
static void make_camera(const gsb_gs *e, const float *c2w, float fx, float fy, float cx, float cy, CamParams &c)
{
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            c.R[i * 3 + j] = c2w[j * 4 + i];
    const float T[3] = {c2w[3], c2w[7], c2w[11]};
    for (int i = 0; i < 3; i++)
    {
        float a = -c.R[i * 3 + 0], b = -c.R[i * 3 + 1], d = -c.R[i * 3 + 2];
        volatile float s = a * T[0];
        volatile float s2 = b * T[1];
        volatile float s3 = s + s2;
        volatile float s4 = d * T[2];
        c.t[i] = s3 + s4;
        c.cam_pos[i] = T[i];
    }
    c.fx = fx, c.fy = fy, c.cx = cx, c.cy = cy;
    c.W = e->W, c.H = e->H;
    c.eps2d = e->cfg.eps2d, c.near_plane = e->cfg.near_plane, c.far_plane = e->cfg.far_plane, c.radius_clip = e->cfg.radius_clip;
    c.max_radii = e->cfg.max_gs_radii;
    cam_limits(c);
}

static RasterIO make_io(const gsb_gs *e, const float *ref_depth, const float *base_color, const float *gt)
{
    RasterIO io;
    memset(&io, 0, sizeof io);
    io.refDepth = ref_depth, io.baseColor = base_color, io.gt = gt;
    io.deltaDepth = e->cfg.delta_depth;
    io.clampRef = 1;
    io.v_out = e->v_out, io.lossTile = e->lossTile;
    return io;
}

// torch::optim::Adam scalars for the current step; the step count is per optimiser and identical for the 6 of them
static AdamStep make_adam_step(const gsb_gs *e)
{
    const double b1 = (double)0.9f, b2 = (double)0.999f; // float betas widened to double (src/raw_gs_model.cpp:662-663)
    const double bc1 = 1.0 - pow(b1, e->adamStep), bc2 = 1.0 - pow(b2, e->adamStep);
    AdamStep s;
    s.a.beta1 = (float)b1, s.a.beta2 = (float)b2;
    s.a.one_m_beta1 = (float)(1.0 - b1), s.a.one_m_beta2 = (float)(1.0 - b2);
    s.a.sqrt_bc2 = (float)sqrt(bc2);
    s.a.eps = 1e-15f;
    const double lr[6] = {(double)e->cfg.lr_means * e->cfg.scene_scale, e->cfg.lr_scales, e->cfg.lr_quats, e->cfg.lr_dc, e->cfg.lr_rest, e->cfg.lr_opac};
    for (int i = 0; i < 6; i++)
        s.step_size[i] = (float)(lr[i] / bc1);
    return s;
}

// RawGaussianModel::forward -> gesForward (src/raw_gs_model.cpp:188-367) without autograd: rgb [H,W,3], depth [H,W], alpha [H,W]
extern "C" int gsb_gs_render(gsb_gs_t *e, const float *c2w, float fx, float fy, float cx, float cy, const float *ref_depth_dev,
                             const float *base_color_dev, float *rgb_dev, float *depth_dev, float *alpha_dev)
{
    if (!e || !c2w || !ref_depth_dev || !base_color_dev || !rgb_dev || !depth_dev || !alpha_dev)
        return gs_set_error(__FILE__, __LINE__, "null argument");
    CamParams cam;
    make_camera(e, c2w, fx, fy, cx, cy, cam);
    project_sh_fwd(e->p, e->nDev, e->nUpper, cam, e->recs, e->grads, e->bins, e->tileW, e->tileH, false, e->stream);
    bin_tiles(e->recs, e->nDev, e->nUpper, e->bins, e->tileW, e->tileH, e->stream);
    RasterIO io = make_io(e, ref_depth_dev, base_color_dev, nullptr);
    io.rgb = rgb_dev, io.depth = depth_dev, io.alphas = alpha_dev;
    if (e->comm)
    {
        // every rank needs the whole render: every tile goes to every rank, each rank composites all of it
        raster_fwd_push(e->recs, e->bins, e->W, e->H, e->tileW, e->tileH, io, e->comm->viewDev, true, e->stream);
        comm_barrier(e->comm, e->stream);
        composite_exchange(RASTER_RENDER, e->comm->view, e->comm->viewDev, e->W, e->H, e->tileW, e->tileH, io, e->stream);
        comm_barrier(e->comm, e->stream); // the gather slots may be overwritten by the peers' next push only after this
    }
    else
        raster_fwd(RASTER_RENDER, e->recs, e->bins, e->W, e->H, e->tileW, e->tileH, io, e->stream);
    GS_CUDA_OK(cudaGetLastError());
    return 0;
}

// One optimiser iteration of SLAMPipeline::localOptimize (slam/slam_pipeline.cpp:222-254): model.forward + computeLoss (L1) +
// backward + optimizersStep + optimizersZeroGrad.
extern "C" int gsb_gs_train_step(gsb_gs_t *e, const float *c2w, float fx, float fy, float cx, float cy, const float *ref_depth_dev,
                                 const float *base_color_dev, const float *gt_rgb_dev)
{
    if (!e || !c2w || !ref_depth_dev || !base_color_dev || !gt_rgb_dev)
        return gs_set_error(__FILE__, __LINE__, "null argument");
    CamParams cam;
    make_camera(e, c2w, fx, fy, cx, cy, cam);
    project_sh_fwd(e->p, e->nDev, e->nUpper, cam, e->recs, e->grads, e->bins, e->tileW, e->tileH, true, e->stream);
    bin_tiles(e->recs, e->nDev, e->nUpper, e->bins, e->tileW, e->tileH, e->stream);
    RasterIO io = make_io(e, ref_depth_dev, base_color_dev, gt_rgb_dev);
    e->lastCam = cam, e->lastIo = io, e->haveLast = true;
    if (e->comm)
    {
        // reduce-scatter by the rasteriser's own stores, owner-side composite + loss, all-gather of dL/d(render) by the composite's stores
        raster_fwd_push(e->recs, e->bins, e->W, e->H, e->tileW, e->tileH, io, e->comm->viewDev, false, e->stream);
        comm_barrier(e->comm, e->stream);
        composite_exchange(RASTER_TRAIN, e->comm->view, e->comm->viewDev, e->W, e->H, e->tileW, e->tileH, io, e->stream);
        comm_barrier(e->comm, e->stream);
    }
    else
        raster_fwd(RASTER_TRAIN, e->recs, e->bins, e->W, e->H, e->tileW, e->tileH, io, e->stream);
    if (e->nUpper > 0)
        raster_bwd(e->recs, e->bins, e->W, e->H, io, nullptr, e->grads, e->stream);
    e->adamStep++;
    AdamStep s = make_adam_step(e);
    bwd_params_adam(e->p, e->m, e->v, e->touched, s, e->nDev, e->nUpper, cam, e->recs, e->grads, e->aux, e->haveDbg ? &e->dbg : nullptr,
                    e->bins.counters, e->stream);
    GS_CUDA_OK(cudaGetLastError());
    return 0;
}

// loss of the last train step: mean |gt - rgb| (synchronises)
extern "C" int gsb_gs_loss(gsb_gs_t *e, double *loss)
{
    reduce_loss(e->lossTile, e->T, 1.0 / (3.0 * e->W * e->H), e->lossDev, e->stream);
    GS_CUDA_OK(cudaMemcpyAsync(e->hostLoss, e->lossDev, sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    GS_CUDA_OK(cudaStreamSynchronize(e->stream));
    *loss = *e->hostLoss;
    return 0;
}

// The same read-back without stalling the host: gsb_gs_loss_begin enqueues the reduction and the 8-byte copy into pinned memory and
// returns; gsb_gs_loss_end waits for that copy only (not for whatever was enqueued since) and returns the value.  A loop that reads
// the loss of cycle k while cycle k+1 is already queued keeps the GPU fed.
extern "C" int gsb_gs_loss_begin(gsb_gs_t *e)
{
    reduce_loss(e->lossTile, e->T, 1.0 / (3.0 * e->W * e->H), e->lossDev, e->stream);
    GS_CUDA_OK(cudaMemcpyAsync(e->hostLoss + 1, e->lossDev, sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    GS_CUDA_OK(cudaEventRecord(e->lossEv, e->stream));
    e->lossPending = true;
    return 0;
}
extern "C" int gsb_gs_loss_end(gsb_gs_t *e, double *loss)
{
    if (!e->lossPending)
        return gs_set_error(__FILE__, __LINE__, "gsb_gs_loss_end without gsb_gs_loss_begin");
    GS_CUDA_OK(cudaEventSynchronize(e->lossEv));
    e->lossPending = false;
    if (e->comm && *e->comm->errHost)
        return gs_set_error(__FILE__, __LINE__, "multi-GPU exchange: a peer rank never reached a barrier (timed out)");
    *loss = e->hostLoss[1];
    return 0;
}

// SLAMPipeline::removeRedundantGs (slam/slam_pipeline.cpp:564-586) -> RawGaussianModel::prunePoints (+ removeFromOptimizer for every
// optimiser, src/raw_gs_model.cpp:744-765: a surviving Gaussian keeps its own Adam moments).  No host round trip: the new count stays
// on the device, the host keeps the old count as launch bound until the next gsb_gs_count.
extern "C" int gsb_gs_prune(gsb_gs_t *e, float min_opac, float min_scale, float max_scale)
{
    if (e->nUpper <= 0)
        return 0;
    PruneBuffers b;
    b.p = e->p, b.m = e->m, b.v = e->v, b.pOut = e->tmp, b.mOut = e->mTmp, b.vOut = e->vTmp;
    b.touched = e->touched, b.touchedOut = e->touchedTmp;
    prune(b, e->nDev, e->nUpper, min_opac, min_scale, max_scale, e->scanTmp, e->bins.counters, e->stream);
    std::swap(e->p, e->tmp);
    std::swap(e->m, e->mTmp);
    std::swap(e->v, e->vTmp);
    std::swap(e->touched, e->touchedTmp);
    GS_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int gsb_gs_enable_grad_dump(gsb_gs_t *e, int on)
{
    if (on && !e->dbg.means)
        if (alloc_params(e, e->dbg, e->cap))
            return 1;
    e->haveDbg = on != 0;
    return 0;
}

extern "C" int gsb_gs_read(gsb_gs_t *e, int what, void *dst, size_t bytes)
{
    const void *src = nullptr;
    size_t avail = 0;
    const size_t P = (size_t)e->W * e->H;
    switch (what)
    {
    case GSB_GS_SPLAT_RECORDS: src = e->recs, avail = (size_t)e->cap * sizeof(SplatRec); break;
    case GSB_GS_SPLAT_GRADS: src = e->grads, avail = (size_t)e->cap * sizeof(SplatGrad); break;
    case GSB_GS_TILE_OFFSETS: src = e->bins.tileOffsets, avail = (size_t)(e->T + 1) * 4; break;
    case GSB_GS_FLATTEN_IDS: src = e->bins.flattenSorted, avail = (size_t)e->bins.isectCap * 4; break;
    case GSB_GS_V_OUT: src = e->v_out, avail = P * 32; break;
    case GSB_GS_COUNTERS: src = e->bins.counters, avail = CNT_TOTAL * 4; break;
    case GSB_GS_GRAD_MEANS: src = e->dbg.means, avail = (size_t)e->cap * 12; break;
    case GSB_GS_GRAD_SCALES: src = e->dbg.scales, avail = (size_t)e->cap * 12; break;
    case GSB_GS_GRAD_QUATS: src = e->dbg.quats, avail = (size_t)e->cap * 16; break;
    case GSB_GS_GRAD_DC: src = e->dbg.dc, avail = (size_t)e->cap * 12; break;
    case GSB_GS_GRAD_REST: src = e->dbg.rest, avail = (size_t)e->cap * 180; break;
    case GSB_GS_GRAD_OPAC: src = e->dbg.opac, avail = (size_t)e->cap * 4; break;
    case GSB_GS_SPAWN_PIXELS: src = e->sb.pixOf, avail = P * 4; break;
    default: return gs_set_error(__FILE__, __LINE__, "bad read id");
    }
    if (!src)
        return gs_set_error(__FILE__, __LINE__, "buffer not allocated (gsb_gs_enable_grad_dump)");
    if (bytes > avail)
        return gs_set_error(__FILE__, __LINE__, "read larger than the buffer");
    GS_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, e->stream));
    GS_CUDA_OK(cudaStreamSynchronize(e->stream));
    return 0;
}

// runRaycastByCam's tensor glue (slam/slam_pipeline.cpp:386-403): from GetFreeVertex()/GetFreeImage() device images to the
// depth_map [H,W] / color_map [H,W,3] / confidence_map [H,W] (nullable) the Gaussian model consumes.  c2w row-major (cam.c2w).
extern "C" int gsb_gs_raycast_maps(gsb_gs_t *e, const void *free_vertex_dev, const void *free_image_dev, const float *c2w, float voxel_size,
                                   float *depth_map_dev, float *color_map_dev, float *conf_map_dev)
{
    if (!e || !free_vertex_dev || !free_image_dev || !c2w || !depth_map_dev || !color_map_dev)
        return gs_set_error(__FILE__, __LINE__, "null argument");
    CamParams cam;
    make_camera(e, c2w, 1.f, 1.f, 0.f, 0.f, cam);
    float w2c[16] = {cam.R[0], cam.R[1], cam.R[2], cam.t[0], cam.R[3], cam.R[4], cam.R[5], cam.t[1], cam.R[6], cam.R[7], cam.R[8], cam.t[2], 0, 0, 0, 1};
    raycast_maps(e->W * e->H, (const float4 *)free_vertex_dev, (const uchar4 *)free_image_dev, w2c, voxel_size, depth_map_dev, color_map_dev,
                 conf_map_dev, e->stream);
    GS_CUDA_OK(cudaGetLastError());
    return 0;
}

// Camera::image (float rgb in [0,1]) and Camera::depth (metres) from the raw frame
extern "C" int gsb_gs_frame_to_float(gsb_gs_t *e, const void *rgba_dev, const void *depth_mm_dev, float *rgb_dev, float *depth_dev)
{
    if (!e || !rgba_dev || !rgb_dev)
        return gs_set_error(__FILE__, __LINE__, "null argument");
    frame_to_float(e->W * e->H, (const uchar4 *)rgba_dev, (const short *)depth_mm_dev, rgb_dev, depth_dev, e->stream);
    GS_CUDA_OK(cudaGetLastError());
    return 0;
}

// SLAMPipeline::initNewGaussians + SLAMGaussianModel::addGaussians (slam/slam_pipeline.cpp:450-526, slam/slam_gs_model.cpp:5-56):
// render the current camera (when Gaussians exist), build the sample mask, sample, and append new Gaussians at the raycast
// vertices.  No host round trip; call gsb_gs_count afterwards to learn the new count.
extern "C" int gsb_gs_spawn(gsb_gs_t *e, const gsb_spawn_config_t *sc, const float *c2w, float fx, float fy, float cx, float cy,
                            const void *free_vertex_dev, float voxel_size, const float *depth_map_dev, const float *color_map_dev,
                            const float *gt_rgb_dev)
{
    if (!e || !sc || !c2w || !free_vertex_dev || !depth_map_dev || !color_map_dev || !gt_rgb_dev)
        return gs_set_error(__FILE__, __LINE__, "null argument");
    if (!(sc->max_init_scale > 0.f))
        return gs_set_error(__FILE__, __LINE__, "gsb_gs_spawn needs max_init_scale > 0 (bounded-radius KNN)");
    const float *rRgb = e->spRgb, *rAlpha = e->spAlpha;
    if (sc->render_rgb_dev && sc->render_alpha_dev)
        rRgb = (const float *)sc->render_rgb_dev, rAlpha = (const float *)sc->render_alpha_dev; // the caller rendered (multi-GPU)
    else if (e->nUpper > 0 || e->comm)   // (with a communicator every rank takes part in the render, even one without Gaussians)
        if (gsb_gs_render(e, c2w, fx, fy, cx, cy, depth_map_dev, color_map_dev, e->spRgb, e->spDepth, e->spAlpha))
            return 1;
    SpawnParams sp;
    sp.W = e->W, sp.H = e->H, sp.P = e->W * e->H;
    sp.voxelSize = voxel_size;
    sp.colorErrorThres = sc->color_error_thres;
    sp.depthMin = sc->depth_vis_min, sp.depthMax = sc->depth_vis_max, sp.alphaMax = sc->alpha_vis_max;
    double r = (double)sc->sample_ratio * 4294967296.0;
    sp.ratioThreshold = r >= 4294967295.0 ? 0xffffffffu : (r <= 0.0 ? 0u : (unsigned)r);
    sp.seed = sc->seed;
    sp.maxScale = sc->max_init_scale, sp.minScale = sc->min_init_scale;
    sp.defaultOpacity = sc->default_opacity;
    sp.rank = sc->rank, sp.world = sc->world;
    sp.forceRender = (sc->render_rgb_dev && sc->render_alpha_dev) ? 1 : 0;
    if (e->comm)
        sp.rank = e->comm->rank, sp.world = e->comm->world, sp.forceRender = 1;
    spawn(sp, e->sb, (const float4 *)free_vertex_dev, depth_map_dev, color_map_dev, gt_rgb_dev, rRgb, rAlpha, e->p, e->nDev, e->cap,
          e->touched, e->bins.counters, e->stream);
    // host-side bound until the next gsb_gs_count
    long long up = (long long)e->nUpper + sp.P;
    e->nUpper = (int)(up > e->cap ? e->cap : up);
    GS_CUDA_OK(cudaGetLastError());
    return 0;
}

// Multi-GPU: from here on gsb_gs_render / gsb_gs_train_step / gsb_gs_spawn exchange the partial images through the communicator's
// peer-mapped segment (gs_comm.h); every rank must issue the same sequence of these calls.  The engine's dL/d(render) image and tile
// losses move into the segment, where the owner ranks' composite kernels store them.
extern "C" int gsb_gs_set_comm(gsb_gs_t *e, gsb_comm_t *comm)
{
    if (!e)
        return gs_set_error(__FILE__, __LINE__, "null argument");
    if (comm && (comm->W != e->W || comm->H != e->H || !comm->attached))
        return gs_set_error(__FILE__, __LINE__, "communicator not attached, or created for another image size");
    if (comm && comm->world == 1)
        comm = nullptr;
    GS_CUDA_OK(cudaStreamSynchronize(e->stream));
    e->comm = comm;
    e->v_out = comm ? comm->view.vout[comm->rank] : e->vOutOwn;
    e->lossTile = comm ? comm->view.lossTile[comm->rank] : e->lossTileOwn;
    return 0;
}

// Single stages on the camera / images of the last gsb_gs_train_step, for per-kernel timing (bench.py roofline, ncu).
// stage: 0 projection+SH (for backward), 1 tile binning, 2 rasteriser forward (train mode), 3 rasteriser backward,
//        4 drop the backward work list (leaves the engine ready for the next train step), 5 parameter backward + Adam with a
//        zero step size.  Stage 1 re-runs stage 0 first.  Run 1,2 before 3 and 5; finish with 4.
extern "C" int gsb_gs_run_stage(gsb_gs_t *e, int stage)
{
    if (!e->haveLast)
        return gs_set_error(__FILE__, __LINE__, "gsb_gs_run_stage needs a previous gsb_gs_train_step");
    switch (stage)
    {
    case 0:
        GS_CUDA_OK(cudaMemsetAsync(e->bins.counters + CNT_ITEMS, 0, sizeof(int), e->stream));
        GS_CUDA_OK(cudaMemsetAsync(e->bins.segCount, 0, sizeof(int) * (size_t)e->T * BIN_CHUNKS, e->stream));
        project_sh_fwd(e->p, e->nDev, e->nUpper, e->lastCam, e->recs, e->grads, e->bins, e->tileW, e->tileH, true, e->stream);
        break;
    case 1: // binning consumes the tile counts, so it is always timed together with the projection that produces them
        GS_CUDA_OK(cudaMemsetAsync(e->bins.counters + CNT_ITEMS, 0, sizeof(int), e->stream));
        GS_CUDA_OK(cudaMemsetAsync(e->bins.segCount, 0, sizeof(int) * (size_t)e->T * BIN_CHUNKS, e->stream));
        project_sh_fwd(e->p, e->nDev, e->nUpper, e->lastCam, e->recs, e->grads, e->bins, e->tileW, e->tileH, true, e->stream);
        bin_tiles(e->recs, e->nDev, e->nUpper, e->bins, e->tileW, e->tileH, e->stream);
        break;
    case 2: raster_fwd(RASTER_TRAIN, e->recs, e->bins, e->W, e->H, e->tileW, e->tileH, e->lastIo, e->stream); break;
    case 3:
        GS_CUDA_OK(cudaMemsetAsync(e->bins.counters + CNT_BWD_CURSOR, 0, sizeof(int), e->stream));
        raster_bwd(e->recs, e->bins, e->W, e->H, e->lastIo, nullptr, e->grads, e->stream);
        break;
    case 6: // rasteriser backward with pair statistics (GSB_GS_COUNTERS ints 8-11: pairs tested / passed as two 64-bit values)
        GS_CUDA_OK(cudaMemsetAsync(e->bins.counters + CNT_BWD_CURSOR, 0, sizeof(int), e->stream));
        raster_bwd_stats(e->recs, e->bins, e->W, e->H, e->lastIo, e->grads, e->stream);
        break;
    case 4: GS_CUDA_OK(cudaMemsetAsync(e->bins.counters + CNT_ITEMS, 0, sizeof(int), e->stream)); break;
    case 5: // parameter backward + Adam with a zero step size (moments advance, parameters do not move)
    {
        AdamStep s;
        s.a.beta1 = 0.9f, s.a.beta2 = 0.999f, s.a.one_m_beta1 = 0.1f, s.a.one_m_beta2 = 0.001f, s.a.sqrt_bc2 = 1.0f, s.a.eps = 1e-15f;
        for (int i = 0; i < 6; i++)
            s.step_size[i] = 0.f;
        bwd_params_adam(e->p, e->m, e->v, e->touched, s, e->nDev, e->nUpper, e->lastCam, e->recs, e->grads, e->aux, nullptr, e->bins.counters,
                        e->stream);
        break;
    }
    default: return gs_set_error(__FILE__, __LINE__, "bad stage id");
    }
    GS_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---- staged entry points: one per gsplat::*_tensor function of the GES path (C = 1), device pointers in and out, caller-owned outputs.
// viewmat: row-major 4x4 world-to-camera, K: row-major 3x3 (the reference passes [1,4,4] / [1,3,3] tensors).
static void make_camera_viewmat(const gsb_gs *e, const float *viewmat, const float *K, int clamp_radii, CamParams &c)
{
    for (int i = 0; i < 3; i++)
    {
        for (int j = 0; j < 3; j++)
            c.R[i * 3 + j] = viewmat[i * 4 + j];
        c.t[i] = viewmat[i * 4 + 3];
        c.cam_pos[i] = 0.f; // view directions are an input of the staged SH functions
    }
    c.fx = K[0], c.fy = K[4], c.cx = K[2], c.cy = K[5];
    c.W = e->W, c.H = e->H;
    c.eps2d = e->cfg.eps2d, c.near_plane = e->cfg.near_plane, c.far_plane = e->cfg.far_plane, c.radius_clip = e->cfg.radius_clip;
    c.max_radii = clamp_radii;
    cam_limits(c);
}

static int check_n(const gsb_gs *e, int n)
{
    if (!e)
        return gs_set_error(__FILE__, __LINE__, "null engine");
    if (n < 0 || n > e->cap)
        return gs_set_error(__FILE__, __LINE__, "Gaussian count exceeds the engine capacity");
    return 0;
}

extern "C" int gsb_gs_projection_fwd(gsb_gs_t *e, int n, const float *means, const float *quats, const float *scales, const float *viewmat,
                                     const float *K, int clamp_radii, int *radii, float *means2d, float *depths, float *conics)
{
    if (check_n(e, n))
        return 1;
    if (!means || !quats || !scales || !viewmat || !K || !radii || !means2d || !depths || !conics)
        return gs_set_error(__FILE__, __LINE__, "null argument");
    CamParams cam;
    make_camera_viewmat(e, viewmat, K, clamp_radii, cam);
    staged_project_fwd(n, means, quats, scales, cam, radii, means2d, depths, conics, e->stream);
    GS_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int gsb_gs_projection_bwd(gsb_gs_t *e, int n, const float *means, const float *quats, const float *scales, const float *viewmat,
                                     const float *K, const int *radii, const float *conics, const float *v_means2d, const float *v_depths,
                                     const float *v_conics, float *v_means, float *v_quats, float *v_scales)
{
    if (check_n(e, n))
        return 1;
    if (!means || !quats || !scales || !viewmat || !K || !radii || !conics || !v_means2d || !v_depths || !v_conics || !v_means || !v_quats ||
        !v_scales)
        return gs_set_error(__FILE__, __LINE__, "null argument");
    CamParams cam;
    make_camera_viewmat(e, viewmat, K, 0, cam);
    staged_project_bwd(n, means, quats, scales, cam, radii, conics, v_means2d, v_depths, v_conics, v_means, v_quats, v_scales, e->stream);
    GS_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int gsb_gs_sh_fwd(gsb_gs_t *e, int n, int degrees_to_use, const float *dirs, const float *coeffs, const unsigned char *masks,
                             float *colors)
{
    if (check_n(e, n))
        return 1;
    if (degrees_to_use != 3)
        return gs_set_error(__FILE__, __LINE__, "only SH degree 3 (16 bases) is built: the SLAM path never uses another (raw_gs_model.cpp:256)");
    if (!dirs || !coeffs || !colors)
        return gs_set_error(__FILE__, __LINE__, "null argument");
    staged_sh_fwd(n, dirs, coeffs, masks, colors, e->stream);
    GS_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int gsb_gs_sh_bwd(gsb_gs_t *e, int n, int degrees_to_use, const float *dirs, const float *coeffs, const unsigned char *masks,
                             const float *v_colors, float *v_coeffs, float *v_dirs)
{
    if (check_n(e, n))
        return 1;
    if (degrees_to_use != 3)
        return gs_set_error(__FILE__, __LINE__, "only SH degree 3 (16 bases) is built");
    if (!dirs || !coeffs || !v_colors || !v_coeffs)
        return gs_set_error(__FILE__, __LINE__, "null argument");
    staged_sh_bwd(n, dirs, coeffs, masks, v_colors, v_coeffs, v_dirs, e->stream);
    GS_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int gsb_gs_isect_tiles(gsb_gs_t *e, int n, const float *means2d, const int *radii, int *tiles_per_gauss, int *n_isects)
{
    if (check_n(e, n))
        return 1;
    if (!means2d || !radii || !n_isects)
        return gs_set_error(__FILE__, __LINE__, "null argument");
    GS_CUDA_OK(cudaMemsetAsync(e->bins.segCount, 0, sizeof(int) * (size_t)e->T * BIN_CHUNKS, e->stream));
    staged_pack(n, means2d, nullptr, nullptr, nullptr, nullptr, radii, e->recs, e->grads, e->bins, e->tileW, e->tileH, e->W, e->H, tiles_per_gauss, true,
                false, e->stream);
    bin_tiles(e->recs, nullptr, n, e->bins, e->tileW, e->tileH, e->stream);
    GS_CUDA_OK(cudaMemcpyAsync(e->hostInts, e->bins.counters, 8 * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    GS_CUDA_OK(cudaStreamSynchronize(e->stream)); // the reference synchronises here too (n_isects sizes the output tensors)
    if (e->hostInts[CNT_OVERFLOW] & 1)
        return gs_set_error(__FILE__, __LINE__, "intersection capacity exceeded (gsb_gs_config.isect_capacity)");
    *n_isects = e->hostInts[CNT_ISECTS];
    return 0;
}

extern "C" int gsb_gs_isect_fetch(gsb_gs_t *e, int n_isects, long long *isect_ids, int *flatten_ids, int *tile_offsets)
{
    if (!e)
        return gs_set_error(__FILE__, __LINE__, "null engine");
    if (n_isects < 0 || n_isects > e->bins.isectCap)
        return gs_set_error(__FILE__, __LINE__, "invalid n_isects");
    if (isect_ids && n_isects)
        staged_isect_ids(e->bins, e->T, isect_ids, e->stream);
    if (flatten_ids && n_isects)
        GS_CUDA_OK(cudaMemcpyAsync(flatten_ids, e->bins.flattenSorted, sizeof(int) * (size_t)n_isects, cudaMemcpyDeviceToDevice, e->stream));
    if (tile_offsets)
        GS_CUDA_OK(cudaMemcpyAsync(tile_offsets, e->bins.tileOffsets, sizeof(int) * (size_t)e->T, cudaMemcpyDeviceToDevice, e->stream));
    GS_CUDA_OK(cudaGetLastError());
    return 0;
}

// bins supplied by the caller (tile_offsets [T], flatten_ids [n_isects]) -> engine-side Bins view
static int staged_bins(gsb_gs *e, const int *tile_offsets, const int *flatten_ids, int n_isects, Bins &b)
{
    b = e->bins;
    GS_CUDA_OK(cudaMemcpyAsync(e->bins.tileOffsets, tile_offsets, sizeof(int) * (size_t)e->T, cudaMemcpyDeviceToDevice, e->stream));
    e->hostInts[0] = n_isects;
    GS_CUDA_OK(cudaMemcpyAsync(e->bins.tileOffsets + e->T, e->hostInts, sizeof(int), cudaMemcpyHostToDevice, e->stream));
    GS_CUDA_OK(cudaStreamSynchronize(e->stream)); // hostInts is reused by later calls
    b.flattenSorted = const_cast<int *>(flatten_ids);
    return 0;
}

extern "C" int gsb_gs_rasterize_ges_fwd(gsb_gs_t *e, int n, const float *means2d, const float *conics, const float *colors4,
                                        const float *opacities, const float *ref_depth, float delta_depth, const int *tile_offsets,
                                        const int *flatten_ids, int n_isects, float *render4, float *alphas)
{
    if (check_n(e, n))
        return 1;
    if (!means2d || !conics || !colors4 || !opacities || !ref_depth || !tile_offsets || (!flatten_ids && n_isects) || !render4 || !alphas)
        return gs_set_error(__FILE__, __LINE__, "null argument");
    // radii are not an input of the reference's forward: any positive value marks the splat live (bins come from the caller)
    Bins b;
    if (staged_bins(e, tile_offsets, flatten_ids, n_isects, b))
        return 1;
    staged_pack(n, means2d, conics, colors4, nullptr, opacities, nullptr, e->recs, e->grads, e->bins, e->tileW, e->tileH, e->W, e->H, nullptr, false, false,
                e->stream);
    RasterIO io = make_io(e, ref_depth, nullptr, nullptr);
    io.deltaDepth = delta_depth;
    io.clampRef = 0; // the caller passes ref_depth_clamped (src/raw_gs_model.cpp:207)
    io.render4 = render4, io.alphas = alphas;
    raster_fwd(RASTER_RAW, e->recs, b, e->W, e->H, e->tileW, e->tileH, io, e->stream);
    GS_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int gsb_gs_rasterize_ges_bwd(gsb_gs_t *e, int n, const float *means2d, const float *conics, const float *colors4,
                                        const float *opacities, const int *radii, const float *ref_depth, float delta_depth,
                                        const float *v_render4, const float *v_alphas, float *v_means2d, float *v_conics, float *v_colors4,
                                        float *v_opacities)
{
    if (check_n(e, n))
        return 1;
    if (!means2d || !conics || !colors4 || !opacities || !radii || !ref_depth || !v_render4 || !v_alphas || !v_means2d || !v_conics ||
        !v_colors4 || !v_opacities)
        return gs_set_error(__FILE__, __LINE__, "null argument");
    const int P = e->W * e->H;
    GS_CUDA_OK(cudaMemsetAsync(e->bins.counters + CNT_ITEMS, 0, sizeof(int), e->stream));
    GS_CUDA_OK(cudaMemsetAsync(e->bins.counters + CNT_BWD_CURSOR, 0, sizeof(int), e->stream));
    staged_pack(n, means2d, conics, colors4, nullptr, opacities, radii, e->recs, e->grads, e->bins, e->tileW, e->tileH, e->W, e->H, nullptr, false, true,
                e->stream);
    pack_v_out(P, v_render4, v_alphas, e->v_out, e->v_depth, e->stream);
    staged_cut(P, ref_depth, delta_depth, e->v_out, e->stream);
    RasterIO io = make_io(e, ref_depth, nullptr, nullptr);
    io.deltaDepth = delta_depth, io.clampRef = 0;
    raster_bwd(e->recs, e->bins, e->W, e->H, io, e->v_depth, e->grads, e->stream);
    staged_unpack_grads(n, e->recs, e->grads, v_means2d, v_conics, v_colors4, v_opacities, e->stream);
    GS_CUDA_OK(cudaMemsetAsync(e->bins.counters + CNT_ITEMS, 0, sizeof(int), e->stream));
    GS_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---- render_method "raw": depth-sorted bins, front-to-back compositing, and the fused SSIM map
extern "C" int gsb_gs_isect_tiles_depth(gsb_gs_t *e, int n, const float *means2d, const int *radii, const float *depths, int *tiles_per_gauss,
                                        int *n_isects)
{
    if (check_n(e, n))
        return 1;
    if (!means2d || !radii || !depths || !n_isects)
        return gs_set_error(__FILE__, __LINE__, "null argument");
    GS_CUDA_OK(cudaMemsetAsync(e->bins.segCount, 0, sizeof(int) * (size_t)e->T * BIN_CHUNKS, e->stream));
    staged_pack(n, means2d, nullptr, nullptr, depths, nullptr, radii, e->recs, e->grads, e->bins, e->tileW, e->tileH, e->W, e->H, tiles_per_gauss,
                true, false, e->stream);
    bin_tiles(e->recs, nullptr, n, e->bins, e->tileW, e->tileH, e->stream);
    sort_tiles_depth(e->recs, e->bins, e->T, e->stream);
    GS_CUDA_OK(cudaMemcpyAsync(e->hostInts, e->bins.counters, 8 * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    GS_CUDA_OK(cudaStreamSynchronize(e->stream));
    if (e->hostInts[CNT_OVERFLOW] & 1)
        return gs_set_error(__FILE__, __LINE__, "intersection capacity exceeded (gsb_gs_config.isect_capacity)");
    *n_isects = e->hostInts[CNT_ISECTS];
    return 0;
}

extern "C" int gsb_gs_isect_fetch_depth(gsb_gs_t *e, int n_isects, long long *isect_ids, int *flatten_ids, int *tile_offsets)
{
    if (!e)
        return gs_set_error(__FILE__, __LINE__, "null engine");
    if (n_isects < 0 || n_isects > e->bins.isectCap)
        return gs_set_error(__FILE__, __LINE__, "invalid n_isects");
    if (isect_ids && n_isects)
        isect_ids_depth(e->recs, e->bins, e->T, isect_ids, e->stream);
    if (flatten_ids && n_isects)
        GS_CUDA_OK(cudaMemcpyAsync(flatten_ids, e->bins.flattenSorted, sizeof(int) * (size_t)n_isects, cudaMemcpyDeviceToDevice, e->stream));
    if (tile_offsets)
        GS_CUDA_OK(cudaMemcpyAsync(tile_offsets, e->bins.tileOffsets, sizeof(int) * (size_t)e->T, cudaMemcpyDeviceToDevice, e->stream));
    GS_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int gsb_gs_rasterize_fwd(gsb_gs_t *e, int n, const float *means2d, const float *conics, const float *colors4, const float *opacities,
                                    const float *background4, const int *tile_offsets, const int *flatten_ids, int n_isects, float *render4,
                                    float *alphas, int *last_ids)
{
    if (check_n(e, n))
        return 1;
    if (!means2d || !conics || !colors4 || !opacities || !tile_offsets || (!flatten_ids && n_isects) || !render4 || !alphas || !last_ids)
        return gs_set_error(__FILE__, __LINE__, "null argument");
    Bins b;
    if (staged_bins(e, tile_offsets, flatten_ids, n_isects, b))
        return 1;
    staged_pack(n, means2d, conics, colors4, nullptr, opacities, nullptr, e->recs, e->grads, e->bins, e->tileW, e->tileH, e->W, e->H, nullptr, false,
                false, e->stream);
    raw_fwd(e->recs, b, e->W, e->H, e->tileW, e->tileH, background4, render4, alphas, last_ids, e->stream);
    GS_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int gsb_gs_rasterize_bwd(gsb_gs_t *e, int n, const float *means2d, const float *conics, const float *colors4, const float *opacities,
                                    const float *background4, const int *tile_offsets, const int *flatten_ids, int n_isects,
                                    const float *render_alphas, const int *last_ids, const float *v_render4, const float *v_alphas,
                                    float *v_means2d, float *v_conics, float *v_colors4, float *v_opacities)
{
    if (check_n(e, n))
        return 1;
    if (!means2d || !conics || !colors4 || !opacities || !tile_offsets || (!flatten_ids && n_isects) || !render_alphas || !last_ids ||
        !v_render4 || !v_alphas || !v_means2d || !v_conics || !v_colors4 || !v_opacities)
        return gs_set_error(__FILE__, __LINE__, "null argument");
    Bins b;
    if (staged_bins(e, tile_offsets, flatten_ids, n_isects, b))
        return 1;
    staged_pack(n, means2d, conics, colors4, nullptr, opacities, nullptr, e->recs, e->grads, e->bins, e->tileW, e->tileH, e->W, e->H, nullptr, false,
                false, e->stream);
    raw_bwd(n, e->recs, b, e->W, e->H, e->tileW, e->tileH, background4, render_alphas, last_ids, v_render4, v_alphas, e->grads, e->stream);
    staged_unpack_grads(n, e->recs, e->grads, v_means2d, v_conics, v_colors4, v_opacities, e->stream);
    GS_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int gsb_gs_ssim_fwd(gsb_gs_t *e, int planes, int height, int width, float C1, float C2, const float *img1, const float *img2,
                               float *ssim_map, float *dm_dmu1, float *dm_dsigma1_sq, float *dm_dsigma12)
{
    if (!e || !img1 || !img2 || !ssim_map || planes <= 0 || height <= 0 || width <= 0)
        return gs_set_error(__FILE__, __LINE__, "invalid argument");
    if ((dm_dmu1 != nullptr) != (dm_dsigma1_sq != nullptr) || (dm_dmu1 != nullptr) != (dm_dsigma12 != nullptr))
        return gs_set_error(__FILE__, __LINE__, "the three derivative maps go together (train = true) or are all NULL");
    ssim_fwd(planes, height, width, C1, C2, img1, img2, ssim_map, dm_dmu1, dm_dsigma1_sq, dm_dsigma12, e->stream);
    GS_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int gsb_gs_ssim_bwd(gsb_gs_t *e, int planes, int height, int width, const float *img1, const float *img2, const float *dL_dmap,
                               const float *dm_dmu1, const float *dm_dsigma1_sq, const float *dm_dsigma12, float *dL_dimg1)
{
    if (!e || !img1 || !img2 || !dL_dmap || !dm_dmu1 || !dm_dsigma1_sq || !dm_dsigma12 || !dL_dimg1 || planes <= 0 || height <= 0 || width <= 0)
        return gs_set_error(__FILE__, __LINE__, "invalid argument");
    ssim_bwd(planes, height, width, img1, img2, dL_dmap, dm_dmu1, dm_dsigma1_sq, dm_dsigma12, dL_dimg1, e->stream);
    GS_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int gsb_gs_adam_step(gsb_gs_t *e, long long n, float *param, const float *grad, float *exp_avg, float *exp_avg_sq, float lr,
                                float beta1, float beta2, float eps, int step)
{
    if (!e || !param || !grad || !exp_avg || !exp_avg_sq)
        return gs_set_error(__FILE__, __LINE__, "null argument");
    if (n < 0 || n > 0x7fffffffLL || step < 1)
        return gs_set_error(__FILE__, __LINE__, "invalid size or step");
    const double b1 = (double)beta1, b2 = (double)beta2;
    const double bc1 = 1.0 - pow(b1, step), bc2 = 1.0 - pow(b2, step);
    AdamScalars a;
    a.beta1 = beta1, a.beta2 = beta2, a.one_m_beta1 = (float)(1.0 - b1), a.one_m_beta2 = (float)(1.0 - b2);
    a.sqrt_bc2 = (float)sqrt(bc2), a.eps = eps;
    staged_adam((int)n, param, grad, exp_avg, exp_avg_sq, a, (float)((double)lr / bc1), e->stream);
    GS_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int gsb_gs_dist_cuda2(gsb_gs_t *e, int n, const float *points_dev, float *mean_dist2_dev)
{
    if (!e || !points_dev || !mean_dist2_dev || n <= 0)
        return gs_set_error(__FILE__, __LINE__, "invalid argument");
    if (!e->knnWs || n > e->knnCap)
    {
        // first use (or a larger point set than ever before): the one allocation this entry point makes; the old block, if any,
        // stays on the engine's list and is released by gsb_gs_destroy
        const int cap = n > e->cap ? n : e->cap;
        GS_CUDA_OK(cudaStreamSynchronize(e->stream));
        char *ws = nullptr;
        if (dev_alloc(e, &ws, knn_workspace_bytes(cap)))
            return 1;
        e->knnWs = ws, e->knnCap = cap;
    }
    knn_mean_dist3(n, points_dev, mean_dist2_dev, e->knnWs, e->stream);
    GS_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---- multi-GPU: the Gaussian set is sharded across ranks.  The GES blend is an order-independent sum, so each
// rank rasterises its own Gaussians over the whole image into partial sums acc5 = render_colors [H*W*4] + alphas [H*W]; the
// caller all-reduces acc5 (NCCL) and every rank finishes redundantly on the summed image.  Backward and Adam are rank-local.
extern "C" int gsb_gs_forward_partial(gsb_gs_t *e, const float *c2w, float fx, float fy, float cx, float cy, const float *ref_depth_dev,
                                      float *acc5_dev, int for_backward)
{
    if (!e || !c2w || !ref_depth_dev || !acc5_dev)
        return gs_set_error(__FILE__, __LINE__, "null argument");
    CamParams cam;
    make_camera(e, c2w, fx, fy, cx, cy, cam);
    project_sh_fwd(e->p, e->nDev, e->nUpper, cam, e->recs, e->grads, e->bins, e->tileW, e->tileH, for_backward != 0, e->stream);
    bin_tiles(e->recs, e->nDev, e->nUpper, e->bins, e->tileW, e->tileH, e->stream);
    RasterIO io = make_io(e, ref_depth_dev, nullptr, nullptr);
    io.render4 = acc5_dev, io.alphas = acc5_dev + (size_t)4 * e->W * e->H;
    e->lastCam = cam;
    raster_fwd(RASTER_RAW, e->recs, e->bins, e->W, e->H, e->tileW, e->tileH, io, e->stream);
    GS_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int gsb_gs_render_finish(gsb_gs_t *e, const float *ref_depth_dev, const float *base_color_dev, const float *acc5_dev, float *rgb_dev,
                                    float *depth_dev, float *alpha_dev)
{
    if (!e || !ref_depth_dev || !base_color_dev || !acc5_dev || !rgb_dev || !depth_dev || !alpha_dev)
        return gs_set_error(__FILE__, __LINE__, "null argument");
    RasterIO io = make_io(e, ref_depth_dev, base_color_dev, nullptr);
    io.rgb = rgb_dev, io.depth = depth_dev, io.alphas = alpha_dev;
    composite(RASTER_RENDER, acc5_dev, e->W, e->H, e->tileW, e->tileH, io, e->stream);
    GS_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int gsb_gs_train_finish(gsb_gs_t *e, const float *ref_depth_dev, const float *base_color_dev, const float *gt_rgb_dev,
                                   const float *acc5_dev)
{
    if (!e || !ref_depth_dev || !base_color_dev || !gt_rgb_dev || !acc5_dev)
        return gs_set_error(__FILE__, __LINE__, "null argument");
    RasterIO io = make_io(e, ref_depth_dev, base_color_dev, gt_rgb_dev);
    e->lastIo = io, e->haveLast = true;
    composite(RASTER_TRAIN, acc5_dev, e->W, e->H, e->tileW, e->tileH, io, e->stream);
    if (e->nUpper > 0)
        raster_bwd(e->recs, e->bins, e->W, e->H, io, nullptr, e->grads, e->stream);
    e->adamStep++;
    AdamStep s = make_adam_step(e);
    bwd_params_adam(e->p, e->m, e->v, e->touched, s, e->nDev, e->nUpper, e->lastCam, e->recs, e->grads, e->aux, e->haveDbg ? &e->dbg : nullptr,
                    e->bins.counters, e->stream);
    GS_CUDA_OK(cudaGetLastError());
    return 0;
}
