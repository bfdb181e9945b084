// TSDF engine object + its C ABI (include/gpsslam_b200.h, section B).
// Host logic mirrors ITMBasicEngine::ProcessFrame / runRaycast (reference InfiniTAM/ITMLib/Core/ITMBasicEngine.tpp:260-385,
// 500-526) and ITMTrackingController::Prepare (Core/ITMTrackingController.h:66-102): same stage order, same pose handling
// (SetInvM + Coerce through se3::Pose), but every stage is an asynchronous launch on one stream with no host round trip.
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/gpsslam_b200.h"
#include <stdlib.h>

#include "common.cuh"
#include "icp.h"
#include "se3.h"
#include "tsdf.h"

static thread_local std::string g_err;
int gs_set_error(const char *file, int line, const char *msg)
{
    char buf[512];
    snprintf(buf, sizeof buf, "%s:%d: %s", file, line, msg);
    g_err = buf;
    return 1;
}
extern "C" const char *gsb_last_error(void) { return g_err.c_str(); }
long long g_gsb_launches = 0;
extern "C" long long gsb_launch_count(void) { return g_gsb_launches; }
extern "C" const char *gsb_version(void) { return "gpsslam_b200 0.1 sm_100a"; }

struct gsb_tsdf
{
    gsb_tsdf_config_t cfg;
    tsdf::Scene scene;
    tsdf::Frame frame;
    tsdf::Camera cam;      // live camera (pose_d + intrinsics_d)
    se3::Pose pose_d, pose_pointCloud;
    cudaStream_t stream, ownStream;
    // device buffers
    short *depth_mm;
    uchar4 *rgba;
    float *depth_f;
    float2 *minmaxLive, *minmaxFree;
    float4 *rayLive, *rayFree, *pointsMap, *normalsMap;
    uchar4 *imageFree;
    icp::Tracker *tracker;
    int framesProcessed;       // ITMBasicEngine::framesProcessed (fused frames)
    int trackingFrames;        // ITMTrackingState::framesProcessed
    bool trackingActive;       // ITMBasicEngine::trackingActive (turnOnTracking / turnOffTracking)
    int agePointCloud;         // ITMTrackingState::age_pointCloud (-1 = no valid point cloud yet)
    bool haveFrame;
    const void *lastRgba;      // RGBA frame consumed by the last ProcessFrame (own upload buffer or the caller's resident frame)
    bool stageTiming;          // gsb_tsdf_enable_stage_timing: CUDA events between the stages of ProcessFrame
    cudaEvent_t stageEv[7];
    std::vector<void *> allocs;
    // ---- voxel-hash sharding (world > 1): everything another rank reads or writes lives in ONE cudaMalloc segment, so that one CUDA IPC
    // handle per rank maps it all: [flags | ICP exchange | visType | rayLive | rayFree | imageFree | pointsMap | normalsMap | voxel blocks]
    int rank, world;
    char *seg;
    size_t segBytes, offXchg, offVis, offRayLive, offRayFree, offImage, offPoints, offNormals, offVba;
    char *peer[tsdf::SHARD_MAX_WORLD];
    bool ipcOpened[tsdf::SHARD_MAX_WORLD];
    bool attached;
    unsigned epoch;      // barriers issued so far (identical on every rank: SPMD call sequence)
    int *errHost;        // pinned, mapped: set by a barrier (or an in-kernel exchange) that timed out
    tsdf::ShardView view;
    int shardMode;       // 0: a block's voxels live on its owner only, raycasts read them over NVLink; 1: the owner integrates and stores the
                         // block into every rank, raycasts read local memory
};

#define E_CUDA(call) GS_CUDA_OK(call)

template <typename T>
static int dev_alloc(gsb_tsdf *e, T **p, size_t n)
{
    E_CUDA(cudaMalloc((void **)p, n * sizeof(T)));
    e->allocs.push_back((void *)*p);
    return 0;
}

extern "C" void gsb_tsdf_default_config(gsb_tsdf_config_t *c)
{
    memset(c, 0, sizeof *c);
    c->width = 1200, c->height = 680;
    c->fx = 600.f, c->fy = 600.f, c->cx = 599.5f, c->cy = 339.5f; // configs/release/replica/office0.yaml:18-20
    c->voxel_size = 0.005f, c->mu = 0.02f, c->view_frustum_min = 0.2f, c->view_frustum_max = 10.0f;
    c->max_w = 100;
    c->num_blocks = SDF_DEFAULT_BLOCK_NUM;
    c->tracker = 0;
    c->device = 0;
    c->integrate_variant = 0;
}

static size_t seg_align(size_t x) { return (x + 4095) / 4096 * 4096; }

static void fill_shard_view(gsb_tsdf *e)
{
    tsdf::ShardView &v = e->view;
    v.rank = e->rank, v.world = e->world;
    v.replicated = e->shardMode == 1;
    v.probe = 0;
    e->scene.nPush = 0;
    for (int q = 0; q < e->world; q++)
        if (e->shardMode == 1 && q != e->rank)
            e->scene.pushVba[e->scene.nPush++] = (Voxel *)(e->peer[q] + e->offVba);
    for (int q = 0; q < e->world; q++)
    {
        char *b = e->peer[q];
        v.flags[q] = (unsigned *)b;
        v.icpXchg[q] = (float *)(b + e->offXchg);
        v.visType[q] = (unsigned char *)(b + e->offVis);
        v.rayLive[q] = (float4 *)(b + e->offRayLive);
        v.rayFree[q] = (float4 *)(b + e->offRayFree);
        v.imageFree[q] = (uchar4 *)(b + e->offImage);
        v.pointsMap[q] = (float4 *)(b + e->offPoints);
        v.normalsMap[q] = (float4 *)(b + e->offNormals);
        v.vba[q] = (const Voxel *)(b + e->offVba);
    }
}

static void shard_barrier(gsb_tsdf *e)
{
    e->epoch++;
    tsdf::shard_barrier(e->view, e->epoch, e->errHost, e->stream);
}

extern "C" int gsb_tsdf_create(const gsb_tsdf_config_t *cfg, gsb_tsdf_t **out) { return gsb_tsdf_create_sharded(cfg, 0, 1, out); }

extern "C" int gsb_tsdf_create_sharded(const gsb_tsdf_config_t *cfg, int rank, int world, gsb_tsdf_t **out)
{
    if (!cfg || !out)
        return gs_set_error(__FILE__, __LINE__, "null argument");
    if (world < 1 || world > tsdf::SHARD_MAX_WORLD || rank < 0 || rank >= world)
        return gs_set_error(__FILE__, __LINE__, "invalid rank / world (world <= 16)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return gs_set_error(__FILE__, __LINE__, "no CUDA device: gpsslam_b200 has no CPU fallback");
    if (cfg->width <= 0 || cfg->height <= 0 || cfg->voxel_size <= 0 || cfg->mu <= 0)
        return gs_set_error(__FILE__, __LINE__, "invalid configuration");
    E_CUDA(cudaSetDevice(cfg->device));
    gsb_tsdf *e = new (std::nothrow) gsb_tsdf();
    if (!e)
        return gs_set_error(__FILE__, __LINE__, "out of host memory");
    e->cfg = *cfg;
    if (e->cfg.integrate_variant != 0)
    {
        delete e;
        return gs_set_error(__FILE__, __LINE__, "integrate_variant must be 0 (the TMA-pipelined kernel; the alternates were removed)");
    }
    if (e->cfg.num_blocks <= 0)
        e->cfg.num_blocks = SDF_DEFAULT_BLOCK_NUM;
    if (e->cfg.max_w <= 0)
        e->cfg.max_w = 100;
    if (e->cfg.max_w > 255)
    {
        delete e;
        return gs_set_error(__FILE__, __LINE__, "max_w > 255: the voxel's depth weight is one byte (ITMVoxel_s_rgb::w_depth)");
    }
    if (world > 1 && e->cfg.num_blocks > (1 << 19))
    {
        delete e;
        return gs_set_error(__FILE__, __LINE__, "sharded scene: num_blocks <= 524288 (28-bit voxel handles)");
    }
    const int W = cfg->width, H = cfg->height, P = W * H;
    e->rank = rank, e->world = world;
    // default: mode 1 (owner integrates, every rank stores; local raycast reads) -- measured on 2 and 8 B200s it is the faster of the two
    // (profiles/r02_shard_modes.md); mode 0 keeps 1 / world of the voxel memory per GPU
    e->shardMode = 1;
    if (const char *v = getenv("GSB_TSDF_SHARD_MODE"))
        e->shardMode = atoi(v) == 0 ? 0 : 1;
    e->scene.nPush = 0;
    e->seg = nullptr, e->errHost = nullptr, e->epoch = 0, e->attached = world == 1;
    memset(e->peer, 0, sizeof e->peer);
    memset(e->ipcOpened, 0, sizeof e->ipcOpened);
    tsdf::Scene &s = e->scene;
    s.rank = rank, s.world = world;
    s.E = SDF_TOTAL_ENTRIES;
    s.numBlocks = e->cfg.num_blocks;
    s.voxelSize = cfg->voxel_size, s.mu = cfg->mu, s.vfmin = cfg->view_frustum_min, s.vfmax = cfg->view_frustum_max;
    s.maxW = e->cfg.max_w;
    E_CUDA(cudaStreamCreateWithFlags(&e->ownStream, cudaStreamNonBlocking));
    e->stream = e->ownStream;
    int rc = 0;
    rc |= dev_alloc(e, &s.table, (size_t)s.E);
    rc |= dev_alloc(e, &s.allocKey, (size_t)s.E);
    rc |= dev_alloc(e, &s.visIds, (size_t)s.numBlocks);
    if (world > 1)
    {
        e->offXchg = 4096;
        e->offVis = e->offXchg + seg_align(2 * tsdf::SHARD_MAX_WORLD * 32 * sizeof(float));
        e->offRayLive = e->offVis + seg_align((size_t)s.E);
        e->offRayFree = e->offRayLive + seg_align((size_t)P * sizeof(float4));
        e->offImage = e->offRayFree + seg_align((size_t)P * sizeof(float4));
        e->offPoints = e->offImage + seg_align((size_t)P * sizeof(uchar4));
        e->offNormals = e->offPoints + seg_align((size_t)P * sizeof(float4));
        e->offVba = e->offNormals + seg_align((size_t)P * sizeof(float4));
        e->segBytes = e->offVba + seg_align((size_t)s.numBlocks * SDF_BLOCK_SIZE3 * sizeof(Voxel));
        rc |= dev_alloc(e, &e->seg, e->segBytes);
        rc |= dev_alloc(e, &s.visIdsOwn, (size_t)s.numBlocks);
        if (!rc)
        {
            cudaMemset(e->seg, 0, e->offVis);
            if (cudaHostAlloc((void **)&e->errHost, sizeof(int), cudaHostAllocMapped) != cudaSuccess)
                rc = gs_set_error(__FILE__, __LINE__, "pinned allocation failed");
            else
                *e->errHost = 0;
            e->peer[rank] = e->seg;
            s.visType = (unsigned char *)(e->seg + e->offVis);
            s.vba = (Voxel *)(e->seg + e->offVba);
            e->rayLive = (float4 *)(e->seg + e->offRayLive), e->rayFree = (float4 *)(e->seg + e->offRayFree);
            e->imageFree = (uchar4 *)(e->seg + e->offImage);
            e->pointsMap = (float4 *)(e->seg + e->offPoints), e->normalsMap = (float4 *)(e->seg + e->offNormals);
        }
    }
    else
    {
        rc |= dev_alloc(e, &s.vba, (size_t)s.numBlocks * SDF_BLOCK_SIZE3);
        rc |= dev_alloc(e, &s.visType, (size_t)s.E);
        s.visIdsOwn = s.visIds;
    }
    rc |= dev_alloc(e, &s.chunkCounts, (size_t)(s.E + 1023) / 1024);
    rc |= dev_alloc(e, &s.state, 8);
    rc |= dev_alloc(e, &e->depth_mm, (size_t)P);
    rc |= dev_alloc(e, &e->rgba, (size_t)P);
    rc |= dev_alloc(e, &e->depth_f, (size_t)P);
    const int mm = ((W + 7) / 8) * ((H + 7) / 8);
    rc |= dev_alloc(e, &e->minmaxLive, (size_t)mm);
    rc |= dev_alloc(e, &e->minmaxFree, (size_t)mm);
    if (world == 1)
    {
        rc |= dev_alloc(e, &e->rayLive, (size_t)P);
        rc |= dev_alloc(e, &e->rayFree, (size_t)P);
        rc |= dev_alloc(e, &e->pointsMap, (size_t)P);
        rc |= dev_alloc(e, &e->normalsMap, (size_t)P);
        rc |= dev_alloc(e, &e->imageFree, (size_t)P);
    }
    if (rc)
    {
        gsb_tsdf_destroy(e);
        return 1;
    }
    e->frame.depth_mm = e->depth_mm, e->frame.rgba = e->rgba, e->frame.depth_f = e->depth_f, e->frame.W = W, e->frame.H = H;
    e->cam.fx = cfg->fx, e->cam.fy = cfg->fy, e->cam.cx = cfg->cx, e->cam.cy = cfg->cy;
    e->tracker = nullptr;
    e->stageTiming = false;
    e->trackingActive = cfg->tracker != 0;
    if (cfg->tracker != 0)
    {
        e->tracker = icp::create_tracker(cfg->tracker, W, H, cfg->view_frustum_min, cfg->view_frustum_max);
        if (!e->tracker)
        {
            gsb_tsdf_destroy(e);
            return gs_set_error(__FILE__, __LINE__, "tracker creation failed");
        }
    }
    *out = e;
    return gsb_tsdf_reset(e);
}

extern "C" void gsb_tsdf_destroy(gsb_tsdf_t *e)
{
    if (!e)
        return;
    cudaStreamSynchronize(e->stream);
    if (e->tracker)
        icp::destroy_tracker(e->tracker);
    if (e->stageTiming)
        for (cudaEvent_t ev : e->stageEv)
            cudaEventDestroy(ev);
    for (int q = 0; q < e->world; q++)
        if (e->ipcOpened[q])
            cudaIpcCloseMemHandle(e->peer[q]);
    if (e->errHost)
        cudaFreeHost(e->errHost);
    for (void *p : e->allocs)
        cudaFree(p);
    cudaStreamDestroy(e->ownStream);
    delete e;
}

// ---- sharded scene: mapping the other ranks' segments (same pattern as gsb_comm_*: CUDA IPC between processes, plain pointers between
// engines of one process)
extern "C" int gsb_tsdf_shard_export(gsb_tsdf_t *e, void *handle64)
{
    if (!e || !handle64 || e->world < 2)
        return gs_set_error(__FILE__, __LINE__, "not a sharded engine");
    cudaIpcMemHandle_t h;
    E_CUDA(cudaIpcGetMemHandle(&h, e->seg));
    memcpy(handle64, &h, 64);
    return 0;
}

extern "C" int gsb_tsdf_shard_attach(gsb_tsdf_t *e, const void *handles)
{
    if (!e || !handles || e->world < 2)
        return gs_set_error(__FILE__, __LINE__, "not a sharded engine");
    E_CUDA(cudaSetDevice(e->cfg.device));
    for (int q = 0; q < e->world; q++)
    {
        if (q == e->rank || e->ipcOpened[q])
            continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)handles + (size_t)q * 64, 64);
        void *p = nullptr;
        E_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        e->peer[q] = (char *)p;
        e->ipcOpened[q] = true;
    }
    fill_shard_view(e);
    e->attached = true;
    if (e->tracker)
        icp::set_shard(e->tracker, e->rank, e->world, e->view.icpXchg, e->errHost);
    return 0;
}

extern "C" int gsb_tsdf_shard_attach_local(gsb_tsdf_t *e, gsb_tsdf_t *const *peers)
{
    if (!e || !peers || e->world < 2)
        return gs_set_error(__FILE__, __LINE__, "not a sharded engine");
    for (int q = 0; q < e->world; q++)
    {
        if (!peers[q] || peers[q]->world != e->world || peers[q]->rank != q || peers[q]->segBytes != e->segBytes)
            return gs_set_error(__FILE__, __LINE__, "peer list does not match this engine");
        if (peers[q]->cfg.device != e->cfg.device)
        {
            int can = 0;
            E_CUDA(cudaDeviceCanAccessPeer(&can, e->cfg.device, peers[q]->cfg.device));
            if (!can)
                return gs_set_error(__FILE__, __LINE__, "no peer access between the devices");
            E_CUDA(cudaSetDevice(e->cfg.device));
            cudaError_t err = cudaDeviceEnablePeerAccess(peers[q]->cfg.device, 0);
            if (err != cudaSuccess && err != cudaErrorPeerAccessAlreadyEnabled)
                return gs_set_error(__FILE__, __LINE__, cudaGetErrorString(err));
            cudaGetLastError();
        }
        e->peer[q] = peers[q]->seg;
    }
    fill_shard_view(e);
    e->attached = true;
    if (e->tracker)
        icp::set_shard(e->tracker, e->rank, e->world, e->view.icpXchg, e->errHost);
    return 0;
}

extern "C" int gsb_tsdf_shard_set_mode(gsb_tsdf_t *e, int mode)
{
    if (!e || e->world < 2 || (mode != 0 && mode != 1))
        return gs_set_error(__FILE__, __LINE__, "not a sharded engine, or mode not 0 / 1");
    if (e->framesProcessed != 0)
        return gs_set_error(__FILE__, __LINE__, "the shard mode can only change on an empty scene");
    e->shardMode = mode;
    if (e->attached)
        fill_shard_view(e);
    return 0;
}

// measurement aid (tools/shard_probe.py): 0 normal, 1 = raycast results stay on this rank, 2 = per-thread instead of bulk peer stores
extern "C" int gsb_tsdf_shard_probe(gsb_tsdf_t *e, int what)
{
    if (!e || e->world < 2)
        return gs_set_error(__FILE__, __LINE__, "not a sharded engine");
    e->view.probe = what;
    return 0;
}

extern "C" int gsb_tsdf_shard_error(gsb_tsdf_t *e) { return (e && e->errHost) ? *e->errHost : 0; }

extern "C" int gsb_tsdf_shard_info(gsb_tsdf_t *e, int *rank, int *world, int *row0, int *row1)
{
    if (!e)
        return gs_set_error(__FILE__, __LINE__, "null argument");
    int y0, y1;
    tsdf::slab_rows(e->cfg.height, e->rank, e->world, y0, y1);
    if (rank) *rank = e->rank;
    if (world) *world = e->world;
    if (row0) *row0 = y0;
    if (row1) *row1 = y1;
    return 0;
}

extern "C" int gsb_tsdf_reset(gsb_tsdf_t *e)
{
    E_CUDA(cudaSetDevice(e->cfg.device));
    tsdf::reset_scene(e->scene, e->stream);
    const int W = e->cfg.width, H = e->cfg.height, P = W * H;
    E_CUDA(cudaMemsetAsync(e->rayLive, 0, sizeof(float4) * P, e->stream));
    E_CUDA(cudaMemsetAsync(e->rayFree, 0, sizeof(float4) * P, e->stream));
    E_CUDA(cudaMemsetAsync(e->pointsMap, 0, sizeof(float4) * P, e->stream));
    E_CUDA(cudaMemsetAsync(e->normalsMap, 0, sizeof(float4) * P, e->stream));
    E_CUDA(cudaMemsetAsync(e->imageFree, 0, sizeof(uchar4) * P, e->stream));
    e->pose_d = se3::Pose();
    e->pose_pointCloud = se3::Pose();
    e->cam.M = e->pose_d.M;
    e->cam.invM = e->pose_d.get_invM();
    e->framesProcessed = 0;
    e->trackingFrames = 0;
    e->agePointCloud = -1;
    e->haveFrame = false;
    e->lastRgba = nullptr;
    E_CUDA(cudaStreamSynchronize(e->stream));
    E_CUDA(cudaGetLastError());
    return 0;
}

// ITMBasicEngine::LoadFromFile (Core/ITMBasicEngine.tpp:137-171) without the file I/O: resetAll(), then the scene arrays of
// ITMScene::LoadFromDirectory (hash.dat, voxel.dat, last.txt, vba.txt; Objects/Scene/ITMVoxelBlockHash.h:143-155, ITMLocalVBA.h:52-66)
// are taken from host memory.  Like the reference, visibility is rebuilt by the next frame.
extern "C" int gsb_tsdf_load_scene(gsb_tsdf_t *e, const void *hash_entries_host, size_t n_entries, const void *voxels_host, size_t n_voxels,
                                   int last_free_block_id, int last_free_excess_list_id)
{
    if (!e || !hash_entries_host || !voxels_host)
        return gs_set_error(__FILE__, __LINE__, "null argument");
    if (n_entries != (size_t)e->scene.E || n_voxels != (size_t)e->scene.numBlocks * SDF_BLOCK_SIZE3)
        return gs_set_error(__FILE__, __LINE__, "scene arrays do not match this engine's hash table / voxel block array sizes");
    if (last_free_block_id < -1 || last_free_block_id >= e->scene.numBlocks || last_free_excess_list_id < -1 ||
        last_free_excess_list_id >= SDF_EXCESS_LIST_SIZE)
        return gs_set_error(__FILE__, __LINE__, "free-list heads out of range");
    if (gsb_tsdf_reset(e))
        return 1;
    E_CUDA(cudaMemcpyAsync(e->scene.table, hash_entries_host, n_entries * sizeof(HashEntry), cudaMemcpyHostToDevice, e->stream));
    E_CUDA(cudaMemcpyAsync(e->scene.vba, voxels_host, n_voxels * sizeof(Voxel), cudaMemcpyHostToDevice, e->stream));
    const int heads[2] = {last_free_block_id, last_free_excess_list_id};
    E_CUDA(cudaMemcpyAsync(e->scene.state, heads, sizeof heads, cudaMemcpyHostToDevice, e->stream));
    E_CUDA(cudaStreamSynchronize(e->stream));
    return 0;
}

extern "C" int gsb_tsdf_set_stream(gsb_tsdf_t *e, void *st)
{
    e->stream = st ? (cudaStream_t)st : e->ownStream;
    return 0;
}
// ITMBasicEngine::turnOnTracking / turnOffTracking (Core/ITMBasicEngine.tpp:400-410): with tracking off ProcessFrame takes the pose
// from gtC2wPoses[frame] (:278)
extern "C" int gsb_tsdf_set_tracking(gsb_tsdf_t *e, int on)
{
    if (!e)
        return gs_set_error(__FILE__, __LINE__, "null argument");
    if (on && !e->tracker)
        return gs_set_error(__FILE__, __LINE__, "engine was created without a tracker (tracker == 0)");
    e->trackingActive = on != 0;
    return 0;
}
extern "C" void *gsb_tsdf_get_stream(gsb_tsdf_t *e) { return (void *)e->stream; }
extern "C" int gsb_tsdf_sync(gsb_tsdf_t *e)
{
    E_CUDA(cudaStreamSynchronize(e->stream));
    E_CUDA(cudaGetLastError());
    return 0;
}

static void refresh_camera(gsb_tsdf *e)
{
    e->cam.M = e->pose_d.M;
    e->cam.invM = e->pose_d.get_invM();
}

// stages after the view is on the device
static int process_resident(gsb_tsdf *e, const float *gt_c2w)
{
    cudaStream_t st = e->stream;
    if (!e->attached)
        return gs_set_error(__FILE__, __LINE__, "sharded engine: the peer segments are not attached (gsb_tsdf_shard_attach)");
    E_CUDA(cudaSetDevice(e->cfg.device));   // several engines of one process may sit on different devices
    auto mark = [&](int i)
    {
        if (e->stageTiming)
            cudaEventRecord(e->stageEv[i], st);
    };
    mark(0);
    // --- tracking (ITMBasicEngine.tpp:273-280)
    if (!e->trackingActive)
    {
        if (!gt_c2w)
            return gs_set_error(__FILE__, __LINE__, "tracking is off (tracker == 0 or gsb_tsdf_set_tracking(0)): a ground-truth camera-to-world pose is needed");
        Mat4 c2w;
        memcpy(c2w.m, gt_c2w, 64);
        e->pose_d.set_invM(c2w);
        e->pose_d.coerce();
    }
    else
    {
        // ITMTrackingController::Track -> tracker->TrackCamera; needs this frame's float depth, which the
        // reference produces in UpdateView.  Ours is produced by the allocation pass, so convert first.
        if (e->agePointCloud != -1)
        {
            if (e->agePointCloud >= 0)
                e->trackingFrames++;
            else
                e->trackingFrames = 0;
            icp::convert_depth(e->frame.depth_mm, e->depth_f, e->cfg.width, e->cfg.height, st);
            Mat4 scenePose = e->pose_pointCloud.M;
            int rc = icp::track_camera(e->tracker, e->depth_f, e->pointsMap, e->normalsMap, e->cam.fx, e->cam.fy, e->cam.cx, e->cam.cy, scenePose,
                                       e->trackingFrames, &e->pose_d, st);
            if (rc)
                return rc;
        }
    }
    refresh_camera(e);
    mark(1);
    // --- fusion (ITMDenseMapper::ProcessFrame)
    tsdf::allocate(e->scene, e->frame, e->cam, st);   // replicated on every rank of a sharded scene (deterministic)
    mark(2);
    tsdf::integrate(e->scene, e->frame, e->cam, e->cfg.integrate_variant, st);   // sharded: the visible blocks this rank owns
    mark(3);
    e->framesProcessed++;
    // --- ITMTrackingController::Prepare (always: requiresPointCloudRendering() is constant true)
    const int W = e->cfg.width, H = e->cfg.height;
    tsdf::expected_depth_live(e->scene, e->cam, W, H, e->minmaxLive, st);
    mark(4);
    if (e->world == 1)
    {
        tsdf::raycast(e->scene, e->cam, W, H, e->minmaxLive, e->rayLive, nullptr, true, st);
        mark(5);
        tsdf::icp_maps(e->scene, e->cam, W, H, e->rayLive, e->pointsMap, e->normalsMap, st);
    }
    else
    {
        // Sharded scene.  Barrier A: every rank has integrated its blocks of this frame (and has finished reading what the raycast is
        // about to overwrite).  The raycast reads the other ranks' voxels (mode 0), stores its visibility marks and its rows into every
        // rank.  Barrier B: all of that has landed, nobody reads voxels any more -> the next frame may integrate; the ICP maps of this
        // rank's slab follow (every rank receives every row while tracking is on, then barrier C).
        shard_barrier(e);
        tsdf::raycast_sharded(e->scene, e->view, e->cam, W, H, e->minmaxLive, true, st);
        shard_barrier(e);
        mark(5);
        tsdf::icp_maps_sharded(e->scene, e->view, e->cam, W, H, e->trackingActive, st);
        if (e->trackingActive)
            shard_barrier(e);
    }
    mark(6);
    e->pose_pointCloud = e->pose_d;
    e->agePointCloud = (e->agePointCloud == -1) ? -2 : 0;
    e->haveFrame = true;
    e->lastRgba = e->frame.rgba;
    E_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int gsb_tsdf_process_frame(gsb_tsdf_t *e, const uint8_t *rgba_host, const int16_t *depth_mm_host, const float *gt_c2w)
{
    if (!e || !rgba_host || !depth_mm_host)
        return gs_set_error(__FILE__, __LINE__, "null argument");
    const size_t P = (size_t)e->cfg.width * e->cfg.height;
    E_CUDA(cudaSetDevice(e->cfg.device));
    // B1: ITMViewBuilder::UpdateView H2D (ITMViewBuilder_CUDA.cu:60-61); async when the host buffers are pinned
    E_CUDA(cudaMemcpyAsync(e->rgba, rgba_host, P * 4, cudaMemcpyHostToDevice, e->stream));
    E_CUDA(cudaMemcpyAsync(e->depth_mm, depth_mm_host, P * 2, cudaMemcpyHostToDevice, e->stream));
    return process_resident(e, gt_c2w);
}

extern "C" int gsb_tsdf_process_frame_device(gsb_tsdf_t *e, const void *rgba_dev, const void *depth_mm_dev, const float *gt_c2w)
{
    if (!e || !rgba_dev || !depth_mm_dev)
        return gs_set_error(__FILE__, __LINE__, "null argument");
    e->frame.rgba = (const uchar4 *)rgba_dev;
    e->frame.depth_mm = (const short *)depth_mm_dev;
    int rc = process_resident(e, gt_c2w);
    e->frame.rgba = e->rgba;
    e->frame.depth_mm = e->depth_mm;
    return rc;
}

extern "C" int gsb_tsdf_run_raycast(gsb_tsdf_t *e, const float *c2w, float fx, float fy, float cx, float cy)
{
    if (!e || !c2w)
        return gs_set_error(__FILE__, __LINE__, "null argument");
    // slam_pipeline.cpp:373-380 builds an SE3Pose with SetInvM(c2w); runRaycast then uses pose->GetM() / GetInvM()
    se3::Pose p;
    Mat4 m;
    memcpy(m.m, c2w, 64);
    p.set_invM(m);
    tsdf::Camera cam;
    cam.M = p.M;
    cam.invM = p.get_invM();
    cam.fx = fx, cam.fy = fy, cam.cx = cx, cam.cy = cy;
    const int W = e->cfg.width, H = e->cfg.height;
    E_CUDA(cudaSetDevice(e->cfg.device));
    if (e->world == 1)
    {
        tsdf::expected_depth_free(e->scene, cam, W, H, e->minmaxFree, e->stream);
        tsdf::raycast(e->scene, cam, W, H, e->minmaxFree, e->rayFree, e->imageFree, false, e->stream);
    }
    else
    {
        // sharded scene: every rank marches its slab of rows and stores the result into every rank's free-view images.  The barrier in
        // front says "every rank is done with the previous free-view images and with integrating"; the one behind says "all rows have
        // landed everywhere and nobody reads voxels any more".
        if (!e->attached)
            return gs_set_error(__FILE__, __LINE__, "sharded engine: the peer segments are not attached (gsb_tsdf_shard_attach)");
        tsdf::expected_depth_free(e->scene, cam, W, H, e->minmaxFree, e->stream);
        shard_barrier(e);
        tsdf::raycast_sharded(e->scene, e->view, cam, W, H, e->minmaxFree, false, e->stream);
        shard_barrier(e);
    }
    E_CUDA(cudaGetLastError());
    return 0;
}

// diagnostic: march statistics of a free-view raycast from c2w (does not touch the free-view outputs); synchronises
extern "C" int gsb_tsdf_raycast_stats(gsb_tsdf_t *e, const float *c2w, float fx, float fy, float cx, float cy, unsigned long long *totals8_host)
{
    if (!e || !c2w || !totals8_host)
        return gs_set_error(__FILE__, __LINE__, "null argument");
    se3::Pose p;
    Mat4 m;
    memcpy(m.m, c2w, 64);
    p.set_invM(m);
    tsdf::Camera cam;
    cam.M = p.M;
    cam.invM = p.get_invM();
    cam.fx = fx, cam.fy = fy, cam.cx = cx, cam.cy = cy;
    const int W = e->cfg.width, H = e->cfg.height;
    unsigned long long *tot = nullptr;
    E_CUDA(cudaMalloc((void **)&tot, 64));
    E_CUDA(cudaMemsetAsync(tot, 0, 64, e->stream));
    tsdf::expected_depth_free(e->scene, cam, W, H, e->minmaxFree, e->stream);
    tsdf::raycast_stats(e->scene, cam, W, H, e->minmaxFree, tot, e->stream);
    E_CUDA(cudaMemcpyAsync(totals8_host, tot, 64, cudaMemcpyDeviceToHost, e->stream));
    E_CUDA(cudaStreamSynchronize(e->stream));
    cudaFree(tot);
    return 0;
}

extern "C" const void *gsb_tsdf_current_rgba_dev(gsb_tsdf_t *e) { return e->lastRgba; }
extern "C" const void *gsb_tsdf_free_image_dev(gsb_tsdf_t *e) { return e->imageFree; }
extern "C" const void *gsb_tsdf_free_vertex_dev(gsb_tsdf_t *e) { return e->rayFree; }
extern "C" const void *gsb_tsdf_live_vertex_dev(gsb_tsdf_t *e) { return e->rayLive; }
extern "C" const void *gsb_tsdf_points_map_dev(gsb_tsdf_t *e) { return e->pointsMap; }
extern "C" const void *gsb_tsdf_normals_map_dev(gsb_tsdf_t *e) { return e->normalsMap; }

extern "C" int gsb_tsdf_get_pose(gsb_tsdf_t *e, float *M, float *invM)
{
    if (M)
        memcpy(M, e->pose_d.M.m, 64);
    if (invM)
    {
        Mat4 i = e->pose_d.get_invM();
        memcpy(invM, i.m, 64);
    }
    return 0;
}
extern "C" float gsb_tsdf_voxel_size(gsb_tsdf_t *e) { return e->cfg.voxel_size; }
extern "C" int gsb_tsdf_frames_processed(gsb_tsdf_t *e) { return e->framesProcessed; }

extern "C" int gsb_tsdf_counter(gsb_tsdf_t *e, int which, int *value)
{
    if (which < 0 || which > 7)
        return gs_set_error(__FILE__, __LINE__, "bad counter id");
    E_CUDA(cudaMemcpyAsync(value, e->scene.state + which, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    E_CUDA(cudaStreamSynchronize(e->stream));
    return 0;
}

extern "C" int gsb_tsdf_read(gsb_tsdf_t *e, int what, void *dst, size_t bytes)
{
    const size_t P = (size_t)e->cfg.width * e->cfg.height;
    const size_t mm = (size_t)((e->cfg.width + 7) / 8) * ((e->cfg.height + 7) / 8);
    const void *src = nullptr;
    size_t avail = 0;
    switch (what)
    {
    case GSB_TSDF_HASH_TABLE: src = e->scene.table, avail = (size_t)e->scene.E * 16; break;
    case GSB_TSDF_VOXELS: src = e->scene.vba, avail = (size_t)e->scene.numBlocks * SDF_BLOCK_SIZE3 * 8; break;
    case GSB_TSDF_VISIBLE_IDS: src = e->scene.visIds, avail = (size_t)e->scene.numBlocks * 4; break;
    case GSB_TSDF_VISIBLE_TYPES: src = e->scene.visType, avail = (size_t)e->scene.E; break;
    case GSB_TSDF_DEPTH_F: src = e->depth_f, avail = P * 4; break;
    case GSB_TSDF_MINMAX_LIVE: src = e->minmaxLive, avail = mm * 8; break;
    case GSB_TSDF_MINMAX_FREE: src = e->minmaxFree, avail = mm * 8; break;
    case GSB_TSDF_RAYCAST_LIVE: src = e->rayLive, avail = P * 16; break;
    case GSB_TSDF_RAYCAST_FREE: src = e->rayFree, avail = P * 16; break;
    case GSB_TSDF_POINTS_MAP: src = e->pointsMap, avail = P * 16; break;
    case GSB_TSDF_NORMALS_MAP: src = e->normalsMap, avail = P * 16; break;
    case GSB_TSDF_IMAGE_FREE: src = e->imageFree, avail = P * 4; break;
    default: return gs_set_error(__FILE__, __LINE__, "bad read id");
    }
    if (bytes > avail)
        return gs_set_error(__FILE__, __LINE__, "read larger than the buffer");
    E_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, e->stream));
    E_CUDA(cudaStreamSynchronize(e->stream));
    return 0;
}

// ITMBasicEngine::SaveSceneToMesh (Core/ITMBasicEngine.tpp:105-117) up to the file: marching cubes over every allocated block.
// tri_dev: device buffer of max_tri triangles, 18 floats each (p0 p1 p2 in metres, c0 c1 c2 in 0..1), or NULL to count only.
// The reference keeps at most max_tri - 1 triangles (noMaxTriangles rule of the CPU mesher); *n_tri = triangles written (or present).
extern "C" int gsb_tsdf_mesh(gsb_tsdf_t *e, float *tri_dev, long long max_tri, long long *n_tri)
{
    if (!e || !n_tri || (tri_dev && max_tri <= 0))
        return gs_set_error(__FILE__, __LINE__, "invalid argument");
    E_CUDA(cudaSetDevice(e->cfg.device));
    void *scratch = nullptr;
    E_CUDA(cudaMalloc(&scratch, tsdf::mesh_scratch_bytes(e->scene)));
    const int rc = tsdf::mesh_scene(e->scene, e->world > 1 ? &e->view : nullptr, scratch, tri_dev, max_tri, n_tri, e->stream);
    cudaFree(scratch);
    return rc;
}

// Measurement aid: CUDA events between the stages of ProcessFrame (track | allocate | integrate | expected depth | raycast | ICP maps)
extern "C" int gsb_tsdf_enable_stage_timing(gsb_tsdf_t *e, int on)
{
    if (!e)
        return gs_set_error(__FILE__, __LINE__, "null argument");
    if (on && !e->stageTiming)
        for (cudaEvent_t &ev : e->stageEv)
            E_CUDA(cudaEventCreate(&ev));
    if (!on && e->stageTiming)
        for (cudaEvent_t ev : e->stageEv)
            cudaEventDestroy(ev);
    e->stageTiming = on != 0;
    return 0;
}
// device times (ms) of the six stages of the last ProcessFrame; synchronises the stream
extern "C" int gsb_tsdf_stage_times(gsb_tsdf_t *e, float *ms6)
{
    if (!e || !ms6 || !e->stageTiming || !e->haveFrame)
        return gs_set_error(__FILE__, __LINE__, "stage timing is not enabled or no frame was processed");
    E_CUDA(cudaEventSynchronize(e->stageEv[6]));
    for (int i = 0; i < 6; i++)
        E_CUDA(cudaEventElapsedTime(&ms6[i], e->stageEv[i], e->stageEv[i + 1]));
    return 0;
}

extern "C" int gsb_tsdf_run_stage(gsb_tsdf_t *e, int stage)
{
    if (!e->haveFrame)
        return gs_set_error(__FILE__, __LINE__, "run_stage needs a processed frame");
    E_CUDA(cudaSetDevice(e->cfg.device));
    const int W = e->cfg.width, H = e->cfg.height;
    switch (stage)
    {
    case 0: tsdf::allocate(e->scene, e->frame, e->cam, e->stream); break;
    case 1: tsdf::integrate(e->scene, e->frame, e->cam, e->cfg.integrate_variant, e->stream); break;
    case 2: tsdf::expected_depth_live(e->scene, e->cam, W, H, e->minmaxLive, e->stream); break;
    case 3:
        if (e->world == 1)
            tsdf::raycast(e->scene, e->cam, W, H, e->minmaxLive, e->rayLive, nullptr, true, e->stream);
        else
            tsdf::raycast_sharded(e->scene, e->view, e->cam, W, H, e->minmaxLive, true, e->stream);   // no barriers: timing aid only
        break;
    case 4:
        if (e->world == 1)
            tsdf::icp_maps(e->scene, e->cam, W, H, e->rayLive, e->pointsMap, e->normalsMap, e->stream);
        else
            tsdf::icp_maps_sharded(e->scene, e->view, e->cam, W, H, false, e->stream);
        break;
    default: return gs_set_error(__FILE__, __LINE__, "bad stage id");
    }
    E_CUDA(cudaGetLastError());
    return 0;
}

// ---- C. ICP tracker access (parity / diagnostics)
extern "C" int gsb_tsdf_icp_eval(gsb_tsdf_t *e, int level, const float *approx_invM, int *n_valid, float *f, float *nabla6, float *hessian36)
{
    if (!e || !e->tracker)
        return gs_set_error(__FILE__, __LINE__, "engine was created without a tracker");
    if (!e->haveFrame)
        return gs_set_error(__FILE__, __LINE__, "icp_eval needs a processed frame");
    Mat4 inv;
    memcpy(inv.m, approx_invM, 64);
    return icp::icp_eval(e->tracker, e->depth_f, e->pointsMap, e->normalsMap, e->cam.fx, e->cam.fy, e->cam.cx, e->cam.cy, e->pose_pointCloud.M,
                         e->trackingFrames, level, inv, n_valid, f, nabla6, hessian36, e->stream);
}
extern "C" int gsb_tsdf_set_tracking_frames(gsb_tsdf_t *e, int n)
{
    e->trackingFrames = n;
    return 0;
}
extern "C" int gsb_tsdf_tracker_result(gsb_tsdf_t *e, int *result, float *score, int *iterations)
{
    if (!e || !e->tracker)
        return gs_set_error(__FILE__, __LINE__, "engine was created without a tracker");
    icp::tracker_result(e->tracker, result, score, iterations);
    return 0;
}
extern "C" int gsb_tsdf_depth_level(gsb_tsdf_t *e, int level, float *dst_host, int *w, int *h)
{
    if (!e || !e->tracker)
        return gs_set_error(__FILE__, __LINE__, "engine was created without a tracker");
    const float *p = icp::level_depth(e->tracker, level, w, h);
    if (level == 0)
        p = e->depth_f;
    if (!p)
        return gs_set_error(__FILE__, __LINE__, "bad pyramid level");
    if (dst_host)
    {
        E_CUDA(cudaMemcpyAsync(dst_host, p, sizeof(float) * (size_t)(*w) * (*h), cudaMemcpyDeviceToHost, e->stream));
        E_CUDA(cudaStreamSynchronize(e->stream));
    }
    return 0;
}

// trackingState->pose_d->SetInvM(invM); Coerce()  -- initial pose when tracking is on (the reference starts from identity)
extern "C" int gsb_tsdf_set_pose(gsb_tsdf_t *e, const float *invM)
{
    if (!e || !invM)
        return gs_set_error(__FILE__, __LINE__, "null argument");
    Mat4 m;
    memcpy(m.m, invM, 64);
    e->pose_d.set_invM(m);
    e->pose_d.coerce();
    refresh_camera(e);
    return 0;
}
