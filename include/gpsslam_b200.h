/* gpsslam_b200 -- C ABI of the B200-native per-frame SLAM compute engine.
 *
 * This is the drop-in boundary for the three hot paths of MisEty/GPS-SLAM:
 *   A  gsplat "GES" rasteriser forward + per-Gaussian backward   (gsb_gs_*)
 *   B  InfiniTAM hashed-voxel TSDF allocate / integrate / raycast (gsb_tsdf_*)
 *   C  ITM depth-tracker ICP reduction + LM solve                 (gsb_icp_*, gsb_tsdf_process_frame with tracking)
 *
 * Conventions
 *   - plain C: opaque handles, raw pointers and sizes, no C++/torch types;
 *   - every function returns 0 on success, non-zero on failure; gsb_last_error() gives the message
 *     (thread-local).  Nothing throws, nothing calls exit();
 *   - pointers named *_dev are device pointers on the engine's CUDA device, *_host are host pointers
 *     (pinned memory makes the copies asynchronous);
 *   - 4x4 matrices are 16 floats in ORUtils::Matrix4 order, i.e. column-major m[col*4+row]
 *     (reference InfiniTAM/ORUtils/Matrix.h:26-36); gsplat-side view matrices are row-major [4][4]
 *     exactly as the reference passes them (gsplat/gsplat_wapper.hpp:100-115);
 *   - all work is enqueued on the engine's stream (gsb_*_set_stream, default: a private non-blocking stream);
 *     calls return without synchronising unless documented otherwise;
 *   - there is NO CPU fallback: without a CUDA device every create() fails.
 */
#ifndef GPSSLAM_B200_H
#define GPSSLAM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char *gsb_last_error(void);
/* library / build identification: "gpsslam_b200 <version> sm_100a" */
const char *gsb_version(void);
/* kernels launched by this library since load (host-side count, all engines) */
long long gsb_launch_count(void);

/* ===================================================================================================
 * B.  TSDF engine  -- replaces ITMLib::ITMBasicEngine<ITMVoxel_s_rgb, ITMVoxelBlockHash>
 *     (reference InfiniTAM/ITMLib/Core/ITMBasicEngine.h:52-110, .tpp:260-385, 500-526)
 * =================================================================================================== */
typedef struct gsb_tsdf gsb_tsdf_t;

typedef struct gsb_tsdf_config
{
    int width, height;            /* ITMRGBDCalib intrinsics_d / intrinsics_rgb (identical, createTsdfEngine) */
    float fx, fy, cx, cy;
    float voxel_size;             /* ITMSceneParams::voxelSize         (office0.yaml:69)                    */
    float mu;                     /* ITMSceneParams::mu  (trunc_dist)                                        */
    float view_frustum_min;       /* ITMSceneParams::viewFrustum_min                                          */
    float view_frustum_max;       /* ITMSceneParams::viewFrustum_max                                          */
    int max_w;                    /* ITMSceneParams::maxW, reference default 100 (ITMLibSettings.cpp:10); 1..255 (one byte per voxel) */
    int num_blocks;               /* SDF_LOCAL_BLOCK_NUM, reference 0x40000; 0 = default                      */
    int tracker;                  /* 0 = ground-truth poses (turnOffTracking), 1 = extended, 2 = icp           */
    int device;                   /* CUDA device ordinal                                                       */
    int integrate_variant;        /* must be 0 (TMA-pipelined kernel); kept so that the struct layout of round 1 callers stays valid */
} gsb_tsdf_config_t;

void gsb_tsdf_default_config(gsb_tsdf_config_t *cfg);
int gsb_tsdf_create(const gsb_tsdf_config_t *cfg, gsb_tsdf_t **out);
/* Voxel hash sharded by spatial block over `world` GPUs of one box (one engine per GPU, one process per GPU or several engines in one
 * process).  The hash table, allocation and visibility lists are computed identically by every rank (every rank is given
 * every frame); the voxel DATA of a block lives on rank hashIndex(blockPos) mod world only (reference hash: ITMRepresentationAccess.h:7-11).
 * ProcessFrame integrates the visible blocks this rank owns (per-block independent, ITMSceneReconstructionEngine_CUDA.tcu:348-383);
 * raycasts are split by image rows, read the voxels from their owner over NVLink peer memory (castRay, ITMVisualisationEngine_Shared.h:
 * 122-221) and store their rows straight into the other ranks' images; cross-GPU ordering is two flag barriers per frame through peer
 * memory.  Every rank must issue the same sequence of process_frame / run_raycast calls.  Results are bit-identical to world == 1:
 * hash table, visible list, poses, live and free-view vertex / colour images on EVERY rank; GSB_TSDF_VOXELS holds the blocks this rank owns
 * (mode 1: all blocks); the ICP maps hold this rank's rows (gsb_tsdf_shard_info) unless tracking is on (then every rank holds all rows).
 * After create_sharded the peers' segments must be mapped once: between processes exchange the 64-byte handles of shard_export (any
 * transport) and call shard_attach(handles of all ranks in rank order); engines inside one process use shard_attach_local. */
int gsb_tsdf_create_sharded(const gsb_tsdf_config_t *cfg, int rank, int world, gsb_tsdf_t **out);
int gsb_tsdf_shard_export(gsb_tsdf_t *e, void *handle64);
int gsb_tsdf_shard_attach(gsb_tsdf_t *e, const void *handles /* world x 64 bytes */);
int gsb_tsdf_shard_attach_local(gsb_tsdf_t *e, gsb_tsdf_t *const *peers /* world engines in rank order */);
/* mode 0: storage-sharded -- a block's voxels exist on its owner only, raycasts read them over NVLink (1/world of the voxel memory
 * per GPU).  mode 1 (default): owner computes, everybody stores -- the integrate kernel writes each block it updated into every rank's array (4 KB TMA
 * bulk stores over NVLink), every rank keeps a complete copy, raycasts (still split by rows) read local memory.  Same results either way;
 * every rank must use the same mode; only on an empty scene. */
int gsb_tsdf_shard_set_mode(gsb_tsdf_t *e, int mode);
int gsb_tsdf_shard_probe(gsb_tsdf_t *e, int what);   /* measurement aid: 0 normal, 1 raycast rows stay local, 2 per-thread peer stores */
int gsb_tsdf_shard_error(gsb_tsdf_t *e);   /* non-zero: a cross-GPU barrier or exchange timed out (a peer never arrived) */
int gsb_tsdf_shard_info(gsb_tsdf_t *e, int *rank, int *world, int *row0, int *row1);   /* rows [row0, row1) of the ICP maps this rank computes */
void gsb_tsdf_destroy(gsb_tsdf_t *e);
int gsb_tsdf_reset(gsb_tsdf_t *e);                               /* ITMBasicEngine::resetAll */
int gsb_tsdf_set_stream(gsb_tsdf_t *e, void *cuda_stream);       /* cudaStream_t; NULL = private stream */
void *gsb_tsdf_get_stream(gsb_tsdf_t *e);
int gsb_tsdf_sync(gsb_tsdf_t *e);

/* ITMBasicEngine::ProcessFrame(rgb, rawDepth): host RGBA8 [h*w*4] + int16 depth in mm [h*w].
 * gt_c2w: camera-to-world for this frame when tracker == 0 (gtC2wPoses[framesProcessed]), else NULL.
 * Runs UpdateView -> (Track) -> AllocateSceneFromDepth -> IntegrateIntoScene -> CreateExpectedDepths -> CreateICPMaps. */
int gsb_tsdf_process_frame(gsb_tsdf_t *e, const uint8_t *rgba_host, const int16_t *depth_mm_host, const float *gt_c2w);
/* same with the frame already resident in HBM */
int gsb_tsdf_process_frame_device(gsb_tsdf_t *e, const void *rgba_dev, const void *depth_mm_dev, const float *gt_c2w);

/* ITMBasicEngine::runRaycast(pose, intrinsics) (free view): FindVisibleBlocks + CreateExpectedDepths + raycast +
 * colour-from-volume.  c2w = pose->GetInvM().  Results stay valid until the next run_raycast. */
int gsb_tsdf_run_raycast(gsb_tsdf_t *e, const float *c2w, float fx, float fy, float cx, float cy);
const void *gsb_tsdf_free_image_dev(gsb_tsdf_t *e);   /* GetFreeImage()->GetData(MEMORYDEVICE_CUDA): uchar4 [h*w]  */
const void *gsb_tsdf_free_vertex_dev(gsb_tsdf_t *e);  /* GetFreeVertex(): float4 [h*w], xyz in voxel units, w=conf+1 */
const void *gsb_tsdf_current_rgba_dev(gsb_tsdf_t *e); /* uchar4 [h*w]: the RGBA frame the last ProcessFrame consumed        */
const void *gsb_tsdf_live_vertex_dev(gsb_tsdf_t *e);  /* GetLiveVertex()                                             */
const void *gsb_tsdf_points_map_dev(gsb_tsdf_t *e);   /* trackingState->pointCloud->locations (metres, w=conf+1/-1)  */
const void *gsb_tsdf_normals_map_dev(gsb_tsdf_t *e);  /* trackingState->pointCloud->colours                          */

/* diagnostic (tools/raycast_stats.py): march statistics of a free-view raycast from c2w -- totals8_host = rays, march steps, steps through
 * unallocated space, trilinear reads, steps that entered another voxel block, sum over warps of the slowest ray's steps, warps, 0.
 * Leaves the free-view outputs untouched; synchronises. */
int gsb_tsdf_raycast_stats(gsb_tsdf_t *e, const float *c2w, float fx, float fy, float cx, float cy, unsigned long long *totals8_host);

/* GetTrackingState()->pose_d: M = GetM() (world->camera), invM = GetInvM() */
int gsb_tsdf_get_pose(gsb_tsdf_t *e, float *M, float *invM);
int gsb_tsdf_set_pose(gsb_tsdf_t *e, const float *invM);          /* pose_d->SetInvM(invM); Coerce() */
float gsb_tsdf_voxel_size(gsb_tsdf_t *e);
int gsb_tsdf_frames_processed(gsb_tsdf_t *e);
/* ITMBasicEngine::turnOnTracking / turnOffTracking (Core/ITMBasicEngine.h:91-92): off = every following ProcessFrame takes gt_c2w
 * (createTsdfEngine with use_gt_pose, slam/InfiniTAM_tools.cpp:59-63); on needs an engine created with tracker != 0 */
int gsb_tsdf_set_tracking(gsb_tsdf_t *e, int on);

/* state read-back (synchronises).  `what`: */
enum
{
    GSB_TSDF_HASH_TABLE = 0,     /* ITMHashEntry[1179648], 16 B each                      */
    GSB_TSDF_VOXELS = 1,         /* ITMVoxel_s_rgb[num_blocks*512], 8 B each              */
    GSB_TSDF_VISIBLE_IDS = 2,    /* int[count]  (count via gsb_tsdf_counter)              */
    GSB_TSDF_VISIBLE_TYPES = 3,  /* uchar[1179648]                                        */
    GSB_TSDF_DEPTH_F = 4,        /* float[h*w]                                            */
    GSB_TSDF_MINMAX_LIVE = 5,    /* float2[ceil(h/8)*ceil(w/8)] (compact 1/8-res image)    */
    GSB_TSDF_MINMAX_FREE = 6,
    GSB_TSDF_RAYCAST_LIVE = 7,   /* float4[h*w]                                           */
    GSB_TSDF_RAYCAST_FREE = 8,
    GSB_TSDF_POINTS_MAP = 9,
    GSB_TSDF_NORMALS_MAP = 10,
    GSB_TSDF_IMAGE_FREE = 11     /* uchar4[h*w]                                           */
};
int gsb_tsdf_read(gsb_tsdf_t *e, int what, void *dst_host, size_t bytes);
/* which: 0 lastFreeBlockId, 1 lastFreeExcessListId, 2 noVisibleEntries, 3 error flag, 6 visible entries owned by this rank (sharded scene);
 * synchronises */
int gsb_tsdf_counter(gsb_tsdf_t *e, int which, int *value);

/* ITMBasicEngine::SaveSceneToMesh (Core/ITMBasicEngine.tpp:105-117) minus the file: marching cubes over every allocated voxel block
 * (ITMMeshingEngine_*::MeshScene, Engines/Meshing/Shared/ITMMeshingEngine_Shared.h:277-470).  tri_dev: device buffer of max_tri triangles,
 * 18 floats each = p0 p1 p2 (metres) c0 c1 c2 (0..1), or NULL to count only.  Deterministic, in the CPU mesher's order (ascending hash entry,
 * z, y, x, case-table order); like the reference at most max_tri - 1 triangles are kept.  *n_tri = triangles written (or present when
 * tri_dev is NULL).  Synchronises.  gps_slam_b200/checkpoint.py writes the reference's ASCII PLY from it. */
int gsb_tsdf_mesh(gsb_tsdf_t *e, float *tri_dev, long long max_tri, long long *n_tri);

/* ITMBasicEngine::LoadFromFile (Core/ITMBasicEngine.tpp:137-171) minus the file I/O: resets the engine, then installs the scene from
 * host arrays in the reference's own layouts (hash.dat / voxel.dat payloads, lastFreeBlockId of vba.txt, lastFreeExcessListId of last.txt).
 * SaveToFile is gsb_tsdf_read(GSB_TSDF_HASH_TABLE / GSB_TSDF_VOXELS) + gsb_tsdf_counter(0 / 1); gps_slam_b200/checkpoint.py writes and
 * reads the reference's Scene/ directory with them. */
int gsb_tsdf_load_scene(gsb_tsdf_t *e, const void *hash_entries_host, size_t n_entries, const void *voxels_host, size_t n_voxels,
                        int last_free_block_id, int last_free_excess_list_id);

/* Single stages on the current frame / pose, for profiling and stage-level parity.
 * stage: 0 allocate (B1-B4), 1 integrate (B5), 2 expected depth (B6), 3 raycast (B7), 4 ICP maps (B8) */
int gsb_tsdf_run_stage(gsb_tsdf_t *e, int stage);
/* Measurement aid (bench.py roofline): CUDA events between the stages of ProcessFrame, so that each stage's device time is read
 * from fresh frames inside the loop.  ms6 = track | allocate | integrate | expected depth | raycast | ICP maps of the last frame;
 * gsb_tsdf_stage_times synchronises the engine's stream. */
int gsb_tsdf_enable_stage_timing(gsb_tsdf_t *e, int on);
int gsb_tsdf_stage_times(gsb_tsdf_t *e, float *ms6);

/* ===================================================================================================
 * C.  ICP tracker  -- replaces ITMExtendedTracker (tracker == 1, the reference's compiled-in default) and ITMDepthTracker
 *     (tracker == 2) (reference InfiniTAM/ITMLib/Trackers/Interface/ITMExtendedTracker.cpp:470-665, ITMDepthTracker.cpp:233-298).
 *     Tracking itself runs inside gsb_tsdf_process_frame (ITMTrackingController::Track); these calls expose its pieces.
 * =================================================================================================== */
/* one evaluation of the ICP normal equations at pyramid `level` for the camera->world estimate approx_invM (16 floats,
 * column-major) against the current raycast maps: ComputeGandH_Depth (extended: f is the un-normalised sum) / ComputeGandH
 * (icp: f / n, or 1e5 when n <= 100).  hessian36 uses the reference's layout hessian[r + c*6]. Synchronises. */
int gsb_tsdf_icp_eval(gsb_tsdf_t *e, int level, const float *approx_invM, int *n_valid, float *f, float *nabla6, float *hessian36);
int gsb_tsdf_set_tracking_frames(gsb_tsdf_t *e, int n);          /* ITMTrackingState::framesProcessed (weights switch on at >= 100) */
/* trackerResult of the last frame: 0 TRACKING_FAILED, 1 TRACKING_POOR, 2 TRACKING_GOOD; trackerScore; LM iterations run */
int gsb_tsdf_tracker_result(gsb_tsdf_t *e, int *result, float *score, int *iterations);
/* depth pyramid level after PrepareForEvaluation (level 0 = the float depth); dst_host may be NULL to query the size */
int gsb_tsdf_depth_level(gsb_tsdf_t *e, int level, float *dst_host, int *w, int *h);

/* ===================================================================================================
 * A.  Gaussian model  -- replaces RawGaussianModel / SLAMGaussianModel with render_method == "ges"
 *     (reference include/raw_gs_model.h:8-298, src/raw_gs_model.cpp:188-417, 654-705; slam/slam_gs_model.cpp:5-56)
 *     and, underneath, the gsplat autograd wrappers it calls (gsplat/gsplat_wapper.hpp:16-620):
 *     FullyFusedProjection, SphericalHarmonicsNew, isectTilesNoDepth, isectOffsetEncodeNoDepth,
 *     RasterizeToPixelsGes_NewParallel, plus 6 x torch::optim::Adam.
 *     Parameters are the reference's tensors: means [N,3], scales [N,3] (log), quats [N,4] (w,x,y,z), featuresDc [N,3],
 *     featuresRest [N,15,3], opacities [N,1] (logit); fp32, row-major, contiguous.
 *     Cameras: c2w = 16 floats ROW-major camera-to-world (cam.c2w_slam as torch stores it), pinhole fx, fy, cx, cy.
 * =================================================================================================== */
typedef struct gsb_gs gsb_gs_t;

typedef struct gsb_gs_config
{
    int width, height;
    int capacity;                 /* maximum number of Gaussians (buffers are allocated once)                          */
    int isect_capacity;           /* maximum Gaussian-tile intersections per render                                    */
    int item_capacity;            /* maximum backward work items per render (one per 2048 box pixels of a splat)       */
    int max_gs_radii;             /* MODEL.max_gs_radii (office0.yaml:83); <= 0 disables the clamp                     */
    float delta_depth;            /* MODEL.delta_depth                                                                 */
    float eps2d, near_plane, far_plane, radius_clip;   /* include/raw_gs_model.h rendering constants                   */
    float lr_means, lr_scales, lr_quats, lr_dc, lr_rest, lr_opac;   /* MODEL.*_lr (office0.yaml:93-100)                */
    float scene_scale;            /* multiplies lr_means (src/raw_gs_model.cpp:666)                                    */
    int device;
} gsb_gs_config_t;

void gsb_gs_default_config(gsb_gs_config_t *cfg);
int gsb_gs_create(const gsb_gs_config_t *cfg, gsb_gs_t **out);
void gsb_gs_destroy(gsb_gs_t *e);
int gsb_gs_set_stream(gsb_gs_t *e, void *cuda_stream);
int gsb_gs_sync(gsb_gs_t *e);

/* parameter I/O; pointers may be host or device memory. rest may be NULL (zeros). */
int gsb_gs_set_params(gsb_gs_t *e, int n, const float *means, const float *scales_log, const float *quats, const float *dc,
                      const float *rest, const float *opac_logit);
int gsb_gs_append(gsb_gs_t *e, int n, const float *means, const float *scales_log, const float *quats, const float *dc,
                  const float *rest, const float *opac_logit);            /* RawGaussianParams::add                     */
int gsb_gs_get_params(gsb_gs_t *e, int n, float *means, float *scales_log, float *quats, float *dc, float *rest, float *opac_logit);
int gsb_gs_count(gsb_gs_t *e, int *n);                                    /* getGaussianNum(); synchronises             */
int gsb_gs_count_upper(gsb_gs_t *e);                                      /* host-side bound, no synchronisation        */

int gsb_gs_init_optimizers(gsb_gs_t *e);                                  /* RawGaussianModel::initOptimizers           */
int gsb_gs_set_learning_rates(gsb_gs_t *e, float means, float scales, float quats, float dc, float rest, float opac);

/* RawGaussianModel::forward (gesForward), no grad: rgb [H,W,3], depth [H,W], alpha (weight sum) [H,W], device pointers.
 * ref_depth [H,W] and base_color [H,W,3] are the TSDF raycast maps (runRaycastByCam depth_map / color_map). */
int gsb_gs_render(gsb_gs_t *e, const float *c2w, float fx, float fy, float cx, float cy, const float *ref_depth_dev,
                  const float *base_color_dev, float *rgb_dev, float *depth_dev, float *alpha_dev);
/* one iteration of SLAMPipeline::localOptimize: forward + L1 loss vs gt_rgb [H,W,3] + backward + Adam step + zero grad */
int gsb_gs_train_step(gsb_gs_t *e, const float *c2w, float fx, float fy, float cx, float cy, const float *ref_depth_dev,
                      const float *base_color_dev, const float *gt_rgb_dev);
int gsb_gs_loss(gsb_gs_t *e, double *loss);                               /* loss["total"] of the last step; synchronises */
/* the same in two halves: _begin enqueues the reduction + the 8-byte copy to pinned memory, _end waits for that copy only */
int gsb_gs_loss_begin(gsb_gs_t *e);
int gsb_gs_loss_end(gsb_gs_t *e, double *loss);
/* SLAMPipeline::removeRedundantGs + prunePoints (remove_configs.low_opac_thres, small_scale_thres, large_scale_thres).
 * Like the reference's removeFromOptimizer (src/raw_gs_model.cpp:744-765) the surviving Gaussians keep their Adam moments, so
 * gsb_gs_train_step may follow directly; gsb_gs_init_optimizers before the prune makes it cheaper (no state to move). */
int gsb_gs_prune(gsb_gs_t *e, float min_opac, float min_scale, float max_scale);

/* SLAMPipeline::initNewGaussians + SLAMGaussianModel::addGaussians (slam/slam_pipeline.cpp:450-526, slam/slam_gs_model.cpp:5-56) */
typedef struct gsb_spawn_config
{
    float color_error_thres;      /* PIPE.color_error_thres (0.05)                        */
    float depth_vis_min, depth_vis_max, alpha_vis_max;   /* PIPE.vis_configs (0, 5, 5)    */
    float sample_ratio;           /* PIPE.new_gs_sample_ratio (0.25)                      */
    float max_init_scale, min_init_scale, default_opacity;   /* MODEL (0.01, -1, 0.5)     */
    unsigned seed;                /* sampling seed (the reference draws torch::randperm)  */
    int rank, world;              /* multi-GPU: spawn only Gaussians whose 4 cm block hashes to `rank` (world <= 1: all)      */
    const void *render_rgb_dev;   /* optional: current render [H,W,3] / weight sum [H,W] supplied by the caller (multi-GPU,   */
    const void *render_alpha_dev; /* after the all-reduce); NULL -> the engine renders the camera itself                      */
} gsb_spawn_config_t;
int gsb_gs_spawn(gsb_gs_t *e, const gsb_spawn_config_t *sc, const float *c2w, float fx, float fy, float cx, float cy,
                 const void *free_vertex_dev, float voxel_size, const float *depth_map_dev, const float *color_map_dev,
                 const float *gt_rgb_dev);
/* runRaycastByCam tensor glue (slam/slam_pipeline.cpp:386-403): GetFreeVertex()/GetFreeImage() -> depth_map, color_map, confidence */
int gsb_gs_raycast_maps(gsb_gs_t *e, const void *free_vertex_dev, const void *free_image_dev, const float *c2w, float voxel_size,
                        float *depth_map_dev, float *color_map_dev, float *conf_map_dev);
/* Camera::image / Camera::depth (float) from the raw RGBA8 / int16-mm frame */
int gsb_gs_frame_to_float(gsb_gs_t *e, const void *rgba_dev, const void *depth_mm_dev, float *rgb_dev, float *depth_dev);

/* Multi-GPU: the Gaussian set is sharded across ranks (one engine per GPU).  The GES blend is an order-independent sum, so every
 * rank rasterises its own Gaussians into partial sums acc5 = render_colors [H*W*4] followed by alphas [H*W]; the caller
 * all-reduces acc5 (ncclAllReduce sum over NVLink) and every rank finishes on the summed image; backward and Adam are local. */
int gsb_gs_forward_partial(gsb_gs_t *e, const float *c2w, float fx, float fy, float cx, float cy, const float *ref_depth_dev,
                           float *acc5_dev, int for_backward);
int gsb_gs_render_finish(gsb_gs_t *e, const float *ref_depth_dev, const float *base_color_dev, const float *acc5_dev, float *rgb_dev,
                         float *depth_dev, float *alpha_dev);
int gsb_gs_train_finish(gsb_gs_t *e, const float *ref_depth_dev, const float *base_color_dev, const float *gt_rgb_dev,
                        const float *acc5_dev);

/* Multi-GPU, exchange below the C ABI (no NCCL call between two entry points): a communicator owns this rank's exchange segment in
 * device memory and maps every peer's (CUDA IPC over NVLink / NVSwitch when the ranks are processes; plain pointers inside one process).
 * With a communicator attached to an engine, gsb_gs_render / gsb_gs_train_step / gsb_gs_spawn do the whole multi-GPU iteration
 * themselves: the rasteriser stores each tile's partial sums into the gather slot of the tile's owner rank, the owner composites and
 * stores dL/d(render) into every rank's gradient image; two flag barriers through peer memory per exchange; backward and Adam local.
 * Every rank must issue the same sequence of these calls.  Bring-up (what a host does once, e.g. over MPI / torch.distributed):
 *     gsb_comm_create(device, rank, world, W, H, &c);  gsb_comm_export(c, handle64);   all-gather the 64-byte handles;
 *     gsb_comm_attach(c, handles);   (handles: world x 64 bytes, rank order)  gsb_gs_set_comm(engine, c);
 * gsb_comm_attach_local takes the peers' communicators directly (several engines in one process). */
typedef struct gsb_comm gsb_comm_t;
int gsb_comm_create(int device, int rank, int world, int width, int height, gsb_comm_t **out);
int gsb_comm_export(gsb_comm_t *c, void *handle64);
int gsb_comm_attach(gsb_comm_t *c, const void *handles);
int gsb_comm_attach_local(gsb_comm_t *c, gsb_comm_t *const *peers);
int gsb_comm_barrier(gsb_comm_t *c, void *cuda_stream);       /* one flag barrier on the given stream (tests, host-level fences) */
int gsb_comm_error(gsb_comm_t *c);                             /* non-zero once a barrier gave up waiting for a peer (~20 s)      */
void gsb_comm_destroy(gsb_comm_t *c);
int gsb_gs_set_comm(gsb_gs_t *e, gsb_comm_t *comm);            /* NULL detaches                                                   */

/* ---- Staged entry points: one per gsplat::*_tensor function that RawGaussianModel::gesForward reaches through the autograd
 * wrappers of gsplat/gsplat_wapper.hpp, so that each wrapper can be re-pointed at this library on its own.
 * One camera (C = 1: the reference unsqueezes a camera dimension of size 1 everywhere, src/raw_gs_model.cpp:187); every pointer
 * is DEVICE memory, fp32 / int32 contiguous, outputs are caller-allocated (the shim allocates them as torch tensors); calls are
 * asynchronous on the engine's stream except gsb_gs_isect_tiles.  n <= gsb_gs_config.capacity.  viewmat row-major [4,4], K [3,3].
 *   gsb_gs_projection_fwd     gsplat::fully_fused_projection_fwd_tensor (rasterizer/fully_fused_projection_fwd.cu:196-273; pinhole,
 *                             quats+scales form, eps2d/near/far/radius_clip from the engine config); clamp_radii > 0 additionally applies
 *                             torch::clamp_max(radii, max_gs_radii) (src/raw_gs_model.cpp:241-242).  Culled Gaussians: radii 0, rest 0.
 *   gsb_gs_projection_bwd     gsplat::fully_fused_projection_bwd_tensor (fully_fused_projection_bwd.cu:288-403): v_means [n,3],
 *                             v_quats [n,4], v_scales [n,3] (w.r.t. the real scales), written (not accumulated).
 *   gsb_gs_sh_fwd / _bwd      gsplat::compute_sh_fwd_tensor / compute_sh_bwd_tensor (compute_sh_fwd.cu:40-72, compute_sh_bwd.cu:56-123):
 *                             dirs [n,3] (not normalised), coeffs [n,16,3], masks uint8 [n] or NULL -> colors [n,3];
 *                             v_coeffs [n,16,3], v_dirs [n,3] or NULL.  degrees_to_use must be 3.
 *   gsb_gs_isect_tiles        gsplat::isect_tiles_tensor_no_depth + isect_offset_encode_tensor_no_depth (isect_tiles_no_depth.cu:132-461):
 *                             bins the splats, returns n_isects (synchronises, as the reference does to size its tensors);
 *                             tiles_per_gauss [n] optional.  gsb_gs_isect_fetch then copies isect_ids int64 [n_isects] (= tile index),
 *                             flatten_ids int32 [n_isects] and tile_offsets int32 [tile_h*tile_w]; any of the three may be NULL.
 *                             The reference's group_gs_ids / group_starts tables (work list of ITS backward) have no counterpart:
 *                             the backward here derives its own work items.
 *   gsb_gs_rasterize_ges_fwd  gsplat::rasterize_to_pixels_fwd_ges_tensor (rasterize_to_pixels_fwd_ges.cu:338-407): colors4 [n,4] = rgb +
 *                             camera depth, opacities [n], ref_depth [H,W] ALREADY clamped by the caller -> render4 [H,W,4], alphas [H,W].
 *   gsb_gs_rasterize_ges_bwd  gsplat::rasterize_to_pixels_bwd_ges_gs_parallel_tensor (rasterize_to_pixels_bwd_ges_new_parallel.cu:304-385):
 *                             v_render4 [H,W,4], v_alphas [H,W] -> v_means2d [n,2], v_conics [n,3], v_colors4 [n,4], v_opacities [n].
 *   gsb_gs_adam_step          torch::optim::Adam::step for one parameter tensor of n floats (no weight decay / amsgrad;
 *                             src/raw_gs_model.cpp:661-672), step = 1-based count after increment. */
int gsb_gs_projection_fwd(gsb_gs_t *e, int n, const float *means, const float *quats, const float *scales, const float *viewmat,
                          const float *K, int clamp_radii, int *radii, float *means2d, float *depths, float *conics);
int gsb_gs_projection_bwd(gsb_gs_t *e, int n, const float *means, const float *quats, const float *scales, const float *viewmat,
                          const float *K, const int *radii, const float *conics, const float *v_means2d, const float *v_depths,
                          const float *v_conics, float *v_means, float *v_quats, float *v_scales);
int gsb_gs_sh_fwd(gsb_gs_t *e, int n, int degrees_to_use, const float *dirs, const float *coeffs, const unsigned char *masks, float *colors);
int gsb_gs_sh_bwd(gsb_gs_t *e, int n, int degrees_to_use, const float *dirs, const float *coeffs, const unsigned char *masks,
                  const float *v_colors, float *v_coeffs, float *v_dirs);
int gsb_gs_isect_tiles(gsb_gs_t *e, int n, const float *means2d, const int *radii, int *tiles_per_gauss, int *n_isects);
int gsb_gs_isect_fetch(gsb_gs_t *e, int n_isects, long long *isect_ids, int *flatten_ids, int *tile_offsets);
int gsb_gs_rasterize_ges_fwd(gsb_gs_t *e, int n, const float *means2d, const float *conics, const float *colors4, const float *opacities,
                             const float *ref_depth, float delta_depth, const int *tile_offsets, const int *flatten_ids, int n_isects,
                             float *render4, float *alphas);
int gsb_gs_rasterize_ges_bwd(gsb_gs_t *e, int n, const float *means2d, const float *conics, const float *colors4, const float *opacities,
                             const int *radii, const float *ref_depth, float delta_depth, const float *v_render4, const float *v_alphas,
                             float *v_means2d, float *v_conics, float *v_colors4, float *v_opacities);
int gsb_gs_adam_step(gsb_gs_t *e, long long n, float *param, const float *grad, float *exp_avg, float *exp_avg_sq, float lr, float beta1,
                     float beta2, float eps, int step);

/* ---- render_method "raw" (RawGaussianModel::rawForward, src/raw_gs_model.cpp:37-186) and the fused SSIM term of computeLoss
 * (src/raw_gs_model.cpp:383-395); same conventions as the staged entry points above.
 *   gsb_gs_isect_tiles_depth  gsplat::isect_tiles_tensor + isect_offset_encode_tensor (rasterizer/isect_tiles.cu:132-430): bins ordered by
 *                             (tile, camera depth); ties keep ascending Gaussian id like the reference's stable radix sort.
 *                             gsb_gs_isect_fetch_depth: isect_ids int64 = tile << 32 | depth bits, flatten_ids, tile_offsets.
 *   gsb_gs_rasterize_fwd      gsplat::rasterize_to_pixels_fwd_tensor (rasterize_to_pixels_fwd.cu:198-376), COLOR_DIM 4, optional
 *                             background [4] (device): front-to-back alpha compositing -> render4 [H,W,4], alphas [H,W] (= 1 - T),
 *                             last_ids [H,W] (bin index of the last splat that contributed).
 *   gsb_gs_rasterize_bwd      gsplat::rasterize_to_pixels_bwd_tensor (rasterize_to_pixels_bwd.cu:289-511), absgrad = false.
 *   gsb_gs_ssim_fwd / _bwd    fusedssim / fusedssim_backward (rasterizer/ssim.cu:368-460): planes = B * CH images of height x width,
 *                             NCHW; the three derivative maps are NULL when train = false. */
int gsb_gs_isect_tiles_depth(gsb_gs_t *e, int n, const float *means2d, const int *radii, const float *depths, int *tiles_per_gauss,
                             int *n_isects);
int gsb_gs_isect_fetch_depth(gsb_gs_t *e, int n_isects, long long *isect_ids, int *flatten_ids, int *tile_offsets);
int gsb_gs_rasterize_fwd(gsb_gs_t *e, int n, const float *means2d, const float *conics, const float *colors4, const float *opacities,
                         const float *background4, const int *tile_offsets, const int *flatten_ids, int n_isects, float *render4,
                         float *alphas, int *last_ids);
int gsb_gs_rasterize_bwd(gsb_gs_t *e, int n, const float *means2d, const float *conics, const float *colors4, const float *opacities,
                         const float *background4, const int *tile_offsets, const int *flatten_ids, int n_isects, const float *render_alphas,
                         const int *last_ids, const float *v_render4, const float *v_alphas, float *v_means2d, float *v_conics,
                         float *v_colors4, float *v_opacities);
int gsb_gs_ssim_fwd(gsb_gs_t *e, int planes, int height, int width, float C1, float C2, const float *img1, const float *img2,
                    float *ssim_map, float *dm_dmu1, float *dm_dsigma1_sq, float *dm_dsigma12);
int gsb_gs_ssim_bwd(gsb_gs_t *e, int planes, int height, int width, const float *img1, const float *img2, const float *dL_dmap,
                    const float *dm_dmu1, const float *dm_dsigma1_sq, const float *dm_dsigma12, float *dL_dimg1);

/* distCUDA2 (gsplat/rasterizer/simple_knn.cu:227-239; RawGaussianParams::init, src/raw_gs_param.cpp:28): points_dev [n,3] ->
 * mean_dist2_dev [n] = mean of the squared distances to the 3 nearest other points (FLT_MAX terms when n < 4, as the reference).
 * Exact (uniform grid + expanding rings instead of the reference's Morton boxes), asynchronous, no host round trip; the first call
 * allocates its workspace (sized for max(n, capacity) points). */
int gsb_gs_dist_cuda2(gsb_gs_t *e, int n, const float *points_dev, float *mean_dist2_dev);

/* state read-back for parity tests (synchronises) */
enum
{
    GSB_GS_SPLAT_RECORDS = 0,   /* [N] 12 floats: mean2d.xy, opacity, radius(int) | conic abc, depth | rgb, flags(int)   */
    GSB_GS_SPLAT_GRADS = 1,     /* [N] 12 floats: v_mean2d.xy, v_opacity, v_depth | v_conic abc, 0 | v_rgb, 0            */
    GSB_GS_TILE_OFFSETS = 2,    /* int[T+1]  (isect_offsets + n_isects)                                                  */
    GSB_GS_FLATTEN_IDS = 3,     /* int[n_isects]                                                                         */
    GSB_GS_V_OUT = 4,           /* 8 floats per pixel [H*W]: dL/d render rgb, dL/d alpha | depth cut, 3 x padding        */
    GSB_GS_COUNTERS = 5,        /* int[16]: n_isects, n_items, overflow bits, -, n_visible, ...; [8..11] pair statistics of stage 6 */
    GSB_GS_GRAD_MEANS = 6, GSB_GS_GRAD_SCALES = 7, GSB_GS_GRAD_QUATS = 8, GSB_GS_GRAD_DC = 9, GSB_GS_GRAD_REST = 10, GSB_GS_GRAD_OPAC = 11,
    GSB_GS_SPAWN_PIXELS = 12    /* int[<= H*W]: pixel index of every Gaussian the last gsb_gs_spawn appended, in append order      */
};
int gsb_gs_read(gsb_gs_t *e, int what, void *dst_host, size_t bytes);
/* single stages on the camera / images of the last train step, for per-kernel timing: 0 projection+SH, 1 tile binning,
 * (1 re-runs 0), 2 rasteriser forward (train), 3 rasteriser backward, 4 drop the backward work list (call last),
 * 5 parameter backward + Adam with a zero step size (consumes the work list), 6 rasteriser backward with its pair counters compiled in
 * ((pixel, splat) pairs evaluated / passed -> GSB_GS_COUNTERS[8..11], two 64-bit values) */
int gsb_gs_run_stage(gsb_gs_t *e, int stage);
int gsb_gs_enable_grad_dump(gsb_gs_t *e, int on);   /* keep the parameter gradients of each train step for GSB_GS_GRAD_* */

/* ===================================================================================================
 * Peer mailbox (new; multi-GPU host plumbing below the C ABI).  A byte segment per rank that every other rank of the box can write into
 * over NVLink, and 256 counters per rank for hand-shakes on streams.  The functional split of the SLAM loop uses it: the rank that owns
 * the TSDF side stores the camera maps of a cycle into the Gaussian ranks' mailboxes (gps_slam_b200/split.py).  Mapping as for gsb_comm_*.
 * =================================================================================================== */
typedef struct gsb_mbox gsb_mbox_t;
int gsb_mbox_create(int device, int rank, int world, size_t bytes, gsb_mbox_t **out);
int gsb_mbox_export(gsb_mbox_t *m, void *handle64);
int gsb_mbox_attach(gsb_mbox_t *m, const void *handles /* world x 64 bytes */);
int gsb_mbox_attach_local(gsb_mbox_t *m, gsb_mbox_t *const *peers);
void *gsb_mbox_local(gsb_mbox_t *m);   /* device pointer of this rank's own mailbox */
/* enqueue on `stream`: copy bytes from src_dev (this rank) to offset dst_offset of rank dst_rank's mailbox */
int gsb_mbox_put(gsb_mbox_t *m, int dst_rank, size_t dst_offset, const void *src_dev, size_t bytes, void *stream);
/* enqueue on `stream`: after everything queued before, counter `flag` of rank dst_rank becomes `value` (release at system scope) */
int gsb_mbox_signal(gsb_mbox_t *m, int dst_rank, int flag, unsigned value, void *stream);
/* enqueue on `stream`: the stream proceeds when this rank's counter `flag` has reached `value` (bounded spin -> gsb_mbox_error) */
int gsb_mbox_wait(gsb_mbox_t *m, int flag, unsigned value, void *stream);
int gsb_mbox_error(gsb_mbox_t *m);
void gsb_mbox_destroy(gsb_mbox_t *m);

#ifdef __cplusplus
}
#endif
#endif /* GPSSLAM_B200_H */
