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
    int max_w;                    /* ITMSceneParams::maxW, reference default 100 (ITMLibSettings.cpp:10)      */
    int num_blocks;               /* SDF_LOCAL_BLOCK_NUM, reference 0x40000; 0 = default                      */
    int tracker;                  /* 0 = ground-truth poses (turnOffTracking), 1 = extended, 2 = icp           */
    int device;                   /* CUDA device ordinal                                                       */
    int integrate_variant;        /* 0 = TMA-pipelined (default), 1 = direct LDG/STG                           */
} gsb_tsdf_config_t;

void gsb_tsdf_default_config(gsb_tsdf_config_t *cfg);
int gsb_tsdf_create(const gsb_tsdf_config_t *cfg, gsb_tsdf_t **out);
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
const void *gsb_tsdf_live_vertex_dev(gsb_tsdf_t *e);  /* GetLiveVertex()                                             */
const void *gsb_tsdf_points_map_dev(gsb_tsdf_t *e);   /* trackingState->pointCloud->locations (metres, w=conf+1/-1)  */
const void *gsb_tsdf_normals_map_dev(gsb_tsdf_t *e);  /* trackingState->pointCloud->colours                          */

/* GetTrackingState()->pose_d: M = GetM() (world->camera), invM = GetInvM() */
int gsb_tsdf_get_pose(gsb_tsdf_t *e, float *M, float *invM);
float gsb_tsdf_voxel_size(gsb_tsdf_t *e);
int gsb_tsdf_frames_processed(gsb_tsdf_t *e);

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
/* which: 0 lastFreeBlockId, 1 lastFreeExcessListId, 2 noVisibleEntries, 3 error flag (synchronises) */
int gsb_tsdf_counter(gsb_tsdf_t *e, int which, int *value);

/* Single stages on the current frame / pose, for profiling and stage-level parity.
 * stage: 0 allocate (B1-B4), 1 integrate (B5), 2 expected depth (B6), 3 raycast (B7), 4 ICP maps (B8) */
int gsb_tsdf_run_stage(gsb_tsdf_t *e, int stage);

#ifdef __cplusplus
}
#endif
#endif /* GPSSLAM_B200_H */
