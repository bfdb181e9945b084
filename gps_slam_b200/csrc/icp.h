// Internal interface of the ICP tracker; implemented in icp_kernels.cu.
#pragma once
#include "common.cuh"
#include "se3.h"

namespace icp
{
struct Tracker;

// kind: 1 = extended (ITMExtendedTracker, the reference's compiled-in default), 2 = icp (ITMDepthTracker)
Tracker *create_tracker(int kind, int W, int H, float vfmin, float vfmax);
void destroy_tracker(Tracker *t);

// B1 stand-alone: short mm -> float m (ITMViewBuilder_Shared.h:27-36)
void convert_depth(const short *depth_mm, float *depth_f, int W, int H, cudaStream_t st);

// TrackCamera: updates *pose_d in place (host pose, final value read back once per frame)
int track_camera(Tracker *t, const float *depth_f, const float4 *pointsMap, const float4 *normalsMap, float fx, float fy, float cx, float cy,
                 const Mat4 &scenePose, int trackingFrames, se3::Pose *pose_d, cudaStream_t st);
// ICP sharded over `world` GPUs (row split + in-kernel all-reduce of the 29 sums through peer memory): xchg[q] = rank q's exchange block
// ([2][16][32] floats, zero-initialised, peer-mapped), err = host-mapped flag raised when a peer does not deliver
void set_shard(Tracker *t, int rank, int world, float *const *xchg, int *err);
// single evaluation at one pyramid level for a given camera->world estimate (ComputeGandH_Depth / ComputeGandH), for parity tests
int icp_eval(Tracker *t, const float *depth_f, const float4 *pointsMap, const float4 *normalsMap, float fx, float fy, float cx, float cy,
             const Mat4 &scenePose, int trackingFrames, int level, const Mat4 &approxInvPose, int *nValid, float *f, float *nabla6, float *hessian36,
             cudaStream_t st);
// device pointer of pyramid level >= 1 (nullptr for level 0 = the caller's depth) and its size
const float *level_depth(Tracker *t, int level, int *w, int *h);
// trackerResult (0 failed, 1 poor, 2 good), trackerScore and LM iterations of the last TrackCamera
void tracker_result(Tracker *t, int *result, float *score, int *iterations);
} // namespace icp
