// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// C-ABI driver around the REFERENCE's own InfiniTAM CPU engine
// (ITMBasicEngine<ITMVoxel_s_rgb, ITMVoxelBlockHash>, deviceType = DEVICE_CPU), compiled from the
// sources where they lie under /root/reference/InfiniTAM by oracle/itm_ref/Makefile into
// oracle/_ref/libitm_ref*.so.  It lets tests/ and bench.py's cpu_baseline / --impl reference leg
// run the reference implementation of SURVEY.md section 8 rows B1-B9 and C1-C5 and read back its
// internal state (hash table, voxel block array, visible list, min/max image, raycast, ICP maps,
// per-evaluation ICP accumulators, tracked pose).
//
// Nothing in here is reference source: it only *calls* the reference classes.  The
// private/protected -> public macro trick below exposes engine internals to this translation unit
// only (access specifiers do not change the class layout under the Itanium ABI used by gcc).

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#define private public
#define protected public
#include "ITMLib/ITMLibDefines.h"
#include "ITMLib/Core/ITMBasicEngine.h"
#include "ITMLib/Objects/RenderStates/ITMRenderState_VH.h"
#include "ITMLib/Trackers/Interface/ITMExtendedTracker.h"
#include "ITMLib/Trackers/Interface/ITMDepthTracker.h"
#include "ITMLib/Trackers/CPU/ITMExtendedTracker_CPU.h"
#include "ITMLib/Trackers/CPU/ITMDepthTracker_CPU.h"
#include "ITMLib/Engines/Meshing/CPU/ITMMeshingEngine_CPU.h"
#include "ITMLib/Engines/Meshing/Shared/ITMMeshingEngine_Shared.h"
#include "ITMLib/Objects/Meshing/ITMMesh.h"
#undef private
#undef protected

using namespace ITMLib;
typedef ITMBasicEngine<ITMVoxel, ITMVoxelIndex> Engine;

struct ItmRef
{
	ITMLibSettings *settings;
	Engine *engine;
	ITMUChar4Image *rgb;
	ITMShortImage *depth;
	int w, h;
	std::vector<ORUtils::Matrix4<float> *> poses;
	ORUtils::SE3Pose freePose;
	ITMIntrinsics freeIntr;
};

extern "C"
{

// tracker_kind: 0 = ground-truth poses (tracking off), 1 = "extended" (the reference's compiled-in
// trackerConfig, Utils/ITMLibSettings.cpp:54-57), 2 = "icp" (ITMDepthTracker, commented-out default)
void *itmref_create(int w, int h, float fx, float fy, float cx, float cy,
					float voxelSize, float mu, float vfmin, float vfmax, int tracker_kind, int omp_threads)
{
#ifdef _OPENMP
	if (omp_threads > 0)
		omp_set_num_threads(omp_threads);
#endif
	ItmRef *r = new ItmRef();
	r->w = w;
	r->h = h;
	ITMRGBDCalib calib;
	calib.intrinsics_rgb.SetFrom(w, h, fx, fy, cx, cy);
	calib.intrinsics_d = calib.intrinsics_rgb;
	calib.disparityCalib.SetStandard();

	r->settings = new ITMLibSettings();
	r->settings->deviceType = ITMLibSettings::DEVICE_CPU;
#ifndef COMPILE_WITHOUT_CUDA
	// oracle/itm_ref_cuda builds this same driver with the reference's CUDA engine compiled in: ITMREF_DEVICE=cuda runs it (timing only --
	// the read-back accessors below return host pointers, which the CUDA engine does not keep up to date)
	if (const char *d = getenv("ITMREF_DEVICE"))
		if (!strcmp(d, "cuda"))
			r->settings->deviceType = ITMLibSettings::DEVICE_CUDA;
#endif
	r->settings->createMeshingEngine = false;
	r->settings->sceneParams.voxelSize = voxelSize;
	r->settings->sceneParams.mu = mu;
	r->settings->sceneParams.viewFrustum_min = vfmin;
	r->settings->sceneParams.viewFrustum_max = vfmax;
	if (tracker_kind == 2)
		r->settings->trackerConfig = "type=icp,levels=rrrbb,minstep=1e-3,outlierC=0.01,outlierF=0.002,numiterC=10,numiterF=2,failureDec=5.0";

	r->engine = new Engine(r->settings, calib, Vector2i(w, h), Vector2i(w, h));
	if (tracker_kind == 0)
		r->engine->turnOffTracking();
	r->rgb = new ITMUChar4Image(Vector2i(w, h), true, false);
	r->depth = new ITMShortImage(Vector2i(w, h), true, false);
	return r;
}

void itmref_destroy(void *h)
{
	ItmRef *r = (ItmRef *)h;
	delete r->engine;
	delete r->rgb;
	delete r->depth;
	for (auto *p : r->poses)
		delete p;
	delete r->settings;
	delete r;
}

// rgba: w*h*4 bytes, depth_mm: w*h int16, c2w: 16 floats in ORUtils column-major order (m[col*4+row]),
// may be NULL when tracking is on.
int itmref_process_frame(void *h, const uint8_t *rgba, const int16_t *depth_mm, const float *c2w)
{
	ItmRef *r = (ItmRef *)h;
	memcpy(r->rgb->GetData(MEMORYDEVICE_CPU), rgba, (size_t)r->w * r->h * 4);
	memcpy(r->depth->GetData(MEMORYDEVICE_CPU), depth_mm, (size_t)r->w * r->h * 2);
	if (c2w)
	{
		ORUtils::Matrix4<float> *m = new ORUtils::Matrix4<float>();
		memcpy(m->m, c2w, 64);
		r->poses.push_back(m);
		// ProcessFrame indexes gtC2wPoses[framesProcessed] (Core/ITMBasicEngine.tpp:278)
		r->engine->gtC2wPoses.resize(r->engine->framesProcessed + 1);
		r->engine->gtC2wPoses[r->engine->framesProcessed] = m;
	}
	return (int)r->engine->ProcessFrame(r->rgb, r->depth, NULL);
}

int itmref_num_hash_entries(void *) { return ITMVoxelBlockHash::noTotalEntries; }
int itmref_num_blocks(void *) { return SDF_LOCAL_BLOCK_NUM; }
const void *itmref_hash_entries(void *h) { return ((ItmRef *)h)->engine->scene->index.GetEntries(); }
const void *itmref_voxels(void *h) { return ((ItmRef *)h)->engine->scene->localVBA.GetVoxelBlocks(); }
int itmref_last_free_block(void *h) { return ((ItmRef *)h)->engine->scene->localVBA.lastFreeBlockId; }
int itmref_last_free_excess(void *h) { return ((ItmRef *)h)->engine->scene->index.GetLastFreeExcessListId(); }

const int *itmref_visible_ids(void *h, int live, int *n)
{
	ItmRef *r = (ItmRef *)h;
	ITMRenderState_VH *rs = (ITMRenderState_VH *)(live ? r->engine->renderState_live : r->engine->renderState_freeview);
	*n = rs->noVisibleEntries;
	return rs->GetVisibleEntryIDs();
}
const uint8_t *itmref_visible_types(void *h)
{
	return ((ITMRenderState_VH *)((ItmRef *)h)->engine->renderState_live)->GetEntriesVisibleType();
}
const float *itmref_depth(void *h) { return ((ItmRef *)h)->engine->view->depth->GetData(MEMORYDEVICE_CPU); }
const float *itmref_minmax(void *h, int live)
{
	ItmRef *r = (ItmRef *)h;
	ITMRenderState *rs = live ? r->engine->renderState_live : r->engine->renderState_freeview;
	return (const float *)rs->renderingRangeImage->GetData(MEMORYDEVICE_CPU);
}
const float *itmref_raycast(void *h, int live)
{
	ItmRef *r = (ItmRef *)h;
	ITMRenderState *rs = live ? r->engine->renderState_live : r->engine->renderState_freeview;
	return (const float *)rs->raycastResult->GetData(MEMORYDEVICE_CPU);
}
const uint8_t *itmref_raycast_image(void *h, int live)
{
	ItmRef *r = (ItmRef *)h;
	ITMRenderState *rs = live ? r->engine->renderState_live : r->engine->renderState_freeview;
	return (const uint8_t *)rs->raycastImage->GetData(MEMORYDEVICE_CPU);
}
const float *itmref_points_map(void *h) { return (const float *)((ItmRef *)h)->engine->trackingState->pointCloud->locations->GetData(MEMORYDEVICE_CPU); }
const float *itmref_normals_map(void *h) { return (const float *)((ItmRef *)h)->engine->trackingState->pointCloud->colours->GetData(MEMORYDEVICE_CPU); }

// M = world->camera, invM = camera->world, both ORUtils column-major
void itmref_pose(void *h, float *M16, float *invM16)
{
	ItmRef *r = (ItmRef *)h;
	const ORUtils::SE3Pose *p = r->engine->trackingState->pose_d;
	memcpy(M16, p->GetM().m, 64);
	ORUtils::Matrix4<float> inv = p->GetInvM();
	memcpy(invM16, inv.m, 64);
}
void itmref_set_pose_invM(void *h, const float *invM16)
{
	ItmRef *r = (ItmRef *)h;
	ORUtils::Matrix4<float> m;
	memcpy(m.m, invM16, 64);
	r->engine->trackingState->pose_d->SetInvM(m);
	r->engine->trackingState->pose_d->Coerce();
}
int itmref_tracker_result(void *h) { return (int)((ItmRef *)h)->engine->trackingState->trackerResult; }
int itmref_frames_processed(void *h) { return ((ItmRef *)h)->engine->framesProcessed; }

// Free-view raycast exactly as slam_pipeline.cpp:362-415 drives it: runRaycast(pose, intrinsics)
// (Core/ITMBasicEngine.tpp:500-526).  c2w column-major; pose built with SetInvM + Coerce like
// tensorToInfiMatrix4 + SE3Pose use in slam/slam_pipeline.cpp.
void itmref_run_raycast(void *h, const float *c2w, float fx, float fy, float cx, float cy)
{
	ItmRef *r = (ItmRef *)h;
	ORUtils::Matrix4<float> m;
	memcpy(m.m, c2w, 64);
	r->freePose.SetInvM(m);
	r->freeIntr.SetFrom(r->w, r->h, fx, fy, cx, cy);
	r->engine->runRaycast(&r->freePose, &r->freeIntr);
}
void itmref_free_pose(void *h, float *M16, float *invM16)
{
	ItmRef *r = (ItmRef *)h;
	memcpy(M16, r->freePose.GetM().m, 64);
	ORUtils::Matrix4<float> inv = r->freePose.GetInvM();
	memcpy(invM16, inv.m, 64);
}

// One ICP evaluation through the reference tracker (Trackers/CPU/ITM{Extended,Depth}Tracker_CPU.cpp
// ComputeGandH_Depth / ComputeGandH) at pyramid level `level`, for camera->world estimate
// approxInvPose (column-major), against the current raycast maps.  Returns the number of valid
// points; f, nabla[6], hessian[36] as the reference returns them (Extended: un-normalised sums).
int itmref_icp_eval(void *h, int level, const float *approxInvPose16, float *f, float *nabla6, float *hessian36)
{
	ItmRef *r = (ItmRef *)h;
	ORUtils::Matrix4<float> inv;
	memcpy(inv.m, approxInvPose16, 64);
	for (int i = 0; i < 36; i++)
		hessian36[i] = 0;
	for (int i = 0; i < 6; i++)
		nabla6[i] = 0;
	*f = 0;
	ITMTracker *t = r->engine->tracker;
	if (ITMExtendedTracker_CPU *ex = dynamic_cast<ITMExtendedTracker_CPU *>(t))
	{
		ex->SetEvaluationData(r->engine->trackingState, r->engine->view);
		ex->PrepareForEvaluation();
		ex->SetEvaluationParams(level);
		return ex->ComputeGandH_Depth(*f, nabla6, hessian36, inv);
	}
	if (ITMDepthTracker_CPU *dt = dynamic_cast<ITMDepthTracker_CPU *>(t))
	{
		dt->SetEvaluationData(r->engine->trackingState, r->engine->view);
		dt->PrepareForEvaluation();
		dt->SetEvaluationParams(level);
		return dt->ComputeGandH(*f, nabla6, hessian36, inv);
	}
	return -1;
}
// override the tracker's notion of "frames tracked so far" (switches useWeights on at >= 100)
void itmref_set_tracking_frames(void *h, int n) { ((ItmRef *)h)->engine->trackingState->framesProcessed = n; }

// level-l depth pyramid image after PrepareForEvaluation (for C1 parity)
const float *itmref_depth_level(void *h, int level, int *w, int *hh)
{
	ItmRef *r = (ItmRef *)h;
	ITMTracker *t = r->engine->tracker;
	ITMFloatImage *img = NULL;
	if (ITMExtendedTracker_CPU *ex = dynamic_cast<ITMExtendedTracker_CPU *>(t))
		img = ex->viewHierarchy_Depth->GetLevel(level)->depth;
	else if (ITMDepthTracker_CPU *dt = dynamic_cast<ITMDepthTracker_CPU *>(t))
		img = dt->viewHierarchy->GetLevel(level)->data;
	if (!img)
		return NULL;
	*w = img->noDims.x;
	*hh = img->noDims.y;
	return img->GetData(MEMORYDEVICE_CPU);
}

// ITMBasicEngine::SaveToFile / LoadFromFile (Core/ITMBasicEngine.tpp:119-171); dir must end with '/'.  Returns 0 on success.
int itmref_save(void *h, const char *dir)
{
	try
	{
		((ItmRef *)h)->engine->SaveToFile(std::string(dir));
		return 0;
	}
	catch (std::exception &e)
	{
		fprintf(stderr, "itmref_save: %s\n", e.what());
		return 1;
	}
}

int itmref_load(void *h, const char *dir)
{
	try
	{
		((ItmRef *)h)->engine->LoadFromFile(std::string(dir));
		return 0;
	}
	catch (std::exception &e)
	{
		fprintf(stderr, "itmref_load: %s\n", e.what());
		return 1;
	}
}


// SaveSceneToMesh (Core/ITMBasicEngine.tpp:105-117) without the file: the reference's own ITMMeshingEngine_CPU::MeshScene gives the triangle
// POSITIONS in hash-entry order; it leaves the per-vertex colours unset (only its CUDA twin fills c0..c2, Engines/Meshing/CUDA/
// ITMMeshingEngine_CUDA.tcu:118-137), so the colours come from a second walk in the same order that calls the reference's buildVertList and
// takes colorList[] exactly like the CUDA kernel does.  out: [max_tri][18] floats = p0 p1 p2 c0 c1 c2; returns the triangle count.
int itmref_mesh(void *h, float *out, int max_tri)
{
	ItmRef *r = (ItmRef *)h;
	ITMScene<ITMVoxel, ITMVoxelIndex> *scene = r->engine->scene;
	ITMMesh mesh(MEMORYDEVICE_CPU, (uint)max_tri);
	ITMMeshingEngine_CPU<ITMVoxel, ITMVoxelIndex> eng;
	eng.MeshScene(&mesh, scene);
	const int n = (int)mesh.noTotalTriangles;
	const ITMMesh::Triangle *tri = mesh.triangles->GetData(MEMORYDEVICE_CPU);
	const ITMVoxel *localVBA = scene->localVBA.GetVoxelBlocks();
	const ITMHashEntry *hashTable = scene->index.GetEntries();
	int k = 0;
	for (int entryId = 0; entryId < scene->index.noTotalEntries && k < n; entryId++)
	{
		const ITMHashEntry &e = hashTable[entryId];
		if (e.ptr < 0)
			continue;
		Vector3i globalPos = e.pos.toInt() * SDF_BLOCK_SIZE;
		for (int z = 0; z < SDF_BLOCK_SIZE; z++)
			for (int y = 0; y < SDF_BLOCK_SIZE; y++)
				for (int x = 0; x < SDF_BLOCK_SIZE; x++)
				{
					Vector3f vertList[12], colorList[12];
					int cubeIndex = buildVertList(vertList, colorList, globalPos, Vector3i(x, y, z), localVBA, hashTable);
					if (cubeIndex < 0)
						continue;
					for (int i = 0; triangleTable[cubeIndex][i] != -1 && k < n; i += 3, k++)
					{
						float *o = out + (size_t)k * 18;
						const Vector3f *p[3] = {&tri[k].p0, &tri[k].p1, &tri[k].p2};
						for (int v = 0; v < 3; v++)
						{
							o[v * 3 + 0] = p[v]->x, o[v * 3 + 1] = p[v]->y, o[v * 3 + 2] = p[v]->z;
							const Vector3f &c = colorList[triangleTable[cubeIndex][i + v]];
							o[9 + v * 3 + 0] = c.x, o[9 + v * 3 + 1] = c.y, o[9 + v * 3 + 2] = c.z;
						}
					}
				}
	}
	return n;
}
} // extern "C"
