// Test driver for the InfiniTAM-facing C++ facade (gps_slam_b200/cxx/InfiniTAM): does, against synthetic frames from a file, what the
// reference's host code does with the engine -- createTsdfEngine (slam/InfiniTAM_tools.cpp:3-67 there), the CLIEngine frame loop
// (slam/TsdfFusion/CLIEngine.cpp:34-58), the pose read-out of SLAMTrainCams (slam/slam_pipeline.cpp:77-82) and the free-view raycasts of
// runRaycastByCam (slam/slam_pipeline.cpp:362-383) -- and dumps what it sees, for tests/test_cxx_itm_gpu.py to compare bit for bit
// with the ctypes route into the same library and with the golden vectors produced by the reference's CPU engine.
//
// With -DGSB_WITH_REFERENCE_CLIENGINE the frame loop is the reference's own CLIEngine, compiled unchanged from
// /root/reference/slam/TsdfFusion/CLIEngine.cpp against the facade headers (recipe: oracle/itm_ref/Makefile, binary in oracle/_ref/).
//
//   input  : int32 n, w, h, tracker(0 gt | 1 extended | 2 icp), n_free; float fx, fy, cx, cy, voxel, mu, vf_min, vf_max;
//            n x float[16] c2w (column-major); n_free x float[16] free-view c2w; n x (uchar4[w*h], int16[w*h])
//   output : n x float[16] GetInvM() per frame; n_free x (uchar4[w*h] free image, float4[w*h] free vertex); 2 x float4[w*h] raycasts at
//            camPoses[0] / camPoses[n-1] via the camPoses / camIntrincs route; float voxel size
//   argv[3]: directory for SaveToFile; the scene is then loaded into a second engine (LoadFromFile), which fuses the last frame once more
//            at its ground-truth pose; its free-view raycast from there is appended to the output
#include <cstdio>
#include <cstdlib>
#include <vector>

#ifdef GSB_WITH_REFERENCE_CLIENGINE
#include "TsdfFusion/CLIEngine.h"
using namespace InfiniTAM::Engine;
#endif
#include "ITMLib/Core/ITMBasicEngine.h"
#include "ITMLib/ITMLibDefines.h"
#include "ITMLib/Utils/ITMLibSettings.h"
#include "ORUtils/Matrix.h"

using namespace ITMLib;

static void readAll(void *dst, size_t bytes, FILE *f)
{
    if (fread(dst, 1, bytes, f) != bytes)
    {
        fprintf(stderr, "short read\n");
        exit(2);
    }
}

template <class T> static void dumpDevice(const ORUtils::Image<T> *img, FILE *out)
{
    std::vector<T> host(img->dataSize);
    if (cudaMemcpy(host.data(), img->GetData(MEMORYDEVICE_CUDA), host.size() * sizeof(T), cudaMemcpyDeviceToHost) != cudaSuccess)
    {
        fprintf(stderr, "cudaMemcpy failed\n");
        exit(3);
    }
    fwrite(host.data(), sizeof(T), host.size(), out);
}

int main(int argc, char **argv)
{
    if (argc < 3)
    {
        fprintf(stderr, "usage: %s frames.bin out.bin [scene_dir]\n", argv[0]);
        return 1;
    }
    FILE *in = fopen(argv[1], "rb"), *out = fopen(argv[2], "wb");
    if (!in || !out)
        return 1;
    int hdr[5];
    float cal[8];
    readAll(hdr, sizeof hdr, in), readAll(cal, sizeof cal, in);
    const int n = hdr[0], w = hdr[1], h = hdr[2], tracker = hdr[3], nFree = hdr[4];
    std::vector<ORUtils::Matrix4<float> *> gtPoses(n);
    std::vector<ORUtils::Matrix4<float>> freePoses(nFree);
    for (int i = 0; i < n; i++)
        gtPoses[i] = new ORUtils::Matrix4<float>(), readAll(gtPoses[i]->m, 64, in);
    for (int i = 0; i < nFree; i++)
        readAll(freePoses[i].m, 64, in);

    try
    {
        // ---- createTsdfEngine
        ITMRGBDCalib calib;
        calib.intrinsics_rgb.SetFrom(w, h, cal[0], cal[1], cal[2], cal[3]);
        calib.intrinsics_d = calib.intrinsics_rgb;
        calib.disparityCalib.SetStandard();
        std::vector<ITMUChar4Image *> rgbImages(n);
        std::vector<ITMShortImage *> depthImages(n);
        for (int i = 0; i < n; i++)
        {
            rgbImages[i] = new ITMUChar4Image(ORUtils::Vector2<int>(w, h), true, false);
            depthImages[i] = new ITMShortImage(ORUtils::Vector2<int>(w, h), true, false);
            readAll(rgbImages[i]->GetData(MEMORYDEVICE_CPU), (size_t)w * h * 4, in);
            readAll(depthImages[i]->GetData(MEMORYDEVICE_CPU), (size_t)w * h * 2, in);
        }
        ITMLibSettings *settings = new ITMLibSettings();
        settings->sceneParams.voxelSize = cal[4], settings->sceneParams.mu = cal[5];
        settings->sceneParams.viewFrustum_min = cal[6], settings->sceneParams.viewFrustum_max = cal[7];
        if (tracker == 2)
            settings->trackerConfig = "type=icp,levels=rrrbb,minstep=1e-3,outlierC=0.01,outlierF=0.002,numiterC=10,numiterF=2,failureDec=5.0";
        ITMMainEngine *mainEngine = new ITMBasicEngine<ITMVoxel, ITMVoxelIndex>(settings, calib, rgbImages[0]->noDims, depthImages[0]->noDims);
        auto *basic = dynamic_cast<ITMBasicEngine<ITMVoxel, ITMVoxelIndex> *>(mainEngine);
        if (tracker == 0)
        {
            basic->turnOffTracking();
            basic->gtC2wPoses = gtPoses;
        }

        // ---- frame loop + pose read-out
#ifdef GSB_WITH_REFERENCE_CLIENGINE
        CLIEngine *cli = CLIEngine::Instance();
        cli->Initialise(rgbImages, depthImages, mainEngine);
        for (int i = 0; i < n; i++)
        {
            if (!cli->ProcessFrame())
                return 4;
            ORUtils::Matrix4<float> est = cli->getMainEngine()->GetTrackingState()->pose_d->GetInvM();
            fwrite(est.m, 4, 16, out);
        }
        if (cli->ProcessFrame())   // the sequence is exhausted
            return 4;
#else
        for (int i = 0; i < n; i++)
        {
            mainEngine->ProcessFrame(rgbImages[i], depthImages[i]);
            ORUtils::Matrix4<float> est = mainEngine->GetTrackingState()->pose_d->GetInvM();
            fwrite(est.m, 4, 16, out);
        }
#endif
        // ---- runRaycastByCam, the "else" branch: a pose that is not one of the processed cameras
        for (int i = 0; i < nFree; i++)
        {
            ORUtils::SE3Pose pose;
            ITMIntrinsics intr;
            pose.SetInvM(freePoses[i]);
            intr.SetFrom(w, h, cal[0], cal[1], cal[2], cal[3]);
            basic->runRaycast(&pose, &intr);
            dumpDevice(basic->GetFreeImage(), out);
            dumpDevice(basic->GetFreeVertex(), out);
        }
        // ---- runRaycastByCam, the camPoses / camIntrincs branch
        const int ids[2] = {0, n - 1};
        for (int k = 0; k < 2; k++)
        {
            ORUtils::SE3Pose pose = basic->camPoses[ids[k]];
            ITMIntrinsics intr = basic->camIntrincs[ids[k]];
            basic->runRaycast(&pose, &intr);
            dumpDevice(basic->GetFreeVertex(), out);
        }
        const float voxel = basic->getVoxelSize();
        fwrite(&voxel, 4, 1, out);

        // ---- SaveToFile / LoadFromFile round trip into a second engine
        if (argc > 3)
        {
            basic->SaveToFile(argv[3]);
            auto *second = new ITMBasicEngine<ITMVoxel, ITMVoxelIndex>(settings, calib, rgbImages[0]->noDims);
            second->turnOffTracking();
            second->LoadFromFile(argv[3]);
            second->gtC2wPoses.assign(1, gtPoses[n - 1]);
            second->ProcessFrame(rgbImages[n - 1], depthImages[n - 1]);
            ORUtils::SE3Pose pose;
            pose.SetInvM(*gtPoses[n - 1]);
            second->runRaycast(&pose, &calib.intrinsics_d);
            dumpDevice(second->GetFreeImage(), out);
            dumpDevice(second->GetFreeVertex(), out);
            delete second;
        }
        delete mainEngine;
    }
    catch (const std::exception &e)
    {
        fprintf(stderr, "driver: %s\n", e.what());
        return 5;
    }
    fclose(in), fclose(out);
    return 0;
}
