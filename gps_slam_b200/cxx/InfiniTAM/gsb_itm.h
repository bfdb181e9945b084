// gps_slam_b200 C++ host layer, TSDF side: the slice of the InfiniTAM interface that the reference's SLAM code compiles against
// (SURVEY.md section 8b-2), re-implemented over the C ABI of libgpsslam_b200.so.  Users: slam/InfiniTAM_tools.cpp:3-67
// (createTsdfEngine), slam/TsdfFusion/CLIEngine.{h,cpp}, slam/slam_pipeline.{h,cpp} (ProcessFrame, GetTrackingState()->pose_d->GetInvM(),
// runRaycast, GetFreeImage / GetFreeVertex, getVoxelSize, camPoses / camIntrincs, SaveToFile / LoadFromFile), src/cv_utils.cpp:216-341
// and src/tensor_math.cpp:5-39 (ORUtils::Image / Matrix4 access) in the reference tree.  The headers of the same names next to this
// file (ITMLib/Core/ITMBasicEngine.h, ORUtils/Matrix.h, ...) only include this one.
//
// Same names, argument meaning and error behaviour: types live in ORUtils:: / ITMLib:: with the reference's global typedefs
// (Vector2i, Vector4u, ITMUChar4Image, ...), ORUtils::Matrix4 is column-major m[col*4+row] with operator()(col,row), host image
// memory is pinned (the reference's MemoryBlock uses cudaMallocHost when built with CUDA), engine errors raise std::runtime_error
// like DIEWITHEXCEPTION.  Everything per-frame runs on the GPU inside the library: there is no CPU engine behind this facade
// (settings->deviceType is ignored) and construction fails loudly without a CUDA device.
//
// Not provided (SURVEY.md section 8 marks them out of scope): GetView / GetImage visualisations, swapping, surfel and multi-scene
// engines, relocaliser, IMU / colour trackers.
#pragma once

#include <cuda_runtime_api.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <sys/stat.h>
#include <vector>

#include "gpsslam_b200.h"

#ifndef DIEWITHEXCEPTION
#define DIEWITHEXCEPTION(x) throw std::runtime_error(x)
#endif

enum MemoryDeviceType { MEMORYDEVICE_CPU, MEMORYDEVICE_CUDA };

namespace ORUtils
{
// ---- small vectors (x, y, z, w members and the r/g/b, width/height aliases the reference's callers use)
template <class T> struct Vector2
{
    union { struct { T x, y; }; struct { T s, t; }; struct { T width, height; }; T v[2]; };
    Vector2() : x(), y() {}
    Vector2(T x_, T y_) : x(x_), y(y_) {}
    explicit Vector2(T t) : x(t), y(t) {}
    T &operator[](int i) { return v[i]; }
    const T &operator[](int i) const { return v[i]; }
    bool operator==(const Vector2 &o) const { return x == o.x && y == o.y; }
    bool operator!=(const Vector2 &o) const { return !(*this == o); }
};
template <class T> struct Vector3
{
    union { struct { T x, y, z; }; struct { T r, g, b; }; T v[3]; };
    Vector3() : x(), y(), z() {}
    Vector3(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
    explicit Vector3(T t) : x(t), y(t), z(t) {}
    T &operator[](int i) { return v[i]; }
    const T &operator[](int i) const { return v[i]; }
    Vector3 operator+(const Vector3 &o) const { return Vector3(x + o.x, y + o.y, z + o.z); }
    Vector3 operator-(const Vector3 &o) const { return Vector3(x - o.x, y - o.y, z - o.z); }
    Vector3 operator*(T k) const { return Vector3(x * k, y * k, z * k); }
};
template <class T> struct Vector4
{
    union { struct { T x, y, z, w; }; struct { T r, g, b, a; }; T v[4]; };
    Vector4() : x(), y(), z(), w() {}
    Vector4(T x_, T y_, T z_, T w_) : x(x_), y(y_), z(z_), w(w_) {}
    explicit Vector4(T t) : x(t), y(t), z(t), w(t) {}
    T &operator[](int i) { return v[i]; }
    const T &operator[](int i) const { return v[i]; }
};

// ---- 4x4 matrix, column-major storage: m[col * 4 + row], operator()(col, row)  (reference ORUtils/Matrix.h:26-36 convention)
template <class T> struct Matrix4
{
    T m[16];
    Matrix4() { std::memset(m, 0, sizeof m); }
    explicit Matrix4(const T *src) { std::memcpy(m, src, sizeof m); }
    T &operator()(int col, int row) { return m[col * 4 + row]; }
    const T &operator()(int col, int row) const { return m[col * 4 + row]; }
    T &at(int col, int row) { return m[col * 4 + row]; }
    const T &at(int col, int row) const { return m[col * 4 + row]; }
    void setIdentity()
    {
        std::memset(m, 0, sizeof m);
        m[0] = m[5] = m[10] = m[15] = (T)1;
    }
    Matrix4 operator*(const Matrix4 &b) const
    {
        Matrix4 r;
        for (int c = 0; c < 4; c++)
            for (int row = 0; row < 4; row++)
            {
                T s = 0;
                for (int k = 0; k < 4; k++)
                    s += (*this)(k, row) * b(c, k);
                r(c, row) = s;
            }
        return r;
    }
    Vector4<T> operator*(const Vector4<T> &p) const
    {
        Vector4<T> r;
        for (int row = 0; row < 4; row++)
            r[row] = (*this)(0, row) * p.x + (*this)(1, row) * p.y + (*this)(2, row) * p.z + (*this)(3, row) * p.w;
        return r;
    }
    bool operator==(const Matrix4 &o) const { return std::memcmp(m, o.m, sizeof m) == 0; }
};

// ---- rigid pose.  The engine is the authority on both M (world -> camera) and invM (camera -> world): a pose that comes out of the
// engine keeps the exact pair it reported; SetInvM / SetM on the host use the rigid inverse [R|t]^-1 = [R^T | -R^T t].
class SE3Pose
{
    Matrix4<float> M, invM;
    static Matrix4<float> rigidInverse(const Matrix4<float> &a)
    {
        Matrix4<float> r;
        for (int c = 0; c < 3; c++)
            for (int row = 0; row < 3; row++)
                r(c, row) = a(row, c);
        for (int row = 0; row < 3; row++)
            r(3, row) = -(r(0, row) * a(3, 0) + r(1, row) * a(3, 1) + r(2, row) * a(3, 2));
        r(3, 3) = 1.f;
        return r;
    }

public:
    SE3Pose() { M.setIdentity(), invM.setIdentity(); }
    explicit SE3Pose(const Matrix4<float> &src) { SetM(src); }
    void SetM(const Matrix4<float> &src) { M = src, invM = rigidInverse(src); }
    void SetInvM(const Matrix4<float> &src) { invM = src, M = rigidInverse(src); }
    void SetBoth(const Matrix4<float> &m_, const Matrix4<float> &invM_) { M = m_, invM = invM_; }
    void SetFrom(const SE3Pose *o) { M = o->M, invM = o->invM; }
    const Matrix4<float> &GetM() const { return M; }
    Matrix4<float> GetInvM() const { return invM; }
    Vector3<float> GetT() const { return Vector3<float>(M(3, 0), M(3, 1), M(3, 2)); }
};

// ---- image: host side pinned, device side cudaMalloc, or a non-owning view of engine memory (GetFreeImage / GetFreeVertex)
template <class T> class Image
{
    T *host_ = nullptr, *dev_ = nullptr;
    bool ownsDev_ = false;
    static void ok(cudaError_t e)
    {
        if (e != cudaSuccess)
            DIEWITHEXCEPTION(std::string("CUDA error in ORUtils::Image: ") + cudaGetErrorString(e));
    }

public:
    Vector2<int> noDims;
    size_t dataSize = 0;

    Image(Vector2<int> dims, bool allocate_CPU, bool allocate_CUDA) : noDims(dims), dataSize((size_t)dims.x * dims.y)
    {
        if (allocate_CPU)
            ok(cudaMallocHost((void **)&host_, dataSize * sizeof(T))), std::memset((void *)host_, 0, dataSize * sizeof(T));
        if (allocate_CUDA)
            ok(cudaMalloc((void **)&dev_, dataSize * sizeof(T))), ok(cudaMemset(dev_, 0, dataSize * sizeof(T))), ownsDev_ = true;
    }
    Image(Vector2<int> dims, MemoryDeviceType where) : Image(dims, where == MEMORYDEVICE_CPU, where == MEMORYDEVICE_CUDA) {}
    Image(const Image &) = delete;
    Image &operator=(const Image &) = delete;
    ~Image()
    {
        if (host_) cudaFreeHost(host_);
        if (dev_ && ownsDev_) cudaFree(dev_);
    }
    T *GetData(MemoryDeviceType where) { return where == MEMORYDEVICE_CPU ? host_ : dev_; }
    const T *GetData(MemoryDeviceType where) const { return where == MEMORYDEVICE_CPU ? host_ : dev_; }
    void Clear(unsigned char byte = 0)
    {
        if (host_) std::memset((void *)host_, byte, dataSize * sizeof(T));
        if (dev_ && ownsDev_) ok(cudaMemset(dev_, byte, dataSize * sizeof(T)));
    }
    void UpdateDeviceFromHost() const
    {
        if (host_ && dev_) ok(cudaMemcpy(dev_, host_, dataSize * sizeof(T), cudaMemcpyHostToDevice));
    }
    void UpdateHostFromDevice() const
    {
        if (host_ && dev_) ok(cudaMemcpy(host_, dev_, dataSize * sizeof(T), cudaMemcpyDeviceToHost));
    }
    // facade only: point the device side at memory the engine owns (zero copy, valid until the next raycast)
    void ViewDevice(const void *enginePtr)
    {
        if (dev_ && ownsDev_) cudaFree(dev_);
        dev_ = (T *)enginePtr, ownsDev_ = false;
    }
};
} // namespace ORUtils

typedef ORUtils::Vector2<int> Vector2i;
typedef ORUtils::Vector2<float> Vector2f;
typedef ORUtils::Vector3<float> Vector3f;
typedef ORUtils::Vector3<unsigned char> Vector3u;
typedef ORUtils::Vector4<float> Vector4f;
typedef ORUtils::Vector4<unsigned char> Vector4u;
typedef ORUtils::Matrix4<float> Matrix4f;
typedef ORUtils::Image<Vector4u> ITMUChar4Image;
typedef ORUtils::Image<Vector4f> ITMFloat4Image;
typedef ORUtils::Image<short> ITMShortImage;
typedef ORUtils::Image<float> ITMFloatImage;

// ---- ORUtils/NVTimer.h subset (CLIEngine's stop watches), on std::chrono
struct StopWatchInterface
{
    std::chrono::steady_clock::time_point t0;
    double totalMs = 0.0;
    int sessions = 0;
    bool running = false;
};
inline bool sdkCreateTimer(StopWatchInterface **t) { *t = new StopWatchInterface(); return true; }
inline bool sdkDeleteTimer(StopWatchInterface **t) { delete *t; *t = nullptr; return true; }
inline bool sdkStartTimer(StopWatchInterface **t) { (*t)->t0 = std::chrono::steady_clock::now(), (*t)->running = true; return true; }
inline bool sdkStopTimer(StopWatchInterface **t)
{
    if ((*t)->running)
        (*t)->totalMs += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - (*t)->t0).count(), (*t)->sessions++;
    (*t)->running = false;
    return true;
}
inline bool sdkResetTimer(StopWatchInterface **t) { (*t)->totalMs = 0.0, (*t)->sessions = 0, (*t)->running = false; return true; }
inline float sdkGetTimerValue(StopWatchInterface **t) { return (float)(*t)->totalMs; }
inline float sdkGetAverageTimerValue(StopWatchInterface **t) { return (*t)->sessions ? (float)((*t)->totalMs / (*t)->sessions) : 0.f; }

namespace ITMLib
{
struct ITMVoxel_s_rgb {};       // 8-byte voxel {short sdf; uchar w_depth; uchar3 clr; uchar w_color}: the layout the library implements
struct ITMVoxelBlockHash {};    // 8x8x8 blocks behind the 0x100000 + 0x20000 entry hash table
class ITMIMUMeasurement {};
class ITMView {};

class ITMIntrinsics
{
public:
    struct ProjectionParamsSimple
    {
        Vector4f all;
        float fx, fy, px, py;
    } projectionParamsSimple;
    Vector2i imgSize;
    void SetFrom(int width, int height, float fx, float fy, float cx, float cy)
    {
        imgSize = Vector2i(width, height);
        projectionParamsSimple.fx = fx, projectionParamsSimple.fy = fy, projectionParamsSimple.px = cx, projectionParamsSimple.py = cy;
        projectionParamsSimple.all = Vector4f(fx, fy, cx, cy);
    }
    ITMIntrinsics() { SetFrom(640, 480, 580.f, 580.f, 320.f, 240.f); }
};

class ITMExtrinsics
{
public:
    Matrix4f calib, calib_inv;
    ITMExtrinsics() { calib.setIdentity(), calib_inv.setIdentity(); }
};

class ITMDisparityCalib
{
public:
    enum TrafoType { TRAFO_KINECT, TRAFO_AFFINE };
    void SetStandard() { SetFrom(1.0f / 1000.0f, 0.0f, TRAFO_AFFINE); }   // depth in millimetres, no offset
    void SetFrom(float a, float b, TrafoType t)
    {
        if (a != 1.0f / 1000.0f || b != 0.0f || t != TRAFO_AFFINE)
            DIEWITHEXCEPTION("gps_slam_b200: only the standard depth calibration (millimetres, affine, no offset) is supported");
        params = ORUtils::Vector2<float>(a, b), type = t;
    }
    const ORUtils::Vector2<float> &GetParams() const { return params; }
    TrafoType GetType() const { return type; }
    ITMDisparityCalib() : params(1.0f / 1000.0f, 0.0f), type(TRAFO_AFFINE) {}

private:
    ORUtils::Vector2<float> params;
    TrafoType type;
};

class ITMRGBDCalib
{
public:
    ITMIntrinsics intrinsics_rgb, intrinsics_d;
    ITMExtrinsics trafo_rgb_to_depth;
    ITMDisparityCalib disparityCalib;
};

class ITMSceneParams
{
public:
    float voxelSize, viewFrustum_min, viewFrustum_max, mu;
    int maxW;
    bool stopIntegratingAtMaxW;
    ITMSceneParams(float mu_, int maxW_, float voxelSize_, float vfMin, float vfMax, bool stopAtMaxW)
        : voxelSize(voxelSize_), viewFrustum_min(vfMin), viewFrustum_max(vfMax), mu(mu_), maxW(maxW_), stopIntegratingAtMaxW(stopAtMaxW) {}
};

class ITMLibSettings
{
public:
    typedef enum { DEVICE_CPU, DEVICE_CUDA, DEVICE_METAL } DeviceType;
    typedef enum { FAILUREMODE_RELOCALISE, FAILUREMODE_IGNORE, FAILUREMODE_STOP_INTEGRATION } FailureMode;
    typedef enum { SWAPPINGMODE_DISABLED, SWAPPINGMODE_ENABLED, SWAPPINGMODE_DELETE } SwappingMode;
    typedef enum { LIBMODE_BASIC, LIBMODE_BASIC_SURFELS, LIBMODE_LOOPCLOSURE } LibMode;
    DeviceType deviceType = DEVICE_CUDA;
    bool useApproximateRaycast = false, useBilateralFilter = false, skipPoints = true, createMeshingEngine = true;
    FailureMode behaviourOnFailure = FAILUREMODE_IGNORE;
    SwappingMode swappingMode = SWAPPINGMODE_DISABLED;
    LibMode libMode = LIBMODE_BASIC;
    // the reference's compiled-in default (Utils/ITMLibSettings.cpp:54-57 there); "type=icp..." selects the plain depth tracker
    const char *trackerConfig = "type=extended,levels=rrbb,useDepth=1,minstep=1e-4,outlierSpaceC=0.1,outlierSpaceF=0.004,"
                                "numiterC=20,numiterF=50,tukeyCutOff=8,framesToSkip=20,framesToWeight=50,failureDec=20.0";
    ITMSceneParams sceneParams{0.02f, 100, 0.005f, 0.2f, 3.0f, false};
    int cudaDevice = 0;   // facade only
    virtual ~ITMLibSettings() {}
    MemoryDeviceType GetMemoryType() const { return MEMORYDEVICE_CUDA; }
};

class ITMTrackingState
{
public:
    enum TrackingResult { TRACKING_GOOD = 2, TRACKING_POOR = 1, TRACKING_FAILED = 0 };
    TrackingResult trackerResult = TRACKING_GOOD;
    float trackerScore = 0.f;
    ORUtils::SE3Pose *pose_d;
    ITMTrackingState() : pose_d(new ORUtils::SE3Pose()) {}
    ITMTrackingState(const ITMTrackingState &) = delete;
    ~ITMTrackingState() { delete pose_d; }
};

class ITMMainEngine
{
public:
    virtual ITMView *GetView(void) = 0;
    virtual ITMTrackingState *GetTrackingState(void) = 0;
    virtual ITMTrackingState::TrackingResult ProcessFrame(ITMUChar4Image *rgbImage, ITMShortImage *rawDepthImage,
                                                          ITMIMUMeasurement *imuMeasurement = NULL) = 0;
    virtual Vector2i GetImageSize(void) const = 0;
    virtual void SaveSceneToMesh(const char *fileName) {}
    virtual void SaveToFile() {}
    virtual void LoadFromFile() {}
    virtual ~ITMMainEngine() {}
};

// ITMLib::ITMBasicEngine<ITMVoxel_s_rgb, ITMVoxelBlockHash>: public surface of Core/ITMBasicEngine.h:52-110 in the reference
template <class TVoxel, class TIndex> class ITMBasicEngine : public ITMMainEngine
{
    gsb_tsdf_t *h_ = nullptr;
    gsb_tsdf_config_t cfg_;
    ITMTrackingState state_;
    ITMIntrinsics intrinsicsD_;
    ITMUChar4Image freeImage_;
    ITMFloat4Image freeVertex_, liveVertex_;
    bool trackingActive_ = true;
    bool syncAfterCalls_ = true;
    int framesProcessed_ = 0;

    static void check(int rc, const char *what)
    {
        if (rc != 0)
            DIEWITHEXCEPTION(std::string("gps_slam_b200 ") + what + ": " + gsb_last_error());
    }
    void create()
    {
        if (h_) gsb_tsdf_destroy(h_), h_ = nullptr;
        check(gsb_tsdf_create(&cfg_, &h_), "gsb_tsdf_create");
    }
    void refreshPose()
    {
        Matrix4f M, invM;
        check(gsb_tsdf_get_pose(h_, M.m, invM.m), "gsb_tsdf_get_pose");
        state_.pose_d->SetBoth(M, invM);
    }
    static void writeBlock(const std::string &path, const void *data, size_t count, size_t elemSize)
    {
        std::ofstream f(path, std::ios::binary);
        if (!f) DIEWITHEXCEPTION("could not open " + path + " for writing");
        f.write((const char *)&count, sizeof(size_t));
        f.write((const char *)data, (std::streamsize)(count * elemSize));
    }
    // expectedCount: the element count this engine's arrays have; a header that says otherwise (truncated / foreign / corrupt file)
    // is rejected before anything is allocated from it
    static size_t readBlock(const std::string &path, std::vector<char> &data, size_t elemSize, size_t expectedCount)
    {
        std::ifstream f(path, std::ios::binary);
        if (!f) DIEWITHEXCEPTION("could not open " + path + " for reading");
        size_t count = 0;
        f.read((char *)&count, sizeof(size_t));
        if ((size_t)f.gcount() != sizeof(size_t)) DIEWITHEXCEPTION(path + " has no header");
        if (count != expectedCount)
            DIEWITHEXCEPTION(path + " holds " + std::to_string(count) + " elements, this engine expects " + std::to_string(expectedCount));
        data.resize(count * elemSize);
        f.read(data.data(), (std::streamsize)data.size());
        if ((size_t)f.gcount() != data.size()) DIEWITHEXCEPTION(path + " is shorter than its header says");
        return count;
    }
    static const size_t kHashEntries = 0x100000 + 0x20000, kExcess = 0x20000;

public:
    std::vector<ORUtils::Matrix4<float> *> gtC2wPoses;   // one per frame when tracking is off (createTsdfEngine)
    std::vector<ORUtils::SE3Pose> camPoses;               // pose of every processed frame
    std::vector<ITMLib::ITMIntrinsics> camIntrincs;

    ITMBasicEngine(const ITMLibSettings *settings, const ITMRGBDCalib &calib, Vector2i imgSize_rgb, Vector2i imgSize_d = Vector2i(-1, -1))
        : freeImage_(Vector2i(0, 0), false, false), freeVertex_(Vector2i(0, 0), false, false), liveVertex_(Vector2i(0, 0), false, false)
    {
        if (imgSize_d.x == -1 || imgSize_d.y == -1) imgSize_d = imgSize_rgb;
        if (imgSize_d != imgSize_rgb) DIEWITHEXCEPTION("gps_slam_b200: colour and depth images must have the same size");
        if (settings->swappingMode != ITMLibSettings::SWAPPINGMODE_DISABLED || settings->libMode != ITMLibSettings::LIBMODE_BASIC)
            DIEWITHEXCEPTION("gps_slam_b200: only LIBMODE_BASIC without swapping is supported");
        gsb_tsdf_default_config(&cfg_);
        const ITMIntrinsics::ProjectionParamsSimple &p = calib.intrinsics_d.projectionParamsSimple;
        cfg_.width = imgSize_d.x, cfg_.height = imgSize_d.y, cfg_.fx = p.fx, cfg_.fy = p.fy, cfg_.cx = p.px, cfg_.cy = p.py;
        cfg_.voxel_size = settings->sceneParams.voxelSize, cfg_.mu = settings->sceneParams.mu, cfg_.max_w = settings->sceneParams.maxW;
        cfg_.view_frustum_min = settings->sceneParams.viewFrustum_min, cfg_.view_frustum_max = settings->sceneParams.viewFrustum_max;
        cfg_.device = settings->cudaDevice;
        const std::string tc = settings->trackerConfig ? settings->trackerConfig : "";
        if (tc.find("type=extended") != std::string::npos || tc.compare(0, 8, "extended") == 0) cfg_.tracker = 1;
        else if (tc.find("type=icp") != std::string::npos) cfg_.tracker = 2;
        else DIEWITHEXCEPTION("gps_slam_b200: trackerConfig must select type=extended or type=icp");
        intrinsicsD_ = calib.intrinsics_d;
        freeImage_.noDims = freeVertex_.noDims = liveVertex_.noDims = imgSize_d;
        freeImage_.dataSize = freeVertex_.dataSize = liveVertex_.dataSize = (size_t)imgSize_d.x * imgSize_d.y;
        create();
    }
    ~ITMBasicEngine() { if (h_) gsb_tsdf_destroy(h_); }

    ITMView *GetView(void) { return nullptr; }
    ITMTrackingState *GetTrackingState(void) { return &state_; }
    Vector2i GetImageSize(void) const { return Vector2i(cfg_.width, cfg_.height); }

    // Core/ITMBasicEngine.tpp:260-385: (track | take gtC2wPoses[frame]) -> allocate -> integrate -> raycast for the next frame
    ITMTrackingState::TrackingResult ProcessFrame(ITMUChar4Image *rgbImage, ITMShortImage *rawDepthImage, ITMIMUMeasurement * = NULL)
    {
        const float *gt = nullptr;
        if (!trackingActive_)
        {
            if ((size_t)framesProcessed_ >= gtC2wPoses.size()) DIEWITHEXCEPTION("gps_slam_b200: tracking is off and gtC2wPoses has no pose for this frame");
            gt = gtC2wPoses[framesProcessed_]->m;
        }
        if (rgbImage->noDims != GetImageSize() || rawDepthImage->noDims != GetImageSize()) DIEWITHEXCEPTION("gps_slam_b200: frame size differs from the engine's");
        check(gsb_tsdf_process_frame(h_, (const uint8_t *)rgbImage->GetData(MEMORYDEVICE_CPU),
                                     (const int16_t *)rawDepthImage->GetData(MEMORYDEVICE_CPU), gt), "gsb_tsdf_process_frame");
        refreshPose();
        state_.trackerResult = ITMTrackingState::TRACKING_GOOD;
        if (trackingActive_)
        {
            int result = 2, iterations = 0;
            check(gsb_tsdf_tracker_result(h_, &result, &state_.trackerScore, &iterations), "gsb_tsdf_tracker_result");
            state_.trackerResult = (ITMTrackingState::TrackingResult)result;
        }
        camPoses.emplace_back();
        camPoses.back().SetFrom(state_.pose_d);
        camIntrincs.push_back(intrinsicsD_);
        framesProcessed_++;
        return state_.trackerResult;
    }

    // Core/ITMBasicEngine.tpp:500-526: free-view raycast + colour render; NULL arguments = the live pose / depth intrinsics
    void runRaycast(ORUtils::SE3Pose *pose = NULL, ITMIntrinsics *intrinsics = NULL)
    {
        const ORUtils::SE3Pose *p = pose ? pose : state_.pose_d;
        const ITMIntrinsics::ProjectionParamsSimple &k = (intrinsics ? intrinsics : &intrinsicsD_)->projectionParamsSimple;
        const Matrix4f invM = p->GetInvM();
        check(gsb_tsdf_run_raycast(h_, invM.m, k.fx, k.fy, k.px, k.py), "gsb_tsdf_run_raycast");
        freeImage_.ViewDevice(gsb_tsdf_free_image_dev(h_));
        freeVertex_.ViewDevice(gsb_tsdf_free_vertex_dev(h_));
        // the reference's engine works on the legacy default stream, so its callers read the images from any stream without further
        // ado (torch::from_blob(...).clone(), src/cv_utils.cpp:322-341 there); the library works on its own stream: wait for it
        if (syncAfterCalls_) sync();
    }
    // device pointers valid until the next runRaycast (the reference's callers clone them at once, src/cv_utils.cpp:322-341 there)
    ORUtils::Image<Vector4u> *GetFreeImage() { return &freeImage_; }
    ORUtils::Image<Vector4f> *GetFreeVertex() { return &freeVertex_; }
    ORUtils::Image<Vector4f> *GetLiveVertex()
    {
        liveVertex_.ViewDevice(gsb_tsdf_live_vertex_dev(h_));
        return &liveVertex_;
    }

    float getVoxelSize() { return gsb_tsdf_voxel_size(h_); }
    void turnOnTracking() { check(gsb_tsdf_set_tracking(h_, 1), "gsb_tsdf_set_tracking"), trackingActive_ = true; }
    // with tracking off every frame takes its pose from gtC2wPoses (Core/ITMBasicEngine.tpp:278 there)
    void turnOffTracking() { check(gsb_tsdf_set_tracking(h_, 0), "gsb_tsdf_set_tracking"), trackingActive_ = false; }
    void resetAll()
    {
        check(gsb_tsdf_reset(h_), "gsb_tsdf_reset");
        framesProcessed_ = 0, camPoses.clear(), camIntrincs.clear();
    }
    void sync() { check(gsb_tsdf_sync(h_), "gsb_tsdf_sync"); }   // facade only
    // facade only: run the engine on the caller's stream (e.g. at::cuda::getCurrentCUDAStream()); results are then ordered with the
    // caller's own work on that stream and runRaycast no longer blocks the host
    void useStream(cudaStream_t s)
    {
        check(gsb_tsdf_set_stream(h_, s ? (void *)s : (void *)cudaStreamLegacy), "gsb_tsdf_set_stream");
        syncAfterCalls_ = false;
    }
    gsb_tsdf_t *handle() { return h_; }                           // facade only

    // Core/ITMBasicEngine.tpp:105-117 + ITMMesh::WritePLY (Objects/Meshing/ITMMesh.h:39-104): marching cubes on the device
    // (gsb_tsdf_mesh: deterministic, the CPU mesher's triangle order, per-vertex colours like the CUDA mesher), then the reference's
    // ASCII PLY: three vertices per triangle with uchar colours, one face per triangle
    void SaveSceneToMesh(const char *fileName)
    {
        long long n = 0;
        if (gsb_tsdf_mesh(h_, nullptr, 0, &n))
            DIEWITHEXCEPTION(gsb_last_error());
        float *dev = nullptr;
        std::vector<float> tri((size_t)(n > 0 ? n : 1) * 18);
        if (n > 0)
        {
            if (cudaMalloc((void **)&dev, (size_t)(n + 1) * 18 * sizeof(float)) != cudaSuccess)
                DIEWITHEXCEPTION("gps_slam_b200: out of device memory for the mesh");
            const int rc = gsb_tsdf_mesh(h_, dev, n + 1, &n);
            if (!rc)
                cudaMemcpy(tri.data(), dev, (size_t)n * 18 * sizeof(float), cudaMemcpyDeviceToHost);
            cudaFree(dev);
            if (rc)
                DIEWITHEXCEPTION(gsb_last_error());
        }
        printf("write ply mesh...\n");
        FILE *f = fopen(fileName, "w");
        if (!f)
            return;
        fprintf(f, "ply\nformat ascii 1.0\nelement vertex %d\n", (int)(n * 3));
        fprintf(f, "property float x\nproperty float y\nproperty float z\nproperty uchar red\nproperty uchar green\nproperty uchar blue\n");
        fprintf(f, "element face %d\nproperty list uchar int vertex_indices\nend_header\n", (int)n);
        for (long long i = 0; i < n; i++)
        {
            const float *t = tri.data() + (size_t)i * 18;
            for (int v = 0; v < 3; v++)
                fprintf(f, "%f %f %f %d %d %d\n", t[v * 3], t[v * 3 + 1], t[v * 3 + 2], static_cast<unsigned char>(t[9 + v * 3] * 255),
                        static_cast<unsigned char>(t[9 + v * 3 + 1] * 255), static_cast<unsigned char>(t[9 + v * 3 + 2] * 255));
        }
        for (long long i = 0; i < n; i++)
            fprintf(f, "3 %d %d %d\n", (int)(i * 3), (int)(i * 3 + 1), (int)(i * 3 + 2));
        fclose(f);
    }

    // Core/ITMBasicEngine.tpp:119-171: the Scene/ directory in the reference's own file layout (see gps_slam_b200/checkpoint.py)
    void SaveToFile(const std::string &saveOutputDirectory)
    {
        const std::string scene = saveOutputDirectory + "/Scene";
        mkdir(saveOutputDirectory.c_str(), 0777), mkdir(scene.c_str(), 0777), mkdir((saveOutputDirectory + "/Relocaliser").c_str(), 0777);
        int lastBlock = 0, lastExcess = 0;
        check(gsb_tsdf_counter(h_, 0, &lastBlock), "gsb_tsdf_counter"), check(gsb_tsdf_counter(h_, 1, &lastExcess), "gsb_tsdf_counter");
        const size_t blocks = (size_t)(cfg_.num_blocks > 0 ? cfg_.num_blocks : 0x40000), voxels = blocks * 512;
        std::vector<char> buf(voxels * 8);
        check(gsb_tsdf_read(h_, GSB_TSDF_VOXELS, buf.data(), buf.size()), "gsb_tsdf_read voxels");
        writeBlock(scene + "/voxel.dat", buf.data(), voxels, 8);
        std::vector<int> ids(blocks > kExcess ? blocks : kExcess);
        for (size_t i = 0; i < ids.size(); i++) ids[i] = (int)i;
        writeBlock(scene + "/alloc.dat", ids.data(), blocks, 4);
        std::ofstream(scene + "/vba.txt") << lastBlock << " " << voxels;
        std::ofstream(scene + "/last.txt") << lastExcess;
        buf.resize(kHashEntries * 16);
        check(gsb_tsdf_read(h_, GSB_TSDF_HASH_TABLE, buf.data(), buf.size()), "gsb_tsdf_read hash table");
        writeBlock(scene + "/hash.dat", buf.data(), kHashEntries, 16);
        writeBlock(scene + "/excess.dat", ids.data(), kExcess, 4);
    }
    void LoadFromFile(const std::string &saveInputDirectory)
    {
        const std::string scene = saveInputDirectory + "/Scene";
        std::vector<char> hash, voxels, list;
        const size_t blocks = (size_t)(cfg_.num_blocks > 0 ? cfg_.num_blocks : 0x40000);
        const size_t nHash = readBlock(scene + "/hash.dat", hash, 16, kHashEntries), nVox = readBlock(scene + "/voxel.dat", voxels, 8, blocks * 512);
        for (int which = 0; which < 2; which++)
        {
            const size_t n = readBlock(scene + (which ? "/excess.dat" : "/alloc.dat"), list, 4, which ? kExcess : blocks);
            for (size_t i = 0; i < n; i++)
                if (((const int *)list.data())[i] != (int)i)
                    DIEWITHEXCEPTION("gps_slam_b200: free lists are not the identity permutation (scene saved with swapping): not supported");
        }
        int lastBlock = 0, lastExcess = 0;
        {
            std::ifstream fv(scene + "/vba.txt"), fl(scene + "/last.txt");
            if (!(fv >> lastBlock)) DIEWITHEXCEPTION("could not read " + scene + "/vba.txt");
            if (!(fl >> lastExcess)) DIEWITHEXCEPTION("could not read " + scene + "/last.txt");
        }
        check(gsb_tsdf_load_scene(h_, hash.data(), nHash, voxels.data(), nVox, lastBlock, lastExcess), "gsb_tsdf_load_scene");
    }
    void SaveToFile() {}
    void LoadFromFile() {}
};
} // namespace ITMLib

typedef ITMLib::ITMVoxel_s_rgb ITMVoxel;
typedef ITMLib::ITMVoxelBlockHash ITMVoxelIndex;
