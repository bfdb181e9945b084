// gps_slam_b200 C++ host layer, gsplat side: the gsplat::*_tensor functions (rasterizer/bindings.h), fusedssim, distCUDA2 and
// the free functions of gsplat_wapper.hpp, each a thin translation of torch tensors onto the staged C-ABI entry points of
// libgpsslam_b200.so (include/gpsslam_b200.h).  This file contains no arithmetic on the data path: it checks shapes, allocates the
// outputs as torch tensors exactly where the reference does (ATen caching allocator, caller owns them), points the engine at
// torch's current CUDA stream (the reference launches on at::cuda::getCurrentCUDAStream()) and calls the library.
//
// Engines: the staged entry points need a gsb_gs_t for workspace (bins, packed splat records, per-pixel records).  One engine is
// kept per (device, image size, projection constants); it is created on first use and lives until process exit, like the
// reference's static CUDA state.  Capacity: GSB_SHIM_CAPACITY Gaussians (environment, default 2^20).  Single host thread, as the
// reference.
#include "gsplat_wapper.hpp"

#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include <cstdlib>
#include <memory>
#include <vector>

#include "gpsslam_b200.h"

namespace
{
using torch::Tensor;

struct Engine
{
    gsb_gs_t *h = nullptr;
    int device = 0, W = 0, H = 0, tileW = 0, tileH = 0, capacity = 0;
    float eps2d = 0.3f, nearPlane = 0.01f, farPlane = 1e10f, radiusClip = 0.f;
    bool constantsPinned = false;   // created by a projection call (constants known) rather than by a size-only call
    // identity of the last binning, so that isectOffsetEncode* can hand back the offsets computed with it
    const void *lastIsectIds = nullptr;
    int64_t lastIsects = -1;
    bool lastWithDepth = false;
    ~Engine() { if (h) gsb_gs_destroy(h); }
};

std::vector<std::unique_ptr<Engine>> &registry()
{
    static std::vector<std::unique_ptr<Engine>> r;
    return r;
}

void fail(const char *what) { TORCH_CHECK(false, "gps_slam_b200: ", what, ": ", gsb_last_error()); }
#define GSB(call) do { if ((call) != 0) fail(#call); } while (0)

int shimCapacity()
{
    const char *s = getenv("GSB_SHIM_CAPACITY");
    long v = s ? atol(s) : 0;
    return v > 0 ? (int)v : (1 << 20);
}

Engine *create(int device, int W, int H, const float *constants)
{
    gsb_gs_config_t c;
    gsb_gs_default_config(&c);
    c.device = device, c.width = W, c.height = H, c.capacity = shimCapacity();
    c.max_gs_radii = 0;   // the wrappers never clamp: RawGaussianModel does it itself (src/raw_gs_model.cpp:241-242)
    if (constants)
        c.eps2d = constants[0], c.near_plane = constants[1], c.far_plane = constants[2], c.radius_clip = constants[3];
    auto e = std::make_unique<Engine>();
    if (gsb_gs_create(&c, &e->h) != 0)
        fail("gsb_gs_create");
    e->device = device, e->W = W, e->H = H, e->tileW = (W + 15) / 16, e->tileH = (H + 15) / 16, e->capacity = c.capacity;
    e->eps2d = c.eps2d, e->nearPlane = c.near_plane, e->farPlane = c.far_plane, e->radiusClip = c.radius_clip;
    e->constantsPinned = constants != nullptr;
    registry().push_back(std::move(e));
    return registry().back().get();
}

// move the entry to the back (= most recently used) and bind it to torch's current stream
Engine *use(size_t i)
{
    auto &r = registry();
    if (i + 1 != r.size())
        std::rotate(r.begin() + i, r.begin() + i + 1, r.end());
    Engine *e = r.back().get();
    cudaStream_t s = at::cuda::getCurrentCUDAStream(e->device).stream();
    GSB(gsb_gs_set_stream(e->h, s ? (void *)s : (void *)cudaStreamLegacy));   // handle 0 means "engine's private stream" to the C ABI
    return e;
}

Engine *engineForProjection(int device, int W, int H, float eps2d, float nearP, float farP, float clip)
{
    auto &r = registry();
    for (size_t i = r.size(); i-- > 0;)
    {
        Engine *e = r[i].get();
        if (e->device == device && e->W == W && e->H == H && e->eps2d == eps2d && e->nearPlane == nearP && e->farPlane == farP &&
            e->radiusClip == clip)
            return use(i);
    }
    const float k[4] = {eps2d, nearP, farP, clip};
    create(device, W, H, k);
    return use(r.size() - 1);
}

// most recently used engine of that image size; eps2d < 0 = any
Engine *engineForSize(int device, int W, int H, float eps2d = -1.f)
{
    auto &r = registry();
    for (size_t i = r.size(); i-- > 0;)
        if (r[i]->device == device && r[i]->W == W && r[i]->H == H && (eps2d < 0.f || r[i]->eps2d == eps2d))
            return use(i);
    if (eps2d >= 0.f)
    {
        const float k[4] = {eps2d, 0.01f, 1e10f, 0.f};
        create(device, W, H, k);
    }
    else
        create(device, W, H, nullptr);
    return use(r.size() - 1);
}

Engine *engineForTiles(int device, int tileW, int tileH)
{
    auto &r = registry();
    for (size_t i = r.size(); i-- > 0;)
        if (r[i]->device == device && r[i]->tileW == tileW && r[i]->tileH == tileH)
            return use(i);
    create(device, tileW * 16, tileH * 16, nullptr);
    return use(r.size() - 1);
}

// calls that only need a stream and scratch (SH, SSIM, KNN): whatever engine was used last on that device
Engine *engineAny(int device)
{
    auto &r = registry();
    for (size_t i = r.size(); i-- > 0;)
        if (r[i]->device == device)
            return use(i);
    create(device, 64, 64, nullptr);
    return use(r.size() - 1);
}

void checkF32(const Tensor &t, const char *name)
{
    TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor");
    TORCH_CHECK(t.is_contiguous(), name, " must be contiguous");
    TORCH_CHECK(t.scalar_type() == torch::kFloat32, name, " must be float32");
}
void checkI32(const Tensor &t, const char *name)
{
    TORCH_CHECK(t.is_cuda() && t.is_contiguous(), name, " must be a contiguous CUDA tensor");
    TORCH_CHECK(t.scalar_type() == torch::kInt32, name, " must be int32");
}
int deviceOf(const Tensor &t) { return t.get_device(); }
const float *F(const Tensor &t) { return t.data_ptr<float>(); }
float *F(Tensor &t) { return t.data_ptr<float>(); }

// 16 + 9 floats of camera to the host (the C ABI takes them as host arrays; the reference reads them on the device)
struct HostCam
{
    Tensor vm, K;
    HostCam(const Tensor &viewmats, const Tensor &Ks)
    {
        TORCH_CHECK(viewmats.numel() == 16 && Ks.numel() == 9, "one camera per call (C = 1): viewmats [1,4,4], Ks [1,3,3]");
        vm = viewmats.detach().to(torch::kCPU, torch::kFloat32).contiguous();
        K = Ks.detach().to(torch::kCPU, torch::kFloat32).contiguous();
    }
};

void noGesExtras(const gsplat::OptT &backgrounds, const gsplat::OptT &mask)
{
    TORCH_CHECK(!backgrounds.has_value(), "GES rasteriser: backgrounds are not supported (the reference's model passes none)");
    TORCH_CHECK(!mask.has_value(), "tile masks are not supported");
}
} // namespace

// ------------------------------------------------------------------------------------------------ gsplat::*_tensor
namespace gsplat
{

std::tuple<T, T, T, T, T> fully_fused_projection_fwd_tensor(const T &means, const OptT &covars, const OptT &quats, const OptT &scales,
                                                            const T &viewmats, const T &Ks, const uint32_t image_width,
                                                            const uint32_t image_height, const float eps2d, const float near_plane,
                                                            const float far_plane, const float radius_clip, const bool calc_compensations,
                                                            const CameraModelType camera_model)
{
    checkF32(means, "means");
    TORCH_CHECK(!covars.has_value() && quats.has_value() && scales.has_value(), "projection takes the quats + scales form only");
    TORCH_CHECK(camera_model == PINHOLE, "only the pinhole camera model is on the SLAM path");
    TORCH_CHECK(!calc_compensations, "calc_compensations is not supported (the reference's model passes false)");
    checkF32(*quats, "quats"), checkF32(*scales, "scales");
    const at::cuda::OptionalCUDAGuard guard(device_of(means));
    const int64_t N = means.size(0);
    TORCH_CHECK(means.dim() == 2 && means.size(1) == 3 && quats->numel() == N * 4 && scales->numel() == N * 3, "means [N,3], quats [N,4], scales [N,3]");
    HostCam cam(viewmats, Ks);
    Engine *e = engineForProjection(deviceOf(means), image_width, image_height, eps2d, near_plane, far_plane, radius_clip);
    TORCH_CHECK(N <= e->capacity, "more Gaussians than GSB_SHIM_CAPACITY");
    T radii = torch::empty({1, N}, means.options().dtype(torch::kInt32));
    T means2d = torch::empty({1, N, 2}, means.options()), depths = torch::empty({1, N}, means.options()),
      conics = torch::empty({1, N, 3}, means.options());
    if (N > 0)
        GSB(gsb_gs_projection_fwd(e->h, (int)N, F(means), F(*quats), F(*scales), F(cam.vm), F(cam.K), 0, radii.data_ptr<int>(), F(means2d),
                                  F(depths), F(conics)));
    return {radii, means2d, depths, conics, T()};
}

std::tuple<T, T, T, T, T> fully_fused_projection_bwd_tensor(const T &means, const OptT &covars, const OptT &quats, const OptT &scales,
                                                            const T &viewmats, const T &Ks, const uint32_t image_width,
                                                            const uint32_t image_height, const float eps2d, const CameraModelType camera_model,
                                                            const T &radii, const T &conics, const OptT &compensations, const T &v_means2d,
                                                            const T &v_depths, const T &v_conics, const OptT &v_compensations,
                                                            const bool viewmats_requires_grad)
{
    checkF32(means, "means"), checkI32(radii, "radii"), checkF32(conics, "conics");
    checkF32(v_means2d, "v_means2d"), checkF32(v_depths, "v_depths"), checkF32(v_conics, "v_conics");
    TORCH_CHECK(!covars.has_value() && quats.has_value() && scales.has_value(), "projection takes the quats + scales form only");
    TORCH_CHECK(camera_model == PINHOLE && !compensations.has_value() && !v_compensations.has_value(), "pinhole, no compensations");
    // viewmats_requires_grad: the reference's wrapper discards v_viewmats (gsplat_wapper.hpp:208 there), so it is not computed
    checkF32(*quats, "quats"), checkF32(*scales, "scales");
    const at::cuda::OptionalCUDAGuard guard(device_of(means));
    const int64_t N = means.size(0);
    HostCam cam(viewmats, Ks);
    Engine *e = engineForSize(deviceOf(means), image_width, image_height, eps2d);
    TORCH_CHECK(N <= e->capacity, "more Gaussians than GSB_SHIM_CAPACITY");
    T v_means = torch::empty({N, 3}, means.options()), v_quats = torch::empty({N, 4}, means.options()),
      v_scales = torch::empty({N, 3}, means.options());
    if (N > 0)
        GSB(gsb_gs_projection_bwd(e->h, (int)N, F(means), F(*quats), F(*scales), F(cam.vm), F(cam.K), radii.data_ptr<int>(), F(conics),
                                  F(v_means2d), F(v_depths), F(v_conics), F(v_means), F(v_quats), F(v_scales)));
    return {v_means, T(), v_quats, v_scales, T()};
}

T compute_sh_fwd_tensor(const uint32_t degrees_to_use, const T &dirs_, const T &coeffs_, const OptT masks)
{
    T dirs = dirs_.contiguous(), coeffs = coeffs_.contiguous();
    checkF32(dirs, "dirs"), checkF32(coeffs, "coeffs");
    TORCH_CHECK(coeffs.size(-1) == 3 && coeffs.size(-2) == 16 && degrees_to_use == 3, "SH degree 3 (16 bases) only, as the reference's model uses");
    const at::cuda::OptionalCUDAGuard guard(device_of(dirs));
    const int64_t N = dirs.numel() / 3;
    TORCH_CHECK(coeffs.numel() == N * 48, "dirs [...,3] and coeffs [...,16,3] disagree");
    T m;
    if (masks.has_value() && masks->defined())
        m = masks->to(torch::kUInt8).contiguous();
    T colors = torch::empty_like(dirs);
    if (N > 0)
        GSB(gsb_gs_sh_fwd(engineAny(deviceOf(dirs))->h, (int)N, 3, F(dirs), F(coeffs), m.defined() ? m.data_ptr<uint8_t>() : nullptr, F(colors)));
    return colors;
}

std::tuple<T, T> compute_sh_bwd_tensor(const uint32_t K, const uint32_t degrees_to_use, const T &dirs_, const T &coeffs_, const OptT masks,
                                       const T &v_colors_, bool compute_v_dirs)
{
    T dirs = dirs_.contiguous(), coeffs = coeffs_.contiguous(), v_colors = v_colors_.contiguous();
    checkF32(dirs, "dirs"), checkF32(coeffs, "coeffs"), checkF32(v_colors, "v_colors");
    TORCH_CHECK(K == 16 && degrees_to_use == 3 && coeffs.size(-2) == 16, "SH degree 3 (16 bases) only");
    const at::cuda::OptionalCUDAGuard guard(device_of(dirs));
    const int64_t N = dirs.numel() / 3;
    T m;
    if (masks.has_value() && masks->defined())
        m = masks->to(torch::kUInt8).contiguous();
    T v_coeffs = torch::empty_like(coeffs), v_dirs = compute_v_dirs ? torch::empty_like(dirs) : T();
    if (N > 0)
        GSB(gsb_gs_sh_bwd(engineAny(deviceOf(dirs))->h, (int)N, 3, F(dirs), F(coeffs), m.defined() ? m.data_ptr<uint8_t>() : nullptr,
                          F(v_colors), F(v_coeffs), compute_v_dirs ? F(v_dirs) : nullptr));
    return {v_coeffs, v_dirs};
}

static std::tuple<T, T, T> binTiles(bool withDepth, const T &means2d, const T &radii, const T &depths, const OptT &camera_ids,
                                    const OptT &gaussian_ids, uint32_t C, uint32_t tile_size, uint32_t tile_width, uint32_t tile_height)
{
    checkF32(means2d, "means2d"), checkI32(radii, "radii");
    TORCH_CHECK(!camera_ids.has_value() && !gaussian_ids.has_value(), "packed mode is not supported");
    TORCH_CHECK(C == 1 && means2d.dim() == 3 && means2d.size(0) == 1, "one camera per call (C = 1)");
    TORCH_CHECK(tile_size == 16, "tile_size must be 16");
    const at::cuda::OptionalCUDAGuard guard(device_of(means2d));
    const int64_t N = means2d.size(1);
    Engine *e = engineForTiles(deviceOf(means2d), tile_width, tile_height);
    TORCH_CHECK(N <= e->capacity, "more Gaussians than GSB_SHIM_CAPACITY");
    T tilesPerGauss = torch::empty({1, N}, radii.options());
    int nIsects = 0;
    if (N > 0)
    {
        if (withDepth)
        {
            checkF32(depths, "depths");
            GSB(gsb_gs_isect_tiles_depth(e->h, (int)N, F(means2d), radii.data_ptr<int>(), F(depths), tilesPerGauss.data_ptr<int>(), &nIsects));
        }
        else
            GSB(gsb_gs_isect_tiles(e->h, (int)N, F(means2d), radii.data_ptr<int>(), tilesPerGauss.data_ptr<int>(), &nIsects));
    }
    T isectIds = torch::empty({nIsects}, radii.options().dtype(torch::kInt64)), flattenIds = torch::empty({nIsects}, radii.options());
    if (nIsects > 0)
    {
        if (withDepth)
            GSB(gsb_gs_isect_fetch_depth(e->h, nIsects, (long long *)isectIds.data_ptr<int64_t>(), flattenIds.data_ptr<int>(), nullptr));
        else
            GSB(gsb_gs_isect_fetch(e->h, nIsects, (long long *)isectIds.data_ptr<int64_t>(), flattenIds.data_ptr<int>(), nullptr));
    }
    e->lastIsectIds = isectIds.data_ptr(), e->lastIsects = nIsects, e->lastWithDepth = withDepth;
    return {tilesPerGauss, isectIds, flattenIds};
}

// offsets of the binning that produced isect_ids: taken from the engine when isect_ids is the tensor the last binning call returned
// (the reference's call order, src/raw_gs_model.cpp:269-277), otherwise recomputed from the ids (tile index = id >> shift)
static T encodeOffsets(bool withDepth, const T &isect_ids, uint32_t C, uint32_t tile_width, uint32_t tile_height)
{
    TORCH_CHECK(C == 1, "one camera per call (C = 1)");
    TORCH_CHECK(isect_ids.is_cuda() && isect_ids.scalar_type() == torch::kInt64, "isect_ids must be a CUDA int64 tensor");
    const at::cuda::OptionalCUDAGuard guard(device_of(isect_ids));
    Engine *e = engineForTiles(deviceOf(isect_ids), tile_width, tile_height);
    T offsets = torch::empty({1, (int64_t)tile_height, (int64_t)tile_width}, isect_ids.options().dtype(torch::kInt32));
    const int64_t n = isect_ids.numel();
    if (n > 0 && e->lastIsectIds == isect_ids.data_ptr() && e->lastIsects == n && e->lastWithDepth == withDepth)
    {
        if (withDepth)
            GSB(gsb_gs_isect_fetch_depth(e->h, (int)n, nullptr, nullptr, offsets.data_ptr<int>()));
        else
            GSB(gsb_gs_isect_fetch(e->h, (int)n, nullptr, nullptr, offsets.data_ptr<int>()));
        return offsets;
    }
    T tiles = withDepth ? isect_ids.bitwise_right_shift(32) : isect_ids;
    T probe = torch::arange((int64_t)tile_width * tile_height, isect_ids.options());
    return torch::searchsorted(tiles.contiguous(), probe, /*out_int32=*/true).view_as(offsets);
}

std::tuple<T, T, T> isect_tiles_tensor(const T &means2d, const T &radii, const T &depths, const OptT &camera_ids, const OptT &gaussian_ids,
                                       const uint32_t C, const uint32_t tile_size, const uint32_t tile_width, const uint32_t tile_height,
                                       const bool, const bool)
{
    return binTiles(true, means2d, radii, depths, camera_ids, gaussian_ids, C, tile_size, tile_width, tile_height);
}

T isect_offset_encode_tensor(const T &isect_ids, const uint32_t C, const uint32_t tile_width, const uint32_t tile_height)
{
    return encodeOffsets(true, isect_ids, C, tile_width, tile_height);
}

std::tuple<T, T, T, T, T> isect_tiles_tensor_no_depth(const T &means2d, const T &radii, const T &depths, const OptT &camera_ids,
                                                      const OptT &gaussian_ids, const uint32_t C, const uint32_t tile_size,
                                                      const uint32_t tile_width, const uint32_t tile_height, const bool, const bool)
{
    auto b = binTiles(false, means2d, radii, depths, camera_ids, gaussian_ids, C, tile_size, tile_width, tile_height);
    T none = torch::empty({0}, radii.options());
    return {std::get<0>(b), std::get<1>(b), std::get<2>(b), none, none};
}

T isect_offset_encode_tensor_no_depth(const T &isect_ids, const uint32_t C, const uint32_t tile_width, const uint32_t tile_height)
{
    return encodeOffsets(false, isect_ids, C, tile_width, tile_height);
}

static void checkSplats(const T &means2d, const T &conics, const T &colors, const T &opacities, int64_t &N)
{
    checkF32(means2d, "means2d"), checkF32(conics, "conics"), checkF32(colors, "colors"), checkF32(opacities, "opacities");
    TORCH_CHECK(means2d.dim() == 3 && means2d.size(0) == 1, "one camera per call (C = 1)");
    N = means2d.size(1);
    TORCH_CHECK(conics.numel() == N * 3 && opacities.numel() == N, "means2d [1,N,2], conics [1,N,3], opacities [N]");
    TORCH_CHECK(colors.numel() == N * 4, "colors must have 4 channels (rgb + camera depth), as RawGaussianModel builds them");
}

std::tuple<T, T, T> rasterize_to_pixels_fwd_ges_tensor(const T &means2d, const T &conics, const T &colors, const T &opacities,
                                                       const T &ref_depth_map, const T &base_color_map, const OptT &backgrounds,
                                                       const OptT &mask, const uint32_t image_width, const uint32_t image_height,
                                                       const uint32_t tile_size, const T &tile_offsets, const T &flatten_ids,
                                                       const float delta_depth)
{
    int64_t N;
    checkSplats(means2d, conics, colors, opacities, N);
    checkF32(ref_depth_map, "ref_depth_map"), checkI32(tile_offsets, "tile_offsets"), checkI32(flatten_ids, "flatten_ids");
    noGesExtras(backgrounds, mask);
    TORCH_CHECK(tile_size == 16, "tile_size must be 16");
    TORCH_CHECK(ref_depth_map.numel() == (int64_t)image_width * image_height, "ref_depth_map must be [1,H,W]");
    const at::cuda::OptionalCUDAGuard guard(device_of(means2d));
    Engine *e = engineForSize(deviceOf(means2d), image_width, image_height);
    TORCH_CHECK(tile_offsets.numel() == (int64_t)e->tileW * e->tileH, "tile_offsets must be [1,tile_height,tile_width]");
    T render = torch::empty({1, (int64_t)image_height, (int64_t)image_width, 4}, means2d.options());
    T alphas = torch::empty({1, (int64_t)image_height, (int64_t)image_width, 1}, means2d.options());
    GSB(gsb_gs_rasterize_ges_fwd(e->h, (int)N, F(means2d), F(conics), F(colors), F(opacities), F(ref_depth_map), delta_depth,
                                 tile_offsets.data_ptr<int>(), flatten_ids.data_ptr<int>(), (int)flatten_ids.numel(), F(render), F(alphas)));
    return {render, alphas, T()};
}

std::tuple<T, T, T, T, T> rasterize_to_pixels_bwd_ges_gs_parallel_tensor(const T &means2d, const T &conics, const T &colors,
                                                                         const T &opacities, const T &radiis, const T &ref_depth_map,
                                                                         const T &base_color_map, const OptT &backgrounds,
                                                                         const uint32_t image_width, const uint32_t image_height,
                                                                         const uint32_t n_isects, const T &group_gs_ids,
                                                                         const T &group_starts, const float delta_depth,
                                                                         const T &render_alphas, const T &v_render_colors,
                                                                         const T &v_render_alphas, bool absgrad)
{
    int64_t N;
    checkSplats(means2d, conics, colors, opacities, N);
    checkI32(radiis, "radiis"), checkF32(ref_depth_map, "ref_depth_map"), checkF32(v_render_colors, "v_render_colors"),
        checkF32(v_render_alphas, "v_render_alphas");
    TORCH_CHECK(!backgrounds.has_value(), "GES rasteriser: backgrounds are not supported");
    TORCH_CHECK(!absgrad, "absgrad is not supported (the SLAM path passes false)");
    const int64_t P = (int64_t)image_width * image_height;
    TORCH_CHECK(radiis.numel() == N && ref_depth_map.numel() == P && v_render_colors.numel() == P * 4 && v_render_alphas.numel() == P,
                "radiis [1,N], ref_depth_map [1,H,W], v_render_colors [1,H,W,4], v_render_alphas [1,H,W,1]");
    const at::cuda::OptionalCUDAGuard guard(device_of(means2d));
    Engine *e = engineForSize(deviceOf(means2d), image_width, image_height);
    T v_means2d = torch::empty_like(means2d), v_conics = torch::empty_like(conics), v_colors = torch::empty_like(colors),
      v_opacities = torch::empty_like(opacities);
    if (N > 0)
        GSB(gsb_gs_rasterize_ges_bwd(e->h, (int)N, F(means2d), F(conics), F(colors), F(opacities), radiis.data_ptr<int>(), F(ref_depth_map),
                                     delta_depth, F(v_render_colors), F(v_render_alphas), F(v_means2d), F(v_conics), F(v_colors),
                                     F(v_opacities)));
    return {T(), v_means2d, v_conics, v_colors, v_opacities};
}

std::tuple<T, T, T> rasterize_to_pixels_fwd_tensor(const T &means2d, const T &conics, const T &colors, const T &opacities,
                                                   const OptT &backgrounds, const OptT &mask, const uint32_t image_width,
                                                   const uint32_t image_height, const uint32_t tile_size, const T &tile_offsets,
                                                   const T &flatten_ids)
{
    int64_t N;
    checkSplats(means2d, conics, colors, opacities, N);
    checkI32(tile_offsets, "tile_offsets"), checkI32(flatten_ids, "flatten_ids");
    TORCH_CHECK(!mask.has_value(), "tile masks are not supported");
    TORCH_CHECK(tile_size == 16, "tile_size must be 16");
    if (backgrounds.has_value())
    {
        checkF32(*backgrounds, "backgrounds");
        TORCH_CHECK(backgrounds->numel() == 4, "backgrounds must be [1,4]");
    }
    const at::cuda::OptionalCUDAGuard guard(device_of(means2d));
    Engine *e = engineForSize(deviceOf(means2d), image_width, image_height);
    TORCH_CHECK(tile_offsets.numel() == (int64_t)e->tileW * e->tileH, "tile_offsets must be [1,tile_height,tile_width]");
    T render = torch::empty({1, (int64_t)image_height, (int64_t)image_width, 4}, means2d.options());
    T alphas = torch::empty({1, (int64_t)image_height, (int64_t)image_width, 1}, means2d.options());
    T lastIds = torch::empty({1, (int64_t)image_height, (int64_t)image_width}, tile_offsets.options());
    GSB(gsb_gs_rasterize_fwd(e->h, (int)N, F(means2d), F(conics), F(colors), F(opacities), backgrounds.has_value() ? F(*backgrounds) : nullptr,
                             tile_offsets.data_ptr<int>(), flatten_ids.data_ptr<int>(), (int)flatten_ids.numel(), F(render), F(alphas),
                             lastIds.data_ptr<int>()));
    return {render, alphas, lastIds};
}

std::tuple<T, T, T, T, T> rasterize_to_pixels_bwd_tensor(const T &means2d_, const T &conics_, const T &colors_, const T &opacities_,
                                                         const OptT &backgrounds, const OptT &mask, const uint32_t image_width,
                                                         const uint32_t image_height, const uint32_t tile_size, const T &tile_offsets_,
                                                         const T &flatten_ids_, const T &render_alphas, const T &last_ids,
                                                         const T &v_render_colors, const T &v_render_alphas, bool absgrad)
{
    T means2d = means2d_.contiguous(), conics = conics_.contiguous(), colors = colors_.contiguous(), opacities = opacities_.contiguous(),
      tile_offsets = tile_offsets_.contiguous(), flatten_ids = flatten_ids_.contiguous();
    int64_t N;
    checkSplats(means2d, conics, colors, opacities, N);
    checkI32(tile_offsets, "tile_offsets"), checkI32(flatten_ids, "flatten_ids"), checkI32(last_ids, "last_ids");
    checkF32(render_alphas, "render_alphas"), checkF32(v_render_colors, "v_render_colors"), checkF32(v_render_alphas, "v_render_alphas");
    TORCH_CHECK(!mask.has_value(), "tile masks are not supported");
    TORCH_CHECK(!absgrad, "absgrad is not supported (the SLAM path passes false)");
    TORCH_CHECK(tile_size == 16, "tile_size must be 16");
    if (backgrounds.has_value())
        checkF32(*backgrounds, "backgrounds");
    const at::cuda::OptionalCUDAGuard guard(device_of(means2d));
    Engine *e = engineForSize(deviceOf(means2d), image_width, image_height);
    T v_means2d = torch::empty_like(means2d), v_conics = torch::empty_like(conics), v_colors = torch::empty_like(colors),
      v_opacities = torch::empty_like(opacities);
    if (N > 0)
        GSB(gsb_gs_rasterize_bwd(e->h, (int)N, F(means2d), F(conics), F(colors), F(opacities),
                                 backgrounds.has_value() ? F(*backgrounds) : nullptr, tile_offsets.data_ptr<int>(), flatten_ids.data_ptr<int>(),
                                 (int)flatten_ids.numel(), F(render_alphas), last_ids.data_ptr<int>(), F(v_render_colors),
                                 F(v_render_alphas), F(v_means2d), F(v_conics), F(v_colors), F(v_opacities)));
    return {T(), v_means2d, v_conics, v_colors, v_opacities};
}

} // namespace gsplat

// ------------------------------------------------------------------------------------------------ ssim.h / simple_knn.h
std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor> fusedssim(float C1, float C2, torch::Tensor &img1_, torch::Tensor &img2_,
                                                                                  bool train)
{
    Tensor img1 = img1_.contiguous(), img2 = img2_.contiguous();
    checkF32(img1, "img1"), checkF32(img2, "img2");
    TORCH_CHECK(img1.dim() == 4 && img1.sizes() == img2.sizes(), "img1, img2 must be [B,CH,H,W]");
    const at::cuda::OptionalCUDAGuard guard(device_of(img1));
    Tensor map = torch::empty_like(img1);
    Tensor d0 = train ? torch::empty_like(img1) : torch::empty({0}), d1 = train ? torch::empty_like(img1) : torch::empty({0}),
           d2 = train ? torch::empty_like(img1) : torch::empty({0});
    GSB(gsb_gs_ssim_fwd(engineAny(deviceOf(img1))->h, (int)(img1.size(0) * img1.size(1)), (int)img1.size(2), (int)img1.size(3), C1, C2, F(img1),
                        F(img2), F(map), train ? F(d0) : nullptr, train ? F(d1) : nullptr, train ? F(d2) : nullptr));
    return {map, d0, d1, d2};
}

torch::Tensor fusedssim_backward(float, float, torch::Tensor &img1_, torch::Tensor &img2_, torch::Tensor &dL_dmap_, torch::Tensor &dm_dmu1,
                                 torch::Tensor &dm_dsigma1_sq, torch::Tensor &dm_dsigma12)
{
    Tensor img1 = img1_.contiguous(), img2 = img2_.contiguous(), dmap = dL_dmap_.contiguous();
    checkF32(img1, "img1"), checkF32(img2, "img2"), checkF32(dmap, "dL_dmap");
    checkF32(dm_dmu1, "dm_dmu1"), checkF32(dm_dsigma1_sq, "dm_dsigma1_sq"), checkF32(dm_dsigma12, "dm_dsigma12");
    const at::cuda::OptionalCUDAGuard guard(device_of(img1));
    Tensor g = torch::empty_like(img1);
    GSB(gsb_gs_ssim_bwd(engineAny(deviceOf(img1))->h, (int)(img1.size(0) * img1.size(1)), (int)img1.size(2), (int)img1.size(3), F(img1), F(img2),
                        F(dmap), F(dm_dmu1), F(dm_dsigma1_sq), F(dm_dsigma12), F(g)));
    return g;
}

torch::Tensor distCUDA2(const torch::Tensor &points_)
{
    Tensor points = points_.contiguous();
    checkF32(points, "points");
    TORCH_CHECK(points.dim() == 2 && points.size(1) == 3, "points must be [P,3]");
    const at::cuda::OptionalCUDAGuard guard(device_of(points));
    Tensor d = torch::empty({points.size(0)}, points.options());
    if (points.size(0) > 0)
        GSB(gsb_gs_dist_cuda2(engineAny(deviceOf(points))->h, (int)points.size(0), F(points), F(d)));
    return d;
}

// ------------------------------------------------------------------------------------------------ gsplat_wapper.hpp free functions
double getDuration(struct timespec start, struct timespec end)
{
    return 1e3 * (double)(end.tv_sec - start.tv_sec) + 1e-6 * (double)(end.tv_nsec - start.tv_nsec);
}

variable_list isectTiles(torch::Tensor means2d, torch::Tensor radii, torch::Tensor depths, int tile_size, int tile_width, int tile_height, bool sort)
{
    at::optional<Tensor> none;
    auto b = gsplat::isect_tiles_tensor(means2d.contiguous(), radii.contiguous(), depths.contiguous(), none, none, (uint32_t)means2d.size(0),
                                        tile_size, tile_width, tile_height, sort, true);
    return {std::get<0>(b), std::get<1>(b), std::get<2>(b)};
}

torch::Tensor isectOffsetEncode(torch::Tensor isect_ids, int n_cameras, int tile_width, int tile_height)
{
    return gsplat::isect_offset_encode_tensor(isect_ids.contiguous(), n_cameras, tile_width, tile_height);
}

variable_list isectTilesNoDepth(torch::Tensor means2d, torch::Tensor radii, torch::Tensor depths, int tile_size, int tile_width, int tile_height,
                                bool sort)
{
    at::optional<Tensor> none;
    auto b = gsplat::isect_tiles_tensor_no_depth(means2d.contiguous(), radii.contiguous(), depths.contiguous(), none, none,
                                                 (uint32_t)means2d.size(0), tile_size, tile_width, tile_height, sort, true);
    return {std::get<0>(b), std::get<1>(b), std::get<2>(b), std::get<3>(b), std::get<4>(b)};
}

torch::Tensor isectOffsetEncodeNoDepth(torch::Tensor isect_ids, int n_cameras, int tile_width, int tile_height)
{
    return gsplat::isect_offset_encode_tensor_no_depth(isect_ids.contiguous(), n_cameras, tile_width, tile_height);
}

torch::Tensor simpleKNN(torch::Tensor points) { return distCUDA2(points); }

// number of SH bases <-> degree (1, 4, 9, 16 <-> 0..3; anything else maps to degree 4 / 25 bases like the reference)
int degFromSh(int numBases)
{
    for (int d = 0; d < 4; d++)
        if ((d + 1) * (d + 1) == numBases)
            return d;
    return 4;
}

int numShBases(int degree) { return degree >= 0 && degree <= 3 ? (degree + 1) * (degree + 1) : 25; }

static const double kShC0 = 0.28209479177387814;   // Y_0^0
torch::Tensor rgb2sh(const torch::Tensor &rgb) { return (rgb - 0.5) / kShC0; }
torch::Tensor sh2rgb(const torch::Tensor &sh) { return torch::clamp(sh * kShC0 + 0.5, 0.0f, 1.0f); }
