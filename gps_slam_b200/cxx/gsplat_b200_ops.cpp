// Registers the C++ host layer (gsplat/gsplat_wapper.hpp over libgpsslam_b200.so) as torch ops under the namespace gsplat_b200, so
// that Python can drive the very classes a C++ caller such as RawGaussianModel::gesForward (reference src/raw_gs_model.cpp:188-367)
// would use -- FullyFusedProjection::apply, SphericalHarmonicsNew::apply, isectTilesNoDepth, RasterizeToPixelsGes_NewParallel::apply,
// ... -- with libtorch autograd.  The op names and schemas equal those oracle/gsplat_ref registers for the reference's own
// wrappers (namespace gsplat_ref), which is what lets tests/test_cxx_shim_gpu.py run one script against both.
//   torch.ops.load_library("gps_slam_b200/libgsplat_b200_torch.so"); torch.ops.gsplat_b200.fully_fused_projection(...)
#include "gsplat/gsplat_wapper.hpp"

#include <torch/library.h>

namespace
{
using torch::Tensor;
typedef std::vector<Tensor> Tensors;
const at::optional<Tensor> kNone;

Tensors fully_fused_projection(Tensor means, Tensor quats, Tensor scales, Tensor viewmats, Tensor Ks, int64_t width, int64_t height,
                               double eps2d, double near_plane, double far_plane, double radius_clip)
{
    return FullyFusedProjection::apply(means, kNone, quats, scales, viewmats, Ks, (int)width, (int)height, (float)eps2d, (float)near_plane,
                                       (float)far_plane, (float)radius_clip, false, std::string("pinhole"));
}

Tensor spherical_harmonics(int64_t degree, Tensor dirs, Tensor coeffs, Tensor masks)
{
    return SphericalHarmonicsNew::apply((int)degree, dirs, coeffs, masks);
}

Tensors isect_tiles_no_depth(Tensor means2d, Tensor radii, Tensor depths, int64_t tile_size, int64_t tile_width, int64_t tile_height)
{
    return isectTilesNoDepth(means2d, radii, depths, (int)tile_size, (int)tile_width, (int)tile_height, true);
}

Tensor isect_offset_encode_no_depth(Tensor isect_ids, int64_t n_cameras, int64_t tile_width, int64_t tile_height)
{
    return isectOffsetEncodeNoDepth(isect_ids, (int)n_cameras, (int)tile_width, (int)tile_height);
}

Tensors isect_tiles(Tensor means2d, Tensor radii, Tensor depths, int64_t tile_size, int64_t tile_width, int64_t tile_height)
{
    return isectTiles(means2d, radii, depths, (int)tile_size, (int)tile_width, (int)tile_height, true);
}

Tensor isect_offset_encode(Tensor isect_ids, int64_t n_cameras, int64_t tile_width, int64_t tile_height)
{
    return isectOffsetEncode(isect_ids, (int)n_cameras, (int)tile_width, (int)tile_height);
}

Tensors rasterize_ges(Tensor means2d, Tensor conics, Tensor colors, Tensor opacities, Tensor radiis, Tensor ref_depth_map, Tensor base_color_map,
                      int64_t width, int64_t height, int64_t tile_size, Tensor isect_offsets, Tensor flatten_ids, Tensor group_gs_ids,
                      Tensor group_starts, bool absgrad, double delta_depth)
{
    return RasterizeToPixelsGes_NewParallel::apply(means2d, conics, colors, opacities, radiis, ref_depth_map, base_color_map, kNone, kNone,
                                                   (int)width, (int)height, (int)tile_size, isect_offsets, flatten_ids, group_gs_ids,
                                                   group_starts, absgrad, (float)delta_depth);
}

Tensors rasterize_ges_fwd(Tensor means2d, Tensor conics, Tensor colors, Tensor opacities, Tensor ref_depth_map, Tensor base_color_map,
                          int64_t width, int64_t height, int64_t tile_size, Tensor isect_offsets, Tensor flatten_ids, double delta_depth)
{
    auto r = gsplat::rasterize_to_pixels_fwd_ges_tensor(means2d.contiguous(), conics.contiguous(), colors.contiguous(), opacities.contiguous(),
                                                        ref_depth_map.contiguous(), base_color_map.contiguous(), kNone, kNone, (uint32_t)width,
                                                        (uint32_t)height, (uint32_t)tile_size, isect_offsets.contiguous(),
                                                        flatten_ids.contiguous(), (float)delta_depth);
    return {std::get<0>(r), std::get<1>(r)};
}

Tensors rasterize_raw(Tensor means2d, Tensor conics, Tensor colors, Tensor opacities, int64_t width, int64_t height, int64_t tile_size,
                      Tensor isect_offsets, Tensor flatten_ids, bool absgrad)
{
    return RasterizeToPixels::apply(means2d, conics, colors, opacities, kNone, kNone, (int)width, (int)height, (int)tile_size, isect_offsets,
                                    flatten_ids, absgrad);
}

Tensors rasterize_raw_bg(Tensor means2d, Tensor conics, Tensor colors, Tensor opacities, Tensor backgrounds, int64_t width, int64_t height,
                         int64_t tile_size, Tensor isect_offsets, Tensor flatten_ids, bool absgrad)
{
    return RasterizeToPixels::apply(means2d, conics, colors, opacities, at::optional<Tensor>(backgrounds), kNone, (int)width, (int)height,
                                    (int)tile_size, isect_offsets, flatten_ids, absgrad);
}

Tensor fused_ssim_map(double C1, double C2, Tensor img1, Tensor img2, std::string padding, bool train)
{
    return FusedSSIMMap::apply((float)C1, (float)C2, img1, img2, padding, train);
}

Tensor simple_knn(Tensor points) { return simpleKNN(points); }

// RawGaussianParams::saveTensor / loadTensor (reference src/raw_gs_param.cpp:220-254): the "model.pt" checkpoint is a
// torch::serialize archive with one named tensor per parameter.  Written with the same libtorch API, so the container format is the
// reference's by construction; tensors are stored as given (the reference stores its CUDA tensors).
const char *const kModelKeys[7] = {"means", "scales", "quats", "featuresDc", "featuresRest", "opacities", "exposure"};

void save_model_pt(std::string filename, Tensors params)
{
    TORCH_CHECK(params.size() == 7, "expected means, scales, quats, featuresDc, featuresRest, opacities, exposure");
    torch::serialize::OutputArchive archive;
    for (int i = 0; i < 7; i++)
        archive.write(kModelKeys[i], params[i]);
    archive.save_to(filename);
}

Tensors load_model_pt(std::string filename)
{
    torch::serialize::InputArchive archive;
    archive.load_from(filename);
    Tensors params(7);
    for (int i = 0; i < 7; i++)
        archive.read(kModelKeys[i], params[i]);
    return params;
}

} // namespace

TORCH_LIBRARY(gsplat_b200, m)
{
    m.def("fully_fused_projection", &fully_fused_projection);
    m.def("spherical_harmonics", &spherical_harmonics);
    m.def("isect_tiles_no_depth", &isect_tiles_no_depth);
    m.def("isect_offset_encode_no_depth", &isect_offset_encode_no_depth);
    m.def("isect_tiles", &isect_tiles);
    m.def("isect_offset_encode", &isect_offset_encode);
    m.def("rasterize_ges", &rasterize_ges);
    m.def("rasterize_ges_fwd", &rasterize_ges_fwd);
    m.def("rasterize_raw", &rasterize_raw);
    m.def("rasterize_raw_bg", &rasterize_raw_bg);
    m.def("fused_ssim_map", &fused_ssim_map);
    m.def("simple_knn", &simple_knn);
    m.def("save_model_pt", &save_model_pt);
    m.def("load_model_pt", &load_model_pt);
}
