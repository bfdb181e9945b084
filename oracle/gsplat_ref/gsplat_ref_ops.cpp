// TEST INFRASTRUCTURE (oracle/): exposes the reference's own gsplat autograd wrappers (reference gsplat/gsplat_wapper.hpp,
// compiled where it lies together with gsplat/rasterizer/*.cu) as torch ops, so that oracle/gsplat_ref.py can replay
// RawGaussianModel::gesForward / computeLoss (reference src/raw_gs_model.cpp:188-417) with the reference's real kernels
// and libtorch autograd on the GPU box.  Used to pin oracle/gs_oracle.py and the CUDA engine against the reference;
// never linked into, imported by or shipped with the product library.
#include "gsplat_wapper.hpp"
#include <torch/library.h>

namespace {

using torch::Tensor;
typedef std::vector<Tensor> TensorList;

// gsplat_wapper.hpp:96-244
TensorList fully_fused_projection(Tensor means, Tensor quats, Tensor scales, Tensor viewmats, Tensor Ks, int64_t width, int64_t height,
                                  double eps2d, double near_plane, double far_plane, double radius_clip)
{
    at::optional<Tensor> covars;
    return FullyFusedProjection::apply(means, covars, quats, scales, viewmats, Ks, (int)width, (int)height, (float)eps2d, (float)near_plane,
                                       (float)far_plane, (float)radius_clip, false, std::string("pinhole"));
}

// gsplat_wapper.hpp:16-94
Tensor spherical_harmonics(int64_t degree, Tensor dirs, Tensor coeffs, Tensor masks)
{
    return SphericalHarmonicsNew::apply((int)degree, dirs, coeffs, masks);
}

// gsplat_wapper.cpp:58-88
TensorList isect_tiles_no_depth(Tensor means2d, Tensor radii, Tensor depths, int64_t tile_size, int64_t tile_width, int64_t tile_height)
{
    return isectTilesNoDepth(means2d, radii, depths, (int)tile_size, (int)tile_width, (int)tile_height, true);
}

Tensor isect_offset_encode_no_depth(Tensor isect_ids, int64_t n_cameras, int64_t tile_width, int64_t tile_height)
{
    return isectOffsetEncodeNoDepth(isect_ids, (int)n_cameras, (int)tile_width, (int)tile_height);
}

// gsplat_wapper.cpp:14-50 (depth-sorted variant used by render_method "raw")
TensorList isect_tiles(Tensor means2d, Tensor radii, Tensor depths, int64_t tile_size, int64_t tile_width, int64_t tile_height)
{
    return isectTiles(means2d, radii, depths, (int)tile_size, (int)tile_width, (int)tile_height, true);
}

Tensor isect_offset_encode(Tensor isect_ids, int64_t n_cameras, int64_t tile_width, int64_t tile_height)
{
    return isectOffsetEncode(isect_ids, (int)n_cameras, (int)tile_width, (int)tile_height);
}

// gsplat_wapper.hpp:489-620
TensorList rasterize_ges(Tensor means2d, Tensor conics, Tensor colors, Tensor opacities, Tensor radiis, Tensor ref_depth_map,
                         Tensor base_color_map, int64_t width, int64_t height, int64_t tile_size, Tensor isect_offsets, Tensor flatten_ids,
                         Tensor group_gs_ids, Tensor group_starts, bool absgrad, double delta_depth)
{
    at::optional<Tensor> none;
    return RasterizeToPixelsGes_NewParallel::apply(means2d, conics, colors, opacities, radiis, ref_depth_map, base_color_map, none, none,
                                                   (int)width, (int)height, (int)tile_size, isect_offsets, flatten_ids, group_gs_ids,
                                                   group_starts, absgrad, (float)delta_depth);
}

// gsplat_wapper.hpp:243-352 (front-to-back alpha compositing, render_method "raw")
TensorList rasterize_raw(Tensor means2d, Tensor conics, Tensor colors, Tensor opacities, int64_t width, int64_t height, int64_t tile_size,
                         Tensor isect_offsets, Tensor flatten_ids, bool absgrad)
{
    at::optional<Tensor> none;
    return RasterizeToPixels::apply(means2d, conics, colors, opacities, none, none, (int)width, (int)height, (int)tile_size, isect_offsets,
                                    flatten_ids, absgrad);
}

TensorList rasterize_raw_bg(Tensor means2d, Tensor conics, Tensor colors, Tensor opacities, Tensor backgrounds, int64_t width, int64_t height,
                            int64_t tile_size, Tensor isect_offsets, Tensor flatten_ids, bool absgrad)
{
    at::optional<Tensor> none, bg = backgrounds;
    return RasterizeToPixels::apply(means2d, conics, colors, opacities, bg, none, (int)width, (int)height, (int)tile_size, isect_offsets,
                                    flatten_ids, absgrad);
}

// the forward kernel alone (returns last_ids too): rasterize_to_pixels_fwd_ges.cu:338-407
TensorList rasterize_ges_fwd(Tensor means2d, Tensor conics, Tensor colors, Tensor opacities, Tensor ref_depth_map, Tensor base_color_map,
                             int64_t width, int64_t height, int64_t tile_size, Tensor isect_offsets, Tensor flatten_ids, double delta_depth)
{
    at::optional<Tensor> none;
    auto r = gsplat::rasterize_to_pixels_fwd_ges_tensor(means2d.contiguous(), conics.contiguous(), colors.contiguous(), opacities.contiguous(),
                                                        ref_depth_map.contiguous(), base_color_map.contiguous(), none, none, (uint32_t)width,
                                                        (uint32_t)height, (uint32_t)tile_size, isect_offsets.contiguous(),
                                                        flatten_ids.contiguous(), (float)delta_depth);
    return {std::get<0>(r), std::get<1>(r), std::get<2>(r)};
}

// the backward kernel alone, WITH the background term (rasterize_to_pixels_bwd.cu:289-511).  The autograd wrapper above never passes
// the background to it (gsplat_wapper.hpp:300-312: `backgrounds` is a fresh empty optional in backward), so the wrapper's gradients
// are those of the background-free composite; this entry exposes what the kernel computes when it is given one.
TensorList rasterize_raw_bwd(Tensor means2d, Tensor conics, Tensor colors, Tensor opacities, Tensor backgrounds, int64_t width, int64_t height,
                             int64_t tile_size, Tensor isect_offsets, Tensor flatten_ids, Tensor render_alphas, Tensor last_ids,
                             Tensor v_render_colors, Tensor v_render_alphas)
{
    at::optional<Tensor> none, bg = backgrounds.contiguous();
    auto r = gsplat::rasterize_to_pixels_bwd_tensor(means2d.contiguous(), conics.contiguous(), colors.contiguous(), opacities.contiguous(), bg, none,
                                                    (uint32_t)width, (uint32_t)height, (uint32_t)tile_size, isect_offsets.contiguous(),
                                                    flatten_ids.contiguous(), render_alphas.contiguous(), last_ids.contiguous(),
                                                    v_render_colors.contiguous(), v_render_alphas.contiguous(), false);
    return {std::get<1>(r), std::get<2>(r), std::get<3>(r), std::get<4>(r)};
}

// the forward kernel alone (returns last_ids too): rasterize_to_pixels_fwd.cu:198-376
TensorList rasterize_raw_fwd(Tensor means2d, Tensor conics, Tensor colors, Tensor opacities, Tensor backgrounds, int64_t width, int64_t height,
                             int64_t tile_size, Tensor isect_offsets, Tensor flatten_ids)
{
    at::optional<Tensor> none, bg = backgrounds.contiguous();
    auto r = gsplat::rasterize_to_pixels_fwd_tensor(means2d.contiguous(), conics.contiguous(), colors.contiguous(), opacities.contiguous(), bg, none,
                                                    (uint32_t)width, (uint32_t)height, (uint32_t)tile_size, isect_offsets.contiguous(),
                                                    flatten_ids.contiguous());
    return {std::get<0>(r), std::get<1>(r), std::get<2>(r)};
}

// gsplat_wapper.hpp:622-676
Tensor fused_ssim_map(double C1, double C2, Tensor img1, Tensor img2, std::string padding, bool train)
{
    return FusedSSIMMap::apply((float)C1, (float)C2, img1, img2, padding, train);
}

// rasterizer/simple_knn.cu:227-239 (called directly at src/raw_gs_param.cpp:28)
Tensor simple_knn(Tensor points) { return simpleKNN(points); }

} // namespace

TORCH_LIBRARY(gsplat_ref, m)
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
    m.def("rasterize_raw_fwd", &rasterize_raw_fwd);
    m.def("rasterize_raw_bwd", &rasterize_raw_bwd);
    m.def("fused_ssim_map", &fused_ssim_map);
    m.def("simple_knn", &simple_knn);
}
