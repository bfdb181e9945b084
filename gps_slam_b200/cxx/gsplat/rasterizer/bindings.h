// gps_slam_b200 C++ host layer -- drop-in for the reference's gsplat/rasterizer/bindings.h (declarations at :34-360 there).
// Only the gsplat::*_tensor functions that the autograd wrappers of gsplat_wapper.hpp reach are declared; each keeps the
// reference's signature and tensor conventions (camera dimension C = 1 explicit, fp32 contiguous CUDA tensors, outputs freshly
// allocated torch tensors, errors as c10::Error through TORCH_CHECK) and is implemented in ../gsplat_b200.cpp as a thin
// translation onto the staged C-ABI entry points of include/gpsslam_b200.h.  No kernel lives on this side of the C ABI.
#pragma once
#include <torch/all.h>
#include <tuple>

namespace gsplat
{
typedef torch::Tensor T;
typedef at::optional<torch::Tensor> OptT;

enum CameraModelType { PINHOLE = 0, ORTHO = 1, FISHEYE = 2 };   // bindings.h:37-42; only PINHOLE is on the SLAM path

// fully_fused_projection_fwd.cu:196-273 -> radii [C,N] i32, means2d [C,N,2], depths [C,N], conics [C,N,3], compensations (undefined)
std::tuple<T, T, T, T, T> fully_fused_projection_fwd_tensor(const T &means, const OptT &covars, const OptT &quats, const OptT &scales,
                                                            const T &viewmats, const T &Ks, const uint32_t image_width,
                                                            const uint32_t image_height, const float eps2d, const float near_plane,
                                                            const float far_plane, const float radius_clip, const bool calc_compensations,
                                                            const CameraModelType camera_model);
// fully_fused_projection_bwd.cu:288-403 -> v_means, v_covars (undefined), v_quats, v_scales, v_viewmats (undefined)
std::tuple<T, T, T, T, T> fully_fused_projection_bwd_tensor(const T &means, const OptT &covars, const OptT &quats, const OptT &scales,
                                                            const T &viewmats, const T &Ks, const uint32_t image_width,
                                                            const uint32_t image_height, const float eps2d,
                                                            const CameraModelType camera_model, const T &radii, const T &conics,
                                                            const OptT &compensations, const T &v_means2d, const T &v_depths,
                                                            const T &v_conics, const OptT &v_compensations,
                                                            const bool viewmats_requires_grad);
// isect_tiles.cu:132-430 (bins ordered by tile, depth) -> tiles_per_gauss, isect_ids i64, flatten_ids i32
std::tuple<T, T, T> isect_tiles_tensor(const T &means2d, const T &radii, const T &depths, const OptT &camera_ids, const OptT &gaussian_ids,
                                       const uint32_t C, const uint32_t tile_size, const uint32_t tile_width, const uint32_t tile_height,
                                       const bool sort, const bool double_buffer);
T isect_offset_encode_tensor(const T &isect_ids, const uint32_t C, const uint32_t tile_width, const uint32_t tile_height);
// isect_tiles_no_depth.cu:132-461 (bins ordered by tile, Gaussian id) -> ..., group_gs_ids, group_starts (both EMPTY here: they are
// the work list of the reference's own backward kernel; gsb_gs_rasterize_ges_bwd derives its work items itself)
std::tuple<T, T, T, T, T> isect_tiles_tensor_no_depth(const T &means2d, const T &radii, const T &depths, const OptT &camera_ids,
                                                      const OptT &gaussian_ids, const uint32_t C, const uint32_t tile_size,
                                                      const uint32_t tile_width, const uint32_t tile_height, const bool sort,
                                                      const bool double_buffer);
T isect_offset_encode_tensor_no_depth(const T &isect_ids, const uint32_t C, const uint32_t tile_width, const uint32_t tile_height);
// rasterize_to_pixels_fwd.cu:198-376 -> render_colors [C,H,W,D], render_alphas [C,H,W,1], last_ids [C,H,W]
std::tuple<T, T, T> rasterize_to_pixels_fwd_tensor(const T &means2d, const T &conics, const T &colors, const T &opacities,
                                                   const OptT &backgrounds, const OptT &mask, const uint32_t image_width,
                                                   const uint32_t image_height, const uint32_t tile_size, const T &tile_offsets,
                                                   const T &flatten_ids);
// rasterize_to_pixels_bwd.cu:289-511 -> v_means2d_abs (undefined), v_means2d, v_conics, v_colors, v_opacities
std::tuple<T, T, T, T, T> rasterize_to_pixels_bwd_tensor(const T &means2d, const T &conics, const T &colors, const T &opacities,
                                                         const OptT &backgrounds, const OptT &mask, const uint32_t image_width,
                                                         const uint32_t image_height, const uint32_t tile_size, const T &tile_offsets,
                                                         const T &flatten_ids, const T &render_alphas, const T &last_ids,
                                                         const T &v_render_colors, const T &v_render_alphas, bool absgrad);
// rasterize_to_pixels_fwd_ges.cu:338-407 -> render_colors [C,H,W,4], render_alphas [C,H,W,1], last_ids (undefined: the GES backward
// never reads it)
std::tuple<T, T, T> rasterize_to_pixels_fwd_ges_tensor(const T &means2d, const T &conics, const T &colors, const T &opacities,
                                                       const T &ref_depth_map, const T &base_color_map, const OptT &backgrounds,
                                                       const OptT &mask, const uint32_t image_width, const uint32_t image_height,
                                                       const uint32_t tile_size, const T &tile_offsets, const T &flatten_ids,
                                                       const float delta_depth);
// rasterize_to_pixels_bwd_ges_new_parallel.cu:304-385 -> v_means2d_abs (undefined), v_means2d, v_conics, v_colors, v_opacities
std::tuple<T, T, T, T, T> rasterize_to_pixels_bwd_ges_gs_parallel_tensor(const T &means2d, const T &conics, const T &colors,
                                                                         const T &opacities, const T &radiis, const T &ref_depth_map,
                                                                         const T &base_color_map, const OptT &backgrounds,
                                                                         const uint32_t image_width, const uint32_t image_height,
                                                                         const uint32_t n_isects, const T &group_gs_ids,
                                                                         const T &group_starts, const float delta_depth,
                                                                         const T &render_alphas, const T &v_render_colors,
                                                                         const T &v_render_alphas, bool absgrad);
// compute_sh_fwd.cu:40-72, compute_sh_bwd.cu:56-123
T compute_sh_fwd_tensor(const uint32_t degrees_to_use, const T &dirs, const T &coeffs, const OptT masks);
std::tuple<T, T> compute_sh_bwd_tensor(const uint32_t K, const uint32_t degrees_to_use, const T &dirs, const T &coeffs, const OptT masks,
                                       const T &v_colors, bool compute_v_dirs);
} // namespace gsplat
