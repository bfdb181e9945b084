// gps_slam_b200 C++ host layer: fused SSIM entry points under the names the reference's gsplat_wapper.hpp calls (FusedSSIMMap::forward /
// backward there); bodies in ../gsplat_b200.cpp over gsb_gs_ssim_fwd / gsb_gs_ssim_bwd of the C ABI.
//   fusedssim          img1, img2 [B,CH,H,W] -> (ssim_map, dm_dmu1, dm_dsigma1_sq, dm_dsigma12); the three derivative maps are empty
//                      tensors when train == false
//   fusedssim_backward dL/d(ssim_map) + the derivative maps -> dL/d(img1)
#pragma once
#include <torch/all.h>

#include <tuple>

namespace gsb_shim
{
typedef torch::Tensor Map;
typedef std::tuple<Map, Map, Map, Map> SsimMaps;
} // namespace gsb_shim

gsb_shim::SsimMaps fusedssim(float C1, float C2, gsb_shim::Map &img1, gsb_shim::Map &img2, bool train);
gsb_shim::Map fusedssim_backward(float C1, float C2, gsb_shim::Map &img1, gsb_shim::Map &img2, gsb_shim::Map &dL_dmap, gsb_shim::Map &dm_dmu1,
                                 gsb_shim::Map &dm_dsigma1_sq, gsb_shim::Map &dm_dsigma12);
