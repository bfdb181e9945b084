// drop-in for the reference's gsplat/rasterizer/ssim.h (fusedssim / fusedssim_backward, ssim.cu:387-460 there); implemented in
// ../gsplat_b200.cpp over gsb_gs_ssim_fwd / gsb_gs_ssim_bwd.
#pragma once
#include <torch/all.h>
#include <tuple>

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor> fusedssim(float C1, float C2, torch::Tensor &img1, torch::Tensor &img2,
                                                                                  bool train);
torch::Tensor fusedssim_backward(float C1, float C2, torch::Tensor &img1, torch::Tensor &img2, torch::Tensor &dL_dmap, torch::Tensor &dm_dmu1,
                                 torch::Tensor &dm_dsigma1_sq, torch::Tensor &dm_dsigma12);
