// drop-in for the reference's gsplat/rasterizer/simple_knn.h (distCUDA2, simple_knn.cu:227-239 there; called directly at
// src/raw_gs_param.cpp:28); implemented in ../gsplat_b200.cpp over gsb_gs_dist_cuda2.
#pragma once
#include <torch/all.h>

torch::Tensor distCUDA2(const torch::Tensor &points);   // [P,3] -> [P] mean squared distance to the 3 nearest other points
