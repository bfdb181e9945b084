// gps_slam_b200 C++ host layer -- drop-in for the reference's gsplat/gsplat_wapper.hpp (the header include/raw_gs_param.h:4 pulls in).
// Same class names, the same static forward(AutogradContext*, ...) argument lists and the same number / order of backward results as
// the reference (forward lists at gsplat_wapper.hpp:19-23, 100-115, 246-259, 358-374, 492-511, 625-631 there), so that
// RawGaussianModel::{rawForward,gesForward,computeLoss} (src/raw_gs_model.cpp:37-417) and RawGaussianParams::init
// (src/raw_gs_param.cpp:28) compile against it unchanged.  Underneath, every class calls the gsplat::*_tensor functions of
// rasterizer/bindings.h, which forward to the staged C-ABI entry points of libgpsslam_b200.so (include/gpsslam_b200.h).
//
// Behaviour kept: inputs made contiguous; camera dimension C = 1 explicit; backward returns one entry per forward argument with
// an undefined Tensor for everything non-differentiable; errors are C++ exceptions.  Not carried over: the LOGBACKWARDTIME
// printouts, `absgrad` (the SLAM path always passes false; true raises), tile masks, packed / ortho / fisheye variants.
#ifndef GSPLAT_NEW_WAPPER_H
#define GSPLAT_NEW_WAPPER_H

#include "rasterizer/bindings.h"
#include "rasterizer/simple_knn.h"
#include "rasterizer/ssim.h"

#include <torch/torch.h>
using namespace torch::autograd;

double getDuration(struct timespec start, struct timespec end);   // milliseconds between two CLOCK_MONOTONIC stamps

namespace gsb_shim
{
inline variable_list undefined(size_t n) { return variable_list(n); }
inline gsplat::CameraModelType cameraModel(const std::string &name)
{
    if (name == "pinhole") return gsplat::PINHOLE;
    if (name == "ortho") return gsplat::ORTHO;
    if (name == "fisheye") return gsplat::FISHEYE;
    throw std::runtime_error("Unknown camera model");
}
} // namespace gsb_shim

// colours = SH(dirs; coeffs) for visible Gaussians.  forward(ctx, sh_degree, dirs [..,3], coeffs [..,K,3], masks [..])
class SphericalHarmonicsNew : public Function<SphericalHarmonicsNew>
{
public:
    static torch::Tensor forward(AutogradContext *ctx, int sh_degree, torch::Tensor dirs, torch::Tensor coeffs, torch::Tensor masks)
    {
        ctx->save_for_backward({dirs, coeffs, masks});
        ctx->saved_data["sh_degree"] = sh_degree;
        ctx->saved_data["num_bases"] = coeffs.size(-2);
        return gsplat::compute_sh_fwd_tensor(sh_degree, dirs, coeffs, masks);
    }
    static tensor_list backward(AutogradContext *ctx, tensor_list grad_outputs)
    {
        const auto kept = ctx->get_saved_variables();
        const bool wantDirs = ctx->needs_input_grad(1);
        auto g = gsplat::compute_sh_bwd_tensor(ctx->saved_data["num_bases"].toInt(), ctx->saved_data["sh_degree"].toInt(), kept[0], kept[1],
                                               kept[2], grad_outputs[0].contiguous(), wantDirs);
        return {torch::Tensor(), wantDirs ? std::get<1>(g) : torch::Tensor(), std::get<0>(g), torch::Tensor()};
    }
};

// world Gaussians -> screen-space splats.  Returns {radii, means2d, depths, conics, compensations}.
class FullyFusedProjection : public Function<FullyFusedProjection>
{
public:
    static variable_list forward(AutogradContext *ctx, torch::Tensor means, at::optional<torch::Tensor> covars, torch::Tensor quats,
                                 torch::Tensor scales, torch::Tensor viewmats, torch::Tensor Ks, int width, int height, float eps2d,
                                 float near_plane, float far_plane, float radius_clip, bool calc_compensations, std::string camera_model)
    {
        const gsplat::CameraModelType model = gsb_shim::cameraModel(camera_model);
        means = means.contiguous(), quats = quats.contiguous(), scales = scales.contiguous(), viewmats = viewmats.contiguous();
        auto out = gsplat::fully_fused_projection_fwd_tensor(means, covars, quats, scales, viewmats, Ks, width, height, eps2d, near_plane,
                                                             far_plane, radius_clip, calc_compensations, model);
        ctx->save_for_backward({means, quats, scales, viewmats, Ks, std::get<0>(out), std::get<3>(out)});
        ctx->saved_data["width"] = width;
        ctx->saved_data["height"] = height;
        ctx->saved_data["eps2d"] = eps2d;
        ctx->saved_data["camera_model_type"] = (int)model;
        torch::Tensor comp = calc_compensations ? std::get<4>(out) : torch::zeros({1}, means.options());
        return {std::get<0>(out), std::get<1>(out), std::get<2>(out), std::get<3>(out), comp};
    }
    static variable_list backward(AutogradContext *ctx, variable_list grad_outputs)
    {
        const auto kept = ctx->get_saved_variables();   // means, quats, scales, viewmats, Ks, radii, conics
        at::optional<torch::Tensor> none;
        auto g = gsplat::fully_fused_projection_bwd_tensor(
            kept[0], none, kept[1], kept[2], kept[3], kept[4], ctx->saved_data["width"].toInt(), ctx->saved_data["height"].toInt(),
            (float)ctx->saved_data["eps2d"].toDouble(), (gsplat::CameraModelType)ctx->saved_data["camera_model_type"].toInt(), kept[5],
            kept[6], none, grad_outputs[1].contiguous(), grad_outputs[2].contiguous(), grad_outputs[3].contiguous(), none,
            ctx->needs_input_grad(4));
        variable_list r = gsb_shim::undefined(14);   // covars and viewmats never receive a gradient, as in the reference
        r[0] = std::get<0>(g), r[2] = std::get<2>(g), r[3] = std::get<3>(g);
        return r;
    }
};

// depth-sorted front-to-back alpha compositing (render_method "raw").  Returns {render_colors, render_alphas}.
class RasterizeToPixels : public Function<RasterizeToPixels>
{
public:
    static variable_list forward(AutogradContext *ctx, torch::Tensor means2d, torch::Tensor conics, torch::Tensor colors,
                                 torch::Tensor opacities, at::optional<torch::Tensor> backgrounds, at::optional<torch::Tensor> masks,
                                 int width, int height, int tile_size, torch::Tensor isect_offsets, torch::Tensor flatten_ids, bool absgrad)
    {
        if (backgrounds.has_value()) backgrounds = backgrounds->contiguous();
        if (masks.has_value()) masks = masks->contiguous();
        auto out = gsplat::rasterize_to_pixels_fwd_tensor(means2d.contiguous(), conics.contiguous(), colors.contiguous(),
                                                          opacities.contiguous(), backgrounds, masks, width, height, tile_size,
                                                          isect_offsets.contiguous(), flatten_ids.contiguous());
        // like the reference, the background is not kept: the splat gradients are those of the background-free composite
        ctx->save_for_backward({means2d, conics, colors, opacities, isect_offsets, flatten_ids, std::get<1>(out), std::get<2>(out)});
        ctx->saved_data["width"] = width;
        ctx->saved_data["height"] = height;
        ctx->saved_data["tile_size"] = tile_size;
        ctx->saved_data["absgrad"] = absgrad;
        return {std::get<0>(out), std::get<1>(out)};
    }
    static variable_list backward(AutogradContext *ctx, variable_list grad_outputs)
    {
        const auto kept = ctx->get_saved_variables();
        at::optional<torch::Tensor> none;
        auto g = gsplat::rasterize_to_pixels_bwd_tensor(kept[0], kept[1], kept[2], kept[3], none, none, ctx->saved_data["width"].toInt(),
                                                        ctx->saved_data["height"].toInt(), ctx->saved_data["tile_size"].toInt(), kept[4],
                                                        kept[5], kept[6], kept[7], grad_outputs[0].contiguous(),
                                                        grad_outputs[1].contiguous(), ctx->saved_data["absgrad"].toBool());
        variable_list r = gsb_shim::undefined(12);
        r[0] = std::get<1>(g), r[1] = std::get<2>(g), r[2] = std::get<3>(g), r[3] = std::get<4>(g);
        if (ctx->needs_input_grad(4))   // d/d background = sum over pixels of v_colour * (1 - alpha)
            r[4] = (grad_outputs[0] * (1.0 - kept[6]).to(torch::kFloat)).sum({1, 2});
        return r;
    }
};

// GES blend (order-independent weighted sum inside the depth window of the TSDF render), tile-parallel backward variant.
// The reference never instantiates this class (src/raw_gs_model.cpp:291 uses the _NewParallel one); the forward is provided,
// the backward raises.
class RasterizeToPixelsGes : public Function<RasterizeToPixelsGes>
{
public:
    static variable_list forward(AutogradContext *ctx, torch::Tensor means2d, torch::Tensor conics, torch::Tensor colors,
                                 torch::Tensor opacities, torch::Tensor ref_depth_map, torch::Tensor base_color_map,
                                 at::optional<torch::Tensor> backgrounds, at::optional<torch::Tensor> masks, int width, int height,
                                 int tile_size, torch::Tensor isect_offsets, torch::Tensor flatten_ids, bool absgrad, float delta_depth)
    {
        auto out = gsplat::rasterize_to_pixels_fwd_ges_tensor(means2d.contiguous(), conics.contiguous(), colors.contiguous(),
                                                              opacities.contiguous(), ref_depth_map.contiguous(),
                                                              base_color_map.contiguous(), backgrounds, masks, width, height, tile_size,
                                                              isect_offsets.contiguous(), flatten_ids.contiguous(), delta_depth);
        return {std::get<0>(out), std::get<1>(out)};
    }
    static variable_list backward(AutogradContext *, variable_list)
    {
        throw std::runtime_error("RasterizeToPixelsGes::backward is not provided by gps_slam_b200 (unused by the reference); "
                                 "use RasterizeToPixelsGes_NewParallel");
    }
};

// GES blend with the Gaussian-parallel backward: what RawGaussianModel::gesForward uses.  Returns {render_colors, render_alphas}.
class RasterizeToPixelsGes_NewParallel : public Function<RasterizeToPixelsGes_NewParallel>
{
public:
    static variable_list forward(AutogradContext *ctx, torch::Tensor means2d, torch::Tensor conics, torch::Tensor colors,
                                 torch::Tensor opacities, torch::Tensor radiis, torch::Tensor ref_depth_map, torch::Tensor base_color_map,
                                 at::optional<torch::Tensor> backgrounds, at::optional<torch::Tensor> masks, int width, int height,
                                 int tile_size, torch::Tensor isect_offsets, torch::Tensor flatten_ids, torch::Tensor group_gs_ids,
                                 torch::Tensor group_starts, bool absgrad, float delta_depth)
    {
        if (backgrounds.has_value()) backgrounds = backgrounds->contiguous();
        if (masks.has_value()) masks = masks->contiguous();
        auto out = gsplat::rasterize_to_pixels_fwd_ges_tensor(means2d.contiguous(), conics.contiguous(), colors.contiguous(),
                                                              opacities.contiguous(), ref_depth_map.contiguous(),
                                                              base_color_map.contiguous(), backgrounds, masks, width, height, tile_size,
                                                              isect_offsets.contiguous(), flatten_ids.contiguous(), delta_depth);
        ctx->save_for_backward({means2d, conics, colors, opacities, radiis, ref_depth_map, base_color_map, std::get<1>(out), group_gs_ids,
                                group_starts});
        ctx->saved_data["width"] = width;
        ctx->saved_data["height"] = height;
        ctx->saved_data["absgrad"] = absgrad;
        ctx->saved_data["delta_depth"] = delta_depth;
        ctx->saved_data["n_isects"] = flatten_ids.size(0);
        return {std::get<0>(out), std::get<1>(out)};
    }
    static variable_list backward(AutogradContext *ctx, variable_list grad_outputs)
    {
        const auto kept = ctx->get_saved_variables();
        at::optional<torch::Tensor> none;
        auto g = gsplat::rasterize_to_pixels_bwd_ges_gs_parallel_tensor(
            kept[0].contiguous(), kept[1].contiguous(), kept[2].contiguous(), kept[3].contiguous(), kept[4].contiguous(),
            kept[5].contiguous(), kept[6], none, ctx->saved_data["width"].toInt(), ctx->saved_data["height"].toInt(),
            ctx->saved_data["n_isects"].toInt(), kept[8], kept[9], (float)ctx->saved_data["delta_depth"].toDouble(), kept[7],
            grad_outputs[0].contiguous(), grad_outputs[1].contiguous(), ctx->saved_data["absgrad"].toBool());
        variable_list r = gsb_shim::undefined(18);
        r[0] = std::get<1>(g), r[1] = std::get<2>(g), r[2] = std::get<3>(g), r[3] = std::get<4>(g);
        return r;
    }
};

// per-pixel SSIM map of img1 against img2 (NCHW) with an 11x11 Gaussian window; gradient w.r.t. img1 only.
class FusedSSIMMap : public Function<FusedSSIMMap>
{
public:
    static torch::Tensor forward(AutogradContext *ctx, float C1, float C2, torch::Tensor img1, torch::Tensor img2, std::string padding,
                                 bool train)
    {
        auto out = fusedssim(C1, C2, img1, img2, train);
        ctx->save_for_backward({img1, img2, std::get<1>(out), std::get<2>(out), std::get<3>(out)});
        ctx->saved_data["C1"] = C1;
        ctx->saved_data["C2"] = C2;
        ctx->saved_data["padding"] = padding;
        torch::Tensor map = std::get<0>(out);
        return padding == "valid" ? map.slice(2, 5, -5).slice(3, 5, -5) : map;
    }
    static tensor_list backward(AutogradContext *ctx, tensor_list grad_outputs)
    {
        auto kept = ctx->get_saved_variables();
        torch::Tensor dmap = grad_outputs[0];
        if (ctx->saved_data["padding"].toStringRef() == "valid")
        {
            torch::Tensor full = torch::zeros_like(kept[0]);
            full.slice(2, 5, -5).slice(3, 5, -5).copy_(dmap);
            dmap = full;
        }
        dmap = dmap.contiguous();
        torch::Tensor g = fusedssim_backward((float)ctx->saved_data["C1"].toDouble(), (float)ctx->saved_data["C2"].toDouble(), kept[0],
                                             kept[1], dmap, kept[2], kept[3], kept[4]);
        return {torch::Tensor(), torch::Tensor(), g, torch::Tensor(), torch::Tensor(), torch::Tensor()};
    }
};

// tile binning: {tiles_per_gauss, isect_ids, flatten_ids}; bins ordered by (tile, depth)
variable_list isectTiles(torch::Tensor means2d, torch::Tensor radii, torch::Tensor depths, int tile_size, int tile_width, int tile_height,
                         bool sort = true);
torch::Tensor isectOffsetEncode(torch::Tensor isect_ids, int n_cameras, int tile_width, int tile_height);
// {tiles_per_gauss, isect_ids, flatten_ids, group_gs_ids, group_starts}; bins ordered by (tile, Gaussian id)
variable_list isectTilesNoDepth(torch::Tensor means2d, torch::Tensor radii, torch::Tensor depths, int tile_size, int tile_width,
                                int tile_height, bool sort = true);
torch::Tensor isectOffsetEncodeNoDepth(torch::Tensor isect_ids, int n_cameras, int tile_width, int tile_height);
torch::Tensor simpleKNN(torch::Tensor points);

int degFromSh(int numBases);
int numShBases(int degree);
torch::Tensor rgb2sh(const torch::Tensor &rgb);
torch::Tensor sh2rgb(const torch::Tensor &sh);

#endif
