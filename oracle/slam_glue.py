"""TEST INFRASTRUCTURE ONLY -- CPU restatement (torch CPU fp32, the same ATen arithmetic the reference's libtorch calls run) of the
tensor glue between the TSDF engine and the Gaussian model, and of the Gaussian spawn.  Never imported by gps_slam_b200.

Follows, function by function:
  uchar4_image_to_tensor / float4_image_to_tensor   src/cv_utils.cpp:322-341 (ITMUChar4ImageToTensor, ITMUFloat4ImageToTensor)
  pose_inv, vertices_transform                       src/tensor_math.cpp:56-81
  raycast_maps                                       slam/slam_pipeline.cpp:386-403 (runRaycastByCam, use_cam_depth = false)
  frame_to_float                                     src/dataset_reader.cpp:269-369 (image / 255, depth / 1000) + Camera::toGPU
  feature_gradient, compute_normal_map               src/tensor_math.cpp:217-248, 278-300
  sample_mask                                        slam/slam_pipeline.cpp:450-503 (initNewGaussians)
  compute_quat, quaternion_from_axis_angle           src/tensor_math.cpp:184-201
  dist2_knn3                                         gsplat/rasterizer/simple_knn.cu:151-239 (distCUDA2: mean squared distance to the 3 nearest)
  init_params                                        src/raw_gs_param.cpp:11-74 (RawGaussianParams::init), gsplat/gsplat_wapper.cpp:127-133 (rgb2sh)
Parity pin: there is no golden vector for these in the reference; the restatement is torch code of the same operator sequence
(the reference's own source is torch code), and distCUDA2 is pinned to the reference's CUDA kernel in tests/test_gs_raw_gpu.py.
"""
import numpy as np
import torch

C0 = 0.28209479177387814


def uchar4_image_to_tensor(img_u8):
    """[H,W,4] uint8 -> [H,W,3] float (clone().to(kFloat).div(255.0).slice(2,0,3))"""
    return torch.from_numpy(np.ascontiguousarray(img_u8)).to(torch.float32).div(255.0)[..., :3].contiguous()


def float4_image_to_tensor(img_f4):
    """[H,W,4] float -> cat(value * (confidence > 0), confidence)"""
    t = torch.from_numpy(np.ascontiguousarray(img_f4)).clone()
    value, conf = t[..., :3], t[..., 3:4]
    return torch.cat([value * conf.gt(0), conf], 2).contiguous()


def pose_inv(c2w):
    c2w = torch.as_tensor(np.asarray(c2w, np.float32))
    R, T = c2w[:3, :3], c2w[:3, 3:4]
    Rinv = R.transpose(0, 1)
    out = torch.eye(4)
    out[:3, :3] = Rinv
    out[:3, 3:4] = torch.matmul(-Rinv, T)
    return out


def vertices_transform(vertex, transform):
    n = vertex.shape[0] * vertex.shape[1]
    hom = torch.ones((n, 4))
    hom[:, :3] = vertex.reshape(n, 3)
    t = transform.matmul(hom.transpose(0, 1)).transpose(0, 1)
    t = t[:, :3] / t[:, 3:4]
    return t.reshape(vertex.shape[0], vertex.shape[1], 3)


def raycast_maps(free_vertex_f4, free_image_u8, c2w, voxel_size):
    """-> dict(color_map [H,W,3], vertex_map [H,W,3], confidence_map [H,W,1], depth_map [H,W,1])"""
    color = uchar4_image_to_tensor(free_image_u8)
    vc = float4_image_to_tensor(free_vertex_f4)
    vertex = vc[..., :3].contiguous() * voxel_size
    conf = vc[..., 3:4].contiguous()
    tv = vertices_transform(vertex, pose_inv(c2w))
    depth = tv[..., 2].unsqueeze(-1).contiguous()
    depth.masked_fill_((vertex.sum(2) == 0).unsqueeze(-1), 0)
    return dict(color_map=color, vertex_map=vertex, confidence_map=conf, depth_map=depth)


def frame_to_float(rgba_u8, depth_mm_i16):
    rgb = torch.from_numpy(np.ascontiguousarray(rgba_u8))[..., :3].to(torch.float32) / 255.0
    depth = torch.from_numpy(np.ascontiguousarray(depth_mm_i16)).to(torch.float32) / 1000.0
    return rgb, depth


def feature_gradient(img):
    H, W, Cc = img.shape
    wx = torch.tensor([[-1., 0., 1.], [-2., 0., 2.], [-1., 0., 1.]]).view(1, 1, 3, 3).to(img)
    wy = torch.tensor([[-1., -2., -1.], [0., 0., 0.], [1., 2., 1.]]).view(1, 1, 3, 3).to(img)
    p = img.permute(2, 0, 1).reshape(-1, 1, H, W)
    pad = torch.nn.functional.pad(p, (1, 1, 1, 1), mode="replicate")
    dx = torch.nn.functional.conv2d(pad, wx).squeeze(1).permute(1, 2, 0)
    dy = torch.nn.functional.conv2d(pad, wy).squeeze(1).permute(1, 2, 0)
    return dx, dy


def compute_normal_map(vertex_map):
    H, W, _ = vertex_map.shape
    dx, dy = feature_gradient(vertex_map)
    n = torch.cross(dy.reshape(-1, 3), dx.reshape(-1, 3), dim=-1).view(H, W, 3)
    n = n / (torch.norm(n, 2, -1, True) + 1e-8)
    invalid = vertex_map[..., 2] <= 0
    return torch.where(invalid.unsqueeze(-1), torch.zeros_like(n), n)


def sample_mask(maps, image, render_rgb, render_alpha, color_error_thres, depth_min, depth_max, alpha_max):
    """initNewGaussians' mask [H,W] bool; render_rgb / render_alpha None when the model is empty"""
    d = maps["depth_map"]
    valid = (d > depth_min) & (d < depth_max)
    valid = valid & ~((maps["vertex_map"].sum(2) == 0).unsqueeze(-1))
    src = maps["color_map"] if render_rgb is None else render_rgb
    err = torch.mean(torch.abs(src - image), -1, True)
    m = (err > color_error_thres) & valid
    if render_alpha is not None:
        m = m & (render_alpha.reshape(d.shape) < alpha_max)
    return m[..., 0]


def quaternion_from_axis_angle(axis, angle):
    na = axis / (torch.norm(axis, 2, -1, True) + 1e-8)
    half = angle / 2
    return torch.cat([torch.cos(half), na * torch.sin(half)], 1)


def compute_quat(init_vec, target_vec):
    axis = torch.cross(init_vec, target_vec, dim=1)
    axis = axis / (torch.norm(axis, 2, -1, True) + 1e-8)
    angle = torch.acos(torch.sum(init_vec * target_vec, 1)).unsqueeze(-1)
    return quaternion_from_axis_angle(axis, angle)


def dist2_knn3(xyz, chunk=2048):
    """mean of the three smallest squared distances to the OTHER points, fp32 (dx*dx + dy*dy + dz*dz), brute force"""
    x = xyz.to(torch.float32)
    n = x.shape[0]
    out = torch.empty(n)
    for s in range(0, n, chunk):
        q = x[s:s + chunk]
        d = q[:, None, :] - x[None, :, :]
        d2 = d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1] + d[..., 2] * d[..., 2]
        d2[torch.arange(q.shape[0]), torch.arange(s, s + q.shape[0])] = float("inf")
        best = torch.topk(d2, 3, dim=1, largest=False).values
        out[s:s + chunk] = (best[:, 0] + best[:, 1] + best[:, 2]) / 3.0
    return out


def init_params(xyz, rgb, normals, init_opac, max_scale, min_scale):
    """RawGaussianParams::init for maxSH = 3 (15 higher-order bases, zero)"""
    n = xyz.shape[0]
    raw_scales = torch.sqrt(dist2_knn3(xyz)).clamp(min_scale, max_scale).unsqueeze(1).repeat(1, 3)
    raw_scales[:, 2] = raw_scales[:, 2] * 0.1
    z = torch.zeros_like(raw_scales)
    z[:, 2] = 1
    quats = compute_quat(z, normals)
    return dict(means=xyz.numpy(), scales=raw_scales.log().numpy(), quats=quats.numpy(), featuresDc=((rgb - 0.5) / C0).numpy(),
                featuresRest=np.zeros((n, 15, 3), np.float32), opacities=torch.logit(init_opac * torch.ones(n, 1)).numpy())
