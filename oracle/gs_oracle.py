"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, fp32) of the reference's gsplat "GES" path, SURVEY.md
section 8 rows A1-A12.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

PARITY UNPINNED: the reference ships no test, golden vector or fixture for any of these kernels
and its .cu files cannot be compiled offline (they include glm, fetched from the network by CMake).  This file follows
the reference sources line by line (citations on every function); floating-point expressions are evaluated in fp32
without FMA contraction, in a fixed order that the CUDA kernels of gps_slam_b200/csrc/gs_*.cu reproduce exactly for
the integer-producing stages (radii, tile ranges, bins).  Transcendentals (exp, and rsqrt which is replaced by
1/sqrt on both sides) are the only sources of non-bit-identical results; tolerances are stated in the tests.

All arrays are float32 / int32 numpy arrays; C (number of cameras) is always 1 and squeezed out.
"""
import numpy as np

F = np.float32


def _f(x):
    return np.asarray(x, dtype=np.float32)


# ---------------------------------------------------------------------------------------------------------------
# A1: activations (include/raw_gs_param.h:80-82)
def real_scales(log_scales):
    return np.exp(_f(log_scales)).astype(np.float32)


def real_opacities(logit):
    x = _f(logit)
    return (F(1.0) / (F(1.0) + np.exp(-x))).astype(np.float32)


# A2: poseInv (src/tensor_math.cpp:56-67): row-major 4x4 camera-to-world -> world-to-camera
def pose_inv(c2w):
    c2w = _f(c2w)
    R = c2w[:3, :3]
    T = c2w[:3, 3]
    Rinv = R.T
    out = np.eye(4, dtype=np.float32)
    out[:3, :3] = Rinv
    # torch.matmul(-Rinv, T): fp32 dot products, accumulated left to right
    nR = -Rinv
    out[:3, 3] = (nR[:, 0] * T[0] + nR[:, 1] * T[1]) + nR[:, 2] * T[2]
    return out


# ---------------------------------------------------------------------------------------------------------------
# A3 + A4: fully_fused_projection_fwd_kernel (gsplat/rasterizer/fully_fused_projection_fwd.cu:43-194) with
# quat_to_rotmat (utils.cuh:14-36), quat_scale_to_covar_preci (:65-96), persp_proj (:253-292), add_blur (:603-610),
# inverse (:582-594); then clamp_max(radii, max_gs_radii) (src/raw_gs_model.cpp:241-242)
def quat_to_rotmat(q):
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    inv_norm = F(1.0) / np.sqrt(x * x + y * y + z * z + w * w)   # reference: rsqrt (approximate); 1/sqrt on both sides here
    x = x * inv_norm
    y = y * inv_norm
    z = z * inv_norm
    w = w * inv_norm
    x2, y2, z2 = x * x, y * y, z * z
    xy, xz, yz = x * y, x * z, y * z
    wx, wy, wz = w * x, w * y, w * z
    R = np.empty((q.shape[0], 3, 3), np.float32)   # row-major R[i][j]
    R[:, 0, 0] = F(1.0) - F(2.0) * (y2 + z2)
    R[:, 1, 0] = F(2.0) * (xy + wz)
    R[:, 2, 0] = F(2.0) * (xz - wy)
    R[:, 0, 1] = F(2.0) * (xy - wz)
    R[:, 1, 1] = F(1.0) - F(2.0) * (x2 + z2)
    R[:, 2, 1] = F(2.0) * (yz + wx)
    R[:, 0, 2] = F(2.0) * (xz + wy)
    R[:, 1, 2] = F(2.0) * (yz - wx)
    R[:, 2, 2] = F(1.0) - F(2.0) * (x2 + y2)
    return R, (w, x, y, z, inv_norm)


def _mm3(a, b):
    """batched 3x3 product, c[i][j] = (a[i][0]*b[0][j] + a[i][1]*b[1][j]) + a[i][2]*b[2][j]"""
    c = np.empty(np.broadcast_shapes(a.shape, b.shape), np.float32)
    for i in range(3):
        for j in range(3):
            c[..., i, j] = (a[..., i, 0] * b[..., 0, j] + a[..., i, 1] * b[..., 1, j]) + a[..., i, 2] * b[..., 2, j]
    return c


def _T(a):
    return np.swapaxes(a, -1, -2)


def quat_scale_to_covar(quats, scales):
    R, _ = quat_to_rotmat(quats)
    M = R * scales[:, None, :]            # M = R * diag(s)
    return _mm3(M, _T(M)), R


def persp_limits(W, H, fx, fy, cx, cy):
    tan_fovx = F(0.5) * F(W) / fx
    tan_fovy = F(0.5) * F(H) / fy
    lim_x_pos = (F(W) - cx) / fx + F(0.3) * tan_fovx
    lim_x_neg = cx / fx + F(0.3) * tan_fovx
    lim_y_pos = (F(H) - cy) / fy + F(0.3) * tan_fovy
    lim_y_neg = cy / fy + F(0.3) * tan_fovy
    return lim_x_pos, lim_x_neg, lim_y_pos, lim_y_neg


def project_fwd(means, quats, scales, viewmat, K, W, H, eps2d=0.3, near=0.01, far=1e10, radius_clip=0.0, max_radii=100):
    """-> dict(radii i32[N] (clamped), means2d[N,2], depths[N], conics[N,3], valid bool[N]); culled rows are zero"""
    means, quats, scales, viewmat, K = _f(means), _f(quats), _f(scales), _f(viewmat), _f(K)
    N = means.shape[0]
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    R = viewmat[:3, :3]
    t = viewmat[:3, 3]
    with np.errstate(all="ignore"):
        mc = np.empty((N, 3), np.float32)
        for i in range(3):
            mc[:, i] = ((R[i, 0] * means[:, 0] + R[i, 1] * means[:, 1]) + R[i, 2] * means[:, 2]) + t[i]
        valid = ~((mc[:, 2] < F(near)) | (mc[:, 2] > F(far)))
        covar, _ = quat_scale_to_covar(quats, scales)
        Rb = np.broadcast_to(R, (N, 3, 3))
        covar_c = _mm3(_mm3(Rb, covar), _T(Rb))

        x, y, z = mc[:, 0], mc[:, 1], mc[:, 2]
        lxp, lxn, lyp, lyn = persp_limits(W, H, fx, fy, cx, cy)
        rz = F(1.0) / z
        rz2 = rz * rz
        tx = z * np.minimum(lxp, np.maximum(-lxn, x * rz))
        ty = z * np.minimum(lyp, np.maximum(-lyn, y * rz))
        J00 = fx * rz
        J11 = fy * rz
        J02 = (-fx) * tx * rz2
        J12 = (-fy) * ty * rz2
        cc = covar_c
        # tmp = J * covar_c (2x3), zero entries of J skipped (exact)
        t00 = J00 * cc[:, 0, 0] + J02 * cc[:, 2, 0]
        t01 = J00 * cc[:, 0, 1] + J02 * cc[:, 2, 1]
        t02 = J00 * cc[:, 0, 2] + J02 * cc[:, 2, 2]
        t10 = J11 * cc[:, 1, 0] + J12 * cc[:, 2, 0]
        t11 = J11 * cc[:, 1, 1] + J12 * cc[:, 2, 1]
        t12 = J11 * cc[:, 1, 2] + J12 * cc[:, 2, 2]
        c00 = t00 * J00 + t02 * J02
        c01 = t01 * J11 + t02 * J12
        c10 = t10 * J00 + t12 * J02
        c11 = t11 * J11 + t12 * J12
        m2x = fx * x * rz + cx
        m2y = fy * y * rz + cy

        c00 = c00 + F(eps2d)
        c11 = c11 + F(eps2d)
        det = c00 * c11 - c01 * c10
        valid &= ~(det <= 0)
        invdet = F(1.0) / det
        con_a = c11 * invdet
        con_b = (-c01) * invdet
        con_c = c00 * invdet
        b = F(0.5) * (c00 + c11)
        v1 = b + np.sqrt(np.maximum(F(0.01), b * b - det))
        radius = np.ceil(F(3.0) * np.sqrt(v1))
        valid &= ~(radius <= F(radius_clip))
        valid &= ~((m2x + radius <= 0) | (m2x - radius >= F(W)) | (m2y + radius <= 0) | (m2y - radius >= F(H)))
    valid &= np.isfinite(radius)
    radii = np.where(valid, radius, 0).astype(np.int32)
    if max_radii > 0:
        radii = np.minimum(radii, np.int32(max_radii))
    z0 = np.zeros(N, np.float32)
    out = dict(radii=radii, valid=valid,
               means2d=np.stack([np.where(valid, m2x, z0), np.where(valid, m2y, z0)], 1).astype(np.float32),
               depths=np.where(valid, z, z0).astype(np.float32),
               conics=np.stack([np.where(valid, con_a, z0), np.where(valid, con_b, z0), np.where(valid, con_c, z0)], 1).astype(np.float32))
    return out


# ---------------------------------------------------------------------------------------------------------------
# A5: sh_coeffs_to_color_fast (gsplat/rasterizer/spherical_harmonics.cuh:17-105), degree <= 3, then
# clamp_min(c + 0.5, 0) (src/raw_gs_model.cpp:257).  dirs = means - cam_pos, NOT normalised by the caller.
def sh_bases(dirs, degree=3):
    """-> (basis [N,16] float32 with the constant factors folded in as the reference does, (x,y,z,inorm))"""
    d = _f(dirs)
    inorm = F(1.0) / np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2])
    x, y, z = d[:, 0] * inorm, d[:, 1] * inorm, d[:, 2] * inorm
    return x, y, z, inorm


def sh_fwd(dirs, coeffs, mask=None, degree=3):
    """coeffs [N,16,3] -> raw SH colour [N,3] (before the +0.5 / clamp)"""
    co = _f(coeffs)
    N = co.shape[0]
    x, y, z, _ = sh_bases(dirs)
    res = np.empty((N, 3), np.float32)
    for c in range(3):
        cf = co[:, :, c]
        r = F(0.2820947917738781) * cf[:, 0]
        if degree >= 1:
            r = r + F(0.48860251190292) * (((-y) * cf[:, 1] + z * cf[:, 2]) - x * cf[:, 3])
        if degree >= 2:
            z2 = z * z
            fTmp0B = F(-1.092548430592079) * z
            fC1 = x * x - y * y
            fS1 = F(2.0) * x * y
            pSH6 = F(0.9461746957575601) * z2 - F(0.3153915652525201)
            pSH7 = fTmp0B * x
            pSH5 = fTmp0B * y
            pSH8 = F(0.5462742152960395) * fC1
            pSH4 = F(0.5462742152960395) * fS1
            r = r + ((((pSH4 * cf[:, 4] + pSH5 * cf[:, 5]) + pSH6 * cf[:, 6]) + pSH7 * cf[:, 7]) + pSH8 * cf[:, 8])
        if degree >= 3:
            fTmp0C = F(-2.285228997322329) * z2 + F(0.4570457994644658)
            fTmp1B = F(1.445305721320277) * z
            fC2 = x * fC1 - y * fS1
            fS2 = x * fS1 + y * fC1
            pSH12 = z * (F(1.865881662950577) * z2 - F(1.119528997770346))
            pSH13 = fTmp0C * x
            pSH11 = fTmp0C * y
            pSH14 = fTmp1B * fC1
            pSH10 = fTmp1B * fS1
            pSH15 = F(-0.5900435899266435) * fC2
            pSH9 = F(-0.5900435899266435) * fS2
            r = r + ((((((pSH9 * cf[:, 9] + pSH10 * cf[:, 10]) + pSH11 * cf[:, 11]) + pSH12 * cf[:, 12]) + pSH13 * cf[:, 13])
                      + pSH14 * cf[:, 14]) + pSH15 * cf[:, 15])
        res[:, c] = r
    if mask is not None:
        res[~mask] = 0
    return res


def sh_colors(dirs, coeffs, mask=None):
    return np.maximum(sh_fwd(dirs, coeffs, mask) + F(0.5), F(0.0)).astype(np.float32)


# ---------------------------------------------------------------------------------------------------------------
# A6: isect_tiles_no_depth (gsplat/rasterizer/isect_tiles_no_depth.cu:57-129), stable sort by tile (:303-344),
# isect_offset_encode_no_depth (:373-425)
def tile_ranges(means2d, radii, tile_size, tile_w, tile_h):
    r = radii.astype(np.float32)
    ts = F(tile_size)
    tile_radius = r / ts
    tx = means2d[:, 0] / ts
    ty = means2d[:, 1] / ts

    def sat_u32(v):   # CUDA float -> uint32 conversion saturates (negative -> 0); see SURVEY.md section 9
        v = np.where(np.isnan(v), 0, v)
        return np.clip(v, 0, 4294967295.0).astype(np.int64)

    x0 = np.minimum(sat_u32(np.floor(tx - tile_radius)), tile_w)
    y0 = np.minimum(sat_u32(np.floor(ty - tile_radius)), tile_h)
    x1 = np.minimum(sat_u32(np.ceil(tx + tile_radius)), tile_w)
    y1 = np.minimum(sat_u32(np.ceil(ty + tile_radius)), tile_h)
    vis = radii > 0
    x0, y0, x1, y1 = [np.where(vis, a, 0) for a in (x0, y0, x1, y1)]
    return x0, y0, x1, y1


def isect_tiles_no_depth(means2d, radii, tile_size, tile_w, tile_h):
    """-> tiles_per_gauss i32[N], groups_per_gauss i32[N], isect_ids i64[I] (sorted), flatten_ids i32[I] (stable order)"""
    means2d = _f(means2d)
    x0, y0, x1, y1 = tile_ranges(means2d, radii, tile_size, tile_w, tile_h)
    tiles_per_gauss = np.where(radii > 0, (y1 - y0) * (x1 - x0), 0).astype(np.int32)
    r = radii.astype(np.float32)
    groups_per_gauss = np.where(radii > 0, ((F(4.0) * r * r + F(32.0) - F(1.0)) / F(32.0)).astype(np.int32), 0).astype(np.int32)
    ids, flat = [], []
    for g in np.nonzero(tiles_per_gauss > 0)[0]:
        yy, xx = np.meshgrid(np.arange(y0[g], y1[g]), np.arange(x0[g], x1[g]), indexing="ij")
        t = (yy * tile_w + xx).reshape(-1)
        ids.append(t)
        flat.append(np.full(t.shape, g, np.int32))
    if ids:
        ids = np.concatenate(ids).astype(np.int64)
        flat = np.concatenate(flat)
        order = np.argsort(ids, kind="stable")
        ids, flat = ids[order], flat[order]
    else:
        ids, flat = np.zeros(0, np.int64), np.zeros(0, np.int32)
    return tiles_per_gauss, groups_per_gauss, ids, flat


def isect_offset_encode(isect_ids, n_tiles):
    """-> tile_offsets i32[n_tiles]: offsets[t] = first index with tile id >= t (= n_isects past the end)"""
    return np.searchsorted(isect_ids, np.arange(n_tiles), side="left").astype(np.int32)


# ---------------------------------------------------------------------------------------------------------------
# A7: rasterize_to_pixels_fwd_ges_kernel (gsplat/rasterizer/rasterize_to_pixels_fwd_ges.cu:47-215)
def raster_fwd_ges(means2d, conics, colors4, opacities, ref_depth, W, H, tile_size, tile_offsets, flatten_ids, delta_depth):
    """colors4 [N,4] = rgb + camera depth; ref_depth [H,W] (already clamped: <0.01 -> 1000).
    -> render [H,W,4], alphas [H,W]; per pixel the sum runs over the tile list in order (ascending Gaussian id)."""
    tile_w = (W + tile_size - 1) // tile_size
    tile_h = (H + tile_size - 1) // tile_size
    render = np.zeros((H, W, 4), np.float32)
    alphas = np.zeros((H, W), np.float32)
    n_isects = len(flatten_ids)
    means2d, conics, colors4, opacities = _f(means2d), _f(conics), _f(colors4), _f(opacities).reshape(-1)
    for ty in range(tile_h):
        for tx in range(tile_w):
            tid = ty * tile_w + tx
            s = tile_offsets[tid]
            e = n_isects if tid == tile_w * tile_h - 1 else tile_offsets[tid + 1]
            if e <= s:
                continue
            y0, x0 = ty * tile_size, tx * tile_size
            y1, x1 = min(y0 + tile_size, H), min(x0 + tile_size, W)
            py, px = np.meshgrid(np.arange(y0, y1, dtype=np.float32) + F(0.5), np.arange(x0, x1, dtype=np.float32) + F(0.5), indexing="ij")
            rd = ref_depth[y0:y1, x0:x1]
            acc = np.zeros((y1 - y0, x1 - x0, 4), np.float32)
            wsum = np.zeros((y1 - y0, x1 - x0), np.float32)
            for g in flatten_ids[s:e]:
                dx = means2d[g, 0] - px
                dy = means2d[g, 1] - py
                sigma = F(0.5) * (conics[g, 0] * dx * dx + conics[g, 2] * dy * dy) + conics[g, 1] * dx * dy
                alpha = np.minimum(F(0.999), opacities[g] * np.exp(-sigma))
                ok = ~(colors4[g, 3] > rd + F(delta_depth)) & ~((sigma < 0) | (alpha < F(1.0) / F(255.0)))
                a = np.where(ok, alpha, F(0)).astype(np.float32)
                acc += colors4[g][None, None, :] * a[..., None]
                wsum += a
            render[y0:y1, x0:x1] = acc
            alphas[y0:y1, x0:x1] = wsum
    return render, alphas


# A8: composite + L1 (src/raw_gs_model.cpp:317-326, 369-417; src/tensor_math.cpp:41-44)
def composite(render, alphas, ref_depth_raw, base_color):
    w = alphas[..., None]
    rgb = (render[..., :3] + base_color * F(1.0)) / (w + F(1.0))
    bw = (ref_depth_raw > 0).astype(np.float32)[..., None]
    depth = (render[..., 3:4] + ref_depth_raw[..., None] * bw) / (w + bw)
    return rgb.astype(np.float32), depth[..., 0].astype(np.float32)


def l1_loss_and_grad(rgb, gt):
    """loss = mean|rgb - gt|; returns loss, dL/drgb"""
    diff = rgb - gt
    n = F(diff.size)
    return np.abs(diff).astype(np.float64).mean(), (np.sign(diff) / n).astype(np.float32)


def composite_bwd(v_rgb, rgb, alphas):
    """-> v_render [H,W,4] (depth channel zero: depth_weight = 0), v_alphas [H,W]"""
    w1 = alphas[..., None] + F(1.0)
    v_render = np.zeros(rgb.shape[:2] + (4,), np.float32)
    v_render[..., :3] = v_rgb / w1
    v_alpha = -(v_rgb * rgb).sum(-1) / w1[..., 0]
    return v_render, v_alpha.astype(np.float32)


# ---------------------------------------------------------------------------------------------------------------
# A9: temp_bwd_kernel (gsplat/rasterizer/rasterize_to_pixels_bwd_ges_new_parallel.cu:62-199): Gaussian-parallel,
# support = the (2r x 2r) pixel box, NOT the forward's tile footprint
def raster_bwd_ges(means2d, conics, colors4, opacities, radii, ref_depth, delta_depth, W, H, v_render, v_alphas):
    N = means2d.shape[0]
    means2d, conics, colors4, opacities = _f(means2d), _f(conics), _f(colors4), _f(opacities).reshape(-1)
    v_means2d = np.zeros((N, 2), np.float32)
    v_conics = np.zeros((N, 3), np.float32)
    v_colors = np.zeros((N, 4), np.float32)
    v_opac = np.zeros(N, np.float32)
    for g in np.nonzero(radii > 0)[0]:
        r = int(radii[g])
        xi, yi = int(means2d[g, 0]), int(means2d[g, 1])      # C truncation toward zero
        x_min, x_max, y_min, y_max = xi - r, xi + r, yi - r, yi + r
        js = np.arange(x_min + 1, x_max + 1)
        is_ = np.arange(y_min + 1, y_max + 1)
        js = js[(js >= 0) & (js < W)]
        is_ = is_[(is_ >= 0) & (is_ < H)]
        if len(js) == 0 or len(is_) == 0:
            continue
        ii, jj = np.meshgrid(is_, js, indexing="ij")
        px = jj.astype(np.float32) + F(0.5)
        py = ii.astype(np.float32) + F(0.5)
        dx = means2d[g, 0] - px
        dy = means2d[g, 1] - py
        a, b, c = conics[g]
        sigma = F(0.5) * (a * dx * dx + c * dy * dy) + b * dx * dy
        vis = np.exp(-sigma).astype(np.float32)
        opac = opacities[g]
        alpha = np.minimum(F(0.999), opac * vis)
        rd = ref_depth[ii, jj]
        valid = ~((sigma < 0) | (alpha < F(1.0) / F(255.0)) | (colors4[g, 3] > rd + F(delta_depth)))
        vc = v_render[ii, jj]                # [..,4]
        va = v_alphas[ii, jj]
        al = np.where(valid, alpha, F(0))
        v_colors[g] = (al[..., None] * vc).reshape(-1, 4).astype(np.float64).sum(0)
        v_alpha = (colors4[g][None, None, :] * vc).sum(-1) + va
        grad_ok = valid & (opac * vis <= F(0.999))
        v_sigma = np.where(grad_ok, -opac * vis * v_alpha, F(0)).astype(np.float32)
        v_conics[g, 0] = (F(0.5) * v_sigma * dx * dx).astype(np.float64).sum()
        v_conics[g, 1] = (v_sigma * dx * dy).astype(np.float64).sum()
        v_conics[g, 2] = (F(0.5) * v_sigma * dy * dy).astype(np.float64).sum()
        v_means2d[g, 0] = (v_sigma * (a * dx + b * dy)).astype(np.float64).sum()
        v_means2d[g, 1] = (v_sigma * (b * dx + c * dy)).astype(np.float64).sum()
        v_opac[g] = np.where(grad_ok, vis * v_alpha, F(0)).astype(np.float64).sum()
    return v_means2d, v_conics, v_colors, v_opac


# ---------------------------------------------------------------------------------------------------------------
# A10: sh_coeffs_to_color_fast_vjp (gsplat/rasterizer/spherical_harmonics.cuh:108-366), degree 3, with v_dirs
def sh_bwd(dirs, coeffs, v_colors, mask=None):
    """v_colors [N,3] = gradient w.r.t. the raw SH colour -> v_coeffs [N,16,3], v_dirs [N,3]"""
    co = _f(coeffs).astype(np.float64)
    vcol = _f(v_colors).astype(np.float64)
    N = co.shape[0]
    x, y, z, inorm = [a.astype(np.float64) for a in sh_bases(dirs)]
    v_coeffs = np.zeros((N, 16, 3), np.float64)
    v_dirs = np.zeros((N, 3), np.float64)
    z2 = z * z
    fTmp0B = -1.092548430592079 * z
    fC1 = x * x - y * y
    fS1 = 2.0 * x * y
    pSH6 = 0.9461746957575601 * z2 - 0.3153915652525201
    pSH7, pSH5 = fTmp0B * x, fTmp0B * y
    pSH8, pSH4 = 0.5462742152960395 * fC1, 0.5462742152960395 * fS1
    fTmp0C = -2.285228997322329 * z2 + 0.4570457994644658
    fTmp1B = 1.445305721320277 * z
    fC2 = x * fC1 - y * fS1
    fS2 = x * fS1 + y * fC1
    pSH12 = z * (1.865881662950577 * z2 - 1.119528997770346)
    pSH13, pSH11 = fTmp0C * x, fTmp0C * y
    pSH14, pSH10 = fTmp1B * fC1, fTmp1B * fS1
    pSH15, pSH9 = -0.5900435899266435 * fC2, -0.5900435899266435 * fS2
    basis = [0.2820947917738781 * np.ones_like(x), -0.48860251190292 * y, 0.48860251190292 * z, -0.48860251190292 * x,
             pSH4, pSH5, pSH6, pSH7, pSH8, pSH9, pSH10, pSH11, pSH12, pSH13, pSH14, pSH15]
    # derivatives of the basis w.r.t. the unit direction
    fTmp0B_z = -1.092548430592079
    fC1_x, fC1_y, fS1_x, fS1_y = 2.0 * x, -2.0 * y, 2.0 * y, 2.0 * x
    pSH6_z = 2.0 * 0.9461746957575601 * z
    pSH7_x, pSH7_z, pSH5_y, pSH5_z = fTmp0B, fTmp0B_z * x, fTmp0B, fTmp0B_z * y
    pSH8_x, pSH8_y = 0.5462742152960395 * fC1_x, 0.5462742152960395 * fC1_y
    pSH4_x, pSH4_y = 0.5462742152960395 * fS1_x, 0.5462742152960395 * fS1_y
    fTmp0C_z = -2.285228997322329 * 2.0 * z
    fTmp1B_z = 1.445305721320277
    fC2_x = fC1 + x * fC1_x - y * fS1_x
    fC2_y = x * fC1_y - fS1 - y * fS1_y
    fS2_x = fS1 + x * fS1_x + y * fC1_x
    fS2_y = x * fS1_y + fC1 + y * fC1_y
    pSH12_z = 3.0 * 1.865881662950577 * z2 - 1.119528997770346
    pSH13_x, pSH13_z, pSH11_y, pSH11_z = fTmp0C, fTmp0C_z * x, fTmp0C, fTmp0C_z * y
    pSH14_x, pSH14_y, pSH14_z = fTmp1B * fC1_x, fTmp1B * fC1_y, fTmp1B_z * fC1
    pSH10_x, pSH10_y, pSH10_z = fTmp1B * fS1_x, fTmp1B * fS1_y, fTmp1B_z * fS1
    pSH15_x, pSH15_y = -0.5900435899266435 * fC2_x, -0.5900435899266435 * fC2_y
    pSH9_x, pSH9_y = -0.5900435899266435 * fS2_x, -0.5900435899266435 * fS2_y
    zero = np.zeros_like(x)
    d_x = [zero, zero, zero, -0.48860251190292 + zero, pSH4_x, zero, zero, pSH7_x, pSH8_x, pSH9_x, pSH10_x, zero, zero, pSH13_x, pSH14_x, pSH15_x]
    d_y = [zero, -0.48860251190292 + zero, zero, zero, pSH4_y, pSH5_y, zero, zero, pSH8_y, pSH9_y, pSH10_y, pSH11_y, zero, zero, pSH14_y, pSH15_y]
    d_z = [zero, zero, 0.48860251190292 + zero, zero, zero, pSH5_z, pSH6_z, pSH7_z, zero, zero, pSH10_z, pSH11_z, pSH12_z, pSH13_z, pSH14_z, zero]
    v_x = np.zeros(N)
    v_y = np.zeros(N)
    v_z = np.zeros(N)
    for c in range(3):
        vc = vcol[:, c]
        for k in range(16):
            v_coeffs[:, k, c] = basis[k] * vc
            v_x += vc * d_x[k] * co[:, k, c]
            v_y += vc * d_y[k] * co[:, k, c]
            v_z += vc * d_z[k] * co[:, k, c]
    dotp = v_x * x + v_y * y + v_z * z
    v_dirs[:, 0] = (v_x - dotp * x) * inorm
    v_dirs[:, 1] = (v_y - dotp * y) * inorm
    v_dirs[:, 2] = (v_z - dotp * z) * inorm
    if mask is not None:
        v_coeffs[~mask] = 0
        v_dirs[~mask] = 0
    return v_coeffs.astype(np.float32), v_dirs.astype(np.float32)


# ---------------------------------------------------------------------------------------------------------------
# A11: fully_fused_projection_bwd_kernel (gsplat/rasterizer/fully_fused_projection_bwd.cu:53-265) with inverse_vjp
# (utils.cuh:596-600), persp_proj_vjp (:294-372), pos/covar_world_to_cam_vjp (:529-580), quat_scale_to_covar_vjp (:98-136),
# quat_to_rotmat_vjp (:38-62).  Evaluated in float64 and rounded: this stage is a true VJP, compared with a tolerance.
def project_bwd(means, quats, scales, viewmat, K, W, H, radii, conics, v_means2d, v_depths, v_conics):
    means, quats, scales = [_f(a).astype(np.float64) for a in (means, quats, scales)]
    viewmat, K = _f(viewmat).astype(np.float64), _f(K).astype(np.float64)
    conics, v_means2d, v_depths, v_conics = [_f(a).astype(np.float64) for a in (conics, v_means2d, v_depths, v_conics)]
    N = means.shape[0]
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    R = viewmat[:3, :3]
    t = viewmat[:3, 3]
    vis = radii > 0
    v_means = np.zeros((N, 3))
    v_quats = np.zeros((N, 4))
    v_scales = np.zeros((N, 3))
    idx = np.nonzero(vis)[0]
    if len(idx) == 0:
        return v_means.astype(np.float32), v_quats.astype(np.float32), v_scales.astype(np.float32)
    m, q, s = means[idx], quats[idx], scales[idx]
    n = len(idx)
    Minv = np.zeros((n, 2, 2))
    Minv[:, 0, 0], Minv[:, 0, 1], Minv[:, 1, 0], Minv[:, 1, 1] = conics[idx, 0], conics[idx, 1], conics[idx, 1], conics[idx, 2]
    vMinv = np.zeros((n, 2, 2))
    vMinv[:, 0, 0], vMinv[:, 0, 1], vMinv[:, 1, 0], vMinv[:, 1, 1] = v_conics[idx, 0], v_conics[idx, 1] * 0.5, v_conics[idx, 1] * 0.5, v_conics[idx, 2]
    v_cov2d = -Minv @ vMinv @ Minv
    # forward recompute
    w_, x_, y_, z_ = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    inv_norm = 1.0 / np.sqrt(x_ * x_ + y_ * y_ + z_ * z_ + w_ * w_)
    x, y, z, w = x_ * inv_norm, y_ * inv_norm, z_ * inv_norm, w_ * inv_norm
    Rq = np.zeros((n, 3, 3))
    Rq[:, 0, 0] = 1 - 2 * (y * y + z * z); Rq[:, 0, 1] = 2 * (x * y - w * z); Rq[:, 0, 2] = 2 * (x * z + w * y)
    Rq[:, 1, 0] = 2 * (x * y + w * z); Rq[:, 1, 1] = 1 - 2 * (x * x + z * z); Rq[:, 1, 2] = 2 * (y * z - w * x)
    Rq[:, 2, 0] = 2 * (x * z - w * y); Rq[:, 2, 1] = 2 * (y * z + w * x); Rq[:, 2, 2] = 1 - 2 * (x * x + y * y)
    Mm = Rq * s[:, None, :]
    covar = Mm @ np.swapaxes(Mm, 1, 2)
    mc = m @ R.T + t
    covar_c = R @ covar @ R.T
    X, Y, Z = mc[:, 0], mc[:, 1], mc[:, 2]
    lxp, lxn, lyp, lyn = [float(a) for a in persp_limits(W, H, F(fx), F(fy), F(cx), F(cy))]
    rz = 1.0 / Z
    rz2 = rz * rz
    tx = Z * np.minimum(lxp, np.maximum(-lxn, X * rz))
    ty = Z * np.minimum(lyp, np.maximum(-lyn, Y * rz))
    J = np.zeros((n, 2, 3))
    J[:, 0, 0] = fx * rz; J[:, 0, 2] = -fx * tx * rz2
    J[:, 1, 1] = fy * rz; J[:, 1, 2] = -fy * ty * rz2
    v_covar_c = np.swapaxes(J, 1, 2) @ v_cov2d @ J
    v_mean_c = np.stack([fx * rz * v_means2d[idx, 0], fy * rz * v_means2d[idx, 1],
                         -(fx * X * v_means2d[idx, 0] + fy * Y * v_means2d[idx, 1]) * rz2], 1)
    rz3 = rz2 * rz
    v_J = v_cov2d @ J @ np.swapaxes(covar_c, 1, 2) + np.swapaxes(v_cov2d, 1, 2) @ J @ covar_c   # [n,2,3]
    inx = (X * rz <= lxp) & (X * rz >= -lxn)
    iny = (Y * rz <= lyp) & (Y * rz >= -lyn)
    v_mean_c[:, 0] += np.where(inx, -fx * rz2 * v_J[:, 0, 2], 0)
    v_mean_c[:, 2] += np.where(inx, 0, -fx * rz3 * v_J[:, 0, 2] * tx)
    v_mean_c[:, 1] += np.where(iny, -fy * rz2 * v_J[:, 1, 2], 0)
    v_mean_c[:, 2] += np.where(iny, 0, -fy * rz3 * v_J[:, 1, 2] * ty)
    v_mean_c[:, 2] += -fx * rz2 * v_J[:, 0, 0] - fy * rz2 * v_J[:, 1, 1] + 2 * fx * tx * rz3 * v_J[:, 0, 2] + 2 * fy * ty * rz3 * v_J[:, 1, 2]
    v_mean_c[:, 2] += v_depths[idx]
    v_means[idx] = v_mean_c @ R
    v_covar = R.T @ v_covar_c @ R
    # quat_scale_to_covar_vjp
    v_M = (v_covar + np.swapaxes(v_covar, 1, 2)) @ Mm
    v_R = v_M * s[:, None, :]
    v_scales[idx] = (Rq * v_M).sum(1)
    # quat_to_rotmat_vjp; glm v_R[c][r] is column-major: v_R[i][j] (glm) = our v_R[:, j, i]
    g = lambda i, j: v_R[:, j, i]
    vq = np.stack([
        2 * (x * (g(1, 2) - g(2, 1)) + y * (g(2, 0) - g(0, 2)) + z * (g(0, 1) - g(1, 0))),
        2 * (-2 * x * (g(1, 1) + g(2, 2)) + y * (g(0, 1) + g(1, 0)) + z * (g(0, 2) + g(2, 0)) + w * (g(1, 2) - g(2, 1))),
        2 * (x * (g(0, 1) + g(1, 0)) - 2 * y * (g(0, 0) + g(2, 2)) + z * (g(1, 2) + g(2, 1)) + w * (g(2, 0) - g(0, 2))),
        2 * (x * (g(0, 2) + g(2, 0)) + y * (g(1, 2) + g(2, 1)) - 2 * z * (g(0, 0) + g(1, 1)) + w * (g(0, 1) - g(1, 0)))], 1)
    qn = np.stack([w, x, y, z], 1)
    v_quats[idx] = (vq - (vq * qn).sum(1, keepdims=True) * qn) * inv_norm[:, None]
    return v_means.astype(np.float32), v_quats.astype(np.float32), v_scales.astype(np.float32)


# ---------------------------------------------------------------------------------------------------------------
# A12: torch::optim::Adam::step as configured by initOptimizers (src/raw_gs_model.cpp:654-675): betas (0.9f, 0.999f)
# widened to double, eps 1e-15f, no weight decay, no amsgrad.  libtorch is third-party (version unpinned);
# this is the update documented for torch.optim.Adam.
BETA1 = float(np.float32(0.9))
BETA2 = float(np.float32(0.999))
EPS = float(np.float32(1e-15))


def adam_step(p, g, m, v, step, lr):
    """in-place on float32 arrays; `step` is the 1-based step count after increment"""
    bc1 = 1.0 - BETA1 ** step
    bc2 = 1.0 - BETA2 ** step
    m *= F(BETA1)
    m += g * F(1.0 - BETA1)
    v *= F(BETA2)
    v += g * g * F(1.0 - BETA2)
    denom = np.sqrt(v) / F(np.sqrt(bc2)) + F(EPS)
    p -= F(lr / bc1) * (m / denom)


# ---------------------------------------------------------------------------------------------------------------
# whole iteration, as RawGaussianModel::gesForward + computeLoss + backward produce it
def ges_iteration(params, c2w, K, W, H, ref_depth_raw, base_color, gt_rgb, delta_depth=0.1, max_radii=100, tile_size=16):
    """params: dict(means, scales(log), quats, featuresDc [N,3], featuresRest [N,15,3], opacities(logit) [N,1])
    -> dict of every intermediate the parity tests compare"""
    means = _f(params["means"])
    scales = real_scales(params["scales"])
    opac = real_opacities(params["opacities"]).reshape(-1)
    viewmat = pose_inv(c2w)
    proj = project_fwd(means, params["quats"], scales, viewmat, K, W, H, max_radii=max_radii)
    radii = proj["radii"]
    shs = np.concatenate([_f(params["featuresDc"])[:, None, :], _f(params["featuresRest"])], 1)
    cam_t = _f(c2w)[:3, 3]
    dirs = means - cam_t[None, :]
    mask = radii > 0
    sh_raw = sh_fwd(dirs, shs, mask)
    colors = np.maximum(sh_raw + F(0.5), F(0.0)).astype(np.float32)
    tile_w, tile_h = (W + tile_size - 1) // tile_size, (H + tile_size - 1) // tile_size
    tpg, gpg, isect_ids, flatten_ids = isect_tiles_no_depth(proj["means2d"], radii, tile_size, tile_w, tile_h)
    offsets = isect_offset_encode(isect_ids, tile_w * tile_h)
    colors4 = np.concatenate([colors, proj["depths"][:, None]], 1).astype(np.float32)
    ref_clamped = np.where(ref_depth_raw < F(0.01), F(1000.0), ref_depth_raw).astype(np.float32)
    render, alphas = raster_fwd_ges(proj["means2d"], proj["conics"], colors4, opac, ref_clamped, W, H, tile_size, offsets, flatten_ids, delta_depth)
    rgb, depth = composite(render, alphas, ref_depth_raw, base_color)
    loss, v_rgb = l1_loss_and_grad(rgb, gt_rgb)
    v_render, v_alphas = composite_bwd(v_rgb, rgb, alphas)
    v_m2d, v_con, v_col4, v_op = raster_bwd_ges(proj["means2d"], proj["conics"], colors4, opac, radii, ref_clamped, delta_depth, W, H, v_render, v_alphas)
    v_sh_raw = v_col4[:, :3] * ((sh_raw + F(0.5)) >= 0)      # clamp_min backward
    v_sh_raw[~mask] = 0
    v_coeffs, v_dirs = sh_bwd(dirs, shs, v_sh_raw, mask)
    v_means, v_quats, v_scales = project_bwd(means, params["quats"], scales, viewmat, K, W, H, radii, proj["conics"], v_m2d, v_col4[:, 3], v_con)
    grads = dict(means=(v_means + v_dirs).astype(np.float32),
                 scales=(v_scales * scales).astype(np.float32),
                 quats=v_quats,
                 featuresDc=v_coeffs[:, 0, :].copy(),
                 featuresRest=v_coeffs[:, 1:, :].copy(),
                 opacities=(v_op * opac * (F(1.0) - opac)).reshape(-1, 1).astype(np.float32))
    return dict(viewmat=viewmat, proj=proj, colors=colors, sh_raw=sh_raw, tiles_per_gauss=tpg, groups_per_gauss=gpg, isect_ids=isect_ids,
                flatten_ids=flatten_ids, tile_offsets=offsets, render=render, alphas=alphas, rgb=rgb, depth=depth, loss=loss,
                v_render=v_render, v_alphas=v_alphas, v_means2d=v_m2d, v_conics=v_con, v_colors=v_col4, v_opacities=v_op, grads=grads)
