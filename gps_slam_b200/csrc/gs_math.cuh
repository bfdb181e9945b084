// Per-Gaussian math of the gsplat "GES" path as device functions, shared by the staged (gsplat::*_tensor-shaped) and the
// fused kernels.  Forward expressions are written in the exact operation order of oracle/gs_oracle.py (no FMA contraction:
// the including translation units are built with -fmad=false) so that radii / tile ranges / bins are bit-identical.
//
// reference sources restated here:
//   quat_to_rotmat, quat_scale_to_covar_preci, persp_proj, add_blur, inverse and their *_vjp   gsplat/rasterizer/utils.cuh
//   fully_fused_projection_{fwd,bwd}_kernel                       gsplat/rasterizer/fully_fused_projection_{fwd,bwd}.cu
//   sh_coeffs_to_color_fast{,_vjp}                                gsplat/rasterizer/spherical_harmonics.cuh
#pragma once
#include <cuda_runtime.h>

namespace gs
{

struct CamParams
{
    float R[9];    // world->camera rotation, row-major
    float t[3];
    float fx, fy, cx, cy;
    int W, H;
    float eps2d, near_plane, far_plane, radius_clip;
    int max_radii; // clamp_max(radii, max_gs_radii) (src/raw_gs_model.cpp:241-242); <= 0 disables
    float lim_x_pos, lim_x_neg, lim_y_pos, lim_y_neg;
    float cam_pos[3]; // c2w translation (view direction origin)
};

__host__ __device__ inline void cam_limits(CamParams &c)
{
    float tan_fovx = 0.5f * (float)c.W / c.fx;
    float tan_fovy = 0.5f * (float)c.H / c.fy;
    c.lim_x_pos = ((float)c.W - c.cx) / c.fx + 0.3f * tan_fovx;
    c.lim_x_neg = c.cx / c.fx + 0.3f * tan_fovx;
    c.lim_y_pos = ((float)c.H - c.cy) / c.fy + 0.3f * tan_fovy;
    c.lim_y_neg = c.cy / c.fy + 0.3f * tan_fovy;
}

struct Proj
{
    float m2x, m2y, depth, ca, cb, cc;
    int radius; // 0 = culled
};

// c[i][j] = (a[i][0]*b[0][j] + a[i][1]*b[1][j]) + a[i][2]*b[2][j], row-major 3x3
__device__ __forceinline__ void mm3(const float *a, const float *b, float *c)
{
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            c[i * 3 + j] = (a[i * 3 + 0] * b[0 * 3 + j] + a[i * 3 + 1] * b[1 * 3 + j]) + a[i * 3 + 2] * b[2 * 3 + j];
}
__device__ __forceinline__ void mm3_bt(const float *a, const float *b, float *c) // c = a * b^T
{
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            c[i * 3 + j] = (a[i * 3 + 0] * b[j * 3 + 0] + a[i * 3 + 1] * b[j * 3 + 1]) + a[i * 3 + 2] * b[j * 3 + 2];
}

__device__ __forceinline__ void quat_to_rotmat(const float *q, float *R, float *nq /* w x y z inv_norm */)
{
    float w = q[0], x = q[1], y = q[2], z = q[3];
    float inv_norm = 1.0f / sqrtf(x * x + y * y + z * z + w * w); // reference: rsqrt (approximate); exact on both sides here
    x *= inv_norm, y *= inv_norm, z *= inv_norm, w *= inv_norm;
    float x2 = x * x, y2 = y * y, z2 = z * z;
    float xy = x * y, xz = x * z, yz = y * z;
    float wx = w * x, wy = w * y, wz = w * z;
    R[0] = 1.f - 2.f * (y2 + z2);
    R[3] = 2.f * (xy + wz);
    R[6] = 2.f * (xz - wy);
    R[1] = 2.f * (xy - wz);
    R[4] = 1.f - 2.f * (x2 + z2);
    R[7] = 2.f * (yz + wx);
    R[2] = 2.f * (xz + wy);
    R[5] = 2.f * (yz - wx);
    R[8] = 1.f - 2.f * (x2 + y2);
    if (nq)
        nq[0] = w, nq[1] = x, nq[2] = y, nq[3] = z, nq[4] = inv_norm;
}

struct ProjIntermediates
{
    float mc[3];
    float covar_c[9];
    float Rq[9];
    float M[9];
    float covar[9];
    float nq[5];
};

// A3 + A4. `scale` = exp(log-scale) already applied.
__device__ __forceinline__ Proj project_one(const float *mean, const float *quat, const float *scale, const CamParams &c, ProjIntermediates *keep)
{
    Proj o;
    o.radius = 0;
    o.m2x = o.m2y = o.depth = o.ca = o.cb = o.cc = 0.f;
    float mc[3];
#pragma unroll
    for (int i = 0; i < 3; i++)
        mc[i] = ((c.R[i * 3 + 0] * mean[0] + c.R[i * 3 + 1] * mean[1]) + c.R[i * 3 + 2] * mean[2]) + c.t[i];
    float Rq[9], M[9], covar[9], tmp[9], cc[9], nq[5];
    quat_to_rotmat(quat, Rq, nq);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            M[i * 3 + j] = Rq[i * 3 + j] * scale[j];
    mm3_bt(M, M, covar);
    mm3(c.R, covar, tmp);
    mm3_bt(tmp, c.R, cc);
    if (keep)
    {
#pragma unroll
        for (int i = 0; i < 9; i++)
            keep->covar_c[i] = cc[i], keep->Rq[i] = Rq[i], keep->M[i] = M[i], keep->covar[i] = covar[i];
#pragma unroll
        for (int i = 0; i < 3; i++)
            keep->mc[i] = mc[i];
#pragma unroll
        for (int i = 0; i < 5; i++)
            keep->nq[i] = nq[i];
    }
    if (mc[2] < c.near_plane || mc[2] > c.far_plane)
        return o;
    float x = mc[0], y = mc[1], z = mc[2];
    float rz = 1.f / z;
    float rz2 = rz * rz;
    float tx = z * fminf(c.lim_x_pos, fmaxf(-c.lim_x_neg, x * rz));
    float ty = z * fminf(c.lim_y_pos, fmaxf(-c.lim_y_neg, y * rz));
    float J00 = c.fx * rz, J11 = c.fy * rz;
    float J02 = (-c.fx) * tx * rz2, J12 = (-c.fy) * ty * rz2;
    float t00 = J00 * cc[0] + J02 * cc[6];
    float t01 = J00 * cc[1] + J02 * cc[7];
    float t02 = J00 * cc[2] + J02 * cc[8];
    float t10 = J11 * cc[3] + J12 * cc[6];
    float t11 = J11 * cc[4] + J12 * cc[7];
    float t12 = J11 * cc[5] + J12 * cc[8];
    float c00 = t00 * J00 + t02 * J02;
    float c01 = t01 * J11 + t02 * J12;
    float c10 = t10 * J00 + t12 * J02;
    float c11 = t11 * J11 + t12 * J12;
    float m2x = c.fx * x * rz + c.cx;
    float m2y = c.fy * y * rz + c.cy;
    c00 = c00 + c.eps2d;
    c11 = c11 + c.eps2d;
    float det = c00 * c11 - c01 * c10;
    if (det <= 0.f)
        return o;
    float invdet = 1.f / det;
    float b = 0.5f * (c00 + c11);
    float v1 = b + sqrtf(fmaxf(0.01f, b * b - det));
    float radius = ceilf(3.f * sqrtf(v1));
    if (radius <= c.radius_clip)
        return o;
    if (m2x + radius <= 0 || m2x - radius >= (float)c.W || m2y + radius <= 0 || m2y - radius >= (float)c.H)
        return o;
    if (!isfinite(radius))
        return o;
    int r = (int)radius;
    if (c.max_radii > 0 && r > c.max_radii)
        r = c.max_radii;
    o.radius = r;
    o.m2x = m2x, o.m2y = m2y, o.depth = z;
    o.ca = c11 * invdet, o.cb = (-c01) * invdet, o.cc = c00 * invdet;
    return o;
}

// A5: degree-3 SH colour of one channel; dir NOT normalised. Returns raw value (before +0.5 / clamp).
struct ShBasis
{
    float x, y, z, inorm;
    float b[16];
};
__device__ __forceinline__ void sh_basis(const float *dir, ShBasis &s)
{
    float inorm = 1.0f / sqrtf(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
    float x = dir[0] * inorm, y = dir[1] * inorm, z = dir[2] * inorm;
    s.x = x, s.y = y, s.z = z, s.inorm = inorm;
}
// coeffs: pointer to [16][3] (stride between k = kstride floats), channel c
__device__ __forceinline__ float sh_eval_channel(const ShBasis &s, const float *cf, int kstride)
{
    float x = s.x, y = s.y, z = s.z;
#define CF(k) cf[(k)*kstride]
    float r = 0.2820947917738781f * CF(0);
    r = r + 0.48860251190292f * (((-y) * CF(1) + z * CF(2)) - x * CF(3));
    float z2 = z * z;
    float fTmp0B = -1.092548430592079f * z;
    float fC1 = x * x - y * y;
    float fS1 = 2.f * x * y;
    float pSH6 = 0.9461746957575601f * z2 - 0.3153915652525201f;
    float pSH7 = fTmp0B * x;
    float pSH5 = fTmp0B * y;
    float pSH8 = 0.5462742152960395f * fC1;
    float pSH4 = 0.5462742152960395f * fS1;
    r = r + ((((pSH4 * CF(4) + pSH5 * CF(5)) + pSH6 * CF(6)) + pSH7 * CF(7)) + pSH8 * CF(8));
    float fTmp0C = -2.285228997322329f * z2 + 0.4570457994644658f;
    float fTmp1B = 1.445305721320277f * z;
    float fC2 = x * fC1 - y * fS1;
    float fS2 = x * fS1 + y * fC1;
    float pSH12 = z * (1.865881662950577f * z2 - 1.119528997770346f);
    float pSH13 = fTmp0C * x;
    float pSH11 = fTmp0C * y;
    float pSH14 = fTmp1B * fC1;
    float pSH10 = fTmp1B * fS1;
    float pSH15 = -0.5900435899266435f * fC2;
    float pSH9 = -0.5900435899266435f * fS2;
    r = r + ((((((pSH9 * CF(9) + pSH10 * CF(10)) + pSH11 * CF(11)) + pSH12 * CF(12)) + pSH13 * CF(13)) + pSH14 * CF(14)) + pSH15 * CF(15));
#undef CF
    return r;
}

// A10: VJP of the degree-3 SH colour for all three channels at once.
// v_col[3]: gradient w.r.t. raw SH colour; coeffs [16][3] via (kstride); outputs v_coeff basis (16 values, multiply by v_col[c]) and v_dir.
__device__ __forceinline__ void sh_vjp(const ShBasis &s, const float *cf /* [16*3] k-major */, int kstride, const float *v_col, float *basis16,
                                       float *v_dir)
{
    float x = s.x, y = s.y, z = s.z;
    float z2 = z * z;
    float fTmp0B = -1.092548430592079f * z;
    float fC1 = x * x - y * y;
    float fS1 = 2.f * x * y;
    float pSH6 = 0.9461746957575601f * z2 - 0.3153915652525201f;
    float pSH7 = fTmp0B * x, pSH5 = fTmp0B * y;
    float pSH8 = 0.5462742152960395f * fC1, pSH4 = 0.5462742152960395f * fS1;
    float fTmp0C = -2.285228997322329f * z2 + 0.4570457994644658f;
    float fTmp1B = 1.445305721320277f * z;
    float fC2 = x * fC1 - y * fS1;
    float fS2 = x * fS1 + y * fC1;
    float pSH12 = z * (1.865881662950577f * z2 - 1.119528997770346f);
    float pSH13 = fTmp0C * x, pSH11 = fTmp0C * y;
    float pSH14 = fTmp1B * fC1, pSH10 = fTmp1B * fS1;
    float pSH15 = -0.5900435899266435f * fC2, pSH9 = -0.5900435899266435f * fS2;
    basis16[0] = 0.2820947917738781f;
    basis16[1] = -0.48860251190292f * y;
    basis16[2] = 0.48860251190292f * z;
    basis16[3] = -0.48860251190292f * x;
    basis16[4] = pSH4, basis16[5] = pSH5, basis16[6] = pSH6, basis16[7] = pSH7, basis16[8] = pSH8;
    basis16[9] = pSH9, basis16[10] = pSH10, basis16[11] = pSH11, basis16[12] = pSH12, basis16[13] = pSH13, basis16[14] = pSH14,
    basis16[15] = pSH15;

    float fTmp0B_z = -1.092548430592079f;
    float fC1_x = 2.f * x, fC1_y = -2.f * y, fS1_x = 2.f * y, fS1_y = 2.f * x;
    float pSH6_z = 2.f * 0.9461746957575601f * z;
    float pSH7_x = fTmp0B, pSH7_z = fTmp0B_z * x, pSH5_y = fTmp0B, pSH5_z = fTmp0B_z * y;
    float pSH8_x = 0.5462742152960395f * fC1_x, pSH8_y = 0.5462742152960395f * fC1_y;
    float pSH4_x = 0.5462742152960395f * fS1_x, pSH4_y = 0.5462742152960395f * fS1_y;
    float fTmp0C_z = -2.285228997322329f * 2.f * z;
    float fTmp1B_z = 1.445305721320277f;
    float fC2_x = fC1 + x * fC1_x - y * fS1_x;
    float fC2_y = x * fC1_y - fS1 - y * fS1_y;
    float fS2_x = fS1 + x * fS1_x + y * fC1_x;
    float fS2_y = x * fS1_y + fC1 + y * fC1_y;
    float pSH12_z = 3.f * 1.865881662950577f * z2 - 1.119528997770346f;
    float pSH13_x = fTmp0C, pSH13_z = fTmp0C_z * x, pSH11_y = fTmp0C, pSH11_z = fTmp0C_z * y;
    float pSH14_x = fTmp1B * fC1_x, pSH14_y = fTmp1B * fC1_y, pSH14_z = fTmp1B_z * fC1;
    float pSH10_x = fTmp1B * fS1_x, pSH10_y = fTmp1B * fS1_y, pSH10_z = fTmp1B_z * fS1;
    float pSH15_x = -0.5900435899266435f * fC2_x, pSH15_y = -0.5900435899266435f * fC2_y;
    float pSH9_x = -0.5900435899266435f * fS2_x, pSH9_y = -0.5900435899266435f * fS2_y;

    float v_x = 0.f, v_y = 0.f, v_z = 0.f;
#pragma unroll
    for (int c = 0; c < 3; c++)
    {
        float vc = v_col[c];
#define CF(k) cf[(k)*kstride + c]
        v_x += -0.48860251190292f * CF(3) * vc;
        v_y += -0.48860251190292f * CF(1) * vc;
        v_z += 0.48860251190292f * CF(2) * vc;
        v_x += vc * (pSH4_x * CF(4) + pSH8_x * CF(8) + pSH7_x * CF(7));
        v_y += vc * (pSH4_y * CF(4) + pSH8_y * CF(8) + pSH5_y * CF(5));
        v_z += vc * (pSH6_z * CF(6) + pSH7_z * CF(7) + pSH5_z * CF(5));
        v_x += vc * (pSH9_x * CF(9) + pSH15_x * CF(15) + pSH10_x * CF(10) + pSH14_x * CF(14) + pSH13_x * CF(13));
        v_y += vc * (pSH9_y * CF(9) + pSH15_y * CF(15) + pSH10_y * CF(10) + pSH14_y * CF(14) + pSH11_y * CF(11));
        v_z += vc * (pSH12_z * CF(12) + pSH13_z * CF(13) + pSH11_z * CF(11) + pSH14_z * CF(14) + pSH10_z * CF(10));
#undef CF
    }
    float dotp = v_x * x + v_y * y + v_z * z;
    v_dir[0] = (v_x - dotp * x) * s.inorm;
    v_dir[1] = (v_y - dotp * y) * s.inorm;
    v_dir[2] = (v_z - dotp * z) * s.inorm;
}

// A11: VJP of the projection. Inputs: forward intermediates (recomputed by project_one with keep), the conic, and the
// raster-side gradients. Outputs v_mean (world), v_quat (raw quaternion), v_scale (w.r.t. real scale).
__device__ __forceinline__ void project_vjp(const float *scale, const CamParams &c, const ProjIntermediates &k, float ca, float cb, float cc,
                                            float v_m2x, float v_m2y, float v_depth, float v_ca, float v_cb, float v_cc, float *v_mean,
                                            float *v_quat, float *v_scale)
{
    // inverse_vjp: v_cov2d = -Minv * v_Minv * Minv with v_Minv = [[v_a, v_b/2],[v_b/2, v_c]]
    float P00 = ca, P01 = cb, P11 = cc;
    float G00 = v_ca, G01 = v_cb * .5f, G11 = v_cc;
    float A00 = P00 * G00 + P01 * G01, A01 = P00 * G01 + P01 * G11;
    float A10 = P01 * G00 + P11 * G01, A11 = P01 * G01 + P11 * G11;
    float V00 = -(A00 * P00 + A01 * P01), V01 = -(A00 * P01 + A01 * P11);
    float V10 = -(A10 * P00 + A11 * P01), V11 = -(A10 * P01 + A11 * P11);

    float x = k.mc[0], y = k.mc[1], z = k.mc[2];
    float rz = 1.f / z, rz2 = rz * rz;
    float tx = z * fminf(c.lim_x_pos, fmaxf(-c.lim_x_neg, x * rz));
    float ty = z * fminf(c.lim_y_pos, fmaxf(-c.lim_y_neg, y * rz));
    float J00 = c.fx * rz, J11 = c.fy * rz, J02 = -c.fx * tx * rz2, J12 = -c.fy * ty * rz2;
    // v_covar_c = J^T * V * J   (3x3); J = [[J00,0,J02],[0,J11,J12]]
    float VJ00 = V00 * J00, VJ01 = V01 * J11, VJ02 = V00 * J02 + V01 * J12;
    float VJ10 = V10 * J00, VJ11 = V11 * J11, VJ12 = V10 * J02 + V11 * J12;
    float vcc[9];
    vcc[0] = J00 * VJ00, vcc[1] = J00 * VJ01, vcc[2] = J00 * VJ02;
    vcc[3] = J11 * VJ10, vcc[4] = J11 * VJ11, vcc[5] = J11 * VJ12;
    vcc[6] = J02 * VJ00 + J12 * VJ10, vcc[7] = J02 * VJ01 + J12 * VJ11, vcc[8] = J02 * VJ02 + J12 * VJ12;

    float vmc[3];
    vmc[0] = c.fx * rz * v_m2x;
    vmc[1] = c.fy * rz * v_m2y;
    vmc[2] = -(c.fx * x * v_m2x + c.fy * y * v_m2y) * rz2;
    // v_J = V * J * covar_c^T + V^T * J * covar_c   (2x3)
    const float *C = k.covar_c;
    float JC[6], JCt[6]; // J*covar_c and J*covar_c^T
#pragma unroll
    for (int j = 0; j < 3; j++)
    {
        JC[0 * 3 + j] = J00 * C[0 * 3 + j] + J02 * C[2 * 3 + j];
        JC[1 * 3 + j] = J11 * C[1 * 3 + j] + J12 * C[2 * 3 + j];
        JCt[0 * 3 + j] = J00 * C[j * 3 + 0] + J02 * C[j * 3 + 2];
        JCt[1 * 3 + j] = J11 * C[j * 3 + 1] + J12 * C[j * 3 + 2];
    }
    float vJ[6];
#pragma unroll
    for (int j = 0; j < 3; j++)
    {
        vJ[0 * 3 + j] = (V00 * JCt[0 * 3 + j] + V01 * JCt[1 * 3 + j]) + (V00 * JC[0 * 3 + j] + V10 * JC[1 * 3 + j]);
        vJ[1 * 3 + j] = (V10 * JCt[0 * 3 + j] + V11 * JCt[1 * 3 + j]) + (V01 * JC[0 * 3 + j] + V11 * JC[1 * 3 + j]);
    }
    float rz3 = rz2 * rz;
    if (x * rz <= c.lim_x_pos && x * rz >= -c.lim_x_neg)
        vmc[0] += -c.fx * rz2 * vJ[2];
    else
        vmc[2] += -c.fx * rz3 * vJ[2] * tx;
    if (y * rz <= c.lim_y_pos && y * rz >= -c.lim_y_neg)
        vmc[1] += -c.fy * rz2 * vJ[5];
    else
        vmc[2] += -c.fy * rz3 * vJ[5] * ty;
    vmc[2] += -c.fx * rz2 * vJ[0] - c.fy * rz2 * vJ[4] + 2.f * c.fx * tx * rz3 * vJ[2] + 2.f * c.fy * ty * rz3 * vJ[5];
    vmc[2] += v_depth;
    // pos_world_to_cam_vjp: v_mean = R^T * v_mean_c
#pragma unroll
    for (int j = 0; j < 3; j++)
        v_mean[j] = c.R[0 * 3 + j] * vmc[0] + c.R[1 * 3 + j] * vmc[1] + c.R[2 * 3 + j] * vmc[2];
    // covar_world_to_cam_vjp: v_covar = R^T * v_covar_c * R
    float t1[9], vcov[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            t1[i * 3 + j] = c.R[0 * 3 + i] * vcc[0 * 3 + j] + c.R[1 * 3 + i] * vcc[1 * 3 + j] + c.R[2 * 3 + i] * vcc[2 * 3 + j];
    mm3(t1, c.R, vcov);
    // quat_scale_to_covar_vjp: v_M = (v_covar + v_covar^T) * M ; v_R = v_M * S ; v_scale[j] = sum_i Rq[i][j] * v_M[i][j]
    float sym[9], vM[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            sym[i * 3 + j] = vcov[i * 3 + j] + vcov[j * 3 + i];
    mm3(sym, k.M, vM);
    float vR[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            vR[i * 3 + j] = vM[i * 3 + j] * scale[j];
#pragma unroll
    for (int j = 0; j < 3; j++)
        v_scale[j] = k.Rq[0 * 3 + j] * vM[0 * 3 + j] + k.Rq[1 * 3 + j] * vM[1 * 3 + j] + k.Rq[2 * 3 + j] * vM[2 * 3 + j];
    // quat_to_rotmat_vjp; the reference indexes glm column-major v_R[c][r]: G(i,j) = our vR[j][i]
#define G(i, j) vR[(j)*3 + (i)]
    float w = k.nq[0], qx = k.nq[1], qy = k.nq[2], qz = k.nq[3], inv_norm = k.nq[4];
    float vq[4];
    vq[0] = 2.f * (qx * (G(1, 2) - G(2, 1)) + qy * (G(2, 0) - G(0, 2)) + qz * (G(0, 1) - G(1, 0)));
    vq[1] = 2.f * (-2.f * qx * (G(1, 1) + G(2, 2)) + qy * (G(0, 1) + G(1, 0)) + qz * (G(0, 2) + G(2, 0)) + w * (G(1, 2) - G(2, 1)));
    vq[2] = 2.f * (qx * (G(0, 1) + G(1, 0)) - 2.f * qy * (G(0, 0) + G(2, 2)) + qz * (G(1, 2) + G(2, 1)) + w * (G(2, 0) - G(0, 2)));
    vq[3] = 2.f * (qx * (G(0, 2) + G(2, 0)) + qy * (G(1, 2) + G(2, 1)) - 2.f * qz * (G(0, 0) + G(1, 1)) + w * (G(0, 1) - G(1, 0)));
#undef G
    float qn[4] = {w, qx, qy, qz};
    float d = vq[0] * qn[0] + vq[1] * qn[1] + vq[2] * qn[2] + vq[3] * qn[3];
#pragma unroll
    for (int i = 0; i < 4; i++)
        v_quat[i] = (vq[i] - d * qn[i]) * inv_norm;
}

// A12: one Adam update (torch::optim::Adam semantics, see oracle/gs_oracle.py adam_step); returns new parameter
struct AdamScalars
{
    float beta1, beta2, one_m_beta1, one_m_beta2, sqrt_bc2, eps;
};
__device__ __forceinline__ float adam_update(float p, float g, float &m, float &v, const AdamScalars &a, float step_size)
{
    m = m * a.beta1 + g * a.one_m_beta1;
    v = v * a.beta2 + g * g * a.one_m_beta2;
    float denom = sqrtf(v) / a.sqrt_bc2 + a.eps;
    return p - step_size * (m / denom);
}

} // namespace gs
