// Host-side rigid-pose arithmetic, operation-for-operation compatible with the reference's ORUtils::SE3Pose and
// ORUtils::Matrix4<float>::inv so that a pose pushed through SetInvM + Coerce yields the same 32 floats
// (reference: InfiniTAM/ORUtils/SE3Pose.cpp:89-151 SetModelViewFromParams, :153-241 SetParamsFromModelView,
// :317-337 GetInvM/SetInvM/Coerce; InfiniTAM/ORUtils/Matrix.h:177-245 inv, :117-123 operator*).
// Compile the including TU with -ffp-contract=off.  All matrices are column-major m[col*4+row].
#pragma once
#include <cmath>
#include <cstring>

#include "common.cuh"

#ifdef __CUDACC__
#define SE3_HD __host__ __device__
#else
#define SE3_HD
#endif

namespace se3
{

SE3_HD inline bool inverse(const Mat4 &A, Mat4 &out)
{
    // cofactor expansion on the transposed source, 2x2 products shared pairwise
    float t[12], s[16], det;
    float *d = out.m;
    for (int i = 0; i < 4; i++)
    {
        s[i] = A.m[i * 4];
        s[i + 4] = A.m[i * 4 + 1];
        s[i + 8] = A.m[i * 4 + 2];
        s[i + 12] = A.m[i * 4 + 3];
    }
    t[0] = s[10] * s[15]; t[1] = s[11] * s[14]; t[2] = s[9] * s[15]; t[3] = s[11] * s[13];
    t[4] = s[9] * s[14];  t[5] = s[10] * s[13]; t[6] = s[8] * s[15]; t[7] = s[11] * s[12];
    t[8] = s[8] * s[14];  t[9] = s[10] * s[12]; t[10] = s[8] * s[13]; t[11] = s[9] * s[12];

    d[0] = (t[0] * s[5] + t[3] * s[6] + t[4] * s[7]) - (t[1] * s[5] + t[2] * s[6] + t[5] * s[7]);
    d[1] = (t[1] * s[4] + t[6] * s[6] + t[9] * s[7]) - (t[0] * s[4] + t[7] * s[6] + t[8] * s[7]);
    d[2] = (t[2] * s[4] + t[7] * s[5] + t[10] * s[7]) - (t[3] * s[4] + t[6] * s[5] + t[11] * s[7]);
    d[3] = (t[5] * s[4] + t[8] * s[5] + t[11] * s[6]) - (t[4] * s[4] + t[9] * s[5] + t[10] * s[6]);

    det = s[0] * d[0] + s[1] * d[1] + s[2] * d[2] + s[3] * d[3];
    if (det == 0.0f)
        return false;

    d[4] = (t[1] * s[1] + t[2] * s[2] + t[5] * s[3]) - (t[0] * s[1] + t[3] * s[2] + t[4] * s[3]);
    d[5] = (t[0] * s[0] + t[7] * s[2] + t[8] * s[3]) - (t[1] * s[0] + t[6] * s[2] + t[9] * s[3]);
    d[6] = (t[3] * s[0] + t[6] * s[1] + t[11] * s[3]) - (t[2] * s[0] + t[7] * s[1] + t[10] * s[3]);
    d[7] = (t[4] * s[0] + t[9] * s[1] + t[10] * s[2]) - (t[5] * s[0] + t[8] * s[1] + t[11] * s[2]);

    t[0] = s[2] * s[7]; t[1] = s[3] * s[6]; t[2] = s[1] * s[7]; t[3] = s[3] * s[5];
    t[4] = s[1] * s[6]; t[5] = s[2] * s[5]; t[6] = s[0] * s[7]; t[7] = s[3] * s[4];
    t[8] = s[0] * s[6]; t[9] = s[2] * s[4]; t[10] = s[0] * s[5]; t[11] = s[1] * s[4];

    d[8] = (t[0] * s[13] + t[3] * s[14] + t[4] * s[15]) - (t[1] * s[13] + t[2] * s[14] + t[5] * s[15]);
    d[9] = (t[1] * s[12] + t[6] * s[14] + t[9] * s[15]) - (t[0] * s[12] + t[7] * s[14] + t[8] * s[15]);
    d[10] = (t[2] * s[12] + t[7] * s[13] + t[10] * s[15]) - (t[3] * s[12] + t[6] * s[13] + t[11] * s[15]);
    d[11] = (t[5] * s[12] + t[8] * s[13] + t[11] * s[14]) - (t[4] * s[12] + t[9] * s[13] + t[10] * s[14]);
    d[12] = (t[2] * s[10] + t[5] * s[11] + t[1] * s[9]) - (t[4] * s[11] + t[0] * s[9] + t[3] * s[10]);
    d[13] = (t[8] * s[11] + t[0] * s[8] + t[7] * s[10]) - (t[6] * s[10] + t[9] * s[11] + t[1] * s[8]);
    d[14] = (t[6] * s[9] + t[11] * s[11] + t[3] * s[8]) - (t[10] * s[11] + t[2] * s[8] + t[7] * s[9]);
    d[15] = (t[10] * s[10] + t[4] * s[8] + t[9] * s[9]) - (t[8] * s[9] + t[11] * s[10] + t[5] * s[8]);

    det = 1.0f / det;
    for (int i = 0; i < 16; i++)
        d[i] *= det;
    return true;
}

// r = a * b with the reference's accumulate-from-zero order
SE3_HD inline Mat4 mul(const Mat4 &a, const Mat4 &b)
{
    Mat4 r;
    for (int i = 0; i < 16; i++)
        r.m[i] = 0.0f;
    for (int x = 0; x < 4; x++)
        for (int y = 0; y < 4; y++)
            for (int k = 0; k < 4; k++)
                r.m[x * 4 + y] += a.m[k * 4 + y] * b.m[x * 4 + k];
    return r;
}

struct Pose
{
    float p[6]; // tx ty tz rx ry rz
    Mat4 M;     // world -> camera

    SE3_HD Pose() { set_params(0, 0, 0, 0, 0, 0); }

    SE3_HD void set_params(float tx, float ty, float tz, float rx, float ry, float rz)
    {
        p[0] = tx, p[1] = ty, p[2] = tz, p[3] = rx, p[4] = ry, p[5] = rz;
        matrix_from_params();
    }

    SE3_HD static void cross3(const float *a, const float *b, float *c)
    {
        c[0] = a[1] * b[2] - a[2] * b[1];
        c[1] = a[2] * b[0] - a[0] * b[2];
        c[2] = a[0] * b[1] - a[1] * b[0];
    }
    SE3_HD static float dot3(const float *a, const float *b)
    {
        float r = 0;
        for (int i = 0; i < 3; i++)
            r += a[i] * b[i];
        return r;
    }

    // exponential map (Rodrigues with Taylor branches)
    SE3_HD void matrix_from_params()
    {
        const float one_6th = 1.0f / 6.0f, one_20th = 1.0f / 20.0f;
        float w[3] = {p[3], p[4], p[5]}, t[3] = {p[0], p[1], p[2]};
        float theta_sq = dot3(w, w);
        float theta = sqrt(theta_sq);
        float A, B;
        float T[3], cr[3];
        cross3(w, t, cr);
        if (theta_sq < 1e-8f)
        {
            A = 1.0f - one_6th * theta_sq;
            B = 0.5f;
            T[0] = t[0] + 0.5f * cr[0], T[1] = t[1] + 0.5f * cr[1], T[2] = t[2] + 0.5f * cr[2];
        }
        else
        {
            float Cc;
            if (theta_sq < 1e-6f)
            {
                Cc = one_6th * (1.0f - one_20th * theta_sq);
                A = 1.0f - theta_sq * Cc;
                B = 0.5f - 0.25f * one_6th * theta_sq;
            }
            else
            {
                float inv_theta = 1.0f / theta;
                A = sinf(theta) * inv_theta;
                B = (1.0f - cosf(theta)) * (inv_theta * inv_theta);
                Cc = (1.0f - A) * (inv_theta * inv_theta);
            }
            float c2[3];
            cross3(w, cr, c2);
            T[0] = t[0] + B * cr[0] + Cc * c2[0], T[1] = t[1] + B * cr[1] + Cc * c2[1], T[2] = t[2] + B * cr[2] + Cc * c2[2];
        }
        float wx2 = w[0] * w[0], wy2 = w[1] * w[1], wz2 = w[2] * w[2];
        float R[9]; // column-major 3x3: R[row + 3*col]
        R[0 + 3 * 0] = 1.0f - B * (wy2 + wz2);
        R[1 + 3 * 1] = 1.0f - B * (wx2 + wz2);
        R[2 + 3 * 2] = 1.0f - B * (wx2 + wy2);
        float a, b;
        a = A * w[2], b = B * (w[0] * w[1]);
        R[0 + 3 * 1] = b - a;
        R[1 + 3 * 0] = b + a;
        a = A * w[1], b = B * (w[0] * w[2]);
        R[0 + 3 * 2] = b + a;
        R[2 + 3 * 0] = b - a;
        a = A * w[0], b = B * (w[1] * w[2]);
        R[1 + 3 * 2] = b - a;
        R[2 + 3 * 1] = b + a;
        for (int c = 0; c < 3; c++)
            for (int r = 0; r < 3; r++)
                M.m[r + 4 * c] = R[r + 3 * c];
        M.m[0 + 4 * 3] = T[0], M.m[1 + 4 * 3] = T[1], M.m[2 + 4 * 3] = T[2];
        M.m[3 + 4 * 0] = 0.0f, M.m[3 + 4 * 1] = 0.0f, M.m[3 + 4 * 2] = 0.0f, M.m[3 + 4 * 3] = 1.0f;
    }

    // logarithm map
    SE3_HD void params_from_matrix()
    {
        float R[9], T[3];
        for (int c = 0; c < 3; c++)
            for (int r = 0; r < 3; r++)
                R[r + 3 * c] = M.m[r + 4 * c];
        T[0] = M.m[12], T[1] = M.m[13], T[2] = M.m[14];
        float rr[3];
        // ORUtils::Matrix3 names members mCR: m00 = m[0], m11 = m[4], m22 = m[8]
        float cos_angle = (R[0] + R[4] + R[8] - 1.0f) * 0.5f;
        rr[0] = (R[2 + 3 * 1] - R[1 + 3 * 2]) * 0.5f;
        rr[1] = (R[0 + 3 * 2] - R[2 + 3 * 0]) * 0.5f;
        rr[2] = (R[1 + 3 * 0] - R[0 + 3 * 1]) * 0.5f;
        float sin_angle_abs = sqrt(dot3(rr, rr));
        if (cos_angle > M_SQRT1_2)
        {
            if (sin_angle_abs)
            {
                float q = asinf(sin_angle_abs) / sin_angle_abs;
                rr[0] *= q, rr[1] *= q, rr[2] *= q;
            }
        }
        else
        {
            if (cos_angle > -M_SQRT1_2)
            {
                float q = acosf(cos_angle) / sin_angle_abs;
                rr[0] *= q, rr[1] *= q, rr[2] *= q;
            }
            else
            {
                float angle = (float)M_PI - asinf(sin_angle_abs);
                float d0 = R[0] - cos_angle, d1 = R[4] - cos_angle, d2 = R[8] - cos_angle;
                float r2[3];
                if (fabsf(d0) > fabsf(d1) && fabsf(d0) > fabsf(d2))
                {
                    r2[0] = d0, r2[1] = (R[1 + 3 * 0] + R[0 + 3 * 1]) * 0.5f, r2[2] = (R[0 + 3 * 2] + R[2 + 3 * 0]) * 0.5f;
                }
                else if (fabsf(d1) > fabsf(d2))
                {
                    r2[0] = (R[1 + 3 * 0] + R[0 + 3 * 1]) * 0.5f, r2[1] = d1, r2[2] = (R[2 + 3 * 1] + R[1 + 3 * 2]) * 0.5f;
                }
                else
                {
                    r2[0] = (R[0 + 3 * 2] + R[2 + 3 * 0]) * 0.5f, r2[1] = (R[2 + 3 * 1] + R[1 + 3 * 2]) * 0.5f, r2[2] = d2;
                }
                if (dot3(r2, rr) < 0.0f)
                    r2[0] *= -1.0f, r2[1] *= -1.0f, r2[2] *= -1.0f;
                float len = sqrt(dot3(r2, r2));
                if (len == 0)
                    r2[0] = r2[1] = r2[2] = 0;
                else
                    r2[0] /= len, r2[1] /= len, r2[2] /= len;
                rr[0] = angle * r2[0], rr[1] = angle * r2[1], rr[2] = angle * r2[2];
            }
        }
        float shtot = 0.5f;
        float theta = sqrt(dot3(rr, rr));
        if (theta > 0.00001f)
            shtot = sinf(theta * 0.5f) / theta;

        Pose half;
        half.set_params(0.0f, 0.0f, 0.0f, rr[0] * -0.5f, rr[1] * -0.5f, rr[2] * -0.5f);
        float rt[3];
        // Matrix3 * Vector3 (Matrix.h:305-311): r[i] = m[i]*v0 + m[3+i]*v1 + m[6+i]*v2
        for (int i = 0; i < 3; i++)
            rt[i] = half.M.m[i] * T[0] + half.M.m[4 + i] * T[1] + half.M.m[8 + i] * T[2];

        if (theta > 0.001f)
        {
            float denom = dot3(rr, rr);
            float param = dot3(T, rr) * (1 - 2 * shtot) / denom;
            rt[0] -= rr[0] * param, rt[1] -= rr[1] * param, rt[2] -= rr[2] * param;
        }
        else
        {
            float param = dot3(T, rr) / 24;
            rt[0] -= rr[0] * param, rt[1] -= rr[1] * param, rt[2] -= rr[2] * param;
        }
        rt[0] /= 2 * shtot, rt[1] /= 2 * shtot, rt[2] /= 2 * shtot;
        p[3] = rr[0], p[4] = rr[1], p[5] = rr[2];
        p[0] = rt[0], p[1] = rt[1], p[2] = rt[2];
    }

    SE3_HD void set_M(const Mat4 &m)
    {
        M = m;
        params_from_matrix();
    }
    SE3_HD void set_invM(const Mat4 &invM)
    {
        inverse(invM, M);
        params_from_matrix();
    }
    SE3_HD Mat4 get_invM() const
    {
        Mat4 r;
        inverse(M, r);
        return r;
    }
    SE3_HD void coerce()
    {
        params_from_matrix();
        matrix_from_params();
    }
};

} // namespace se3
