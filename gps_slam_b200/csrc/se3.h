// Rigid-pose arithmetic of the tracker / engine state (host and device), written in this repository's own form.
//
// WHAT must match the reference is the *sequence of fp32 roundings* -- a pose that went through SetInvM + Coerce in the reference
// has to come out as the same 32 floats here, otherwise every voxel integrated with it differs (behaviour studied in
// InfiniTAM/ORUtils/SE3Pose.cpp:89-241, :317-337 and InfiniTAM/ORUtils/Matrix.h:117-123, :177-245).  HOW it is written is ours:
//   * the 4x4 inverse is the classical adjugate from twelve shared 2x2 products per half, driven by an index table
//     (kAdj below) instead of sixteen spelled-out expressions;
//   * exp / log on SO(3) are written from the Rodrigues formulas over cyclic index triples, with the series coefficients in one
//     helper (so3_coeffs) shared by the exponential map and by the half-rotation the logarithm needs;
//   * the rounding order every formula commits to is stated next to it.
// Compile the including TU with -ffp-contract=off (host) / -fmad=false (device).  Matrices are column-major m[col*4+row].
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

struct V3
{
    float v[3];
    SE3_HD float &operator[](int i) { return v[i]; }
    SE3_HD float operator[](int i) const { return v[i]; }
};

// rounding order: ((0 + a0 b0) + a1 b1) + a2 b2
SE3_HD inline float dot(const V3 &a, const V3 &b)
{
    float acc = 0.0f;
    for (int i = 0; i < 3; i++)
        acc += a[i] * b[i];
    return acc;
}

// c_k = a_i b_j - a_j b_i over the cyclic triples (k, i, j)
SE3_HD inline V3 cross(const V3 &a, const V3 &b)
{
    V3 c;
    for (int k = 0; k < 3; k++)
    {
        const int i = (k + 1) % 3, j = (k + 2) % 3;
        c[k] = a[i] * b[j] - a[j] * b[i];
    }
    return c;
}

SE3_HD inline V3 scaled(const V3 &a, float s)
{
    V3 r;
    for (int i = 0; i < 3; i++)
        r[i] = a[i] * s;
    return r;
}

// ---- 4x4 inverse ---------------------------------------------------------------------------------------------------------------
// Adjugate over the transposed source S (S[4 r + c] = A(c, r)).  Two sets of twelve 2x2 products are formed, one from rows 2-3 of S
// (feeding adjugate entries 0-7) and one from rows 0-1 (entries 8-15); product k of a set is S[base + kPairA[k]] * S[base + kPairB[k]].
// Entry e is (p0 s0 + p1 s1 + p2 s2) - (n0 t0 + n1 t1 + n2 t2), sums left to right; kAdj[e] lists, as (product, element) index
// pairs, the three positive and the three negative terms.  det = sum_{e<4} S[e] adj[e]; the result is adj * (1 / det).
struct AdjTerm
{
    unsigned char pos[3][2], neg[3][2];
};

SE3_HD inline bool inverse(const Mat4 &A, Mat4 &out)
{
    constexpr unsigned char kPairA[12] = {2, 3, 1, 3, 1, 2, 0, 3, 0, 2, 0, 1};
    constexpr unsigned char kPairB[12] = {7, 6, 7, 5, 6, 5, 7, 4, 6, 4, 5, 4};
    constexpr AdjTerm kAdj[16] = {
        {{{0, 5}, {3, 6}, {4, 7}}, {{1, 5}, {2, 6}, {5, 7}}},         {{{1, 4}, {6, 6}, {9, 7}}, {{0, 4}, {7, 6}, {8, 7}}},
        {{{2, 4}, {7, 5}, {10, 7}}, {{3, 4}, {6, 5}, {11, 7}}},       {{{5, 4}, {8, 5}, {11, 6}}, {{4, 4}, {9, 5}, {10, 6}}},
        {{{1, 1}, {2, 2}, {5, 3}}, {{0, 1}, {3, 2}, {4, 3}}},         {{{0, 0}, {7, 2}, {8, 3}}, {{1, 0}, {6, 2}, {9, 3}}},
        {{{3, 0}, {6, 1}, {11, 3}}, {{2, 0}, {7, 1}, {10, 3}}},       {{{4, 0}, {9, 1}, {10, 2}}, {{5, 0}, {8, 1}, {11, 2}}},
        {{{0, 13}, {3, 14}, {4, 15}}, {{1, 13}, {2, 14}, {5, 15}}},   {{{1, 12}, {6, 14}, {9, 15}}, {{0, 12}, {7, 14}, {8, 15}}},
        {{{2, 12}, {7, 13}, {10, 15}}, {{3, 12}, {6, 13}, {11, 15}}}, {{{5, 12}, {8, 13}, {11, 14}}, {{4, 12}, {9, 13}, {10, 14}}},
        {{{2, 10}, {5, 11}, {1, 9}}, {{4, 11}, {0, 9}, {3, 10}}},     {{{8, 11}, {0, 8}, {7, 10}}, {{6, 10}, {9, 11}, {1, 8}}},
        {{{6, 9}, {11, 11}, {3, 8}}, {{10, 11}, {2, 8}, {7, 9}}},     {{{10, 10}, {4, 8}, {9, 9}}, {{8, 9}, {11, 10}, {5, 8}}}};
    float S[16], prod[12], adj[16];
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++)
            S[4 * r + c] = A.m[4 * c + r];
    auto entry = [&](int e)
    {
        const AdjTerm &t = kAdj[e];
        const float plus = prod[t.pos[0][0]] * S[t.pos[0][1]] + prod[t.pos[1][0]] * S[t.pos[1][1]] + prod[t.pos[2][0]] * S[t.pos[2][1]];
        const float minus = prod[t.neg[0][0]] * S[t.neg[0][1]] + prod[t.neg[1][0]] * S[t.neg[1][1]] + prod[t.neg[2][0]] * S[t.neg[2][1]];
        adj[e] = plus - minus;
    };
    for (int k = 0; k < 12; k++)
        prod[k] = S[8 + kPairA[k]] * S[8 + kPairB[k]];
    for (int e = 0; e < 4; e++)
        entry(e);
    const float det = S[0] * adj[0] + S[1] * adj[1] + S[2] * adj[2] + S[3] * adj[3];
    if (det == 0.0f)
        return false;
    for (int e = 4; e < 8; e++)
        entry(e);
    for (int k = 0; k < 12; k++)
        prod[k] = S[kPairA[k]] * S[kPairB[k]];
    for (int e = 8; e < 16; e++)
        entry(e);
    const float rdet = 1.0f / det;
    for (int e = 0; e < 16; e++)
        out.m[e] = adj[e] * rdet;
    return true;
}

// r = a b; every entry is accumulated from zero over k = 0..3
SE3_HD inline Mat4 mul(const Mat4 &a, const Mat4 &b)
{
    Mat4 r;
    for (int col = 0; col < 4; col++)
        for (int row = 0; row < 4; row++)
        {
            float acc = 0.0f;
            for (int k = 0; k < 4; k++)
                acc += a.m[k * 4 + row] * b.m[col * 4 + k];
            r.m[col * 4 + row] = acc;
        }
    return r;
}

// ---- SO(3) / SE(3) -------------------------------------------------------------------------------------------------------------
// Series coefficients of the exponential map for |w|^2 = th2:  A = sin th / th,  B = (1 - cos th) / th^2,  C = (1 - A) / th^2,
// by Taylor expansion below 1e-6 (and, below 1e-8, with the C term dropped altogether: hasC = false).
struct So3Coeffs
{
    float A, B, C;
    bool hasC;
};

SE3_HD inline So3Coeffs so3_coeffs(float th2)
{
    const float sixth = 1.0f / 6.0f, twentieth = 1.0f / 20.0f;
    So3Coeffs k;
    k.hasC = !(th2 < 1e-8f);
    if (!k.hasC)
    {
        k.A = 1.0f - sixth * th2;
        k.B = 0.5f;
        k.C = 0.0f;
    }
    else if (th2 < 1e-6f)
    {
        k.C = sixth * (1.0f - twentieth * th2);
        k.A = 1.0f - th2 * k.C;
        k.B = 0.5f - 0.25f * sixth * th2;
    }
    else
    {
        const float th = sqrt(th2), rth = 1.0f / th;
        k.A = sinf(th) * rth;
        k.B = (1.0f - cosf(th)) * (rth * rth);
        k.C = (1.0f - k.A) * (rth * rth);
    }
    return k;
}

// R = I + A [w]x + B [w]x^2 written into the upper-left 3x3 of M:
//   R(k,k) = 1 - B (w_i^2 + w_j^2),  R(i,j) = B (w_i w_j) - A w_k,  R(j,i) = B (w_i w_j) + A w_k   over cyclic (k, i, j)
SE3_HD inline void so3_matrix(const V3 &w, const So3Coeffs &k, Mat4 &M)
{
    float sq[3];
    for (int i = 0; i < 3; i++)
        sq[i] = w[i] * w[i];
    for (int c = 0; c < 3; c++)
    {
        const int i = (c + 1) % 3, j = (c + 2) % 3;
        M.m[c + 4 * c] = 1.0f - k.B * (sq[i] + sq[j]);
        const float skew = k.A * w[c], sym = k.B * (w[i] * w[j]);
        M.m[i + 4 * j] = sym - skew;
        M.m[j + 4 * i] = sym + skew;
    }
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

    // exponential map: translation = t + B (w x t) + C (w x (w x t)), summed in that order
    SE3_HD void matrix_from_params()
    {
        const V3 t = {{p[0], p[1], p[2]}}, w = {{p[3], p[4], p[5]}};
        const So3Coeffs k = so3_coeffs(dot(w, w));
        const V3 wt = cross(w, t);
        V3 trans;
        for (int i = 0; i < 3; i++)
            trans[i] = t[i] + k.B * wt[i];
        if (k.hasC)
        {
            const V3 wwt = cross(w, wt);
            for (int i = 0; i < 3; i++)
                trans[i] += k.C * wwt[i];
        }
        so3_matrix(w, k, M);
        for (int i = 0; i < 3; i++)
        {
            M.m[12 + i] = trans[i];
            M.m[4 * i + 3] = 0.0f;
        }
        M.m[15] = 1.0f;
    }

    // logarithm map.  Rotation vector from the antisymmetric part (scaled by asin / acos of the angle, or, within 45 degrees of
    // pi, from the dominant column of the symmetric part); translation = V^-1 T through the half rotation:
    //   u = R(-w/2) T - w (w.T) (1 - 2 s) / |w|^2,   s = sin(|w|/2) / |w|,   result u / (2 s)
    SE3_HD void params_from_matrix()
    {
        auto R = [&](int row, int col) { return M.m[row + 4 * col]; };
        const V3 T = {{M.m[12], M.m[13], M.m[14]}};
        const float cosAng = (R(0, 0) + R(1, 1) + R(2, 2) - 1.0f) * 0.5f;
        V3 w;
        for (int k = 0; k < 3; k++)
        {
            const int i = (k + 1) % 3, j = (k + 2) % 3;
            w[k] = (R(j, i) - R(i, j)) * 0.5f;
        }
        const float sinAbs = sqrt(dot(w, w));
        if (cosAng > M_SQRT1_2)
        {
            if (sinAbs)
                w = scaled(w, asinf(sinAbs) / sinAbs);
        }
        else if (cosAng > -M_SQRT1_2)
            w = scaled(w, acosf(cosAng) / sinAbs);
        else
        {
            const float angle = (float)M_PI - asinf(sinAbs);
            float dg[3];
            for (int i = 0; i < 3; i++)
                dg[i] = R(i, i) - cosAng;
            const int lead = (fabsf(dg[0]) > fabsf(dg[1]) && fabsf(dg[0]) > fabsf(dg[2])) ? 0 : (fabsf(dg[1]) > fabsf(dg[2]) ? 1 : 2);
            V3 axis;
            for (int i = 0; i < 3; i++)
                axis[i] = (i == lead) ? dg[lead] : (R(i, lead) + R(lead, i)) * 0.5f;
            if (dot(axis, w) < 0.0f)
                axis = scaled(axis, -1.0f);
            const float len = sqrt(dot(axis, axis));
            for (int i = 0; i < 3; i++)
                axis[i] = (len == 0) ? 0.0f : axis[i] / len;
            w = scaled(axis, angle);
        }
        const float th2 = dot(w, w);
        const float th = sqrt(th2);
        float s = 0.5f;
        if (th > 0.00001f)
            s = sinf(th * 0.5f) / th;

        const V3 wHalf = scaled(w, -0.5f);
        Mat4 half;
        so3_matrix(wHalf, so3_coeffs(dot(wHalf, wHalf)), half);
        V3 u;
        for (int i = 0; i < 3; i++)
            u[i] = half.m[i] * T[0] + half.m[4 + i] * T[1] + half.m[8 + i] * T[2];
        const float along = (th > 0.001f) ? dot(T, w) * (1 - 2 * s) / dot(w, w) : dot(T, w) / 24;
        for (int i = 0; i < 3; i++)
        {
            u[i] -= w[i] * along;
            u[i] /= 2 * s;
            p[i] = u[i];
            p[3 + i] = w[i];
        }
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
    // re-orthonormalise: log then exp
    SE3_HD void coerce()
    {
        params_from_matrix();
        matrix_from_params();
    }
};

} // namespace se3
