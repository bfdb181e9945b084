// TEST INFRASTRUCTURE (oracle/): a minimal stand-in for the subset of glm 1.0.1 that the reference's gsplat kernels use
// (reference CMakeLists.txt:54-57 fetches glm from the network; it is not in this image).  Written from the GLSL / glm
// semantics, not copied from glm: column-major mat<C,R,T> = C columns of vec<R,T>, m[c][r]; mat*mat and mat*vec accumulate
// over the inner index in ascending order like glm's generic implementations.  Only used to compile the reference .cu
// files where they lie into oracle/_ref/libgsplat_ref.so (see oracle/gsplat_ref/Makefile); never part of the product.
#pragma once
#include <cmath>
#include <cuda_runtime.h>

#define GLM_HD __host__ __device__ __forceinline__

namespace glm {

typedef int length_t;

template <length_t L, typename T> struct vec;

template <typename T> struct vec<2, T> {
    T x, y;
    vec() = default;
    GLM_HD explicit vec(T s) : x(s), y(s) {}
    template <typename U> GLM_HD vec(const vec<2, U> &o) : x(T(o.x)), y(T(o.y)) {}   // converting constructor (e.g. half -> float)
    template <typename A, typename B> GLM_HD vec(A a, B b) : x(T(a)), y(T(b)) {}
    GLM_HD T &operator[](length_t i) { return (&x)[i]; }
    GLM_HD const T &operator[](length_t i) const { return (&x)[i]; }
    static GLM_HD constexpr length_t length() { return 2; }
};

template <typename T> struct vec<3, T> {
    T x, y, z;
    vec() = default;
    GLM_HD explicit vec(T s) : x(s), y(s), z(s) {}
    template <typename U> GLM_HD vec(const vec<3, U> &o) : x(T(o.x)), y(T(o.y)), z(T(o.z)) {}   // converting constructor (e.g. half -> float)
    template <typename A, typename B, typename C> GLM_HD vec(A a, B b, C c) : x(T(a)), y(T(b)), z(T(c)) {}
    GLM_HD T &operator[](length_t i) { return (&x)[i]; }
    GLM_HD const T &operator[](length_t i) const { return (&x)[i]; }
    static GLM_HD constexpr length_t length() { return 3; }
};

template <typename T> struct vec<4, T> {
    T x, y, z, w;
    vec() = default;
    GLM_HD explicit vec(T s) : x(s), y(s), z(s), w(s) {}
    template <typename U> GLM_HD vec(const vec<4, U> &o) : x(T(o.x)), y(T(o.y)), z(T(o.z)), w(T(o.w)) {}   // converting constructor (e.g. half -> float)
    template <typename A, typename B, typename C, typename D> GLM_HD vec(A a, B b, C c, D d) : x(T(a)), y(T(b)), z(T(c)), w(T(d)) {}
    GLM_HD T &operator[](length_t i) { return (&x)[i]; }
    GLM_HD const T &operator[](length_t i) const { return (&x)[i]; }
    static GLM_HD constexpr length_t length() { return 4; }
};

typedef vec<2, float> vec2;
typedef vec<3, float> vec3;
typedef vec<4, float> vec4;

// ---- vector arithmetic (component-wise) ----
#define GLM_VEC_BINOP(OP)                                                                                               \
    template <length_t L, typename T> GLM_HD vec<L, T> operator OP(const vec<L, T> &a, const vec<L, T> &b) {            \
        vec<L, T> r;                                                                                                    \
        for (length_t i = 0; i < L; ++i) r[i] = a[i] OP b[i];                                                           \
        return r;                                                                                                       \
    }                                                                                                                   \
    template <length_t L, typename T> GLM_HD vec<L, T> operator OP(const vec<L, T> &a, T s) {                           \
        vec<L, T> r;                                                                                                    \
        for (length_t i = 0; i < L; ++i) r[i] = a[i] OP s;                                                              \
        return r;                                                                                                       \
    }                                                                                                                   \
    template <length_t L, typename T> GLM_HD vec<L, T> operator OP(T s, const vec<L, T> &b) {                           \
        vec<L, T> r;                                                                                                    \
        for (length_t i = 0; i < L; ++i) r[i] = s OP b[i];                                                              \
        return r;                                                                                                       \
    }                                                                                                                   \
    template <length_t L, typename T> GLM_HD vec<L, T> &operator OP##=(vec<L, T> &a, const vec<L, T> &b) {              \
        for (length_t i = 0; i < L; ++i) a[i] OP## = b[i];                                                              \
        return a;                                                                                                       \
    }                                                                                                                   \
    template <length_t L, typename T> GLM_HD vec<L, T> &operator OP##=(vec<L, T> &a, T s) {                             \
        for (length_t i = 0; i < L; ++i) a[i] OP## = s;                                                                 \
        return a;                                                                                                       \
    }
GLM_VEC_BINOP(+)
GLM_VEC_BINOP(-)
GLM_VEC_BINOP(*)
GLM_VEC_BINOP(/)
#undef GLM_VEC_BINOP

template <length_t L, typename T> GLM_HD vec<L, T> operator-(const vec<L, T> &a) {
    vec<L, T> r;
    for (length_t i = 0; i < L; ++i) r[i] = -a[i];
    return r;
}

template <length_t L, typename T> GLM_HD T dot(const vec<L, T> &a, const vec<L, T> &b) {
    T s = a[0] * b[0];
    for (length_t i = 1; i < L; ++i) s += a[i] * b[i];
    return s;
}

template <length_t L, typename T> GLM_HD T length(const vec<L, T> &a) { return sqrt(dot(a, a)); }

GLM_HD float atan(float y, float x) { return ::atan2f(y, x); }
GLM_HD double atan(double y, double x) { return ::atan2(y, x); }

// ---- matrices: C columns, R rows, stored as columns ----
template <length_t C, length_t R, typename T> struct mat {
    typedef vec<R, T> col_type;
    col_type c[C];
    mat() = default;
    // diagonal constructor (glm: mat(s) = s * identity)
    GLM_HD explicit mat(T s) {
        for (length_t i = 0; i < C; ++i)
            for (length_t j = 0; j < R; ++j) c[i][j] = (i == j) ? s : T(0);
    }
    template <typename U> GLM_HD mat(const mat<C, R, U> &o) {   // converting constructor
        for (length_t i = 0; i < C; ++i)
            for (length_t j = 0; j < R; ++j) c[i][j] = T(o[i][j]);
    }
    // column constructors
    GLM_HD mat(const col_type &a, const col_type &b) { static_assert(C == 2, "2 columns"); c[0] = a; c[1] = b; }
    GLM_HD mat(const col_type &a, const col_type &b, const col_type &d) { static_assert(C == 3, "3 columns"); c[0] = a; c[1] = b; c[2] = d; }
    // scalar constructors, column-major order (mixed scalar types allowed, like glm's templated ones)
    template <typename A0, typename A1, typename A2, typename A3>
    GLM_HD mat(A0 a0, A1 a1, A2 a2, A3 a3) {
        static_assert(C * R == 4, "4 scalars");
        const T v[4] = {T(a0), T(a1), T(a2), T(a3)};
        for (length_t i = 0; i < C; ++i)
            for (length_t j = 0; j < R; ++j) c[i][j] = v[i * R + j];
    }
    template <typename A0, typename A1, typename A2, typename A3, typename A4, typename A5>
    GLM_HD mat(A0 a0, A1 a1, A2 a2, A3 a3, A4 a4, A5 a5) {
        static_assert(C * R == 6, "6 scalars");
        const T v[6] = {T(a0), T(a1), T(a2), T(a3), T(a4), T(a5)};
        for (length_t i = 0; i < C; ++i)
            for (length_t j = 0; j < R; ++j) c[i][j] = v[i * R + j];
    }
    template <typename A0, typename A1, typename A2, typename A3, typename A4, typename A5, typename A6, typename A7, typename A8>
    GLM_HD mat(A0 a0, A1 a1, A2 a2, A3 a3, A4 a4, A5 a5, A6 a6, A7 a7, A8 a8) {
        static_assert(C * R == 9, "9 scalars");
        const T v[9] = {T(a0), T(a1), T(a2), T(a3), T(a4), T(a5), T(a6), T(a7), T(a8)};
        for (length_t i = 0; i < C; ++i)
            for (length_t j = 0; j < R; ++j) c[i][j] = v[i * R + j];
    }
    GLM_HD col_type &operator[](length_t i) { return c[i]; }
    GLM_HD const col_type &operator[](length_t i) const { return c[i]; }
};

typedef mat<2, 2, float> mat2;
typedef mat<3, 3, float> mat3;
typedef mat<4, 4, float> mat4;

template <length_t C, length_t R, typename T> GLM_HD mat<C, R, T> operator+(const mat<C, R, T> &a, const mat<C, R, T> &b) {
    mat<C, R, T> r;
    for (length_t i = 0; i < C; ++i) r[i] = a[i] + b[i];
    return r;
}
template <length_t C, length_t R, typename T> GLM_HD mat<C, R, T> operator-(const mat<C, R, T> &a, const mat<C, R, T> &b) {
    mat<C, R, T> r;
    for (length_t i = 0; i < C; ++i) r[i] = a[i] - b[i];
    return r;
}
template <length_t C, length_t R, typename T> GLM_HD mat<C, R, T> &operator+=(mat<C, R, T> &a, const mat<C, R, T> &b) {
    for (length_t i = 0; i < C; ++i) a[i] += b[i];
    return a;
}
template <length_t C, length_t R, typename T> GLM_HD mat<C, R, T> &operator-=(mat<C, R, T> &a, const mat<C, R, T> &b) {
    for (length_t i = 0; i < C; ++i) a[i] -= b[i];
    return a;
}
template <length_t C, length_t R, typename T> GLM_HD mat<C, R, T> operator*(const mat<C, R, T> &a, T s) {
    mat<C, R, T> r;
    for (length_t i = 0; i < C; ++i) r[i] = a[i] * s;
    return r;
}
template <length_t C, length_t R, typename T> GLM_HD mat<C, R, T> operator*(T s, const mat<C, R, T> &a) {
    mat<C, R, T> r;
    for (length_t i = 0; i < C; ++i) r[i] = a[i] * s;
    return r;
}
template <length_t C, length_t R, typename T> GLM_HD mat<C, R, T> operator-(const mat<C, R, T> &a) {
    mat<C, R, T> r;
    for (length_t i = 0; i < C; ++i) r[i] = -a[i];
    return r;
}

// (R x C) * (C x C2) -> (R x C2):   r[j][i] = sum_k a[k][i] * b[j][k]
template <length_t C, length_t R, length_t C2, typename T>
GLM_HD mat<C2, R, T> operator*(const mat<C, R, T> &a, const mat<C2, C, T> &b) {
    mat<C2, R, T> r;
    for (length_t j = 0; j < C2; ++j)
        for (length_t i = 0; i < R; ++i) {
            T s = a[0][i] * b[j][0];
            for (length_t k = 1; k < C; ++k) s += a[k][i] * b[j][k];
            r[j][i] = s;
        }
    return r;
}
// matrix * column vector
template <length_t C, length_t R, typename T> GLM_HD vec<R, T> operator*(const mat<C, R, T> &a, const vec<C, T> &v) {
    vec<R, T> r;
    for (length_t i = 0; i < R; ++i) {
        T s = a[0][i] * v[0];
        for (length_t k = 1; k < C; ++k) s += a[k][i] * v[k];
        r[i] = s;
    }
    return r;
}
// row vector * matrix
template <length_t C, length_t R, typename T> GLM_HD vec<C, T> operator*(const vec<R, T> &v, const mat<C, R, T> &a) {
    vec<C, T> r;
    for (length_t j = 0; j < C; ++j) r[j] = dot(a[j], v);
    return r;
}

template <length_t C, length_t R, typename T> GLM_HD mat<R, C, T> transpose(const mat<C, R, T> &a) {
    mat<R, C, T> r;
    for (length_t i = 0; i < C; ++i)
        for (length_t j = 0; j < R; ++j) r[j][i] = a[i][j];
    return r;
}

// outerProduct(c, r) = c * r^T : column i = c * r[i]
template <length_t RC, length_t RR, typename T> GLM_HD mat<RR, RC, T> outerProduct(const vec<RC, T> &c, const vec<RR, T> &r) {
    mat<RR, RC, T> m;
    for (length_t i = 0; i < RR; ++i) m[i] = c * r[i];
    return m;
}

} // namespace glm
