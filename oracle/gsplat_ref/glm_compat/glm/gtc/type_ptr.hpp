// TEST INFRASTRUCTURE (oracle/): glm/gtc/type_ptr.hpp stand-in (make_vecN / make_matN from a scalar pointer, column-major).
#pragma once
#include "../glm.hpp"

namespace glm {
template <typename T> GLM_HD vec<2, T> make_vec2(const T *p) { return vec<2, T>(p[0], p[1]); }
template <typename T> GLM_HD vec<3, T> make_vec3(const T *p) { return vec<3, T>(p[0], p[1], p[2]); }
template <typename T> GLM_HD vec<4, T> make_vec4(const T *p) { return vec<4, T>(p[0], p[1], p[2], p[3]); }
template <typename T> GLM_HD mat<2, 2, T> make_mat2(const T *p) { return mat<2, 2, T>(p[0], p[1], p[2], p[3]); }
template <typename T> GLM_HD mat<3, 3, T> make_mat3(const T *p) { return mat<3, 3, T>(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8]); }
} // namespace glm
