// Minimal GLM-compatible header for compiling the REFERENCE's gsplat CUDA sources in this
// container (GLM itself is a vcpkg dependency that is not vendored under /root/reference and is
// not installed here; see oracle/gut_oracle.c header and DESIGN.md "Oracle").
//
// TEST INFRASTRUCTURE ONLY -- used by oracle/build_ref.py to build oracle/_ref/libgsplat_ref.so,
// the reference-kernel checker / same-box GPU baseline.  The product never includes it.
//
// It provides exactly the subset of GLM 1.0's public API that /root/reference/gsplat uses
// (vec2/3/4, mat2/3/4, mat3x2, qua; dot, cross, length, normalize, transpose, inverse,
// outerProduct, make_vec*, quat_cast, mat3_cast, rotate, slerp) with GLM's documented semantics:
// column-major matrices (m[col][row]), quaternion constructor order (w, x, y, z), q * v =
// v + 2 (w (u x v) + u x (u x v)), quat_cast selecting the largest of the four 4q^2-1 terms,
// slerp falling back to lerp when cos(theta) > 1 - epsilon.  Written from GLM's documentation
// of those operations, not copied from GLM.
#pragma once

#include <cmath>
#include <cstddef>
#include <limits>

#if defined(__CUDACC__)
#define GLMS_FN __host__ __device__ inline
#else
#define GLMS_FN inline
#endif

namespace glm {

enum qualifier { packed_highp = 0, defaultp = 0 };
typedef int length_t;

template <length_t L, typename T, qualifier Q = defaultp> struct vec;
template <length_t C, length_t R, typename T, qualifier Q = defaultp> struct mat;
template <typename T, qualifier Q = defaultp> struct qua;

// ------------------------------------------------------------------ vec2
template <typename T, qualifier Q> struct vec<2, T, Q> {
    T x, y;
    vec() = default;
    GLMS_FN explicit vec(T s) : x(s), y(s) {}
    GLMS_FN vec(T a, T b) : x(a), y(b) {}
    template <typename A, typename B> GLMS_FN vec(A a, B b) : x(static_cast<T>(a)), y(static_cast<T>(b)) {}
    template <typename U, qualifier P> GLMS_FN vec(vec<2, U, P> const &v) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)) {}
    GLMS_FN T &operator[](length_t i) { return (&x)[i]; }
    GLMS_FN T const &operator[](length_t i) const { return (&x)[i]; }
    GLMS_FN vec &operator+=(vec const &o) { x += o.x; y += o.y; return *this; }
    GLMS_FN vec &operator-=(vec const &o) { x -= o.x; y -= o.y; return *this; }
    GLMS_FN vec &operator*=(T s) { x *= s; y *= s; return *this; }
};
// ------------------------------------------------------------------ vec3
template <typename T, qualifier Q> struct vec<3, T, Q> {
    T x, y, z;
    vec() = default;
    GLMS_FN explicit vec(T s) : x(s), y(s), z(s) {}
    GLMS_FN vec(T a, T b, T c) : x(a), y(b), z(c) {}
    template <typename A, typename B, typename C_>
    GLMS_FN vec(A a, B b, C_ c) : x(static_cast<T>(a)), y(static_cast<T>(b)), z(static_cast<T>(c)) {}
    template <typename U, qualifier P>
    GLMS_FN vec(vec<3, U, P> const &v) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(v.z)) {}
    GLMS_FN T &operator[](length_t i) { return (&x)[i]; }
    GLMS_FN T const &operator[](length_t i) const { return (&x)[i]; }
    GLMS_FN vec &operator+=(vec const &o) { x += o.x; y += o.y; z += o.z; return *this; }
    GLMS_FN vec &operator-=(vec const &o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    GLMS_FN vec &operator*=(T s) { x *= s; y *= s; z *= s; return *this; }
};
// ------------------------------------------------------------------ vec4
template <typename T, qualifier Q> struct vec<4, T, Q> {
    T x, y, z, w;
    vec() = default;
    GLMS_FN explicit vec(T s) : x(s), y(s), z(s), w(s) {}
    GLMS_FN vec(T a, T b, T c, T d) : x(a), y(b), z(c), w(d) {}
    template <typename A, typename B, typename C_, typename D>
    GLMS_FN vec(A a, B b, C_ c, D d)
        : x(static_cast<T>(a)), y(static_cast<T>(b)), z(static_cast<T>(c)), w(static_cast<T>(d)) {}
    template <typename U, qualifier P>
    GLMS_FN vec(vec<4, U, P> const &v)
        : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(v.z)), w(static_cast<T>(v.w)) {}
    GLMS_FN T &operator[](length_t i) { return (&x)[i]; }
    GLMS_FN T const &operator[](length_t i) const { return (&x)[i]; }
    GLMS_FN vec &operator+=(vec const &o) { x += o.x; y += o.y; z += o.z; w += o.w; return *this; }
    GLMS_FN vec &operator-=(vec const &o) { x -= o.x; y -= o.y; z -= o.z; w -= o.w; return *this; }
    GLMS_FN vec &operator*=(T s) { x *= s; y *= s; z *= s; w *= s; return *this; }
};

// ---- component-wise vector operators (generated for L = 2, 3, 4)
#define GLMS_VEC_OPS(L, ...)                                                                                         \
    template <typename T, qualifier Q> GLMS_FN vec<L, T, Q> operator+(vec<L, T, Q> const &a, vec<L, T, Q> const &b) { \
        vec<L, T, Q> r; for (int i = 0; i < L; ++i) r[i] = a[i] + b[i]; return r; }                                  \
    template <typename T, qualifier Q> GLMS_FN vec<L, T, Q> operator-(vec<L, T, Q> const &a, vec<L, T, Q> const &b) { \
        vec<L, T, Q> r; for (int i = 0; i < L; ++i) r[i] = a[i] - b[i]; return r; }                                  \
    template <typename T, qualifier Q> GLMS_FN vec<L, T, Q> operator*(vec<L, T, Q> const &a, vec<L, T, Q> const &b) { \
        vec<L, T, Q> r; for (int i = 0; i < L; ++i) r[i] = a[i] * b[i]; return r; }                                  \
    template <typename T, qualifier Q> GLMS_FN vec<L, T, Q> operator/(vec<L, T, Q> const &a, vec<L, T, Q> const &b) { \
        vec<L, T, Q> r; for (int i = 0; i < L; ++i) r[i] = a[i] / b[i]; return r; }                                  \
    template <typename T, qualifier Q> GLMS_FN vec<L, T, Q> operator*(vec<L, T, Q> const &a, T s) {                  \
        vec<L, T, Q> r; for (int i = 0; i < L; ++i) r[i] = a[i] * s; return r; }                                     \
    template <typename T, qualifier Q> GLMS_FN vec<L, T, Q> operator*(T s, vec<L, T, Q> const &a) {                  \
        vec<L, T, Q> r; for (int i = 0; i < L; ++i) r[i] = s * a[i]; return r; }                                     \
    template <typename T, qualifier Q> GLMS_FN vec<L, T, Q> operator/(vec<L, T, Q> const &a, T s) {                  \
        vec<L, T, Q> r; for (int i = 0; i < L; ++i) r[i] = a[i] / s; return r; }                                     \
    template <typename T, qualifier Q> GLMS_FN vec<L, T, Q> operator-(vec<L, T, Q> const &a) {                       \
        vec<L, T, Q> r; for (int i = 0; i < L; ++i) r[i] = -a[i]; return r; }
GLMS_VEC_OPS(2)
GLMS_VEC_OPS(3)
GLMS_VEC_OPS(4)
#undef GLMS_VEC_OPS

template <typename T, qualifier Q> GLMS_FN T dot(vec<2, T, Q> const &a, vec<2, T, Q> const &b) { return a.x * b.x + a.y * b.y; }
template <typename T, qualifier Q> GLMS_FN T dot(vec<3, T, Q> const &a, vec<3, T, Q> const &b) {
    return a.x * b.x + a.y * b.y + a.z * b.z;
}
template <typename T, qualifier Q> GLMS_FN T dot(vec<4, T, Q> const &a, vec<4, T, Q> const &b) {
    return (a.x * b.x + a.y * b.y) + (a.z * b.z + a.w * b.w);
}
template <typename T, qualifier Q> GLMS_FN vec<3, T, Q> cross(vec<3, T, Q> const &x, vec<3, T, Q> const &y) {
    return vec<3, T, Q>(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y);
}
template <length_t L, typename T, qualifier Q> GLMS_FN T length(vec<L, T, Q> const &v) { return std::sqrt(dot(v, v)); }
template <length_t L, typename T, qualifier Q> GLMS_FN vec<L, T, Q> normalize(vec<L, T, Q> const &v) {
    return v * (static_cast<T>(1) / std::sqrt(dot(v, v)));
}

// ------------------------------------------------------------------ matrices (column-major)
#define GLMS_MAT_COMMON(C_, R_)                                                              \
    typedef vec<R_, T, Q> col_type;                                                          \
    col_type value[C_];                                                                      \
    mat() = default;                                                                         \
    GLMS_FN col_type &operator[](length_t i) { return value[i]; }                            \
    GLMS_FN col_type const &operator[](length_t i) const { return value[i]; }

template <typename T, qualifier Q> struct mat<2, 2, T, Q> {
    GLMS_MAT_COMMON(2, 2)
    GLMS_FN explicit mat(T s) { value[0] = col_type(s, 0); value[1] = col_type(0, s); }
    GLMS_FN mat(T a, T b, T c, T d) { value[0] = col_type(a, b); value[1] = col_type(c, d); }
    GLMS_FN mat(col_type const &a, col_type const &b) { value[0] = a; value[1] = b; }
};
template <typename T, qualifier Q> struct mat<3, 3, T, Q> {
    GLMS_MAT_COMMON(3, 3)
    GLMS_FN explicit mat(T s) { value[0] = col_type(s, 0, 0); value[1] = col_type(0, s, 0); value[2] = col_type(0, 0, s); }
    GLMS_FN mat(T a, T b, T c, T d, T e, T f, T g, T h, T i) {
        value[0] = col_type(a, b, c); value[1] = col_type(d, e, f); value[2] = col_type(g, h, i);
    }
    GLMS_FN mat(col_type const &a, col_type const &b, col_type const &c) { value[0] = a; value[1] = b; value[2] = c; }
};
template <typename T, qualifier Q> struct mat<4, 4, T, Q> {
    GLMS_MAT_COMMON(4, 4)
    GLMS_FN explicit mat(T s) {
        value[0] = col_type(s, 0, 0, 0); value[1] = col_type(0, s, 0, 0);
        value[2] = col_type(0, 0, s, 0); value[3] = col_type(0, 0, 0, s);
    }
};
template <typename T, qualifier Q> struct mat<3, 2, T, Q> { // 3 columns of 2 rows
    GLMS_MAT_COMMON(3, 2)
};
#undef GLMS_MAT_COMMON

template <length_t N, typename T, qualifier Q> GLMS_FN mat<N, N, T, Q> operator+(mat<N, N, T, Q> const &a, mat<N, N, T, Q> const &b) {
    mat<N, N, T, Q> r; for (int i = 0; i < N; ++i) r[i] = a[i] + b[i]; return r;
}
template <length_t N, typename T, qualifier Q> GLMS_FN mat<N, N, T, Q> &operator+=(mat<N, N, T, Q> &a, mat<N, N, T, Q> const &b) {
    for (int i = 0; i < N; ++i) a[i] += b[i]; return a;
}
template <length_t N, typename T, qualifier Q> GLMS_FN mat<N, N, T, Q> operator-(mat<N, N, T, Q> const &a) {
    mat<N, N, T, Q> r; for (int i = 0; i < N; ++i) r[i] = -a[i]; return r;
}
template <length_t N, typename T, qualifier Q> GLMS_FN mat<N, N, T, Q> operator*(T s, mat<N, N, T, Q> const &a) {
    mat<N, N, T, Q> r; for (int i = 0; i < N; ++i) r[i] = s * a[i]; return r;
}
template <length_t N, typename T, qualifier Q> GLMS_FN mat<N, N, T, Q> operator*(mat<N, N, T, Q> const &a, T s) {
    mat<N, N, T, Q> r; for (int i = 0; i < N; ++i) r[i] = a[i] * s; return r;
}
// matrix * column vector: sum_c m[c] * v[c]
// (the vector parameter is a non-deduced context, like GLM's row_type, so vec<N,double> converts)
template <length_t N, typename T, qualifier Q>
GLMS_FN vec<N, T, Q> operator*(mat<N, N, T, Q> const &m, typename mat<N, N, T, Q>::col_type const &v) {
    vec<N, T, Q> r = m[0] * v[0];
    for (int c = 1; c < N; ++c) r += m[c] * v[c];
    return r;
}
template <length_t N, typename T, qualifier Q> GLMS_FN mat<N, N, T, Q> operator*(mat<N, N, T, Q> const &a, mat<N, N, T, Q> const &b) {
    mat<N, N, T, Q> r;
    for (int c = 0; c < N; ++c) r[c] = a * b[c];
    return r;
}
template <length_t N, typename T, qualifier Q> GLMS_FN mat<N, N, T, Q> transpose(mat<N, N, T, Q> const &m) {
    mat<N, N, T, Q> r;
    for (int c = 0; c < N; ++c)
        for (int rr = 0; rr < N; ++rr) r[c][rr] = m[rr][c];
    return r;
}
// outerProduct(c, r) = c * r^T : column i is c * r[i]
template <length_t N, typename T, qualifier Q> GLMS_FN mat<N, N, T, Q> outerProduct(vec<N, T, Q> const &c, vec<N, T, Q> const &r) {
    mat<N, N, T, Q> m;
    for (int i = 0; i < N; ++i) m[i] = c * r[i];
    return m;
}
template <typename T, qualifier Q> GLMS_FN mat<2, 2, T, Q> inverse(mat<2, 2, T, Q> const &m) {
    T ood = static_cast<T>(1) / (m[0][0] * m[1][1] - m[1][0] * m[0][1]);
    return mat<2, 2, T, Q>(m[1][1] * ood, -m[0][1] * ood, -m[1][0] * ood, m[0][0] * ood);
}

// ------------------------------------------------------------------ quaternion (w, x, y, z ctor order)
template <typename T, qualifier Q> struct qua {
    T x, y, z, w;
    qua() = default;
    GLMS_FN qua(T w_, T x_, T y_, T z_) : x(x_), y(y_), z(z_), w(w_) {}
};
template <typename T, qualifier Q> GLMS_FN qua<T, Q> operator+(qua<T, Q> const &a, qua<T, Q> const &b) {
    return qua<T, Q>(a.w + b.w, a.x + b.x, a.y + b.y, a.z + b.z);
}
template <typename T, qualifier Q> GLMS_FN qua<T, Q> operator-(qua<T, Q> const &a) { return qua<T, Q>(-a.w, -a.x, -a.y, -a.z); }
template <typename T, qualifier Q> GLMS_FN qua<T, Q> operator*(qua<T, Q> const &a, T s) {
    return qua<T, Q>(a.w * s, a.x * s, a.y * s, a.z * s);
}
template <typename T, qualifier Q> GLMS_FN qua<T, Q> operator*(T s, qua<T, Q> const &a) { return a * s; }
template <typename T, qualifier Q> GLMS_FN qua<T, Q> operator/(qua<T, Q> const &a, T s) {
    return qua<T, Q>(a.w / s, a.x / s, a.y / s, a.z / s);
}
template <typename T, qualifier Q> GLMS_FN T dot(qua<T, Q> const &a, qua<T, Q> const &b) {
    return (a.w * b.w + a.x * b.x) + (a.y * b.y + a.z * b.z);
}
template <typename T, qualifier Q> GLMS_FN T length(qua<T, Q> const &q) { return std::sqrt(dot(q, q)); }
template <typename T, qualifier Q> GLMS_FN qua<T, Q> normalize(qua<T, Q> const &q) {
    T len = length(q);
    if (len <= static_cast<T>(0)) return qua<T, Q>(static_cast<T>(1), static_cast<T>(0), static_cast<T>(0), static_cast<T>(0));
    T ool = static_cast<T>(1) / len;
    return qua<T, Q>(q.w * ool, q.x * ool, q.y * ool, q.z * ool);
}
template <typename T, qualifier Q> GLMS_FN qua<T, Q> conjugate(qua<T, Q> const &q) { return qua<T, Q>(q.w, -q.x, -q.y, -q.z); }
template <typename T, qualifier Q> GLMS_FN qua<T, Q> inverse(qua<T, Q> const &q) { return conjugate(q) / dot(q, q); }
// rotate a vector by a quaternion
template <typename T, qualifier Q> GLMS_FN vec<3, T, Q> operator*(qua<T, Q> const &q, vec<3, T, Q> const &v) {
    vec<3, T, Q> const u(q.x, q.y, q.z);
    vec<3, T, Q> const uv(cross(u, v));
    vec<3, T, Q> const uuv(cross(u, uv));
    return v + ((uv * q.w) + uuv) * static_cast<T>(2);
}
template <typename T, qualifier Q> GLMS_FN vec<3, T, Q> rotate(qua<T, Q> const &q, vec<3, T, Q> const &v) { return q * v; }

template <typename T, qualifier Q> GLMS_FN mat<3, 3, T, Q> mat3_cast(qua<T, Q> const &q) {
    mat<3, 3, T, Q> r(static_cast<T>(1));
    T qxx(q.x * q.x), qyy(q.y * q.y), qzz(q.z * q.z), qxz(q.x * q.z), qxy(q.x * q.y), qyz(q.y * q.z);
    T qwx(q.w * q.x), qwy(q.w * q.y), qwz(q.w * q.z);
    r[0][0] = T(1) - T(2) * (qyy + qzz); r[0][1] = T(2) * (qxy + qwz); r[0][2] = T(2) * (qxz - qwy);
    r[1][0] = T(2) * (qxy - qwz); r[1][1] = T(1) - T(2) * (qxx + qzz); r[1][2] = T(2) * (qyz + qwx);
    r[2][0] = T(2) * (qxz + qwy); r[2][1] = T(2) * (qyz - qwx); r[2][2] = T(1) - T(2) * (qxx + qyy);
    return r;
}
template <typename T, qualifier Q> GLMS_FN qua<T, Q> quat_cast(mat<3, 3, T, Q> const &m) {
    T fourX = m[0][0] - m[1][1] - m[2][2];
    T fourY = m[1][1] - m[0][0] - m[2][2];
    T fourZ = m[2][2] - m[0][0] - m[1][1];
    T fourW = m[0][0] + m[1][1] + m[2][2];
    int biggest = 0;
    T fourBiggest = fourW;
    if (fourX > fourBiggest) { fourBiggest = fourX; biggest = 1; }
    if (fourY > fourBiggest) { fourBiggest = fourY; biggest = 2; }
    if (fourZ > fourBiggest) { fourBiggest = fourZ; biggest = 3; }
    T biggestVal = std::sqrt(fourBiggest + static_cast<T>(1)) * static_cast<T>(0.5);
    T mult = static_cast<T>(0.25) / biggestVal;
    switch (biggest) {
    case 0: return qua<T, Q>(biggestVal, (m[1][2] - m[2][1]) * mult, (m[2][0] - m[0][2]) * mult, (m[0][1] - m[1][0]) * mult);
    case 1: return qua<T, Q>((m[1][2] - m[2][1]) * mult, biggestVal, (m[0][1] + m[1][0]) * mult, (m[2][0] + m[0][2]) * mult);
    case 2: return qua<T, Q>((m[2][0] - m[0][2]) * mult, (m[0][1] + m[1][0]) * mult, biggestVal, (m[1][2] + m[2][1]) * mult);
    default: return qua<T, Q>((m[0][1] - m[1][0]) * mult, (m[2][0] + m[0][2]) * mult, (m[1][2] + m[2][1]) * mult, biggestVal);
    }
}
template <typename T> GLMS_FN T mix(T x, T y, T a) { return x * (static_cast<T>(1) - a) + y * a; }
template <typename T, qualifier Q> GLMS_FN qua<T, Q> slerp(qua<T, Q> const &x, qua<T, Q> const &y, T a) {
    qua<T, Q> z = y;
    T cosTheta = dot(x, y);
    if (cosTheta < static_cast<T>(0)) { z = -y; cosTheta = -cosTheta; }
    if (cosTheta > static_cast<T>(1) - std::numeric_limits<T>::epsilon()) {
        return qua<T, Q>(mix(x.w, z.w, a), mix(x.x, z.x, a), mix(x.y, z.y, a), mix(x.z, z.z, a));
    }
    T angle = std::acos(cosTheta);
    return (std::sin((static_cast<T>(1) - a) * angle) * x + std::sin(a * angle) * z) / std::sin(angle);
}

// ------------------------------------------------------------------ type_ptr.hpp helpers
template <typename T> GLMS_FN vec<2, T, defaultp> make_vec2(T const *p) { return vec<2, T, defaultp>(p[0], p[1]); }
template <typename T> GLMS_FN vec<3, T, defaultp> make_vec3(T const *p) { return vec<3, T, defaultp>(p[0], p[1], p[2]); }
template <typename T> GLMS_FN vec<4, T, defaultp> make_vec4(T const *p) { return vec<4, T, defaultp>(p[0], p[1], p[2], p[3]); }

typedef vec<2, float> vec2; typedef vec<3, float> vec3; typedef vec<4, float> vec4;
typedef vec<2, float> fvec2; typedef vec<3, float> fvec3; typedef vec<4, float> fvec4;
typedef mat<2, 2, float> mat2; typedef mat<3, 3, float> mat3; typedef mat<4, 4, float> mat4;
typedef mat<2, 2, float> fmat2; typedef mat<3, 3, float> fmat3; typedef mat<4, 4, float> fmat4;
typedef qua<float> quat; typedef qua<float> fquat;

} // namespace glm
