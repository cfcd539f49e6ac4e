/*
 * glm stand-in for building the UNMODIFIED reference hot path into oracle/_ref.
 * TEST INFRASTRUCTURE ONLY.
 *
 * The reference depends on glm 0.9.7.6 (conanfile.txt:2), a header-only Conan
 * package that is not under /root/reference and cannot be fetched offline.
 * This header restates, from glm's published sources, exactly the subset the
 * reference calls (call sites: geometry.cpp:24,30,32,40,49,101,126;
 * drawing.cpp:61,71-74,86-92,111,131; shading.cpp:21,31-33), preserving the
 * operation order of each glm function (type_mat4x4.inl, func_matrix.inl,
 * func_geometric.inl, func_common.inl, gtc/matrix_transform.inl).
 * It is written against the API, independently of oracle/oracle.c, so the two
 * restatements cross-check each other; neither is a checkout of glm, hence
 * "parity unpinned" for the glm layer.
 */
#ifndef GLM_STANDIN_HPP
#define GLM_STANDIN_HPP

#include <cmath>
#include <cstddef>

namespace glm {

template <typename T> struct tvec2 {
    union { T x, r, s; };
    union { T y, g, t; };
    tvec2() : x(0), y(0) {}
    explicit tvec2(T v) : x(v), y(v) {}
    template <typename A, typename B> tvec2(A a, B b) : x(static_cast<T>(a)), y(static_cast<T>(b)) {}
    template <typename U> explicit tvec2(const tvec2<U> &v) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)) {}
    T &operator[](std::size_t i) { return i == 0 ? x : y; }
    const T &operator[](std::size_t i) const { return i == 0 ? x : y; }
};

template <typename T> struct tvec3 {
    union { T x, r, s; };
    union { T y, g, t; };
    union { T z, b, p; };
    tvec3() : x(0), y(0), z(0) {}
    explicit tvec3(T v) : x(v), y(v), z(v) {}
    template <typename A, typename B, typename C> tvec3(A a, B b_, C c) : x(static_cast<T>(a)), y(static_cast<T>(b_)), z(static_cast<T>(c)) {}
    template <typename U> explicit tvec3(const tvec3<U> &v) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(v.z)) {}
    T &operator[](std::size_t i) { return i == 0 ? x : (i == 1 ? y : z); }
    const T &operator[](std::size_t i) const { return i == 0 ? x : (i == 1 ? y : z); }
};

template <typename T> struct tvec4 {
    union { T x, r, s; };
    union { T y, g, t; };
    union { T z, b, p; };
    union { T w, a, q; };
    tvec4() : x(0), y(0), z(0), w(0) {}
    explicit tvec4(T v) : x(v), y(v), z(v), w(v) {}
    template <typename A, typename B, typename C, typename D> tvec4(A a_, B b_, C c, D d) : x(static_cast<T>(a_)), y(static_cast<T>(b_)), z(static_cast<T>(c)), w(static_cast<T>(d)) {}
    T &operator[](std::size_t i) { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    const T &operator[](std::size_t i) const { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
};

typedef tvec2<float> vec2;
typedef tvec3<float> vec3;
typedef tvec4<float> vec4;
typedef tvec2<unsigned int> uvec2;
typedef tvec3<unsigned int> uvec3;

/* componentwise arithmetic */
template <typename T> inline tvec2<T> operator+(const tvec2<T> &a, const tvec2<T> &b) { return tvec2<T>(a.x + b.x, a.y + b.y); }
template <typename T> inline tvec2<T> operator-(const tvec2<T> &a, const tvec2<T> &b) { return tvec2<T>(a.x - b.x, a.y - b.y); }
template <typename T> inline tvec2<T> operator*(const tvec2<T> &a, const tvec2<T> &b) { return tvec2<T>(a.x * b.x, a.y * b.y); }
template <typename T> inline tvec2<T> operator*(const tvec2<T> &a, T s) { return tvec2<T>(a.x * s, a.y * s); }
template <typename T> inline tvec2<T> operator*(T s, const tvec2<T> &a) { return tvec2<T>(s * a.x, s * a.y); }

template <typename T> inline tvec3<T> operator+(const tvec3<T> &a, const tvec3<T> &b) { return tvec3<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename T> inline tvec3<T> operator-(const tvec3<T> &a, const tvec3<T> &b) { return tvec3<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename T> inline tvec3<T> operator-(const tvec3<T> &a) { return tvec3<T>(-a.x, -a.y, -a.z); }
template <typename T> inline tvec3<T> operator*(const tvec3<T> &a, const tvec3<T> &b) { return tvec3<T>(a.x * b.x, a.y * b.y, a.z * b.z); }
template <typename T> inline tvec3<T> operator*(const tvec3<T> &a, T s) { return tvec3<T>(a.x * s, a.y * s, a.z * s); }
template <typename T> inline tvec3<T> operator*(T s, const tvec3<T> &a) { return tvec3<T>(s * a.x, s * a.y, s * a.z); }

template <typename T> inline tvec4<T> operator+(const tvec4<T> &a, const tvec4<T> &b) { return tvec4<T>(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
template <typename T> inline tvec4<T> operator-(const tvec4<T> &a, const tvec4<T> &b) { return tvec4<T>(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
template <typename T> inline tvec4<T> operator*(const tvec4<T> &a, const tvec4<T> &b) { return tvec4<T>(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
template <typename T> inline tvec4<T> operator*(const tvec4<T> &a, T s) { return tvec4<T>(a.x * s, a.y * s, a.z * s, a.w * s); }
template <typename T> inline tvec4<T> operator*(T s, const tvec4<T> &a) { return tvec4<T>(s * a.x, s * a.y, s * a.z, s * a.w); }

/* func_common.inl */
template <typename T> inline T min(T x, T y) { return x < y ? x : y; }
template <typename T> inline T max(T x, T y) { return x > y ? x : y; }
template <typename T> inline tvec2<T> min(const tvec2<T> &a, const tvec2<T> &b) { return tvec2<T>(min(a.x, b.x), min(a.y, b.y)); }
template <typename T> inline tvec2<T> max(const tvec2<T> &a, const tvec2<T> &b) { return tvec2<T>(max(a.x, b.x), max(a.y, b.y)); }
template <typename T> inline tvec3<T> min(const tvec3<T> &a, const tvec3<T> &b) { return tvec3<T>(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
template <typename T> inline tvec3<T> max(const tvec3<T> &a, const tvec3<T> &b) { return tvec3<T>(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
/* gtx/extented_min_max.inl: 3-argument forms */
template <typename V> inline V min(const V &x, const V &y, const V &z) { return glm::min(glm::min(x, y), z); }
template <typename V> inline V max(const V &x, const V &y, const V &z) { return glm::max(glm::max(x, y), z); }
template <typename T> inline tvec2<T> clamp(const tvec2<T> &x, const tvec2<T> &lo, const tvec2<T> &hi) { return min(max(x, lo), hi); }
inline vec2 ceil(const vec2 &v) { return vec2(std::ceil(v.x), std::ceil(v.y)); }

/* func_vector_relational.inl */
struct bvec3 { bool x, y, z; };
inline bvec3 greaterThanEqual(const vec3 &a, const vec3 &b) { bvec3 r = {a.x >= b.x, a.y >= b.y, a.z >= b.z}; return r; }
inline bool all(const bvec3 &v) { return v.x && v.y && v.z; }

/* func_geometric.inl */
inline float dot(const vec3 &a, const vec3 &b) { vec3 tmp(a * b); return tmp.x + tmp.y + tmp.z; }
inline float inversesqrt(float x) { return 1.f / std::sqrt(x); }
inline vec3 normalize(const vec3 &v) { return v * inversesqrt(dot(v, v)); }

/* func_trigonometric.inl, gtc/constants.inl */
inline float radians(float degrees) { return degrees * 0.01745329251994329576923690768489f; }
template <typename T> inline T one_over_pi() { return T(0.318309886183790671537767526745028724); }

/* type_mat4x4 (column-major: m[c][r]) */
struct mat4 {
    vec4 value[4];
    mat4() { value[0] = vec4(1.f, 0.f, 0.f, 0.f); value[1] = vec4(0.f, 1.f, 0.f, 0.f); value[2] = vec4(0.f, 0.f, 1.f, 0.f); value[3] = vec4(0.f, 0.f, 0.f, 1.f); }
    explicit mat4(float s) { value[0] = vec4(s, 0.f, 0.f, 0.f); value[1] = vec4(0.f, s, 0.f, 0.f); value[2] = vec4(0.f, 0.f, s, 0.f); value[3] = vec4(0.f, 0.f, 0.f, s); }
    mat4(const vec4 &a, const vec4 &b, const vec4 &c, const vec4 &d) { value[0] = a; value[1] = b; value[2] = c; value[3] = d; }
    vec4 &operator[](std::size_t i) { return value[i]; }
    const vec4 &operator[](std::size_t i) const { return value[i]; }
};

inline vec4 operator*(const mat4 &m, const vec4 &v) {
    vec4 const Mov0(v[0]);
    vec4 const Mov1(v[1]);
    vec4 const Mul0 = m[0] * Mov0;
    vec4 const Mul1 = m[1] * Mov1;
    vec4 const Add0 = Mul0 + Mul1;
    vec4 const Mov2(v[2]);
    vec4 const Mov3(v[3]);
    vec4 const Mul2 = m[2] * Mov2;
    vec4 const Mul3 = m[3] * Mov3;
    vec4 const Add1 = Mul2 + Mul3;
    vec4 const Add2 = Add0 + Add1;
    return Add2;
}

inline mat4 operator*(const mat4 &m1, const mat4 &m2) {
    vec4 const SrcA0 = m1[0], SrcA1 = m1[1], SrcA2 = m1[2], SrcA3 = m1[3];
    vec4 const SrcB0 = m2[0], SrcB1 = m2[1], SrcB2 = m2[2], SrcB3 = m2[3];
    mat4 Result(0.f);
    Result[0] = SrcA0 * SrcB0[0] + SrcA1 * SrcB0[1] + SrcA2 * SrcB0[2] + SrcA3 * SrcB0[3];
    Result[1] = SrcA0 * SrcB1[0] + SrcA1 * SrcB1[1] + SrcA2 * SrcB1[2] + SrcA3 * SrcB1[3];
    Result[2] = SrcA0 * SrcB2[0] + SrcA1 * SrcB2[1] + SrcA2 * SrcB2[2] + SrcA3 * SrcB2[3];
    Result[3] = SrcA0 * SrcB3[0] + SrcA1 * SrcB3[1] + SrcA2 * SrcB3[2] + SrcA3 * SrcB3[3];
    return Result;
}

inline mat4 operator*(const mat4 &m, float s) { return mat4(m[0] * s, m[1] * s, m[2] * s, m[3] * s); }

/* gtc/matrix_transform.inl */
inline mat4 translate(const mat4 &m, const vec3 &v) {
    mat4 Result(m);
    Result[3] = m[0] * v[0] + m[1] * v[1] + m[2] * v[2] + m[3];
    return Result;
}

inline mat4 rotate(const mat4 &m, float angle, const vec3 &v) {
    float const a = angle;
    float const c = std::cos(a);
    float const s = std::sin(a);
    vec3 axis(normalize(v));
    vec3 temp((1.f - c) * axis);
    mat4 Rotate(0.f);
    Rotate[0][0] = c + temp[0] * axis[0];
    Rotate[0][1] = 0 + temp[0] * axis[1] + s * axis[2];
    Rotate[0][2] = 0 + temp[0] * axis[2] - s * axis[1];
    Rotate[1][0] = 0 + temp[1] * axis[0] - s * axis[2];
    Rotate[1][1] = c + temp[1] * axis[1];
    Rotate[1][2] = 0 + temp[1] * axis[2] + s * axis[0];
    Rotate[2][0] = 0 + temp[2] * axis[0] + s * axis[1];
    Rotate[2][1] = 0 + temp[2] * axis[1] - s * axis[0];
    Rotate[2][2] = c + temp[2] * axis[2];
    mat4 Result(0.f);
    Result[0] = m[0] * Rotate[0][0] + m[1] * Rotate[0][1] + m[2] * Rotate[0][2];
    Result[1] = m[0] * Rotate[1][0] + m[1] * Rotate[1][1] + m[2] * Rotate[1][2];
    Result[2] = m[0] * Rotate[2][0] + m[1] * Rotate[2][1] + m[2] * Rotate[2][2];
    Result[3] = m[3];
    return Result;
}

inline mat4 scale(const mat4 &m, const vec3 &v) {
    mat4 Result(0.f);
    Result[0] = m[0] * v[0];
    Result[1] = m[1] * v[1];
    Result[2] = m[2] * v[2];
    Result[3] = m[3];
    return Result;
}

inline mat4 perspective(float fovy, float aspect, float zNear, float zFar) {
    float const tanHalfFovy = std::tan(fovy / 2.f);
    mat4 Result(0.f);
    Result[0][0] = 1.f / (aspect * tanHalfFovy);
    Result[1][1] = 1.f / (tanHalfFovy);
    Result[2][2] = -(zFar + zNear) / (zFar - zNear);
    Result[2][3] = -1.f;
    Result[3][2] = -(2.f * zFar * zNear) / (zFar - zNear);
    return Result;
}

/* func_matrix.inl */
inline mat4 transpose(const mat4 &m) {
    mat4 r(0.f);
    for (int c = 0; c < 4; ++c)
        for (int rr = 0; rr < 4; ++rr) r[c][rr] = m[rr][c];
    return r;
}

inline mat4 inverse(const mat4 &m) {
    float Coef00 = m[2][2] * m[3][3] - m[3][2] * m[2][3];
    float Coef02 = m[1][2] * m[3][3] - m[3][2] * m[1][3];
    float Coef03 = m[1][2] * m[2][3] - m[2][2] * m[1][3];

    float Coef04 = m[2][1] * m[3][3] - m[3][1] * m[2][3];
    float Coef06 = m[1][1] * m[3][3] - m[3][1] * m[1][3];
    float Coef07 = m[1][1] * m[2][3] - m[2][1] * m[1][3];

    float Coef08 = m[2][1] * m[3][2] - m[3][1] * m[2][2];
    float Coef10 = m[1][1] * m[3][2] - m[3][1] * m[1][2];
    float Coef11 = m[1][1] * m[2][2] - m[2][1] * m[1][2];

    float Coef12 = m[2][0] * m[3][3] - m[3][0] * m[2][3];
    float Coef14 = m[1][0] * m[3][3] - m[3][0] * m[1][3];
    float Coef15 = m[1][0] * m[2][3] - m[2][0] * m[1][3];

    float Coef16 = m[2][0] * m[3][2] - m[3][0] * m[2][2];
    float Coef18 = m[1][0] * m[3][2] - m[3][0] * m[1][2];
    float Coef19 = m[1][0] * m[2][2] - m[2][0] * m[1][2];

    float Coef20 = m[2][0] * m[3][1] - m[3][0] * m[2][1];
    float Coef22 = m[1][0] * m[3][1] - m[3][0] * m[1][1];
    float Coef23 = m[1][0] * m[2][1] - m[2][0] * m[1][1];

    vec4 Fac0(Coef00, Coef00, Coef02, Coef03);
    vec4 Fac1(Coef04, Coef04, Coef06, Coef07);
    vec4 Fac2(Coef08, Coef08, Coef10, Coef11);
    vec4 Fac3(Coef12, Coef12, Coef14, Coef15);
    vec4 Fac4(Coef16, Coef16, Coef18, Coef19);
    vec4 Fac5(Coef20, Coef20, Coef22, Coef23);

    vec4 Vec0(m[1][0], m[0][0], m[0][0], m[0][0]);
    vec4 Vec1(m[1][1], m[0][1], m[0][1], m[0][1]);
    vec4 Vec2(m[1][2], m[0][2], m[0][2], m[0][2]);
    vec4 Vec3(m[1][3], m[0][3], m[0][3], m[0][3]);

    vec4 Inv0(Vec1 * Fac0 - Vec2 * Fac1 + Vec3 * Fac2);
    vec4 Inv1(Vec0 * Fac0 - Vec2 * Fac3 + Vec3 * Fac4);
    vec4 Inv2(Vec0 * Fac1 - Vec1 * Fac3 + Vec3 * Fac5);
    vec4 Inv3(Vec0 * Fac2 - Vec1 * Fac4 + Vec2 * Fac5);

    vec4 SignA(+1.f, -1.f, +1.f, -1.f);
    vec4 SignB(-1.f, +1.f, -1.f, +1.f);
    mat4 Inverse(Inv0 * SignA, Inv1 * SignB, Inv2 * SignA, Inv3 * SignB);

    vec4 Row0(Inverse[0][0], Inverse[1][0], Inverse[2][0], Inverse[3][0]);

    vec4 Dot0(m[0] * Row0);
    float Dot1 = (Dot0.x + Dot0.y) + (Dot0.z + Dot0.w);

    float OneOverDeterminant = 1.f / Dot1;

    return Inverse * OneOverDeterminant;
}

} // namespace glm

#endif
