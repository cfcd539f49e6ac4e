// hostmath.h -- host-side matrix builders of the frame path, in glm 0.9.7.6's operation order.
//
// The reference builds two matrices per frame on the CPU with glm (drawing.cpp:222-229,
// geometry.cpp:22-33) plus transpose(inverse(modelview)) for normals (geometry.cpp:101) and the
// normalised light directions (geometry.cpp:124-133).  These stay on the host here as well (they
// are O(1) per frame and use libm's sinf/cosf/tanf); only their results go to the device.  Every
// expression keeps the association glm uses so the fp32 bits equal the reference's.  Compile with
// -ffp-contract=off (nvcc host pass: -Xcompiler -ffp-contract=off).
#pragma once

#include <cmath>

namespace hostmath {

struct Vec4 {
    float x, y, z, w;
};

inline Vec4 operator+(Vec4 a, Vec4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline Vec4 operator-(Vec4 a, Vec4 b) { return {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
inline Vec4 operator*(Vec4 a, Vec4 b) { return {a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w}; }
inline Vec4 operator*(Vec4 a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }

// column-major like glm: c[i] is column i
struct Mat4 {
    Vec4 c[4];
    static Mat4 identity() { return {{{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}}}; }
    static Mat4 zero() { return {{{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}}}; }
    float at(int col, int row) const { return (&c[col].x)[row]; }
    float &at(int col, int row) { return (&c[col].x)[row]; }
    void store(float out[16]) const {
        for (int i = 0; i < 4; ++i) { out[4 * i] = c[i].x; out[4 * i + 1] = c[i].y; out[4 * i + 2] = c[i].z; out[4 * i + 3] = c[i].w; }
    }
    static Mat4 load(const float in[16]) {
        Mat4 m;
        for (int i = 0; i < 4; ++i) m.c[i] = {in[4 * i], in[4 * i + 1], in[4 * i + 2], in[4 * i + 3]};
        return m;
    }
};

// glm operator*(mat4, vec4): (m0*x + m1*y) + (m2*z + m3*w)
inline Vec4 mul(const Mat4 &m, Vec4 v) { return (m.c[0] * v.x + m.c[1] * v.y) + (m.c[2] * v.z + m.c[3] * v.w); }

// glm operator*(mat4, mat4): column i = ((a0*b[i].x + a1*b[i].y) + a2*b[i].z) + a3*b[i].w
inline Mat4 mul(const Mat4 &a, const Mat4 &b) {
    Mat4 r;
    for (int i = 0; i < 4; ++i) r.c[i] = ((a.c[0] * b.c[i].x + a.c[1] * b.c[i].y) + a.c[2] * b.c[i].z) + a.c[3] * b.c[i].w;
    return r;
}

// glm::translate(mat4(1), d)
inline Mat4 translation(float dx, float dy, float dz) {
    Mat4 m = Mat4::identity(), r = m;
    r.c[3] = ((m.c[0] * dx + m.c[1] * dy) + m.c[2] * dz) + m.c[3];
    return r;
}

// glm::scale(mat4(1), (f,f,f))
inline Mat4 scaling(float f) {
    Mat4 m = Mat4::identity(), r;
    r.c[0] = m.c[0] * f;
    r.c[1] = m.c[1] * f;
    r.c[2] = m.c[2] * f;
    r.c[3] = m.c[3];
    return r;
}

// glm::rotate(mat4(1), angle, axis) -- gtc/matrix_transform.inl, including the literal "0 +" terms
inline Mat4 rotation(float angle, float ax, float ay, float az) {
    const float c = std::cos(angle), s = std::sin(angle);
    const float inv_len = 1.f / std::sqrt((ax * ax + ay * ay) + az * az); // normalize = v * inversesqrt(dot)
    const float a[3] = {ax * inv_len, ay * inv_len, az * inv_len};
    const float t[3] = {(1.f - c) * a[0], (1.f - c) * a[1], (1.f - c) * a[2]};
    float rot[3][3];
    rot[0][0] = c + t[0] * a[0];
    rot[0][1] = 0 + t[0] * a[1] + s * a[2];
    rot[0][2] = 0 + t[0] * a[2] - s * a[1];
    rot[1][0] = 0 + t[1] * a[0] - s * a[2];
    rot[1][1] = c + t[1] * a[1];
    rot[1][2] = 0 + t[1] * a[2] + s * a[0];
    rot[2][0] = 0 + t[2] * a[0] + s * a[1];
    rot[2][1] = 0 + t[2] * a[1] - s * a[0];
    rot[2][2] = c + t[2] * a[2];
    const Mat4 m = Mat4::identity();
    Mat4 r;
    for (int j = 0; j < 3; ++j) r.c[j] = (m.c[0] * rot[j][0] + m.c[1] * rot[j][1]) + m.c[2] * rot[j][2];
    r.c[3] = m.c[3];
    return r;
}

// transformation_matrix (geometry.cpp:22-25): T * Ry * Rx * Rz * S, multiplied left to right
inline Mat4 transformation_matrix(float factor, const float disp[3], const float tait_bryan[3]) {
    Mat4 m = translation(disp[0], disp[1], disp[2]);
    m = mul(m, rotation(tait_bryan[1], 0.f, 1.f, 0.f));
    m = mul(m, rotation(tait_bryan[0], 1.f, 0.f, 0.f));
    m = mul(m, rotation(tait_bryan[2], 0.f, 0.f, 1.f));
    return mul(m, scaling(factor));
}

// glm::perspective(fovy, aspect, near, far), right-handed, depth -1..1
inline Mat4 perspective(float fovy, float aspect, float z_near, float z_far) {
    const float tan_half = std::tan(fovy / 2.f);
    Mat4 r = Mat4::zero();
    r.at(0, 0) = 1.f / (aspect * tan_half);
    r.at(1, 1) = 1.f / (tan_half);
    r.at(2, 2) = -(z_far + z_near) / (z_far - z_near);
    r.at(2, 3) = -1.f;
    r.at(3, 2) = -(2.f * z_far * z_near) / (z_far - z_near);
    return r;
}

// camera_matrix (geometry.cpp:27-33): perspective(radians(45), aspect, 0.1, 6) * modelview
inline Mat4 camera_matrix(const Mat4 &modelview, float aspect) {
    const float fovy = 45.0f * 0.01745329251994329576923690768489f; // glm::radians
    return mul(perspective(fovy, aspect, 0.1f, 6.f), modelview);
}

// glm::inverse(mat4): cofactor scheme of func_matrix.inl (compute_inverse<tmat4x4>)
inline Mat4 inverse(const Mat4 &m) {
    auto sub2 = [&](int c0, int r0, int c1, int r1, int c2, int r2, int c3, int r3) { return m.at(c0, r0) * m.at(c1, r1) - m.at(c2, r2) * m.at(c3, r3); };
    const float k00 = sub2(2, 2, 3, 3, 3, 2, 2, 3), k02 = sub2(1, 2, 3, 3, 3, 2, 1, 3), k03 = sub2(1, 2, 2, 3, 2, 2, 1, 3);
    const float k04 = sub2(2, 1, 3, 3, 3, 1, 2, 3), k06 = sub2(1, 1, 3, 3, 3, 1, 1, 3), k07 = sub2(1, 1, 2, 3, 2, 1, 1, 3);
    const float k08 = sub2(2, 1, 3, 2, 3, 1, 2, 2), k10 = sub2(1, 1, 3, 2, 3, 1, 1, 2), k11 = sub2(1, 1, 2, 2, 2, 1, 1, 2);
    const float k12 = sub2(2, 0, 3, 3, 3, 0, 2, 3), k14 = sub2(1, 0, 3, 3, 3, 0, 1, 3), k15 = sub2(1, 0, 2, 3, 2, 0, 1, 3);
    const float k16 = sub2(2, 0, 3, 2, 3, 0, 2, 2), k18 = sub2(1, 0, 3, 2, 3, 0, 1, 2), k19 = sub2(1, 0, 2, 2, 2, 0, 1, 2);
    const float k20 = sub2(2, 0, 3, 1, 3, 0, 2, 1), k22 = sub2(1, 0, 3, 1, 3, 0, 1, 1), k23 = sub2(1, 0, 2, 1, 2, 0, 1, 1);

    const Vec4 f0{k00, k00, k02, k03}, f1{k04, k04, k06, k07}, f2{k08, k08, k10, k11};
    const Vec4 f3{k12, k12, k14, k15}, f4{k16, k16, k18, k19}, f5{k20, k20, k22, k23};
    const Vec4 v0{m.at(1, 0), m.at(0, 0), m.at(0, 0), m.at(0, 0)}, v1{m.at(1, 1), m.at(0, 1), m.at(0, 1), m.at(0, 1)};
    const Vec4 v2{m.at(1, 2), m.at(0, 2), m.at(0, 2), m.at(0, 2)}, v3{m.at(1, 3), m.at(0, 3), m.at(0, 3), m.at(0, 3)};

    const Vec4 i0 = (v1 * f0 - v2 * f1) + v3 * f2;
    const Vec4 i1 = (v0 * f0 - v2 * f3) + v3 * f4;
    const Vec4 i2 = (v0 * f1 - v1 * f3) + v3 * f5;
    const Vec4 i3 = (v0 * f2 - v1 * f4) + v2 * f5;

    const Vec4 sign_a{+1.f, -1.f, +1.f, -1.f}, sign_b{-1.f, +1.f, -1.f, +1.f};
    Mat4 inv{{i0 * sign_a, i1 * sign_b, i2 * sign_a, i3 * sign_b}};

    const Vec4 row0{inv.c[0].x, inv.c[1].x, inv.c[2].x, inv.c[3].x};
    const Vec4 d = m.c[0] * row0;
    const float det = (d.x + d.y) + (d.z + d.w);
    const float one_over_det = 1.f / det;
    for (int i = 0; i < 4; ++i) inv.c[i] = inv.c[i] * one_over_det;
    return inv;
}

inline Mat4 transpose(const Mat4 &m) {
    Mat4 r;
    for (int col = 0; col < 4; ++col)
        for (int row = 0; row < 4; ++row) r.at(col, row) = m.at(row, col);
    return r;
}

// Light::transform (geometry.cpp:124-127): normalize(xyz(view * (dir, 0)))
inline void light_direction(const Mat4 &view, const float dir[3], float out[3]) {
    const Vec4 t = mul(view, Vec4{dir[0], dir[1], dir[2], 0.f});
    const float inv_len = 1.f / std::sqrt((t.x * t.x + t.y * t.y) + t.z * t.z);
    out[0] = t.x * inv_len;
    out[1] = t.y * inv_len;
    out[2] = t.z * inv_len;
}

} // namespace hostmath
