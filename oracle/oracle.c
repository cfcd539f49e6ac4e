/*
 * oracle.c -- CPU restatement of the canmom/rasteriser frame path.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Each function cites the
 * reference file:line it follows; glm 0.9.7.6 (not in /root/reference) is
 * restated from its published algorithm, operation order preserved.
 *
 * Build: gcc -std=c11 -O2 -ffp-contract=off -fPIC -shared oracle.c -lm -lpthread
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

/* ------------------------------------------------------------------ */
/* glm 0.9.7.6 restated (column-major: m[c*4 + r] == glm m[c][r])      */
/* ------------------------------------------------------------------ */

typedef struct { float x, y, z, w; } v4;

static inline v4 v4_make(float x, float y, float z, float w) { v4 r = {x, y, z, w}; return r; }
static inline v4 v4_add(v4 a, v4 b) { return v4_make(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
static inline v4 v4_sub(v4 a, v4 b) { return v4_make(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
static inline v4 v4_mul(v4 a, v4 b) { return v4_make(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
static inline v4 v4_scale(v4 a, float s) { return v4_make(a.x * s, a.y * s, a.z * s, a.w * s); }
static inline v4 col(const float *m, int c) { return v4_make(m[c * 4], m[c * 4 + 1], m[c * 4 + 2], m[c * 4 + 3]); }
static inline void set_col(float *m, int c, v4 v) { m[c * 4] = v.x; m[c * 4 + 1] = v.y; m[c * 4 + 2] = v.z; m[c * 4 + 3] = v.w; }

/* glm func_common.inl: min(x,y) = x < y ? x : y ; max(x,y) = x > y ? x : y */
static inline float glm_min(float x, float y) { return x < y ? x : y; }
static inline float glm_max(float x, float y) { return x > y ? x : y; }

static void mat_identity(float *m) { memset(m, 0, 16 * sizeof(float)); m[0] = m[5] = m[10] = m[15] = 1.f; }

/* glm type_mat4x4.inl operator*(mat4,mat4): R[i] = ((a0*b[i][0] + a1*b[i][1]) + a2*b[i][2]) + a3*b[i][3] */
static void mat_mul(const float *a, const float *b, float *out) {
    float r[16];
    for (int i = 0; i < 4; ++i) {
        v4 acc = v4_scale(col(a, 0), b[i * 4 + 0]);
        acc = v4_add(acc, v4_scale(col(a, 1), b[i * 4 + 1]));
        acc = v4_add(acc, v4_scale(col(a, 2), b[i * 4 + 2]));
        acc = v4_add(acc, v4_scale(col(a, 3), b[i * 4 + 3]));
        set_col(r, i, acc);
    }
    memcpy(out, r, sizeof r);
}

/* glm operator*(mat4,vec4): (m0*x + m1*y) + (m2*z + m3*w) */
static inline v4 mat_vec(const float *m, v4 v) {
    v4 add0 = v4_add(v4_scale(col(m, 0), v.x), v4_scale(col(m, 1), v.y));
    v4 add1 = v4_add(v4_scale(col(m, 2), v.z), v4_scale(col(m, 3), v.w));
    return v4_add(add0, add1);
}

/* glm::translate(m, v): R = m; R[3] = m0*v.x + m1*v.y + m2*v.z + m3 */
static void mat_translate(const float *m, const float v[3], float *out) {
    float r[16];
    memcpy(r, m, sizeof r);
    v4 t = v4_scale(col(m, 0), v[0]);
    t = v4_add(t, v4_scale(col(m, 1), v[1]));
    t = v4_add(t, v4_scale(col(m, 2), v[2]));
    t = v4_add(t, col(m, 3));
    set_col(r, 3, t);
    memcpy(out, r, sizeof r);
}

/* glm::scale(m, v): R[k] = m[k]*v[k], R[3] = m[3] */
static void mat_scale(const float *m, const float v[3], float *out) {
    float r[16];
    set_col(r, 0, v4_scale(col(m, 0), v[0]));
    set_col(r, 1, v4_scale(col(m, 1), v[1]));
    set_col(r, 2, v4_scale(col(m, 2), v[2]));
    set_col(r, 3, col(m, 3));
    memcpy(out, r, sizeof r);
}

/* glm::rotate(m, angle, axis) (gtc/matrix_transform.inl) */
static void mat_rotate(const float *m, float angle, const float axis_in[3], float *out) {
    const float a = angle;
    const float c = cosf(a);
    const float s = sinf(a);
    /* normalize(v) = v * inversesqrt(dot(v,v)); dot = (x*x + y*y) + z*z; inversesqrt = 1/sqrt */
    float d = (axis_in[0] * axis_in[0] + axis_in[1] * axis_in[1]) + axis_in[2] * axis_in[2];
    float inv = 1.f / sqrtf(d);
    float axis[3] = {axis_in[0] * inv, axis_in[1] * inv, axis_in[2] * inv};
    float temp[3] = {(1.f - c) * axis[0], (1.f - c) * axis[1], (1.f - c) * axis[2]};
    float R[3][3];
    R[0][0] = c + temp[0] * axis[0];
    R[0][1] = 0 + temp[0] * axis[1] + s * axis[2];
    R[0][2] = 0 + temp[0] * axis[2] - s * axis[1];
    R[1][0] = 0 + temp[1] * axis[0] - s * axis[2];
    R[1][1] = c + temp[1] * axis[1];
    R[1][2] = 0 + temp[1] * axis[2] + s * axis[0];
    R[2][0] = 0 + temp[2] * axis[0] + s * axis[1];
    R[2][1] = 0 + temp[2] * axis[1] - s * axis[0];
    R[2][2] = c + temp[2] * axis[2];
    float r[16];
    for (int j = 0; j < 3; ++j) {
        v4 t = v4_scale(col(m, 0), R[j][0]);
        t = v4_add(t, v4_scale(col(m, 1), R[j][1]));
        t = v4_add(t, v4_scale(col(m, 2), R[j][2]));
        set_col(r, j, t);
    }
    set_col(r, 3, col(m, 3));
    memcpy(out, r, sizeof r);
}

/* glm::perspective (RH, depth -1..1) */
static void mat_perspective(float fovy, float aspect, float z_near, float z_far, float *out) {
    const float tan_half_fovy = tanf(fovy / 2.f);
    memset(out, 0, 16 * sizeof(float));
    out[0 * 4 + 0] = 1.f / (aspect * tan_half_fovy);
    out[1 * 4 + 1] = 1.f / (tan_half_fovy);
    out[2 * 4 + 2] = -(z_far + z_near) / (z_far - z_near);
    out[2 * 4 + 3] = -1.f;
    out[3 * 4 + 2] = -(2.f * z_far * z_near) / (z_far - z_near);
}

/* glm::inverse(mat4) (func_matrix.inl compute_inverse<tmat4x4>) */
static void mat_inverse(const float *m_, float *out) {
#define M(c, r) m_[(c) * 4 + (r)]
    float Coef00 = M(2, 2) * M(3, 3) - M(3, 2) * M(2, 3);
    float Coef02 = M(1, 2) * M(3, 3) - M(3, 2) * M(1, 3);
    float Coef03 = M(1, 2) * M(2, 3) - M(2, 2) * M(1, 3);
    float Coef04 = M(2, 1) * M(3, 3) - M(3, 1) * M(2, 3);
    float Coef06 = M(1, 1) * M(3, 3) - M(3, 1) * M(1, 3);
    float Coef07 = M(1, 1) * M(2, 3) - M(2, 1) * M(1, 3);
    float Coef08 = M(2, 1) * M(3, 2) - M(3, 1) * M(2, 2);
    float Coef10 = M(1, 1) * M(3, 2) - M(3, 1) * M(1, 2);
    float Coef11 = M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2);
    float Coef12 = M(2, 0) * M(3, 3) - M(3, 0) * M(2, 3);
    float Coef14 = M(1, 0) * M(3, 3) - M(3, 0) * M(1, 3);
    float Coef15 = M(1, 0) * M(2, 3) - M(2, 0) * M(1, 3);
    float Coef16 = M(2, 0) * M(3, 2) - M(3, 0) * M(2, 2);
    float Coef18 = M(1, 0) * M(3, 2) - M(3, 0) * M(1, 2);
    float Coef19 = M(1, 0) * M(2, 2) - M(2, 0) * M(1, 2);
    float Coef20 = M(2, 0) * M(3, 1) - M(3, 0) * M(2, 1);
    float Coef22 = M(1, 0) * M(3, 1) - M(3, 0) * M(1, 1);
    float Coef23 = M(1, 0) * M(2, 1) - M(2, 0) * M(1, 1);

    v4 Fac0 = v4_make(Coef00, Coef00, Coef02, Coef03);
    v4 Fac1 = v4_make(Coef04, Coef04, Coef06, Coef07);
    v4 Fac2 = v4_make(Coef08, Coef08, Coef10, Coef11);
    v4 Fac3 = v4_make(Coef12, Coef12, Coef14, Coef15);
    v4 Fac4 = v4_make(Coef16, Coef16, Coef18, Coef19);
    v4 Fac5 = v4_make(Coef20, Coef20, Coef22, Coef23);

    v4 Vec0 = v4_make(M(1, 0), M(0, 0), M(0, 0), M(0, 0));
    v4 Vec1 = v4_make(M(1, 1), M(0, 1), M(0, 1), M(0, 1));
    v4 Vec2 = v4_make(M(1, 2), M(0, 2), M(0, 2), M(0, 2));
    v4 Vec3 = v4_make(M(1, 3), M(0, 3), M(0, 3), M(0, 3));

    v4 Inv0 = v4_add(v4_sub(v4_mul(Vec1, Fac0), v4_mul(Vec2, Fac1)), v4_mul(Vec3, Fac2));
    v4 Inv1 = v4_add(v4_sub(v4_mul(Vec0, Fac0), v4_mul(Vec2, Fac3)), v4_mul(Vec3, Fac4));
    v4 Inv2 = v4_add(v4_sub(v4_mul(Vec0, Fac1), v4_mul(Vec1, Fac3)), v4_mul(Vec3, Fac5));
    v4 Inv3 = v4_add(v4_sub(v4_mul(Vec0, Fac2), v4_mul(Vec1, Fac4)), v4_mul(Vec2, Fac5));

    v4 SignA = v4_make(+1.f, -1.f, +1.f, -1.f);
    v4 SignB = v4_make(-1.f, +1.f, -1.f, +1.f);
    float inv[16];
    set_col(inv, 0, v4_mul(Inv0, SignA));
    set_col(inv, 1, v4_mul(Inv1, SignB));
    set_col(inv, 2, v4_mul(Inv2, SignA));
    set_col(inv, 3, v4_mul(Inv3, SignB));

    v4 Row0 = v4_make(inv[0 * 4 + 0], inv[1 * 4 + 0], inv[2 * 4 + 0], inv[3 * 4 + 0]);
    v4 Dot0 = v4_mul(col(m_, 0), Row0);
    float Dot1 = (Dot0.x + Dot0.y) + (Dot0.z + Dot0.w);
    float OneOverDeterminant = 1.f / Dot1;
    for (int i = 0; i < 16; ++i) out[i] = inv[i] * OneOverDeterminant;
#undef M
}

static void mat_transpose(const float *m, float *out) {
    float r[16];
    for (int c = 0; c < 4; ++c)
        for (int rr = 0; rr < 4; ++rr) r[c * 4 + rr] = m[rr * 4 + c];
    memcpy(out, r, sizeof r);
}

/* ------------------------------------------------------------------ */
/* geometry.cpp                                                        */
/* ------------------------------------------------------------------ */

/* geometry.cpp:22-25: translate(I,d) * rotate(I,ry,Y) * rotate(I,rx,X) * rotate(I,rz,Z) * scale(I,f), left to right */
void orc_transformation_matrix(float factor, const float disp[3], const float tb[3], float out[16]) {
    float I[16], T[16], Ry[16], Rx[16], Rz[16], S[16], acc[16];
    const float ax_x[3] = {1.f, 0.f, 0.f}, ax_y[3] = {0.f, 1.f, 0.f}, ax_z[3] = {0.f, 0.f, 1.f};
    const float f3[3] = {factor, factor, factor};
    mat_identity(I);
    mat_translate(I, disp, T);
    mat_rotate(I, tb[1], ax_y, Ry);
    mat_rotate(I, tb[0], ax_x, Rx);
    mat_rotate(I, tb[2], ax_z, Rz);
    mat_scale(I, f3, S);
    mat_mul(T, Ry, acc);
    mat_mul(acc, Rx, acc);
    mat_mul(acc, Rz, acc);
    mat_mul(acc, S, out);
}

/* geometry.cpp:27-33 */
void orc_camera_matrix(const float modelview[16], float aspect_ratio, float out[16]) {
    float persp[16];
    /* glm::radians(45.0f) = 45 * 0.01745329251994329576923690768489f */
    float fovy = 45.0f * 0.01745329251994329576923690768489f;
    mat_perspective(fovy, aspect_ratio, 0.1f, 6.f, persp);
    mat_mul(persp, modelview, out);
}

/* drawing.cpp:222-229 and geometry.cpp:101 */
void orc_frame_matrices(const orc_args *args, float modelview[16], float camera[16], float normal_matrix[16], float view[16]) {
    float model[16], inv[16];
    const float view_disp[3] = {0.f, 0.f, -3.f}, zero3[3] = {0.f, 0.f, 0.f};
    orc_transformation_matrix(args->scale, args->displacement, args->tait_bryan_angles, model);
    orc_transformation_matrix(1.f, view_disp, zero3, view);
    mat_mul(view, model, modelview);
    orc_camera_matrix(modelview, args->aspect_ratio, camera);
    mat_inverse(modelview, inv);
    mat_transpose(inv, normal_matrix);
}

/* geometry.cpp:35-42 */
void orc_transform_direction(const float m[16], const float v[3], float out[3]) {
    v4 t = mat_vec(m, v4_make(v[0], v[1], v[2], 0.f));
    out[0] = t.x; out[1] = t.y; out[2] = t.z;
}

/* geometry.cpp:124-133; glm::normalize = v * (1/sqrt(dot(v,v))) */
void orc_transform_lights(const float view[16], orc_light *lights, uint32_t n_lights) {
    for (uint32_t i = 0; i < n_lights; ++i) {
        float t[3];
        orc_transform_direction(view, lights[i].direction, t);
        float d = (t[0] * t[0] + t[1] * t[1]) + t[2] * t[2];
        float inv = 1.f / sqrtf(d);
        lights[i].trans_dir[0] = t[0] * inv;
        lights[i].trans_dir[1] = t[1] * inv;
        lights[i].trans_dir[2] = t[2] * inv;
    }
}

/* geometry.cpp:44-50 (transform_point), :52-60 (z_divide), :62-74 (remap_ndc, ndc_to_raster) */
void orc_raster_vertex(const float camera[16], uint32_t width, uint32_t height, const float pos[3], float out[4]) {
    v4 clip = mat_vec(camera, v4_make(pos[0], pos[1], pos[2], 1.f));
    float nx = clip.x / clip.w, ny = clip.y / clip.w, nz = clip.z / clip.w, nw = 1.f / clip.w;
    /* remap_ndc(value, high) = 0.5f*(value + 1.0f)*high ; width/height are int -> float */
    float fw = (float)(int)width, fh = (float)(int)height;
    out[0] = 0.5f * (nx + 1.0f) * fw;
    out[1] = 0.5f * (-ny + 1.0f) * fh;
    out[2] = nz;
    out[3] = nw;
}

/* geometry.cpp:76-83 */
float orc_signed_area_2d(const float v0[4], const float v1[4], const float v2[4]) {
    return -0.5f * (v0[0] * v1[1] - v1[0] * v0[1] +
                    v1[0] * v2[1] - v2[0] * v1[1] +
                    v2[0] * v0[1] - v0[0] * v2[1]);
}

/* ------------------------------------------------------------------ */
/* shading.cpp / material.cpp                                          */
/* ------------------------------------------------------------------ */

/* The reference casts vec3 -> uvec3 (shading.cpp:33); a negative float to
 * unsigned is undefined in C++, so the oracle fixes what x86-64 gcc emits
 * (cvttss2si to 64 bits, keep the low 32) and the device path does the same. */
static inline uint32_t float_to_uint(float f) { return (uint32_t)(long long)f; }

/* shading.cpp:20-34 */
void orc_shade(const float normal[3], const float albedo[3], const orc_light *lights, uint32_t n_lights, uint32_t out_rgb[3]) {
    float sum[3] = {0.f, 0.f, 0.f};
    for (uint32_t l = 0; l < n_lights; ++l) {
        const orc_light *L = &lights[l];
        /* dot(normal, -trans_dir): tmp = a*b ; (tmp.x + tmp.y) + tmp.z */
        float tx = normal[0] * (-L->trans_dir[0]);
        float ty = normal[1] * (-L->trans_dir[1]);
        float tz = normal[2] * (-L->trans_dir[2]);
        float k = glm_max(0.f, (tx + ty) + tz);
        for (int c = 0; c < 3; ++c) {
            float v = L->intensity * L->colour[c];
            v = v * albedo[c];
            v = v * k;
            v = v * 0.318309886183790671537767526745028724f;
            sum[c] = sum[c] + v;
        }
    }
    for (int c = 0; c < 3; ++c) out_rgb[c] = float_to_uint(glm_min(sum[c], 255.f));
}

/* CImg.h:5184-5186 cimg::cut */
static inline float cimg_cut(float val, float lo, float hi) { return val < lo ? lo : val > hi ? hi : val; }

/* CImg.h:13475-13492 _linear_atXY on one channel plane */
static float linear_at_xy(const float *plane, int w, int h, float fx, float fy) {
    const float nfx = cimg_cut(fx, (float)0, (float)(w - 1));
    const float nfy = cimg_cut(fy, (float)0, (float)(h - 1));
    const unsigned int x = float_to_uint(nfx), y = float_to_uint(nfy);
    const float dx = nfx - (float)x, dy = nfy - (float)y;
    const unsigned int nx = dx > 0 ? x + 1 : x, ny = dy > 0 ? y + 1 : y;
    const float Icc = plane[x + (size_t)y * w], Inc = plane[nx + (size_t)y * w];
    const float Icn = plane[x + (size_t)ny * w], Inn = plane[nx + (size_t)ny * w];
    return Icc + dx * (Inc - Icc + dy * (Icc + Inn - Icn - Inc)) + dy * (Icn - Icc);
}

/* material.cpp:11-26 */
void orc_material_sample(const orc_material *m, const float uv[2], float out[3]) {
    if (m->has_texture) {
        float u = uv[0] * (float)m->tex_w;
        float v = (1.f - uv[1]) * (float)m->tex_h;
        size_t plane = (size_t)m->tex_w * m->tex_h;
        out[0] = linear_at_xy(m->texels, m->tex_w, m->tex_h, u, v);
        out[1] = linear_at_xy(m->texels + plane, m->tex_w, m->tex_h, u, v);
        out[2] = linear_at_xy(m->texels + 2 * plane, m->tex_w, m->tex_h, u, v);
        /* EXTENSION (not in the reference, which drops Kd of a textured material: fileloader.cpp:47-58 / material.cpp:19-21):
         * has_texture bit 1 = the texel modulates the material's Kd, one rounded product per channel */
        if (m->has_texture & 2) { out[0] = m->kd[0] * out[0]; out[1] = m->kd[1] * out[1]; out[2] = m->kd[2] * out[2]; }
    } else {
        out[0] = m->kd[0]; out[1] = m->kd[1]; out[2] = m->kd[2];
    }
}

/* ------------------------------------------------------------------ */
/* drawing.cpp                                                         */
/* ------------------------------------------------------------------ */

/* drawing.cpp:36-39 */
static inline float edge(float px, float py, const float *a, const float *b) {
    return (b[0] - a[0]) * (py - a[1]) - (b[1] - a[1]) * (px - a[0]);
}

/* Unpinned corner: material index -1 (no .mtl found) indexes
 * materials[-1] in the reference (drawing.cpp:173, undefined behaviour); both
 * oracle and device define it as untextured white like add_square's material
 * (renderer.cpp:49). */
static const orc_material k_default_material = {{1.f, 1.f, 1.f}, 0, 0, 0, 0};

typedef struct {
    const orc_scene *scene;
    const orc_light *lights;
    uint32_t n_lights;
    uint32_t width, height;
    int wind_clockwise;
    const float *raster;  /* 4 floats per vertex */
    const float *cnormal; /* 3 floats per normal */
    const float *campos;  /* 3 floats per vertex: xyz(modelview * (v,1)) (drawing.cpp:232-233); NULL unless flat == ORC_FLAT_FACE */
    uint8_t *frame;
    float *depth;
    uint32_t *tri_id;
} raster_job;

/* draw_triangle + update_pixel (drawing.cpp:96-203) for rows [y0,y1) */
static void raster_band(const raster_job *J, uint32_t y0, uint32_t y1, orc_counters *cnt) {
    const orc_scene *S = J->scene;
    const uint32_t W = J->width, H = J->height;
    const size_t plane = (size_t)W * H;
    const float zero3[3] = {0.f, 0.f, 0.f};
    orc_counters c = {0, 0, 0, 0};

    for (uint64_t t = 0; t < S->n_tris; ++t) {
        const int32_t *face = S->tris + t * 10;
        /* de-index (drawing.cpp:165-173) */
        const float *v[3], *n[3];
        float uv[3][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
        for (int k = 0; k < 3; ++k) {
            v[k] = J->raster + (size_t)face[k] * 4;
            n[k] = face[3 + k] >= 0 ? J->cnormal + (size_t)face[3 + k] * 3 : zero3;
            if (face[6 + k] >= 0) { uv[k][0] = S->uvs[(size_t)face[6 + k] * 2]; uv[k][1] = S->uvs[(size_t)face[6 + k] * 2 + 1]; }
        }
        const orc_material *mat = (face[9] >= 0 && (uint32_t)face[9] < S->n_materials) ? &S->materials[face[9]] : &k_default_material;

        /* cull (drawing.cpp:176-180) */
        float area2d = orc_signed_area_2d(v[0], v[1], v[2]);
        int front = (area2d > 0) ^ (J->wind_clockwise != 0);
        if (!front) continue;
        c.front_facing++;

        /* bounding_box (drawing.cpp:77-93): glm::min/max 3-arg = min(min(x,y),z); clamp = min(max(x,lo),hi) */
        float brx = (float)(W - 1u), bry = (float)(H - 1u);
        float minx = glm_min(glm_min(v[0][0], v[1][0]), v[2][0]);
        float miny = glm_min(glm_min(v[0][1], v[1][1]), v[2][1]);
        float maxx = ceilf(glm_max(glm_max(v[0][0], v[1][0]), v[2][0]));
        float maxy = ceilf(glm_max(glm_max(v[0][1], v[1][1]), v[2][1]));
        uint32_t tlx = float_to_uint(glm_min(glm_max(minx, 0.f), brx));
        uint32_t tly = float_to_uint(glm_min(glm_max(miny, 0.f), bry));
        uint32_t rbx = float_to_uint(glm_min(glm_max(maxx, 0.f), brx));
        uint32_t rby = float_to_uint(glm_min(glm_max(maxy, 0.f), bry));

        /* band restriction (not in the reference; pixels are independent) */
        uint32_t ys = tly > y0 ? tly : y0;
        for (uint32_t y = ys; y <= rby && y < y1; ++y) {
            for (uint32_t x = tlx; x <= rbx; ++x) {
                c.bbox_tests++;
                /* barycentric (drawing.cpp:41-49) */
                float px = (float)x, py = (float)y;
                float area = edge(v[2][0], v[2][1], v[0], v[1]);
                float b0 = edge(px, py, v[1], v[2]) / area;
                float b1 = edge(px, py, v[2], v[0]) / area;
                float b2 = edge(px, py, v[0], v[1]) / area;
                if (!(b0 >= 0.f && b1 >= 0.f && b2 >= 0.f)) continue; /* :111 */
                c.covered++;
                /* screen_interpolate (drawing.cpp:51-56,115-116) */
                float ndcdepth = v[0][2] * b0 + v[1][2] * b1 + v[2][2] * b2;
                size_t idx = (size_t)x + (size_t)y * W;
                if (!(ndcdepth < J->depth[idx])) continue; /* :119 */
                c.depth_passes++;
                J->depth[idx] = ndcdepth;
                if (J->tri_id) J->tri_id[idx] = (uint32_t)t;
                /* interpolation_coords / perspective depth (drawing.cpp:125-128) */
                float i0 = v[0][3] * b0, i1 = v[1][3] * b1, i2 = v[2][3] * b2;
                float d = 1.f / (i0 + i1 + i2);
                /* perspective_interpolate normals + normalize (drawing.cpp:64-75,131-132) */
                float m[3];
                for (int k = 0; k < 3; ++k) m[k] = d * (i0 * n[0][k] + i1 * n[1][k] + i2 * n[2][k]);
                if (J->campos) { /* extension: face normal = cross(c1 - c0, c2 - c0) (glm::cross order), then normalize */
                    const float *c0 = J->campos + (size_t)face[0] * 3, *c1 = J->campos + (size_t)face[1] * 3, *c2 = J->campos + (size_t)face[2] * 3;
                    float ea[3] = {c1[0] - c0[0], c1[1] - c0[1], c1[2] - c0[2]}, eb[3] = {c2[0] - c0[0], c2[1] - c0[1], c2[2] - c0[2]};
                    m[0] = ea[1] * eb[2] - eb[1] * ea[2];
                    m[1] = ea[2] * eb[0] - eb[2] * ea[0];
                    m[2] = ea[0] * eb[1] - eb[0] * ea[1];
                }
                float inv = 1.f / sqrtf((m[0] * m[0] + m[1] * m[1]) + m[2] * m[2]);
                float normal[3] = {m[0] * inv, m[1] * inv, m[2] * inv};
                if (J->wind_clockwise) { normal[0] = -normal[0]; normal[1] = -normal[1]; normal[2] = -normal[2]; }
                /* uv (drawing.cpp:135) */
                float tuv[2];
                for (int k = 0; k < 2; ++k) tuv[k] = d * (i0 * uv[0][k] + i1 * uv[1][k] + i2 * uv[2][k]);
                /* shade (drawing.cpp:138-146) */
                float albedo[3];
                uint32_t rgb[3];
                orc_material_sample(mat, tuv, albedo);
                orc_shade(normal, albedo, J->lights, J->n_lights, rgb);
                J->frame[idx] = (uint8_t)rgb[0];
                J->frame[idx + plane] = (uint8_t)rgb[1];
                J->frame[idx + 2 * plane] = (uint8_t)rgb[2];
            }
        }
    }
    if (cnt) *cnt = c;
}

/* the per-vertex passes of draw_frame (drawing.cpp:222-247) */
static int prepare(const orc_scene *scene, orc_light *lights, uint32_t n_lights, const orc_args *args,
                   float **raster_out, float **cnormal_out, float **campos_out) {
    float modelview[16], camera[16], normal_matrix[16], view[16];
    orc_frame_matrices(args, modelview, camera, normal_matrix, view);
    orc_transform_lights(view, lights, n_lights);
    float *raster = (float *)malloc((size_t)(scene->n_positions ? scene->n_positions : 1) * 4 * sizeof(float));
    float *cnormal = (float *)malloc((size_t)(scene->n_normals ? scene->n_normals : 1) * 3 * sizeof(float));
    if (!raster || !cnormal) { free(raster); free(cnormal); return -1; }
    for (uint32_t i = 0; i < scene->n_positions; ++i)
        orc_raster_vertex(camera, args->image_width, args->image_height, scene->positions + (size_t)i * 3, raster + (size_t)i * 4);
    for (uint32_t i = 0; i < scene->n_normals; ++i)
        orc_transform_direction(normal_matrix, scene->normals + (size_t)i * 3, cnormal + (size_t)i * 3);
    *campos_out = 0;
    if (args->flat == ORC_FLAT_FACE) { /* transform_vertices(modelview) + xyz_all (drawing.cpp:232-233) */
        float *campos = (float *)malloc((size_t)(scene->n_positions ? scene->n_positions : 1) * 3 * sizeof(float));
        if (!campos) { free(raster); free(cnormal); return -1; }
        for (uint32_t i = 0; i < scene->n_positions; ++i) {
            const float *p = scene->positions + (size_t)i * 3;
            v4 c = mat_vec(modelview, v4_make(p[0], p[1], p[2], 1.f));
            campos[(size_t)i * 3] = c.x; campos[(size_t)i * 3 + 1] = c.y; campos[(size_t)i * 3 + 2] = c.z;
        }
        *campos_out = campos;
    }
    *raster_out = raster;
    *cnormal_out = cnormal;
    return 0;
}

void orc_draw_frame(const orc_scene *scene, orc_light *lights, uint32_t n_lights,
                    const orc_args *args, uint8_t *frame, float *depth, uint32_t *tri_id,
                    uint32_t band_y0, uint32_t band_y1, orc_counters *counters) {
    float *raster, *cnormal, *campos;
    if (prepare(scene, lights, n_lights, args, &raster, &cnormal, &campos)) return;
    raster_job J = {scene, lights, n_lights, args->image_width, args->image_height, args->wind_clockwise,
                    raster, cnormal, campos, frame, depth, tri_id};
    if (band_y1 > args->image_height) band_y1 = args->image_height;
    raster_band(&J, band_y0, band_y1, counters);
    free(raster);
    free(cnormal);
    free(campos);
}

typedef struct {
    const raster_job *job;
    uint32_t n_bands;
    uint32_t *next_band;       /* shared cursor, fetched atomically */
    orc_counters counters;     /* per worker, summed after join */
    int spawned;               /* 1 if it runs on its own pthread */
} band_worker;

static void *band_worker_main(void *arg) {
    band_worker *w = (band_worker *)arg;
    const uint32_t H = w->job->height;
    for (;;) {
        uint32_t b = __atomic_fetch_add(w->next_band, 1u, __ATOMIC_RELAXED);
        if (b >= w->n_bands) break;
        uint32_t y0 = (uint32_t)((uint64_t)H * b / w->n_bands), y1 = (uint32_t)((uint64_t)H * (b + 1) / w->n_bands);
        orc_counters c;
        raster_band(w->job, y0, y1, &c);
        w->counters.bbox_tests += c.bbox_tests;
        w->counters.covered += c.covered;
        w->counters.depth_passes += c.depth_passes;
        w->counters.front_facing = c.front_facing; /* every band walks every triangle */
    }
    return 0;
}

void orc_draw_frame_mt(const orc_scene *scene, orc_light *lights, uint32_t n_lights,
                       const orc_args *args, uint8_t *frame, float *depth, uint32_t *tri_id,
                       int n_threads, orc_counters *counters) {
    float *raster, *cnormal, *campos;
    if (prepare(scene, lights, n_lights, args, &raster, &cnormal, &campos)) return;
    raster_job J = {scene, lights, n_lights, args->image_width, args->image_height, args->wind_clockwise,
                    raster, cnormal, campos, frame, depth, tri_id};
    const uint32_t H = args->image_height;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 1024) n_threads = 1024;
    /* many thin bands pulled from a shared cursor, so a centred model balances */
    uint32_t n_bands = (uint32_t)n_threads * 8u;
    if (n_bands > H) n_bands = H ? H : 1;
    uint32_t next_band = 0;
    band_worker *workers = (band_worker *)calloc((size_t)n_threads, sizeof(band_worker));
    pthread_t *threads = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    orc_counters total = {0, 0, 0, 0};
    if (workers && threads) {
        for (int i = 0; i < n_threads; ++i) {
            workers[i].job = &J;
            workers[i].n_bands = n_bands;
            workers[i].next_band = &next_band;
        }
        /* workers 1.. on their own threads, worker 0 on the caller (also drains if a spawn failed) */
        for (int i = 1; i < n_threads; ++i)
            workers[i].spawned = pthread_create(&threads[i], 0, band_worker_main, &workers[i]) == 0;
        band_worker_main(&workers[0]);
        for (int i = 0; i < n_threads; ++i) {
            if (workers[i].spawned) pthread_join(threads[i], 0);
            total.bbox_tests += workers[i].counters.bbox_tests;
            total.covered += workers[i].counters.covered;
            total.depth_passes += workers[i].counters.depth_passes;
            if (workers[i].counters.front_facing) total.front_facing = workers[i].counters.front_facing;
        }
    }
    if (counters) *counters = total;
    free(workers);
    free(threads);
    free(raster);
    free(cnormal);
    free(campos);
}

void orc_clear(uint32_t width, uint32_t height, uint8_t *frame, float *depth, uint32_t *tri_id) {
    size_t p = (size_t)width * height;
    if (frame) memset(frame, 0, 3 * p);
    if (depth) for (size_t i = 0; i < p; ++i) depth[i] = 1.f;
    if (tri_id) for (size_t i = 0; i < p; ++i) tri_id[i] = ORC_NO_TRIANGLE;
}

/* CImg.h:26786-26794 normalize(0,255) then uchar truncation in the PNM writer (CImg.h:52410) */
void orc_depth_to_u8(const float *depth, uint64_t n, uint8_t *out) {
    if (!n) return;
    float m = depth[0], M = depth[0];
    for (uint64_t i = 0; i < n; ++i) { /* max_min, CImg.h:23715-23729 */
        float val = depth[i];
        if (val > M) M = val;
        if (val < m) m = val;
    }
    const float a = 0.f, b = 255.f;
    if (m == M) { memset(out, 0, n); return; } /* fill(min_value) */
    for (uint64_t i = 0; i < n; ++i) {
        float v = depth[i];
        if (m != a || M != b) v = (v - m) / (M - m) * (b - a) + a;
        out[i] = (uint8_t)v;
    }
}

void orc_normalize_texture(float *texels, uint64_t n) {
    if (!n) return;
    float m = texels[0], M = texels[0];
    for (uint64_t i = 0; i < n; ++i) {
        float val = texels[i];
        if (val > M) M = val;
        if (val < m) m = val;
    }
    const float a = 0.f, b = 1.f;
    if (m == M) { for (uint64_t i = 0; i < n; ++i) texels[i] = a; return; }
    if (m != a || M != b)
        for (uint64_t i = 0; i < n; ++i) texels[i] = (texels[i] - m) / (M - m) * (b - a) + a;
}

float orc_spin_angle(float ry0, uint32_t k, uint32_t n_frames) {
    return ry0 + (float)k * (6.2831853f / (float)n_frames);
}

uint64_t orc_fnv1a64(const void *data, uint64_t n_bytes) {
    const uint8_t *p = (const uint8_t *)data;
    uint64_t h = 0xcbf29ce484222325ull;
    for (uint64_t i = 0; i < n_bytes; ++i) { h ^= p[i]; h *= 0x100000001b3ull; }
    return h;
}
