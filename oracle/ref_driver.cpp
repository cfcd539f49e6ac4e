/*
 * ref_driver.cpp -- C-callable wrapper around the UNMODIFIED reference hot
 * path (TEST INFRASTRUCTURE ONLY).  Compiled by oracle/build_ref.sh together
 * with /root/reference/{geometry,drawing,shading,material,fileloader}.cpp,
 * from where those sources lie, into oracle/_ref/libref.so.  No reference
 * source is copied into this repository.  It replaces renderer.cpp /
 * arguments.cpp (which need TCLAP and X11) with a handle API the Python
 * tests drive through ctypes: same vectors, same draw_frame call
 * (headers/drawing.h:16-18), same buffer initialisation (renderer.cpp:85-86).
 */
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include <glm/vec2.hpp>
#include <glm/vec3.hpp>
#include <glm/vec4.hpp>

#include "CImg.h"

#include "arguments.h"
#include "drawing.h"
#include "face.h"
#include "fileloader.h"
#include "geometry.h"
#include "light.h"
#include "material.h"
#include "shading.h"

/* arguments.cpp is not linked (TCLAP is unavailable); fields are set by the caller. */
Args::Args(int, char **)
    : image_width(540u), image_height(304u), aspect_ratio(540.f / 304.f), spin(false), flat(false),
      wind_clockwise(false), scale(1.f), displacement(0.f), tait_bryan_angles(0.f) {}

namespace {

struct RefScene {
    std::vector<glm::vec3> vertices;
    std::vector<glm::vec3> normals;
    std::vector<glm::vec2> uvs;
    std::vector<Triangle> faces;
    std::vector<Material> materials;
};

Args make_args(uint32_t w, uint32_t h, float scale, const float *disp, const float *angles, int wind_clockwise, int flat) {
    Args a(0, 0);
    a.image_width = w;
    a.image_height = h;
    a.aspect_ratio = (float)w / (float)h; /* arguments.cpp:39 */
    a.scale = scale;
    a.displacement = glm::vec3(disp[0], disp[1], disp[2]);
    a.tait_bryan_angles = glm::vec3(angles[0], angles[1], angles[2]);
    a.wind_clockwise = wind_clockwise != 0;
    a.flat = flat != 0;
    return a;
}

void mat_out(const glm::mat4 &m, float *out) {
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) out[c * 4 + r] = m[c][r];
}

glm::mat4 mat_in(const float *in) {
    glm::mat4 m(0.f);
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) m[c][r] = in[c * 4 + r];
    return m;
}

} // namespace

extern "C" {

/* kd: 3 floats per material; tex_paths[i] NULL/"" => untextured (material.h:19), else a file CImg can
 * read natively (binary PPM) => textured ctor (material.h:20-23, loads + normalize(0,1)). */
void *ref_scene_create(const float *pos, uint32_t n_pos, const float *nrm, uint32_t n_nrm,
                       const float *uv, uint32_t n_uv, const int32_t *tris, uint64_t n_tris,
                       const float *kd, const char *const *tex_paths, uint32_t n_mats) {
    RefScene *s = new RefScene();
    for (uint32_t i = 0; i < n_pos; ++i) s->vertices.push_back(glm::vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]));
    for (uint32_t i = 0; i < n_nrm; ++i) s->normals.push_back(glm::vec3(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]));
    for (uint32_t i = 0; i < n_uv; ++i) s->uvs.push_back(glm::vec2(uv[2 * i], uv[2 * i + 1]));
    for (uint64_t t = 0; t < n_tris; ++t) {
        const int32_t *f = tris + 10 * t;
        s->faces.push_back(Triangle({f[0], f[1], f[2]}, {f[3], f[4], f[5]}, {f[6], f[7], f[8]}, f[9]));
    }
    try {
        for (uint32_t m = 0; m < n_mats; ++m) {
            glm::vec3 dc(kd[3 * m], kd[3 * m + 1], kd[3 * m + 2]);
            if (tex_paths && tex_paths[m] && tex_paths[m][0]) s->materials.push_back(Material(dc, std::string(tex_paths[m])));
            else s->materials.push_back(Material(dc));
        }
    } catch (...) {
        delete s;
        return 0;
    }
    return s;
}

void ref_scene_destroy(void *scene) { delete static_cast<RefScene *>(scene); }

/* lights: 10 floats each = direction[3], intensity, colour[3], trans_dir[3] (trans_dir written back).
 * frame/depth are allocated and initialised exactly as renderer.cpp:85-86 does, drawn by the
 * reference's draw_frame, and copied out in CImg's planar layout. */
int ref_scene_draw(void *scene, float *lights, uint32_t n_lights, uint32_t w, uint32_t h, float scale,
                   const float *disp, const float *angles, int wind_clockwise, int flat,
                   uint8_t *frame_out, float *depth_out) {
    RefScene *s = static_cast<RefScene *>(scene);
    if (!s) return -1;
    Args args = make_args(w, h, scale, disp, angles, wind_clockwise, flat);
    std::vector<Light> lv;
    for (uint32_t i = 0; i < n_lights; ++i) {
        const float *l = lights + 10 * i;
        lv.push_back(Light(glm::vec3(l[0], l[1], l[2]), l[3], glm::vec3(l[4], l[5], l[6])));
    }
    cimg_library::CImg<unsigned char> frame_buffer(w, h, 1, 3, 0);
    cimg_library::CImg<float> depth_buffer(w, h, 1, 1, 1.f);
    draw_frame(s->vertices, s->faces, s->normals, s->uvs, lv, s->materials, args, &frame_buffer, &depth_buffer);
    for (uint32_t i = 0; i < n_lights; ++i) {
        lights[10 * i + 7] = lv[i].trans_dir.x;
        lights[10 * i + 8] = lv[i].trans_dir.y;
        lights[10 * i + 9] = lv[i].trans_dir.z;
    }
    std::memcpy(frame_out, frame_buffer.data(), (size_t)w * h * 3);
    std::memcpy(depth_out, depth_buffer.data(), (size_t)w * h * sizeof(float));
    return 0;
}

/* depth_buffer.normalize(0,255) + the uchar truncation of CImg's PNM writer (renderer.cpp:93). */
void ref_depth_to_u8(const float *depth, uint32_t w, uint32_t h, uint8_t *out) {
    cimg_library::CImg<float> d(depth, w, h, 1, 1);
    d.normalize(0, 255);
    for (size_t i = 0; i < (size_t)w * h; ++i) out[i] = (unsigned char)d.data()[i];
}

/* known-answer hooks: each forwards to one reference function */
void ref_transformation_matrix(float factor, const float *disp, const float *tb, float *out) {
    mat_out(transformation_matrix(factor, glm::vec3(disp[0], disp[1], disp[2]), glm::vec3(tb[0], tb[1], tb[2])), out);
}
void ref_camera_matrix(const float *modelview, float aspect, float *out) { mat_out(camera_matrix(mat_in(modelview), aspect), out); }
void ref_normal_matrix(const float *modelview, float *out) { mat_out(glm::transpose(glm::inverse(mat_in(modelview))), out); } /* geometry.cpp:101 */
void ref_transform_direction(const float *m, const float *v, float *out) {
    glm::vec3 r = transform_direction(mat_in(m), glm::vec3(v[0], v[1], v[2]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void ref_raster_vertex(const float *camera, int w, int h, const float *p, float *out) {
    glm::vec4 r = ndc_to_raster(w, h, z_divide(transform_point(mat_in(camera), glm::vec3(p[0], p[1], p[2]))));
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}
float ref_signed_area_2d(const float *v0, const float *v1, const float *v2) {
    std::array<glm::vec4, 3> v = {glm::vec4(v0[0], v0[1], v0[2], v0[3]), glm::vec4(v1[0], v1[1], v1[2], v1[3]), glm::vec4(v2[0], v2[1], v2[2], v2[3])};
    return signed_area_2d(v);
}
void ref_transform_lights(const float *view, float *lights, uint32_t n_lights) {
    std::vector<Light> lv;
    for (uint32_t i = 0; i < n_lights; ++i) {
        const float *l = lights + 10 * i;
        lv.push_back(Light(glm::vec3(l[0], l[1], l[2]), l[3], glm::vec3(l[4], l[5], l[6])));
    }
    transform_lights(mat_in(view), lv);
    for (uint32_t i = 0; i < n_lights; ++i) {
        lights[10 * i + 7] = lv[i].trans_dir.x;
        lights[10 * i + 8] = lv[i].trans_dir.y;
        lights[10 * i + 9] = lv[i].trans_dir.z;
    }
}
void ref_shade(const float *normal, const float *albedo, const float *lights, uint32_t n_lights, uint32_t *out) {
    std::vector<Light> lv;
    for (uint32_t i = 0; i < n_lights; ++i) {
        const float *l = lights + 10 * i;
        Light L(glm::vec3(l[0], l[1], l[2]), l[3], glm::vec3(l[4], l[5], l[6]));
        L.trans_dir = glm::vec3(l[7], l[8], l[9]);
        lv.push_back(L);
    }
    glm::uvec3 r = shade(glm::vec3(normal[0], normal[1], normal[2]), glm::vec3(albedo[0], albedo[1], albedo[2]), lv);
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void ref_material_sample(void *scene, uint32_t material, const float *uv, float *out) {
    RefScene *s = static_cast<RefScene *>(scene);
    glm::vec3 r = s->materials[material].sample(glm::vec2(uv[0], uv[1]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}

/* loaders (fileloader.cpp:79-133): returns a RefScene built by the reference's load_obj.
 * mats_dir must end in '/' (paths are concatenated, fileloader.cpp:55). Textures must be
 * readable by CImg without external tools (binary PPM). */
void *ref_load_obj(const char *obj_file, const char *mats_dir) {
    RefScene *s = new RefScene();
    Args args(0, 0);
    args.obj_file = obj_file;
    args.materials_directory = mats_dir ? mats_dir : "";
    try {
        load_obj(args, s->vertices, s->faces, s->normals, s->uvs, s->materials);
    } catch (...) {
        delete s;
        return 0;
    }
    return s;
}
void ref_scene_sizes(void *scene, uint64_t *out5) {
    RefScene *s = static_cast<RefScene *>(scene);
    out5[0] = s->vertices.size(); out5[1] = s->normals.size(); out5[2] = s->uvs.size();
    out5[3] = s->faces.size(); out5[4] = s->materials.size();
}
void ref_scene_copy(void *scene, float *pos, float *nrm, float *uv, int32_t *tris) {
    RefScene *s = static_cast<RefScene *>(scene);
    for (size_t i = 0; i < s->vertices.size(); ++i) { pos[3 * i] = s->vertices[i].x; pos[3 * i + 1] = s->vertices[i].y; pos[3 * i + 2] = s->vertices[i].z; }
    for (size_t i = 0; i < s->normals.size(); ++i) { nrm[3 * i] = s->normals[i].x; nrm[3 * i + 1] = s->normals[i].y; nrm[3 * i + 2] = s->normals[i].z; }
    for (size_t i = 0; i < s->uvs.size(); ++i) { uv[2 * i] = s->uvs[i].x; uv[2 * i + 1] = s->uvs[i].y; }
    for (size_t t = 0; t < s->faces.size(); ++t) {
        const Triangle &f = s->faces[t];
        for (int k = 0; k < 3; ++k) { tris[10 * t + k] = f.vertices[k]; tris[10 * t + 3 + k] = f.normals[k]; tris[10 * t + 6 + k] = f.uvs[k]; }
        tris[10 * t + 9] = f.material;
    }
}
/* returns number of lights; out (7 floats each) may be NULL to query the count */
uint32_t ref_load_lights(const char *file, float *out, uint32_t capacity) {
    std::vector<Light> lv;
    load_lights(file, lv);
    for (uint32_t i = 0; out && i < lv.size() && i < capacity; ++i) {
        out[7 * i] = lv[i].direction.x; out[7 * i + 1] = lv[i].direction.y; out[7 * i + 2] = lv[i].direction.z;
        out[7 * i + 3] = lv[i].intensity;
        out[7 * i + 4] = lv[i].colour.x; out[7 * i + 5] = lv[i].colour.y; out[7 * i + 6] = lv[i].colour.z;
    }
    return (uint32_t)lv.size();
}

} /* extern "C" */
