// rast_draw_frame.hpp -- C++ shim with the reference's draw_frame signature on top of the C ABI.
//
// The reference's only entry into the frame path is (headers/drawing.h:16-18):
//
//   void draw_frame(const std::vector<glm::vec3>& model_vertices, const std::vector<Triangle>& faces,
//                   const std::vector<glm::vec3>& model_vertnormals, const std::vector<glm::vec2>& vertuvs,
//                   std::vector<Light>& lights, const std::vector<Material>& materials, const Args& arguments,
//                   cimg_library::CImg<unsigned char>* frame_buffer, cimg_library::CImg<float>* depth_buffer);
//
// rast::draw_frame below takes the same arguments in the same order, as templates, so it binds to the
// reference's own types (glm::vec3 = 3 packed floats, struct Triangle = 10 ints, CImg<T>::data()) and to
// this repository's host types alike.  Materials need an adapter because the reference's class keeps its
// fields private: pass anything with kd / has_texture / tex_w / tex_h / texels members (host::MaterialData)
// -- INTEGRATION.md shows the three accessor lines to add to headers/material.h.
//
// Semantics: the reference reads its vectors on every call.  Here the scene is uploaded on the first call and again
// whenever it is not the same scene any more: the vectors' addresses, sizes AND contents are compared (a 64-bit hash of
// every array per call -- a few GB/s, microseconds for a Suzanne-sized scene -- so vertices edited in place are seen).
// A caller that redraws one huge static scene (the spin loop, renderer.cpp:105-111) can skip the hashing with
// Session::assume_unchanged(true) -- then only addresses and sizes are compared and Session::invalidate() forces the
// next upload.  The finished frame and depth overwrite the caller's buffers (the reference's callers always pass cleared
// buffers: renderer.cpp:85-86,107-108), lights[i].trans_dir is written like Light::transform
// (geometry.cpp:126).  Errors throw std::runtime_error on the C++ side of the ABI.
#pragma once

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "rast.h"

namespace rast {

class Session {
public:
    explicit Session(int device = 0) {
        if (rast_create(device, &ctx_) != RAST_OK) throw std::runtime_error(std::string("rast_create: ") + rast_last_error(nullptr));
    }
    ~Session() { unpin_outputs(); rast_destroy(ctx_); }
    Session(const Session &) = delete;
    Session &operator=(const Session &) = delete;
    rast_ctx *ctx() const { return ctx_; }
    void check(int rc, const char *what) const {
        if (rc != RAST_OK) throw std::runtime_error(std::string(what) + ": " + rast_last_error(ctx_));
    }

    template <class Vec3s, class Faces, class Vec2s, class Materials>
    void upload(const Vec3s &vertices, const Faces &faces, const Vec3s &normals, const Vec2s &uvs, const Materials &materials) {
        static_assert(sizeof(typename Vec3s::value_type) == 12 && sizeof(typename Vec2s::value_type) == 8, "vec3 / vec2 must be packed floats");
        static_assert(sizeof(typename Faces::value_type) == 40, "Triangle must be 10 x int32 (headers/face.h:6-13)");
        check(rast_upload_mesh(ctx_, reinterpret_cast<const float *>(vertices.data()), (uint32_t)vertices.size(),
                               reinterpret_cast<const float *>(normals.data()), (uint32_t)normals.size(),
                               reinterpret_cast<const float *>(uvs.data()), (uint32_t)uvs.size(),
                               reinterpret_cast<const int32_t *>(faces.data()), (uint64_t)faces.size()), "rast_upload_mesh");
        std::vector<rast_material> m(materials.size());
        for (size_t i = 0; i < materials.size(); ++i) {
            m[i].kd[0] = materials[i].kd[0]; m[i].kd[1] = materials[i].kd[1]; m[i].kd[2] = materials[i].kd[2];
            m[i].has_texture = materials[i].has_texture ? (1 | texture_bits_) : 0;
            m[i].tex_w = materials[i].tex_w; m[i].tex_h = materials[i].tex_h;
            m[i].texels = materials[i].has_texture ? &materials[i].texels[0] : nullptr;
        }
        check(rast_upload_materials(ctx_, m.data(), (uint32_t)m.size()), "rast_upload_materials");
        key_ = make_key(vertices, faces, normals, uvs, materials, true);
        uploaded_ = true;
    }

    // Extension (default off, the reference ignores Kd of a textured material: material.cpp:19-21): RAST_TEXTURE_MODULATE_KD makes the
    // texel modulate Kd.  Takes effect at the next upload.
    void set_texture_bits(int bits) { if (bits != texture_bits_) { texture_bits_ = bits; uploaded_ = false; } }

    // Opt-in: page-lock the frame / depth buffers handed to draw_frame (rast_host_register) the first time they are seen -- draws into
    // pageable memory go through the driver's staging copies and are several times slower.  The caller promises that the buffers
    // outlive the Session or calls unpin_outputs() before freeing them (the reference's frame_buffer / depth_buffer live for the whole
    // run: renderer.cpp:85-86).  retained_outputs(true) adds the promise of rast_set_retained_outputs (the spin loop's own behaviour).
    void pin_outputs(bool on) { pin_outputs_ = on; if (!on) unpin_outputs(); }
    void retained_outputs(bool on) { check(rast_set_retained_outputs(ctx_, on ? 1 : 0), "rast_set_retained_outputs"); }
    void unpin_outputs() {
        for (void *p : pinned_) rast_host_unregister(p);
        pinned_.clear();
    }
    void note_output(void *p, size_t bytes) {
        if (!pin_outputs_ || !p || !bytes) return;
        for (void *q : pinned_) if (q == p) return;
        if (rast_host_register(p, bytes) == RAST_OK) pinned_.push_back(p);
    }

    // The caller promises that the scene arrays do not change between draws (no per-call content hash); invalidate() when they do.
    void assume_unchanged(bool on) { assume_unchanged_ = on; }
    void invalidate() { uploaded_ = false; }

    template <class Vec3s, class Faces, class Vec2s, class Materials>
    bool holds(const Vec3s &vertices, const Faces &faces, const Vec3s &normals, const Vec2s &uvs, const Materials &materials) const {
        if (!uploaded_) return false;
        const Key k = make_key(vertices, faces, normals, uvs, materials, !assume_unchanged_);
        return k.v == key_.v && k.f == key_.f && k.n == key_.n && k.u == key_.u && k.m == key_.m && k.nv == key_.nv && k.nf == key_.nf && k.nn == key_.nn &&
               k.nu == key_.nu && k.nm == key_.nm && (assume_unchanged_ || k.hash == key_.hash);
    }

private:
    struct Key { const void *v, *f, *n, *u, *m; size_t nv, nf, nn, nu, nm; uint64_t hash; };
    // 64-bit multiply-xorshift hash over 8-byte words (4 independent lanes so that it runs at memory speed)
    static uint64_t hash_bytes(const void *p, size_t bytes, uint64_t h) {
        const unsigned char *b = static_cast<const unsigned char *>(p);
        uint64_t l[4] = {h ^ 0x9E3779B97F4A7C15ull, h ^ 0xC2B2AE3D27D4EB4Full, h ^ 0x165667B19E3779F9ull, h ^ 0x27D4EB2F165667C5ull};
        size_t i = 0;
        for (; i + 32 <= bytes; i += 32) {
            uint64_t w[4];
            std::memcpy(w, b + i, 32);
            for (int k = 0; k < 4; ++k) { l[k] = (l[k] ^ w[k]) * 0x100000001B3ull; l[k] ^= l[k] >> 29; }
        }
        uint64_t tail[4] = {0, 0, 0, 0};
        if (i < bytes) std::memcpy(tail, b + i, bytes - i);
        for (int k = 0; k < 4; ++k) { l[k] = (l[k] ^ tail[k]) * 0x100000001B3ull; l[k] ^= l[k] >> 29; }
        uint64_t out = bytes;
        for (int k = 0; k < 4; ++k) { out = (out ^ l[k]) * 0xFF51AFD7ED558CCDull; out ^= out >> 33; }
        return out;
    }
    template <class Vec3s, class Faces, class Vec2s, class Materials>
    static Key make_key(const Vec3s &vertices, const Faces &faces, const Vec3s &normals, const Vec2s &uvs, const Materials &materials, bool with_hash) {
        Key k = {vertices.data(), faces.data(), normals.data(), uvs.data(), materials.data(), vertices.size(), faces.size(), normals.size(), uvs.size(), materials.size(), 0};
        if (with_hash) {
            uint64_t h = hash_bytes(vertices.data(), vertices.size() * sizeof(typename Vec3s::value_type), 1);
            h = hash_bytes(faces.data(), faces.size() * sizeof(typename Faces::value_type), h);
            h = hash_bytes(normals.data(), normals.size() * sizeof(typename Vec3s::value_type), h);
            h = hash_bytes(uvs.data(), uvs.size() * sizeof(typename Vec2s::value_type), h);
            for (size_t i = 0; i < materials.size(); ++i) {
                const float kd[3] = {(float)materials[i].kd[0], (float)materials[i].kd[1], (float)materials[i].kd[2]};
                h = hash_bytes(kd, sizeof kd, h);
                if (materials[i].has_texture)
                    h = hash_bytes(&materials[i].texels[0], (size_t)materials[i].tex_w * (size_t)materials[i].tex_h * 3 * sizeof(float), h ^ (uint64_t)materials[i].tex_w);
            }
            k.hash = h;
        }
        return k;
    }
    rast_ctx *ctx_ = nullptr;
    Key key_{};
    bool uploaded_ = false, assume_unchanged_ = false, pin_outputs_ = false;
    int texture_bits_ = 0;
    std::vector<void *> pinned_;
};

template <class ArgsT> inline rast_args to_rast_args(const ArgsT &a) {
    rast_args r;
    r.image_width = a.image_width;
    r.image_height = a.image_height;
    r.aspect_ratio = a.aspect_ratio;
    r.scale = a.scale;
    for (int k = 0; k < 3; ++k) {
        r.displacement[k] = a.displacement[k];
        r.tait_bryan_angles[k] = a.tait_bryan_angles[k];
    }
    r.wind_clockwise = a.wind_clockwise ? 1 : 0;
    r.flat = a.flat ? 1 : 0; // like the reference, the path ignores it; the host program's --flat-mode face sets RAST_FLAT_FACE itself
    return r;
}

// Same argument order as the reference's draw_frame.  Lights: any vector whose elements are laid out as
// struct Light (direction[3], intensity, colour[3], trans_dir[3] = 10 floats, headers/light.h:7-14).
template <class Vec3s, class Faces, class Vec2s, class Lights, class Materials, class ArgsT, class FrameImage, class DepthImage>
void draw_frame(Session &session, const Vec3s &model_vertices, const Faces &faces, const Vec3s &model_vertnormals, const Vec2s &vertuvs,
                Lights &lights, const Materials &materials, const ArgsT &arguments, FrameImage *frame_buffer, DepthImage *depth_buffer) {
    static_assert(sizeof(typename Lights::value_type) == sizeof(rast_light), "Light must be 10 packed floats");
    if (!session.holds(model_vertices, faces, model_vertnormals, vertuvs, materials))
        session.upload(model_vertices, faces, model_vertnormals, vertuvs, materials);
    rast_light *l = reinterpret_cast<rast_light *>(lights.data());
    session.check(rast_set_lights(session.ctx(), l, (uint32_t)lights.size()), "rast_set_lights");
    const rast_args a = to_rast_args(arguments);
    session.note_output(frame_buffer->data(), (size_t)a.image_width * a.image_height * 3);
    if (depth_buffer) session.note_output(depth_buffer->data(), (size_t)a.image_width * a.image_height * sizeof(float));
    session.check(rast_draw_frame(session.ctx(), &a, frame_buffer->data(), depth_buffer ? depth_buffer->data() : nullptr, l), "rast_draw_frame");
}

} // namespace rast
