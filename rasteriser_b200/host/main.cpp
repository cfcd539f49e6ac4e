// main.cpp -- reference-compatible command line on the B200 frame path.
//
// Same flow as the reference's main (renderer.cpp:52-129): parse flags, load lights, load the model (or the
// built-in square), allocate frame (RGB8, 0) and depth (f32, 1.0) buffers, draw, write frame.png and the
// normalised depth.png.  What differs: draw_frame runs on the GPU through the rast_* ABI, and -s (spin)
// -- an X11 window rotated by wall-clock time in the reference (renderer.cpp:94-126) -- is a headless,
// deterministic sequence: frame k of N is rotated by ry + k * 2*pi/N; frames are rendered in
// batches, optionally written as PNG, and the measured frame rate is printed.
#include <chrono>
#include <cstring>
#include <cstdio>
#include <iostream>
#include <string>
#include <vector>

#include "../../include/rast.h"
#include "../../include/rast_draw_frame.hpp"
#include "args.hpp"
#include "image.hpp"
#include "loaders.hpp"
#include "png.hpp"

using host::Args;
using host::Image;

namespace {

struct Tri { int32_t i[10]; }; // struct Triangle (headers/face.h:6-13)
struct Vec3 { float x, y, z; };
struct Vec2 { float x, y; };

// The loader's flat arrays seen as the reference's vector<vec3> / vector<vec2> / vector<Triangle>: same bytes, no copy
// (a 50 M-triangle scene is 2 GB of indices).
template <class T> struct View {
    typedef T value_type;
    const T *p;
    size_t n;
    const T *data() const { return p; }
    size_t size() const { return n; }
};
template <class T, class U> View<T> view(const host::Pod<U> &flat) { return View<T>{reinterpret_cast<const T *>(flat.data()), flat.size() * sizeof(U) / sizeof(T)}; }

} // namespace

int main(int argc, char **argv) {
    Args arguments;
    std::string message;
    switch (host::parse_args(argc, argv, arguments, message)) {
        case host::ParseResult::Help: std::cout << host::usage_text(argv[0]); return 0;
        case host::ParseResult::Version: std::cout << argv[0] << "  version: " << rast_version() << std::endl; return 0;
        case host::ParseResult::Error:
            std::cerr << message << "\n\nBrief USAGE: \n   " << argv[0] << "  -l <lights.csv> [-o <model.obj>] ...\n\nFor complete USAGE and HELP type: \n   " << argv[0] << " --help\n" << std::endl;
            return 1;
        case host::ParseResult::Ok: break;
    }
    const bool verbose = !arguments.quiet;
    auto t_stage = std::chrono::steady_clock::now();
    auto stage = [&](const char *name) { // --timing: wall time since the previous stage
        const auto now = std::chrono::steady_clock::now();
        if (arguments.timing) std::cerr << "[timing] " << name << " " << std::chrono::duration<double>(now - t_stage).count() << " s" << std::endl;
        t_stage = now;
    };

    host::Model model;
    std::vector<host::Light> lights;
    std::string err;
    if (!host::load_lights(arguments.lights_file, lights, err)) { std::cerr << err << std::endl; return 1; } // renderer.cpp:75
    if (!arguments.obj_file.empty()) {                                                                        // renderer.cpp:78-82
        bool cached = false;
        if (!arguments.mesh_cache.empty()) {
            std::string cerr_text;
            cached = host::load_mesh_cache(arguments.mesh_cache, model, cerr_text, verbose);
            if (!cached) model = host::Model();
        }
        if (!cached) {
            const bool ok = host::load_obj(arguments.obj_file, arguments.materials_directory, model, err, verbose, arguments.load_threads);
            if (!err.empty()) std::cerr << err << std::endl; // fileloader.cpp:95-97
            if (!ok) return 1;
            if (!arguments.mesh_cache.empty() && !host::save_mesh_cache(arguments.mesh_cache, model, err)) std::cerr << err << std::endl;
        }
    } else {
        host::add_square(model);
    }
    stage("load scene");
    const View<Vec3> vertices = view<Vec3>(model.positions), normals = view<Vec3>(model.normals);
    const View<Vec2> uvs = view<Vec2>(model.uvs);
    const View<Tri> faces = view<Tri>(model.tris);

    // renderer.cpp:85-86
    Image<unsigned char> frame_buffer(arguments.image_width, arguments.image_height, 3, 0);
    Image<float> depth_buffer(arguments.image_width, arguments.image_height, 1, 1.f);

    try {
        rast::Session session(arguments.device);
        if (arguments.modulate_kd) session.set_texture_bits(RAST_TEXTURE_MODULATE_KD);
        stage("create context");
        const int flat_code = (arguments.flat && arguments.flat_face) ? RAST_FLAT_FACE : (arguments.flat ? 1 : 0);
        if (!arguments.spin) {
            if (flat_code == RAST_FLAT_FACE) { // extension path: same calls as the shim, with the extension code in rast_args.flat
                session.upload(vertices, faces, normals, uvs, model.materials);
                rast_light *l = reinterpret_cast<rast_light *>(lights.data());
                session.check(rast_set_lights(session.ctx(), l, (uint32_t)lights.size()), "rast_set_lights");
                rast_args a = rast::to_rast_args(arguments);
                a.flat = flat_code;
                session.check(rast_draw_frame(session.ctx(), &a, frame_buffer.data(), depth_buffer.data(), l), "rast_draw_frame");
            } else
            rast::draw_frame(session, vertices, faces, normals, uvs, lights, model.materials, arguments, &frame_buffer, &depth_buffer);
            stage("upload + draw_frame");
            // renderer.cpp:92-93: frame.png, and depth.normalize(0,255) saved as 8-bit grey
            err = host::png_write_planar(arguments.frame_out, frame_buffer.data(), arguments.image_width, arguments.image_height, 3);
            if (!err.empty()) { std::cerr << err << std::endl; return 1; }
            std::vector<unsigned char> depth8((size_t)arguments.image_width * arguments.image_height);
            session.check(rast_depth_to_u8(session.ctx(), depth8.data()), "rast_depth_to_u8"); // normalize + uchar cast on the device
            err = host::png_write(arguments.depth_out, depth8.data(), arguments.image_width, arguments.image_height, 1);
            if (!err.empty()) { std::cerr << err << std::endl; return 1; }
            stage("frame.png + depth.png");
        } else {
            // headless spin: N frames, ry_k = ry + k * 2*pi/N, rendered in batches through rast_draw_frames
            session.upload(vertices, faces, normals, uvs, model.materials);
            session.check(rast_set_lights(session.ctx(), reinterpret_cast<rast_light *>(lights.data()), (uint32_t)lights.size()), "rast_set_lights");
            const unsigned n = arguments.frames ? arguments.frames : 1u;
            const size_t P = (size_t)arguments.image_width * arguments.image_height;
            const unsigned chunk = n < 64u ? n : 64u;
            unsigned char *frames = static_cast<unsigned char *>(rast_host_alloc((uint64_t)chunk * 3 * P));
            if (!frames) { std::cerr << "out of pinned host memory" << std::endl; return 1; }
            std::vector<rast_args> poses(chunk);
            host::ApngWriter recording;
            if (!arguments.record.empty()) {
                err = recording.open(arguments.record, arguments.image_width, arguments.image_height, 3, n, arguments.record_delay_ms);
                if (!err.empty()) { std::cerr << err << std::endl; return 1; }
            }
            const auto t0 = std::chrono::steady_clock::now();
            for (unsigned first = 0; first < n; first += chunk) {
                const unsigned count = n - first < chunk ? n - first : chunk;
                for (unsigned i = 0; i < count; ++i) {
                    poses[i] = rast::to_rast_args(arguments);
                    poses[i].flat = flat_code;
                    poses[i].tait_bryan_angles[1] = rast_spin_angle(arguments.tait_bryan_angles[1], first + i, n);
                }
                session.check(rast_draw_frames(session.ctx(), poses.data(), count, frames, nullptr, 0), "rast_draw_frames");
                if (!arguments.record.empty()) {
                    for (unsigned i = 0; i < count; ++i) {
                        err = recording.add_frame_planar(frames + (size_t)i * 3 * P);
                        if (!err.empty()) { std::cerr << err << std::endl; return 1; }
                    }
                }
                if (!arguments.save_frames.empty()) {
                    for (unsigned i = 0; i < count; ++i) {
                        char name[4096];
                        std::snprintf(name, sizeof name, arguments.save_frames.c_str(), first + i);
                        err = host::png_write_planar(name, frames + (size_t)i * 3 * P, arguments.image_width, arguments.image_height, 3);
                        if (!err.empty()) { std::cerr << err << std::endl; return 1; }
                    }
                }
            }
            const std::chrono::duration<float> dt = std::chrono::steady_clock::now() - t0;
            if (verbose) std::cout << n << " frames in " << dt.count() << " s: " << std::to_string((float)n / dt.count()) << " frames/s" << std::endl; // renderer.cpp:120
            if (!arguments.record.empty()) {
                err = recording.close();
                if (!err.empty()) { std::cerr << err << std::endl; return 1; }
            }
            // the last frame is left in frame.png so the run has a visible result
            err = host::png_write_planar(arguments.frame_out, frames + (size_t)((n - 1) % chunk) * 3 * P, arguments.image_width, arguments.image_height, 3);
            rast_host_free(frames);
            if (!err.empty()) { std::cerr << err << std::endl; return 1; }
        }
    } catch (const std::exception &e) {
        std::cerr << "Error: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
