// Boundary proof against the reference's REAL headers.  TEST INFRASTRUCTURE ONLY, dev container only:
// built by tests/orc.py::build_shim_real_headers() from where the reference lies (/root/reference) into oracle/_ref/.
//
//   * includes the reference's own headers/drawing.h (-> light.h, arguments.h, face.h, material.h, vendored CImg.h) and
//     headers/fileloader.h; glm / text-csv come from the stand-ins in oracle/glm_stub (un-vendored Conan packages);
//   * material.h is the reference's file with the ONE line INTEGRATION.md section 3 asks a maintainer to add
//     (`friend struct RastMaterialView;`), applied to a temporary copy at build time -- nothing of it is committed;
//   * the scene is read by the reference's own load_obj / load_lights (fileloader.cpp, compiled unmodified), so the vectors
//     handed over are the reference's std::vector<glm::vec3>, std::vector<Triangle>, std::vector<Light>, std::vector<Material>
//     (with real CImg<float> textures), the buffers real CImg<unsigned char> / CImg<float>;
//   * the call is INTEGRATION.md's: rast::draw_frame(gpu, <the reference's nine arguments>).
//
// Prints FNV-1a-64 of the frame and depth planes (compared with tests/golden/cases.json by the GPU test), then edits the
// vertices in place and draws again: the second frame must differ (no stale scene).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "drawing.h"      // the reference's declaration of draw_frame and all its types
#include "fileloader.h"   // the reference's loaders

// --- added after class Material -------------------------------------------------------------
struct RastMaterialView {                      // what include/rast_draw_frame.hpp reads from a material
    float kd[3]; bool has_texture; int tex_w, tex_h; const float* texels;
    explicit RastMaterialView(const Material& m)
      : kd{m.diffuse_colour.x, m.diffuse_colour.y, m.diffuse_colour.z}, has_texture(m.has_texture),
        tex_w(m.diffuse_texture.width()), tex_h(m.diffuse_texture.height()), texels(m.diffuse_texture.data()) {}
};

#include "rast_draw_frame.hpp"

// arguments.cpp needs TCLAP (un-vendored): the constructor the header declares is defined here with arguments.cpp:15-33's defaults
Args::Args(int, char **)
    : image_width(540u), image_height(304u), aspect_ratio(540.f / 304.f), spin(false), flat(false), wind_clockwise(false), scale(1.f),
      displacement(0.f), tait_bryan_angles(0.f) {}

static uint64_t fnv(const void *p, size_t n) {
    uint64_t h = 14695981039346656037ull;
    const unsigned char *b = static_cast<const unsigned char *>(p);
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

int main(int argc, char **argv) {
    if (argc < 6) { std::fprintf(stderr, "usage: %s model.obj materials_dir/ lights.csv width height [opt-ins]\n", argv[0]); return 2; }
    Args arguments(0, nullptr);
    arguments.obj_file = argv[1];
    arguments.materials_directory = argv[2];
    arguments.lights_file = argv[3];
    arguments.image_width = (unsigned)std::atoi(argv[4]);
    arguments.image_height = (unsigned)std::atoi(argv[5]);
    arguments.aspect_ratio = float(arguments.image_width) / float(arguments.image_height); // arguments.cpp:39

    std::vector<glm::vec3> model_vertices, model_vertnormals;
    std::vector<glm::vec2> vertuvs;
    std::vector<Triangle> faces;
    std::vector<Material> materials;
    std::vector<Light> lights;
    load_obj(arguments, model_vertices, faces, model_vertnormals, vertuvs, materials); // renderer.cpp:80
    load_lights(arguments.lights_file, lights);                                         // renderer.cpp:81

    try {
        rast::Session gpu(0);
        const bool opt_ins = argc > 6; // any 6th argument: the two optional lines of INTEGRATION.md section 3
        if (opt_ins) { gpu.pin_outputs(true); gpu.retained_outputs(true); }
        std::vector<RastMaterialView> mats(materials.begin(), materials.end());
        cimg_library::CImg<unsigned char> frame_buffer(arguments.image_width, arguments.image_height, 1, 3, 0); // renderer.cpp:85
        cimg_library::CImg<float> depth_buffer(arguments.image_width, arguments.image_height, 1, 1, 1.f);       // renderer.cpp:86
        rast::draw_frame(gpu, model_vertices, faces, model_vertnormals, vertuvs, lights, mats, arguments, &frame_buffer, &depth_buffer);
        std::printf("RESULT %016llx %016llx %zu %zu %.9g\n", (unsigned long long)fnv(frame_buffer.data(), frame_buffer.size()),
                    (unsigned long long)fnv(depth_buffer.data(), depth_buffer.size() * sizeof(float)), faces.size(), lights.size(), (double)lights[0].trans_dir.x);
        for (glm::vec3 &v : model_vertices) v.x = v.x * 0.5f; // edited in place: same vector, same address
        frame_buffer.fill(0); depth_buffer.fill(1.f);
        rast::draw_frame(gpu, model_vertices, faces, model_vertnormals, vertuvs, lights, mats, arguments, &frame_buffer, &depth_buffer);
        std::printf("EDITED %016llx %016llx\n", (unsigned long long)fnv(frame_buffer.data(), frame_buffer.size()),
                    (unsigned long long)fnv(depth_buffer.data(), depth_buffer.size() * sizeof(float)));
        for (glm::vec3 &v : model_vertices) v.x = v.x * 2.0f; // back to the original (exact in binary32)
        if (!opt_ins) { frame_buffer.fill(0); depth_buffer.fill(1.f); } // with retained outputs the buffers keep the previous draw: the library resets what it must
        rast::draw_frame(gpu, model_vertices, faces, model_vertnormals, vertuvs, lights, mats, arguments, &frame_buffer, &depth_buffer);
        std::printf("AGAIN %016llx %016llx\n", (unsigned long long)fnv(frame_buffer.data(), frame_buffer.size()),
                    (unsigned long long)fnv(depth_buffer.data(), depth_buffer.size() * sizeof(float)));
    } catch (const std::exception &e) {
        std::fprintf(stderr, "shim_real_headers: %s\n", e.what());
        return 1;
    }
    return 0;
}
