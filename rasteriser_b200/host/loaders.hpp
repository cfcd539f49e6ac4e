// loaders.hpp -- model, material and light loading of the host program.
//
// Mirrors fileloader.cpp:79-133 of the reference (load_obj, load_lights), which wraps tinyobjloader
// 1.0.5 and text-csv.  The arrays produced are the flat equivalents of the reference's vectors:
// positions xyz, normals xyz, uvs uv, and 10 x int32 per triangle in struct Triangle's order
// (headers/face.h:6-13).  OBJ parsing restates tinyobjloader's documented behaviour -- its own float
// reader (not strtod), 1-based / relative indices, "-1 = absent", triangle-fan triangulation,
// per-face material from usemtl, shapes split on o / g -- so that the vertex bits and the triangle
// order equal what the reference would feed draw_frame.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

namespace host {

struct MaterialData {       // class Material (headers/material.h:11-25)
    float kd[3] = {0.f, 0.f, 0.f};
    bool has_texture = false;
    std::string texture_file;   // materials_directory + map_Kd name (fileloader.cpp:55)
    int tex_w = 0, tex_h = 0;
    std::vector<float> texels;  // planar [3][h][w], normalised by normalize(0,1) (material.h:22)
};

struct Light {              // struct Light (headers/light.h:7-14)
    float direction[3];
    float intensity;
    float colour[3];
    float trans_dir[3];
};

struct Model {
    std::vector<float> positions, normals, uvs;
    std::vector<int32_t> tris; // 10 per triangle
    std::vector<MaterialData> materials;
    size_t n_tris() const { return tris.size() / 10; }
};

// load_obj (fileloader.cpp:79-121).  Prints the reference's progress lines to stdout.  Returns false with
// `error` set when the .obj cannot be read; warnings (missing .mtl) go to `error` with a true return.
bool load_obj(const std::string &obj_file, const std::string &materials_directory, Model &model, std::string &error, bool verbose = true);

// load_lights (fileloader.cpp:123-133): rows of direction_x,dir_y,dir_z,intensity,red,green,blue.
bool load_lights(const std::string &file, std::vector<Light> &lights, std::string &error);

// add_square (renderer.cpp:32-50): the scene used when no -o is given.
void add_square(Model &model);

// Material(dc, path) (material.h:20-23): load an image, convert to float planes, normalize(0,1).
bool load_texture(const std::string &path, MaterialData &m, std::string &error, bool verbose = true);

// tinyobjloader's float reader (tiny_obj_loader.h:463-586): exposed for tests.
float parse_obj_float(const char *begin, const char *end, double default_value = 0.0);

} // namespace host
