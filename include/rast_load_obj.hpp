// rast_load_obj.hpp -- the reference's load_obj (fileloader.cpp:79-121, declared in headers/fileloader.h:16) on the
// parallel OBJ reader of rasteriser_b200/host/loaders.cpp.
//
//   void load_obj(const Args& arguments, std::vector<glm::vec3>& vertices, std::vector<Triangle>& triangles,
//                 std::vector<glm::vec3>& vertnormals, std::vector<glm::vec2>& vertuvs, std::vector<Material>& materials);
//
// Same arguments, same resulting vectors (bit-identical floats, same triangle order and material ids: the reader
// restates tinyobjloader 1.0.5, tests/test_host_loaders.py compares it with the reference's loader), same stdout /
// stderr lines and exit(1) on an unreadable file.  Materials are built with the reference's own constructors
// (headers/material.h:18-23), so textures still load through the reference's image code.  The types are template
// parameters so that this header needs neither glm nor CImg; link rasteriser_b200/host/loaders.cpp and png.cpp (-lz).
#pragma once

#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <type_traits>
#include <vector>

#include "../rasteriser_b200/host/loaders.hpp"

namespace rast {

template <class Args, class Vec3, class Triangle, class Vec2, class Material>
void load_obj(const Args &arguments, std::vector<Vec3> &vertices, std::vector<Triangle> &triangles, std::vector<Vec3> &vertnormals,
              std::vector<Vec2> &vertuvs, std::vector<Material> &materials, unsigned threads = 0) {
    static_assert(sizeof(Vec3) == 12 && sizeof(Vec2) == 8, "vec3 / vec2 must be packed floats");
    static_assert(sizeof(Triangle) == 40, "Triangle must be 10 x int32 (headers/face.h:6-13)");
    host::Model model;
    std::string err;
    const bool ok = host::load_obj(arguments.obj_file, arguments.materials_directory, model, err, /*verbose=*/false, threads, nullptr,
                                   /*decode_textures=*/false);
    if (!err.empty()) std::cerr << err << std::endl; // fileloader.cpp:95-97
    if (!ok) std::exit(1);                           // fileloader.cpp:98-100
    auto append = [](auto &dst, const auto &flat, size_t per_element) { // the reference appends to whatever the vectors hold
        typedef typename std::remove_reference<decltype(dst)>::type::value_type T;
        const T *first = reinterpret_cast<const T *>(flat.data());
        dst.insert(dst.end(), first, first + flat.size() / per_element);
    };
    append(vertices, model.positions, 3);   // components_to_vec3s (fileloader.cpp:103)
    append(vertnormals, model.normals, 3);  // :106
    append(vertuvs, model.uvs, 2);          // :109
    for (const host::MaterialData &m : model.materials) { // load_materials (fileloader.cpp:47-58)
        const Vec3 kd(m.kd[0], m.kd[1], m.kd[2]);
        if (!m.has_texture) materials.push_back(Material(kd));
        else materials.push_back(Material(kd, m.texture_file));
    }
    for (uint64_t n : model.shape_triangles) std::cout << "Loading " << n << " triangles..." << std::endl; // load_triangles (:68)
    append(triangles, model.tris, 10);
    std::cout << "Loaded model " << arguments.obj_file << "." << std::endl; // :120
}

} // namespace rast
