// Compile-and-run check of include/rast_load_obj.hpp with stand-ins for the reference's types (glm::vec3 / vec2, Triangle,
// Material, Args): prints counts and an FNV-1a-64 of the arrays for the Python test to compare with the golden scenes.
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../include/rast_load_obj.hpp"

struct vec3 { float x, y, z; vec3() : x(0), y(0), z(0) {} vec3(float a, float b, float c) : x(a), y(b), z(c) {} };
struct vec2 { float x, y; };
struct Triangle { int vertices[3], normals[3], uvs[3], material; }; // headers/face.h:6-13
struct Material {                                                   // headers/material.h:11-25 (constructors only)
    vec3 kd; bool has_texture; std::string file;
    Material(const vec3 &dc) : kd(dc), has_texture(false) {}
    Material(const vec3 &dc, const std::string &dtf) : kd(dc), has_texture(true), file(dtf) { std::cout << "Loaded texture " << file << "." << std::endl; }
};
struct Args { std::string obj_file, materials_directory; };

static uint64_t fnv(const void *p, size_t n, uint64_t h = 1469598103934665603ull) {
    const unsigned char *b = static_cast<const unsigned char *>(p);
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

int main(int argc, char **argv) {
    Args a{argv[1], argc > 2 ? argv[2] : ""};
    std::vector<vec3> v, vn;
    std::vector<vec2> vt;
    std::vector<Triangle> t;
    std::vector<Material> m;
    rast::load_obj(a, v, t, vn, vt, m, 3);
    std::printf("RESULT %zu %zu %zu %zu %zu %016llx %016llx %s\n", v.size(), vn.size(), vt.size(), t.size(), m.size(),
                (unsigned long long)fnv(v.data(), v.size() * 12), (unsigned long long)fnv(t.data(), t.size() * 40), m.empty() ? "-" : m[0].file.c_str());
    return 0;
}
