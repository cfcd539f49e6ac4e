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
#include <cstdlib>
#include <cstring>
#include <new>
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

// Growable array of a trivially-copyable type whose resize leaves new elements UNINITIALISED: the parallel
// OBJ loader sizes the output once and lets each worker thread first-touch the part it writes (a
// std::vector would zero gigabytes on one thread first).
template <class T> class Pod {
public:
    Pod() = default;
    Pod(const Pod &) = delete;
    Pod &operator=(const Pod &) = delete;
    Pod(Pod &&o) noexcept : p_(o.p_), n_(o.n_), cap_(o.cap_) { o.p_ = nullptr; o.n_ = o.cap_ = 0; }
    Pod &operator=(Pod &&o) noexcept { if (this != &o) { std::free(p_); p_ = o.p_; n_ = o.n_; cap_ = o.cap_; o.p_ = nullptr; o.n_ = o.cap_ = 0; } return *this; }
    ~Pod() { std::free(p_); }
    T *data() { return p_; }
    const T *data() const { return p_; }
    size_t size() const { return n_; }
    bool empty() const { return n_ == 0; }
    T &operator[](size_t i) { return p_[i]; }
    const T &operator[](size_t i) const { return p_[i]; }
    T *begin() { return p_; }
    T *end() { return p_ + n_; }
    const T *begin() const { return p_; }
    const T *end() const { return p_ + n_; }
    void clear() { n_ = 0; }
    void resize_uninitialized(size_t n) {
        if (n > cap_) {
            T *q = static_cast<T *>(std::realloc(p_, n * sizeof(T)));
            if (!q) throw std::bad_alloc();
            p_ = q; cap_ = n;
        }
        n_ = n;
    }
    void assign(const T *first, const T *last) {
        resize_uninitialized((size_t)(last - first));
        if (n_) std::memcpy(p_, first, n_ * sizeof(T));
    }
private:
    T *p_ = nullptr;
    size_t n_ = 0, cap_ = 0;
};

struct Model {
    Pod<float> positions, normals, uvs;
    Pod<int32_t> tris; // 10 per triangle
    std::vector<MaterialData> materials;
    std::vector<uint64_t> shape_triangles; // triangles of each exported shape, in file order (the "Loading N triangles..." lines)
    size_t n_tris() const { return tris.size() / 10; }
};

struct LoadStats {          // filled by load_obj when non-null (tools/bench_loader, tests)
    unsigned threads = 0;
    size_t file_bytes = 0;
    double read_s = 0, scan_s = 0, resolve_s = 0, parse_s = 0, total_s = 0;
};

// load_obj (fileloader.cpp:79-121).  Prints the reference's progress lines to stdout.  Returns false with
// `error` set when the .obj cannot be read; warnings (missing .mtl) go to `error` with a true return.
// The file is read into memory once and parsed by `threads` workers (0 = all hardware threads) in two passes:
// count, then parse straight into the final arrays.  The result is independent of the thread count.
// decode_textures = false leaves MaterialData::texels empty (texture_file, has_texture are still set): for callers that
// construct the reference's own Material objects from the file names (include/rast_load_obj.hpp).
bool load_obj(const std::string &obj_file, const std::string &materials_directory, Model &model, std::string &error, bool verbose = true,
              unsigned threads = 0, LoadStats *stats = nullptr, bool decode_textures = true);

void set_obj_piece_bytes(size_t bytes); // test hook: 0 = automatic

// Binary mesh cache (host extension, opt-in with --mesh-cache): the arrays load_obj produced, written verbatim
// with a small header, so a 50 M-triangle scene is read back at memory-copy speed instead of re-parsed.
// Textures are not cached (they are re-read from materials_directory through the stored names).
bool save_mesh_cache(const std::string &file, const Model &model, std::string &error);
bool load_mesh_cache(const std::string &file, Model &model, std::string &error, bool verbose = true);

// load_lights (fileloader.cpp:123-133): rows of direction_x,dir_y,dir_z,intensity,red,green,blue.
bool load_lights(const std::string &file, std::vector<Light> &lights, std::string &error);

// add_square (renderer.cpp:32-50): the scene used when no -o is given.
void add_square(Model &model);

// Material(dc, path) (material.h:20-23): load an image, convert to float planes, normalize(0,1).
bool load_texture(const std::string &path, MaterialData &m, std::string &error, bool verbose = true);

// tinyobjloader's float reader (tiny_obj_loader.h:463-586): exposed for tests.
float parse_obj_float(const char *begin, const char *end, double default_value = 0.0);

} // namespace host
