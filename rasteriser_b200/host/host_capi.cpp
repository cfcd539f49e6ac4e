// host_capi.cpp -- C entry points over the host-side loaders / PNG codec / flag parser, so the Python
// tests (ctypes) can check them against the fixtures the reference's own loader produced.  Built into
// librast_host.so; not needed by the renderer executable.
#include <cstring>
#include <string>
#include <vector>

#include "args.hpp"
#include "loaders.hpp"
#include "png.hpp"

extern "C" {

void *rasth_load_obj(const char *obj_file, const char *mats_dir, char *err, int err_cap) {
    host::Model *m = new host::Model();
    std::string e;
    const bool ok = host::load_obj(obj_file, mats_dir ? mats_dir : "", *m, e, false);
    if (err && err_cap > 0) { std::strncpy(err, e.c_str(), (size_t)err_cap - 1); err[err_cap - 1] = 0; }
    if (!ok) { delete m; return nullptr; }
    return m;
}
// the same with an explicit thread count; stats = {threads, file_bytes, read_s, scan_s, resolve_s, parse_s, total_s}
void *rasth_load_obj_mt(const char *obj_file, const char *mats_dir, unsigned threads, double stats[7], char *err, int err_cap) {
    host::Model *m = new host::Model();
    std::string e;
    host::LoadStats st;
    const bool ok = host::load_obj(obj_file, mats_dir ? mats_dir : "", *m, e, false, threads, &st);
    if (err && err_cap > 0) { std::strncpy(err, e.c_str(), (size_t)err_cap - 1); err[err_cap - 1] = 0; }
    if (stats) { stats[0] = st.threads; stats[1] = (double)st.file_bytes; stats[2] = st.read_s; stats[3] = st.scan_s; stats[4] = st.resolve_s; stats[5] = st.parse_s; stats[6] = st.total_s; }
    if (!ok) { delete m; return nullptr; }
    return m;
}
void rasth_set_obj_piece_bytes(uint64_t bytes) { host::set_obj_piece_bytes((size_t)bytes); }
int rasth_save_mesh_cache(void *h, const char *file) {
    std::string e;
    return host::save_mesh_cache(file, *static_cast<host::Model *>(h), e) ? 0 : -1;
}
void *rasth_load_mesh_cache(const char *file) {
    host::Model *m = new host::Model();
    std::string e;
    if (!host::load_mesh_cache(file, *m, e, false)) { delete m; return nullptr; }
    return m;
}
void rasth_model_free(void *h) { delete static_cast<host::Model *>(h); }
void rasth_model_sizes(void *h, uint64_t out[5]) {
    host::Model *m = static_cast<host::Model *>(h);
    out[0] = m->positions.size() / 3; out[1] = m->normals.size() / 3; out[2] = m->uvs.size() / 2; out[3] = m->n_tris(); out[4] = m->materials.size();
}
void rasth_model_copy(void *h, float *pos, float *nrm, float *uv, int32_t *tris) {
    host::Model *m = static_cast<host::Model *>(h);
    if (!m->positions.empty()) std::memcpy(pos, m->positions.data(), m->positions.size() * 4);
    if (!m->normals.empty()) std::memcpy(nrm, m->normals.data(), m->normals.size() * 4);
    if (!m->uvs.empty()) std::memcpy(uv, m->uvs.data(), m->uvs.size() * 4);
    if (!m->tris.empty()) std::memcpy(tris, m->tris.data(), m->tris.size() * 4);
}
// material i: kd[3], has_texture, tex_w, tex_h; texels copied if out_texels != NULL
void rasth_model_material(void *h, uint32_t i, float kd[3], int32_t info[3], float *out_texels) {
    const host::MaterialData &md = static_cast<host::Model *>(h)->materials[i];
    kd[0] = md.kd[0]; kd[1] = md.kd[1]; kd[2] = md.kd[2];
    info[0] = md.has_texture ? 1 : 0; info[1] = md.tex_w; info[2] = md.tex_h;
    if (out_texels && md.has_texture) std::memcpy(out_texels, md.texels.data(), md.texels.size() * 4);
}
int rasth_load_lights(const char *file, float *out7, int capacity) {
    std::vector<host::Light> l;
    std::string e;
    if (!host::load_lights(file, l, e)) return -1;
    for (int i = 0; i < (int)l.size() && i < capacity; ++i) {
        std::memcpy(out7 + 7 * i, l[i].direction, 12);
        out7[7 * i + 3] = l[i].intensity;
        std::memcpy(out7 + 7 * i + 4, l[i].colour, 12);
    }
    return (int)l.size();
}
float rasth_parse_float(const char *s) { return host::parse_obj_float(s, s + std::strlen(s)); }
void rasth_png_set_threads(unsigned threads) { host::png_set_threads(threads); }
int rasth_png_write(const char *path, const uint8_t *planar, uint32_t w, uint32_t h, uint32_t c) { return host::png_write_planar(path, planar, w, h, c).empty() ? 0 : -1; }
// frames: n planar [c][h][w] images back to back -> one animated PNG
int rasth_apng_write(const char *path, const uint8_t *frames, uint32_t n, uint32_t w, uint32_t h, uint32_t c, uint32_t delay_ms) {
    host::ApngWriter a;
    if (!a.open(path, w, h, c, n, delay_ms).empty()) return -1;
    for (uint32_t i = 0; i < n; ++i)
        if (!a.add_frame_planar(frames + (size_t)i * c * w * h).empty()) return -1;
    return a.close().empty() ? 0 : -1;
}
int rasth_png_read(const char *path, uint32_t dims[3], uint8_t *out, uint64_t cap) {
    host::PngImage img;
    if (!host::png_read(path, img).empty()) return -1;
    dims[0] = img.width; dims[1] = img.height; dims[2] = img.channels;
    if (out && cap >= img.pixels.size()) std::memcpy(out, img.pixels.data(), img.pixels.size());
    return 0;
}
// returns 0 ok, 1 help, 2 version, 3 error; numeric fields: w,h,spin,flat,cw,frames ; floats: aspect,scale,dx,dy,dz,rx,ry,rz
int rasth_parse_args(int argc, const char *const *argv, uint32_t u[6], float f[8], char *strings, int cap) {
    host::Args a;
    std::string msg;
    const host::ParseResult r = host::parse_args(argc, argv, a, msg);
    u[0] = a.image_width; u[1] = a.image_height; u[2] = a.spin; u[3] = a.flat; u[4] = a.wind_clockwise; u[5] = a.frames;
    f[0] = a.aspect_ratio; f[1] = a.scale;
    for (int k = 0; k < 3; ++k) { f[2 + k] = a.displacement[k]; f[5 + k] = a.tait_bryan_angles[k]; }
    const std::string s = a.obj_file + "\n" + a.lights_file + "\n" + a.materials_directory + "\n" + msg;
    if (strings && cap > 0) { std::strncpy(strings, s.c_str(), (size_t)cap - 1); strings[cap - 1] = 0; }
    return (int)r;
}

} // extern "C"
