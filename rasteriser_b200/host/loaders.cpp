#include "loaders.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>

#include "image.hpp"
#include "png.hpp"

namespace host {
namespace {

inline bool is_blank(char c) { return c == ' ' || c == '\t'; }
inline bool is_digit(char c) { return (unsigned)(c - '0') < 10u; }
inline bool is_eol(char c) { return c == '\r' || c == '\n' || c == '\0'; }

// One logical line: LF, CR and CRLF all end a line; a last line without terminator counts.
bool next_line(std::istream &in, std::string &line) {
    line.clear();
    std::streambuf *sb = in.rdbuf();
    bool got_any = false;
    for (;;) {
        const int c = sb->sbumpc();
        if (c == EOF) {
            if (!got_any) in.setstate(std::ios::eofbit);
            return got_any;
        }
        got_any = true;
        if (c == '\n') return true;
        if (c == '\r') {
            if (sb->sgetc() == '\n') sb->sbumpc();
            return true;
        }
        line.push_back((char)c);
    }
}

// Cursor over one line, with the token rules tinyobjloader applies.
struct Cursor {
    const char *p;
    void skip_blanks() { p += std::strspn(p, " \t"); }
    const char *token_end() const { return p + std::strcspn(p, " \t\r"); }
    float number(double dflt = 0.0) {
        skip_blanks();
        const char *e = token_end();
        const float f = parse_obj_float(p, e, dflt);
        p = e;
        return f;
    }
    std::string word() { // like sscanf("%s")
        p += std::strspn(p, " \t\r\n");
        const size_t n = std::strcspn(p, " \t\r\n");
        std::string s(p, n);
        p += n;
        return s;
    }
};

struct Corner { int v, vt, vn; };

// "make index zero-base, and also support relative index" (tiny_obj_loader.h:414-418)
inline int resolve_index(int idx, int count) { return idx > 0 ? idx - 1 : (idx == 0 ? 0 : count + idx); }

// i, i/j, i/j/k, i//k (tiny_obj_loader.h:680-711)
Corner parse_corner(Cursor &c, int nv, int nvn, int nvt) {
    Corner r = {-1, -1, -1};
    r.v = resolve_index(std::atoi(c.p), nv);
    c.p += std::strcspn(c.p, "/ \t\r");
    if (*c.p != '/') return r;
    ++c.p;
    if (*c.p == '/') { // i//k
        ++c.p;
        r.vn = resolve_index(std::atoi(c.p), nvn);
        c.p += std::strcspn(c.p, "/ \t\r");
        return r;
    }
    r.vt = resolve_index(std::atoi(c.p), nvt);
    c.p += std::strcspn(c.p, "/ \t\r");
    if (*c.p != '/') return r;
    ++c.p;
    r.vn = resolve_index(std::atoi(c.p), nvn);
    c.p += std::strcspn(c.p, "/ \t\r");
    return r;
}

struct MtlEntry { std::string name; float kd[3]; std::string map_kd; };

// map_Kd [options] filename (tiny_obj_loader.h:746-838): options are skipped with their arguments,
// the last bare token is the file name.
std::string texture_name(const char *s) {
    static const struct { const char *opt; int args; } kOptions[] = {
        {"-blendu", 1}, {"-blendv", 1}, {"-clamp", 1}, {"-boost", 1}, {"-bm", 1}, {"-o", 3}, {"-s", 3}, {"-t", 3}, {"-imfchan", 1}, {"-mm", 2}};
    std::string name;
    Cursor c{s};
    while (!is_eol(*c.p)) {
        bool matched = false;
        for (const auto &o : kOptions) {
            const size_t n = std::strlen(o.opt);
            if (!std::strncmp(c.p, o.opt, n) && is_blank(c.p[n])) {
                c.p += n + 1;
                for (int k = 0; k < o.args; ++k) { c.skip_blanks(); c.p = c.token_end(); }
                matched = true;
                break;
            }
        }
        if (!matched && !std::strncmp(c.p, "-type", 5) && is_blank(c.p[5])) {
            c.p += 5;
            c.skip_blanks();
            c.p = c.token_end();
            matched = true;
        }
        if (!matched) {
            c.skip_blanks();
            const char *e = c.token_end();
            name.assign(c.p, e);
            c.p = e;
            c.skip_blanks();
        }
    }
    return name;
}

// LoadMtl (tiny_obj_loader.h:954-1316), the fields the renderer uses: newmtl, Kd, map_Kd.
void parse_mtl(std::istream &in, std::vector<MtlEntry> &materials, std::map<std::string, int> &by_name) {
    MtlEntry cur{"", {0.f, 0.f, 0.f}, ""};
    std::string line;
    while (in.peek() != -1) {
        next_line(in, line);
        if (!line.empty()) line = line.substr(0, line.find_last_not_of(" \t") + 1);
        if (line.empty()) continue;
        Cursor c{line.c_str()};
        c.skip_blanks();
        if (*c.p == '\0' || *c.p == '#') continue;
        if (!std::strncmp(c.p, "newmtl", 6) && is_blank(c.p[6])) {
            if (!cur.name.empty()) { // flush previous material
                by_name.insert(std::make_pair(cur.name, (int)materials.size()));
                materials.push_back(cur);
            }
            cur = MtlEntry{"", {0.f, 0.f, 0.f}, ""};
            c.p += 7;
            cur.name = c.word();
        } else if (c.p[0] == 'K' && c.p[1] == 'd' && is_blank(c.p[2])) {
            c.p += 2;
            cur.kd[0] = c.number(); cur.kd[1] = c.number(); cur.kd[2] = c.number();
        } else if (!std::strncmp(c.p, "map_Kd", 6) && is_blank(c.p[6])) {
            const std::string n = texture_name(c.p + 7);
            if (!n.empty()) cur.map_kd = n;
        }
    }
    // the last material is flushed even when unnamed (tiny_obj_loader.h:1311-1314)
    by_name.insert(std::make_pair(cur.name, (int)materials.size()));
    materials.push_back(cur);
}

} // namespace

// tryParseDouble + parseFloat (tiny_obj_loader.h:463-586): sign, integer digits accumulated as
// mantissa*10 + d, decimals as d * 10^-k (table for the first 7, pow beyond), optional exponent applied
// as ldexp(mantissa * 5^e, e); the double is then narrowed to float.  A malformed token yields the default.
float parse_obj_float(const char *s, const char *end, double dflt) {
    static const double kNegPow10[] = {1.0, 0.1, 0.01, 0.001, 0.0001, 0.00001, 0.000001, 0.0000001};
    double value = dflt;
    do {
        if (s >= end) break;
        const char *c = s;
        char sign = '+';
        if (*c == '+' || *c == '-') sign = *c++;
        else if (!is_digit(*c)) break;

        double mantissa = 0.0;
        int n_read = 0;
        while (c != end && is_digit(*c)) { mantissa *= 10; mantissa += (int)(*c - '0'); ++c; ++n_read; }
        if (n_read == 0) break;

        int exponent = 0;
        bool ok = true;
        if (c != end) {
            if (*c == '.') {
                ++c;
                n_read = 1;
                while (c != end && is_digit(*c)) {
                    mantissa += (int)(*c - '0') * (n_read < 8 ? kNegPow10[n_read] : std::pow(10.0, -n_read));
                    ++n_read;
                    ++c;
                }
            } else if (*c != 'e' && *c != 'E') {
                c = end; // anything else ends the number
            }
            if (c != end && (*c == 'e' || *c == 'E')) {
                ++c;
                char esign = '+';
                if (c != end && (*c == '+' || *c == '-')) esign = *c++;
                else if (!is_digit(*c)) ok = false;
                if (ok) {
                    n_read = 0;
                    while (c != end && is_digit(*c)) { exponent *= 10; exponent += (int)(*c - '0'); ++c; ++n_read; }
                    exponent *= (esign == '+' ? 1 : -1);
                    if (n_read == 0) ok = false;
                }
            }
        }
        if (!ok) break;
        value = (sign == '+' ? 1 : -1) * (exponent ? std::ldexp(mantissa * std::pow(5.0, exponent), exponent) : mantissa);
    } while (false);
    return (float)value;
}

bool load_texture(const std::string &path, MaterialData &m, std::string &error, bool verbose) {
    PngImage png;
    const std::string e = png_read(path, png);
    if (!e.empty()) { error = e; return false; }
    // CImg<float>(file): one float plane per channel holding the 8-bit values; grey images are expanded so
    // that channels 0..2 exist (the reference samples channels 0,1,2: material.cpp:19-21)
    Image<float> img(png.width, png.height, 3, 0.f);
    for (unsigned y = 0; y < png.height; ++y)
        for (unsigned x = 0; x < png.width; ++x)
            for (unsigned c = 0; c < 3; ++c) {
                const unsigned src = png.channels >= 3 ? c : 0;
                img(x, y, c) = (float)png.pixels[((size_t)y * png.width + x) * png.channels + src];
            }
    if (verbose) std::cout << "Loaded texture " << path << "." << std::endl; // material.h:21
    img.normalize(0.f, 1.f);                                                  // material.h:22
    m.has_texture = true;
    m.texture_file = path;
    m.tex_w = (int)png.width;
    m.tex_h = (int)png.height;
    m.texels.assign(img.data(), img.data() + img.size());
    return true;
}

bool load_obj(const std::string &obj_file, const std::string &materials_directory, Model &model, std::string &error, bool verbose) {
    std::ifstream in(obj_file.c_str());
    if (!in) { error = "Cannot open file [" + obj_file + "]\n"; return false; }

    std::vector<MtlEntry> mtl;
    std::map<std::string, int> mtl_by_name;
    std::vector<std::vector<Corner>> pending;   // faces since the last flush ("faceGroup")
    std::vector<int32_t> shape;                 // triangles of the shape being built
    std::vector<std::vector<int32_t>> shapes;   // finished shapes, in file order
    int material = -1;
    std::ostringstream warn;

    // exportFaceGroupToShape (tiny_obj_loader.h:879-940): fan (f0, f[k-1], f[k]), one material id per triangle
    auto flush_faces = [&]() -> bool {
        if (pending.empty()) return false;
        for (const std::vector<Corner> &face : pending) {
            if (face.size() < 2) continue;
            const Corner c0 = face[0];
            Corner c2 = face[1];
            for (size_t k = 2; k < face.size(); ++k) {
                const Corner c1 = c2;
                c2 = face[k];
                const int32_t t[10] = {c0.v, c1.v, c2.v, c0.vn, c1.vn, c2.vn, c0.vt, c1.vt, c2.vt, material};
                shape.insert(shape.end(), t, t + 10);
            }
        }
        return true;
    };

    std::string line;
    while (in.peek() != -1) {
        next_line(in, line);
        if (line.empty()) continue;
        Cursor c{line.c_str()};
        c.skip_blanks();
        if (*c.p == '\0' || *c.p == '#') continue;
        if (c.p[0] == 'v' && is_blank(c.p[1])) {
            c.p += 2;
            for (int k = 0; k < 3; ++k) model.positions.push_back(c.number());
        } else if (c.p[0] == 'v' && c.p[1] == 'n' && is_blank(c.p[2])) {
            c.p += 3;
            for (int k = 0; k < 3; ++k) model.normals.push_back(c.number());
        } else if (c.p[0] == 'v' && c.p[1] == 't' && is_blank(c.p[2])) {
            c.p += 3;
            for (int k = 0; k < 2; ++k) model.uvs.push_back(c.number());
        } else if (c.p[0] == 'f' && is_blank(c.p[1])) {
            c.p += 2;
            c.skip_blanks();
            std::vector<Corner> face;
            while (!is_eol(*c.p)) {
                face.push_back(parse_corner(c, (int)(model.positions.size() / 3), (int)(model.normals.size() / 3), (int)(model.uvs.size() / 2)));
                c.p += std::strspn(c.p, " \t\r");
            }
            pending.push_back(face);
        } else if (!std::strncmp(c.p, "usemtl", 6) && is_blank(c.p[6])) {
            c.p += 7;
            const std::string name = c.word();
            const auto it = mtl_by_name.find(name);
            const int id = it != mtl_by_name.end() ? it->second : -1;
            if (id != material) { // per-face materials: flush into the current shape, keep the shape open
                flush_faces();
                pending.clear();
                material = id;
            }
        } else if (!std::strncmp(c.p, "mtllib", 6) && is_blank(c.p[6])) {
            std::istringstream names(std::string(c.p + 7));
            std::string fn;
            bool any = false, found = false;
            while (std::getline(names, fn, ' ')) {
                any = true;
                const std::string path = materials_directory.empty() ? fn : materials_directory + fn; // plain concatenation
                std::ifstream mf(path.c_str());
                if (!mf) { warn << "WARN: Material file [ " << path << " ] not found." << std::endl; continue; }
                parse_mtl(mf, mtl, mtl_by_name);
                found = true;
                break;
            }
            if (!any) warn << "WARN: Looks like empty filename for mtllib. Use default material. \n";
            else if (!found) warn << "WARN: Failed to load material file(s). Use default material.\n";
        } else if ((c.p[0] == 'g' || c.p[0] == 'o') && is_blank(c.p[1])) {
            // a new group / object closes the shape; as in tinyobjloader 1.0.5 the shape is kept only if faces
            // were still pending at this point
            if (flush_faces()) shapes.push_back(shape);
            shape.clear();
            pending.clear();
        }
    }
    const bool flushed = flush_faces();
    if (flushed || !shape.empty()) shapes.push_back(shape);

    error = warn.str();
    // load_materials (fileloader.cpp:47-58)
    for (const MtlEntry &e : mtl) {
        MaterialData m;
        m.kd[0] = e.kd[0]; m.kd[1] = e.kd[1]; m.kd[2] = e.kd[2];
        if (!e.map_kd.empty()) {
            std::string terr;
            if (!load_texture(materials_directory + e.map_kd, m, terr, verbose)) { error += terr; return false; }
        }
        model.materials.push_back(m);
    }
    // load_triangles per shape (fileloader.cpp:60-77,116-118)
    for (const std::vector<int32_t> &s : shapes) {
        if (verbose) std::cout << "Loading " << s.size() / 10 << " triangles..." << std::endl;
        model.tris.insert(model.tris.end(), s.begin(), s.end());
    }
    if (verbose) std::cout << "Loaded model " << obj_file << "." << std::endl;
    return true;
}

// The reference reads with text-csv's csv_istream: `while (csv_stream) { csv >> 7 floats; push_back }`
// (fileloader.cpp:129-132).  Defined here as: every non-blank line is one light of 7 comma-separated
// numbers; blank lines and a missing final newline are ignored.
bool load_lights(const std::string &file, std::vector<Light> &lights, std::string &error) {
    std::ifstream in(file.c_str());
    if (!in) { error = "Cannot open lights file [" + file + "]"; return false; }
    std::string line;
    while (in.peek() != -1) {
        next_line(in, line);
        if (line.find_first_not_of(" \t") == std::string::npos) continue;
        float v[7] = {0, 0, 0, 0, 0, 0, 0};
        std::istringstream fields(line);
        std::string field;
        for (int k = 0; k < 7 && std::getline(fields, field, ','); ++k) v[k] = std::strtof(field.c_str(), nullptr);
        Light l;
        l.direction[0] = v[0]; l.direction[1] = v[1]; l.direction[2] = v[2];
        l.intensity = v[3];
        l.colour[0] = v[4]; l.colour[1] = v[5]; l.colour[2] = v[6];
        l.trans_dir[0] = l.trans_dir[1] = l.trans_dir[2] = 0.f;
        lights.push_back(l);
    }
    return true;
}

void add_square(Model &model) {
    const float pos[] = {-0.5f, -0.5f, 0.f, 0.5f, -0.5f, 0.f, -0.5f, 0.5f, 0.f, 0.5f, 0.5f, 0.f};
    const float nrm[] = {0.f, 0.f, 1.f};
    const float uv[] = {0.f, 0.f, 1.f, 0.f, 0.f, 1.f, 1.f, 1.f};
    const int32_t tris[] = {0, 1, 2, 0, 0, 0, 0, 1, 2, 0, 2, 1, 3, 0, 0, 0, 2, 1, 3, 0};
    model.positions.assign(pos, pos + 12);
    model.normals.assign(nrm, nrm + 3);
    model.uvs.assign(uv, uv + 8);
    model.tris.assign(tris, tris + 20);
    MaterialData white;
    white.kd[0] = white.kd[1] = white.kd[2] = 1.f;
    model.materials.assign(1, white);
}

} // namespace host
