#include "loaders.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <thread>

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
    // beyond the literal table tinyobjloader calls pow(10.0, -n) for every digit; the same calls, made once
    constexpr int kPowTable = 64;
    static const struct Pow { double v[kPowTable]; Pow() { for (int i = 0; i < kPowTable; ++i) v[i] = std::pow(10.0, -i); } } kPow;
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
                    mantissa += (int)(*c - '0') * (n_read < 8 ? kNegPow10[n_read] : (n_read < kPowTable ? kPow.v[n_read] : std::pow(10.0, -n_read)));
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

// ---- parallel OBJ reader ----------------------------------------------------------------------------------
// LoadObj + exportFaceGroupToShape (tiny_obj_loader.h:1384-1650, :879-940) followed by load_obj's flattening of
// all shapes (fileloader.cpp:103-118), restated for files of tens of millions of triangles: the file is held in
// memory, cut into pieces at line boundaries, and every piece is visited twice by a pool of threads -- once to
// count, once to parse straight into the final arrays.  What tinyobjloader does with state carried from line to
// line (current material, faces pending since the last flush, the shape being built, the v / vn / vt counts
// relative indices refer to) is resolved between the passes from per-piece counts and the few structural lines
// (usemtl, mtllib, g, o), so the arrays equal the sequential reading whatever the thread count.
namespace {

struct Line { const char *b, *e; }; // [b, e): one logical line without its terminator, cut at an embedded NUL

// next logical line at or after p (LF, CR and CRLF end a line); returns the start of the following line
inline const char *scan_line(const char *p, const char *end, Line &ln) {
    const char *q = p, *nul = nullptr;
    while (q < end && *q != '\n' && *q != '\r') {
        if (*q == '\0' && !nul) nul = q;
        ++q;
    }
    ln.b = p;
    ln.e = nul ? nul : q;
    if (q < end) q += (*q == '\r' && q + 1 < end && q[1] == '\n') ? 2 : 1;
    return q;
}

enum class Kind { Skip, V, VN, VT, F, UseMtl, MtlLib, Group };

inline char at(const char *p, const char *e, int i) { return p + i < e ? p[i] : '\0'; }

// the dispatch of LoadObj's line loop; `body` is where the payload starts
inline Kind classify(const Line &ln, const char *&body) {
    const char *p = ln.b, *e = ln.e;
    while (p < e && is_blank(*p)) ++p;
    if (p >= e || *p == '#') return Kind::Skip;
    const char c0 = p[0], c1 = at(p, e, 1), c2 = at(p, e, 2);
    if (c0 == 'v') {
        if (is_blank(c1)) { body = p + 2; return Kind::V; }
        if (c1 == 'n' && is_blank(c2)) { body = p + 3; return Kind::VN; }
        if (c1 == 't' && is_blank(c2)) { body = p + 3; return Kind::VT; }
        return Kind::Skip;
    }
    if (c0 == 'f' && is_blank(c1)) { body = p + 2; return Kind::F; }
    if (e - p >= 7 && is_blank(p[6])) {
        if (!std::memcmp(p, "usemtl", 6)) { body = p + 7; return Kind::UseMtl; }
        if (!std::memcmp(p, "mtllib", 6)) { body = p + 7; return Kind::MtlLib; }
    }
    if ((c0 == 'g' || c0 == 'o') && is_blank(c1)) return Kind::Group;
    return Kind::Skip;
}

inline float number(const char *&p, const char *e) { // parseFloat on the next blank-delimited token
    while (p < e && is_blank(*p)) ++p;
    const char *t = p;
    while (t < e && !is_blank(*t) && *t != '\r') ++t;
    const float f = parse_obj_float(p, t, 0.0);
    p = t;
    return f;
}

// atoi on [p, e): leading isspace characters, an optional sign, decimal digits
inline int bounded_atoi(const char *p, const char *e) {
    while (p < e && (*p == ' ' || (*p >= '\t' && *p <= '\r'))) ++p;
    bool neg = false;
    if (p < e && (*p == '+' || *p == '-')) neg = *p++ == '-';
    unsigned v = 0;
    while (p < e && is_digit(*p)) v = v * 10u + (unsigned)(*p++ - '0');
    return neg ? (int)(0u - v) : (int)v;
}
inline const char *corner_stop(const char *p, const char *e) { // strcspn(p, "/ \t\r")
    while (p < e && *p != '/' && !is_blank(*p) && *p != '\r') ++p;
    return p;
}

// i, i/j, i/j/k, i//k (tiny_obj_loader.h:680-711).  PARSE = false only walks the token the same way (pass 1
// needs the corner count, and the walk does not depend on the numbers).
template <bool PARSE> inline Corner parse_corner(const char *&p, const char *e, int nv, int nvn, int nvt) {
    Corner r = {-1, -1, -1};
    if (PARSE) r.v = resolve_index(bounded_atoi(p, e), nv);
    p = corner_stop(p, e);
    if (p >= e || *p != '/') return r;
    ++p;
    if (p < e && *p == '/') { // i//k
        ++p;
        if (PARSE) r.vn = resolve_index(bounded_atoi(p, e), nvn);
        p = corner_stop(p, e);
        return r;
    }
    if (PARSE) r.vt = resolve_index(bounded_atoi(p, e), nvt);
    p = corner_stop(p, e);
    if (p >= e || *p != '/') return r;
    ++p;
    if (PARSE) r.vn = resolve_index(bounded_atoi(p, e), nvn);
    p = corner_stop(p, e);
    return r;
}
inline void skip_corner_gap(const char *&p, const char *e) { // strspn(p, " \t\r")
    while (p < e && (is_blank(*p) || *p == '\r')) ++p;
}

// A run of lines between two structural lines (or piece boundaries): its faces share one material and one fate.
struct Segment {
    uint64_t faces = 0, tris = 0; // 'f' lines, triangles they fan into
    int material = -1;
    bool keep = false;
    uint64_t tri_offset = 0;      // where its triangles go in Model::tris (triangles, not ints)
};
struct Event { Kind kind; Line line; const char *body; };
struct Piece {
    const char *b, *e;
    uint64_t nv = 0, nvn = 0, nvt = 0;          // lines of each kind in the piece
    uint64_t v0 = 0, vn0 = 0, vt0 = 0;          // counts before the piece (prefix sums)
    std::vector<Event> events;                   // structural lines, in order
    size_t first_segment = 0;                    // segments of the piece: events.size() + 1, in order
};

template <class Fn> void run_parallel(unsigned threads, size_t n, Fn fn) {
    if (threads <= 1 || n <= 1) {
        for (size_t i = 0; i < n; ++i) fn(i);
        return;
    }
    std::atomic<size_t> next(0);
    std::vector<std::thread> pool;
    const unsigned nt = (unsigned)std::min<size_t>(threads, n);
    for (unsigned t = 0; t < nt; ++t)
        pool.emplace_back([&]() {
            for (size_t i = next.fetch_add(1); i < n; i = next.fetch_add(1)) fn(i);
        });
    for (std::thread &t : pool) t.join();
}

size_t g_piece_bytes = 0; // test hook: forces the piece size so that small files exercise the piece boundaries

double seconds_since(const std::chrono::steady_clock::time_point &t0) {
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

} // namespace

void set_obj_piece_bytes(size_t bytes) { g_piece_bytes = bytes; }

bool load_obj(const std::string &obj_file, const std::string &materials_directory, Model &model, std::string &error, bool verbose,
              unsigned threads, LoadStats *stats, bool decode_textures) {
    const auto t_start = std::chrono::steady_clock::now();
    if (threads == 0) threads = std::max(1u, std::thread::hardware_concurrency());
    const int fd = ::open(obj_file.c_str(), O_RDONLY);
    struct stat st;
    if (fd < 0 || ::fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) {
        if (fd >= 0) ::close(fd);
        error = "Cannot open file [" + obj_file + "]\n";
        return false;
    }
    const size_t size = (size_t)st.st_size;
    // The text is mapped, not copied: the pages are faulted in by the worker threads of pass 1.  The parser may look
    // one byte past a token, so a NUL must follow the text -- the zero fill of the last mapped page provides it,
    // except when the size is an exact multiple of the page size; then the file is read into a buffer instead.
    struct Text {
        const char *p = nullptr; size_t mapped = 0; Pod<char> owned;
        ~Text() { if (mapped) ::munmap(const_cast<char *>(p), mapped); }
    } text;
    const size_t page = (size_t)::sysconf(_SC_PAGESIZE);
    if (size > 0 && size % page != 0) {
        void *m = ::mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m != MAP_FAILED) { text.p = static_cast<const char *>(m); text.mapped = size; ::madvise(m, size, MADV_WILLNEED); }
    }
    if (!text.p) {
        text.owned.resize_uninitialized(size + 1);
        text.owned[size] = '\0';
        text.p = text.owned.data();
        const size_t slice = 64u << 20;
        const size_t n_slices = (size + slice - 1) / slice;
        std::atomic<bool> bad(false);
        run_parallel(threads, n_slices, [&](size_t i) {
            size_t off = i * slice;
            const size_t stop = std::min(size, off + slice);
            while (off < stop) {
                const ssize_t got = ::pread(fd, text.owned.data() + off, stop - off, (off_t)off);
                if (got <= 0) { bad = true; return; }
                off += (size_t)got;
            }
        });
        if (bad) { ::close(fd); error = "Cannot read file [" + obj_file + "]\n"; return false; }
    }
    ::close(fd);
    const double t_read = seconds_since(t_start);
    const char *begin = text.p, *end = begin + size;

    // pieces of at least 1 MB (8 per thread), each starting at a line start
    std::vector<Piece> pieces;
    {
        const size_t target = g_piece_bytes ? g_piece_bytes : std::max<size_t>(size / (threads * 8u) + 1, 1u << 20);
        const char *p = begin;
        while (p < end) {
            const char *q = (size_t)(end - p) > target ? p + target : end;
            if (q < end) { Line ln; q = scan_line(q, end, ln); } // finish the line q falls into
            Piece pc;
            pc.b = p; pc.e = q;
            pieces.push_back(pc);
            p = q;
        }
    }

    // pass 1: counts per piece and per segment; the structural lines
    std::vector<std::vector<Segment>> piece_segments(pieces.size());
    run_parallel(threads, pieces.size(), [&](size_t i) {
        Piece &pc = pieces[i];
        std::vector<Segment> &segs = piece_segments[i];
        segs.emplace_back();
        Line ln;
        for (const char *p = pc.b; p < pc.e;) {
            p = scan_line(p, pc.e, ln);
            const char *body = nullptr;
            switch (classify(ln, body)) {
                case Kind::V: ++pc.nv; break;
                case Kind::VN: ++pc.nvn; break;
                case Kind::VT: ++pc.nvt; break;
                case Kind::F: {
                    const char *c = body;
                    while (c < ln.e && is_blank(*c)) ++c;
                    uint64_t corners = 0;
                    while (c < ln.e) {
                        parse_corner<false>(c, ln.e, 0, 0, 0);
                        skip_corner_gap(c, ln.e);
                        ++corners;
                    }
                    ++segs.back().faces;
                    if (corners >= 3) segs.back().tris += corners - 2;
                    break;
                }
                case Kind::UseMtl: pc.events.push_back(Event{Kind::UseMtl, ln, body}); segs.emplace_back(); break;
                case Kind::MtlLib: pc.events.push_back(Event{Kind::MtlLib, ln, body}); segs.emplace_back(); break;
                case Kind::Group: pc.events.push_back(Event{Kind::Group, ln, nullptr}); segs.emplace_back(); break;
                case Kind::Skip: break;
            }
        }
    });
    const double t_scan = seconds_since(t_start);

    // between the passes (sequential, a few structural lines): materials, which segments survive, where they go
    std::vector<MtlEntry> mtl;
    std::map<std::string, int> mtl_by_name;
    std::ostringstream warn;
    std::vector<Segment> segments;
    std::vector<uint64_t> shape_tris; // triangles per exported shape, for the progress lines
    {
        uint64_t v = 0, vn = 0, vt = 0;
        for (size_t i = 0; i < pieces.size(); ++i) {
            pieces[i].v0 = v; pieces[i].vn0 = vn; pieces[i].vt0 = vt;
            v += pieces[i].nv; vn += pieces[i].nvn; vt += pieces[i].nvt;
            pieces[i].first_segment = segments.size();
            segments.insert(segments.end(), piece_segments[i].begin(), piece_segments[i].end());
        }
        model.positions.resize_uninitialized(v * 3);
        model.normals.resize_uninitialized(vn * 3);
        model.uvs.resize_uninitialized(vt * 2);

        int material = -1;
        std::vector<size_t> pending, shape; // segment ids: faces since the last flush / flushed into the open shape
        uint64_t pending_faces = 0;
        auto flush = [&]() { // exportFaceGroupToShape: true iff any face line was pending
            if (pending_faces == 0) { pending.clear(); return false; }
            shape.insert(shape.end(), pending.begin(), pending.end());
            pending.clear();
            pending_faces = 0;
            return true;
        };
        auto close_shape = [&](bool keep) {
            uint64_t n = 0;
            for (size_t id : shape) { segments[id].keep = keep; n += segments[id].tris; }
            if (keep) shape_tris.push_back(n);
            shape.clear();
        };
        for (size_t i = 0; i < pieces.size(); ++i) {
            const Piece &pc = pieces[i];
            for (size_t k = 0; k <= pc.events.size(); ++k) {
                if (k > 0) { // the structural line that opens segment k of the piece
                    const Event &ev = pc.events[k - 1];
                    if (ev.kind == Kind::UseMtl) {
                        const char *w = ev.body;
                        while (w < ev.line.e && (*w == ' ' || *w == '\t' || *w == '\r' || *w == '\n')) ++w;
                        const char *we = w;
                        while (we < ev.line.e && !(*we == ' ' || *we == '\t' || *we == '\r' || *we == '\n')) ++we;
                        const auto it = mtl_by_name.find(std::string(w, we));
                        const int id = it != mtl_by_name.end() ? it->second : -1;
                        if (id != material) { // per-face materials: flush into the current shape, keep the shape open
                            flush();
                            material = id;
                        }
                    } else if (ev.kind == Kind::MtlLib) {
                        std::istringstream names(std::string(ev.body, ev.line.e));
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
                    } else { // g / o: the shape is exported only if faces were still pending (tinyobjloader 1.0.5)
                        close_shape(flush());
                    }
                }
                const size_t id = pc.first_segment + k;
                segments[id].material = material;
                pending.push_back(id);
                pending_faces += segments[id].faces;
            }
        }
        const bool flushed = flush();
        uint64_t in_shape = 0;
        for (size_t id : shape) in_shape += segments[id].tris;
        if (flushed || in_shape > 0) close_shape(true);
        uint64_t t = 0;
        for (Segment &sg : segments)
            if (sg.keep) { sg.tri_offset = t; t += sg.tris; }
        model.tris.resize_uninitialized(t * 10);
    }
    const double t_resolve = seconds_since(t_start);

    // pass 2: every piece parses its lines into place
    run_parallel(threads, pieces.size(), [&](size_t i) {
        const Piece &pc = pieces[i];
        float *pos = model.positions.data() + pc.v0 * 3, *nrm = model.normals.data() + pc.vn0 * 3, *uv = model.uvs.data() + pc.vt0 * 2;
        int nv = (int)pc.v0, nvn = (int)pc.vn0, nvt = (int)pc.vt0;
        size_t seg = pc.first_segment;
        int32_t *out = model.tris.data() + segments[seg].tri_offset * 10;
        Line ln;
        for (const char *p = pc.b; p < pc.e;) {
            p = scan_line(p, pc.e, ln);
            const char *body = nullptr;
            switch (classify(ln, body)) {
                case Kind::V: for (int k = 0; k < 3; ++k) *pos++ = number(body, ln.e); ++nv; break;
                case Kind::VN: for (int k = 0; k < 3; ++k) *nrm++ = number(body, ln.e); ++nvn; break;
                case Kind::VT: for (int k = 0; k < 2; ++k) *uv++ = number(body, ln.e); ++nvt; break;
                case Kind::F: {
                    const Segment &sg = segments[seg];
                    if (!sg.keep) break;
                    const char *c = body;
                    while (c < ln.e && is_blank(*c)) ++c;
                    Corner c0 = {-1, -1, -1}, c2 = c0;
                    for (uint64_t k = 0; c < ln.e; ++k) { // fan (f0, f[k-1], f[k]), one material id per triangle
                        const Corner cur = parse_corner<true>(c, ln.e, nv, nvn, nvt);
                        skip_corner_gap(c, ln.e);
                        if (k == 0) { c0 = cur; continue; }
                        const Corner c1 = c2;
                        c2 = cur;
                        if (k >= 2) {
                            const int32_t t[10] = {c0.v, c1.v, c2.v, c0.vn, c1.vn, c2.vn, c0.vt, c1.vt, c2.vt, sg.material};
                            std::memcpy(out, t, sizeof t);
                            out += 10;
                        }
                    }
                    break;
                }
                case Kind::UseMtl: case Kind::MtlLib: case Kind::Group:
                    ++seg;
                    out = model.tris.data() + segments[seg].tri_offset * 10;
                    break;
                case Kind::Skip: break;
            }
        }
    });
    const double t_parse = seconds_since(t_start);

    error = warn.str();
    // load_materials (fileloader.cpp:47-58)
    for (const MtlEntry &e : mtl) {
        MaterialData m;
        m.kd[0] = e.kd[0]; m.kd[1] = e.kd[1]; m.kd[2] = e.kd[2];
        if (!e.map_kd.empty()) {
            std::string terr;
            if (decode_textures) {
                if (!load_texture(materials_directory + e.map_kd, m, terr, verbose)) { error += terr; return false; }
            } else {
                m.has_texture = true;
                m.texture_file = materials_directory + e.map_kd; // fileloader.cpp:55
            }
        }
        model.materials.push_back(m);
    }
    model.shape_triangles = shape_tris;
    // load_triangles per shape (fileloader.cpp:60-77,116-118)
    if (verbose) {
        for (uint64_t n : shape_tris) std::cout << "Loading " << n << " triangles..." << std::endl;
        std::cout << "Loaded model " << obj_file << "." << std::endl;
    }
    if (stats) {
        stats->threads = threads;
        stats->file_bytes = size;
        stats->read_s = t_read; stats->scan_s = t_scan - t_read; stats->resolve_s = t_resolve - t_scan; stats->parse_s = t_parse - t_resolve;
        stats->total_s = seconds_since(t_start);
    }
    return true;
}

// ---- binary mesh cache ---------------------------------------------------------------------------------------
namespace {
struct CacheHeader {
    char magic[8];          // "RASTMESH"
    uint32_t version, n_materials;
    uint64_t n_pos, n_nrm, n_uv, n_tri_ints;
};
bool write_all(std::FILE *f, const void *p, size_t n) { return n == 0 || std::fwrite(p, 1, n, f) == n; }
bool read_all(std::FILE *f, void *p, size_t n) { return n == 0 || std::fread(p, 1, n, f) == n; }
} // namespace

bool save_mesh_cache(const std::string &file, const Model &model, std::string &error) {
    std::FILE *f = std::fopen(file.c_str(), "wb");
    if (!f) { error = "Cannot write mesh cache [" + file + "]"; return false; }
    CacheHeader h = {{'R', 'A', 'S', 'T', 'M', 'E', 'S', 'H'}, 1u, (uint32_t)model.materials.size(),
                     model.positions.size(), model.normals.size(), model.uvs.size(), model.tris.size()};
    bool ok = write_all(f, &h, sizeof h) && write_all(f, model.positions.data(), model.positions.size() * 4) &&
              write_all(f, model.normals.data(), model.normals.size() * 4) && write_all(f, model.uvs.data(), model.uvs.size() * 4) &&
              write_all(f, model.tris.data(), model.tris.size() * 4);
    for (const MaterialData &m : model.materials) {
        const uint32_t len = m.has_texture ? (uint32_t)m.texture_file.size() : 0u;
        ok = ok && write_all(f, m.kd, 12) && write_all(f, &len, 4) && write_all(f, m.texture_file.data(), len);
    }
    ok = (std::fclose(f) == 0) && ok;
    if (!ok) error = "Cannot write mesh cache [" + file + "]";
    return ok;
}

bool load_mesh_cache(const std::string &file, Model &model, std::string &error, bool verbose) {
    std::FILE *f = std::fopen(file.c_str(), "rb");
    if (!f) { error = "Cannot open mesh cache [" + file + "]"; return false; }
    CacheHeader h;
    bool ok = read_all(f, &h, sizeof h) && !std::memcmp(h.magic, "RASTMESH", 8) && h.version == 1u;
    if (ok) {
        model.positions.resize_uninitialized(h.n_pos); model.normals.resize_uninitialized(h.n_nrm);
        model.uvs.resize_uninitialized(h.n_uv); model.tris.resize_uninitialized(h.n_tri_ints);
        ok = read_all(f, model.positions.data(), h.n_pos * 4) && read_all(f, model.normals.data(), h.n_nrm * 4) &&
             read_all(f, model.uvs.data(), h.n_uv * 4) && read_all(f, model.tris.data(), h.n_tri_ints * 4);
    }
    for (uint32_t i = 0; ok && i < h.n_materials; ++i) {
        MaterialData m;
        uint32_t len = 0;
        ok = read_all(f, m.kd, 12) && read_all(f, &len, 4) && len < (1u << 20);
        std::string tex(len, '\0');
        ok = ok && read_all(f, &tex[0], len);
        if (ok && len) {
            std::string terr;
            if (!load_texture(tex, m, terr, verbose)) { error = terr; std::fclose(f); return false; }
        }
        if (ok) model.materials.push_back(m);
    }
    std::fclose(f);
    if (!ok) { error = "Bad mesh cache [" + file + "]"; return false; }
    if (verbose) std::cout << "Loaded model " << file << "." << std::endl;
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
