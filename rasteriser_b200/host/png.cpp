#include "png.hpp"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

#include <zlib.h>

namespace host {
namespace {

const uint8_t kSignature[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};

uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

} // namespace

std::string png_read(const std::string &path, PngImage &out) {
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) return "cannot open " + path;
    std::vector<uint8_t> file;
    uint8_t buf[65536];
    size_t n;
    while ((n = std::fread(buf, 1, sizeof buf, f)) > 0) file.insert(file.end(), buf, buf + n);
    std::fclose(f);
    if (file.size() < 8 || std::memcmp(file.data(), kSignature, 8) != 0) return path + ": not a PNG file";

    unsigned width = 0, height = 0, depth = 0, colour = 0, interlace = 0;
    std::vector<uint8_t> idat, palette;
    size_t pos = 8;
    bool seen_end = false;
    while (pos + 12 <= file.size() && !seen_end) {
        const uint32_t len = be32(&file[pos]);
        const char *type = reinterpret_cast<const char *>(&file[pos + 4]);
        if (pos + 12 + (size_t)len > file.size()) return path + ": truncated chunk";
        const uint8_t *data = &file[pos + 8];
        if (!std::memcmp(type, "IHDR", 4) && len >= 13) {
            width = be32(data); height = be32(data + 4); depth = data[8]; colour = data[9]; interlace = data[12];
        } else if (!std::memcmp(type, "PLTE", 4)) {
            palette.assign(data, data + len);
        } else if (!std::memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), data, data + len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            seen_end = true;
        }
        pos += 12 + (size_t)len;
    }
    if (!width || !height) return path + ": missing IHDR";
    if (depth != 8 || interlace != 0) return path + ": only 8-bit non-interlaced PNG is supported";
    unsigned src_channels;
    switch (colour) {
        case 0: src_channels = 1; break;
        case 2: src_channels = 3; break;
        case 3: src_channels = 1; break;
        case 4: src_channels = 2; break;
        case 6: src_channels = 4; break;
        default: return path + ": unknown colour type";
    }
    const size_t stride = (size_t)width * src_channels;
    std::vector<uint8_t> raw((stride + 1) * height);
    uLongf raw_len = (uLongf)raw.size();
    if (uncompress(raw.data(), &raw_len, idat.data(), (uLong)idat.size()) != Z_OK || raw_len != raw.size()) return path + ": zlib inflate failed";

    // undo the per-row filters (PNG spec section 9)
    std::vector<uint8_t> img(stride * height);
    const unsigned bpp = src_channels;
    for (unsigned y = 0; y < height; ++y) {
        const uint8_t filter = raw[(stride + 1) * y];
        const uint8_t *in = &raw[(stride + 1) * y + 1];
        uint8_t *cur = &img[stride * y];
        const uint8_t *up = y ? &img[stride * (y - 1)] : nullptr;
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= bpp ? cur[i - bpp] : 0, b = up ? up[i] : 0, c = (up && i >= bpp) ? up[i - bpp] : 0;
            int v = in[i];
            switch (filter) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) >> 1; break;
                case 4: v += paeth(a, b, c); break;
                default: return path + ": bad filter type";
            }
            cur[i] = (uint8_t)v;
        }
    }
    out.width = width;
    out.height = height;
    if (colour == 3) { // palette -> RGB
        out.channels = 3;
        out.pixels.resize((size_t)width * height * 3);
        for (size_t i = 0; i < (size_t)width * height; ++i) {
            const size_t p = (size_t)img[i] * 3;
            for (int k = 0; k < 3; ++k) out.pixels[3 * i + k] = p + k < palette.size() ? palette[p + k] : 0;
        }
    } else {
        out.channels = src_channels;
        out.pixels.swap(img);
    }
    return "";
}

// ---- writer ------------------------------------------------------------------------------------------------------
// The image is cut into bands of rows, one per worker thread.  Each band is filtered (type 0) and deflated on its
// own as a raw deflate stream that ends on a byte boundary (Z_SYNC_FLUSH; the last band ends the stream with
// Z_FINISH), so the concatenation of the band streams between one zlib header and the Adler-32 of the whole image
// is ONE valid zlib stream -- the pigz construction.  Every band becomes its own IDAT chunk (a decoder
// concatenates IDAT payloads, PNG spec 11.2.4).
namespace {

unsigned g_png_threads = 0; // 0 = hardware threads

struct Band {
    unsigned y0 = 0, y1 = 0;
    std::vector<uint8_t> data;  // this band's part of the zlib stream (band 0 starts with the zlib header, the last band
                                // ends with the Adler-32 of the whole image)
    uLong adler = 1;            // Adler-32 of the band's filtered bytes
    uLong raw_len = 0;
    bool ok = false;
};

// rows(y, dst): writes the `stride` interleaved bytes of row y.  Returns the zlib stream of the filtered image in
// pieces, one per band, or an empty vector on failure.
template <class RowFn>
std::vector<Band> deflate_image(unsigned width, unsigned height, unsigned channels, RowFn rows) {
    const size_t stride = (size_t)width * channels;
    unsigned threads = g_png_threads ? g_png_threads : std::max(1u, std::thread::hardware_concurrency());
    // at least 64 KB of pixels per band, so that restarting the deflate window costs nothing measurable
    const size_t min_rows = std::max<size_t>(1, (64u << 10) / std::max<size_t>(stride, 1));
    const unsigned n_bands = (unsigned)std::max<size_t>(1, std::min<size_t>(threads, height / min_rows));
    std::vector<Band> bands(n_bands);
    for (unsigned b = 0; b < n_bands; ++b) {
        bands[b].y0 = (unsigned)((uint64_t)height * b / n_bands);
        bands[b].y1 = (unsigned)((uint64_t)height * (b + 1) / n_bands);
    }
    auto work = [&](unsigned b) {
        Band &bd = bands[b];
        const size_t n_rows = bd.y1 - bd.y0;
        std::vector<uint8_t> raw((stride + 1) * n_rows);
        for (size_t r = 0; r < n_rows; ++r) { // filter type 0 (None) per row: fastest, values are what matters
            raw[(stride + 1) * r] = 0;
            rows(bd.y0 + (unsigned)r, &raw[(stride + 1) * r + 1]);
        }
        bd.raw_len = (uLong)raw.size();
        bd.adler = adler32(1L, raw.data(), (uInt)raw.size());
        z_stream zs;
        std::memset(&zs, 0, sizeof zs);
        if (deflateInit2(&zs, 1, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return;
        const size_t head = b == 0 ? 2 : 0, tail = b + 1 == n_bands ? 4 : 0; // zlib header / Adler-32 slot
        bd.data.resize(head + deflateBound(&zs, (uLong)raw.size()) + 16 + tail);
        if (head) { bd.data[0] = 0x78; bd.data[1] = 0x01; } // deflate, 32 KB window, fastest level, no dictionary
        zs.next_in = raw.data();
        zs.avail_in = (uInt)raw.size();
        zs.next_out = bd.data.data() + head;
        zs.avail_out = (uInt)(bd.data.size() - head - tail);
        const int rc = deflate(&zs, b + 1 == n_bands ? Z_FINISH : Z_SYNC_FLUSH);
        const bool done = (b + 1 == n_bands) ? rc == Z_STREAM_END : (rc == Z_OK && zs.avail_in == 0 && zs.avail_out != 0);
        const size_t clen = head + zs.total_out;
        deflateEnd(&zs);
        if (!done) return;
        bd.data.resize(clen + tail); // the tail is filled in below
        bd.ok = true;
    };
    if (n_bands == 1) work(0);
    else {
        std::vector<std::thread> pool;
        for (unsigned b = 0; b < n_bands; ++b) pool.emplace_back(work, b);
        for (std::thread &t : pool) t.join();
    }
    uLong adler = 1;
    for (unsigned b = 0; b < n_bands; ++b) {
        if (!bands[b].ok) return std::vector<Band>();
        adler = b == 0 ? bands[0].adler : adler32_combine(adler, bands[b].adler, (z_off_t)bands[b].raw_len);
    }
    std::vector<uint8_t> &last = bands[n_bands - 1].data;
    const size_t n = last.size();
    last[n - 4] = (uint8_t)(adler >> 24); last[n - 3] = (uint8_t)(adler >> 16); last[n - 2] = (uint8_t)(adler >> 8); last[n - 1] = (uint8_t)adler;
    return bands;
}

// one chunk straight to the file: big-endian length, type, [4-byte prefix,] payload, CRC over type + prefix + payload
bool write_chunk(FILE *f, const char type[4], const uint8_t *prefix, size_t prefix_len, const uint8_t *payload, size_t len) {
    uint8_t head[8];
    const uint32_t total = (uint32_t)(prefix_len + len);
    head[0] = (uint8_t)(total >> 24); head[1] = (uint8_t)(total >> 16); head[2] = (uint8_t)(total >> 8); head[3] = (uint8_t)total;
    std::memcpy(head + 4, type, 4);
    uLong crc = crc32(0L, head + 4, 4);
    if (prefix_len) crc = crc32(crc, prefix, (uInt)prefix_len);
    if (len) crc = crc32(crc, payload, (uInt)len);
    const uint8_t tail[4] = {(uint8_t)(crc >> 24), (uint8_t)(crc >> 16), (uint8_t)(crc >> 8), (uint8_t)crc};
    return std::fwrite(head, 1, 8, f) == 8 && (prefix_len == 0 || std::fwrite(prefix, 1, prefix_len, f) == prefix_len) &&
           (len == 0 || std::fwrite(payload, 1, len, f) == len) && std::fwrite(tail, 1, 4, f) == 4;
}

bool write_header(FILE *f, unsigned width, unsigned height, unsigned channels) {
    static const uint8_t colour_of[5] = {0, 0, 4, 2, 6};
    uint8_t ihdr[13] = {(uint8_t)(width >> 24), (uint8_t)(width >> 16), (uint8_t)(width >> 8), (uint8_t)width,
                        (uint8_t)(height >> 24), (uint8_t)(height >> 16), (uint8_t)(height >> 8), (uint8_t)height, 8, colour_of[channels], 0, 0, 0};
    return std::fwrite(kSignature, 1, 8, f) == 8 && write_chunk(f, "IHDR", nullptr, 0, ihdr, 13);
}

template <class RowFn>
std::string write_png_file(const std::string &path, unsigned width, unsigned height, unsigned channels, RowFn rows) {
    if (channels < 1 || channels > 4) return "png_write: 1..4 channels";
    const std::vector<Band> bands = deflate_image(width, height, channels, rows);
    if (bands.empty()) return "png_write: deflate failed";
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) return "cannot write " + path;
    bool ok = write_header(f, width, height, channels);
    for (size_t b = 0; ok && b < bands.size(); ++b) ok = write_chunk(f, "IDAT", nullptr, 0, bands[b].data.data(), bands[b].data.size());
    ok = ok && write_chunk(f, "IEND", nullptr, 0, nullptr, 0);
    ok = (std::fclose(f) == 0) && ok;
    return ok ? "" : "short write to " + path;
}

template <class Dst> void interleave_row(const uint8_t *planar, size_t plane, unsigned width, unsigned channels, unsigned y, Dst *dst) {
    const uint8_t *row = planar + (size_t)y * width;
    if (channels == 3) {
        const uint8_t *r = row, *g = row + plane, *b = row + 2 * plane;
        for (unsigned x = 0; x < width; ++x) { dst[3 * x] = r[x]; dst[3 * x + 1] = g[x]; dst[3 * x + 2] = b[x]; }
    } else {
        for (unsigned c = 0; c < channels; ++c)
            for (unsigned x = 0; x < width; ++x) dst[(size_t)x * channels + c] = row[c * plane + x];
    }
}

void put_u32(uint8_t *p, uint32_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; }

} // namespace

void png_set_threads(unsigned threads) { g_png_threads = threads; }

std::string png_write(const std::string &path, const uint8_t *interleaved, unsigned width, unsigned height, unsigned channels) {
    const size_t stride = (size_t)width * channels;
    return write_png_file(path, width, height, channels, [=](unsigned y, uint8_t *dst) { std::memcpy(dst, interleaved + stride * y, stride); });
}

// planar [c][h][w] (CImg layout, CImg.h:11715-11721) -> file; the interleaving happens row by row inside the bands
std::string png_write_planar(const std::string &path, const uint8_t *planar, unsigned width, unsigned height, unsigned channels) {
    const size_t plane = (size_t)width * height;
    return write_png_file(path, width, height, channels, [=](unsigned y, uint8_t *dst) { interleave_row(planar, plane, width, channels, y, dst); });
}

// ---- animated PNG (APNG 1.0): the recorded form of the spin sequence ---------------------------------------------
// signature, IHDR, acTL(frames, plays), then per frame fcTL + the frame's zlib stream -- IDAT chunks for frame 0,
// fdAT chunks (sequence number + data) for the others -- and IEND.  Viewers without APNG support show frame 0.
ApngWriter::~ApngWriter() { if (f_) std::fclose(static_cast<FILE *>(f_)); }

std::string ApngWriter::open(const std::string &path, unsigned width, unsigned height, unsigned channels, unsigned n_frames, unsigned delay_ms) {
    if (channels < 1 || channels > 4 || n_frames == 0) return "apng: bad arguments";
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) return "cannot write " + path;
    f_ = f; path_ = path; w_ = width; h_ = height; c_ = channels; frames_ = n_frames; delay_ms_ = delay_ms; written_ = 0; seq_ = 0;
    uint8_t actl[8];
    put_u32(actl, n_frames);
    put_u32(actl + 4, 0); // loop forever
    if (!write_header(f, width, height, channels) || !write_chunk(f, "acTL", nullptr, 0, actl, 8)) return "short write to " + path;
    return "";
}

std::string ApngWriter::add_frame_planar(const uint8_t *planar) {
    FILE *f = static_cast<FILE *>(f_);
    if (!f || written_ >= frames_) return "apng: not open or too many frames";
    const size_t plane = (size_t)w_ * h_;
    const unsigned width = w_, channels = c_;
    const std::vector<Band> bands = deflate_image(w_, h_, c_, [=](unsigned y, uint8_t *dst) { interleave_row(planar, plane, width, channels, y, dst); });
    if (bands.empty()) return "apng: deflate failed";
    uint8_t fctl[26];
    put_u32(fctl, seq_++);
    put_u32(fctl + 4, w_); put_u32(fctl + 8, h_); put_u32(fctl + 12, 0); put_u32(fctl + 16, 0);
    fctl[20] = (uint8_t)(delay_ms_ >> 8); fctl[21] = (uint8_t)delay_ms_; // delay numerator
    fctl[22] = (uint8_t)(1000 >> 8); fctl[23] = (uint8_t)(1000 & 0xFF);   // denominator: milliseconds
    fctl[24] = 0; fctl[25] = 0;                                           // dispose: none, blend: source
    bool ok = write_chunk(f, "fcTL", nullptr, 0, fctl, 26);
    for (size_t b = 0; ok && b < bands.size(); ++b) {
        if (written_ == 0) ok = write_chunk(f, "IDAT", nullptr, 0, bands[b].data.data(), bands[b].data.size());
        else {
            uint8_t seq[4];
            put_u32(seq, seq_++);
            ok = write_chunk(f, "fdAT", seq, 4, bands[b].data.data(), bands[b].data.size());
        }
    }
    ++written_;
    return ok ? "" : "short write to " + path_;
}

std::string ApngWriter::close() {
    FILE *f = static_cast<FILE *>(f_);
    if (!f) return "";
    bool ok = written_ == frames_ && write_chunk(f, "IEND", nullptr, 0, nullptr, 0);
    ok = (std::fclose(f) == 0) && ok;
    f_ = nullptr;
    return ok ? "" : "apng: incomplete file " + path_;
}

} // namespace host
