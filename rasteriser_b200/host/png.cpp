#include "png.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <zlib.h>

namespace host {
namespace {

const uint8_t kSignature[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};

uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
void put_be32(std::vector<uint8_t> &v, uint32_t x) {
    v.push_back((uint8_t)(x >> 24)); v.push_back((uint8_t)(x >> 16)); v.push_back((uint8_t)(x >> 8)); v.push_back((uint8_t)x);
}

void put_chunk(std::vector<uint8_t> &file, const char type[4], const std::vector<uint8_t> &payload) {
    put_be32(file, (uint32_t)payload.size());
    const size_t start = file.size();
    file.insert(file.end(), type, type + 4);
    file.insert(file.end(), payload.begin(), payload.end());
    put_be32(file, (uint32_t)crc32(0L, file.data() + start, (uInt)(file.size() - start)));
}

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

std::string png_write(const std::string &path, const uint8_t *interleaved, unsigned width, unsigned height, unsigned channels) {
    if (channels < 1 || channels > 4) return "png_write: 1..4 channels";
    static const uint8_t colour_of[5] = {0, 0, 4, 2, 6};
    const size_t stride = (size_t)width * channels;
    std::vector<uint8_t> raw((stride + 1) * height);
    for (unsigned y = 0; y < height; ++y) { // filter type 0 (None) per row: fastest, values are what matters
        raw[(stride + 1) * y] = 0;
        std::memcpy(&raw[(stride + 1) * y + 1], interleaved + stride * y, stride);
    }
    uLongf clen = compressBound((uLong)raw.size());
    std::vector<uint8_t> comp(clen);
    if (compress2(comp.data(), &clen, raw.data(), (uLong)raw.size(), 1) != Z_OK) return "png_write: deflate failed";
    comp.resize(clen);

    std::vector<uint8_t> file(kSignature, kSignature + 8), ihdr;
    put_be32(ihdr, width);
    put_be32(ihdr, height);
    ihdr.push_back(8);
    ihdr.push_back(colour_of[channels]);
    ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    put_chunk(file, "IHDR", ihdr);
    put_chunk(file, "IDAT", comp);
    put_chunk(file, "IEND", std::vector<uint8_t>());
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) return "cannot write " + path;
    const bool ok = std::fwrite(file.data(), 1, file.size(), f) == file.size();
    std::fclose(f);
    return ok ? "" : "short write to " + path;
}

std::string png_write_planar(const std::string &path, const uint8_t *planar, unsigned width, unsigned height, unsigned channels) {
    const size_t plane = (size_t)width * height;
    std::vector<uint8_t> inter(plane * channels);
    for (unsigned c = 0; c < channels; ++c)
        for (size_t i = 0; i < plane; ++i) inter[i * channels + c] = planar[c * plane + i];
    return png_write(path, inter.data(), width, height, channels);
}

} // namespace host
