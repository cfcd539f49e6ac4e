// png.hpp -- minimal PNG reader / writer on zlib (8-bit, non-interlaced).
//
// Replaces the PNG path of the reference, which shells out to ImageMagick through CImg
// (frame_buffer.save("frame.png"), renderer.cpp:92-93; texture load, material.h:20) and therefore
// cannot run offline.  Pixel values are identical on decode; the file bytes are not (different encoder).
#pragma once

#include <cstdint>
#include <string>
#include <vector>

namespace host {

struct PngImage {
    unsigned width = 0, height = 0, channels = 0; // channels: 1 grey, 2 grey+alpha, 3 RGB, 4 RGBA
    std::vector<uint8_t> pixels;                   // interleaved, row-major
};

// Returns an empty string on success, else an error message.
std::string png_read(const std::string &path, PngImage &out);
std::string png_write(const std::string &path, const uint8_t *interleaved, unsigned width, unsigned height, unsigned channels);
// The writer deflates bands of rows on several threads (0 = all hardware threads, the default).
void png_set_threads(unsigned threads);
// planar [c][h][w] (CImg layout) -> file
std::string png_write_planar(const std::string &path, const uint8_t *planar, unsigned width, unsigned height, unsigned channels);

// Animated PNG: frames of one size appended one at a time (planar CImg layout in), `delay_ms` between frames.
class ApngWriter {
public:
    ApngWriter() = default;
    ~ApngWriter();
    ApngWriter(const ApngWriter &) = delete;
    ApngWriter &operator=(const ApngWriter &) = delete;
    std::string open(const std::string &path, unsigned width, unsigned height, unsigned channels, unsigned n_frames, unsigned delay_ms);
    std::string add_frame_planar(const uint8_t *planar);
    std::string close(); // fails unless exactly n_frames were added
private:
    void *f_ = nullptr;
    std::string path_;
    unsigned w_ = 0, h_ = 0, c_ = 0, frames_ = 0, delay_ms_ = 0, written_ = 0, seq_ = 0;
};

} // namespace host
