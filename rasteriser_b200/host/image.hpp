// image.hpp -- the small planar image container the host program needs.
//
// The reference keeps its frame / depth buffers and textures in CImg<T> (vendor/cimg/CImg.h, 60k
// lines, PNG only through ImageMagick).  The host side here needs four things from it: planar
// storage x + y*W + c*W*H (CImg.h:11715-11721), fill (:25804-25809), normalize (:26786-26794) and
// saving as 8-bit PNG -- restated below in a page instead of vendoring the library.
#pragma once

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace host {

template <typename T> class Image {
public:
    Image() : w_(0), h_(0), c_(0) {}
    // CImg<T>(w, h, 1, c, value)
    Image(unsigned w, unsigned h, unsigned channels, T value) : w_(w), h_(h), c_(channels), data_((size_t)w * h * channels, value) {}

    unsigned width() const { return w_; }
    unsigned height() const { return h_; }
    unsigned spectrum() const { return c_; }
    size_t size() const { return data_.size(); }
    bool is_empty() const { return data_.empty(); }
    T *data() { return data_.data(); }
    const T *data() const { return data_.data(); }
    T &operator()(unsigned x, unsigned y, unsigned c = 0) { return data_[x + (size_t)y * w_ + (size_t)c * w_ * h_]; }
    const T &operator()(unsigned x, unsigned y, unsigned c = 0) const { return data_[x + (size_t)y * w_ + (size_t)c * w_ * h_]; }

    Image &fill(T value) {
        for (T &v : data_) v = value;
        return *this;
    }

    // CImg::normalize(min_value, max_value): (v - m) / (M - m) * (b - a) + a over ALL channels jointly, in T's
    // float type; if the image is constant it is filled with min_value; untouched if already [a, b].
    Image &normalize(T min_value, T max_value) {
        if (data_.empty()) return *this;
        const T a = min_value < max_value ? min_value : max_value, b = min_value < max_value ? max_value : min_value;
        T m = data_[0], M = data_[0];
        for (const T &v : data_) {
            if (v > M) M = v;
            if (v < m) m = v;
        }
        if (m == M) return fill(min_value);
        if (m != a || M != b)
            for (T &v : data_) v = (T)((v - m) / (M - m) * (b - a) + a);
        return *this;
    }

private:
    unsigned w_, h_, c_;
    std::vector<T> data_;
};

} // namespace host
