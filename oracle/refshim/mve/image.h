// refshim: MVE mve::Image stand-in (interleaved, row major; see ../README.md)
#pragma once
#include <algorithm>
#include <cstdint>
#include <memory>
#include <vector>
#include "math/functions.h"

namespace mve {

enum ImageType { IMAGE_TYPE_UNKNOWN, IMAGE_TYPE_UINT8, IMAGE_TYPE_FLOAT };

template <typename T>
class Image {
public:
    typedef std::shared_ptr<Image<T> > Ptr;
    typedef std::shared_ptr<Image<T> const> ConstPtr;
    typedef std::vector<T> ImageData;

    Image() : w(0), h(0), c(0) {}
    Image(int width, int height, int channels) { allocate(width, height, channels); }
    static Ptr create() { return Ptr(new Image<T>()); }
    static Ptr create(int width, int height, int channels) { return Ptr(new Image<T>(width, height, channels)); }
    static Ptr create(Image<T> const& o) { return Ptr(new Image<T>(o)); }
    Ptr duplicate() const { return Ptr(new Image<T>(*this)); }

    void allocate(int width, int height, int channels) {
        w = width; h = height; c = channels;
        data.assign(static_cast<std::size_t>(w) * h * c, T(0));
    }
    void fill(T const& value) { std::fill(data.begin(), data.end(), value); }
    void fill_color(T const* color) {
        for (std::size_t i = 0; i < data.size(); i += c) for (int k = 0; k < c; ++k) data[i + k] = color[k];
    }
    int width() const { return w; }
    int height() const { return h; }
    int channels() const { return c; }
    int get_pixel_amount() const { return w * h; }
    int get_value_amount() const { return w * h * c; }
    std::size_t get_byte_size() const { return data.size() * sizeof(T); }
    T* get_data_pointer() { return data.data(); }
    T const* get_data_pointer() const { return data.data(); }
    ImageData& get_data() { return data; }
    ImageData const& get_data() const { return data; }
    T* begin() { return data.data(); }
    T* end() { return data.data() + data.size(); }
    T const* begin() const { return data.data(); }
    T const* end() const { return data.data() + data.size(); }
    bool valid() const { return w && h && c; }

    T& at(int index) { return data[index]; }
    T const& at(int index) const { return data[index]; }
    T& at(int index, int channel) { return data[index * c + channel]; }
    T const& at(int index, int channel) const { return data[index * c + channel]; }
    T& at(int x, int y, int channel) { return data[(y * w + x) * c + channel]; }
    T const& at(int x, int y, int channel) const { return data[(y * w + x) * c + channel]; }
    T& operator[](int index) { return data[index]; }
    T const& operator[](int index) const { return data[index]; }

    // clamp, truncate, weights w0..w3, math::interpolate over the four neighbours
    T linear_at(float x, float y, int channel) const {
        x = std::max(0.0f, std::min(static_cast<float>(w - 1), x));
        y = std::max(0.0f, std::min(static_cast<float>(h - 1), y));
        int const floor_x = static_cast<int>(x);
        int const floor_y = static_cast<int>(y);
        int const floor_xp1 = std::min(floor_x + 1, w - 1);
        int const floor_yp1 = std::min(floor_y + 1, h - 1);
        float const w1 = x - static_cast<float>(floor_x);
        float const w0 = 1.0f - w1;
        float const w3 = y - static_cast<float>(floor_y);
        float const w2 = 1.0f - w3;
        int const rowstride = w * c;
        int const row1 = floor_y * rowstride;
        int const row2 = floor_yp1 * rowstride;
        int const col1 = floor_x * c;
        int const col2 = floor_xp1 * c;
        return math::interpolate<T>(data[row1 + col1 + channel], data[row1 + col2 + channel],
                                    data[row2 + col1 + channel], data[row2 + col2 + channel],
                                    w0 * w2, w1 * w2, w0 * w3, w1 * w3);
    }
    void linear_at(float x, float y, T* px) const { for (int k = 0; k < c; ++k) px[k] = linear_at(x, y, k); }

    void delete_channel(int channel) {
        std::vector<T> nd;
        nd.reserve(static_cast<std::size_t>(w) * h * (c - 1));
        for (std::size_t i = 0; i < data.size(); ++i) if (static_cast<int>(i % c) != channel) nd.push_back(data[i]);
        data.swap(nd); c -= 1;
    }

private:
    int w, h, c;
    std::vector<T> data;
};

typedef Image<std::uint8_t> ByteImage;
typedef Image<float> FloatImage;
typedef Image<double> DoubleImage;
typedef Image<int> IntImage;

}  // namespace mve
