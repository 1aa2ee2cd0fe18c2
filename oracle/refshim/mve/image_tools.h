// refshim: MVE mve/image_tools.h stand-in: the handful of operations libs/tex calls, restated as in
// oracle/imgprep.c (see ../README.md)
#pragma once
#include <cmath>
#include "mve/image.h"

namespace mve { namespace image {

enum DesaturateType { DESATURATE_MAXIMUM, DESATURATE_LIGHTNESS, DESATURATE_LUMINOSITY, DESATURATE_LUMINANCE, DESATURATE_AVERAGE };

template <typename T>
inline typename Image<T>::Ptr desaturate(typename Image<T>::ConstPtr img, DesaturateType /*luminance*/) {
    typename Image<T>::Ptr out = Image<T>::create(img->width(), img->height(), 1);
    int const n = img->get_pixel_amount();
    for (int i = 0; i < n; ++i)
        out->at(i) = math::interpolate<T>(img->at(i, 0), img->at(i, 1), img->at(i, 2), 0.21f, 0.72f, 0.07f);
    return out;
}

template <typename T>
inline typename Image<T>::Ptr sobel_edge(typename Image<T>::ConstPtr img) {
    int const w = img->width(), h = img->height(), c = img->channels();
    typename Image<T>::Ptr out = Image<T>::create(w, h, c);
    double const max_value = 255.0;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
            for (int cc = 0; cc < c; ++cc) {
                if (y == 0 || y == h - 1 || x == 0 || x == w - 1) { out->at(x, y, cc) = T(0); continue; }
                double gx = 1.0 * img->at(x + 1, y - 1, cc) - 1.0 * img->at(x - 1, y - 1, cc) + 2.0 * img->at(x + 1, y, cc)
                    - 2.0 * img->at(x - 1, y, cc) + 1.0 * img->at(x + 1, y + 1, cc) - 1.0 * img->at(x - 1, y + 1, cc);
                double gy = 1.0 * img->at(x - 1, y + 1, cc) - 1.0 * img->at(x - 1, y - 1, cc) + 2.0 * img->at(x, y + 1, cc)
                    - 2.0 * img->at(x, y - 1, cc) + 1.0 * img->at(x + 1, y + 1, cc) - 1.0 * img->at(x + 1, y - 1, cc);
                double g = std::sqrt(gx * gx + gy * gy);
                out->at(x, y, cc) = static_cast<T>(g < max_value ? g : max_value);
            }
    return out;
}

// crop with out-of-range pixels set to fill_color; left/top may be negative
template <typename T>
inline typename Image<T>::Ptr crop(typename Image<T>::ConstPtr img, int width, int height, int left, int top, T const* fill_color) {
    int const c = img->channels();
    typename Image<T>::Ptr out = Image<T>::create(width, height, c);
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x) {
            int const sx = x + left, sy = y + top;
            bool const in = sx >= 0 && sx < img->width() && sy >= 0 && sy < img->height();
            for (int k = 0; k < c; ++k) out->at(x, y, k) = in ? img->at(sx, sy, k) : fill_color[k];
        }
    return out;
}

inline FloatImage::Ptr byte_to_float_image(ByteImage::ConstPtr img) {
    FloatImage::Ptr out = FloatImage::create(img->width(), img->height(), img->channels());
    for (int i = 0; i < img->get_value_amount(); ++i) out->at(i) = static_cast<float>(img->at(i)) / 255.0f;
    return out;
}
inline ByteImage::Ptr float_to_byte_image(FloatImage::ConstPtr img, float vmin = 0.0f, float vmax = 1.0f) {
    ByteImage::Ptr out = ByteImage::create(img->width(), img->height(), img->channels());
    for (int i = 0; i < img->get_value_amount(); ++i) {
        float v = std::min(vmax, std::max(vmin, img->at(i)));
        out->at(i) = static_cast<std::uint8_t>(255.0f * (v - vmin) / (vmax - vmin) + 0.5f);
    }
    return out;
}
template <typename T> inline void gamma_correct(typename Image<T>::Ptr img, T const& power) {
    for (T* p = img->begin(); p != img->end(); ++p) *p = std::pow(*p, power);
}

} }  // namespace mve::image
