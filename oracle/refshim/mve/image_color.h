// refshim: MVE mve/image_color.h stand-in (see ../README.md)
#pragma once

namespace mve { namespace image {

// float overload of MVE's colour conversion, operation order as in oracle/datacosts.c
template <typename T>
inline void color_rgb_to_ycbcr(T* v) {
    T out[3];
    out[0] = (v[0] * T(0.299) + v[1] * T(0.587)) + v[2] * T(0.114);
    out[1] = ((v[0] * T(-0.168736) + v[1] * T(-0.331264)) + v[2] * T(0.5)) + T(0.5);
    out[2] = ((v[0] * T(0.5) + v[1] * T(-0.418688)) + v[2] * T(-0.081312)) + T(0.5);
    v[0] = out[0]; v[1] = out[1]; v[2] = out[2];
}

} }  // namespace mve::image
