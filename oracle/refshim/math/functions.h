// refshim: MVE math/functions.h stand-in (see ../README.md)
#pragma once
#include <algorithm>
#include "math/defines.h"

namespace math {

template <typename T> inline T const& clamp(T const& v, T const& lo = T(0), T const& hi = T(1)) {
    return v < lo ? lo : (v > hi ? hi : v);
}
template <typename T> inline T fastpow(T const& base, unsigned int p) {
    T r = T(1); for (unsigned int i = 0; i < p; ++i) r *= base; return r;
}

// weighted sums, left to right; the unsigned char overloads round with + 0.5f
template <typename T> inline T interpolate(T const& v1, T const& v2, float w1, float w2) { return v1 * w1 + v2 * w2; }
template <typename T> inline T interpolate(T const& v1, T const& v2, T const& v3, float w1, float w2, float w3) {
    return v1 * w1 + v2 * w2 + v3 * w3;
}
template <typename T> inline T interpolate(T const& v1, T const& v2, T const& v3, T const& v4, float w1, float w2, float w3, float w4) {
    return v1 * w1 + v2 * w2 + v3 * w3 + v4 * w4;
}
template <> inline unsigned char interpolate(unsigned char const& v1, unsigned char const& v2, float w1, float w2) {
    return (unsigned char)((float)v1 * w1 + (float)v2 * w2 + 0.5f);
}
template <> inline unsigned char interpolate(unsigned char const& v1, unsigned char const& v2, unsigned char const& v3, float w1, float w2, float w3) {
    return (unsigned char)((float)v1 * w1 + (float)v2 * w2 + (float)v3 * w3 + 0.5f);
}
template <> inline unsigned char interpolate(unsigned char const& v1, unsigned char const& v2, unsigned char const& v3, unsigned char const& v4,
                                             float w1, float w2, float w3, float w4) {
    return (unsigned char)((float)v1 * w1 + (float)v2 * w2 + (float)v3 * w3 + (float)v4 * w4 + 0.5f);
}

}  // namespace math
