// refshim: MVE math::Accum stand-in (weighted accumulator; see ../README.md)
#pragma once
#include "math/vector.h"

namespace math {

template <typename T>
class Accum {
public:
    T v;
    float w;
    Accum() : w(0.0f) {}
    explicit Accum(T const& init) : v(init), w(0.0f) {}
    void add(T const& value, float weight) { v += value * weight; w += weight; }
    void sub(T const& value, float weight) { v -= value * weight; w -= weight; }
    T normalized(float weight) const { return v / weight; }
    T normalized() const { return v / w; }
};

}  // namespace math
