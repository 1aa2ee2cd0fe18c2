// refshim: MVE math::Vector stand-in (see ../README.md).  Operation order follows the oracle's
// restatement: sums run left to right from T(0); normalisation divides by the norm.
#pragma once
#include <algorithm>
#include <cmath>
#include <cassert>
#include <iostream>
#include <limits>
#include <stdexcept>
#include <string>
#include "math/defines.h"

namespace math {

template <typename T, int N>
class Vector {
public:
    T v[N];
    Vector() {}
    explicit Vector(T const* values) { for (int i = 0; i < N; ++i) v[i] = values[i]; }
    explicit Vector(T const& value) { for (int i = 0; i < N; ++i) v[i] = value; }
    Vector(T const& a, T const& b) { static_assert(N == 2, "dim"); v[0] = a; v[1] = b; }
    Vector(T const& a, T const& b, T const& c) { static_assert(N == 3, "dim"); v[0] = a; v[1] = b; v[2] = c; }
    Vector(T const& a, T const& b, T const& c, T const& d) { static_assert(N == 4, "dim"); v[0] = a; v[1] = b; v[2] = c; v[3] = d; }
    Vector(Vector<T, N - 1> const& o, T const& last) { for (int i = 0; i < N - 1; ++i) v[i] = o.v[i]; v[N - 1] = last; }
    template <typename O> Vector(Vector<O, N> const& o) { for (int i = 0; i < N; ++i) v[i] = static_cast<T>(o.v[i]); }

    Vector& fill(T const& value) { for (int i = 0; i < N; ++i) v[i] = value; return *this; }
    T* begin() { return v; }
    T* end() { return v + N; }
    T const* begin() const { return v; }
    T const* end() const { return v + N; }
    T* operator*() { return v; }
    T const* operator*() const { return v; }
    T& operator[](int i) { return v[i]; }
    T const& operator[](int i) const { return v[i]; }
    T& operator()(int i) { return v[i]; }
    T const& operator()(int i) const { return v[i]; }

    T square_norm() const { T s = T(0); for (int i = 0; i < N; ++i) s = s + v[i] * v[i]; return s; }
    T norm() const { return std::sqrt(square_norm()); }
    Vector& normalize() { T n = norm(); for (int i = 0; i < N; ++i) v[i] = v[i] / n; return *this; }
    Vector normalized() const { return Vector(*this).normalize(); }
    T dot(Vector const& o) const { T s = T(0); for (int i = 0; i < N; ++i) s = s + v[i] * o.v[i]; return s; }
    Vector cross(Vector const& o) const {
        static_assert(N == 3, "dim");
        return Vector(v[1] * o.v[2] - v[2] * o.v[1], v[2] * o.v[0] - v[0] * o.v[2], v[0] * o.v[1] - v[1] * o.v[0]);
    }
    T minimum() const { return *std::min_element(v, v + N); }
    T maximum() const { return *std::max_element(v, v + N); }
    T sum() const { T s = T(0); for (int i = 0; i < N; ++i) s = s + v[i]; return s; }
    Vector cw_mult(Vector const& o) const { Vector r; for (int i = 0; i < N; ++i) r.v[i] = v[i] * o.v[i]; return r; }
    Vector cw_div(Vector const& o) const { Vector r; for (int i = 0; i < N; ++i) r.v[i] = v[i] / o.v[i]; return r; }

    bool operator==(Vector const& o) const { for (int i = 0; i < N; ++i) if (!(v[i] == o.v[i])) return false; return true; }
    bool operator!=(Vector const& o) const { return !(*this == o); }
    Vector operator-() const { Vector r; for (int i = 0; i < N; ++i) r.v[i] = -v[i]; return r; }
    Vector& operator+=(Vector const& o) { for (int i = 0; i < N; ++i) v[i] = v[i] + o.v[i]; return *this; }
    Vector& operator-=(Vector const& o) { for (int i = 0; i < N; ++i) v[i] = v[i] - o.v[i]; return *this; }
    Vector& operator+=(T const& s) { for (int i = 0; i < N; ++i) v[i] = v[i] + s; return *this; }
    Vector& operator-=(T const& s) { for (int i = 0; i < N; ++i) v[i] = v[i] - s; return *this; }
    Vector& operator*=(T const& s) { for (int i = 0; i < N; ++i) v[i] = v[i] * s; return *this; }
    Vector& operator/=(T const& s) { for (int i = 0; i < N; ++i) v[i] = v[i] / s; return *this; }
    Vector operator+(Vector const& o) const { return Vector(*this) += o; }
    Vector operator-(Vector const& o) const { return Vector(*this) -= o; }
    Vector operator+(T const& s) const { return Vector(*this) += s; }
    Vector operator-(T const& s) const { return Vector(*this) -= s; }
    Vector operator*(T const& s) const { return Vector(*this) *= s; }
    Vector operator/(T const& s) const { return Vector(*this) /= s; }
};

template <typename T, int N>
inline Vector<T, N> operator*(T const& s, Vector<T, N> const& v) { return v * s; }

template <typename T, int N>
inline std::ostream& operator<<(std::ostream& os, Vector<T, N> const& v) {
    for (int i = 0; i < N; ++i) os << v[i] << (i + 1 < N ? " " : "");
    return os;
}

typedef Vector<float, 2> Vec2f;
typedef Vector<float, 3> Vec3f;
typedef Vector<float, 4> Vec4f;
typedef Vector<double, 2> Vec2d;
typedef Vector<double, 3> Vec3d;
typedef Vector<double, 4> Vec4d;
typedef Vector<int, 2> Vec2i;
typedef Vector<int, 3> Vec3i;
typedef Vector<unsigned int, 3> Vec3ui;
typedef Vector<unsigned char, 3> Vec3uc;
typedef Vector<unsigned char, 4> Vec4uc;
typedef Vector<std::size_t, 2> Vec2st;
typedef Vector<std::size_t, 3> Vec3st;

}  // namespace math
