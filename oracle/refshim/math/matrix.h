// refshim: MVE math::Matrix stand-in (row major; see ../README.md)
#pragma once
#include "math/vector.h"

namespace math {

template <typename T, int N, int M>
class Matrix {
public:
    T m[N * M];
    Matrix() {}
    explicit Matrix(T const* values) { for (int i = 0; i < N * M; ++i) m[i] = values[i]; }
    explicit Matrix(T const& value) { for (int i = 0; i < N * M; ++i) m[i] = value; }
    T* operator*() { return m; }
    T const* operator*() const { return m; }
    T& operator()(int r, int c) { return m[r * M + c]; }
    T const& operator()(int r, int c) const { return m[r * M + c]; }
    T& operator[](int i) { return m[i]; }
    T const& operator[](int i) const { return m[i]; }
    Matrix& fill(T const& value) { for (int i = 0; i < N * M; ++i) m[i] = value; return *this; }

    // inner product from T(0), left to right
    Vector<T, N> operator*(Vector<T, M> const& rhs) const {
        Vector<T, N> r;
        for (int i = 0; i < N; ++i) {
            T s = T(0);
            for (int k = 0; k < M; ++k) s = s + m[i * M + k] * rhs[k];
            r[i] = s;
        }
        return r;
    }
    // (this * [rhs, v]) without the last row: inner product over M-1 columns, then "+ v * last column"
    Vector<T, N - 1> mult(Vector<T, M - 1> const& rhs, T const& v) const {
        Vector<T, N - 1> r;
        for (int i = 0; i < N - 1; ++i) {
            T s = T(0);
            for (int k = 0; k < M - 1; ++k) s = s + m[i * M + k] * rhs[k];
            r[i] = s + v * m[i * M + M - 1];
        }
        return r;
    }
    template <int U>
    Matrix<T, N, U> operator*(Matrix<T, M, U> const& rhs) const {
        Matrix<T, N, U> r;
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < U; ++j) {
                T s = T(0);
                for (int k = 0; k < M; ++k) s = s + m[i * M + k] * rhs(k, j);
                r(i, j) = s;
            }
        return r;
    }
    Matrix<T, M, N> transposed() const {
        Matrix<T, M, N> r;
        for (int i = 0; i < N; ++i) for (int j = 0; j < M; ++j) r(j, i) = (*this)(i, j);
        return r;
    }
};

typedef Matrix<float, 2, 2> Matrix2f;
typedef Matrix<float, 3, 3> Matrix3f;
typedef Matrix<float, 4, 4> Matrix4f;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 4, 4> Matrix4d;

}  // namespace math
