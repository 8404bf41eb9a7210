// TEST INFRASTRUCTURE ONLY -- a small fixed-size dense matrix type with the part of Eigen's / Sophus' interface that the bodies of the reference's
// inertial g2o types use (src/G2oTypes.cc, src/CameraModels/Pinhole.cpp: comma initialisers, +, -, *, / with scalars, products, transpose, cast,
// Identity, setZero, block<3,3>(r, c) = ..., operator() / []), so that those bodies compile verbatim into oracle/_ref/libref_inertial.so without Eigen.
// Every operation is evaluated eagerly, row by row, left to right: Eigen's expression templates may sum in another order, so what this pins is the
// reference's FORMULAS (signs, indices, terms) to rounding (1e-12 relative), not its last bit.  Lives in its own namespaces (aliased to Eigen / Sophus
// only inside ref_wrap_inertial.cpp) and its own shared library, apart from the float stand-ins of ref_slam_types.hpp.
#pragma once
#include <cmath>
#include <type_traits>

namespace EigenMini {

template <typename T, int R, int C> struct Matrix;

template <typename T, int R, int C> struct CommaInit {
    Matrix<T, R, C>& m; int k;
    CommaInit& operator,(T v) { m.m[k++] = v; return *this; }
    template <int N> CommaInit& operator,(const Matrix<T, N, 1>& v) { for (int i = 0; i < N; ++i) m.m[k++] = v.m[i]; return *this; }
};
template <typename T, int N> struct DiagonalMatrix;
template <typename M, int BR, int BC> struct BlockRef {
    M& m; int r0, c0;
    Matrix<typename M::Scalar, BR, BC> eval() const { Matrix<typename M::Scalar, BR, BC> o; for (int i = 0; i < BR; ++i) for (int j = 0; j < BC; ++j) o(i, j) = m(r0 + i, c0 + j); return o; }
    template <typename T> BlockRef& operator=(const DiagonalMatrix<T, BR>& d) { for (int i = 0; i < BR; ++i) for (int j = 0; j < BC; ++j) m(r0 + i, c0 + j) = i == j ? d.d[i] : 0; return *this; }
    template <typename T, int N> BlockRef& operator+=(const DiagonalMatrix<T, N>& d) { for (int i = 0; i < N; ++i) m(r0 + i, c0 + i) += d.d[i]; return *this; }
    template <typename T> BlockRef& operator=(const Matrix<T, BR, BC>& v) { for (int i = 0; i < BR; ++i) for (int j = 0; j < BC; ++j) m(r0 + i, c0 + j) = v(i, j); return *this; }
};

template <typename T, int R, int C> struct Matrix {
    typedef T Scalar;
    T m[R * C];   // row-major
    Matrix() { for (int i = 0; i < R * C; ++i) m[i] = 0; }
    static Matrix Identity() { Matrix r; for (int i = 0; i < (R < C ? R : C); ++i) r.m[i * C + i] = 1; return r; }
    static Matrix Zero() { return Matrix(); }
    T& operator()(int i, int j) { return m[i * C + j]; }
    const T& operator()(int i, int j) const { return m[i * C + j]; }
    T& operator()(int i) { return m[i]; }
    const T& operator()(int i) const { return m[i]; }
    T& operator[](int i) { return m[i]; }
    const T& operator[](int i) const { return m[i]; }
    void setZero() { for (int i = 0; i < R * C; ++i) m[i] = 0; }
    void fill(T v) { for (int i = 0; i < R * C; ++i) m[i] = v; }
    void setIdentity() { *this = Identity(); }
    Matrix<T, C, R> transpose() const { Matrix<T, C, R> r; for (int i = 0; i < R; ++i) for (int j = 0; j < C; ++j) r.m[j * R + i] = m[i * C + j]; return r; }
    template <typename U> Matrix<U, R, C> cast() const { Matrix<U, R, C> r; for (int i = 0; i < R * C; ++i) r.m[i] = (U)m[i]; return r; }
    Matrix<T, 1, C> row(int i) const { Matrix<T, 1, C> r; for (int j = 0; j < C; ++j) r.m[j] = m[i * C + j]; return r; }
    template <int RR = R, int CC = C, typename = typename std::enable_if<RR == 1 && CC == 1>::type> operator T() const { return m[0]; }   // a 1 x 1 product is a scalar
    T norm() const { T s = 0; for (int i = 0; i < R * C; ++i) s += m[i] * m[i]; return std::sqrt(s); }
    CommaInit<T, R, C> operator<<(T v) { m[0] = v; return CommaInit<T, R, C>{*this, 1}; }
    template <int N> CommaInit<T, R, C> operator<<(const Matrix<T, N, 1>& v) { for (int i = 0; i < N; ++i) m[i] = v.m[i]; return CommaInit<T, R, C>{*this, N}; }
    template <int BR, int BC> BlockRef<Matrix, BR, BC> block(int r, int c) { return BlockRef<Matrix, BR, BC>{*this, r, c}; }
    template <int BR, int BC> Matrix<T, BR, BC> blockCopy(int r, int c) const { Matrix<T, BR, BC> o; for (int i = 0; i < BR; ++i) for (int j = 0; j < BC; ++j) o.m[i * BC + j] = m[(r + i) * C + c + j]; return o; }
    Matrix& operator+=(const Matrix& o) { for (int i = 0; i < R * C; ++i) m[i] += o.m[i]; return *this; }
    Matrix& operator-=(const Matrix& o) { for (int i = 0; i < R * C; ++i) m[i] -= o.m[i]; return *this; }
};

template <typename T, int N> struct DiagonalMatrix {
    T d[N];
    DiagonalMatrix() { for (int i = 0; i < N; ++i) d[i] = 0; }
    DiagonalMatrix(T a, T b, T c) { static_assert(N == 3, "three-argument form"); d[0] = a; d[1] = b; d[2] = c; }
    Matrix<T, N, 1>& diagonal() { return *reinterpret_cast<Matrix<T, N, 1>*>(d); }
};
template <typename T, int R, int N> Matrix<T, R, N> operator*(const Matrix<T, R, N>& a, const DiagonalMatrix<T, N>& d) {
    Matrix<T, R, N> r; for (int i = 0; i < R; ++i) for (int j = 0; j < N; ++j) r.m[i * N + j] = a.m[i * N + j] * d.d[j]; return r;
}
template <typename T, int R, int C> Matrix<T, R, C> operator+(const Matrix<T, R, C>& a, const Matrix<T, R, C>& b) { Matrix<T, R, C> r; for (int i = 0; i < R * C; ++i) r.m[i] = a.m[i] + b.m[i]; return r; }
template <typename T, int R, int C> Matrix<T, R, C> operator-(const Matrix<T, R, C>& a, const Matrix<T, R, C>& b) { Matrix<T, R, C> r; for (int i = 0; i < R * C; ++i) r.m[i] = a.m[i] - b.m[i]; return r; }
template <typename T, int R, int C> Matrix<T, R, C> operator-(const Matrix<T, R, C>& a) { Matrix<T, R, C> r; for (int i = 0; i < R * C; ++i) r.m[i] = -a.m[i]; return r; }
template <typename T, int R, int K, int C> Matrix<T, R, C> operator*(const Matrix<T, R, K>& a, const Matrix<T, K, C>& b) {
    Matrix<T, R, C> r;
    for (int i = 0; i < R; ++i) for (int j = 0; j < C; ++j) { T s = 0; for (int k = 0; k < K; ++k) s += a.m[i * K + k] * b.m[k * C + j]; r.m[i * C + j] = s; }
    return r;
}
template <typename T, int R, int K, typename M, int C> Matrix<T, R, C> operator*(const Matrix<T, R, K>& a, const BlockRef<M, K, C>& b) { return a * b.eval(); }
template <typename T, int R, int C, typename U, typename = typename std::enable_if<std::is_arithmetic<U>::value>::type>
Matrix<T, R, C> operator*(const Matrix<T, R, C>& a, U s) { Matrix<T, R, C> r; for (int i = 0; i < R * C; ++i) r.m[i] = a.m[i] * (T)s; return r; }
template <typename T, int R, int C, typename U, typename = typename std::enable_if<std::is_arithmetic<U>::value>::type>
Matrix<T, R, C> operator*(U s, const Matrix<T, R, C>& a) { Matrix<T, R, C> r; for (int i = 0; i < R * C; ++i) r.m[i] = (T)s * a.m[i]; return r; }
template <typename T, int R, int C, typename U, typename = typename std::enable_if<std::is_arithmetic<U>::value>::type>
Matrix<T, R, C> operator/(const Matrix<T, R, C>& a, U s) { Matrix<T, R, C> r; for (int i = 0; i < R * C; ++i) r.m[i] = a.m[i] / (T)s; return r; }

template <typename T, typename U, typename = typename std::enable_if<std::is_arithmetic<U>::value>::type> T operator+(const Matrix<T, 1, 1>& a, U s) { return a.m[0] + (T)s; }   // (row * vector) + scalar

typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<float, 3, 3> Matrix3f;
typedef Matrix<float, 3, 1> Vector3f;

// Eigen::Quaternion<double>, the operations g2o's SE3Quat uses; algorithms as in Eigen/src/Geometry/Quaternion.h (quaternion from a rotation matrix: Shepperd's
// branches; v' = v + w t + q x t with t = 2 q x v; Hamilton product; toRotationMatrix)
struct Coeffs4 { double* p; Coeffs4& operator*=(double s) { for (int i = 0; i < 4; ++i) p[i] *= s; return *this; } };
struct Quaterniond {
    double c[4];   // x, y, z, w (Eigen's coefficient order)
    Quaterniond() : c{0, 0, 0, 1} {}
    Quaterniond(double w, double x, double y, double z) : c{x, y, z, w} {}
    explicit Quaterniond(const Matrix<double, 3, 3>& mat) {
        double t = mat(0, 0) + mat(1, 1) + mat(2, 2);
        if (t > 0) {
            t = std::sqrt(t + 1.0);
            c[3] = 0.5 * t; t = 0.5 / t;
            c[0] = (mat(2, 1) - mat(1, 2)) * t; c[1] = (mat(0, 2) - mat(2, 0)) * t; c[2] = (mat(1, 0) - mat(0, 1)) * t;
        } else {
            int i = 0;
            if (mat(1, 1) > mat(0, 0)) i = 1;
            if (mat(2, 2) > mat(i, i)) i = 2;
            const int j = (i + 1) % 3, k = (j + 1) % 3;
            t = std::sqrt(mat(i, i) - mat(j, j) - mat(k, k) + 1.0);
            c[i] = 0.5 * t; t = 0.5 / t;
            c[3] = (mat(k, j) - mat(j, k)) * t;
            c[j] = (mat(j, i) + mat(i, j)) * t;
            c[k] = (mat(k, i) + mat(i, k)) * t;
        }
    }
    double w() const { return c[3]; } double x() const { return c[0]; } double y() const { return c[1]; } double z() const { return c[2]; }
    Coeffs4 coeffs() { return Coeffs4{c}; }
    void normalize() { const double n = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2] + c[3] * c[3]); for (double& v : c) v /= n; }
    Matrix<double, 3, 1> operator*(const Matrix<double, 3, 1>& v) const {
        Matrix<double, 3, 1> uv, r;
        uv[0] = c[1] * v[2] - c[2] * v[1]; uv[1] = c[2] * v[0] - c[0] * v[2]; uv[2] = c[0] * v[1] - c[1] * v[0];
        uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
        r[0] = v[0] + c[3] * uv[0] + (c[1] * uv[2] - c[2] * uv[1]);
        r[1] = v[1] + c[3] * uv[1] + (c[2] * uv[0] - c[0] * uv[2]);
        r[2] = v[2] + c[3] * uv[2] + (c[0] * uv[1] - c[1] * uv[0]);
        return r;
    }
    Quaterniond operator*(const Quaterniond& b) const {
        return Quaterniond(w() * b.w() - x() * b.x() - y() * b.y() - z() * b.z(), w() * b.x() + x() * b.w() + y() * b.z() - z() * b.y(),
                           w() * b.y() + y() * b.w() + z() * b.x() - x() * b.z(), w() * b.z() + z() * b.w() + x() * b.y() - y() * b.x());
    }
    Quaterniond& operator*=(const Quaterniond& b) { *this = *this * b; return *this; }
    Matrix<double, 3, 3> toRotationMatrix() const {
        Matrix<double, 3, 3> res;
        const double tx = 2 * x(), ty = 2 * y(), tz = 2 * z(), twx = tx * w(), twy = ty * w(), twz = tz * w(), txx = tx * x(), txy = ty * x(), txz = tz * x(), tyy = ty * y(), tyz = tz * y(), tzz = tz * z();
        res(0, 0) = 1 - (tyy + tzz); res(0, 1) = txy - twz; res(0, 2) = txz + twy;
        res(1, 0) = txy + twz; res(1, 1) = 1 - (txx + tzz); res(1, 2) = tyz - twx;
        res(2, 0) = txz - twy; res(2, 1) = tyz + twx; res(2, 2) = 1 - (txx + tyy);
        return res;
    }
};

// g2o's per-vertex Jacobian blocks are dynamic-size maps; the bodies only call setZero() and block<3,3>(r, c) = ...
struct DynJacobian {
    typedef double Scalar;
    int rows, cols; double m[15 * 6];
    DynJacobian(int r = 0, int c = 0) : rows(r), cols(c) { for (double& v : m) v = 0; }
    void setZero() { for (double& v : m) v = 0; }
    double& operator()(int i, int j) { return m[i * cols + j]; }
    template <int BR, int BC> BlockRef<DynJacobian, BR, BC> block(int r, int c) { return BlockRef<DynJacobian, BR, BC>{*this, r, c}; }
};

}  // namespace EigenMini

namespace SophusMini {
template <typename T> struct SO3 {
    static EigenMini::Matrix<T, 3, 3> hat(const EigenMini::Matrix<T, 3, 1>& w) {   // so3.hpp:631-640
        EigenMini::Matrix<T, 3, 3> W;
        W(0, 1) = -w[2]; W(0, 2) = w[1]; W(1, 0) = w[2]; W(1, 2) = -w[0]; W(2, 0) = -w[1]; W(2, 1) = w[0];
        return W;
    }
};
typedef SO3<double> SO3d;
}  // namespace SophusMini
