// TEST INFRASTRUCTURE ONLY -- the handful of Eigen / Sophus operators behind ref_slam_types.hpp, in the evaluation order
// oracle/matcher_oracle.cpp documents.  Compiled with -ffp-contract=off (oracle/Makefile): each operation rounds once.
#include "ref_slam_types.hpp"

namespace Eigen {
Vector3f Vector3f::operator+(const Vector3f& o) const { return Vector3f(v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]); }
Vector3f Vector3f::operator-(const Vector3f& o) const { return Vector3f(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
Vector3f Vector3f::operator/(float d) const { return Vector3f(v[0] / d, v[1] / d, v[2] / d); }
float Vector3f::dot(const Vector3f& o) const { return (v[0] * o.v[0] + v[1] * o.v[1]) + v[2] * o.v[2]; }
float Vector3f::norm() const { return sqrtf((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]); }
Vector3f Matrix3f::operator*(const Vector3f& p) const {
    return Vector3f((m[0] * p.v[0] + m[1] * p.v[1]) + m[2] * p.v[2], (m[3] * p.v[0] + m[4] * p.v[1]) + m[5] * p.v[2],
                    (m[6] * p.v[0] + m[7] * p.v[1]) + m[8] * p.v[2]);
}
Matrix3f Matrix3f::operator*(const Matrix3f& o) const {
    Matrix3f r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.m[3 * i + j] = (m[3 * i] * o.m[j] + m[3 * i + 1] * o.m[3 + j]) + m[3 * i + 2] * o.m[6 + j];
    return r;
}
Matrix3f Matrix3f::transpose() const {
    Matrix3f r;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[3 * i + j] = m[3 * j + i];
    return r;
}
Matrix3f Matrix3f::inverse() const {   // cofactor (i, j) of the transpose over the determinant expanded along the first row
    const float c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
    const float det = (m[0] * c00 + m[1] * c01) + m[2] * c02;
    const float id = 1.0f / det;
    Matrix3f r;
    r.m[0] = c00 * id; r.m[1] = (m[2] * m[7] - m[1] * m[8]) * id; r.m[2] = (m[1] * m[5] - m[2] * m[4]) * id;
    r.m[3] = c01 * id; r.m[4] = (m[0] * m[8] - m[2] * m[6]) * id; r.m[5] = (m[2] * m[3] - m[0] * m[5]) * id;
    r.m[6] = c02 * id; r.m[7] = (m[1] * m[6] - m[0] * m[7]) * id; r.m[8] = (m[0] * m[4] - m[1] * m[3]) * id;
    return r;
}
}  // namespace Eigen

namespace Sophus {
// Thirdparty/Sophus/sophus/so3.hpp:358-367: uv = q.vec x p; uv += uv; p + w*uv + q.vec x uv
Eigen::Vector3f SE3f::rotate(const Eigen::Vector3f& p) const {
    float ux = qy * p.v[2] - qz * p.v[1], uy = qz * p.v[0] - qx * p.v[2], uz = qx * p.v[1] - qy * p.v[0];
    ux += ux; uy += uy; uz += uz;
    const float cx = qy * uz - qz * uy, cy = qz * ux - qx * uz, cz = qx * uy - qy * ux;
    return Eigen::Vector3f((p.v[0] + qw * ux) + cx, (p.v[1] + qw * uy) + cy, (p.v[2] + qw * uz) + cz);
}
// se3.hpp:321-324
Eigen::Vector3f SE3f::operator*(const Eigen::Vector3f& p) const { return rotate(p) + t; }
SE3f SE3f::inverse() const {
    SE3f r; r.qw = qw; r.qx = -qx; r.qy = -qy; r.qz = -qz;
    Eigen::Vector3f rt = r.rotate(t);
    r.t = Eigen::Vector3f(-rt.v[0], -rt.v[1], -rt.v[2]);
    return r;
}
SE3f::SE3(const Eigen::Matrix3f& R, const Eigen::Vector3f& trans) : t(trans) {     // Eigen::Quaternion from a rotation matrix (quaternionbase_assign_impl)
    float tr = (R(0, 0) + R(1, 1)) + R(2, 2);
    if (tr > 0) {
        tr = sqrtf(tr + 1.0f);
        qw = 0.5f * tr; tr = 0.5f / tr;
        qx = (R(2, 1) - R(1, 2)) * tr; qy = (R(0, 2) - R(2, 0)) * tr; qz = (R(1, 0) - R(0, 1)) * tr;
    } else {
        int i = 0;
        if (R(1, 1) > R(0, 0)) i = 1;
        if (R(2, 2) > R(i, i)) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        tr = sqrtf(((R(i, i) - R(j, j)) - R(k, k)) + 1.0f);
        float q[3];
        q[i] = 0.5f * tr; tr = 0.5f / tr;
        qw = (R(k, j) - R(j, k)) * tr;
        q[j] = (R(j, i) + R(i, j)) * tr; q[k] = (R(k, i) + R(i, k)) * tr;
        qx = q[0]; qy = q[1]; qz = q[2];
    }
}
SE3f SE3f::operator*(const SE3f& o) const {
    SE3f r;
    r.qw = ((qw * o.qw - qx * o.qx) - qy * o.qy) - qz * o.qz;
    r.qx = ((qw * o.qx + qx * o.qw) + qy * o.qz) - qz * o.qy;
    r.qy = ((qw * o.qy + qy * o.qw) + qz * o.qx) - qx * o.qz;
    r.qz = ((qw * o.qz + qz * o.qw) + qx * o.qy) - qy * o.qx;
    r.t = rotate(o.t) + t;
    return r;
}
Eigen::Matrix3f SE3f::rotationMatrix() const {     // Eigen::QuaternionBase::toRotationMatrix
    const float tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;
    const float twx = tx * qw, twy = ty * qw, twz = tz * qw, txx = tx * qx, txy = ty * qx, txz = tz * qx, tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
    Eigen::Matrix3f R;
    R(0, 0) = 1 - (tyy + tzz); R(0, 1) = txy - twz; R(0, 2) = txz + twy;
    R(1, 0) = txy + twz; R(1, 1) = 1 - (txx + tzz); R(1, 2) = tyz - twx;
    R(2, 0) = txz - twy; R(2, 1) = tyz + twx; R(2, 2) = 1 - (txx + tyy);
    return R;
}
static SE3f unit_of(const Sim3f& S) { SE3f r; r.qw = S.qw; r.qx = S.qx; r.qy = S.qy; r.qz = S.qz; return r; }
Eigen::Vector3f Sim3f::operator*(const Eigen::Vector3f& p) const {
    const Eigen::Vector3f r = unit_of(*this).rotate(p);
    return Eigen::Vector3f(s * r.v[0] + t.v[0], s * r.v[1] + t.v[1], s * r.v[2] + t.v[2]);
}
Sim3f Sim3f::inverse() const {
    Sim3f r; r.s = 1.0f / s; r.qw = qw; r.qx = -qx; r.qy = -qy; r.qz = -qz;
    const Eigen::Vector3f rt = unit_of(r).rotate(t);
    r.t = Eigen::Vector3f(-(r.s * rt.v[0]), -(r.s * rt.v[1]), -(r.s * rt.v[2]));
    return r;
}
Eigen::Matrix3f Sim3f::rotationMatrix() const { return unit_of(*this).rotationMatrix(); }
Eigen::Matrix3f SO3f::hat(const Eigen::Vector3f& w) {
    Eigen::Matrix3f O;
    O(0, 0) = 0; O(0, 1) = -w(2); O(0, 2) = w(1); O(1, 0) = w(2); O(1, 1) = 0; O(1, 2) = -w(0); O(2, 0) = -w(1); O(2, 1) = w(0); O(2, 2) = 0;
    return O;
}
}  // namespace Sophus
