// TEST INFRASTRUCTURE ONLY -- the handful of Eigen / Sophus operators behind ref_slam_types.hpp, in the evaluation order
// oracle/matcher_oracle.cpp documents.  Compiled with -ffp-contract=off (oracle/Makefile): each operation rounds once.
#include "ref_slam_types.hpp"

namespace Eigen {
Vector3f Vector3f::operator+(const Vector3f& o) const { return Vector3f(v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]); }
Vector3f Vector3f::operator-(const Vector3f& o) const { return Vector3f(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
float Vector3f::dot(const Vector3f& o) const { return (v[0] * o.v[0] + v[1] * o.v[1]) + v[2] * o.v[2]; }
float Vector3f::norm() const { return sqrtf((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]); }
Vector3f Matrix3f::operator*(const Vector3f& p) const {
    return Vector3f((m[0] * p.v[0] + m[1] * p.v[1]) + m[2] * p.v[2], (m[3] * p.v[0] + m[4] * p.v[1]) + m[5] * p.v[2],
                    (m[6] * p.v[0] + m[7] * p.v[1]) + m[8] * p.v[2]);
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
}  // namespace Sophus
