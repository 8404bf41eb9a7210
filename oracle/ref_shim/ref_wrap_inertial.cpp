// TEST INFRASTRUCTURE ONLY -- C entry points over the REFERENCE's own inertial g2o types: the bodies of
//   ImuCamPose::Project / isDepthPositive / Update                      src/G2oTypes.cc:170-175, 187-190, 192-220
//   EdgeMono::linearizeOplus, EdgeMonoOnlyPose::linearizeOplus          :349-395
//   EdgeInertial::computeError / linearizeOplus                         :514-594
//   EdgePriorPoseImu::computeError / linearizeOplus                     :731-760
//   ExpSO3 / LogSO3 / InverseRightJacobianSO3 / RightJacobianSO3 / Skew :777-861
//   Pinhole::project(Vector3d) / projectJac                             src/CameraModels/Pinhole.cpp:35-41, 71-81
//   EdgeSE3ProjectXYZOnlyPose / EdgeSE3ProjectXYZ::linearizeOplus       src/OptimizableTypes.cpp:49-63, 139-160
//   g2o::RobustKernelHuber::robustify                                   Thirdparty/g2o/g2o/core/robust_kernel_impl.cpp:78-91
//   g2o::SE3Quat::operator* / map / exp / normalizeRotation, skew       Thirdparty/g2o/g2o/types/se3quat.h:104-110, 217-220, 223-261, 284-289, se3_ops.hpp:27-38
// are cut out of the reference at build time (extract_ranges.py -> oracle/_ref/gen/*.inc) and compiled as they are against mini_eigen.hpp and the
// class shells below (exactly the members those bodies touch, under the reference's names: include/G2oTypes.h).  What is NOT the reference's code:
// IMU::Preintegrated::GetDelta* (float; Sophus::SO3f::exp and Eigen::JacobiSVD are unavailable -- they forward to the oracle's restatement) and
// NormalizeRotation (Eigen::JacobiSVD -- forwards to the oracle's polar factor).  Everything in double around them is the reference's text.
#include <cmath>
#include <cstring>
#include <vector>

#include "mini_eigen.hpp"
namespace Eigen = EigenMini;
namespace Sophus = SophusMini;

extern "C" {
void orbo_imu_delta(const float* P, const double* bg, const double* ba, double* dR9, double* dV3, double* dP3);
void orbo_so3(int what, const double* in, double* out);
}
enum { P_DT = 0, P_DR = 1, P_DV = 10, P_DP = 13, P_JRG = 16, P_JVG = 25, P_JVA = 34, P_JPG = 43, P_JPA = 52, P_B = 61, P_C = 67 };

using namespace std;

namespace ORB_SLAM3 {

namespace IMU {
const float GRAVITY_VALUE = 9.81;                      // include/ImuTypes.h:40
class Bias {
public:
    Bias() : bax(0), bay(0), baz(0), bwx(0), bwy(0), bwz(0) {}
    Bias(const float& b_acc_x, const float& b_acc_y, const float& b_acc_z, const float& b_ang_vel_x, const float& b_ang_vel_y, const float& b_ang_vel_z)
        : bax(b_acc_x), bay(b_acc_y), baz(b_acc_z), bwx(b_ang_vel_x), bwy(b_ang_vel_y), bwz(b_ang_vel_z) {}
    float bax, bay, baz, bwx, bwy, bwz;
};
class Preintegrated {                                   // only what EdgeInertial calls; the float terms come from the oracle (see the header)
public:
    const float* P;
    Bias b;
    Bias GetDeltaBias(const Bias& b_) { return Bias(b_.bax - b.bax, b_.bay - b.bay, b_.baz - b.baz, b_.bwx - b.bwx, b_.bwy - b.bwy, b_.bwz - b.bwz); }   // src/ImuTypes.cc:276-280
    void delta(const Bias& b_, double* dR, double* dV, double* dP) {
        const double bg[3] = {b_.bwx, b_.bwy, b_.bwz}, ba[3] = {b_.bax, b_.bay, b_.baz};
        orbo_imu_delta(P, bg, ba, dR, dV, dP);
    }
    Eigen::Matrix3f GetDeltaRotation(const Bias& b_) { double dR[9], dV[3], dP[3]; delta(b_, dR, dV, dP); Eigen::Matrix3f r; for (int i = 0; i < 9; ++i) r.m[i] = (float)dR[i]; return r; }
    Eigen::Vector3f GetDeltaVelocity(const Bias& b_) { double dR[9], dV[3], dP[3]; delta(b_, dR, dV, dP); Eigen::Vector3f r; for (int i = 0; i < 3; ++i) r.m[i] = (float)dV[i]; return r; }
    Eigen::Vector3f GetDeltaPosition(const Bias& b_) { double dR[9], dV[3], dP[3]; delta(b_, dR, dV, dP); Eigen::Vector3f r; for (int i = 0; i < 3; ++i) r.m[i] = (float)dP[i]; return r; }
};
}  // namespace IMU

template <typename T = double> Eigen::Matrix<T, 3, 3> NormalizeRotation(const Eigen::Matrix<T, 3, 3>& R) {   // include/G2oTypes.h:67-71 (JacobiSVD: stand-in)
    Eigen::Matrix<T, 3, 3> out;
    orbo_so3(4, R.m, out.m);
    return out;
}
// include/G2oTypes.h:55-65
Eigen::Matrix3d ExpSO3(const double x, const double y, const double z);
Eigen::Matrix3d ExpSO3(const Eigen::Vector3d& w);
Eigen::Vector3d LogSO3(const Eigen::Matrix3d& R);
Eigen::Matrix3d InverseRightJacobianSO3(const Eigen::Vector3d& v);
Eigen::Matrix3d RightJacobianSO3(const Eigen::Vector3d& v);
Eigen::Matrix3d RightJacobianSO3(const double x, const double y, const double z);
Eigen::Matrix3d Skew(const Eigen::Vector3d& w);
Eigen::Matrix3d InverseRightJacobianSO3(const double x, const double y, const double z);

class GeometricCamera {
public:
    virtual ~GeometricCamera() {}
    virtual Eigen::Vector2d project(const Eigen::Vector3d& v3D) = 0;
    virtual Eigen::Matrix<double, 2, 3> projectJac(const Eigen::Vector3d& v3D) = 0;
    std::vector<float> mvParameters;
};
class Pinhole : public GeometricCamera {
public:
    Eigen::Vector2d project(const Eigen::Vector3d& v3D);
    Eigen::Matrix<double, 2, 3> projectJac(const Eigen::Vector3d& v3D);
};
#include "pinhole_project_d.inc"
#include "pinhole_project_jac.inc"

class ImuCamPose {                                      // include/G2oTypes.h:74-113
public:
    ImuCamPose() : its(0) {}
    void Update(const double* pu);
    Eigen::Vector2d Project(const Eigen::Vector3d& Xw, int cam_idx = 0) const;
    bool isDepthPositive(const Eigen::Vector3d& Xw, int cam_idx = 0) const;
    Eigen::Vector3d twb;
    Eigen::Matrix3d Rwb;
    std::vector<Eigen::Matrix3d> Rcw, Rcb, Rbc;
    std::vector<Eigen::Vector3d> tcw, tcb, tbc;
    double bf;
    std::vector<GeometricCamera*> pCamera;
    int its;
};
#include "g2o_imucampose_project.inc"
#include "g2o_imucampose_depth.inc"
#include "g2o_imucampose_update.inc"

struct VertexBase { virtual ~VertexBase() {} };
class VertexPose : public VertexBase { public: ImuCamPose _estimate; const ImuCamPose& estimate() const { return _estimate; } };
template <class V> class Vertex3 : public VertexBase { public: Eigen::Vector3d _estimate; const Eigen::Vector3d& estimate() const { return _estimate; } };
class VertexVelocity : public Vertex3<VertexVelocity> {};
class VertexGyroBias : public Vertex3<VertexGyroBias> {};
class VertexAccBias : public Vertex3<VertexAccBias> {};
}  // namespace ORB_SLAM3
namespace g2o {
class VertexSBAPointXYZ : public ORB_SLAM3::Vertex3<VertexSBAPointXYZ> {};
using Eigen::Matrix3d; using Eigen::Vector3d; using Eigen::Quaterniond; using Eigen::Matrix;
typedef Eigen::Matrix<double, 6, 1> Vector6d;
#include "g2o_se3_skew.inc"
class SE3Quat {                                         // Thirdparty/g2o/g2o/types/se3quat.h: the members the bodies below touch; the bodies are the reference's
protected:
    Quaterniond _r;
    Vector3d _t;
public:
    SE3Quat() {}
    SE3Quat(const Quaterniond& q, const Vector3d& t) : _r(q), _t(t) { normalizeRotation(); }   // :62-64
    inline const Vector3d& translation() const { return _t; }
    inline const Quaterniond& rotation() const { return _r; }
#include "g2o_se3quat_mul.inc"
#include "g2o_se3quat_map.inc"
#include "g2o_se3quat_exp.inc"
#include "g2o_se3quat_normalize.inc"
};
class VertexSE3Expmap : public ORB_SLAM3::VertexBase {   // types_six_dof_expmap.h:55-77
public:
    SE3Quat _estimate;
    const SE3Quat& estimate() const { return _estimate; }
    void setEstimate(const SE3Quat& e) { _estimate = e; }
    void oplusImpl(const double* update_) {              // :73-76 (Eigen::Map replaced by a copy)
        Vector6d update; for (int i = 0; i < 6; ++i) update[i] = update_[i];
        setEstimate(SE3Quat::exp(update) * estimate());
    }
};
class RobustKernelHuber {                               // core/robust_kernel_impl.h:54-67; setDelta as robust_kernel_impl.cpp:70-76
public:
    void setDelta(double delta) { dsqr = delta * delta; _delta = delta; }
    void robustify(double e, Eigen::Vector3d& rho) const;
    double _delta, dsqr;
};
#include "g2o_huber.inc"
}  // namespace g2o
namespace ORB_SLAM3 {

class EdgeMono {                                        // include/G2oTypes.h:342-385
public:
    EdgeMono(int cam_idx_ = 0) : cam_idx(cam_idx_) {}
    void computeError() {                               // :353-358 (header-inline in the reference)
        const g2o::VertexSBAPointXYZ* VPoint = static_cast<const g2o::VertexSBAPointXYZ*>(_vertices[0]);
        const VertexPose* VPose = static_cast<const VertexPose*>(_vertices[1]);
        const Eigen::Vector2d obs(_measurement);
        _error = obs - VPose->estimate().Project(VPoint->estimate(), cam_idx);
    }
    void linearizeOplus();
    std::vector<VertexBase*> _vertices;
    Eigen::Vector2d _measurement, _error;
    Eigen::Matrix<double, 2, 3> _jacobianOplusXi;
    Eigen::Matrix<double, 2, 6> _jacobianOplusXj;
    const int cam_idx;
};
class EdgeMonoOnlyPose {                                // include/G2oTypes.h:389-420
public:
    EdgeMonoOnlyPose(const Eigen::Vector3d& Xw_, int cam_idx_ = 0) : Xw(Xw_), cam_idx(cam_idx_) {}
    void linearizeOplus();
    std::vector<VertexBase*> _vertices;
    Eigen::Matrix<double, 2, 6> _jacobianOplusXi;
    const Eigen::Vector3d Xw;
    const int cam_idx;
};
#include "g2o_edge_mono.inc"

class EdgeSE3ProjectXYZOnlyPose {                      // include/OptimizableTypes.h:31-57
public:
    void computeError() {                               // :41-45 (header-inline in the reference)
        const g2o::VertexSE3Expmap* v1 = static_cast<const g2o::VertexSE3Expmap*>(_vertices[0]);
        Eigen::Vector2d obs(_measurement);
        _error = obs - pCamera->project(v1->estimate().map(Xw));
    }
    bool isDepthPositive() { const g2o::VertexSE3Expmap* v1 = static_cast<const g2o::VertexSE3Expmap*>(_vertices[0]); return (v1->estimate().map(Xw))(2) > 0.0; }   // :47-50
    void linearizeOplus();
    std::vector<VertexBase*> _vertices;
    Eigen::Vector2d _measurement, _error;
    Eigen::Matrix<double, 2, 6> _jacobianOplusXi;
    Eigen::Vector3d Xw;
    GeometricCamera* pCamera;
};
class EdgeSE3ProjectXYZ {                               // include/OptimizableTypes.h:89-115
public:
    void computeError() {                               // :99-104
        const g2o::VertexSE3Expmap* v1 = static_cast<const g2o::VertexSE3Expmap*>(_vertices[1]);
        const g2o::VertexSBAPointXYZ* v2 = static_cast<const g2o::VertexSBAPointXYZ*>(_vertices[0]);
        Eigen::Vector2d obs(_measurement);
        _error = obs - pCamera->project(v1->estimate().map(v2->estimate()));
    }
    void linearizeOplus();
    std::vector<VertexBase*> _vertices;
    Eigen::Vector2d _measurement, _error;
    Eigen::Matrix<double, 2, 3> _jacobianOplusXi;
    Eigen::Matrix<double, 2, 6> _jacobianOplusXj;
    GeometricCamera* pCamera;
};
#include "edge_se3_only_pose.inc"
#include "edge_se3_xyz.inc"

class EdgeInertial {                                    // include/G2oTypes.h:488-560
public:
    EdgeInertial(IMU::Preintegrated* pInt, const float* P) : mpInt(pInt), dt((double)P[P_DT]) {
        for (int i = 0; i < 9; ++i) { JRg.m[i] = (double)P[P_JRG + i]; JVg.m[i] = (double)P[P_JVG + i]; JPg.m[i] = (double)P[P_JPG + i]; JVa.m[i] = (double)P[P_JVA + i]; JPa.m[i] = (double)P[P_JPA + i]; }
        g << 0, 0, -IMU::GRAVITY_VALUE;                 // src/G2oTypes.cc:497
        _jacobianOplus = {Eigen::DynJacobian(9, 6), Eigen::DynJacobian(9, 3), Eigen::DynJacobian(9, 3), Eigen::DynJacobian(9, 3), Eigen::DynJacobian(9, 6), Eigen::DynJacobian(9, 3)};
    }
    void computeError();
    void linearizeOplus();
    std::vector<VertexBase*> _vertices;
    Eigen::Matrix<double, 9, 1> _error;
    std::vector<Eigen::DynJacobian> _jacobianOplus;
    Eigen::Matrix3d JRg, JVg, JPg, JVa, JPa;
    IMU::Preintegrated* mpInt;
    const double dt;
    Eigen::Vector3d g;
};
#include "g2o_edge_inertial.inc"

class EdgePriorPoseImu {                                // include/G2oTypes.h:770-800
public:
    EdgePriorPoseImu() { _jacobianOplus = {Eigen::DynJacobian(15, 6), Eigen::DynJacobian(15, 3), Eigen::DynJacobian(15, 3), Eigen::DynJacobian(15, 3)}; }
    void computeError();
    void linearizeOplus();
    std::vector<VertexBase*> _vertices;
    Eigen::Matrix<double, 15, 1> _error;
    std::vector<Eigen::DynJacobian> _jacobianOplus;
    Eigen::Matrix3d Rwb;
    Eigen::Vector3d twb, vwb, bg, ba;
};
#include "g2o_edge_prior.inc"
#include "g2o_so3.inc"

static void set3(Eigen::Vector3d& v, const double* p) { for (int i = 0; i < 3; ++i) v.m[i] = p[i]; }
static void set9(Eigen::Matrix3d& v, const double* p) { for (int i = 0; i < 9; ++i) v.m[i] = p[i]; }
static void make_pose(ImuCamPose& c, Pinhole* cam, const double* Rwb, const double* twb, const double* extr24) {
    set9(c.Rwb, Rwb); set3(c.twb, twb);
    c.Rcb.resize(1); c.tcb.resize(1); c.Rbc.resize(1); c.tbc.resize(1); c.Rcw.resize(1); c.tcw.resize(1);
    set9(c.Rcb[0], extr24); set3(c.tcb[0], extr24 + 9); set9(c.Rbc[0], extr24 + 12); set3(c.tbc[0], extr24 + 21);
    c.pCamera.assign(1, cam);
    const Eigen::Matrix3d Rbw = c.Rwb.transpose();      // what ImuCamPose::Update derives (:213-221); the oracle's edge functions take the pose in this form
    const Eigen::Vector3d tbw = -Rbw * c.twb;
    c.Rcw[0] = c.Rcb[0] * Rbw; c.tcw[0] = c.Rcb[0] * tbw + c.tcb[0];
}

}  // namespace ORB_SLAM3

using namespace ORB_SLAM3;

extern "C" {

void ref_so3(int what, const double* in, double* out) {   // 0 Exp, 1 Log, 2 RightJacobian, 3 InverseRightJacobian, 5 Skew
    Eigen::Vector3d v; Eigen::Matrix3d M;
    if (what == 1) { set9(M, in); const Eigen::Vector3d w = LogSO3(M); for (int i = 0; i < 3; ++i) out[i] = w.m[i]; return; }
    set3(v, in);
    const Eigen::Matrix3d R = what == 0 ? ExpSO3(v) : what == 2 ? RightJacobianSO3(v) : what == 3 ? InverseRightJacobianSO3(v) : Skew(v);
    for (int i = 0; i < 9; ++i) out[i] = R.m[i];
}

// EdgeInertial::computeError + linearizeOplus; J [9][24]: pose 1 (6) | v1 | gyro bias | acc bias | pose 2 (6) | v2
void ref_edge_inertial(const float* P, const double* Rwb1, const double* twb1, const double* v1, const double* bg, const double* ba, const double* Rwb2, const double* twb2,
                       const double* v2, double* err9, double* J9x24) {
    IMU::Preintegrated pint; pint.P = P; pint.b = IMU::Bias(P[P_B], P[P_B + 1], P[P_B + 2], P[P_B + 3], P[P_B + 4], P[P_B + 5]);
    VertexPose VP1, VP2; VertexVelocity VV1, VV2; VertexGyroBias VG; VertexAccBias VA;
    set9(VP1._estimate.Rwb, Rwb1); set3(VP1._estimate.twb, twb1); set9(VP2._estimate.Rwb, Rwb2); set3(VP2._estimate.twb, twb2);
    set3(VV1._estimate, v1); set3(VV2._estimate, v2); set3(VG._estimate, bg); set3(VA._estimate, ba);
    EdgeInertial e(&pint, P);
    e._vertices = {&VP1, &VV1, &VG, &VA, &VP2, &VV2};
    e.computeError();
    for (int i = 0; i < 9; ++i) err9[i] = e._error.m[i];
    if (!J9x24) return;
    e.linearizeOplus();
    const int col0[6] = {0, 6, 9, 12, 15, 21};
    for (int v = 0; v < 6; ++v) for (int r = 0; r < 9; ++r) for (int c = 0; c < e._jacobianOplus[v].cols; ++c) J9x24[r * 24 + col0[v] + c] = e._jacobianOplus[v](r, c);
}

// EdgeMono: error, Jacobians wrt the point (2 x 3) and the pose (2 x 6), isDepthPositive; and EdgeMonoOnlyPose's 2 x 6 Jacobian
void ref_edge_mono(const double* Rwb, const double* twb, const double* extr24, const float* cam4, const double* Xw, const double* obs, double* err2, double* Jpoint, double* Jpose,
                   double* JposeOnly, int* depthPositive) {
    Pinhole cam; cam.mvParameters.assign(cam4, cam4 + 4);
    VertexPose VP; make_pose(VP._estimate, &cam, Rwb, twb, extr24);
    g2o::VertexSBAPointXYZ VX; set3(VX._estimate, Xw);
    EdgeMono e(0);
    e._vertices = {&VX, &VP};
    e._measurement[0] = obs[0]; e._measurement[1] = obs[1];
    e.computeError();
    err2[0] = e._error[0]; err2[1] = e._error[1];
    e.linearizeOplus();
    for (int i = 0; i < 6; ++i) Jpoint[i] = e._jacobianOplusXi.m[i];
    for (int i = 0; i < 12; ++i) Jpose[i] = e._jacobianOplusXj.m[i];
    EdgeMonoOnlyPose eo(VX._estimate, 0);
    eo._vertices = {&VP};
    eo.linearizeOplus();
    for (int i = 0; i < 12; ++i) JposeOnly[i] = eo._jacobianOplusXi.m[i];
    *depthPositive = VP._estimate.isDepthPositive(VX._estimate, 0) ? 1 : 0;
}

// ImuCamPose::Update: Rwb / twb in/out, the derived camera pose out; `times` consecutive updates with the same pu
void ref_pose_update(double* Rwb, double* twb, const double* extr24, const double* pu, int times, double* Rcw, double* tcw) {
    Pinhole cam; cam.mvParameters.assign(4, 1.f);
    ImuCamPose c; make_pose(c, &cam, Rwb, twb, extr24);
    for (int k = 0; k < times; ++k) c.Update(pu);
    for (int i = 0; i < 9; ++i) { Rwb[i] = c.Rwb.m[i]; Rcw[i] = c.Rcw[0].m[i]; }
    for (int i = 0; i < 3; ++i) { twb[i] = c.twb.m[i]; tcw[i] = c.tcw[0].m[i]; }
}

// EdgePriorPoseImu: e [15], J [15][15] over (pose 6 | v | bg | ba); prior21 / st21: Rwb 9 | twb 3 | v 3 | bg 3 | ba 3
void ref_edge_prior(const double* prior21, const double* st21, double* e15, double* J225) {
    VertexPose VP; VertexVelocity VV; VertexGyroBias VG; VertexAccBias VA;
    set9(VP._estimate.Rwb, st21); set3(VP._estimate.twb, st21 + 9); set3(VV._estimate, st21 + 12); set3(VG._estimate, st21 + 15); set3(VA._estimate, st21 + 18);
    EdgePriorPoseImu e;
    set9(e.Rwb, prior21); set3(e.twb, prior21 + 9); set3(e.vwb, prior21 + 12); set3(e.bg, prior21 + 15); set3(e.ba, prior21 + 18);
    e._vertices = {&VP, &VV, &VG, &VA};
    e.computeError(); e.linearizeOplus();
    for (int i = 0; i < 15; ++i) e15[i] = e._error.m[i];
    const int col0[4] = {0, 6, 9, 12};
    for (int i = 0; i < 225; ++i) J225[i] = 0;
    for (int v = 0; v < 4; ++v) for (int r = 0; r < 15; ++r) for (int c = 0; c < e._jacobianOplus[v].cols; ++c) J225[r * 15 + col0[v] + c] = e._jacobianOplus[v](r, c);
}

// EdgeSE3ProjectXYZ / EdgeSE3ProjectXYZOnlyPose (LocalBundleAdjustment, PoseOptimization): error, Jacobians, depth sign; pose7 = qw qx qy qz tx ty tz
void ref_lba_edge(const double* pose7, const float* cam4, const double* X3, const double* obs2, double* err2, double* Jpoint, double* Jpose, double* JposeOnly, int* depthPositive) {
    Pinhole cam; cam.mvParameters.assign(cam4, cam4 + 4);
    g2o::VertexSE3Expmap VP; Eigen::Vector3d t; set3(t, pose7 + 4);
    VP._estimate = g2o::SE3Quat(Eigen::Quaterniond(pose7[0], pose7[1], pose7[2], pose7[3]), t);
    g2o::VertexSBAPointXYZ VX; set3(VX._estimate, X3);
    EdgeSE3ProjectXYZ e; e.pCamera = &cam; e._vertices = {&VX, &VP}; e._measurement[0] = obs2[0]; e._measurement[1] = obs2[1];
    e.computeError(); e.linearizeOplus();
    err2[0] = e._error[0]; err2[1] = e._error[1];
    for (int i = 0; i < 6; ++i) Jpoint[i] = e._jacobianOplusXi.m[i];
    for (int i = 0; i < 12; ++i) Jpose[i] = e._jacobianOplusXj.m[i];
    EdgeSE3ProjectXYZOnlyPose eo; eo.pCamera = &cam; eo._vertices = {&VP}; eo.Xw = VX._estimate; eo._measurement = e._measurement;
    eo.computeError(); eo.linearizeOplus();
    for (int i = 0; i < 12; ++i) JposeOnly[i] = eo._jacobianOplusXi.m[i];
    *depthPositive = eo.isDepthPositive() ? 1 : 0;
}
void ref_huber(double delta, double e, double* rho3) {
    g2o::RobustKernelHuber k; k.setDelta(delta);
    Eigen::Vector3d rho; k.robustify(e, rho);
    for (int i = 0; i < 3; ++i) rho3[i] = rho[i];
}
// VertexSE3Expmap::oplusImpl = SE3Quat::exp(update) * estimate()
void ref_lba_pose_oplus(double* pose7, const double* update6) {
    g2o::VertexSE3Expmap VP; Eigen::Vector3d t; set3(t, pose7 + 4);
    VP._estimate = g2o::SE3Quat(Eigen::Quaterniond(pose7[0], pose7[1], pose7[2], pose7[3]), t);
    VP.oplusImpl(update6);
    const Eigen::Quaterniond& q = VP._estimate.rotation();
    pose7[0] = q.w(); pose7[1] = q.x(); pose7[2] = q.y(); pose7[3] = q.z();
    for (int i = 0; i < 3; ++i) pose7[4 + i] = VP._estimate.translation()[i];
}

}  // extern "C"
