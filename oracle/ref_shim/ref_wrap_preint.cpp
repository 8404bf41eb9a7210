// TEST INFRASTRUCTURE ONLY -- C entry point over the REFERENCE's own IMU preintegration: the bodies of
//   IntegratedRotation::IntegratedRotation                 src/ImuTypes.cc:84-105
//   Preintegrated::Initialize                              :147-166
//   Preintegrated::IntegrateNewMeasurement                 :177-236
// are cut out of the reference at build time and compiled as they are against mini_eigen.hpp and the class shells below (the members of
// include/ImuTypes.h those bodies touch).  NOT the reference's code: NormalizeRotation (Eigen::JacobiSVD -> the oracle's polar factor) and the noise
// set-up of Calib::Set (:397-406, four squares written out here: the original needs Sophus::SE3).  The unqualified sqrt / sin / cos of those bodies
// resolve to the float overloads here (using namespace std), the choice the oracle makes as well.
#include <cmath>
#include <iostream>
#include <list>
#include <mutex>
#include <vector>

#include "mini_eigen.hpp"
namespace Eigen = EigenMini;
namespace SophusMini { typedef SO3<float> SO3f; }
namespace Sophus = SophusMini;

extern "C" void orbo_normalize_rotation_f(const float* R, float* out);

using namespace std;

namespace ORB_SLAM3 {
namespace IMU {

const float eps = 1e-4;                                 // src/ImuTypes.cc:32

Eigen::Matrix3f NormalizeRotation(const Eigen::Matrix3f& R) { Eigen::Matrix3f out; orbo_normalize_rotation_f(R.m, out.m); return out; }   // :34-37 (JacobiSVD: stand-in)

class Bias {
public:
    Bias() : bax(0), bay(0), baz(0), bwx(0), bwy(0), bwz(0) {}
    Bias(const float& b_acc_x, const float& b_acc_y, const float& b_acc_z, const float& b_ang_vel_x, const float& b_ang_vel_y, const float& b_ang_vel_z)
        : bax(b_acc_x), bay(b_acc_y), baz(b_acc_z), bwx(b_ang_vel_x), bwy(b_ang_vel_y), bwz(b_ang_vel_z) {}
    float bax, bay, baz, bwx, bwy, bwz;
};
class IntegratedRotation {                              // include/ImuTypes.h:127-140
public:
    IntegratedRotation() {}
    IntegratedRotation(const Eigen::Vector3f& angVel, const Bias& imuBias, const float& time);
    float deltaT;
    Eigen::Matrix3f deltaR;
    Eigen::Matrix3f rightJ;
};
#include "imu_integrated_rotation.inc"

class Calib;
class Preintegrated {                                   // include/ImuTypes.h:143-240
public:
    Preintegrated() {}
    Preintegrated(const Bias& b_, const Calib& calib);   // src/ImuTypes.cc:107-112 (defined below the Calib shell)
    void Initialize(const Bias& b_);
    void IntegrateNewMeasurement(const Eigen::Vector3f& acceleration, const Eigen::Vector3f& angVel, const float& dt);
    float dT;
    Eigen::Matrix<float, 15, 15> C;
    Eigen::Matrix<float, 15, 15> Info;
    Eigen::DiagonalMatrix<float, 6> Nga, NgaWalk;
    Bias b;
    Eigen::Matrix3f dR;
    Eigen::Vector3f dV, dP;
    Eigen::Matrix3f JRg, JVg, JVa, JPg, JPa;
    Eigen::Vector3f avgA, avgW;
    Bias bu;
    Eigen::Matrix<float, 6, 1> db;
    struct integrable {
        integrable() {}
        integrable(const Eigen::Vector3f& a_, const Eigen::Vector3f& w_, const float& t_) : a(a_), w(w_), t(t_) {}
        Eigen::Vector3f a, w;
        float t;
    };
    std::vector<integrable> mvMeasurements;
    std::mutex mMutex;
};
#include "imu_initialize.inc"
#include "imu_integrate.inc"

class Calib { public: Eigen::DiagonalMatrix<float, 6> Cov, CovWalk; };
class Point {                                           // include/ImuTypes.h:46-59
public:
    Point(const float& acc_x, const float& acc_y, const float& acc_z, const float& ang_vel_x, const float& ang_vel_y, const float& ang_vel_z, const double& timestamp) : t(timestamp) {
        a << acc_x, acc_y, acc_z; w << ang_vel_x, ang_vel_y, ang_vel_z;
    }
    Eigen::Vector3f a, w;
    double t;
};
}  // namespace IMU

// Tracking::PreintegrateIMU (src/Tracking.cc:1628-1738): the queue selection (:1646-1678) and the integration steps (:1680-1729) are the reference's text; the
// members they touch are the shells below
namespace IMU { inline Preintegrated::Preintegrated(const Bias& b_, const Calib& calib) { Nga = calib.Cov; NgaWalk = calib.CovWalk; Initialize(b_); } }
struct FrameShell { double mTimeStamp; FrameShell* mpPrevFrame; IMU::Bias mImuBias; IMU::Calib mImuCalib; };
inline void usleep(int) {}
struct TrackingShell {
    std::list<IMU::Point> mlQueueImuData;
    std::vector<IMU::Point> mvImuFromLastFrame;
    std::mutex mMutexImuQueue;
    FrameShell mCurrentFrame, mLastFrame, mPrev;
    double mImuPer = 0.001;                             // :609
    IMU::Preintegrated* mpImuPreintegratedFromLastKF = nullptr;
    IMU::Preintegrated* lastFrame = nullptr;
    void select() {
        mvImuFromLastFrame.clear();
#include "tracking_preintegrate_select.inc"
    }
    void steps() {
#include "tracking_preintegrate_steps.inc"
        lastFrame = pImuPreintegratedFromLastFrame;
    }
};
}  // namespace ORB_SLAM3

extern "C" void ref_imu_preintegrate(int n, const float* acc, const float* gyr, const float* dts, const float* bias6, const float* noise4, float* P) {
    using namespace ORB_SLAM3::IMU;
    Preintegrated p;
    const float ng2 = noise4[0] * noise4[0], na2 = noise4[1] * noise4[1], ngw2 = noise4[2] * noise4[2], naw2 = noise4[3] * noise4[3];   // Calib::Set :397-406
    p.Nga.diagonal() << ng2, ng2, ng2, na2, na2, na2;
    p.NgaWalk.diagonal() << ngw2, ngw2, ngw2, naw2, naw2, naw2;
    p.Initialize(Bias(bias6[0], bias6[1], bias6[2], bias6[3], bias6[4], bias6[5]));
    for (int i = 0; i < n; ++i) {
        Eigen::Vector3f a, w;
        for (int k = 0; k < 3; ++k) { a[k] = acc[3 * i + k]; w[k] = gyr[3 * i + k]; }
        p.IntegrateNewMeasurement(a, w, dts[i]);
    }
    // the 292-float record of include/orb_b200.h: dT | dR | dV | dP | JRg | JVg | JVa | JPg | JPa | b | C
    int o = 0;
    P[o++] = p.dT;
    for (int i = 0; i < 9; ++i) P[o++] = p.dR.m[i];
    for (int i = 0; i < 3; ++i) P[o++] = p.dV.m[i];
    for (int i = 0; i < 3; ++i) P[o++] = p.dP.m[i];
    for (const Eigen::Matrix3f* M : {&p.JRg, &p.JVg, &p.JVa, &p.JPg, &p.JPa}) for (int i = 0; i < 9; ++i) P[o++] = M->m[i];
    P[o++] = p.b.bax; P[o++] = p.b.bay; P[o++] = p.b.baz; P[o++] = p.b.bwx; P[o++] = p.b.bwy; P[o++] = p.b.bwz;
    for (int i = 0; i < 225; ++i) P[o++] = p.C.m[i];
}

static void dump_record(const ORB_SLAM3::IMU::Preintegrated& p, float* P) {
    int o = 0;
    P[o++] = p.dT;
    for (int i = 0; i < 9; ++i) P[o++] = p.dR.m[i];
    for (int i = 0; i < 3; ++i) P[o++] = p.dV.m[i];
    for (int i = 0; i < 3; ++i) P[o++] = p.dP.m[i];
    for (const Eigen::Matrix3f* M : {&p.JRg, &p.JVg, &p.JVa, &p.JPg, &p.JPa}) for (int i = 0; i < 9; ++i) P[o++] = M->m[i];
    P[o++] = p.b.bax; P[o++] = p.b.bay; P[o++] = p.b.baz; P[o++] = p.b.bwx; P[o++] = p.b.bwy; P[o++] = p.b.bwz;
    for (int i = 0; i < 225; ++i) P[o++] = p.C.m[i];
}

// Tracking::PreintegrateIMU over a sequence of frames: the IMU samples (t, acc, gyr) with index < queuedUpTo[f] have been handed over (GrabImuData) when frame f arrives.
// For every frame f >= 1: selCount[f] / selFirst[f] = size and first index of mvImuFromLastFrame, PfromLastFrame [f][292] = pImuPreintegratedFromLastFrame (fresh per
// frame).  Samples must be identifiable by their time stamps (strictly increasing).
extern "C" void ref_tracking_preintegrate(int nImu, const double* tImu, const float* acc, const float* gyr, int nFrames, const double* tFrames, const int* queuedUpTo,
                                          const float* bias6, const float* noise4, int* selFirst, int* selCount, float* PfromLastFrame) {
    using namespace ORB_SLAM3;
    TrackingShell T;
    IMU::Calib calib;
    const float ng2 = noise4[0] * noise4[0], na2 = noise4[1] * noise4[1], ngw2 = noise4[2] * noise4[2], naw2 = noise4[3] * noise4[3];
    calib.Cov.diagonal() << ng2, ng2, ng2, na2, na2, na2;
    calib.CovWalk.diagonal() << ngw2, ngw2, ngw2, naw2, naw2, naw2;
    const IMU::Bias bias(bias6[0], bias6[1], bias6[2], bias6[3], bias6[4], bias6[5]);
    IMU::Preintegrated fromLastKF(bias, calib);
    T.mpImuPreintegratedFromLastKF = &fromLastKF;
    int pushed = 0;
    for (int f = 0; f < nFrames; ++f) {
        for (; pushed < queuedUpTo[f]; ++pushed)
            T.mlQueueImuData.push_back(IMU::Point(acc[3 * pushed], acc[3 * pushed + 1], acc[3 * pushed + 2], gyr[3 * pushed], gyr[3 * pushed + 1], gyr[3 * pushed + 2], tImu[pushed]));
        selFirst[f] = -1; selCount[f] = 0;
        if (f == 0) continue;                            // no previous frame: PreintegrateIMU returns at :1631-1636
        T.mPrev.mTimeStamp = tFrames[f - 1];
        T.mCurrentFrame.mTimeStamp = tFrames[f]; T.mCurrentFrame.mpPrevFrame = &T.mPrev; T.mCurrentFrame.mImuCalib = calib;
        T.mLastFrame.mImuBias = bias;
        T.select();
        selCount[f] = (int)T.mvImuFromLastFrame.size();
        if (selCount[f]) for (int k = 0; k < nImu; ++k) if (tImu[k] == T.mvImuFromLastFrame[0].t) { selFirst[f] = k; break; }
        T.lastFrame = nullptr;
        if (selCount[f] > 0) T.steps();
        if (T.lastFrame) { dump_record(*T.lastFrame, PfromLastFrame + 292 * (size_t)f); delete T.lastFrame; }
    }
}
