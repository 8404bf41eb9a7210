// TEST INFRASTRUCTURE ONLY -- declarations that let LINE RANGES of the reference's matcher-side sources
// (src/ORBmatcher.cc, src/Frame.cc, src/MapPoint.cc, src/CameraModels/Pinhole.cpp; the ranges are listed in
// extract_ranges.py and written, at build time only, to oracle/_ref/gen/) compile verbatim without Eigen, Sophus,
// DBoW2 or the rest of the SLAM object graph.  The classes below carry exactly the members those function bodies
// touch, under the reference's member names (include/Frame.h, include/MapPoint.h, include/ORBmatcher.h,
// include/CameraModels/Pinhole.h); no function body of the reference is restated here.
//
// Arithmetic that the reference gets from Eigen / Sophus expression templates is written out once, in the order the
// oracle documents (oracle/matcher_oracle.cpp header): Matrix3f * Vector3f row-wise left to right, norm / dot left to
// right, SO3f * point in Sophus' quaternion form (Thirdparty/Sophus/sophus/so3.hpp:358-367).  Those few operators live out of
// line in ref_slam_types.cpp, compiled with -ffp-contract=off so they stay individually rounded; the reference's own text
// (everything that includes this header) is compiled with the reference's flags, contraction included.
#pragma once
#include <climits>
#include <cmath>
#include <cstdint>
#include <map>
#include <mutex>
#include <set>
#include <tuple>
#include <vector>
#include "minicv.hpp"
#include "DBoW2/FeatureVector.h"   // the reference's own DBoW2 classes (-I $(REF)/Thirdparty/DBoW2): SearchByBoW walks two FeatureVectors
#include "ORBextractor.h"   // the reference's own header (resolved through -I $(REF)/include): Frame::ComputeStereoMatches reads mvImagePyramid

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW

namespace Eigen {
template <typename T, int R, int C> struct Matrix;
template <> struct Matrix<float, 3, 1> {
    float v[3];
    Matrix() : v{0, 0, 0} {}
    Matrix(float a, float b, float c) : v{a, b, c} {}
    float& operator()(int i) { return v[i]; }
    const float& operator()(int i) const { return v[i]; }
    float& operator[](int i) { return v[i]; }
    const float& operator[](int i) const { return v[i]; }
    // out of line in ref_slam_types.cpp (compiled with -ffp-contract=off, see the header comment)
    Matrix operator+(const Matrix& o) const;
    Matrix operator-(const Matrix& o) const;
    Matrix operator/(float d) const;
    float dot(const Matrix& o) const;
    float norm() const;
};
template <> struct Matrix<float, 2, 1> {
    float v[2];
    Matrix() : v{0, 0} {}
    float& operator()(int i) { return v[i]; }
    const float& operator()(int i) const { return v[i]; }
    float& operator[](int i) { return v[i]; }
    const float& operator[](int i) const { return v[i]; }
};
template <> struct Matrix<float, 3, 3> {
    float m[9];   // row-major
    Matrix() : m{1, 0, 0, 0, 1, 0, 0, 0, 1} {}
    Matrix<float, 3, 1> operator*(const Matrix<float, 3, 1>& p) const;
    const float& operator()(int i, int j) const { return m[3 * i + j]; }
    float& operator()(int i, int j) { return m[3 * i + j]; }
    // what Pinhole::epipolarConstrain's fundamental-matrix line needs (out of line, individually rounded; the C-ABI takes F12 as an
    // INPUT computed by the caller's own Eigen code, so this stand-in only has to be one fixed arithmetic, not Eigen's)
    Matrix operator*(const Matrix& o) const;   // row-major triple loop, left to right
    Matrix transpose() const;
    Matrix inverse() const;                    // cofactors / determinant
};
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<float, 2, 1> Vector2f;
typedef Matrix<float, 3, 3> Matrix3f;
}  // namespace Eigen

namespace Sophus {
template <typename T> struct SE3;
template <> struct SE3<float> {
    float qw, qx, qy, qz;   // unit quaternion
    Eigen::Vector3f t;
    SE3() : qw(1), qx(0), qy(0), qz(0) {}
    SE3(const Eigen::Matrix<float, 3, 3>& R, const Eigen::Vector3f& trans);   // se3.hpp:481: rotation matrix -> unit quaternion (Eigen's conversion)
    // so3.hpp:358-367 then se3.hpp:321-324 (+ translation)
    Eigen::Vector3f rotate(const Eigen::Vector3f& p) const;
    Eigen::Vector3f operator*(const Eigen::Vector3f& p) const;
    SE3 inverse() const;
    const Eigen::Vector3f& translation() const { return t; }
    SE3 operator*(const SE3& o) const;         // quaternion product (not renormalised, so3.hpp:316-331), translation t + R o.t
    Eigen::Matrix3f rotationMatrix() const;    // Eigen::Quaternion::toRotationMatrix
};
typedef SE3<float> SE3f;
template <typename T> struct SO3;
template <> struct SO3<float> { static Eigen::Matrix3f hat(const Eigen::Vector3f& w); };   // so3.hpp:631-640
typedef SO3<float> SO3f;
template <typename T> struct Sim3;
template <> struct Sim3<float> {       // x -> s R x + t with R the rotation of the unit quaternion (Sophus keeps s inside the quaternion's norm)
    float s, qw, qx, qy, qz;
    Eigen::Vector3f t;
    Sim3() : s(1), qw(1), qx(0), qy(0), qz(0) {}
    Sim3 inverse() const;
    Eigen::Vector3f operator*(const Eigen::Vector3f& p) const;
    Eigen::Matrix<float, 3, 3> rotationMatrix() const;
    const Eigen::Vector3f& translation() const { return t; }
    float scale() const { return s; }
};
typedef Sim3<float> Sim3f;
}  // namespace Sophus

#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64

namespace ORB_SLAM3 {
using std::vector;
using std::pair;

class Frame;
class KeyFrame;
class MapPoint;
class ORBmatcher;

class GeometricCamera {
public:
    virtual ~GeometricCamera() {}
    virtual Eigen::Vector2f project(const Eigen::Vector3f& v3D) = 0;
    virtual Eigen::Matrix3f toK_() = 0;
    virtual bool epipolarConstrain(GeometricCamera* pCamera2, const cv::KeyPoint& kp1, const cv::KeyPoint& kp2, const Eigen::Matrix3f& R12,
                                   const Eigen::Vector3f& t12, const float sigmaLevel, const float unc) = 0;
    std::vector<float> mvParameters;
};
class Pinhole : public GeometricCamera {
public:
    Eigen::Vector2f project(const Eigen::Vector3f& v3D);   // body: src/CameraModels/Pinhole.cpp:43-49
    Eigen::Matrix3f toK_() {                                // src/CameraModels/Pinhole.cpp:100-104 (a comma initialiser, no arithmetic)
        Eigen::Matrix3f K;
        K(0, 0) = mvParameters[0]; K(0, 1) = 0.f; K(0, 2) = mvParameters[2]; K(1, 0) = 0.f; K(1, 1) = mvParameters[1]; K(1, 2) = mvParameters[3];
        K(2, 0) = 0.f; K(2, 1) = 0.f; K(2, 2) = 1.f;
        return K;
    }
    bool epipolarConstrain(GeometricCamera* pCamera2, const cv::KeyPoint& kp1, const cv::KeyPoint& kp2, const Eigen::Matrix3f& R12,
                           const Eigen::Vector3f& t12, const float sigmaLevel, const float unc);   // body: src/CameraModels/Pinhole.cpp:107-129
};

class KeyFrame {
public:
    std::vector<MapPoint*> mvpMapPoints;
    DBoW2::FeatureVector mFeatVec;
    cv::Mat mDescriptors;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight, mvKeysUn;
    GeometricCamera *mpCamera = nullptr, *mpCamera2 = nullptr;
    int NLeft = -1;
    bool mbBad = false;
    std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
    bool isBad() { return mbBad; }
    // --- what ORBmatcher::Fuse reads (include/KeyFrame.h) ---
    int N = 0, mnScaleLevels = 0;
    int mnGridCols = FRAME_GRID_COLS, mnGridRows = FRAME_GRID_ROWS;
    float fx = 0, fy = 0, cx = 0, cy = 0, mbf = 0, mfLogScaleFactor = 0;
    float mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0, mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
    std::vector<float> mvScaleFactors, mvInvLevelSigma2, mvLevelSigma2, mvuRight;
    std::vector<std::vector<std::vector<size_t>>> mGrid, mGridRight;
    Sophus::SE3f mTcw;
    Eigen::Vector3f mOw;
    Sophus::SE3f GetPose() { return mTcw; }
    Sophus::SE3f GetRightPose() { return mTcw; }
    Sophus::SE3f GetPoseInverse() { return mTcw.inverse(); }        // the reference caches mTwc = mTcw.inverse() (src/KeyFrame.cc:115)
    Sophus::SE3f GetRightPoseInverse() { return mTcw.inverse(); }
    Eigen::Vector3f GetCameraCenter() { return mOw; }
    Eigen::Vector3f GetRightCameraCenter() { return mOw; }
    bool IsInImage(const float& x, const float& y) const;                       // body: src/KeyFrame.cc:750-753
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const bool bRight = false) const;   // body: src/KeyFrame.cc:704-748
    MapPoint* GetMapPoint(const size_t& idx) { return mvpMapPoints[idx]; }
    std::set<MapPoint*> GetMapPoints();                       // the good map points of the keyframe (src/KeyFrame.cc:317-330)
    void AddMapPoint(MapPoint* pMP, const size_t& idx) { mvpMapPoints[idx] = pMP; }
};

class MapPoint {
public:
    std::map<KeyFrame*, std::tuple<int, int>> mObservations;
    std::mutex mMutexFeatures;
    void ComputeDistinctiveDescriptors();                      // body: src/MapPoint.cc:329-403
    std::tuple<int, int> GetIndexInKeyFrame(KeyFrame* pKF);    // body: src/MapPoint.cc:411-418
    // flattened state (what the reference reads under mutexes)
    Eigen::Vector3f mWorldPos, mNormalVector;
    cv::Mat mDescriptor;
    int nObs = 0;
    bool mbBad = false;
    float mfMinDistance = 0, mfMaxDistance = 0;
    std::mutex mMutexPos;
    int index = -1;   // position in the wrapper's flat arrays

    Eigen::Vector3f GetWorldPos() { return mWorldPos; }
    Eigen::Vector3f GetNormal() { return mNormalVector; }
    cv::Mat GetDescriptor() { return mDescriptor.clone(); }
    int Observations() { return nObs; }
    bool isBad() { return mbBad; }
    float GetMinDistanceInvariance();                          // bodies: src/MapPoint.cc:502-512
    float GetMaxDistanceInvariance();
    int PredictScale(const float& currentDist, Frame* pF);     // body: src/MapPoint.cc:531-546
    int PredictScale(const float& currentDist, KeyFrame* pKF); // body: src/MapPoint.cc:514-529
    // Fuse's view of the map: which keyframes observe the point, and a log of what Fuse did with it (the wrapper reads the log back)
    std::set<KeyFrame*> inKF;
    // slot: the keypoint of the (one) keyframe under Fuse that holds this point, -1 if none.  fuseAction / fuseIdx log what Fuse did for the
    // searched map point: 1 = AddObservation(pKF, idx); 2 = it was Replace()d by the keyframe's point at idx; 3 = it Replace()d the point at idx.
    int slot = -1, fuseAction = 0, fuseIdx = -1;
    KeyFrame* slotKF = nullptr;
    bool IsInKeyFrame(KeyFrame* pKF) { return inKF.count(pKF) != 0; }
    void AddObservation(KeyFrame* pKF, int idx) { fuseAction = 1; fuseIdx = idx; slot = idx; slotKF = pKF; inKF.insert(pKF); ++nObs; }
    // MapPoint::Replace (src/MapPoint.cc:232-291) as Fuse sees it: this point goes bad; where the keyframe held it, it now holds pMP
    void Replace(MapPoint* pMP) {
        mbBad = true;
        if (slot >= 0) { pMP->fuseAction = 3; pMP->fuseIdx = slot; pMP->slot = slot; pMP->slotKF = slotKF; pMP->inKF.insert(slotKF); ++pMP->nObs; replace_in_keyframe(pMP); slot = -1; }
        else { fuseAction = 2; fuseIdx = pMP->slot; }
    }
    void replace_in_keyframe(MapPoint* pMP);

    float mTrackProjX = 0, mTrackProjY = 0, mTrackDepth = 0, mTrackDepthR = 0, mTrackProjXR = 0, mTrackProjYR = 0;
    bool mbTrackInView = false, mbTrackInViewR = false;
    int mnTrackScaleLevel = 0, mnTrackScaleLevelR = 0;
    float mTrackViewCos = 0, mTrackViewCosR = 0;
};

inline std::set<MapPoint*> KeyFrame::GetMapPoints() { std::set<MapPoint*> s; for (MapPoint* p : mvpMapPoints) if (p && !p->isBad()) s.insert(p); return s; }
inline void MapPoint::replace_in_keyframe(MapPoint* pMP) { slotKF->mvpMapPoints[slot] = pMP; }

class Frame {
public:
    int N = 0, Nleft = -1;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight, mvKeysUn;
    std::vector<float> mvuRight, mvDepth;
    cv::Mat mDescriptorsRight;
    std::vector<float> mvInvScaleFactors;
    ORBextractor *mpORBextractorLeft = nullptr, *mpORBextractorRight = nullptr;
    cv::Mat mDescriptors;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    std::vector<int> mvLeftToRightMatch, mvRightToLeftMatch;
    DBoW2::FeatureVector mFeatVec;
    GeometricCamera* mpCamera2 = nullptr;
    float mb = 0, mbf = 0;
    std::vector<float> mvScaleFactors;
    int mnScaleLevels = 0;
    float mfLogScaleFactor = 0;
    float mnMinX = 0, mnMaxX = 0, mnMinY = 0, mnMaxY = 0;      // static in the reference; per object here
    float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
    std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    std::vector<std::size_t> mGridRight[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    GeometricCamera* mpCamera = nullptr;
    Eigen::Matrix<float, 3, 1> mOw;
    Eigen::Matrix<float, 3, 3> mRcw;
    Eigen::Matrix<float, 3, 1> mtcw;
    Sophus::SE3<float> mTcw;

    Sophus::SE3<float> GetPose() const { return mTcw; }
    Sophus::SE3f GetRelativePoseTrl() { return Sophus::SE3f(); }
    void AssignFeaturesToGrid();                                // body: src/Frame.cc:385-416
    void ComputeStereoMatches();                                // body: src/Frame.cc:811-982
    bool isInFrustum(MapPoint* pMP, float viewingCosLimit);     // body: src/Frame.cc:512-573 (monocular branch)
    bool PosInGrid(const cv::KeyPoint& kp, int& posX, int& posY);   // body: src/Frame.cc:725-735
    vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1, const int maxLevel = -1,
                                     const bool bRight = false) const;   // body: src/Frame.cc:657-723
};

class ORBmatcher {
public:
    ORBmatcher(float nnratio = 0.6, bool checkOri = true);
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
    int SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th = 3, const bool bFarPoints = false,
                           const float thFarPoints = 50.0f);
    int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono);
    int SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches);
    int SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12);
    int Fuse(KeyFrame* pKF, const vector<MapPoint*>& vpMapPoints, const float th = 3.0, const bool bRight = false);
    int Fuse(KeyFrame* pKF, Sophus::Sim3f& Scw, const std::vector<MapPoint*>& vpPoints, float th, vector<MapPoint*>& vpReplacePoint);
    int SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12, const Sophus::Sim3f& S12, const float th);
    int SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<pair<size_t, size_t>>& vMatchedPairs, const bool bOnlyStereo, const bool bCoarse = false);
    int SearchForInitialization(Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12, int windowSize = 10);
    static const int TH_LOW;
    static const int TH_HIGH;
    static const int HISTO_LENGTH;
    float RadiusByViewingCos(const float& viewCos);
    void ComputeThreeMaxima(std::vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3);
    float mfNNratio;
    bool mbCheckOrientation;
};

}  // namespace ORB_SLAM3
