// Host-side C++ mirror of the three reference class surfaces over the C-ABI (orb_b200.h).  Header only, no OpenCV/Eigen:
// cv::KeyPoint / cv::Mat are replaced by layout-compatible PODs.  Method names, argument meaning and
// error behaviour follow the reference so that Frame.cc / Tracking.cc / LocalMapping.cc-style callers read the same.
//   ORB_SLAM3::ORBextractor   reference include/ORBextractor.h:43-109
//   ORB_SLAM3::ORBmatcher     reference include/ORBmatcher.h:36-103   (per-frame hot functions)
//   ORB_SLAM3::Optimizer      reference include/Optimizer.h:46-102    (LocalBundleAdjustment numeric core)
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <cmath>
#include <vector>

#include "../orb_b200.h"

namespace ORB_SLAM3 {

using KeyPoint = OrbKeyPoint;                    // byte-compatible with cv::KeyPoint
struct Descriptors {                             // stands in for a CV_8UC1 cv::Mat with 32 columns
    int rows = 0;
    std::vector<uint8_t> data;
    uint8_t* ptr(int r) { return data.data() + (size_t)r * 32; }
    const uint8_t* ptr(int r) const { return data.data() + (size_t)r * 32; }
    void release() { rows = 0; data.clear(); }
};
struct Image {                                   // stands in for a CV_8UC1 cv::Mat
    const uint8_t* data = nullptr;
    int rows = 0, cols = 0;
    size_t step = 0;
    bool empty() const { return !data || rows <= 0 || cols <= 0; }
};

inline void orb_check(int rc, const char* where) {
    if (rc != ORB_OK) throw std::runtime_error(std::string(where) + ": " + orb_last_error());
}

class ORBextractor {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };
    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, int maxWidth = 1920, int maxHeight = 1080,
                 int device = 0)
        : nlevels_(nlevels) {
        orb_check(orbx_create(&h_, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, maxWidth, maxHeight, 1, device), "orbx_create");
        mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels); mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
        orbx_get_tables(h_, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(), mvInvLevelSigma2.data(), nullptr);
    }
    ~ORBextractor() { orbx_destroy(h_); }
    ORBextractor(const ORBextractor&) = delete;
    ORBextractor& operator=(const ORBextractor&) = delete;

    // int operator()(InputArray image, InputArray mask, vector<KeyPoint>&, OutputArray descriptors, vector<int>& vLappingArea)
    // Returns monoIndex, or -1 for an empty image (src/ORBextractor.cc:1090-1091).  The mask is ignored, as in the reference.
    int operator()(const Image& image, const Image& /*mask*/, std::vector<KeyPoint>& keypoints, Descriptors& descriptors,
                   std::vector<int>& vLappingArea) {
        if (image.empty()) return -1;
        const int cap = orbx_max_keypoints(h_);
        keypoints.resize(cap);
        descriptors.data.resize((size_t)cap * 32);
        int n = 0, mono = 0;
        const int rc = orbx_extract(h_, image.data, image.rows, image.cols, image.step, vLappingArea.at(0), vLappingArea.at(1),
                                    keypoints.data(), descriptors.data.data(), cap, &n, &mono);
        if (rc == ORB_ERR_EMPTY) return -1;
        orb_check(rc, "orbx_extract");
        keypoints.resize(n);
        descriptors.rows = n;
        descriptors.data.resize((size_t)n * 32);
        if (n == 0) descriptors.release();         // _descriptors.release() (:1108-1109)
        return mono;
    }
    int GetLevels() { return nlevels_; }
    float GetScaleFactor() { return nlevels_ > 1 ? mvScaleFactor[1] : 1.0f; }
    std::vector<float> GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }
    // mvImagePyramid[level] of the last call (public member in the reference, read by the stereo matcher only)
    std::vector<uint8_t> ImagePyramidLevel(int level, int& width, int& height) {
        orb_check(orbx_get_level_size(h_, level, &width, &height), "orbx_get_level_size");
        std::vector<uint8_t> out((size_t)width * height);
        orb_check(orbx_copy_level(h_, 0, level, 0, out.data()), "orbx_copy_level");
        return out;
    }
    orbx_handle* handle() { return h_; }

private:
    orbx_handle* h_ = nullptr;
    int nlevels_;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
};

// The slice of Frame the matchers read (mvKeysUn, mDescriptors, image bounds, mvScaleFactors) plus the in/out
// mvpMapPoints state as (index of the map point in the caller's array, Observations()>0 flag).
struct FrameView {
    std::vector<KeyPoint> mvKeysUn;
    Descriptors mDescriptors;
    float mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0;
    std::vector<float> mvScaleFactors;
    std::vector<int32_t> mvpMapPoints;     // -1 == NULL
    std::vector<uint8_t> mvbClaimed;
    OrbmFrame c_struct() const {
        OrbmFrame f;
        f.K = (int)mvKeysUn.size(); f.keypoints = mvKeysUn.data(); f.descriptors = mDescriptors.data.data();
        f.minX = mnMinX; f.minY = mnMinY; f.maxX = mnMaxX; f.maxY = mnMaxY;
        f.scaleFactors = mvScaleFactors.data(); f.nlevels = (int)mvScaleFactors.size();
        return f;
    }
};

class ORBmatcher {
public:
    static const int TH_LOW = 50, TH_HIGH = 100, HISTO_LENGTH = 30;   // src/ORBmatcher.cc:35-37
    ORBmatcher(float nnratio = 0.6, bool checkOri = true, int maxKeypoints = 4096, int maxMapPoints = 16384, int device = 0)
        : mfNNratio(nnratio), mbCheckOrientation(checkOri) {
        orb_check(orbm_create(&h_, 1, maxKeypoints, maxMapPoints, device), "orbm_create");
    }
    ~ORBmatcher() { orbm_destroy(h_); }
    ORBmatcher(const ORBmatcher&) = delete;
    ORBmatcher& operator=(const ORBmatcher&) = delete;

    // static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b) -- needs a handle here (device popcount)
    int DescriptorDistance(const uint8_t* a, const uint8_t* b) {
        int32_t d = 0;
        orb_check(orbm_descriptor_distance(h_, a, b, 1, &d), "orbm_descriptor_distance");
        return d;
    }
    // int SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, const float th, const bool bFarPoints, const float thFarPoints)
    int SearchByProjection(FrameView& F, const OrbmLocalPoints& vpMapPoints, const float th = 3, const bool bFarPoints = false,
                           const float thFarPoints = 50.0f) {
        prepare(F);
        const OrbmFrame f = F.c_struct();
        int n = 0;
        orb_check(orbm_search_local_map(h_, &f, &vpMapPoints, th, mfNNratio, bFarPoints, thFarPoints, F.mvpMapPoints.data(), F.mvbClaimed.data(), &n),
                  "orbm_search_local_map");
        return n;
    }
    // int SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono)  (mono branch)
    int SearchByProjection(FrameView& CurrentFrame, const OrbmLastFrame& LastFrame, const float Tcw[7], const float cam[4], const float th,
                           const bool bMono) {
        if (!bMono) throw std::invalid_argument("only the monocular branch is accelerated");
        prepare(CurrentFrame);
        const OrbmFrame f = CurrentFrame.c_struct();
        int n = 0;
        orb_check(orbm_search_last_frame(h_, &f, &LastFrame, Tcw, cam, th, mbCheckOrientation, CurrentFrame.mvpMapPoints.data(),
                                         CurrentFrame.mvbClaimed.data(), &n), "orbm_search_last_frame");
        return n;
    }
    // bool Frame::isInFrustum(MapPoint* pMP, float viewingCosLimit) for all local map points of a frame (Tracking::SearchLocalPoints,
    // src/Tracking.cc:3346): fills the per-point tracking fields that the local-map SearchByProjection above reads.
    struct TrackFields { std::vector<uint8_t> inView; std::vector<float> projX, projY, projXR, depth, viewCos; std::vector<int32_t> level; };
    void isInFrustum(const OrbmFrustumIn& points, TrackFields& out) {
        const size_t M = (size_t)(points.M > 0 ? points.M : 0);
        out.inView.assign(M, 0); out.projX.assign(M, -1.f); out.projY.assign(M, -1.f); out.projXR.assign(M, 0.f); out.depth.assign(M, 0.f);
        out.viewCos.assign(M, 0.f); out.level.assign(M, -1);
        if (!M) return;
        orb_check(orbm_frustum_project(h_, &points, out.inView.data(), out.projX.data(), out.projY.data(), out.projXR.data(), out.depth.data(),
                                       out.level.data(), out.viewCos.data()), "orbm_frustum_project");
    }
    // cv::BFMatcher(NORM_HAMMING).knnMatch(query, train, matches, 2)
    void knnMatch(const Descriptors& query, const Descriptors& train, std::vector<int32_t>& idx, std::vector<int32_t>& dist) {
        idx.assign((size_t)query.rows * 2, -1); dist.assign((size_t)query.rows * 2, -1);
        orb_check(orbm_bf_knn2(h_, query.data.data(), query.rows, train.data.data(), train.rows, idx.data(), dist.data()), "orbm_bf_knn2");
    }

protected:
    static void prepare(FrameView& F) {
        if (F.mvpMapPoints.size() != F.mvKeysUn.size()) F.mvpMapPoints.assign(F.mvKeysUn.size(), -1);
        if (F.mvbClaimed.size() != F.mvKeysUn.size()) F.mvbClaimed.assign(F.mvKeysUn.size(), 0);
    }
    float mfNNratio;
    bool mbCheckOrientation;
    orbm_handle* h_ = nullptr;
};

class Optimizer {
public:
    // void static LocalBundleAdjustment(KeyFrame* pKF, bool* pbStopFlag, Map* pMap, int& num_fixedKF, int& num_OptKF, int& num_MPs, int& num_edges)
    // The pointer-graph walk (src/Optimizer.cc:1125-1403) and the write-back (:1464-1497) stay with the caller; this is
    // optimizer.initializeOptimization(); optimizer.optimize(10); and the values the outlier test at :1417-1430 reads.
    static void LocalBundleAdjustment(const LbaProblem& graph, LbaResult& out, int device = 0) {
        if (graph.stopFlag && *graph.stopFlag) return;               // :1406-1408
        lba_handle* h = nullptr;
        orb_check(lba_create(&h, graph.nPoses, graph.nPoints > 0 ? graph.nPoints : 1, graph.nEdges > 0 ? graph.nEdges : 1, device), "lba_create");
        const int rc = lba_solve(h, &graph, &out);
        lba_destroy(h);
        orb_check(rc, "lba_solve");
    }
    // int static PoseOptimization(Frame* pFrame)  (src/Optimizer.cc:814-1113, monocular branch): the caller flattens the frame's map
    // point associations (:862-905); returns nInitialCorrespondences - nBad, writes the pose and the outlier flags.
    static int PoseOptimization(int N, const double pose7[7], const float cam4[4], const double* Xw, const double* obs, const float* invSigma2,
                                double poseOut[7], uint8_t* mvbOutlier, int device = 0) {
        int32_t n = N, inliers = 0;
        orb_check(pose_optimization_batch(1, N > 0 ? N : 1, &n, pose7, cam4, Xw, obs, invSigma2, (double)(float)std::sqrt(5.991) /* deltaMono, :852 */, poseOut, mvbOutlier, &inliers, device),
                  "pose_optimization_batch");
        return inliers;
    }
};

}  // namespace ORB_SLAM3
