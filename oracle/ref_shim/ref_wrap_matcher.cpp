// TEST INFRASTRUCTURE ONLY -- C entry points over the REFERENCE's own matcher-side function bodies.  The bodies are
// #included verbatim from oracle/_ref/gen/*.inc, which oracle/Makefile generates at build time from line ranges of
// /root/reference/src/{ORBmatcher.cc,Frame.cc,MapPoint.cc,CameraModels/Pinhole.cpp} (extract_ranges.py; nothing of it is
// committed).  This file only flattens arrays into the stand-in Frame / MapPoint objects (ref_slam_types.hpp) and back,
// with the same flat signatures as the orbo_* functions of oracle/matcher_oracle.cpp so that tests can compare 1:1.
#include "ref_slam_types.hpp"

using namespace std;

namespace ORB_SLAM3 {
#include "matcher_consts.inc"
#include "matcher_local_map.inc"
#include "matcher_bow_kf_frame.inc"
#include "matcher_init.inc"
#include "matcher_fuse.inc"
#include "matcher_fuse_sim3.inc"
#include "matcher_sim3.inc"
#include "matcher_triangulation.inc"
#include "matcher_bow_kf_kf.inc"
#include "matcher_last_frame.inc"
#include "matcher_maxima_distance.inc"
#include "pinhole_project.inc"
#include "pinhole_epipolar.inc"
}  // namespace ORB_SLAM3
namespace ORB_SLAM3 {
#include "frame_assign_grid.inc"
#include "frame_in_frustum_mono.inc"
    return false;   // the reference continues with the two-camera branch (src/Frame.cc:574-...), not on the monocular path
}
#include "frame_features_in_area.inc"
#include "frame_stereo_matches.inc"
#include "mappoint_distinctive.inc"
#include "mappoint_index_in_kf.inc"
#include "keyframe_features_in_area.inc"
#include "mappoint_predict_scale_kf.inc"
#include "mappoint_invariance.inc"
#include "mappoint_predict_scale.inc"
}  // namespace ORB_SLAM3

using namespace ORB_SLAM3;

namespace {
void fill_frame(Frame& F, int K, const cv::KeyPoint* kps, const uint8_t* desc, const float* bounds, const float* scaleFactors, int nlevels) {
    F.N = K; F.Nleft = -1;
    F.mvKeys.assign(kps, kps + K);
    F.mvKeysUn.assign(kps, kps + K);           // zero distortion: mvKeysUn == mvKeys (src/Frame.cc:749)
    F.mvuRight.assign(K, -1.f);                // monocular (src/Frame.cc:326)
    F.mDescriptors = K ? cv::Mat(K, 32, CV_8UC1, (void*)desc, 32) : cv::Mat();
    F.mvpMapPoints.assign(K, (MapPoint*)nullptr);
    F.mvbOutlier.assign(K, false);
    F.mnMinX = bounds[0]; F.mnMinY = bounds[1]; F.mnMaxX = bounds[2]; F.mnMaxY = bounds[3];
    // src/Frame.cc:342-343
    F.mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(F.mnMaxX - F.mnMinX);
    F.mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(F.mnMaxY - F.mnMinY);
    if (scaleFactors) F.mvScaleFactors.assign(scaleFactors, scaleFactors + nlevels);
    F.AssignFeaturesToGrid();
}
// state carried in from earlier searches: keypoint k holds map point curMatch[k], whose Observations()>0 iff curClaimed[k]
void load_state(Frame& F, std::vector<MapPoint>& prior, const int* curMatch, const uint8_t* curClaimed) {
    for (int k = 0; k < F.N; ++k)
        if (curMatch[k] >= 0) {
            prior[k].index = curMatch[k];
            prior[k].nObs = curClaimed[k] ? 1 : 0;
            F.mvpMapPoints[k] = &prior[k];
        }
}
void store_state(const Frame& F, int* curMatch, uint8_t* curClaimed) {
    for (int k = 0; k < F.N; ++k) {
        MapPoint* p = F.mvpMapPoints[k];
        curMatch[k] = p ? p->index : -1;
        curClaimed[k] = p ? (p->nObs > 0) : 0;
    }
}
}  // namespace

extern "C" {

int ref_descriptor_distance(const uint8_t* a, const uint8_t* b) {
    return ORBmatcher::DescriptorDistance(cv::Mat(1, 32, CV_8UC1, (void*)a, 32), cv::Mat(1, 32, CV_8UC1, (void*)b, 32));
}

void ref_compute_three_maxima(const int* histoSizes, int L, int* ind3) {
    std::vector<std::vector<int>> h(L);
    for (int i = 0; i < L; ++i) h[i].assign(histoSizes[i], 0);
    ORBmatcher m;
    ind3[0] = ind3[1] = ind3[2] = -1;
    m.ComputeThreeMaxima(h.data(), L, ind3[0], ind3[1], ind3[2]);
}

int ref_features_in_area(int K, const cv::KeyPoint* kps, const float* bounds, float x, float y, float r, int minLevel, int maxLevel,
                         int* out, int cap) {
    Frame F;
    fill_frame(F, K, kps, nullptr, bounds, nullptr, 0);
    vector<size_t> v = F.GetFeaturesInArea(x, y, r, minLevel, maxLevel);
    for (size_t i = 0; i < v.size() && (int)i < cap; ++i) out[i] = (int)v[i];
    return (int)v.size();
}

int ref_search_local_map(int K, const cv::KeyPoint* kps, const uint8_t* desc, const float* bounds, const float* scaleFactors, int M,
                         const uint8_t* inView, const uint8_t* bad, const float* depth, const float* projX, const float* projY,
                         const int* level, const float* viewCos, const uint8_t* hasObs, const uint8_t* mpDesc, float th, float nnratio,
                         int bFarPoints, float thFarPoints, int* curMatch, uint8_t* curClaimed) {
    Frame F;
    fill_frame(F, K, kps, desc, bounds, scaleFactors, 64);   // nlevels is not part of this flat signature: take what is there
    std::vector<MapPoint> prior(K);   // MapPoint holds a mutex: sized once, never moved
    load_state(F, prior, curMatch, curClaimed);
    std::vector<MapPoint> mps(M);
    std::vector<MapPoint*> vp(M);
    for (int i = 0; i < M; ++i) {
        MapPoint& p = mps[i];
        p.index = i; p.mbTrackInView = inView[i]; p.mbBad = bad[i]; p.mTrackDepth = depth[i]; p.mTrackProjX = projX[i];
        p.mTrackProjY = projY[i]; p.mnTrackScaleLevel = level[i]; p.mTrackViewCos = viewCos[i]; p.nObs = hasObs[i] ? 1 : 0;
        p.mDescriptor = cv::Mat(1, 32, CV_8UC1, (void*)(mpDesc + (size_t)i * 32), 32);
        vp[i] = &p;
    }
    ORBmatcher matcher(nnratio, true);
    int n = matcher.SearchByProjection(F, vp, th, bFarPoints != 0, thFarPoints);
    store_state(F, curMatch, curClaimed);
    return n;
}

int ref_search_last_frame(int K, const cv::KeyPoint* kps, const uint8_t* desc, const float* bounds, const float* scaleFactors,
                          const float* Tcw, const float* cam, int M, const uint8_t* valid, const float* xyz, const int* lastOctave,
                          const float* lastAngle, const uint8_t* hasObs, const uint8_t* mpDesc, float th, int checkOrientation,
                          int* curMatch, uint8_t* curClaimed) {
    Frame Cur, Last;
    fill_frame(Cur, K, kps, desc, bounds, scaleFactors, 64);
    std::vector<MapPoint> prior(K);
    load_state(Cur, prior, curMatch, curClaimed);
    Pinhole camera;
    camera.mvParameters.assign(cam, cam + 4);
    Cur.mpCamera = &camera;
    Cur.mTcw.qw = Tcw[0]; Cur.mTcw.qx = Tcw[1]; Cur.mTcw.qy = Tcw[2]; Cur.mTcw.qz = Tcw[3];
    Cur.mTcw.t = Eigen::Vector3f(Tcw[4], Tcw[5], Tcw[6]);
    Last.N = M; Last.Nleft = -1;
    Last.mvKeys.resize(M); Last.mvKeysUn.resize(M);
    Last.mvpMapPoints.assign(M, (MapPoint*)nullptr);
    Last.mvbOutlier.assign(M, false);
    std::vector<MapPoint> mps(M);
    for (int i = 0; i < M; ++i) {
        Last.mvKeys[i].octave = lastOctave[i]; Last.mvKeysUn[i].octave = lastOctave[i];
        Last.mvKeys[i].angle = lastAngle[i]; Last.mvKeysUn[i].angle = lastAngle[i];
        MapPoint& p = mps[i];
        p.index = i; p.nObs = hasObs[i] ? 1 : 0;
        p.mWorldPos = Eigen::Vector3f(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        p.mDescriptor = cv::Mat(1, 32, CV_8UC1, (void*)(mpDesc + (size_t)i * 32), 32);
        if (valid[i]) Last.mvpMapPoints[i] = &p;
    }
    ORBmatcher matcher(0.9f, checkOrientation != 0);
    int n = matcher.SearchByProjection(Cur, Last, th, true);
    store_state(Cur, curMatch, curClaimed);
    return n;
}

int ref_search_for_initialization(int K1, const cv::KeyPoint* kps1, const uint8_t* desc1, int K2, const cv::KeyPoint* kps2,
                                  const uint8_t* desc2, const float* bounds, const float* scaleFactors, float* prevMatched, int windowSize,
                                  float nnratio, int checkOrientation, int* matches12) {
    Frame F1, F2;
    fill_frame(F1, K1, kps1, desc1, bounds, scaleFactors, 64);
    fill_frame(F2, K2, kps2, desc2, bounds, scaleFactors, 64);
    std::vector<cv::Point2f> prev(K1);
    for (int i = 0; i < K1; ++i) prev[i] = cv::Point2f(prevMatched[2 * i], prevMatched[2 * i + 1]);
    std::vector<int> m12;
    ORBmatcher matcher(nnratio, checkOrientation != 0);
    int n = matcher.SearchForInitialization(F1, F2, prev, m12, windowSize);
    for (int i = 0; i < K1; ++i) { matches12[i] = m12[i]; prevMatched[2 * i] = prev[i].x; prevMatched[2 * i + 1] = prev[i].y; }
    return n;
}

// Frame::ComputeStereoMatches (src/Frame.cc:811-982) on two of the reference's own extractors (handles of ref_orbx_create, each after its
// operator() on the left / right image of the pair): mvKeys / mvKeysRight + descriptors in, mvuRight / mvDepth out.
void ref_stereo_matches(void* exLeft, void* exRight, int N, const cv::KeyPoint* kl, const uint8_t* dl, int Nr, const cv::KeyPoint* kr, const uint8_t* dr,
                        const float* scaleFactors, const float* invScaleFactors, int nlevels, float mb, float mbf, float* uRight, float* depth) {
    Frame F;
    F.N = N;
    F.mvKeys.assign(kl, kl + N); F.mvKeysRight.assign(kr, kr + Nr);
    F.mDescriptors = N ? cv::Mat(N, 32, CV_8UC1, (void*)dl, 32) : cv::Mat();
    F.mDescriptorsRight = Nr ? cv::Mat(Nr, 32, CV_8UC1, (void*)dr, 32) : cv::Mat();
    F.mvScaleFactors.assign(scaleFactors, scaleFactors + nlevels);
    F.mvInvScaleFactors.assign(invScaleFactors, invScaleFactors + nlevels);
    F.mb = mb; F.mbf = mbf;
    F.mpORBextractorLeft = (ORBextractor*)exLeft; F.mpORBextractorRight = (ORBextractor*)exRight;
    F.ComputeStereoMatches();
    for (int i = 0; i < N; ++i) { uRight[i] = F.mvuRight[i]; depth[i] = F.mvDepth[i]; }
}

// int ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches) (src/ORBmatcher.cc:223-425).  Feature vectors as
// parallel (node id, feature index) arrays; kfPoint[i]: 0 = no map point, 1 = map point, 2 = bad map point.  matchF[nF] = KF feature index or -1.
int ref_search_by_bow(int nKF, const cv::KeyPoint* kpsKF, const uint8_t* descKF, const uint8_t* kfPoint, int eKF, const int* fvNodeKF, const int* fvFeatKF,
                      int nF, const cv::KeyPoint* kpsF, const uint8_t* descF, int eF, const int* fvNodeF, const int* fvFeatF, float nnratio, int checkOri,
                      int* matchF) {
    KeyFrame KF;
    std::vector<MapPoint> mps(nKF);
    KF.mvpMapPoints.assign(nKF, (MapPoint*)nullptr);
    for (int i = 0; i < nKF; ++i) { mps[i].index = i; mps[i].mbBad = kfPoint[i] == 2; if (kfPoint[i]) KF.mvpMapPoints[i] = &mps[i]; }
    KF.mDescriptors = nKF ? cv::Mat(nKF, 32, CV_8UC1, (void*)descKF, 32) : cv::Mat();
    KF.mvKeysUn.assign(kpsKF, kpsKF + nKF); KF.mvKeys = KF.mvKeysUn;
    for (int e = 0; e < eKF; ++e) KF.mFeatVec.addFeature(fvNodeKF[e], fvFeatKF[e]);
    Frame F;
    F.N = nF; F.Nleft = -1;
    F.mDescriptors = nF ? cv::Mat(nF, 32, CV_8UC1, (void*)descF, 32) : cv::Mat();
    F.mvKeys.assign(kpsF, kpsF + nF);
    for (int e = 0; e < eF; ++e) F.mFeatVec.addFeature(fvNodeF[e], fvFeatF[e]);
    std::vector<MapPoint*> out;
    ORBmatcher matcher(nnratio, checkOri != 0);
    const int n = matcher.SearchByBoW(&KF, F, out);
    for (int i = 0; i < nF; ++i) matchF[i] = out[i] ? out[i]->index : -1;
    return n;
}

// int ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12) (src/ORBmatcher.cc:765-905; loop closing / merging).
// point1 / point2 [N]: 0 no map point, 1 map point, 2 bad.  match12 [N1]: feature of KF2 whose map point is vpMatches12[idx1], -1 = NULL.
int ref_search_by_bow_kf(int N1, const cv::KeyPoint* kps1, const uint8_t* desc1, const uint8_t* point1, int E1, const int* fvNode1, const int* fvFeat1,
                         int N2, const cv::KeyPoint* kps2, const uint8_t* desc2, const uint8_t* point2, int E2, const int* fvNode2, const int* fvFeat2,
                         float nnratio, int checkOri, int* match12) {
    KeyFrame KF1, KF2;
    std::vector<MapPoint> mps1(N1), mps2(N2);
    auto fill = [](KeyFrame& KF, std::vector<MapPoint>& mps, int N, const cv::KeyPoint* kps, const uint8_t* desc, const uint8_t* point, int E, const int* fvNode, const int* fvFeat) {
        KF.N = N; KF.NLeft = -1;
        KF.mvpMapPoints.assign(N, (MapPoint*)nullptr);
        for (int i = 0; i < N; ++i) { mps[i].index = i; mps[i].mbBad = point[i] == 2; if (point[i]) KF.mvpMapPoints[i] = &mps[i]; }
        KF.mDescriptors = N ? cv::Mat(N, 32, CV_8UC1, (void*)desc, 32) : cv::Mat();
        KF.mvKeysUn.assign(kps, kps + N); KF.mvKeys = KF.mvKeysUn;
        for (int e = 0; e < E; ++e) KF.mFeatVec.addFeature(fvNode[e], fvFeat[e]);
    };
    fill(KF1, mps1, N1, kps1, desc1, point1, E1, fvNode1, fvFeat1);
    fill(KF2, mps2, N2, kps2, desc2, point2, E2, fvNode2, fvFeat2);
    std::vector<MapPoint*> out;
    ORBmatcher matcher(nnratio, checkOri != 0);
    const int n = matcher.SearchByBoW(&KF1, &KF2, out);
    for (int i = 0; i < N1; ++i) match12[i] = out[i] ? out[i]->index : -1;
    return n;
}

// int ORBmatcher::SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, vMatchedPairs, bOnlyStereo = false, bCoarse) (src/ORBmatcher.cc:907-1146),
// monocular keyframes (Pinhole::epipolarConstrain, src/CameraModels/Pinhole.cpp:107-129).  hasMP[i]: the keyframe holds a map point at feature i.
// matches12 [N1]: vMatchedPairs scattered back (-1 = none).  epOut / F12Out: the epipole and fundamental matrix the body works with, recomputed here
// with the same stand-in operators in the same order (the C-ABI takes both as inputs: the caller forms them with its own Eigen code).
static void fill_tri_kf(KeyFrame& KF, std::vector<MapPoint>& mps, Pinhole& camera, int N, const cv::KeyPoint* kps, const uint8_t* desc, const uint8_t* hasMP, int E,
                        const int* fvNode, const int* fvFeat, const float* scaleFactors, const float* levelSigma2, int nlevels, const float* Tcw, const float* cam) {
    KF.N = N; KF.NLeft = -1;
    KF.mvpMapPoints.assign(N, (MapPoint*)nullptr);                 // mps: N stand-in map points, pre-sized by the caller (not movable)
    for (int i = 0; i < N; ++i) if (hasMP[i]) KF.mvpMapPoints[i] = &mps[i];
    KF.mDescriptors = N ? cv::Mat(N, 32, CV_8UC1, (void*)desc, 32) : cv::Mat();
    KF.mvKeysUn.assign(kps, kps + N); KF.mvKeys = KF.mvKeysUn;
    KF.mvuRight.assign(N, -1.f);
    for (int e = 0; e < E; ++e) KF.mFeatVec.addFeature(fvNode[e], fvFeat[e]);
    KF.mvScaleFactors.assign(scaleFactors, scaleFactors + nlevels); KF.mvLevelSigma2.assign(levelSigma2, levelSigma2 + nlevels);
    KF.mTcw.qw = Tcw[0]; KF.mTcw.qx = Tcw[1]; KF.mTcw.qy = Tcw[2]; KF.mTcw.qz = Tcw[3]; KF.mTcw.t = Eigen::Vector3f(Tcw[4], Tcw[5], Tcw[6]);
    KF.mOw = KF.mTcw.inverse().translation();             // GetCameraCenter() = mTwc.translation() (src/KeyFrame.cc:143-147)
    camera.mvParameters.assign(cam, cam + 4); KF.mpCamera = &camera;
}
int ref_search_for_triangulation(int N1, const cv::KeyPoint* kps1, const uint8_t* desc1, const uint8_t* hasMP1, int E1, const int* fvNode1, const int* fvFeat1,
                                 int N2, const cv::KeyPoint* kps2, const uint8_t* desc2, const uint8_t* hasMP2, int E2, const int* fvNode2, const int* fvFeat2,
                                 const float* scaleFactors, const float* levelSigma2, int nlevels, const float* T1w, const float* T2w, const float* cam1,
                                 const float* cam2, int bCoarse, int checkOri, int* matches12, float* epOut, float* F12Out) {
    KeyFrame KF1, KF2;
    std::vector<MapPoint> mps1(N1), mps2(N2);
    Pinhole c1, c2;
    fill_tri_kf(KF1, mps1, c1, N1, kps1, desc1, hasMP1, E1, fvNode1, fvFeat1, scaleFactors, levelSigma2, nlevels, T1w, cam1);
    fill_tri_kf(KF2, mps2, c2, N2, kps2, desc2, hasMP2, E2, fvNode2, fvFeat2, scaleFactors, levelSigma2, nlevels, T2w, cam2);
    {
        const Eigen::Vector3f C2 = KF2.GetPose() * KF1.GetCameraCenter();
        const Eigen::Vector2f ep = KF2.mpCamera->project(C2);
        epOut[0] = ep(0); epOut[1] = ep(1);
        const Sophus::SE3f T12 = KF1.GetPose() * KF2.GetPoseInverse();
        const Eigen::Matrix3f F12 = c1.toK_().transpose().inverse() * Sophus::SO3f::hat(T12.translation()) * T12.rotationMatrix() * c2.toK_().inverse();
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) F12Out[3 * i + j] = F12(i, j);
    }
    std::vector<std::pair<size_t, size_t>> pairs;
    ORBmatcher matcher(0.6f, checkOri != 0);
    const int n = matcher.SearchForTriangulation(&KF1, &KF2, pairs, false, bCoarse != 0);
    for (int i = 0; i < N1; ++i) matches12[i] = -1;
    for (auto& pr : pairs) matches12[pr.first] = (int)pr.second;
    return n;
}

// int ORBmatcher::Fuse(KeyFrame* pKF, const vector<MapPoint*>& vpMapPoints, th, bRight = false) (src/ORBmatcher.cc:1148-1338), monocular keyframe.
// Map points: state[i] (0 NULL, 1 ok, 2 bad, 3 already observed by the keyframe), world position, normal, RAW min / max distance, descriptor, nObs.
// kfPoint[k]: -1 = keyframe keypoint k has no map point, else the Observations() of the map point it holds (>= 0); kfPointBad[k].
// Out per map point: action (0 none, 1 added as observation of keypoint idx, 2 replaced by the keyframe's point at idx, 3 replaces the keyframe's
// point at idx, 4 counted but nothing done: the keyframe's point at that keypoint is bad) and idx; returns nFused.
int ref_fuse(int K, const cv::KeyPoint* kps, const uint8_t* desc, const float* bounds, const float* scaleFactors, const float* invLevelSigma2, int nlevels,
             float logScaleFactor, const float* Tcw, const float* Ow, const float* cam, const int* kfPoint, const uint8_t* kfPointBad, int M,
             const uint8_t* state, const float* xyz, const float* normal, const float* minDistance, const float* maxDistance, const uint8_t* mpDesc,
             const int* mpObs, float th, int* action, int* actionIdx) {
    KeyFrame KF;
    KF.N = K; KF.NLeft = -1;
    KF.mvKeysUn.assign(kps, kps + K); KF.mvKeys = KF.mvKeysUn;
    KF.mvuRight.assign(K, -1.f);
    KF.mDescriptors = K ? cv::Mat(K, 32, CV_8UC1, (void*)desc, 32) : cv::Mat();
    KF.mnMinX = bounds[0]; KF.mnMinY = bounds[1]; KF.mnMaxX = bounds[2]; KF.mnMaxY = bounds[3];
    KF.mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / (KF.mnMaxX - KF.mnMinX);
    KF.mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / (KF.mnMaxY - KF.mnMinY);
    KF.mvScaleFactors.assign(scaleFactors, scaleFactors + nlevels); KF.mvInvLevelSigma2.assign(invLevelSigma2, invLevelSigma2 + nlevels);
    KF.mnScaleLevels = nlevels; KF.mfLogScaleFactor = logScaleFactor;
    KF.fx = cam[0]; KF.fy = cam[1]; KF.cx = cam[2]; KF.cy = cam[3]; KF.mbf = 0.f;
    Pinhole camera; camera.mvParameters.assign(cam, cam + 4); KF.mpCamera = &camera;
    KF.mTcw.qw = Tcw[0]; KF.mTcw.qx = Tcw[1]; KF.mTcw.qy = Tcw[2]; KF.mTcw.qz = Tcw[3]; KF.mTcw.t = Eigen::Vector3f(Tcw[4], Tcw[5], Tcw[6]);
    KF.mOw = Eigen::Vector3f(Ow[0], Ow[1], Ow[2]);
    // the keyframe's grid = the frame's (KeyFrame copies Frame::mGrid, src/KeyFrame.cc:60-75): built with the reference's AssignFeaturesToGrid
    {
        Frame F;
        fill_frame(F, K, kps, desc, bounds, scaleFactors, nlevels);
        KF.mGrid.assign(FRAME_GRID_COLS, std::vector<std::vector<size_t>>(FRAME_GRID_ROWS));
        for (int i = 0; i < FRAME_GRID_COLS; ++i) for (int j = 0; j < FRAME_GRID_ROWS; ++j) KF.mGrid[i][j] = F.mGrid[i][j];
        KF.mGridRight = KF.mGrid;
    }
    std::vector<MapPoint> kfMps(K);
    KF.mvpMapPoints.assign(K, (MapPoint*)nullptr);
    for (int k = 0; k < K; ++k) if (kfPoint[k] >= 0) { kfMps[k].index = k; kfMps[k].slot = k; kfMps[k].slotKF = &KF; kfMps[k].inKF.insert(&KF); kfMps[k].nObs = kfPoint[k]; kfMps[k].mbBad = kfPointBad[k] != 0; KF.mvpMapPoints[k] = &kfMps[k]; }
    std::vector<MapPoint> mps(M);
    std::vector<MapPoint*> vp(M, (MapPoint*)nullptr);
    for (int i = 0; i < M; ++i) {
        MapPoint& p = mps[i];
        p.index = i; p.mbBad = state[i] == 2; p.nObs = mpObs[i];
        if (state[i] == 3) p.inKF.insert(&KF);
        p.mWorldPos = Eigen::Vector3f(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        p.mNormalVector = Eigen::Vector3f(normal[3 * i], normal[3 * i + 1], normal[3 * i + 2]);
        p.mfMinDistance = minDistance[i]; p.mfMaxDistance = maxDistance[i];
        p.mDescriptor = cv::Mat(1, 32, CV_8UC1, (void*)(mpDesc + (size_t)i * 32), 32);
        if (state[i]) vp[i] = &p;
    }
    ORBmatcher matcher(0.6f, true);
    const int nFused = matcher.Fuse(&KF, vp, th, false);
    for (int i = 0; i < M; ++i) { action[i] = 0; actionIdx[i] = -1; }
    for (int i = 0; i < M; ++i) { action[i] = mps[i].fuseAction; actionIdx[i] = mps[i].fuseIdx; }
    return nFused;
}

// ---- the two Sim3 members (loop closing / map merging) ----
static void fill_grid_kf(KeyFrame& KF, Pinhole& camera, int K, const cv::KeyPoint* kps, const uint8_t* desc, const float* bounds, const float* scaleFactors, int nlevels,
                         float logScaleFactor, const float* cam) {
    KF.N = K; KF.NLeft = -1;
    KF.mvKeysUn.assign(kps, kps + K); KF.mvKeys = KF.mvKeysUn;
    KF.mvuRight.assign(K, -1.f);
    KF.mDescriptors = K ? cv::Mat(K, 32, CV_8UC1, (void*)desc, 32) : cv::Mat();
    KF.mnMinX = bounds[0]; KF.mnMinY = bounds[1]; KF.mnMaxX = bounds[2]; KF.mnMaxY = bounds[3];
    KF.mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / (KF.mnMaxX - KF.mnMinX);
    KF.mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / (KF.mnMaxY - KF.mnMinY);
    KF.mvScaleFactors.assign(scaleFactors, scaleFactors + nlevels);
    KF.mnScaleLevels = nlevels; KF.mfLogScaleFactor = logScaleFactor;
    KF.fx = cam[0]; KF.fy = cam[1]; KF.cx = cam[2]; KF.cy = cam[3]; KF.mbf = 0.f;
    camera.mvParameters.assign(cam, cam + 4); KF.mpCamera = &camera;
    Frame F;
    fill_frame(F, K, kps, desc, bounds, scaleFactors, nlevels);
    KF.mGrid.assign(FRAME_GRID_COLS, std::vector<std::vector<size_t>>(FRAME_GRID_ROWS));
    for (int i = 0; i < FRAME_GRID_COLS; ++i) for (int j = 0; j < FRAME_GRID_ROWS; ++j) KF.mGrid[i][j] = F.mGrid[i][j];
    KF.mGridRight = KF.mGrid;
    KF.mvpMapPoints.assign(K, (MapPoint*)nullptr);
}
static void set_pose(Sophus::SE3f& T, const float* p) { T.qw = p[0]; T.qx = p[1]; T.qy = p[2]; T.qz = p[3]; T.t = Eigen::Vector3f(p[4], p[5], p[6]); }
static void fill_points(std::vector<MapPoint>& mps, int M, const uint8_t* state, const float* xyz, const float* normal, const float* minD, const float* maxD, const uint8_t* desc) {
    for (int i = 0; i < M; ++i) {
        MapPoint& p = mps[i];
        p.index = i; p.mbBad = state[i] == 2;
        p.mWorldPos = Eigen::Vector3f(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        if (normal) p.mNormalVector = Eigen::Vector3f(normal[3 * i], normal[3 * i + 1], normal[3 * i + 2]);
        p.mfMinDistance = minD[i]; p.mfMaxDistance = maxD[i];
        p.mDescriptor = cv::Mat(1, 32, CV_8UC1, (void*)(desc + (size_t)i * 32), 32);
    }
}
// int ORBmatcher::SearchBySim3(pKF1, pKF2, vpMatches12, S12, th) (src/ORBmatcher.cc:1457-1674).  The keyframes share camera and level tables; map
// point i of keyframe k sits at its feature i (state 0 none, 1 good, 2 bad).  pre12 [N1]: -1 or the KF2 feature whose map point already is
// vpMatches12[i].  S12 = (s, qw, qx, qy, qz, tx, ty, tz).  Out: match12 [N1] = KF2 feature of vpMatches12[i] afterwards; pc2of1 [N1][3] /
// pc1of2 [N2][3]: the map points in the OTHER keyframe's camera frame as the body computes them (inputs of the C-ABI); returns nFound.
int ref_search_by_sim3(int nlevels, const float* scaleFactors, float logScaleFactor, const float* bounds, const float* cam,
                       int N1, const cv::KeyPoint* kps1, const uint8_t* desc1, const float* T1w, const uint8_t* state1, const float* xyz1, const float* min1, const float* max1, const uint8_t* mpDesc1,
                       int N2, const cv::KeyPoint* kps2, const uint8_t* desc2, const float* T2w, const uint8_t* state2, const float* xyz2, const float* min2, const float* max2, const uint8_t* mpDesc2,
                       const float* S12v, float th, const int* pre12, int* match12, float* pc2of1, float* pc1of2) {
    KeyFrame KF1, KF2; Pinhole c1, c2;
    fill_grid_kf(KF1, c1, N1, kps1, desc1, bounds, scaleFactors, nlevels, logScaleFactor, cam);
    fill_grid_kf(KF2, c2, N2, kps2, desc2, bounds, scaleFactors, nlevels, logScaleFactor, cam);
    set_pose(KF1.mTcw, T1w); set_pose(KF2.mTcw, T2w);
    std::vector<MapPoint> mps1(N1), mps2(N2);
    fill_points(mps1, N1, state1, xyz1, nullptr, min1, max1, mpDesc1);
    fill_points(mps2, N2, state2, xyz2, nullptr, min2, max2, mpDesc2);
    for (int i = 0; i < N1; ++i) if (state1[i]) { KF1.mvpMapPoints[i] = &mps1[i]; mps1[i].mObservations[&KF1] = std::make_tuple(i, -1); }
    for (int i = 0; i < N2; ++i) if (state2[i]) { KF2.mvpMapPoints[i] = &mps2[i]; mps2[i].mObservations[&KF2] = std::make_tuple(i, -1); }
    Sophus::Sim3f S12;
    S12.s = S12v[0]; S12.qw = S12v[1]; S12.qx = S12v[2]; S12.qy = S12v[3]; S12.qz = S12v[4]; S12.t = Eigen::Vector3f(S12v[5], S12v[6], S12v[7]);
    {
        const Sophus::Sim3f S21 = S12.inverse();
        for (int i = 0; i < N1; ++i) { const Eigen::Vector3f p = S21 * (KF1.GetPose() * mps1[i].mWorldPos); pc2of1[3 * i] = p(0); pc2of1[3 * i + 1] = p(1); pc2of1[3 * i + 2] = p(2); }
        for (int i = 0; i < N2; ++i) { const Eigen::Vector3f p = S12 * (KF2.GetPose() * mps2[i].mWorldPos); pc1of2[3 * i] = p(0); pc1of2[3 * i + 1] = p(1); pc1of2[3 * i + 2] = p(2); }
    }
    std::vector<MapPoint*> vpMatches12(N1, (MapPoint*)nullptr);
    for (int i = 0; i < N1; ++i) if (pre12[i] >= 0) vpMatches12[i] = &mps2[pre12[i]];
    ORBmatcher matcher(0.75f, true);
    const int n = matcher.SearchBySim3(&KF1, &KF2, vpMatches12, S12, th);
    for (int i = 0; i < N1; ++i) match12[i] = vpMatches12[i] ? vpMatches12[i]->index : -1;
    return n;
}
// int ORBmatcher::Fuse(KeyFrame* pKF, Sophus::Sim3f& Scw, const vector<MapPoint*>& vpPoints, float th, vector<MapPoint*>& vpReplacePoint)
// (src/ORBmatcher.cc:1340-1455).  Points: state 1 ok, 2 bad, 3 already among the keyframe's map points (as kfPoint index: the point IS the
// keyframe's point at feature found3[i]).  kfPoint [K]: 0 none, 1 good, 2 bad map point at that keyframe feature.  Out: action [M] 0 none,
// 1 AddObservation at idx, 2 vpReplacePoint[i] = the keyframe's point at idx, 4 counted only (bad point there); Tcw7 / Ow: the decomposition the
// body works with; returns nFused.
int ref_fuse_sim3(int K, const cv::KeyPoint* kps, const uint8_t* desc, const float* bounds, const float* scaleFactors, int nlevels, float logScaleFactor, const float* cam,
                  const uint8_t* kfPoint, const float* Scwv, int M, const uint8_t* state, const float* xyz, const float* normal, const float* minD, const float* maxD,
                  const uint8_t* mpDesc, float th, int* action, int* actionIdx, float* Tcw7, float* Ow3) {
    KeyFrame KF; Pinhole cam0;
    fill_grid_kf(KF, cam0, K, kps, desc, bounds, scaleFactors, nlevels, logScaleFactor, cam);
    std::vector<MapPoint> kfMps(K), mps(M);
    for (int k = 0; k < K; ++k) if (kfPoint[k]) { kfMps[k].index = k; kfMps[k].mbBad = kfPoint[k] == 2; KF.mvpMapPoints[k] = &kfMps[k]; }
    fill_points(mps, M, state, xyz, normal, minD, maxD, mpDesc);
    std::vector<MapPoint*> vp(M), repl(M, (MapPoint*)nullptr);
    int nextFree = 0;
    for (int i = 0; i < M; ++i) {
        vp[i] = &mps[i];
        if (state[i] == 3) {                                    // the very object the keyframe already holds: hand in the keyframe's own point
            while (nextFree < K && kfPoint[nextFree] != 1) ++nextFree;
            if (nextFree < K) vp[i] = &kfMps[nextFree++];
        }
    }
    Sophus::Sim3f Scw;
    Scw.s = Scwv[0]; Scw.qw = Scwv[1]; Scw.qx = Scwv[2]; Scw.qy = Scwv[3]; Scw.qz = Scwv[4]; Scw.t = Eigen::Vector3f(Scwv[5], Scwv[6], Scwv[7]);
    {
        const Sophus::SE3f Tcw = Sophus::SE3f(Scw.rotationMatrix(), Scw.translation() / Scw.scale());
        const Eigen::Vector3f Ow = Tcw.inverse().translation();
        Tcw7[0] = Tcw.qw; Tcw7[1] = Tcw.qx; Tcw7[2] = Tcw.qy; Tcw7[3] = Tcw.qz; Tcw7[4] = Tcw.t(0); Tcw7[5] = Tcw.t(1); Tcw7[6] = Tcw.t(2);
        Ow3[0] = Ow(0); Ow3[1] = Ow(1); Ow3[2] = Ow(2);
    }
    ORBmatcher matcher(0.75f, true);
    const int nFused = matcher.Fuse(&KF, Scw, vp, th, repl);
    for (int i = 0; i < M; ++i) {
        action[i] = 0; actionIdx[i] = -1;
        if (vp[i] != &mps[i]) continue;
        if (mps[i].fuseAction == 1) { action[i] = 1; actionIdx[i] = mps[i].fuseIdx; }
        else if (repl[i]) { action[i] = 2; actionIdx[i] = (repl[i] >= kfMps.data() && repl[i] < kfMps.data() + K) ? repl[i]->index : repl[i]->fuseIdx; }
    }
    return nFused;
}

// void MapPoint::ComputeDistinctiveDescriptors() (src/MapPoint.cc:329-403) for one map point with n observations (descriptor rows); returns the
// index of the chosen observation (the reference stores a clone of that row in mDescriptor).
int ref_distinctive_descriptor(int n, const uint8_t* desc) {
    std::vector<KeyFrame> kfs(n);
    MapPoint mp;
    for (int i = 0; i < n; ++i) {
        kfs[i].mDescriptors = cv::Mat(1, 32, CV_8UC1, (void*)(desc + (size_t)i * 32), 32);
        mp.mObservations[&kfs[i]] = std::make_tuple(0, -1);       // the map iterates in pointer order = array order
    }
    mp.mDescriptor = cv::Mat();
    mp.ComputeDistinctiveDescriptors();
    if (mp.mDescriptor.empty()) return -1;
    for (int i = 0; i < n; ++i) if (!memcmp(mp.mDescriptor.data, desc + (size_t)i * 32, 32)) return i;   // first row with the chosen bytes
    return -2;
}

// minDistance / maxDistance are the RAW mfMinDistance / mfMaxDistance: the 0.8f / 1.2f invariance factors are applied by the
// reference's own getters (src/MapPoint.cc:502-512).
void ref_is_in_frustum(int M, const float* P, const float* N, const float* minDistance, const float* maxDistance,
                       const float* Rcw, const float* tcw, const float* Ow, const float* cam, const float* bounds, float mbf,
                       float logScaleFactor, int nScaleLevels, float viewingCosLimit, uint8_t* inView, float* projX, float* projY,
                       float* projXR, float* depth, int* level, float* viewCos) {
    Frame F;
    F.Nleft = -1;
    Pinhole camera;
    camera.mvParameters.assign(cam, cam + 4);
    F.mpCamera = &camera;
    for (int i = 0; i < 9; ++i) F.mRcw.m[i] = Rcw[i];
    F.mtcw = Eigen::Vector3f(tcw[0], tcw[1], tcw[2]);
    F.mOw = Eigen::Vector3f(Ow[0], Ow[1], Ow[2]);
    F.mnMinX = bounds[0]; F.mnMinY = bounds[1]; F.mnMaxX = bounds[2]; F.mnMaxY = bounds[3];
    F.mbf = mbf; F.mfLogScaleFactor = logScaleFactor; F.mnScaleLevels = nScaleLevels;
    for (int i = 0; i < M; ++i) {
        MapPoint p;
        p.mWorldPos = Eigen::Vector3f(P[3 * i], P[3 * i + 1], P[3 * i + 2]);
        p.mNormalVector = Eigen::Vector3f(N[3 * i], N[3 * i + 1], N[3 * i + 2]);
        p.mfMaxDistance = maxDistance[i];
        p.mfMinDistance = minDistance[i];
        // fields the reference leaves untouched for points out of view: documented defaults of orbm_frustum_project
        p.mTrackProjXR = 0.f; p.mTrackDepth = 0.f; p.mnTrackScaleLevel = -1; p.mTrackViewCos = 0.f;
        inView[i] = F.isInFrustum(&p, viewingCosLimit) ? 1 : 0;
        projX[i] = p.mTrackProjX; projY[i] = p.mTrackProjY; projXR[i] = p.mTrackProjXR; depth[i] = p.mTrackDepth;
        level[i] = p.mnTrackScaleLevel; viewCos[i] = p.mTrackViewCos;
    }
}

}  // extern "C"
