// TEST INFRASTRUCTURE ONLY (see oracle_common.h).  CPU restatement of the per-frame matchers:
//   ORBmatcher::DescriptorDistance                       reference src/ORBmatcher.cc:2058-2074
//   ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th, bFar, thFar)   :43-213 (mono branch)
//   ORBmatcher::RadiusByViewingCos                       :215-221
//   ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono)                     :1676-1887 (mono branch)
//   ORBmatcher::ComputeThreeMaxima                       :2012-2053
//   Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea   src/Frame.cc:385-416,725-735,657-723
//   Pinhole::project(Vector3f)                           src/CameraModels/Pinhole.cpp:43-49
//   Sophus SO3f * point, SE3f * point                    Thirdparty/Sophus/sophus/so3.hpp:358-367, se3.hpp:321-324
//   cv::BFMatcher(NORM_HAMMING).knnMatch(k=2)            use at src/Frame.cc:1144 (pinned vs cv2 in tests)
//   Frame::isInFrustum + MapPoint::PredictScale          src/Frame.cc:512-574, src/MapPoint.cc:531-546
// PARITY UNPINNED by the reference (no vectors, not buildable here) for everything but the brute-force matcher.
//
// Floating point contract: every float operation below rounds individually, in the written order
// (-ffp-contract=off).  The reference evaluates the pose transform through Eigen expression templates
// whose contraction under -O3 -march=native cannot be reproduced without Eigen; that last-ulp choice is
// "parity unpinned" (DESIGN.md).  Everything after the projection is integer / comparison logic.
#include "oracle_common.h"

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

namespace orbo {

static const int TH_HIGH = 100;       // src/ORBmatcher.cc:35
static const int HISTO_LENGTH = 30;   // :37
static const int GRID_COLS = 64, GRID_ROWS = 48;   // include/Frame.h:44-45

static int descriptor_distance(const uint8_t* a, const uint8_t* b) {
    const uint32_t* pa = (const uint32_t*)a;
    const uint32_t* pb = (const uint32_t*)b;
    int dist = 0;
    for (int i = 0; i < 8; ++i) {
        uint32_t v = pa[i] ^ pb[i];
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

struct FrameView {
    int K;
    const KeyPoint* kps;     // mvKeysUn (== mvKeys with zero distortion)
    const uint8_t* desc;     // K x 32
    float minX, minY, maxX, maxY, gridWInv, gridHInv;
    const float* scaleFactors;
    std::vector<int> grid[GRID_COLS][GRID_ROWS];

    void build_grid() {   // Frame::AssignFeaturesToGrid
        for (int i = 0; i < K; ++i) {
            int px = (int)std::round((kps[i].x - minX) * gridWInv);
            int py = (int)std::round((kps[i].y - minY) * gridHInv);
            if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) continue;
            grid[px][py].push_back(i);
        }
    }
    // Frame::GetFeaturesInArea (mono)
    void features_in_area(float x, float y, float r, int minLevel, int maxLevel, std::vector<int>& out) const {
        out.clear();
        const float factorX = r, factorY = r;
        const int nMinCellX = std::max(0, (int)std::floor((x - minX - factorX) * gridWInv));
        if (nMinCellX >= GRID_COLS) return;
        const int nMaxCellX = std::min(GRID_COLS - 1, (int)std::ceil((x - minX + factorX) * gridWInv));
        if (nMaxCellX < 0) return;
        const int nMinCellY = std::max(0, (int)std::floor((y - minY - factorY) * gridHInv));
        if (nMinCellY >= GRID_ROWS) return;
        const int nMaxCellY = std::min(GRID_ROWS - 1, (int)std::ceil((y - minY + factorY) * gridHInv));
        if (nMaxCellY < 0) return;
        const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
        for (int ix = nMinCellX; ix <= nMaxCellX; ++ix)
            for (int iy = nMinCellY; iy <= nMaxCellY; ++iy)
                for (int idx : grid[ix][iy]) {
                    const KeyPoint& kp = kps[idx];
                    if (bCheckLevels) {
                        if (kp.octave < minLevel) continue;
                        if (maxLevel >= 0 && kp.octave > maxLevel) continue;
                    }
                    const float dx = kp.x - x, dy = kp.y - y;
                    if (std::fabs(dx) < factorX && std::fabs(dy) < factorY) out.push_back(idx);
                }
    }
};

}  // namespace orbo

using namespace orbo;

extern "C" {

int orbo_descriptor_distance(const uint8_t* a, const uint8_t* b) { return descriptor_distance(a, b); }

// Local-map matcher.  Per map point: inView (mbTrackInView), bad, depth (mTrackDepth), projX/projY,
// level (mnTrackScaleLevel), viewCos, hasObs (Observations()>0), descriptor.
// curMatch[K] in/out: map-point index assigned to each keypoint (-1 none); curClaimed[K] in/out: the assigned
// map point has Observations()>0 (such keypoints are skipped, :84-86).
int orbo_search_local_map(int K, const KeyPoint* kps, const uint8_t* desc, const float* bounds /*minX,minY,maxX,maxY*/,
                          const float* scaleFactors, int M, const uint8_t* inView, const uint8_t* bad, const float* depth,
                          const float* projX, const float* projY, const int* level, const float* viewCos, const uint8_t* hasObs,
                          const uint8_t* mpDesc, float th, float nnratio, int bFarPoints, float thFarPoints,
                          int* curMatch, uint8_t* curClaimed) {
    FrameView F;
    F.K = K; F.kps = kps; F.desc = desc;
    F.minX = bounds[0]; F.minY = bounds[1]; F.maxX = bounds[2]; F.maxY = bounds[3];
    F.gridWInv = (float)GRID_COLS / (F.maxX - F.minX);
    F.gridHInv = (float)GRID_ROWS / (F.maxY - F.minY);
    F.scaleFactors = scaleFactors;
    F.build_grid();
    int nmatches = 0;
    const bool bFactor = th != 1.0;
    std::vector<int> vIndices;
    for (int i = 0; i < M; ++i) {
        if (!inView[i]) continue;
        if (bFarPoints && depth[i] > thFarPoints) continue;
        if (bad[i]) continue;
        const int lvl = level[i];
        float r = viewCos[i] > 0.998 ? 2.5f : 4.0f;   // RadiusByViewingCos (float vs double literal compare)
        if (bFactor) r *= th;
        F.features_in_area(projX[i], projY[i], r * scaleFactors[lvl], lvl - 1, lvl, vIndices);
        if (vIndices.empty()) continue;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int idx : vIndices) {
            if (curMatch[idx] >= 0 && curClaimed[idx]) continue;
            const int dist = descriptor_distance(mpDesc + (size_t)i * 32, desc + (size_t)idx * 32);
            if (dist < bestDist) {
                bestDist2 = bestDist; bestDist = dist;
                bestLevel2 = bestLevel; bestLevel = kps[idx].octave;
                bestIdx = idx;
            } else if (dist < bestDist2) {
                bestLevel2 = kps[idx].octave;
                bestDist2 = dist;
            }
        }
        if (bestDist <= TH_HIGH) {
            if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
            curMatch[bestIdx] = i;
            curClaimed[bestIdx] = hasObs[i];
            ++nmatches;
        }
    }
    return nmatches;
}

// Last-frame matcher (mono).  Tcw = (qw,qx,qy,qz,tx,ty,tz), cam = (fx,fy,cx,cy).
// Per last-frame index i: valid (pMP && !outlier), world position, octave, angle, hasObs, descriptor of the map point.
int orbo_search_last_frame(int K, const KeyPoint* kps, const uint8_t* desc, const float* bounds, const float* scaleFactors,
                           const float* Tcw, const float* cam, int M, const uint8_t* valid, const float* xyz,
                           const int* lastOctave, const float* lastAngle, const uint8_t* hasObs, const uint8_t* mpDesc,
                           float th, int checkOrientation, int* curMatch, uint8_t* curClaimed) {
    FrameView F;
    F.K = K; F.kps = kps; F.desc = desc;
    F.minX = bounds[0]; F.minY = bounds[1]; F.maxX = bounds[2]; F.maxY = bounds[3];
    F.gridWInv = (float)GRID_COLS / (F.maxX - F.minX);
    F.gridHInv = (float)GRID_ROWS / (F.maxY - F.minY);
    F.scaleFactors = scaleFactors;
    F.build_grid();
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    const float qw = Tcw[0], qx = Tcw[1], qy = Tcw[2], qz = Tcw[3];
    std::vector<int> vIndices;
    for (int i = 0; i < M; ++i) {
        if (!valid[i]) continue;
        const float px = xyz[3 * i], py = xyz[3 * i + 1], pz = xyz[3 * i + 2];
        // SO3 * p: uv = q.vec x p; uv += uv; p + w*uv + q.vec x uv   (so3.hpp:358-367)
        float ux = qy * pz - qz * py, uy = qz * px - qx * pz, uz = qx * py - qy * px;
        ux += ux; uy += uy; uz += uz;
        const float cx_ = qy * uz - qz * uy, cy_ = qz * ux - qx * uz, cz_ = qx * uy - qy * ux;
        const float xc = (px + qw * ux) + cx_ + Tcw[4];
        const float yc = (py + qw * uy) + cy_ + Tcw[5];
        const float zc = (pz + qw * uz) + cz_ + Tcw[6];
        const float invzc = (float)(1.0 / zc);
        if (invzc < 0) continue;
        const float u = cam[0] * xc / zc + cam[2];
        const float v = cam[1] * yc / zc + cam[3];
        if (u < F.minX || u > F.maxX) continue;
        if (v < F.minY || v > F.maxY) continue;
        const int oct = lastOctave[i];
        const float radius = th * scaleFactors[oct];
        F.features_in_area(u, v, radius, oct - 1, oct + 1, vIndices);
        if (vIndices.empty()) continue;
        int bestDist = 256, bestIdx2 = -1;
        for (int i2 : vIndices) {
            if (curMatch[i2] >= 0 && curClaimed[i2]) continue;
            const int dist = descriptor_distance(mpDesc + (size_t)i * 32, desc + (size_t)i2 * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= TH_HIGH) {
            curMatch[bestIdx2] = i;
            curClaimed[bestIdx2] = hasObs[i];
            ++nmatches;
            if (checkOrientation) {
                float rot = lastAngle[i] - kps[bestIdx2].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)std::round(rot * factor);
                if (bin == HISTO_LENGTH) bin = 0;
                rotHist[bin].push_back(bestIdx2);
            }
        }
    }
    if (checkOrientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1, max1 = 0, max2 = 0, max3 = 0;   // ComputeThreeMaxima
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            const int s = (int)rotHist[i].size();
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
            else if (s > max3) { max3 = s; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
        for (int i = 0; i < HISTO_LENGTH; ++i)
            if (i != ind1 && i != ind2 && i != ind3)
                for (int idx : rotHist[i]) { curMatch[idx] = -1; curClaimed[idx] = 0; --nmatches; }
    }
    return nmatches;
}

// cv::BFMatcher(NORM_HAMMING).knnMatch(query, train, k=2): exact top-2 by (distance, lower train index).
void orbo_bf_knn2(const uint8_t* q, int Q, const uint8_t* t, int T, int* idx, int* dist) {
    for (int i = 0; i < Q; ++i) {
        int b1 = 1 << 30, i1 = -1, b2 = 1 << 30, i2 = -1;
        for (int j = 0; j < T; ++j) {
            const int d = descriptor_distance(q + (size_t)i * 32, t + (size_t)j * 32);
            if (d < b1) { b2 = b1; i2 = i1; b1 = d; i1 = j; }
            else if (d < b2) { b2 = d; i2 = j; }
        }
        idx[2 * i] = i1; idx[2 * i + 1] = i2;
        dist[2 * i] = i1 >= 0 ? b1 : -1; dist[2 * i + 1] = i2 >= 0 ? b2 : -1;
    }
}

// Frame::GetFeaturesInArea alone (for unit tests of the grid)
int orbo_features_in_area(int K, const KeyPoint* kps, const float* bounds, float x, float y, float r, int minLevel, int maxLevel,
                          int* out, int cap) {
    FrameView F;
    F.K = K; F.kps = kps; F.desc = nullptr;
    F.minX = bounds[0]; F.minY = bounds[1]; F.maxX = bounds[2]; F.maxY = bounds[3];
    F.gridWInv = (float)GRID_COLS / (F.maxX - F.minX);
    F.gridHInv = (float)GRID_ROWS / (F.maxY - F.minY);
    F.build_grid();
    std::vector<int> v;
    F.features_in_area(x, y, r, minLevel, maxLevel, v);
    for (size_t i = 0; i < v.size() && (int)i < cap; ++i) out[i] = v[i];
    return (int)v.size();
}

// ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:648-763): windowed search between the two frames of the monocular
// initialiser (only level-0 keypoints of F1; candidates of F2 at level 0 inside windowSize around vbPrevMatched[i1]; best / second
// best among candidates whose current match is not at least as good (vMatchedDistance), TH_LOW, ratio test, re-assignment of an
// already matched F2 keypoint, rotation histogram on F1 indices).  prevMatched (K1 x 2) is updated like vbPrevMatched (:757-759).
// Oracle only so far: the sm_100a kernel for this init-time function is the first item of the next round.
int orbo_search_for_initialization(int K1, const KeyPoint* kps1, const uint8_t* desc1, int K2, const KeyPoint* kps2, const uint8_t* desc2,
                                   const float* bounds, const float* scaleFactors, float* prevMatched, int windowSize, float nnratio,
                                   int checkOrientation, int* matches12) {
    static const int TH_LOW = 50;   // src/ORBmatcher.cc:36
    FrameView F2;
    F2.K = K2; F2.kps = kps2; F2.desc = desc2;
    F2.minX = bounds[0]; F2.minY = bounds[1]; F2.maxX = bounds[2]; F2.maxY = bounds[3];
    F2.gridWInv = (float)GRID_COLS / (F2.maxX - F2.minX);
    F2.gridHInv = (float)GRID_ROWS / (F2.maxY - F2.minY);
    F2.scaleFactors = scaleFactors;
    F2.build_grid();
    int nmatches = 0;
    for (int i = 0; i < K1; ++i) matches12[i] = -1;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    std::vector<int> vMatchedDistance(K2, INT32_MAX), vnMatches21(K2, -1), vIndices2;
    for (int i1 = 0; i1 < K1; ++i1) {
        const int level1 = kps1[i1].octave;
        if (level1 > 0) continue;
        F2.features_in_area(prevMatched[2 * i1], prevMatched[2 * i1 + 1], (float)windowSize, level1, level1, vIndices2);
        if (vIndices2.empty()) continue;
        int bestDist = INT32_MAX, bestDist2 = INT32_MAX, bestIdx2 = -1;
        for (int i2 : vIndices2) {
            const int dist = descriptor_distance(desc1 + (size_t)i1 * 32, desc2 + (size_t)i2 * 32);
            if (vMatchedDistance[i2] <= dist) continue;
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist <= TH_LOW && bestDist < (float)bestDist2 * nnratio) {
            if (vnMatches21[bestIdx2] >= 0) { matches12[vnMatches21[bestIdx2]] = -1; --nmatches; }
            matches12[i1] = bestIdx2;
            vnMatches21[bestIdx2] = i1;
            vMatchedDistance[bestIdx2] = bestDist;
            ++nmatches;
            if (checkOrientation) {
                float rot = kps1[i1].angle - kps2[bestIdx2].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)std::round(rot * factor);
                if (bin == HISTO_LENGTH) bin = 0;
                rotHist[bin].push_back(i1);
            }
        }
    }
    if (checkOrientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1, max1 = 0, max2 = 0, max3 = 0;   // ComputeThreeMaxima
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            const int sz = (int)rotHist[i].size();
            if (sz > max1) { max3 = max2; max2 = max1; max1 = sz; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (sz > max2) { max3 = max2; max2 = sz; ind3 = ind2; ind2 = i; }
            else if (sz > max3) { max3 = sz; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int idx1 : rotHist[i])
                if (matches12[idx1] >= 0) { matches12[idx1] = -1; --nmatches; }
        }
    }
    for (int i1 = 0; i1 < K1; ++i1)
        if (matches12[i1] >= 0) { prevMatched[2 * i1] = kps2[matches12[i1]].x; prevMatched[2 * i1 + 1] = kps2[matches12[i1]].y; }
    return nmatches;
}

// Frame::isInFrustum, monocular branch (src/Frame.cc:512-574), + MapPoint::PredictScale(const float&, Frame*) (src/MapPoint.cc:531-546)
// for M map points.  Eigen's 3-vector products are written out in one fixed order with every operation rounded individually
// (the reference's -O3 -march=native build may contract them differently: unpinned at the last ulp, like the last-frame
// transform above); `log(ratio)` is std::log(float) = glibc logf; Pinhole::project(Vector3f) is src/CameraModels/Pinhole.cpp:35-41.
// `uv(0) - mbf*invz` (:565) is one fused multiply-add in the reference's -O3 -march=native build (found by comparing with
// oracle/_ref, which compiles the reference's own text with contraction on): written as fmaf here.
// Outputs as documented in include/orb_b200.h (orbm_frustum_project).
void orbo_is_in_frustum(int M, const float* P, const float* N, const float* minDistInv, const float* maxDistInv, const float* maxDistance,
                        const float* Rcw, const float* tcw, const float* Ow, const float* cam, const float* bounds, float mbf,
                        float logScaleFactor, int nScaleLevels, float viewingCosLimit,
                        uint8_t* inView, float* projX, float* projY, float* projXR, float* depth, int* level, float* viewCos) {
    for (int i = 0; i < M; ++i) {
        inView[i] = 0; projX[i] = -1.f; projY[i] = -1.f; projXR[i] = 0.f; depth[i] = 0.f; level[i] = -1; viewCos[i] = 0.f;   // :515-517
        const float X = P[3 * i], Y = P[3 * i + 1], Z = P[3 * i + 2];
        const float xc = ((Rcw[0] * X + Rcw[1] * Y) + Rcw[2] * Z) + tcw[0];                 // :523 Pc = mRcw * P + mtcw
        const float yc = ((Rcw[3] * X + Rcw[4] * Y) + Rcw[5] * Z) + tcw[1];
        const float zc = ((Rcw[6] * X + Rcw[7] * Y) + Rcw[8] * Z) + tcw[2];
        const float pcDist = sqrtf((xc * xc + yc * yc) + zc * zc);                          // :524
        const float invz = 1.0f / zc;                                                       // :528
        if (zc < 0.0f) continue;                                                            // :529
        const float u = cam[0] * xc / zc + cam[2], v = cam[1] * yc / zc + cam[3];           // :532 mpCamera->project(Pc)
        if (u < bounds[0] || u > bounds[2]) continue;                                       // :534
        if (v < bounds[1] || v > bounds[3]) continue;                                       // :536
        projX[i] = u; projY[i] = v;                                                         // :539-540
        const float ox = X - Ow[0], oy = Y - Ow[1], oz = Z - Ow[2];                         // :545
        const float dist = sqrtf((ox * ox + oy * oy) + oz * oz);                            // :546
        if (dist < minDistInv[i] || dist > maxDistInv[i]) continue;                         // :548
        const float c = ((ox * N[3 * i] + oy * N[3 * i + 1]) + oz * N[3 * i + 2]) / dist;   // :554
        if (c < viewingCosLimit) continue;                                                  // :556
        const float ratio = maxDistance[i] / dist;                                          // MapPoint.cc:536
        int nScale = (int)std::ceil(std::log(ratio) / logScaleFactor);                      // :539 (float overloads)
        if (nScale < 0) nScale = 0;
        else if (nScale >= nScaleLevels) nScale = nScaleLevels - 1;
        inView[i] = 1; projXR[i] = fmaf(-mbf, invz, u); depth[i] = pcDist; level[i] = nScale; viewCos[i] = c;   // :563-571
    }
}

// ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches) (src/ORBmatcher.cc:223-425), monocular branch
// (F.Nleft == -1, pKF->mpCamera2 == NULL): matching restricted to features of the same vocabulary node; a frame feature that already holds a
// match is skipped by later keyframe features (:275-276); TH_LOW, ratio test, rotation histogram on frame indices.
// Feature vectors: parallel (node id, feature index) arrays in DBoW2::FeatureVector order.  kfPoint[i]: 0 no map point, 1 map point, 2 bad.
int orbo_search_by_bow(int nKF, const KeyPoint* kpsKF, const uint8_t* descKF, const uint8_t* kfPoint, int eKF, const int* fvNodeKF, const int* fvFeatKF,
                       int nF, const KeyPoint* kpsF, const uint8_t* descF, int eF, const int* fvNodeF, const int* fvFeatF, float nnratio, int checkOri,
                       int* matchF) {
    static const int TH_LOW = 50;
    for (int i = 0; i < nF; ++i) matchF[i] = -1;
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    int a = 0, b = 0;
    while (a < eKF && b < eF) {
        if (fvNodeKF[a] == fvNodeF[b]) {
            int a1 = a, b1 = b;
            while (a1 < eKF && fvNodeKF[a1] == fvNodeKF[a]) ++a1;
            while (b1 < eF && fvNodeF[b1] == fvNodeF[b]) ++b1;
            for (int iKF = a; iKF < a1; ++iKF) {
                const int realIdxKF = fvFeatKF[iKF];
                if (kfPoint[realIdxKF] != 1) continue;
                int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
                for (int iF = b; iF < b1; ++iF) {
                    const int realIdxF = fvFeatF[iF];
                    if (matchF[realIdxF] >= 0) continue;
                    const int dist = descriptor_distance(descKF + (size_t)realIdxKF * 32, descF + (size_t)realIdxF * 32);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = realIdxF; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
                if (bestDist1 <= TH_LOW && static_cast<float>(bestDist1) < nnratio * static_cast<float>(bestDist2)) {
                    matchF[bestIdxF] = realIdxKF;
                    if (checkOri) {
                        float rot = kpsKF[realIdxKF].angle - kpsF[bestIdxF].angle;
                        if (rot < 0.0) rot += 360.0f;
                        int bin = (int)std::round(rot * factor);
                        if (bin == HISTO_LENGTH) bin = 0;
                        rotHist[bin].push_back(bestIdxF);
                    }
                    ++nmatches;
                }
            }
            a = a1; b = b1;
        } else if (fvNodeKF[a] < fvNodeF[b]) { const int t = fvNodeF[b]; while (a < eKF && fvNodeKF[a] < t) ++a; }
        else { const int t = fvNodeKF[a]; while (b < eF && fvNodeF[b] < t) ++b; }
    }
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1, max1 = 0, max2 = 0, max3 = 0;
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            const int sz = (int)rotHist[i].size();
            if (sz > max1) { max3 = max2; max2 = max1; max1 = sz; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (sz > max2) { max3 = max2; max2 = sz; ind3 = ind2; ind2 = i; }
            else if (sz > max3) { max3 = sz; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int idx : rotHist[i]) { matchF[idx] = -1; --nmatches; }
        }
    }
    return nmatches;
}

// MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:329-403) for one map point: n observed descriptors -> index of the one with the
// least median Hamming distance to the others (first minimum; the median is element (int)(0.5 * (n - 1)) of the sorted row).
int orbo_distinctive_descriptor(int n, const uint8_t* desc) {
    if (n <= 0) return -1;
    int BestMedian = INT32_MAX, BestIdx = 0;
    std::vector<int> vDists(n);
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < n; ++j) vDists[j] = i == j ? 0 : descriptor_distance(desc + (size_t)i * 32, desc + (size_t)j * 32);
        std::sort(vDists.begin(), vDists.end());
        const int median = vDists[(size_t)(0.5 * (n - 1))];
        if (median < BestMedian) { BestMedian = median; BestIdx = i; }
    }
    return BestIdx;
}

}  // extern "C"
// mode 0: world point + keyframe pose + normal test (the two Fuse overloads; `gate` = the chi-square test of the plain one);
// mode 2: the point is given in the keyframe's camera frame (SearchBySim3, :1506-1560): depth, projection `fx*x+cx` with x = X * (float)(1.0/Z)
//         (one FMA in the reference build), distance = its norm, no normal test, no gate.
static void project_search_impl(int K, const KeyPoint* kps, const uint8_t* desc, const float* bounds, const float* scaleFactors, const float* invLevelSigma2, int nlevels,
                                float logScaleFactor, const float* Tcw, const float* Ow, const float* cam, int M, const uint8_t* state, const float* xyz,
                                const float* normal, const float* minDistance, const float* maxDistance, const uint8_t* mpDesc, float th, int mode, bool gate,
                                int* bestIdxOut, int* bestDistOut) {
    FrameView F;
    F.K = K; F.kps = kps; F.desc = desc;
    F.minX = bounds[0]; F.minY = bounds[1]; F.maxX = bounds[2]; F.maxY = bounds[3];
    F.gridWInv = (float)GRID_COLS / (F.maxX - F.minX);
    F.gridHInv = (float)GRID_ROWS / (F.maxY - F.minY);
    F.scaleFactors = scaleFactors;
    F.build_grid();
    std::vector<int> vIndices;
    for (int i = 0; i < M; ++i) {
        bestIdxOut[i] = -1; bestDistOut[i] = 256;
        if (state[i] != 1) continue;
        const float px = xyz[3 * i], py = xyz[3 * i + 1], pz = xyz[3 * i + 2];
        float u, v, dist3D;
        if (mode == 0) {
            const float qw = Tcw[0], qx = Tcw[1], qy = Tcw[2], qz = Tcw[3];
            float ux = qy * pz - qz * py, uy = qz * px - qx * pz, uz = qx * py - qy * px;
            ux += ux; uy += uy; uz += uz;
            const float cx_ = qy * uz - qz * uy, cy_ = qz * ux - qx * uz, cz_ = qx * uy - qy * ux;
            const float xc = (px + qw * ux) + cx_ + Tcw[4], yc = (py + qw * uy) + cy_ + Tcw[5], zc = (pz + qw * uz) + cz_ + Tcw[6];
            if (zc < 0.0f) continue;
            u = cam[0] * xc / zc + cam[2]; v = cam[1] * yc / zc + cam[3];
        } else {
            if (pz < 0.0) continue;
            const float invz = 1.0 / pz;
            const float x = px * invz, y = py * invz;
            u = fmaf(cam[0], x, cam[2]); v = fmaf(cam[1], y, cam[3]);
        }
        if (!(u >= F.minX && u < F.maxX && v >= F.minY && v < F.maxY)) continue;                    // KeyFrame::IsInImage
        const float maxD = 1.2f * maxDistance[i], minD = 0.8f * minDistance[i];
        if (mode == 0) {
            const float ox = px - Ow[0], oy = py - Ow[1], oz = pz - Ow[2];
            dist3D = sqrtf((ox * ox + oy * oy) + oz * oz);
            if (dist3D < minD || dist3D > maxD) continue;
            const float dot = (ox * normal[3 * i] + oy * normal[3 * i + 1]) + oz * normal[3 * i + 2];
            if (dot < 0.5 * dist3D) continue;
        } else {
            dist3D = sqrtf((px * px + py * py) + pz * pz);
            if (dist3D < minD || dist3D > maxD) continue;
        }
        const float ratio = maxDistance[i] / dist3D;
        int lvl = (int)std::ceil(std::log(ratio) / logScaleFactor);
        if (lvl < 0) lvl = 0; else if (lvl >= nlevels) lvl = nlevels - 1;
        const float radius = th * scaleFactors[lvl];
        F.features_in_area(u, v, radius, -1, -1, vIndices);          // KeyFrame::GetFeaturesInArea has no level filter
        int bestDist = 256, bestIdx = -1;
        for (int idx : vIndices) {
            const int kpLevel = kps[idx].octave;
            if (kpLevel < lvl - 1 || kpLevel > lvl) continue;
            if (gate) {
                const float ex = u - kps[idx].x, ey = v - kps[idx].y;
                const float e2 = fmaf(ex, ex, ey * ey);
                if (e2 * invLevelSigma2[kpLevel] > 5.99) continue;
            }
            const int dist = descriptor_distance(mpDesc + (size_t)i * 32, desc + (size_t)idx * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
        }
        bestIdxOut[i] = bestIdx; bestDistOut[i] = bestDist;
    }
}
extern "C" {
// ORBmatcher::Fuse(KeyFrame* pKF, const vector<MapPoint*>& vpMapPoints, th, bRight = false) (src/ORBmatcher.cc:1148-1338), monocular keyframe:
// the SEARCH of every map point (projection, image / distance / viewing-angle tests, scale prediction, radius search with the chi-square gate,
// best Hamming distance).  What Fuse then does with a hit (AddObservation / Replace, :1310-1330) mutates the pointer graph and stays with the
// caller; no map point's search depends on it.  state[i]: 0 NULL, 1 ok, 2 bad, 3 already in the keyframe.  bestIdx -1 / bestDist 256: no candidate.
// `ex*ex+ey*ey` (:1292) is one FMA in the reference's -O3 -march=native build (checked against oracle/_ref).
void orbo_fuse_search(int K, const KeyPoint* kps, const uint8_t* desc, const float* bounds, const float* scaleFactors, const float* invLevelSigma2, int nlevels,
                      float logScaleFactor, const float* Tcw, const float* Ow, const float* cam, int M, const uint8_t* state, const float* xyz,
                      const float* normal, const float* minDistance, const float* maxDistance, const uint8_t* mpDesc, float th, int* bestIdxOut, int* bestDistOut) {
    project_search_impl(K, kps, desc, bounds, scaleFactors, invLevelSigma2, nlevels, logScaleFactor, Tcw, Ow, cam, M, state, xyz, normal, minDistance, maxDistance, mpDesc, th,
                        0, true, bestIdxOut, bestDistOut);
}
// ORBmatcher::Fuse(KeyFrame* pKF, Sophus::Sim3f& Scw, vpPoints, th, vpReplacePoint) (src/ORBmatcher.cc:1340-1455): the search of the plain overload
// without the chi-square gate; Tcw / Ow are the caller's decomposition of Scw (:1349-1350).  The walk that fills vpReplacePoint / adds
// observations (:1436-1450) stays with the caller.
void orbo_fuse_search_sim3(int K, const KeyPoint* kps, const uint8_t* desc, const float* bounds, const float* scaleFactors, int nlevels, float logScaleFactor, const float* Tcw,
                           const float* Ow, const float* cam, int M, const uint8_t* state, const float* xyz, const float* normal, const float* minDistance,
                           const float* maxDistance, const uint8_t* mpDesc, float th, int* bestIdxOut, int* bestDistOut) {
    project_search_impl(K, kps, desc, bounds, scaleFactors, nullptr, nlevels, logScaleFactor, Tcw, Ow, cam, M, state, xyz, normal, minDistance, maxDistance, mpDesc, th,
                        0, false, bestIdxOut, bestDistOut);
}
// ORBmatcher::SearchBySim3 (src/ORBmatcher.cc:1457-1674): both directions + the agreement test.  pc2of1 [N1][3]: the map points of KF1 in KF2's
// camera frame (S21 * (T1w * p3Dw), :1507-1508), pc1of2 likewise (:1586-1587) -- Sophus expressions of the caller.  state: 0 no map point, 1 good,
// 2 bad; pre12 [N1]: -1 or the KF2 feature already matched (vbAlreadyMatched1/2, :1478-1491).  match12 [N1] = pre12 plus the new agreements; returns nFound.
int orbo_search_by_sim3(int nlevels, const float* scaleFactors, float logScaleFactor, const float* bounds, const float* cam,
                        int N1, const KeyPoint* kps1, const uint8_t* desc1, const uint8_t* state1, const float* pc2of1, const float* min1, const float* max1, const uint8_t* mpDesc1,
                        int N2, const KeyPoint* kps2, const uint8_t* desc2, const uint8_t* state2, const float* pc1of2, const float* min2, const float* max2, const uint8_t* mpDesc2,
                        float th, const int* pre12, int* match12) {
    std::vector<uint8_t> s1(N1), s2(N2);
    for (int i = 0; i < N1; ++i) s1[i] = state1[i] == 1 && pre12[i] < 0;
    for (int i = 0; i < N2; ++i) s2[i] = state2[i] == 1;
    for (int i = 0; i < N1; ++i) if (pre12[i] >= 0 && pre12[i] < N2) s2[pre12[i]] = 0;
    std::vector<int> m1(N1), d1(N1), m2(N2), d2(N2);
    project_search_impl(N2, kps2, desc2, bounds, scaleFactors, nullptr, nlevels, logScaleFactor, nullptr, nullptr, cam, N1, s1.data(), pc2of1, nullptr, min1, max1, mpDesc1, th,
                        2, false, m1.data(), d1.data());
    project_search_impl(N1, kps1, desc1, bounds, scaleFactors, nullptr, nlevels, logScaleFactor, nullptr, nullptr, cam, N2, s2.data(), pc1of2, nullptr, min2, max2, mpDesc2, th,
                        2, false, m2.data(), d2.data());
    int nFound = 0;
    for (int i1 = 0; i1 < N1; ++i1) {
        match12[i1] = pre12[i1];
        const int idx2 = d1[i1] <= TH_HIGH ? m1[i1] : -1;
        if (idx2 >= 0) {
            const int idx1 = d2[idx2] <= TH_HIGH ? m2[idx2] : -1;
            if (idx1 == i1) { match12[i1] = idx2; ++nFound; }
        }
    }
    return nFound;
}

// ORBmatcher::SearchForTriangulation(pKF1, pKF2, vMatchedPairs, bOnlyStereo = false, bCoarse) (src/ORBmatcher.cc:907-1146), monocular keyframes,
// with Pinhole::epipolarConstrain (src/CameraModels/Pinhole.cpp:107-129).  The epipole `ep` (:919) and the fundamental matrix F12 (Pinhole.cpp:112,
// constant per keyframe pair although the reference recomputes it per candidate) are INPUTS: they come out of Eigen/Sophus expression templates
// in the caller.  Only features WITHOUT a map point take part; a KF2 feature may be matched by several KF1 features (vbMatched2 is never set in
// the reference).  Among the candidates of the same vocabulary node that pass the epipole-distance and epipolar tests the smallest distance
// <= TH_LOW wins, the LAST one on ties (`dist>bestDist` is strict, :1008).  The scalar float expressions are contracted by the reference's
// -O3 -march=native build exactly as written here with fmaf (checked against oracle/_ref).
int orbo_search_for_triangulation(int N1, const KeyPoint* kps1, const uint8_t* desc1, const uint8_t* hasMP1, int E1, const int* fvNode1, const int* fvFeat1,
                                  int N2, const KeyPoint* kps2, const uint8_t* desc2, const uint8_t* hasMP2, int E2, const int* fvNode2, const int* fvFeat2,
                                  const float* scaleFactors2, const float* levelSigma2_2, const float* ep, const float* F12, int bCoarse, int checkOri,
                                  int* matches12) {
    static const int TH_LOW = 50;
    (void)N2;
    for (int i = 0; i < N1; ++i) matches12[i] = -1;
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    int a = 0, b = 0;
    while (a < E1 && b < E2) {
        if (fvNode1[a] == fvNode2[b]) {
            int a1 = a, b1 = b;
            while (a1 < E1 && fvNode1[a1] == fvNode1[a]) ++a1;
            while (b1 < E2 && fvNode2[b1] == fvNode2[b]) ++b1;
            for (int i1 = a; i1 < a1; ++i1) {
                const int idx1 = fvFeat1[i1];
                if (hasMP1[idx1]) continue;
                const KeyPoint& kp1 = kps1[idx1];
                int bestDist = TH_LOW, bestIdx2 = -1;
                for (int i2 = b; i2 < b1; ++i2) {
                    const int idx2 = fvFeat2[i2];
                    if (hasMP2[idx2]) continue;
                    const int dist = descriptor_distance(desc1 + (size_t)idx1 * 32, desc2 + (size_t)idx2 * 32);
                    if (dist > TH_LOW || dist > bestDist) continue;
                    const KeyPoint& kp2 = kps2[idx2];
                    const float distex = ep[0] - kp2.x, distey = ep[1] - kp2.y;
                    if (fmaf(distex, distex, distey * distey) < 100 * scaleFactors2[kp2.octave]) continue;
                    bool ok = bCoarse != 0;
                    if (!ok) {
                        const float ea = fmaf(kp1.x, F12[0], kp1.y * F12[3]) + F12[6];
                        const float eb = fmaf(kp1.x, F12[1], kp1.y * F12[4]) + F12[7];
                        const float ec = fmaf(kp1.x, F12[2], kp1.y * F12[5]) + F12[8];
                        const float num = fmaf(ea, kp2.x, eb * kp2.y) + ec;
                        const float den = fmaf(ea, ea, eb * eb);
                        if (den != 0) {
                            const float dsqr = num * num / den;
                            ok = dsqr < 3.84 * levelSigma2_2[kp2.octave];
                        }
                    }
                    if (ok) { bestIdx2 = idx2; bestDist = dist; }
                }
                if (bestIdx2 >= 0) {
                    matches12[idx1] = bestIdx2;
                    ++nmatches;
                    if (checkOri) {
                        float rot = kp1.angle - kps2[bestIdx2].angle;
                        if (rot < 0.0) rot += 360.0f;
                        int bin = (int)std::round(rot * factor);
                        if (bin == HISTO_LENGTH) bin = 0;
                        rotHist[bin].push_back(idx1);
                    }
                }
            }
            a = a1; b = b1;
        } else if (fvNode1[a] < fvNode2[b]) { const int t = fvNode2[b]; while (a < E1 && fvNode1[a] < t) ++a; }
        else { const int t = fvNode1[a]; while (b < E2 && fvNode2[b] < t) ++b; }
    }
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1, max1 = 0, max2 = 0, max3 = 0;
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            const int sz = (int)rotHist[i].size();
            if (sz > max1) { max3 = max2; max2 = max1; max1 = sz; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (sz > max2) { max3 = max2; max2 = sz; ind3 = ind2; ind2 = i; }
            else if (sz > max3) { max3 = sz; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int idx : rotHist[i]) { matches12[idx] = -1; --nmatches; }
        }
    }
    return nmatches;
}

// ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12) (src/ORBmatcher.cc:765-905), NLeft == -1: like the
// (KeyFrame, Frame) overload but both sides need a good map point, a KF2 feature is claimed once (vbMatched2), the threshold is STRICT
// (bestDist1 < TH_LOW, :859) and the result is indexed by the KF1 feature.  point1 / point2: 0 no map point, 1 map point, 2 bad.
int orbo_search_by_bow_kf(int N1, const KeyPoint* kps1, const uint8_t* desc1, const uint8_t* point1, int E1, const int* fvNode1, const int* fvFeat1,
                          int N2, const KeyPoint* kps2, const uint8_t* desc2, const uint8_t* point2, int E2, const int* fvNode2, const int* fvFeat2,
                          float nnratio, int checkOri, int* match12) {
    static const int TH_LOW = 50;
    for (int i = 0; i < N1; ++i) match12[i] = -1;
    std::vector<uint8_t> vbMatched2(N2, 0);
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    int a = 0, b = 0;
    while (a < E1 && b < E2) {
        if (fvNode1[a] == fvNode2[b]) {
            int a1 = a, b1 = b;
            while (a1 < E1 && fvNode1[a1] == fvNode1[a]) ++a1;
            while (b1 < E2 && fvNode2[b1] == fvNode2[b]) ++b1;
            for (int i1 = a; i1 < a1; ++i1) {
                const int idx1 = fvFeat1[i1];
                if (point1[idx1] != 1) continue;
                int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
                for (int i2 = b; i2 < b1; ++i2) {
                    const int idx2 = fvFeat2[i2];
                    if (vbMatched2[idx2] || point2[idx2] != 1) continue;
                    const int dist = descriptor_distance(desc1 + (size_t)idx1 * 32, desc2 + (size_t)idx2 * 32);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = idx2; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
                if (bestDist1 < TH_LOW && static_cast<float>(bestDist1) < nnratio * static_cast<float>(bestDist2)) {
                    match12[idx1] = bestIdx2;
                    vbMatched2[bestIdx2] = 1;
                    if (checkOri) {
                        float rot = kps1[idx1].angle - kps2[bestIdx2].angle;
                        if (rot < 0.0) rot += 360.0f;
                        int bin = (int)std::round(rot * factor);
                        if (bin == HISTO_LENGTH) bin = 0;
                        rotHist[bin].push_back(idx1);
                    }
                    ++nmatches;
                }
            }
            a = a1; b = b1;
        } else if (fvNode1[a] < fvNode2[b]) { const int t = fvNode2[b]; while (a < E1 && fvNode1[a] < t) ++a; }
        else { const int t = fvNode1[a]; while (b < E2 && fvNode2[b] < t) ++b; }
    }
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1, max1 = 0, max2 = 0, max3 = 0;
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            const int sz = (int)rotHist[i].size();
            if (sz > max1) { max3 = max2; max2 = max1; max1 = sz; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (sz > max2) { max3 = max2; max2 = sz; ind3 = ind2; ind2 = i; }
            else if (sz > max3) { max3 = sz; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int idx : rotHist[i]) { match12[idx] = -1; --nmatches; }
        }
    }
    return nmatches;
}

}  // extern "C"
