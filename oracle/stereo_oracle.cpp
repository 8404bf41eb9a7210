// TEST INFRASTRUCTURE ONLY (see oracle_common.h).  CPU restatement of Frame::ComputeStereoMatches (reference src/Frame.cc:811-982),
// the stereo consumer of ORBextractor::mvImagePyramid.  Pinned by oracle/_ref, which compiles the reference's
// own function body (tests/test_ref_pins_oracle_cpu.py).
#include "oracle_common.h"

#include <algorithm>
#include <climits>
#include <cmath>
#include <utility>
#include <vector>

using namespace orbo;

extern "C" {

// planesL / planesR: the unbordered pyramid planes mvImagePyramid[l] of the left / right extractor (tightly packed, widths w[l], heights h[l]).
// The 11 x 11 SAD windows never leave the planes for keypoints the extractor can produce (x, y >= 19 at their own level): returns -1 if one does.
int orbo_stereo_matches(int nlevels, const uint8_t* const* planesL, const uint8_t* const* planesR, const int* w, const int* h, int N, const KeyPoint* kl,
                        const uint8_t* dl, int Nr, const KeyPoint* kr, const uint8_t* dr, const float* scaleFactors, const float* invScaleFactors, float mb,
                        float mbf, float* uRight, float* depth) {
    static const int TH_HIGH = 100, TH_LOW = 50;
    for (int i = 0; i < N; ++i) { uRight[i] = -1.0f; depth[i] = -1.0f; }
    const int thOrbDist = (TH_HIGH + TH_LOW) / 2;
    const int nRows = h[0];
    std::vector<std::vector<size_t>> vRowIndices(nRows);
    for (int iR = 0; iR < Nr; ++iR) {
        const float kpY = kr[iR].y;
        const float r = 2.0f * scaleFactors[kr[iR].octave];
        const int maxr = (int)std::ceil(kpY + r), minr = (int)std::floor(kpY - r);
        for (int yi = minr; yi <= maxr; ++yi) {
            if (yi < 0 || yi >= nRows) return -1;      // the reference indexes without a check
            vRowIndices[yi].push_back(iR);
        }
    }
    const float minZ = mb, minD = 0, maxD = mbf / minZ;
    std::vector<std::pair<int, int>> vDistIdx;
    auto hamming = [](const uint8_t* a, const uint8_t* b) { int d = 0; for (int i = 0; i < 32; ++i) d += __builtin_popcount(a[i] ^ b[i]); return d; };
    for (int iL = 0; iL < N; ++iL) {
        const KeyPoint& kpL = kl[iL];
        const int levelL = kpL.octave;
        const float vL = kpL.y, uL = kpL.x;
        const std::vector<size_t>& vCandidates = vRowIndices[(size_t)vL];
        if (vCandidates.empty()) continue;
        const float minU = uL - maxD, maxU = uL - minD;
        if (maxU < 0) continue;
        int bestDist = TH_HIGH;
        size_t bestIdxR = 0;
        for (size_t iC = 0; iC < vCandidates.size(); ++iC) {
            const size_t iR = vCandidates[iC];
            const KeyPoint& kpR = kr[iR];
            if (kpR.octave < levelL - 1 || kpR.octave > levelL + 1) continue;
            const float uR = kpR.x;
            if (uR >= minU && uR <= maxU) {
                const int dist = hamming(dl + (size_t)iL * 32, dr + iR * 32);
                if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
            }
        }
        if (bestDist < thOrbDist) {
            const float uR0 = kr[bestIdxR].x;
            const float scaleFactor = invScaleFactors[kpL.octave];
            const float scaleduL = std::round(kpL.x * scaleFactor), scaledvL = std::round(kpL.y * scaleFactor), scaleduR0 = std::round(uR0 * scaleFactor);
            const int wdw = 5, L = 5;
            const int lw = w[kpL.octave], lh = h[kpL.octave];
            const uint8_t* IL = planesL[kpL.octave]; const uint8_t* IR = planesR[kpL.octave];
            const int y0 = (int)(scaledvL - wdw), xL0 = (int)(scaleduL - wdw);
            int best = INT_MAX, bestincR = 0;
            float vDists[2 * 5 + 1];
            const float iniu = scaleduR0 + L - wdw, endu = scaleduR0 + L + wdw + 1;
            if (iniu < 0 || endu >= lw) continue;
            if (y0 < 0 || y0 + 2 * wdw + 1 > lh || xL0 < 0 || xL0 + 2 * wdw + 1 > lw || (int)(scaleduR0 - L - wdw) < 0) return -1;
            for (int incR = -L; incR <= +L; ++incR) {
                const int xR0 = (int)(scaleduR0 + incR - wdw);
                long s = 0;
                for (int r = 0; r < 2 * wdw + 1; ++r)
                    for (int c = 0; c < 2 * wdw + 1; ++c) {
                        const int a = IL[(size_t)(y0 + r) * lw + xL0 + c], b = IR[(size_t)(y0 + r) * lw + xR0 + c];
                        s += a > b ? a - b : b - a;
                    }
                const float dist = (float)(double)s;            // float dist = cv::norm(IL, IR, cv::NORM_L1)
                if (dist < best) { best = (int)dist; bestincR = incR; }
                vDists[L + incR] = dist;
            }
            if (bestincR == -L || bestincR == L) continue;
            const float dist1 = vDists[L + bestincR - 1], dist2 = vDists[L + bestincR], dist3 = vDists[L + bestincR + 1];
            const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
            if (deltaR < -1 || deltaR > 1) continue;
            float bestuR = scaleFactors[kpL.octave] * ((float)scaleduR0 + (float)bestincR + deltaR);
            float disparity = (uL - bestuR);
            if (disparity >= minD && disparity < maxD) {
                if (disparity <= 0) { disparity = 0.01; bestuR = uL - 0.01; }
                depth[iL] = mbf / disparity;
                uRight[iL] = bestuR;
                vDistIdx.push_back(std::pair<int, int>(best, iL));
            }
        }
    }
    if (vDistIdx.empty()) return 0;                         // the reference reads vDistIdx[0] of an empty vector here (undefined)
    std::sort(vDistIdx.begin(), vDistIdx.end());
    const float median = vDistIdx[vDistIdx.size() / 2].first;
    const float thDist = 1.5f * 1.4f * median;
    for (int i = (int)vDistIdx.size() - 1; i >= 0; --i) {
        if (vDistIdx[i].first < thDist) break;
        uRight[vDistIdx[i].second] = -1;
        depth[vDistIdx[i].second] = -1;
    }
    return (int)vDistIdx.size();
}

}  // extern "C"
