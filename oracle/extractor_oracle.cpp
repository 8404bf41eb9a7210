// TEST INFRASTRUCTURE ONLY (see oracle_common.h).  CPU restatement of
// ORBextractor (reference src/ORBextractor.cc, include/ORBextractor.h) plus the
// five OpenCV primitives it calls.  Every function cites the
// reference lines it follows.  Pinned by (tests/): the primitives bit for bit
// against cv2 4.13 (test_oracle_cpu.py + golden fixtures); the pyramid chain and the
// per-cell FAST loop against a cv2 transcription (test_cells_vs_cv2_cpu.py); IC_Angle
// against the definition + cv2.fastAtan2 (test_orientation_cpu.py); the steered BRIEF
// against numpy and cv2.ORB (test_brief_vs_cv2_cpu.py); DistributeOctTree against a second
// transcription with the real std::sort (test_quadtree_transcription_cpu.py).  PARITY
// UNPINNED by the reference (it ships no vectors and cannot be built here).  Build: see oracle/Makefile (-O3 -ffp-contract=off; the one
// place where the reference's -march=native build fuses a multiply-add, the
// BRIEF sample rotation, is written with an explicit fmaf()).
#include "oracle_common.h"

#include <algorithm>
#include <cstring>
#include <list>
#include <utility>

namespace orbo {

static const int PATCH_SIZE = 31;       // src/ORBextractor.cc:71
static const int HALF_PATCH_SIZE = 15;  // :72
static const int EDGE_THRESHOLD = 19;   // :73

static const signed char kPattern[1024] = {
#include "../orb_slam3_modified_b200/csrc/brief_pattern.inc"
};

struct Image {
    int w = 0, h = 0;
    std::vector<uint8_t> d;
    uint8_t at(int y, int x) const { return d[(size_t)y * w + x]; }
};

// ---------------------------------------------------------------------------
// cv::resize(INTER_LINEAR), CV_8UC1 (call site src/ORBextractor.cc:1183).
// OpenCV imgproc/resize.cpp: fixed-point coefficients (11 bits), horizontal pass
// in int32 without rounding, vertical pass with the >>4, >>16, +2, >>2 chain.
// Exact 2x decimation is promoted to INTER_AREA by cv::resize itself.
// ---------------------------------------------------------------------------
static void resize_linear(const uint8_t* src, int sw, int sh, int sstep, uint8_t* dst, int dw, int dh, int dstep) {
    if (sw == 2 * dw && sh == 2 * dh) {  // is_area_fast && iscale == 2 -> INTER_AREA 2x2 mean
        for (int y = 0; y < dh; ++y)
            for (int x = 0; x < dw; ++x) {
                const uint8_t* p = src + (size_t)(2 * y) * sstep + 2 * x;
                dst[(size_t)y * dstep + x] = (uint8_t)((p[0] + p[1] + p[sstep] + p[sstep + 1] + 2) >> 2);
            }
        return;
    }
    const double scale_x = (double)sw / dw, scale_y = (double)sh / dh;
    std::vector<int> xofs(dw), yofs(dh);
    std::vector<short> ialpha(2 * dw), ibeta(2 * dh);
    for (int dx = 0; dx < dw; ++dx) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = cvFloor(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        ialpha[2 * dx] = (short)cvRoundf((1.f - fx) * 2048);
        ialpha[2 * dx + 1] = (short)cvRoundf(fx * 2048);
    }
    for (int dy = 0; dy < dh; ++dy) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = cvFloor(fy);
        fy -= sy;
        yofs[dy] = sy;
        ibeta[2 * dy] = (short)cvRoundf((1.f - fy) * 2048);
        ibeta[2 * dy + 1] = (short)cvRoundf(fy * 2048);
    }
    std::vector<int> row0(dw), row1(dw);
    for (int dy = 0; dy < dh; ++dy) {
        int sy0 = std::min(std::max(yofs[dy], 0), sh - 1);
        int sy1 = std::min(std::max(yofs[dy] + 1, 0), sh - 1);
        const uint8_t* s0 = src + (size_t)sy0 * sstep;
        const uint8_t* s1 = src + (size_t)sy1 * sstep;
        for (int dx = 0; dx < dw; ++dx) {
            int sx = xofs[dx];
            int sx1 = std::min(sx + 1, sw - 1);
            int a0 = ialpha[2 * dx], a1 = ialpha[2 * dx + 1];
            row0[dx] = s0[sx] * a0 + s0[sx1] * a1;
            row1[dx] = s1[sx] * a0 + s1[sx1] * a1;
        }
        int b0 = ibeta[2 * dy], b1 = ibeta[2 * dy + 1];
        for (int dx = 0; dx < dw; ++dx) {
            int v = (((b0 * (row0[dx] >> 4)) >> 16) + ((b1 * (row1[dx] >> 4)) >> 16) + 2) >> 2;
            dst[(size_t)dy * dstep + dx] = (uint8_t)v;
        }
    }
}

static inline int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        else i = 2 * n - 2 - i;
    }
    return i;
}

// ---------------------------------------------------------------------------
// cv::GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) on CV_8UC1 (call site
// src/ORBextractor.cc:1133).  OpenCV's 8-bit path uses the fixed-point kernel
// [18,34,48,56,48,34,18]/256; horizontal pass exact (8.8), vertical 16.16 with
// round-to-nearest (+32768 >> 16).
// ---------------------------------------------------------------------------
static void blur7(const uint8_t* src, int w, int h, int sstep, uint8_t* dst, int dstep) {
    static const int k[7] = {18, 34, 48, 56, 48, 34, 18};
    std::vector<uint16_t> hbuf((size_t)w * h);
    for (int y = 0; y < h; ++y) {
        const uint8_t* r = src + (size_t)y * sstep;
        uint16_t* hb = &hbuf[(size_t)y * w];
        for (int x = 0; x < w; ++x) {
            if (x >= 3 && x + 3 < w) {   // interior: no reflection
                hb[x] = (uint16_t)(18 * (r[x - 3] + r[x + 3]) + 34 * (r[x - 2] + r[x + 2]) + 48 * (r[x - 1] + r[x + 1]) + 56 * r[x]);
            } else {
                int s = 0;
                for (int i = 0; i < 7; ++i) s += k[i] * r[reflect101(x + i - 3, w)];
                hb[x] = (uint16_t)s;
            }
        }
    }
    for (int y = 0; y < h; ++y) {
        const uint16_t* rows[7];
        for (int i = 0; i < 7; ++i) rows[i] = &hbuf[(size_t)reflect101(y + i - 3, h) * w];
        uint8_t* d = dst + (size_t)y * dstep;
        for (int x = 0; x < w; ++x) {
            const uint32_t s = 18u * (rows[0][x] + rows[6][x]) + 34u * (rows[1][x] + rows[5][x]) + 48u * (rows[2][x] + rows[4][x]) + 56u * rows[3][x];
            d[x] = (uint8_t)((s + 32768u) >> 16);
        }
    }
}

// ---------------------------------------------------------------------------
// cv::FAST(roi, kps, T, nonmaxSuppression=true), TYPE_9_16 (call sites
// src/ORBextractor.cc:826,845).  OpenCV features2d/fast.cpp + fast_score.cpp.
// ---------------------------------------------------------------------------
static const int kRing[16][2] = {{0, 3}, {1, 3}, {2, 2}, {3, 1}, {3, 0}, {3, -1}, {2, -2}, {1, -3},
                                 {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};

// max over the 16 cyclic 9-arcs of max(min(d), min(-d)); corner at T <=> result > T; score = result-1.
static int fast_arc_max(const uint8_t* p, int step) {
    int d[25];
    int v = p[0];
    for (int k = 0; k < 16; ++k) d[k] = v - p[kRing[k][1] * step + kRing[k][0]];
    for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
    int best = -1000;
    for (int k = 0; k < 16; ++k) {
        int mn = d[k], mx = d[k];
        for (int i = 1; i < 9; ++i) { mn = std::min(mn, d[k + i]); mx = std::max(mx, d[k + i]); }
        best = std::max(best, std::max(mn, -mx));
    }
    return best;
}

struct Corner { int x, y, score; };

static void fast9_nms(const uint8_t* roi, int w, int h, int step, int T, std::vector<Corner>& out) {
    out.clear();
    if (w < 7 || h < 7) return;
    std::vector<uint8_t> sc((size_t)w * h, 0);
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            // early rejection as in OpenCV's FAST_t (a 9-arc contains ring pixel k or k+8 for every k): exact, only saves time
            const uint8_t* p = roi + (size_t)y * step + x;
            const int v = p[0], lo = v - T, hi = v + T;
            const int p0 = p[3 * step], p8 = p[-3 * step], p4 = p[3], p12 = p[-3];
            const bool dark = (p0 < lo || p8 < lo) && (p4 < lo || p12 < lo);
            const bool brig = (p0 > hi || p8 > hi) && (p4 > hi || p12 > hi);
            if (!dark && !brig) continue;
            int m = fast_arc_max(p, step);
            if (m > T) sc[(size_t)y * w + x] = (uint8_t)(m - 1);
        }
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            int s = sc[(size_t)y * w + x];
            if (!s) continue;  // T >= 1 here so score >= 1 for every corner
            bool keep = true;
            for (int dy = -1; dy <= 1 && keep; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    if (!dx && !dy) continue;
                    if (s <= sc[(size_t)(y + dy) * w + x + dx]) { keep = false; break; }
                }
            if (keep) out.push_back({x, y, s});
        }
}

// ---------------------------------------------------------------------------
// cv::fastAtan2 (call site src/ORBextractor.cc:102).  OpenCV core/mathfuncs_core:
// degree-7 odd polynomial, float32, each operation individually rounded.
// ---------------------------------------------------------------------------
static float fast_atan2(float y, float x) {
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    const float eps = (float)2.2204460492503131e-16;
    float ax = std::fabs(x), ay = std::fabs(y), a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + eps);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + eps);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// ---------------------------------------------------------------------------
// ORBextractor
// ---------------------------------------------------------------------------
struct Node {  // ExtractorNode, include/ORBextractor.h:31-41
    int ULx, ULy, URx, URy, BLx, BLy, BRx, BRy;
    std::vector<KeyPoint> keys;
    std::list<Node>::iterator lit;
    bool noMore = false;
};

// ExtractorNode::DivideNode, src/ORBextractor.cc:480-536
static void divide_node(const Node& n, Node& n1, Node& n2, Node& n3, Node& n4) {
    const int halfX = (int)std::ceil(static_cast<float>(n.URx - n.ULx) / 2);
    const int halfY = (int)std::ceil(static_cast<float>(n.BRy - n.ULy) / 2);
    n1.ULx = n.ULx; n1.ULy = n.ULy;
    n1.URx = n.ULx + halfX; n1.URy = n.ULy;
    n1.BLx = n.ULx; n1.BLy = n.ULy + halfY;
    n1.BRx = n.ULx + halfX; n1.BRy = n.ULy + halfY;
    n2.ULx = n1.URx; n2.ULy = n1.URy;
    n2.URx = n.URx; n2.URy = n.URy;
    n2.BLx = n1.BRx; n2.BLy = n1.BRy;
    n2.BRx = n.URx; n2.BRy = n.ULy + halfY;
    n3.ULx = n1.BLx; n3.ULy = n1.BLy;
    n3.URx = n1.BRx; n3.URy = n1.BRy;
    n3.BLx = n.BLx; n3.BLy = n.BLy;
    n3.BRx = n1.BRx; n3.BRy = n.BLy;
    n4.ULx = n3.URx; n4.ULy = n3.URy;
    n4.URx = n2.BRx; n4.URy = n2.BRy;
    n4.BLx = n3.BRx; n4.BLy = n3.BRy;
    n4.BRx = n.BRx; n4.BRy = n.BRy;
    for (const KeyPoint& kp : n.keys) {
        if (kp.x < n1.URx) {
            if (kp.y < n1.BRy) n1.keys.push_back(kp);
            else n3.keys.push_back(kp);
        } else if (kp.y < n1.BRy) n2.keys.push_back(kp);
        else n4.keys.push_back(kp);
    }
    n1.noMore = n1.keys.size() == 1;
    n2.noMore = n2.keys.size() == 1;
    n3.noMore = n3.keys.size() == 1;
    n4.noMore = n4.keys.size() == 1;
}

typedef std::pair<int, Node*> SizeNode;
// compareNodes, src/ORBextractor.cc:538-553 (ties left to std::sort's order)
static bool compare_nodes(SizeNode& a, SizeNode& b) {
    if (a.first < b.first) return true;
    if (a.first > b.first) return false;
    return a.second->ULx < b.second->ULx;
}

struct Extractor {
    int nfeatures, nlevels, iniTh, minTh;
    double scaleFactor;  // the reference stores the float argument in a double member (ORBextractor.h:92)
    std::vector<float> scale, invScale, sigma2, invSigma2;
    std::vector<int> featPerLevel, umax;
    std::vector<Image> pyr;
    std::vector<std::vector<KeyPoint>> cand, kept;  // debug taps (per level)

    // ORBextractor::ORBextractor, src/ORBextractor.cc:409-469
    Extractor(int nf, float sf, int nl, int ini, int mn) : nfeatures(nf), nlevels(nl), iniTh(ini), minTh(mn), scaleFactor(sf) {
        scale.resize(nl); sigma2.resize(nl); invScale.resize(nl); invSigma2.resize(nl);
        scale[0] = 1.0f; sigma2[0] = 1.0f;
        for (int i = 1; i < nl; ++i) {
            scale[i] = (float)(scale[i - 1] * scaleFactor);
            sigma2[i] = scale[i] * scale[i];
        }
        for (int i = 0; i < nl; ++i) { invScale[i] = 1.0f / scale[i]; invSigma2[i] = 1.0f / sigma2[i]; }
        featPerLevel.resize(nl);
        float factor = (float)(1.0f / scaleFactor);
        float nDesired = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels));
        int sum = 0;
        for (int l = 0; l < nl - 1; ++l) {
            featPerLevel[l] = cvRoundf(nDesired);
            sum += featPerLevel[l];
            nDesired *= factor;
        }
        featPerLevel[nl - 1] = std::max(nfeatures - sum, 0);
        umax.resize(HALF_PATCH_SIZE + 1);
        int v, v0, vmax = cvFloor(HALF_PATCH_SIZE * std::sqrt(2.f) / 2 + 1);
        int vmin = cvCeil(HALF_PATCH_SIZE * std::sqrt(2.f) / 2);
        const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
        for (v = 0; v <= vmax; ++v) umax[v] = cvRound(std::sqrt(hp2 - v * v));
        for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
            while (umax[v0] == umax[v0 + 1]) ++v0;
            umax[v] = v0;
            ++v0;
        }
        pyr.resize(nl);
    }

    // ORBextractor::ComputePyramid, src/ORBextractor.cc:1170-1195.  The reflected
    // 19-px border the reference also writes is never read by operator() (SURVEY.md
    // section 8a "reach analysis"), so the planes are stored unbordered.
    void compute_pyramid(const uint8_t* img, int rows, int cols, int step) {
        for (int l = 0; l < nlevels; ++l) {
            float s = invScale[l];
            int w = cvRoundf((float)cols * s), h = cvRoundf((float)rows * s);
            pyr[l].w = w; pyr[l].h = h;
            pyr[l].d.assign((size_t)w * h, 0);
            if (l == 0) {
                for (int y = 0; y < rows; ++y) std::memcpy(&pyr[0].d[(size_t)y * w], img + (size_t)y * step, cols);
            } else {
                resize_linear(pyr[l - 1].d.data(), pyr[l - 1].w, pyr[l - 1].h, pyr[l - 1].w, pyr[l].d.data(), w, h, w);
            }
        }
    }

    // ORBextractor::DistributeOctTree, src/ORBextractor.cc:555-779
    std::vector<KeyPoint> distribute(const std::vector<KeyPoint>& in, int minX, int maxX, int minY, int maxY, int N) {
        const int nIni = (int)std::round(static_cast<float>(maxX - minX) / (maxY - minY));
        const float hX = static_cast<float>(maxX - minX) / nIni;
        std::list<Node> nodes;
        std::vector<Node*> ini(nIni);
        for (int i = 0; i < nIni; ++i) {
            Node ni;
            ni.ULx = (int)(hX * static_cast<float>(i)); ni.ULy = 0;
            ni.URx = (int)(hX * static_cast<float>(i + 1)); ni.URy = 0;
            ni.BLx = ni.ULx; ni.BLy = maxY - minY;
            ni.BRx = ni.URx; ni.BRy = maxY - minY;
            nodes.push_back(ni);
            ini[i] = &nodes.back();
        }
        for (const KeyPoint& kp : in) ini[(size_t)(kp.x / hX)]->keys.push_back(kp);
        for (auto it = nodes.begin(); it != nodes.end();) {
            if (it->keys.size() == 1) { it->noMore = true; ++it; }
            else if (it->keys.empty()) it = nodes.erase(it);
            else ++it;
        }
        bool finish = false;
        std::vector<SizeNode> sizeNode;
        auto push_child = [&](Node& c, int* nToExpand) {
            if (c.keys.size() > 0) {
                nodes.push_front(c);
                if (c.keys.size() > 1) {
                    if (nToExpand) ++*nToExpand;
                    sizeNode.push_back(std::make_pair((int)c.keys.size(), &nodes.front()));
                    nodes.front().lit = nodes.begin();
                }
            }
        };
        while (!finish) {
            int prevSize = (int)nodes.size();
            auto it = nodes.begin();
            int nToExpand = 0;
            sizeNode.clear();
            while (it != nodes.end()) {
                if (it->noMore) { ++it; continue; }
                Node n1, n2, n3, n4;
                divide_node(*it, n1, n2, n3, n4);
                push_child(n1, &nToExpand); push_child(n2, &nToExpand);
                push_child(n3, &nToExpand); push_child(n4, &nToExpand);
                it = nodes.erase(it);
            }
            if ((int)nodes.size() >= N || (int)nodes.size() == prevSize) {
                finish = true;
            } else if ((int)nodes.size() + nToExpand * 3 > N) {
                while (!finish) {
                    prevSize = (int)nodes.size();
                    std::vector<SizeNode> prev = sizeNode;
                    sizeNode.clear();
                    std::sort(prev.begin(), prev.end(), compare_nodes);
                    for (int j = (int)prev.size() - 1; j >= 0; --j) {
                        Node n1, n2, n3, n4;
                        divide_node(*prev[j].second, n1, n2, n3, n4);
                        push_child(n1, nullptr); push_child(n2, nullptr);
                        push_child(n3, nullptr); push_child(n4, nullptr);
                        nodes.erase(prev[j].second->lit);
                        if ((int)nodes.size() >= N) break;
                    }
                    if ((int)nodes.size() >= N || (int)nodes.size() == prevSize) finish = true;
                }
            }
        }
        std::vector<KeyPoint> out;
        out.reserve(nodes.size());
        for (Node& n : nodes) {
            const KeyPoint* best = &n.keys[0];
            float maxResp = best->response;
            for (size_t k = 1; k < n.keys.size(); ++k)
                if (n.keys[k].response > maxResp) { best = &n.keys[k]; maxResp = n.keys[k].response; }
            out.push_back(*best);
        }
        return out;
    }

    // IC_Angle, src/ORBextractor.cc:76-103
    float ic_angle(const Image& im, float px, float py) const {
        int m01 = 0, m10 = 0;
        const int cx = cvRoundf(px), cy = cvRoundf(py), step = im.w;
        const uint8_t* c = &im.d[(size_t)cy * step + cx];
        for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m10 += u * c[u];
        for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
            int vsum = 0, d = umax[v];
            for (int u = -d; u <= d; ++u) {
                int vp = c[u + v * step], vm = c[u - v * step];
                vsum += vp - vm;
                m10 += u * (vp + vm);
            }
            m01 += v * vsum;
        }
        return fast_atan2((float)m01, (float)m10);
    }

    // ORBextractor::ComputeKeyPointsOctTree, src/ORBextractor.cc:781-896
    void compute_keypoints() {
        cand.assign(nlevels, {}); kept.assign(nlevels, {});
        const float W = 35;
        std::vector<Corner> cell;
        for (int l = 0; l < nlevels; ++l) {
            const Image& im = pyr[l];
            const int minBX = EDGE_THRESHOLD - 3, minBY = minBX;
            const int maxBX = im.w - EDGE_THRESHOLD + 3, maxBY = im.h - EDGE_THRESHOLD + 3;
            std::vector<KeyPoint>& toDist = cand[l];
            const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
            const int nCols = (int)(width / W), nRows = (int)(height / W);
            const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
            for (int i = 0; i < nRows; ++i) {
                const float iniY = (float)(minBY + i * hCell);
                float maxY = iniY + hCell + 6;
                if (iniY >= maxBY - 3) continue;
                if (maxY > maxBY) maxY = (float)maxBY;
                for (int j = 0; j < nCols; ++j) {
                    const float iniX = (float)(minBX + j * wCell);
                    float maxX = iniX + wCell + 6;
                    if (iniX >= maxBX - 6) continue;
                    if (maxX > maxBX) maxX = (float)maxBX;
                    const int x0 = (int)iniX, y0 = (int)iniY, rw = (int)maxX - x0, rh = (int)maxY - y0;
                    const uint8_t* roi = &im.d[(size_t)y0 * im.w + x0];
                    fast9_nms(roi, rw, rh, im.w, iniTh, cell);
                    if (cell.empty()) fast9_nms(roi, rw, rh, im.w, minTh, cell);
                    for (const Corner& c : cell) {
                        KeyPoint kp;
                        kp.x = (float)c.x + j * wCell; kp.y = (float)c.y + i * hCell;
                        kp.size = 7.f; kp.angle = -1.f; kp.response = (float)c.score; kp.octave = 0; kp.class_id = -1;
                        toDist.push_back(kp);
                    }
                }
            }
            std::vector<KeyPoint>& kps = kept[l];
            kps = distribute(toDist, minBX, maxBX, minBY, maxBY, featPerLevel[l]);
            const int scaledPatch = (int)(PATCH_SIZE * scale[l]);
            for (KeyPoint& k : kps) {
                k.x += minBX; k.y += minBY; k.octave = l; k.size = (float)scaledPatch;
            }
        }
        for (int l = 0; l < nlevels; ++l)
            for (KeyPoint& k : kept[l]) k.angle = ic_angle(pyr[l], k.x, k.y);
    }

    // computeOrbDescriptor, src/ORBextractor.cc:107-146.  a,b = (float)cos/sin of a
    // float => one glibc sincosf; the reference's -O3 -march=native build contracts
    // x*b + y*a into fma(x,b,y*a) and x*a - y*b into fma(x,a,-(y*b)).
    static void orb_descriptor(const KeyPoint& kp, const Image& img, uint8_t* desc) {
        const float factorPI = (float)(3.14159265358979323846 / 180.f);
        float angle = kp.angle * factorPI;
        float a, b;
        sincosf(angle, &b, &a);
        const int cx = cvRoundf(kp.x), cy = cvRoundf(kp.y), step = img.w;
        const uint8_t* c = &img.d[(size_t)cy * step + cx];
        const signed char* p = kPattern;
        auto get = [&](int idx) -> int {
            float x = (float)p[2 * idx], y = (float)p[2 * idx + 1];
            int r = cvRoundf(fmaf(x, b, y * a));
            int q = cvRoundf(fmaf(x, a, -(y * b)));
            return c[r * step + q];
        };
        for (int i = 0; i < 32; ++i, p += 32) {
            int val = 0;
            for (int k = 0; k < 8; ++k) val |= (get(2 * k) < get(2 * k + 1)) << k;
            desc[i] = (uint8_t)val;
        }
    }

    // ORBextractor::operator(), src/ORBextractor.cc:1086-1168
    int run(const uint8_t* img, int rows, int cols, int step, int lap0, int lap1, KeyPoint* out, uint8_t* desc, int cap, int* nOut) {
        *nOut = 0;
        if (!img || rows <= 0 || cols <= 0) return -1;
        compute_pyramid(img, rows, cols, step);
        compute_keypoints();
        int n = 0;
        for (int l = 0; l < nlevels; ++l) n += (int)kept[l].size();
        *nOut = n;
        if (n > cap) return -2;
        int mono = 0, stereo = n - 1;
        Image blurred;
        for (int l = 0; l < nlevels; ++l) {
            std::vector<KeyPoint>& kps = kept[l];
            if (kps.empty()) continue;
            blurred.w = pyr[l].w; blurred.h = pyr[l].h;
            blurred.d.resize(pyr[l].d.size());
            blur7(pyr[l].d.data(), pyr[l].w, pyr[l].h, pyr[l].w, blurred.d.data(), pyr[l].w);
            const float s = scale[l];
            uint8_t d[32];
            for (const KeyPoint& k0 : kps) {
                orb_descriptor(k0, blurred, d);
                KeyPoint k = k0;
                if (l != 0) { k.x *= s; k.y *= s; }
                int at;
                if (k.x >= lap0 && k.x <= lap1) at = stereo--;
                else at = mono++;
                out[at] = k;
                std::memcpy(desc + (size_t)at * 32, d, 32);
            }
        }
        return mono;
    }
};

}  // namespace orbo

using namespace orbo;

extern "C" {

void* orbo_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh) {
    return new Extractor(nfeatures, scaleFactor, nlevels, iniTh, minTh);
}
void orbo_destroy(void* h) { delete (Extractor*)h; }

int orbo_extract(void* h, const uint8_t* img, int rows, int cols, int step, int lap0, int lap1, KeyPoint* kps, uint8_t* desc, int cap, int* n) {
    return ((Extractor*)h)->run(img, rows, cols, step, lap0, lap1, kps, desc, cap, n);
}
void orbo_tables(void* h, float* scale, float* invScale, float* sigma2, float* invSigma2, int* featPerLevel, int* umax16) {
    Extractor* e = (Extractor*)h;
    for (int i = 0; i < e->nlevels; ++i) {
        scale[i] = e->scale[i]; invScale[i] = e->invScale[i]; sigma2[i] = e->sigma2[i]; invSigma2[i] = e->invSigma2[i];
        featPerLevel[i] = e->featPerLevel[i];
    }
    for (int i = 0; i < 16; ++i) umax16[i] = e->umax[i];
}
void orbo_level_size(void* h, int level, int* w, int* hh) {
    Extractor* e = (Extractor*)h;
    *w = e->pyr[level].w; *hh = e->pyr[level].h;
}
void orbo_level_copy(void* h, int level, uint8_t* dst) {
    Extractor* e = (Extractor*)h;
    std::memcpy(dst, e->pyr[level].d.data(), e->pyr[level].d.size());
}
// debug taps: candidates fed to the quadtree (coords relative to minBorder) and kept keypoints (level coords)
int orbo_level_candidates(void* h, int level, KeyPoint* out, int cap) {
    Extractor* e = (Extractor*)h;
    int n = (int)e->cand[level].size();
    for (int i = 0; i < n && i < cap; ++i) out[i] = e->cand[level][i];
    return n;
}
int orbo_level_keypoints(void* h, int level, KeyPoint* out, int cap) {
    Extractor* e = (Extractor*)h;
    int n = (int)e->kept[level].size();
    for (int i = 0; i < n && i < cap; ++i) out[i] = e->kept[level][i];
    return n;
}

// primitives, exposed for the cv2 pinning tests
void orbo_resize_linear(const uint8_t* src, int sw, int sh, int sstep, uint8_t* dst, int dw, int dh, int dstep) {
    resize_linear(src, sw, sh, sstep, dst, dw, dh, dstep);
}
void orbo_blur7(const uint8_t* src, int w, int h, int sstep, uint8_t* dst, int dstep) { blur7(src, w, h, sstep, dst, dstep); }
int orbo_fast(const uint8_t* roi, int w, int h, int step, int T, int* xys, int cap) {
    std::vector<Corner> c;
    fast9_nms(roi, w, h, step, T, c);
    for (size_t i = 0; i < c.size() && (int)i < cap; ++i) { xys[3 * i] = c[i].x; xys[3 * i + 1] = c[i].y; xys[3 * i + 2] = c[i].score; }
    return (int)c.size();
}
float orbo_fast_atan2(float y, float x) { return fast_atan2(y, x); }
void orbo_fast_atan2_n(const float* y, const float* x, float* out, int n) {
    for (int i = 0; i < n; ++i) out[i] = fast_atan2(y[i], x[i]);
}
void orbo_sincosf_n(const float* a, float* s, float* c, int n) {
    for (int i = 0; i < n; ++i) sincosf(a[i], &s[i], &c[i]);
}
// distribute alone (for quadtree parity tests on arbitrary candidate sets)
int orbo_distribute(void* h, const KeyPoint* in, int n, int minX, int maxX, int minY, int maxY, int N, KeyPoint* out, int cap) {
    std::vector<KeyPoint> v(in, in + n);
    std::vector<KeyPoint> r = ((Extractor*)h)->distribute(v, minX, maxX, minY, maxY, N);
    for (size_t i = 0; i < r.size() && (int)i < cap; ++i) out[i] = r[i];
    return (int)r.size();
}

}  // extern "C"
