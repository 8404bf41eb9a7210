// Compiles against the C++ host mirror (include/orb_b200/orb_slam3.hpp) like a reference caller would and drives every surface:
//   extract   Frame::ExtractORB (src/Frame.cc:418-425)                      -> ORBextractor::operator()
//   track     Tracking::TrackWithMotionModel (src/Tracking.cc:2876-2897)    -> ORBmatcher(0.9,true).SearchByProjection(Current, Last, th, mono)
//             + Optimizer::PoseOptimization (src/Tracking.cc:2919)
//   lba       LocalMapping::Run (src/LocalMapping.cc:154)                    -> Optimizer::LocalBundleAdjustment(..., &mbAbortBA, ...)
//   liba      LocalMapping::Run (src/LocalMapping.cc:129-151)                -> Optimizer::LocalInertialBA(..., bLarge, bRecInit)
// Inputs are raw arrays written by tests/test_cpp_shim.py; results are printed as text for the test to compare with the oracle.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "orb_b200/orb_slam3.hpp"

template <class T>
static std::vector<T> load(const std::string& dir, const char* name) {
    std::vector<T> v;
    FILE* f = fopen((dir + "/" + name).c_str(), "rb");
    if (!f) { fprintf(stderr, "missing %s\n", name); exit(3); }
    fseek(f, 0, SEEK_END); const long n = ftell(f); fseek(f, 0, SEEK_SET);
    v.resize((size_t)n / sizeof(T));
    if (n && fread(v.data(), 1, (size_t)n, f) != (size_t)n) exit(3);
    fclose(f);
    return v;
}
static unsigned long long checksum(const void* p, size_t n) {
    unsigned long long s = 0;
    for (size_t i = 0; i < n; ++i) s = s * 1315423911ull + ((const uint8_t*)p)[i];
    return s;
}

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    const std::string dir = argv[1];
    const int rows = atoi(argv[2]), cols = atoi(argv[3]);
    using namespace ORB_SLAM3;
    // ---- extract ----
    std::vector<uint8_t> img = load<uint8_t>(dir, "img.raw");
    ORBextractor extractor(1000, 1.2f, 8, 20, 7, cols, rows);
    Image im; im.data = img.data(); im.rows = rows; im.cols = cols; im.step = cols;
    FrameView Cur;
    std::vector<int> vLapping = {0, 1000};
    const int monoLeft = extractor(im, Image(), Cur.mvKeysUn, Cur.mDescriptors, vLapping);
    printf("extract %d %zu %llu\n", monoLeft, Cur.mvKeysUn.size(), checksum(Cur.mDescriptors.data.data(), Cur.mDescriptors.data.size()));
    Image empty;
    std::vector<KeyPoint> k2; Descriptors d2;
    if (extractor(empty, empty, k2, d2, vLapping) != -1) return 4;
    // ---- track with the motion model: two ORBmatcher temporaries (like the th and 2*th attempts), then PoseOptimization ----
    Cur.mnMinX = 0; Cur.mnMinY = 0; Cur.mnMaxX = (float)cols; Cur.mnMaxY = (float)rows;
    Cur.mvScaleFactors = extractor.GetScaleFactors();
    std::vector<uint8_t> valid = load<uint8_t>(dir, "last_valid.raw"), hasObs = load<uint8_t>(dir, "last_hasobs.raw"), mpDesc = load<uint8_t>(dir, "last_desc.raw");
    std::vector<float> xyz = load<float>(dir, "last_xyz.raw"), angle = load<float>(dir, "last_angle.raw"), Tcw = load<float>(dir, "tcw.raw"), cam = load<float>(dir, "cam.raw");
    std::vector<int32_t> octave = load<int32_t>(dir, "last_octave.raw");
    OrbmLastFrame Last; Last.M = (int)valid.size(); Last.valid = valid.data(); Last.xyz = xyz.data(); Last.octave = octave.data(); Last.angle = angle.data();
    Last.hasObs = hasObs.data(); Last.descriptors = mpDesc.data();
    int nmatches = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        ORBmatcher matcher(0.9f, true);                                  // stack temporary per use, as in the reference
        std::fill(Cur.mvpMapPoints.begin(), Cur.mvpMapPoints.end(), -1);  // fill(mvpMapPoints, NULL), src/Tracking.cc:2876
        std::fill(Cur.mvbClaimed.begin(), Cur.mvbClaimed.end(), 0);
        nmatches = matcher.SearchByProjection(Cur, Last, Tcw.data(), cam.data(), attempt == 0 ? 15.f : 30.f, true);
        printf("match%d %d %llu\n", attempt, nmatches, checksum(Cur.mvpMapPoints.data(), 4 * Cur.mvpMapPoints.size()));
    }
    {
        std::vector<double> Xw, obs; std::vector<float> isg;
        for (size_t i = 0; i < Cur.mvpMapPoints.size(); ++i) {
            const int m = Cur.mvpMapPoints[i];
            if (m < 0) continue;
            for (int c = 0; c < 3; ++c) Xw.push_back((double)xyz[3 * (size_t)m + c]);
            obs.push_back((double)Cur.mvKeysUn[i].x); obs.push_back((double)Cur.mvKeysUn[i].y);
            isg.push_back(extractor.GetInverseScaleSigmaSquares()[Cur.mvKeysUn[i].octave]);
        }
        const int N = (int)isg.size();
        double pose7[7], poseOut[7];
        for (int i = 0; i < 7; ++i) pose7[i] = (double)Tcw[i];
        std::vector<uint8_t> outlier((size_t)(N > 0 ? N : 1));
        const int inl = Optimizer::PoseOptimization(N, pose7, cam.data(), Xw.data(), obs.data(), isg.data(), poseOut, outlier.data());
        printf("poseopt %d %d %.9f %.9f %.9f %.9f %.9f %.9f %.9f\n", N, inl, poseOut[0], poseOut[1], poseOut[2], poseOut[3], poseOut[4], poseOut[5], poseOut[6]);
    }
    // ---- LocalBundleAdjustment, twice (the per-thread arena is created once), then with the abort flag already set ----
    {
        std::vector<double> poses = load<double>(dir, "lba_poses.raw"), points = load<double>(dir, "lba_points.raw"), lobs = load<double>(dir, "lba_obs.raw");
        std::vector<uint8_t> fixed = load<uint8_t>(dir, "lba_fixed.raw");
        std::vector<float> lcam = load<float>(dir, "lba_cam.raw"), is2 = load<float>(dir, "lba_is2.raw");
        std::vector<int32_t> ep = load<int32_t>(dir, "lba_ep.raw"), ek = load<int32_t>(dir, "lba_ek.raw");
        bool mbAbortBA = false;
        LbaProblem g; memset(&g, 0, sizeof(g));
        g.nPoses = (int)fixed.size(); g.poses = poses.data(); g.poseFixed = fixed.data(); g.cam = lcam.data();
        g.nPoints = (int)points.size() / 3; g.points = points.data();
        g.nEdges = (int)ep.size(); g.edgePoint = ep.data(); g.edgePose = ek.data(); g.obs = lobs.data(); g.invSigma2 = is2.data();
        g.huberDelta = (double)(float)std::sqrt(5.991); g.iterations = 10; g.userLambdaInit = 0;
        g.stopFlag = reinterpret_cast<const volatile uint8_t*>(&mbAbortBA);   // bool* pbStopFlag handed over as it is
        for (int rep = 0; rep < 3; ++rep) {
            std::vector<double> oposes(poses.size()), opoints(points.size()), chi2(ep.size());
            std::vector<uint8_t> dpos(ep.size());
            LbaResult r; memset(&r, 0, sizeof(r));
            r.poses = oposes.data(); r.points = opoints.data(); r.edgeChi2 = chi2.data(); r.edgeDepthPositive = dpos.data();
            r.iterations = -1;
            mbAbortBA = rep == 2;
            Optimizer::LocalBundleAdjustment(g, r);
            double sp = 0; for (double v : oposes) sp += v;
            double sx = 0; for (double v : opoints) sx += v;
            printf("lba%d %d %d %.12f %.9f %.9f\n", rep, r.iterations, r.trials, r.chi2, sp, sx);
        }
    }
    // ---- LocalInertialBA (LocalMapping::Run once the IMU is initialised, src/LocalMapping.cc:129-151): plain, then with bRecInit ----
    {
        std::vector<double> st = load<double>(dir, "liba_state.raw"), tc = load<double>(dir, "liba_tcw.raw"), ex = load<double>(dir, "liba_extr.raw"),
                            pts = load<double>(dir, "liba_points.raw"), lobs = load<double>(dir, "liba_obs.raw");
        std::vector<float> lcam = load<float>(dir, "liba_cam.raw"), pre = load<float>(dir, "liba_preint.raw"), td = load<float>(dir, "liba_td.raw"), is2 = load<float>(dir, "liba_is2.raw");
        std::vector<int32_t> k1 = load<int32_t>(dir, "liba_k1.raw"), k2 = load<int32_t>(dir, "liba_k2.raw"), ep = load<int32_t>(dir, "liba_ep.raw"), ek = load<int32_t>(dir, "liba_ek.raw");
        LocalInertialBAProblem g; memset(&g, 0, sizeof(g));
        g.nKF = (int)st.size() / 21; g.nOpt = (int)k1.size(); g.kfState21 = st.data(); g.kfTcw12 = tc.data(); g.cam4 = lcam.data(); g.extrinsics24 = ex.data();
        g.nInertial = (int)k1.size(); g.ieKf1 = k1.data(); g.ieKf2 = k2.data(); g.preint = pre.data();
        g.nPoints = (int)pts.size() / 3; g.points3 = pts.data(); g.trackDepth = td.data();
        g.nEdges = (int)ep.size(); g.edgePoint = ep.data(); g.edgeKf = ek.data(); g.obs2 = lobs.data(); g.invSigma2 = is2.data();
        for (int rec = 0; rec < 2; ++rec) {
            std::vector<double> ost(st.size()), otc(tc.size()), opts(pts.size()), chi2(ep.size());
            std::vector<uint8_t> erase(ep.size());
            double stats[8];
            LocalInertialBAResult r; memset(&r, 0, sizeof(r));
            r.kfState21 = ost.data(); r.kfTcw12 = otc.data(); r.points3 = opts.data(); r.erase = erase.data(); r.edgeChi2 = chi2.data(); r.stats8 = stats;
            const bool ok = Optimizer::LocalInertialBA(g, r, /*bLarge=*/false, /*bRecInit=*/rec == 1);
            double ss = 0; for (int k = 0; k < g.nOpt * 21; ++k) ss += ost[k];
            int ne = 0; for (uint8_t v : erase) ne += v;
            printf("liba%d %d %d %d %d %.9f %.6f\n", rec, ok ? 1 : 0, (int)stats[5], (int)stats[4], ne, ss, stats[1]);
        }
    }
    return 0;
}
