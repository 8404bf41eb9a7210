// C entry points over the host mirror include/orb_b200/euroc_io.hpp, for tests/test_euroc_io_cpu.py.
#include <cstring>

#include "orb_b200/euroc_io.hpp"

using namespace ORB_SLAM3::euroc;

extern "C" {
int mirror_euroc_load_images(const char* imagePath, const char* timesPath, int cap, int stride, char* names, double* stamps) {
    std::vector<std::string> v; std::vector<double> t;
    LoadImages(imagePath, timesPath, v, t);
    for (int i = 0; i < (int)v.size() && i < cap; ++i) { strncpy(names + (size_t)i * stride, v[i].c_str(), stride - 1); names[(size_t)i * stride + stride - 1] = 0; stamps[i] = t[i]; }
    return (int)v.size();
}
int mirror_euroc_load_imu(const char* imuPath, int cap, double* stamps, float* acc3, float* gyr3) {
    std::vector<double> t; std::vector<Point3f> a, g;
    LoadIMU(imuPath, t, a, g);
    for (int i = 0; i < (int)t.size() && i < cap; ++i) { stamps[i] = t[i]; memcpy(acc3 + 3 * i, &a[i], 12); memcpy(gyr3 + 3 * i, &g[i], 12); }
    return (int)t.size();
}
// the main loop's hand-over for consecutive frame times (spanBegin / spanEnd) + the flattening of those spans taken as they are (mirror_tracking_flatten below applies
// Tracking's queue selection first, which is what PreintegrateIMU integrates)
int mirror_euroc_flatten(const char* imuPath, int nFrames, const double* tFrames, int maxMeas, float* acc, float* gyr, float* dt, int* nMeas, int* spanBegin, int* spanEnd) {
    std::vector<double> t; std::vector<Point3f> a, g;
    LoadIMU(imuPath, t, a, g);
    std::size_t first = 0;
    std::vector<std::size_t> b(nFrames, 0), e(nFrames, 0);
    for (int f = 1; f < nFrames; ++f) { ImuSince(t, tFrames[f], first, b[f], e[f]); spanBegin[f] = (int)b[f]; spanEnd[f] = (int)e[f]; }
    const int count = nFrames - 1;
    std::vector<const std::vector<double>*> pt(count, &t);
    std::vector<const std::vector<Point3f>*> pa(count, &a), pg(count, &g);
    return FlattenForPreintegration(count, pt.data(), pa.data(), pg.data(), b.data() + 1, e.data() + 1, tFrames, tFrames + 1, maxMeas, acc, gyr, dt, nMeas) ? 1 : 0;
}
// Tracking::PreintegrateIMU's front end with the mirror: per frame f >= 1 the queue selection (SelectImuFromQueue) and the flattened integration steps
int mirror_tracking_flatten(int nImu, const double* tImu, const float* acc, const float* gyr, int nFrames, const double* tFrames, const int* queuedUpTo, int maxMeas,
                            int* selFirst, int* selCount, float* A, float* G, float* DT, int* NM) {
    std::vector<double> t(tImu, tImu + nImu);
    std::vector<Point3f> a(nImu), g(nImu);
    for (int i = 0; i < nImu; ++i) { memcpy(&a[i], acc + 3 * i, 12); memcpy(&g[i], gyr + 3 * i, 12); }
    std::size_t front = 0;
    std::vector<std::size_t> sel;
    for (int f = 0; f < nFrames; ++f) {
        selFirst[f] = -1; selCount[f] = 0; NM[f] = 0;
        if (f == 0) continue;
        SelectImuFromQueue(t, (std::size_t)queuedUpTo[f], front, tFrames[f - 1], tFrames[f], 0.001, sel);
        selCount[f] = (int)sel.size();
        if (sel.empty()) continue;
        selFirst[f] = (int)sel[0];
        const std::vector<double>* pt = &t; const std::vector<Point3f>*pa = &a, *pg = &g;
        const std::size_t b = sel[0], e = sel.back() + 1;
        if (!FlattenForPreintegration(1, &pt, &pa, &pg, &b, &e, tFrames + f - 1, tFrames + f, maxMeas, A + (std::size_t)f * maxMeas * 3, G + (std::size_t)f * maxMeas * 3,
                                      DT + (std::size_t)f * maxMeas, NM + f)) return 0;
    }
    return 1;
}
}
