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
// the main loop's hand-over for consecutive frame times + the flattening for imu_preintegrate_batch of frame pairs (f-1, f), one "stream" per pair
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
}
