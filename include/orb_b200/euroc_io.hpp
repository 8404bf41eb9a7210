// EuRoC sequence readers of the reference's monocular-inertial example (the data formats either side of the path), host C++:
//   void LoadImages(const string& strImagePath, const string& strPathTimes, vector<string>& vstrImages, vector<double>& vTimeStamps)
//   void LoadIMU(const string& strImuPath, vector<double>& vTimeStamps, vector<cv::Point3f>& vAcc, vector<cv::Point3f>& vGyro)
//        (reference Examples/Monocular-Inertial/mono_inertial_euroc.cc:252-310, declared :34-37)
//   the per-frame hand-over of IMU samples of its main loop (:170-183): every sample with t <= tframe that was not handed over yet
// plus the flattening of those samples into what imu_preintegrate_batch (include/orb_b200.h) takes for a batch of streams.
// Same parsing rules as the reference: one time stamp (ns) per non-empty line of the times file, image name = path + "/" + line + ".png",
// seconds = ns / 1e9; IMU csv lines "t, wx, wy, wz, ax, ay, az", lines starting with '#' skipped.
#ifndef ORB_B200_EUROC_IO_HPP
#define ORB_B200_EUROC_IO_HPP
#include <cstddef>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

namespace ORB_SLAM3 {
namespace euroc {

struct Point3f { float x, y, z; };   // cv::Point3f layout

inline void LoadImages(const std::string& strImagePath, const std::string& strPathTimes, std::vector<std::string>& vstrImages, std::vector<double>& vTimeStamps) {
    std::ifstream in(strPathTimes.c_str());
    std::string line;
    while (std::getline(in, line)) {
        if (line.empty()) continue;
        vstrImages.push_back(strImagePath + "/" + line + ".png");
        std::istringstream ss(line);
        double ns = 0;                       // (the reference leaves `t` uninitialised when the line is not a number)
        ss >> ns;
        vTimeStamps.push_back(ns / 1e9);
    }
}

inline void LoadIMU(const std::string& strImuPath, std::vector<double>& vTimeStamps, std::vector<Point3f>& vAcc, std::vector<Point3f>& vGyro) {
    std::ifstream in(strImuPath.c_str());
    std::string line;
    while (std::getline(in, line)) {
        if (line.empty() || line[0] == '#') continue;
        double field[7] = {0, 0, 0, 0, 0, 0, 0};
        std::size_t from = 0;
        for (int k = 0; k < 7; ++k) {        // six commas separate the seven fields; the last one runs to the end of the line
            const std::size_t comma = k < 6 ? line.find(',', from) : std::string::npos;
            field[k] = std::stod(line.substr(from, comma == std::string::npos ? std::string::npos : comma - from));
            if (comma == std::string::npos) break;
            from = comma + 1;
        }
        vTimeStamps.push_back(field[0] / 1e9);
        const Point3f a = {(float)field[4], (float)field[5], (float)field[6]}, g = {(float)field[1], (float)field[2], (float)field[3]};
        vAcc.push_back(a);
        vGyro.push_back(g);
    }
}

// The samples handed to TrackMonocular with frame ni > 0 (:170-183): [first_imu, end) with t <= tframe; first_imu advances.  (The reference does not
// test the end of the IMU vector; this does.)
inline void ImuSince(const std::vector<double>& vTimestampsImu, double tframe, std::size_t& first_imu, std::size_t& begin, std::size_t& end) {
    begin = first_imu;
    while (first_imu < vTimestampsImu.size() && vTimestampsImu[first_imu] <= tframe) ++first_imu;
    end = first_imu;
}

// The selection at the head of Tracking::PreintegrateIMU (src/Tracking.cc:1646-1672) on the queue Tracking::GrabImuData fills: the queue holds the samples
// [front, queued) of the stream (queued = one past the last sample handed over so far); samples older than tPrev - imuPer are dropped, those before tCur - imuPer are
// taken and consumed, the first one at or after tCur - imuPer is taken but STAYS at the front of the queue (it opens the next frame's interval).
// sel receives the indices of mvImuFromLastFrame; imuPer is Tracking's mImuPer (0.001, :609).
inline void SelectImuFromQueue(const std::vector<double>& t, std::size_t queued, std::size_t& front, double tPrev, double tCur, double imuPer, std::vector<std::size_t>& sel) {
    sel.clear();
    while (front < queued) {
        if (t[front] < tPrev - imuPer) { ++front; continue; }
        if (t[front] < tCur - imuPer) { sel.push_back(front); ++front; continue; }
        sel.push_back(front);
        break;
    }
}

// Flatten the IMU samples between consecutive frames of `count` streams into imu_preintegrate_batch's arrays: stream s contributes the samples
// [begin[s], end[s]) of its vectors (= mvImuFromLastFrame, a contiguous run: see SelectImuFromQueue); the integration steps follow Tracking::PreintegrateIMU (src/Tracking.cc:1680-1729): n-1 steps between
// consecutive samples, the first and last interpolated to the frame times tPrev[s] / tCur[s].
// acc / gyr [count][maxMeas][3], dt [count][maxMeas], nMeas [count]; returns false when a stream has more than maxMeas steps.
inline bool FlattenForPreintegration(int count, const std::vector<double>* const* tImu, const std::vector<Point3f>* const* vAcc, const std::vector<Point3f>* const* vGyro,
                                     const std::size_t* begin, const std::size_t* end, const double* tPrev, const double* tCur, int maxMeas,
                                     float* acc, float* gyr, float* dt, int* nMeas) {
    for (int s = 0; s < count; ++s) {
        const std::vector<double>& t = *tImu[s];
        const std::vector<Point3f>&a = *vAcc[s], &w = *vGyro[s];
        const int n = (int)(end[s] - begin[s]);
        nMeas[s] = n > 1 ? n - 1 : 0;
        if (nMeas[s] > maxMeas) return false;
        for (int i = 0; i < n - 1; ++i) {
            const std::size_t k = begin[s] + (std::size_t)i;
            float tstep;
            Point3f am, wm;
            if (i == 0 && i < n - 2) {
                const float tab = (float)(t[k + 1] - t[k]), tini = (float)(t[k] - tPrev[s]);
                am.x = (a[k].x + a[k + 1].x - (a[k + 1].x - a[k].x) * (tini / tab)) * 0.5f; am.y = (a[k].y + a[k + 1].y - (a[k + 1].y - a[k].y) * (tini / tab)) * 0.5f;
                am.z = (a[k].z + a[k + 1].z - (a[k + 1].z - a[k].z) * (tini / tab)) * 0.5f;
                wm.x = (w[k].x + w[k + 1].x - (w[k + 1].x - w[k].x) * (tini / tab)) * 0.5f; wm.y = (w[k].y + w[k + 1].y - (w[k + 1].y - w[k].y) * (tini / tab)) * 0.5f;
                wm.z = (w[k].z + w[k + 1].z - (w[k + 1].z - w[k].z) * (tini / tab)) * 0.5f;
                tstep = (float)(t[k + 1] - tPrev[s]);
            } else if (i < n - 2) {
                am.x = (a[k].x + a[k + 1].x) * 0.5f; am.y = (a[k].y + a[k + 1].y) * 0.5f; am.z = (a[k].z + a[k + 1].z) * 0.5f;
                wm.x = (w[k].x + w[k + 1].x) * 0.5f; wm.y = (w[k].y + w[k + 1].y) * 0.5f; wm.z = (w[k].z + w[k + 1].z) * 0.5f;
                tstep = (float)(t[k + 1] - t[k]);
            } else if (i > 0 && i == n - 2) {
                const float tab = (float)(t[k + 1] - t[k]), tend = (float)(t[k + 1] - tCur[s]);
                am.x = (a[k].x + a[k + 1].x - (a[k + 1].x - a[k].x) * (tend / tab)) * 0.5f; am.y = (a[k].y + a[k + 1].y - (a[k + 1].y - a[k].y) * (tend / tab)) * 0.5f;
                am.z = (a[k].z + a[k + 1].z - (a[k + 1].z - a[k].z) * (tend / tab)) * 0.5f;
                wm.x = (w[k].x + w[k + 1].x - (w[k + 1].x - w[k].x) * (tend / tab)) * 0.5f; wm.y = (w[k].y + w[k + 1].y - (w[k + 1].y - w[k].y) * (tend / tab)) * 0.5f;
                wm.z = (w[k].z + w[k + 1].z - (w[k + 1].z - w[k].z) * (tend / tab)) * 0.5f;
                tstep = (float)(tCur[s] - t[k]);
            } else {                            // i == 0 && i == n - 2: a single step spanning the whole interval
                am = a[k]; wm = w[k];
                tstep = (float)(tCur[s] - tPrev[s]);
            }
            float* A = acc + ((std::size_t)s * maxMeas + i) * 3; float* G = gyr + ((std::size_t)s * maxMeas + i) * 3;
            A[0] = am.x; A[1] = am.y; A[2] = am.z; G[0] = wm.x; G[1] = wm.y; G[2] = wm.z;
            dt[(std::size_t)s * maxMeas + i] = tstep;
        }
    }
    return true;
}

}  // namespace euroc
}  // namespace ORB_SLAM3
#endif
