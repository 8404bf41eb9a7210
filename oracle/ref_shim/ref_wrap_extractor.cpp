// TEST INFRASTRUCTURE ONLY -- C entry points over the REFERENCE's own ORBextractor, compiled from
// /root/reference/src/ORBextractor.cc where it lies (the Makefile passes the path as REF_ORBEXTRACTOR_CC; the
// source is #included so that its file-static helpers IC_Angle / computeOrbDescriptor are callable too).
// Output: oracle/_ref/libref_orb.so (git-ignored).  Nothing of the reference is copied into the repo.
#include REF_ORBEXTRACTOR_CC

namespace {
struct Probe : ORB_SLAM3::ORBextractor {   // opens the protected members the stage tests look at
    using ORB_SLAM3::ORBextractor::ORBextractor;
    using ORB_SLAM3::ORBextractor::DistributeOctTree;
    using ORB_SLAM3::ORBextractor::mnFeaturesPerLevel;
    using ORB_SLAM3::ORBextractor::umax;
    using ORB_SLAM3::ORBextractor::mvScaleFactor;
    using ORB_SLAM3::ORBextractor::mvInvScaleFactor;
    using ORB_SLAM3::ORBextractor::mvLevelSigma2;
    using ORB_SLAM3::ORBextractor::mvInvLevelSigma2;
    using ORB_SLAM3::ORBextractor::pattern;
};
}

extern "C" {

void* ref_orbx_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh) {
    return new Probe(nfeatures, scaleFactor, nlevels, iniTh, minTh);
}
void ref_orbx_destroy(void* h) { delete (Probe*)h; }

// int ORBextractor::operator()(image, mask, keypoints, descriptors, vLappingArea); returns its return value
int ref_orbx_extract(void* h, const uint8_t* img, int rows, int cols, int step, int lap0, int lap1, cv::KeyPoint* kps, uint8_t* desc,
                     int cap, int* n) {
    Probe* e = (Probe*)h;
    cv::Mat image = (rows > 0 && cols > 0) ? cv::Mat(rows, cols, CV_8UC1, (void*)img, (size_t)step) : cv::Mat();
    std::vector<cv::KeyPoint> keys;
    cv::Mat descriptors;
    std::vector<int> lap = {lap0, lap1};
    int mono = (*e)(image, cv::Mat(), keys, descriptors, lap);
    *n = (int)keys.size();
    for (int i = 0; i < (int)keys.size() && i < cap; ++i) {
        kps[i] = keys[i];
        std::memcpy(desc + (size_t)i * 32, descriptors.ptr<uchar>(i), 32);
    }
    return mono;
}

void ref_orbx_tables(void* h, float* scale, float* invScale, float* sigma2, float* invSigma2, int* featPerLevel, int* umax16) {
    Probe* e = (Probe*)h;
    for (int i = 0; i < e->GetLevels(); ++i) {
        scale[i] = e->mvScaleFactor[i]; invScale[i] = e->mvInvScaleFactor[i];
        sigma2[i] = e->mvLevelSigma2[i]; invSigma2[i] = e->mvInvLevelSigma2[i];
        featPerLevel[i] = e->mnFeaturesPerLevel[i];
    }
    for (int i = 0; i < 16; ++i) umax16[i] = e->umax[i];
}

// mvImagePyramid[level] of the last call (public member, include/ORBextractor.h:84); border = 0: the ROI, border = 19: the
// reflected frame ComputePyramid writes around it (src/ORBextractor.cc:1185-1191), read by Frame::ComputeStereoMatches.
void ref_orbx_level_size(void* h, int level, int* w, int* hh) {
    Probe* e = (Probe*)h;
    *w = e->mvImagePyramid[level].cols; *hh = e->mvImagePyramid[level].rows;
}
void ref_orbx_level_copy(void* h, int level, int border, uint8_t* dst) {
    Probe* e = (Probe*)h;
    const cv::Mat& m = e->mvImagePyramid[level];
    const int W = m.cols + 2 * border;
    for (int r = -border; r < m.rows + border; ++r)
        std::memcpy(dst + (size_t)(r + border) * W, m.data + (ptrdiff_t)r * (ptrdiff_t)(size_t)m.step - border, (size_t)W);
}

int ref_orbx_distribute(void* h, const cv::KeyPoint* in, int n, int minX, int maxX, int minY, int maxY, int N, cv::KeyPoint* out, int cap) {
    std::vector<cv::KeyPoint> v(in, in + n);
    std::vector<cv::KeyPoint> r = ((Probe*)h)->DistributeOctTree(v, minX, maxX, minY, maxY, N, 0);
    for (size_t i = 0; i < r.size() && (int)i < cap; ++i) out[i] = r[i];
    return (int)r.size();
}

// file-static helpers of the reference (src/ORBextractor.cc:76-146)
float ref_ic_angle(void* h, const uint8_t* img, int rows, int cols, int step, float x, float y) {
    cv::Mat image(rows, cols, CV_8UC1, (void*)img, (size_t)step);
    return ORB_SLAM3::IC_Angle(image, cv::Point2f(x, y), ((Probe*)h)->umax);
}
void ref_orb_descriptor(void* h, const uint8_t* img, int rows, int cols, int step, float x, float y, float angle, uint8_t* desc32) {
    cv::Mat image(rows, cols, CV_8UC1, (void*)img, (size_t)step);
    cv::KeyPoint kp(x, y, 31.f, angle);
    ORB_SLAM3::computeOrbDescriptor(kp, image, &((Probe*)h)->pattern[0], desc32);
}

}  // extern "C"
