// TEST INFRASTRUCTURE ONLY -- the five OpenCV image primitives behind minicv.hpp, forwarded to the primitives of
// liborb_oracle.so that tests/test_oracle_cpu.py pins bit-for-bit against the cv2 4.13 wheel (call sites in the
// reference: cv::FAST src/ORBextractor.cc:826,845; cv::resize :1183; cv::copyMakeBorder :1185,1190;
// cv::GaussianBlur :1133; cv::fastAtan2 :102).
#include "minicv.hpp"

extern "C" {
void orbo_resize_linear(const uint8_t* src, int sw, int sh, int sstep, uint8_t* dst, int dw, int dh, int dstep);
void orbo_blur7(const uint8_t* src, int w, int h, int sstep, uint8_t* dst, int dstep);
int orbo_fast(const uint8_t* roi, int w, int h, int step, int T, int* xys, int cap);
float orbo_fast_atan2(float y, float x);
}

namespace cv {

void FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression) {
    assert(nonmaxSuppression);
    Mat img = image.getMat();
    keypoints.clear();
    if (img.empty()) return;
    std::vector<int> xys((size_t)3 * img.rows * img.cols + 3);
    int n = orbo_fast(img.data, img.cols, img.rows, (int)img.step, threshold, xys.data(), img.rows * img.cols);
    keypoints.reserve(n);
    // OpenCV features2d/fast.cpp: KeyPoint((float)j, (float)(i-1), 7.f, -1, (float)score)
    for (int i = 0; i < n; ++i) keypoints.push_back(KeyPoint((float)xys[3 * i], (float)xys[3 * i + 1], 7.f, -1, (float)xys[3 * i + 2]));
}

void resize(InputArray src, OutputArray dst, Size dsize, double fx, double fy, int interpolation) {
    assert(interpolation == INTER_LINEAR && fx == 0 && fy == 0 && dsize.width > 0 && dsize.height > 0);
    Mat s = src.getMat();
    dst.create(dsize.height, dsize.width, CV_8UC1);
    Mat d = dst.getMat();
    orbo_resize_linear(s.data, s.cols, s.rows, (int)s.step, d.data, d.cols, d.rows, (int)d.step);
}

static inline int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}

void copyMakeBorder(InputArray src, OutputArray dst, int top, int bottom, int left, int right, int borderType) {
    assert((borderType & ~BORDER_ISOLATED) == BORDER_REFLECT_101);
    Mat s = src.getMat();
    dst.create(s.rows + top + bottom, s.cols + left + right, CV_8UC1);
    Mat d = dst.getMat();
    // src may be the interior ROI of dst (src/ORBextractor.cc:1178,1185): interior writes are then identities and
    // the border only reads interior pixels, so any order is safe.
    for (int r = 0; r < d.rows; ++r) {
        const uchar* srow = s.data + (size_t)reflect101(r - top, s.rows) * s.step;
        uchar* drow = d.data + (size_t)r * d.step;
        for (int c = 0; c < d.cols; ++c) drow[c] = srow[reflect101(c - left, s.cols)];
    }
}

void GaussianBlur(InputArray src, OutputArray dst, Size ksize, double sigmaX, double sigmaY, int borderType) {
    assert(ksize.width == 7 && ksize.height == 7 && sigmaX == 2 && sigmaY == 2 && borderType == BORDER_REFLECT_101);
    Mat s = src.getMat();
    Mat tmp(s.rows, s.cols, CV_8UC1);
    orbo_blur7(s.data, s.cols, s.rows, (int)s.step, tmp.data, (int)tmp.step);
    tmp.copyTo(dst);
}

float fastAtan2(float y, float x) { return orbo_fast_atan2(y, x); }

void KeyPointsFilter::retainBest(std::vector<KeyPoint>& keypoints, int npoints) {
    // dead in the reference (only ComputeKeyPointsOld uses it); kept so that the file links
    if (npoints < 0 || (int)keypoints.size() <= npoints) return;
    std::stable_sort(keypoints.begin(), keypoints.end(), [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
    keypoints.resize(npoints);
}

}  // namespace cv
