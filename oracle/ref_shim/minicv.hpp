// TEST INFRASTRUCTURE ONLY -- a minimal stand-in for the OpenCV *types* the reference's hot-path sources use,
// so that /root/reference/src/ORBextractor.cc (and line ranges of src/ORBmatcher.cc) compile VERBATIM into
// oracle/_ref/ without OpenCV C++ (absent from this image).  Nothing here restates reference code: the
// containers (Mat, KeyPoint, Point_, Size, Rect, Input/OutputArray) follow OpenCV's documented layout and value
// semantics, and the five image primitives (resize, FAST, GaussianBlur, copyMakeBorder, fastAtan2) forward to
// the cv2-pinned primitives of liborb_oracle.so (tests/test_oracle_cpu.py, tests/golden/primitives.npz).
// Restrictions (asserted): CV_8UC1 only; resize INTER_LINEAR with an explicit dsize; GaussianBlur 7x7 sigma 2
// BORDER_REFLECT_101; copyMakeBorder BORDER_REFLECT_101 (the ISOLATED flag is implied: sources passed by the
// wrapper are never ROIs of larger images whose surroundings OpenCV would read).
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <memory>
#include <sstream>   // the real opencv2/core pulls these in; DBoW2's TemplatedVocabulary.h relies on it
#include <string>
#include <vector>

typedef unsigned char uchar;

#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0

// cvRound: round-half-to-even (SSE cvtss2si / cvtsd2si), OpenCV core/fast_math.hpp
static inline int cvRound(double v) { return (int)std::nearbyint(v); }
static inline int cvRound(float v) { return (int)std::nearbyintf(v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
static inline int cvFloor(float v) { int i = (int)v; return i - (i > v); }
static inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }
static inline int cvCeil(float v) { int i = (int)v; return i + (i < v); }

namespace cv {

enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };
enum { BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };

template <typename T>
struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T _x, T _y) : x(_x), y(_y) {}
    template <typename U>
    Point_& operator*=(U s) { x = (T)(x * s); y = (T)(y * s); return *this; }
};
typedef Point_<int> Point2i;
typedef Point2i Point;
typedef Point_<float> Point2f;
template <typename T>
struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T _x, T _y, T _z) : x(_x), y(_y), z(_z) {}
};
typedef Point3_<float> Point3f;

template <typename T>
struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size;

struct Rect {
    int x, y, width, height;
    Rect() : x(0), y(0), width(0), height(0) {}
    Rect(int _x, int _y, int w, int h) : x(_x), y(_y), width(w), height(h) {}
};

// 28 bytes, OpenCV core/types.hpp
class KeyPoint {
public:
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(Point2f _pt, float _size, float _angle = -1, float _response = 0, int _octave = 0, int _class_id = -1)
        : pt(_pt), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
    KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0, int _class_id = -1)
        : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

struct MatStep {
    size_t v;
    MatStep() : v(0) {}
    operator size_t() const { return v; }
    MatStep& operator=(size_t s) { v = s; return *this; }
};

class _InputArray;
class _OutputArray;

// Reference-counted 8-bit single-channel matrix header with ROI views, OpenCV's cv::Mat value semantics.
class Mat {
public:
    int rows, cols;
    uchar* data;
    MatStep step;
    std::shared_ptr<uchar> buf;

    Mat() : rows(0), cols(0), data(nullptr) {}
    Mat(int r, int c, int type) : rows(0), cols(0), data(nullptr) { create(r, c, type); }
    Mat(Size s, int type) : rows(0), cols(0), data(nullptr) { create(s.height, s.width, type); }
    Mat(int r, int c, int type, void* ext, size_t st = 0) : rows(r), cols(c), data((uchar*)ext) {
        assert(type == CV_8UC1);
        step = st ? st : (size_t)c;
    }
    void create(int r, int c, int type) {
        assert(type == CV_8UC1);   // CV_32F (FORB::toMat32F) parses but is never executed here
        if (data && r == rows && c == cols) return;
        rows = r; cols = c; step = (size_t)c;
        buf = std::shared_ptr<uchar>(new uchar[(size_t)std::max(r, 0) * std::max(c, 0) + 1], std::default_delete<uchar[]>());
        data = buf.get();
    }
    void release() { rows = cols = 0; data = nullptr; buf.reset(); step = 0; }
    static Mat zeros(int r, int c, int type) {
        Mat m(r, c, type);
        if (m.data) std::memset(m.data, 0, (size_t)r * c);
        return m;
    }
    int type() const { return CV_8UC1; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    size_t step1() const { return step; }
    Size size() const { return Size(cols, rows); }
    template <typename T> T& at(int r, int c) { return *(T*)(data + (size_t)r * step + c * sizeof(T)); }
    template <typename T> const T& at(int r, int c) const { return *(const T*)(data + (size_t)r * step + c * sizeof(T)); }
    uchar* ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar* ptr(int r = 0) const { return data + (size_t)r * step; }
    template <typename T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step); }
    template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step); }
    Mat operator()(const Rect& r) const {
        Mat m(*this);
        m.data = data + (size_t)r.y * step + r.x;
        m.rows = r.height; m.cols = r.width;
        return m;
    }
    Mat rowRange(int a, int b) const { return (*this)(Rect(0, a, cols, b - a)); }
    Mat colRange(int a, int b) const { return (*this)(Rect(a, 0, b - a, rows)); }
    Mat row(int r) const { return rowRange(r, r + 1); }
    Mat clone() const {
        Mat m(rows, cols, CV_8UC1);
        for (int r = 0; r < rows; ++r) std::memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, (size_t)cols);
        return m;
    }
    inline void copyTo(const _OutputArray& dst) const;
};

class _InputArray {
public:
    const Mat* m;
    Mat none;
    _InputArray() : m(&none) {}
    _InputArray(const Mat& mm) : m(&mm) {}
    bool empty() const { return m->empty(); }
    Mat getMat() const { return *m; }
};
class _OutputArray {
public:
    Mat* m;
    _OutputArray(Mat& mm) : m(&mm) {}
    _OutputArray(const Mat& mm) : m(const_cast<Mat*>(&mm)) {}   // OpenCV has the same overload for temporaries (row(), ROI)
    void create(int r, int c, int type) const { m->create(r, c, type); }
    void release() const { m->release(); }
    Mat getMat() const { return *m; }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

inline void Mat::copyTo(const _OutputArray& dst) const {
    dst.create(rows, cols, CV_8UC1);
    Mat d = dst.getMat();
    for (int r = 0; r < rows; ++r) std::memmove(d.data + (size_t)r * d.step, data + (size_t)r * step, (size_t)cols);
}

enum { NORM_L1 = 2 };
// cv::norm(a, b, NORM_L1) for two 8-bit single-channel matrices of the same size: sum of absolute differences (exact integer)
inline double norm(const Mat& a, const Mat& b, int normType) {
    assert(normType == NORM_L1 && a.rows == b.rows && a.cols == b.cols);
    long s = 0;
    for (int r = 0; r < a.rows; ++r) {
        const uchar* pa = a.data + (size_t)r * a.step; const uchar* pb = b.data + (size_t)r * b.step;
        for (int c = 0; c < a.cols; ++c) s += pa[c] > pb[c] ? pa[c] - pb[c] : pb[c] - pa[c];
    }
    return (double)s;
}

// cv::FileStorage / cv::FileNode: DBoW2's TemplatedVocabulary.h (compiled verbatim into oracle/_ref for its transform()) names them in its
// YAML save / load members.  Those members are never called here; these declarations only let the header parse.
struct FileNode {
    enum { NONE = 0, SEQ = 5, MAP = 6 };
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](const std::string&) const { return FileNode(); }
    FileNode operator[](int) const { return FileNode(); }
    int type() const { return NONE; }
    size_t size() const { return 0; }
    operator int() const { return 0; }
    operator double() const { return 0; }
    operator float() const { return 0; }
    operator std::string() const { return std::string(); }
};
struct FileStorage {
    enum { READ = 0, WRITE = 1 };
    FileStorage() {}
    FileStorage(const std::string&, int) {}
    bool isOpened() const { return false; }
    void release() {}
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](const std::string&) const { return FileNode(); }
};
template <class T> inline FileStorage& operator<<(FileStorage& fs, const T&) { return fs; }
#define CV_32F 5

// the five primitives: implemented in minicv.cpp on top of the cv2-pinned oracle primitives
void FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true);
void resize(InputArray src, OutputArray dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void copyMakeBorder(InputArray src, OutputArray dst, int top, int bottom, int left, int right, int borderType);
void GaussianBlur(InputArray src, OutputArray dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT);
float fastAtan2(float y, float x);

// only referenced by the reference's dead ComputeKeyPointsOld (src/ORBextractor.cc:898-1075, never called)
struct KeyPointsFilter {
    static void retainBest(std::vector<KeyPoint>& keypoints, int npoints);
};

}  // namespace cv
