/*
 * oracle/cvshim/cvshim.h -- minimal stand-in for the OpenCV C++ API surface that
 * /root/reference/SingleRobotScenario/src/ORBextractor.cc uses.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  OpenCV's C++ headers are not in this image, so the
 * reference extractor cannot be built as shipped.  This header lets the reference's OWN, UNMODIFIED
 * ORBextractor.cc / ORBextractor.h be compiled from where they lie (oracle/Makefile -> oracle/_ref/): containers
 * (cv::Mat, cv::KeyPoint, cv::Point_, ...) are re-declared here with OpenCV's member names, and the five OpenCV
 * *algorithms* the file calls (cv::FAST, cv::resize, cv::GaussianBlur, cv::copyMakeBorder, cv::fastAtan2) forward to
 * the primitives of oracle/orb_oracle.c, each of which is pinned bit-exact against cv2 4.13.0 by
 * tests/test_oracle_opencv_pin.py.  Everything else -- the cell loop, DistributeOctTree, IC_Angle, the steered
 * BRIEF sampling, the pyramid bookkeeping -- is the reference's own object code.
 */
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <iterator>
#include <list>
#include <memory>
#include <vector>

extern "C" {
void oracle_resize_linear_u8(const uint8_t *src, int sw, int sh, int sstride, uint8_t *dst, int dw, int dh, int dstride);
void oracle_gaussian_blur7_u8(const uint8_t *src, int w, int h, int sstride, uint8_t *dst, int dstride);
int oracle_fast9_16(const uint8_t *img, int w, int h, int stride, int threshold, int nms, int *out_xys, int cap);
float oracle_fast_atan2(float y, float x);
}

typedef unsigned char uchar;
#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0

static inline int cvRound(double v) { return (int)lrint(v); }   /* round-half-even like SSE2 cvtsd2si */
static inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
static inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }

namespace cv {

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T _x, T _y) : x(_x), y(_y) {}
    template <typename U> Point_(const Point_<U> &o) : x((T)o.x), y((T)o.y) {}
};
template <typename T> static inline Point_<T> &operator*=(Point_<T> &a, float b) { a.x = (T)(a.x * b); a.y = (T)(a.y * b); return a; }
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;

struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };
struct Rect { int x, y, width, height; Rect(int _x, int _y, int w, int h) : x(_x), y(_y), width(w), height(h) {} };

struct KeyPoint {
    Point2f pt; float size, angle, response; int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0, int _class_id = -1)
        : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
};

struct KeyPointsFilter {   /* only used by the dead ComputeKeyPointsOld */
    static void retainBest(std::vector<KeyPoint> &k, int n) {
        if (n >= 0 && (int)k.size() > n) {
            std::stable_sort(k.begin(), k.end(), [](const KeyPoint &a, const KeyPoint &b) { return a.response > b.response; });
            k.resize(n);
        }
    }
};

enum { BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16 };
enum { INTER_LINEAR = 1 };

struct MatStep { size_t v; MatStep() : v(0) {} operator size_t() const { return v; } };

struct MatExpr { int rows, cols; };   /* only Mat::zeros */

class Mat {   /* single channel u8 only */
public:
    int rows, cols; uchar *data; MatStep step;
    Mat() : rows(0), cols(0), data(nullptr) {}
    Mat(int r, int c, int type) { alloc(r, c); (void)type; }
    Mat(Size s, int type) { alloc(s.height, s.width); (void)type; }
    Mat(int r, int c, int type, void *ext, size_t stride) : rows(r), cols(c), data((uchar *)ext) { step.v = stride; (void)type; }
    /* Mat::zeros returns a MatExpr in OpenCV; assigning it to a Mat that already has that size and type fills the existing
     * buffer IN PLACE (MatOp_Initializer::assign -> create() is a no-op -> setTo(0)).  computeDescriptors
     * (ORBextractor.cc:1037) relies on this to write into a rowRange of the caller's descriptor matrix. */
    static MatExpr zeros(int r, int c, int type) { MatExpr e; e.rows = r; e.cols = c; (void)type; return e; }
    Mat &operator=(const MatExpr &e) { create(e.rows, e.cols, CV_8UC1); for (int y = 0; y < rows; y++) std::memset(data + (size_t)y * step.v, 0, cols); return *this; }
    Mat(const MatExpr &e) { alloc(e.rows, e.cols); std::memset(data, 0, (size_t)rows * cols); }
    int type() const { return CV_8UC1; }
    bool empty() const { return data == nullptr || rows * cols == 0; }
    size_t step1() const { return step.v; }
    void release() { buf.reset(); data = nullptr; rows = cols = 0; step.v = 0; }
    void create(int r, int c, int type) { if (r != rows || c != cols || !data) alloc(r, c); (void)type; }
    Mat clone() const { Mat m(rows, cols, CV_8UC1); for (int y = 0; y < rows; y++) std::memcpy(m.data + (size_t)y * m.step.v, data + (size_t)y * step.v, cols); return m; }
    Mat operator()(const Rect &r) const { Mat m = *this; m.data = data + (size_t)r.y * step.v + r.x; m.rows = r.height; m.cols = r.width; return m; }
    Mat rowRange(int a, int b) const { Mat m = *this; m.data = data + (size_t)a * step.v; m.rows = b - a; return m; }
    Mat colRange(int a, int b) const { Mat m = *this; m.data = data + a; m.cols = b - a; return m; }
    template <typename T> T &at(int r, int c) { return *(T *)(data + (size_t)r * step.v + c * sizeof(T)); }
    template <typename T> const T &at(int r, int c) const { return *(const T *)(data + (size_t)r * step.v + c * sizeof(T)); }
    uchar *ptr(int r = 0) { return data + (size_t)r * step.v; }
    const uchar *ptr(int r = 0) const { return data + (size_t)r * step.v; }
    template <typename T> T *ptr(int r = 0) { return (T *)(data + (size_t)r * step.v); }
private:
    std::shared_ptr<std::vector<uchar>> buf;
    void alloc(int r, int c) { rows = r; cols = c; buf = std::make_shared<std::vector<uchar>>((size_t)r * c + 64); data = buf->data(); step.v = (size_t)c; }
};

class _InputArray {
public:
    _InputArray(const Mat &m) : m_(&m) {}
    bool empty() const { return m_->empty(); }
    Mat getMat() const { return *m_; }
private:
    const Mat *m_;
};
typedef const _InputArray &InputArray;

class _OutputArray {
public:
    _OutputArray(Mat &m) : m_(&m) {}
    void create(int r, int c, int type) const { m_->create(r, c, type); }
    void create(Size s, int type) const { m_->create(s.height, s.width, type); }
    void release() const { m_->release(); }
    Mat getMat() const { return *m_; }
    Mat &ref() const { return *m_; }
private:
    Mat *m_;
};
typedef const _OutputArray &OutputArray;

static inline float fastAtan2(float y, float x) { return oracle_fast_atan2(y, x); }

static inline int borderReflect101(int p, int len) { if (len == 1) return 0; while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p; return p; }

/* cv::FAST(image, keypoints, threshold, nonmaxSuppression): TYPE_9_16; keypoints in raster order, size 7, angle -1 */
static inline void FAST(InputArray image, std::vector<KeyPoint> &kps, int threshold, bool nms = true)
{
    Mat m = image.getMat();
    kps.clear();
    if (m.cols < 7 || m.rows < 7) return;
    const int cap = m.cols * m.rows;
    std::vector<int> xys((size_t)3 * cap);
    const int n = oracle_fast9_16(m.data, m.cols, m.rows, (int)m.step.v, threshold, nms ? 1 : 0, xys.data(), cap);
    for (int i = 0; i < n; i++) kps.push_back(KeyPoint((float)xys[3 * i], (float)xys[3 * i + 1], 7.f, -1, (float)xys[3 * i + 2]));
}

/* cv::resize(src, dst, dsize, 0, 0, INTER_LINEAR) on CV_8UC1; dst keeps its buffer when it already has dsize (ROI write) */
static inline void resize(InputArray src, OutputArray dst, Size dsize, double fx = 0, double fy = 0, int interp = INTER_LINEAR)
{
    assert(interp == INTER_LINEAR && fx == 0 && fy == 0);
    Mat s = src.getMat();
    dst.create(dsize, CV_8UC1);
    Mat d = dst.getMat();
    oracle_resize_linear_u8(s.data, s.cols, s.rows, (int)s.step.v, d.data, d.cols, d.rows, (int)d.step.v);
}

/* cv::GaussianBlur(src, dst, Size(7,7), 2, 2, BORDER_REFLECT_101) on CV_8UC1 (in place allowed) */
static inline void GaussianBlur(InputArray src, OutputArray dst, Size ksize, double sx, double sy = 0, int border = BORDER_REFLECT_101)
{
    assert(ksize.width == 7 && ksize.height == 7 && sx == 2 && sy == 2 && border == BORDER_REFLECT_101);
    Mat s = src.getMat().clone();
    dst.create(s.rows, s.cols, CV_8UC1);
    Mat d = dst.getMat();
    oracle_gaussian_blur7_u8(s.data, s.cols, s.rows, (int)s.step.v, d.data, (int)d.step.v);
}

/* cv::copyMakeBorder(src, dst, t, b, l, r, BORDER_REFLECT_101 [+ BORDER_ISOLATED]); src may be the interior ROI of dst */
static inline void copyMakeBorder(InputArray src, OutputArray dst, int top, int bottom, int left, int right, int type)
{
    assert((type & ~BORDER_ISOLATED) == BORDER_REFLECT_101);
    Mat s = src.getMat();
    dst.create(s.rows + top + bottom, s.cols + left + right, CV_8UC1);
    Mat d = dst.getMat();
    Mat inner = s.clone();
    for (int y = 0; y < d.rows; y++) {
        const uchar *sr = inner.ptr(borderReflect101(y - top, s.rows));
        uchar *dr = d.ptr(y);
        for (int x = 0; x < d.cols; x++) dr[x] = sr[borderReflect101(x - left, s.cols)];
    }
}

}  // namespace cv
