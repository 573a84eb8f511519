/*
 * oracle/slamshim/slamshim_cv.h -- minimal stand-in for the OpenCV C++ surface that the reference's ORBmatcher.cc uses.
 *
 * TEST INFRASTRUCTURE ONLY.  Lets /root/reference/SingleRobotScenario/src/ORBmatcher.cc be compiled UNMODIFIED (oracle/Makefile ->
 * oracle/_ref/libref_orbmatcher.so) although neither OpenCV nor the reference's data model (Frame / KeyFrame / MapPoint -> DBoW2,
 * g2o, Eigen) can be built in this image.  cv::Mat here is a small float / byte matrix; its algebra follows what OpenCV 4.13 does
 * for these sizes (matrix product = the small-gemm path: fp32 products added left to right, pinned against cv2.gemm by
 * tests/test_oracle_opencv_pin.py; cv::norm and Mat::dot accumulate in double).
 */
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <vector>

typedef unsigned char uchar;
#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_32FC1 5
#define CV_PI 3.1415926535897932384626433832795
static inline int cvRound(double v) { return (int)lrint(v); }

namespace cv {

template <typename T> struct Point_ { T x, y; Point_() : x(0), y(0) {} Point_(T a, T b) : x(a), y(b) {} };
typedef Point_<float> Point2f;
template <typename T> struct Point3_ { T x, y, z; Point3_() : x(0), y(0), z(0) {} Point3_(T a, T b, T c) : x(a), y(b), z(c) {} };
typedef Point3_<float> Point3f;
typedef Point_<int> Point;

struct KeyPoint {
    Point2f pt; float size, angle, response; int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
};

struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };

class Mat {
public:
    int rows, cols;
    Mat() : rows(0), cols(0), type_(CV_32F), step_(0), data(nullptr) {}
    Mat(int r, int c, int type) { alloc(r, c, type); }
    Mat(Size sz, int type) { alloc(sz.height, sz.width, type); }
    Size size() const { return Size(cols, rows); }
    void create(int r, int c, int type) { if (data && rows == r && cols == c && type_ == type) return; alloc(r, c, type); }
    /* assignment: an lvalue rebinds (cv::Mat header semantics); a temporary view (P.col(i) = expr, OpenCV's MatExpr assignment) is written through */
    Mat(const Mat &) = default;
    Mat &operator=(const Mat &o) & { rows = o.rows; cols = o.cols; type_ = o.type_; step_ = o.step_; data = o.data; buf = o.buf; return *this; }
    Mat &operator=(const Mat &o) && { write_from(o); return *this; }
    static Mat zeros(int r, int c, int type) { Mat m(r, c, type); std::memset(m.data, 0, m.step_ * r); return m; }
    static Mat eye(int r, int c, int type) { Mat m = zeros(r, c, type); for (int i = 0; i < std::min(r, c); i++) m.at<float>(i, i) = 1.f; return m; }
    int type() const { return type_; }
    bool empty() const { return data == nullptr || rows * cols == 0; }
    size_t elem() const { return type_ == CV_32F ? 4 : 1; }
    template <typename T> T &at(int i, int j) { return *(T *)(data + (size_t)i * step_ + (size_t)j * sizeof(T)); }
    template <typename T> const T &at(int i, int j) const { return *(const T *)(data + (size_t)i * step_ + (size_t)j * sizeof(T)); }
    template <typename T> T &at(int i) { return cols == 1 ? at<T>(i, 0) : at<T>(0, i); }
    template <typename T> const T &at(int i) const { return cols == 1 ? at<T>(i, 0) : at<T>(0, i); }
    template <typename T> T *ptr(int r = 0) { return (T *)(data + (size_t)r * step_); }
    template <typename T> const T *ptr(int r = 0) const { return (const T *)(data + (size_t)r * step_); }
    Mat rowRange(int a, int b) const { Mat m = *this; m.data = data + (size_t)a * step_; m.rows = b - a; return m; }
    Mat colRange(int a, int b) const { Mat m = *this; m.data = data + (size_t)a * elem(); m.cols = b - a; return m; }
    Mat row(int i) const { return rowRange(i, i + 1); }
    Mat col(int j) const { return colRange(j, j + 1); }
    Mat clone() const { Mat m(rows, cols, type_); for (int i = 0; i < rows; i++) std::memcpy(m.data + (size_t)i * m.step_, data + (size_t)i * step_, (size_t)cols * elem()); return m; }
    void copyTo(Mat &o) const { if (o.data && o.rows == rows && o.cols == cols && o.type_ == type_) o.write_from(*this); else o = clone(); }
    void copyTo(Mat &&o) const { o.write_from(*this); }                       /* into a view: sR.copyTo(T.rowRange(0,3).colRange(0,3)) */
    void write_from(const Mat &o) { assert(rows == o.rows && cols == o.cols); for (int i = 0; i < rows; i++) std::memcpy(data + (size_t)i * step_, o.data + (size_t)i * o.step_, (size_t)cols * elem()); }
    Mat t() const { Mat m(cols, rows, type_); for (int i = 0; i < rows; i++) for (int j = 0; j < cols; j++) m.at<float>(j, i) = at<float>(i, j); return m; }
    double dot(const Mat &o) const
    {
        double s = 0;
        for (int i = 0; i < rows; i++) for (int j = 0; j < cols; j++) s += (double)at<float>(i, j) * (double)o.at<float>(i, j);
        return s;
    }
private:
    int type_; size_t step_;
public:
    uchar *data;
private:
    std::shared_ptr<std::vector<uchar>> buf;
    void alloc(int r, int c, int type) { rows = r; cols = c; type_ = type; step_ = (size_t)c * (type == CV_32F ? 4 : 1); buf = std::make_shared<std::vector<uchar>>(step_ * r + 16); data = buf->data(); }
};

/* A * B for float matrices: OpenCV's small-matrix gemm, fp32 products accumulated left to right */
static inline Mat operator*(const Mat &a, const Mat &b)
{
    assert(a.cols == b.rows);
    Mat m(a.rows, b.cols, CV_32F);
    for (int i = 0; i < a.rows; i++)
        for (int j = 0; j < b.cols; j++) {
            float s = a.at<float>(i, 0) * b.at<float>(0, j);
            for (int k = 1; k < a.cols; k++) s = s + a.at<float>(i, k) * b.at<float>(k, j);
            m.at<float>(i, j) = s;
        }
    return m;
}
#define SLAMSHIM_EW(op) \
    static inline Mat operator op(const Mat &a, const Mat &b) { Mat m(a.rows, a.cols, CV_32F); for (int i = 0; i < a.rows; i++) for (int j = 0; j < a.cols; j++) m.at<float>(i, j) = a.at<float>(i, j) op b.at<float>(i, j); return m; }
SLAMSHIM_EW(+)
SLAMSHIM_EW(-)
static inline Mat operator-(const Mat &a) { Mat m(a.rows, a.cols, CV_32F); for (int i = 0; i < a.rows; i++) for (int j = 0; j < a.cols; j++) m.at<float>(i, j) = -a.at<float>(i, j); return m; }
static inline Mat operator*(const Mat &a, double s) { Mat m(a.rows, a.cols, CV_32F); for (int i = 0; i < a.rows; i++) for (int j = 0; j < a.cols; j++) m.at<float>(i, j) = (float)(a.at<float>(i, j) * s); return m; }
static inline Mat operator*(double s, const Mat &a) { return a * s; }
static inline Mat operator/(const Mat &a, double s) { return a * (1. / s); }

static inline double norm(const Mat &a)
{
    double s = 0;
    for (int i = 0; i < a.rows; i++) for (int j = 0; j < a.cols; j++) { const double v = a.at<float>(i, j); s += v * v; }
    return std::sqrt(s);
}

/* cv::Mat_<float>(r, c) << a, b, ...  (MatCommaInitializer_) */
template <typename T> class Mat_;
template <typename T> struct MatCommaInitializer_ {
    Mat m; int k;
    MatCommaInitializer_(const Mat &mm, T first) : m(mm), k(0) { put(first); }
    void put(T v) { m.at<T>(k / m.cols, k % m.cols) = v; k++; }
    MatCommaInitializer_ &operator,(double v) { put((T)v); return *this; }
    operator Mat() const { return m; }
};
template <typename T> class Mat_ : public Mat {
public:
    Mat_(int r, int c) : Mat(r, c, CV_32F) {}
    MatCommaInitializer_<T> operator<<(double v) const { return MatCommaInitializer_<T>(*this, (T)v); }
};

#define CV_REDUCE_SUM 0
/* The four functions below exist so that the reference's Sim3Solver.cc compiles; ComputeSim3, the only caller, is NOT part of what oracle/_ref pins
 * (the pinned members are CheckInliers / Project / FromCameraToImage and the constructor), so these are plain double-precision stand-ins, not
 * restatements of OpenCV's float algorithms. */
static inline void reduce(const Mat &src, Mat &dst, int dim, int)
{
    assert(dim == 1);
    dst = Mat(src.rows, 1, CV_32F);
    for (int i = 0; i < src.rows; i++) { float s = src.at<float>(i, 0); for (int j = 1; j < src.cols; j++) s = s + src.at<float>(i, j); dst.at<float>(i, 0) = s; }
}
static inline void pow(const Mat &src, double p, Mat &dst)
{
    Mat m(src.rows, src.cols, CV_32F);
    for (int i = 0; i < src.rows; i++) for (int j = 0; j < src.cols; j++) m.at<float>(i, j) = (float)std::pow((double)src.at<float>(i, j), p);
    dst = m;
}
static inline bool eigen(const Mat &A, Mat &evals, Mat &evecs)      /* symmetric Jacobi in double, eigenvalues descending, eigenvectors as rows */
{
    const int n = A.rows;
    std::vector<double> a((size_t)n * n), v((size_t)n * n, 0.0);
    for (int i = 0; i < n; i++) { for (int j = 0; j < n; j++) a[(size_t)i * n + j] = A.at<float>(i, j); v[(size_t)i * n + i] = 1; }
    for (int sweep = 0; sweep < 64; sweep++) {
        double off = 0;
        for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) off += a[(size_t)i * n + j] * a[(size_t)i * n + j];
        if (off < 1e-300) break;
        for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) {
            if (a[(size_t)p * n + q] == 0) continue;
            const double th = (a[(size_t)q * n + q] - a[(size_t)p * n + p]) / (2 * a[(size_t)p * n + q]);
            const double t = (th >= 0 ? 1 : -1) / (std::fabs(th) + std::sqrt(th * th + 1)), c = 1 / std::sqrt(t * t + 1), s = t * c;
            for (int k = 0; k < n; k++) { const double x = a[(size_t)k * n + p], y = a[(size_t)k * n + q]; a[(size_t)k * n + p] = c * x - s * y; a[(size_t)k * n + q] = s * x + c * y; }
            for (int k = 0; k < n; k++) { const double x = a[(size_t)p * n + k], y = a[(size_t)q * n + k]; a[(size_t)p * n + k] = c * x - s * y; a[(size_t)q * n + k] = s * x + c * y; }
            for (int k = 0; k < n; k++) { const double x = v[(size_t)p * n + k], y = v[(size_t)q * n + k]; v[(size_t)p * n + k] = c * x - s * y; v[(size_t)q * n + k] = s * x + c * y; }
        }
    }
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int x, int y) { return a[(size_t)x * n + x] > a[(size_t)y * n + y]; });
    evals = Mat(n, 1, CV_32F); evecs = Mat(n, n, CV_32F);
    for (int i = 0; i < n; i++) { evals.at<float>(i, 0) = (float)a[(size_t)order[i] * n + order[i]]; for (int k = 0; k < n; k++) evecs.at<float>(i, k) = (float)v[(size_t)order[i] * n + k]; }
    return true;
}
static inline void Rodrigues(const Mat &rvec, Mat &R)
{
    const double r[3] = {rvec.at<float>(0), rvec.at<float>(1), rvec.at<float>(2)};
    const double theta = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    R = Mat::eye(3, 3, CV_32F);
    if (theta < 2.220446049250313e-16) return;
    const double c = std::cos(theta), s = std::sin(theta), c1 = 1 - c, x = r[0] / theta, y = r[1] / theta, z = r[2] / theta;
    const double M[9] = {c + c1 * x * x, c1 * x * y - s * z, c1 * x * z + s * y, c1 * x * y + s * z, c + c1 * y * y, c1 * y * z - s * x,
                         c1 * x * z - s * y, c1 * y * z + s * x, c + c1 * z * z};
    for (int i = 0; i < 9; i++) R.at<float>(i / 3, i % 3) = (float)M[i];
}

}  // namespace cv
