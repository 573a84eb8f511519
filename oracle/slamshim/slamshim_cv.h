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
typedef Point_<int> Point;

struct KeyPoint {
    Point2f pt; float size, angle, response; int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
};

class Mat {
public:
    int rows, cols;
    Mat() : rows(0), cols(0), type_(CV_32F), step_(0), data(nullptr) {}
    Mat(int r, int c, int type) { alloc(r, c, type); }
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
    void copyTo(Mat &o) const { o = clone(); }
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

}  // namespace cv
