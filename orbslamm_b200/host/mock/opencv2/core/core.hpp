// Minimal stand-in for the OpenCV types the ORBSLAMM hot-path classes touch (cv::Mat, cv::KeyPoint, ...).
// ONLY for compiling and testing the host shims in an image without OpenCV headers; a real build includes the
// real <opencv2/core/core.hpp> instead (put the OpenCV include dir before host/mock on the include path).
#pragma once
#include <cstring>
#include <memory>
#include <vector>
#include <cmath>

#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5

namespace cv {

struct Point2f { float x = 0, y = 0; Point2f() {} Point2f(float a, float b) : x(a), y(b) {} };
struct Point2i { int x = 0, y = 0; Point2i() {} Point2i(int a, int b) : x(a), y(b) {} };
typedef Point2i Point;

struct KeyPoint {
    Point2f pt; float size = 0, angle = -1, response = 0; int octave = 0, class_id = -1;
    KeyPoint() {}
    KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1) : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
};

class Mat {
public:
    int rows = 0, cols = 0;
    size_t step = 0;
    unsigned char *data = nullptr;
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    void create(int r, int c, int type)
    {
        type_ = type; rows = r; cols = c;
        step = (size_t)c * elemSize();
        buf_ = std::shared_ptr<unsigned char>(new unsigned char[step * (r > 0 ? r : 1)], std::default_delete<unsigned char[]>());
        data = buf_.get();
    }
    void release() { buf_.reset(); data = nullptr; rows = cols = 0; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return type_; }
    size_t elemSize() const { return type_ == CV_32F ? 4 : 1; }
    bool isContinuous() const { return step == (size_t)cols * elemSize(); }
    template <typename T> T &at(int r, int c) { return *reinterpret_cast<T *>(data + step * r + sizeof(T) * c); }
    template <typename T> const T &at(int r, int c) const { return *reinterpret_cast<const T *>(data + step * r + sizeof(T) * c); }
    template <typename T> T &at(int i) { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    template <typename T> const T &at(int i) const { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    template <typename T> T *ptr(int r = 0) { return reinterpret_cast<T *>(data + step * r); }
    template <typename T> const T *ptr(int r = 0) const { return reinterpret_cast<const T *>(data + step * r); }
    unsigned char *ptr(int r = 0) { return data + step * r; }
    const unsigned char *ptr(int r = 0) const { return data + step * r; }
    Mat row(int r) const { Mat m; m.type_ = type_; m.rows = 1; m.cols = cols; m.step = step; m.data = data + step * r; m.buf_ = buf_; return m; }
    Mat clone() const { Mat m(rows, cols, type_); for (int r = 0; r < rows; r++) std::memcpy(m.ptr(r), ptr(r), (size_t)cols * elemSize()); return m; }
    Mat getMat() const { return *this; }
    static Mat eye(int r, int c, int type) { Mat m(r, c, type); std::memset(m.data, 0, m.step * r); for (int i = 0; i < r && i < c; i++) m.at<float>(i, i) = 1.f; return m; }
private:
    int type_ = CV_8U;
    std::shared_ptr<unsigned char> buf_;
};

// Proxy types like OpenCV's own (core/mat.hpp): InputArray / OutputArray are NOT cv::Mat -- no ptr(), no rows -- so that API misuse in the drop-in sources fails
// in this mock build exactly as it would against real OpenCV.
class _InputArray {
public:
    _InputArray() {}
    _InputArray(const Mat &m) : m_(const_cast<Mat *>(&m)) {}
    Mat getMat() const { return m_ ? *m_ : Mat(); }
    bool empty() const { return !m_ || m_->empty(); }
protected:
    Mat *m_ = nullptr;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray() {}
    _OutputArray(Mat &m) : _InputArray(m) {}
    void create(int r, int c, int type) const { if (m_) m_->create(r, c, type); }
    void release() const { if (m_) m_->release(); }
};
typedef const _InputArray &InputArray;
typedef const _OutputArray &OutputArray;

}  // namespace cv
