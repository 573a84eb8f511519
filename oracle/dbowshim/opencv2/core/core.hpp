/* oracle/dbowshim -- the OpenCV surface the reference's DBoW2 (S/Thirdparty/DBoW2/DBoW2/*.{h,cpp}) needs, so that TemplatedVocabulary /
 * FORB / BowVector / FeatureVector / ScoringObject compile UNMODIFIED into oracle/_ref/libref_dbow2.so (oracle/Makefile).  cv::Mat is a byte /
 * float matrix; cv::FileStorage (the YAML save / load path, never exercised: ORBSLAMM loads ORBvoc.txt with loadFromTextFile) is an inert stub.
 * TEST INFRASTRUCTURE ONLY. */
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <iostream>
#include <sstream>
#include <memory>
#include <string>
#include <vector>
#define CV_8U 0
#define CV_32F 5
namespace cv {
class Mat {
public:
    int rows = 0, cols = 0;
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    void create(int r, int c, int type)
    {
        if (buf && rows == r && cols == c && type_ == type) return;
        rows = r; cols = c; type_ = type;
        buf = std::make_shared<std::vector<unsigned char>>((size_t)r * c * (type == CV_32F ? 4 : 1) + 16, 0);
    }
    static Mat zeros(int r, int c, int type) { Mat m; m.create(r, c, type); std::memset(m.buf->data(), 0, m.buf->size()); return m; }
    void release() { buf.reset(); rows = cols = 0; }
    bool empty() const { return !buf || rows * cols == 0; }
    Mat clone() const { Mat m; m.rows = rows; m.cols = cols; m.type_ = type_; if (buf) m.buf = std::make_shared<std::vector<unsigned char>>(*buf); return m; }
    template <typename T> T *ptr(int r = 0) { return (T *)(buf->data() + (size_t)r * cols * (type_ == CV_32F ? 4 : 1)); }
    template <typename T> const T *ptr(int r = 0) const { return (const T *)(buf->data() + (size_t)r * cols * (type_ == CV_32F ? 4 : 1)); }
    template <typename T> T &at(int r, int c) { return ptr<T>(r)[c]; }
private:
    int type_ = CV_8U;
    std::shared_ptr<std::vector<unsigned char>> buf;
};
class FileNode {
public:
    FileNode operator[](const char *) const { return FileNode(); }
    FileNode operator[](const std::string &) const { return FileNode(); }
    FileNode operator[](int) const { return FileNode(); }
    size_t size() const { return 0; }
    operator int() const { return 0; }
    operator double() const { return 0; }
    operator std::string() const { return std::string(); }
};
class FileStorage {
public:
    enum { READ = 0, WRITE = 1 };
    FileStorage() {}
    FileStorage(const std::string &, int) {}
    bool isOpened() const { return false; }
    void release() {}
    FileNode operator[](const char *) const { return FileNode(); }
    FileNode operator[](const std::string &) const { return FileNode(); }
};
template <typename T> static inline FileStorage &operator<<(FileStorage &fs, const T &) { return fs; }
}  // namespace cv
