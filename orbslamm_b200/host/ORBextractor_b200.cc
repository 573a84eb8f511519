// ORBextractor_b200.cc -- replaces S/src/ORBextractor.cc in libORBSLAMM.so: marshals cv::Mat / cv::KeyPoint to the
// flat-array C-ABI (include/orbslamm_b200.h).  All arithmetic happens in liborbslamm_b200.so on the GPU.
#include <cassert>
#include <cstring>
#include <stdexcept>
#include <string>
#include "ORBextractor.h"
#include "orbslamm_b200.h"

namespace iORB_SLAM
{

static int g_device = 0;
void ORBextractor::SetDevice(int device) { g_device = device; }

static void check(int rc, const char *what)
{
    // the reference has no error channel; a CUDA failure here is unrecoverable for tracking, so fail loudly
    if (rc != ORBS_OK) throw std::runtime_error(std::string(what) + ": " + orbs_last_error());
}

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST), minThFAST(_minThFAST),
      mHandle(nullptr), mLastW(0), mLastH(0)
{
    check(orbx_create(&mHandle, _nfeatures, _scaleFactor, _nlevels, _iniThFAST, _minThFAST, g_device), "orbx_create");
    mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels); mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
    mnFeaturesPerLevel.resize(nlevels);
    check(orbx_get_tables(mHandle, nullptr, nullptr, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(),
                          mvInvLevelSigma2.data(), mnFeaturesPerLevel.data()), "orbx_get_tables");
    mvImagePyramid.resize(nlevels);
}

ORBextractor::~ORBextractor() { orbx_destroy(mHandle); }

void ORBextractor::operator()(cv::InputArray _image, cv::InputArray /*mask*/, std::vector<cv::KeyPoint> &_keypoints, cv::OutputArray _descriptors)
{
    if (_image.empty()) return;                                    // ORBextractor.cc:1046-1047
    cv::Mat image = _image.getMat();
    assert(image.type() == CV_8UC1);                               // ORBextractor.cc:1050
    int cap = 0;
    check(orbx_max_keypoints(mHandle, image.cols, image.rows, &cap), "orbx_max_keypoints");
    mLastW = image.cols; mLastH = image.rows;
    mXY.resize(2 * (size_t)cap); mAngle.resize(cap); mResponse.resize(cap); mSize.resize(cap); mOctave.resize(cap);
    cv::Mat desc(cap, 32, CV_8U);
    int32_t n = 0;
    check(orbx_extract(mHandle, image.ptr(0), 1, image.cols, image.rows, (int)image.step, image.step * image.rows, mXY.data(), mAngle.data(),
                       mResponse.data(), mOctave.data(), mSize.data(), desc.ptr(0), cap, &n), "orbx_extract");
    _keypoints.clear();
    if (n == 0) { _descriptors.release(); return; }                // ORBextractor.cc:1064-1065
    _keypoints.reserve(n);
    for (int i = 0; i < n; i++) _keypoints.push_back(cv::KeyPoint(mXY[2 * i], mXY[2 * i + 1], mSize[i], mAngle[i], mResponse[i], mOctave[i], -1));
    _descriptors.create(n, 32, CV_8U);                             // ORBextractor.cc:1068
    cv::Mat out = _descriptors.getMat();                           // (OutputArray is a proxy: no ptr() of its own)
    for (int i = 0; i < n; i++) std::memcpy(out.ptr(i), desc.ptr(i), 32);
}

void ORBextractor::FetchImagePyramid()
{
    if (!mLastW) return;
    for (int l = 0; l < nlevels; l++) {
        int lw = 0, lh = 0;
        check(orbx_level_size(mHandle, mLastW, mLastH, l, &lw, &lh), "orbx_level_size");
        cv::Mat m(lh, lw, CV_8UC1);
        check(orbx_get_pyramid_level(mHandle, 0, l, 0, m.ptr(0), (int)m.step), "orbx_get_pyramid_level");
        mvImagePyramid[l] = m;
    }
}

}  // namespace iORB_SLAM
