// ORBextractor.h -- drop-in for the reference's S/include/ORBextractor.h:45-111 (namespace iORB_SLAM, same public
// members incl. the public mvImagePyramid), implemented over the orbslamm_b200 C-ABI (orbx_*).  Frame::ExtractORB
// (S/src/Frame.cc:247-253) and the Frame constructors (S/src/Frame.cc:69-75) call it unchanged.
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <list>
#include <vector>
#include <opencv2/core/core.hpp>

struct orbx_handle;

namespace iORB_SLAM
{

// kept for source compatibility (ORBextractor.h:32-43); the quad-tree itself runs on the device
class ExtractorNode
{
public:
    ExtractorNode() : bNoMore(false) {}
    std::vector<cv::KeyPoint> vKeys;
    cv::Point2i UL, UR, BL, BR;
    std::list<ExtractorNode>::iterator lit;
    bool bNoMore;
};

class ORBextractor
{
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
    ~ORBextractor();
    ORBextractor(const ORBextractor &) = delete;
    ORBextractor &operator=(const ORBextractor &) = delete;

    // Compute the ORB features and descriptors on an image (mask is ignored, as in the reference).
    void operator()(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint> &keypoints, cv::OutputArray descriptors);

    int inline GetLevels() { return nlevels; }
    float inline GetScaleFactor() { return scaleFactor; }
    std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

    // Filled lazily: call FetchImagePyramid() after operator() when the pyramid is needed on the host (stereo matching
    // in Frame.cc:471,561-578 is the only reader; the monocular path never touches it).
    std::vector<cv::Mat> mvImagePyramid;
    void FetchImagePyramid();

    // CUDA device ordinal used by extractors created afterwards (default 0); one per process / per System is typical
    static void SetDevice(int device);

protected:
    int nfeatures;
    double scaleFactor;
    int nlevels;
    int iniThFAST;
    int minThFAST;
    std::vector<int> mnFeaturesPerLevel;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;

    orbx_handle *mHandle;
    int mLastW, mLastH;
    // pinned-size scratch reused across frames
    std::vector<float> mXY, mAngle, mResponse, mSize;
    std::vector<int> mOctave;
};

}  // namespace iORB_SLAM

#endif
