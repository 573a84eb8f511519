#pragma once
#include "MapPoint.h"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"

#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64

namespace iORB_SLAM
{
class Frame
{
public:
    Frame() : mTcw(cv::Mat::eye(4, 4, CV_32F)) {}
    void SetPose(cv::Mat Tcw) { mTcw = Tcw.clone(); }

    static float fx, fy, cx, cy;
    static float mnMinX, mnMaxX, mnMinY, mnMaxY;
    float mb = 0, mbf = 0;
    int N = 0;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    std::vector<float> mvuRight;
    cv::Mat mDescriptors;
    std::vector<MapPoint *> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    cv::Mat mTcw;
    std::vector<float> mvScaleFactors, mvInvLevelSigma2;
    float mfLogScaleFactor = 0;
    DBoW2::FeatureVector mFeatVec;
};
}  // namespace iORB_SLAM
