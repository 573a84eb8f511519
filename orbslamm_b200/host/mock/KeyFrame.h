#pragma once
#include "MapPoint.h"

namespace iORB_SLAM
{
class KeyFrame
{
public:
    KeyFrame() : Tcw(cv::Mat::eye(4, 4, CV_32F)) {}
    cv::Mat GetPose() { return Tcw.clone(); }
    void SetPose(const cv::Mat &T) { Tcw = T.clone(); }
    bool isBad() { return mbBad; }
    bool isNotFixed() { return false; }                          // M/include/KeyFrame.h:113-115 (multi-robot tree)
    std::vector<KeyFrame *> GetVectorCovisibleKeyFrames() { return mvpOrderedConnectedKeyFrames; }
    std::vector<MapPoint *> GetMapPointMatches() { return mvpMapPoints; }
    void EraseMapPointMatch(MapPoint *pMP) { for (auto &p : mvpMapPoints) if (p == pMP) p = nullptr; }

    long unsigned int mnId = 0, mnBALocalForKF = ~0ul, mnBAFixedForKF = ~0ul, mnBAGlobalForKF = 0;
    float fx = 0, fy = 0, cx = 0, cy = 0;
    std::vector<cv::KeyPoint> mvKeysUn;
    std::vector<float> mvuRight, mvInvLevelSigma2;
    std::vector<MapPoint *> mvpMapPoints;
    std::vector<KeyFrame *> mvpOrderedConnectedKeyFrames;
    cv::Mat Tcw, mTcwGBA;
    bool mbBad = false;
};
}  // namespace iORB_SLAM
