// Mock of S/include/ORBmatcher.h:37-102 restricted to the members the accelerated path defines.
#pragma once
#include <vector>
#include "Frame.h"
#include "KeyFrame.h"

namespace iORB_SLAM
{
class ORBmatcher
{
public:
    ORBmatcher(float nnratio = 0.6, bool checkOri = true);
    static int DescriptorDistance(const cv::Mat &a, const cv::Mat &b);
    int SearchByProjection(Frame &F, const std::vector<MapPoint *> &vpMapPoints, const float th = 3);
    int SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono);
    static const int TH_LOW;
    static const int TH_HIGH;
    static const int HISTO_LENGTH;
protected:
    float RadiusByViewingCos(const float &viewCos);
    float mfNNratio;
    bool mbCheckOrientation;
};
}  // namespace iORB_SLAM
