// Mock of the one thing Optimizer.h needs from S/include/LoopClosing.h:49-51: the KeyFrameAndPose map (the reference uses an Eigen aligned allocator).
#pragma once
#include <map>
#include "KeyFrame.h"
#include "Thirdparty/g2o/g2o/types/sim3.h"
namespace iORB_SLAM
{
class LoopClosing
{
public:
    typedef std::map<KeyFrame *, g2o::Sim3, std::less<KeyFrame *>> KeyFrameAndPose;
};
}  // namespace iORB_SLAM
