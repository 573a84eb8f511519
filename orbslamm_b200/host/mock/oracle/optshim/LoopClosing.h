// Reference build of Optimizer.cc: the one thing Optimizer.h needs from S/include/LoopClosing.h:49-51, with the reference's own g2o::Sim3.
// Test infrastructure only.
#pragma once
#include <map>
#include "KeyFrame.h"
#include "Thirdparty/g2o/g2o/types/types_seven_dof_expmap.h"
namespace iORB_SLAM
{
class LoopClosing
{
public:
    typedef std::pair<std::set<KeyFrame *>, int> ConsistentGroup;
    typedef std::map<KeyFrame *, g2o::Sim3, std::less<KeyFrame *>, Eigen::aligned_allocator<std::pair<KeyFrame *const, g2o::Sim3> > > KeyFrameAndPose;
};
}  // namespace iORB_SLAM
