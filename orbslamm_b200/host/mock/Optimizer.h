// Mock of S/include/Optimizer.h:37-68 restricted to the functions the accelerated path defines.
#pragma once
#include "Frame.h"
#include "KeyFrame.h"
#include "Map.h"
#include "Thirdparty/g2o/g2o/types/sim3.h"
#include "LoopClosing.h"

namespace iORB_SLAM
{
class Optimizer
{
public:
    void static BundleAdjustment(const std::vector<KeyFrame *> &vpKF, const std::vector<MapPoint *> &vpMP, int nIterations = 5,
                                 bool *pbStopFlag = NULL, const unsigned long nLoopKF = 0, const bool bRobust = true);
    void static GlobalBundleAdjustemnt(Map *pMap, int nIterations = 5, bool *pbStopFlag = NULL, const unsigned long nLoopKF = 0,
                                       const bool bRobust = true);
#ifdef ORBSLAMM_MULTI_ROBOT
    void static MMGlobalBundleAdjustemnt(Map *pMap, int nIterations = 5, bool *pbStopFlag = NULL, const unsigned long nLoopKF = 0, const bool bRobust = true);
#endif
    void static LocalBundleAdjustment(KeyFrame *pKF, bool *pbStopFlag, Map *pMap);
    int static PoseOptimization(Frame *pFrame);
    void static OptimizeEssentialGraph(Map *pMap, KeyFrame *pLoopKF, KeyFrame *pCurKF, const LoopClosing::KeyFrameAndPose &NonCorrectedSim3,
                                       const LoopClosing::KeyFrameAndPose &CorrectedSim3, const std::map<KeyFrame *, std::set<KeyFrame *>> &LoopConnections,
                                       const bool &bFixScale);
#ifdef ORBSLAMM_MULTI_ROBOT
    void static MMOptimizeEssentialGraph(Map *pMap, KeyFrame *pLoopKF, KeyFrame *pCurKF, const LoopClosing::KeyFrameAndPose &NonCorrectedSim3,
                                         const LoopClosing::KeyFrameAndPose &CorrectedSim3, const std::map<KeyFrame *, std::set<KeyFrame *>> &LoopConnections,
                                         const bool &bFixScale);
#endif
    static int OptimizeSim3(KeyFrame *pKF1, KeyFrame *pKF2, std::vector<MapPoint *> &vpMatches1, g2o::Sim3 &g2oS12, const float th2, const bool bFixScale);
};
}  // namespace iORB_SLAM
