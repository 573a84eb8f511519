#pragma once
#include <mutex>
#include "KeyFrame.h"

namespace iORB_SLAM
{
class Map
{
public:
    std::vector<KeyFrame *> GetAllKeyFrames() { return mvKFs; }
    std::vector<MapPoint *> GetAllMapPoints() { return mvMPs; }
    long unsigned int GetMaxKFid() { long unsigned int m = 0; for (KeyFrame *k : mvKFs) m = std::max(m, k->mnId); return m; }   // Map.cc:120-124 (mnMaxKFid)
    bool isAttached() { return !mvAttached.empty(); }                    // M/include/Map.h (MultiMapper)
    std::vector<Map *> getAttachedMaps() { return mvAttached; }
    std::vector<Map *> mvAttached;
    std::mutex mMutexMapUpdate;
    std::vector<KeyFrame *> mvKFs;
    std::vector<MapPoint *> mvMPs;
};
}  // namespace iORB_SLAM
