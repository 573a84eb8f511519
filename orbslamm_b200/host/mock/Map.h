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
    bool isAttached() { return !mvAttached.empty(); }                    // M/include/Map.h (MultiMapper)
    std::vector<Map *> getAttachedMaps() { return mvAttached; }
    std::vector<Map *> mvAttached;
    std::mutex mMutexMapUpdate;
    std::vector<KeyFrame *> mvKFs;
    std::vector<MapPoint *> mvMPs;
};
}  // namespace iORB_SLAM
