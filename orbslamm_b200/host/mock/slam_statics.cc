// definitions of the mock data model's statics (test scaffolding)
#include "Frame.h"
#include "KeyFrame.h"
namespace iORB_SLAM
{
std::mutex MapPoint::mGlobalMutex;
float Frame::fx, Frame::fy, Frame::cx, Frame::cy;
float Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY;

// MapPoint::Replace (S/src/MapPoint.cc:187-228) without the map / descriptor bookkeeping
void MapPoint::Replace(MapPoint *pMP)
{
    if (pMP == this) return;
    std::map<KeyFrame *, size_t> obs = mObservations;
    mObservations.clear();
    mbBad = true;
    mpReplaced = pMP;
    for (auto &o : obs) {
        if (!pMP->IsInKeyFrame(o.first)) { o.first->ReplaceMapPointMatch(o.second, pMP); pMP->AddObservation(o.first, o.second); }
        else o.first->EraseMapPointMatch(o.second);
    }
}
}
