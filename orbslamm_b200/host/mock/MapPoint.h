// Mock of the reference data model (S/include/{MapPoint,KeyFrame,Frame,Map}.h): only the members the hot-path shims
// touch, with the reference's names.  Test scaffolding for an image without the reference's dependencies; in the real
// tree the shims are compiled against the reference's own headers.
#pragma once
#include <map>
#include <mutex>
#include <vector>
#include <opencv2/core/core.hpp>

namespace iORB_SLAM
{
class KeyFrame;

class MapPoint
{
public:
    MapPoint() : mWorldPos(3, 1, CV_32F), mDescriptor(1, 32, CV_8U) {}
    cv::Mat GetWorldPos() { return mWorldPos.clone(); }
    void SetWorldPos(const cv::Mat &p) { mWorldPos = p.clone(); }
    cv::Mat GetDescriptor() { return mDescriptor.clone(); }
    bool isBad() { return mbBad; }
    int Observations() { return nObs; }
    std::map<KeyFrame *, size_t> GetObservations() { return mObservations; }
    void EraseObservation(KeyFrame *pKF) { if (mObservations.erase(pKF)) nObs--; }
    void UpdateNormalAndDepth() { nNormalUpdates++; }

    long unsigned int mnId = 0;
    float mTrackProjX = 0, mTrackProjY = 0, mTrackViewCos = 1;
    bool mbTrackInView = false;
    int mnTrackScaleLevel = 0;
    long unsigned int mnBALocalForKF = ~0ul, mnBAGlobalForKF = 0;
    cv::Mat mPosGBA;
    static std::mutex mGlobalMutex;

    cv::Mat mWorldPos, mDescriptor;
    std::map<KeyFrame *, size_t> mObservations;
    int nObs = 1, nNormalUpdates = 0;
    bool mbBad = false;
};
}  // namespace iORB_SLAM
