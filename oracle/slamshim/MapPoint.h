/* oracle/slamshim/MapPoint.h -- data-only stand-in for iORB_SLAM::MapPoint with the member names ORBmatcher.cc uses
 * (S/include/MapPoint.h); TEST INFRASTRUCTURE ONLY. */
#pragma once
#include <map>
#include "slamshim_cv.h"
namespace iORB_SLAM {
class KeyFrame;
class MapPoint {
public:
    MapPoint() : mnId(0), mTrackProjX(0), mTrackProjY(0), mTrackProjXR(0), mbTrackInView(false), mnTrackScaleLevel(0), mTrackViewCos(0), mnLastFrameSeen(0),
                 mnFuseCandidateForKF(0), mfMinDistance(0), mfMaxDistance(0), nObs(1), bad(false), mpReplaced(nullptr) {}
    cv::Mat GetWorldPos() { return mWorldPos.clone(); }
    cv::Mat GetNormal() { return mNormalVector.clone(); }
    cv::Mat GetDescriptor() { return mDescriptor.clone(); }
    int Observations() { return nObs; }
    bool isBad() { return bad; }
    float GetMinDistanceInvariance() { return 0.8f * mfMinDistance; }       /* MapPoint.cc:373-383 */
    float GetMaxDistanceInvariance() { return 1.2f * mfMaxDistance; }
    int PredictScale(const float &currentDist, const float &logScaleFactor) { float ratio = mfMaxDistance / currentDist; return std::ceil(std::log(ratio) / logScaleFactor); }
    bool IsInKeyFrame(KeyFrame *pKF) { return mObservations.count(pKF) != 0; }
    int GetIndexInKeyFrame(KeyFrame *pKF) { return mObservations.count(pKF) ? (int)mObservations[pKF] : -1; }
    void AddObservation(KeyFrame *pKF, size_t idx) { if (!mObservations.count(pKF)) { mObservations[pKF] = idx; nObs++; } }
    void Replace(MapPoint *pMP) { mpReplaced = pMP; bad = true; }
    long unsigned int mnId;
    float mTrackProjX, mTrackProjY, mTrackProjXR;
    bool mbTrackInView;
    int mnTrackScaleLevel;
    float mTrackViewCos;
    long unsigned int mnLastFrameSeen, mnFuseCandidateForKF;
    cv::Mat mWorldPos, mNormalVector, mDescriptor;
    float mfMinDistance, mfMaxDistance;
    int nObs; bool bad; MapPoint *mpReplaced;
    std::map<KeyFrame *, size_t> mObservations;
};
}
