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
    MapPoint() : mWorldPos(3, 1, CV_32F), mDescriptor(1, 32, CV_8U), mNormalVector(3, 1, CV_32F) {}
    cv::Mat GetWorldPos() { return mWorldPos.clone(); }
    void SetWorldPos(const cv::Mat &p) { mWorldPos = p.clone(); }
    cv::Mat GetDescriptor() { return mDescriptor.clone(); }
    // added by the drop-in (mDescriptor is protected in S/include/MapPoint.h and written under mMutexFeatures, MapPoint.cc:304-307): see INTEGRATION.md
    void SetDescriptor(const cv::Mat &d) { mDescriptor = d.clone(); }
    bool isBad() { return mbBad; }
    int Observations() { return nObs; }
    std::map<KeyFrame *, size_t> GetObservations() { return mObservations; }
    void EraseObservation(KeyFrame *pKF) { if (mObservations.erase(pKF)) nObs--; }
    void UpdateNormalAndDepth() { nNormalUpdates++; }
    cv::Mat GetNormal() { return mNormalVector.clone(); }
    float GetMinDistanceInvariance() { return 0.8f * mfMinDistance; }        // MapPoint.cc:373-383
    float GetMaxDistanceInvariance() { return 1.2f * mfMaxDistance; }
    // The two accessors the drop-in needs added to S/include/MapPoint.h (mfMinDistance / mfMaxDistance are protected there and
    // PredictScale, MapPoint.cc:385-394, reads the raw mfMaxDistance): see INTEGRATION.md
    float GetMinDistance() { return mfMinDistance; }
    float GetMaxDistance() { return mfMaxDistance; }
    bool IsInKeyFrame(KeyFrame *pKF) { return mObservations.count(pKF) != 0; }
    int GetIndexInKeyFrame(KeyFrame *pKF) { return mObservations.count(pKF) ? (int)mObservations[pKF] : -1; }
    void AddObservation(KeyFrame *pKF, size_t idx) { if (mObservations.count(pKF)) return; mObservations[pKF] = idx; nObs++; }
    void Replace(MapPoint *pMP);                                             // MapPoint.cc:187-228, defined in slam_statics.cc

    long unsigned int mnId = 0;
    float mTrackProjX = 0, mTrackProjY = 0, mTrackViewCos = 1;
    bool mbTrackInView = false;
    int mnTrackScaleLevel = 0;
    long unsigned int mnBALocalForKF = ~0ul, mnBAGlobalForKF = 0, mnCorrectedByKF = ~0ul, mnCorrectedReference = 0;
    KeyFrame *GetReferenceKeyFrame() { return mpRefKF; }
    KeyFrame *mpRefKF = nullptr;
    cv::Mat mPosGBA;
    static std::mutex mGlobalMutex;

    cv::Mat mWorldPos, mDescriptor, mNormalVector;
    float mfMinDistance = 0, mfMaxDistance = 0;
    MapPoint *mpReplaced = nullptr;
    std::map<KeyFrame *, size_t> mObservations;
    int nObs = 1, nNormalUpdates = 0;
    bool mbBad = false;
};
}  // namespace iORB_SLAM
