/* oracle/slamshim/KeyFrame.h -- data-only stand-in for iORB_SLAM::KeyFrame (S/include/KeyFrame.h) so that the KeyFrame variants in
 * ORBmatcher.cc compile; they are not exercised by the parity tests.  TEST INFRASTRUCTURE ONLY. */
#pragma once
#include <set>
#include "MapPoint.h"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"
namespace iORB_SLAM {
class KeyFrame {
public:
    KeyFrame() : mnId(0), N(0), fx(0), fy(0), cx(0), cy(0), mbf(0), mb(0), mnScaleLevels(8), mfLogScaleFactor(0), mnMinX(0), mnMinY(0), mnMaxX(0), mnMaxY(0) {}
    std::vector<MapPoint *> GetMapPointMatches() { return mvpMapPoints; }
    std::set<MapPoint *> GetMapPoints() { std::set<MapPoint *> s; for (MapPoint *p : mvpMapPoints) if (p && !p->isBad()) s.insert(p); return s; }
    MapPoint *GetMapPoint(const size_t &idx) { return mvpMapPoints[idx]; }
    void AddMapPoint(MapPoint *pMP, const size_t &idx) { mvpMapPoints[idx] = pMP; }
    cv::Mat GetRotation() { return Tcw.rowRange(0, 3).colRange(0, 3).clone(); }
    cv::Mat GetTranslation() { return Tcw.rowRange(0, 3).col(3).clone(); }
    cv::Mat GetCameraCenter() { return Ow.clone(); }
    bool IsInImage(const float &x, const float &y) const { return (x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY); }
    std::vector<size_t> GetFeaturesInArea(const float &x, const float &y, const float &r) const
    {
        std::vector<size_t> v;
        for (size_t i = 0; i < mvKeysUn.size(); i++) if (std::fabs(mvKeysUn[i].pt.x - x) < r && std::fabs(mvKeysUn[i].pt.y - y) < r) v.push_back(i);
        return v;
    }
    long unsigned int mnId;
    int N;
    float fx, fy, cx, cy, mbf, mb;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    std::vector<float> mvuRight, mvDepth;
    cv::Mat mDescriptors;
    DBoW2::FeatureVector mFeatVec;
    int mnScaleLevels; float mfLogScaleFactor;
    std::vector<float> mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
    int mnMinX, mnMinY, mnMaxX, mnMaxY;
    cv::Mat Tcw, Ow;
    std::vector<MapPoint *> mvpMapPoints;
};
}
