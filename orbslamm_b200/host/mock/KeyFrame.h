#pragma once
#include <algorithm>
#include <map>
#include <set>
#include "MapPoint.h"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"

namespace iORB_SLAM
{
class KeyFrame
{
public:
    KeyFrame() : Tcw(cv::Mat::eye(4, 4, CV_32F)) {}
    cv::Mat GetPose() { return Tcw.clone(); }
    void SetPose(const cv::Mat &T) { Tcw = T.clone(); }
    bool isBad() { return mbBad; }
    KeyFrame *GetParent() { return mpParent; }
    bool hasChild(KeyFrame *pKF) { return mspChildrens.count(pKF) != 0; }
    std::set<KeyFrame *> GetLoopEdges() { return mspLoopEdges; }
    int GetWeight(KeyFrame *pKF) { return mConnectedKeyFrameWeights.count(pKF) ? mConnectedKeyFrameWeights[pKF] : 0; }
    std::vector<KeyFrame *> GetCovisiblesByWeight(const int &w)     // KeyFrame.cc:208-225: ordered by weight, descending
    {
        std::vector<std::pair<int, KeyFrame *>> v;
        for (auto &kv : mConnectedKeyFrameWeights) if (kv.second >= w) v.push_back(std::make_pair(kv.second, kv.first));
        std::stable_sort(v.begin(), v.end(), [](const std::pair<int, KeyFrame *> &a, const std::pair<int, KeyFrame *> &b) { return a.first > b.first; });
        std::vector<KeyFrame *> out;
        for (auto &x : v) out.push_back(x.second);
        return out;
    }
    bool isNotFixed() { return false; }                          // M/include/KeyFrame.h:113-115 (multi-robot tree)
    std::vector<KeyFrame *> GetVectorCovisibleKeyFrames() { return mvpOrderedConnectedKeyFrames; }
    std::vector<MapPoint *> GetMapPointMatches() { return mvpMapPoints; }
    void EraseMapPointMatch(MapPoint *pMP) { for (auto &p : mvpMapPoints) if (p == pMP) p = nullptr; }
    void EraseMapPointMatch(const size_t &idx) { mvpMapPoints[idx] = nullptr; }
    void ReplaceMapPointMatch(const size_t &idx, MapPoint *pMP) { mvpMapPoints[idx] = pMP; }
    MapPoint *GetMapPoint(const size_t &idx) { return mvpMapPoints[idx]; }
    void AddMapPoint(MapPoint *pMP, const size_t &idx) { mvpMapPoints[idx] = pMP; }
    std::set<MapPoint *> GetMapPoints() { std::set<MapPoint *> s; for (MapPoint *p : mvpMapPoints) if (p && !p->isBad()) s.insert(p); return s; }
    cv::Mat GetRotation() { cv::Mat R(3, 3, CV_32F); for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) R.at<float>(r, c) = Tcw.at<float>(r, c); return R; }
    cv::Mat GetTranslation() { cv::Mat t(3, 1, CV_32F); for (int r = 0; r < 3; r++) t.at<float>(r) = Tcw.at<float>(r, 3); return t; }
    cv::Mat GetCameraCenter()          // Ow = -Rcw^T tcw (KeyFrame::SetPose, KeyFrame.cc:72-76), OpenCV small-gemm order
    {
        cv::Mat O(3, 1, CV_32F);
        for (int r = 0; r < 3; r++) {
            float s = -Tcw.at<float>(0, r) * Tcw.at<float>(0, 3);
            s = s + -Tcw.at<float>(1, r) * Tcw.at<float>(1, 3);
            s = s + -Tcw.at<float>(2, r) * Tcw.at<float>(2, 3);
            O.at<float>(r) = s;
        }
        return O;
    }

    long unsigned int mnId = 0, mnBALocalForKF = ~0ul, mnBAFixedForKF = ~0ul, mnBAGlobalForKF = 0;
    float fx = 0, fy = 0, cx = 0, cy = 0, mbf = 0;
    int N = 0;
    int mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0;           // ints in S/include/KeyFrame.h
    float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0, mfLogScaleFactor = 0;
    std::vector<float> mvScaleFactors, mvLevelSigma2;
    DBoW2::FeatureVector mFeatVec;
    cv::Mat mDescriptors;
    std::vector<cv::KeyPoint> mvKeysUn;
    std::vector<float> mvuRight, mvInvLevelSigma2;
    std::vector<MapPoint *> mvpMapPoints;
    std::vector<KeyFrame *> mvpOrderedConnectedKeyFrames;
    KeyFrame *mpParent = nullptr;
    std::set<KeyFrame *> mspChildrens, mspLoopEdges;
    std::map<KeyFrame *, int> mConnectedKeyFrameWeights;
    cv::Mat Tcw, mTcwGBA, mK;
    bool mbBad = false;
};
}  // namespace iORB_SLAM
