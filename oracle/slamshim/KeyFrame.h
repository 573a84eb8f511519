/* oracle/slamshim/KeyFrame.h -- data-only stand-in for iORB_SLAM::KeyFrame (S/include/KeyFrame.h) with the members the KeyFrame variants
 * of ORBmatcher.cc use.  The grid is the Frame's (KeyFrame.cc:54-60 copies F.mGrid and the cell sizes) while mnMinX.. are INTEGERS
 * (S/include/KeyFrame.h) and GetFeaturesInArea / IsInImage follow KeyFrame.cc:618-662.  TEST INFRASTRUCTURE ONLY. */
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
        std::vector<size_t> vIndices;
        vIndices.reserve(N);
        const int nMinCellX = std::max(0, (int)floor((x - mnMinX - r) * mfGridElementWidthInv));
        if (nMinCellX >= mnGridCols) return vIndices;
        const int nMaxCellX = std::min((int)mnGridCols - 1, (int)ceil((x - mnMinX + r) * mfGridElementWidthInv));
        if (nMaxCellX < 0) return vIndices;
        const int nMinCellY = std::max(0, (int)floor((y - mnMinY - r) * mfGridElementHeightInv));
        if (nMinCellY >= mnGridRows) return vIndices;
        const int nMaxCellY = std::min((int)mnGridRows - 1, (int)ceil((y - mnMinY + r) * mfGridElementHeightInv));
        if (nMaxCellY < 0) return vIndices;
        for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
            for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
                const std::vector<size_t> vCell = mGrid[ix][iy];
                for (size_t j = 0, jend = vCell.size(); j < jend; j++) {
                    const cv::KeyPoint &kpUn = mvKeysUn[vCell[j]];
                    const float distx = kpUn.pt.x - x, disty = kpUn.pt.y - y;
                    if (fabs(distx) < r && fabs(disty) < r) vIndices.push_back(vCell[j]);
                }
            }
        return vIndices;
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
    cv::Mat Tcw, Ow, mK;
    int mnGridCols = 64, mnGridRows = 48;
    float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
    std::vector<std::vector<std::vector<size_t>>> mGrid;
    std::vector<MapPoint *> mvpMapPoints;
};
}
