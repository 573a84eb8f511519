/* oracle/slamshim/Frame.h -- stand-in for iORB_SLAM::Frame (S/include/Frame.h) with the members ORBmatcher.cc reads.  The feature
 * grid follows Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea (S/src/Frame.cc:230-245, 327-392).
 * TEST INFRASTRUCTURE ONLY. */
#pragma once
#include "KeyFrame.h"
#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64
namespace iORB_SLAM {
class Frame {
public:
    Frame() : N(0), mb(0), mbf(0), mfLogScaleFactor(0), mnId(0) {}
    bool PosInGrid(const cv::KeyPoint &kp, int &posX, int &posY)
    {
        posX = round((kp.pt.x - mnMinX) * mfGridElementWidthInv);
        posY = round((kp.pt.y - mnMinY) * mfGridElementHeightInv);
        if (posX < 0 || posX >= FRAME_GRID_COLS || posY < 0 || posY >= FRAME_GRID_ROWS) return false;
        return true;
    }
    void AssignFeaturesToGrid()
    {
        for (int i = 0; i < FRAME_GRID_COLS; i++) for (int j = 0; j < FRAME_GRID_ROWS; j++) mGrid[i][j].clear();
        for (int i = 0; i < N; i++) { int x, y; if (PosInGrid(mvKeysUn[i], x, y)) mGrid[x][y].push_back(i); }
    }
    std::vector<size_t> GetFeaturesInArea(const float &x, const float &y, const float &r, const int minLevel = -1, const int maxLevel = -1) const
    {
        std::vector<size_t> vIndices;
        vIndices.reserve(N);
        const int nMinCellX = std::max(0, (int)floor((x - mnMinX - r) * mfGridElementWidthInv));
        if (nMinCellX >= FRAME_GRID_COLS) return vIndices;
        const int nMaxCellX = std::min((int)FRAME_GRID_COLS - 1, (int)ceil((x - mnMinX + r) * mfGridElementWidthInv));
        if (nMaxCellX < 0) return vIndices;
        const int nMinCellY = std::max(0, (int)floor((y - mnMinY - r) * mfGridElementHeightInv));
        if (nMinCellY >= FRAME_GRID_ROWS) return vIndices;
        const int nMaxCellY = std::min((int)FRAME_GRID_ROWS - 1, (int)ceil((y - mnMinY + r) * mfGridElementHeightInv));
        if (nMaxCellY < 0) return vIndices;
        const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
        for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
            for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
                const std::vector<size_t> vCell = mGrid[ix][iy];
                if (vCell.empty()) continue;
                for (size_t j = 0, jend = vCell.size(); j < jend; j++) {
                    const cv::KeyPoint &kpUn = mvKeysUn[vCell[j]];
                    if (bCheckLevels) {
                        if (kpUn.octave < minLevel) continue;
                        if (maxLevel >= 0) if (kpUn.octave > maxLevel) continue;
                    }
                    const float distx = kpUn.pt.x - x, disty = kpUn.pt.y - y;
                    if (fabs(distx) < r && fabs(disty) < r) vIndices.push_back(vCell[j]);
                }
            }
        return vIndices;
    }
    int N;
    float mb, mbf;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    std::vector<float> mvuRight, mvDepth;
    cv::Mat mDescriptors;
    DBoW2::FeatureVector mFeatVec;
    std::vector<MapPoint *> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    cv::Mat mTcw;
    float mfLogScaleFactor;
    std::vector<float> mvScaleFactors, mvInvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
    long unsigned int mnId;
    static float fx, fy, cx, cy, invfx, invfy;
    static float mnMinX, mnMaxX, mnMinY, mnMaxY;
    static float mfGridElementWidthInv, mfGridElementHeightInv;
    std::vector<size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];
};
}
