// S/include/Sim3Solver.h:33-129 for the mock data model (same members, same order, same types -- mvnMaxError1 / 2 are std::vector<size_t>).
#pragma once
#include <vector>
#include <opencv2/core/core.hpp>
#include "KeyFrame.h"

namespace iORB_SLAM
{
class Sim3Solver
{
public:
    Sim3Solver(KeyFrame *pKF1, KeyFrame *pKF2, const std::vector<MapPoint *> &vpMatched12, const bool bFixScale = true);
    void SetRansacParameters(double probability = 0.99, int minInliers = 6, int maxIterations = 300);
    cv::Mat find(std::vector<bool> &vbInliers12, int &nInliers);
    cv::Mat iterate(int nIterations, bool &bNoMore, std::vector<bool> &vbInliers, int &nInliers);
    cv::Mat GetEstimatedRotation();
    cv::Mat GetEstimatedTranslation();
    float GetEstimatedScale();

protected:
    void ComputeCentroid(cv::Mat &P, cv::Mat &Pr, cv::Mat &C);
    void ComputeSim3(cv::Mat &P1, cv::Mat &P2);
    void CheckInliers();
    void Project(const std::vector<cv::Mat> &vP3Dw, std::vector<cv::Mat> &vP2D, cv::Mat Tcw, cv::Mat K);
    void FromCameraToImage(const std::vector<cv::Mat> &vP3Dc, std::vector<cv::Mat> &vP2D, cv::Mat K);

protected:
    KeyFrame *mpKF1; KeyFrame *mpKF2;
    std::vector<cv::Mat> mvX3Dc1, mvX3Dc2;
    std::vector<MapPoint *> mvpMapPoints1, mvpMapPoints2, mvpMatches12;
    std::vector<size_t> mvnIndices1, mvSigmaSquare1, mvSigmaSquare2, mvnMaxError1, mvnMaxError2;
    int N, mN1;
    cv::Mat mR12i, mt12i; float ms12i; cv::Mat mT12i, mT21i;
    std::vector<bool> mvbInliersi; int mnInliersi;
    int mnIterations; std::vector<bool> mvbBestInliers; int mnBestInliers;
    cv::Mat mBestT12, mBestRotation, mBestTranslation; float mBestScale;
    bool mbFixScale;
    std::vector<size_t> mvAllIndices;
    std::vector<cv::Mat> mvP1im1, mvP2im2;
    double mRansacProb; int mRansacMinInliers, mRansacMaxIts;
    float mTh, mSigma2;
    cv::Mat mK1, mK2;
};
}  // namespace iORB_SLAM
