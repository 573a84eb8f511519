// Test scaffolding: the members of Sim3Solver that STAY in the reference's Sim3Solver.cc in a real build (constructor :35-103, SetRansacParameters
// :105-137, find :204-208, the getters, FromCameraToImage :419-437), written for the mock data model (whose cv::Mat has no algebra), plus a
// double-precision Horn solve standing in for ComputeSim3 (:226-338; the real one is OpenCV float code and is not what this repo replaces).
#include <algorithm>
#include <cmath>
#include "Sim3Solver.h"

namespace iORB_SLAM
{
static cv::Mat col3(float a, float b, float c) { cv::Mat m(3, 1, CV_32F); m.at<float>(0) = a; m.at<float>(1) = b; m.at<float>(2) = c; return m; }

Sim3Solver::Sim3Solver(KeyFrame *pKF1, KeyFrame *pKF2, const std::vector<MapPoint *> &vpMatched12, const bool bFixScale) : mnIterations(0), mnBestInliers(0), mbFixScale(bFixScale)
{
    mpKF1 = pKF1; mpKF2 = pKF2;
    std::vector<MapPoint *> vpKeyFrameMP1 = pKF1->GetMapPointMatches();
    mN1 = (int)vpMatched12.size();
    mvpMatches12 = vpMatched12;
    const cv::Mat T1 = pKF1->GetPose(), T2 = pKF2->GetPose();
    auto to_cam = [](const cv::Mat &T, const cv::Mat &X) {                 // Rcw*X3Dw + tcw, OpenCV small-gemm order
        float o[3];
        for (int r = 0; r < 3; r++) { float s = T.at<float>(r, 0) * X.at<float>(0); s = s + T.at<float>(r, 1) * X.at<float>(1); s = s + T.at<float>(r, 2) * X.at<float>(2); o[r] = s + T.at<float>(r, 3); }
        return col3(o[0], o[1], o[2]);
    };
    size_t idx = 0;
    for (int i1 = 0; i1 < mN1; i1++) {
        if (!vpMatched12[i1]) continue;
        MapPoint *pMP1 = vpKeyFrameMP1[i1], *pMP2 = vpMatched12[i1];
        if (!pMP1) continue;
        if (pMP1->isBad() || pMP2->isBad()) continue;
        const int indexKF1 = pMP1->GetIndexInKeyFrame(pKF1), indexKF2 = pMP2->GetIndexInKeyFrame(pKF2);
        if (indexKF1 < 0 || indexKF2 < 0) continue;
        const cv::KeyPoint &kp1 = pKF1->mvKeysUn[indexKF1], &kp2 = pKF2->mvKeysUn[indexKF2];
        const float sigmaSquare1 = pKF1->mvLevelSigma2[kp1.octave], sigmaSquare2 = pKF2->mvLevelSigma2[kp2.octave];
        mvnMaxError1.push_back(9.210 * sigmaSquare1);
        mvnMaxError2.push_back(9.210 * sigmaSquare2);
        mvpMapPoints1.push_back(pMP1); mvpMapPoints2.push_back(pMP2); mvnIndices1.push_back(i1);
        mvX3Dc1.push_back(to_cam(T1, pMP1->GetWorldPos())); mvX3Dc2.push_back(to_cam(T2, pMP2->GetWorldPos()));
        mvAllIndices.push_back(idx); idx++;
    }
    mK1 = pKF1->mK; mK2 = pKF2->mK;
    FromCameraToImage(mvX3Dc1, mvP1im1, mK1);
    FromCameraToImage(mvX3Dc2, mvP2im2, mK2);
    SetRansacParameters();
}

void Sim3Solver::SetRansacParameters(double probability, int minInliers, int maxIterations)
{
    mRansacProb = probability; mRansacMinInliers = minInliers; mRansacMaxIts = maxIterations;
    N = (int)mvpMapPoints1.size();
    mvbInliersi.resize(N);
    float epsilon = (float)mRansacMinInliers / N;
    int nIterations;
    if (mRansacMinInliers == N) nIterations = 1;
    else nIterations = std::ceil(std::log(1 - mRansacProb) / std::log(1 - std::pow(epsilon, 3)));
    mRansacMaxIts = std::max(1, std::min(nIterations, mRansacMaxIts));
    mnIterations = 0;
}

cv::Mat Sim3Solver::find(std::vector<bool> &vbInliers12, int &nInliers) { bool bFlag; return iterate(mRansacMaxIts, bFlag, vbInliers12, nInliers); }
cv::Mat Sim3Solver::GetEstimatedRotation() { return mBestRotation.clone(); }
cv::Mat Sim3Solver::GetEstimatedTranslation() { return mBestTranslation.clone(); }
float Sim3Solver::GetEstimatedScale() { return mBestScale; }

void Sim3Solver::FromCameraToImage(const std::vector<cv::Mat> &vP3Dc, std::vector<cv::Mat> &vP2D, cv::Mat K)
{
    const float &fx = K.at<float>(0, 0), &fy = K.at<float>(1, 1), &cx = K.at<float>(0, 2), &cy = K.at<float>(1, 2);
    vP2D.clear();
    for (size_t i = 0; i < vP3Dc.size(); i++) {
        const float invz = 1 / (vP3Dc[i].at<float>(2));
        const float x = vP3Dc[i].at<float>(0) * invz, y = vP3Dc[i].at<float>(1) * invz;
        cv::Mat p(2, 1, CV_32F); p.at<float>(0) = fx * x + cx; p.at<float>(1) = fy * y + cy;
        vP2D.push_back(p);
    }
}
void Sim3Solver::ComputeCentroid(cv::Mat &, cv::Mat &, cv::Mat &) {}
void Sim3Solver::Project(const std::vector<cv::Mat> &, std::vector<cv::Mat> &, cv::Mat, cv::Mat) {}

// Horn's closed form in double (power iteration for the dominant eigenvector of N); fills mR12i, mt12i, ms12i, mT12i, mT21i
void Sim3Solver::ComputeSim3(cv::Mat &P1, cv::Mat &P2)
{
    double O1[3] = {0, 0, 0}, O2[3] = {0, 0, 0}, A[3][3], B[3][3];
    for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) { O1[r] += P1.at<float>(r, c) / 3.0; O2[r] += P2.at<float>(r, c) / 3.0; } }
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { A[r][c] = P1.at<float>(r, c) - O1[r]; B[r][c] = P2.at<float>(r, c) - O2[r]; }
    double M[3][3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { M[i][j] = 0; for (int k = 0; k < 3; k++) M[i][j] += B[i][k] * A[j][k]; }
    double Nm[4][4] = {{M[0][0] + M[1][1] + M[2][2], M[1][2] - M[2][1], M[2][0] - M[0][2], M[0][1] - M[1][0]},
                       {0, M[0][0] - M[1][1] - M[2][2], M[0][1] + M[1][0], M[2][0] + M[0][2]},
                       {0, 0, -M[0][0] + M[1][1] - M[2][2], M[1][2] + M[2][1]},
                       {0, 0, 0, -M[0][0] - M[1][1] + M[2][2]}};
    for (int i = 0; i < 4; i++) for (int j = 0; j < i; j++) Nm[i][j] = Nm[j][i];
    double tr = 0; for (int i = 0; i < 4; i++) tr += std::fabs(Nm[i][i]);
    double q[4] = {1, 0.1, 0.1, 0.1};
    for (int it = 0; it < 200; it++) {                                      // power iteration on N + shift (dominant = largest eigenvalue)
        double y[4];
        for (int i = 0; i < 4; i++) { y[i] = (tr + 1e-9) * q[i]; for (int j = 0; j < 4; j++) y[i] += Nm[i][j] * q[j]; }
        const double n = std::sqrt(y[0] * y[0] + y[1] * y[1] + y[2] * y[2] + y[3] * y[3]);
        for (int i = 0; i < 4; i++) q[i] = y[i] / n;
    }
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    const double R[3][3] = {{1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)}, {2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)},
                            {2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)}};
    double nom = 0, den = 0, P3[3][3];
    for (int i = 0; i < 3; i++) for (int c = 0; c < 3; c++) { P3[i][c] = 0; for (int k = 0; k < 3; k++) P3[i][c] += R[i][k] * B[k][c]; nom += A[i][c] * P3[i][c]; den += P3[i][c] * P3[i][c]; }
    ms12i = mbFixScale ? 1.0f : (float)(nom / den);
    mR12i = cv::Mat(3, 3, CV_32F); mt12i = cv::Mat(3, 1, CV_32F);
    mT12i = cv::Mat::eye(4, 4, CV_32F); mT21i = cv::Mat::eye(4, 4, CV_32F);
    double t[3];
    for (int i = 0; i < 3; i++) { t[i] = O1[i]; for (int k = 0; k < 3; k++) t[i] -= ms12i * R[i][k] * O2[k]; }
    for (int i = 0; i < 3; i++) {
        mt12i.at<float>(i) = (float)t[i]; mT12i.at<float>(i, 3) = (float)t[i];
        double ti = 0;
        for (int k = 0; k < 3; k++) { mR12i.at<float>(i, k) = (float)R[i][k]; mT12i.at<float>(i, k) = (float)(ms12i * R[i][k]); mT21i.at<float>(i, k) = (float)(R[k][i] / ms12i); ti -= R[k][i] / ms12i * t[k]; }
        mT21i.at<float>(i, 3) = (float)ti;
    }
}
}  // namespace iORB_SLAM
