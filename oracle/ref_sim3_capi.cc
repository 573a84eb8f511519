/*
 * oracle/ref_sim3_capi.cc -- C entry points around the REFERENCE's own Sim3Solver (/root/reference/SingleRobotScenario/src/Sim3Solver.cc compiled
 * unmodified against oracle/slamshim by oracle/Makefile into oracle/_ref/libref_sim3solver.so).  What runs as the reference's object code and is
 * compared with the oracle / the CUDA path: the constructor's bookkeeping (camera-frame points, integer-truncated error thresholds,
 * Sim3Solver.cc:35-103), FromCameraToImage (:419-437), Project (:392-417) and CheckInliers (:340-365).  ComputeSim3 is NOT pinned (its OpenCV
 * calls -- cv::eigen, cv::Rodrigues, MatExpr scaling -- are double-precision stand-ins in the shim).
 *
 * TEST INFRASTRUCTURE ONLY (tests/test_oracle_vs_reference.py).
 */
#include <cstring>
#include <vector>
#include "Sim3Solver.h"

using namespace iORB_SLAM;

extern "C" {

/* N correspondences (all valid); X1 / X2: the map points in the camera frames of KF1 / KF2 (the keyframes get identity poses, so the constructor's
 * Rcw*X+tcw reproduces them exactly); oct1 / oct2: keypoint octaves; level_sigma2[nlevels]; K1 / K2: fx fy cx cy.
 * n_hyp hypotheses T12 / T21 (row-major 4x4).  Out: inliers u8[n_hyp, N], n_inliers[n_hyp], max_err1 / max_err2 i32[N] (the truncated thresholds),
 * p1im1 / p2im2 f32[N,2] (FromCameraToImage). */
void ref_sim3_check_inliers(int N, const float *X1, const float *X2, const int *oct1, const int *oct2, const float *level_sigma2, int nlevels, const float *K1,
                            const float *K2, int n_hyp, const float *T12, const float *T21, unsigned char *inliers, int *n_inliers, int *max_err1, int *max_err2,
                            float *p1im1, float *p2im2)
{
    KeyFrame A, B;
    KeyFrame *kfs[2] = {&A, &B};
    const float *Ks[2] = {K1, K2}; const int *octs[2] = {oct1, oct2}; const float *Xs[2] = {X1, X2};
    std::vector<MapPoint> mp[2];
    for (int s = 0; s < 2; s++) {
        KeyFrame &K = *kfs[s];
        K.N = N; K.mvKeysUn.resize(N); K.mvpMapPoints.assign(N, (MapPoint *)nullptr);
        K.mvLevelSigma2.assign(level_sigma2, level_sigma2 + nlevels);
        K.Tcw = cv::Mat::eye(4, 4, CV_32F);
        K.mK = cv::Mat::eye(3, 3, CV_32F);
        K.mK.at<float>(0, 0) = Ks[s][0]; K.mK.at<float>(1, 1) = Ks[s][1]; K.mK.at<float>(0, 2) = Ks[s][2]; K.mK.at<float>(1, 2) = Ks[s][3];
        mp[s].assign(N, MapPoint());
        for (int i = 0; i < N; i++) {
            K.mvKeysUn[i].octave = octs[s][i];
            mp[s][i].mWorldPos = cv::Mat(3, 1, CV_32F);
            for (int k = 0; k < 3; k++) mp[s][i].mWorldPos.at<float>(k) = Xs[s][3 * i + k];
            mp[s][i].mObservations[&K] = i;
            K.mvpMapPoints[i] = &mp[s][i];
        }
    }
    std::vector<MapPoint *> matched(N);
    for (int i = 0; i < N; i++) matched[i] = &mp[1][i];
    Sim3Solver solver(&A, &B, matched, false);
    for (int i = 0; i < N; i++) {
        max_err1[i] = (int)solver.mvnMaxError1[i]; max_err2[i] = (int)solver.mvnMaxError2[i];
        p1im1[2 * i] = solver.mvP1im1[i].at<float>(0); p1im1[2 * i + 1] = solver.mvP1im1[i].at<float>(1);
        p2im2[2 * i] = solver.mvP2im2[i].at<float>(0); p2im2[2 * i + 1] = solver.mvP2im2[i].at<float>(1);
    }
    for (int h = 0; h < n_hyp; h++) {
        solver.mT12i = cv::Mat(4, 4, CV_32F); solver.mT21i = cv::Mat(4, 4, CV_32F);
        std::memcpy(solver.mT12i.data, T12 + 16 * h, 64); std::memcpy(solver.mT21i.data, T21 + 16 * h, 64);
        solver.CheckInliers();
        n_inliers[h] = solver.mnInliersi;
        for (int i = 0; i < N; i++) inliers[(size_t)h * N + i] = solver.mvbInliersi[i] ? 1 : 0;
    }
}

}
