// Sim3Solver_b200.cc -- GPU-backed Sim3Solver::iterate and Sim3Solver::CheckInliers (S/src/Sim3Solver.cc:140-201, 340-365).  Everything else of the
// class (constructor, SetRansacParameters, find, ComputeCentroid, ComputeSim3, Project, FromCameraToImage, the getters) stays in the reference's
// Sim3Solver.cc: compile that file with the two replaced bodies wrapped in #ifndef ORBSLAMM_B200.
//
// The reference checks one hypothesis at a time: ComputeSim3 on three random correspondences, then CheckInliers = two cv::Mat projections per
// correspondence through freshly allocated matrices.  The random triples do not depend on the inlier counts, so iterate() first generates ALL
// hypotheses of the call with the reference's own ComputeSim3 and the reference's own random draws, checks them in ONE device call
// (orbo_sim3_check_inliers) and then walks the counts in order under the reference's update / return rule.  Per hypothesis the result is
// identical to the reference's.  One difference in the random stream: when the call returns at hypothesis j, the draws of the hypotheses after j
// have already been taken from rand() (the reference would take them in its next call).
#include <algorithm>
#include <stdexcept>
#include <string>
#include <vector>
#include "Sim3Solver.h"
#include "Thirdparty/DBoW2/DUtils/Random.h"
#include "orbslamm_b200.h"

namespace iORB_SLAM
{

namespace
{
struct S3TLS {
    orbo_handle *h = nullptr;
    ~S3TLS() { orbo_destroy(h); }
};
orbo_handle *s3_handle()
{
    static thread_local S3TLS tls;       // LoopClosing and MultiMapper run their solvers on different threads
    if (!tls.h && orbo_create(&tls.h, 0) != ORBS_OK) throw std::runtime_error(std::string("orbo_create: ") + orbs_last_error());
    return tls.h;
}
void s3_check(int rc, const char *what) { if (rc != ORBS_OK) throw std::runtime_error(std::string(what) + ": " + orbs_last_error()); }
void flat44(const cv::Mat &T, float *o) { for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) o[4 * r + c] = T.at<float>(r, c); }

// CheckInliers for a batch of hypotheses of one solver; flat views of the solver's per-correspondence members
struct S3Flat {
    std::vector<float> X1, X2, P1, P2; std::vector<int32_t> e1, e2; float K1[4], K2[4]; int N;
    S3Flat(const std::vector<cv::Mat> &vX1, const std::vector<cv::Mat> &vX2, const std::vector<cv::Mat> &vP1, const std::vector<cv::Mat> &vP2,
           const std::vector<size_t> &m1, const std::vector<size_t> &m2, const cv::Mat &mK1, const cv::Mat &mK2)
    {
        N = (int)vX1.size();
        X1.resize(3 * (size_t)N); X2.resize(3 * (size_t)N); P1.resize(2 * (size_t)N); P2.resize(2 * (size_t)N); e1.resize(N); e2.resize(N);
        for (int i = 0; i < N; i++) {
            for (int k = 0; k < 3; k++) { X1[3 * i + k] = vX1[i].at<float>(k); X2[3 * i + k] = vX2[i].at<float>(k); }
            for (int k = 0; k < 2; k++) { P1[2 * i + k] = vP1[i].at<float>(k); P2[2 * i + k] = vP2[i].at<float>(k); }
            e1[i] = (int32_t)m1[i]; e2[i] = (int32_t)m2[i];
        }
        K1[0] = mK1.at<float>(0, 0); K1[1] = mK1.at<float>(1, 1); K1[2] = mK1.at<float>(0, 2); K1[3] = mK1.at<float>(1, 2);
        K2[0] = mK2.at<float>(0, 0); K2[1] = mK2.at<float>(1, 1); K2[2] = mK2.at<float>(0, 2); K2[3] = mK2.at<float>(1, 2);
    }
    void check(int n_hyp, const float *T12, const float *T21, uint8_t *inliers, int32_t *counts)
    {
        s3_check(orbo_sim3_check_inliers(s3_handle(), n_hyp, T12, T21, N, X1.data(), X2.data(), P1.data(), P2.data(), e1.data(), e2.data(), K1, K2, inliers, counts,
                                         ORBS_MEM_HOST), "orbo_sim3_check_inliers");
    }
};
}  // namespace

void Sim3Solver::CheckInliers()
{
    S3Flat F(mvX3Dc1, mvX3Dc2, mvP1im1, mvP2im2, mvnMaxError1, mvnMaxError2, mK1, mK2);
    if (!F.N) { mnInliersi = 0; return; }
    float T12[16], T21[16];
    flat44(mT12i, T12); flat44(mT21i, T21);
    std::vector<uint8_t> in(F.N);
    int32_t n = 0;
    F.check(1, T12, T21, in.data(), &n);
    mnInliersi = n;
    for (int i = 0; i < F.N; i++) mvbInliersi[i] = in[i] != 0;
}

cv::Mat Sim3Solver::iterate(int nIterations, bool &bNoMore, std::vector<bool> &vbInliers, int &nInliers)
{
    bNoMore = false;
    vbInliers = std::vector<bool>(mN1, false);
    nInliers = 0;
    if (N < mRansacMinInliers) { bNoMore = true; return cv::Mat(); }

    // 1. all hypotheses of this call: the reference's own sampling (:160-172, including its `vAvailableIndices[idx] = back()` bookkeeping) and ComputeSim3
    const int nHyp = std::max(0, std::min(nIterations, mRansacMaxIts - mnIterations));
    struct Hyp { cv::Mat T12, T21, R, t; float s; };
    std::vector<Hyp> hyp(nHyp);
    std::vector<float> T12(16 * (size_t)nHyp), T21(16 * (size_t)nHyp);
    std::vector<size_t> vAvailableIndices;
    cv::Mat P3Dc1i(3, 3, CV_32F), P3Dc2i(3, 3, CV_32F);
#ifdef ORBSLAMM_DEVICE_COMPUTE_SIM3
    std::vector<float> minset1(9 * (size_t)nHyp + 1), minset2(9 * (size_t)nHyp + 1);
#endif
    for (int h = 0; h < nHyp; h++) {
        vAvailableIndices = mvAllIndices;
        for (short i = 0; i < 3; ++i) {
            int randi = DUtils::Random::RandomInt(0, vAvailableIndices.size() - 1);
            int idx = vAvailableIndices[randi];
            for (int r = 0; r < 3; r++) { P3Dc1i.at<float>(r, i) = mvX3Dc1[idx].at<float>(r); P3Dc2i.at<float>(r, i) = mvX3Dc2[idx].at<float>(r); }   // copyTo(col(i))
            vAvailableIndices[idx] = vAvailableIndices.back();
            vAvailableIndices.pop_back();
        }
#ifdef ORBSLAMM_DEVICE_COMPUTE_SIM3
        for (int i = 0; i < 3; i++) for (int r = 0; r < 3; r++) { minset1[9 * (size_t)h + 3 * i + r] = P3Dc1i.at<float>(r, i); minset2[9 * (size_t)h + 3 * i + r] = P3Dc2i.at<float>(r, i); }
#else
        ComputeSim3(P3Dc1i, P3Dc2i);
        hyp[h].T12 = mT12i.clone(); hyp[h].T21 = mT21i.clone(); hyp[h].R = mR12i.clone(); hyp[h].t = mt12i.clone(); hyp[h].s = ms12i;
        flat44(mT12i, &T12[16 * (size_t)h]); flat44(mT21i, &T21[16 * (size_t)h]);
#endif
    }
#ifdef ORBSLAMM_DEVICE_COMPUTE_SIM3
    // optional: Horn's closed form for all min sets in one launch (orbo_sim3_compute).  Float-tolerance parity with cv::eigen / cv::Rodrigues, not bit-exactness
    // (include/orbslamm_b200.h), which is why the default build keeps the reference's ComputeSim3 above.
    if (nHyp > 0) {
        std::vector<float> Rts(13 * (size_t)nHyp);
        s3_check(orbo_sim3_compute(s3_handle(), nHyp, minset1.data(), minset2.data(), mbFixScale ? 1 : 0, T12.data(), T21.data(), Rts.data(), ORBS_MEM_HOST), "orbo_sim3_compute");
        for (int h = 0; h < nHyp; h++) {
            hyp[h].T12 = cv::Mat(4, 4, CV_32F); hyp[h].T21 = cv::Mat(4, 4, CV_32F); hyp[h].R = cv::Mat(3, 3, CV_32F); hyp[h].t = cv::Mat(3, 1, CV_32F);
            for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) { hyp[h].T12.at<float>(r, c) = T12[16 * (size_t)h + 4 * r + c]; hyp[h].T21.at<float>(r, c) = T21[16 * (size_t)h + 4 * r + c]; }
            for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) hyp[h].R.at<float>(r, c) = Rts[13 * (size_t)h + 3 * r + c]; hyp[h].t.at<float>(r) = Rts[13 * (size_t)h + 9 + r]; }
            hyp[h].s = Rts[13 * (size_t)h + 12];
        }
    }
#endif

    // 2. CheckInliers of all of them in one device call
    std::vector<uint8_t> in((size_t)nHyp * N + 1);
    std::vector<int32_t> cnt(nHyp + 1, 0);
    if (nHyp > 0) {
        S3Flat F(mvX3Dc1, mvX3Dc2, mvP1im1, mvP2im2, mvnMaxError1, mvnMaxError2, mK1, mK2);
        F.check(nHyp, T12.data(), T21.data(), in.data(), cnt.data());
    }

    // 3. the reference's loop over the hypotheses (:156-194)
    for (int h = 0; h < nHyp; h++) {
        mnIterations++;
        mT12i = hyp[h].T12; mT21i = hyp[h].T21; mR12i = hyp[h].R; mt12i = hyp[h].t; ms12i = hyp[h].s;
        mnInliersi = cnt[h];
        for (int i = 0; i < N; i++) mvbInliersi[i] = in[(size_t)h * N + i] != 0;
        if (mnInliersi >= mnBestInliers) {
            mvbBestInliers = mvbInliersi;
            mnBestInliers = mnInliersi;
            mBestT12 = mT12i.clone();
            mBestRotation = mR12i.clone();
            mBestTranslation = mt12i.clone();
            mBestScale = ms12i;
            if (mnInliersi > mRansacMinInliers) {
                nInliers = mnInliersi;
                for (int i = 0; i < N; i++)
                    if (mvbInliersi[i]) vbInliers[mvnIndices1[i]] = true;
                return mBestT12;
            }
        }
    }
    if (mnIterations >= mRansacMaxIts) bNoMore = true;
    return cv::Mat();
}

}  // namespace iORB_SLAM
