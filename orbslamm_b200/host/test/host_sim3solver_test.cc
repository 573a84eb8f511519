// host_sim3solver_test.cc -- drives the drop-in Sim3Solver::iterate (one device call per batch of hypotheses) next to a reference-style sequential
// loop (one hypothesis at a time, CheckInliers through the CPU oracle) with the same random seed and the same ComputeSim3, and reports whether they
// agree.  Inputs are written by tests/test_host_shim_gpu.py.  Usage: host_sim3solver_test <dir>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>
#include "Sim3Solver.h"
#include "Thirdparty/DBoW2/DUtils/Random.h"

using namespace iORB_SLAM;

extern "C" void oracle_sim3_check_inliers(int n_hyp, const float *T12, const float *T21, int N, const float *X3Dc1, const float *X3Dc2, const float *P1im1,
                                          const float *P2im2, const int *max_err1, const int *max_err2, const float *K1, const float *K2, unsigned char *inliers,
                                          int *n_inliers);

template <typename T> std::vector<T> rd(const std::string &p)
{
    std::ifstream f(p, std::ios::binary | std::ios::ate);
    if (!f) { fprintf(stderr, "cannot open %s\n", p.c_str()); exit(2); }
    const size_t n = (size_t)f.tellg();
    std::vector<T> v(n / sizeof(T));
    f.seekg(0); f.read((char *)v.data(), n);
    return v;
}

struct Result { std::vector<float> T; int nInliers = 0, iterations = 0, calls = 0, best = 0; std::vector<unsigned char> inl; };

// the reference's iterate() loop (Sim3Solver.cc:140-201), one hypothesis at a time, inliers from the CPU oracle
struct Probe : public Sim3Solver {
    using Sim3Solver::Sim3Solver;
    cv::Mat iterate_sequential(int nIterations, bool &bNoMore, std::vector<bool> &vbInliers, int &nInliers)
    {
        bNoMore = false; vbInliers = std::vector<bool>(mN1, false); nInliers = 0;
        if (N < mRansacMinInliers) { bNoMore = true; return cv::Mat(); }
        std::vector<float> X1(3 * N), X2(3 * N), P1(2 * N), P2(2 * N); std::vector<int> e1(N), e2(N);
        for (int i = 0; i < N; i++) {
            for (int k = 0; k < 3; k++) { X1[3 * i + k] = mvX3Dc1[i].at<float>(k); X2[3 * i + k] = mvX3Dc2[i].at<float>(k); }
            for (int k = 0; k < 2; k++) { P1[2 * i + k] = mvP1im1[i].at<float>(k); P2[2 * i + k] = mvP2im2[i].at<float>(k); }
            e1[i] = (int)mvnMaxError1[i]; e2[i] = (int)mvnMaxError2[i];
        }
        const float K1[4] = {mK1.at<float>(0, 0), mK1.at<float>(1, 1), mK1.at<float>(0, 2), mK1.at<float>(1, 2)};
        const float K2[4] = {mK2.at<float>(0, 0), mK2.at<float>(1, 1), mK2.at<float>(0, 2), mK2.at<float>(1, 2)};
        std::vector<size_t> vAvailableIndices;
        cv::Mat P3Dc1i(3, 3, CV_32F), P3Dc2i(3, 3, CV_32F);
        int nCurrentIterations = 0;
        while (mnIterations < mRansacMaxIts && nCurrentIterations < nIterations) {
            nCurrentIterations++; mnIterations++;
            vAvailableIndices = mvAllIndices;
            for (short i = 0; i < 3; ++i) {
                int randi = DUtils::Random::RandomInt(0, vAvailableIndices.size() - 1);
                int idx = vAvailableIndices[randi];
                for (int r = 0; r < 3; r++) { P3Dc1i.at<float>(r, i) = mvX3Dc1[idx].at<float>(r); P3Dc2i.at<float>(r, i) = mvX3Dc2[idx].at<float>(r); }
                vAvailableIndices[idx] = vAvailableIndices.back();
                vAvailableIndices.pop_back();
            }
            ComputeSim3(P3Dc1i, P3Dc2i);
            float T12[16], T21[16];
            for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) { T12[4 * r + c] = mT12i.at<float>(r, c); T21[4 * r + c] = mT21i.at<float>(r, c); }
            std::vector<unsigned char> in(N);
            oracle_sim3_check_inliers(1, T12, T21, N, X1.data(), X2.data(), P1.data(), P2.data(), e1.data(), e2.data(), K1, K2, in.data(), &mnInliersi);
            for (int i = 0; i < N; i++) mvbInliersi[i] = in[i] != 0;
            if (mnInliersi >= mnBestInliers) {
                mvbBestInliers = mvbInliersi; mnBestInliers = mnInliersi; mBestT12 = mT12i.clone(); mBestRotation = mR12i.clone(); mBestTranslation = mt12i.clone(); mBestScale = ms12i;
                if (mnInliersi > mRansacMinInliers) {
                    nInliers = mnInliersi;
                    for (int i = 0; i < N; i++) if (mvbInliersi[i]) vbInliers[mvnIndices1[i]] = true;
                    return mBestT12;
                }
            }
        }
        if (mnIterations >= mRansacMaxIts) bNoMore = true;
        return cv::Mat();
    }
    Result run(bool device, int minInliers, int chunk)
    {
        SetRansacParameters(0.99, minInliers, 300);
        Result r;
        bool noMore = false;
        std::vector<bool> inl; int n = 0;
        cv::Mat T;
        while (!noMore && T.empty()) { T = device ? iterate(chunk, noMore, inl, n) : iterate_sequential(chunk, noMore, inl, n); r.calls++; }
        r.nInliers = n; r.iterations = mnIterations; r.best = mnBestInliers;
        const cv::Mat &B = T.empty() ? mBestT12 : T;
        for (int i = 0; i < 16; i++) r.T.push_back(B.empty() ? 0.f : B.at<float>(i / 4, i % 4));
        for (size_t i = 0; i < inl.size(); i++) r.inl.push_back(inl[i] ? 1 : 0);
        return r;
    }
};

int main(int argc, char **argv)
{
    if (argc < 2) return 2;
    const std::string d = argv[1];
    const std::vector<float> X1 = rd<float>(d + "/s3r_X1.bin"), X2 = rd<float>(d + "/s3r_X2.bin"), K1 = rd<float>(d + "/s3r_K1.bin"), K2 = rd<float>(d + "/s3r_K2.bin"),
                             ls2 = rd<float>(d + "/s3r_ls2.bin");
    const std::vector<int> o1 = rd<int>(d + "/s3r_oct1.bin"), o2 = rd<int>(d + "/s3r_oct2.bin");
    const int N = (int)o1.size();
    int bad = 0;
    FILE *out = fopen((d + "/out_sim3solver.txt").c_str(), "w");
    for (int scenario = 0; scenario < 3; scenario++) {          // 0: succeeds after a few failed hypotheses (chunks of 5); 1: the same in chunks of 1; 2: the inlier bar is out of reach, every iteration is consumed
        Result res[2];
        for (int device = 0; device < 2; device++) {
            KeyFrame A, B;
            KeyFrame *kfs[2] = {&A, &B};
            std::vector<MapPoint> mp[2];
            for (int s = 0; s < 2; s++) {
                KeyFrame &K = *kfs[s];
                const std::vector<float> &X = s ? X2 : X1, &Kk = s ? K2 : K1; const std::vector<int> &oc = s ? o2 : o1;
                K.N = N; K.mvKeysUn.resize(N); K.mvpMapPoints.assign(N, nullptr); K.mvLevelSigma2 = ls2;
                K.mK = cv::Mat::eye(3, 3, CV_32F);
                K.mK.at<float>(0, 0) = Kk[0]; K.mK.at<float>(1, 1) = Kk[1]; K.mK.at<float>(0, 2) = Kk[2]; K.mK.at<float>(1, 2) = Kk[3];
                mp[s].assign(N, MapPoint());
                for (int i = 0; i < N; i++) {
                    K.mvKeysUn[i].octave = oc[i];
                    mp[s][i].mWorldPos = cv::Mat(3, 1, CV_32F);            // vector::assign copied the header: give every point its own storage
                    for (int k = 0; k < 3; k++) mp[s][i].mWorldPos.at<float>(k) = X[3 * i + k];
                    mp[s][i].mObservations[&K] = i; K.mvpMapPoints[i] = &mp[s][i];
                }
            }
            std::vector<MapPoint *> matched(N);
            for (int i = 0; i < N; i++) matched[i] = &mp[1][i];
            Probe solver(&A, &B, matched, false);
            DUtils::Random::SeedRand(scenario == 2 ? 99 : 1235);
            res[device] = solver.run(device != 0, scenario == 2 ? (4 * N) / 5 + 3 : 20, scenario == 1 ? 1 : 5);
        }
        const bool same = res[0].T == res[1].T && res[0].nInliers == res[1].nInliers && res[0].iterations == res[1].iterations && res[0].calls == res[1].calls &&
                          res[0].best == res[1].best && res[0].inl == res[1].inl;
        if (!same) {
            bad++;
            for (int k = 0; k < 2; k++)
                printf("scenario %d %s: nInliers %d iterations %d calls %d best %d T[3]=%g T[0]=%g ninl_flags %d\n", scenario, k ? "device" : "sequential", res[k].nInliers,
                       res[k].iterations, res[k].calls, res[k].best, res[k].T[3], res[k].T[0], (int)res[k].inl.size());
        }
        fprintf(out, "scenario %d same %d nInliers %d iterations %d calls %d best %d N %d\n", scenario, same ? 1 : 0, res[1].nInliers, res[1].iterations, res[1].calls, res[1].best, N);
    }
    fclose(out);
    printf("sim3solver host shim: %d mismatching scenario(s)\n", bad);
    return bad ? 1 : 0;
}
