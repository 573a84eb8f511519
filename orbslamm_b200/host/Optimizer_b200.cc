// Optimizer_b200.cc -- GPU-backed definitions of Optimizer::PoseOptimization / LocalBundleAdjustment /
// BundleAdjustment / GlobalBundleAdjustemnt.  In the reference tree these replace the same-named functions of
// S/src/Optimizer.cc (:60-65, :68-260, :262-474, :476-801) and OptimizeSim3 (:1348-1543); the essential-graph optimisers stay on
// g2o ("next" rows of SURVEY.md 8(f)).  The graph *construction* below follows the reference line by line (which
// keyframes are local / fixed, which observations become edges, which locks are taken); only the numerical solve is
// delegated to orbo_* (include/orbslamm_b200.h).  Monocular observations only (mvuRight < 0).
#include <list>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>
#include "Optimizer.h"
#include "orbslamm_b200.h"

namespace iORB_SLAM
{

namespace
{
struct OptTLS {
    orbo_handle *h = nullptr;
    ~OptTLS() { orbo_destroy(h); }
};
orbo_handle *handle()
{
    static thread_local OptTLS tls;       // Optimizer statics are called concurrently from 4 threads
    if (!tls.h && orbo_create(&tls.h, 0) != ORBS_OK) throw std::runtime_error(std::string("orbo_create: ") + orbs_last_error());
    return tls.h;
}
void check(int rc, const char *what) { if (rc < 0) throw std::runtime_error(std::string(what) + ": " + orbs_last_error()); }
void pose_to_flat(const cv::Mat &T, float *o) { for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) o[4 * r + c] = T.at<float>(r, c); }
cv::Mat flat_to_pose(const float *p) { cv::Mat T(4, 4, CV_32F); for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) T.at<float>(r, c) = p[4 * r + c]; return T; }
}  // namespace

int Optimizer::PoseOptimization(Frame *pFrame)
{
    const int N = pFrame->N;
    std::vector<float> Xw, obs, w;
    std::vector<size_t> index;
    Xw.reserve(3 * (size_t)N); obs.reserve(2 * (size_t)N); w.reserve(N); index.reserve(N);
    {
        std::unique_lock<std::mutex> lock(MapPoint::mGlobalMutex);          // Optimizer.cc:301
        for (int i = 0; i < N; i++) {
            MapPoint *pMP = pFrame->mvpMapPoints[i];
            if (!pMP) continue;
            if (!(pFrame->mvuRight[i] < 0)) throw std::runtime_error("orbslamm_b200: stereo observations are not supported");
            pFrame->mvbOutlier[i] = false;                                   // Optimizer.cc:309
            const cv::KeyPoint &kpUn = pFrame->mvKeysUn[i];
            obs.push_back(kpUn.pt.x); obs.push_back(kpUn.pt.y);
            w.push_back(pFrame->mvInvLevelSigma2[kpUn.octave]);
            cv::Mat X = pMP->GetWorldPos();
            Xw.push_back(X.at<float>(0)); Xw.push_back(X.at<float>(1)); Xw.push_back(X.at<float>(2));
            index.push_back(i);
        }
    }
    const int32_t M = (int32_t)index.size();
    if (M < 3) return 0;                                                     // Optimizer.cc:387-388
    float Tcw[16];
    pose_to_flat(pFrame->mTcw, Tcw);
    const float K4[4] = {Frame::fx, Frame::fy, Frame::cx, Frame::cy};
    std::vector<uint8_t> outlier(M);
    int32_t nInliers = 0;
    check(orbo_pose_optimization(handle(), 1, Tcw, K4, Xw.data(), obs.data(), w.data(), &M, M, outlier.data(), &nInliers, ORBS_MEM_HOST),
          "orbo_pose_optimization");
    for (int32_t e = 0; e < M; e++) pFrame->mvbOutlier[index[e]] = outlier[e] != 0;
    pFrame->SetPose(flat_to_pose(Tcw));                                      // Optimizer.cc:468-471
    return nInliers;
}

void Optimizer::GlobalBundleAdjustemnt(Map *pMap, int nIterations, bool *pbStopFlag, const unsigned long nLoopKF, const bool bRobust)
{
    std::vector<KeyFrame *> vpKFs = pMap->GetAllKeyFrames();
    std::vector<MapPoint *> vpMP = pMap->GetAllMapPoints();
    BundleAdjustment(vpKFs, vpMP, nIterations, pbStopFlag, nLoopKF, bRobust);
}

#ifdef ORBSLAMM_MULTI_ROBOT
// MultipleRobotsScenario/src/Optimizer.cc:40-57: the map plus every map attached to it by MultiMapper, one bundle adjustment over the union
void Optimizer::MMGlobalBundleAdjustemnt(Map *pMap, int nIterations, bool *pbStopFlag, const unsigned long nLoopKF, const bool bRobust)
{
    std::vector<KeyFrame *> vpKFs = pMap->GetAllKeyFrames();
    std::vector<MapPoint *> vpMPs = pMap->GetAllMapPoints();
    if (pMap->isAttached()) {
        std::vector<Map *> vpAttachedMaps = pMap->getAttachedMaps();
        for (std::vector<Map *>::iterator it = vpAttachedMaps.begin(), itend = vpAttachedMaps.end(); it != itend; it++) {
            Map *pmMap = *it;
            std::vector<KeyFrame *> vpAttachedKFs = pmMap->GetAllKeyFrames();
            std::vector<MapPoint *> vpAttachedMPs = pmMap->GetAllMapPoints();
            vpKFs.insert(vpKFs.end(), vpAttachedKFs.begin(), vpAttachedKFs.end());
            vpMPs.insert(vpMPs.end(), vpAttachedMPs.begin(), vpAttachedMPs.end());
        }
    }
    BundleAdjustment(vpKFs, vpMPs, nIterations, pbStopFlag, nLoopKF, bRobust);
}
#endif

namespace
{
// bool* stop flag of the reference -> the int flag the C-ABI polls
struct StopBridge {
    bool *src; volatile int flag = 0;
    explicit StopBridge(bool *p) : src(p) { if (src && *src) flag = 1; }
};
}  // namespace

void Optimizer::BundleAdjustment(const std::vector<KeyFrame *> &vpKFs, const std::vector<MapPoint *> &vpMP, int nIterations, bool *pbStopFlag,
                                 const unsigned long nLoopKF, const bool bRobust)
{
    std::map<KeyFrame *, int> kfIndex;
    std::vector<KeyFrame *> kfs;
    std::vector<float> poses;
    std::vector<uint8_t> fixed;
    std::vector<double> intr;
    for (size_t i = 0; i < vpKFs.size(); i++) {                              // Optimizer.cc:89-104
        KeyFrame *pKF = vpKFs[i];
        if (pKF->isBad()) continue;
        kfIndex[pKF] = (int)kfs.size(); kfs.push_back(pKF);
        poses.resize(poses.size() + 16);
        pose_to_flat(pKF->GetPose(), &poses[poses.size() - 16]);
#ifdef ORBSLAMM_MULTI_ROBOT
        fixed.push_back((pKF->mnId == 0 && !pKF->isNotFixed()) ? 1 : 0);        // M/src/Optimizer.cc:99: the first keyframe of an attached map is free
#else
        fixed.push_back(pKF->mnId == 0 ? 1 : 0);                                 // S/src/Optimizer.cc:100
#endif
        intr.push_back(pKF->fx); intr.push_back(pKF->fy); intr.push_back(pKF->cx); intr.push_back(pKF->cy);
    }
    std::vector<MapPoint *> mps;
    std::vector<float> points, uv, w;
    std::vector<int32_t> ekf, ept;
    for (size_t i = 0; i < vpMP.size(); i++) {                               // Optimizer.cc:109-178
        MapPoint *pMP = vpMP[i];
        if (pMP->isBad()) continue;
        const int pi = (int)mps.size();
        int nEdges = 0;
        const std::map<KeyFrame *, size_t> observations = pMP->GetObservations();
        for (std::map<KeyFrame *, size_t>::const_iterator mit = observations.begin(); mit != observations.end(); mit++) {
            KeyFrame *pKF = mit->first;
            std::map<KeyFrame *, int>::iterator it = kfIndex.find(pKF);
            if (pKF->isBad() || it == kfIndex.end()) continue;
            const cv::KeyPoint &kpUn = pKF->mvKeysUn[mit->second];
            if (!(pKF->mvuRight[mit->second] < 0)) throw std::runtime_error("orbslamm_b200: stereo observations are not supported");
            ekf.push_back(it->second); ept.push_back(pi);
            uv.push_back(kpUn.pt.x); uv.push_back(kpUn.pt.y); w.push_back(pKF->mvInvLevelSigma2[kpUn.octave]);
            nEdges++;
        }
        if (nEdges == 0) continue;                                           // vbNotIncludedMP, Optimizer.cc:180-189
        mps.push_back(pMP);
        cv::Mat X = pMP->GetWorldPos();
        points.push_back(X.at<float>(0)); points.push_back(X.at<float>(1)); points.push_back(X.at<float>(2));
    }
    if (kfs.empty() || mps.empty() || ekf.empty()) return;
    StopBridge stop(pbStopFlag);
    const int rc = orbo_bundle_adjust(handle(), (int)kfs.size(), poses.data(), fixed.data(), intr.data(), (int)mps.size(), points.data(), (int)ekf.size(),
                                      ekf.data(), ept.data(), uv.data(), w.data(), 0, nIterations, 0, bRobust ? 1 : 0, pbStopFlag ? &stop.flag : nullptr,
                                      nullptr, nullptr, nullptr, nullptr);
    check(rc, "orbo_bundle_adjust");
    for (size_t i = 0; i < kfs.size(); i++) {                                // Optimizer.cc:196-231
        cv::Mat T = flat_to_pose(&poses[16 * i]);
        if (nLoopKF == 0) kfs[i]->SetPose(T);
        else { kfs[i]->mTcwGBA = T; kfs[i]->mnBAGlobalForKF = nLoopKF; }
    }
    for (size_t i = 0; i < mps.size(); i++) {                                // Optimizer.cc:234-258
        cv::Mat X(3, 1, CV_32F);
        for (int c = 0; c < 3; c++) X.at<float>(c) = points[3 * i + c];
        if (nLoopKF == 0) { mps[i]->SetWorldPos(X); mps[i]->UpdateNormalAndDepth(); }
        else { mps[i]->mPosGBA = X; mps[i]->mnBAGlobalForKF = nLoopKF; }
    }
}

void Optimizer::LocalBundleAdjustment(KeyFrame *pKF, bool *pbStopFlag, Map *pMap)
{
    // Local KeyFrames: first breadth search from the current keyframe (Optimizer.cc:479-491)
    std::list<KeyFrame *> lLocalKeyFrames;
    lLocalKeyFrames.push_back(pKF);
    pKF->mnBALocalForKF = pKF->mnId;
    const std::vector<KeyFrame *> vNeighKFs = pKF->GetVectorCovisibleKeyFrames();
    for (size_t i = 0; i < vNeighKFs.size(); i++) {
        KeyFrame *pKFi = vNeighKFs[i];
        pKFi->mnBALocalForKF = pKF->mnId;
        if (!pKFi->isBad()) lLocalKeyFrames.push_back(pKFi);
    }
    // Local MapPoints seen in Local KeyFrames (Optimizer.cc:494-509)
    std::list<MapPoint *> lLocalMapPoints;
    for (std::list<KeyFrame *>::iterator lit = lLocalKeyFrames.begin(); lit != lLocalKeyFrames.end(); lit++) {
        std::vector<MapPoint *> vpMPs = (*lit)->GetMapPointMatches();
        for (size_t k = 0; k < vpMPs.size(); k++) {
            MapPoint *pMP = vpMPs[k];
            if (pMP && !pMP->isBad() && pMP->mnBALocalForKF != pKF->mnId) { lLocalMapPoints.push_back(pMP); pMP->mnBALocalForKF = pKF->mnId; }
        }
    }
    // Fixed KeyFrames: see local points but are not local (Optimizer.cc:512-527)
    std::list<KeyFrame *> lFixedCameras;
    for (std::list<MapPoint *>::iterator lit = lLocalMapPoints.begin(); lit != lLocalMapPoints.end(); lit++) {
        std::map<KeyFrame *, size_t> observations = (*lit)->GetObservations();
        for (std::map<KeyFrame *, size_t>::iterator mit = observations.begin(); mit != observations.end(); mit++) {
            KeyFrame *pKFi = mit->first;
            if (pKFi->mnBALocalForKF != pKF->mnId && pKFi->mnBAFixedForKF != pKF->mnId) {
                pKFi->mnBAFixedForKF = pKF->mnId;
                if (!pKFi->isBad()) lFixedCameras.push_back(pKFi);
            }
        }
    }
    // flat graph
    std::map<KeyFrame *, int> kfIndex;
    std::vector<KeyFrame *> kfs;
    std::vector<float> poses;
    std::vector<uint8_t> fixed;
    std::vector<double> intr;
    auto add_kf = [&](KeyFrame *k, uint8_t fx) {
        kfIndex[k] = (int)kfs.size(); kfs.push_back(k);
        poses.resize(poses.size() + 16);
        pose_to_flat(k->GetPose(), &poses[poses.size() - 16]);
        fixed.push_back(fx);
        intr.push_back(k->fx); intr.push_back(k->fy); intr.push_back(k->cx); intr.push_back(k->cy);
    };
    for (std::list<KeyFrame *>::iterator lit = lLocalKeyFrames.begin(); lit != lLocalKeyFrames.end(); lit++)
        add_kf(*lit, (*lit)->mnId == 0 ? 1 : 0);                             // setFixed(pKFi->mnId==0), Optimizer.cc:553
    for (std::list<KeyFrame *>::iterator lit = lFixedCameras.begin(); lit != lFixedCameras.end(); lit++) add_kf(*lit, 2);
    std::vector<MapPoint *> mps;
    std::vector<float> points, uv, w;
    std::vector<int32_t> ekf, ept;
    std::vector<KeyFrame *> vpEdgeKF;
    std::vector<MapPoint *> vpEdgeMP;
    for (std::list<MapPoint *>::iterator lit = lLocalMapPoints.begin(); lit != lLocalMapPoints.end(); lit++) {   // Optimizer.cc:595-675
        MapPoint *pMP = *lit;
        const int pi = (int)mps.size();
        mps.push_back(pMP);
        cv::Mat X = pMP->GetWorldPos();
        points.push_back(X.at<float>(0)); points.push_back(X.at<float>(1)); points.push_back(X.at<float>(2));
        const std::map<KeyFrame *, size_t> observations = pMP->GetObservations();
        for (std::map<KeyFrame *, size_t>::const_iterator mit = observations.begin(); mit != observations.end(); mit++) {
            KeyFrame *pKFi = mit->first;
            if (pKFi->isBad()) continue;
            std::map<KeyFrame *, int>::iterator it = kfIndex.find(pKFi);
            if (it == kfIndex.end()) continue;
            const cv::KeyPoint &kpUn = pKFi->mvKeysUn[mit->second];
            if (!(pKFi->mvuRight[mit->second] < 0)) throw std::runtime_error("orbslamm_b200: stereo observations are not supported");
            ekf.push_back(it->second); ept.push_back(pi);
            uv.push_back(kpUn.pt.x); uv.push_back(kpUn.pt.y); w.push_back(pKFi->mvInvLevelSigma2[kpUn.octave]);
            vpEdgeKF.push_back(pKFi); vpEdgeMP.push_back(pMP);
        }
    }
    if (pbStopFlag && *pbStopFlag) return;                                   // Optimizer.cc:678-680
    if (mps.empty() || ekf.empty()) return;
    StopBridge stop(pbStopFlag);
    std::vector<uint8_t> outlier(ekf.size());
    const int rc = orbo_bundle_adjust(handle(), (int)kfs.size(), poses.data(), fixed.data(), intr.data(), (int)mps.size(), points.data(), (int)ekf.size(),
                                      ekf.data(), ept.data(), uv.data(), w.data(), 1, 5, 10, 1, pbStopFlag ? &stop.flag : nullptr, nullptr, nullptr,
                                      outlier.data(), nullptr);
    check(rc, "orbo_bundle_adjust");
    if (rc == 1) return;
    std::unique_lock<std::mutex> lock(pMap->mMutexMapUpdate);                // Optimizer.cc:769
    for (size_t e = 0; e < outlier.size(); e++) {                            // Optimizer.cc:771-780
        if (!outlier[e] || vpEdgeMP[e]->isBad()) continue;
        vpEdgeKF[e]->EraseMapPointMatch(vpEdgeMP[e]);
        vpEdgeMP[e]->EraseObservation(vpEdgeKF[e]);
    }
    size_t i = 0;
    for (std::list<KeyFrame *>::iterator lit = lLocalKeyFrames.begin(); lit != lLocalKeyFrames.end(); lit++, i++)   // Optimizer.cc:785-791
        (*lit)->SetPose(flat_to_pose(&poses[16 * i]));
    for (size_t p = 0; p < mps.size(); p++) {                                // Optimizer.cc:794-800
        cv::Mat X(3, 1, CV_32F);
        for (int c = 0; c < 3; c++) X.at<float>(c) = points[3 * p + c];
        mps[p]->SetWorldPos(X);
        mps[p]->UpdateNormalAndDepth();
    }
}

// LoopClosing::ComputeSim3 / MultiMapper -> Optimizer::OptimizeSim3(mpCurrentKF, pKF, vpMapPointMatches, gScm, 10, mbFixScale)   (LoopClosing.cc:332)
int Optimizer::OptimizeSim3(KeyFrame *pKF1, KeyFrame *pKF2, std::vector<MapPoint *> &vpMatches1, g2o::Sim3 &g2oS12, const float th2, const bool bFixScale)
{
    const int32_t N = (int32_t)vpMatches1.size();
    if (!N) return 0;
    float R1w[9], t1w[3], R2w[9], t2w[3];
    {
        const cv::Mat R1 = pKF1->GetRotation(), T1 = pKF1->GetTranslation(), R2 = pKF2->GetRotation(), T2 = pKF2->GetTranslation();
        for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) { R1w[3 * r + c] = R1.at<float>(r, c); R2w[3 * r + c] = R2.at<float>(r, c); } t1w[r] = T1.at<float>(r); t2w[r] = T2.at<float>(r); }
    }
    const std::vector<MapPoint *> vpMapPoints1 = pKF1->GetMapPointMatches();
    std::vector<uint8_t> valid(N, 0), inlier(N, 0);
    std::vector<float> P1c(3 * (size_t)N, 0.f), P2c(3 * (size_t)N, 0.f), obs1(2 * (size_t)N, 0.f), obs2(2 * (size_t)N, 0.f), w1(N, 0.f), w2(N, 0.f);
    auto to_camera = [](const float *R, const float *t, const cv::Mat &X, float *out) {   // R*P3Dw + t: cv::Mat small gemm (fp32, left to right), then add
        for (int r = 0; r < 3; r++) {
            float s = R[3 * r] * X.at<float>(0);
            s = s + R[3 * r + 1] * X.at<float>(1);
            s = s + R[3 * r + 2] * X.at<float>(2);
            out[r] = s + t[r];
        }
    };
    for (int i = 0; i < N; i++) {                                               // Optimizer.cc:1398-1463
        if (!vpMatches1[i]) continue;
        MapPoint *pMP1 = vpMapPoints1[i], *pMP2 = vpMatches1[i];
        const int i2 = pMP2->GetIndexInKeyFrame(pKF2);
        if (!pMP1 || !pMP2) continue;
        if (pMP1->isBad() || pMP2->isBad() || i2 < 0) continue;
        to_camera(R1w, t1w, pMP1->GetWorldPos(), &P1c[3 * i]);
        to_camera(R2w, t2w, pMP2->GetWorldPos(), &P2c[3 * i]);
        const cv::KeyPoint &kpUn1 = pKF1->mvKeysUn[i], &kpUn2 = pKF2->mvKeysUn[i2];
        obs1[2 * i] = kpUn1.pt.x; obs1[2 * i + 1] = kpUn1.pt.y; obs2[2 * i] = kpUn2.pt.x; obs2[2 * i + 1] = kpUn2.pt.y;
        w1[i] = pKF1->mvInvLevelSigma2[kpUn1.octave]; w2[i] = pKF2->mvInvLevelSigma2[kpUn2.octave];
        valid[i] = 1;
    }
    double S[8] = {g2oS12.rotation().x(), g2oS12.rotation().y(), g2oS12.rotation().z(), g2oS12.rotation().w(),
                   g2oS12.translation()[0], g2oS12.translation()[1], g2oS12.translation()[2], g2oS12.scale()};
    const float K1[4] = {pKF1->fx, pKF1->fy, pKF1->cx, pKF1->cy}, K2[4] = {pKF2->fx, pKF2->fy, pKF2->cx, pKF2->cy};   // mK(0,0), (1,1), (0,2), (1,2)
    int32_t nIn = 0;
    check(orbo_optimize_sim3(handle(), 1, S, valid.data(), P1c.data(), P2c.data(), obs1.data(), obs2.data(), w1.data(), w2.data(), K1, K2, &N, N, th2,
                             bFixScale ? 1 : 0, inlier.data(), &nIn, nullptr, ORBS_MEM_HOST), "orbo_optimize_sim3");
    for (int i = 0; i < N; i++) if (valid[i] && !inlier[i]) vpMatches1[i] = static_cast<MapPoint *>(NULL);          // :1476, :1516
    g2oS12.rotation().x() = S[0]; g2oS12.rotation().y() = S[1]; g2oS12.rotation().z() = S[2]; g2oS12.rotation().w() = S[3];
    g2oS12.translation()[0] = S[4]; g2oS12.translation()[1] = S[5]; g2oS12.translation()[2] = S[6]; g2oS12.scale() = S[7];
    return nIn;
}

}  // namespace iORB_SLAM
