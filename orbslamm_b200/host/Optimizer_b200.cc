// Optimizer_b200.cc -- GPU-backed definitions of Optimizer::PoseOptimization / LocalBundleAdjustment /
// BundleAdjustment / GlobalBundleAdjustemnt.  In the reference tree these replace the same-named functions of
// S/src/Optimizer.cc (:60-65, :68-260, :262-474, :476-801); OptimizeSim3 and the essential-graph optimisers stay on
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
        fixed.push_back(pKF->mnId == 0 ? 1 : 0);
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

}  // namespace iORB_SLAM
