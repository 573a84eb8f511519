// Optimizer_b200.cc -- GPU-backed definitions of Optimizer::PoseOptimization / LocalBundleAdjustment /
// BundleAdjustment / GlobalBundleAdjustemnt.  In the reference tree these replace the same-named functions of
// S/src/Optimizer.cc (:60-65, :68-260, :262-474, :476-801) OptimizeSim3 (:1348-1543) and (MM)OptimizeEssentialGraph (:804-1067, M: :1069-1346):
// every member of the class.  The graph *construction* below follows the reference line by line (which
// keyframes are local / fixed, which observations become edges, which locks are taken); only the numerical solve is
// delegated to orbo_* (include/orbslamm_b200.h).  Monocular observations only (mvuRight < 0).
#include <cmath>
#include <list>
#include <set>
#include <map>
#include <unordered_map>
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
// the reference's `bool *pbStopFlag` (mbAbortBA / mbStopGBA, raised by another thread while the optimisation runs) goes to the C-ABI as is:
// the library watches that byte for the whole call
static_assert(sizeof(bool) == 1, "the C-ABI stop flag is one byte");
inline const volatile uint8_t *stop_ptr(bool *p) { return reinterpret_cast<const volatile uint8_t *>(p); }
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
    const int rc = orbo_bundle_adjust(handle(), (int)kfs.size(), poses.data(), fixed.data(), intr.data(), (int)mps.size(), points.data(), (int)ekf.size(),
                                      ekf.data(), ept.data(), uv.data(), w.data(), 0, nIterations, 0, bRobust ? 1 : 0, stop_ptr(pbStopFlag),
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
    std::vector<uint8_t> outlier(ekf.size());
    const int rc = orbo_bundle_adjust(handle(), (int)kfs.size(), poses.data(), fixed.data(), intr.data(), (int)mps.size(), points.data(), (int)ekf.size(),
                                      ekf.data(), ept.data(), uv.data(), w.data(), 1, 5, 10, 1, stop_ptr(pbStopFlag), nullptr, nullptr,
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

// ---------------------------------------------------------------------------------------------------------------------------------------
// Essential graph.  The graph assembly below follows Optimizer.cc:820-1000 line by line (which keyframes become vertices, which edges are
// inserted and with which measurement); the solve is orbo_optimize_pose_graph.  g2o::Sim3 values are read and written through rotation() /
// translation() / scale(); the Sim3 products / inverses of the assembly are the formulas of g2o/types/sim3.h on plain doubles.
namespace
{
struct S8 { double v[8]; };          // r (x y z w), t, s
void q_rot(const double *q, const double *x, double *o)          // Eigen::Quaterniond * Vector3d (_transformVector)
{
    double uv[3] = {q[1] * x[2] - q[2] * x[1], q[2] * x[0] - q[0] * x[2], q[0] * x[1] - q[1] * x[0]};
    uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
    o[0] = x[0] + q[3] * uv[0] + (q[1] * uv[2] - q[2] * uv[1]);
    o[1] = x[1] + q[3] * uv[1] + (q[2] * uv[0] - q[0] * uv[2]);
    o[2] = x[2] + q[3] * uv[2] + (q[0] * uv[1] - q[1] * uv[0]);
}
S8 s8_mul(const S8 &a, const S8 &b)                               // Sim3::operator*, sim3.h:214-220
{
    S8 r;
    const double *p = a.v, *q = b.v;
    r.v[3] = p[3] * q[3] - p[0] * q[0] - p[1] * q[1] - p[2] * q[2];
    r.v[0] = p[3] * q[0] + p[0] * q[3] + p[1] * q[2] - p[2] * q[1];
    r.v[1] = p[3] * q[1] + p[1] * q[3] + p[2] * q[0] - p[0] * q[2];
    r.v[2] = p[3] * q[2] + p[2] * q[3] + p[0] * q[1] - p[1] * q[0];
    double rt[3];
    q_rot(a.v, b.v + 4, rt);
    for (int k = 0; k < 3; k++) r.v[4 + k] = a.v[7] * rt[k] + a.v[4 + k];
    r.v[7] = a.v[7] * b.v[7];
    return r;
}
S8 s8_inv(const S8 &a)                                            // Sim3::inverse, sim3.h:184-187
{
    S8 r;
    r.v[0] = -a.v[0]; r.v[1] = -a.v[1]; r.v[2] = -a.v[2]; r.v[3] = a.v[3];
    const double f = -1. / a.v[7], st[3] = {f * a.v[4], f * a.v[5], f * a.v[6]};
    q_rot(r.v, st, r.v + 4);
    r.v[7] = 1. / a.v[7];
    return r;
}
void s8_map(const S8 &a, const double *x, double *o) { double r[3]; q_rot(a.v, x, r); for (int k = 0; k < 3; k++) o[k] = a.v[7] * r[k] + a.v[4 + k]; }   // s*(r*xyz) + t
S8 s8_of(const g2o::Sim3 &s) { S8 r; r.v[0] = s.rotation().x(); r.v[1] = s.rotation().y(); r.v[2] = s.rotation().z(); r.v[3] = s.rotation().w();
                        for (int k = 0; k < 3; k++) { r.v[4 + k] = s.translation()[k]; }
                        r.v[7] = s.scale(); return r; }
S8 s8_of_pose(KeyFrame *pKF)                                      // g2o::Sim3 Siw(Converter::toMatrix3d(Rcw), Converter::toVector3d(tcw), 1.0): Quaterniond(Matrix3d)
{
    const cv::Mat Rm = pKF->GetRotation(), tm = pKF->GetTranslation();
    double R[9];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) R[3 * r + c] = Rm.at<float>(r, c);
    S8 o;
    double *q = o.v, t = R[0] + R[4] + R[8];
    if (t > 0) { t = std::sqrt(t + 1.0); q[3] = 0.5 * t; t = 0.5 / t; q[0] = (R[7] - R[5]) * t; q[1] = (R[2] - R[6]) * t; q[2] = (R[3] - R[1]) * t; }
    else {
        int i = 0;
        if (R[4] > R[0]) i = 1;
        if (R[8] > R[4 * i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
        q[i] = 0.5 * t; t = 0.5 / t;
        q[3] = (R[3 * k + j] - R[3 * j + k]) * t; q[j] = (R[3 * j + i] + R[3 * i + j]) * t; q[k] = (R[3 * k + i] + R[3 * i + k]) * t;
    }
    for (int k = 0; k < 3; k++) o.v[4 + k] = tm.at<float>(k);
    o.v[7] = 1.0;
    return o;
}

void essential_graph(const std::vector<KeyFrame *> &vpKFs, const std::vector<MapPoint *> &vpMPs, Map *pMap, KeyFrame *pLoopKF, KeyFrame *pCurKF,
                     const LoopClosing::KeyFrameAndPose &NonCorrectedSim3, const LoopClosing::KeyFrameAndPose &CorrectedSim3,
                     const std::map<KeyFrame *, std::set<KeyFrame *>> &LoopConnections, const bool &bFixScale)
{
    const int minFeat = 100;
    std::map<KeyFrame *, int> vid;                       // keyframe -> vertex (the reference indexes arrays by mnId)
    std::vector<S8> vScw;
    std::vector<uint8_t> fixed;
    for (size_t i = 0, iend = vpKFs.size(); i < iend; i++) {                    // vertices, :829-862
        KeyFrame *pKF = vpKFs[i];
        if (pKF->isBad()) continue;
        LoopClosing::KeyFrameAndPose::const_iterator it = CorrectedSim3.find(pKF);
        vid[pKF] = (int)vScw.size();
        vScw.push_back(it != CorrectedSim3.end() ? s8_of(it->second) : s8_of_pose(pKF));
        fixed.push_back(pKF == pLoopKF ? 1 : 0);
    }
    std::vector<int32_t> ei, ej;
    std::vector<double> meas;
    auto add_edge = [&](KeyFrame *a, KeyFrame *b, const S8 &Sba) {             // vertex[0] = a, vertex[1] = b, measurement Sba
        if (!vid.count(a) || !vid.count(b)) return;                             // g2o refuses an edge with a missing vertex
        ei.push_back(vid[a]); ej.push_back(vid[b]);
        meas.insert(meas.end(), Sba.v, Sba.v + 8);
    };
    auto non_corrected = [&](KeyFrame *pKF) -> S8 {                             // NonCorrectedSim3 if present, else the vertex estimate
        LoopClosing::KeyFrameAndPose::const_iterator it = NonCorrectedSim3.find(pKF);
        if (it != NonCorrectedSim3.end()) return s8_of(it->second);
        return vid.count(pKF) ? vScw[vid[pKF]] : s8_of_pose(pKF);
    };
    std::set<std::pair<long unsigned int, long unsigned int>> sInsertedEdges;
    for (std::map<KeyFrame *, std::set<KeyFrame *>>::const_iterator mit = LoopConnections.begin(), mend = LoopConnections.end(); mit != mend; mit++) {   // loop edges, :868-893
        KeyFrame *pKF = mit->first;
        if (!vid.count(pKF)) continue;
        const long unsigned int nIDi = pKF->mnId;
        const S8 Swi = s8_inv(vScw[vid[pKF]]);
        for (std::set<KeyFrame *>::const_iterator sit = mit->second.begin(), send = mit->second.end(); sit != send; sit++) {
            const long unsigned int nIDj = (*sit)->mnId;
            if ((nIDi != pCurKF->mnId || nIDj != pLoopKF->mnId) && pKF->GetWeight(*sit) < minFeat) continue;
            if (!vid.count(*sit)) continue;
            add_edge(pKF, *sit, s8_mul(vScw[vid[*sit]], Swi));
            sInsertedEdges.insert(std::make_pair(std::min(nIDi, nIDj), std::max(nIDi, nIDj)));
        }
    }
    for (size_t i = 0, iend = vpKFs.size(); i < iend; i++) {                    // normal edges, :896-993
        KeyFrame *pKF = vpKFs[i];
        if (!vid.count(pKF)) continue;
        const S8 Swi = s8_inv(non_corrected(pKF));
        KeyFrame *pParentKF = pKF->GetParent();
        if (pParentKF) add_edge(pKF, pParentKF, s8_mul(non_corrected(pParentKF), Swi));                              // spanning tree
        const std::set<KeyFrame *> sLoopEdges = pKF->GetLoopEdges();
        for (std::set<KeyFrame *>::const_iterator sit = sLoopEdges.begin(), send = sLoopEdges.end(); sit != send; sit++) {   // earlier loop closures
            KeyFrame *pLKF = *sit;
            if (pLKF->mnId < pKF->mnId) add_edge(pKF, pLKF, s8_mul(non_corrected(pLKF), Swi));
        }
        const std::vector<KeyFrame *> vpConnectedKFs = pKF->GetCovisiblesByWeight(minFeat);                           // strong covisibility
        for (std::vector<KeyFrame *>::const_iterator vit = vpConnectedKFs.begin(); vit != vpConnectedKFs.end(); vit++) {
            KeyFrame *pKFn = *vit;
            if (pKFn && pKFn != pParentKF && !pKF->hasChild(pKFn) && !sLoopEdges.count(pKFn)) {
                if (!pKFn->isBad() && pKFn->mnId < pKF->mnId) {
                    if (sInsertedEdges.count(std::make_pair(std::min(pKF->mnId, pKFn->mnId), std::max(pKF->mnId, pKFn->mnId)))) continue;
                    add_edge(pKF, pKFn, s8_mul(non_corrected(pKFn), Swi));
                }
            }
        }
    }
    const int32_t K = (int32_t)vScw.size(), E = (int32_t)ei.size();
    std::vector<double> est(8 * (size_t)std::max(K, 1));
    for (int k = 0; k < K; k++) for (int c = 0; c < 8; c++) est[8 * (size_t)k + c] = vScw[k].v[c];
    if (K > 0 && E > 0)                                                                                               // optimizer.optimize(20), lambda init 1e-16 (:812, :996-997)
        check(orbo_optimize_pose_graph(handle(), K, est.data(), fixed.data(), E, ei.data(), ej.data(), meas.data(), bFixScale ? 1 : 0, 20, 1e-16, nullptr),
              "orbo_optimize_pose_graph");

    std::unique_lock<std::mutex> lock(pMap->mMutexMapUpdate);                                                         // :999
    std::vector<S8> vCorrectedSwc(K);
    for (size_t i = 0; i < vpKFs.size(); i++) {                                 // SE3 pose recovering: Sim3 [sR t; 0 1] -> SE3 [R t/s; 0 1], :1002-1018
        KeyFrame *pKFi = vpKFs[i];
        if (!vid.count(pKFi)) continue;
        const int k = vid[pKFi];
        S8 C;
        for (int c = 0; c < 8; c++) C.v[c] = est[8 * (size_t)k + c];
        vCorrectedSwc[k] = s8_inv(C);
        const double *q = C.v;                                                  // Quaterniond::toRotationMatrix
        const double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2], twx = tx * q[3], twy = ty * q[3], twz = tz * q[3], txx = tx * q[0], txy = ty * q[0], txz = tz * q[0],
                     tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
        const double R[9] = {1 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1 - (txx + tzz), tyz - twx, txz - twy, tyz + twx, 1 - (txx + tyy)};
        const double is = 1. / C.v[7];
        cv::Mat Tiw = cv::Mat::eye(4, 4, CV_32F);                               // Converter::toCvSE3
        for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) Tiw.at<float>(r, c) = (float)R[3 * r + c]; Tiw.at<float>(r, 3) = (float)(C.v[4 + r] * is); }
        pKFi->SetPose(Tiw);
    }
    std::unordered_map<unsigned long, int> vertex_of_id;                        // mnId -> vertex (the reference indexes vScw by mnId, :1032-1041)
    vertex_of_id.reserve(vid.size());
    for (std::map<KeyFrame *, int>::const_iterator it = vid.begin(); it != vid.end(); ++it) vertex_of_id[it->first->mnId] = it->second;
    for (size_t i = 0, iend = vpMPs.size(); i < iend; i++) {                    // map points: corrected by their reference keyframe, :1021-1045
        MapPoint *pMP = vpMPs[i];
        if (pMP->isBad()) continue;
        const unsigned long nIDr = pMP->mnCorrectedByKF == pCurKF->mnId ? pMP->mnCorrectedReference : pMP->GetReferenceKeyFrame()->mnId;
        const std::unordered_map<unsigned long, int>::const_iterator f = vertex_of_id.find(nIDr);
        const int k = f == vertex_of_id.end() ? -1 : f->second;
        cv::Mat P3Dw = pMP->GetWorldPos();
        if (k < 0) {
            // the reference keyframe is not a vertex (bad keyframe): the reference reads default-constructed (identity) Sim3s from its mnId-indexed
            // vectors, i.e. the position is kept but still re-set and the normal / depth refreshed
            pMP->SetWorldPos(P3Dw);
            pMP->UpdateNormalAndDepth();
            continue;
        }
        const double X[3] = {P3Dw.at<float>(0), P3Dw.at<float>(1), P3Dw.at<float>(2)};
        double Xr[3], Xc[3];
        s8_map(vScw[k], X, Xr);
        s8_map(vCorrectedSwc[k], Xr, Xc);
        cv::Mat out(3, 1, CV_32F);
        for (int c = 0; c < 3; c++) out.at<float>(c) = (float)Xc[c];
        pMP->SetWorldPos(out);
        pMP->UpdateNormalAndDepth();
    }
}
}  // namespace

void Optimizer::OptimizeEssentialGraph(Map *pMap, KeyFrame *pLoopKF, KeyFrame *pCurKF, const LoopClosing::KeyFrameAndPose &NonCorrectedSim3,
                                       const LoopClosing::KeyFrameAndPose &CorrectedSim3, const std::map<KeyFrame *, std::set<KeyFrame *>> &LoopConnections,
                                       const bool &bFixScale)
{
    essential_graph(pMap->GetAllKeyFrames(), pMap->GetAllMapPoints(), pMap, pLoopKF, pCurKF, NonCorrectedSim3, CorrectedSim3, LoopConnections, bFixScale);
}

#ifdef ORBSLAMM_MULTI_ROBOT
// MultipleRobotsScenario/src/Optimizer.cc:1069-1346: the same optimisation over the map and every map attached to it by MultiMapper
void Optimizer::MMOptimizeEssentialGraph(Map *pMap, KeyFrame *pLoopKF, KeyFrame *pCurKF, const LoopClosing::KeyFrameAndPose &NonCorrectedSim3,
                                         const LoopClosing::KeyFrameAndPose &CorrectedSim3, const std::map<KeyFrame *, std::set<KeyFrame *>> &LoopConnections,
                                         const bool &bFixScale)
{
    std::vector<KeyFrame *> vpKFs = pMap->GetAllKeyFrames();
    std::vector<MapPoint *> vpMPs = pMap->GetAllMapPoints();
    if (pMap->isAttached()) {
        std::vector<Map *> vpAttachedMaps = pMap->getAttachedMaps();
        for (std::vector<Map *>::iterator it = vpAttachedMaps.begin(), itend = vpAttachedMaps.end(); it != itend; it++) {
            std::vector<KeyFrame *> vpAttachedKFs = (*it)->GetAllKeyFrames();
            std::vector<MapPoint *> vpAttachedMPs = (*it)->GetAllMapPoints();
            vpKFs.insert(vpKFs.end(), vpAttachedKFs.begin(), vpAttachedKFs.end());
            vpMPs.insert(vpMPs.end(), vpAttachedMPs.begin(), vpAttachedMPs.end());
        }
    }
    essential_graph(vpKFs, vpMPs, pMap, pLoopKF, pCurKF, NonCorrectedSim3, CorrectedSim3, LoopConnections, bFixScale);
}
#endif

}  // namespace iORB_SLAM
