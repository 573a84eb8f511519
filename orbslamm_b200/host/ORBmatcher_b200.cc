// ORBmatcher_b200.cc -- GPU-backed definitions of the ORBmatcher members on the tracking hot path and of the KeyFrame / Sim3
// projection family.  In the reference tree these replace the same-named functions of S/src/ORBmatcher.cc (:41-43 ctor, :45-129
// and :1330-1472 SearchByProjection, :131-137 RadiusByViewingCos, :1649-1665 DescriptorDistance, :292-405 SearchByProjection(KF, Scw),
// :827-977 and :979-1102 Fuse, :1104-1328 SearchBySim3, :1474-1601 SearchByProjection(Frame, KF), :159-290 and :524-657 SearchByBoW,
// :407-522 SearchForInitialization, :659-825 SearchForTriangulation) -- every member of the class: ORBmatcher.cc leaves the build.
// (ComputeThreeMaxima and CheckDistEpipolarLine, the two protected helpers, run inside the kernels.)
//
// The shim only marshals: Frame/MapPoint fields -> flat arrays -> orbm_* (include/orbslamm_b200.h) -> pointers written
// back into Frame::mvpMapPoints.  ORBmatcher objects are stack-constructed from several threads in the reference, so
// the device workspace is a thread_local handle.
#include <cmath>
#include <cstring>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>
#include "ORBmatcher.h"
#include "orbslamm_b200.h"

namespace iORB_SLAM
{

const int ORBmatcher::TH_HIGH = 100;
const int ORBmatcher::TH_LOW = 50;
const int ORBmatcher::HISTO_LENGTH = 30;

namespace
{
struct MatcherTLS {
    orbm_handle *h = nullptr;
    ~MatcherTLS() { orbm_destroy(h); }
};
orbm_handle *handle()
{
    static thread_local MatcherTLS tls;
    if (!tls.h && orbm_create(&tls.h, 0) != ORBS_OK) throw std::runtime_error(std::string("orbm_create: ") + orbs_last_error());
    return tls.h;
}
void check(int rc, const char *what) { if (rc != ORBS_OK) throw std::runtime_error(std::string(what) + ": " + orbs_last_error()); }

// flat view of a frame's features; descriptors are 16-byte aligned through the vector<uint4>-like storage
struct FrameArrays {
    std::vector<float> xy, angle;
    std::vector<int32_t> octave, match;
    std::vector<uint64_t> desc64;          // 4 x u64 per descriptor: guarantees the 16-byte alignment the C-ABI wants
    int32_t n = 0;
    explicit FrameArrays(Frame &F)
    {
        n = F.N;
        xy.resize(2 * (size_t)n); angle.resize(n); octave.resize(n); match.assign(n, -1); desc64.resize(4 * (size_t)n + 2);
        uint8_t *d = desc();
        for (int i = 0; i < n; i++) {
            const cv::KeyPoint &kp = F.mvKeysUn[i];
            xy[2 * i] = kp.pt.x; xy[2 * i + 1] = kp.pt.y; angle[i] = kp.angle; octave[i] = kp.octave;
            std::memcpy(d + 32 * (size_t)i, F.mDescriptors.ptr(i), 32);
            // a feature that already holds a map point with observations is skipped by the search (ORBmatcher.cc:86-88,1406-1408)
            MapPoint *p = F.mvpMapPoints[i];
            if (p && p->Observations() > 0) match[i] = 0x40000000;
        }
    }
    uint8_t *desc() { uintptr_t p = (uintptr_t)desc64.data(); p = (p + 15) & ~(uintptr_t)15; return (uint8_t *)p; }
};
}  // namespace

ORBmatcher::ORBmatcher(float nnratio, bool checkOri) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}

float ORBmatcher::RadiusByViewingCos(const float &viewCos) { return viewCos > 0.998 ? 2.5 : 4.0; }

int ORBmatcher::DescriptorDistance(const cv::Mat &a, const cv::Mat &b)
{
    // single pair through the batched device entry point; bulk callers should call orbm_descriptor_distance directly
    alignas(16) uint8_t da[32], db[32];
    std::memcpy(da, a.ptr(0), 32); std::memcpy(db, b.ptr(0), 32);
    int32_t d = 0;
    check(orbm_descriptor_distance(handle(), da, 1, db, 1, &d, ORBS_MEM_HOST), "orbm_descriptor_distance");
    return d;
}

// MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:242-307) for all the map points of one caller loop in ONE device call.  The reference calls the
// member point by point (LocalMapping.cc:163, 436, 628; LoopClosing.cc:500; MultiMapper.cc:607), each call doing O(N^2) DescriptorDistance evaluations on the
// host; the loops become `ORBmatcher::ComputeDistinctiveDescriptors(points)` after them.
void ORBmatcher::ComputeDistinctiveDescriptors(const std::vector<MapPoint *> &vpMPs)
{
    std::vector<MapPoint *> pts;
    std::vector<int32_t> start(1, 0);
    std::vector<cv::Mat> rows;
    for (size_t i = 0; i < vpMPs.size(); i++) {
        MapPoint *pMP = vpMPs[i];
        if (!pMP || pMP->isBad()) continue;                                      // MapPoint.cc:251-253
        const std::map<KeyFrame *, size_t> observations = pMP->GetObservations();
        const size_t before = rows.size();
        for (std::map<KeyFrame *, size_t>::const_iterator mit = observations.begin(); mit != observations.end(); ++mit)
            if (!mit->first->isBad()) rows.push_back(mit->first->mDescriptors.row((int)mit->second));
        if (rows.size() == before) continue;                                     // :257-258, :269-270: nothing to choose from
        pts.push_back(pMP); start.push_back((int32_t)rows.size());
    }
    if (pts.empty()) return;
    std::vector<uint64_t> buf(rows.size() * 4 + 2);
    uintptr_t a = ((uintptr_t)buf.data() + 15) & ~(uintptr_t)15;
    uint8_t *flat = (uint8_t *)a;
    for (size_t r = 0; r < rows.size(); r++) std::memcpy(flat + 32 * r, rows[r].ptr(0), 32);
    std::vector<int32_t> best(pts.size(), 0);
    check(orbm_distinctive_descriptors(handle(), (int)pts.size(), flat, start.data(), best.data(), ORBS_MEM_HOST), "orbm_distinctive_descriptors");
    for (size_t p = 0; p < pts.size(); p++) pts[p]->SetDescriptor(rows[start[p] + best[p]].clone());
}

// Tracking::SearchLocalPoints -> ORBmatcher(0.8).SearchByProjection(mCurrentFrame, mvpLocalMapPoints, th)   (Tracking.cc:1249)
int ORBmatcher::SearchByProjection(Frame &F, const std::vector<MapPoint *> &vpMapPoints, const float th)
{
    const bool bFactor = th != 1.0;
    FrameArrays fa(F);
    const size_t M = vpMapPoints.size();
    if (!M || !fa.n) return 0;
    std::vector<uint8_t> valid(M, 0);
    std::vector<float> uv(2 * M), radius(M), qangle(M, 0.f);
    std::vector<int32_t> minl(M), maxl(M);
    std::vector<uint64_t> qd64(4 * M + 2);
    uint8_t *qd = (uint8_t *)(((uintptr_t)qd64.data() + 15) & ~(uintptr_t)15);
    for (size_t i = 0; i < M; i++) {
        MapPoint *pMP = vpMapPoints[i];
        if (!pMP->mbTrackInView || pMP->isBad()) continue;          // ORBmatcher.cc:54-58
        const int lvl = pMP->mnTrackScaleLevel;
        float r = RadiusByViewingCos(pMP->mTrackViewCos);
        if (bFactor) r *= th;
        valid[i] = 1;
        uv[2 * i] = pMP->mTrackProjX; uv[2 * i + 1] = pMP->mTrackProjY;
        radius[i] = r * F.mvScaleFactors[lvl];
        minl[i] = lvl - 1; maxl[i] = lvl;
        cv::Mat d = pMP->GetDescriptor();
        std::memcpy(qd + 32 * i, d.ptr(0), 32);
    }
    const float bounds[4] = {Frame::mnMinX, Frame::mnMinY, Frame::mnMaxX, Frame::mnMaxY};
    const int32_t qn = (int32_t)M;
    int32_t nmatches = 0;
    check(orbm_search_by_projection(handle(), 1, bounds, fa.xy.data(), fa.octave.data(), fa.angle.data(), fa.desc(), &fa.n, fa.n, valid.data(),
                                    uv.data(), radius.data(), minl.data(), maxl.data(), qangle.data(), qd, &qn, qn, TH_HIGH, mfNNratio, 0,
                                    fa.match.data(), &nmatches, ORBS_MEM_HOST), "orbm_search_by_projection");
    for (int k = 0; k < fa.n; k++)
        if (fa.match[k] >= 0 && fa.match[k] < qn) F.mvpMapPoints[k] = vpMapPoints[fa.match[k]];
    return nmatches;
}

// Tracking::TrackWithMotionModel -> ORBmatcher(0.9,true).SearchByProjection(mCurrentFrame, mLastFrame, th, mono)   (Tracking.cc:930)
int ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono)
{
    if (!bMono) throw std::runtime_error("orbslamm_b200: only the monocular SearchByProjection is accelerated");
    FrameArrays fa(CurrentFrame);
    const int M = LastFrame.N;
    if (!M || !fa.n) return 0;
    std::vector<uint8_t> valid(M, 0);
    std::vector<float> Xw(3 * (size_t)M, 0.f), qangle(M), uv(2 * (size_t)M), radius(M);
    std::vector<int32_t> loct(M), minl(M), maxl(M);
    std::vector<uint64_t> qd64(4 * (size_t)M + 2);
    uint8_t *qd = (uint8_t *)(((uintptr_t)qd64.data() + 15) & ~(uintptr_t)15);
    for (int i = 0; i < M; i++) {
        MapPoint *pMP = LastFrame.mvpMapPoints[i];
        loct[i] = LastFrame.mvKeys[i].octave; qangle[i] = LastFrame.mvKeysUn[i].angle;
        if (!pMP || LastFrame.mvbOutlier[i]) continue;              // ORBmatcher.cc:1355-1359
        valid[i] = 1;
        cv::Mat x3Dw = pMP->GetWorldPos();
        Xw[3 * i] = x3Dw.at<float>(0); Xw[3 * i + 1] = x3Dw.at<float>(1); Xw[3 * i + 2] = x3Dw.at<float>(2);
        cv::Mat d = pMP->GetDescriptor();
        std::memcpy(qd + 32 * (size_t)i, d.ptr(0), 32);
    }
    float Tcw[16];
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) Tcw[4 * r + c] = CurrentFrame.mTcw.at<float>(r, c);
    const float K4[4] = {Frame::fx, Frame::fy, Frame::cx, Frame::cy};
    const float bounds[4] = {Frame::mnMinX, Frame::mnMinY, Frame::mnMaxX, Frame::mnMaxY};
    const int32_t qn = M;
    check(orbm_project_last_frame(handle(), 1, Tcw, K4, bounds, CurrentFrame.mvScaleFactors.data(), (int)CurrentFrame.mvScaleFactors.size(), Xw.data(),
                                  loct.data(), &qn, qn, th, valid.data(), uv.data(), radius.data(), minl.data(), maxl.data(), ORBS_MEM_HOST),
          "orbm_project_last_frame");
    int32_t nmatches = 0;
    check(orbm_search_by_projection(handle(), 1, bounds, fa.xy.data(), fa.octave.data(), fa.angle.data(), fa.desc(), &fa.n, fa.n, valid.data(),
                                    uv.data(), radius.data(), minl.data(), maxl.data(), qangle.data(), qd, &qn, qn, TH_HIGH, 0.f,
                                    mbCheckOrientation ? 1 : 0, fa.match.data(), &nmatches, ORBS_MEM_HOST), "orbm_search_by_projection");
    for (int k = 0; k < fa.n; k++) {
        if (fa.match[k] >= 0 && fa.match[k] < qn) CurrentFrame.mvpMapPoints[k] = LastFrame.mvpMapPoints[fa.match[k]];
        else if (fa.match[k] < 0 && CurrentFrame.mvpMapPoints[k] && !(CurrentFrame.mvpMapPoints[k]->Observations() > 0)) { /* untouched */ }
    }
    return nmatches;
}

// ---------------------------------------------------------------------------------------------------------------------------------------
// KeyFrame / Sim3 projection family.  Each member = host marshalling -> orbm_project_points -> one window search -> the reference's own
// pointer-graph updates on the host, in the reference's order.
namespace
{
// cv::Mat algebra that precedes the loops, written out in OpenCV's evaluation order so that it does not depend on which cv::Mat is
// compiled in (Mat / s = Mat * (1.0 / s) in double, rounded to float; A * B = small gemm, fp32 products added left to right)
void mat33(const cv::Mat &M, float *R) { for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) R[3 * r + c] = M.at<float>(r, c); }
void vec3(const cv::Mat &v, float *t) { for (int r = 0; r < 3; r++) t[r] = v.at<float>(r); }
void neg_Rt_times(const float *R, const float *t, float *out)        // -R^T * t
{
    for (int r = 0; r < 3; r++) {
        float s = -R[r] * t[0];
        s = s + -R[3 + r] * t[1];
        s = s + -R[6 + r] * t[2];
        out[r] = s;
    }
}
void decompose_scw(const cv::Mat &Scw, float *Rcw, float *tcw, float *Ow)      // ORBmatcher.cc:300-305 / :987-992
{
    double d = 0;
    for (int c = 0; c < 3; c++) d += (double)Scw.at<float>(0, c) * (double)Scw.at<float>(0, c);
    const float scw = (float)std::sqrt(d);
    const double inv = 1. / (double)scw;
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) Rcw[3 * r + c] = (float)((double)Scw.at<float>(r, c) * inv);
        tcw[r] = (float)((double)Scw.at<float>(r, 3) * inv);
    }
    neg_Rt_times(Rcw, tcw, Ow);
}

struct KeyFrameArrays {                      // flat view of the target keyframe's features
    std::vector<float> xy, angle; std::vector<int32_t> octave; std::vector<uint64_t> desc64; int32_t n = 0;
    float grid_bounds[4], win_origin[2], kf_bounds[4];
    explicit KeyFrameArrays(KeyFrame *pKF)
    {
        n = pKF->N;
        xy.resize(2 * (size_t)n + 2); angle.resize(n + 1); octave.resize(n + 1); desc64.resize(4 * (size_t)n + 2);
        uint8_t *d = desc();
        for (int i = 0; i < n; i++) {
            const cv::KeyPoint &kp = pKF->mvKeysUn[i];
            xy[2 * i] = kp.pt.x; xy[2 * i + 1] = kp.pt.y; angle[i] = kp.angle; octave[i] = kp.octave;
            std::memcpy(d + 32 * (size_t)i, pKF->mDescriptors.ptr(i), 32);
        }
        // KeyFrame keeps the Frame's grid (cell sizes copied, KeyFrame.cc:54-60) but truncated integer bounds (S/include/KeyFrame.h); the
        // grid itself was filled with the Frame's float bounds, which are still the current Frame statics
        grid_bounds[0] = Frame::mnMinX; grid_bounds[1] = Frame::mnMinY; grid_bounds[2] = Frame::mnMaxX; grid_bounds[3] = Frame::mnMaxY;
        kf_bounds[0] = (float)pKF->mnMinX; kf_bounds[1] = (float)pKF->mnMinY; kf_bounds[2] = (float)pKF->mnMaxX; kf_bounds[3] = (float)pKF->mnMaxY;
        win_origin[0] = kf_bounds[0]; win_origin[1] = kf_bounds[1];
    }
    uint8_t *desc() { return (uint8_t *)(((uintptr_t)desc64.data() + 15) & ~(uintptr_t)15); }
};

struct PointArrays {                         // flat view of the candidate map points
    std::vector<float> Xw, normal, mf_min, mf_max, uv, radius; std::vector<int32_t> minl, maxl; std::vector<uint8_t> valid; std::vector<uint64_t> desc64;
    int32_t n = 0;
    explicit PointArrays(size_t M) : Xw(3 * M + 3), normal(3 * M + 3), mf_min(M + 1), mf_max(M + 1), uv(2 * M + 2), radius(M + 1), minl(M + 1), maxl(M + 1),
                                     valid(M + 1, 0), desc64(4 * M + 2), n((int32_t)M) {}
    uint8_t *desc() { return (uint8_t *)(((uintptr_t)desc64.data() + 15) & ~(uintptr_t)15); }
    void set(size_t i, MapPoint *pMP)
    {
        valid[i] = 1;
        cv::Mat X = pMP->GetWorldPos(), Nn = pMP->GetNormal(), d = pMP->GetDescriptor();
        for (int k = 0; k < 3; k++) { Xw[3 * i + k] = X.at<float>(k); normal[3 * i + k] = Nn.at<float>(k); }
        mf_min[i] = pMP->GetMinDistance(); mf_max[i] = pMP->GetMaxDistance();
        std::memcpy(desc() + 32 * i, d.ptr(0), 32);
    }
};

void fill_view(orbm_projection &V, const float *R, const float *t, const float *Ow, float fx, float fy, float cx, float cy, const float *bounds4, float log_sf,
               float th, int flags)
{
    std::memset(&V, 0, sizeof V);
    std::memcpy(V.R, R, sizeof V.R); std::memcpy(V.t, t, sizeof V.t);
    if (Ow) std::memcpy(V.Ow, Ow, sizeof V.Ow);
    V.fx = fx; V.fy = fy; V.cx = cx; V.cy = cy;
    V.min_x = bounds4[0]; V.min_y = bounds4[1]; V.max_x = bounds4[2]; V.max_y = bounds4[3];
    V.log_scale_factor = log_sf; V.th = th; V.flags = flags;
}

void project(const orbm_projection &V, const std::vector<float> &scale_factors, PointArrays &P)
{
    if (!P.n) return;
    check(orbm_project_points(handle(), 1, &V, scale_factors.data(), (int)scale_factors.size(), P.Xw.data(), P.normal.data(), P.mf_min.data(), P.mf_max.data(),
                              &P.n, P.n, P.valid.data(), P.uv.data(), P.radius.data(), P.minl.data(), P.maxl.data(), nullptr, ORBS_MEM_HOST), "orbm_project_points");
}

// best feature per point, no claims (Fuse, SearchBySim3)
std::vector<int32_t> best_in_window(KeyFrameArrays &K, PointArrays &P, int th_dist, const std::vector<float> *inv_sigma2)
{
    std::vector<int32_t> idx(P.n + 1, -1), dist(P.n + 1, -1);
    if (!P.n || !K.n) return idx;
    check(orbm_search_best_in_window(handle(), 1, K.grid_bounds, K.win_origin, K.xy.data(), K.octave.data(), K.desc(), &K.n, K.n, P.valid.data(), P.uv.data(),
                                     P.radius.data(), P.minl.data(), P.maxl.data(), P.desc(), &P.n, P.n, th_dist, inv_sigma2 ? inv_sigma2->data() : nullptr,
                                     inv_sigma2 ? (int)inv_sigma2->size() : 0, 5.99f, idx.data(), dist.data(), ORBS_MEM_HOST), "orbm_search_best_in_window");
    return idx;
}
}  // namespace

// LoopClosing::ComputeSim3 / MultiMapper -> matcher.SearchByProjection(mpCurrentKF, mScw, mvpLoopMapPoints, mvpCurrentMatchedPoints, 10)
int ORBmatcher::SearchByProjection(KeyFrame *pKF, cv::Mat Scw, const std::vector<MapPoint *> &vpPoints, std::vector<MapPoint *> &vpMatched, int th)
{
    float Rcw[9], tcw[3], Ow[3];
    decompose_scw(Scw, Rcw, tcw, Ow);
    std::set<MapPoint *> spAlreadyFound(vpMatched.begin(), vpMatched.end());
    spAlreadyFound.erase(static_cast<MapPoint *>(NULL));
    KeyFrameArrays K(pKF);
    PointArrays P(vpPoints.size());
    if (!P.n || !K.n) return 0;
    for (size_t i = 0; i < vpPoints.size(); i++) {
        MapPoint *pMP = vpPoints[i];
        if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;          // ORBmatcher.cc:319-320
        P.set(i, pMP);
    }
    orbm_projection V;
    fill_view(V, Rcw, tcw, Ow, pKF->fx, pKF->fy, pKF->cx, pKF->cy, K.kf_bounds, pKF->mfLogScaleFactor, (float)th, ORBM_PROJ_CHECK_NORMAL);
    project(V, pKF->mvScaleFactors, P);
    std::vector<int32_t> match(K.n, -1);
    for (int k = 0; k < K.n; k++) if (vpMatched[k]) match[k] = 0x40000000;   // "if(vpMatched[idx]) continue;" (:373-374)
    int32_t nmatches = 0;
    check(orbm_search_by_projection_kf(handle(), 1, K.grid_bounds, K.win_origin, K.xy.data(), K.octave.data(), nullptr, K.desc(), &K.n, K.n, P.valid.data(),
                                       P.uv.data(), P.radius.data(), P.minl.data(), P.maxl.data(), nullptr, P.desc(), &P.n, P.n, TH_LOW, 0.f, 0, match.data(),
                                       &nmatches, ORBS_MEM_HOST), "orbm_search_by_projection_kf");
    for (int k = 0; k < K.n; k++) if (match[k] >= 0 && match[k] < P.n) vpMatched[k] = vpPoints[match[k]];
    return nmatches;
}

// Tracking::Relocalization -> matcher2.SearchByProjection(mCurrentFrame, vpCandidateKFs[i], sFound, 10, 100) / (.., 3, 64)   (Tracking.cc:1454,1482)
int ORBmatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const std::set<MapPoint *> &sAlreadyFound, const float th, const int ORBdist)
{
    float Rcw[9], tcw[3], Ow[3];
    for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) Rcw[3 * r + c] = CurrentFrame.mTcw.at<float>(r, c); tcw[r] = CurrentFrame.mTcw.at<float>(r, 3); }
    neg_Rt_times(Rcw, tcw, Ow);
    const std::vector<MapPoint *> vpMPs = pKF->GetMapPointMatches();
    PointArrays P(vpMPs.size());
    const int N = CurrentFrame.N;
    if (!P.n || !N) return 0;
    std::vector<float> qangle(P.n, 0.f);
    for (size_t i = 0; i < vpMPs.size(); i++) {
        MapPoint *pMP = vpMPs[i];
        qangle[i] = pKF->mvKeysUn[i].angle;
        if (!pMP || pMP->isBad() || sAlreadyFound.count(pMP)) continue;   // :1494-1498
        P.set(i, pMP);
    }
    const float bounds[4] = {Frame::mnMinX, Frame::mnMinY, Frame::mnMaxX, Frame::mnMaxY};
    orbm_projection V;
    fill_view(V, Rcw, tcw, Ow, Frame::fx, Frame::fy, Frame::cx, Frame::cy, bounds, CurrentFrame.mfLogScaleFactor, th,
              ORBM_PROJ_NO_DEPTH | ORBM_PROJ_FRAME_BOUNDS | ORBM_PROJ_FRAME_UV | ORBM_PROJ_LEVEL_PLUS1);
    project(V, CurrentFrame.mvScaleFactors, P);
    FrameArrays fa(CurrentFrame);
    for (int k = 0; k < N; k++) fa.match[k] = CurrentFrame.mvpMapPoints[k] ? 0x40000000 : -1;      // "if(CurrentFrame.mvpMapPoints[i2]) continue;" (:1547-1548)
    int32_t nmatches = 0;
    check(orbm_search_by_projection(handle(), 1, bounds, fa.xy.data(), fa.octave.data(), fa.angle.data(), fa.desc(), &fa.n, fa.n, P.valid.data(), P.uv.data(),
                                    P.radius.data(), P.minl.data(), P.maxl.data(), qangle.data(), P.desc(), &P.n, P.n, ORBdist, 0.f, mbCheckOrientation ? 1 : 0,
                                    fa.match.data(), &nmatches, ORBS_MEM_HOST), "orbm_search_by_projection");
    for (int k = 0; k < N; k++) if (fa.match[k] >= 0 && fa.match[k] < P.n) CurrentFrame.mvpMapPoints[k] = vpMPs[fa.match[k]];
    return nmatches;
}

// LocalMapping::SearchInNeighbors -> matcher.Fuse(pKFi, vpMapPointMatches) / (mpCurrentKeyFrame, vpFuseCandidates)   (LocalMapping.cc:483 ff.)
int ORBmatcher::Fuse(KeyFrame *pKF, const std::vector<MapPoint *> &vpMapPoints, const float th)
{
    float Rcw[9], tcw[3], Ow[3];
    mat33(pKF->GetRotation(), Rcw); vec3(pKF->GetTranslation(), tcw); vec3(pKF->GetCameraCenter(), Ow);
    KeyFrameArrays K(pKF);
    PointArrays P(vpMapPoints.size());
    if (!P.n || !K.n) return 0;
    for (size_t i = 0; i < vpMapPoints.size(); i++) {
        MapPoint *pMP = vpMapPoints[i];
        if (!pMP || pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;
        P.set(i, pMP);
    }
    orbm_projection V;
    fill_view(V, Rcw, tcw, Ow, pKF->fx, pKF->fy, pKF->cx, pKF->cy, K.kf_bounds, pKF->mfLogScaleFactor, th, ORBM_PROJ_CHECK_NORMAL);
    project(V, pKF->mvScaleFactors, P);
    const std::vector<int32_t> best = best_in_window(K, P, TH_LOW, &pKF->mvInvLevelSigma2);
    // the search result of a point depends only on the point and the keyframe's features, neither of which the loop changes; the skip tests
    // and the map updates do depend on earlier iterations and run here in the reference's order (:840-850, :940-972)
    int nFused = 0;
    for (size_t i = 0; i < vpMapPoints.size(); i++) {
        MapPoint *pMP = vpMapPoints[i];
        if (!pMP || !P.valid[i] || best[i] < 0) continue;
        if (pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;
        MapPoint *pMPinKF = pKF->GetMapPoint(best[i]);
        if (pMPinKF) {
            if (!pMPinKF->isBad()) {
                if (pMPinKF->Observations() > pMP->Observations()) pMP->Replace(pMPinKF);
                else pMPinKF->Replace(pMP);
            }
        } else {
            pMP->AddObservation(pKF, best[i]);
            pKF->AddMapPoint(pMP, best[i]);
        }
        nFused++;
    }
    return nFused;
}

// LoopClosing::SearchAndFuse / MultiMapper -> matcher.Fuse(pKF, cvScw, mvpLoopMapPoints, 4, vpReplacePoints)   (LoopClosing.cc:603)
int ORBmatcher::Fuse(KeyFrame *pKF, cv::Mat Scw, const std::vector<MapPoint *> &vpPoints, float th, std::vector<MapPoint *> &vpReplacePoint)
{
    float Rcw[9], tcw[3], Ow[3];
    decompose_scw(Scw, Rcw, tcw, Ow);
    const std::set<MapPoint *> spAlreadyFound = pKF->GetMapPoints();
    KeyFrameArrays K(pKF);
    PointArrays P(vpPoints.size());
    if (!P.n || !K.n) return 0;
    for (size_t i = 0; i < vpPoints.size(); i++) {
        MapPoint *pMP = vpPoints[i];
        if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;
        P.set(i, pMP);
    }
    orbm_projection V;
    fill_view(V, Rcw, tcw, Ow, pKF->fx, pKF->fy, pKF->cx, pKF->cy, K.kf_bounds, pKF->mfLogScaleFactor, th, ORBM_PROJ_CHECK_NORMAL);
    project(V, pKF->mvScaleFactors, P);
    const std::vector<int32_t> best = best_in_window(K, P, TH_LOW, nullptr);
    int nFused = 0;
    for (size_t i = 0; i < vpPoints.size(); i++) {
        if (!P.valid[i] || best[i] < 0) continue;
        MapPoint *pMP = vpPoints[i];
        MapPoint *pMPinKF = pKF->GetMapPoint(best[i]);
        if (pMPinKF) { if (!pMPinKF->isBad()) vpReplacePoint[i] = pMPinKF; }
        else { pMP->AddObservation(pKF, best[i]); pKF->AddMapPoint(pMP, best[i]); }
        nFused++;
    }
    return nFused;
}

// LoopClosing::ComputeSim3 / MultiMapper -> matcher.SearchBySim3(mpCurrentKF, pKF, vpMapPointMatches, s, R, t, 7.5)   (LoopClosing.cc:332 ff.)
int ORBmatcher::SearchBySim3(KeyFrame *pKF1, KeyFrame *pKF2, std::vector<MapPoint *> &vpMatches12, const float &s12, const cv::Mat &R12, const cv::Mat &t12,
                             const float th)
{
    float R1w[9], t1w[3], R2w[9], t2w[3], r12[9], T12[3], sR12[9], sR21[9], t21[3];
    mat33(pKF1->GetRotation(), R1w); vec3(pKF1->GetTranslation(), t1w); mat33(pKF2->GetRotation(), R2w); vec3(pKF2->GetTranslation(), t2w);
    mat33(R12, r12); vec3(t12, T12);
    const double s = (double)s12, is = 1.0 / s12;
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { sR12[3 * r + c] = (float)((double)r12[3 * r + c] * s); sR21[3 * r + c] = (float)((double)r12[3 * c + r] * is); }
    for (int r = 0; r < 3; r++) {          // t21 = -sR21 * t12
        float a = -sR21[3 * r] * T12[0];
        a = a + -sR21[3 * r + 1] * T12[1];
        a = a + -sR21[3 * r + 2] * T12[2];
        t21[r] = a;
    }
    const std::vector<MapPoint *> vpMapPoints1 = pKF1->GetMapPointMatches(), vpMapPoints2 = pKF2->GetMapPointMatches();
    const int N1 = (int)vpMapPoints1.size(), N2 = (int)vpMapPoints2.size();
    std::vector<bool> vbAlreadyMatched1(N1, false), vbAlreadyMatched2(N2, false);
    for (int i = 0; i < N1; i++) {
        MapPoint *pMP = vpMatches12[i];
        if (pMP) {
            vbAlreadyMatched1[i] = true;
            const int idx2 = pMP->GetIndexInKeyFrame(pKF2);
            if (idx2 >= 0 && idx2 < N2) vbAlreadyMatched2[idx2] = true;
        }
    }
    KeyFrameArrays K1(pKF1), K2(pKF2);
    PointArrays P1(N1), P2(N2);
    for (int i = 0; i < N1; i++) { MapPoint *p = vpMapPoints1[i]; if (p && !vbAlreadyMatched1[i] && !p->isBad()) P1.set(i, p); }
    for (int i = 0; i < N2; i++) { MapPoint *p = vpMapPoints2[i]; if (p && !vbAlreadyMatched2[i] && !p->isBad()) P2.set(i, p); }
    orbm_projection V;
    const int flags = ORBM_PROJ_TWO_STEP | ORBM_PROJ_DIST_CAMERA;
    fill_view(V, R1w, t1w, nullptr, pKF1->fx, pKF1->fy, pKF1->cx, pKF1->cy, K2.kf_bounds, pKF2->mfLogScaleFactor, th, flags);
    std::memcpy(V.R2, sR21, sizeof V.R2); std::memcpy(V.t2, t21, sizeof V.t2);
    project(V, pKF2->mvScaleFactors, P1);
    const std::vector<int32_t> vnMatch1 = best_in_window(K2, P1, TH_HIGH, nullptr);
    fill_view(V, R2w, t2w, nullptr, pKF1->fx, pKF1->fy, pKF1->cx, pKF1->cy, K1.kf_bounds, pKF1->mfLogScaleFactor, th, flags);
    std::memcpy(V.R2, sR12, sizeof V.R2); std::memcpy(V.t2, T12, sizeof V.t2);
    project(V, pKF1->mvScaleFactors, P2);
    const std::vector<int32_t> vnMatch2 = best_in_window(K1, P2, TH_HIGH, nullptr);
    int nFound = 0;
    for (int i1 = 0; i1 < N1; i1++) {      // :1311-1325
        const int idx2 = vnMatch1[i1];
        if (idx2 >= 0 && vnMatch2[idx2] == i1) { vpMatches12[i1] = vpMapPoints2[idx2]; nFound++; }
    }
    return nFound;
}

// ---------------------------------------------------------------------------------------------------------------------------------------
// Vocabulary-bucket matchers and SearchForInitialization
namespace
{
struct FeatVecCSR {
    std::vector<int32_t> nodes, start, items; int32_t n = 0;
    FeatVecCSR(const DBoW2::FeatureVector &fv, size_t n_features)
    {
        items.reserve(n_features + 1);
        start.push_back(0);
        for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it) {
            nodes.push_back((int32_t)it->first);
            for (size_t k = 0; k < it->second.size(); k++) items.push_back((int32_t)it->second[k]);
            start.push_back((int32_t)items.size());
        }
        n = (int32_t)nodes.size();
        if (nodes.empty()) { nodes.push_back(0); start.push_back(0); }
        items.resize(n_features + 1, 0);
    }
};
struct DescAngles {
    std::vector<uint64_t> desc64; std::vector<float> angle; std::vector<uint8_t> elig; int32_t n = 0;
    DescAngles(const cv::Mat &D, const std::vector<cv::KeyPoint> &keys, int N) : desc64(4 * (size_t)N + 2), angle(N + 1, 0.f), elig(N + 1, 0), n(N)
    {
        for (int i = 0; i < N; i++) { std::memcpy(desc() + 32 * (size_t)i, D.ptr(i), 32); angle[i] = keys[i].angle; }
    }
    uint8_t *desc() { return (uint8_t *)(((uintptr_t)desc64.data() + 15) & ~(uintptr_t)15); }
};
int bow_call(int mode, DescAngles &A, FeatVecCSR &fa, DescAngles &B, FeatVecCSR &fb, float ratio, bool ori, const orbm_epipolar *epi, std::vector<int32_t> &match12)
{
    match12.assign(A.n + 1, -1);
    if (!A.n || !B.n || !fa.n || !fb.n) return 0;
    int32_t nmatches = 0;
    check(orbm_search_by_bow(handle(), 1, mode, A.desc(), A.angle.data(), A.elig.data(), &A.n, A.n, fa.nodes.data(), fa.start.data(), fa.items.data(), &fa.n,
                             (int)fa.nodes.size(), B.desc(), B.angle.data(), B.elig.data(), &B.n, B.n, fb.nodes.data(), fb.start.data(), fb.items.data(), &fb.n,
                             (int)fb.nodes.size(), ratio, ori ? 1 : 0, epi, match12.data(), &nmatches, ORBS_MEM_HOST), "orbm_search_by_bow");
    return nmatches;
}
}  // namespace

// Tracking::TrackReferenceKeyFrame / Relocalization -> matcher.SearchByBoW(mpReferenceKF, mCurrentFrame, vpMapPointMatches)   (Tracking.cc:809, 1415)
int ORBmatcher::SearchByBoW(KeyFrame *pKF, Frame &F, std::vector<MapPoint *> &vpMapPointMatches)
{
    const std::vector<MapPoint *> vpMapPointsKF = pKF->GetMapPointMatches();
    vpMapPointMatches = std::vector<MapPoint *>(F.N, static_cast<MapPoint *>(NULL));
    DescAngles A(pKF->mDescriptors, pKF->mvKeysUn, (int)vpMapPointsKF.size()), B(F.mDescriptors, F.mvKeys, F.N);      // kp.angle - F.mvKeys[..].angle (:243)
    for (int i = 0; i < A.n; i++) { MapPoint *p = vpMapPointsKF[i]; A.elig[i] = (p && !p->isBad()) ? 1 : 0; }
    for (int i = 0; i < B.n; i++) B.elig[i] = 1;
    FeatVecCSR fa(pKF->mFeatVec, A.n), fb(F.mFeatVec, B.n);
    std::vector<int32_t> m;
    const int n = bow_call(ORBM_BOW_MATCH, A, fa, B, fb, mfNNratio, mbCheckOrientation, nullptr, m);
    for (int i = 0; i < A.n; i++) if (m[i] >= 0) vpMapPointMatches[m[i]] = vpMapPointsKF[i];
    return n;
}

// LoopClosing::ComputeSim3 / MultiMapper -> matcher.SearchByBoW(mpCurrentKF, pKF, vvpMapPointMatches[i])   (LoopClosing.cc:245 ff.)
int ORBmatcher::SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, std::vector<MapPoint *> &vpMatches12)
{
    const std::vector<MapPoint *> vpMapPoints1 = pKF1->GetMapPointMatches(), vpMapPoints2 = pKF2->GetMapPointMatches();
    vpMatches12 = std::vector<MapPoint *>(vpMapPoints1.size(), static_cast<MapPoint *>(NULL));
    DescAngles A(pKF1->mDescriptors, pKF1->mvKeysUn, (int)vpMapPoints1.size()), B(pKF2->mDescriptors, pKF2->mvKeysUn, (int)vpMapPoints2.size());
    for (int i = 0; i < A.n; i++) { MapPoint *p = vpMapPoints1[i]; A.elig[i] = (p && !p->isBad()) ? 1 : 0; }
    for (int i = 0; i < B.n; i++) { MapPoint *p = vpMapPoints2[i]; B.elig[i] = (p && !p->isBad()) ? 1 : 0; }
    FeatVecCSR fa(pKF1->mFeatVec, A.n), fb(pKF2->mFeatVec, B.n);
    std::vector<int32_t> m;
    const int n = bow_call(ORBM_BOW_MATCH, A, fa, B, fb, mfNNratio, mbCheckOrientation, nullptr, m);
    for (int i = 0; i < A.n; i++) if (m[i] >= 0) vpMatches12[i] = vpMapPoints2[m[i]];
    return n;
}

// LocalMapping::CreateNewMapPoints -> matcher.SearchForTriangulation(mpCurrentKeyFrame, pKF2, F12, vMatchedIndices, false)   (LocalMapping.cc:215 ff.)
int ORBmatcher::SearchForTriangulation(KeyFrame *pKF1, KeyFrame *pKF2, cv::Mat F12, std::vector<std::pair<size_t, size_t>> &vMatchedPairs, const bool bOnlyStereo)
{
    if (bOnlyStereo) throw std::runtime_error("orbslamm_b200: only the monocular SearchForTriangulation is accelerated");
    // epipole of camera 1 in image 2 (:665-673), cv::Mat algebra in OpenCV's order
    float R2w[9], t2w[3], Cw[3], C2[3];
    mat33(pKF2->GetRotation(), R2w); vec3(pKF2->GetTranslation(), t2w); vec3(pKF1->GetCameraCenter(), Cw);
    for (int r = 0; r < 3; r++) {
        float s = R2w[3 * r] * Cw[0];
        s = s + R2w[3 * r + 1] * Cw[1];
        s = s + R2w[3 * r + 2] * Cw[2];
        C2[r] = s + t2w[r];
    }
    const float invz = 1.0f / C2[2];
    const float epipole[2] = {pKF2->fx * C2[0] * invz + pKF2->cx, pKF2->fy * C2[1] * invz + pKF2->cy};
    DescAngles A(pKF1->mDescriptors, pKF1->mvKeysUn, pKF1->N), B(pKF2->mDescriptors, pKF2->mvKeysUn, pKF2->N);
    std::vector<float> xy1(2 * (size_t)A.n + 2), xy2(2 * (size_t)B.n + 2); std::vector<int32_t> oct2(B.n + 1);
    for (int i = 0; i < A.n; i++) {
        if (!(pKF1->mvuRight[i] < 0)) throw std::runtime_error("orbslamm_b200: stereo keyframes are not supported");
        A.elig[i] = pKF1->GetMapPoint(i) ? 0 : 1; xy1[2 * i] = pKF1->mvKeysUn[i].pt.x; xy1[2 * i + 1] = pKF1->mvKeysUn[i].pt.y;
    }
    for (int i = 0; i < B.n; i++) {
        if (!(pKF2->mvuRight[i] < 0)) throw std::runtime_error("orbslamm_b200: stereo keyframes are not supported");
        B.elig[i] = pKF2->GetMapPoint(i) ? 0 : 1; xy2[2 * i] = pKF2->mvKeysUn[i].pt.x; xy2[2 * i + 1] = pKF2->mvKeysUn[i].pt.y; oct2[i] = pKF2->mvKeysUn[i].octave;
    }
    float F[9];
    mat33(F12, F);
    orbm_epipolar epi;
    epi.xy1 = xy1.data(); epi.xy2 = xy2.data(); epi.octave2 = oct2.data(); epi.F12 = F; epi.epipole = epipole;
    epi.scale_factors2 = pKF2->mvScaleFactors.data(); epi.level_sigma2_2 = pKF2->mvLevelSigma2.data(); epi.nlevels = (int32_t)pKF2->mvScaleFactors.size();
    FeatVecCSR fa(pKF1->mFeatVec, A.n), fb(pKF2->mFeatVec, B.n);
    std::vector<int32_t> m;
    const int n = bow_call(ORBM_BOW_TRIANGULATION, A, fa, B, fb, mfNNratio, mbCheckOrientation, &epi, m);
    vMatchedPairs.clear();
    vMatchedPairs.reserve(n);
    for (int i = 0; i < A.n; i++) if (m[i] >= 0) vMatchedPairs.push_back(std::make_pair((size_t)i, (size_t)m[i]));
    return n;
}

// Tracking::MonocularInitialization -> matcher.SearchForInitialization(mInitialFrame, mCurrentFrame, mvbPrevMatched, mvIniMatches, 100)   (Tracking.cc:639)
int ORBmatcher::SearchForInitialization(Frame &F1, Frame &F2, std::vector<cv::Point2f> &vbPrevMatched, std::vector<int> &vnMatches12, int windowSize)
{
    const int32_t N1 = (int32_t)F1.mvKeysUn.size(), N2 = (int32_t)F2.mvKeysUn.size();
    vnMatches12 = std::vector<int>(N1, -1);
    if (!N1 || !N2) return 0;
    DescAngles A(F1.mDescriptors, F1.mvKeysUn, N1), B(F2.mDescriptors, F2.mvKeysUn, N2);
    std::vector<int32_t> oct1(N1), oct2(N2), m(N1, -1);
    std::vector<float> xy2(2 * (size_t)N2), prev(2 * (size_t)N1);
    for (int i = 0; i < N1; i++) { oct1[i] = F1.mvKeysUn[i].octave; prev[2 * i] = vbPrevMatched[i].x; prev[2 * i + 1] = vbPrevMatched[i].y; }
    for (int i = 0; i < N2; i++) { oct2[i] = F2.mvKeysUn[i].octave; xy2[2 * i] = F2.mvKeysUn[i].pt.x; xy2[2 * i + 1] = F2.mvKeysUn[i].pt.y; }
    const float bounds[4] = {Frame::mnMinX, Frame::mnMinY, Frame::mnMaxX, Frame::mnMaxY};
    int32_t nmatches = 0;
    check(orbm_search_for_initialization(handle(), 1, bounds, oct1.data(), A.angle.data(), A.desc(), &N1, N1, xy2.data(), oct2.data(), B.angle.data(), B.desc(), &N2, N2,
                                         prev.data(), windowSize, mfNNratio, mbCheckOrientation ? 1 : 0, m.data(), &nmatches, ORBS_MEM_HOST),
          "orbm_search_for_initialization");
    for (int i = 0; i < N1; i++) { vnMatches12[i] = m[i]; vbPrevMatched[i].x = prev[2 * i]; vbPrevMatched[i].y = prev[2 * i + 1]; }
    return nmatches;
}

}  // namespace iORB_SLAM
