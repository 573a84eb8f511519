// ORBmatcher_b200.cc -- GPU-backed definitions of the ORBmatcher members on the tracking hot path.  In the
// reference tree these replace the same-named functions of S/src/ORBmatcher.cc (:41-43 ctor, :45-129 and :1330-1472
// SearchByProjection, :131-137 RadiusByViewingCos, :1649-1665 DescriptorDistance); the remaining members
// (SearchByBoW, SearchForInitialization, SearchForTriangulation, SearchBySim3, Fuse, the KeyFrame/Sim3 projection
// variants) stay in ORBmatcher.cc -- they are "next" rows of SURVEY.md 8(f).
//
// The shim only marshals: Frame/MapPoint fields -> flat arrays -> orbm_* (include/orbslamm_b200.h) -> pointers written
// back into Frame::mvpMapPoints.  ORBmatcher objects are stack-constructed from several threads in the reference, so
// the device workspace is a thread_local handle.
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

}  // namespace iORB_SLAM
