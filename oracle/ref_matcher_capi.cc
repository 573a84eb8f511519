/*
 * oracle/ref_matcher_capi.cc -- C entry points around the REFERENCE's own ORBmatcher class
 * (/root/reference/SingleRobotScenario/src/ORBmatcher.cc compiled unmodified against oracle/slamshim by oracle/Makefile into
 * oracle/_ref/libref_orbmatcher.so).  Frames and map points are rebuilt from flat arrays as the stand-in data model of
 * oracle/slamshim; the search loops, thresholds, rotation histogram, ComputeThreeMaxima and DescriptorDistance run as the
 * reference's object code.
 *
 * TEST INFRASTRUCTURE ONLY (tests/test_oracle_vs_reference.py).
 */
#include <cstring>
#include <vector>
#include "ORBmatcher.h"

namespace iORB_SLAM {
float Frame::fx, Frame::fy, Frame::cx, Frame::cy, Frame::invfx, Frame::invfy;
float Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY;
float Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv;
}
using namespace iORB_SLAM;

static void fill_frame(Frame &F, int N, const float *xy, const int *octave, const float *angle, const unsigned char *desc, const float *Tcw,
                       const float *scale_factors, int nlevels)
{
    F.N = N;
    F.mvKeys.resize(N); F.mvKeysUn.resize(N);
    for (int i = 0; i < N; i++) {
        cv::KeyPoint kp; kp.pt.x = xy[2 * i]; kp.pt.y = xy[2 * i + 1]; kp.octave = octave[i]; kp.angle = angle[i];
        F.mvKeys[i] = kp; F.mvKeysUn[i] = kp;                     /* no distortion: mvKeysUn = mvKeys (Frame.cc:406-410) */
    }
    F.mvuRight.assign(N, -1.f); F.mvDepth.assign(N, -1.f);
    F.mDescriptors = cv::Mat(N > 0 ? N : 1, 32, CV_8U);
    if (N > 0) std::memcpy(F.mDescriptors.data, desc, (size_t)32 * N);
    F.mvpMapPoints.assign(N, (MapPoint *)nullptr);
    F.mvbOutlier.assign(N, false);
    F.mTcw = cv::Mat(4, 4, CV_32F);
    std::memcpy(F.mTcw.data, Tcw, 16 * sizeof(float));
    F.mvScaleFactors.assign(scale_factors, scale_factors + nlevels);
    F.mfLogScaleFactor = std::log(scale_factors[1]);
    F.AssignFeaturesToGrid();
}

extern "C" {

/* Frame statics (Frame.cc:78-92): intrinsics, image bounds, grid cell sizes */
void ref_orbm_set_camera(float fx, float fy, float cx, float cy, float min_x, float min_y, float max_x, float max_y)
{
    Frame::fx = fx; Frame::fy = fy; Frame::cx = cx; Frame::cy = cy; Frame::invfx = 1.0f / fx; Frame::invfy = 1.0f / fy;
    Frame::mnMinX = min_x; Frame::mnMinY = min_y; Frame::mnMaxX = max_x; Frame::mnMaxY = max_y;
    Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(max_x - min_x);
    Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(max_y - min_y);
}

int ref_orbm_descriptor_distance(const unsigned char *a, const unsigned char *b)
{
    cv::Mat A(1, 32, CV_8U), B(1, 32, CV_8U);
    std::memcpy(A.data, a, 32); std::memcpy(B.data, b, 32);
    return ORBmatcher::DescriptorDistance(A, B);
}

/* ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono = true), ORBmatcher.cc:1330-1472.
 * last-frame slot i holds a map point iff valid[i]; feat_match[k] = last-frame slot matched to current feature k, or -1. */
int ref_orbm_search_last_frame(float nnratio, int check_ori, float th, const float *Tcw, const float *scale_factors, int nlevels,
                               int N, const float *f_xy, const int *f_octave, const float *f_angle, const unsigned char *f_desc,
                               int M, const unsigned char *valid, const float *Xw, const int *l_octave, const float *l_angle, const unsigned char *l_desc,
                               int *feat_match)
{
    Frame cur, last;
    fill_frame(cur, N, f_xy, f_octave, f_angle, f_desc, Tcw, scale_factors, nlevels);
    std::vector<float> lxy(2 * (size_t)(M > 0 ? M : 1), 0.f);
    const float eye[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    fill_frame(last, M, lxy.data(), l_octave, l_angle, l_desc, eye, scale_factors, nlevels);
    std::vector<MapPoint> mps(M);
    for (int i = 0; i < M; i++) {
        if (!valid[i]) continue;
        MapPoint &p = mps[i];
        p.mWorldPos = cv::Mat(3, 1, CV_32F);
        for (int k = 0; k < 3; k++) p.mWorldPos.at<float>(k) = Xw[3 * i + k];
        p.mDescriptor = cv::Mat(1, 32, CV_8U);
        std::memcpy(p.mDescriptor.data, l_desc + (size_t)32 * i, 32);
        last.mvpMapPoints[i] = &p;
    }
    ORBmatcher matcher(nnratio, check_ori != 0);
    const int n = matcher.SearchByProjection(cur, last, th, true);
    for (int k = 0; k < N; k++) feat_match[k] = cur.mvpMapPoints[k] ? (int)(cur.mvpMapPoints[k] - mps.data()) : -1;
    return n;
}

/* ORBmatcher::SearchByProjection(F, vpMapPoints, th), ORBmatcher.cc:45-129 (local-map tracking).  Map point i is a candidate iff
 * in_view[i] (mbTrackInView, as set by Frame::isInFrustum); held[k] != 0: feature k already holds a map point with observations. */
int ref_orbm_search_local_points(float nnratio, float th, const float *scale_factors, int nlevels,
                                 int N, const float *f_xy, const int *f_octave, const float *f_angle, const unsigned char *f_desc, const unsigned char *held,
                                 int M, const unsigned char *in_view, const float *proj_xy, const int *level, const float *view_cos, const unsigned char *q_desc,
                                 int *feat_match)
{
    Frame F;
    const float eye[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    fill_frame(F, N, f_xy, f_octave, f_angle, f_desc, eye, scale_factors, nlevels);
    MapPoint holder;                       /* a map point with Observations() > 0 for the features that are already taken */
    for (int k = 0; k < N; k++) if (held && held[k]) F.mvpMapPoints[k] = &holder;
    std::vector<MapPoint> mps(M);
    std::vector<MapPoint *> vp(M);
    for (int i = 0; i < M; i++) {
        MapPoint &p = mps[i];
        p.mbTrackInView = in_view[i] != 0;
        p.mTrackProjX = proj_xy[2 * i]; p.mTrackProjY = proj_xy[2 * i + 1];
        p.mnTrackScaleLevel = level[i]; p.mTrackViewCos = view_cos[i];
        p.mDescriptor = cv::Mat(1, 32, CV_8U);
        std::memcpy(p.mDescriptor.data, q_desc + (size_t)32 * i, 32);
        vp[i] = &p;
    }
    ORBmatcher matcher(nnratio, true);
    const int n = matcher.SearchByProjection(F, vp, th);
    for (int k = 0; k < N; k++) {
        MapPoint *p = F.mvpMapPoints[k];
        feat_match[k] = (p && p != &holder) ? (int)(p - mps.data()) : (p == &holder ? -2 : -1);
    }
    return n;
}


/* ---- KeyFrame / Sim3 projection family ---------------------------------------------------------------------------------------------- */
struct RefKF {                /* flat description of a keyframe; grid filled with the Frame's float bounds, window offsets are the ints */
    int N; const float *xy; const int *octave; const float *angle; const unsigned char *desc;
    const float *scale_factors; const float *inv_level_sigma2; int nlevels;
    const float *K4; const float *grid_bounds4; const float *Tcw;
};

static void fill_keyframe(KeyFrame &K, const RefKF &a)
{
    K.N = a.N;
    K.mvKeys.resize(a.N); K.mvKeysUn.resize(a.N);
    for (int i = 0; i < a.N; i++) {
        cv::KeyPoint kp; kp.pt.x = a.xy[2 * i]; kp.pt.y = a.xy[2 * i + 1]; kp.octave = a.octave[i]; kp.angle = a.angle ? a.angle[i] : 0.f;
        K.mvKeys[i] = kp; K.mvKeysUn[i] = kp;
    }
    K.mvuRight.assign(a.N, -1.f); K.mvDepth.assign(a.N, -1.f);
    K.mDescriptors = cv::Mat(a.N > 0 ? a.N : 1, 32, CV_8U);
    if (a.N > 0) std::memcpy(K.mDescriptors.data, a.desc, (size_t)32 * a.N);
    K.mvpMapPoints.assign(a.N, (MapPoint *)nullptr);
    K.fx = a.K4[0]; K.fy = a.K4[1]; K.cx = a.K4[2]; K.cy = a.K4[3];
    K.mvScaleFactors.assign(a.scale_factors, a.scale_factors + a.nlevels);
    if (a.inv_level_sigma2) K.mvInvLevelSigma2.assign(a.inv_level_sigma2, a.inv_level_sigma2 + a.nlevels);
    K.mnScaleLevels = a.nlevels;
    K.mfLogScaleFactor = std::log(a.scale_factors[1]);
    /* KeyFrame::KeyFrame(Frame&): mnMinX(F.mnMinX) etc. truncate the Frame's float bounds to int; the cell sizes and mGrid are copied */
    const float *b = a.grid_bounds4;
    K.mnMinX = (int)b[0]; K.mnMinY = (int)b[1]; K.mnMaxX = (int)b[2]; K.mnMaxY = (int)b[3];
    K.mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(b[2] - b[0]);
    K.mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(b[3] - b[1]);
    K.mGrid.assign(FRAME_GRID_COLS, std::vector<std::vector<size_t>>(FRAME_GRID_ROWS));
    for (int i = 0; i < a.N; i++) {            /* Frame::AssignFeaturesToGrid / PosInGrid, Frame.cc:230-245, 382-392 */
        const int px = round((a.xy[2 * i] - b[0]) * K.mfGridElementWidthInv), py = round((a.xy[2 * i + 1] - b[1]) * K.mfGridElementHeightInv);
        if (px < 0 || px >= FRAME_GRID_COLS || py < 0 || py >= FRAME_GRID_ROWS) continue;
        K.mGrid[px][py].push_back(i);
    }
    if (a.Tcw) {
        K.Tcw = cv::Mat(4, 4, CV_32F);
        std::memcpy(K.Tcw.data, a.Tcw, 16 * sizeof(float));
        cv::Mat Rcw = K.Tcw.rowRange(0, 3).colRange(0, 3), tcw = K.Tcw.rowRange(0, 3).col(3);
        K.Ow = -Rcw.t() * tcw;                  /* KeyFrame::SetPose, KeyFrame.cc:72-76 */
    }
}

static void fill_points(std::vector<MapPoint> &mps, int M, const float *Xw, const float *normal, const float *mf_min, const float *mf_max,
                        const unsigned char *desc)
{
    mps.assign(M, MapPoint());
    for (int i = 0; i < M; i++) {
        MapPoint &p = mps[i];
        p.mWorldPos = cv::Mat(3, 1, CV_32F); p.mNormalVector = cv::Mat(3, 1, CV_32F);
        for (int k = 0; k < 3; k++) { p.mWorldPos.at<float>(k) = Xw[3 * i + k]; p.mNormalVector.at<float>(k) = normal ? normal[3 * i + k] : 0.f; }
        p.mfMinDistance = mf_min[i]; p.mfMaxDistance = mf_max[i];
        p.mDescriptor = cv::Mat(1, 32, CV_8U);
        std::memcpy(p.mDescriptor.data, desc + (size_t)32 * i, 32);
    }
}

static cv::Mat mat44(const float *T) { cv::Mat m(4, 4, CV_32F); std::memcpy(m.data, T, 16 * sizeof(float)); return m; }

/* ORBmatcher::SearchByProjection(pKF, Scw, vpPoints, vpMatched, th), ORBmatcher.cc:292-405.  skip[i] != 0: the point is bad or already in
 * vpMatched; held[k] != 0: vpMatched[k] is non-null on entry.  feat_match[k] = index of the point now matched to feature k, -2 = held, -1 none. */
int ref_orbm_search_kf_sim3(const RefKF *kf, const float *Scw, int th, int M, const unsigned char *skip, const float *Xw, const float *normal,
                            const float *mf_min, const float *mf_max, const unsigned char *q_desc, const unsigned char *held, int *feat_match)
{
    KeyFrame K; fill_keyframe(K, *kf);
    std::vector<MapPoint> mps; fill_points(mps, M, Xw, normal, mf_min, mf_max, q_desc);
    std::vector<MapPoint *> vp(M);
    for (int i = 0; i < M; i++) { vp[i] = &mps[i]; mps[i].bad = skip[i] != 0; }
    MapPoint holder;
    std::vector<MapPoint *> matched(K.N, (MapPoint *)nullptr);
    for (int k = 0; k < K.N; k++) if (held && held[k]) matched[k] = &holder;
    ORBmatcher matcher(0.75f, true);
    const int n = matcher.SearchByProjection(&K, mat44(Scw), vp, matched, th);
    for (int k = 0; k < K.N; k++) feat_match[k] = matched[k] == &holder ? -2 : (matched[k] ? (int)(matched[k] - mps.data()) : -1);
    return n;
}

/* decode what Fuse did with point i: the KeyFrame slot it chose */
static int fused_slot(std::vector<MapPoint> &mps, std::vector<MapPoint> &holders, KeyFrame *pKF, MapPoint *p, std::vector<MapPoint *> *replace, int i)
{
    for (int hop = 0; hop < 8 && p; hop++) {
        if (replace && (*replace)[i]) { MapPoint *r = (*replace)[i]; if (r >= holders.data() && r < holders.data() + holders.size()) return (int)(r - holders.data()); p = r; replace = nullptr; continue; }
        if (p->mpReplaced) { MapPoint *r = p->mpReplaced; if (r >= holders.data() && r < holders.data() + holders.size()) return (int)(r - holders.data()); p = r; continue; }
        return p->GetIndexInKeyFrame(pKF);
    }
    return -1;
}

/* ORBmatcher::Fuse(pKF, vpMapPoints, th), ORBmatcher.cc:827-977.  skip[i]: NULL / bad / already in the keyframe.  occupied[k]: the
 * keyframe slot holds a (well observed) map point -> the reference replaces; else it adds the observation.  slot[i] = chosen slot or -1. */
int ref_orbm_fuse_kf(const RefKF *kf, float th, int M, const unsigned char *skip, const float *Xw, const float *normal, const float *mf_min,
                     const float *mf_max, const unsigned char *q_desc, const unsigned char *occupied, int *slot)
{
    KeyFrame K; fill_keyframe(K, *kf);
    std::vector<MapPoint> mps; fill_points(mps, M, Xw, normal, mf_min, mf_max, q_desc);
    std::vector<MapPoint> holders(K.N);
    for (int k = 0; k < K.N; k++) { holders[k].nObs = 1000000; if (occupied && occupied[k]) K.mvpMapPoints[k] = &holders[k]; }
    std::vector<MapPoint *> vp(M);
    for (int i = 0; i < M; i++) vp[i] = skip[i] ? (MapPoint *)nullptr : &mps[i];
    ORBmatcher matcher(0.6f, true);
    const int n = matcher.Fuse(&K, vp, th);
    for (int i = 0; i < M; i++) slot[i] = skip[i] ? -1 : fused_slot(mps, holders, &K, &mps[i], nullptr, i);
    return n;
}

/* ORBmatcher::Fuse(pKF, Scw, vpPoints, th, vpReplacePoint), ORBmatcher.cc:979-1102 */
int ref_orbm_fuse_sim3(const RefKF *kf, const float *Scw, float th, int M, const unsigned char *skip, const float *Xw, const float *normal,
                       const float *mf_min, const float *mf_max, const unsigned char *q_desc, const unsigned char *occupied, int *slot)
{
    KeyFrame K; fill_keyframe(K, *kf);
    std::vector<MapPoint> mps; fill_points(mps, M, Xw, normal, mf_min, mf_max, q_desc);
    std::vector<MapPoint> holders(K.N);
    for (int k = 0; k < K.N; k++) if (occupied && occupied[k]) K.mvpMapPoints[k] = &holders[k];
    std::vector<MapPoint *> vp(M), replace(M, (MapPoint *)nullptr);
    for (int i = 0; i < M; i++) { vp[i] = &mps[i]; mps[i].bad = skip[i] != 0; }
    ORBmatcher matcher(0.8f, true);
    const int n = matcher.Fuse(&K, mat44(Scw), vp, th, replace);
    for (int i = 0; i < M; i++) slot[i] = skip[i] ? -1 : fused_slot(mps, holders, &K, &mps[i], &replace, i);
    return n;
}

/* ORBmatcher::SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th), ORBmatcher.cc:1104-1328.  has1[i] / has2[i]: the keyframe slot
 * holds a (good) map point, described by the X / mf / desc arrays at the same index.  matches12[N1] in: index into KF2 of an existing
 * match or -1; out: the same, updated. */
int ref_orbm_search_by_sim3(const RefKF *kf1, const RefKF *kf2, float s12, const float *R12, const float *t12, float th,
                            const unsigned char *has1, const float *X1, const float *mfmin1, const float *mfmax1, const unsigned char *d1,
                            const unsigned char *has2, const float *X2, const float *mfmin2, const float *mfmax2, const unsigned char *d2,
                            int *matches12)
{
    KeyFrame K1, K2; fill_keyframe(K1, *kf1); fill_keyframe(K2, *kf2);
    std::vector<MapPoint> p1, p2;
    fill_points(p1, K1.N, X1, nullptr, mfmin1, mfmax1, d1); fill_points(p2, K2.N, X2, nullptr, mfmin2, mfmax2, d2);
    for (int i = 0; i < K1.N; i++) if (has1[i]) { K1.mvpMapPoints[i] = &p1[i]; p1[i].mObservations[&K1] = i; }
    for (int i = 0; i < K2.N; i++) if (has2[i]) { K2.mvpMapPoints[i] = &p2[i]; p2[i].mObservations[&K2] = i; }
    std::vector<MapPoint *> m12(K1.N, (MapPoint *)nullptr);
    for (int i = 0; i < K1.N; i++) if (matches12[i] >= 0) m12[i] = &p2[matches12[i]];
    cv::Mat R(3, 3, CV_32F), t(3, 1, CV_32F);
    std::memcpy(R.data, R12, 9 * sizeof(float)); std::memcpy(t.data, t12, 3 * sizeof(float));
    ORBmatcher matcher(0.75f, true);
    const int n = matcher.SearchBySim3(&K1, &K2, m12, s12, R, t, th);
    for (int i = 0; i < K1.N; i++) matches12[i] = m12[i] ? (int)(m12[i] - p2.data()) : -1;
    return n;
}

/* ORBmatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist), ORBmatcher.cc:1474-1601 (relocalisation).  The keyframe
 * contributes its map points (has[i], skip[i] = bad or in sAlreadyFound) and keypoint angles; held[k]: CurrentFrame.mvpMapPoints[k] is set. */
int ref_orbm_search_frame_kf(int check_ori, float th, int orb_dist, const float *Tcw, const float *scale_factors, int nlevels,
                             int N, const float *f_xy, const int *f_octave, const float *f_angle, const unsigned char *f_desc, const unsigned char *held,
                             int M, const unsigned char *has, const unsigned char *skip, const float *Xw, const float *mf_min, const float *mf_max,
                             const float *kf_angle, const unsigned char *q_desc, int *feat_match)
{
    Frame F;
    fill_frame(F, N, f_xy, f_octave, f_angle, f_desc, Tcw, scale_factors, nlevels);
    MapPoint holder;
    for (int k = 0; k < N; k++) if (held && held[k]) F.mvpMapPoints[k] = &holder;
    KeyFrame K;
    K.N = M; K.mvKeysUn.resize(M); K.mvpMapPoints.assign(M, (MapPoint *)nullptr);
    std::vector<MapPoint> mps; fill_points(mps, M, Xw, nullptr, mf_min, mf_max, q_desc);
    std::set<MapPoint *> found;
    for (int i = 0; i < M; i++) {
        K.mvKeysUn[i].angle = kf_angle[i];
        if (has[i]) K.mvpMapPoints[i] = &mps[i];
        if (skip[i]) found.insert(&mps[i]);
    }
    ORBmatcher matcher(0.9f, check_ori != 0);
    const int n = matcher.SearchByProjection(F, &K, found, th, orb_dist);
    for (int k = 0; k < N; k++) { MapPoint *p = F.mvpMapPoints[k]; feat_match[k] = p == &holder ? -2 : (p ? (int)(p - mps.data()) : -1); }
    return n;
}


/* ---- vocabulary-bucket matchers and SearchForInitialization -------------------------------------------------------------------------- */
static void fill_featvec(DBoW2::FeatureVector &fv, int n_nodes, const int *nodes, const int *start, const int *items)
{
    for (int a = 0; a < n_nodes; a++) {
        std::vector<unsigned int> &v = fv[(unsigned)nodes[a]];
        for (int u = start[a]; u < start[a + 1]; u++) v.push_back((unsigned)items[u]);
    }
}

/* ORBmatcher::SearchByBoW(pKF, F, vpMapPointMatches), ORBmatcher.cc:159-290.  has1[i]: keyframe feature i holds a good map point.
 * match[k] (per FRAME feature) = keyframe feature whose map point was assigned, or -1. */
int ref_orbm_search_by_bow_kf_frame(float nnratio, int check_ori, const RefKF *kf, const unsigned char *has1, int nn1, const int *nodes1, const int *start1,
                                    const int *items1, int N2, const float *f_angle, const unsigned char *f_desc, int nn2, const int *nodes2, const int *start2,
                                    const int *items2, int *match)
{
    KeyFrame K; fill_keyframe(K, *kf);
    fill_featvec(K.mFeatVec, nn1, nodes1, start1, items1);
    std::vector<MapPoint> mps(K.N);
    for (int i = 0; i < K.N; i++) if (has1[i]) K.mvpMapPoints[i] = &mps[i];
    Frame F;
    std::vector<float> xy(2 * (size_t)(N2 > 0 ? N2 : 1), 0.f); std::vector<int> oct(N2 > 0 ? N2 : 1, 0);
    const float eye[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    fill_frame(F, N2, xy.data(), oct.data(), f_angle, f_desc, eye, kf->scale_factors, kf->nlevels);
    fill_featvec(F.mFeatVec, nn2, nodes2, start2, items2);
    std::vector<MapPoint *> out;
    ORBmatcher matcher(nnratio, check_ori != 0);
    const int n = matcher.SearchByBoW(&K, F, out);
    for (int k = 0; k < N2; k++) match[k] = out[k] ? (int)(out[k] - mps.data()) : -1;
    return n;
}

/* ORBmatcher::SearchByBoW(pKF1, pKF2, vpMatches12), ORBmatcher.cc:524-657.  match12[i1] = feature of KF2 whose map point was matched, or -1. */
int ref_orbm_search_by_bow_kf_kf(float nnratio, int check_ori, const RefKF *kf1, const unsigned char *has1, int nn1, const int *nodes1, const int *start1,
                                 const int *items1, const RefKF *kf2, const unsigned char *has2, int nn2, const int *nodes2, const int *start2,
                                 const int *items2, int *match12)
{
    KeyFrame K1, K2; fill_keyframe(K1, *kf1); fill_keyframe(K2, *kf2);
    fill_featvec(K1.mFeatVec, nn1, nodes1, start1, items1); fill_featvec(K2.mFeatVec, nn2, nodes2, start2, items2);
    std::vector<MapPoint> p1(K1.N), p2(K2.N);
    for (int i = 0; i < K1.N; i++) if (has1[i]) K1.mvpMapPoints[i] = &p1[i];
    for (int i = 0; i < K2.N; i++) if (has2[i]) K2.mvpMapPoints[i] = &p2[i];
    std::vector<MapPoint *> out;
    ORBmatcher matcher(nnratio, check_ori != 0);
    const int n = matcher.SearchByBoW(&K1, &K2, out);
    for (int i = 0; i < K1.N; i++) match12[i] = out[i] ? (int)(out[i] - p2.data()) : -1;
    return n;
}

/* ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo = false), ORBmatcher.cc:659-825 (monocular keyframes).
 * has1 / has2: the feature already holds a map point (skipped).  match12[i1] = i2 or -1 (the pairs). */
int ref_orbm_search_for_triangulation(float nnratio, int check_ori, const RefKF *kf1, const unsigned char *has1, int nn1, const int *nodes1, const int *start1,
                                      const int *items1, const RefKF *kf2, const unsigned char *has2, int nn2, const int *nodes2, const int *start2,
                                      const int *items2, const float *level_sigma2, const float *F12, int *match12)
{
    KeyFrame K1, K2; fill_keyframe(K1, *kf1); fill_keyframe(K2, *kf2);
    K2.mvLevelSigma2.assign(level_sigma2, level_sigma2 + kf2->nlevels); K1.mvLevelSigma2 = K2.mvLevelSigma2;
    fill_featvec(K1.mFeatVec, nn1, nodes1, start1, items1); fill_featvec(K2.mFeatVec, nn2, nodes2, start2, items2);
    MapPoint holder;
    for (int i = 0; i < K1.N; i++) if (has1[i]) K1.mvpMapPoints[i] = &holder;
    for (int i = 0; i < K2.N; i++) if (has2[i]) K2.mvpMapPoints[i] = &holder;
    cv::Mat F(3, 3, CV_32F);
    std::memcpy(F.data, F12, 9 * sizeof(float));
    std::vector<std::pair<size_t, size_t>> pairs;
    ORBmatcher matcher(nnratio, check_ori != 0);
    const int n = matcher.SearchForTriangulation(&K1, &K2, F, pairs, false);
    for (int i = 0; i < K1.N; i++) match12[i] = -1;
    for (auto &pr : pairs) match12[pr.first] = (int)pr.second;
    return n;
}

/* ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize), ORBmatcher.cc:407-522 */
int ref_orbm_search_for_initialization(float nnratio, int check_ori, int window, const float *scale_factors, int nlevels,
                                       int N1, const float *xy1, const int *oct1, const float *ang1, const unsigned char *desc1,
                                       int N2, const float *xy2, const int *oct2, const float *ang2, const unsigned char *desc2,
                                       float *prev_matched, int *matches12)
{
    Frame F1, F2;
    const float eye[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    fill_frame(F1, N1, xy1, oct1, ang1, desc1, eye, scale_factors, nlevels);
    fill_frame(F2, N2, xy2, oct2, ang2, desc2, eye, scale_factors, nlevels);
    std::vector<cv::Point2f> prev(N1);
    for (int i = 0; i < N1; i++) { prev[i].x = prev_matched[2 * i]; prev[i].y = prev_matched[2 * i + 1]; }
    std::vector<int> m12;
    ORBmatcher matcher(nnratio, check_ori != 0);
    const int n = matcher.SearchForInitialization(F1, F2, prev, m12, window);
    for (int i = 0; i < N1; i++) { matches12[i] = m12[i]; prev_matched[2 * i] = prev[i].x; prev_matched[2 * i + 1] = prev[i].y; }
    return n;
}

}
