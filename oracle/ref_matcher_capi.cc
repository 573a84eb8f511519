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

}
