// oracle/ref_optimizer_capi.cc -- TEST INFRASTRUCTURE ONLY.  A flat C API over the reference's OWN Optimizer.cc / Converter.cc and its vendored g2o
// (all compiled unmodified from /root/reference into oracle/_ref/libref_optimizer.so, against oracle/eigenshim + oracle/optshim + oracle/slamshim),
// so that tests can pin oracle/slam_oracle.c and the CUDA path against the reference's object code:
//   ref_opt_*  build the stand-in map (KeyFrame / MapPoint / Map / Frame of orbslamm_b200/host/mock) from flat arrays and call
//              Optimizer::PoseOptimization / LocalBundleAdjustment / BundleAdjustment / OptimizeSim3 / OptimizeEssentialGraph;
//   ref_g2o_*  single g2o primitives (SE3Quat::exp, Sim3 log / exp / inverse / product, EdgeSE3ProjectXYZ(OnlyPose) error + Jacobians, EdgeSim3 error
//              + numeric Jacobians, EdgeSim3ProjectXYZ / EdgeInverseSim3ProjectXYZ, RobustKernelHuber) for random-input comparisons.
#include <cstring>
#include <map>
#include <set>
#include <vector>
#include "Optimizer.h"
#include "Converter.h"
#include "Thirdparty/g2o/g2o/core/robust_kernel_impl.h"
#include "Thirdparty/g2o/g2o/core/jacobian_workspace.h"
#include "Thirdparty/g2o/g2o/core/block_solver.h"
#include "Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.h"
#include "Thirdparty/g2o/g2o/solvers/linear_solver_eigen.h"
#include "Thirdparty/g2o/g2o/types/types_six_dof_expmap.h"
#include "Thirdparty/g2o/g2o/types/types_seven_dof_expmap.h"

namespace iORB_SLAM
{
std::mutex MapPoint::mGlobalMutex;
float Frame::fx, Frame::fy, Frame::cx, Frame::cy;
float Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY;
void MapPoint::Replace(MapPoint *) {}
}  // namespace iORB_SLAM

using namespace iORB_SLAM;

namespace {
void set_pose(cv::Mat &T, const float *p) { for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) T.at<float>(r, c) = p[4 * r + c]; }
void get_pose(const cv::Mat &T, float *p) { for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) p[4 * r + c] = T.at<float>(r, c); }

// the mock map of a flat BA graph.  fixed[k]: 0 free (local keyframe), 1 mnId == 0 (fixed by the reference's rule, local), 2 fixed camera (outside
// the covisibility of the current keyframe).  The per-edge information weight travels through the level table: the edge's keypoint gets octave = its
// slot in the keyframe and mvInvLevelSigma2[slot] = the weight (Optimizer.cc reads pKF->mvInvLevelSigma2[kpUn.octave]).
struct BaMap {
    std::vector<KeyFrame> kfs; std::vector<MapPoint> pts; Map map;
    BaMap(int K, const float *poses, const uint8_t *fixed, const double *intr, int P, const float *points, int E, const int *e_kf, const int *e_pt,
          const float *e_uv, const float *e_w)
        : kfs(K), pts(P)
    {
        long unsigned int next_id = 1;
        for (int k = 0; k < K; k++) {
            KeyFrame &kf = kfs[k];
            kf.mnId = fixed[k] == 1 ? 0 : next_id++;
            kf.fx = (float)intr[4 * k]; kf.fy = (float)intr[4 * k + 1]; kf.cx = (float)intr[4 * k + 2]; kf.cy = (float)intr[4 * k + 3];
            set_pose(kf.Tcw, poses + 16 * k);
            map.mvKFs.push_back(&kf);
        }
        for (int p = 0; p < P; p++) { pts[p].mnId = p; pts[p].nObs = 0; for (int c = 0; c < 3; c++) pts[p].mWorldPos.at<float>(c) = points[3 * p + c]; map.mvMPs.push_back(&pts[p]); }
        for (int e = 0; e < E; e++) {
            KeyFrame &kf = kfs[e_kf[e]];
            const size_t idx = kf.mvKeysUn.size();
            cv::KeyPoint kp; kp.pt.x = e_uv[2 * e]; kp.pt.y = e_uv[2 * e + 1]; kp.octave = (int)idx;
            kf.mvKeysUn.push_back(kp); kf.mvuRight.push_back(-1.f); kf.mvInvLevelSigma2.push_back(e_w[e]);
            kf.mvpMapPoints.push_back(&pts[e_pt[e]]);
            pts[e_pt[e]].mObservations[&kf] = idx; pts[e_pt[e]].nObs++;
        }
    }
    void read(float *poses, float *points, int32_t *nobs)
    {
        for (size_t k = 0; k < kfs.size(); k++) get_pose(kfs[k].Tcw, poses + 16 * k);
        for (size_t p = 0; p < pts.size(); p++) { for (int c = 0; c < 3; c++) points[3 * p + c] = pts[p].mWorldPos.at<float>(c); if (nobs) nobs[p] = pts[p].nObs; }
    }
};
}  // namespace

extern "C" {

// Optimizer::PoseOptimization on one frame: N correspondences (Xw, uv, invSigma2).  Returns nInitialCorrespondences - nBad; Tcw in/out, outlier out.
int ref_opt_pose_optimization(float *Tcw, const float *K4, int N, const float *Xw, const float *uv, const float *inv_sigma2, uint8_t *outlier)
{
    Frame F;
    Frame::fx = K4[0]; Frame::fy = K4[1]; Frame::cx = K4[2]; Frame::cy = K4[3];
    std::vector<MapPoint> mps(N);
    F.N = N; F.mvKeysUn.resize(N); F.mvuRight.assign(N, -1.f); F.mvpMapPoints.assign(N, nullptr); F.mvbOutlier.assign(N, false); F.mvInvLevelSigma2.resize(N);
    for (int i = 0; i < N; i++) {
        for (int c = 0; c < 3; c++) mps[i].mWorldPos.at<float>(c) = Xw[3 * i + c];
        F.mvKeysUn[i].pt.x = uv[2 * i]; F.mvKeysUn[i].pt.y = uv[2 * i + 1]; F.mvKeysUn[i].octave = i; F.mvInvLevelSigma2[i] = inv_sigma2[i];
        F.mvpMapPoints[i] = &mps[i];
    }
    set_pose(F.mTcw, Tcw);
    const int n = Optimizer::PoseOptimization(&F);
    get_pose(F.mTcw, Tcw);
    for (int i = 0; i < N; i++) outlier[i] = F.mvbOutlier[i] ? 1 : 0;
    return n;
}

// Optimizer::LocalBundleAdjustment.  The current keyframe is the last free one; every other keyframe with fixed != 2 is in its covisibility list.
// nobs (optional): observations left per point after the reference erased the outlier observations (Optimizer.cc:771-780).
void ref_opt_local_ba(int K, float *poses, const uint8_t *fixed, const double *intr, int P, float *points, int E, const int *e_kf, const int *e_pt,
                      const float *e_uv, const float *e_w, int32_t *nobs)
{
    BaMap M(K, poses, fixed, intr, P, points, E, e_kf, e_pt, e_uv, e_w);
    int cur = -1;
    for (int k = 0; k < K; k++) if (fixed[k] == 0) cur = k;
    if (cur < 0) return;
    for (int k = 0; k < K; k++) if (k != cur && fixed[k] != 2) M.kfs[cur].mvpOrderedConnectedKeyFrames.push_back(&M.kfs[k]);
    bool stop = false;
    Optimizer::LocalBundleAdjustment(&M.kfs[cur], &stop, &M.map);
    M.read(poses, points, nobs);
}

// Optimizer::BundleAdjustment (GlobalBundleAdjustemnt's core) with nLoopKF = 0: results written to the keyframes / points.  Only the mnId == 0 keyframe is
// fixed by the reference; fixed == 2 keyframes are not representable here and must not be passed.
void ref_opt_bundle_adjust(int K, float *poses, const uint8_t *fixed, const double *intr, int P, float *points, int E, const int *e_kf, const int *e_pt,
                           const float *e_uv, const float *e_w, int n_iterations, int robust)
{
    BaMap M(K, poses, fixed, intr, P, points, E, e_kf, e_pt, e_uv, e_w);
    std::vector<KeyFrame *> vk; std::vector<MapPoint *> vp;
    for (auto &k : M.kfs) vk.push_back(&k);
    for (auto &p : M.pts) vp.push_back(&p);
    bool stop = false;
    Optimizer::BundleAdjustment(vk, vp, n_iterations, &stop, 0, robust != 0);
    M.read(poses, points, nullptr);
}

// Optimizer::OptimizeSim3 for one keyframe pair.  T1 / T2: poses of the two keyframes (f32[16]); K1 / K2: fx fy cx cy; N matched features of KF1:
// Xw1[i] = world position of KF1's map point i, Xw2[i] = world position of the matched map point (vpMatches1[i]) or valid[i] = 0; uv1 / uv2 + level
// inverse sigma2 (information weights) of the keypoints; sim3 (r xyzw, t, s) in/out; inlier[i] out (vpMatches1[i] kept).  Returns the reference's return value (nIn).
int ref_opt_optimize_sim3(const float *T1, const float *T2, const float *K1, const float *K2, int N, const uint8_t *valid, const float *Xw1, const float *Xw2,
                          const float *uv1, const float *uv2, const float *inv_sigma2_1, const float *inv_sigma2_2, double *sim3, float th2, int fix_scale, uint8_t *inlier)
{
    KeyFrame kf1, kf2;
    kf1.mnId = 1; kf2.mnId = 2;
    set_pose(kf1.Tcw, T1); set_pose(kf2.Tcw, T2);
    kf1.mK = cv::Mat::eye(3, 3, CV_32F); kf2.mK = cv::Mat::eye(3, 3, CV_32F);
    kf1.mK.at<float>(0, 0) = K1[0]; kf1.mK.at<float>(1, 1) = K1[1]; kf1.mK.at<float>(0, 2) = K1[2]; kf1.mK.at<float>(1, 2) = K1[3];
    kf2.mK.at<float>(0, 0) = K2[0]; kf2.mK.at<float>(1, 1) = K2[1]; kf2.mK.at<float>(0, 2) = K2[2]; kf2.mK.at<float>(1, 2) = K2[3];
    std::vector<MapPoint> p1(N), p2(N);
    std::vector<MapPoint *> matches(N, nullptr);
    kf1.mvKeysUn.resize(N); kf2.mvKeysUn.resize(N); kf1.mvInvLevelSigma2.resize(N); kf2.mvInvLevelSigma2.resize(N);
    kf1.mvpMapPoints.assign(N, nullptr); kf2.mvpMapPoints.assign(N, nullptr);
    for (int i = 0; i < N; i++) {
        for (int c = 0; c < 3; c++) { p1[i].mWorldPos.at<float>(c) = Xw1[3 * i + c]; p2[i].mWorldPos.at<float>(c) = Xw2[3 * i + c]; }
        p1[i].mnId = i; p2[i].mnId = N + i;
        kf1.mvKeysUn[i].pt.x = uv1[2 * i]; kf1.mvKeysUn[i].pt.y = uv1[2 * i + 1]; kf1.mvKeysUn[i].octave = i; kf1.mvInvLevelSigma2[i] = inv_sigma2_1[i];
        kf2.mvKeysUn[i].pt.x = uv2[2 * i]; kf2.mvKeysUn[i].pt.y = uv2[2 * i + 1]; kf2.mvKeysUn[i].octave = i; kf2.mvInvLevelSigma2[i] = inv_sigma2_2[i];
        kf1.mvpMapPoints[i] = &p1[i];
        if (valid[i]) { kf2.mvpMapPoints[i] = &p2[i]; p2[i].mObservations[&kf2] = (size_t)i; matches[i] = &p2[i]; }
    }
    g2o::Sim3 S(Eigen::Quaterniond(sim3[3], sim3[0], sim3[1], sim3[2]), Eigen::Vector3d(sim3[4], sim3[5], sim3[6]), sim3[7]);
    const int n = Optimizer::OptimizeSim3(&kf1, &kf2, matches, S, th2, fix_scale != 0);
    sim3[0] = S.rotation().x(); sim3[1] = S.rotation().y(); sim3[2] = S.rotation().z(); sim3[3] = S.rotation().w();
    for (int c = 0; c < 3; c++) sim3[4 + c] = S.translation()[c];
    sim3[7] = S.scale();
    for (int i = 0; i < N; i++) inlier[i] = matches[i] ? 1 : 0;
    return n;
}

// ---- g2o primitives --------------------------------------------------------------------------------------------
void ref_g2o_se3_exp(const double *u6, double *q_xyzw, double *t3)
{
    Eigen::Matrix<double, 6, 1> u; for (int i = 0; i < 6; i++) u[i] = u6[i];
    const g2o::SE3Quat T = g2o::SE3Quat::exp(u);
    q_xyzw[0] = T.rotation().x(); q_xyzw[1] = T.rotation().y(); q_xyzw[2] = T.rotation().z(); q_xyzw[3] = T.rotation().w();
    for (int c = 0; c < 3; c++) t3[c] = T.translation()[c];
}
// VertexSE3Expmap::oplusImpl: estimate <- exp(update) * estimate
void ref_g2o_se3_oplus(double *q_xyzw, double *t3, const double *u6)
{
    g2o::VertexSE3Expmap v;
    v.setEstimate(g2o::SE3Quat(Eigen::Quaterniond(q_xyzw[3], q_xyzw[0], q_xyzw[1], q_xyzw[2]), Eigen::Vector3d(t3[0], t3[1], t3[2])));
    v.oplus(u6);
    const g2o::SE3Quat &T = v.estimate();
    q_xyzw[0] = T.rotation().x(); q_xyzw[1] = T.rotation().y(); q_xyzw[2] = T.rotation().z(); q_xyzw[3] = T.rotation().w();
    for (int c = 0; c < 3; c++) t3[c] = T.translation()[c];
}
// Converter::toSE3Quat / toCvMat round trip pieces: float Tcw -> (q, t), and back
void ref_converter_to_se3quat(const float *Tcw, double *q_xyzw, double *t3)
{
    cv::Mat T(4, 4, CV_32F); set_pose(T, Tcw);
    const g2o::SE3Quat S = Converter::toSE3Quat(T);
    q_xyzw[0] = S.rotation().x(); q_xyzw[1] = S.rotation().y(); q_xyzw[2] = S.rotation().z(); q_xyzw[3] = S.rotation().w();
    for (int c = 0; c < 3; c++) t3[c] = S.translation()[c];
}
void ref_converter_to_cvmat(const double *q_xyzw, const double *t3, float *Tcw)
{
    const g2o::SE3Quat S(Eigen::Quaterniond(q_xyzw[3], q_xyzw[0], q_xyzw[1], q_xyzw[2]), Eigen::Vector3d(t3[0], t3[1], t3[2]));
    get_pose(Converter::toCvMat(S), Tcw);
}
// EdgeSE3ProjectXYZ: error, chi2, depth test, analytic Jacobians (point 2x3, pose 2x6; row-major out)
void ref_g2o_edge_se3_project_xyz(const double *q_xyzw, const double *t3, const double *Xw, const double *obs, double inv_sigma2, const double *K4,
                                  double *err2, double *chi2, int *depth_ok, double *Jpoint6, double *Jpose12)
{
    g2o::VertexSBAPointXYZ vp; g2o::VertexSE3Expmap vs;
    vp.setId(0); vs.setId(1);
    vp.setEstimate(Eigen::Vector3d(Xw[0], Xw[1], Xw[2]));
    vs.setEstimate(g2o::SE3Quat(Eigen::Quaterniond(q_xyzw[3], q_xyzw[0], q_xyzw[1], q_xyzw[2]), Eigen::Vector3d(t3[0], t3[1], t3[2])));
    g2o::EdgeSE3ProjectXYZ e;
    e.setVertex(0, &vp); e.setVertex(1, &vs);
    Eigen::Vector2d o(obs[0], obs[1]);
    e.setMeasurement(o);
    e.setInformation(Eigen::Matrix2d::Identity() * inv_sigma2);
    e.fx = K4[0]; e.fy = K4[1]; e.cx = K4[2]; e.cy = K4[3];
    e.computeError();
    err2[0] = e.error()[0]; err2[1] = e.error()[1];
    *chi2 = e.chi2(); *depth_ok = e.isDepthPositive() ? 1 : 0;
    g2o::JacobianWorkspace ws; ws.updateSize(&e); ws.allocate();                  // the Jacobian maps of an edge point into the workspace: keep it alive while they are read
    static_cast<g2o::OptimizableGraph::Edge &>(e).linearizeOplus(ws);
    for (int r = 0; r < 2; r++) { for (int c = 0; c < 3; c++) Jpoint6[3 * r + c] = e.jacobianOplusXi()(r, c); for (int c = 0; c < 6; c++) Jpose12[6 * r + c] = e.jacobianOplusXj()(r, c); }
}
void ref_g2o_edge_se3_only_pose(const double *q_xyzw, const double *t3, const double *Xw, const double *obs, const double *K4, double *err2, double *Jpose12)
{
    g2o::VertexSE3Expmap vs;
    vs.setId(0);
    vs.setEstimate(g2o::SE3Quat(Eigen::Quaterniond(q_xyzw[3], q_xyzw[0], q_xyzw[1], q_xyzw[2]), Eigen::Vector3d(t3[0], t3[1], t3[2])));
    g2o::EdgeSE3ProjectXYZOnlyPose e;
    e.setVertex(0, &vs);
    e.setMeasurement(Eigen::Vector2d(obs[0], obs[1]));
    e.setInformation(Eigen::Matrix2d::Identity());
    e.fx = K4[0]; e.fy = K4[1]; e.cx = K4[2]; e.cy = K4[3];
    e.Xw = Eigen::Vector3d(Xw[0], Xw[1], Xw[2]);
    e.computeError();
    err2[0] = e.error()[0]; err2[1] = e.error()[1];
    g2o::JacobianWorkspace ws; ws.updateSize(&e); ws.allocate();                  // the Jacobian maps of an edge point into the workspace: keep it alive while they are read
    static_cast<g2o::OptimizableGraph::Edge &>(e).linearizeOplus(ws);
    for (int r = 0; r < 2; r++) for (int c = 0; c < 6; c++) Jpose12[6 * r + c] = e.jacobianOplusXi()(r, c);
}
// RobustKernelHuber::robustify -> rho[3]
void ref_g2o_huber(double e2, double delta, double *rho3)
{
    g2o::RobustKernelHuber k; k.setDelta(delta);
    Eigen::Vector3d rho; k.robustify(e2, rho);
    for (int i = 0; i < 3; i++) rho3[i] = rho[i];
}
static g2o::Sim3 mk_sim3(const double *s8) { return g2o::Sim3(Eigen::Quaterniond(s8[3], s8[0], s8[1], s8[2]), Eigen::Vector3d(s8[4], s8[5], s8[6]), s8[7]); }
static void put_sim3(const g2o::Sim3 &S, double *s8)
{
    s8[0] = S.rotation().x(); s8[1] = S.rotation().y(); s8[2] = S.rotation().z(); s8[3] = S.rotation().w();
    for (int c = 0; c < 3; c++) s8[4 + c] = S.translation()[c];
    s8[7] = S.scale();
}
void ref_g2o_sim3_exp(const double *u7, double *s8) { Eigen::Matrix<double, 7, 1> u; for (int i = 0; i < 7; i++) u[i] = u7[i]; put_sim3(g2o::Sim3(u), s8); }
void ref_g2o_sim3_log(const double *s8, double *u7) { const Eigen::Matrix<double, 7, 1> u = mk_sim3(s8).log(); for (int i = 0; i < 7; i++) u7[i] = u[i]; }
void ref_g2o_sim3_inverse(const double *s8, double *o8) { put_sim3(mk_sim3(s8).inverse(), o8); }
void ref_g2o_sim3_mul(const double *a8, const double *b8, double *o8) { put_sim3(mk_sim3(a8) * mk_sim3(b8), o8); }
void ref_g2o_sim3_map(const double *s8, const double *x3, double *o3) { const Eigen::Vector3d r = mk_sim3(s8).map(Eigen::Vector3d(x3[0], x3[1], x3[2])); for (int c = 0; c < 3; c++) o3[c] = r[c]; }
// VertexSim3Expmap::oplusImpl (with / without fixed scale)
void ref_g2o_sim3_oplus(double *s8, const double *u7, int fix_scale)
{
    g2o::VertexSim3Expmap v; v._fix_scale = fix_scale != 0; v.setEstimate(mk_sim3(s8));
    double u[7]; for (int i = 0; i < 7; i++) u[i] = u7[i];
    v.oplus(u);
    put_sim3(v.estimate(), s8);
}
// EdgeSim3: error = log(meas * Si * Sj^-1) and the numeric Jacobians g2o computes for it (base_binary_edge.hpp:131-205), row-major 7x7 each
void ref_g2o_edge_sim3(const double *meas8, const double *si8, const double *sj8, int fix_scale, double *err7, double *Ji49, double *Jj49)
{
    g2o::VertexSim3Expmap vi, vj;
    vi.setId(0); vj.setId(1); vi._fix_scale = vj._fix_scale = fix_scale != 0;
    vi.setEstimate(mk_sim3(si8)); vj.setEstimate(mk_sim3(sj8));
    g2o::EdgeSim3 e;
    e.setVertex(0, &vi); e.setVertex(1, &vj);
    e.setMeasurement(mk_sim3(meas8));
    e.information() = Eigen::Matrix<double, 7, 7>::Identity();
    e.computeError();
    for (int i = 0; i < 7; i++) err7[i] = e.error()[i];
    g2o::JacobianWorkspace ws; ws.updateSize(&e); ws.allocate();                  // the Jacobian maps of an edge point into the workspace: keep it alive while they are read
    static_cast<g2o::OptimizableGraph::Edge &>(e).linearizeOplus(ws);
    for (int r = 0; r < 7; r++) for (int c = 0; c < 7; c++) { Ji49[7 * r + c] = e.jacobianOplusXi()(r, c); Jj49[7 * r + c] = e.jacobianOplusXj()(r, c); }
}
// EdgeSim3ProjectXYZ / EdgeInverseSim3ProjectXYZ (OptimizeSim3): error and numeric Jacobians wrt the point (2x3) and the Sim3 vertex (2x7)
void ref_g2o_edge_sim3_project(int inverse, const double *s8, const double *K1, const double *K2, const double *X3, const double *obs, int fix_scale,
                               double *err2, double *Jpoint6, double *Jsim14)
{
    g2o::VertexSBAPointXYZ vp; g2o::VertexSim3Expmap vs;
    vp.setId(1); vs.setId(0); vs._fix_scale = fix_scale != 0;
    vp.setEstimate(Eigen::Vector3d(X3[0], X3[1], X3[2]));
    vs.setEstimate(mk_sim3(s8));
    vs._principle_point1[0] = K1[2]; vs._principle_point1[1] = K1[3]; vs._focal_length1[0] = K1[0]; vs._focal_length1[1] = K1[1];
    vs._principle_point2[0] = K2[2]; vs._principle_point2[1] = K2[3]; vs._focal_length2[0] = K2[0]; vs._focal_length2[1] = K2[1];
    Eigen::Vector2d o(obs[0], obs[1]);
    if (!inverse) {
        g2o::EdgeSim3ProjectXYZ e;
        e.setVertex(0, &vp); e.setVertex(1, &vs); e.setMeasurement(o); e.setInformation(Eigen::Matrix2d::Identity());
        e.computeError(); err2[0] = e.error()[0]; err2[1] = e.error()[1];
        g2o::JacobianWorkspace ws; ws.updateSize(&e); ws.allocate();                  // the Jacobian maps of an edge point into the workspace: keep it alive while they are read
    static_cast<g2o::OptimizableGraph::Edge &>(e).linearizeOplus(ws);
        for (int r = 0; r < 2; r++) { for (int c = 0; c < 3; c++) Jpoint6[3 * r + c] = e.jacobianOplusXi()(r, c); for (int c = 0; c < 7; c++) Jsim14[7 * r + c] = e.jacobianOplusXj()(r, c); }
    } else {
        g2o::EdgeInverseSim3ProjectXYZ e;
        e.setVertex(0, &vp); e.setVertex(1, &vs); e.setMeasurement(o); e.setInformation(Eigen::Matrix2d::Identity());
        e.computeError(); err2[0] = e.error()[0]; err2[1] = e.error()[1];
        g2o::JacobianWorkspace ws; ws.updateSize(&e); ws.allocate();                  // the Jacobian maps of an edge point into the workspace: keep it alive while they are read
    static_cast<g2o::OptimizableGraph::Edge &>(e).linearizeOplus(ws);
        for (int r = 0; r < 2; r++) { for (int c = 0; c < 3; c++) Jpoint6[3 * r + c] = e.jacobianOplusXi()(r, c); for (int c = 0; c < 7; c++) Jsim14[7 * r + c] = e.jacobianOplusXj()(r, c); }
    }
}

// The numeric core of Optimizer::OptimizeEssentialGraph: the same g2o graph the reference sets up in Optimizer.cc:808-1008 (BlockSolver_7_3 over
// LinearSolverEigen, Levenberg with setUserLambdaInit(1e-16), VertexSim3Expmap with _fix_scale, EdgeSim3 with identity information, edge vertex 0 = i,
// vertex 1 = j, measurement Sji), built from flat arrays, initializeOptimization() + optimize(iterations).  sim3 [K][8] in/out.
int ref_g2o_pose_graph(int K, double *sim3, const uint8_t *fixed, int E, const int *e_i, const int *e_j, const double *e_meas, int fix_scale, int iterations)
{
    g2o::SparseOptimizer optimizer;
    optimizer.setVerbose(false);
    g2o::BlockSolver_7_3::LinearSolverType *linearSolver = new g2o::LinearSolverEigen<g2o::BlockSolver_7_3::PoseMatrixType>();
    g2o::BlockSolver_7_3 *solver_ptr = new g2o::BlockSolver_7_3(linearSolver);
    g2o::OptimizationAlgorithmLevenberg *solver = new g2o::OptimizationAlgorithmLevenberg(solver_ptr);
    solver->setUserLambdaInit(1e-16);
    optimizer.setAlgorithm(solver);
    std::vector<g2o::VertexSim3Expmap *> vs(K);
    for (int k = 0; k < K; k++) {
        g2o::VertexSim3Expmap *v = new g2o::VertexSim3Expmap();
        v->setEstimate(mk_sim3(sim3 + 8 * k));
        if (fixed[k]) v->setFixed(true);
        v->setId(k);
        v->setMarginalized(false);
        v->_fix_scale = fix_scale != 0;
        optimizer.addVertex(v);
        vs[k] = v;
    }
    const Eigen::Matrix<double, 7, 7> matLambda = Eigen::Matrix<double, 7, 7>::Identity();
    for (int e = 0; e < E; e++) {
        g2o::EdgeSim3 *ed = new g2o::EdgeSim3();
        ed->setVertex(1, dynamic_cast<g2o::OptimizableGraph::Vertex *>(optimizer.vertex(e_j[e])));
        ed->setVertex(0, dynamic_cast<g2o::OptimizableGraph::Vertex *>(optimizer.vertex(e_i[e])));
        ed->setMeasurement(mk_sim3(e_meas + 8 * e));
        ed->information() = matLambda;
        optimizer.addEdge(ed);
    }
    optimizer.initializeOptimization();
    const int its = optimizer.optimize(iterations);
    for (int k = 0; k < K; k++) put_sim3(vs[k]->estimate(), sim3 + 8 * k);
    return its;
}

}  // extern "C"
