"""The KeyFrame / Sim3 members of ORBmatcher composed from the two C-ABI halves (projection, window search) exactly like the C++ drop-in
(orbslamm_b200/host/ORBmatcher_b200.cc) composes them; parametrised by a backend so that the same composition runs on the oracle
(CPU tests, against the reference's object code) and on the CUDA library (GPU tests).

Reference: S/src/ORBmatcher.cc:292-405 (SearchByProjection(KF, Scw)), :827-977 and :979-1102 (Fuse), :1104-1328 (SearchBySim3),
:1474-1601 (SearchByProjection(Frame, KF)).  The host-side cv::Mat algebra that precedes the loops (Scw decomposition, sR12 / sR21 / t21)
is restated here in numpy float32 with OpenCV's evaluation order (Mat / scalar = multiply by the double reciprocal; small gemm)."""
import numpy as np

import orbslamm_b200 as _ob
from orbslamm_b200 import synth


class _LazyOracle:
    """the CPU oracle is loaded on first use: the CUDA-only users of this module (bench.py's map-merge leg) never touch it"""
    def __getattr__(self, name):
        import oracle as o
        return getattr(o, name)


oracle = _LazyOracle()

F32 = np.float32
TH_LOW, TH_HIGH = 50, 100


# ------------------------------------------------------------------ cv::Mat algebra on the host
def gemm32(A, B):
    """OpenCV small-matrix product: fp32 products added left to right."""
    A = np.asarray(A, F32); B = np.asarray(B, F32)
    out = np.zeros((A.shape[0], B.shape[1]), F32)
    for i in range(A.shape[0]):
        for j in range(B.shape[1]):
            s = F32(A[i, 0] * B[0, j])
            for k in range(1, A.shape[1]):
                s = F32(s + F32(A[i, k] * B[k, j]))
            out[i, j] = s
    return out


def scale32(A, s):
    """Mat * double -> float"""
    return (np.asarray(A, F32).astype(np.float64) * np.float64(s)).astype(F32)


def decompose_scw(Scw):
    """ORBmatcher.cc:300-305"""
    Scw = np.asarray(Scw, F32).reshape(4, 4)
    sR = Scw[:3, :3]
    r0 = sR[0].astype(np.float64)
    scw = F32(np.sqrt(r0[0] * r0[0] + r0[1] * r0[1] + r0[2] * r0[2]))
    Rcw = scale32(sR, 1.0 / np.float64(scw))
    tcw = scale32(Scw[:3, 3:4], 1.0 / np.float64(scw))
    Ow = gemm32(-(Rcw.T.copy()), tcw)
    return Rcw, tcw.ravel(), Ow.ravel()


def camera_centre(Tcw):
    Tcw = np.asarray(Tcw, F32).reshape(4, 4)
    return gemm32(-(Tcw[:3, :3].T.copy()), Tcw[:3, 3:4]).ravel()


# ------------------------------------------------------------------ backends
class OracleBackend:

    def project(self, R, t, K4, bounds4, log_sf, th, flags, sf, pts, valid, Ow=None, R2=None, t2=None, use_normal=True):
        V = oracle.make_projection(R, t, K4, bounds4, log_sf, th, flags, Ow=Ow, R2=R2, t2=t2)
        return oracle.project_points(V, sf, pts["Xw"], pts["normal"] if use_normal else None, pts["mf_min"], pts["mf_max"], valid)

    def best(self, kf, qv, uv, rad, l0, l1, q_desc, th_dist, gate=False):
        g = oracle.grid_params(*[float(x) for x in kf["grid_bounds4"]])
        return oracle.search_best_in_window(g, kf["win_origin2"], kf["xy"], kf["octave"], kf["desc"], qv, uv, rad, l0, l1, q_desc, th_dist,
                                            kf["inv_level_sigma2"] if gate else None, 5.99)

    def claim(self, kf, qv, uv, rad, l0, l1, q_angle, q_desc, th_dist, ori, fm_in, frame=False):
        g = oracle.grid_params(*[float(x) for x in kf["grid_bounds4"]])
        return oracle.search_by_projection_kf(g, None if frame else kf["win_origin2"], kf["xy"], kf["octave"], kf["angle"], kf["desc"], qv, uv, rad,
                                              l0, l1, q_angle, q_desc, th_dist, 0.0, ori, fm_in)


class CudaBackend:
    def __init__(self):
        import orbslamm_b200 as ob
        self.ob = ob
        self.m = ob.ORBmatcher(0.75, True)

    def project(self, R, t, K4, bounds4, log_sf, th, flags, sf, pts, valid, Ow=None, R2=None, t2=None, use_normal=True):
        V = self.ob.make_projection(R, t, K4, bounds4, log_sf, th, flags, Ow=Ow, R2=R2, t2=t2)
        M = len(valid)
        r = self.m.project_points([V], sf, pts["Xw"][None], pts["normal"][None] if use_normal else None, pts["mf_min"][None], pts["mf_max"][None],
                                  np.array([M], np.int32), np.asarray(valid, np.uint8)[None])
        return tuple(a[0] for a in r)

    def best(self, kf, qv, uv, rad, l0, l1, q_desc, th_dist, gate=False):
        N, M = len(kf["xy"]), len(qv)
        bi, bd = self.m.search_best_in_window(kf["grid_bounds4"], kf["win_origin2"], kf["xy"][None], kf["octave"][None], kf["desc"][None],
                                              np.array([N], np.int32), qv[None], uv[None], rad[None], l0[None], l1[None], q_desc[None],
                                              np.array([M], np.int32), th_dist, kf["inv_level_sigma2"] if gate else None, 5.99)
        return bi[0], bd[0]

    def claim(self, kf, qv, uv, rad, l0, l1, q_angle, q_desc, th_dist, ori, fm_in, frame=False):
        N, M = len(kf["xy"]), len(qv)
        nm, fm = self.m.SearchByProjectionKF(kf["grid_bounds4"], None if frame else kf["win_origin2"], kf["xy"][None], kf["octave"][None],
                                             kf["angle"][None], kf["desc"][None], np.array([N], np.int32), qv[None], uv[None], rad[None], l0[None],
                                             l1[None], np.asarray(q_angle, F32)[None], q_desc[None], np.array([M], np.int32), th_dist, 0.0, ori,
                                             None if fm_in is None else fm_in[None])
        return int(nm[0]), fm[0]


# ------------------------------------------------------------------ the five members
P = _ob               # PROJ_* flags (include/orbslamm_b200.h; the oracle's values are the same and tests/test_cabi_symbols.py keeps them so)


def kf_bounds(kf):
    """KeyFrame::mnMinX.. are ints (KeyFrame.h); IsInImage and GetFeaturesInArea use them"""
    return np.trunc(np.asarray(kf["grid_bounds4"], F32)).astype(F32)


def search_kf_sim3(B, kf, Scw, th, pts, skip, held):
    """SearchByProjection(pKF, Scw, vpPoints, vpMatched, th): returns (nmatches, feat_match[N]: point index, -2 held, -1 none)."""
    Rcw, tcw, Ow = decompose_scw(Scw)
    qv, uv, rad, l0, l1, _ = B.project(Rcw, tcw, kf["K4"], kf_bounds(kf), kf["log_sf"], F32(int(th)), P.PROJ_CHECK_NORMAL, kf["scale_factors"], pts,
                                       1 - np.asarray(skip, np.uint8), Ow=Ow)
    fm_in = np.where(np.asarray(held) > 0, 1 << 30, -1).astype(np.int32)
    n, fm = B.claim(kf, qv, uv, rad, l0, l1, np.zeros(len(qv), F32), pts["desc"], TH_LOW, False, fm_in)
    return n, np.where(fm == 1 << 30, -2, fm)


def fuse_search(B, kf, th, pts, skip, Scw=None):
    """Search half of Fuse: slot[i] = the keyframe feature the point would be fused into, or -1."""
    if Scw is None:
        T = np.asarray(kf["Tcw"], F32).reshape(4, 4)
        Rcw, tcw, Ow = T[:3, :3], T[:3, 3], camera_centre(T)
    else:
        Rcw, tcw, Ow = decompose_scw(Scw)
    qv, uv, rad, l0, l1, _ = B.project(Rcw, tcw, kf["K4"], kf_bounds(kf), kf["log_sf"], F32(th), P.PROJ_CHECK_NORMAL, kf["scale_factors"], pts,
                                       1 - np.asarray(skip, np.uint8), Ow=Ow)
    bi, _ = B.best(kf, qv, uv, rad, l0, l1, pts["desc"], TH_LOW, gate=Scw is None)
    return bi


def search_by_sim3(B, kf1, kf2, s12, R12, t12, th, has1, pts1, has2, pts2, matches12):
    """SearchBySim3: returns (nFound, matches12 updated)."""
    R12 = np.asarray(R12, F32).reshape(3, 3); t12 = np.asarray(t12, F32).reshape(3, 1)
    sR12 = scale32(R12, np.float64(F32(s12)))
    sR21 = scale32(R12.T.copy(), 1.0 / np.float64(F32(s12)))
    t21 = gemm32(-sR21, t12)
    T1 = np.asarray(kf1["Tcw"], F32).reshape(4, 4); T2 = np.asarray(kf2["Tcw"], F32).reshape(4, 4)
    m12 = np.asarray(matches12, np.int32).copy()
    N1, N2 = len(kf1["xy"]), len(kf2["xy"])
    done1 = m12 >= 0
    done2 = np.zeros(N2, bool); done2[m12[done1]] = True        # GetIndexInKeyFrame(pKF2) of an existing match = its feature index
    flags = P.PROJ_TWO_STEP | P.PROJ_DIST_CAMERA
    v1 = ((np.asarray(has1) > 0) & ~done1).astype(np.uint8)
    qv, uv, rad, l0, l1, _ = B.project(T1[:3, :3], T1[:3, 3], kf1["K4"], kf_bounds(kf2), kf2["log_sf"], F32(th), flags, kf2["scale_factors"], pts1, v1,
                                       R2=sR21, t2=t21.ravel(), use_normal=False)
    match1, _ = B.best(kf2, qv, uv, rad, l0, l1, pts1["desc"], TH_HIGH)
    v2 = ((np.asarray(has2) > 0) & ~done2).astype(np.uint8)
    qv, uv, rad, l0, l1, _ = B.project(T2[:3, :3], T2[:3, 3], kf1["K4"], kf_bounds(kf1), kf1["log_sf"], F32(th), flags, kf1["scale_factors"], pts2, v2,
                                       R2=sR12, t2=t12.ravel(), use_normal=False)
    match2, _ = B.best(kf1, qv, uv, rad, l0, l1, pts2["desc"], TH_HIGH)
    n = 0
    for i1 in range(N1):
        j = match1[i1]
        if j >= 0 and match2[j] == i1:
            m12[i1] = j; n += 1
    return n, m12


def search_frame_kf(B, cur, Tcw, K4, bounds4, log_sf, sf, held, has, skip, pts, kf_angle, th, orb_dist, check_ori):
    """SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist): returns (nmatches, feat_match[N])."""
    T = np.asarray(Tcw, F32).reshape(4, 4)
    Ow = camera_centre(T)
    flags = P.PROJ_NO_DEPTH | P.PROJ_FRAME_BOUNDS | P.PROJ_FRAME_UV | P.PROJ_LEVEL_PLUS1
    v = ((np.asarray(has) > 0) & ~(np.asarray(skip) > 0)).astype(np.uint8)
    qv, uv, rad, l0, l1, _ = B.project(T[:3, :3], T[:3, 3], K4, bounds4, log_sf, F32(th), flags, sf, pts, v, Ow=Ow, use_normal=False)
    fm_in = np.where(np.asarray(held) > 0, 1 << 30, -1).astype(np.int32)
    n, fm = B.claim(cur, qv, uv, rad, l0, l1, kf_angle, pts["desc"], orb_dist, check_ori, fm_in, frame=True)
    return n, np.where(fm == 1 << 30, -2, fm)


# ------------------------------------------------------------------ synthetic cases
def make_case(cam, stream_id, distorted_bounds=False, seed=0):
    """Two synthetic frames: the current frame's features act as the target keyframe (pose Tcw), the last frame's keypoints as map points
    with the MapPoint members the family reads (normal, mfMinDistance, mfMaxDistance)."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import make_tracking_case
    k = make_tracking_case(cam, stream_id)
    rng = np.random.default_rng(100 + seed + stream_id)
    cur, last = k["cur"], k["last"]
    sf = np.array(list(k["P"].scale)[:8], F32); inv2 = np.array(list(k["P"].inv_sigma2)[:8], F32)
    gb = np.array([-27.3, -20.7, cam["w"] + 25.6, cam["h"] + 18.2], F32) if distorted_bounds else k["bounds"].astype(F32)
    kf = dict(xy=np.stack([cur["x"], cur["y"]], 1).astype(F32), octave=cur["octave"].astype(np.int32), angle=cur["angle"].astype(F32),
              desc=np.ascontiguousarray(cur["desc"]), scale_factors=sf, inv_level_sigma2=inv2, K4=k["K4"], grid_bounds4=gb, Tcw=k["Tcw"],
              log_sf=F32(np.log(sf[1])))
    kf["win_origin2"] = kf_bounds(kf)[:2].copy()
    M = len(last["x"])
    Xw = k["Xw"]
    # reference keyframe of every point = the identity-pose camera; mfMaxDistance = dist * scale[level], mfMinDistance = max / scale[nlevels-1]
    # (MapPoint::UpdateNormalAndDepth, MapPoint.cc:331-371); normal = mean viewing direction, perturbed so that a few fail the 60 degree test
    d = np.linalg.norm(Xw, axis=1)
    lvl = last["octave"]
    mf_max = (d * sf[lvl] * rng.uniform(0.8, 1.25, M)).astype(F32)
    # MapPoint::PredictScale is not clamped in this version of the reference: a level outside [0, 8) indexes mvScaleFactors out of range
    # (undefined behaviour, whatever the heap holds).  Keep the predicted level of every point inside the pyramid.
    Tk = k["Tcw"].astype(np.float64)
    dt = np.linalg.norm(Xw - (-Tk[:3, :3].T @ Tk[:3, 3]), axis=1)
    lv = np.ceil(np.log(mf_max / dt) / np.log(1.2))
    bad = (lv < 0.001) | (lv > 6.999)
    mf_max = np.where(bad, dt * 1.2 ** (np.clip(lvl, 1, 7) - 0.5), mf_max).astype(F32)
    mf_min = (mf_max / sf[7]).astype(F32)
    nrm = Xw / d[:, None] + rng.normal(0, 0.45, (M, 3))
    nrm = (nrm / np.linalg.norm(nrm, axis=1)[:, None]).astype(F32)
    pts = dict(Xw=Xw.astype(F32), normal=nrm, mf_min=mf_min, mf_max=mf_max, desc=np.ascontiguousarray(last["desc"]))
    skip = (rng.random(M) < 0.1).astype(np.uint8)
    held = (rng.random(len(cur["x"])) < 0.15).astype(np.uint8)
    return dict(k=k, kf=kf, pts=pts, skip=skip, held=held, sf=sf, last=last, cur=cur)


def sim3_of(Tcw, s):
    S = np.asarray(Tcw, F32).reshape(4, 4).copy()
    S[:3, :3] = (S[:3, :3].astype(np.float64) * s).astype(F32)
    S[:3, 3] = (S[:3, 3].astype(np.float64) * s).astype(F32)
    return S


def make_sim3_pair(cam, stream_id, s12=1.02):
    """KF1 = last frame, KF2 = current frame of a synthetic stream whose frames differ by an integer image shift: a fronto-parallel plane at
    depth z0 seen from two cameras that differ by a translation parallel to it reproduces the shift exactly.  KF2 lives in its own map,
    1 / s12 the size of KF1's.  Returns everything SearchBySim3 reads."""
    c = make_case(cam, stream_id)
    k, last, cur = c["k"], c["last"], c["cur"]
    rng = np.random.default_rng(7 + stream_id)
    sf, inv2 = c["sf"], c["kf"]["inv_level_sigma2"]
    fx, fy, cx, cy = [float(v) for v in k["K4"]]
    dx, dy = k["shift"]
    z0 = 20.0
    a = np.deg2rad(11.0)
    R1 = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]); t1 = np.array([0.4, -0.2, 1.1])
    T1 = np.eye(4); T1[:3, :3] = R1; T1[:3, 3] = t1
    Tsh = np.eye(4); Tsh[:3, 3] = [dx * z0 / fx, dy * z0 / fy, 0.0]
    T2 = Tsh @ T1                                                   # camera 2 from world (KF1's scale)

    def backproject(f, T, noise):
        n = len(f["x"])
        z = z0 + rng.normal(0, noise, n)
        Xc = np.stack([(f["x"] - cx) * z / fx, (f["y"] - cy) * z / fy, z], 1)
        return (Xc - T[:3, 3]) @ T[:3, :3]

    def points(f, X, T, desc, metric):
        # T: the camera the points get projected INTO (the other keyframe); metric: its map scale relative to KF1's.  The predicted level
        # stays inside the pyramid (see make_case): level = octave or octave - 1
        n = len(X)
        d = np.linalg.norm(X @ T[:3, :3].T + T[:3, 3], axis=1) * metric
        o = f["octave"].astype(np.float64)
        u = np.where(o > 0, rng.uniform(-1.4, -0.1, n), rng.uniform(-0.9, -0.1, n))
        mf_max = (d * 1.2 ** (o + u)).astype(F32)
        return dict(Xw=X.astype(F32), normal=np.zeros((n, 3), F32), mf_min=(mf_max / sf[7]).astype(F32), mf_max=mf_max, desc=np.ascontiguousarray(desc))

    def keyframe(f, T):
        kf = dict(xy=np.stack([f["x"], f["y"]], 1).astype(F32), octave=f["octave"].astype(np.int32), angle=f["angle"].astype(F32),
                  desc=np.ascontiguousarray(f["desc"]), scale_factors=sf, inv_level_sigma2=inv2, K4=k["K4"], grid_bounds4=c["kf"]["grid_bounds4"],
                  Tcw=T.astype(F32), log_sf=c["kf"]["log_sf"])
        kf["win_origin2"] = kf_bounds(kf)[:2].copy()
        return kf

    X1 = backproject(last, T1, 0.05); X2 = backproject(cur, T2, 0.05)
    pts1 = points(last, X1, T2, last["desc"], 1.0 / float(s12))    # PredictScale sees |p| in the target camera, in the target's metric
    pts2 = points(cur, X2, T1, cur["desc"], 1.0)
    pts2["Xw"] = (X2 / float(s12)).astype(F32)
    T2s = T2.copy(); T2s[:3, 3] = T2[:3, 3] / float(s12)            # pose of KF2 in its own map: p2 = (R2 X + t2) / s12
    kf1, kf2 = keyframe(last, T1), keyframe(cur, T2s)
    has1 = (rng.random(len(last["x"])) < 0.9).astype(np.uint8); has2 = (rng.random(len(cur["x"])) < 0.9).astype(np.uint8)
    # S12: p1 = s12 R12 p2 + t12 with T12 = T1 T2^-1 = Tsh^-1 (in KF1's metric)
    T12 = T1 @ np.linalg.inv(T2)
    R12 = T12[:3, :3].astype(F32); t12 = T12[:3, 3].astype(F32)
    m12 = np.full(len(last["x"]), -1, np.int32)
    # a few matches that already exist on entry (vpMatches12): nearest shifted keypoint pairs
    used = set()
    for i in range(0, len(last["x"]), 40):
        dd = np.abs(cur["x"] - last["x"][i] - dx) + np.abs(cur["y"] - last["y"][i] - dy)
        j = int(np.argmin(dd))
        if dd[j] < 1.5 and has1[i] and has2[j] and j not in used:
            m12[i] = j; used.add(j)
    return dict(kf1=kf1, kf2=kf2, pts1=pts1, pts2=pts2, has1=has1, has2=has2, s12=F32(s12), R12=R12, t12=t12, m12=m12)


def quat_of(R):
    """Eigen::Quaterniond(Matrix3d) for a proper rotation (w > 0 branch or not), x y z w"""
    from scipy.spatial.transform import Rotation
    q = Rotation.from_matrix(np.asarray(R, np.float64)).as_quat()
    return q if q[3] >= 0 else -q


def make_sim3_opt_case(cam, stream_id, n_outliers=12, s12=1.02, seed=0):
    """Correspondences for Optimizer::OptimizeSim3 from the SearchBySim3 pair: true feature pairs (nearest shifted keypoints), a few wrong
    pairs, camera-frame points in float as the reference computes them, and a perturbed initial Sim3."""
    p = make_sim3_pair(cam, stream_id, s12=s12)
    kf1, kf2 = p["kf1"], p["kf2"]
    rng = np.random.default_rng(11 + seed + stream_id)
    x1, x2 = kf1["xy"], kf2["xy"]
    T1, T2 = kf1["Tcw"], kf2["Tcw"]
    P1c_all = np.stack([gemm32(T1[:3, :3], X.reshape(3, 1)).ravel() + T1[:3, 3] for X in p["pts1"]["Xw"]]).astype(F32)
    P2c_all = np.stack([gemm32(T2[:3, :3], X.reshape(3, 1)).ravel() + T2[:3, 3] for X in p["pts2"]["Xw"]]).astype(F32)
    # true correspondences: project P1c into KF2 with the true S21 and take the nearest KF2 keypoint
    fx, fy, cx, cy = [float(v) for v in kf1["K4"]]
    R12 = p["R12"].astype(np.float64); t12 = p["t12"].astype(np.float64); s = float(p["s12"])
    q2 = (P1c_all.astype(np.float64) - t12) @ R12 / s
    u2 = np.stack([fx * q2[:, 0] / q2[:, 2] + cx, fy * q2[:, 1] / q2[:, 2] + cy], 1)
    N1 = len(x1)
    pair = np.full(N1, -1)
    for i in range(0, N1, 3):
        d = np.abs(x2 - u2[i]).sum(1)
        j = int(np.argmin(d))
        if d[j] < 1.0:
            pair[i] = j
    idx = np.where(pair >= 0)[0]
    bad = rng.choice(idx, size=min(n_outliers, len(idx) // 4), replace=False)
    pair[bad] = rng.integers(0, len(x2), len(bad))
    valid = (pair >= 0).astype(np.uint8)
    j = np.where(pair >= 0, pair, 0)
    inv2 = kf1["inv_level_sigma2"]
    case = dict(valid=valid, P1c=P1c_all, P2c=P2c_all[j], obs1=x1, obs2=x2[j], w1=inv2[kf1["octave"]], w2=inv2[kf2["octave"][j]], K1=kf1["K4"], K2=kf2["K4"],
                true=np.concatenate([quat_of(R12), t12, [s]]), bad=bad)
    from scipy.spatial.transform import Rotation
    dR = Rotation.from_rotvec(rng.normal(0, 0.004, 3)).as_matrix()
    case["init"] = np.concatenate([quat_of(dR @ R12), t12 + rng.normal(0, 0.02, 3), [s * 1.01]])
    return case


# ------------------------------------------------------------------ vocabulary-bucket matchers
def synthetic_nodes(desc, n_bits=6):
    """Stand-in for the vocabulary node of a descriptor (DBoW2 transform at levelsup): one bit per 5-byte group = 'more than half of its
    bits set'.  Descriptors that differ in a few bits mostly share the node, like neighbours in the vocabulary tree."""
    d = np.ascontiguousarray(desc, np.uint8)
    bits = np.unpackbits(d, axis=1)
    node = np.zeros(len(d), np.int64)
    for k in range(n_bits):
        node |= (bits[:, 40 * k:40 * k + 40].sum(1) > 20).astype(np.int64) << k
    return node * 3 + 7            # node ids are sparse in the real tree


def epipole_and_F(kf1, kf2):
    """Epipole of camera 1 in image 2 as SearchForTriangulation computes it (:665-673) and F12 = K1^-T [t12]x R12 K2^-1 (LocalMapping::ComputeF12)."""
    T1 = np.asarray(kf1["Tcw"], F32).reshape(4, 4); T2 = np.asarray(kf2["Tcw"], F32).reshape(4, 4)
    Cw = camera_centre(T1)
    C2 = (gemm32(T2[:3, :3], Cw.reshape(3, 1)).ravel() + T2[:3, 3]).astype(F32)
    with np.errstate(divide="ignore", invalid="ignore"):          # cameras that differ by an in-plane translation: the epipole is at infinity, as in the reference
        invz = F32(1.0) / C2[2]
        ex = F32(F32(F32(F32(kf2["K4"][0]) * C2[0]) * invz) + F32(kf2["K4"][2])); ey = F32(F32(F32(F32(kf2["K4"][1]) * C2[1]) * invz) + F32(kf2["K4"][3]))
    R1, t1, R2, t2 = T1[:3, :3].astype(np.float64), T1[:3, 3].astype(np.float64), T2[:3, :3].astype(np.float64), T2[:3, 3].astype(np.float64)
    R12 = R1 @ R2.T; t12 = -R12 @ t2 + t1
    tx = np.array([[0, -t12[2], t12[1]], [t12[2], 0, -t12[0]], [-t12[1], t12[0], 0]])
    K = lambda k: np.array([[k[0], 0, k[2]], [0, k[1], k[3]], [0, 0, 1]], np.float64)
    F12 = np.linalg.inv(K(kf1["K4"])).T @ tx @ R12 @ np.linalg.inv(K(kf2["K4"]))
    return ex, ey, F12.astype(F32)


def make_bow_case(cam, stream_id, perturb=False):
    """Two keyframes with synthetic vocabulary nodes, map-point masks and (for SearchForTriangulation) F12 / epipole.  perturb: the second pose
    used for F12 is moved off the true one, which gives a finite epipole and makes the epipolar gates reject part of the candidates."""
    p = make_sim3_pair(cam, stream_id, s12=1.0)
    kf1, kf2 = dict(p["kf1"]), dict(p["kf2"])
    rng = np.random.default_rng(23 + stream_id)
    if perturb:
        T = kf2["Tcw"].copy(); T[2, 3] += F32(0.35); T[0, 3] += F32(0.05)
        kf2["Tcw"] = T
    fv1 = oracle.feature_vector(synthetic_nodes(kf1["desc"])); fv2 = oracle.feature_vector(synthetic_nodes(kf2["desc"]))
    ex, ey, F12 = epipole_and_F(kf1, kf2)
    ls2 = (kf1["scale_factors"].astype(np.float64) ** 2).astype(F32)
    tri1 = (rng.random(len(kf1["xy"])) < 0.55).astype(np.uint8); tri2 = (rng.random(len(kf2["xy"])) < 0.55).astype(np.uint8)   # features WITH a map point
    return dict(kf1=kf1, kf2=kf2, fv1=fv1, fv2=fv2, has1=p["has1"], has2=p["has2"], tri1=tri1, tri2=tri2, ex=ex, ey=ey, F12=F12, ls2=ls2)


# ------------------------------------------------------------------ golden vectors (tools/make_golden.py, tests/test_golden_*.py)
GOLDEN_CAM = dict(synth.TUM); GOLDEN_CAM.update(w=400, h=300, nfeatures=400, cx=200.0, cy=150.0)


def _bow(B, mode, d1, a1, e1, fv1, d2, a2, e2, fv2, ratio, ori, epi=None):
    if isinstance(B, OracleBackend):
        return oracle.search_by_bow(mode, d1, a1, e1, fv1, d2, a2, e2, fv2, ratio, ori, epi)
    m = B.ob.ORBmatcher(ratio, ori)
    tri = None
    if epi is not None:
        tri = dict(xy1=epi["xy1"][None], xy2=epi["xy2"][None], octave2=epi["octave2"][None], F12=np.asarray(epi["F12"], F32).reshape(1, 9),
                   epipole=np.array([[epi["ex"], epi["ey"]]], F32), scale_factors2=epi["scale_factors2"], level_sigma2_2=epi["level_sigma2_2"])
    nm, mt = m.SearchByBoW(d1[None], a1[None], np.asarray(e1, np.uint8)[None], [len(d1)], [fv1], d2[None], a2[None], np.asarray(e2, np.uint8)[None], [len(d2)], [fv2], tri)
    return int(nm[0]), mt[0]


def golden_outputs(B):
    """Every member of the widened path on one 400x300 / 400-feature case, for backend B (oracle or CUDA): a flat dict of arrays."""
    from orbslamm_b200 import vocabulary as V
    out = {}
    c = make_case(GOLDEN_CAM, 21, distorted_bounds=True)
    kf, pts, skip, held, k = c["kf"], c["pts"], c["skip"], c["held"], c["k"]
    Scw = sim3_of(kf["Tcw"], 1.21)
    n, fm = search_kf_sim3(B, kf, Scw, 10, pts, skip, held)
    out["search_kf_sim3"] = np.concatenate([fm, [n]])
    out["fuse_kf"] = fuse_search(B, kf, 3.0, pts, skip)
    out["fuse_sim3"] = fuse_search(B, kf, 4.0, pts, skip, Scw=Scw)
    has = (np.random.default_rng(3).random(len(skip)) < 0.85).astype(np.uint8)
    cur = dict(kf); cur["grid_bounds4"] = k["bounds"].astype(F32)
    n, fm = search_frame_kf(B, cur, k["Tcw"], k["K4"], k["bounds"], kf["log_sf"], c["sf"], held, has, skip, pts, c["last"]["angle"], 10.0, 100, True)
    out["search_frame_kf"] = np.concatenate([fm, [n]])
    p = make_sim3_pair(GOLDEN_CAM, 21)
    n, m12 = search_by_sim3(B, p["kf1"], p["kf2"], p["s12"], p["R12"], p["t12"], 7.5, p["has1"], p["pts1"], p["has2"], p["pts2"], p["m12"])
    out["search_by_sim3"] = np.concatenate([m12, [n]])
    b = make_bow_case(GOLDEN_CAM, 21, True)
    k1, k2 = b["kf1"], b["kf2"]
    n, m = _bow(B, 0, k1["desc"], k1["angle"], b["has1"], b["fv1"], k2["desc"], k2["angle"], b["has2"], b["fv2"], 0.75, True)
    out["bow_kf_kf"] = np.concatenate([m, [n]])
    epi = dict(xy1=k1["xy"], xy2=k2["xy"], octave2=k2["octave"], F12=b["F12"], ex=b["ex"], ey=b["ey"], scale_factors2=k2["scale_factors"], level_sigma2_2=b["ls2"])
    n, m = _bow(B, 1, k1["desc"], k1["angle"], 1 - b["tri1"], b["fv1"], k2["desc"], k2["angle"], 1 - b["tri2"], b["fv2"], 0.6, False, epi)
    out["triangulation"] = np.concatenate([m, [n]])
    l, cu = k["last"], k["cur"]
    pm = np.stack([l["x"], l["y"]], 1).astype(F32)
    if isinstance(B, OracleBackend):
        n, m, pm2 = oracle.search_for_initialization(oracle.grid_params(*k["bounds"]), l, cu, pm, 100, 0.9, True)
    else:
        mi = B.ob.ORBmatcher(0.9, True)
        nm, mt, pmo = mi.SearchForInitialization(k["bounds"], l["octave"][None], l["angle"][None], l["desc"][None], [len(l["x"])],
                                                 np.stack([cu["x"], cu["y"]], 1).astype(F32)[None], cu["octave"][None], cu["angle"][None], cu["desc"][None], [len(cu["x"])],
                                                 pm[None], 100)
        n, m, pm2 = int(nm[0]), mt[0], pmo[0]
    out["init"] = np.concatenate([m, [n]]); out["init_prev"] = pm2
    voc = V.synthetic(6, 3, seed=9)
    if isinstance(B, OracleBackend):
        t = oracle.vocab_transform(voc, cu["desc"], 2)
    else:
        t = V.ORBVocabulary(voc).transform(cu["desc"][None], [len(cu["desc"])], 2)[0]
    out["voc_bow_ids"], out["voc_bow_vals"] = t["bow_ids"], t["bow_vals"]
    out["voc_fv_nodes"], out["voc_fv_start"], out["voc_fv_items"] = t["fv"]["nodes"], t["fv"]["start"], t["fv"]["items"]
    s = make_sim3_opt_case(GOLDEN_CAM, 21, n_outliers=6)
    if isinstance(B, OracleBackend):
        r = oracle.optimize_sim3(s["init"], s["valid"], s["P1c"], s["P2c"], s["obs1"], s["obs2"], s["w1"], s["w2"], s["K1"], s["K2"], 10.0, False)
        out["sim3"], out["sim3_inlier"] = r["sim3"], np.concatenate([r["inlier"], [r["n_in"]]])
    else:
        S, inl, nin, _ = B.ob.Optimizer().OptimizeSim3(s["init"][None], s["valid"][None], s["P1c"][None], s["P2c"][None], s["obs1"][None], s["obs2"][None], s["w1"][None],
                                                       s["w2"][None], s["K1"][None], s["K2"][None], [len(s["valid"])], 10.0, False)
        out["sim3"], out["sim3_inlier"] = S[0], np.concatenate([inl[0], [nin[0]]])
    return out


# ------------------------------------------------------------------ Sim3Solver RANSAC hypotheses
def make_sim3_ransac_case(cam, stream_id, n_hyp=300, seed=0):
    """Correspondences of make_sim3_opt_case plus n_hyp hypotheses (T12, T21) spread around the true Sim3 from exact to far off, as a RANSAC produces
    them; built in double and rounded to float like ComputeSim3's outputs."""
    from scipy.spatial.transform import Rotation
    s = make_sim3_opt_case(cam, stream_id)
    v = s["valid"] > 0
    p = make_sim3_pair(cam, stream_id)
    rng = np.random.default_rng(31 + seed + stream_id)
    R12 = p["R12"].astype(np.float64); t12 = p["t12"].astype(np.float64); s12 = float(p["s12"])
    T12 = np.zeros((n_hyp, 4, 4), F32); T21 = np.zeros((n_hyp, 4, 4), F32)
    for h in range(n_hyp):
        mag = [0.0, 1e-4, 1e-3, 5e-3, 2e-2, 0.1][h % 6]
        R = Rotation.from_rotvec(rng.normal(0, mag, 3)).as_matrix() @ R12
        t = t12 + rng.normal(0, 4 * mag, 3); sc = s12 * (1 + rng.normal(0, mag))
        A = np.eye(4); A[:3, :3] = sc * R; A[:3, 3] = t
        T12[h] = A.astype(F32); T21[h] = np.linalg.inv(A).astype(F32)
    oct1 = np.rint(-np.log(s["w1"][v].astype(np.float64)) / (2 * np.log(1.2))).astype(np.int32)
    oct2 = np.rint(-np.log(s["w2"][v].astype(np.float64)) / (2 * np.log(1.2))).astype(np.int32)
    ls2 = (p["kf1"]["scale_factors"].astype(np.float64) ** 2).astype(F32)
    return dict(X1=s["P1c"][v], X2=s["P2c"][v], oct1=oct1, oct2=oct2, ls2=ls2, K1=s["K1"], K2=s["K2"], T12=T12, T21=T21)


# ------------------------------------------------------------------ essential graph (Sim3 pose graph)
def sim3_compose(a, b):
    """g2o::Sim3 product on (qx qy qz qw tx ty tz s) rows, double"""
    from scipy.spatial.transform import Rotation
    Ra, Rb = Rotation.from_quat(a[:4]), Rotation.from_quat(b[:4])
    q = (Ra * Rb).as_quat()
    return np.concatenate([q if q[3] >= 0 else -q, a[7] * Ra.apply(b[4:7]) + a[4:7], [a[7] * b[7]]])


def sim3_inverse(a):
    from scipy.spatial.transform import Rotation
    Ri = Rotation.from_quat(a[:4]).inv()
    q = Ri.as_quat()
    return np.concatenate([q if q[3] >= 0 else -q, Ri.apply(-a[4:7] / a[7]), [1.0 / a[7]]])


def make_pose_graph(K=60, seed=0, n_loops=8, scale_drift=0.01):
    """A loop-closure situation like the one OptimizeEssentialGraph sees: K keyframes on a closed trajectory, their odometry edges (spanning tree +
    covisibility to the next but one) measured from drifted poses, a few loop edges measured from the true poses; vertices start at the drifted
    poses; keyframe 0 is fixed.  Returns (sim3 [K,8], fixed, e_i, e_j, e_meas [E,8], true [K,8])."""
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(seed)
    true = []
    for k in range(K):
        a = 2 * np.pi * k / K
        Rwc = Rotation.from_euler("y", -a + np.pi / 2) * Rotation.from_rotvec(rng.normal(0, 0.02, 3))
        C = np.array([10 * np.cos(a), 0.2 * np.sin(3 * a), 10 * np.sin(a)])
        Rcw = Rwc.inv()
        q = Rcw.as_quat()
        true.append(np.concatenate([q if q[3] >= 0 else -q, -Rcw.apply(C), [1.0]]))
    true = np.array(true)
    drift = [true[0].copy()]                       # Siw with accumulated drift (rotation, translation and scale)
    for k in range(1, K):
        rel = sim3_compose(true[k], sim3_inverse(true[k - 1]))                # S_k,k-1
        noise = np.concatenate([Rotation.from_rotvec(rng.normal(0, 0.004, 3)).as_quat(), rng.normal(0, 0.02, 3), [1 + rng.normal(0, scale_drift)]])
        drift.append(sim3_compose(sim3_compose(noise, rel), drift[k - 1]))
    drift = np.array(drift)
    ei, ej, meas = [], [], []
    def edge(i, j, src):                            # vertex[0] = i, vertex[1] = j, measurement Sji = Sjw * Swi (Optimizer.cc:895-905)
        ei.append(i); ej.append(j); meas.append(sim3_compose(src[j], sim3_inverse(src[i])))
    for k in range(1, K):
        edge(k, k - 1, drift)
        if k >= 2:
            edge(k, k - 2, drift)
    for _ in range(n_loops):                        # loop edges between the end and the start of the trajectory, from the true poses
        i = int(rng.integers(K - 8, K)); j = int(rng.integers(0, max(1, min(6, K - 8))))
        edge(i, j, true)
    fixed = np.zeros(K, np.uint8); fixed[0] = 1
    return drift, fixed, np.array(ei, np.int32), np.array(ej, np.int32), np.array(meas), true


# ------------------------------------------------------------------ essential graph through the C++ drop-in (tests/test_host_shim_gpu.py)
def make_essential_graph_scene(K=40, seed=4):
    """A mock map for Optimizer::OptimizeEssentialGraph: keyframe SE3 poses with drift, spanning tree k -> k-1, covisibility weights, one earlier loop
    edge, the loop connections and corrected / non-corrected Sim3 of the current keyframe and its two neighbours, map points with reference keyframes."""
    from scipy.spatial.transform import Rotation
    S, fixed, _, _, _, true = make_pose_graph(K, seed=seed, scale_drift=0.0)
    S = S.copy(); S[:, 7] = 1.0                                   # keyframe poses are SE3
    rng = np.random.default_rng(seed)
    def T_of(s):
        T = np.eye(4, dtype=F32); T[:3, :3] = Rotation.from_quat(s[:4]).as_matrix().astype(F32); T[:3, 3] = s[4:7].astype(F32); return T
    poses = np.stack([T_of(s) for s in S])
    parent = np.arange(-1, K - 1, dtype=np.int32)
    cov = []
    for k in range(1, K):
        cov += [k, k - 1, 200]
        if k >= 2: cov += [k, k - 2, 150]
        if k >= 3: cov += [k, k - 3, 90]
    cur, loop = K - 1, 0
    cov += [cur, 1, 120, cur, 0, 30, K - 2, 0, 50]
    loopedges = [K // 2, 5]
    loopconn = [cur, 0, cur, 1, K - 2, 0]
    corr_idx = [cur, K - 2, K - 3]
    corr = np.stack([np.concatenate([true[i][:7] * np.array([1, 1, 1, 1, 1.05, 1.05, 1.05]), [1.05]]) for i in corr_idx])     # [sR st] with s = 1.05
    noncorr = np.stack([S[i] for i in corr_idx])
    P = 3 * K
    pref = np.repeat(np.arange(K), 3).astype(np.int32)
    pts = rng.uniform(-12, 12, (P, 3)).astype(F32)
    pcorr = np.full(P, -1, np.int32); pcorr[-3:] = K - 2          # points already corrected by the current keyframe (mnCorrectedByKF / mnCorrectedReference)
    return dict(K=K, loop=loop, cur=cur, poses=poses, parent=parent, cov=np.array(cov, np.int32), loopedges=np.array(loopedges, np.int32),
                loopconn=np.array(loopconn, np.int32), corr_idx=np.array(corr_idx, np.int32), corr=corr, noncorr=noncorr, pref=pref, pts=pts, pcorr=pcorr)


def essential_graph_expected(sc, fix_scale, backend):
    """The reference's assembly (Optimizer.cc:820-1000) on the scene + backend(sim3, fixed, ei, ej, meas, fix_scale) -> corrected poses [K,4,4] f32, points f32."""
    from scipy.spatial.transform import Rotation
    K = sc["K"]
    def s_of_pose(T):
        q = Rotation.from_matrix(T[:3, :3].astype(np.float64)).as_quat()
        return np.concatenate([q if q[3] >= 0 else -q, T[:3, 3].astype(np.float64), [1.0]])
    corr = {int(i): sc["corr"][n] for n, i in enumerate(sc["corr_idx"])}; nonc = {int(i): sc["noncorr"][n] for n, i in enumerate(sc["corr_idx"])}
    vScw = np.stack([corr[k] if k in corr else s_of_pose(sc["poses"][k]) for k in range(K)])
    w = {}
    for a, b, ww in sc["cov"].reshape(-1, 3): w[(int(a), int(b))] = int(ww); w[(int(b), int(a))] = int(ww)
    ledge = {k: set() for k in range(K)}
    for a, b in sc["loopedges"].reshape(-1, 2): ledge[int(a)].add(int(b)); ledge[int(b)].add(int(a))
    conn = {}
    for a, b in sc["loopconn"].reshape(-1, 2): conn.setdefault(int(a), set()).add(int(b))
    ei, ej, meas = [], [], []
    inserted = set()
    for i in sorted(conn):                                         # std::map over KeyFrame*: the mock keyframes live in one array, so pointer order = index order
        Swi = sim3_inverse(vScw[i])
        for j in sorted(conn[i]):
            if (i != sc["cur"] or j != sc["loop"]) and w.get((i, j), 0) < 100: continue
            ei.append(i); ej.append(j); meas.append(sim3_compose(vScw[j], Swi)); inserted.add((min(i, j), max(i, j)))
    nc = lambda k: nonc[k] if k in nonc else vScw[k]
    for i in range(K):
        Swi = sim3_inverse(nc(i))
        p = int(sc["parent"][i])
        if p >= 0:
            ei.append(i); ej.append(p); meas.append(sim3_compose(nc(p), Swi))
        for l in sorted(ledge[i]):
            if l < i: ei.append(i); ej.append(l); meas.append(sim3_compose(nc(l), Swi))
        neigh = sorted([(ww, j) for (a, j), ww in w.items() if a == i and ww >= 100], key=lambda x: -x[0])
        children = {k for k in range(K) if int(sc["parent"][k]) == i}
        for ww, j in neigh:
            if j != p and j not in children and j not in ledge[i] and j < i and (min(i, j), max(i, j)) not in inserted:
                ei.append(i); ej.append(j); meas.append(sim3_compose(nc(j), Swi))
    fixed = np.zeros(K, np.uint8); fixed[sc["loop"]] = 1
    est = backend(vScw, fixed, np.array(ei, np.int32), np.array(ej, np.int32), np.array(meas), fix_scale)
    poses = np.zeros((K, 4, 4), F32)
    for k in range(K):
        T = np.eye(4); T[:3, :3] = Rotation.from_quat(est[k][:4]).as_matrix(); T[:3, 3] = est[k][4:7] / est[k][7]
        poses[k] = T.astype(F32)
    pts = sc["pts"].astype(np.float64).copy()
    for p in range(len(pts)):
        r = int(sc["pcorr"][p]) if sc["pcorr"][p] >= 0 else int(sc["pref"][p])
        Srw, Swr = vScw[r], sim3_inverse(est[r])
        x = Srw[7] * Rotation.from_quat(Srw[:4]).apply(pts[p]) + Srw[4:7]
        pts[p] = Swr[7] * Rotation.from_quat(Swr[:4]).apply(x) + Swr[4:7]
    return poses, pts.astype(F32), len(ei)
