"""CPU oracle for the ORBSLAMM hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package; the product (orbslamm_b200) never does.

Pieces:
  * orb_oracle.c   restatement of ORBextractor.cc on restated OpenCV primitives
  * slam_oracle.c  restatement of ORBmatcher / Optimizer(+g2o) arithmetic
  * orb_cv2.py     the same extractor but calling the *real* OpenCV primitives
                   through cv2 (what the reference itself calls) -- used to pin
                   orb_oracle.c and as the reference-faithful CPU baseline.

Parity status: the reference repository holds no golden vectors or tests for
this path.  Extractor: PINNED -- orb_oracle.c equals the reference's
own ORBextractor.cc compiled unmodified against an OpenCV stand-in (oracle/_ref,
ref_build.py, tests/test_oracle_vs_reference.py) and its OpenCV primitives equal
cv2 4.13.0 bit for bit.  Matcher: PINNED -- slam_oracle.c's search equals the reference's
own ORBmatcher.cc compiled unmodified against a stand-in data model (oracle/slamshim,
oracle/_ref/libref_orbmatcher.so).  Optimizer: Optimizer.cc + g2o cannot be compiled here
(Eigen headers absent); the oracle is pinned against an independent numpy/scipy twin,
beyond that "parity unpinned" by the reference.
"""
import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    """Compile oracle/*.c into oracle/_build/liboracle.so (gcc, no FMA contraction)."""
    out_dir = os.path.join(_HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("orb_oracle.c", "slam_oracle.c")]
    srcs = [s for s in srcs if os.path.exists(s)]
    deps = srcs + [os.path.join(_HERE, "..", "include", "orb_brief_pattern.inc")]
    if (not force) and os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(d) for d in deps):
        return so
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", so] + srcs + ["-lm"]
    subprocess.check_call(cmd)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _declare(_LIB)
    return _LIB


class OrbParams(ctypes.Structure):
    _fields_ = [("nfeatures", ctypes.c_int), ("nlevels", ctypes.c_int), ("ini_th", ctypes.c_int),
                ("min_th", ctypes.c_int), ("scale_factor", ctypes.c_double),
                ("scale", ctypes.c_float * 16), ("inv_scale", ctypes.c_float * 16),
                ("sigma2", ctypes.c_float * 16), ("inv_sigma2", ctypes.c_float * 16),
                ("features_per_level", ctypes.c_int * 16), ("umax", ctypes.c_int * 16)]


class Cand(ctypes.Structure):
    _fields_ = [("x", ctypes.c_float), ("y", ctypes.c_float), ("response", ctypes.c_float)]


_u8p = ctypes.POINTER(ctypes.c_uint8)
_i32p = ctypes.POINTER(ctypes.c_int)
_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)


def _p(a, t):
    return a.ctypes.data_as(t)


def _declare(L):
    L.oracle_fast_atan2.restype = ctypes.c_float
    L.oracle_fast_atan2.argtypes = [ctypes.c_float, ctypes.c_float]
    L.oracle_fast9_16.restype = ctypes.c_int
    L.oracle_detect_cells.restype = ctypes.c_int
    L.oracle_distribute_octree.restype = ctypes.c_int
    L.oracle_orb_extract.restype = ctypes.c_int
    L.oracle_orb_params_init.restype = ctypes.c_int
    L.oracle_orb_params_init.argtypes = [ctypes.POINTER(OrbParams), ctypes.c_int, ctypes.c_float,
                                         ctypes.c_int, ctypes.c_int, ctypes.c_int]


def orb_params(nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
    P = OrbParams()
    rc = lib().oracle_orb_params_init(ctypes.byref(P), nfeatures, scale_factor, nlevels, ini_th, min_th)
    if rc != 0:
        raise ValueError("bad extractor parameters")
    return P


def resize_linear(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty((dh, dw), np.uint8)
    lib().oracle_resize_linear_u8(_p(src, _u8p), src.shape[1], src.shape[0], src.strides[0],
                                  _p(dst, _u8p), dw, dh, dw)
    return dst


def gaussian_blur7(src):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty_like(src)
    lib().oracle_gaussian_blur7_u8(_p(src, _u8p), src.shape[1], src.shape[0], src.strides[0],
                                   _p(dst, _u8p), dst.strides[0])
    return dst


def fast9_16(img, threshold, nms=True):
    img = np.ascontiguousarray(img, np.uint8)
    cap = img.size + 1
    out = np.empty((cap, 3), np.int32)
    n = lib().oracle_fast9_16(_p(img, _u8p), img.shape[1], img.shape[0], img.strides[0],
                              int(threshold), int(bool(nms)), _p(out, _i32p), cap)
    return out[:n].copy()


def fast_atan2(y, x):
    return float(lib().oracle_fast_atan2(float(y), float(x)))


def level_size(P, w, h, level):
    lw, lh = ctypes.c_int(), ctypes.c_int()
    lib().oracle_level_size(ctypes.byref(P), w, h, level, ctypes.byref(lw), ctypes.byref(lh))
    return lw.value, lh.value


def pyramid(P, image):
    """Un-bordered pyramid levels (ORBextractor.cc:1107-1132 minus the border)."""
    image = np.ascontiguousarray(image, np.uint8)
    h, w = image.shape
    levels = [image]
    for l in range(1, P.nlevels):
        lw, lh = level_size(P, w, h, l)
        levels.append(resize_linear(levels[-1], lw, lh))
    return levels


def detect_cells(level_img, ini_th, min_th):
    img = np.ascontiguousarray(level_img, np.uint8)
    cap = (img.shape[1] // 2 + 1) * (img.shape[0] // 2 + 1) + 16
    out = np.empty((cap, 3), np.float32)
    n = lib().oracle_detect_cells(_p(img, _u8p), img.shape[1], img.shape[0], img.strides[0],
                                  int(ini_th), int(min_th), out.ctypes.data_as(ctypes.POINTER(Cand)), cap)
    assert n <= cap
    return out[:n].copy()


def distribute_octree(cands, min_x, max_x, min_y, max_y, n_target):
    cands = np.ascontiguousarray(cands, np.float32).reshape(-1, 3)
    cap = len(cands) + 8
    sel = np.empty(cap, np.int32)
    n = lib().oracle_distribute_octree(cands.ctypes.data_as(ctypes.POINTER(Cand)), len(cands),
                                       int(min_x), int(max_x), int(min_y), int(max_y), int(n_target),
                                       _p(sel, _i32p), cap)
    if n < 0:
        raise ValueError("unsupported level shape for DistributeOctTree")
    return sel[:n].copy()


def ic_moments(P, level_img, x, y):
    img = np.ascontiguousarray(level_img, np.uint8)
    m01, m10 = ctypes.c_int(), ctypes.c_int()
    lib().oracle_ic_moments(_p(img, _u8p), img.strides[0], int(x), int(y), P.umax,
                            ctypes.byref(m01), ctypes.byref(m10))
    return m01.value, m10.value


def orb_descriptor(blur_img, x, y, angle_deg):
    img = np.ascontiguousarray(blur_img, np.uint8)
    d = np.empty(32, np.uint8)
    lib().oracle_orb_descriptor(_p(img, _u8p), img.strides[0], int(x), int(y),
                                ctypes.c_float(angle_deg), _p(d, _u8p))
    return d


def orb_extract(P, image):
    """Full extractor on restated primitives.  Returns dict of SoA arrays."""
    image = np.ascontiguousarray(image, np.uint8)
    if image.size == 0:
        return _empty_result(P.nlevels)
    h, w = image.shape
    cap = P.nfeatures * 2 + 64 * P.nlevels
    while True:
        kx = np.empty(cap, np.float32); ky = np.empty(cap, np.float32)
        ka = np.empty(cap, np.float32); kr = np.empty(cap, np.float32)
        ko = np.empty(cap, np.int32); ks = np.empty(cap, np.float32)
        desc = np.empty((cap, 32), np.uint8)
        lc = np.zeros(P.nlevels, np.int32)
        n = lib().oracle_orb_extract(ctypes.byref(P), _p(image, _u8p), w, h, image.strides[0],
                                     _p(kx, _f32p), _p(ky, _f32p), _p(ka, _f32p), _p(kr, _f32p),
                                     _p(ko, _i32p), _p(ks, _f32p), _p(desc, _u8p), cap, _p(lc, _i32p))
        if n < 0:
            raise ValueError(f"oracle_orb_extract status {n}")
        if n <= cap:
            break
        cap = n
    return dict(x=kx[:n].copy(), y=ky[:n].copy(), angle=ka[:n].copy(), response=kr[:n].copy(),
                octave=ko[:n].copy(), size=ks[:n].copy(), desc=desc[:n].copy(), level_counts=lc)


def _empty_result(nlevels):
    z = np.zeros(0, np.float32)
    return dict(x=z, y=z, angle=z, response=z, octave=np.zeros(0, np.int32), size=z,
                desc=np.zeros((0, 32), np.uint8), level_counts=np.zeros(nlevels, np.int32))


# ------------------------------------------------------------------------------------------
# matcher / optimizer (slam_oracle.c)
class GridParams(ctypes.Structure):
    _fields_ = [("min_x", ctypes.c_float), ("min_y", ctypes.c_float), ("max_x", ctypes.c_float),
                ("max_y", ctypes.c_float), ("w_inv", ctypes.c_float), ("h_inv", ctypes.c_float)]


def grid_params(min_x, min_y, max_x, max_y):
    g = GridParams()
    lib().oracle_grid_params_init(ctypes.byref(g), ctypes.c_float(min_x), ctypes.c_float(min_y),
                                  ctypes.c_float(max_x), ctypes.c_float(max_y))
    return g


def descriptor_distance(a, b):
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    return int(lib().oracle_descriptor_distance(_p(a, _u8p), _p(b, _u8p)))


def distinctive_descriptor(desc):
    """MapPoint::ComputeDistinctiveDescriptors (S/src/MapPoint.cc:271-303) on the descriptors of one map point's observations, u8[N, 32]: the N x N
    DescriptorDistance table (float entries, zero diagonal), every row sorted, its element (int)(0.5 * (N - 1)) = the median, the FIRST row with the
    strictly smallest median wins.  Returns BestIdx (-1 for N = 0: the reference returns before choosing)."""
    d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32); N = len(d)
    if N == 0:
        return -1
    D = np.zeros((N, N), np.float32)
    for i in range(N):
        for j in range(i + 1, N):
            D[i, j] = D[j, i] = descriptor_distance(d[i], d[j])
    best_median, best = 2 ** 31 - 1, 0
    for i in range(N):
        v = np.sort(D[i].astype(np.int32))
        median = int(v[int(0.5 * (N - 1))])
        if median < best_median:
            best_median, best = median, i
    return best


def project_last_frame(Tcw, K4, g, scale_factors, Xw, last_octave, th, valid):
    """Projection block of SearchByProjection(Cur, Last).  Returns (valid, uv, radius, minl, maxl)."""
    Tcw = np.ascontiguousarray(Tcw, np.float32).reshape(16); K4 = np.ascontiguousarray(K4, np.float32)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    Xw = np.ascontiguousarray(Xw, np.float32).reshape(-1, 3); M = len(Xw)
    lo = np.ascontiguousarray(last_octave, np.int32)
    qv = np.ascontiguousarray(valid, np.uint8).copy()
    uv = np.zeros((M, 2), np.float32); rad = np.zeros(M, np.float32)
    mn = np.zeros(M, np.int32); mx = np.zeros(M, np.int32)
    lib().oracle_project_last_frame(_p(Tcw, _f32p), _p(K4, _f32p), ctypes.byref(g), _p(sf, _f32p), M, _p(Xw, _f32p),
                                    _p(lo, _i32p), ctypes.c_float(th), _p(qv, _u8p), _p(uv, _f32p), _p(rad, _f32p),
                                    _p(mn, _i32p), _p(mx, _i32p))
    return qv, uv, rad, mn, mx


def search_by_projection(g, f_xy, f_octave, f_angle, f_desc, q_valid, q_uv, q_radius, q_minl, q_maxl, q_angle,
                         q_desc, th_dist=100, ratio=0.0, check_ori=True, feat_match=None):
    f_xy = np.ascontiguousarray(f_xy, np.float32).reshape(-1, 2); N = len(f_xy)
    f_octave = np.ascontiguousarray(f_octave, np.int32); f_angle = np.ascontiguousarray(f_angle, np.float32)
    f_desc = np.ascontiguousarray(f_desc, np.uint8)
    q_valid = np.ascontiguousarray(q_valid, np.uint8); M = len(q_valid)
    q_uv = np.ascontiguousarray(q_uv, np.float32); q_radius = np.ascontiguousarray(q_radius, np.float32)
    q_minl = np.ascontiguousarray(q_minl, np.int32); q_maxl = np.ascontiguousarray(q_maxl, np.int32)
    q_angle = np.ascontiguousarray(q_angle, np.float32); q_desc = np.ascontiguousarray(q_desc, np.uint8)
    fm = np.full(N, -1, np.int32) if feat_match is None else np.ascontiguousarray(feat_match, np.int32).copy()
    lib().oracle_search_by_projection.restype = ctypes.c_int
    n = lib().oracle_search_by_projection(ctypes.byref(g), N, _p(f_xy, _f32p), _p(f_octave, _i32p), _p(f_angle, _f32p),
                                          _p(f_desc, _u8p), M, _p(q_valid, _u8p), _p(q_uv, _f32p), _p(q_radius, _f32p),
                                          _p(q_minl, _i32p), _p(q_maxl, _i32p), _p(q_angle, _f32p), _p(q_desc, _u8p),
                                          int(th_dist), ctypes.c_float(ratio), int(bool(check_ori)), _p(fm, _i32p))
    return n, fm


def pose_optimization(Tcw, Xw, obs, inv_sigma2, K4):
    """Optimizer::PoseOptimization.  Returns (Tcw_out f32[4,4], outlier u8[M], n_inliers)."""
    T = np.ascontiguousarray(Tcw, np.float32).reshape(16).copy()
    Xw = np.ascontiguousarray(Xw, np.float32).reshape(-1, 3); M = len(Xw)
    obs = np.ascontiguousarray(obs, np.float32).reshape(-1, 2)
    w = np.ascontiguousarray(inv_sigma2, np.float32); K4 = np.ascontiguousarray(K4, np.float32)
    out = np.zeros(max(M, 1), np.uint8)
    lib().oracle_pose_optimization.restype = ctypes.c_int
    n = lib().oracle_pose_optimization(_p(T, _f32p), M, _p(Xw, _f32p), _p(obs, _f32p), _p(w, _f32p), _p(K4, _f32p),
                                       _p(out, _u8p))
    return T.reshape(4, 4), out[:M], n


def bundle_adjust(poses, fixed, intr, points, e_kf, e_pt, e_uv, e_inv_sigma2, two_stage=True, its0=5, its1=10,
                  robust=True):
    """LocalBundleAdjustment (two_stage) / BundleAdjustment over flat arrays.  Returns dict."""
    poses = np.ascontiguousarray(poses, np.float32).reshape(-1, 16).copy(); K = len(poses)
    fixed = np.ascontiguousarray(fixed, np.uint8)
    intr = np.ascontiguousarray(intr, np.float64)
    if intr.ndim == 1:
        intr = np.tile(intr, (K, 1))
    intr = np.ascontiguousarray(intr)
    points = np.ascontiguousarray(points, np.float32).reshape(-1, 3).copy(); P = len(points)
    e_kf = np.ascontiguousarray(e_kf, np.int32); e_pt = np.ascontiguousarray(e_pt, np.int32); E = len(e_kf)
    e_uv = np.ascontiguousarray(e_uv, np.float32); w = np.ascontiguousarray(e_inv_sigma2, np.float32)
    chi2 = np.zeros(max(E, 1)); dok = np.zeros(max(E, 1), np.uint8); outl = np.zeros(max(E, 1), np.uint8)
    stats = np.zeros(2, np.int32)
    lib().oracle_bundle_adjust.restype = ctypes.c_int
    rc = lib().oracle_bundle_adjust(K, _p(poses, _f32p), _p(fixed, _u8p), _p(intr, _f64p), P, _p(points, _f32p), E,
                                    _p(e_kf, _i32p), _p(e_pt, _i32p), _p(e_uv, _f32p), _p(w, _f32p),
                                    int(bool(two_stage)), int(its0), int(its1), int(bool(robust)), None,
                                    _p(chi2, _f64p), _p(dok, _u8p), _p(outl, _u8p), _p(stats, _i32p))
    return dict(poses=poses.reshape(K, 4, 4), points=points, chi2=chi2[:E], depth_ok=dok[:E], outlier=outl[:E],
                lm_iterations=int(stats[0]), lm_trials=int(stats[1]), rc=rc)


def undistort_points(K4, dist5, xy):
    """Frame::UndistortKeyPoints (Frame.cc:404-434): cv::undistortPoints(pts, K, dist, R=I, P=K)."""
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    out = np.empty_like(xy)
    K4 = np.ascontiguousarray(K4, np.float32); d = np.zeros(5, np.float32); d[:len(dist5)] = dist5
    lib().oracle_undistort_points(_p(K4, _f32p), _p(d, _f32p), len(xy), _p(xy, _f32p), _p(out, _f32p))
    return out


def is_in_frustum(Tcw, Ow, K4, bounds4, log_scale_factor, cos_limit, Xw, normal, mf_min_dist, mf_max_dist):
    """Frame::isInFrustum (Frame.cc:269-325) over flat arrays -> (in_view u8[M], proj f32[M,2], level i32[M], view_cos f32[M])."""
    Xw = np.ascontiguousarray(Xw, np.float32).reshape(-1, 3); M = len(Xw)
    normal = np.ascontiguousarray(normal, np.float32).reshape(-1, 3)
    T = np.ascontiguousarray(Tcw, np.float32).reshape(16); Ow = np.ascontiguousarray(Ow, np.float32)
    K4 = np.ascontiguousarray(K4, np.float32); b = np.ascontiguousarray(bounds4, np.float32)
    mn = np.ascontiguousarray(mf_min_dist, np.float32); mx = np.ascontiguousarray(mf_max_dist, np.float32)
    iv = np.zeros(M, np.uint8); uv = np.zeros((M, 2), np.float32); lv = np.zeros(M, np.int32); vc = np.zeros(M, np.float32)
    L = lib()
    L.oracle_is_in_frustum.argtypes = [_f32p, _f32p, _f32p, _f32p, ctypes.c_float, ctypes.c_float, ctypes.c_int, _f32p, _f32p, _f32p, _f32p,
                                       _u8p, _f32p, _i32p, _f32p]
    L.oracle_is_in_frustum(_p(T, _f32p), _p(Ow, _f32p), _p(K4, _f32p), _p(b, _f32p), float(log_scale_factor), float(cos_limit), M, _p(Xw, _f32p),
                           _p(normal, _f32p), _p(mn, _f32p), _p(mx, _f32p), _p(iv, _u8p), _p(uv, _f32p), _p(lv, _i32p), _p(vc, _f32p))
    return iv, uv, lv, vc


# ---------------------------------------------------------------------------------------------------------------------
# KeyFrame / Sim3 projection family (slam_oracle.c: oracle_project_points, oracle_search_best_in_window, oracle_search_by_projection_kf)
PROJ_TWO_STEP, PROJ_NO_DEPTH, PROJ_FRAME_BOUNDS, PROJ_FRAME_UV, PROJ_DIST_CAMERA, PROJ_CHECK_NORMAL, PROJ_LEVEL_PLUS1 = 1, 2, 4, 8, 16, 32, 64


class Projection(ctypes.Structure):
    """= oracle_projection = orbm_projection (include/orbslamm_b200.h)"""
    _fields_ = [("R", ctypes.c_float * 9), ("t", ctypes.c_float * 3), ("R2", ctypes.c_float * 9), ("t2", ctypes.c_float * 3),
                ("Ow", ctypes.c_float * 3), ("fx", ctypes.c_float), ("fy", ctypes.c_float), ("cx", ctypes.c_float), ("cy", ctypes.c_float),
                ("min_x", ctypes.c_float), ("min_y", ctypes.c_float), ("max_x", ctypes.c_float), ("max_y", ctypes.c_float),
                ("log_scale_factor", ctypes.c_float), ("th", ctypes.c_float), ("flags", ctypes.c_int32)]


def make_projection(R, t, K4, bounds4, log_scale_factor, th, flags, Ow=None, R2=None, t2=None):
    V = Projection()
    V.R[:] = [float(x) for x in np.asarray(R, np.float32).ravel()]; V.t[:] = [float(x) for x in np.asarray(t, np.float32).ravel()]
    if R2 is not None:
        V.R2[:] = [float(x) for x in np.asarray(R2, np.float32).ravel()]; V.t2[:] = [float(x) for x in np.asarray(t2, np.float32).ravel()]
    if Ow is not None:
        V.Ow[:] = [float(x) for x in np.asarray(Ow, np.float32).ravel()]
    V.fx, V.fy, V.cx, V.cy = [float(x) for x in np.asarray(K4, np.float32)]
    V.min_x, V.min_y, V.max_x, V.max_y = [float(x) for x in np.asarray(bounds4, np.float32)]
    V.log_scale_factor = float(np.float32(log_scale_factor)); V.th = float(np.float32(th)); V.flags = int(flags)
    return V


def project_points(V, scale_factors, Xw, normal, mf_min, mf_max, valid):
    """Returns (valid, uv, radius, minl, maxl, level)."""
    sf = np.ascontiguousarray(scale_factors, np.float32)
    Xw = np.ascontiguousarray(Xw, np.float32).reshape(-1, 3); M = len(Xw)
    nrm = None if normal is None else np.ascontiguousarray(normal, np.float32).reshape(M, 3)
    mn = np.ascontiguousarray(mf_min, np.float32); mx = np.ascontiguousarray(mf_max, np.float32)
    qv = np.ascontiguousarray(valid, np.uint8).copy()
    uv = np.zeros((M, 2), np.float32); rad = np.zeros(M, np.float32)
    l0 = np.zeros(M, np.int32); l1 = np.zeros(M, np.int32); lv = np.zeros(M, np.int32)
    L = lib()
    L.oracle_project_points.restype = None
    L.oracle_project_points.argtypes = [ctypes.c_void_p, _f32p, ctypes.c_int, ctypes.c_int, _f32p, ctypes.c_void_p, _f32p, _f32p, _u8p, _f32p, _f32p,
                                        _i32p, _i32p, _i32p]
    L.oracle_project_points(ctypes.byref(V), _p(sf, _f32p), len(sf), M, _p(Xw, _f32p), None if nrm is None else nrm.ctypes.data, _p(mn, _f32p),
                            _p(mx, _f32p), _p(qv, _u8p), _p(uv, _f32p), _p(rad, _f32p), _p(l0, _i32p), _p(l1, _i32p), _p(lv, _i32p))
    return qv, uv, rad, l0, l1, lv


def search_best_in_window(g, win_origin2, f_xy, f_octave, f_desc, q_valid, q_uv, q_radius, q_minl, q_maxl, q_desc, th_dist,
                          inv_level_sigma2=None, chi2_gate=5.99):
    """Returns (best_idx, best_dist) per query (Fuse / SearchBySim3 candidate loop)."""
    f_xy = np.ascontiguousarray(f_xy, np.float32).reshape(-1, 2); N = len(f_xy)
    f_octave = np.ascontiguousarray(f_octave, np.int32); f_desc = np.ascontiguousarray(f_desc, np.uint8)
    q_valid = np.ascontiguousarray(q_valid, np.uint8); M = len(q_valid)
    q_uv = np.ascontiguousarray(q_uv, np.float32); q_radius = np.ascontiguousarray(q_radius, np.float32)
    q_minl = np.ascontiguousarray(q_minl, np.int32); q_maxl = np.ascontiguousarray(q_maxl, np.int32); q_desc = np.ascontiguousarray(q_desc, np.uint8)
    wo = None if win_origin2 is None else np.ascontiguousarray(win_origin2, np.float32)
    inv = None if inv_level_sigma2 is None else np.ascontiguousarray(inv_level_sigma2, np.float32)
    bi = np.zeros(M, np.int32); bd = np.zeros(M, np.int32)
    L = lib()
    L.oracle_search_best_in_window.restype = None
    L.oracle_search_best_in_window.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, _f32p, _i32p, _u8p, ctypes.c_int, _u8p, _f32p, _f32p,
                                               _i32p, _i32p, _u8p, ctypes.c_int, ctypes.c_void_p, ctypes.c_double, _i32p, _i32p]
    L.oracle_search_best_in_window(ctypes.byref(g), None if wo is None else wo.ctypes.data, N, _p(f_xy, _f32p), _p(f_octave, _i32p), _p(f_desc, _u8p),
                                   M, _p(q_valid, _u8p), _p(q_uv, _f32p), _p(q_radius, _f32p), _p(q_minl, _i32p), _p(q_maxl, _i32p), _p(q_desc, _u8p),
                                   int(th_dist), None if inv is None else inv.ctypes.data, float(chi2_gate), _p(bi, _i32p), _p(bd, _i32p))
    return bi, bd


def search_by_projection_kf(g, win_origin2, f_xy, f_octave, f_angle, f_desc, q_valid, q_uv, q_radius, q_minl, q_maxl, q_angle,
                            q_desc, th_dist=100, ratio=0.0, check_ori=True, feat_match=None):
    """search_by_projection with a KeyFrame's integer window origin (None = the grid's)."""
    f_xy = np.ascontiguousarray(f_xy, np.float32).reshape(-1, 2); N = len(f_xy)
    f_octave = np.ascontiguousarray(f_octave, np.int32); f_angle = np.ascontiguousarray(f_angle, np.float32)
    f_desc = np.ascontiguousarray(f_desc, np.uint8)
    q_valid = np.ascontiguousarray(q_valid, np.uint8); M = len(q_valid)
    q_uv = np.ascontiguousarray(q_uv, np.float32); q_radius = np.ascontiguousarray(q_radius, np.float32)
    q_minl = np.ascontiguousarray(q_minl, np.int32); q_maxl = np.ascontiguousarray(q_maxl, np.int32)
    q_angle = np.ascontiguousarray(q_angle, np.float32); q_desc = np.ascontiguousarray(q_desc, np.uint8)
    wo = None if win_origin2 is None else np.ascontiguousarray(win_origin2, np.float32)
    fm = np.full(N, -1, np.int32) if feat_match is None else np.ascontiguousarray(feat_match, np.int32).copy()
    L = lib()
    L.oracle_search_by_projection_kf.restype = ctypes.c_int
    L.oracle_search_by_projection_kf.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, _f32p, _i32p, _f32p, _u8p, ctypes.c_int, _u8p, _f32p,
                                                 _f32p, _i32p, _i32p, _f32p, _u8p, ctypes.c_int, ctypes.c_float, ctypes.c_int, _i32p]
    n = L.oracle_search_by_projection_kf(ctypes.byref(g), None if wo is None else wo.ctypes.data, N, _p(f_xy, _f32p), _p(f_octave, _i32p),
                                         _p(f_angle, _f32p), _p(f_desc, _u8p), M, _p(q_valid, _u8p), _p(q_uv, _f32p), _p(q_radius, _f32p),
                                         _p(q_minl, _i32p), _p(q_maxl, _i32p), _p(q_angle, _f32p), _p(q_desc, _u8p), int(th_dist), float(ratio),
                                         int(bool(check_ori)), _p(fm, _i32p))
    return n, fm


def optimize_sim3(sim3, valid, P1c, P2c, obs1, obs2, w1, w2, K1, K2, th2=10.0, fix_scale=False):
    """Optimizer::OptimizeSim3 (Optimizer.cc:1348-1543).  sim3 = (qx, qy, qz, qw, tx, ty, tz, s) of g2oS12.
    Returns dict(sim3, inlier u8[N], n_in, lm_iterations, lm_trials)."""
    S = np.ascontiguousarray(sim3, np.float64).copy()
    v = np.ascontiguousarray(valid, np.uint8); N = len(v)
    a = [np.ascontiguousarray(x, np.float32) for x in (P1c, P2c, obs1, obs2, w1, w2, K1, K2)]
    inl = np.zeros(N, np.uint8); st = np.zeros(2, np.int32)
    L = lib()
    L.oracle_optimize_sim3.restype = ctypes.c_int
    L.oracle_optimize_sim3.argtypes = [_f64p, ctypes.c_int, _u8p] + [_f32p] * 8 + [ctypes.c_float, ctypes.c_int, _u8p, _i32p]
    n = L.oracle_optimize_sim3(_p(S, _f64p), N, _p(v, _u8p), *[_p(x, _f32p) for x in a], float(th2), int(bool(fix_scale)), _p(inl, _u8p), _p(st, _i32p))
    return dict(sim3=S, inlier=inl, n_in=int(n), lm_iterations=int(st[0]), lm_trials=int(st[1]))


# ---------------------------------------------------------------------------------------------------------------------
# vocabulary-bucket matchers and SearchForInitialization
class FeatVec(ctypes.Structure):
    _fields_ = [("n_nodes", ctypes.c_int), ("nodes", ctypes.c_void_p), ("start", ctypes.c_void_p), ("items", ctypes.c_void_p)]


class Epipolar(ctypes.Structure):
    _fields_ = [("xy1", ctypes.c_void_p), ("xy2", ctypes.c_void_p), ("octave2", ctypes.c_void_p), ("F12", ctypes.c_void_p), ("ex", ctypes.c_float),
                ("ey", ctypes.c_float), ("scale_factors2", ctypes.c_void_p), ("level_sigma2_2", ctypes.c_void_p)]


def feature_vector(node_of_feature):
    """DBoW2::FeatureVector as CSR from the node id of every feature (-1 = none): ascending node ids, features in index order."""
    n = np.asarray(node_of_feature, np.int64)
    idx = np.where(n >= 0)[0]
    order = idx[np.argsort(n[idx], kind="stable")]
    nodes, counts = np.unique(n[order], return_counts=True)
    start = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    return dict(nodes=nodes.astype(np.int32), start=start, items=order.astype(np.int32))


def _fv(fv):
    keep = [np.ascontiguousarray(fv[k], np.int32) for k in ("nodes", "start", "items")]
    return FeatVec(len(keep[0]), keep[0].ctypes.data, keep[1].ctypes.data, keep[2].ctypes.data), keep


def search_by_bow(mode, desc1, angle1, elig1, fv1, desc2, angle2, elig2, fv2, ratio, check_ori, epi=None):
    """mode 0: SearchByBoW (both variants); mode 1: SearchForTriangulation (epi = dict xy1, xy2, octave2, F12, ex, ey, scale_factors2, level_sigma2_2).
    Returns (nmatches, match12[N1])."""
    d1 = np.ascontiguousarray(desc1, np.uint8); d2 = np.ascontiguousarray(desc2, np.uint8)
    a1 = np.ascontiguousarray(angle1, np.float32); a2 = np.ascontiguousarray(angle2, np.float32)
    e1 = np.ascontiguousarray(elig1, np.uint8); e2 = np.ascontiguousarray(elig2, np.uint8)
    f1, k1 = _fv(fv1); f2, k2 = _fv(fv2)
    m = np.full(len(d1), -1, np.int32)
    ep = None
    if epi is not None:
        ka = [np.ascontiguousarray(epi["xy1"], np.float32), np.ascontiguousarray(epi["xy2"], np.float32), np.ascontiguousarray(epi["octave2"], np.int32),
              np.ascontiguousarray(epi["F12"], np.float32), np.ascontiguousarray(epi["scale_factors2"], np.float32), np.ascontiguousarray(epi["level_sigma2_2"], np.float32)]
        ep = Epipolar(ka[0].ctypes.data, ka[1].ctypes.data, ka[2].ctypes.data, ka[3].ctypes.data, float(epi["ex"]), float(epi["ey"]), ka[4].ctypes.data, ka[5].ctypes.data)
    L = lib()
    L.oracle_search_by_bow.restype = ctypes.c_int
    L.oracle_search_by_bow.argtypes = [ctypes.c_int, ctypes.c_int, _u8p, _f32p, _u8p, ctypes.c_void_p, ctypes.c_int, _u8p, _f32p, _u8p, ctypes.c_void_p,
                                       ctypes.c_float, ctypes.c_int, ctypes.c_void_p, _i32p]
    n = L.oracle_search_by_bow(int(mode), len(d1), _p(d1, _u8p), _p(a1, _f32p), _p(e1, _u8p), ctypes.byref(f1), len(d2), _p(d2, _u8p), _p(a2, _f32p),
                               _p(e2, _u8p), ctypes.byref(f2), float(ratio), int(bool(check_ori)), None if ep is None else ctypes.byref(ep), _p(m, _i32p))
    return n, m


def search_for_initialization(g, f1, f2, prev_matched, window=100, ratio=0.9, check_ori=True):
    """f1 / f2: dicts x, y, octave, angle, desc.  Returns (nmatches, matches12, prev_matched updated)."""
    xy1 = np.ascontiguousarray(np.stack([f1["x"], f1["y"]], 1), np.float32); xy2 = np.ascontiguousarray(np.stack([f2["x"], f2["y"]], 1), np.float32)
    o1 = np.ascontiguousarray(f1["octave"], np.int32); o2 = np.ascontiguousarray(f2["octave"], np.int32)
    a1 = np.ascontiguousarray(f1["angle"], np.float32); a2 = np.ascontiguousarray(f2["angle"], np.float32)
    d1 = np.ascontiguousarray(f1["desc"], np.uint8); d2 = np.ascontiguousarray(f2["desc"], np.uint8)
    pm = np.ascontiguousarray(prev_matched, np.float32).copy()
    m = np.full(len(xy1), -1, np.int32)
    L = lib()
    L.oracle_search_for_initialization.restype = ctypes.c_int
    L.oracle_search_for_initialization.argtypes = [ctypes.c_void_p, ctypes.c_int, _f32p, _i32p, _f32p, _u8p, ctypes.c_int, _f32p, _i32p, _f32p, _u8p, _f32p,
                                                   ctypes.c_int, ctypes.c_float, ctypes.c_int, _i32p]
    n = L.oracle_search_for_initialization(ctypes.byref(g), len(xy1), _p(xy1, _f32p), _p(o1, _i32p), _p(a1, _f32p), _p(d1, _u8p), len(xy2), _p(xy2, _f32p),
                                           _p(o2, _i32p), _p(a2, _f32p), _p(d2, _u8p), _p(pm, _f32p), int(window), float(ratio), int(bool(check_ori)), _p(m, _i32p))
    return n, m, pm


def grid_build(g, f_xy):
    """Frame::AssignFeaturesToGrid: (cell_start[3073], cell_items[N]) with cell = ix * 48 + iy, ascending index inside a cell."""
    xy = np.ascontiguousarray(f_xy, np.float32).reshape(-1, 2); N = len(xy)
    cs = np.zeros(64 * 48 + 1, np.int32); ci = np.full(max(N, 1), -1, np.int32)
    L = lib()
    L.oracle_grid_build.restype = None
    L.oracle_grid_build.argtypes = [ctypes.c_void_p, ctypes.c_int, _f32p, _i32p, _i32p]
    L.oracle_grid_build(ctypes.byref(g), N, _p(xy, _f32p), _p(cs, _i32p), _p(ci, _i32p))
    return cs, ci[:N]


def features_in_area(g, cell_start, cell_items, f_xy, f_octave, x, y, r, min_level=-1, max_level=-1):
    """Frame::GetFeaturesInArea: indices in the reference's visiting order."""
    xy = np.ascontiguousarray(f_xy, np.float32).reshape(-1, 2); oc = np.ascontiguousarray(f_octave, np.int32)
    cs = np.ascontiguousarray(cell_start, np.int32); ci = np.ascontiguousarray(cell_items, np.int32)
    out = np.zeros(max(len(xy), 1), np.int32)
    L = lib()
    L.oracle_features_in_area.restype = ctypes.c_int
    L.oracle_features_in_area.argtypes = [ctypes.c_void_p, _i32p, _i32p, _f32p, _i32p, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_int, _i32p]
    n = L.oracle_features_in_area(ctypes.byref(g), _p(cs, _i32p), _p(ci, _i32p), _p(xy, _f32p), _p(oc, _i32p), float(x), float(y), float(r), int(min_level),
                                  int(max_level), _p(out, _i32p))
    return out[:n].copy()


# ---------------------------------------------------------------------------------------------------------------------
# DBoW2 vocabulary transform
def vocab_transform(vocab, desc, levelsup=4):
    """vocab: dict L, node_desc [n,32] u8, child_start [n+1], child_ids [n-1], word_id [n], weight [n] f64 (orbslamm_b200.vocabulary format).
    Returns dict(word_of, node_of, bow_ids, bow_vals, fv=dict(nodes, start, items))."""
    d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32); N = len(d)
    nd = np.ascontiguousarray(vocab["node_desc"], np.uint8); cs = np.ascontiguousarray(vocab["child_start"], np.int32)
    ci = np.ascontiguousarray(vocab["child_ids"], np.int32); wi = np.ascontiguousarray(vocab["word_id"], np.int32); ww = np.ascontiguousarray(vocab["weight"], np.float64)
    M = max(N, 1)
    wo = np.zeros(M, np.int32); no = np.zeros(M, np.int32); bi = np.zeros(M, np.int32); bv = np.zeros(M, np.float64)
    fn = np.zeros(M, np.int32); fs = np.zeros(M + 1, np.int32); fi = np.zeros(M, np.int32); fc = ctypes.c_int(0)
    L = lib()
    L.oracle_vocab_transform.restype = ctypes.c_int
    L.oracle_vocab_transform.argtypes = [ctypes.c_int, _u8p, _i32p, _i32p, _i32p, _f64p, ctypes.c_int, _u8p, ctypes.c_int, _i32p, _i32p, _i32p, _f64p, _i32p, _i32p,
                                         _i32p, ctypes.POINTER(ctypes.c_int)]
    nb = L.oracle_vocab_transform(int(vocab["L"]), _p(nd, _u8p), _p(cs, _i32p), _p(ci, _i32p), _p(wi, _i32p), _p(ww, _f64p), N, _p(d, _u8p), int(levelsup),
                                  _p(wo, _i32p), _p(no, _i32p), _p(bi, _i32p), _p(bv, _f64p), _p(fn, _i32p), _p(fs, _i32p), _p(fi, _i32p), ctypes.byref(fc))
    nf = fc.value
    return dict(word_of=wo[:N], node_of=no[:N], bow_ids=bi[:nb].copy(), bow_vals=bv[:nb].copy(),
                fv=dict(nodes=fn[:nf].copy(), start=fs[:nf + 1].copy(), items=fi[:fs[nf]].copy()))


# ---------------------------------------------------------------------------------------------------------------------
# Sim3Solver: thresholds, FromCameraToImage, CheckInliers over a batch of RANSAC hypotheses
def sim3_prepare(X3Dc1, X3Dc2, oct1, oct2, level_sigma2, K1, K2):
    """The constructor's per-correspondence data: (max_err1, max_err2 i32[N], P1im1, P2im2 f32[N,2])."""
    X1 = np.ascontiguousarray(X3Dc1, np.float32); X2 = np.ascontiguousarray(X3Dc2, np.float32); N = len(X1)
    o1 = np.ascontiguousarray(oct1, np.int32); o2 = np.ascontiguousarray(oct2, np.int32); ls = np.ascontiguousarray(level_sigma2, np.float32)
    k1 = np.ascontiguousarray(K1, np.float32); k2 = np.ascontiguousarray(K2, np.float32)
    m1 = np.zeros(N, np.int32); m2 = np.zeros(N, np.int32); p1 = np.zeros((N, 2), np.float32); p2 = np.zeros((N, 2), np.float32)
    L = lib()
    L.oracle_sim3_max_error(N, _p(o1, _i32p), _p(ls, _f32p), _p(m1, _i32p)); L.oracle_sim3_max_error(N, _p(o2, _i32p), _p(ls, _f32p), _p(m2, _i32p))
    L.oracle_sim3_from_camera_to_image(N, _p(X1, _f32p), _p(k1, _f32p), _p(p1, _f32p)); L.oracle_sim3_from_camera_to_image(N, _p(X2, _f32p), _p(k2, _f32p), _p(p2, _f32p))
    return m1, m2, p1, p2


def sim3_check_inliers(T12, T21, X3Dc1, X3Dc2, P1im1, P2im2, max_err1, max_err2, K1, K2):
    """Sim3Solver::CheckInliers for n_hyp hypotheses: returns (inliers u8[n_hyp, N], n_inliers i32[n_hyp])."""
    a = [np.ascontiguousarray(x, np.float32) for x in (T12, T21, X3Dc1, X3Dc2, P1im1, P2im2)]
    nh = a[0].reshape(-1, 16).shape[0]; N = len(a[2])
    m1 = np.ascontiguousarray(max_err1, np.int32); m2 = np.ascontiguousarray(max_err2, np.int32)
    k1 = np.ascontiguousarray(K1, np.float32); k2 = np.ascontiguousarray(K2, np.float32)
    inl = np.zeros((nh, N), np.uint8); n = np.zeros(nh, np.int32)
    lib().oracle_sim3_check_inliers(nh, _p(a[0], _f32p), _p(a[1], _f32p), N, _p(a[2], _f32p), _p(a[3], _f32p), _p(a[4], _f32p), _p(a[5], _f32p), _p(m1, _i32p),
                                    _p(m2, _i32p), _p(k1, _f32p), _p(k2, _f32p), _p(inl, _u8p), _p(n, _i32p))
    return inl, n


def optimize_pose_graph(sim3, fixed, e_i, e_j, e_meas, fix_scale=False, iterations=20, lambda_init=1e-16):
    """Numeric core of Optimizer::OptimizeEssentialGraph (Optimizer.cc:804-1067).  sim3 [K,8], e_meas [E,8] = (qx qy qz qw tx ty tz s).
    Returns dict(sim3, lm_iterations, lm_trials, chol_failures)."""
    S = np.ascontiguousarray(sim3, np.float64).reshape(-1, 8).copy(); K = len(S)
    fx = np.ascontiguousarray(fixed, np.uint8); ei = np.ascontiguousarray(e_i, np.int32); ej = np.ascontiguousarray(e_j, np.int32)
    em = np.ascontiguousarray(e_meas, np.float64).reshape(-1, 8); st = np.zeros(3, np.int32)
    L = lib()
    L.oracle_optimize_pose_graph.restype = ctypes.c_int
    L.oracle_optimize_pose_graph.argtypes = [ctypes.c_int, _f64p, _u8p, ctypes.c_int, _i32p, _i32p, _f64p, ctypes.c_int, ctypes.c_int, ctypes.c_double, _i32p]
    L.oracle_optimize_pose_graph(K, _p(S, _f64p), _p(fx, _u8p), len(ei), _p(ei, _i32p), _p(ej, _i32p), _p(em, _f64p), int(bool(fix_scale)), int(iterations),
                                 float(lambda_init), _p(st, _i32p))
    return dict(sim3=S, lm_iterations=int(st[0]), lm_trials=int(st[1]), chol_failures=int(st[2]))
