"""oracle/ref_build.py -- build and load oracle/_ref/libref_orbextractor.so: the REFERENCE's own ORBextractor.cc
compiled unmodified (from /root/reference, never copied) against the OpenCV stand-in oracle/cvshim.

TEST INFRASTRUCTURE ONLY.  /root/reference exists only in the build container; on the GPU box the prebuilt .so
travels with the snapshot (oracle/_ref/ is git-ignored, not gpurun-ignored) and `available()` reports what is there.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libref_orbextractor.so")
REF_ROOT = "/root/reference/SingleRobotScenario"
_LIB = None


def build():
    """Run oracle/Makefile when the reference sources are present; returns the .so path or None."""
    if os.path.exists(os.path.join(REF_ROOT, "src", "ORBextractor.cc")):
        subprocess.check_call(["make", "-s", "-C", _HERE, f"REF={REF_ROOT}"])
    return _SO if os.path.exists(_SO) else None


def available():
    return os.path.exists(_SO)


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(_SO)
        L.ref_orbx_create.restype = ctypes.c_void_p
        L.ref_orbx_create.argtypes = [ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.ref_orbx_destroy.argtypes = [ctypes.c_void_p]
        L.ref_orbx_extract.restype = ctypes.c_int
        L.ref_orbx_extract.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 7
        L.ref_orbx_tables.argtypes = [ctypes.c_void_p] + [ctypes.c_void_p] * 4
        L.ref_orbx_pyramid_level.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.ref_bump_nodes.restype = ctypes.c_long
        _LIB = L
    return _LIB


class RefORBextractor:
    """iORB_SLAM::ORBextractor of the reference (ORBextractor.h:45-111), object code of the reference's own source."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, ascending_heap=True):
        """ascending_heap: quad-tree nodes get ascending, never reused addresses (oracle/ref_bump_alloc.cc), which makes the
        reference's heap-address tie-break (ORBextractor.cc:684) equal to 'later-created node first'."""
        self.L = lib()
        self.L.ref_bump_enable(1 if ascending_heap else 0)
        self.nlevels = nlevels
        self.cap = 4 * nfeatures + 64
        self.h = self.L.ref_orbx_create(nfeatures, scale_factor, nlevels, ini_th, min_th)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_orbx_destroy(self.h)
            self.h = None

    def __call__(self, image):
        img = np.ascontiguousarray(image, np.uint8)
        h, w = img.shape
        c = self.cap
        x = np.zeros(c, np.float32); y = np.zeros(c, np.float32); ang = np.zeros(c, np.float32); resp = np.zeros(c, np.float32)
        octv = np.zeros(c, np.int32); size = np.zeros(c, np.float32); desc = np.zeros((c, 32), np.uint8)
        n = self.L.ref_orbx_extract(self.h, img.ctypes.data, w, h, w, c, x.ctypes.data, y.ctypes.data, ang.ctypes.data, resp.ctypes.data,
                                    octv.ctypes.data, size.ctypes.data, desc.ctypes.data)
        assert n <= c
        return dict(x=x[:n], y=y[:n], angle=ang[:n], response=resp[:n], octave=octv[:n], size=size[:n], desc=desc[:n])

    def tables(self):
        a = [np.zeros(self.nlevels, np.float32) for _ in range(4)]
        self.L.ref_orbx_tables(self.h, *[t.ctypes.data for t in a])
        return dict(scale=a[0], inv_scale=a[1], sigma2=a[2], inv_sigma2=a[3])

    def pyramid_level(self, level):
        w = ctypes.c_int(); h = ctypes.c_int()
        self.L.ref_orbx_pyramid_level(self.h, level, ctypes.byref(w), ctypes.byref(h), None, 0)
        out = np.zeros((h.value, w.value), np.uint8)
        self.L.ref_orbx_pyramid_level(self.h, level, ctypes.byref(w), ctypes.byref(h), out.ctypes.data, w.value)
        return out


# ---------------------------------------------------------------------------------------------------------------------
# the reference's own ORBmatcher.cc (oracle/_ref/libref_orbmatcher.so, built by the same Makefile against oracle/slamshim)
_SOM = os.path.join(_HERE, "_ref", "libref_orbmatcher.so")
_LIBM = None


def matcher_available():
    return os.path.exists(_SOM)


def matcher_lib():
    global _LIBM
    if _LIBM is None:
        L = ctypes.CDLL(_SOM)
        f, i, vp = ctypes.c_float, ctypes.c_int, ctypes.c_void_p
        L.ref_orbm_set_camera.argtypes = [f] * 8
        L.ref_orbm_descriptor_distance.argtypes = [vp, vp]
        L.ref_orbm_search_last_frame.argtypes = [f, i, f, vp, vp, i, i, vp, vp, vp, vp, i, vp, vp, vp, vp, vp, vp]
        L.ref_orbm_search_local_points.argtypes = [f, f, vp, i, i, vp, vp, vp, vp, vp, i, vp, vp, vp, vp, vp, vp]
        _LIBM = L
    return _LIBM


def _c(a, dt):
    return np.ascontiguousarray(a, dt)


def ref_descriptor_distance(a, b):
    """ORBmatcher::DescriptorDistance (ORBmatcher.cc:1649-1665), reference object code."""
    a = _c(a, np.uint8); b = _c(b, np.uint8)
    return matcher_lib().ref_orbm_descriptor_distance(a.ctypes.data, b.ctypes.data)


def ref_search_last_frame(K4, bounds4, Tcw, scale_factors, cur, last, Xw, valid, th=15.0, nnratio=0.9, check_ori=True):
    """ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono=true) (ORBmatcher.cc:1330-1472), reference object code.
    cur / last: dicts x, y, octave, angle, desc.  Returns (nmatches, feat_match[N] = last-frame slot or -1)."""
    L = matcher_lib()
    L.ref_orbm_set_camera(*[float(v) for v in K4], float(bounds4[0]), float(bounds4[1]), float(bounds4[2]), float(bounds4[3]))
    T = _c(Tcw, np.float32).reshape(16); sf = _c(scale_factors, np.float32)
    fxy = _c(np.stack([cur["x"], cur["y"]], 1), np.float32); N = len(fxy); M = len(last["x"])
    fo = _c(cur["octave"], np.int32); fa = _c(cur["angle"], np.float32); fd = _c(cur["desc"], np.uint8)
    v = _c(valid, np.uint8); X = _c(Xw, np.float32); lo = _c(last["octave"], np.int32); la = _c(last["angle"], np.float32); ld = _c(last["desc"], np.uint8)
    fm = np.full(N, -1, np.int32)
    n = L.ref_orbm_search_last_frame(float(nnratio), int(bool(check_ori)), float(th), T.ctypes.data, sf.ctypes.data, len(sf), N, fxy.ctypes.data, fo.ctypes.data,
                                     fa.ctypes.data, fd.ctypes.data, M, v.ctypes.data, X.ctypes.data, lo.ctypes.data, la.ctypes.data, ld.ctypes.data, fm.ctypes.data)
    return n, fm


def ref_search_local_points(K4, bounds4, scale_factors, cur, in_view, proj_xy, level, view_cos, q_desc, th=1.0, nnratio=0.8, held=None):
    """ORBmatcher::SearchByProjection(F, vpMapPoints, th) (ORBmatcher.cc:45-129), reference object code.
    Returns (nmatches, feat_match[N] = map point index, -1 = free, -2 = feature that already held a map point)."""
    L = matcher_lib()
    L.ref_orbm_set_camera(*[float(v) for v in K4], float(bounds4[0]), float(bounds4[1]), float(bounds4[2]), float(bounds4[3]))
    sf = _c(scale_factors, np.float32)
    fxy = _c(np.stack([cur["x"], cur["y"]], 1), np.float32); N = len(fxy); M = len(in_view)
    fo = _c(cur["octave"], np.int32); fa = _c(cur["angle"], np.float32); fd = _c(cur["desc"], np.uint8)
    hv = _c(held if held is not None else np.zeros(N, np.uint8), np.uint8)
    iv = _c(in_view, np.uint8); uv = _c(proj_xy, np.float32); lv = _c(level, np.int32); vc = _c(view_cos, np.float32); qd = _c(q_desc, np.uint8)
    fm = np.full(N, -1, np.int32)
    n = L.ref_orbm_search_local_points(float(nnratio), float(th), sf.ctypes.data, len(sf), N, fxy.ctypes.data, fo.ctypes.data, fa.ctypes.data, fd.ctypes.data,
                                       hv.ctypes.data, M, iv.ctypes.data, uv.ctypes.data, lv.ctypes.data, vc.ctypes.data, qd.ctypes.data, fm.ctypes.data)
    return n, fm


# ---- KeyFrame / Sim3 projection family (reference object code; see oracle/ref_matcher_capi.cc) -------------------------------------------
class _RefKF(ctypes.Structure):
    _fields_ = [("N", ctypes.c_int), ("xy", ctypes.c_void_p), ("octave", ctypes.c_void_p), ("angle", ctypes.c_void_p), ("desc", ctypes.c_void_p),
                ("scale_factors", ctypes.c_void_p), ("inv_level_sigma2", ctypes.c_void_p), ("nlevels", ctypes.c_int),
                ("K4", ctypes.c_void_p), ("grid_bounds4", ctypes.c_void_p), ("Tcw", ctypes.c_void_p)]


def _refkf(kf):
    """kf: dict xy [N,2], octave, angle, desc, scale_factors, inv_level_sigma2, K4, grid_bounds4, Tcw (optional).  Returns (struct, keepalive)."""
    keep = dict(xy=_c(kf["xy"], np.float32), octave=_c(kf["octave"], np.int32), angle=_c(kf["angle"], np.float32), desc=_c(kf["desc"], np.uint8),
                sf=_c(kf["scale_factors"], np.float32), inv=_c(kf["inv_level_sigma2"], np.float32), K4=_c(kf["K4"], np.float32),
                gb=_c(kf["grid_bounds4"], np.float32), T=_c(kf.get("Tcw", np.eye(4)), np.float32).reshape(16))
    s = _RefKF(len(keep["xy"]), keep["xy"].ctypes.data, keep["octave"].ctypes.data, keep["angle"].ctypes.data, keep["desc"].ctypes.data,
               keep["sf"].ctypes.data, keep["inv"].ctypes.data, len(keep["sf"]), keep["K4"].ctypes.data, keep["gb"].ctypes.data, keep["T"].ctypes.data)
    return s, keep


def _pts(pts):
    """pts: dict Xw [M,3], normal [M,3], mf_min, mf_max, desc [M,32]"""
    return (_c(pts["Xw"], np.float32), _c(pts["normal"], np.float32), _c(pts["mf_min"], np.float32), _c(pts["mf_max"], np.float32), _c(pts["desc"], np.uint8))


def ref_search_kf_sim3(kf, Scw, th, pts, skip, held):
    """ORBmatcher::SearchByProjection(pKF, Scw, vpPoints, vpMatched, th) (ORBmatcher.cc:292-405). Returns (n, feat_match[N])."""
    L = matcher_lib(); s, keep = _refkf(kf); X, nr, mn, mx, qd = _pts(pts)
    S = _c(Scw, np.float32).reshape(16); sk = _c(skip, np.uint8); hd = _c(held, np.uint8); fm = np.full(s.N, -1, np.int32)
    L.ref_orbm_search_kf_sim3.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 8
    n = L.ref_orbm_search_kf_sim3(ctypes.byref(s), S.ctypes.data, int(th), len(X), sk.ctypes.data, X.ctypes.data, nr.ctypes.data, mn.ctypes.data,
                                  mx.ctypes.data, qd.ctypes.data, hd.ctypes.data, fm.ctypes.data)
    return n, fm


def ref_fuse_kf(kf, th, pts, skip, occupied):
    """ORBmatcher::Fuse(pKF, vpMapPoints, th) (ORBmatcher.cc:827-977).  kf['Tcw'] is the keyframe pose.  Returns (nFused, slot[M])."""
    L = matcher_lib(); s, keep = _refkf(kf); X, nr, mn, mx, qd = _pts(pts)
    sk = _c(skip, np.uint8); oc = _c(occupied, np.uint8); slot = np.full(len(X), -1, np.int32)
    L.ref_orbm_fuse_kf.argtypes = [ctypes.c_void_p, ctypes.c_float, ctypes.c_int] + [ctypes.c_void_p] * 8
    n = L.ref_orbm_fuse_kf(ctypes.byref(s), float(th), len(X), sk.ctypes.data, X.ctypes.data, nr.ctypes.data, mn.ctypes.data, mx.ctypes.data,
                           qd.ctypes.data, oc.ctypes.data, slot.ctypes.data)
    return n, slot


def ref_fuse_sim3(kf, Scw, th, pts, skip, occupied):
    """ORBmatcher::Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) (ORBmatcher.cc:979-1102).  Returns (nFused, slot[M])."""
    L = matcher_lib(); s, keep = _refkf(kf); X, nr, mn, mx, qd = _pts(pts)
    S = _c(Scw, np.float32).reshape(16); sk = _c(skip, np.uint8); oc = _c(occupied, np.uint8); slot = np.full(len(X), -1, np.int32)
    L.ref_orbm_fuse_sim3.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_int] + [ctypes.c_void_p] * 8
    n = L.ref_orbm_fuse_sim3(ctypes.byref(s), S.ctypes.data, float(th), len(X), sk.ctypes.data, X.ctypes.data, nr.ctypes.data, mn.ctypes.data,
                             mx.ctypes.data, qd.ctypes.data, oc.ctypes.data, slot.ctypes.data)
    return n, slot


def ref_search_by_sim3(kf1, kf2, s12, R12, t12, th, has1, pts1, has2, pts2, matches12):
    """ORBmatcher::SearchBySim3 (ORBmatcher.cc:1104-1328).  pts1 / pts2 are indexed like the keyframes' features.  Returns (nFound, matches12)."""
    L = matcher_lib(); a, k1 = _refkf(kf1); b, k2 = _refkf(kf2)
    X1, _, mn1, mx1, d1 = _pts(pts1); X2, _, mn2, mx2, d2 = _pts(pts2)
    R = _c(R12, np.float32).reshape(9); t = _c(t12, np.float32).reshape(3); h1 = _c(has1, np.uint8); h2 = _c(has2, np.uint8)
    m = _c(matches12, np.int32).copy()
    L.ref_orbm_search_by_sim3.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float] + [ctypes.c_void_p] * 11
    n = L.ref_orbm_search_by_sim3(ctypes.byref(a), ctypes.byref(b), float(s12), R.ctypes.data, t.ctypes.data, float(th), h1.ctypes.data, X1.ctypes.data,
                                  mn1.ctypes.data, mx1.ctypes.data, d1.ctypes.data, h2.ctypes.data, X2.ctypes.data, mn2.ctypes.data, mx2.ctypes.data,
                                  d2.ctypes.data, m.ctypes.data)
    return n, m


def ref_search_frame_kf(K4, bounds4, Tcw, scale_factors, cur, held, has, skip, pts, kf_angle, th, orb_dist, check_ori=True):
    """ORBmatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) (ORBmatcher.cc:1474-1601).  Returns (n, feat_match[N])."""
    L = matcher_lib()
    L.ref_orbm_set_camera(*[float(v) for v in K4], float(bounds4[0]), float(bounds4[1]), float(bounds4[2]), float(bounds4[3]))
    T = _c(Tcw, np.float32).reshape(16); sf = _c(scale_factors, np.float32)
    fxy = _c(np.stack([cur["x"], cur["y"]], 1), np.float32); N = len(fxy)
    fo = _c(cur["octave"], np.int32); fa = _c(cur["angle"], np.float32); fd = _c(cur["desc"], np.uint8); hd = _c(held, np.uint8)
    X, _, mn, mx, qd = _pts(pts); hs = _c(has, np.uint8); sk = _c(skip, np.uint8); ka = _c(kf_angle, np.float32)
    fm = np.full(N, -1, np.int32)
    L.ref_orbm_search_frame_kf.argtypes = [ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + \
        [ctypes.c_void_p] * 5 + [ctypes.c_int] + [ctypes.c_void_p] * 8
    n = L.ref_orbm_search_frame_kf(int(bool(check_ori)), float(th), int(orb_dist), T.ctypes.data, sf.ctypes.data, len(sf), N, fxy.ctypes.data, fo.ctypes.data,
                                   fa.ctypes.data, fd.ctypes.data, hd.ctypes.data, len(X), hs.ctypes.data, sk.ctypes.data, X.ctypes.data, mn.ctypes.data,
                                   mx.ctypes.data, ka.ctypes.data, qd.ctypes.data, fm.ctypes.data)
    return n, fm


# ---- vocabulary-bucket matchers and SearchForInitialization (reference object code) ----------------------------------------------------
def _fvargs(fv):
    k = [_c(fv["nodes"], np.int32), _c(fv["start"], np.int32), _c(fv["items"], np.int32)]
    return k, [len(k[0]), k[0].ctypes.data, k[1].ctypes.data, k[2].ctypes.data]


def ref_search_by_bow_kf_frame(kf, has1, fv1, f_angle, f_desc, fv2, nnratio=0.7, check_ori=True):
    """ORBmatcher::SearchByBoW(pKF, F, vpMapPointMatches) (ORBmatcher.cc:159-290).  Returns (n, match[N_frame] = keyframe feature or -1)."""
    L = matcher_lib(); s, keep = _refkf(kf)
    h = _c(has1, np.uint8); k1, a1 = _fvargs(fv1); k2, a2 = _fvargs(fv2); fa = _c(f_angle, np.float32); fd = _c(f_desc, np.uint8)
    m = np.full(len(fa), -1, np.int32)
    vp, i = ctypes.c_void_p, ctypes.c_int
    L.ref_orbm_search_by_bow_kf_frame.argtypes = [ctypes.c_float, i, vp, vp, i, vp, vp, vp, i, vp, vp, i, vp, vp, vp, vp]
    n = L.ref_orbm_search_by_bow_kf_frame(float(nnratio), int(bool(check_ori)), ctypes.byref(s), h.ctypes.data, *a1, len(fa), fa.ctypes.data, fd.ctypes.data, *a2,
                                          m.ctypes.data)
    return n, m


def ref_search_by_bow_kf_kf(kf1, has1, fv1, kf2, has2, fv2, nnratio=0.75, check_ori=True):
    """ORBmatcher::SearchByBoW(pKF1, pKF2, vpMatches12) (ORBmatcher.cc:524-657).  Returns (n, match12[N1])."""
    L = matcher_lib(); a, ka = _refkf(kf1); b, kb = _refkf(kf2)
    h1 = _c(has1, np.uint8); h2 = _c(has2, np.uint8); k1, a1 = _fvargs(fv1); k2, a2 = _fvargs(fv2)
    m = np.full(a.N, -1, np.int32)
    vp, i = ctypes.c_void_p, ctypes.c_int
    L.ref_orbm_search_by_bow_kf_kf.argtypes = [ctypes.c_float, i, vp, vp, i, vp, vp, vp, vp, vp, i, vp, vp, vp, vp]
    n = L.ref_orbm_search_by_bow_kf_kf(float(nnratio), int(bool(check_ori)), ctypes.byref(a), h1.ctypes.data, *a1, ctypes.byref(b), h2.ctypes.data, *a2, m.ctypes.data)
    return n, m


def ref_search_for_triangulation(kf1, has1, fv1, kf2, has2, fv2, level_sigma2, F12, nnratio=0.6, check_ori=False):
    """ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, false) (ORBmatcher.cc:659-825).  Returns (n, match12[N1])."""
    L = matcher_lib(); a, ka = _refkf(kf1); b, kb = _refkf(kf2)
    h1 = _c(has1, np.uint8); h2 = _c(has2, np.uint8); k1, a1 = _fvargs(fv1); k2, a2 = _fvargs(fv2)
    ls = _c(level_sigma2, np.float32); F = _c(F12, np.float32).reshape(9)
    m = np.full(a.N, -1, np.int32)
    vp, i = ctypes.c_void_p, ctypes.c_int
    L.ref_orbm_search_for_triangulation.argtypes = [ctypes.c_float, i, vp, vp, i, vp, vp, vp, vp, vp, i, vp, vp, vp, vp, vp, vp]
    n = L.ref_orbm_search_for_triangulation(float(nnratio), int(bool(check_ori)), ctypes.byref(a), h1.ctypes.data, *a1, ctypes.byref(b), h2.ctypes.data, *a2,
                                            ls.ctypes.data, F.ctypes.data, m.ctypes.data)
    return n, m


def ref_search_for_initialization(K4, bounds4, scale_factors, f1, f2, prev_matched, window=100, nnratio=0.9, check_ori=True):
    """ORBmatcher::SearchForInitialization (ORBmatcher.cc:407-522).  Returns (n, matches12, prev_matched updated)."""
    L = matcher_lib()
    L.ref_orbm_set_camera(*[float(v) for v in K4], float(bounds4[0]), float(bounds4[1]), float(bounds4[2]), float(bounds4[3]))
    sf = _c(scale_factors, np.float32)
    arr = []
    for f in (f1, f2):
        arr += [_c(np.stack([f["x"], f["y"]], 1), np.float32), _c(f["octave"], np.int32), _c(f["angle"], np.float32), _c(f["desc"], np.uint8)]
    pm = _c(prev_matched, np.float32).copy(); m = np.full(len(arr[0]), -1, np.int32)
    vp, i = ctypes.c_void_p, ctypes.c_int
    L.ref_orbm_search_for_initialization.argtypes = [ctypes.c_float, i, i, vp, i, i, vp, vp, vp, vp, i, vp, vp, vp, vp, vp, vp]
    n = L.ref_orbm_search_for_initialization(float(nnratio), int(bool(check_ori)), int(window), sf.ctypes.data, len(sf), len(arr[0]), arr[0].ctypes.data,
                                             arr[1].ctypes.data, arr[2].ctypes.data, arr[3].ctypes.data, len(arr[4]), arr[4].ctypes.data, arr[5].ctypes.data,
                                             arr[6].ctypes.data, arr[7].ctypes.data, pm.ctypes.data, m.ctypes.data)
    return n, m, pm


# ---------------------------------------------------------------------------------------------------------------------
# the reference's own DBoW2 (oracle/_ref/libref_dbow2.so: TemplatedVocabulary / FORB / BowVector / FeatureVector / ScoringObject unmodified)
_SOD = os.path.join(_HERE, "_ref", "libref_dbow2.so")
_LIBD = None


def dbow2_available():
    return os.path.exists(_SOD)


class RefVocabulary:
    """ORBVocabulary of the reference: loadFromTextFile + transform(features, BowVector, FeatureVector, levelsup), reference object code."""

    def __init__(self, path):
        global _LIBD
        if _LIBD is None:
            _LIBD = ctypes.CDLL(_SOD)
            _LIBD.ref_dbow2_load_text.restype = ctypes.c_void_p; _LIBD.ref_dbow2_load_text.argtypes = [ctypes.c_char_p]
            _LIBD.ref_dbow2_destroy.argtypes = [ctypes.c_void_p]; _LIBD.ref_dbow2_size.argtypes = [ctypes.c_void_p]
            _LIBD.ref_dbow2_transform.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 6
        self.L = _LIBD
        self.h = self.L.ref_dbow2_load_text(path.encode())
        if not self.h:
            raise RuntimeError("reference loadFromTextFile failed: " + path)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_dbow2_destroy(self.h); self.h = None

    def size(self):
        return self.L.ref_dbow2_size(self.h)

    def transform(self, desc, levelsup=4):
        d = _c(desc, np.uint8).reshape(-1, 32); N = len(d); M = max(N, 1)
        bi = np.zeros(M, np.int32); bv = np.zeros(M, np.float64); fn = np.zeros(M, np.int32); fs = np.zeros(M + 1, np.int32); fi = np.zeros(M, np.int32)
        fc = ctypes.c_int(0)
        nb = self.L.ref_dbow2_transform(self.h, N, d.ctypes.data, int(levelsup), bi.ctypes.data, bv.ctypes.data, fn.ctypes.data, fs.ctypes.data, fi.ctypes.data,
                                        ctypes.addressof(fc))
        nf = fc.value
        return dict(bow_ids=bi[:nb].copy(), bow_vals=bv[:nb].copy(), fv=dict(nodes=fn[:nf].copy(), start=fs[:nf + 1].copy(), items=fi[:fs[nf]].copy()))


# ---------------------------------------------------------------------------------------------------------------------
# the reference's own Sim3Solver.cc (oracle/_ref/libref_sim3solver.so): constructor bookkeeping, FromCameraToImage, Project, CheckInliers
_SOS = os.path.join(_HERE, "_ref", "libref_sim3solver.so")


def sim3solver_available():
    return os.path.exists(_SOS)


def ref_sim3_check_inliers(X3Dc1, X3Dc2, oct1, oct2, level_sigma2, K1, K2, T12, T21):
    """Returns (inliers u8[n_hyp, N], n_inliers, max_err1, max_err2, P1im1, P2im2) from the reference's object code."""
    L = ctypes.CDLL(_SOS)
    X1 = _c(X3Dc1, np.float32); X2 = _c(X3Dc2, np.float32); N = len(X1)
    o1 = _c(oct1, np.int32); o2 = _c(oct2, np.int32); ls = _c(level_sigma2, np.float32); k1 = _c(K1, np.float32); k2 = _c(K2, np.float32)
    a = _c(T12, np.float32).reshape(-1, 16); b = _c(T21, np.float32).reshape(-1, 16); nh = len(a)
    inl = np.zeros((nh, N), np.uint8); n = np.zeros(nh, np.int32); m1 = np.zeros(N, np.int32); m2 = np.zeros(N, np.int32)
    p1 = np.zeros((N, 2), np.float32); p2 = np.zeros((N, 2), np.float32)
    vp, i = ctypes.c_void_p, ctypes.c_int
    L.ref_sim3_check_inliers.argtypes = [i, vp, vp, vp, vp, vp, i, vp, vp, i, vp, vp, vp, vp, vp, vp, vp, vp]
    L.ref_sim3_check_inliers(N, X1.ctypes.data, X2.ctypes.data, o1.ctypes.data, o2.ctypes.data, ls.ctypes.data, len(ls), k1.ctypes.data, k2.ctypes.data, nh,
                             a.ctypes.data, b.ctypes.data, inl.ctypes.data, n.ctypes.data, m1.ctypes.data, m2.ctypes.data, p1.ctypes.data, p2.ctypes.data)
    return inl, n, m1, m2, p1, p2


# ------------------------------------------------------------------------------------------------
# the reference's own Optimizer.cc + Converter.cc + vendored g2o (oracle/_ref/libref_optimizer.so, same Makefile: compiled unmodified against the Eigen
# stand-in oracle/eigenshim, the data-model stand-in oracle/optshim and oracle/slamshim's cv::Mat); C API in oracle/ref_optimizer_capi.cc
_OPT_SO = os.path.join(_HERE, "_ref", "libref_optimizer.so")
_OPT_LIB = None


def optimizer_available():
    return os.path.exists(_OPT_SO)


def optimizer_lib():
    global _OPT_LIB
    if _OPT_LIB is None:
        _OPT_LIB = ctypes.CDLL(_OPT_SO)
    return _OPT_LIB


def _vp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _ba_args(poses, fixed, intr, points, e_kf, e_pt, e_uv, e_w):
    poses = np.ascontiguousarray(poses, np.float32).reshape(-1, 16).copy(); K = len(poses)
    points = np.ascontiguousarray(points, np.float32).reshape(-1, 3).copy()
    intr = np.ascontiguousarray(intr, np.float64)
    if intr.ndim == 1:
        intr = np.tile(intr, (K, 1))
    return (poses, np.ascontiguousarray(fixed, np.uint8), np.ascontiguousarray(intr), points, np.ascontiguousarray(e_kf, np.int32), np.ascontiguousarray(e_pt, np.int32),
            np.ascontiguousarray(e_uv, np.float32), np.ascontiguousarray(e_w, np.float32))


def ref_local_ba(poses, fixed, intr, points, e_kf, e_pt, e_uv, e_inv_sigma2):
    """Optimizer::LocalBundleAdjustment of the reference on the stand-in map of a flat graph (current keyframe = last free one).  Returns dict(poses [K,4,4],
    points [P,3], nobs [P] = observations left per point after the reference erased its outliers)."""
    p, fx, it, pt, kf, ept, uv, w = _ba_args(poses, fixed, intr, points, e_kf, e_pt, e_uv, e_inv_sigma2)
    nobs = np.zeros(len(pt), np.int32)
    optimizer_lib().ref_opt_local_ba(len(p), _vp(p), _vp(fx), _vp(it), len(pt), _vp(pt), len(kf), _vp(kf), _vp(ept), _vp(uv), _vp(w), _vp(nobs))
    return dict(poses=p.reshape(-1, 4, 4), points=pt, nobs=nobs)


def ref_bundle_adjust(poses, fixed, intr, points, e_kf, e_pt, e_uv, e_inv_sigma2, n_iterations=20, robust=False):
    """Optimizer::BundleAdjustment of the reference (nLoopKF = 0).  Only the fixed == 1 (mnId == 0) keyframe is held."""
    p, fx, it, pt, kf, ept, uv, w = _ba_args(poses, fixed, intr, points, e_kf, e_pt, e_uv, e_inv_sigma2)
    optimizer_lib().ref_opt_bundle_adjust(len(p), _vp(p), _vp(fx), _vp(it), len(pt), _vp(pt), len(kf), _vp(kf), _vp(ept), _vp(uv), _vp(w), int(n_iterations), int(bool(robust)))
    return dict(poses=p.reshape(-1, 4, 4), points=pt)


def ref_pose_optimization(Tcw, Xw, obs, inv_sigma2, K4):
    """Optimizer::PoseOptimization of the reference: (Tcw [4,4], outlier u8[N], n_inliers)."""
    T = np.ascontiguousarray(Tcw, np.float32).reshape(16).copy()
    Xw = np.ascontiguousarray(Xw, np.float32); obs = np.ascontiguousarray(obs, np.float32); w = np.ascontiguousarray(inv_sigma2, np.float32)
    k4 = np.ascontiguousarray(K4, np.float32); out = np.zeros(len(w), np.uint8)
    n = optimizer_lib().ref_opt_pose_optimization(_vp(T), _vp(k4), len(w), _vp(Xw), _vp(obs), _vp(w), _vp(out))
    return T.reshape(4, 4), out, int(n)


def ref_optimize_sim3(sim3, valid, P1c, P2c, obs1, obs2, w1, w2, K1, K2, th2=10.0, fix_scale=False):
    """Optimizer::OptimizeSim3 of the reference; the keyframes sit at the identity, so the camera-frame points the reference computes (R*X + t in float) are
    exactly the given ones.  Returns dict(sim3 [8], inlier u8[N], n_in)."""
    f = lambda x: np.ascontiguousarray(x, np.float32)
    I4 = np.eye(4, dtype=np.float32).reshape(16)
    s = np.ascontiguousarray(sim3, np.float64).copy(); v = np.ascontiguousarray(valid, np.uint8); inl = np.zeros(len(v), np.uint8)
    a = [f(K1), f(K2), f(P1c), f(P2c), f(obs1), f(obs2), f(w1), f(w2)]
    n = optimizer_lib().ref_opt_optimize_sim3(_vp(I4), _vp(I4), _vp(a[0]), _vp(a[1]), len(v), _vp(v), _vp(a[2]), _vp(a[3]), _vp(a[4]), _vp(a[5]), _vp(a[6]), _vp(a[7]), _vp(s),
                                              ctypes.c_float(th2), int(bool(fix_scale)), _vp(inl))
    return dict(sim3=s, inlier=inl, n_in=int(n))


def ref_pose_graph(sim3, fixed, e_i, e_j, e_meas, fix_scale=False, iterations=20):
    """The g2o graph of Optimizer::OptimizeEssentialGraph (BlockSolver_7_3 / LinearSolverEigen / Levenberg, lambda init 1e-16) built by the reference's object code."""
    s = np.ascontiguousarray(sim3, np.float64).reshape(-1, 8).copy()
    fx = np.ascontiguousarray(fixed, np.uint8); ei = np.ascontiguousarray(e_i, np.int32); ej = np.ascontiguousarray(e_j, np.int32); em = np.ascontiguousarray(e_meas, np.float64)
    its = optimizer_lib().ref_g2o_pose_graph(len(s), _vp(s), _vp(fx), len(ei), _vp(ei), _vp(ej), _vp(em), int(bool(fix_scale)), int(iterations))
    return dict(sim3=s, lm_iterations=int(its))
