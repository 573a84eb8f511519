"""orbslamm_b200 -- B200-native (sm_100a) implementation of the ORBSLAMM hot path.

Python here is only a thin ctypes binding over the C-ABI in include/orbslamm_b200.h
(the product is liborbslamm_b200.so: hand-written CUDA + a C++ host driver).  There is
no CPU fallback: if the shared library is missing or no CUDA device is usable, the calls
raise.  The CPU checker package (tests only) is never imported from here.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liborbslamm_b200.so")
_lib = None

ORBS_OK = 0


class OrbsError(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"orbslamm_b200 error {code}: {text}")
        self.code = code


def load():
    """Load the in-tree CUDA library.  Raises if it has not been built (python -m orbslamm_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OSError(f"{LIB_PATH} not built; run `python orbslamm_b200/build.py` (there is no CPU fallback)")
        _lib = ctypes.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


def _check(rc):
    if rc != ORBS_OK:
        raise OrbsError(rc, load().orbs_last_error().decode())


class DeviceView(ctypes.Structure):
    _fields_ = [("n_frames", ctypes.c_int), ("slab", ctypes.c_int),
                ("kp_xy", ctypes.c_void_p), ("kp_angle", ctypes.c_void_p), ("kp_response", ctypes.c_void_p),
                ("kp_octave", ctypes.c_void_p), ("kp_size", ctypes.c_void_p), ("desc", ctypes.c_void_p),
                ("counts", ctypes.c_void_p), ("level_counts", ctypes.c_void_p)]


def _declare(L):
    c = ctypes
    L.orbs_last_error.restype = c.c_char_p
    L.orbs_version.restype = c.c_int
    L.orbs_device_count.restype = c.c_int
    vp, i, f, sz = c.c_void_p, c.c_int, c.c_float, c.c_size_t
    L.orbx_create.argtypes = [c.POINTER(vp), i, f, i, i, i, i]
    L.orbx_destroy.argtypes = [vp]
    L.orbx_get_tables.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
    L.orbx_max_keypoints.argtypes = [vp, i, i, c.POINTER(i)]
    L.orbx_extract.argtypes = [vp, vp, i, i, i, i, sz, vp, vp, vp, vp, vp, vp, i, vp]
    L.orbx_extract_device.argtypes = [vp, vp, i, i, i, i, sz]
    L.orbx_device_results.argtypes = [vp, c.POINTER(DeviceView)]
    L.orbx_download.argtypes = [vp, vp, vp, vp, vp, vp, vp, i, vp]
    L.orbx_level_size.argtypes = [vp, i, i, i, c.POINTER(i), c.POINTER(i)]
    L.orbx_get_pyramid_level.argtypes = [vp, i, i, i, vp, i]
    L.orbx_get_candidates.argtypes = [vp, i, i, vp, i, c.POINTER(i)]
    L.orbx_get_candidates.restype = c.c_int
    L.orbx_stream.argtypes = [vp]
    L.orbx_stream.restype = vp
    L.orbx_synchronize.argtypes = [vp]
    L.orbx_kernel_launches.argtypes = [vp]
    L.orbx_kernel_launches.restype = c.c_longlong
    for name in ("orbx_create", "orbx_destroy", "orbx_get_tables", "orbx_max_keypoints", "orbx_extract",
                 "orbx_extract_device", "orbx_device_results", "orbx_download", "orbx_level_size",
                 "orbx_get_pyramid_level", "orbx_synchronize"):
        getattr(L, name).restype = c.c_int
    from . import _bind_more
    _bind_more.declare(L)


def _ptr(a):
    return None if a is None else a.ctypes.data


class ORBextractor:
    """Mirror of iORB_SLAM::ORBextractor (reference S/include/ORBextractor.h:45-111).

    __call__(image) plays operator()(image, mask, keypoints, descriptors): it returns a dict of
    SoA keypoint arrays (cv::KeyPoint fields) and the N x 32 descriptor matrix.  `mask` is ignored
    like in the reference.  extract_batch() is the batched form for many frames of one shape."""

    HARRIS_SCORE = 0
    FAST_SCORE = 1

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, device=0):
        self._L = load()
        self._h = ctypes.c_void_p()
        _check(self._L.orbx_create(ctypes.byref(self._h), int(nfeatures), float(scaleFactor), int(nlevels),
                                   int(iniThFAST), int(minThFAST), int(device)))
        self.nlevels = int(nlevels)
        n = ctypes.c_int()
        sf = ctypes.c_float()
        self._scale = np.zeros(nlevels, np.float32); self._inv_scale = np.zeros(nlevels, np.float32)
        self._sigma2 = np.zeros(nlevels, np.float32); self._inv_sigma2 = np.zeros(nlevels, np.float32)
        self._nfeat = np.zeros(nlevels, np.int32)
        _check(self._L.orbx_get_tables(self._h, ctypes.addressof(n), ctypes.addressof(sf), _ptr(self._scale),
                                       _ptr(self._inv_scale), _ptr(self._sigma2), _ptr(self._inv_sigma2),
                                       _ptr(self._nfeat)))
        self._scale_factor = sf.value

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.orbx_destroy(h)

    # getters of the reference class
    def GetLevels(self): return self.nlevels
    def GetScaleFactor(self): return self._scale_factor
    def GetScaleFactors(self): return self._scale.copy()
    def GetInverseScaleFactors(self): return self._inv_scale.copy()
    def GetScaleSigmaSquares(self): return self._sigma2.copy()
    def GetInverseScaleSigmaSquares(self): return self._inv_sigma2.copy()
    def features_per_level(self): return self._nfeat.copy()

    @property
    def handle(self): return self._h

    def max_keypoints(self, width, height):
        n = ctypes.c_int()
        _check(self._L.orbx_max_keypoints(self._h, int(width), int(height), ctypes.byref(n)))
        return n.value

    def extract_batch(self, images):
        """images: u8 array [B, H, W] (host).  Returns list of per-frame dicts."""
        images = np.ascontiguousarray(images, np.uint8)
        assert images.ndim == 3
        B, H, W = images.shape
        if H == 0 or W == 0:
            return [_empty() for _ in range(B)]
        cap = self.max_keypoints(W, H)
        xy = np.empty((B, cap, 2), np.float32); ang = np.empty((B, cap), np.float32)
        resp = np.empty((B, cap), np.float32); octv = np.empty((B, cap), np.int32)
        size = np.empty((B, cap), np.float32); desc = np.empty((B, cap, 32), np.uint8)
        counts = np.zeros(B, np.int32)
        _check(self._L.orbx_extract(self._h, _ptr(images), B, W, H, images.strides[1], images.strides[0], _ptr(xy), _ptr(ang),
                                    _ptr(resp), _ptr(octv), _ptr(size), _ptr(desc), cap, _ptr(counts)))
        out = []
        for f in range(B):
            n = int(counts[f])
            out.append(dict(x=xy[f, :n, 0].copy(), y=xy[f, :n, 1].copy(), angle=ang[f, :n].copy(),
                            response=resp[f, :n].copy(), octave=octv[f, :n].copy(), size=size[f, :n].copy(),
                            desc=desc[f, :n].copy()))
        return out

    def __call__(self, image, mask=None):
        image = np.asarray(image)
        if image.size == 0:
            return _empty()
        assert image.dtype == np.uint8 and image.ndim == 2, "CV_8UC1 image expected (ORBextractor.cc:1050)"
        return self.extract_batch(image[None])[0]

    def extract_device(self, dev_ptr, n_frames, width, height, stride, frame_stride):
        _check(self._L.orbx_extract_device(self._h, ctypes.c_void_p(dev_ptr), n_frames, width, height, stride, frame_stride))

    def device_view(self):
        v = DeviceView()
        _check(self._L.orbx_device_results(self._h, ctypes.byref(v)))
        return v

    def download(self, n_frames, cap):
        xy = np.empty((n_frames, cap, 2), np.float32); ang = np.empty((n_frames, cap), np.float32)
        resp = np.empty((n_frames, cap), np.float32); octv = np.empty((n_frames, cap), np.int32)
        size = np.empty((n_frames, cap), np.float32); desc = np.empty((n_frames, cap, 32), np.uint8)
        counts = np.zeros(n_frames, np.int32)
        _check(self._L.orbx_download(self._h, _ptr(xy), _ptr(ang), _ptr(resp), _ptr(octv), _ptr(size), _ptr(desc), cap, _ptr(counts)))
        return dict(xy=xy, angle=ang, response=resp, octave=octv, size=size, desc=desc, counts=counts)

    def pyramid_level(self, frame, level, width, height, border=False):
        """mvImagePyramid[level] of a frame of the last call (optionally with the 19-px reflect-101 border)."""
        lw, lh = ctypes.c_int(), ctypes.c_int()
        _check(self._L.orbx_level_size(self._h, width, height, level, ctypes.byref(lw), ctypes.byref(lh)))
        b = 19 if border else 0
        out = np.empty((lh.value + 2 * b, lw.value + 2 * b), np.uint8)
        _check(self._L.orbx_get_pyramid_level(self._h, frame, level, int(bool(border)), _ptr(out), out.strides[0]))
        return out

    def candidates(self, frame, level):
        """FAST candidates (x, y, response) of one level, relative to minBorder, unordered (stage inspection)."""
        n = ctypes.c_int()
        _check(self._L.orbx_get_candidates(self._h, frame, level, None, 0, ctypes.byref(n)))
        out = np.empty((max(n.value, 1), 3), np.int32)
        _check(self._L.orbx_get_candidates(self._h, frame, level, _ptr(out), len(out), ctypes.byref(n)))
        return out[:n.value]

    KERNELS = ("resize_level", "fast_cells", "blur7", "octree", "orient_describe")

    def set_profiling(self, on):
        self._L.orbx_set_profiling.argtypes = [ctypes.c_void_p, ctypes.c_int]
        _check(self._L.orbx_set_profiling(self._h, int(bool(on))))

    def kernel_times(self):
        """{kernel: (total_ms, launches)} since set_profiling(True)."""
        ms = np.zeros(len(self.KERNELS)); cnt = np.zeros(len(self.KERNELS), np.int64)
        self._L.orbx_get_kernel_times.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        _check(self._L.orbx_get_kernel_times(self._h, _ptr(ms), _ptr(cnt), len(ms)))
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(self.KERNELS)}

    def stream(self): return self._L.orbx_stream(self._h)

    def set_stream(self, cuda_stream):
        self._L.orbx_set_stream.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        _check(self._L.orbx_set_stream(self._h, ctypes.c_void_p(cuda_stream)))
    def synchronize(self): _check(self._L.orbx_synchronize(self._h))
    def kernel_launches(self): return int(self._L.orbx_kernel_launches(self._h))


def _empty():
    z = np.zeros(0, np.float32)
    return dict(x=z, y=z, angle=z, response=z, octave=np.zeros(0, np.int32), size=z, desc=np.zeros((0, 32), np.uint8))


def _dp(a):
    """host ndarray -> pointer, int -> device pointer, None -> NULL"""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return ctypes.c_void_p(int(a))
    return a.ctypes.data


class ORBmatcher:
    """Mirror of iORB_SLAM::ORBmatcher (reference S/include/ORBmatcher.h:37-102) over flat arrays.

    nnratio / checkOri as in the reference constructor (ORBmatcher.cc:41-43).  Frames are described by SoA
    arrays instead of Frame objects; see include/orbslamm_b200.h for the layouts."""

    TH_LOW, TH_HIGH, HISTO_LENGTH = 50, 100, 30

    def __init__(self, nnratio=0.6, checkOri=True, device=0):
        self._L = load()
        self._h = ctypes.c_void_p()
        _check(self._L.orbm_create(ctypes.byref(self._h), int(device)))
        self.mfNNratio = float(nnratio)
        self.mbCheckOrientation = bool(checkOri)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.orbm_destroy(h)

    @property
    def handle(self): return self._h
    def stream(self): return self._L.orbm_stream(self._h)

    def set_stream(self, cuda_stream):
        self._L.orbm_set_stream.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        _check(self._L.orbm_set_stream(self._h, ctypes.c_void_p(cuda_stream)))
    def synchronize(self): _check(self._L.orbm_synchronize(self._h))
    def kernel_launches(self): return int(self._L.orbm_kernel_launches(self._h))

    def DescriptorDistance(self, a, b):
        """All-pairs Hamming distances of descriptor rows a [n,32] and b [m,32] (ORBmatcher.cc:1649-1665)."""
        a = np.ascontiguousarray(a, np.uint8).reshape(-1, 32); b = np.ascontiguousarray(b, np.uint8).reshape(-1, 32)
        out = np.zeros((len(a), len(b)), np.int32)
        _check(self._L.orbm_descriptor_distance(self._h, _ptr(a), len(a), _ptr(b), len(b), _ptr(out), 0))
        return out

    def ComputeDistinctiveDescriptors(self, desc_lists):
        """MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:242-307) for a batch of map points: desc_lists[p] = u8[N_p, 32], the descriptors of the point's
        observations in map order.  Returns best_idx i32[n_points] (-1 for an empty list)."""
        n = len(desc_lists)
        start = np.zeros(n + 1, np.int32)
        for p, d in enumerate(desc_lists): start[p + 1] = start[p] + len(d)
        flat = np.zeros((max(int(start[n]), 1), 32), np.uint8)
        for p, d in enumerate(desc_lists):
            if len(d): flat[start[p]:start[p + 1]] = np.asarray(d, np.uint8).reshape(-1, 32)
        out = np.zeros(n, np.int32)
        self._L.orbm_distinctive_descriptors.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        self._L.orbm_distinctive_descriptors.restype = ctypes.c_int
        _check(self._L.orbm_distinctive_descriptors(self._h, n, _ptr(flat), _ptr(start), _ptr(out), 0))
        return out

    def project_last_frame(self, Tcw, K4, bounds4, scale_factors, Xw, last_octave, q_counts, th, q_valid):
        """Projection block of SearchByProjection(Cur, Last) for n_frames frames (host arrays, slab layout)."""
        Tcw = np.ascontiguousarray(Tcw, np.float32).reshape(-1, 16); nfr = len(Tcw)
        K4 = np.ascontiguousarray(K4, np.float32); b4 = np.ascontiguousarray(bounds4, np.float32)
        sf = np.ascontiguousarray(scale_factors, np.float32)
        Xw = np.ascontiguousarray(Xw, np.float32).reshape(nfr, -1, 3); slab = Xw.shape[1]
        lo = np.ascontiguousarray(last_octave, np.int32).reshape(nfr, slab)
        qc = np.ascontiguousarray(q_counts, np.int32)
        qv = np.ascontiguousarray(q_valid, np.uint8).reshape(nfr, slab).copy()
        uv = np.zeros((nfr, slab, 2), np.float32); rad = np.zeros((nfr, slab), np.float32)
        mn = np.zeros((nfr, slab), np.int32); mx = np.zeros((nfr, slab), np.int32)
        _check(self._L.orbm_project_last_frame(self._h, nfr, _ptr(Tcw), _ptr(K4), _ptr(b4), _ptr(sf), len(sf), _ptr(Xw), _ptr(lo),
                                               _ptr(qc), slab, float(th), _ptr(qv), _ptr(uv), _ptr(rad), _ptr(mn), _ptr(mx), 0))
        return qv, uv, rad, mn, mx

    def UndistortKeyPoints(self, kp_xy, counts, K4, dist_coef):
        """Frame::UndistortKeyPoints (Frame.cc:404-434) for n_frames frames: kp_xy [n_frames, slab, 2] -> undistorted."""
        xy = np.ascontiguousarray(kp_xy, np.float32); nfr, slab = xy.shape[0], xy.shape[1]
        cnt = np.ascontiguousarray(counts, np.int32); K4 = np.ascontiguousarray(K4, np.float32)
        d = np.zeros(5, np.float32); d[:len(dist_coef)] = np.asarray(dist_coef, np.float32).ravel()
        out = np.zeros_like(xy)
        _check(self._L.orbm_undistort_keypoints(self._h, nfr, _ptr(xy), _ptr(cnt), slab, _ptr(K4), _ptr(d), _ptr(out), 0))
        return out

    def isInFrustum(self, Tcw, Ow, K4, bounds4, log_scale_factor, Xw, normal, mf_min_distance, mf_max_distance, counts, viewing_cos_limit=0.5):
        """Frame::isInFrustum (Frame.cc:269-325) for counts[f] map points of n_frames frames (slab layout).
        Returns (in_view u8, proj_xy f32[.,2], pred_level i32, view_cos f32)."""
        T = np.ascontiguousarray(Tcw, np.float32).reshape(-1, 16); nfr = len(T)
        Ow = np.ascontiguousarray(Ow, np.float32).reshape(nfr, 3)
        K4 = np.ascontiguousarray(K4, np.float32); b4 = np.ascontiguousarray(bounds4, np.float32)
        Xw = np.ascontiguousarray(Xw, np.float32).reshape(nfr, -1, 3); slab = Xw.shape[1]
        nrm = np.ascontiguousarray(normal, np.float32).reshape(nfr, slab, 3)
        mn = np.ascontiguousarray(mf_min_distance, np.float32).reshape(nfr, slab); mx = np.ascontiguousarray(mf_max_distance, np.float32).reshape(nfr, slab)
        cnt = np.ascontiguousarray(counts, np.int32)
        iv = np.zeros((nfr, slab), np.uint8); uv = np.zeros((nfr, slab, 2), np.float32)
        lv = np.zeros((nfr, slab), np.int32); vc = np.zeros((nfr, slab), np.float32)
        _check(self._L.orbm_is_in_frustum(self._h, nfr, _ptr(T), _ptr(Ow), _ptr(K4), _ptr(b4), float(log_scale_factor), float(viewing_cos_limit),
                                          _ptr(Xw), _ptr(nrm), _ptr(mn), _ptr(mx), _ptr(cnt), slab, _ptr(iv), _ptr(uv), _ptr(lv), _ptr(vc), 0))
        return iv, uv, lv, vc

    def AssignFeaturesToGrid(self, bounds4, f_xy, f_counts):
        """Frame::AssignFeaturesToGrid (Frame.cc:230-245): returns (cell_start [n, 3073], cell_items [n, f_slab]) with cell = ix * 48 + iy."""
        xy = np.ascontiguousarray(f_xy, np.float32); n, fs = xy.shape[0], xy.shape[1]
        cnt = np.ascontiguousarray(f_counts, np.int32); b4 = np.ascontiguousarray(bounds4, np.float32)
        cs = np.zeros((n, 64 * 48 + 1), np.int32); ci = np.zeros((n, fs), np.int32)
        _check(self._L.orbm_assign_features_to_grid(self._h, n, _ptr(b4), _ptr(xy), _ptr(cnt), fs, _ptr(cs), _ptr(ci), 0))
        return cs, ci

    def GetFeaturesInArea(self, bounds4, f_xy, f_octave, f_counts, q_xyr, q_minl, q_maxl, q_counts, cap=256, win_origin2=None):
        """Frame / KeyFrame::GetFeaturesInArea for q_counts[f] queries per frame: returns (idx [n, q_slab, cap], count [n, q_slab])."""
        xy = np.ascontiguousarray(f_xy, np.float32); n, fs = xy.shape[0], xy.shape[1]
        oc = None if f_octave is None else np.ascontiguousarray(f_octave, np.int32)
        cnt = np.ascontiguousarray(f_counts, np.int32); b4 = np.ascontiguousarray(bounds4, np.float32)
        q = np.ascontiguousarray(q_xyr, np.float32); qs = q.shape[1]
        mn = None if q_minl is None else np.ascontiguousarray(q_minl, np.int32); mx = None if q_maxl is None else np.ascontiguousarray(q_maxl, np.int32)
        qc = np.ascontiguousarray(q_counts, np.int32); wo = None if win_origin2 is None else np.ascontiguousarray(win_origin2, np.float32)
        idx = np.zeros((n, qs, cap), np.int32); out_n = np.zeros((n, qs), np.int32)
        _check(self._L.orbm_get_features_in_area(self._h, n, _ptr(b4), _ptr(wo), _ptr(xy), _ptr(oc), _ptr(cnt), fs, _ptr(q), _ptr(mn), _ptr(mx), _ptr(qc), qs, int(cap),
                                                 _ptr(idx), _ptr(out_n), 0))
        return idx, out_n

    def SearchByProjection(self, bounds4, f_xy, f_octave, f_angle, f_desc, f_counts, q_valid, q_uv, q_radius, q_minl,
                           q_maxl, q_angle, q_desc, q_counts, th_dist=100, use_ratio=False, feat_match=None):
        """Generic projection search for n_frames frames (host arrays in slab layout [n_frames, slab, ...]).
        use_ratio=False: the (Cur, Last) variant (orientation check per mbCheckOrientation);
        use_ratio=True: the (Frame, MapPoints) variant with mfNNratio and no orientation check.
        Returns (nmatches[n_frames], feat_match[n_frames, f_slab])."""
        f_xy = np.ascontiguousarray(f_xy, np.float32); nfr, fs = f_xy.shape[0], f_xy.shape[1]
        f_octave = np.ascontiguousarray(f_octave, np.int32); f_angle = np.ascontiguousarray(f_angle, np.float32)
        f_desc = np.ascontiguousarray(f_desc, np.uint8); f_counts = np.ascontiguousarray(f_counts, np.int32)
        q_valid = np.ascontiguousarray(q_valid, np.uint8); qs = q_valid.shape[1]
        q_uv = np.ascontiguousarray(q_uv, np.float32); q_radius = np.ascontiguousarray(q_radius, np.float32)
        q_minl = np.ascontiguousarray(q_minl, np.int32); q_maxl = np.ascontiguousarray(q_maxl, np.int32)
        q_angle = np.ascontiguousarray(q_angle, np.float32); q_desc = np.ascontiguousarray(q_desc, np.uint8)
        q_counts = np.ascontiguousarray(q_counts, np.int32)
        fm = np.full((nfr, fs), -1, np.int32) if feat_match is None else np.ascontiguousarray(feat_match, np.int32).copy()
        nm = np.zeros(nfr, np.int32)
        b4 = np.ascontiguousarray(bounds4, np.float32)
        ratio = self.mfNNratio if use_ratio else 0.0
        ori = (not use_ratio) and self.mbCheckOrientation
        _check(self._L.orbm_search_by_projection(self._h, nfr, _ptr(b4), _ptr(f_xy), _ptr(f_octave), _ptr(f_angle), _ptr(f_desc),
                                                 _ptr(f_counts), fs, _ptr(q_valid), _ptr(q_uv), _ptr(q_radius), _ptr(q_minl),
                                                 _ptr(q_maxl), _ptr(q_angle), _ptr(q_desc), _ptr(q_counts), qs, int(th_dist),
                                                 float(ratio), int(ori), _ptr(fm), _ptr(nm), 0))
        return nm, fm


    # ---- KeyFrame / Sim3 projection family (ORBmatcher.cc:292-405, 827-977, 979-1102, 1104-1328, 1474-1601) -----------------------
    PROJ_TWO_STEP, PROJ_NO_DEPTH, PROJ_FRAME_BOUNDS, PROJ_FRAME_UV, PROJ_DIST_CAMERA, PROJ_CHECK_NORMAL, PROJ_LEVEL_PLUS1 = 1, 2, 4, 8, 16, 32, 64

    def project_points(self, views, scale_factors, Xw, normal, mf_min_distance, mf_max_distance, q_counts, q_valid):
        """orbm_project_points for len(views) target views (slab layout [n_views, slab, ...]); views: list of Projection.
        Returns (valid, uv, radius, minl, maxl, level)."""
        nv = len(views)
        arr = (Projection * nv)(*views)
        sf = np.ascontiguousarray(scale_factors, np.float32)
        Xw = np.ascontiguousarray(Xw, np.float32).reshape(nv, -1, 3); slab = Xw.shape[1]
        nrm = None if normal is None else np.ascontiguousarray(normal, np.float32).reshape(nv, slab, 3)
        mn = np.ascontiguousarray(mf_min_distance, np.float32).reshape(nv, slab); mx = np.ascontiguousarray(mf_max_distance, np.float32).reshape(nv, slab)
        qc = np.ascontiguousarray(q_counts, np.int32)
        qv = np.ascontiguousarray(q_valid, np.uint8).reshape(nv, slab).copy()
        uv = np.zeros((nv, slab, 2), np.float32); rad = np.zeros((nv, slab), np.float32)
        l0 = np.zeros((nv, slab), np.int32); l1 = np.zeros((nv, slab), np.int32); lv = np.zeros((nv, slab), np.int32)
        _check(self._L.orbm_project_points(self._h, nv, ctypes.cast(arr, ctypes.c_void_p), _ptr(sf), len(sf), _ptr(Xw), None if nrm is None else _ptr(nrm),
                                           _ptr(mn), _ptr(mx), _ptr(qc), slab, _ptr(qv), _ptr(uv), _ptr(rad), _ptr(l0), _ptr(l1), _ptr(lv), 0))
        return qv, uv, rad, l0, l1, lv

    def search_best_in_window(self, grid_bounds4, win_origin2, f_xy, f_octave, f_desc, f_counts, q_valid, q_uv, q_radius, q_minl, q_maxl, q_desc,
                              q_counts, th_dist, inv_level_sigma2=None, chi2_gate=5.99):
        """orbm_search_best_in_window (candidate loop of Fuse / SearchBySim3): returns (best_idx, best_dist) [n_frames, q_slab]."""
        f_xy = np.ascontiguousarray(f_xy, np.float32); nfr, fs = f_xy.shape[0], f_xy.shape[1]
        f_octave = np.ascontiguousarray(f_octave, np.int32); f_desc = np.ascontiguousarray(f_desc, np.uint8); f_counts = np.ascontiguousarray(f_counts, np.int32)
        q_valid = np.ascontiguousarray(q_valid, np.uint8); qs = q_valid.shape[1]
        q_uv = np.ascontiguousarray(q_uv, np.float32); q_radius = np.ascontiguousarray(q_radius, np.float32)
        q_minl = np.ascontiguousarray(q_minl, np.int32); q_maxl = np.ascontiguousarray(q_maxl, np.int32)
        q_desc = np.ascontiguousarray(q_desc, np.uint8); q_counts = np.ascontiguousarray(q_counts, np.int32)
        gb = np.ascontiguousarray(grid_bounds4, np.float32)
        wo = None if win_origin2 is None else np.ascontiguousarray(win_origin2, np.float32)
        inv = None if inv_level_sigma2 is None else np.ascontiguousarray(inv_level_sigma2, np.float32)
        bi = np.zeros((nfr, qs), np.int32); bd = np.zeros((nfr, qs), np.int32)
        _check(self._L.orbm_search_best_in_window(self._h, nfr, _ptr(gb), None if wo is None else _ptr(wo), _ptr(f_xy), _ptr(f_octave), _ptr(f_desc),
                                                  _ptr(f_counts), fs, _ptr(q_valid), _ptr(q_uv), _ptr(q_radius), _ptr(q_minl), _ptr(q_maxl), _ptr(q_desc),
                                                  _ptr(q_counts), qs, int(th_dist), None if inv is None else _ptr(inv), 0 if inv is None else len(inv),
                                                  float(chi2_gate), _ptr(bi), _ptr(bd), 0))
        return bi, bd

    def SearchByProjectionKF(self, grid_bounds4, win_origin2, f_xy, f_octave, f_angle, f_desc, f_counts, q_valid, q_uv, q_radius, q_minl,
                             q_maxl, q_angle, q_desc, q_counts, th_dist=50, ratio=0.0, check_ori=False, feat_match=None):
        """orbm_search_by_projection_kf: the claim loop against a KeyFrame target (integer window origin)."""
        f_xy = np.ascontiguousarray(f_xy, np.float32); nfr, fs = f_xy.shape[0], f_xy.shape[1]
        f_octave = np.ascontiguousarray(f_octave, np.int32); f_angle = np.ascontiguousarray(f_angle, np.float32)
        f_desc = np.ascontiguousarray(f_desc, np.uint8); f_counts = np.ascontiguousarray(f_counts, np.int32)
        q_valid = np.ascontiguousarray(q_valid, np.uint8); qs = q_valid.shape[1]
        q_uv = np.ascontiguousarray(q_uv, np.float32); q_radius = np.ascontiguousarray(q_radius, np.float32)
        q_minl = np.ascontiguousarray(q_minl, np.int32); q_maxl = np.ascontiguousarray(q_maxl, np.int32)
        q_angle = np.ascontiguousarray(q_angle, np.float32); q_desc = np.ascontiguousarray(q_desc, np.uint8)
        q_counts = np.ascontiguousarray(q_counts, np.int32)
        fm = np.full((nfr, fs), -1, np.int32) if feat_match is None else np.ascontiguousarray(feat_match, np.int32).copy()
        nm = np.zeros(nfr, np.int32)
        gb = np.ascontiguousarray(grid_bounds4, np.float32)
        wo = None if win_origin2 is None else np.ascontiguousarray(win_origin2, np.float32)
        _check(self._L.orbm_search_by_projection_kf(self._h, nfr, _ptr(gb), None if wo is None else _ptr(wo), _ptr(f_xy), _ptr(f_octave), _ptr(f_angle),
                                                    _ptr(f_desc), _ptr(f_counts), fs, _ptr(q_valid), _ptr(q_uv), _ptr(q_radius), _ptr(q_minl),
                                                    _ptr(q_maxl), _ptr(q_angle), _ptr(q_desc), _ptr(q_counts), qs, int(th_dist),
                                                    float(ratio), int(bool(check_ori)), _ptr(fm), _ptr(nm), 0))
        return nm, fm


    # ---- vocabulary-bucket matchers (ORBmatcher.cc:159-290, 524-657, 659-825) and SearchForInitialization (:407-522) ----------------------
    @staticmethod
    def _fv_slabs(fvs, slab):
        """list of CSR FeatureVectors (dicts nodes, start, items) -> slab arrays (nodes, start, items, counts, node_slab)"""
        n = len(fvs); ns = max(1, max(len(f["nodes"]) for f in fvs))
        nodes = np.zeros((n, ns), np.int32); start = np.zeros((n, ns + 1), np.int32); items = np.zeros((n, slab), np.int32); cnt = np.zeros(n, np.int32)
        for i, f in enumerate(fvs):
            k = len(f["nodes"]); cnt[i] = k
            nodes[i, :k] = f["nodes"]; start[i, :k + 1] = f["start"]; items[i, :len(f["items"])] = f["items"]
        return nodes, start, items, cnt, ns

    def SearchByBoW(self, desc1, angle1, elig1, counts1, fvs1, desc2, angle2, elig2, counts2, fvs2, triangulation=None):
        """orbm_search_by_bow for n_pairs keyframe pairs (slab layout).  fvs1 / fvs2: lists of CSR FeatureVectors.  triangulation: None for
        SearchByBoW (mfNNratio, mbCheckOrientation), or dict(xy1, xy2, octave2, F12 [n,9], epipole [n,2], scale_factors2, level_sigma2_2) for
        SearchForTriangulation.  Returns (nmatches [n_pairs], match12 [n_pairs, slab1])."""
        d1 = np.ascontiguousarray(desc1, np.uint8); d2 = np.ascontiguousarray(desc2, np.uint8); n, s1, s2 = d1.shape[0], d1.shape[1], d2.shape[1]
        a1 = np.ascontiguousarray(angle1, np.float32); a2 = np.ascontiguousarray(angle2, np.float32)
        e1 = np.ascontiguousarray(elig1, np.uint8); e2 = np.ascontiguousarray(elig2, np.uint8)
        c1 = np.ascontiguousarray(counts1, np.int32); c2 = np.ascontiguousarray(counts2, np.int32)
        n1, st1, it1, nc1, ns1 = self._fv_slabs(fvs1, s1); n2, st2, it2, nc2, ns2 = self._fv_slabs(fvs2, s2)
        m = np.zeros((n, s1), np.int32); nm = np.zeros(n, np.int32)
        ep = None; keep = []
        if triangulation is not None:
            t = triangulation
            keep = [np.ascontiguousarray(t["xy1"], np.float32), np.ascontiguousarray(t["xy2"], np.float32), np.ascontiguousarray(t["octave2"], np.int32),
                    np.ascontiguousarray(t["F12"], np.float32), np.ascontiguousarray(t["epipole"], np.float32),
                    np.ascontiguousarray(t["scale_factors2"], np.float32), np.ascontiguousarray(t["level_sigma2_2"], np.float32)]
            ep = Epipolar(*[k.ctypes.data for k in keep], len(keep[5]))
        _check(self._L.orbm_search_by_bow(self._h, n, 0 if ep is None else 1, _ptr(d1), _ptr(a1), _ptr(e1), _ptr(c1), s1, _ptr(n1), _ptr(st1), _ptr(it1), _ptr(nc1),
                                          ns1, _ptr(d2), _ptr(a2), _ptr(e2), _ptr(c2), s2, _ptr(n2), _ptr(st2), _ptr(it2), _ptr(nc2), ns2, self.mfNNratio,
                                          int(self.mbCheckOrientation), None if ep is None else ctypes.byref(ep), _ptr(m), _ptr(nm), 0))
        return nm, m

    def SearchForInitialization(self, bounds4, octave1, angle1, desc1, counts1, xy2, octave2, angle2, desc2, counts2, prev_matched, windowSize=10):
        """ORBmatcher::SearchForInitialization for n_pairs frame pairs (slab layout).  Returns (nmatches, matches12, prev_matched updated)."""
        o1 = np.ascontiguousarray(octave1, np.int32); n, s1 = o1.shape
        a1 = np.ascontiguousarray(angle1, np.float32); d1 = np.ascontiguousarray(desc1, np.uint8); c1 = np.ascontiguousarray(counts1, np.int32)
        x2 = np.ascontiguousarray(xy2, np.float32); s2 = x2.shape[1]
        o2 = np.ascontiguousarray(octave2, np.int32); a2 = np.ascontiguousarray(angle2, np.float32); d2 = np.ascontiguousarray(desc2, np.uint8)
        c2 = np.ascontiguousarray(counts2, np.int32); pm = np.ascontiguousarray(prev_matched, np.float32).copy(); b4 = np.ascontiguousarray(bounds4, np.float32)
        m = np.zeros((n, s1), np.int32); nm = np.zeros(n, np.int32)
        _check(self._L.orbm_search_for_initialization(self._h, n, _ptr(b4), _ptr(o1), _ptr(a1), _ptr(d1), _ptr(c1), s1, _ptr(x2), _ptr(o2), _ptr(a2), _ptr(d2),
                                                      _ptr(c2), s2, _ptr(pm), int(windowSize), self.mfNNratio, int(self.mbCheckOrientation), _ptr(m), _ptr(nm), 0))
        return nm, m, pm


class Epipolar(ctypes.Structure):
    """= orbm_epipolar (include/orbslamm_b200.h)"""
    _fields_ = [("xy1", ctypes.c_void_p), ("xy2", ctypes.c_void_p), ("octave2", ctypes.c_void_p), ("F12", ctypes.c_void_p), ("epipole", ctypes.c_void_p),
                ("scale_factors2", ctypes.c_void_p), ("level_sigma2_2", ctypes.c_void_p), ("nlevels", ctypes.c_int32)]


class Projection(ctypes.Structure):
    """= orbm_projection (include/orbslamm_b200.h)"""
    _fields_ = [("R", ctypes.c_float * 9), ("t", ctypes.c_float * 3), ("R2", ctypes.c_float * 9), ("t2", ctypes.c_float * 3),
                ("Ow", ctypes.c_float * 3), ("fx", ctypes.c_float), ("fy", ctypes.c_float), ("cx", ctypes.c_float), ("cy", ctypes.c_float),
                ("min_x", ctypes.c_float), ("min_y", ctypes.c_float), ("max_x", ctypes.c_float), ("max_y", ctypes.c_float),
                ("log_scale_factor", ctypes.c_float), ("th", ctypes.c_float), ("flags", ctypes.c_int32)]


# orbm_projection flags (include/orbslamm_b200.h)
PROJ_TWO_STEP, PROJ_NO_DEPTH, PROJ_FRAME_BOUNDS, PROJ_FRAME_UV, PROJ_DIST_CAMERA, PROJ_CHECK_NORMAL, PROJ_LEVEL_PLUS1 = 1, 2, 4, 8, 16, 32, 64


def make_projection(R, t, K4, bounds4, log_scale_factor, th, flags, Ow=None, R2=None, t2=None):
    """Fill an orbm_projection: the view a batch of map points is projected into (rotation / translation, optional second transform, camera centre,
    intrinsics, image bounds, log of the pyramid scale factor, window factor, PROJ_* flags)."""
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


class Optimizer:
    """Mirror of iORB_SLAM::Optimizer (reference S/include/Optimizer.h:37-68) over flat arrays: the static functions
    PoseOptimization / LocalBundleAdjustment / BundleAdjustment become methods of a handle that owns the device
    workspaces (the reference's statics are re-entrant; create one Optimizer per calling thread)."""

    def __init__(self, device=0):
        self._L = load()
        self._h = ctypes.c_void_p()
        _check(self._L.orbo_create(ctypes.byref(self._h), int(device)))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.orbo_destroy(h)

    @property
    def handle(self): return self._h
    def stream(self): return self._L.orbo_stream(self._h)
    def set_stream(self, s): _check(self._L.orbo_set_stream(self._h, ctypes.c_void_p(s)))
    def synchronize(self): _check(self._L.orbo_synchronize(self._h))
    def kernel_launches(self): return int(self._L.orbo_kernel_launches(self._h))

    @staticmethod
    def comm_unique_id():
        """128-byte NCCL unique id (create on rank 0, hand to every rank's comm_init)."""
        out = np.zeros(128, np.uint8)
        _check(load().orbo_comm_unique_id(_ptr(out)))
        return out

    def comm_init(self, nranks, rank, unique_id):
        """Collective: afterwards LocalBundleAdjustment / BundleAdjustment run sharded over `nranks` GPUs (map points
        sharded, poses replicated, one NCCL all-reduce of the reduced pose system per LM trial)."""
        uid = np.ascontiguousarray(unique_id, np.uint8)
        _check(self._L.orbo_comm_init(self._h, int(nranks), int(rank), _ptr(uid)))

    def PoseOptimization(self, Tcw, K4, Xw, obs, inv_sigma2, counts):
        """Batched Optimizer::PoseOptimization.  Tcw [n,4,4]; Xw [n,slab,3]; obs [n,slab,2]; inv_sigma2 [n,slab];
        counts [n].  Returns (Tcw_out [n,4,4], outlier u8[n,slab], n_inliers i32[n])."""
        T = np.ascontiguousarray(Tcw, np.float32).reshape(-1, 16).copy(); n = len(T)
        K4 = np.ascontiguousarray(K4, np.float32)
        Xw = np.ascontiguousarray(Xw, np.float32).reshape(n, -1, 3); slab = Xw.shape[1]
        obs = np.ascontiguousarray(obs, np.float32).reshape(n, slab, 2)
        w = np.ascontiguousarray(inv_sigma2, np.float32).reshape(n, slab)
        counts = np.ascontiguousarray(counts, np.int32)
        outl = np.zeros((n, slab), np.uint8); ninl = np.zeros(n, np.int32)
        _check(self._L.orbo_pose_optimization(self._h, n, _ptr(T), _ptr(K4), _ptr(Xw), _ptr(obs), _ptr(w), _ptr(counts), slab,
                                              _ptr(outl), _ptr(ninl), 0))
        return T.reshape(n, 4, 4), outl, ninl

    def PoseOptimizationMatched(self, Tcw, K4, f_xy, f_octave, f_counts, feat_match, q_Xw, q_counts, inv_level_sigma2):
        """PoseOptimization with the edges gathered on the device from a SearchByProjection result (host arrays here).
        Returns (Tcw_out, outlier u8[n, f_slab], n_inliers, n_edges)."""
        T = np.ascontiguousarray(Tcw, np.float32).reshape(-1, 16).copy(); n = len(T)
        K4 = np.ascontiguousarray(K4, np.float32)
        f_xy = np.ascontiguousarray(f_xy, np.float32); fs = f_xy.shape[1]
        f_octave = np.ascontiguousarray(f_octave, np.int32); f_counts = np.ascontiguousarray(f_counts, np.int32)
        fm = np.ascontiguousarray(feat_match, np.int32)
        q_Xw = np.ascontiguousarray(q_Xw, np.float32).reshape(n, -1, 3); qs = q_Xw.shape[1]
        q_counts = np.ascontiguousarray(q_counts, np.int32); ils = np.ascontiguousarray(inv_level_sigma2, np.float32)
        outl = np.zeros((n, fs), np.uint8); ninl = np.zeros(n, np.int32); ne = np.zeros(n, np.int32)
        _check(self._L.orbo_pose_optimization_matched(self._h, n, _ptr(T), _ptr(K4), _ptr(f_xy), _ptr(f_octave), _ptr(f_counts), fs, _ptr(fm),
                                                      _ptr(q_Xw), _ptr(q_counts), qs, _ptr(ils), len(ils), _ptr(outl), _ptr(ninl), _ptr(ne), 0))
        return T.reshape(n, 4, 4), outl, ninl, ne

    def OptimizeSim3(self, sim3, valid, P1c, P2c, obs1, obs2, inv_sigma2_1, inv_sigma2_2, K1, K2, counts, th2=10.0, bFixScale=False):
        """Optimizer::OptimizeSim3 (Optimizer.cc:1348-1543) for n_pairs keyframe pairs (slab layout [n_pairs, slab, ...]).
        sim3 [n_pairs, 8] = r (x y z w), t, s of g2oS12.  Returns (sim3, inlier u8, n_inliers, lm_stats [n_pairs, 2])."""
        S = np.ascontiguousarray(sim3, np.float64).reshape(-1, 8).copy(); n = len(S)
        v = np.ascontiguousarray(valid, np.uint8).reshape(n, -1); slab = v.shape[1]
        a = [np.ascontiguousarray(x, np.float32) for x in (P1c, P2c, obs1, obs2, inv_sigma2_1, inv_sigma2_2)]
        K1 = np.ascontiguousarray(K1, np.float32).reshape(n, 4); K2 = np.ascontiguousarray(K2, np.float32).reshape(n, 4)
        cnt = np.ascontiguousarray(counts, np.int32)
        inl = np.zeros((n, slab), np.uint8); nin = np.zeros(n, np.int32); st = np.zeros((n, 2), np.int32)
        _check(self._L.orbo_optimize_sim3(self._h, n, _ptr(S), _ptr(v), *[_ptr(x) for x in a], _ptr(K1), _ptr(K2), _ptr(cnt), slab, float(th2),
                                          int(bool(bFixScale)), _ptr(inl), _ptr(nin), _ptr(st), 0))
        return S, inl, nin, st

    def Sim3Prepare(self, X3Dc, octave, level_sigma2, K4):
        """Sim3Solver constructor data for one keyframe: (max_err i32[N], p2d f32[N,2])."""
        X = np.ascontiguousarray(X3Dc, np.float32); N = len(X)
        o = np.ascontiguousarray(octave, np.int32); ls = np.ascontiguousarray(level_sigma2, np.float32); K = np.ascontiguousarray(K4, np.float32)
        me = np.zeros(N, np.int32); p = np.zeros((N, 2), np.float32)
        _check(self._L.orbo_sim3_prepare(self._h, N, _ptr(X), _ptr(o), _ptr(ls), len(ls), _ptr(K), _ptr(me), _ptr(p), 0))
        return me, p

    def Sim3Compute(self, X1, X2, bFixScale=False):
        """Sim3Solver::ComputeSim3 for a batch of min sets: X1, X2 f32[n_hyp, 3, 3] (three points each).  Returns (T12, T21 [n_hyp,4,4], R12 [n_hyp,3,3], t12, s12)."""
        a = np.ascontiguousarray(X1, np.float32).reshape(-1, 9); b = np.ascontiguousarray(X2, np.float32).reshape(-1, 9); n = len(a)
        T12 = np.zeros((n, 16), np.float32); T21 = np.zeros((n, 16), np.float32); R = np.zeros((n, 13), np.float32)
        self._L.orbo_sim3_compute.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        self._L.orbo_sim3_compute.restype = ctypes.c_int
        _check(self._L.orbo_sim3_compute(self._h, n, _ptr(a), _ptr(b), int(bool(bFixScale)), _ptr(T12), _ptr(T21), _ptr(R), 0))
        return T12.reshape(n, 4, 4), T21.reshape(n, 4, 4), R[:, :9].reshape(n, 3, 3), R[:, 9:12], R[:, 12]

    def Sim3CheckInliers(self, T12, T21, X3Dc1, X3Dc2, P1im1, P2im2, max_err1, max_err2, K1, K2):
        """Sim3Solver::CheckInliers for a batch of RANSAC hypotheses: (inliers u8[n_hyp, N], n_inliers i32[n_hyp])."""
        a = [np.ascontiguousarray(x, np.float32) for x in (T12, T21, X3Dc1, X3Dc2, P1im1, P2im2)]
        nh = a[0].reshape(-1, 16).shape[0]; N = len(a[2])
        m1 = np.ascontiguousarray(max_err1, np.int32); m2 = np.ascontiguousarray(max_err2, np.int32)
        k1 = np.ascontiguousarray(K1, np.float32); k2 = np.ascontiguousarray(K2, np.float32)
        inl = np.zeros((nh, N), np.uint8); n = np.zeros(nh, np.int32)
        _check(self._L.orbo_sim3_check_inliers(self._h, nh, _ptr(a[0]), _ptr(a[1]), N, _ptr(a[2]), _ptr(a[3]), _ptr(a[4]), _ptr(a[5]), _ptr(m1), _ptr(m2), _ptr(k1),
                                               _ptr(k2), _ptr(inl), _ptr(n), 0))
        return inl, n

    def OptimizePoseGraph(self, sim3, fixed, e_i, e_j, e_meas, bFixScale=False, iterations=20, lambda_init=1e-16):
        """Numeric core of Optimizer::OptimizeEssentialGraph (Optimizer.cc:804-1067): sim3 [K,8], e_meas [E,8] = Sji.  Returns dict(sim3, lm_iterations,
        lm_trials, chol_failures)."""
        S = np.ascontiguousarray(sim3, np.float64).reshape(-1, 8).copy()
        fx = np.ascontiguousarray(fixed, np.uint8); ei = np.ascontiguousarray(e_i, np.int32); ej = np.ascontiguousarray(e_j, np.int32)
        em = np.ascontiguousarray(e_meas, np.float64).reshape(-1, 8); st = np.zeros(3, np.int32)
        _check(self._L.orbo_optimize_pose_graph(self._h, len(S), _ptr(S), _ptr(fx), len(ei), _ptr(ei), _ptr(ej), _ptr(em), int(bool(bFixScale)), int(iterations),
                                                float(lambda_init), _ptr(st)))
        return dict(sim3=S, lm_iterations=int(st[0]), lm_trials=int(st[1]), chol_failures=int(st[2]))

    def _ba(self, poses, fixed, intr, points, e_kf, e_pt, e_uv, e_inv_sigma2, two_stage, its0, its1, robust, stop_flag=None):
        poses = np.ascontiguousarray(poses, np.float32).reshape(-1, 16).copy(); K = len(poses)
        fixed = np.ascontiguousarray(fixed, np.uint8)
        intr = np.ascontiguousarray(intr, np.float64)
        if intr.ndim == 1:
            intr = np.tile(intr, (K, 1))
        intr = np.ascontiguousarray(intr)
        points = np.ascontiguousarray(points, np.float32).reshape(-1, 3).copy(); P = len(points)
        e_kf = np.ascontiguousarray(e_kf, np.int32); e_pt = np.ascontiguousarray(e_pt, np.int32); E = len(e_kf)
        e_uv = np.ascontiguousarray(e_uv, np.float32); w = np.ascontiguousarray(e_inv_sigma2, np.float32)
        chi2 = np.zeros(E); dok = np.zeros(E, np.uint8); outl = np.zeros(E, np.uint8); stats = np.zeros(4, np.int32)
        if stop_flag is not None:
            assert stop_flag.dtype == np.uint8, "stop_flag is one byte (C++ bool)"
        sp = None if stop_flag is None else stop_flag.ctypes.data
        rc = self._L.orbo_bundle_adjust(self._h, K, _ptr(poses), _ptr(fixed), _ptr(intr), P, _ptr(points), E, _ptr(e_kf), _ptr(e_pt),
                                        _ptr(e_uv), _ptr(w), int(two_stage), int(its0), int(its1), int(robust), sp, _ptr(chi2),
                                        _ptr(dok), _ptr(outl), _ptr(stats))
        if rc < 0:
            _check(rc)
        return dict(poses=poses.reshape(K, 4, 4), points=points, chi2=chi2, depth_ok=dok, outlier=outl,
                    lm_iterations=int(stats[0]), lm_trials=int(stats[1]), chol_failures=int(stats[2]), aborted=bool(rc == 1))

    KERNELS = ("errors", "build_points", "build_poses", "point_prep", "schur", "reduced_solve", "backsub", "update", "lm_decide", "exchange")

    def comm_mode(self):
        """0 = single GPU, 1 = NCCL all-reduce, 2 = NVLink peer-memory exchange"""
        self._L.orbo_comm_mode.argtypes = [ctypes.c_void_p]; self._L.orbo_comm_mode.restype = ctypes.c_int
        return int(self._L.orbo_comm_mode(self._h))

    def set_profiling(self, on):
        self._L.orbo_set_profiling.argtypes = [ctypes.c_void_p, ctypes.c_int]
        _check(self._L.orbo_set_profiling(self._h, int(bool(on))))

    def kernel_times(self):
        ms = np.zeros(len(self.KERNELS)); cnt = np.zeros(len(self.KERNELS), np.int64)
        self._L.orbo_get_kernel_times.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        _check(self._L.orbo_get_kernel_times(self._h, _ptr(ms), _ptr(cnt), len(ms)))
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(self.KERNELS)}

    def last_ba_timing(self):
        out = np.zeros(4)
        self._L.orbo_last_ba_timing.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        _check(self._L.orbo_last_ba_timing(self._h, _ptr(out)))
        sk = np.zeros(3, np.int64)
        self._L.orbo_last_ba_structure.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        _check(self._L.orbo_last_ba_structure(self._h, _ptr(sk)))
        return dict(lm_loop_s=float(out[0]), total_s=float(out[1]), setup_s=float(out[2]), ld=int(out[3]), tile_rows=int(sk[0]), l_tiles=int(sk[1]),
                    levels=int(sk[2]))

    def LocalBundleAdjustment(self, poses, fixed, intr, points, e_kf, e_pt, e_uv, e_inv_sigma2, stop_flag=None):
        """Optimizer::LocalBundleAdjustment schedule: 5 robust LM iterations, chi2/depth gating, 10 non-robust."""
        return self._ba(poses, fixed, intr, points, e_kf, e_pt, e_uv, e_inv_sigma2, True, 5, 10, True, stop_flag)

    def BundleAdjustment(self, poses, fixed, intr, points, e_kf, e_pt, e_uv, e_inv_sigma2, nIterations=5, bRobust=True, stop_flag=None):
        """Optimizer::BundleAdjustment (GlobalBundleAdjustemnt / MMGlobalBundleAdjustemnt core)."""
        return self._ba(poses, fixed, intr, points, e_kf, e_pt, e_uv, e_inv_sigma2, False, nIterations, 0, bRobust, stop_flag)


class FrontEnd:
    """Fused per-frame front end over one extractor / matcher / optimizer triple (orbf_*): host images in,
    keypoints + matches + optimised poses out, everything in between stays on the device."""

    def __init__(self, extractor, matcher, optimizer, device=0):
        self._L = load()
        self.ex, self.mt, self.po = extractor, matcher, optimizer       # keep them alive
        self._h = ctypes.c_void_p()
        _check(self._L.orbf_create(ctypes.byref(self._h), extractor.handle, matcher.handle, optimizer.handle, int(device)))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.orbf_destroy(h)

    @property
    def handle(self): return self._h

    def track_frames(self, images, K4, Tcw, q_Xw, q_octave, q_angle, q_desc, q_valid, q_counts, th_proj=15.0, th_dist=100, check_ori=True):
        images = np.ascontiguousarray(images, np.uint8); B, H, W = images.shape
        cap = self.ex.max_keypoints(W, H)
        K4 = np.ascontiguousarray(K4, np.float32); T = np.ascontiguousarray(Tcw, np.float32).reshape(B, 16).copy()
        sf = self.ex.GetScaleFactors(); ils = self.ex.GetInverseScaleSigmaSquares()
        q_Xw = np.ascontiguousarray(q_Xw, np.float32).reshape(B, -1, 3); qs = q_Xw.shape[1]
        q_octave = np.ascontiguousarray(q_octave, np.int32); q_angle = np.ascontiguousarray(q_angle, np.float32)
        q_desc = np.ascontiguousarray(q_desc, np.uint8); q_valid = np.ascontiguousarray(q_valid, np.uint8)
        q_counts = np.ascontiguousarray(q_counts, np.int32)
        xy = np.empty((B, cap, 2), np.float32); ang = np.empty((B, cap), np.float32); resp = np.empty((B, cap), np.float32)
        octv = np.empty((B, cap), np.int32); size = np.empty((B, cap), np.float32); desc = np.empty((B, cap, 32), np.uint8)
        counts = np.zeros(B, np.int32); fm = np.empty((B, cap), np.int32); nm = np.zeros(B, np.int32)
        outl = np.empty((B, cap), np.uint8); ninl = np.zeros(B, np.int32)
        _check(self._L.orbf_track_frames(self._h, _ptr(images), B, W, H, images.strides[1], images.strides[0], _ptr(K4), _ptr(sf), _ptr(ils), len(sf),
                                         _ptr(q_Xw), _ptr(q_octave), _ptr(q_angle), _ptr(q_desc), _ptr(q_valid), _ptr(q_counts), qs, float(th_proj),
                                         int(th_dist), int(bool(check_ori)), _ptr(T), _ptr(xy), _ptr(ang), _ptr(resp), _ptr(octv), _ptr(size),
                                         _ptr(desc), cap, _ptr(counts), _ptr(fm), _ptr(nm), _ptr(outl), _ptr(ninl)))
        return dict(Tcw=T.reshape(B, 4, 4), xy=xy, angle=ang, response=resp, octave=octv, size=size, desc=desc, counts=counts,
                    feat_match=fm, nmatches=nm, outlier=outl, n_inliers=ninl)
