"""Pins the oracle's restated OpenCV primitives against the real cv2 (the primitives the reference calls)."""
import numpy as np
import pytest

import oracle
from oracle import orb_cv2
from orbslamm_b200 import synth

cv2 = pytest.importorskip("cv2")


def _images():
    rng = np.random.default_rng(0)
    yield synth.base_image(640, 480, 5)
    yield synth.base_image(1241, 376, 6)
    yield rng.integers(0, 256, (97, 131), dtype=np.uint8)
    yield (rng.random((200, 333)) > 0.5).astype(np.uint8) * 255
    yield np.full((64, 80), 77, np.uint8)


def test_resize_linear_bit_exact():
    rng = np.random.default_rng(1)
    for img in _images():
        h, w = img.shape
        for _ in range(3):
            dw, dh = int(rng.integers(max(2, w // 2), w)), int(rng.integers(max(2, h // 2), h))
            ref = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)
            assert np.array_equal(oracle.resize_linear(img, dw, dh), ref)
        dw, dh = int(np.rint(np.float32(w) * np.float32(1 / 1.2))), int(np.rint(np.float32(h) * np.float32(1 / 1.2)))
        assert np.array_equal(oracle.resize_linear(img, dw, dh), cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR))


def test_gaussian_blur_bit_exact():
    for img in _images():
        ref = cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
        assert np.array_equal(oracle.gaussian_blur7(img), ref)


@pytest.mark.parametrize("th", [20, 7])
def test_fast_bit_exact(th):
    det = cv2.FastFeatureDetector_create(th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    rng = np.random.default_rng(2)
    imgs = list(_images())
    for _ in range(40):       # cell-sized sub-images like the reference's per-cell calls
        h, w = int(rng.integers(7, 45)), int(rng.integers(7, 45))
        imgs.append(cv2.GaussianBlur(rng.integers(0, 256, (h, w), dtype=np.uint8), (5, 5), 1.2))
    for img in imgs:
        ref = np.array([(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in det.detect(img)], np.int32).reshape(-1, 3)
        assert np.array_equal(oracle.fast9_16(img, th, True), ref)


def test_fast_atan2_bit_exact():
    rng = np.random.default_rng(3)
    ys = rng.integers(-300000, 300000, 20000); xs = rng.integers(-300000, 300000, 20000)
    for y, x in zip(ys, xs):
        assert np.float32(oracle.fast_atan2(float(y), float(x))) == np.float32(cv2.fastAtan2(float(y), float(x)))
    for y, x in [(0, 0), (0, -5), (-3, 0), (3, 0), (0, 5), (1, 1), (-1, -1)]:
        assert np.float32(oracle.fast_atan2(float(y), float(x))) == np.float32(cv2.fastAtan2(float(y), float(x)))


def test_small_gemm_is_sequential_fp32():
    """cv::Mat Rcw*x3Dw+tcw (ORBmatcher.cc:1361) = cv::gemm small path: fp32 products added left to right."""
    rng = np.random.default_rng(4)
    g = oracle.grid_params(-1e9, -1e9, 1e9, 1e9)
    for _ in range(300):
        R = rng.normal(size=(3, 3)).astype(np.float32); x = (rng.normal(size=(3, 1)) * 10).astype(np.float32)
        t = rng.normal(size=(3, 1)).astype(np.float32)
        ref = cv2.gemm(R, x, 1, t, 1).ravel()
        if ref[2] <= 0:
            continue
        T = np.eye(4, dtype=np.float32); T[:3, :3] = R; T[:3, 3] = t.ravel()
        K4 = np.array([1, 1, 0, 0], np.float32)       # u = xc * invz, v = yc * invz
        qv, uv, *_ = oracle.project_last_frame(T, K4, g, np.ones(8, np.float32), x.T.copy(), np.zeros(1, np.int32), 1.0, np.ones(1, np.uint8))
        invz = np.float32(1.0 / np.float64(ref[2]))
        assert qv[0] == 1
        assert uv[0, 0] == np.float32(np.float32(np.float32(1) * ref[0]) * invz) and uv[0, 1] == np.float32(np.float32(np.float32(1) * ref[1]) * invz)


@pytest.mark.parametrize("cam", ["TUM", "KITTI"])
def test_full_extractor_c_equals_cv2_path(cam):
    """orb_oracle.c (restated primitives) == the same extractor driven by real cv2 primitives."""
    c = getattr(synth, cam)
    P = oracle.orb_params(c["nfeatures"], 1.2, 8, 20, 7)
    img = synth.stream(c["w"], c["h"], 2, stream_id=9)[0][1]
    a, b = oracle.orb_extract(P, img), orb_cv2.extract(P, img)
    assert len(a["x"]) > 0.9 * c["nfeatures"]
    for k in ("x", "y", "angle", "response", "octave", "size", "desc", "level_counts"):
        assert np.array_equal(a[k], b[k]), k
    lv, bordered = orb_cv2.compute_pyramid(P, img, with_border=True)
    for x, y in zip(oracle.pyramid(P, img), lv):
        assert np.array_equal(x, y)


def test_undistort_points_matches_cv2():
    """Frame::UndistortKeyPoints -> cv::undistortPoints(pts, K, dist, R=I, P=K) (Frame.cc:423): restated iteration == cv2, bit for bit."""
    rng = np.random.default_rng(11)
    K4 = np.array([517.306408, 516.469215, 318.643040, 255.313989], np.float32)
    K = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1]], np.float32)
    for dist in (np.array([0.262383, -0.953104, -0.005358, 0.002628, 1.163314], np.float32),
                 np.array([0.262383, -0.953104, -0.005358, 0.002628], np.float32), np.array([-0.28, 0.07, 0.0002, -0.0003], np.float32)):
        xy = np.stack([rng.uniform(0, 640, 5000), rng.uniform(0, 480, 5000)], 1).astype(np.float32)
        ref = cv2.undistortPoints(xy.reshape(-1, 1, 2), K, dist.reshape(-1, 1), None, K).reshape(-1, 2)
        assert np.array_equal(oracle.undistort_points(K4, dist, xy), ref)
    xy = rng.uniform(0, 400, (50, 2)).astype(np.float32)
    assert np.array_equal(oracle.undistort_points(K4, np.zeros(4, np.float32), xy), xy)      # Frame.cc:406-410


def test_norm_accumulates_in_double():
    """cv::norm(PO) in Frame::isInFrustum (Frame.cc:300): L2 norm of a float 3-vector accumulated in double."""
    rng = np.random.default_rng(12)
    v = rng.normal(0, 10, (2000, 3)).astype(np.float32)
    ref = np.array([np.float32(cv2.norm(x)) for x in v])
    assert np.array_equal(ref, np.sqrt((v.astype(np.float64) ** 2).sum(1)).astype(np.float32))
