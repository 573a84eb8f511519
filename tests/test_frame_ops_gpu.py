"""GPU parity of the Frame steps either side of the matcher: UndistortKeyPoints and isInFrustum."""
import numpy as np
import pytest

import oracle
from orbslamm_b200 import synth

pytestmark = pytest.mark.gpu

TUM1_K4 = np.array([517.306408, 516.469215, 318.643040, 255.313989], np.float32)            # S/Examples/Monocular/TUM1.yaml:8-11
TUM1_DIST = np.array([0.262383, -0.953104, -0.005358, 0.002628, 1.163314], np.float32)      # :13-17


def test_undistort_keypoints_matches_oracle(lib):
    import orbslamm_b200 as ob
    rng = np.random.default_rng(3)
    m = ob.ORBmatcher(0.9, True)
    nfr, slab = 3, 1500
    counts = np.array([1500, 777, 0], np.int32)
    xy = np.stack([rng.uniform(0, 640, (nfr, slab)), rng.uniform(0, 480, (nfr, slab))], 2).astype(np.float32)
    for dist in (TUM1_DIST, TUM1_DIST[:4], np.zeros(4, np.float32)):
        got = m.UndistortKeyPoints(xy, counts, TUM1_K4, dist)
        for f in range(nfr):
            ref = oracle.undistort_points(TUM1_K4, dist, xy[f, :counts[f]])
            assert np.array_equal(got[f, :counts[f]], ref), f"frame {f}, {len(dist)} coefficients"


def test_is_in_frustum_matches_oracle(lib):
    import orbslamm_b200 as ob
    rng = np.random.default_rng(4)
    cam = synth.KITTI
    K4 = np.array([cam["fx"], cam["fy"], cam["cx"], cam["cy"]], np.float32)
    bounds = np.array([0, 0, cam["w"], cam["h"]], np.float32)
    nfr, slab = 2, 4000
    counts = np.array([4000, 2500], np.int32)
    Tcw = np.zeros((nfr, 4, 4), np.float32); Ow = np.zeros((nfr, 3), np.float32)
    for f in range(nfr):
        a = np.deg2rad(3.0 * (f + 1))
        R = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]], np.float32)
        t = np.array([0.3 * f, -0.1, 0.2], np.float32)
        Tcw[f, :3, :3] = R; Tcw[f, :3, 3] = t; Tcw[f, 3, 3] = 1
        Ow[f] = -(R.T @ t)
    # points around the camera: in front / behind / outside the image / too near / too far / seen from behind
    Xw = np.stack([rng.uniform(-40, 40, (nfr, slab)), rng.uniform(-8, 8, (nfr, slab)), rng.uniform(-10, 80, (nfr, slab))], 2).astype(np.float32)
    # normals = mean viewing direction (camera -> point) with noise, some flipped; scale-invariance ranges around the true distance
    po = Xw - Ow[:, None, :]
    dist = np.linalg.norm(po, axis=2, keepdims=True)
    nrm = po / dist + rng.normal(0, 0.5, (nfr, slab, 3))
    nrm[rng.random((nfr, slab)) < 0.15] *= -1
    nrm = (nrm / np.linalg.norm(nrm, axis=2, keepdims=True)).astype(np.float32)
    mf_max = (dist[..., 0] * rng.uniform(0.5, 3.0, (nfr, slab))).astype(np.float32); mf_min = (mf_max / rng.uniform(2, 4.3, (nfr, slab))).astype(np.float32)
    log_sf = float(np.log(np.float32(1.2)))
    m = ob.ORBmatcher(0.8, True)
    iv, uv, lv, vc = m.isInFrustum(Tcw, Ow, K4, bounds, log_sf, Xw, nrm, mf_min, mf_max, counts, 0.5)
    seen = 0
    for f in range(nfr):
        n = counts[f]
        r_iv, r_uv, r_lv, r_vc = oracle.is_in_frustum(Tcw[f], Ow[f], K4, bounds, log_sf, 0.5, Xw[f, :n], nrm[f, :n], mf_min[f, :n], mf_max[f, :n])
        assert np.array_equal(iv[f, :n], r_iv)
        k = r_iv > 0
        seen += int(k.sum())
        assert np.array_equal(uv[f, :n][k], r_uv[k]) and np.array_equal(vc[f, :n][k], r_vc[k]) and np.array_equal(lv[f, :n][k], r_lv[k])
        assert not iv[f, n:].any()
    assert 200 < seen < counts.sum() - 200          # the case exercises both outcomes


def test_assign_features_to_grid_and_get_features_in_area(lib):
    """Frame::AssignFeaturesToGrid (Frame.cc:230-245) and GetFeaturesInArea (Frame.cc:327-380, KeyFrame.cc:618-657) as stand-alone calls:
    the grid equals the oracle's cell by cell, the returned index lists are identical including their order; distorted (non-integer) bounds,
    a keyframe's integer window origin, level limits, truncation at cap and windows outside the image are covered."""
    import orbslamm_b200 as ob
    from helpers import make_tracking_case, slab
    cases = [make_tracking_case(synth.TUM, 2), make_tracking_case(synth.KITTI, 3)]
    m = ob.ORBmatcher()
    rng = np.random.default_rng(4)
    for bounds in (np.array([0, 0, 1241, 480], np.float32), np.array([-27.3, -20.7, 1266.6, 498.2], np.float32)):
        g = oracle.grid_params(*[float(b) for b in bounds])
        fs = max(len(k["cur"]["x"]) for k in cases) + 3
        fxy = slab([np.stack([k["cur"]["x"], k["cur"]["y"]], 1) for k in cases], fs, np.float32, (2,))
        foc = slab([k["cur"]["octave"] for k in cases], fs, np.int32)
        fc = np.array([len(k["cur"]["x"]) for k in cases], np.int32)
        cs, ci = m.AssignFeaturesToGrid(bounds, fxy, fc)
        Q = 300
        q = np.zeros((2, Q, 3), np.float32)
        q[:, :, 0] = rng.uniform(bounds[0] - 30, bounds[2] + 30, (2, Q)); q[:, :, 1] = rng.uniform(bounds[1] - 30, bounds[3] + 30, (2, Q))
        q[:, :, 2] = rng.choice([2.5, 7.5, 15.0, 40.0, 100.0], (2, Q))
        lv = rng.integers(0, 8, (2, Q)).astype(np.int32)
        mn = np.where(rng.random((2, Q)) < 0.3, -1, lv - 1).astype(np.int32); mx = np.where(mn < 0, -1, lv + rng.integers(0, 2, (2, Q))).astype(np.int32)
        idx, cnt = m.GetFeaturesInArea(bounds, fxy, foc, fc, q, mn, mx, [Q, Q - 7], cap=64)
        wo = np.trunc(bounds[:2]).astype(np.float32)
        idx_k, cnt_k = m.GetFeaturesInArea(bounds, fxy, None, fc, q, None, None, [Q, Q - 7], cap=512, win_origin2=wo)
        gw = oracle.grid_params(*[float(b) for b in bounds]); gw.min_x, gw.min_y = float(wo[0]), float(wo[1])
        for f, k in enumerate(cases):
            n = fc[f]
            ocs, oci = oracle.grid_build(g, fxy[f, :n])
            assert np.array_equal(cs[f], ocs) and np.array_equal(ci[f, :ocs[-1]], oci[:ocs[-1]])
            nq = [Q, Q - 7][f]
            for i in range(nq):
                ref = oracle.features_in_area(g, ocs, oci, fxy[f, :n], foc[f, :n], q[f, i, 0], q[f, i, 1], q[f, i, 2], mn[f, i], mx[f, i])
                assert cnt[f, i] == len(ref) and np.array_equal(idx[f, i, :min(len(ref), 64)], ref[:64])
                ref = oracle.features_in_area(gw, ocs, oci, fxy[f, :n], foc[f, :n], q[f, i, 0], q[f, i, 1], q[f, i, 2], -1, -1)
                assert cnt_k[f, i] == len(ref) and np.array_equal(idx_k[f, i, :len(ref)], ref)
            assert (cnt[f, nq:] == 0).all()
        assert cnt.max() > 64                                     # truncation exercised
