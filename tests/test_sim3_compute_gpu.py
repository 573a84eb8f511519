"""orbo_sim3_compute (Sim3Solver::ComputeSim3, S/src/Sim3Solver.cc:226-338, for a batch of RANSAC min sets) against the cv2-backed restatement of the reference's
OpenCV call sequence (tests/map_merge.py::compute_sim3: cv2.reduce / gemm / eigen / norm / Rodrigues / pow on float matrices).

Tolerance, not bit-exactness: cv::eigen is OpenCV-build dependent (Jacobi in the 2.4 / 3.0 the reference names, Eigen's solver when OpenCV is built with Eigen) and
its result passes through atan2 / Rodrigues; the two sides agree to a few float32 steps.  What the RANSAC consumes is the inlier count of every hypothesis: those
must agree except where a correspondence sits within rounding of its threshold."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import kf_family as kff
import map_merge as M
import oracle
from orbslamm_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cam,sid,fix", [("TUM", 1, False), ("KITTI", 2, False), ("KITTI", 2, True)])
def test_compute_sim3_matches_the_opencv_call_sequence(lib, cam, sid, fix):
    import orbslamm_b200 as ob
    r = kff.make_sim3_ransac_case(getattr(synth, cam), sid)
    X1, X2 = r["X1"], r["X2"]; N = len(X1)
    rng = np.random.default_rng(5)
    tri = M.sample_triples(N, 300, rng)
    opt = ob.Optimizer()
    T12, T21, R, t, s = opt.Sim3Compute(X1[tri], X2[tri], fix)
    ref = [M.compute_sim3(X1[q].T.copy(), X2[q].T.copy()) if not fix else None for q in tri]
    if fix:                                                                 # fixed scale: s12 = 1 (Sim3Solver.cc:303-304); compare against the free-scale rotation
        assert np.all(s == 1.0)
        Tf = opt.Sim3Compute(X1[tri], X2[tri], False)
        assert np.abs(R - Tf[2]).max() == 0.0
        return
    worst = 0.0
    for h in range(len(tri)):
        T12r, T21r, Rr, tr, sr = ref[h]
        scale = max(np.abs(T12r).max(), 1.0)
        worst = max(worst, float(np.abs(T12[h] - T12r).max() / scale), float(np.abs(T21[h] - T21r).max() / max(np.abs(T21r).max(), 1.0)))
        assert abs(float(s[h]) - float(sr)) < 1e-4 * abs(float(sr)) and np.abs(R[h] - Rr).max() < 1e-4
    assert worst < 1e-4, worst
    # what the RANSAC sees: inlier counts per hypothesis from the device hypotheses vs from the reference-sequence hypotheses
    m1, m2, p1, p2 = oracle.sim3_prepare(X1, X2, r["oct1"], r["oct2"], r["ls2"], r["K1"], r["K2"])
    _, cnt_d = oracle.sim3_check_inliers(T12, T21, X1, X2, p1, p2, m1, m2, r["K1"], r["K2"])
    _, cnt_r = oracle.sim3_check_inliers(np.stack([x[0] for x in ref]), np.stack([x[1] for x in ref]), X1, X2, p1, p2, m1, m2, r["K1"], r["K2"])
    assert np.abs(cnt_d - cnt_r).max() <= 2 and (cnt_d == cnt_r).mean() > 0.9 and cnt_r.max() > 20


def test_compute_sim3_recovers_an_exact_similarity(lib):
    """property: three exact correspondences X1 = s R X2 + t give back (s, R, t) and T21 = T12^-1"""
    import orbslamm_b200 as ob
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(1)
    n = 64
    X2 = rng.uniform(-5, 5, (n, 3, 3)) + np.array([0, 0, 12.0])
    Rs = Rotation.from_rotvec(rng.normal(0, 0.6, (n, 3))).as_matrix(); ts = rng.normal(0, 2, (n, 3)); ss = rng.uniform(0.5, 2.0, n)
    X1 = ss[:, None, None] * np.einsum("nij,nkj->nki", Rs, X2) + ts[:, None, :]
    T12, T21, R, t, s = ob.Optimizer().Sim3Compute(X1.astype(np.float32), X2.astype(np.float32))
    assert np.abs(s - ss).max() < 2e-4 and np.abs(R - Rs).max() < 2e-4 and np.abs(t - ts).max() < 5e-3
    I = np.einsum("nij,njk->nik", T12.astype(np.float64), T21.astype(np.float64))
    assert np.abs(I - np.eye(4)).max() < 1e-4
