"""Fused front end (orbf_track_frames): extract -> SearchByProjection(Cur, Last) -> PoseOptimization in one host-buffer call,
checked stage by stage against the oracle; also exercises chunked uploads (batch > 16 frames)."""
import numpy as np
import pytest

import oracle
from orbslamm_b200 import synth
from helpers import make_tracking_case, slab

pytestmark = pytest.mark.gpu


def test_track_frames_matches_oracle(lib):
    import orbslamm_b200 as ob
    cam = dict(synth.TUM); cam.update(w=360, h=280, nfeatures=300, cx=180.0, cy=140.0)
    cases = [make_tracking_case(cam, sid) for sid in range(40, 59)]          # 19 frames -> two upload chunks
    P = cases[0]["P"]
    ex = ob.ORBextractor(cam["nfeatures"], 1.2, 8, 20, 7); mt = ob.ORBmatcher(0.9, True); po = ob.Optimizer()
    fe = ob.FrontEnd(ex, mt, po)
    nQ = max(len(c["last"]["x"]) for c in cases)
    T0 = []
    for c in cases:
        T = c["Tcw"].copy(); T[:3, 3] += np.array([0.03, -0.02, 0.05], np.float32); T0.append(T)
    r = fe.track_frames(np.stack([c["frames"][1] for c in cases]), cases[0]["K4"], np.stack(T0), slab([c["Xw"] for c in cases], nQ, np.float32, (3,)),
                        slab([c["last"]["octave"] for c in cases], nQ, np.int32), slab([c["last"]["angle"] for c in cases], nQ, np.float32),
                        slab([c["last"]["desc"] for c in cases], nQ, np.uint8, (32,)), slab([c["valid"] for c in cases], nQ, np.uint8),
                        np.array([len(c["last"]["x"]) for c in cases], np.int32))
    sf = np.array(list(P.scale)[:8], np.float32); ils = np.array(list(P.inv_sigma2)[:8], np.float32)
    g = oracle.grid_params(0, 0, cam["w"], cam["h"])
    for i, c in enumerate(cases):
        cur, last = c["cur"], c["last"]
        n = len(cur["x"])
        assert r["counts"][i] == n
        assert np.array_equal(r["xy"][i, :n, 0], cur["x"]) and np.array_equal(r["xy"][i, :n, 1], cur["y"]) and np.array_equal(r["desc"][i, :n], cur["desc"])
        assert np.array_equal(r["angle"][i, :n], cur["angle"]) and np.array_equal(r["octave"][i, :n], cur["octave"])
        q = oracle.project_last_frame(T0[i], c["K4"], g, sf, c["Xw"], last["octave"], 15.0, c["valid"])
        fxy = np.stack([cur["x"], cur["y"]], 1)
        n_ref, fm_ref = oracle.search_by_projection(g, fxy, cur["octave"], cur["angle"], cur["desc"], q[0], q[1], q[2], q[3], q[4], last["angle"], last["desc"], 100, 0.0, True)
        assert r["nmatches"][i] == n_ref and np.array_equal(r["feat_match"][i, :n], fm_ref)
        m = fm_ref >= 0
        Tr, outr, nr = oracle.pose_optimization(T0[i], c["Xw"][fm_ref[m]], fxy[m], ils[cur["octave"][m]], c["K4"])
        assert r["n_inliers"][i] == nr and np.array_equal(r["outlier"][i, :n][m], outr)
        assert np.abs(r["Tcw"][i] - Tr).max() < 1e-5 * np.abs(Tr).max()
