"""The oracle reproduces the committed golden vectors (guards the checker itself against drift)."""
import os

import numpy as np

import oracle

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_extractor_golden():
    d = np.load(os.path.join(G, "extract_400x300.npz"))
    P = oracle.orb_params(int(d["params"][0]), float(d["scale_factor"]), int(d["params"][1]), int(d["params"][2]), int(d["params"][3]))
    for i in (0, 1):
        r = oracle.orb_extract(P, d[f"frame{i}"])
        for k in ("x", "y", "angle", "response", "octave", "size", "desc"):
            assert np.array_equal(r[k], d[f"f{i}_{k}"]), (i, k)


def test_matcher_golden():
    e = np.load(os.path.join(G, "extract_400x300.npz")); d = np.load(os.path.join(G, "match_400x300.npz"))
    g = oracle.grid_params(*d["bounds"])
    qv, uv, rad, mn, mx = oracle.project_last_frame(d["Tcw"], d["K4"], g, d["scale_factors"], d["Xw"], e["f0_octave"], 15.0, d["valid"])
    assert np.array_equal(qv, d["q_valid"]) and np.array_equal(uv[qv > 0], d["q_uv"][qv > 0]) and np.array_equal(rad[qv > 0], d["q_radius"][qv > 0])
    fxy = np.stack([e["f1_x"], e["f1_y"]], 1)
    n1, fm1 = oracle.search_by_projection(g, fxy, e["f1_octave"], e["f1_angle"], e["f1_desc"], qv, uv, rad, mn, mx, e["f0_angle"], e["f0_desc"], 100, 0.0, True)
    assert n1 == int(d["n_frames"]) and np.array_equal(fm1, d["fm_frames"])
    n2, fm2 = oracle.search_by_projection(g, fxy, e["f1_octave"], e["f1_angle"], e["f1_desc"], qv, uv, rad * 2, mn, mx - 1, e["f0_angle"], e["f0_desc"], 100, 0.8, False)
    assert n2 == int(d["n_local"]) and np.array_equal(fm2, d["fm_local"])


def test_optimizer_golden():
    d = np.load(os.path.join(G, "optimize_small.npz"))
    T, outl, n = oracle.pose_optimization(d["po_T0"], d["po_Xw"], d["po_obs"], d["po_w"], d["po_K4"])
    assert n == int(d["po_ninl"]) and np.array_equal(outl, d["po_outlier"]) and np.allclose(T, d["po_T"], rtol=0, atol=1e-7)
    r = oracle.bundle_adjust(d["ba_poses0"], d["ba_fixed"], d["ba_intr"], d["ba_points0"], d["ba_kf"], d["ba_pt"], d["ba_uv"], d["ba_w"], True, 5, 10, True)
    assert [r["lm_iterations"], r["lm_trials"]] == d["ba_iters"].tolist()
    assert np.allclose(r["poses"], d["ba_poses"], rtol=0, atol=1e-7) and np.allclose(r["points"], d["ba_points"], rtol=0, atol=1e-6)


def test_kf_family_golden():
    """The widened rows (rest of ORBmatcher, OptimizeSim3, DBoW2 transform): the oracle reproduces tests/golden/kf_family_400x300.npz, which was
    written only after the reference's own object code (oracle/_ref) agreed with it member by member (tools/make_golden.py)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import kf_family as kff
    d = np.load(os.path.join(G, "kf_family_400x300.npz"))
    got = kff.golden_outputs(kff.OracleBackend())
    assert set(got) == set(d.files)
    for k in d.files:
        if k == "sim3":
            assert np.allclose(got[k], d[k], rtol=0, atol=1e-9), k
        else:
            assert np.array_equal(got[k], d[k]), k
    assert d["search_kf_sim3"][-1] > 10 and d["fuse_kf"].max() >= 0 and d["bow_kf_kf"][-1] > 20 and d["init"][-1] > 20 and d["sim3_inlier"][-1] >= 10


def test_sim3_chain_golden():
    """Sim3Solver inlier check (exact; the file was written only after the reference's Sim3Solver.cc object code agreed) and the fixed-scale essential-graph core."""
    d = np.load(os.path.join(G, "sim3_chain_small.npz"))
    m1, m2, p1, p2 = oracle.sim3_prepare(d["X1"], d["X2"], d["oct1"], d["oct2"], d["ls2"], d["K1"], d["K2"])
    assert np.array_equal(m1, d["max_err1"]) and np.array_equal(m2, d["max_err2"]) and np.array_equal(p1, d["p1im1"]) and np.array_equal(p2, d["p2im2"])
    inl, n = oracle.sim3_check_inliers(d["T12"], d["T21"], d["X1"], d["X2"], p1, p2, m1, m2, d["K1"], d["K2"])
    assert np.array_equal(np.packbits(inl, axis=1), d["inliers"]) and np.array_equal(n, d["n_inliers"]) and n.max() > 20
    r = oracle.optimize_pose_graph(d["pg_sim3"], d["pg_fixed"], d["pg_ei"], d["pg_ej"], d["pg_meas"], True, 20, 1e-16)
    assert [r["lm_iterations"], r["lm_trials"]] == d["pg_iters"].tolist() and np.allclose(r["sim3"], d["pg_out"], rtol=0, atol=1e-9)


def test_round2_golden():
    """tests/golden/round2_small.npz (tools/make_golden_r2.py): batched ComputeDistinctiveDescriptors, the OpenCV call sequence of ComputeSim3 through cv2, and every
    decision + the merged poses of the two-map merge chain driven by the oracle."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import map_merge as M
    d = np.load(os.path.join(G, "round2_small.npz"))
    st = d["dd_start"]
    best = [oracle.distinctive_descriptor(d["dd_flat"][st[p]:st[p + 1]]) for p in range(len(st) - 1)]
    assert np.array_equal(best, d["dd_best"]) and d["dd_best"][0] == -1
    for h in range(len(d["s3_X1"])):
        T12, T21 = M.compute_sim3(d["s3_X1"][h].T.copy(), d["s3_X2"][h].T.copy())[:2]
        assert np.array_equal(T12, d["s3_T12"][h]) and np.array_equal(T21, d["s3_T21"][h])            # same cv2 build as the generator (4.13): bit for bit
    sc = M.make_scene(seed=0, Ka=10, Kb=10, n_world=1800)
    out = M.run_merge(sc, M.Stages("oracle", sc["voc"]))
    dec = out["candidates"] + out["bow_matches"] + list(out["ransac"][-1]) + list(out["sim3_inliers"][-1]) + [out["total_matches"], out["fused"], out["essential_edges"],
                                                                                                           out["loop_connections"], out["gba"]["lm_iterations"]]
    assert dec == d["mm_decisions"].tolist()
    assert np.abs(out["poses"] - d["mm_poses"]).max() < 1e-6 * np.abs(d["mm_poses"]).max()
