"""CUDA path (through the C-ABI) against the committed golden vectors."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_extractor_and_matcher_golden(lib):
    import orbslamm_b200 as ob
    e = np.load(os.path.join(G, "extract_400x300.npz")); d = np.load(os.path.join(G, "match_400x300.npz"))
    ex = ob.ORBextractor(int(e["params"][0]), float(e["scale_factor"]), int(e["params"][1]), int(e["params"][2]), int(e["params"][3]))
    outs = ex.extract_batch(np.stack([e["frame0"], e["frame1"]]))
    for i in (0, 1):
        for k in ("x", "y", "angle", "response", "octave", "size", "desc"):
            assert np.array_equal(outs[i][k], e[f"f{i}_{k}"]), (i, k)
    # public pyramid (mvImagePyramid) incl. the 19-px reflect-101 border
    lvl = ex.pyramid_level(1, 2, 400, 300, border=True)
    assert lvl.shape == (208 + 38, 278 + 38) and np.array_equal(lvl[19:-19, 19:-19], ex.pyramid_level(1, 2, 400, 300))
    assert np.array_equal(lvl[0, 19:-19], lvl[38, 19:-19]) and np.array_equal(lvl[19:-19, 0], lvl[19:-19, 38])
    m = ob.ORBmatcher(0.8, True)
    n = len(e["f0_x"])
    qv, uv, rad, mn, mx = m.project_last_frame(d["Tcw"][None], d["K4"], d["bounds"], d["scale_factors"], d["Xw"][None], e["f0_octave"][None],
                                               np.array([n], np.int32), 15.0, d["valid"][None])
    assert np.array_equal(qv[0], d["q_valid"]) and np.array_equal(uv[0][qv[0] > 0], d["q_uv"][d["q_valid"] > 0])
    cur = outs[1]
    args = (d["bounds"], np.stack([cur["x"], cur["y"]], 1)[None], cur["octave"][None], cur["angle"][None], cur["desc"][None], np.array([len(cur["x"])], np.int32))
    nm, fm = m.SearchByProjection(*args, qv, uv, rad, mn, mx, e["f0_angle"][None], e["f0_desc"][None], np.array([n], np.int32), 100)
    assert nm[0] == int(d["n_frames"]) and np.array_equal(fm[0], d["fm_frames"])
    nm, fm = m.SearchByProjection(*args, qv, uv, rad * 2, mn, mx - 1, e["f0_angle"][None], e["f0_desc"][None], np.array([n], np.int32), 100, use_ratio=True)
    assert nm[0] == int(d["n_local"]) and np.array_equal(fm[0], d["fm_local"])


def test_optimizer_golden(lib):
    import orbslamm_b200 as ob
    d = np.load(os.path.join(G, "optimize_small.npz"))
    opt = ob.Optimizer()
    T, outl, n = opt.PoseOptimization(d["po_T0"][None], d["po_K4"], d["po_Xw"][None], d["po_obs"][None], d["po_w"][None], np.array([len(d["po_w"])], np.int32))
    assert n[0] == int(d["po_ninl"]) and np.array_equal(outl[0], d["po_outlier"])
    assert np.abs(T[0] - d["po_T"]).max() < 1e-5 * np.abs(d["po_T"]).max()
    r = opt.LocalBundleAdjustment(d["ba_poses0"], d["ba_fixed"], d["ba_intr"], d["ba_points0"], d["ba_kf"], d["ba_pt"], d["ba_uv"], d["ba_w"])
    assert [r["lm_iterations"], r["lm_trials"]] == d["ba_iters"].tolist()
    assert np.abs(r["poses"] - d["ba_poses"]).max() < 1e-5 * np.abs(d["ba_poses"]).max()
    assert np.abs(r["points"] - d["ba_points"]).max() < 1e-5 * np.abs(d["ba_points"]).max()


def test_kf_family_golden_gpu(lib):
    """CUDA path against tests/golden/kf_family_400x300.npz (rest of ORBmatcher, OptimizeSim3, DBoW2 transform): exact, Sim3 to 1e-5 relative."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import kf_family as kff
    d = np.load(os.path.join(G, "kf_family_400x300.npz"))
    got = kff.golden_outputs(kff.CudaBackend())
    assert set(got) == set(d.files)
    for k in d.files:
        if k == "sim3":
            assert np.abs(got[k] - d[k]).max() < 1e-5 * np.abs(d[k]).max(), k
        else:
            assert np.array_equal(got[k], d[k]), k


def test_sim3_chain_golden_gpu(lib):
    """CUDA path against tests/golden/sim3_chain_small.npz: the Sim3Solver inlier check exactly, the fixed-scale essential-graph core to 1e-5 relative."""
    import orbslamm_b200 as ob
    d = np.load(os.path.join(G, "sim3_chain_small.npz"))
    o = ob.Optimizer()
    m1, p1 = o.Sim3Prepare(d["X1"], d["oct1"], d["ls2"], d["K1"]); m2, p2 = o.Sim3Prepare(d["X2"], d["oct2"], d["ls2"], d["K2"])
    assert np.array_equal(m1, d["max_err1"]) and np.array_equal(m2, d["max_err2"]) and np.array_equal(p1, d["p1im1"]) and np.array_equal(p2, d["p2im2"])
    inl, n = o.Sim3CheckInliers(d["T12"], d["T21"], d["X1"], d["X2"], p1, p2, m1, m2, d["K1"], d["K2"])
    assert np.array_equal(np.packbits(inl, axis=1), d["inliers"]) and np.array_equal(n, d["n_inliers"])
    r = o.OptimizePoseGraph(d["pg_sim3"], d["pg_fixed"], d["pg_ei"], d["pg_ej"], d["pg_meas"], True, 20, 1e-16)
    assert np.abs(r["sim3"] - d["pg_out"]).max() < 1e-5 * np.abs(d["pg_out"]).max()


def test_round2_golden_gpu(lib):
    """CUDA path against tests/golden/round2_small.npz: distinctive descriptors exactly, ComputeSim3 to the float tolerance of its header comment, the two-map merge
    chain with identical decisions and the merged poses to 1e-4."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import map_merge as M
    import orbslamm_b200 as ob
    d = np.load(os.path.join(G, "round2_small.npz"))
    st = d["dd_start"]
    got = ob.ORBmatcher(0.6, True).ComputeDistinctiveDescriptors([d["dd_flat"][st[p]:st[p + 1]] for p in range(len(st) - 1)])
    assert np.array_equal(got, d["dd_best"])
    T12, T21, _, _, _ = ob.Optimizer().Sim3Compute(d["s3_X1"], d["s3_X2"])
    for h in range(len(T12)):
        assert np.abs(T12[h] - d["s3_T12"][h]).max() < 1e-4 * max(np.abs(d["s3_T12"][h]).max(), 1.0)
        assert np.abs(T21[h] - d["s3_T21"][h]).max() < 1e-4 * max(np.abs(d["s3_T21"][h]).max(), 1.0)
    sc = M.make_scene(seed=0, Ka=10, Kb=10, n_world=1800)
    out = M.run_merge(sc, M.Stages("cuda", sc["voc"]))
    dec = out["candidates"] + out["bow_matches"] + list(out["ransac"][-1]) + list(out["sim3_inliers"][-1]) + [out["total_matches"], out["fused"], out["essential_edges"],
                                                                                                           out["loop_connections"], out["gba"]["lm_iterations"]]
    assert dec == d["mm_decisions"].tolist()
    assert np.abs(out["poses"] - d["mm_poses"]).max() < 1e-4 * np.abs(d["mm_poses"]).max()
