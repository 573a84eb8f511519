"""GPU parity of the Optimizer rows against THE REFERENCE'S OWN OBJECT CODE (oracle/_ref/libref_optimizer.so: Optimizer.cc + Converter.cc + the vendored g2o,
compiled unmodified against the Eigen stand-in; it travels prebuilt to the GPU box).  The CUDA path is called through the C-ABI; the reference runs on the host.
Tolerance: north_star's 1e-5 relative for poses / points, outlier and inlier sets exact."""
import numpy as np
import pytest

import oracle
from oracle import ref_build
from orbslamm_b200 import synth

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_build.optimizer_available(), reason="oracle/_ref/libref_optimizer.so did not travel")]

RTOL = 1e-5


def _rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - b).max() / max(np.abs(b).max(), 1e-12))


@pytest.mark.parametrize("K,P,seed", [(10, 200, 42), (40, 2500, 42), (100, 10000, 3)])
def test_local_ba_matches_reference_object_code(lib, K, P, seed):
    """orbo_bundle_adjust (two-stage schedule) vs Optimizer::LocalBundleAdjustment run by the reference's object code on the same graph: poses and in-window
    points to 1e-5 (in fact ~1e-7: one float32 step), and the observations the reference erases are exactly the ones flagged as outliers."""
    import orbslamm_b200 as ob
    g = synth.ba_graph(K=K, P=P, seed=seed)
    got = ob.Optimizer().LocalBundleAdjustment(g["poses"], g["fixed"], g["intr"], g["points"], g["kf"], g["pt"], g["uv"], g["inv_sigma2"])
    ref = ref_build.ref_local_ba(g["poses"], g["fixed"], g["intr"], g["points"], g["kf"], g["pt"], g["uv"], g["inv_sigma2"])
    assert got["lm_iterations"] == 15
    assert _rel(got["poses"], ref["poses"]) < RTOL
    in_window = np.zeros(P, bool); in_window[g["pt"][(g["fixed"] != 2)[g["kf"]]]] = True      # the reference's graph holds the points a LOCAL keyframe sees
    assert _rel(got["points"][in_window], ref["points"][in_window]) < RTOL
    assert _rel(got["poses"], ref["poses"]) < 5e-7 and _rel(got["points"][in_window], ref["points"][in_window]) < 5e-7     # one float32 step of the largest entry
    erased = np.bincount(g["pt"], minlength=P) - ref["nobs"]
    e_in = in_window[g["pt"]]
    assert np.array_equal(np.bincount(g["pt"][e_in], weights=got["outlier"][e_in], minlength=P).astype(int)[in_window], erased[in_window])


@pytest.mark.parametrize("robust", [True, False])
def test_global_ba_matches_reference_object_code(lib, robust):
    """orbo_bundle_adjust (single stage, 20 iterations) vs Optimizer::BundleAdjustment: the thHuber2D = sqrt(5.99) kernel of that function (LocalBA uses 5.991)."""
    import orbslamm_b200 as ob
    g = synth.ba_graph(K=14, P=400, seed=3)
    g["fixed"] = g["fixed"].copy(); g["fixed"][1] = 0
    got = ob.Optimizer().BundleAdjustment(g["poses"], g["fixed"], g["intr"], g["points"], g["kf"], g["pt"], g["uv"], g["inv_sigma2"], 20, robust)
    ref = ref_build.ref_bundle_adjust(g["poses"], g["fixed"], g["intr"], g["points"], g["kf"], g["pt"], g["uv"], g["inv_sigma2"], 20, robust)
    assert _rel(got["poses"], ref["poses"]) < RTOL and _rel(got["points"], ref["points"]) < RTOL


def test_pose_optimization_matches_reference_object_code(lib):
    import orbslamm_b200 as ob
    from helpers import slab
    g = synth.ba_graph(K=24, P=1500, seed=0, min_obs=5, max_obs=12)
    K4 = np.array(g["intr"], np.float32)
    cases = []
    for k in range(2, 12):
        m = g["kf"] == k
        cases.append(dict(T=g["poses"][k], Xw=g["gt_points"][g["pt"][m]].astype(np.float32), obs=g["uv"][m], w=g["inv_sigma2"][m]))
    S = max(len(c["Xw"]) for c in cases)
    T, outl, ninl = ob.Optimizer().PoseOptimization(np.stack([c["T"] for c in cases]), K4, slab([c["Xw"] for c in cases], S, np.float32, (3,)),
                                                    slab([c["obs"] for c in cases], S, np.float32, (2,)), slab([c["w"] for c in cases], S, np.float32),
                                                    np.array([len(c["Xw"]) for c in cases], np.int32))
    for i, c in enumerate(cases):
        Tr, outr, nr = ref_build.ref_pose_optimization(c["T"], c["Xw"], c["obs"], c["w"], K4)
        assert ninl[i] == nr and np.array_equal(outl[i, :len(outr)], outr) and _rel(T[i], Tr) < RTOL


@pytest.mark.parametrize("fix_scale", [False, True])
def test_optimize_sim3_matches_reference_object_code(lib, fix_scale):
    import orbslamm_b200 as ob
    import kf_family as kff
    from helpers import slab
    cases = [kff.make_sim3_opt_case(synth.TUM, 1), kff.make_sim3_opt_case(synth.KITTI, 2)]
    W = max(len(c["valid"]) for c in cases)
    S, inl, nin, st = ob.Optimizer().OptimizeSim3(np.stack([c["init"] for c in cases]), slab([c["valid"] for c in cases], W, np.uint8),
                                                  slab([c["P1c"] for c in cases], W, np.float32, (3,)), slab([c["P2c"] for c in cases], W, np.float32, (3,)),
                                                  slab([c["obs1"] for c in cases], W, np.float32, (2,)), slab([c["obs2"] for c in cases], W, np.float32, (2,)),
                                                  slab([c["w1"] for c in cases], W, np.float32), slab([c["w2"] for c in cases], W, np.float32),
                                                  np.stack([c["K1"] for c in cases]), np.stack([c["K2"] for c in cases]), [len(c["valid"]) for c in cases], 10.0, fix_scale)
    for k, c in enumerate(cases):
        r = ref_build.ref_optimize_sim3(c["init"], c["valid"], c["P1c"], c["P2c"], c["obs1"], c["obs2"], c["w1"], c["w2"], c["K1"], c["K2"], 10.0, fix_scale)
        n = len(c["valid"])
        assert nin[k] == r["n_in"] and np.array_equal(inl[k, :n], r["inlier"])
        assert np.abs(S[k] - r["sim3"]).max() < RTOL * np.abs(r["sim3"]).max()
