"""Driver for compute-sanitizer over the kernels that changed in round 2: the extractor (new FAST NMS pass, descriptor pattern fetch), the projection search
(new candidate kernel), PoseOptimization (one pass per trial), a small LocalBA (per-target Schur assembly, the persistent tile solver with the blocked DMMA
factorisation, device-resident LM control), the batched distinctive-descriptor kernel and the device ComputeSim3.  Every result is checked against the oracle.
    compute-sanitizer --tool memcheck  python tools/sanitize_round2.py
    compute-sanitizer --tool racecheck python tools/sanitize_round2.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
import orbslamm_b200 as ob  # noqa: E402
from orbslamm_b200 import synth  # noqa: E402
from helpers import make_tracking_case  # noqa: E402
import kf_family as kff  # noqa: E402

bad = []
cam = dict(synth.TUM); cam.update(w=400, h=300, nfeatures=400, cx=200.0, cy=150.0)
case = make_tracking_case(cam, 21)
ex = ob.ORBextractor(400, 1.2, 8, 20, 7)
got = ex(case["frames"][1]); ref = case["cur"]
for k in ("x", "y", "angle", "response", "octave", "size", "desc"):
    if not np.array_equal(got[k], ref[k]): bad.append("extract " + k)
m = ob.ORBmatcher(0.9, True)
sf = np.array(list(case["P"].scale)[:8], np.float32)
n = len(case["last"]["x"])
qv, uv, rad, mn, mx = m.project_last_frame(case["Tcw"][None], case["K4"], case["bounds"], sf, case["Xw"][None], case["last"]["octave"][None], np.array([n], np.int32), 15.0,
                                           case["valid"][None])
nm, fm = m.SearchByProjection(case["bounds"], np.stack([got["x"], got["y"]], 1)[None], got["octave"][None], got["angle"][None], got["desc"][None],
                              np.array([len(got["x"])], np.int32), qv, uv, rad, mn, mx, case["last"]["angle"][None], case["last"]["desc"][None], np.array([n], np.int32), 100)
g = oracle.grid_params(*case["bounds"])
r = oracle.project_last_frame(case["Tcw"], case["K4"], g, sf, case["Xw"], case["last"]["octave"], 15.0, case["valid"])
n_ref, fm_ref = oracle.search_by_projection(g, np.stack([ref["x"], ref["y"]], 1), ref["octave"], ref["angle"], ref["desc"], r[0], r[1], r[2], r[3], r[4],
                                            case["last"]["angle"], case["last"]["desc"], 100, 0.0, True)
if int(nm[0]) != n_ref or not np.array_equal(fm[0], fm_ref): bad.append("search_by_projection")
opt = ob.Optimizer()
inv_s2 = np.array(list(case["P"].inv_sigma2)[:8], np.float32)
sel = fm_ref >= 0
fxy = np.stack([ref["x"], ref["y"]], 1)[sel]
T0 = case["Tcw"].copy(); T0[:3, 3] += np.array([0.03, -0.02, 0.05], np.float32)
Tp, outl, ninl = opt.PoseOptimization(T0[None], case["K4"], case["Xw"][fm_ref[sel]][None], fxy[None], inv_s2[ref["octave"][sel]][None], np.array([sel.sum()], np.int32))
Tr, outr, nr = oracle.pose_optimization(T0, case["Xw"][fm_ref[sel]], fxy, inv_s2[ref["octave"][sel]], case["K4"])
if int(ninl[0]) != nr or not np.array_equal(outl[0], outr) or np.abs(Tp[0] - Tr).max() > 1e-5 * np.abs(Tr).max(): bad.append("pose_optimization")
g2 = synth.ba_graph(K=24, P=700, seed=42)
ba = opt.LocalBundleAdjustment(g2["poses"], g2["fixed"], g2["intr"], g2["points"], g2["kf"], g2["pt"], g2["uv"], g2["inv_sigma2"])
bar = oracle.bundle_adjust(g2["poses"], g2["fixed"], g2["intr"], g2["points"], g2["kf"], g2["pt"], g2["uv"], g2["inv_sigma2"], True, 5, 10, True)
if ba["lm_iterations"] != bar["lm_iterations"] or np.abs(ba["poses"] - bar["poses"]).max() > 1e-5 * np.abs(bar["poses"]).max(): bad.append("local_ba")
rng = np.random.default_rng(3)
lists = [rng.integers(0, 256, (int(k), 32), dtype=np.uint8) for k in (0, 1, 2, 5, 33, 40, 7)]
if not np.array_equal(m.ComputeDistinctiveDescriptors(lists), [oracle.distinctive_descriptor(d) for d in lists]): bad.append("distinctive")
rc = kff.make_sim3_ransac_case(kff.GOLDEN_CAM, 21, n_hyp=8)
tri = rng.integers(0, len(rc["X1"]), (40, 3))
T12, T21, R, t, s = opt.Sim3Compute(rc["X1"][tri], rc["X2"][tri])
if not np.isfinite(T12).all(): bad.append("sim3_compute")
print("sanitize driver done: mismatches vs oracle:", bad, "| keypoints", len(got["x"]), "matches", n_ref, "BA iterations", ba["lm_iterations"])
sys.exit(1 if bad else 0)
