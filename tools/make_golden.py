#!/usr/bin/env python3
"""Generate tests/golden/*.npz: small known-answer vectors for the hot path.

The reference repository ships no golden vectors, and its C++ cannot be built here, so these come
from the oracle.  At generation time the extractor vector is produced THREE times -- by orb_oracle.c, by the cv2-primitive
twin (the OpenCV calls the reference makes) and by the reference's own ORBextractor.cc object code (oracle/_ref, ascending-heap
tie-break) -- and the script refuses to write if any two disagree.
    python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle                      # noqa: E402
from oracle import orb_cv2, ref_build  # noqa: E402
from orbslamm_b200 import synth    # noqa: E402
from helpers import make_tracking_case  # noqa: E402

out = os.path.join(ROOT, "tests", "golden")
os.makedirs(out, exist_ok=True)

# 1. extractor: two consecutive 400x300 frames, 400 features
cam = dict(synth.TUM); cam.update(w=400, h=300, nfeatures=400, cx=200.0, cy=150.0)
case = make_tracking_case(cam, 21)
P = case["P"]
assert ref_build.build(), "oracle/_ref (the reference's own ORBextractor.cc) must be buildable to write golden vectors"
R = ref_build.RefORBextractor(400, 1.2, 8, 20, 7)
for img, feats in ((case["frames"][0], case["last"]), (case["frames"][1], case["cur"])):
    b = orb_cv2.extract(P, img)
    for k in ("x", "y", "angle", "response", "octave", "size", "desc"):
        assert np.array_equal(feats[k], b[k]), f"oracle and cv2 twin disagree on {k}"
    r = R(img)
    for k in ("x", "y", "angle", "response", "octave", "size", "desc"):
        assert np.array_equal(feats[k], r[k]), f"oracle and the reference's object code disagree on {k}"
ex = {f"f{i}_{k}": v for i, f in enumerate((case["last"], case["cur"])) for k, v in f.items()}
np.savez_compressed(os.path.join(out, "extract_400x300.npz"), frame0=case["frames"][0], frame1=case["frames"][1],
                    params=np.array([400, 8, 20, 7]), scale_factor=np.float32(1.2), **ex)

# 2. matcher: projection + both search variants on those frames
g = oracle.grid_params(*case["bounds"])
sf = np.array(list(P.scale)[:8], np.float32)
qv, uv, rad, mn, mx = oracle.project_last_frame(case["Tcw"], case["K4"], g, sf, case["Xw"], case["last"]["octave"], 15.0, case["valid"])
cur, last = case["cur"], case["last"]
fxy = np.stack([cur["x"], cur["y"]], 1)
n1, fm1 = oracle.search_by_projection(g, fxy, cur["octave"], cur["angle"], cur["desc"], qv, uv, rad, mn, mx, last["angle"], last["desc"], 100, 0.0, True)
n2, fm2 = oracle.search_by_projection(g, fxy, cur["octave"], cur["angle"], cur["desc"], qv, uv, rad * 2, mn, mx - 1, last["angle"], last["desc"], 100, 0.8, False)
np.savez_compressed(os.path.join(out, "match_400x300.npz"), Tcw=case["Tcw"], K4=case["K4"], bounds=case["bounds"], Xw=case["Xw"], valid=case["valid"],
                    scale_factors=sf, q_valid=qv, q_uv=uv, q_radius=rad, q_minl=mn, q_maxl=mx, n_frames=n1, fm_frames=fm1, n_local=n2, fm_local=fm2)

# 3. optimizer: pose optimisation on the accepted matches + a 10 KF / 200 point LocalBA
m = fm1 >= 0
Xw_m = case["Xw"][fm1[m]]; obs_m = fxy[m]; w_m = np.array(list(P.inv_sigma2)[:8], np.float32)[cur["octave"][m]]
T0 = case["Tcw"].copy(); T0[:3, 3] += np.array([0.05, -0.03, 0.08], np.float32)
Tp, outl, ninl = oracle.pose_optimization(T0, Xw_m, obs_m, w_m, case["K4"])
gb = synth.ba_graph(K=10, P=200, seed=42)
rb = oracle.bundle_adjust(gb["poses"], gb["fixed"], gb["intr"], gb["points"], gb["kf"], gb["pt"], gb["uv"], gb["inv_sigma2"], True, 5, 10, True)
# ... written only if the reference's own Optimizer.cc + g2o object code (oracle/_ref/libref_optimizer.so) agrees: exact flags, values within one float32 step
assert ref_build.optimizer_available(), "oracle/_ref/libref_optimizer.so (the reference's Optimizer.cc + g2o) must be built to write these golden vectors"
_step = lambda a, b: float(np.abs(np.asarray(a, np.float64) - b).max() / np.spacing(np.float32(np.abs(b).max())))
Tr_, outr_, nr_ = ref_build.ref_pose_optimization(T0, Xw_m, obs_m, w_m, case["K4"])
assert nr_ == ninl and np.array_equal(outr_, outl) and _step(Tp, Tr_) <= 1.0, "oracle and the reference's object code disagree on PoseOptimization"
rr_ = ref_build.ref_local_ba(gb["poses"], gb["fixed"], gb["intr"], gb["points"], gb["kf"], gb["pt"], gb["uv"], gb["inv_sigma2"])
_inw = np.zeros(len(gb["points"]), bool); _inw[gb["pt"][(gb["fixed"] != 2)[gb["kf"]]]] = True
assert _step(rb["poses"], rr_["poses"]) <= 1.0 and _step(rb["points"][_inw], rr_["points"][_inw]) <= 1.0, "oracle and the reference's object code disagree on LocalBA"
np.savez_compressed(os.path.join(out, "optimize_small.npz"), po_T0=T0, po_Xw=Xw_m, po_obs=obs_m, po_w=w_m, po_K4=case["K4"], po_T=Tp, po_outlier=outl,
                    po_ninl=ninl, ba_poses0=gb["poses"], ba_fixed=gb["fixed"], ba_intr=gb["intr"], ba_points0=gb["points"], ba_kf=gb["kf"], ba_pt=gb["pt"],
                    ba_uv=gb["uv"], ba_w=gb["inv_sigma2"], ba_poses=rb["poses"], ba_points=rb["points"], ba_chi2=rb["chi2"], ba_outlier=rb["outlier"],
                    ba_iters=np.array([rb["lm_iterations"], rb["lm_trials"]]))
for f in sorted(os.listdir(out)):
    print(f, os.path.getsize(os.path.join(out, f)))

# 4. the widened rows (rest of ORBmatcher, OptimizeSim3, DBoW2 transform) on a 400x300 case: outputs only -- the inputs are regenerated from the seeded
#    synthetic stream through the (golden-checked) oracle extractor.  Written only if the reference's object code agrees where it exists.
import kf_family as kff  # noqa: E402
gold = kff.golden_outputs(kff.OracleBackend())
assert ref_build.matcher_available() and ref_build.dbow2_available(), "oracle/_ref matcher / DBoW2 must be built to write these golden vectors"
c = kff.make_case(kff.GOLDEN_CAM, 21, distorted_bounds=True)
Scw = kff.sim3_of(c["kf"]["Tcw"], 1.21)
n, fm = ref_build.ref_search_kf_sim3(c["kf"], Scw, 10, c["pts"], c["skip"], c["held"])
assert np.array_equal(np.concatenate([fm, [n]]), gold["search_kf_sim3"])
assert np.array_equal(ref_build.ref_fuse_kf(c["kf"], 3.0, c["pts"], c["skip"], c["held"])[1], gold["fuse_kf"])
assert np.array_equal(ref_build.ref_fuse_sim3(c["kf"], Scw, 4.0, c["pts"], c["skip"], c["held"])[1], gold["fuse_sim3"])
p = kff.make_sim3_pair(kff.GOLDEN_CAM, 21)
n, m12 = ref_build.ref_search_by_sim3(p["kf1"], p["kf2"], p["s12"], p["R12"], p["t12"], 7.5, p["has1"], p["pts1"], p["has2"], p["pts2"], p["m12"])
assert np.array_equal(np.concatenate([m12, [n]]), gold["search_by_sim3"])
b = kff.make_bow_case(kff.GOLDEN_CAM, 21, True)
n, m = ref_build.ref_search_by_bow_kf_kf(b["kf1"], b["has1"], b["fv1"], b["kf2"], b["has2"], b["fv2"], 0.75, True)
assert np.array_equal(np.concatenate([m, [n]]), gold["bow_kf_kf"])
n, m = ref_build.ref_search_for_triangulation(b["kf1"], b["tri1"], b["fv1"], b["kf2"], b["tri2"], b["fv2"], b["ls2"], b["F12"], 0.6, False)
assert np.array_equal(np.concatenate([m, [n]]), gold["triangulation"])
k = c["k"]
n, m, pm = ref_build.ref_search_for_initialization(k["K4"], k["bounds"], c["sf"], k["last"], k["cur"], np.stack([k["last"]["x"], k["last"]["y"]], 1), 100, 0.9, True)
assert np.array_equal(np.concatenate([m, [n]]), gold["init"]) and np.array_equal(pm, gold["init_prev"])
import tempfile  # noqa: E402
from orbslamm_b200 import vocabulary as V  # noqa: E402
vp = os.path.join(tempfile.mkdtemp(), "voc.txt")
V.save_text(V.synthetic(6, 3, seed=9), vp)
t = ref_build.RefVocabulary(vp).transform(k["cur"]["desc"], 2)
assert np.array_equal(t["bow_ids"], gold["voc_bow_ids"]) and np.array_equal(t["bow_vals"], gold["voc_bow_vals"]) and np.array_equal(t["fv"]["items"], gold["voc_fv_items"])
np.savez_compressed(os.path.join(out, "kf_family_400x300.npz"), **gold)
for f in sorted(os.listdir(out)):
    print(f, os.path.getsize(os.path.join(out, f)))

# 5. Sim3Solver inlier check (exact; written only if the reference's Sim3Solver.cc object code agrees) and the essential-graph core (oracle output)
assert ref_build.sim3solver_available(), "oracle/_ref Sim3Solver must be built to write these golden vectors"
r3 = kff.make_sim3_ransac_case(kff.GOLDEN_CAM, 21, n_hyp=60)
m1, m2, p1, p2 = oracle.sim3_prepare(r3["X1"], r3["X2"], r3["oct1"], r3["oct2"], r3["ls2"], r3["K1"], r3["K2"])
inl, n = oracle.sim3_check_inliers(r3["T12"], r3["T21"], r3["X1"], r3["X2"], p1, p2, m1, m2, r3["K1"], r3["K2"])
ri, rn, rm1, rm2, rp1, rp2 = ref_build.ref_sim3_check_inliers(r3["X1"], r3["X2"], r3["oct1"], r3["oct2"], r3["ls2"], r3["K1"], r3["K2"], r3["T12"], r3["T21"])
assert np.array_equal(inl, ri) and np.array_equal(n, rn) and np.array_equal(m1, rm1) and np.array_equal(p2, rp2)
Sg, fxg, eig, ejg, emg, _ = kff.make_pose_graph(16, seed=2, n_loops=3)
pg = oracle.optimize_pose_graph(Sg, fxg, eig, ejg, emg, True, 20, 1e-16)
_rpg = ref_build.ref_pose_graph(Sg, fxg, eig, ejg, emg, True, 20)          # same LM iterations; fixed-scale values within the reference's own order-of-elimination spread
assert _rpg["lm_iterations"] == pg["lm_iterations"] and np.abs(_rpg["sim3"] - pg["sim3"]).max() < 2e-5 * np.abs(pg["sim3"]).max()
for _cam, _sid in ((kff.GOLDEN_CAM, 1),):
    _c = kff.make_sim3_opt_case(_cam, _sid)
    for _fix in (False, True):
        _a = (_c["init"], _c["valid"], _c["P1c"], _c["P2c"], _c["obs1"], _c["obs2"], _c["w1"], _c["w2"], _c["K1"], _c["K2"], 10.0, _fix)
        _o, _r = oracle.optimize_sim3(*_a), ref_build.ref_optimize_sim3(*_a)
        assert _o["n_in"] == _r["n_in"] and np.array_equal(_o["sim3"], _r["sim3"]), "oracle and the reference's object code disagree on OptimizeSim3"
np.savez_compressed(os.path.join(out, "sim3_chain_small.npz"), X1=r3["X1"], X2=r3["X2"], oct1=r3["oct1"], oct2=r3["oct2"], ls2=r3["ls2"], K1=r3["K1"], K2=r3["K2"],
                    T12=r3["T12"], T21=r3["T21"], max_err1=m1, max_err2=m2, p1im1=p1, p2im2=p2, inliers=np.packbits(inl, axis=1), n_inliers=n,
                    pg_sim3=Sg, pg_fixed=fxg, pg_ei=eig, pg_ej=ejg, pg_meas=emg, pg_out=pg["sim3"], pg_iters=np.array([pg["lm_iterations"], pg["lm_trials"]]))
for f in sorted(os.listdir(out)):
    print(f, os.path.getsize(os.path.join(out, f)))
