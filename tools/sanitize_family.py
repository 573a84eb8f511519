"""Small driver for compute-sanitizer: every kernel of the widened rows once, on the 400x300 golden case (plus the Sim3Solver check and the grid calls).
    compute-sanitizer --tool memcheck python tools/sanitize_family.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import kf_family as kff  # noqa: E402
import orbslamm_b200 as ob  # noqa: E402

d = np.load(os.path.join(ROOT, "tests", "golden", "kf_family_400x300.npz"))
got = kff.golden_outputs(kff.CudaBackend())
bad = [k for k in d.files if k != "sim3" and not np.array_equal(got[k], d[k])]
r = kff.make_sim3_ransac_case(kff.GOLDEN_CAM, 21, n_hyp=40)
o = ob.Optimizer()
m1, p1 = o.Sim3Prepare(r["X1"], r["oct1"], r["ls2"], r["K1"]); m2, p2 = o.Sim3Prepare(r["X2"], r["oct2"], r["ls2"], r["K2"])
inl, n = o.Sim3CheckInliers(r["T12"], r["T21"], r["X1"], r["X2"], p1, p2, m1, m2, r["K1"], r["K2"])
c = kff.make_case(kff.GOLDEN_CAM, 21)
m = ob.ORBmatcher()
kf = c["kf"]
cs, ci = m.AssignFeaturesToGrid(kf["grid_bounds4"], kf["xy"][None], [len(kf["xy"])])
q = np.concatenate([kf["xy"][:50], np.full((50, 1), 30.0, np.float32)], 1)[None]
idx, cnt = m.GetFeaturesInArea(kf["grid_bounds4"], kf["xy"][None], kf["octave"][None], [len(kf["xy"])], q, None, None, [50], cap=32)
print("sanitize driver done: mismatches vs golden:", bad, "sim3 inliers", int(n.max()), "area hits", int(cnt.max()))
sys.exit(1 if bad else 0)
