"""Profiling aid: one 500 KF / 50k point LocalBA with ORBS_RS_TRACE set -> per-task timeline of the persistent reduced-system solver
(k_rs_solve): task durations by type and the critical path through the elimination levels."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
out = os.path.join(ROOT, "gpurun_out", "rs_trace.txt")
os.environ["ORBS_RS_TRACE"] = out
import orbslamm_b200 as ob  # noqa: E402
from orbslamm_b200 import synth  # noqa: E402

K, P = int(os.environ.get("BA_K", 500)), int(os.environ.get("BA_P", 50000))
g = synth.ba_graph(K=K, P=P, seed=42)
opt = ob.Optimizer(device=0)
for _ in range(2):
    r = opt.LocalBundleAdjustment(g["poses"], g["fixed"], g["intr"], g["points"], g["kf"], g["pt"], g["uv"], g["inv_sigma2"])
t = np.loadtxt(out, dtype=np.int64)
t0 = t[:, 5].min()
names = {0: "factor", 1: "forward", 2: "backward"}
print(f"tasks {len(t)}, span {(t[:, 7].max() - t0) / 1e3:.1f} us, SMs used {len(set(t[:, 8]))}")
for ty in (0, 1, 2):
    m = t[:, 1] == ty
    diag = m & (t[:, 2] == t[:, 3])
    if ty == 0:
        for nm, mm in (("factor diag", diag), ("factor panel", m & ~diag)):
            d = t[mm]
            print(f"{nm:14s} n={len(d):4d}  wait+gemm {np.mean(d[:, 6] - d[:, 5]) / 1e3:6.2f} us  tail {np.mean(d[:, 7] - d[:, 6]) / 1e3:6.2f} us  "
                  f"first start {(d[:, 5].min() - t0) / 1e3:7.1f}  last end {(d[:, 7].max() - t0) / 1e3:7.1f}")
    else:
        d = t[m]
        print(f"{names[ty]:14s} n={len(d):4d}  total {np.mean(d[:, 7] - d[:, 5]) / 1e3:6.2f} us  first start {(d[:, 5].min() - t0) / 1e3:7.1f}  last end {(d[:, 7].max() - t0) / 1e3:7.1f}")
# diagonal tasks in time order: the elimination chain
d = t[(t[:, 1] == 0) & (t[:, 2] == t[:, 3])]
d = d[np.argsort(d[:, 7])]
print("last 12 diagonal tasks (i, ndep, start, deps_done, end in us):")
for row in d[-12:]:
    print(f"  i={row[2]:3d} ndep={row[4]:3d}  {(row[5] - t0) / 1e3:7.1f} {(row[6] - t0) / 1e3:7.1f} {(row[7] - t0) / 1e3:7.1f}")
p = t[(t[:, 1] == 0) & (t[:, 2] != t[:, 3])]
p = p[np.argsort(p[:, 7])]
print("last 8 panel tasks (i, j, ndep, start, deps_done, end):")
for row in p[-8:]:
    print(f"  ({row[2]:3d},{row[3]:3d}) ndep={row[4]:3d}  {(row[5] - t0) / 1e3:7.1f} {(row[6] - t0) / 1e3:7.1f} {(row[7] - t0) / 1e3:7.1f}")
