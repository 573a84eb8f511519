import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, orbslamm_b200 as ob
from orbslamm_b200 import synth
g = synth.ba_graph(K=500, P=50000, seed=42)
opt = ob.Optimizer()
for i in range(6):
    t0=time.perf_counter()
    r = opt.LocalBundleAdjustment(g["poses"], g["fixed"], g["intr"], g["points"], g["kf"], g["pt"], g["uv"], g["inv_sigma2"])
    w=time.perf_counter()-t0
    tm = opt.last_ba_timing()
    print(f"wall {w*1e3:.2f} ms  total {tm['total_s']*1e3:.2f}  setup {tm['setup_s']*1e3:.2f}  loop {tm['lm_loop_s']*1e3:.2f}  tail {(tm['total_s']-tm['setup_s']-tm['lm_loop_s'])*1e3:.2f}")
