import sys, time, os
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,R); sys.path.insert(0,os.path.join(R,'tests'))
import numpy as np, orbslamm_b200 as ob
import kf_family as kff
S, fixed, ei, ej, meas, true = kff.make_pose_graph(K=60, seed=0)
opt = ob.Optimizer()
for i in range(6):
    t0=time.perf_counter(); r = opt.OptimizePoseGraph(S, fixed, ei, ej, meas, False, 20); w=time.perf_counter()-t0
    print(f"K=60 free-scale: {w*1e3:.2f} ms  its {r['lm_iterations']} trials {r['lm_trials']} chol_fail {r['chol_failures']}")
import map_merge as M
sc = M.make_scene(seed=0, Ka=24, Kb=24, n_world=4200)
st = M.Stages("cuda", sc["voc"])
for i in range(3):
    out = M.run_merge(sc, st)
    print({k: round(v,2) for k,v in out["stage_ms"].items()})
