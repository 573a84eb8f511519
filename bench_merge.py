"""Map-merge leg of bench.py (BASELINE.json configs[3]: "Multi-robot 2 streams, Sim3 map-merge + GlobalBA, 2 GPUs").

Two synthetic robot maps (tests/map_merge.py: feature-level scene, 2 x 24 keyframes) are merged the way MultiMapper does it (M/src/MultiMapper.cc:209-662): vocabulary
transform, SearchByBoW, Sim3Solver RANSAC with ONE batched CheckInliers launch, SearchBySim3, OptimizeSim3, SearchByProjection(KF, Scw), Fuse for every keyframe of the
merged-in map, MMOptimizeEssentialGraph, then MMGlobalBundleAdjustemnt(20 iterations, not robust).  merge_ms = wall time of the chain up to the global BA (CUDA stages +
the numpy host glue that stands in for the reference's map bookkeeping); gba = LM iterations/s of the global BA.  At N > 1 the map points of the merged map stay on
the GPUs of the robot whose map they come from (half of the GPUs each), keyframes replicated, one all-reduce of the reduced pose system per LM trial; the sharded result
is checked against the single-GPU one before timing."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
KA = KB = 24
N_WORLD = 4200


def _scene():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import map_merge as M
    return M, M.make_scene(seed=0, Ka=KA, Kb=KB, n_world=N_WORLD)


def _record(out, gba, extra):
    g = out["gba_graph"]
    rec = {"metric": "two-map Sim3 merge ms; MMGlobalBA LM iters/s", "merge_ms": round(sum(v for k, v in out["stage_ms"].items() if k != "MMGlobalBundleAdjustemnt"), 2),
           "stage_ms": {k: round(v, 2) for k, v in out["stage_ms"].items()},
           "config": {"workload": f"2 synthetic robot maps ({KA} + {KB} keyframes, ~1200 features each, KITTI shape), Sim3 map merge + essential graph + global BA of the merged map "
                                  f"({len(g['poses'])} keyframes / {len(g['points'])} points / {len(g['kf'])} observations, 20 LM iterations)",
                      "decisions": {"candidates": out["candidates"], "bow_matches": out["bow_matches"], "sim3_inliers": out["sim3_inliers"][-1][2],
                                    "total_matches": out["total_matches"], "fused": out["fused"], "essential_edges": out["essential_edges"]}},
           "gba": gba}
    rec.update(extra)
    return rec


def bench_merge(args, rank, world):
    import torch
    import orbslamm_b200 as ob
    from orbslamm_b200 import sharding
    dev = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(dev)
    M, sc = _scene()
    st = M.Stages("cuda", sc["voc"])
    for _ in range(2):
        M.run_merge(sc, st)                           # warm-up: handles, staging pools and the pose-graph buffers reach their final sizes after two calls
    out = M.run_merge(sc, st)
    g, single = out["gba_graph"], out["gba"]
    opt = ob.Optimizer(device=dev)
    parity = None
    gr = g
    if world > 1:
        uid = [ob.Optimizer.comm_unique_id() if rank == 0 else None]
        torch.distributed.broadcast_object_list(uid, src=0)
        opt.comm_init(world, rank, uid[0])
        gr = sharding.shard_graph_by_owner(g, M.owner_by_origin(g["origin"], world), rank)
    run = lambda: opt.BundleAdjustment(gr["poses"], gr["fixed"], gr["intr"], gr["points"], gr["kf"], gr["pt"], gr["uv"], gr["inv_sigma2"], 20, False)
    r = run()
    if world > 1:
        rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
        ok = r["lm_iterations"] == single["lm_iterations"] and rel(r["poses"], single["poses"]) < 1e-6 and rel(r["points"], single["points"][gr["local_points"]]) < 1e-6
        t = torch.tensor([1 if ok else 0], device="cuda", dtype=torch.int32)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MIN)
        parity = bool(int(t.item()))
        assert parity, "global BA sharded by origin map differs from the single-GPU result"
    for _ in range(2):
        run()
    steps = max(5, min(args.steps, 20))
    loop_s = 0.0; iters = 0
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    for _ in range(steps):
        r = run()
        loop_s += opt.last_ba_timing()["lm_loop_s"]; iters += r["lm_iterations"]
    if world > 1:
        t = torch.tensor([loop_s], device="cuda", dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        loop_s = float(t.item())
    gba = {"value": round(iters / loop_s, 2), "unit": "LM iterations/s", "ms_per_lm_iteration": round(loop_s * 1e3 / max(iters, 1), 4), "steps": steps,
           "lm_iterations": int(iters), "n_gpus": world, "scaling": "strong" if world > 1 else "weak",
           "parallelism": (f"map points on the GPUs of their origin map ({world // 2} + {world - world // 2}), keyframes replicated, one all-reduce of the reduced system per LM trial"
                           if world > 1 else "1 GPU")}
    if parity is not None:
        gba["sharded_parity"] = parity
    return _record(out, gba, {"n_gpus": world, "centre_error_vs_truth_m": round(M.centre_error_vs_truth(sc, out["poses"]), 4)})


def reference_merge(args):
    """the same chain on the CPU oracle (1 thread)"""
    M, sc = _scene()
    st = M.Stages("oracle", sc["voc"])
    t0 = time.perf_counter()
    out = M.run_merge(sc, st)
    wall = time.perf_counter() - t0
    ba_ms = out["stage_ms"]["MMGlobalBundleAdjustemnt"]
    gba = {"value": round(out["gba"]["lm_iterations"] / (ba_ms * 1e-3), 2), "unit": "LM iterations/s", "lm_iterations": int(out["gba"]["lm_iterations"]), "cores": 1, "kind": "port"}
    return _record(out, gba, {"impl": "reference", "wall_s": round(wall, 2)})
