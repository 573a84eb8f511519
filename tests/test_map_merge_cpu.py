"""configs[3] (two maps -> Sim3 merge -> essential graph -> global BA) composed on the CPU oracle: the chain of MultiMapper.cc:209-662 runs through, every acceptance
gate of the reference is passed, and the merged map lands on the scene's ground truth.  (The CUDA side of the same composition: tests/test_map_merge_gpu.py.)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import map_merge as M


def test_merge_chain_on_the_oracle_recovers_the_ground_truth():
    sc = M.make_scene(seed=0, Ka=10, Kb=10, n_world=1800)
    out = M.run_merge(sc, M.Stages("oracle", sc["voc"]))
    assert max(out["bow_matches"]) >= 15 and out["total_matches"] >= 40 and out["sim3_inliers"][-1][2] >= 20
    assert out["fused"] > 0 and out["loop_connections"] > 0
    # before the merge the keyframes of map B are in another frame and 1.3 x smaller; afterwards they are within centimetres of the truth
    before = M.centre_error_vs_truth(sc, [k["Tcw"] for k in sc["A"]["kfs"]] + [k["Tcw"] for k in sc["B"]["kfs"]])
    after = M.centre_error_vs_truth(sc, out["poses"])
    assert before > 1.0 and after < 0.08, (before, after)
    # Sim3 of the merge: scale of map B relative to map A
    assert abs(1.0 / out["merged"]["Scm"][7] - 1.0 / sc["s_b"]) < 0.02 or abs(out["merged"]["Scm"][7] - 1.0 / sc["s_b"]) < 0.02


def test_shard_by_origin_partitions_the_merged_graph():
    from orbslamm_b200 import sharding
    sc = M.make_scene(seed=1, Ka=8, Kb=8, n_world=1200)
    g = M.run_merge(sc, M.Stages("oracle", sc["voc"]))["gba_graph"]
    sh = [sharding.shard_graph_by_owner(g, g["origin"], r) for r in range(2)]
    assert sum(len(s["points"]) for s in sh) == len(g["points"]) and sum(len(s["kf"]) for s in sh) == len(g["kf"])
    for r, s in enumerate(sh):
        assert np.all(g["origin"][s["local_points"]] == r)
        assert np.array_equal(s["local_points"][s["pt"]], g["pt"][s["local_edges"]])
        assert np.array_equal(s["points"], g["points"][s["local_points"]])
