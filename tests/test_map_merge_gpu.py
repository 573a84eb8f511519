"""BASELINE.json configs[3] on the GPU: two synthetic maps -> DBoW2 transform -> SearchByBoW -> Sim3Solver RANSAC (batched CheckInliers) -> SearchBySim3 -> OptimizeSim3
-> SearchByProjection(KF, Scw) -> Fuse -> MMOptimizeEssentialGraph -> MMGlobalBundleAdjustemnt(20), composed as MultiMapper.cc:209-662 composes them
(tests/map_merge.py).  Integer stages are compared exactly, fp64 stages to 1e-5 with identical inlier flags / LM iteration counts."""
import os
import subprocess
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import map_merge as M

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def scene():
    return M.make_scene(seed=0, Ka=10, Kb=10, n_world=1800)


def test_every_stage_of_the_merge_equals_the_oracle(lib, scene):
    """the oracle drives the chain; at every stage the CUDA entry point runs on the same inputs and must agree (asserts inside run_merge)"""
    out = M.run_merge(scene, M.Stages("oracle", scene["voc"]), check=M.Stages("cuda", scene["voc"]))
    assert out["total_matches"] >= 40 and out["fused"] > 0


def test_merge_end_to_end_on_cuda(lib, scene):
    """the CUDA results are fed forward through the whole chain: same decisions as the oracle-driven chain, merged map equal to 1e-4, on the ground truth"""
    oc = M.run_merge(scene, M.Stages("cuda", scene["voc"]))
    oo = M.run_merge(scene, M.Stages("oracle", scene["voc"]))
    for k in ("candidates", "bow_matches", "ransac", "sim3_inliers", "total_matches", "fused", "essential_edges", "loop_connections"):
        assert oc[k] == oo[k], (k, oc[k], oo[k])
    rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
    assert oc["gba"]["lm_iterations"] == oo["gba"]["lm_iterations"]
    assert rel(oc["poses"], oo["poses"]) < 1e-4 and rel(oc["points"], oo["points"]) < 1e-4
    assert M.centre_error_vs_truth(scene, oc["poses"]) < 0.08


def test_global_ba_sharded_by_origin_map_equals_single_gpu(lib):
    if lib.orbs_device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29617",
           os.path.join(ROOT, "tests", "mgpu_merge_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "MGPU_MERGE_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
