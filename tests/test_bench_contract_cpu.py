"""bench.py contract pieces that need no GPU: both arms print the same static `config` per record (the reference arm runs on this arm's config),
and the default / reference command lines parse."""
import argparse
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _args(**kw):
    a = argparse.Namespace(gpus=1, steps=40, warmup=3, streams=128, instances=2, mode="pipeline", e2e_workers=4, workload="all", camera="kitti",
                           impl="ours", cpu_frames=100, ba_kf=500, ba_pts=50000)
    for k, v in kw.items():
        setattr(a, k, v)
    return a


def test_frontend_config_is_static_and_shared_by_both_arms():
    bench = importlib.import_module("bench")
    for n in (1, 2, 8):
        ours = bench.frontend_config(_args(gpus=n), n)                 # our arm: world size from torchrun
        ref = bench.frontend_config(_args(gpus=n, impl="reference"), n)   # reference arm: --gpus N, rank 0 only
        assert ours == ref
        assert ours["frames_per_step"] == 128 * n and "1241x376" in ours["workload"] and "L2" in ours["l2"]
        assert not any(k in ours for k in ("keypoints_per_frame", "matches_per_frame"))     # measured values live in `stats`
    split = bench.frontend_config(_args(mode="split"), 1)
    assert "flush" in split["l2"]


def test_ba_config_is_static_and_shared_by_both_arms():
    bench_ba = importlib.import_module("bench_ba")
    c1 = bench_ba.ba_config(500, 50000, 1)
    assert c1 == bench_ba.ba_config(500, 50000, 1) and "500 KF / 50000 points" in c1["workload"] and c1["parallelism"] == "1 GPU"
    c4 = bench_ba.ba_config(500, 50000, 4)
    assert "sharded x4" in c4["parallelism"] and "flush" in c4["l2"]
