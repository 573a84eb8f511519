"""Run under torchrun with N >= 2 GPUs: sharded LocalBA (NCCL all-reduce on the reduced pose system) must equal the
single-GPU result.  Launched by tests/test_multigpu_gpu.py."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import orbslamm_b200 as ob                      # noqa: E402
from orbslamm_b200 import sharding, synth       # noqa: E402


def main():
    import faulthandler
    faulthandler.dump_traceback_later(100, exit=True)
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo")
    g = synth.ba_graph(K=40, P=3000, seed=13)
    opt = ob.Optimizer(device=dev)
    single = opt.LocalBundleAdjustment(g["poses"], g["fixed"], g["intr"], g["points"], g["kf"], g["pt"], g["uv"], g["inv_sigma2"]) if rank == 0 else None
    uid = [ob.Optimizer.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    opt.comm_init(world, rank, uid[0])
    sh = sharding.shard_graph(g, world, rank)
    r = opt.LocalBundleAdjustment(sh["poses"], sh["fixed"], sh["intr"], sh["points"], sh["kf"], sh["pt"], sh["uv"], sh["inv_sigma2"])
    pts = [None] * world; out = [None] * world; poses = [None] * world
    dist.all_gather_object(pts, (sh["local_points"], r["points"]))
    dist.all_gather_object(out, (sh["local_edges"], r["outlier"]))
    dist.all_gather_object(poses, r["poses"])
    if rank == 0:
        full_pts = sharding.merge_points(len(g["points"]), pts); full_out = sharding.merge_edges(len(g["kf"]), out)
        for p in poses[1:]:
            assert np.array_equal(p, poses[0]), "poses differ between ranks"
        rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
        assert (r["lm_iterations"], r["lm_trials"]) == (single["lm_iterations"], single["lm_trials"])
        assert rel(poses[0], single["poses"]) < 1e-6 and rel(full_pts, single["points"]) < 1e-6, (rel(poses[0], single["poses"]), rel(full_pts, single["points"]))
        assert (full_out != single["outlier"]).sum() <= 2
        print(f"MGPU_BA_OK world={world} iters={r['lm_iterations']} rel_pose={rel(poses[0], single['poses']):.2e} rel_pts={rel(full_pts, single['points']):.2e}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
