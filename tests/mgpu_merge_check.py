"""torchrun worker (2 ranks): the merged map's global BA with the map points sharded by ORIGIN MAP (robot A's points on rank 0, robot B's on rank 1) equals the
single-GPU solve.  Printed token: MGPU_MERGE_OK."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import orbslamm_b200 as ob
from orbslamm_b200 import sharding
import map_merge as M


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dist.init_process_group("nccl")
    dev = int(os.environ.get("LOCAL_RANK", 0))
    sc = M.make_scene(seed=0, Ka=10, Kb=10, n_world=1800)
    st = M.Stages("cuda", sc["voc"])
    out = M.run_merge(sc, st)
    g, single = out["gba_graph"], out["gba"]
    opt = ob.Optimizer(device=dev)
    uid = [ob.Optimizer.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    opt.comm_init(world, rank, uid[0])
    owner = M.owner_by_origin(g["origin"], world)
    s = sharding.shard_graph_by_owner(g, owner, rank)
    r = opt.BundleAdjustment(s["poses"], s["fixed"], s["intr"], s["points"], s["kf"], s["pt"], s["uv"], s["inv_sigma2"], 20, False)
    rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
    ok = r["lm_iterations"] == single["lm_iterations"] and rel(r["poses"], single["poses"]) < 1e-6 and rel(r["points"], single["points"][s["local_points"]]) < 1e-6
    t = torch.tensor([1 if ok else 0], device="cuda", dtype=torch.int32)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MGPU_MERGE_OK" if int(t.item()) else f"MGPU_MERGE_FAIL iterations {r['lm_iterations']} vs {single['lm_iterations']}, poses {rel(r['poses'], single['poses'])}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
