"""world_size-2 gloo tests of the host-side multi-GPU logic (no GPU): BA graph sharding covers every point and edge
exactly once, shards reassemble, and the bench's max-over-ranks reduction works over torch.distributed."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from orbslamm_b200 import sharding, synth


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = synth.ba_graph(K=12, P=301, seed=9)
    sh = sharding.shard_graph(g, world, rank)
    # every rank sees all keyframes, only its points / edges
    assert sh["poses"].shape == g["poses"].shape and len(sh["points"]) == len(sh["local_points"])
    assert np.array_equal(g["pt"][sh["local_edges"]], sh["local_points"][sh["pt"]])
    # "update" the shard (stand-in for the BA result) and all-gather it back
    upd = sh["points"] + np.float32(rank + 1)
    objs = [None] * world
    dist.all_gather_object(objs, (sh["local_points"], upd))
    full = sharding.merge_points(len(g["points"]), objs)
    expect = g["points"] + (np.arange(len(g["points"])) % world + 1).astype(np.float32)[:, None]
    assert np.array_equal(full, expect)
    eobjs = [None] * world
    dist.all_gather_object(eobjs, (sh["local_edges"], sh["uv"]))
    assert np.array_equal(sharding.merge_edges(len(g["kf"]), eobjs), g["uv"])
    counts = torch.tensor([len(sh["local_points"]), len(sh["local_edges"])])
    dist.all_reduce(counts)
    assert counts.tolist() == [len(g["points"]), len(g["kf"])]
    # timing reduction used by bench.py: max over ranks
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert float(t) == float(world)
    ret[rank] = 1
    dist.destroy_process_group()


def test_ba_sharding_world2():
    world = 2
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert sorted(ret.keys()) == [0, 1]


def test_shard_partition_properties():
    rng = np.random.default_rng(0)
    e_pt = rng.integers(0, 1000, 7000)
    for world in (1, 2, 3, 8):
        seen_p, seen_e = [], []
        for r in range(world):
            lp, le, el = sharding.shard_points(1000, e_pt, world, r)
            assert np.array_equal(lp[el], e_pt[le])
            seen_p.append(lp); seen_e.append(le)
        assert np.array_equal(np.sort(np.concatenate(seen_p)), np.arange(1000))
        assert np.array_equal(np.sort(np.concatenate(seen_e)), np.arange(7000))


def _owner_worker(rank, world, port, ret):
    """sharding by an explicit owner (configs[3]: points stay on the GPUs of their origin map): the per-keyframe observation counts -- a stand-in for the
    partial Hpp blocks every rank contributes to the reduced system -- summed over the ranks equal the full graph's"""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = synth.ba_graph(K=14, P=333, seed=4)
    origin = (np.arange(len(g["points"])) * 7 % 5 < 2).astype(np.int64)            # an uneven split into two "maps"
    sh = sharding.shard_graph_by_owner(g, origin, rank)
    assert np.all(origin[sh["local_points"]] == rank) and np.array_equal(sh["local_points"][sh["pt"]], g["pt"][sh["local_edges"]])
    part = torch.from_numpy(np.bincount(sh["kf"], minlength=len(g["poses"])).astype(np.int64))
    dist.all_reduce(part)
    assert np.array_equal(part.numpy(), np.bincount(g["kf"], minlength=len(g["poses"])))
    objs = [None] * world
    dist.all_gather_object(objs, (sh["local_points"], sh["points"]))
    assert np.array_equal(sharding.merge_points(len(g["points"]), objs), g["points"])
    ret[rank] = 1
    dist.destroy_process_group()


def test_ba_sharding_by_owner_world2():
    world = 2
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_owner_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert sorted(ret.keys()) == [0, 1]


def test_owner_by_origin_splits_the_gpus_between_the_two_maps():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import map_merge as M
    rng = np.random.default_rng(1)
    origin = (rng.random(1000) < 0.4).astype(np.int64)
    assert np.all(M.owner_by_origin(origin, 1) == 0)
    for world in (2, 3, 4, 8):
        own = M.owner_by_origin(origin, world)
        half = world // 2
        assert own[origin == 0].max() < half and own[origin == 1].min() >= half and own.max() < world
        for o, lo, hi in ((0, 0, half), (1, half, world)):                           # round-robin inside a map's GPUs: balanced to within one point
            c = np.bincount(own[origin == o], minlength=world)[lo:hi]
            assert c.max() - c.min() <= 1
