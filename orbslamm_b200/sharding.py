"""Host-side sharding of a bundle-adjustment graph over ranks: map points, with all their
observations, are dealt round-robin to the ranks; keyframes are replicated.  Pure numpy: used by bench.py, the
multi-GPU test and the gloo CPU tests."""
import numpy as np


def shard_points(n_points, e_pt, world, rank):
    """Returns (local_point_ids [Pl] global ids ascending, local_edge_ids [El] indices into the edge arrays,
    e_pt_local [El] point index within the shard)."""
    e_pt = np.asarray(e_pt)
    local_points = np.arange(rank, n_points, world, dtype=np.int64)
    owner = e_pt % world
    local_edges = np.nonzero(owner == rank)[0]
    e_pt_local = (e_pt[local_edges] // world).astype(np.int32)
    return local_points, local_edges, e_pt_local


def shard_graph(g, world, rank):
    """g: dict with poses, fixed, intr, points, kf, pt, uv, inv_sigma2 (synth.ba_graph layout) -> the rank's shard."""
    lp, le, ept = shard_points(len(g["points"]), g["pt"], world, rank)
    return dict(poses=g["poses"], fixed=g["fixed"], intr=g["intr"], points=g["points"][lp], kf=g["kf"][le], pt=ept,
                uv=g["uv"][le], inv_sigma2=g["inv_sigma2"][le], local_points=lp, local_edges=le)


def shard_graph_by_owner(g, owner, rank):
    """Shard with an explicit owner rank per map point (BASELINE.json configs[3]: after a map merge the points stay on the GPU of the robot whose map
    they come from; keyframes are replicated).  owner: int [P]."""
    owner = np.asarray(owner)
    lp = np.nonzero(owner == rank)[0].astype(np.int64)
    local_index = np.full(len(owner), -1, np.int64); local_index[lp] = np.arange(len(lp))
    le = np.nonzero(owner[np.asarray(g["pt"])] == rank)[0]
    return dict(poses=g["poses"], fixed=g["fixed"], intr=g["intr"], points=g["points"][lp], kf=g["kf"][le], pt=local_index[np.asarray(g["pt"])[le]].astype(np.int32),
                uv=g["uv"][le], inv_sigma2=g["inv_sigma2"][le], local_points=lp, local_edges=le)


def merge_points(n_points, shards):
    """shards: list over ranks of (local_point_ids, values [Pl, ...]) -> full array [n_points, ...]."""
    first = shards[0][1]
    out = np.zeros((n_points,) + first.shape[1:], first.dtype)
    for ids, vals in shards:
        out[ids] = vals
    return out


def merge_edges(n_edges, shards):
    first = shards[0][1]
    out = np.zeros((n_edges,) + first.shape[1:], first.dtype)
    for ids, vals in shards:
        out[ids] = vals
    return out
