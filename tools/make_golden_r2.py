#!/usr/bin/env python3
"""tests/golden/round2_small.npz: known answers for the entry points added in round 2 and for the composed two-map merge.

  * MapPoint::ComputeDistinctiveDescriptors: ragged descriptor lists and the chosen rows (oracle.distinctive_descriptor: the loop of MapPoint.cc:271-303 restated
    over the DescriptorDistance that is pinned by the reference's ORBmatcher.cc object code);
  * Sim3Solver::ComputeSim3: 40 min sets and the T12 / T21 the reference's OpenCV call sequence gives through cv2 4.13 (float tolerance on the device side);
  * configs[3] merge chain on the 10 + 10 keyframe scene, driven by the oracle: every decision the chain takes and the merged keyframe poses.
    python tools/make_golden_r2.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle            # noqa: E402
import kf_family as kff  # noqa: E402
import map_merge as M    # noqa: E402

rng = np.random.default_rng(20)
lists = []
for n in [0, 1, 2, 3, 6, 9, 33, 41] + [int(x) for x in rng.integers(2, 12, 40)]:
    base = rng.integers(0, 256, 32, dtype=np.uint8)
    d = np.tile(base, (n, 1))
    for i in range(n):
        for b in rng.integers(0, 256, int(rng.integers(0, 50))): d[i, b >> 3] ^= np.uint8(1 << (b & 7))
    lists.append(d)
start = np.concatenate([[0], np.cumsum([len(d) for d in lists])]).astype(np.int32)
flat = np.concatenate([d for d in lists if len(d)]).astype(np.uint8)
best = np.array([oracle.distinctive_descriptor(d) for d in lists], np.int32)

r = kff.make_sim3_ransac_case(kff.GOLDEN_CAM, 21)
tri = M.sample_triples(len(r["X1"]), 40, np.random.default_rng(21))
ref = [M.compute_sim3(r["X1"][q].T.copy(), r["X2"][q].T.copy()) for q in tri]

sc = M.make_scene(seed=0, Ka=10, Kb=10, n_world=1800)
out = M.run_merge(sc, M.Stages("oracle", sc["voc"]))
dec = np.array(out["candidates"] + out["bow_matches"] + list(out["ransac"][-1]) + list(out["sim3_inliers"][-1]) + [out["total_matches"], out["fused"], out["essential_edges"],
               out["loop_connections"], out["gba"]["lm_iterations"]], np.int64)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "round2_small.npz"), dd_flat=flat, dd_start=start, dd_best=best,
                    s3_X1=r["X1"][tri], s3_X2=r["X2"][tri], s3_T12=np.stack([x[0] for x in ref]), s3_T21=np.stack([x[1] for x in ref]),
                    mm_decisions=dec, mm_poses=out["poses"].astype(np.float32), mm_points_head=out["points"][:200].astype(np.float32))
print("round2_small.npz", os.path.getsize(os.path.join(ROOT, "tests", "golden", "round2_small.npz")), "decisions", dec.tolist())
