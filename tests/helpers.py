"""Shared builders for matcher / optimizer parity tests (CPU side only uses numpy + the oracle)."""
import numpy as np

import oracle
from orbslamm_b200 import synth


def make_tracking_case(cam, stream_id, nfeatures=None, seed=0):
    """Two consecutive synthetic frames -> (cur features, last-frame 'map points', pose, intrinsics).

    The last frame's keypoints become map points at depth U(5, 50) m placed so that, under the returned current
    pose, they project onto their shifted position (SURVEY.md 8d 'match workload')."""
    nf = nfeatures or cam["nfeatures"]
    frames, shifts = synth.stream(cam["w"], cam["h"], 2, stream_id=stream_id)
    P = oracle.orb_params(nf, 1.2, 8, 20, 7)
    last = oracle.orb_extract(P, frames[0])
    cur = oracle.orb_extract(P, frames[1])
    dx, dy = shifts[1]
    rng = np.random.default_rng(seed + 17 * stream_id)
    fx, fy, cx, cy = cam["fx"], cam["fy"], cam["cx"], cam["cy"]
    n = len(last["x"])
    z = rng.uniform(5, 50, n)
    # current pose: small rotation about y plus a translation; Xw = Rcw^T (Xc - tcw)
    a = np.deg2rad(0.7)
    Rcw = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    tcw = np.array([0.11, -0.04, 0.3])
    Xc = np.stack([(last["x"] + dx - cx) * z / fx, (last["y"] + dy - cy) * z / fy, z], 1)
    Xw = (Xc - tcw) @ Rcw            # rows: Rcw^T (Xc - t)
    Tcw = np.eye(4, dtype=np.float32)
    Tcw[:3, :3] = Rcw.astype(np.float32); Tcw[:3, 3] = tcw.astype(np.float32)
    K4 = np.array([fx, fy, cx, cy], np.float32)
    bounds = np.array([0, 0, cam["w"], cam["h"]], np.float32)
    valid = (rng.random(n) < 0.9).astype(np.uint8)       # some slots hold no map point / are outliers
    return dict(P=P, last=last, cur=cur, Tcw=Tcw, K4=K4, bounds=bounds, Xw=Xw.astype(np.float32), valid=valid,
                frames=frames, shift=(dx, dy))


def slab(arrs, width, dtype, tail=()):
    """Stack ragged per-frame arrays into a [n_frames, width, *tail] slab."""
    out = np.zeros((len(arrs), width) + tuple(tail), dtype)
    for i, a in enumerate(arrs):
        out[i, :len(a)] = a
    return out
