"""Seeded synthetic inputs for tests and bench.  numpy only.

Image streams: uniform noise -> Gaussian sigma 1.5 -> min-max normalise -> 300 filled
axis-aligned rectangles (+ a few flat 64x64 patches so the minThFAST fallback runs);
consecutive frames are the previous frame translated by an integer (dx, dy) in [-8, 8]^2
with a fresh noise border.  BA graphs: K keyframes on a smooth trajectory, P points in a
corridor, 3..9 observations per point, 2 % gross outliers, perturbed float32 initial state.
"""
import numpy as np

KITTI = dict(w=1241, h=376, nfeatures=2000, fx=718.856, fy=718.856, cx=607.1928, cy=185.2157)
TUM = dict(w=640, h=480, nfeatures=1000, fx=517.306408, fy=516.469215, cx=318.643040, cy=255.313989)


def _gauss_blur_np(a, sigma):
    r = int(3 * sigma + 0.5)
    k = np.exp(-0.5 * (np.arange(-r, r + 1) / sigma) ** 2)
    k /= k.sum()
    p = np.pad(a, ((0, 0), (r, r)), mode="reflect")
    a = sum(k[i] * p[:, i:i + a.shape[1]] for i in range(2 * r + 1))
    p = np.pad(a, ((r, r), (0, 0)), mode="reflect")
    return sum(k[i] * p[i:i + a.shape[0], :] for i in range(2 * r + 1))


def base_image(w, h, seed, n_rect=300, flat_frac=0.05):
    rng = np.random.default_rng(seed)
    img = _gauss_blur_np(rng.random((h, w)), 1.5)
    img = (img - img.min()) / (img.max() - img.min()) * 255.0
    img = img.astype(np.uint8)
    for _ in range(n_rect):
        x0, y0 = int(rng.integers(0, w)), int(rng.integers(0, h))
        sx, sy = int(rng.integers(8, 61)), int(rng.integers(8, 61))
        img[y0:y0 + sy, x0:x0 + sx] = int(rng.integers(0, 256))
    n_flat = int(flat_frac * (w * h) / (64 * 64))
    for _ in range(n_flat):
        x0, y0 = int(rng.integers(0, max(1, w - 64))), int(rng.integers(0, max(1, h - 64)))
        img[y0:y0 + 64, x0:x0 + 64] = int(rng.integers(0, 256))
    return img


def stream(w, h, n_frames, stream_id=0):
    """Yields (image u8 HxW, (dx, dy) shift from the previous frame)."""
    frames, shifts = [], []
    cur = base_image(w, h, 1000 * stream_id)
    frames.append(cur); shifts.append((0, 0))
    for f in range(1, n_frames):
        rng = np.random.default_rng(1000 * stream_id + f)
        dx, dy = int(rng.integers(-8, 9)), int(rng.integers(-8, 9))
        nxt = rng.integers(0, 256, (h, w), dtype=np.uint8)
        xs0, xs1 = max(0, -dx), min(w, w - dx)
        ys0, ys1 = max(0, -dy), min(h, h - dy)
        nxt[ys0 + dy:ys1 + dy, xs0 + dx:xs1 + dx] = cur[ys0:ys1, xs0:xs1]
        cur = nxt
        frames.append(cur); shifts.append((dx, dy))
    return frames, shifts


def _rot_y(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])


def _rodrigues(w):
    th = np.linalg.norm(w)
    if th < 1e-12:
        return np.eye(3)
    k = w / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def ba_graph(K=100, P=10000, seed=42, cam=KITTI, min_obs=3, max_obs=9, outlier_frac=0.02):
    """Synthetic covisibility graph.  Returns dict of float32/int32 arrays in the C-ABI layout:
    poses [K,4,4] f32 (Tcw, perturbed), fixed [K] u8, intr [4] f64 shared, points [P,3] f32 (perturbed),
    edges: kf i32[E], pt i32[E], uv f32[E,2], inv_sigma2 f32[E]; plus ground truth."""
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy, w, h = cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["w"], cam["h"]
    yaw = np.cumsum(rng.normal(0, np.deg2rad(2.0), K))
    pos = np.zeros((K, 3))
    for k in range(1, K):
        pos[k] = pos[k - 1] + _rot_y(yaw[k - 1]) @ np.array([0, 0, 1.0])
    Rwc = np.stack([_rot_y(a) for a in yaw])
    Rcw = np.transpose(Rwc, (0, 2, 1))
    tcw = -np.einsum("kij,kj->ki", Rcw, pos)
    # points: ahead of a random keyframe, inside a 40 m wide corridor
    anchor = rng.integers(0, K, P)
    local = np.stack([rng.uniform(-20, 20, P), rng.uniform(-3, 3, P), rng.uniform(4, 40, P)], 1)
    Xw = pos[anchor] + np.einsum("pij,pj->pi", Rwc[anchor], local)
    nobs_target = rng.integers(min_obs, max_obs + 1, P)
    # candidate keyframes of a point: anchor, anchor-1, anchor+1, anchor-2, ... (nearest first); keep the first
    # nobs_target of them that see the point in front of the camera and inside the image
    C = 4 * max_obs
    offs = np.zeros(C, np.int64)
    offs[1::2] = -(np.arange(1, C, 2) // 2 + 1)
    offs[2::2] = np.arange(2, C, 2) // 2
    cand = anchor[:, None] + offs[None, :]                                   # [P, C]
    inside = (cand >= 0) & (cand < K)
    ck = np.clip(cand, 0, K - 1)
    Xc = np.einsum("pcij,pj->pci", Rcw[ck], Xw) + tcw[ck]
    z = Xc[..., 2]
    zs = np.where(z > 0.5, z, 1.0)
    u = fx * Xc[..., 0] / zs + cx
    v = fy * Xc[..., 1] / zs + cy
    vis = inside & (z > 0.5) & (u >= 0) & (u < w) & (v >= 0) & (v < h)
    take = vis & (np.cumsum(vis, 1) <= nobs_target[:, None])
    pi, ci = np.nonzero(take)
    kf = ck[pi, ci].astype(np.int32); pt = pi.astype(np.int32)
    uv = np.stack([u[pi, ci], v[pi, ci]], 1)
    E = len(kf)
    octave = rng.integers(0, 8, E)
    sig = 1.2 ** octave
    uv += rng.normal(0, 1.0, (E, 2)) * sig[:, None]
    out = rng.random(E) < outlier_frac
    ang = rng.uniform(0, 2 * np.pi, E)
    mag = rng.uniform(10, 50, E)
    uv[out] += (mag[out, None] * np.stack([np.cos(ang[out]), np.sin(ang[out])], 1))
    inv_sigma2 = (1.2 ** (-2.0 * octave))
    # perturbed initial state
    poses = np.zeros((K, 4, 4), np.float32)
    for k in range(K):
        dR = _rodrigues(rng.normal(0, np.deg2rad(0.5), 3)) if k > 1 else np.eye(3)
        dt = rng.normal(0, 0.05, 3) if k > 1 else np.zeros(3)
        poses[k, :3, :3] = (dR @ Rcw[k]).astype(np.float32)
        poses[k, :3, 3] = (dR @ tcw[k] + dt).astype(np.float32)
        poses[k, 3, 3] = 1
    points = (Xw + rng.normal(0, 0.10, (P, 3))).astype(np.float32)
    # KF 0 is fixed because mnId==0 (written back); KF 1 plays a fixed camera outside the local window: together
    # they pin the monocular scale gauge like lFixedCameras does in the reference (Optimizer.cc:512-527)
    fixed = np.zeros(K, np.uint8); fixed[0] = 1
    if K > 2:
        fixed[1] = 2
    gt_poses = np.zeros((K, 4, 4)); gt_poses[:, :3, :3] = Rcw; gt_poses[:, :3, 3] = tcw; gt_poses[:, 3, 3] = 1
    return dict(poses=poses, fixed=fixed, intr=np.array([fx, fy, cx, cy], np.float64), points=points,
                kf=kf, pt=pt, uv=uv.astype(np.float32), inv_sigma2=inv_sigma2.astype(np.float32),
                gt_poses=gt_poses, gt_points=Xw, is_outlier=out)


def permute_keyframes(g, seed=0):
    """Same problem with the keyframes renumbered at random: the reduced pose system loses its band structure (what a loop
    closure / merged multi-robot map looks like to the solver)."""
    rng = np.random.default_rng(seed)
    K = len(g["poses"])
    perm = rng.permutation(K)                  # new index of old keyframe k
    inv = np.argsort(perm)
    out = dict(g)
    out["poses"] = g["poses"][inv].copy(); out["fixed"] = g["fixed"][inv].copy(); out["gt_poses"] = g["gt_poses"][inv].copy()
    out["kf"] = perm[g["kf"]].astype(np.int32)
    out["perm"] = perm
    return out
