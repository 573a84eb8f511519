"""BASELINE.json configs[3]: two robots' maps -> Sim3 map merge -> MMOptimizeEssentialGraph -> MMGlobalBundleAdjustemnt, composed from the C-ABI entry points in the
order MultiMapper runs them (M/src/MultiMapper.cc:209-306 candidate loop, :328-423 acceptance, :451-662 UpdatePosesAndAdd, :664-690 SearchAndFuse; M/src/LoopClosing.cc:818
the global BA; M/src/Optimizer.cc:40-57, :1068-1346 the MM variants = the single-map algorithms over the union of the attached maps).

The control flow, the map mutations (Replace / AddObservation / UpdateConnections) and the tiny host algebra stay host code in the reference and here (numpy); every
numeric stage goes through a backend -- the CPU oracle or the CUDA library -- so that the same composition is the test (stage by stage on identical inputs, and end to
end) and the bench leg (CUDA only).  Test infrastructure: the product never imports this file.

Synthetic scene (feature level, no images): one world of 3-D points, robot A drives x = -14 .. -1, robot B x = -4 .. 9; every keyframe observes the points in its
frustum (projection + noise, octave from depth, a per-point orientation, the point's 256-bit descriptor with a few flipped bits).  Map A lives in the world frame, map B
in its own frame and scale (monocular maps have arbitrary scale): X_b = (R_bw (X_w - c_b)) / s_b."""
import numpy as np

from orbslamm_b200 import synth
from orbslamm_b200 import vocabulary as V
import kf_family as F

oracle = F.oracle            # lazy: loaded on first use by the oracle stages only

F32 = np.float32
CAM = synth.KITTI
NL = 8
SF = (1.2 ** np.arange(NL)).astype(F32)
LS2 = (SF.astype(np.float64) ** 2).astype(F32)
INV2 = (1.0 / LS2.astype(np.float64)).astype(F32)
K4 = np.array([CAM["fx"], CAM["fy"], CAM["cx"], CAM["cy"]], F32)
BOUNDS = np.array([0, 0, CAM["w"], CAM["h"]], F32)


def _rot(axis, deg):
    from scipy.spatial.transform import Rotation
    return Rotation.from_euler(axis, deg, degrees=True).as_matrix()


def _pose(R, c):
    T = np.eye(4); T[:3, :3] = R; T[:3, 3] = -R @ c
    return T


# ------------------------------------------------------------------------------------------------ scene
def make_scene(seed=0, Ka=14, Kb=14, n_world=2600, s_b=1.3):
    rng = np.random.default_rng(seed)
    x_hi = max(45.0, -4.0 + max(Ka - 10, Kb) + 36.0)                 # the world extends with the trajectories (n_world sets the density)
    Xw = np.stack([rng.uniform(-45, x_hi, n_world), rng.uniform(-4, 3, n_world), rng.uniform(6, 45, n_world)], 1)
    base_desc = rng.integers(0, 256, (n_world, 32), dtype=np.uint8)
    twins = rng.permutation(n_world)[:n_world // 8]                  # repetitive texture: an eighth of the points share their descriptor with another point
    base_desc[twins] = base_desc[(twins + 7) % n_world]
    base_ang = rng.uniform(0, 360, n_world)
    R_bw = _rot("y", 8.0) @ _rot("x", 2.0); c_b = np.array([3.0, 0.5, -2.0])

    def trajectory(x0, K, yaw0, zoff):
        Ts = []
        for k in range(K):
            R = _rot("y", yaw0 + rng.normal(0, 0.7)) @ _rot("x", rng.normal(0, 0.4))
            Ts.append(_pose(R, np.array([x0 + 1.0 * k, rng.normal(0, 0.05), zoff + rng.normal(0, 0.05)])))
        return Ts

    def build_map(Ts_world, to_map, scale):
        """to_map(Xw) -> map coordinates; scale: map units per world unit.  Returns the map dict (keyframes with features, map points)."""
        K = len(Ts_world)
        obs = []                                    # per keyframe: (world point ids, uv, depth)
        for T in Ts_world:
            Xc = Xw @ T[:3, :3].T + T[:3, 3]
            z = Xc[:, 2]
            ok = z > 4.0
            u = CAM["fx"] * Xc[:, 0] / np.where(ok, z, 1) + CAM["cx"] + rng.normal(0, 0.4, n_world)
            v = CAM["fy"] * Xc[:, 1] / np.where(ok, z, 1) + CAM["cy"] + rng.normal(0, 0.4, n_world)
            ok &= (u > 20) & (u < CAM["w"] - 20) & (v > 20) & (v < CAM["h"] - 20)
            ids = np.where(ok)[0]
            ids = ids[rng.permutation(len(ids))]
            assoc = rng.random(len(ids)) < 0.75          # a quarter of the features has no map point (yet)
            obs.append((ids, np.stack([u[ids], v[ids]], 1), np.linalg.norm(Xc[ids], axis=1), assoc))
        seen = np.zeros(n_world, int)
        for ids, _, _, assoc in obs: seen[ids[assoc]] += 1
        mp_of_world = np.full(n_world, -1); wid = np.where(seen >= 2)[0]
        mp_of_world[wid] = np.arange(len(wid))
        P = len(wid)
        Xm = to_map(Xw[wid] + rng.normal(0, 0.04, (P, 3)))
        mp = dict(Xw=Xm.astype(F32), world=wid, ref=np.full(P, -1), desc=np.zeros((P, 32), np.uint8), normal=np.zeros((P, 3)), mf_max=np.zeros(P, F32),
                  mf_min=np.zeros(P, F32), obs=[dict() for _ in range(P)])
        kfs = []
        for k, (T, (ids, uv, d, has_mp)) in enumerate(zip(Ts_world, obs)):
            n = len(ids)
            octave = np.clip(np.rint(np.log(45.0 / d) / np.log(1.2)), 0, NL - 1).astype(np.int32)
            ang = ((base_ang[ids] + rng.normal(0, 1.5, n)) % 360.0).astype(F32)
            desc = base_desc[ids].copy()
            for i in range(n):
                nflip = int(rng.integers(0, 6)) if rng.random() < 0.8 else int(rng.integers(20, 70))
                for b in rng.integers(0, 256, nflip): desc[i, b >> 3] ^= np.uint8(1 << (b & 7))
            Tm = np.eye(4); Tm[:3, :3] = T[:3, :3] @ np.linalg.inv(to_map.R); Tm[:3, 3] = scale * (T[:3, :3] @ to_map.c + T[:3, 3])
            assoc = np.where(has_mp, mp_of_world[ids], -1)
            kf = dict(xy=uv.astype(F32), octave=octave, angle=ang, desc=np.ascontiguousarray(desc), scale_factors=SF, inv_level_sigma2=INV2, K4=K4,
                      grid_bounds4=BOUNDS.copy(), Tcw=Tm.astype(F32), log_sf=F32(np.log(SF[1])), mp=assoc, world=ids)
            kf["win_origin2"] = F.kf_bounds(kf)[:2].copy()
            for i in np.where(kf["mp"] >= 0)[0]:
                p = kf["mp"][i]
                mp["obs"][p][k] = int(i)
                if mp["ref"][p] < 0:
                    mp["ref"][p] = k; mp["desc"][p] = desc[i]
                    mp["mf_max"][p] = F32(d[i] * scale * 1.2 ** (np.clip(octave[i], 1, 6) - 0.5)); mp["mf_min"][p] = F32(mp["mf_max"][p] / SF[NL - 1])
                c = -Tm[:3, :3].T @ Tm[:3, 3]
                n_i = Xm[p] - c
                mp["normal"][p] += n_i / np.linalg.norm(n_i)
            kfs.append(kf)
        mp["normal"] = (mp["normal"] / np.linalg.norm(mp["normal"], axis=1)[:, None]).astype(F32)
        return dict(kfs=kfs, mp=mp, centres_world=np.stack([-T[:3, :3].T @ T[:3, 3] for T in Ts_world]))

    class ToMap:
        def __init__(self, R, c, s): self.R, self.c, self.s = R, c, s
        def __call__(self, X): return (X - self.c) @ self.R.T * self.s
    A = build_map(trajectory(-14.0, Ka, 0.0, 0.0), ToMap(np.eye(3), np.zeros(3), 1.0), 1.0)
    B = build_map(trajectory(-4.0, Kb, 3.0, 0.5), ToMap(R_bw, c_b, 1.0 / s_b), 1.0 / s_b)
    return dict(A=A, B=B, voc=V.synthetic(10, 3, seed=seed + 5), s_b=s_b, seed=seed)


def covisibility(m):
    """KeyFrame::UpdateConnections weights (KeyFrame.cc:310-377): shared map points per keyframe pair."""
    K = len(m["kfs"])
    W = np.zeros((K, K), int)
    for o in m["mp"]["obs"]:
        ks = sorted(o)
        for a in range(len(ks)):
            for b in range(a + 1, len(ks)): W[ks[a], ks[b]] += 1; W[ks[b], ks[a]] += 1
    return W


def feature_points(m, k):
    """the map point behind every feature of keyframe k, as the per-feature arrays SearchBySim3 / Sim3Solver read"""
    kf, mp = m["kfs"][k], m["mp"]
    i = np.where(kf["mp"] >= 0, kf["mp"], 0)
    return dict(Xw=mp["Xw"][i], normal=mp["normal"][i], mf_min=mp["mf_min"][i], mf_max=mp["mf_max"][i], desc=mp["desc"][i]), (kf["mp"] >= 0).astype(np.uint8)


# ------------------------------------------------------------------------------------------------ Sim3Solver host side
def compute_sim3(P1, P2):
    """Sim3Solver::ComputeSim3 (Sim3Solver.cc:226-338) with the OpenCV calls the reference makes, through cv2 (float path).  P1, P2: 3x3 float, points in columns.
    Returns (T12, T21) 4x4 float32."""
    import cv2
    P1 = np.ascontiguousarray(P1, F32); P2 = np.ascontiguousarray(P2, F32)
    def centroid(P):
        C = cv2.reduce(P, 1, cv2.REDUCE_SUM)
        C = F.scale32(C, 1.0 / np.float64(P.shape[1]))
        return P - C, C
    Pr1, O1 = centroid(P1); Pr2, O2 = centroid(P2)
    M = cv2.gemm(Pr2, Pr1, 1.0, None, 0.0, flags=cv2.GEMM_2_T)          # Pr2 * Pr1.t(): a MatExpr gemm with a transposed operand
    m = lambda r, c: np.float64(M[r, c])
    N11 = m(0, 0) + m(1, 1) + m(2, 2); N12 = m(1, 2) - m(2, 1); N13 = m(2, 0) - m(0, 2); N14 = m(0, 1) - m(1, 0)
    N22 = m(0, 0) - m(1, 1) - m(2, 2); N23 = m(0, 1) + m(1, 0); N24 = m(2, 0) + m(0, 2)
    N33 = -m(0, 0) + m(1, 1) - m(2, 2); N34 = m(1, 2) + m(2, 1); N44 = -m(0, 0) - m(1, 1) + m(2, 2)
    N = np.array([[N11, N12, N13, N14], [N12, N22, N23, N24], [N13, N23, N33, N34], [N14, N24, N34, N44]], F32)
    _, _, evec = cv2.eigen(N)
    vec = evec[0:1, 1:4].copy()
    ang = np.arctan2(cv2.norm(vec), np.float64(evec[0, 0]))
    vec = (vec.astype(np.float64) * (2 * ang) / cv2.norm(vec)).astype(F32)
    R12, _ = cv2.Rodrigues(vec)
    R12 = R12.astype(F32)
    P3 = cv2.gemm(R12, Pr2, 1.0, None, 0.0)
    nom = float(np.sum(Pr1.astype(np.float64) * P3.astype(np.float64)))
    den = 0.0
    for v in cv2.pow(P3, 2).ravel(): den += float(v)
    s12 = F32(nom / den)
    t12 = (O1 - F.scale32(cv2.gemm(R12, O2, 1.0, None, 0.0), np.float64(s12))).astype(F32)
    T12 = np.eye(4, dtype=F32); T12[:3, :3] = F.scale32(R12, np.float64(s12)); T12[:3, 3] = t12.ravel()
    sRinv = F.scale32(R12.T.copy(), 1.0 / np.float64(s12))
    T21 = np.eye(4, dtype=F32); T21[:3, :3] = sRinv; T21[:3, 3] = (-cv2.gemm(sRinv, t12, 1.0, None, 0.0)).ravel()
    return T12, T21, R12, t12.ravel(), s12


def ransac_iterations(N, prob=0.99, min_inliers=10, max_its=300):
    """Sim3Solver::SetRansacParameters (:114-137)"""
    eps = F32(min_inliers) / F32(N)
    n = 1 if min_inliers == N else int(np.ceil(np.log(1 - prob) / np.log(1 - float(eps) ** 3)))
    return max(1, min(n, max_its))


def sample_triples(N, n_its, rng):
    """the min-set draws of Sim3Solver::iterate (:161-175) incl. the reference's vAvailableIndices[idx] quirk; the draws do not depend on the inlier counts"""
    out = np.zeros((n_its, 3), int)
    for it in range(n_its):
        avail = list(range(N))
        for i in range(3):
            randi = int(rng.random() * len(avail))
            idx = avail[randi]
            out[it, i] = idx
            if idx < len(avail): avail[idx] = avail[-1]
            avail.pop()
    return out


# ------------------------------------------------------------------------------------------------ backends for the stages kf_family does not cover
class Stages:
    """kind = 'oracle' | 'cuda'"""
    def __init__(self, kind, voc):
        self.kind = kind
        self.B = F.OracleBackend() if kind == "oracle" else F.CudaBackend()
        self.voc = voc
        if kind == "cuda":
            import orbslamm_b200 as ob
            self.ob = ob; self.opt = ob.Optimizer(); self.vh = V.ORBVocabulary(voc)

    def transform(self, descs):
        if self.kind == "oracle":
            return [oracle.vocab_transform(self.voc, d, 2) for d in descs]
        S = max(len(d) for d in descs)
        slab = np.zeros((len(descs), S, 32), np.uint8)
        for i, d in enumerate(descs): slab[i, :len(d)] = d
        return self.vh.transform(slab, [len(d) for d in descs], 2)

    def bow(self, k1, e1, fv1, k2, e2, fv2):
        return F._bow(self.B, 0, k1["desc"], k1["angle"], e1, fv1, k2["desc"], k2["angle"], e2, fv2, 0.75, True)

    def sim3_check(self, T12, T21, X1, X2, o1, o2):
        if self.kind == "oracle":
            m1, m2, p1, p2 = oracle.sim3_prepare(X1, X2, o1, o2, LS2, K4, K4)
            return oracle.sim3_check_inliers(T12, T21, X1, X2, p1, p2, m1, m2, K4, K4)
        m1, p1 = self.opt.Sim3Prepare(X1, o1, LS2, K4); m2, p2 = self.opt.Sim3Prepare(X2, o2, LS2, K4)
        return self.opt.Sim3CheckInliers(T12, T21, X1, X2, p1, p2, m1, m2, K4, K4)

    def optimize_sim3(self, init, valid, P1c, P2c, obs1, obs2, w1, w2):
        if self.kind == "oracle":
            r = oracle.optimize_sim3(init, valid, P1c, P2c, obs1, obs2, w1, w2, K4, K4, 10.0, False)
            return r["sim3"], r["inlier"], int(r["n_in"])
        S, inl, nin, _ = self.opt.OptimizeSim3(init[None], valid[None], P1c[None], P2c[None], obs1[None], obs2[None], w1[None], w2[None], K4[None], K4[None],
                                               [len(valid)], 10.0, False)
        return S[0], inl[0], int(nin[0])

    def pose_graph(self, S, fixed, ei, ej, meas):
        f = oracle.optimize_pose_graph if self.kind == "oracle" else self.opt.OptimizePoseGraph
        return f(S, fixed, ei, ej, meas, False, 20)

    def global_ba(self, g, opt=None):
        if self.kind == "oracle":
            return oracle.bundle_adjust(g["poses"], g["fixed"], g["intr"], g["points"], g["kf"], g["pt"], g["uv"], g["inv_sigma2"], False, 20, 0, False)
        return (opt or self.opt).BundleAdjustment(g["poses"], g["fixed"], g["intr"], g["points"], g["kf"], g["pt"], g["uv"], g["inv_sigma2"], 20, False)


def _same(name, a, b, tol=None):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, f"{name}: shapes {a.shape} vs {b.shape}"
    if tol is None:
        assert np.array_equal(a, b), f"{name}: CUDA differs from the oracle ({int((a != b).sum())} entries)"
    else:
        err = float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
        assert err < tol, f"{name}: relative difference {err:.3g} >= {tol}"


# ------------------------------------------------------------------------------------------------ the merge
def sim3_row(R, t, s):
    return np.concatenate([F.quat_of(R), np.asarray(t, np.float64), [float(s)]])


def sim3_apply(S, X):
    from scipy.spatial.transform import Rotation
    return S[7] * Rotation.from_quat(S[:4]).apply(X) + S[4:7]


def sim3_matrix32(S):
    """Converter::toCvMat(g2o::Sim3): [sR t; 0 1] in float"""
    from scipy.spatial.transform import Rotation
    T = np.eye(4, dtype=F32); T[:3, :3] = (S[7] * Rotation.from_quat(S[:4]).as_matrix()).astype(F32); T[:3, 3] = S[4:7].astype(F32)
    return T


def cam_points32(Tcw, X):
    """Rcw * X + tcw per point in float (cv::Mat small gemm), Sim3Solver.cc:93-97 / Optimizer.cc:1394-1419"""
    T = np.asarray(Tcw, F32)
    return np.stack([F.gemm32(T[:3, :3], x.reshape(3, 1)).ravel() + T[:3, 3] for x in np.asarray(X, F32)]).astype(F32) if len(X) else np.zeros((0, 3), F32)


def run_merge(scene, drive, check=None, out=None, gba_opt=None):
    """MultiMapper's merge of map B (pMap, the robot that detects) into map A (pMapBase).  `drive`: Stages whose results are fed forward; `check`: optional second Stages
    run on the SAME inputs at every stage and asserted equal (integer stages exactly, fp64 stages to 1e-5 / identical flags).  Returns the merged graph and results."""
    A, Bm = scene["A"], scene["B"]
    out = {} if out is None else out
    import time
    timing = out.setdefault("stage_ms", {})
    def both(name, fn, cmp):
        t0 = time.perf_counter()
        r = fn(drive)
        key = name.split(" cand ")[0].split(" KF ")[0]
        timing[key] = timing.get(key, 0.0) + (time.perf_counter() - t0) * 1e3
        if check is not None:
            cmp(name, fn(check), r)
        return r
    def cmp_exact(name, rc, rd):
        for i, (x, y) in enumerate(zip(rc, rd)): _same(f"{name}[{i}]", x, y)
    cur = len(Bm["kfs"]) // 4                      # pKeyFrame: a keyframe of map B inside the overlap
    kfc = Bm["kfs"][cur]
    ptsc, hasc = feature_points(Bm, cur)
    # ---- candidates: DetectRelocalizationCandidates is KeyFrameDatabase scoring (out of scope); here: the map-A keyframes that share most vocabulary words
    Wa = covisibility(A)
    def st_transform(S):
        t = S.transform([kfc["desc"]] + [k["desc"] for k in A["kfs"]])
        return [np.concatenate([x["bow_ids"], x["fv"]["nodes"], x["fv"]["start"], x["fv"]["items"]]) for x in t] + [np.concatenate([x["bow_vals"] for x in t])], t
    tr = both("vocabulary transform", st_transform, lambda n, rc, rd: cmp_exact(n, rc[0], rd[0]))[1]
    def l1_score(a, b):                             # DBoW2 L1Scoring::score on the normalised BowVectors (host code, KeyFrameDatabase.cc:203-260)
        da = dict(zip(a["bow_ids"].tolist(), a["bow_vals"].tolist()))
        return sum(abs(da[w]) + abs(v) - abs(da[w] - v) for w, v in zip(b["bow_ids"].tolist(), b["bow_vals"].tolist()) if w in da) / 2
    score = [l1_score(tr[0], t) for t in tr[1:]]
    cands = [int(i) for i in np.argsort(score)[::-1][:3]]
    out["candidates"] = cands
    # ---- SearchByBoW(pKeyFrame, pKF) per candidate (MultiMapper.cc:209), discard < 15
    def st_bow(S):
        res = []
        for c in cands:
            n, m = S.bow(kfc, hasc, tr[0]["fv"], A["kfs"][c], (A["kfs"][c]["mp"] >= 0).astype(np.uint8), tr[1 + c]["fv"])
            res.append(np.concatenate([m, [n]]))
        return res
    bow = both("SearchByBoW", st_bow, cmp_exact)
    out["bow_matches"] = [int(b[-1]) for b in bow]
    rng = np.random.default_rng(scene["seed"] + 77)
    merged = None
    for ci, c in enumerate(cands):
        if bow[ci][-1] < 15: continue
        kfm = A["kfs"][c]
        ptsm, hasm = feature_points(A, c)
        m12 = bow[ci][:-1].astype(int)             # feature of pKF matched to feature i1 of pKeyFrame
        idx1 = np.where((m12 >= 0) & (kfc["mp"] >= 0))[0]
        # ---- Sim3Solver(pKeyFrame, pKF, matches, fixScale = false), SetRansacParameters(0.99, 10, 300), iterate(5) until a result
        X1 = cam_points32(kfc["Tcw"], ptsc["Xw"][idx1]); X2 = cam_points32(kfm["Tcw"], ptsm["Xw"][m12[idx1]])
        o1 = kfc["octave"][idx1]; o2 = kfm["octave"][m12[idx1]]
        N = len(idx1)
        if N < 10: continue
        n_its = ransac_iterations(N)
        tri = sample_triples(N, n_its, rng)
        hyp = [compute_sim3(X1[t].T.copy(), X2[t].T.copy()) for t in tri]
        T12 = np.stack([h[0] for h in hyp]); T21 = np.stack([h[1] for h in hyp])
        inl, cnt = both(f"Sim3Solver::CheckInliers cand {c}", lambda S: S.sim3_check(T12, T21, X1, X2, o1, o2), cmp_exact)
        best, hit = 0, -1
        for h in range(n_its):                     # iterate(): >= best updates, > minInliers returns (Sim3Solver.cc:176-194)
            if cnt[h] >= best:
                best = int(cnt[h])
                if cnt[h] > 10: hit = h; break
        out.setdefault("ransac", []).append((c, N, n_its, hit, best))
        if hit < 0: continue
        _, _, R12, t12, s12 = hyp[hit]
        m_in = np.full(len(m12), -1, np.int32); m_in[idx1[inl[hit] > 0]] = m12[idx1[inl[hit] > 0]]
        # ---- SearchBySim3(pKeyFrame, pKF, vpMapPointMatches, s, R, t, 7.5) (:298)
        ns3, m_s3 = both(f"SearchBySim3 cand {c}", lambda S: F.search_by_sim3(S.B, kfc, kfm, s12, R12, t12, 7.5, hasc, ptsc, hasm, ptsm, m_in),
                         lambda n, rc, rd: (_same(n, rc[1], rd[1]), _same(n + " count", rc[0], rd[0])))
        # ---- OptimizeSim3(pKeyFrame, pKF, vpMapPointMatches, gScm, 10, false) (:303)
        v = ((m_s3 >= 0) & (kfc["mp"] >= 0)).astype(np.uint8)
        j = np.where(m_s3 >= 0, m_s3, 0)
        P1c = cam_points32(kfc["Tcw"], ptsc["Xw"]); P2c = cam_points32(kfm["Tcw"], ptsm["Xw"][j])
        init = sim3_row(R12.astype(np.float64), t12.astype(np.float64), float(s12))
        def cmp_s3(n, rc, rd):
            _same(n + " inliers", rc[1], rd[1]); _same(n + " count", rc[2], rd[2]); _same(n + " sim3", rc[0], rd[0], 1e-5)
        Scm, inl3, n_in = both(f"OptimizeSim3 cand {c}", lambda S: S.optimize_sim3(init, v, P1c, P2c, kfc["xy"], kfm["xy"][j], INV2[kfc["octave"]], INV2[kfm["octave"][j]]),
                               cmp_s3)
        out.setdefault("sim3_inliers", []).append((c, int(ns3), n_in))
        if n_in < 20: continue
        matched = np.where((inl3 > 0) & (v > 0), kfm["mp"][j], -1)       # mvpCurrentMatchedPoints: map-A points per feature of pKeyFrame
        # mg2oScw = gScm * gSmw
        Smw = sim3_row(kfm["Tcw"][:3, :3].astype(np.float64), kfm["Tcw"][:3, 3].astype(np.float64), 1.0)
        Scw = F.sim3_compose(Scm, Smw)
        merged = dict(c=c, Scw=Scw, matched=matched, Scm=Scm)
        break
    assert merged is not None, f"no map-merge candidate passed ({out})"
    c, Scw, matched = merged["c"], merged["Scw"], merged["matched"]
    # ---- loop map points: the matched keyframe's neighbourhood (:328-350), then SearchByProjection(pKeyFrame, mScw, loop points, matched, 10) (:353)
    neigh = [int(k) for k in np.argsort(-Wa[c], kind="stable") if Wa[c, k] >= 15] + [c]
    loop_ids, seen = [], set()
    for k in neigh:
        for p in A["kfs"][k]["mp"]:
            if p >= 0 and p not in seen: seen.add(int(p)); loop_ids.append(int(p))
    loop_ids = np.array(loop_ids)
    mpA = A["mp"]
    lpts = dict(Xw=mpA["Xw"][loop_ids], normal=mpA["normal"][loop_ids], mf_min=mpA["mf_min"][loop_ids], mf_max=mpA["mf_max"][loop_ids], desc=mpA["desc"][loop_ids])
    pos_of = {int(p): i for i, p in enumerate(loop_ids)}
    skip = np.zeros(len(loop_ids), np.uint8)
    for p in matched[matched >= 0]: skip[pos_of[int(p)]] = 1            # spAlreadyFound
    held = (matched >= 0).astype(np.uint8)
    cvScw = sim3_matrix32(Scw)
    n_more, fm = both("SearchByProjection(KF, Scw)", lambda S: F.search_kf_sim3(S.B, kfc, cvScw, 10, lpts, skip, held),
                      lambda n, rc, rd: (_same(n, rc[1], rd[1]), _same(n + " count", rc[0], rd[0])))
    matched = matched.copy()
    new = fm >= 0
    matched[new] = loop_ids[fm[new]]
    total = int((matched >= 0).sum())
    out["total_matches"] = total
    assert total >= 40, f"merge rejected: {total} matches"
    # ---- UpdatePosesAndAdd: corrected Sim3 of every keyframe of map B by propagation, map points moved, poses rewritten (:516-590)
    from scipy.spatial.transform import Rotation
    Kb = len(Bm["kfs"])
    def s_of(T): return sim3_row(np.asarray(T, np.float64)[:3, :3], np.asarray(T, np.float64)[:3, 3], 1.0)
    Twc = np.linalg.inv(kfc["Tcw"].astype(np.float64))
    corrected, noncorr = {}, {}
    for k in range(Kb):
        Tiw = Bm["kfs"][k]["Tcw"].astype(np.float64)
        corrected[k] = Scw if k == cur else F.sim3_compose(s_of(Tiw @ Twc), Scw)
        noncorr[k] = s_of(Tiw)
    ptsB = Bm["mp"]["Xw"].astype(np.float64).copy()
    done = np.zeros(len(ptsB), bool)
    for k in range(Kb):                              # std::map<KeyFrame*, Sim3> order = allocation order of the keyframes
        Swi = F.sim3_inverse(corrected[k])
        for p in Bm["kfs"][k]["mp"]:
            if p >= 0 and not done[p]:
                ptsB[p] = sim3_apply(Swi, sim3_apply(noncorr[k], ptsB[p].astype(F32).astype(np.float64))); done[p] = True
    posesB = []
    for k in range(Kb):
        S = corrected[k]
        T = np.eye(4); T[:3, :3] = Rotation.from_quat(S[:4]).as_matrix(); T[:3, 3] = S[4:7] / S[7]
        posesB.append(T.astype(F32))
    ptsB = ptsB.astype(F32)
    # ---- merged map: keyframes [A | B], points [A | B]; fusion bookkeeping = union-find "Replace" + observation lists
    Ka, Pa, Pb = len(A["kfs"]), len(mpA["Xw"]), len(ptsB)
    rep = np.arange(Pa + Pb)
    def find(p):
        while rep[p] != p: p = rep[p]
        return p
    obs = [dict(o) for o in mpA["obs"]] + [{Ka + k: i for k, i in o.items()} for o in Bm["mp"]["obs"]]
    kf_mp = [k["mp"].copy() for k in A["kfs"]] + [np.where(k["mp"] >= 0, k["mp"] + Pa, -1) for k in Bm["kfs"]]
    def replace(old, new_):                          # MapPoint::Replace (MapPoint.cc:185-231): observations move unless the keyframe already sees `new_`
        old, new_ = find(old), find(new_)
        if old == new_: return
        for k, i in obs[old].items():
            if k not in obs[new_]: obs[new_][k] = i; kf_mp[k][i] = new_
            else: kf_mp[k][i] = -1
        obs[old] = {}; rep[old] = new_
    for i in np.where(matched >= 0)[0]:              # loop fusion of the current keyframe (:593-612)
        lp = int(matched[i]); cp = kf_mp[Ka + cur][i]
        if cp >= 0: replace(cp, lp)
        else: kf_mp[Ka + cur][i] = lp; obs[lp][Ka + cur] = int(i)
    # ---- SearchAndFuse: Fuse(pKF, cvScw, loop points, 4, vpReplacePoints) for every corrected keyframe (:664-690)
    fused = 0
    fuse_log = []
    for k in range(Kb):
        kf = dict(Bm["kfs"][k])
        sk = np.array([1 if (Ka + k) in obs[find(int(p))] else 0 for p in loop_ids], np.uint8)       # pMP->IsInKeyFrame(pKF)
        slot = both(f"Fuse KF {k}", lambda S: F.fuse_search(S.B, kf, 4.0, lpts, sk, Scw=sim3_matrix32(corrected[k])), lambda n, rc, rd: _same(n, rc, rd))
        fuse_log.append(slot)
        for q in np.where(slot >= 0)[0]:
            lp = find(int(loop_ids[q])); i = int(slot[q]); cp = kf_mp[Ka + k][i]
            if cp >= 0:
                if find(cp) != lp: replace(cp, lp); fused += 1
            elif (Ka + k) not in obs[lp]:
                kf_mp[Ka + k][i] = lp; obs[lp][Ka + k] = i; fused += 1
    out["fused"] = fused
    # ---- connections after the fusion, loop connections (:699-716), MMOptimizeEssentialGraph (Optimizer.cc:1068-1346)
    Kt = Ka + Kb
    W = np.zeros((Kt, Kt), int)
    for p in range(Pa + Pb):
        ks = sorted(obs[p])
        for a in range(len(ks)):
            for b in range(a + 1, len(ks)): W[ks[a], ks[b]] += 1; W[ks[b], ks[a]] += 1
    Wb = covisibility(Bm)
    loopconn = {}
    for k in range(Kb):
        s = {int(j) for j in range(Ka) if W[Ka + k, j] >= 15}                                     # new connections that are neither old neighbours nor map-B keyframes
        if s: loopconn[Ka + k] = s
    vScw = np.zeros((Kt, 8)); nonc = np.zeros((Kt, 8))
    for k in range(Ka): vScw[k] = nonc[k] = s_of(A["kfs"][k]["Tcw"])
    for k in range(Kb): vScw[Ka + k] = corrected[k]; nonc[Ka + k] = noncorr[k]
    parent = [-1] + list(range(Ka - 1)) + [-1] + [Ka + k for k in range(Kb - 1)]                   # spanning trees of the two maps (first connection = previous keyframe)
    ei, ej, meas, ins = [], [], [], set()
    curg, loopg = Ka + cur, c
    for i in sorted(loopconn):
        Swi = F.sim3_inverse(vScw[i])
        for j in sorted(loopconn[i]):
            if (i != curg or j != loopg) and W[i, j] < 100: continue
            ei.append(i); ej.append(j); meas.append(F.sim3_compose(vScw[j], Swi)); ins.add((min(i, j), max(i, j)))
    for i in range(Kt):
        Swi = F.sim3_inverse(nonc[i])
        p = parent[i]
        if p >= 0: ei.append(i); ej.append(p); meas.append(F.sim3_compose(nonc[p], Swi))
        children = {k for k in range(Kt) if parent[k] == i}
        for wgt, j in sorted([(W[i, j], j) for j in range(Kt) if j != i and W[i, j] >= 100], key=lambda x: (-x[0], x[1])):
            if j != p and j not in children and j < i and (min(i, j), max(i, j)) not in ins:
                ei.append(i); ej.append(j); meas.append(F.sim3_compose(nonc[j], Swi))
    fixed = np.zeros(Kt, np.uint8); fixed[loopg] = 1
    ei = np.array(ei, np.int32); ej = np.array(ej, np.int32); meas = np.array(meas)
    out["essential_edges"] = len(ei); out["loop_connections"] = sum(len(v) for v in loopconn.values())
    def cmp_pg(n, rc, rd):
        _same(n + " LM iterations", rc["lm_iterations"], rd["lm_iterations"]); _same(n, rc["sim3"], rd["sim3"], 1e-5)
    pg = both("MMOptimizeEssentialGraph", lambda S: S.pose_graph(vScw, fixed, ei, ej, meas), cmp_pg)
    est = pg["sim3"]
    poses = np.zeros((Kt, 4, 4), F32)
    for k in range(Kt):
        T = np.eye(4); T[:3, :3] = Rotation.from_quat(est[k][:4]).as_matrix(); T[:3, 3] = est[k][4:7] / est[k][7]
        poses[k] = T.astype(F32)
    pts = np.concatenate([mpA["Xw"], ptsB]).astype(np.float64)
    ref = np.concatenate([mpA["ref"], Bm["mp"]["ref"] + Ka])
    for p in range(Pa + Pb):                         # map points follow their reference keyframe (Optimizer.cc:1318-1343)
        r = int(ref[p])
        pts[p] = sim3_apply(F.sim3_inverse(est[r]), sim3_apply(vScw[r], pts[p]))
    # ---- MMGlobalBundleAdjustemnt(pMap, 20, stop, nLoopKF, false) (LoopClosing.cc:818): every keyframe and live point of both maps, keyframe 0 of the base map fixed
    live = np.array([p for p in range(Pa + Pb) if rep[p] == p and len(obs[p]) >= 1])
    new_id = np.full(Pa + Pb, -1); new_id[live] = np.arange(len(live))
    e_kf, e_pt, e_uv, e_w, origin = [], [], [], [], []
    allk = A["kfs"] + Bm["kfs"]
    for p in live:
        for k, i in sorted(obs[p].items()):
            e_kf.append(k); e_pt.append(new_id[p]); e_uv.append(allk[k]["xy"][i]); e_w.append(INV2[allk[k]["octave"][i]])
    fx = np.zeros(Kt, np.uint8); fx[0] = 1
    g = dict(poses=poses, fixed=fx, intr=np.tile(K4.astype(np.float64), (Kt, 1)), points=pts[live].astype(F32), kf=np.array(e_kf, np.int32), pt=np.array(e_pt, np.int32),
             uv=np.array(e_uv, F32), inv_sigma2=np.array(e_w, F32), origin=(live >= Pa).astype(np.int32))
    out["gba_graph"] = g
    def cmp_ba(n, rc, rd):
        _same(n + " LM iterations", rc["lm_iterations"], rd["lm_iterations"])
        _same(n + " poses", rc["poses"], rd["poses"], 1e-5); _same(n + " points", rc["points"], rd["points"], 1e-5)
    ba = both("MMGlobalBundleAdjustemnt", lambda S: S.global_ba(g, gba_opt), cmp_ba)
    out["gba"] = ba
    out["poses"], out["points"] = ba["poses"], ba["points"]
    out["merged"] = merged
    return out


def centre_error_vs_truth(scene, poses):
    """largest distance (map-A units = world metres) between the merged map's keyframe centres and the scene's ground truth: sanity of the whole composition"""
    truth = np.concatenate([scene["A"]["centres_world"], scene["B"]["centres_world"]])
    c = np.stack([-T[:3, :3].astype(np.float64).T @ T[:3, 3].astype(np.float64) for T in poses])
    return float(np.linalg.norm(c - truth, axis=1).max())


def owner_by_origin(origin, world):
    """rank that owns a point of the merged map: the GPUs are split between the two robots (half each; one GPU: everything on it), a robot's points are dealt
    round-robin over its GPUs"""
    origin = np.asarray(origin)
    if world == 1:
        return np.zeros(len(origin), np.int64)
    half = world // 2
    idx = np.zeros(len(origin), np.int64)
    for o in (0, 1):
        sel = np.nonzero(origin == o)[0]
        n = half if o == 0 else world - half
        idx[sel] = (0 if o == 0 else half) + np.arange(len(sel)) % n
    return idx
