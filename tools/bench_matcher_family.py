"""Timing of the KeyFrame / Sim3 / vocabulary entry points added after the headline path: batched C-ABI calls with HOST buffers (copies in the
timed region) on the GPU next to the reference's own object code (oracle/_ref/libref_orbmatcher.so, one call per keyframe pair, 1 thread) or
the oracle port where the reference cannot be built (OptimizeSim3, DBoW2 transform).  Writes gpurun_out/matcher_family.json.
Not the driver's bench (bench.py); the numbers go into DESIGN.md section 7."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from oracle import ref_build  # noqa: E402
import orbslamm_b200 as ob  # noqa: E402
from orbslamm_b200 import synth, vocabulary as V  # noqa: E402
import kf_family as kff  # noqa: E402
from helpers import slab  # noqa: E402

B = int(os.environ.get("PAIRS", "32"))
REP = int(os.environ.get("REP", "5"))
HAVE_REF = ref_build.matcher_available()


def timed(fn, rep=REP):
    fn()
    t = []
    for _ in range(rep):
        t0 = time.perf_counter(); fn(); t.append(time.perf_counter() - t0)
    return float(np.median(t))


def main():
    res = {"pairs_per_call": B, "gpu": "host buffers through the C-ABI, median of %d" % REP, "cpu": "reference object code, 1 thread" if HAVE_REF else "oracle port, 1 thread"}
    m = ob.ORBmatcher(0.75, True)
    c = kff.make_case(synth.KITTI, 2)
    kf, pts, skip, held = c["kf"], c["pts"], c["skip"], c["held"]
    N, M = len(kf["xy"]), len(skip)
    T = kf["Tcw"]
    view = ob.Projection.from_buffer_copy(bytes(oracle.make_projection(T[:3, :3], T[:3, 3], kf["K4"], kff.kf_bounds(kf), kf["log_sf"], 3.0, oracle.PROJ_CHECK_NORMAL,
                                                                         Ow=kff.camera_centre(T))))
    rep = lambda a: np.repeat(np.asarray(a)[None], B, 0)
    Xw, Nn, mn, mx, val = rep(pts["Xw"]), rep(pts["normal"]), rep(pts["mf_min"]), rep(pts["mf_max"]), rep(1 - skip)
    fxy, foc, fan, fds, qds = rep(kf["xy"]), rep(kf["octave"]), rep(kf["angle"]), rep(kf["desc"]), rep(pts["desc"])
    qc = np.full(B, M, np.int32); fc = np.full(B, N, np.int32)

    def fuse_gpu():
        qv, uv, rad, l0, l1, _ = m.project_points([view] * B, c["sf"], Xw, Nn, mn, mx, qc, val)
        return m.search_best_in_window(kf["grid_bounds4"], None, fxy, foc, fds, fc, qv, uv, rad, l0, l1, qds, qc, 50, kf["inv_level_sigma2"], 5.99)
    res["Fuse(KF, vpMapPoints) search, %d points -> %d features" % (M, N)] = {"gpu_ms_per_call": 1e3 * timed(fuse_gpu), "calls_cover_keyframes": B}
    if HAVE_REF:
        res[list(res)[-1]]["cpu_ms_per_keyframe"] = 1e3 * timed(lambda: ref_build.ref_fuse_kf(kf, 3.0, pts, skip, held))

    Scw = kff.sim3_of(T, 1.37)
    R, t, Ow = kff.decompose_scw(Scw)
    view2 = ob.Projection.from_buffer_copy(bytes(oracle.make_projection(R, t, kf["K4"], kff.kf_bounds(kf), kf["log_sf"], 10.0, oracle.PROJ_CHECK_NORMAL, Ow=Ow)))
    fm_in = rep(np.where(held > 0, 1 << 30, -1).astype(np.int32))

    def loop_gpu():
        qv, uv, rad, l0, l1, _ = m.project_points([view2] * B, c["sf"], Xw, Nn, mn, mx, qc, val)
        return m.SearchByProjectionKF(kf["grid_bounds4"], None, fxy, foc, fan, fds, fc, qv, uv, rad, l0, l1, np.zeros((B, M), np.float32), qds, qc, 50, 0.0, False, fm_in)
    res["SearchByProjection(KF, Scw)"] = {"gpu_ms_per_call": 1e3 * timed(loop_gpu), "calls_cover_keyframes": B}
    if HAVE_REF:
        res["SearchByProjection(KF, Scw)"]["cpu_ms_per_keyframe"] = 1e3 * timed(lambda: ref_build.ref_search_kf_sim3(kf, Scw, 10, pts, skip, held))

    b = kff.make_bow_case(synth.KITTI, 2, True)
    k1, k2 = b["kf1"], b["kf2"]
    n1, n2 = len(k1["desc"]), len(k2["desc"])
    mb = ob.ORBmatcher(0.75, True)
    d1, a1, d2, a2 = rep(k1["desc"]), rep(k1["angle"]), rep(k2["desc"]), rep(k2["angle"])
    res["SearchByBoW(KF, KF), %d x %d features" % (n1, n2)] = {
        "gpu_ms_per_call": 1e3 * timed(lambda: mb.SearchByBoW(d1, a1, rep(b["has1"]), [n1] * B, [b["fv1"]] * B, d2, a2, rep(b["has2"]), [n2] * B, [b["fv2"]] * B)),
        "calls_cover_pairs": B}
    tri = dict(xy1=rep(k1["xy"]), xy2=rep(k2["xy"]), octave2=rep(k2["octave"]), F12=rep(b["F12"].ravel()), epipole=rep(np.array([b["ex"], b["ey"]], np.float32)),
               scale_factors2=k2["scale_factors"], level_sigma2_2=b["ls2"])
    mt = ob.ORBmatcher(0.6, False)
    res["SearchForTriangulation"] = {
        "gpu_ms_per_call": 1e3 * timed(lambda: mt.SearchByBoW(d1, a1, rep(1 - b["tri1"]), [n1] * B, [b["fv1"]] * B, d2, a2, rep(1 - b["tri2"]), [n2] * B, [b["fv2"]] * B, tri)),
        "calls_cover_pairs": B}
    if HAVE_REF:
        res[list(res)[-2]]["cpu_ms_per_pair"] = 1e3 * timed(lambda: ref_build.ref_search_by_bow_kf_kf(k1, b["has1"], b["fv1"], k2, b["has2"], b["fv2"], 0.75, True))
        res["SearchForTriangulation"]["cpu_ms_per_pair"] = 1e3 * timed(lambda: ref_build.ref_search_for_triangulation(k1, b["tri1"], b["fv1"], k2, b["tri2"], b["fv2"], b["ls2"],
                                                                                                                        b["F12"], 0.6, False))

    from helpers import make_tracking_case
    ki = make_tracking_case(synth.KITTI, 3)
    l, cu = ki["last"], ki["cur"]
    mi = ob.ORBmatcher(0.9, True)
    pm = rep(np.stack([l["x"], l["y"]], 1).astype(np.float32))
    res["SearchForInitialization, window 100"] = {
        "gpu_ms_per_call": 1e3 * timed(lambda: mi.SearchForInitialization(ki["bounds"], rep(l["octave"]), rep(l["angle"]), rep(l["desc"]), [len(l["x"])] * B,
                                                                           rep(np.stack([cu["x"], cu["y"]], 1).astype(np.float32)), rep(cu["octave"]), rep(cu["angle"]),
                                                                           rep(cu["desc"]), [len(cu["x"])] * B, pm, 100)), "calls_cover_pairs": B}
    if HAVE_REF:
        sf = np.array(list(ki["P"].scale)[:8], np.float32)
        res["SearchForInitialization, window 100"]["cpu_ms_per_pair"] = 1e3 * timed(lambda: ref_build.ref_search_for_initialization(ki["K4"], ki["bounds"], sf, l, cu, pm[0], 100,
                                                                                                                                    0.9, True))

    s = kff.make_sim3_opt_case(synth.KITTI, 2)
    o = ob.Optimizer()
    W = len(s["valid"])
    res["OptimizeSim3, %d correspondences" % int(s["valid"].sum())] = {
        "gpu_ms_per_call": 1e3 * timed(lambda: o.OptimizeSim3(rep(s["init"]), rep(s["valid"]), rep(s["P1c"]), rep(s["P2c"]), rep(s["obs1"]), rep(s["obs2"]), rep(s["w1"]),
                                                              rep(s["w2"]), rep(s["K1"]), rep(s["K2"]), [W] * B, 10.0, False)), "calls_cover_pairs": B,
        "cpu_ms_per_pair_oracle_port": 1e3 * timed(lambda: oracle.optimize_sim3(s["init"], s["valid"], s["P1c"], s["P2c"], s["obs1"], s["obs2"], s["w1"], s["w2"], s["K1"],
                                                                                 s["K2"], 10.0, False))}

    # vocabulary of ORBvoc's shape (k = 10, L = 6: 1 111 111 nodes), random contents
    rng = np.random.default_rng(0)
    k, L = 10, 6
    n = (k ** (L + 1) - 1) // (k - 1)
    parent = np.zeros(n, np.int64); parent[1:] = (np.arange(1, n) - 1) // k
    leaf = np.zeros(n, bool); leaf[n - k ** L:] = True
    desc = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    weight = np.where(leaf, rng.uniform(0.5, 9.0, n), 0.0)
    voc = V.from_nodes(k, L, parent, leaf, desc, weight)
    vh = V.ORBVocabulary(voc)
    fr = rng.integers(0, 256, (B, 2000, 32), dtype=np.uint8)
    res["DBoW2 transform (ComputeBoW), 2000 features, k=10 L=6"] = {
        "gpu_ms_per_call": 1e3 * timed(lambda: vh.transform(fr, [2000] * B, 4)), "calls_cover_frames": B,
        "cpu_ms_per_frame_oracle_port": 1e3 * timed(lambda: oracle.vocab_transform(voc, fr[0], 4), 3)}
    # Sim3Solver: 300 RANSAC hypotheses of one solver in one call (reference: one CheckInliers per hypothesis)
    r3 = kff.make_sim3_ransac_case(synth.KITTI, 2, n_hyp=300)
    m1, p1 = o.Sim3Prepare(r3["X1"], r3["oct1"], r3["ls2"], r3["K1"]); m2, p2 = o.Sim3Prepare(r3["X2"], r3["oct2"], r3["ls2"], r3["K2"])
    res["Sim3Solver::CheckInliers, 300 hypotheses x %d correspondences" % len(r3["X1"])] = {
        "gpu_ms_per_call": 1e3 * timed(lambda: o.Sim3CheckInliers(r3["T12"], r3["T21"], r3["X1"], r3["X2"], p1, p2, m1, m2, r3["K1"], r3["K2"])), "calls_cover_hypotheses": 300}
    if ref_build.sim3solver_available():
        res[list(res)[-1]]["cpu_ms_300_hypotheses_reference_object_code"] = 1e3 * timed(
            lambda: ref_build.ref_sim3_check_inliers(r3["X1"], r3["X2"], r3["oct1"], r3["oct2"], r3["ls2"], r3["K1"], r3["K2"], r3["T12"], r3["T21"]), 3)
    # essential graph: 200-keyframe loop (the oracle port factors the normal equations densely, so the CPU column is an upper bound for a sparse g2o solve)
    Sg, fxg, eig, ejg, emg, _ = kff.make_pose_graph(200, seed=1)
    res["OptimizeEssentialGraph core, 200 keyframes / %d edges" % len(eig)] = {
        "gpu_ms_per_call": 1e3 * timed(lambda: o.OptimizePoseGraph(Sg, fxg, eig, ejg, emg, False, 20, 1e-16), 3),
        "cpu_ms_oracle_port_dense_ldlt": 1e3 * timed(lambda: oracle.optimize_pose_graph(Sg, fxg, eig, ejg, emg, False, 20, 1e-16), 1)}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "matcher_family.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
