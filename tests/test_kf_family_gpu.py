"""GPU parity of the KeyFrame / Sim3 members of ORBmatcher (ORBmatcher.cc:292-405, 827-977, 979-1102, 1104-1328, 1474-1601): the CUDA halves
(orbm_project_points, orbm_search_best_in_window, orbm_search_by_projection_kf) composed as the drop-in composes them, against the oracle
and -- where oracle/_ref travelled with the snapshot -- against the reference's own object code.  All results are exact (indices, counts)."""
import numpy as np
import pytest

import oracle
from oracle import ref_build
from orbslamm_b200 import synth
import kf_family as kff
from helpers import slab

pytestmark = pytest.mark.gpu
HAVE_REF = ref_build.matcher_available()


@pytest.mark.parametrize("cam,sid,dist", [("TUM", 1, False), ("KITTI", 2, False), ("TUM", 3, True), ("KITTI", 5, True)])
def test_kf_family_cuda_equals_oracle_and_reference(lib, cam, sid, dist):
    c = kff.make_case(getattr(synth, cam), sid, distorted_bounds=dist)
    Bo, Bc = kff.OracleBackend(), kff.CudaBackend()
    kf, pts, skip, held, k = c["kf"], c["pts"], c["skip"], c["held"], c["k"]
    Scw = kff.sim3_of(kf["Tcw"], 1.37)
    # projection half alone, every flag combination the five members use: bit-exact float outputs
    Rcw, tcw, Ow = kff.decompose_scw(Scw)
    P = oracle
    for flags, kw in ((P.PROJ_CHECK_NORMAL, dict(Ow=Ow)),
                      (P.PROJ_NO_DEPTH | P.PROJ_FRAME_BOUNDS | P.PROJ_FRAME_UV | P.PROJ_LEVEL_PLUS1, dict(Ow=Ow, use_normal=False)),
                      (P.PROJ_TWO_STEP | P.PROJ_DIST_CAMERA, dict(R2=np.eye(3, dtype=np.float32) * np.float32(0.98), t2=np.array([0.01, -0.02, 0.03], np.float32),
                                                                 use_normal=False))):
        a = Bo.project(Rcw, tcw, kf["K4"], kff.kf_bounds(kf), kf["log_sf"], np.float32(4.0), flags, kf["scale_factors"], pts, 1 - skip, **kw)
        b = Bc.project(Rcw, tcw, kf["K4"], kff.kf_bounds(kf), kf["log_sf"], np.float32(4.0), flags, kf["scale_factors"], pts, 1 - skip, **kw)
        assert a[0].sum() > 100
        for x, y in zip(a, b):
            v = a[0] > 0
            assert np.array_equal(x[v], y[v])
        assert np.array_equal(a[0], b[0])
    # SearchByProjection(pKF, Scw, vpPoints, vpMatched, th)
    n_o, fm_o = kff.search_kf_sim3(Bo, kf, Scw, 10, pts, skip, held)
    n_c, fm_c = kff.search_kf_sim3(Bc, kf, Scw, 10, pts, skip, held)
    assert n_o > 30 and n_c == n_o and np.array_equal(fm_c, fm_o)
    if HAVE_REF:
        n_r, fm_r = ref_build.ref_search_kf_sim3(kf, Scw, 10, pts, skip, held)
        assert n_c == n_r and np.array_equal(fm_c, fm_r)
    # Fuse (both variants)
    for th, S in ((3.0, None), (6.0, None), (4.0, Scw)):
        so, sc = kff.fuse_search(Bo, kf, th, pts, skip, Scw=S), kff.fuse_search(Bc, kf, th, pts, skip, Scw=S)
        assert (so >= 0).sum() > 30 and np.array_equal(so, sc)
        if HAVE_REF:
            n_r, sr = (ref_build.ref_fuse_kf(kf, th, pts, skip, held) if S is None else ref_build.ref_fuse_sim3(kf, S, th, pts, skip, held))
            assert np.array_equal(sc, sr) and n_r == int((sc >= 0).sum())
    # SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist)
    has = (np.random.default_rng(3).random(len(skip)) < 0.85).astype(np.uint8)
    cur = dict(kf); cur["grid_bounds4"] = k["bounds"].astype(np.float32)
    for th, od, ori in ((10.0, 100, True), (3.0, 64, True), (10.0, 100, False)):
        a = kff.search_frame_kf(Bo, cur, k["Tcw"], k["K4"], k["bounds"], kf["log_sf"], c["sf"], held, has, skip, pts, c["last"]["angle"], th, od, ori)
        b = kff.search_frame_kf(Bc, cur, k["Tcw"], k["K4"], k["bounds"], kf["log_sf"], c["sf"], held, has, skip, pts, c["last"]["angle"], th, od, ori)
        assert a[0] > 20 and a[0] == b[0] and np.array_equal(a[1], b[1])
        if HAVE_REF:
            r = ref_build.ref_search_frame_kf(k["K4"], k["bounds"], k["Tcw"], c["sf"], c["cur"], held, has, skip, pts, c["last"]["angle"], th, od, ori)
            assert r[0] == b[0] and np.array_equal(r[1], b[1])


@pytest.mark.parametrize("cam,sid", [("TUM", 1), ("KITTI", 2)])
def test_search_by_sim3_cuda_equals_oracle_and_reference(lib, cam, sid):
    p = kff.make_sim3_pair(getattr(synth, cam), sid)
    Bo, Bc = kff.OracleBackend(), kff.CudaBackend()
    for th in (7.5, 3.0):
        args = (p["kf1"], p["kf2"], p["s12"], p["R12"], p["t12"], th, p["has1"], p["pts1"], p["has2"], p["pts2"], p["m12"])
        n_o, m_o = kff.search_by_sim3(Bo, *args)
        n_c, m_c = kff.search_by_sim3(Bc, *args)
        assert n_o > 20 and n_c == n_o and np.array_equal(m_c, m_o)
        if HAVE_REF:
            n_r, m_r = ref_build.ref_search_by_sim3(*args)
            assert n_c == n_r and np.array_equal(m_c, m_r)


def test_kf_family_batched_views_and_edge_cases(lib):
    """Several keyframes per launch (ragged counts, slab layout), an empty candidate list, and a keyframe without features."""
    import orbslamm_b200 as ob
    cases = [kff.make_case(synth.TUM, sid) for sid in (1, 2, 4)]
    m = ob.ORBmatcher()
    Bo = kff.OracleBackend()
    nQ = max(len(c["skip"]) for c in cases) + 5; nF = max(len(c["kf"]["xy"]) for c in cases) + 3
    views, refs = [], []
    for c in cases:
        kf = c["kf"]; T = kf["Tcw"]
        o = oracle.make_projection(T[:3, :3], T[:3, 3], kf["K4"], kff.kf_bounds(kf), kf["log_sf"], 3.0, oracle.PROJ_CHECK_NORMAL, Ow=kff.camera_centre(T))
        views.append(ob.Projection.from_buffer_copy(bytes(o)))
        refs.append(Bo.project(T[:3, :3], T[:3, 3], kf["K4"], kff.kf_bounds(kf), kf["log_sf"], np.float32(3.0), oracle.PROJ_CHECK_NORMAL, kf["scale_factors"],
                               c["pts"], 1 - c["skip"], Ow=kff.camera_centre(T)))
    qc = np.array([len(c["skip"]) for c in cases], np.int32); fc = np.array([len(c["kf"]["xy"]) for c in cases], np.int32)
    qc[2] = 0                                                # third keyframe: no candidates at all
    X = slab([c["pts"]["Xw"] for c in cases], nQ, np.float32, (3,)); Nn = slab([c["pts"]["normal"] for c in cases], nQ, np.float32, (3,))
    mn = slab([c["pts"]["mf_min"] for c in cases], nQ, np.float32); mx = slab([c["pts"]["mf_max"] for c in cases], nQ, np.float32)
    val = slab([1 - c["skip"] for c in cases], nQ, np.uint8)
    qv, uv, rad, l0, l1, lv = m.project_points(views, cases[0]["sf"], X, Nn, mn, mx, qc, val)
    for i in (0, 1):
        n = qc[i]; v = refs[i][0] > 0
        assert np.array_equal(qv[i, :n], refs[i][0]) and np.array_equal(uv[i, :n][v], refs[i][1][v]) and np.array_equal(rad[i, :n][v], refs[i][2][v])
        assert np.array_equal(l0[i, :n][v], refs[i][3][v]) and np.array_equal(l1[i, :n][v], refs[i][4][v]) and np.array_equal(lv[i, :n][v], refs[i][5][v])
    fxy = slab([c["kf"]["xy"] for c in cases], nF, np.float32, (2,)); foc = slab([c["kf"]["octave"] for c in cases], nF, np.int32)
    fds = slab([c["kf"]["desc"] for c in cases], nF, np.uint8, (32,)); qds = slab([c["pts"]["desc"] for c in cases], nQ, np.uint8, (32,))
    fc2 = fc.copy(); fc2[1] = 0                              # second keyframe: no features
    bi, bd = m.search_best_in_window(cases[0]["kf"]["grid_bounds4"], None, fxy, foc, fds, fc2, qv, uv, rad, l0, l1, qds, qc, 50,
                                     cases[0]["kf"]["inv_level_sigma2"], 5.99)
    so = kff.fuse_search(Bo, cases[0]["kf"], 3.0, cases[0]["pts"], cases[0]["skip"])
    assert np.array_equal(bi[0, :qc[0]], so) and (bi[1, :qc[1]] == -1).all() and (bd[1, :qc[1]] == -1).all()


@pytest.mark.parametrize("ori", [True, False])
def test_bow_family_cuda_equals_oracle_and_reference(lib, ori):
    """SearchByBoW (both variants) and SearchForTriangulation: three keyframe pairs per launch (ragged), exact match arrays and counts."""
    import orbslamm_b200 as ob
    cases = [kff.make_bow_case(synth.TUM, 1, False), kff.make_bow_case(synth.KITTI, 2, True), kff.make_bow_case(synth.TUM, 4, True)]
    n = len(cases)
    s1 = max(len(b["kf1"]["desc"]) for b in cases) + 2; s2 = max(len(b["kf2"]["desc"]) for b in cases) + 5
    d1 = slab([b["kf1"]["desc"] for b in cases], s1, np.uint8, (32,)); d2 = slab([b["kf2"]["desc"] for b in cases], s2, np.uint8, (32,))
    a1 = slab([b["kf1"]["angle"] for b in cases], s1, np.float32); a2 = slab([b["kf2"]["angle"] for b in cases], s2, np.float32)
    c1 = [len(b["kf1"]["desc"]) for b in cases]; c2 = [len(b["kf2"]["desc"]) for b in cases]
    fv1 = [b["fv1"] for b in cases]; fv2 = [b["fv2"] for b in cases]
    # SearchByBoW(KF, KF)
    m = ob.ORBmatcher(0.75, ori)
    nm, mt = m.SearchByBoW(d1, a1, slab([b["has1"] for b in cases], s1, np.uint8), c1, fv1, d2, a2, slab([b["has2"] for b in cases], s2, np.uint8), c2, fv2)
    for k, b in enumerate(cases):
        n_o, m_o = oracle.search_by_bow(0, b["kf1"]["desc"], b["kf1"]["angle"], b["has1"], b["fv1"], b["kf2"]["desc"], b["kf2"]["angle"], b["has2"], b["fv2"], 0.75, ori)
        assert n_o > 100 and nm[k] == n_o and np.array_equal(mt[k, :c1[k]], m_o)
        if HAVE_REF:
            n_r, m_r = ref_build.ref_search_by_bow_kf_kf(b["kf1"], b["has1"], b["fv1"], b["kf2"], b["has2"], b["fv2"], 0.75, ori)
            assert nm[k] == n_r and np.array_equal(mt[k, :c1[k]], m_r)
    # SearchByBoW(KF, Frame): every frame feature is a candidate
    m = ob.ORBmatcher(0.7, ori)
    nm, mt = m.SearchByBoW(d1, a1, slab([b["has1"] for b in cases], s1, np.uint8), c1, fv1, d2, a2, np.ones((n, s2), np.uint8), c2, fv2)
    for k, b in enumerate(cases):
        n_o, m_o = oracle.search_by_bow(0, b["kf1"]["desc"], b["kf1"]["angle"], b["has1"], b["fv1"], b["kf2"]["desc"], b["kf2"]["angle"], np.ones(c2[k], np.uint8),
                                        b["fv2"], 0.7, ori)
        assert nm[k] == n_o and np.array_equal(mt[k, :c1[k]], m_o)
        if HAVE_REF:
            n_r, m_r = ref_build.ref_search_by_bow_kf_frame(b["kf1"], b["has1"], b["fv1"], b["kf2"]["angle"], b["kf2"]["desc"], b["fv2"], 0.7, ori)
            inv = np.full(c2[k], -1, np.int32); g = mt[k, :c1[k]]; inv[g[g >= 0]] = np.where(g >= 0)[0]
            assert nm[k] == n_r and np.array_equal(inv, m_r)
    # SearchForTriangulation
    m = ob.ORBmatcher(0.6, ori)
    tri = dict(xy1=slab([b["kf1"]["xy"] for b in cases], s1, np.float32, (2,)), xy2=slab([b["kf2"]["xy"] for b in cases], s2, np.float32, (2,)),
               octave2=slab([b["kf2"]["octave"] for b in cases], s2, np.int32), F12=np.stack([b["F12"].ravel() for b in cases]),
               epipole=np.array([[b["ex"], b["ey"]] for b in cases], np.float32), scale_factors2=cases[0]["kf2"]["scale_factors"], level_sigma2_2=cases[0]["ls2"])
    nm, mt = m.SearchByBoW(d1, a1, slab([1 - b["tri1"] for b in cases], s1, np.uint8), c1, fv1, d2, a2, slab([1 - b["tri2"] for b in cases], s2, np.uint8), c2, fv2, tri)
    for k, b in enumerate(cases):
        epi = dict(xy1=b["kf1"]["xy"], xy2=b["kf2"]["xy"], octave2=b["kf2"]["octave"], F12=b["F12"], ex=b["ex"], ey=b["ey"], scale_factors2=b["kf2"]["scale_factors"],
                   level_sigma2_2=b["ls2"])
        n_o, m_o = oracle.search_by_bow(1, b["kf1"]["desc"], b["kf1"]["angle"], 1 - b["tri1"], b["fv1"], b["kf2"]["desc"], b["kf2"]["angle"], 1 - b["tri2"], b["fv2"],
                                        0.6, ori, epi)
        assert n_o > 5 and nm[k] == n_o and np.array_equal(mt[k, :c1[k]], m_o)
        if HAVE_REF:
            n_r, m_r = ref_build.ref_search_for_triangulation(b["kf1"], b["tri1"], b["fv1"], b["kf2"], b["tri2"], b["fv2"], b["ls2"], b["F12"], 0.6, ori)
            assert nm[k] == n_r and np.array_equal(mt[k, :c1[k]], m_r)


def test_search_for_initialization_cuda_equals_oracle_and_reference(lib):
    import orbslamm_b200 as ob
    from helpers import make_tracking_case
    cases = [make_tracking_case(synth.TUM, 1), make_tracking_case(synth.TUM, 6), make_tracking_case(synth.TUM, 9, nfeatures=2000)]
    s1 = max(len(k["last"]["x"]) for k in cases) + 1; s2 = max(len(k["cur"]["x"]) for k in cases) + 4
    g = oracle.grid_params(*cases[0]["bounds"]); sf = np.array(list(cases[0]["P"].scale)[:8], np.float32)
    for win, ori in ((100, True), (100, False), (20, True)):
        m = ob.ORBmatcher(0.9, ori)
        pm = slab([np.stack([k["last"]["x"], k["last"]["y"]], 1) for k in cases], s1, np.float32, (2,))
        nm, mt, pm2 = m.SearchForInitialization(cases[0]["bounds"], slab([k["last"]["octave"] for k in cases], s1, np.int32),
                                                slab([k["last"]["angle"] for k in cases], s1, np.float32), slab([k["last"]["desc"] for k in cases], s1, np.uint8, (32,)),
                                                [len(k["last"]["x"]) for k in cases],
                                                slab([np.stack([k["cur"]["x"], k["cur"]["y"]], 1) for k in cases], s2, np.float32, (2,)),
                                                slab([k["cur"]["octave"] for k in cases], s2, np.int32), slab([k["cur"]["angle"] for k in cases], s2, np.float32),
                                                slab([k["cur"]["desc"] for k in cases], s2, np.uint8, (32,)), [len(k["cur"]["x"]) for k in cases], pm, win)
        for i, k in enumerate(cases):
            n1 = len(k["last"]["x"])
            a = oracle.search_for_initialization(g, k["last"], k["cur"], pm[i, :n1], win, 0.9, ori)
            assert a[0] > 50 and nm[i] == a[0] and np.array_equal(mt[i, :n1], a[1]) and np.array_equal(pm2[i, :n1], a[2])
            if HAVE_REF:
                b = ref_build.ref_search_for_initialization(k["K4"], k["bounds"], sf, k["last"], k["cur"], pm[i, :n1], win, 0.9, ori)
                assert nm[i] == b[0] and np.array_equal(mt[i, :n1], b[1]) and np.array_equal(pm2[i, :n1], b[2])


def test_vocabulary_transform_cuda_equals_oracle(lib):
    """orbv_transform (Frame::ComputeBoW): per-feature words / nodes, BowVector (ids exact, fp64 values bit-exact: same summation order) and
    FeatureVector for a ragged batch of frames; stopped words, sibling ties and levelsup variants; the FeatureVector feeds SearchByBoW."""
    import orbslamm_b200 as ob
    from orbslamm_b200 import vocabulary as V
    from helpers import make_tracking_case
    v = V.synthetic(10, 4, seed=3)
    voc = V.ORBVocabulary(v)
    rng = np.random.default_rng(2)
    leaves = np.where(v["word_id"] >= 0)[0]
    k = make_tracking_case(synth.TUM, 5)
    frames = [k["cur"]["desc"], k["last"]["desc"],
              v["node_desc"][rng.choice(leaves, 700)] ^ np.packbits(rng.random((700, 256)) < 0.06, axis=1), np.zeros((0, 32), np.uint8)]
    W = max(len(f) for f in frames) + 9
    for levelsup in (4, 2, 0, 7):
        out = voc.transform(slab(frames, W, np.uint8, (32,)), [len(f) for f in frames], levelsup)
        for f, o in zip(frames, out):
            r = oracle.vocab_transform(v, f, levelsup)
            assert np.array_equal(o["word_of"], r["word_of"]) and np.array_equal(o["node_of"], r["node_of"])
            assert np.array_equal(o["bow_ids"], r["bow_ids"]) and np.array_equal(o["bow_vals"], r["bow_vals"])
            for key in ("nodes", "start", "items"):
                assert np.array_equal(o["fv"][key], r["fv"][key])
    if ref_build.dbow2_available():                      # and directly against the reference's own DBoW2 object code
        import tempfile, os
        path = os.path.join(tempfile.mkdtemp(), "voc.txt")
        V.save_text(v, path)
        R = ref_build.RefVocabulary(path)
        out = voc.transform(slab(frames, W, np.uint8, (32,)), [len(f) for f in frames], 4)
        for f, o in zip(frames, out):
            r = R.transform(f, 4)
            assert np.array_equal(o["bow_ids"], r["bow_ids"]) and np.array_equal(o["bow_vals"], r["bow_vals"])
            for key in ("nodes", "start", "items"):
                assert np.array_equal(o["fv"][key], r["fv"][key])
    # the device FeatureVectors drive SearchByBoW (Tracking::TrackReferenceKeyFrame: ComputeBoW, then SearchByBoW(KF, F))
    out = voc.transform(slab(frames[:2], W, np.uint8, (32,)), [len(frames[0]), len(frames[1])], 2)
    m = ob.ORBmatcher(0.7, True)
    has = (rng.random(len(frames[1])) < 0.9).astype(np.uint8)
    nm, mt = m.SearchByBoW(frames[1][None], k["last"]["angle"][None], has[None], [len(frames[1])], [out[1]["fv"]],
                           frames[0][None], k["cur"]["angle"][None], np.ones((1, len(frames[0])), np.uint8), [len(frames[0])], [out[0]["fv"]])
    n_o, m_o = oracle.search_by_bow(0, frames[1], k["last"]["angle"], has, out[1]["fv"], frames[0], k["cur"]["angle"], np.ones(len(frames[0]), np.uint8), out[0]["fv"],
                                    0.7, True)
    assert n_o > 20 and nm[0] == n_o and np.array_equal(mt[0], m_o)


def test_new_entry_points_reject_bad_arguments(lib):
    """No exception crosses the C boundary: null pointers, missing optional blocks and oversize slabs return ORBS_E_INVALID with a message."""
    import ctypes
    import orbslamm_b200 as ob
    m = ob.ORBmatcher()
    h = m.handle
    L = lib
    z = np.zeros(64, np.int32)
    assert L.orbm_project_points(h, 1, None, None, 8, None, None, None, None, None, 4, None, None, None, None, None, None, 0) == -1
    assert L.orbm_search_best_in_window(h, 1, None, None, None, None, None, None, 4, None, None, None, None, None, None, None, 4, 50, None, 0, 5.99, None, None, 0) == -1
    assert L.orbm_search_by_bow(h, 1, 7, *([None] * 4), 4, *([None] * 4), 4, *([None] * 4), 4, *([None] * 4), 4, 0.7, 0, None, None, None, 0) == -1
    assert L.orbm_search_for_initialization(h, 0, *([None] * 5), 4, *([None] * 5), 4, None, 100, 0.9, 1, None, None, 0) == -1
    assert L.orbm_get_features_in_area(h, 1, None, None, None, None, None, 4, None, None, None, None, 4, 0, None, None, 0) == -1
    assert L.orbm_assign_features_to_grid(h, 1, None, None, None, 4, None, None, 0) == -1
    o = ob.Optimizer()
    assert L.orbo_optimize_sim3(o.handle, 1, *([None] * 11), 4, 10.0, 0, None, None, None, 0) == -1
    assert b"" != L.orbs_last_error()
    # triangulation mode without the epipolar block
    d = np.zeros((4, 32), np.uint8); e = np.ones(4, np.uint8); c = np.array([4], np.int32); nodes = np.array([1], np.int32); st = np.array([0, 4], np.int32)
    it = np.arange(4, dtype=np.int32); out = np.zeros(4, np.int32); nm = np.zeros(1, np.int32)
    p = lambda a: a.ctypes.data
    assert L.orbm_search_by_bow(h, 1, 1, p(d), None, p(e), p(c), 4, p(nodes), p(st), p(it), p(c), 1, p(d), None, p(e), p(c), 4, p(nodes), p(st), p(it), p(c), 1,
                                0.6, 0, None, p(out), p(nm), 0) == -1
    # ... and a well-formed minimal call succeeds (identical descriptors in one node: best == second -> the ratio test rejects)
    nc = np.array([1], np.int32)
    assert L.orbm_search_by_bow(h, 1, 0, p(d), None, p(e), p(c), 4, p(nodes), p(st), p(it), p(nc), 1, p(d), None, p(e), p(c), 4, p(nodes), p(st), p(it), p(nc), 1,
                                0.6, 0, None, p(out), p(nm), 0) == 0
    assert nm[0] == 0 and (out == -1).all()


def test_device_resident_chain_equals_host_path(lib):
    """ORBS_MEM_DEVICE: orbm_project_points -> orbm_search_best_in_window and orbv_transform -> orbm_search_by_bow chained on device buffers (no host
    copies between the calls, one stream) give exactly what the host-buffer calls give."""
    import ctypes
    import torch
    import orbslamm_b200 as ob
    from orbslamm_b200 import vocabulary as V
    c = kff.make_case(synth.KITTI, 2)
    kf, pts, skip = c["kf"], c["pts"], c["skip"]
    Bc = kff.CudaBackend()
    host = kff.fuse_search(Bc, kf, 3.0, pts, skip)                          # host-buffer path (already checked against the oracle)
    m = ob.ORBmatcher()
    h, L = m.handle, lib
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    ptr = lambda t: ctypes.c_void_p(t.data_ptr())
    N, M = len(kf["xy"]), len(skip)
    T = kf["Tcw"]
    view = ob.Projection.from_buffer_copy(bytes(oracle.make_projection(T[:3, :3], T[:3, 3], kf["K4"], kff.kf_bounds(kf), kf["log_sf"], 3.0, oracle.PROJ_CHECK_NORMAL,
                                                                         Ow=kff.camera_centre(T))))
    d_sf, d_X, d_N, d_mn, d_mx = dev(c["sf"]), dev(pts["Xw"]), dev(pts["normal"]), dev(pts["mf_min"]), dev(pts["mf_max"])
    d_qc, d_fc = dev(np.array([M], np.int32)), dev(np.array([N], np.int32))
    d_val = dev((1 - skip).astype(np.uint8))
    d_uv = torch.zeros(M, 2, dtype=torch.float32, device="cuda"); d_rad = torch.zeros(M, dtype=torch.float32, device="cuda")
    d_l0 = torch.zeros(M, dtype=torch.int32, device="cuda"); d_l1 = torch.zeros(M, dtype=torch.int32, device="cuda")
    d_fxy, d_foc, d_fd, d_qd = dev(kf["xy"]), dev(kf["octave"]), dev(kf["desc"]), dev(pts["desc"])
    d_bi = torch.zeros(M, dtype=torch.int32, device="cuda"); d_bd = torch.zeros(M, dtype=torch.int32, device="cuda")
    gb = np.ascontiguousarray(kf["grid_bounds4"], np.float32); wo = np.ascontiguousarray(kf["win_origin2"], np.float32)
    inv = np.ascontiguousarray(kf["inv_level_sigma2"], np.float32)
    assert L.orbm_project_points(h, 1, ctypes.byref(view), ptr(d_sf), 8, ptr(d_X), ptr(d_N), ptr(d_mn), ptr(d_mx), ptr(d_qc), M, ptr(d_val), ptr(d_uv), ptr(d_rad),
                                 ptr(d_l0), ptr(d_l1), None, 1) == 0
    assert L.orbm_search_best_in_window(h, 1, gb.ctypes.data, wo.ctypes.data, ptr(d_fxy), ptr(d_foc), ptr(d_fd), ptr(d_fc), N, ptr(d_val), ptr(d_uv), ptr(d_rad),
                                        ptr(d_l0), ptr(d_l1), ptr(d_qd), ptr(d_qc), M, 50, inv.ctypes.data, 8, 5.99, ptr(d_bi), ptr(d_bd), 1) == 0
    m.synchronize()
    assert np.array_equal(d_bi.cpu().numpy(), host)
    # vocabulary transform of two frames on the device, its FeatureVectors straight into SearchByBoW
    v = V.synthetic(10, 3, seed=8)
    voc = V.ORBVocabulary(v)
    b = kff.make_bow_case(synth.TUM, 1)
    k1, k2 = b["kf1"], b["kf2"]
    S = max(len(k1["desc"]), len(k2["desc"]))
    desc = slab([k1["desc"], k2["desc"]], S, np.uint8, (32,)); cnt = np.array([len(k1["desc"]), len(k2["desc"])], np.int32)
    ho = voc.transform(desc, cnt, 2)                                         # host-buffer path
    d_desc, d_cnt = dev(desc), dev(cnt)
    i32 = lambda *s: torch.zeros(*s, dtype=torch.int32, device="cuda")
    d_bi2, d_bc, d_fn, d_fs, d_fi, d_fcn = i32(2, S), i32(2), i32(2, S), i32(2, S + 1), i32(2, S), i32(2)
    d_bv = torch.zeros(2, S, dtype=torch.float64, device="cuda")
    assert L.orbv_transform(voc._h, 2, ptr(d_desc), ptr(d_cnt), S, 2, None, None, ptr(d_bi2), ptr(d_bv), ptr(d_bc), ptr(d_fn), ptr(d_fs), ptr(d_fi), ptr(d_fcn), 1) == 0
    torch.cuda.synchronize()
    ang = dev(slab([k1["angle"], k2["angle"]], S, np.float32)); el = dev(slab([b["has1"], np.ones(len(k2["desc"]), np.uint8)], S, np.uint8))
    d_m = i32(S); d_nm = i32(1)
    # the vocabulary handle has its own stream: the matcher call is issued after the synchronise above
    assert L.orbm_search_by_bow(h, 1, 0, ptr(d_desc[0]), ptr(ang[0]), ptr(el[0]), ptr(d_cnt[0:1]), S, ptr(d_fn[0]), ptr(d_fs[0]), ptr(d_fi[0]), ptr(d_fcn[0:1]), S,
                                ptr(d_desc[1]), ptr(ang[1]), ptr(el[1]), ptr(d_cnt[1:2]), S, ptr(d_fn[1]), ptr(d_fs[1]), ptr(d_fi[1]), ptr(d_fcn[1:2]), S,
                                0.7, 1, None, ptr(d_m), ptr(d_nm), 1) == 0
    m.synchronize()
    n_o, m_o = oracle.search_by_bow(0, k1["desc"], k1["angle"], b["has1"], ho[0]["fv"], k2["desc"], k2["angle"], np.ones(len(k2["desc"]), np.uint8), ho[1]["fv"], 0.7, True)
    assert n_o > 50 and int(d_nm.cpu()[0]) == n_o and np.array_equal(d_m.cpu().numpy()[:len(k1["desc"])], m_o)


def test_degenerate_inputs(lib):
    """Empty and disjoint inputs: keyframes without features or without common vocabulary nodes, frames without level-0 keypoints, a vocabulary
    transform of zero descriptors -- every call succeeds and reports zero matches."""
    import orbslamm_b200 as ob
    from orbslamm_b200 import vocabulary as V
    rng = np.random.default_rng(0)
    d1 = rng.integers(0, 256, (2, 40, 32), dtype=np.uint8); d2 = rng.integers(0, 256, (2, 50, 32), dtype=np.uint8)
    a1 = np.zeros((2, 40), np.float32); a2 = np.zeros((2, 50), np.float32)
    fvA = oracle.feature_vector(np.arange(40) % 4 * 2)              # nodes 0, 2, 4, 6
    fvB = oracle.feature_vector(np.arange(50) % 5 * 2 + 1)          # nodes 1, 3, 5, 7, 9: nothing in common
    empty = dict(nodes=np.zeros(0, np.int32), start=np.zeros(1, np.int32), items=np.zeros(0, np.int32))
    m = ob.ORBmatcher(0.9, True)
    nm, mt = m.SearchByBoW(d1, a1, np.ones((2, 40), np.uint8), [40, 0], [fvA, empty], d2, a2, np.ones((2, 50), np.uint8), [50, 50], [fvB, fvB])
    assert nm.tolist() == [0, 0] and (mt == -1).all()
    # same nodes on both sides but nothing eligible on side 1
    nm, mt = m.SearchByBoW(d1[:1], a1[:1], np.zeros((1, 40), np.uint8), [40], [fvA], d1[:1], a1[:1], np.ones((1, 40), np.uint8), [40], [fvA])
    assert nm.tolist() == [0] and (mt == -1).all()
    # identical keyframes: every feature matches itself (distance 0 < ratio * second best)
    nm, mt = m.SearchByBoW(d1[:1], a1[:1], np.ones((1, 40), np.uint8), [40], [fvA], d1[:1], a1[:1], np.ones((1, 40), np.uint8), [40], [fvA])
    n_o, m_o = oracle.search_by_bow(0, d1[0], a1[0], np.ones(40, np.uint8), fvA, d1[0], a1[0], np.ones(40, np.uint8), fvA, 0.9, True)
    assert nm[0] == n_o == 40 and np.array_equal(mt[0], m_o) and np.array_equal(m_o, np.arange(40))
    # SearchForInitialization: no level-0 keypoints in F1 / an empty F2
    oc1 = np.ones((2, 40), np.int32); oc1[1] = 0
    xy2 = rng.uniform(10, 300, (2, 50, 2)).astype(np.float32)
    nm, mt, pm = m.SearchForInitialization([0, 0, 400, 300], oc1, a1, d1, [40, 40], xy2, np.zeros((2, 50), np.int32), a2, d2, [50, 0], xy2[:, :40].copy(), 100)
    assert nm.tolist() == [0, 0] and (mt == -1).all() and np.array_equal(pm, xy2[:, :40])
    voc = V.ORBVocabulary(V.synthetic(4, 2, seed=1))
    t = voc.transform(np.zeros((1, 8, 32), np.uint8), [0], 4)[0]
    assert len(t["bow_ids"]) == 0 and len(t["fv"]["nodes"]) == 0 and t["fv"]["start"].tolist() == [0]
