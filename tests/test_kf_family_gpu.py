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
