"""GPU parity: Hamming distances, last-frame projection and both SearchByProjection variants vs the oracle."""
import numpy as np
import pytest

import oracle
from orbslamm_b200 import synth
from helpers import make_tracking_case, slab

pytestmark = pytest.mark.gpu


def test_descriptor_distance_kat(lib):
    import orbslamm_b200 as ob
    m = ob.ORBmatcher()
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (37, 32), dtype=np.uint8); b = rng.integers(0, 256, (53, 32), dtype=np.uint8)
    a[0] = 0; b[0] = 0xFF; a[1] = b[1]
    d = m.DescriptorDistance(a, b)
    assert d[0, 0] == 256 and d[1, 1] == 0
    ref = np.array([[oracle.descriptor_distance(x, y) for y in b] for x in a])
    assert np.array_equal(d, ref)
    assert np.array_equal(d, np.unpackbits(a[:, None, :] ^ b[None, :, :], axis=2).sum(2))


@pytest.mark.parametrize("cam", ["TUM", "KITTI"])
def test_search_by_projection_frames(lib, cam):
    import orbslamm_b200 as ob
    c = getattr(synth, cam)
    cases = [make_tracking_case(c, sid) for sid in (1, 2, 3)]
    m = ob.ORBmatcher(0.9, True)
    sf = np.array(list(cases[0]["P"].scale)[:8], np.float32)
    nF = max(len(k["cur"]["x"]) for k in cases); nQ = max(len(k["last"]["x"]) for k in cases)
    g = oracle.grid_params(*cases[0]["bounds"])
    # projection
    Tcw = np.stack([k["Tcw"] for k in cases]); Xw = slab([k["Xw"] for k in cases], nQ, np.float32, (3,))
    loct = slab([k["last"]["octave"] for k in cases], nQ, np.int32); valid = slab([k["valid"] for k in cases], nQ, np.uint8)
    qc = np.array([len(k["last"]["x"]) for k in cases], np.int32)
    qv, uv, rad, mn, mx = m.project_last_frame(Tcw, cases[0]["K4"], cases[0]["bounds"], sf, Xw, loct, qc, 15.0, valid)
    refs = []
    for i, k in enumerate(cases):
        n = qc[i]
        r = oracle.project_last_frame(k["Tcw"], k["K4"], g, sf, k["Xw"], k["last"]["octave"], 15.0, k["valid"])
        refs.append(r)
        assert np.array_equal(qv[i, :n], r[0])
        ok = r[0].astype(bool)
        assert ok.sum() > 100
        assert np.array_equal(uv[i, :n][ok], r[1][ok]) and np.array_equal(rad[i, :n][ok], r[2][ok])
        assert np.array_equal(mn[i, :n][ok], r[3][ok]) and np.array_equal(mx[i, :n][ok], r[4][ok])
    # search
    f_xy = slab([np.stack([k["cur"]["x"], k["cur"]["y"]], 1) for k in cases], nF, np.float32, (2,))
    f_oct = slab([k["cur"]["octave"] for k in cases], nF, np.int32); f_ang = slab([k["cur"]["angle"] for k in cases], nF, np.float32)
    f_desc = slab([k["cur"]["desc"] for k in cases], nF, np.uint8, (32,)); fc = np.array([len(k["cur"]["x"]) for k in cases], np.int32)
    q_ang = slab([k["last"]["angle"] for k in cases], nQ, np.float32); q_desc = slab([k["last"]["desc"] for k in cases], nQ, np.uint8, (32,))
    for th_dist in (100, 50):
        nm, fm = m.SearchByProjection(cases[0]["bounds"], f_xy, f_oct, f_ang, f_desc, fc, qv, uv, rad, mn, mx, q_ang, q_desc, qc, th_dist)
        for i, k in enumerate(cases):
            r = refs[i]
            n_ref, fm_ref = oracle.search_by_projection(g, f_xy[i, :fc[i]], k["cur"]["octave"], k["cur"]["angle"], k["cur"]["desc"],
                                                        r[0], r[1], r[2], r[3], r[4], k["last"]["angle"], k["last"]["desc"], th_dist, 0.0, True)
            assert n_ref > 50
            assert nm[i] == n_ref
            assert np.array_equal(fm[i, :fc[i]], fm_ref)


def test_search_by_projection_local_map(lib):
    """(Frame, MapPoints) variant: ratio test, two-level window, pre-assigned features, no orientation check."""
    import orbslamm_b200 as ob
    c = synth.TUM
    cases = [make_tracking_case(c, sid) for sid in (4, 5)]
    m = ob.ORBmatcher(0.8, True)
    sf = np.array(list(cases[0]["P"].scale)[:8], np.float32)
    nF = max(len(k["cur"]["x"]) for k in cases); nQ = max(len(k["last"]["x"]) for k in cases)
    g = oracle.grid_params(*cases[0]["bounds"])
    rng = np.random.default_rng(5)
    per = []
    for k in cases:
        n = len(k["last"]["x"])
        dx, dy = k["shift"]
        uv = np.stack([k["last"]["x"] + dx + rng.normal(0, 1.5, n), k["last"]["y"] + dy + rng.normal(0, 1.5, n)], 1).astype(np.float32)
        lvl = k["last"]["octave"]
        viewcos = rng.uniform(0.99, 1.0, n).astype(np.float32)
        r = np.where(viewcos > np.float32(0.998), np.float32(2.5), np.float32(4.0)).astype(np.float32) * np.float32(3.0)   # RadiusByViewingCos * th
        rad = (r * sf[lvl]).astype(np.float32)
        taken = np.full(len(k["cur"]["x"]), -1, np.int32); taken[rng.random(len(taken)) < 0.15] = 7
        per.append(dict(valid=(rng.random(n) < 0.85).astype(np.uint8), uv=uv, rad=rad, mn=(lvl - 1).astype(np.int32), mx=lvl.astype(np.int32), taken=taken))
    f_xy = slab([np.stack([k["cur"]["x"], k["cur"]["y"]], 1) for k in cases], nF, np.float32, (2,))
    f_oct = slab([k["cur"]["octave"] for k in cases], nF, np.int32); f_ang = slab([k["cur"]["angle"] for k in cases], nF, np.float32)
    f_desc = slab([k["cur"]["desc"] for k in cases], nF, np.uint8, (32,)); fc = np.array([len(k["cur"]["x"]) for k in cases], np.int32)
    q_ang = slab([k["last"]["angle"] for k in cases], nQ, np.float32); q_desc = slab([k["last"]["desc"] for k in cases], nQ, np.uint8, (32,))
    qc = np.array([len(k["last"]["x"]) for k in cases], np.int32)
    fm_in = slab([p["taken"] for p in per], nF, np.int32); fm_in[fm_in == 0] = -1
    for i in range(len(cases)):
        fm_in[i, fc[i]:] = -1
    nm, fm = m.SearchByProjection(cases[0]["bounds"], f_xy, f_oct, f_ang, f_desc, fc, slab([p["valid"] for p in per], nQ, np.uint8),
                                  slab([p["uv"] for p in per], nQ, np.float32, (2,)), slab([p["rad"] for p in per], nQ, np.float32),
                                  slab([p["mn"] for p in per], nQ, np.int32), slab([p["mx"] for p in per], nQ, np.int32), q_ang, q_desc, qc,
                                  100, use_ratio=True, feat_match=fm_in)
    for i, (k, p) in enumerate(zip(cases, per)):
        n_ref, fm_ref = oracle.search_by_projection(g, f_xy[i, :fc[i]], k["cur"]["octave"], k["cur"]["angle"], k["cur"]["desc"], p["valid"], p["uv"],
                                                    p["rad"], p["mn"], p["mx"], k["last"]["angle"], k["last"]["desc"], 100, 0.8, False, feat_match=p["taken"])
        assert n_ref > 50
        assert nm[i] == n_ref
        assert np.array_equal(fm[i, :fc[i]], fm_ref)


def test_search_by_projection_4000_features(lib):
    """The monocular initialisation extractor uses 2 x nFeatures (Tracking.cc:126): more than 2048 feature slots per frame -> 64-bit
    candidate keys in k_search_candidates, many over-full windows (warp rescans in k_search_resolve)."""
    import orbslamm_b200 as ob
    k = make_tracking_case(synth.KITTI, 9, nfeatures=4000)
    m = ob.ORBmatcher(0.9, True)
    sf = np.array(list(k["P"].scale)[:8], np.float32)
    g = oracle.grid_params(*k["bounds"])
    cur, last = k["cur"], k["last"]
    nF, nQ = len(cur["x"]), len(last["x"])
    assert nF > 2048
    r = oracle.project_last_frame(k["Tcw"], k["K4"], g, sf, k["Xw"], last["octave"], 15.0, k["valid"])
    fxy = np.stack([cur["x"], cur["y"]], 1)
    for th, ratio, ori, scale_r in ((100, 0.0, True, 1.0), (100, 0.8, False, 2.0)):
        n_ref, fm_ref = oracle.search_by_projection(g, fxy, cur["octave"], cur["angle"], cur["desc"], r[0], r[1], r[2] * scale_r, r[3], r[4],
                                                    last["angle"], last["desc"], th, ratio, ori)
        mm = ob.ORBmatcher(ratio if ratio > 0 else 0.9, ori)
        nm, fm = mm.SearchByProjection(k["bounds"], fxy[None], cur["octave"][None], cur["angle"][None], cur["desc"][None], np.array([nF], np.int32),
                                       r[0][None], r[1][None], (r[2] * scale_r)[None], r[3][None], r[4][None], last["angle"][None], last["desc"][None],
                                       np.array([nQ], np.int32), th, use_ratio=ratio > 0)
        assert n_ref > 500 and int(nm[0]) == n_ref and np.array_equal(fm[0], fm_ref)


def test_search_matches_reference_object_code(lib):
    """CUDA projection search vs the reference's own ORBmatcher.cc object code (oracle/_ref/libref_orbmatcher.so, prebuilt; see
    tests/test_oracle_vs_reference.py for what is and is not the reference's code in that library)."""
    from oracle import ref_build
    if not ref_build.matcher_available():
        pytest.skip("oracle/_ref matcher not built")
    import orbslamm_b200 as ob
    k = make_tracking_case(synth.KITTI, 6)
    sf = np.array(list(k["P"].scale)[:8], np.float32)
    cur, last = k["cur"], k["last"]
    nF, nQ = len(cur["x"]), len(last["x"])
    m = ob.ORBmatcher(0.9, True)
    qv, uv, rad, mn, mx = m.project_last_frame(k["Tcw"][None], k["K4"], k["bounds"], sf, k["Xw"][None], last["octave"][None], np.array([nQ], np.int32), 15.0,
                                               k["valid"][None])
    fxy = np.stack([cur["x"], cur["y"]], 1)
    nm, fm = m.SearchByProjection(k["bounds"], fxy[None], cur["octave"][None], cur["angle"][None], cur["desc"][None], np.array([nF], np.int32), qv, uv, rad, mn, mx,
                                  last["angle"][None], last["desc"][None], np.array([nQ], np.int32), 100)
    n_r, fm_r = ref_build.ref_search_last_frame(k["K4"], k["bounds"], k["Tcw"], sf, cur, last, k["Xw"], k["valid"], 15.0, 0.9, True)
    assert int(nm[0]) == n_r and n_r > 500 and np.array_equal(fm[0], fm_r)


def test_distinctive_descriptors_batch(lib):
    """orbm_distinctive_descriptors == MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:242-307) per point: ragged batch incl. empty, single, pair (the
    (int)(0.5 * (N - 1)) = 0 median of a pair is the zero self-distance: first row wins), more rows than lanes, exact duplicates (ties)."""
    import orbslamm_b200 as ob
    rng = np.random.default_rng(8)
    lists = []
    for n in [0, 1, 2, 3, 4, 7, 33, 70, 5, 0, 12] + [int(x) for x in rng.integers(1, 20, 300)]:
        base = rng.integers(0, 256, 32, dtype=np.uint8)
        d = np.tile(base, (n, 1))
        for i in range(n):
            for b in rng.integers(0, 256, int(rng.integers(0, 60))): d[i, b >> 3] ^= np.uint8(1 << (b & 7))
        if n >= 4: d[n - 1] = d[1]                                          # a duplicate row
        lists.append(d)
    got = ob.ORBmatcher(0.6, True).ComputeDistinctiveDescriptors(lists)
    ref = np.array([oracle.distinctive_descriptor(d) for d in lists], np.int32)
    assert np.array_equal(got, ref)
    assert got[0] == -1 and got[1] == 0 and got[2] == 0


def test_search_by_projection_wide_windows_32bit_keys(lib):
    """k_search_candidates with <= 2048 feature slots (32-bit keys) and windows three times the usual radius: more than 32 grid cells and more than 32 features
    per window, i.e. several cell chunks and several sorted batches merged into the running top-8 per query -- must still equal the sequential scan."""
    import orbslamm_b200 as ob
    k = make_tracking_case(synth.TUM, 5)
    sf = np.array(list(k["P"].scale)[:8], np.float32)
    g = oracle.grid_params(*k["bounds"])
    cur, last = k["cur"], k["last"]
    nF, nQ = len(cur["x"]), len(last["x"])
    assert nF <= 2048
    r = oracle.project_last_frame(k["Tcw"], k["K4"], g, sf, k["Xw"], last["octave"], 15.0, k["valid"])
    fxy = np.stack([cur["x"], cur["y"]], 1)
    for scale_r, lvl in ((3.0, False), (6.0, True)):
        mn, mx = (np.full_like(r[3], -1), np.full_like(r[4], -1)) if lvl else (r[3], r[4])       # lvl: no octave gate -> every feature of the window is a candidate
        n_ref, fm_ref = oracle.search_by_projection(g, fxy, cur["octave"], cur["angle"], cur["desc"], r[0], r[1], r[2] * scale_r, mn, mx,
                                                    last["angle"], last["desc"], 100, 0.0, True)
        nm, fm = ob.ORBmatcher(0.9, True).SearchByProjection(k["bounds"], fxy[None], cur["octave"][None], cur["angle"][None], cur["desc"][None], np.array([nF], np.int32),
                                                             r[0][None], r[1][None], (r[2] * scale_r)[None], mn[None], mx[None], last["angle"][None],
                                                             last["desc"][None], np.array([nQ], np.int32), 100)
        assert n_ref > 100 and int(nm[0]) == n_ref and np.array_equal(fm[0], fm_ref), (scale_r, lvl, n_ref, int(nm[0]))
