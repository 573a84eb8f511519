"""GPU parity: CUDA ORB extractor (through the C-ABI) vs the CPU oracle, bit-exact."""
import numpy as np
import pytest

import oracle
from orbslamm_b200 import synth

pytestmark = pytest.mark.gpu

FIELDS = ("x", "y", "angle", "response", "octave", "size", "desc")


def _assert_same(got, ref, tag):
    assert len(got["x"]) == len(ref["x"]), f"{tag}: count {len(got['x'])} vs {len(ref['x'])}"
    for k in FIELDS:
        assert np.array_equal(got[k], ref[k]), f"{tag}: field {k} differs at {np.nonzero(np.atleast_1d(got[k] != ref[k]))[0][:5]}"


@pytest.mark.parametrize("cam", ["KITTI", "TUM"])
def test_extract_matches_oracle(lib, cam):
    import orbslamm_b200 as ob
    c = getattr(synth, cam)
    frames, _ = synth.stream(c["w"], c["h"], 3, stream_id=3)
    ex = ob.ORBextractor(c["nfeatures"], 1.2, 8, 20, 7)
    P = oracle.orb_params(c["nfeatures"], 1.2, 8, 20, 7)
    outs = ex.extract_batch(np.stack(frames))
    for f, (img, got) in enumerate(zip(frames, outs)):
        _assert_same(got, oracle.orb_extract(P, img), f"{cam} frame {f}")


def test_extract_matches_reference_object_code(lib):
    """CUDA extractor vs the reference's own ORBextractor.cc object code (oracle/_ref, prebuilt in the build container; see
    tests/test_oracle_vs_reference.py for what is and is not the reference's code in that library)."""
    from oracle import ref_build
    if not ref_build.available():
        pytest.skip("oracle/_ref not built")
    import orbslamm_b200 as ob
    for cam, nf in (("KITTI", 2000), ("TUM", 1000)):
        c = getattr(synth, cam)
        frames, _ = synth.stream(c["w"], c["h"], 2, stream_id=9)
        ex = ob.ORBextractor(nf, 1.2, 8, 20, 7)
        R = ref_build.RefORBextractor(nf, 1.2, 8, 20, 7)
        for f, (img, got) in enumerate(zip(frames, ex.extract_batch(np.stack(frames)))):
            _assert_same(got, R(img), f"{cam} frame {f} vs reference")


@pytest.mark.parametrize("pitch_kind", ["packed_odd", "aligned64", "misaligned_base"])
def test_extract_device_resident_images(lib, pitch_kind):
    """orbx_extract_device on caller-owned device memory: 16-byte aligned buffers take the TMA-engine bulk-copy path, a packed odd
    pitch (1241) or a misaligned base pointer takes the plain-load fallback of level 0 -- same result either way."""
    import torch
    import orbslamm_b200 as ob
    c = synth.KITTI
    w, h = c["w"], c["h"]
    frames, _ = synth.stream(w, h, 2, stream_id=4)
    pitch = w if pitch_kind == "packed_odd" else (w + 63) // 64 * 64
    off = 3 if pitch_kind == "misaligned_base" else 0
    buf = torch.zeros(off + 2 * h * pitch + 64, dtype=torch.uint8, device="cuda")
    host = np.zeros((2, h, pitch), np.uint8)
    host[:, :, :w] = np.stack(frames)
    buf[off:off + host.size] = torch.from_numpy(host.reshape(-1)).cuda()
    ex = ob.ORBextractor(c["nfeatures"], 1.2, 8, 20, 7)
    ex.extract_device(buf.data_ptr() + off, 2, w, h, pitch, h * pitch)
    slab = ex.max_keypoints(w, h)
    r = ex.download(2, slab)
    P = oracle.orb_params(c["nfeatures"], 1.2, 8, 20, 7)
    for f in range(2):
        ref = oracle.orb_extract(P, frames[f])
        n = int(r["counts"][f])
        assert n == len(ref["x"])
        assert np.array_equal(r["xy"][f, :n, 0], ref["x"]) and np.array_equal(r["xy"][f, :n, 1], ref["y"])
        assert np.array_equal(r["angle"][f, :n], ref["angle"]) and np.array_equal(r["response"][f, :n], ref["response"])
        assert np.array_equal(r["octave"][f, :n], ref["octave"]) and np.array_equal(r["desc"][f, :n], ref["desc"])


@pytest.mark.parametrize("shape,nf,levels,sf", [((97, 131), 100, 4, 1.2), ((241, 322), 500, 5, 1.5), ((120, 500), 200, 3, 1.3), ((480, 640), 1000, 8, 1.2),
                                                 ((200, 300), 150, 3, 2.0)])
def test_extract_other_shapes(lib, shape, nf, levels, sf):
    """Small / wide images, other level counts and scale factors (large FAST cells on coarse levels, one-tile blur rows, ragged
    batches through the chunked host path) == oracle.  Scale factor 2.0: the four source columns of a resize thread no longer fit
    the three aligned words of k_resize_level<true>, the byte-load variant runs."""
    import orbslamm_b200 as ob
    n = 19                                                     # not a multiple of the 16-frame upload chunk
    frames = [synth.stream(shape[1], shape[0], 1, stream_id=100 + i)[0][0] for i in range(3)]
    batch = np.stack([frames[i % 3] for i in range(n)])
    ex = ob.ORBextractor(nf, sf, levels, 20, 7)
    P = oracle.orb_params(nf, sf, levels, 20, 7)
    refs = [oracle.orb_extract(P, f) for f in frames]
    outs = ex.extract_batch(batch)
    for i in range(n):
        _assert_same(outs[i], refs[i % 3], f"{shape} frame {i}")


def test_two_extractors_of_different_size_interleaved(lib):
    """Monocular Tracking owns mpIniORBextractor(2 * nFeatures) and mpORBextractorLeft(nFeatures) (Tracking.cc:115-121) and uses them alternately after a tracking
    loss; the opt-in shared-memory limits of the kernels are per-device function attributes, so a second, smaller handle must not lower what the first one needs."""
    import orbslamm_b200 as ob
    c = synth.KITTI
    frames, _ = synth.stream(c["w"], c["h"], 2, stream_id=12)
    big = ob.ORBextractor(2 * c["nfeatures"], 1.2, 8, 20, 7)
    first = big(frames[0])
    small = ob.ORBextractor(c["nfeatures"] // 4, 1.2, 8, 20, 7)
    Pb, Ps = oracle.orb_params(2 * c["nfeatures"], 1.2, 8, 20, 7), oracle.orb_params(c["nfeatures"] // 4, 1.2, 8, 20, 7)
    for rnd in range(2):
        _assert_same(small(frames[rnd]), oracle.orb_extract(Ps, frames[rnd]), f"small extractor, round {rnd}")
        _assert_same(big(frames[rnd]), oracle.orb_extract(Pb, frames[rnd]), f"big extractor after the small one, round {rnd}")
    _assert_same(first, oracle.orb_extract(Pb, frames[0]), "big extractor, first call")
