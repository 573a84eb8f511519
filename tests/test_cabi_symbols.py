"""The C-ABI library loads without a GPU and exports every symbol include/orbslamm_b200.h declares."""
import ctypes
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "orbslamm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(orb[smxofv]_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    names = _declared()
    assert len(names) >= 55 and "orbv_transform" in names and "orbf_track_frames" in names and "orbm_search_by_bow" in names
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/orbslamm_b200.h but not exported"


def test_no_compute_without_gpu_but_clean_errors(lib):
    import orbslamm_b200 as ob
    assert lib.orbs_version() >= 100
    if lib.orbs_device_count() > 0:
        return
    # no CUDA device here: creating any handle must fail loudly with ORBS_E_CUDA (no CPU fallback exists)
    for ctor in (lambda: ob.ORBextractor(1000, 1.2, 8, 20, 7), lambda: ob.ORBmatcher(0.9, True), lambda: ob.Optimizer()):
        try:
            ctor()
        except ob.OrbsError as e:
            assert e.code == -2
        else:
            raise AssertionError("handle creation succeeded without a GPU")


def test_argument_validation_needs_no_gpu(lib):
    h = ctypes.c_void_p()
    assert lib.orbx_create(ctypes.byref(h), 0, 1.2, 8, 20, 7, 0) == -1          # nfeatures must be positive
    assert lib.orbx_create(ctypes.byref(h), 1000, 1.2, 17, 20, 7, 0) == -1      # nlevels <= 16
    assert lib.orbx_create(ctypes.byref(h), 1000, 1.2, 8, 0, 7, 0) == -1        # thresholds >= 1
    assert b"" != lib.orbs_last_error()
    lib.orbv_create.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 5
    assert lib.orbv_create(ctypes.byref(h), 0, 10, 6, 100, None, None, None, None, None) == -1     # vocabulary arrays are required
    cs = np.array([0, 1, 1], np.int32); z = np.zeros(64, np.uint8); w = np.zeros(2, np.float64); ids = np.zeros(2, np.int32)
    assert lib.orbv_create(ctypes.byref(h), 0, 10, 6, 1, z.ctypes.data, cs.ctypes.data, ids.ctypes.data, ids.ctypes.data, w.ctypes.data) == -1   # a root alone is no tree
    assert lib.orbv_create(ctypes.byref(h), 0, 10, 6, 2, z.ctypes.data, np.array([0, 2, 2], np.int32).ctypes.data, ids.ctypes.data, ids.ctypes.data,
                           w.ctypes.data) == -1                                                                                  # child lists must cover n - 1 nodes


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "orbslamm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                code = "\n".join(l for l in txt.splitlines() if not l.lstrip().startswith(("#", "//", "*", '"')))
                assert not re.search(r"^\s*(import oracle|from oracle)", code, flags=re.M), f
                assert "liboracle" not in code and "oracle." not in code, f


def test_host_shims_compile_for_both_trees():
    """The C++ drop-in sources compile against the mock data model for the single-robot tree and, with -DORBSLAMM_MULTI_ROBOT, for the multi-robot
    tree (MMGlobalBundleAdjustemnt, MMOptimizeEssentialGraph, the isNotFixed() rule) -- no GPU needed; running them is tests/test_host_shim_gpu.py."""
    import subprocess
    host = os.path.join(ROOT, "orbslamm_b200", "host")
    srcs = [os.path.join(host, f) for f in ("ORBextractor_b200.cc", "ORBmatcher_b200.cc", "Optimizer_b200.cc", "Sim3Solver_b200.cc", "mock/slam_statics.cc",
                                            "mock/Sim3Solver_rest.cc", "test/host_shim_test.cc", "test/host_kf_family_test.cc", "test/host_sim3solver_test.cc",
                                            "test/host_essential_graph_test.cc")]
    base = ["g++", "-std=c++14", "-fsyntax-only", "-I" + os.path.join(host, "mock"), "-I" + host, "-I" + os.path.join(ROOT, "include")]
    for extra in ([], ["-DORBSLAMM_MULTI_ROBOT"], ["-DORBSLAMM_DEVICE_COMPUTE_SIM3"]):
        out = subprocess.run(base + extra + srcs, capture_output=True, text=True)
        assert out.returncode == 0, out.stderr[-3000:]


def test_projection_flags_and_struct_match_the_oracle():
    """orbm_projection and its PROJ_* flags are declared twice on the Python side (product: orbslamm_b200, checker: oracle): same values, same layout"""
    import ctypes
    import orbslamm_b200 as ob
    import oracle
    for n in ("PROJ_TWO_STEP", "PROJ_NO_DEPTH", "PROJ_FRAME_BOUNDS", "PROJ_FRAME_UV", "PROJ_DIST_CAMERA", "PROJ_CHECK_NORMAL", "PROJ_LEVEL_PLUS1"):
        assert getattr(ob, n) == getattr(oracle, n)
    assert ctypes.sizeof(ob.Projection) == ctypes.sizeof(oracle.Projection)
    a = ob.make_projection(range(9), range(3), [1, 2, 3, 4], [5, 6, 7, 8], 0.18, 10, 33, Ow=[1, 2, 3], R2=range(9, 18), t2=[7, 8, 9])
    b = oracle.make_projection(range(9), range(3), [1, 2, 3, 4], [5, 6, 7, 8], 0.18, 10, 33, Ow=[1, 2, 3], R2=range(9, 18), t2=[7, 8, 9])
    assert bytes(a) == bytes(b)
