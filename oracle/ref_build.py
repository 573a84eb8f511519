"""oracle/ref_build.py -- build and load oracle/_ref/libref_orbextractor.so: the REFERENCE's own ORBextractor.cc
compiled unmodified (from /root/reference, never copied) against the OpenCV stand-in oracle/cvshim.

TEST INFRASTRUCTURE ONLY.  /root/reference exists only in the build container; on the GPU box the prebuilt .so
travels with the snapshot (oracle/_ref/ is git-ignored, not gpurun-ignored) and `available()` reports what is there.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libref_orbextractor.so")
REF_ROOT = "/root/reference/SingleRobotScenario"
_LIB = None


def build():
    """Run oracle/Makefile when the reference sources are present; returns the .so path or None."""
    if os.path.exists(os.path.join(REF_ROOT, "src", "ORBextractor.cc")):
        subprocess.check_call(["make", "-s", "-C", _HERE, f"REF={REF_ROOT}"])
    return _SO if os.path.exists(_SO) else None


def available():
    return os.path.exists(_SO)


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(_SO)
        L.ref_orbx_create.restype = ctypes.c_void_p
        L.ref_orbx_create.argtypes = [ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.ref_orbx_destroy.argtypes = [ctypes.c_void_p]
        L.ref_orbx_extract.restype = ctypes.c_int
        L.ref_orbx_extract.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 7
        L.ref_orbx_tables.argtypes = [ctypes.c_void_p] + [ctypes.c_void_p] * 4
        L.ref_orbx_pyramid_level.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.ref_bump_nodes.restype = ctypes.c_long
        _LIB = L
    return _LIB


class RefORBextractor:
    """iORB_SLAM::ORBextractor of the reference (ORBextractor.h:45-111), object code of the reference's own source."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, ascending_heap=True):
        """ascending_heap: quad-tree nodes get ascending, never reused addresses (oracle/ref_bump_alloc.cc), which makes the
        reference's heap-address tie-break (ORBextractor.cc:684) equal to 'later-created node first'."""
        self.L = lib()
        self.L.ref_bump_enable(1 if ascending_heap else 0)
        self.nlevels = nlevels
        self.cap = 4 * nfeatures + 64
        self.h = self.L.ref_orbx_create(nfeatures, scale_factor, nlevels, ini_th, min_th)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_orbx_destroy(self.h)
            self.h = None

    def __call__(self, image):
        img = np.ascontiguousarray(image, np.uint8)
        h, w = img.shape
        c = self.cap
        x = np.zeros(c, np.float32); y = np.zeros(c, np.float32); ang = np.zeros(c, np.float32); resp = np.zeros(c, np.float32)
        octv = np.zeros(c, np.int32); size = np.zeros(c, np.float32); desc = np.zeros((c, 32), np.uint8)
        n = self.L.ref_orbx_extract(self.h, img.ctypes.data, w, h, w, c, x.ctypes.data, y.ctypes.data, ang.ctypes.data, resp.ctypes.data,
                                    octv.ctypes.data, size.ctypes.data, desc.ctypes.data)
        assert n <= c
        return dict(x=x[:n], y=y[:n], angle=ang[:n], response=resp[:n], octave=octv[:n], size=size[:n], desc=desc[:n])

    def tables(self):
        a = [np.zeros(self.nlevels, np.float32) for _ in range(4)]
        self.L.ref_orbx_tables(self.h, *[t.ctypes.data for t in a])
        return dict(scale=a[0], inv_scale=a[1], sigma2=a[2], inv_sigma2=a[3])

    def pyramid_level(self, level):
        w = ctypes.c_int(); h = ctypes.c_int()
        self.L.ref_orbx_pyramid_level(self.h, level, ctypes.byref(w), ctypes.byref(h), None, 0)
        out = np.zeros((h.value, w.value), np.uint8)
        self.L.ref_orbx_pyramid_level(self.h, level, ctypes.byref(w), ctypes.byref(h), out.ctypes.data, w.value)
        return out
