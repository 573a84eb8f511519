"""ctypes declarations + thin Python mirrors of the matcher / optimizer entry points."""
import ctypes

import numpy as np

MEM_HOST, MEM_DEVICE = 0, 1


def declare(L):
    c = ctypes
    vp, i, f = c.c_void_p, c.c_int, c.c_float
    L.orbm_create.argtypes = [c.POINTER(vp), i]
    L.orbm_destroy.argtypes = [vp]
    L.orbm_stream.argtypes = [vp]; L.orbm_stream.restype = vp
    L.orbm_synchronize.argtypes = [vp]
    L.orbm_kernel_launches.argtypes = [vp]; L.orbm_kernel_launches.restype = c.c_longlong
    L.orbm_descriptor_distance.argtypes = [vp, vp, i, vp, i, vp, i]
    L.orbm_project_last_frame.argtypes = [vp, i, vp, vp, vp, vp, i, vp, vp, vp, i, f, vp, vp, vp, vp, vp, i]
    L.orbm_search_by_projection.argtypes = [vp, i, vp, vp, vp, vp, vp, vp, i, vp, vp, vp, vp, vp, vp, vp, vp, i, i, f, i, vp, vp, i]
    for n in ("orbm_create", "orbm_destroy", "orbm_synchronize", "orbm_descriptor_distance", "orbm_project_last_frame",
              "orbm_search_by_projection"):
        getattr(L, n).restype = c.c_int
