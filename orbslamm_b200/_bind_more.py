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
    L.orbm_undistort_keypoints.argtypes = [vp, i, vp, vp, i, vp, vp, vp, i]
    L.orbm_is_in_frustum.argtypes = [vp, i, vp, vp, vp, vp, f, f, vp, vp, vp, vp, vp, i, vp, vp, vp, vp, i]
    L.orbm_project_points.argtypes = [vp, i, vp, vp, i, vp, vp, vp, vp, vp, i, vp, vp, vp, vp, vp, vp, i]
    L.orbm_search_by_projection_kf.argtypes = [vp, i, vp, vp, vp, vp, vp, vp, vp, i, vp, vp, vp, vp, vp, vp, vp, vp, i, i, f, i, vp, vp, i]
    L.orbm_search_best_in_window.argtypes = [vp, i, vp, vp, vp, vp, vp, vp, i, vp, vp, vp, vp, vp, vp, vp, i, i, vp, i, f, vp, vp, i]
    for n in ("orbm_project_points", "orbm_search_by_projection_kf", "orbm_search_best_in_window"):
        getattr(L, n).restype = c.c_int
    L.orbm_search_by_bow.argtypes = [vp, i, i, vp, vp, vp, vp, i, vp, vp, vp, vp, i, vp, vp, vp, vp, i, vp, vp, vp, vp, i, f, i, vp, vp, vp, i]
    L.orbm_search_by_bow.restype = c.c_int
    L.orbm_search_for_initialization.argtypes = [vp, i, vp, vp, vp, vp, vp, i, vp, vp, vp, vp, vp, i, vp, i, f, i, vp, vp, i]
    L.orbm_search_for_initialization.restype = c.c_int
    L.orbm_assign_features_to_grid.argtypes = [vp, i, vp, vp, vp, i, vp, vp, i]; L.orbm_assign_features_to_grid.restype = c.c_int
    L.orbm_get_features_in_area.argtypes = [vp, i, vp, vp, vp, vp, vp, i, vp, vp, vp, vp, i, i, vp, vp, i]; L.orbm_get_features_in_area.restype = c.c_int
    L.orbo_create.argtypes = [c.POINTER(vp), i]
    L.orbo_destroy.argtypes = [vp]
    L.orbo_stream.argtypes = [vp]; L.orbo_stream.restype = vp
    L.orbo_set_stream.argtypes = [vp, vp]
    L.orbo_synchronize.argtypes = [vp]
    L.orbo_kernel_launches.argtypes = [vp]; L.orbo_kernel_launches.restype = c.c_longlong
    L.orbo_comm_unique_id.argtypes = [vp]; L.orbo_comm_unique_id.restype = c.c_int
    L.orbo_comm_init.argtypes = [vp, i, i, vp]; L.orbo_comm_init.restype = c.c_int
    L.orbo_pose_optimization.argtypes = [vp, i, vp, vp, vp, vp, vp, vp, i, vp, vp, i]
    L.orbo_pose_optimization_matched.argtypes = [vp, i, vp, vp, vp, vp, vp, i, vp, vp, vp, i, vp, i, vp, vp, vp, i]
    L.orbo_pose_optimization_matched.restype = c.c_int
    L.orbo_optimize_sim3.argtypes = [vp, i, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i, f, i, vp, vp, vp, i]; L.orbo_optimize_sim3.restype = c.c_int
    L.orbo_sim3_prepare.argtypes = [vp, i, vp, vp, vp, i, vp, vp, vp, i]; L.orbo_sim3_prepare.restype = c.c_int
    L.orbo_sim3_check_inliers.argtypes = [vp, i, vp, vp, i, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i]; L.orbo_sim3_check_inliers.restype = c.c_int
    L.orbo_optimize_pose_graph.argtypes = [vp, i, vp, vp, i, vp, vp, vp, i, i, c.c_double, vp]; L.orbo_optimize_pose_graph.restype = c.c_int
    L.orbo_bundle_adjust.argtypes = [vp, i, vp, vp, vp, i, vp, i, vp, vp, vp, vp, i, i, i, i, vp, vp, vp, vp, vp]
    for n in ("orbo_create", "orbo_destroy", "orbo_set_stream", "orbo_synchronize", "orbo_pose_optimization", "orbo_bundle_adjust"):
        getattr(L, n).restype = c.c_int
    L.orbx_extract_host_async.argtypes = [vp, vp, i, i, i, i, c.c_size_t]; L.orbx_extract_host_async.restype = c.c_int
    L.orbf_create.argtypes = [c.POINTER(vp), vp, vp, vp, i]; L.orbf_create.restype = c.c_int
    L.orbf_destroy.argtypes = [vp]; L.orbf_destroy.restype = c.c_int
    L.orbf_track_frames.argtypes = [vp, vp, i, i, i, i, c.c_size_t, vp, vp, vp, i, vp, vp, vp, vp, vp, vp, i, f, i, i, vp, vp, vp, vp, vp, vp, vp, i, vp, vp, vp, vp, vp]
    L.orbf_track_frames.restype = c.c_int
    for n in ("orbm_create", "orbm_destroy", "orbm_synchronize", "orbm_descriptor_distance", "orbm_project_last_frame",
              "orbm_search_by_projection"):
        getattr(L, n).restype = c.c_int
