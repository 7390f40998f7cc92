"""ctypes loader of semantic_slam_b200/libssb.so (the C-ABI of include/ssb.h)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.environ.get("SSB_LIB") or os.path.join(_HERE, "libssb.so")   # SSB_LIB: an A/B build of the library (experiments)
_LIB = None

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


class SsbError(RuntimeError):
    pass


class GraphOpts(C.Structure):
    _fields_ = [("device", C.c_int), ("verbose", C.c_int), ("max_pcg_iters", C.c_int), ("pcg_tol", C.c_double),
                ("preconditioner", C.c_int), ("coarse_group", C.c_int), ("reserved", C.c_int * 4)]


class LmStats(C.Structure):
    _fields_ = [("iterations", C.c_int), ("terminated", C.c_int), ("total_trials", C.c_int),
                ("total_pcg_iters", C.c_int), ("chi2_initial", C.c_double), ("chi2_final", C.c_double),
                ("lambda_final", C.c_double), ("ms_prepare", C.c_double), ("ms_device", C.c_double),
                ("ms_total", C.c_double), ("kernel_launches", C.c_longlong), ("ms_pcg", C.c_double)]


class CloudLayoutC(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("point_step", C.c_int), ("row_step", C.c_int),
                ("off_x", C.c_int), ("off_y", C.c_int), ("off_z", C.c_int), ("off_rgb", C.c_int)]


class RansacOpts(C.Structure):
    _fields_ = [("threshold", C.c_double), ("refine", C.c_int), ("mode", C.c_int), ("max_iterations", C.c_int),
                ("probability", C.c_double), ("device", C.c_int), ("reserved", C.c_int * 3)]


class OrganizedOpts(C.Structure):
    _fields_ = [("max_depth_change_factor", C.c_float), ("normal_smoothing_size", C.c_float), ("min_inliers", C.c_int),
                ("angular_threshold", C.c_float), ("distance_threshold", C.c_float), ("maximum_curvature", C.c_float),
                ("norm_point_thres", C.c_int), ("reserved", C.c_int * 3)]


class ClusterOpts(C.Structure):
    _fields_ = [("num_centroids_normals", C.c_int), ("num_centroids_distance", C.c_int), ("kmeans_attempts", C.c_int),
                ("kmeans_max_count", C.c_int), ("kmeans_epsilon", C.c_double), ("min_cluster_points", C.c_int),
                ("centroid_tolerance", C.c_float), ("ransac_hypotheses", C.c_int), ("ransac_seed", C.c_uint),
                ("reserved", C.c_int * 4)]


def build(verbose: bool = False) -> str:
    """Compile libssb.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-s"]
    subprocess.check_call(cmd)
    return _SO


# every symbol include/ssb.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "ssb_graph_default_opts", "ssb_graph_create", "ssb_graph_destroy", "ssb_graph_add_se3_node",
    "ssb_graph_add_point_xyz_node", "ssb_graph_add_se3_edge", "ssb_graph_add_se3_point_xyz_edge",
    "ssb_graph_add_point_xyz_point_xyz_edge", "ssb_graph_add_plane_node", "ssb_graph_add_se3_plane_edge",
    "ssb_graph_get_plane", "ssb_graph_set_plane", "ssb_graph_num_vertices", "ssb_graph_num_edges", "ssb_graph_get_se3",
    "ssb_graph_get_point_xyz", "ssb_graph_set_se3", "ssb_graph_set_point_xyz", "ssb_graph_set_fixed",
    "ssb_graph_hessian_index", "ssb_graph_get_all", "ssb_graph_set_all", "ssb_graph_invalidate", "ssb_graph_chi2",
    "ssb_graph_prepare", "ssb_graph_optimize", "ssb_graph_optimize_resident", "ssb_graph_get_history",
    "ssb_graph_landmark_marginals", "ssb_graph_save_g2o", "ssb_graph_load_g2o", "ssb_graph_edge_linearize",
    "ssb_graph_solve_once", "ssb_comm_unique_id", "ssb_graph_attach_comm", "ssb_graph_attach_local", "ssb_shard_ranges",
    "ssb_graph_shard_info", "ssb_shard_plan", "ssb_ransac_default_opts",
    "ssb_ransac_create", "ssb_ransac_destroy", "ssb_ransac_plane_batch", "ssb_ransac_upload",
    "ssb_organized_default_opts", "ssb_organized_planes", "ssb_organized_last_ms", "ssb_ransac_run_resident", "ssb_ransac_fetch", "ssb_ransac_stream", "ssb_ransac_launch_count", "ssb_ransac_timing", "ssb_crop_bbox",
    "ssb_ransac_pcl_samples",
    "ssb_kmeans", "ssb_project_hull", "ssb_cluster_default_opts", "ssb_cluster_planes",
    "ssb_segment_planar_surfaces", "ssb_assoc_default_opts", "ssb_assoc_create", "ssb_assoc_destroy", "ssb_assoc_find_matches",
    "ssb_assoc_set_landmark_estimate", "ssb_assoc_set_landmark_cov", "ssb_assoc_num_landmarks", "ssb_assoc_get_landmark",
    "ssb_assoc_inverse3",
    "ssb_last_error", "ssb_build_info", "ssb_graph_stream", "ssb_graph_snapshot", "ssb_graph_restore",
]


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(_SO):
        raise SsbError(f"{_SO} is missing: build the CUDA extension first (python -c 'import __graft_entry__ as g; "
                       "g.build()'); there is no CPU fallback")
    # shards of one graph driven by threads of this process (attach_local) each own three streams: give every stream
    # its own hardware queue (effective only if CUDA is not initialised yet; the library does not depend on it)
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
    L = C.CDLL(_SO)
    vp = C.c_void_p
    L.ssb_last_error.restype = C.c_char_p
    L.ssb_build_info.restype = C.c_char_p
    L.ssb_graph_default_opts.argtypes = [C.POINTER(GraphOpts)]
    L.ssb_graph_create.argtypes = [C.POINTER(GraphOpts)]
    L.ssb_graph_create.restype = vp
    L.ssb_graph_destroy.argtypes = [vp]
    L.ssb_graph_add_se3_node.argtypes = [vp, dp]
    L.ssb_graph_add_point_xyz_node.argtypes = [vp, dp]
    L.ssb_graph_add_se3_edge.argtypes = [vp, C.c_int, C.c_int, dp, dp]
    L.ssb_graph_add_se3_point_xyz_edge.argtypes = [vp, C.c_int, C.c_int, dp, dp]
    L.ssb_graph_add_point_xyz_point_xyz_edge.argtypes = [vp, C.c_int, C.c_int, dp, dp]
    L.ssb_graph_add_plane_node.argtypes = [vp, dp]
    L.ssb_graph_add_se3_plane_edge.argtypes = [vp, C.c_int, C.c_int, dp, dp]
    L.ssb_graph_get_plane.argtypes = [vp, C.c_int, dp]
    L.ssb_graph_set_plane.argtypes = [vp, C.c_int, dp]
    L.ssb_graph_num_vertices.argtypes = [vp]
    L.ssb_graph_num_edges.argtypes = [vp]
    L.ssb_graph_get_se3.argtypes = [vp, C.c_int, dp]
    L.ssb_graph_get_point_xyz.argtypes = [vp, C.c_int, dp]
    L.ssb_graph_set_se3.argtypes = [vp, C.c_int, dp]
    L.ssb_graph_set_point_xyz.argtypes = [vp, C.c_int, dp]
    L.ssb_graph_set_fixed.argtypes = [vp, C.c_int, C.c_int]
    L.ssb_graph_hessian_index.argtypes = [vp, C.c_int]
    L.ssb_graph_get_all.argtypes = [vp, dp, dp]
    L.ssb_graph_set_all.argtypes = [vp, dp, dp]
    L.ssb_graph_invalidate.argtypes = [vp]
    L.ssb_graph_chi2.argtypes = [vp, dp]
    L.ssb_graph_prepare.argtypes = [vp]
    L.ssb_graph_optimize.argtypes = [vp, C.c_int, C.POINTER(LmStats)]
    L.ssb_graph_optimize_resident.argtypes = [vp, C.c_int, C.POINTER(LmStats)]
    L.ssb_graph_get_history.argtypes = [vp, dp, C.c_int]
    L.ssb_graph_landmark_marginals.argtypes = [vp, ip, C.c_int, dp]
    L.ssb_graph_save_g2o.argtypes = [vp, C.c_char_p]
    L.ssb_graph_load_g2o.argtypes = [vp, C.c_char_p]
    L.ssb_graph_edge_linearize.argtypes = [vp, C.c_int, dp, dp, dp]
    L.ssb_graph_solve_once.argtypes = [vp, C.c_double, dp, C.c_int]
    L.ssb_graph_stream.argtypes = [vp]
    L.ssb_graph_stream.restype = vp
    L.ssb_graph_snapshot.argtypes = [vp]
    L.ssb_graph_restore.argtypes = [vp]
    L.ssb_comm_unique_id.argtypes = [C.c_char_p]
    L.ssb_graph_attach_comm.argtypes = [vp, C.c_int, C.c_int, C.c_char_p]
    L.ssb_shard_ranges.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, ip]
    L.ssb_graph_attach_local.argtypes = [vp, C.c_int, C.c_int, C.c_char_p, C.c_int]
    L.ssb_graph_shard_info.argtypes = [vp, C.c_int, C.c_int, ip]
    L.ssb_shard_plan.argtypes = [C.c_int, C.c_int, ip, ip, C.c_int, ip, ip, C.c_int, C.c_int, C.c_int, ip, ip, ip]
    L.ssb_ransac_default_opts.argtypes = [C.POINTER(RansacOpts)]
    L.ssb_ransac_create.argtypes = [C.c_int]
    L.ssb_ransac_create.restype = vp
    L.ssb_ransac_destroy.argtypes = [vp]
    L.ssb_ransac_plane_batch.argtypes = [vp, vp, C.POINTER(CloudLayoutC), vp, C.c_int, vp, C.c_int,
                                         C.POINTER(RansacOpts), vp, vp, vp]
    L.ssb_ransac_upload.argtypes = [vp, vp, C.POINTER(CloudLayoutC), vp, C.c_int, vp, C.c_int, C.POINTER(RansacOpts)]
    L.ssb_ransac_run_resident.argtypes = [vp]
    L.ssb_ransac_fetch.argtypes = [vp, vp, vp, vp]
    L.ssb_ransac_stream.argtypes = [vp]
    L.ssb_ransac_stream.restype = vp
    L.ssb_ransac_launch_count.argtypes = [vp]
    L.ssb_ransac_launch_count.restype = C.c_longlong
    L.ssb_ransac_timing.argtypes = [vp, dp]
    L.ssb_crop_bbox.argtypes = [vp, vp, C.POINTER(CloudLayoutC), vp, vp]
    L.ssb_ransac_pcl_samples.argtypes = [C.c_int, C.c_int, C.c_uint, vp]
    L.ssb_organized_default_opts.argtypes = [C.POINTER(OrganizedOpts)]
    L.ssb_organized_planes.argtypes = [vp, vp, C.POINTER(CloudLayoutC), vp, C.c_int, C.POINTER(OrganizedOpts), C.c_int, vp, vp, vp, vp,
                                       vp, vp]
    L.ssb_kmeans.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.POINTER(C.c_ulonglong), vp, vp, dp]
    L.ssb_project_hull.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, C.c_int, ip]
    L.ssb_cluster_default_opts.argtypes = [C.POINTER(ClusterOpts)]
    L.ssb_cluster_planes.argtypes = [vp, vp, vp, C.c_int, vp, C.POINTER(ClusterOpts), C.POINTER(C.c_ulonglong), vp, C.c_int, ip, vp,
                                     C.c_int, ip, vp, vp]
    L.ssb_organized_last_ms.argtypes = [vp]
    L.ssb_organized_last_ms.restype = C.c_double
    _LIB = L
    return L


def last_error() -> str:
    return lib().ssb_last_error().decode("utf-8", "replace")


def check(rc: int, what: str):
    if rc < 0:
        raise SsbError(f"{what} failed ({rc}): {last_error()}")
    return rc
