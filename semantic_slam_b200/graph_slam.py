"""Python mirror of ``ps_graph_slam::GraphSLAM`` (/root/reference/include/ps_graph_slam/graph_slam.hpp:27-150,
src/ps_graph_slam/graph_slam.cpp:40-239) over the C-ABI.  Same method names, argument meaning and
error behaviour; vertices/edges are integer handles instead of g2o pointers, SE3 values are 3x4
[R|t] arrays (the top rows of the reference's Eigen::Isometry3d)."""
from __future__ import annotations

import ctypes as C
import numpy as np

from . import _lib
from ._lib import dp, ip, check


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(dp)


class GraphSLAM:
    def __init__(self, verbose: bool = False, device: int = -1, pcg_tol: float = 1e-8, max_pcg_iters: int = 20000,
                 preconditioner: int = 0, coarse_group: int = 32, force_generic: bool = False,
                 coarse_refresh: int = 1, tag_seq_start: int = 0):
        self._L = _lib.lib()
        o = _lib.GraphOpts()
        self._L.ssb_graph_default_opts(C.byref(o))
        o.device = device
        o.verbose = int(verbose)
        o.pcg_tol = pcg_tol
        o.max_pcg_iters = max_pcg_iters
        o.preconditioner = preconditioner
        o.coarse_group = coarse_group
        o.reserved[1] = int(coarse_refresh)  # re-invert the coarse matrix only every n-th damped solve
        o.reserved[3] = int(tag_seq_start)   # test hook: initial launch counter of the data-flow cell tags
        o.reserved[0] = int(force_generic)   # 1 = always use the streaming PCG kernel (no on-chip residency)
        h = self._L.ssb_graph_create(C.byref(o))
        if not h:
            raise _lib.SsbError("ssb_graph_create failed: " + _lib.last_error())
        self._h = C.c_void_p(h)
        self.verbose_ = verbose
        self.stats = None
        self.history = None

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.ssb_graph_destroy(self._h)
            self._h = None

    # ---- reference call surface ------------------------------------------------------------
    def add_se3_node(self, pose34) -> int:
        a, p = _d(pose34)
        return check(self._L.ssb_graph_add_se3_node(self._h, p), "add_se3_node")

    def add_point_xyz_node(self, xyz) -> int:
        a, p = _d(xyz)
        return check(self._L.ssb_graph_add_point_xyz_node(self._h, p), "add_point_xyz_node")

    def add_se3_edge(self, v1: int, v2: int, relative_pose34, information_matrix) -> int:
        a, p = _d(relative_pose34)
        b, q = _d(information_matrix)
        if b.size != 36:
            raise ValueError("information matrix must be 6x6")
        return check(self._L.ssb_graph_add_se3_edge(self._h, v1, v2, p, q), "add_se3_edge")

    def add_se3_point_xyz_edge(self, v_se3: int, v_xyz: int, xyz, information_matrix) -> int:
        a, p = _d(xyz)
        b, q = _d(information_matrix)
        if b.size != 9:
            raise ValueError("information matrix must be 3x3")
        return check(self._L.ssb_graph_add_se3_point_xyz_edge(self._h, v_se3, v_xyz, p, q), "add_se3_point_xyz_edge")

    def add_point_xyz_point_xyz_edge(self, v1: int, v2: int, xyz, information_matrix) -> int:
        a, p = _d(xyz)
        b, q = _d(information_matrix)
        return check(self._L.ssb_graph_add_point_xyz_point_xyz_edge(self._h, v1, v2, p, q),
                     "add_point_xyz_point_xyz_edge")

    # plane landmarks: the API the reference keeps commented out (graph_slam.hpp:44,74-75) with its own edge type
    # include/g2o/edge_se3_plane.hpp
    def add_plane_node(self, plane_coeffs) -> int:
        a, p = _d(plane_coeffs)
        if a.size != 4:
            raise ValueError("plane coefficients must be a 4-vector")
        return check(self._L.ssb_graph_add_plane_node(self._h, p), "add_plane_node")

    def add_se3_plane_edge(self, v_se3: int, v_plane: int, plane_coeffs, information_matrix) -> int:
        a, p = _d(plane_coeffs)
        b, q = _d(information_matrix)
        if a.size != 4 or b.size != 9:
            raise ValueError("plane must be a 4-vector and the information matrix 3x3")
        return check(self._L.ssb_graph_add_se3_plane_edge(self._h, v_se3, v_plane, p, q), "add_se3_plane_edge")

    def get_plane(self, vid: int):
        out = np.zeros(4)
        check(self._L.ssb_graph_get_plane(self._h, vid, out.ctypes.data_as(dp)), "get_plane")
        return out

    def set_plane(self, vid: int, coeffs):
        a, p = _d(coeffs)
        check(self._L.ssb_graph_set_plane(self._h, vid, p), "set_plane")

    def optimize(self, max_iterations: int = 1024) -> bool:
        """GraphSLAM::optimize: False when the graph has fewer than 10 edges (graph_slam.cpp:184-186)."""
        st = _lib.LmStats()
        r = check(self._L.ssb_graph_optimize(self._h, max_iterations, C.byref(st)), "optimize")
        self._store(st)
        return bool(r)

    def computeLandmarkMarginals(self, vids):
        vids = np.ascontiguousarray(vids, dtype=np.int32)
        out = np.zeros((vids.size, 3, 3))
        r = check(self._L.ssb_graph_landmark_marginals(self._h, vids.ctypes.data_as(ip), vids.size,
                                                       out.ctypes.data_as(dp)), "computeLandmarkMarginals")
        return out if r == 1 else None

    def save(self, filename: str):
        check(self._L.ssb_graph_save_g2o(self._h, filename.encode()), "save")

    def load(self, filename: str):
        check(self._L.ssb_graph_load_g2o(self._h, filename.encode()), "load")

    # ---- vertex access (node->estimate(), node->hessianIndex()) ------------------------------
    def num_vertices(self):
        return self._L.ssb_graph_num_vertices(self._h)

    def num_edges(self):
        return self._L.ssb_graph_num_edges(self._h)

    def get_se3(self, vid):
        out = np.zeros((3, 4))
        check(self._L.ssb_graph_get_se3(self._h, vid, out.ctypes.data_as(dp)), "get_se3")
        return out

    def get_point_xyz(self, vid):
        out = np.zeros(3)
        check(self._L.ssb_graph_get_point_xyz(self._h, vid, out.ctypes.data_as(dp)), "get_point_xyz")
        return out

    def set_se3(self, vid, T):
        a, p = _d(T)
        check(self._L.ssb_graph_set_se3(self._h, vid, p), "set_se3")

    def set_point_xyz(self, vid, x):
        a, p = _d(x)
        check(self._L.ssb_graph_set_point_xyz(self._h, vid, p), "set_point_xyz")

    def set_fixed(self, vid, fixed=True):
        check(self._L.ssb_graph_set_fixed(self._h, vid, int(fixed)), "set_fixed")

    def hessian_index(self, vid):
        return self._L.ssb_graph_hessian_index(self._h, vid)

    def get_all(self, n_se3, n_xyz):
        a = np.zeros((max(n_se3, 1), 3, 4))
        b = np.zeros((max(n_xyz, 1), 3))
        check(self._L.ssb_graph_get_all(self._h, a.ctypes.data_as(dp), b.ctypes.data_as(dp)), "get_all")
        return a[:n_se3], b[:n_xyz]

    def set_all(self, se3, xyz):
        a, p = _d(se3)
        b, q = _d(xyz)
        check(self._L.ssb_graph_set_all(self._h, p, q), "set_all")

    def chi2(self) -> float:
        v = C.c_double(0)
        check(self._L.ssb_graph_chi2(self._h, C.byref(v)), "chi2")
        return v.value

    # ---- benchmark / test hooks --------------------------------------------------------------
    def prepare(self):
        check(self._L.ssb_graph_prepare(self._h), "prepare")

    def invalidate(self):
        check(self._L.ssb_graph_invalidate(self._h), "invalidate")

    def snapshot(self):
        check(self._L.ssb_graph_snapshot(self._h), "snapshot")

    def restore(self):
        check(self._L.ssb_graph_restore(self._h), "restore")

    def optimize_resident(self, max_iterations: int = 1024) -> bool:
        st = _lib.LmStats()
        r = check(self._L.ssb_graph_optimize_resident(self._h, max_iterations, C.byref(st)), "optimize_resident")
        self._store(st)
        return bool(r)

    def _store(self, st):
        self.stats = {k: getattr(st, k) for k, _ in st._fields_}
        n = self._L.ssb_graph_get_history(self._h, None, 0)
        hist = np.zeros((max(n, 1), 6))
        self._L.ssb_graph_get_history(self._h, hist.ctypes.data_as(dp), n)
        self.history = hist[:n]
        self.iterations = st.iterations
        self.terminated = bool(st.terminated)

    def edge_linearize(self, eid, D, di, dj):
        err = np.zeros(6)
        Ji = np.zeros(36)
        Jj = np.zeros(36)
        check(self._L.ssb_graph_edge_linearize(self._h, eid, err.ctypes.data_as(dp), Ji.ctypes.data_as(dp),
                                               Jj.ctypes.data_as(dp)), "edge_linearize")
        return err[:D].copy(), Ji[: D * di].reshape(D, di).copy(), Jj[: D * dj].reshape(D, dj).copy()

    def solve_once(self, lam, n):
        x = np.zeros(n)
        its = check(self._L.ssb_graph_solve_once(self._h, float(lam), x.ctypes.data_as(dp), n), "solve_once")
        return its, x

    def stream(self):
        return self._L.ssb_graph_stream(self._h)

    def attach_comm(self, rank, world, unique_id: bytes):
        check(self._L.ssb_graph_attach_comm(self._h, rank, world, unique_id), "attach_comm")

    def attach_local(self, rank: int, world: int, group_key: str, cta_per_rank: int = 0):
        """Shard this graph with `world - 1` other handles driven by host threads of this process (one handle per
        thread; every handle must replay the same add_* calls).  cta_per_rank = 74 / 37: the shards share one GPU."""
        check(self._L.ssb_graph_attach_local(self._h, rank, world, group_key.encode(), cta_per_rank), "attach_local")

    def shard_info(self, world: int, rank: int):
        """(own keyframe range, local keyframes incl. ghosts, owned landmarks, touched landmarks, local edges)."""
        out = np.zeros(6, dtype=np.int32)
        check(self._L.ssb_graph_shard_info(self._h, world, rank, out.ctypes.data_as(ip)), "shard_info")
        return tuple(int(v) for v in out)
