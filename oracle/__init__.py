"""ORACLE — test infrastructure only (ctypes bindings of oracle/liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this package; the
product package (semantic_slam_b200) never does.  See oracle/oracle_graph.cpp and
oracle/oracle_ransac.cpp for the reference file:line each function follows.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_fp = C.POINTER(C.c_float)


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle_graph.cpp", "oracle_ransac.cpp", "oracle_segment.cpp", "oracle_cluster.cpp")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.orc_graph_create.restype = C.c_void_p
        L.orc_graph_destroy.argtypes = [C.c_void_p]
        L.orc_graph_set_threads.argtypes = [C.c_void_p, C.c_int]
        L.orc_graph_add_se3_node.argtypes = [C.c_void_p, _dp]
        L.orc_graph_add_point_xyz_node.argtypes = [C.c_void_p, _dp]
        L.orc_graph_add_se3_edge.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp]
        L.orc_graph_add_se3_point_xyz_edge.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp]
        L.orc_graph_add_point_xyz_point_xyz_edge.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp]
        L.orc_graph_add_plane_node.argtypes = [C.c_void_p, _dp]
        L.orc_graph_add_se3_plane_edge.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp]
        L.orc_graph_get_plane.argtypes = [C.c_void_p, C.c_int, _dp]
        L.orc_graph_set_plane.argtypes = [C.c_void_p, C.c_int, _dp]
        L.orc_plane_oplus.argtypes = [_dp, _dp]
        L.orc_graph_num_vertices.argtypes = [C.c_void_p]
        L.orc_graph_num_edges.argtypes = [C.c_void_p]
        L.orc_graph_get_se3.argtypes = [C.c_void_p, C.c_int, _dp]
        L.orc_graph_set_se3.argtypes = [C.c_void_p, C.c_int, _dp]
        L.orc_graph_get_point_xyz.argtypes = [C.c_void_p, C.c_int, _dp]
        L.orc_graph_set_point_xyz.argtypes = [C.c_void_p, C.c_int, _dp]
        L.orc_graph_set_fixed.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_graph_get_all.argtypes = [C.c_void_p, _dp, _dp]
        L.orc_graph_chi2.argtypes = [C.c_void_p]
        L.orc_graph_chi2.restype = C.c_double
        L.orc_graph_optimize.argtypes = [C.c_void_p, C.c_int, _ip, _ip, _dp, C.c_int]
        L.orc_graph_last_stats.argtypes = [C.c_void_p, _dp]
        L.orc_graph_edge_linearize.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp]
        L.orc_graph_dense_system.argtypes = [C.c_void_p, _dp, _dp, _ip]
        L.orc_graph_sparse_system.argtypes = [C.c_void_p, _ip, _ip, _dp, _dp, _ip]
        L.orc_graph_sparse_system.restype = C.c_longlong
        L.orc_graph_num_scalar.argtypes = [C.c_void_p]
        L.orc_graph_solve_once.argtypes = [C.c_void_p, C.c_double, _dp]
        L.orc_graph_landmark_marginals.argtypes = [C.c_void_p, _ip, C.c_int, _dp, C.c_int]
        L.orc_graph_landmark_marginals_g2o.argtypes = [C.c_void_p, _ip, C.c_int, _dp, C.c_int, C.POINTER(C.c_longlong)]
        L.orc_to_vector_mqt.argtypes = [_dp, _dp]
        L.orc_from_vector_mqt.argtypes = [_dp, _dp]
        L.orc_se3_oplus.argtypes = [_dp, _dp, _dp]
        L.orc_ransac_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, _ip, _ip, C.c_int, _ip, C.c_int,
                                       C.c_double, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p,
                                       C.c_void_p]
        L.orc_ransac_batch.restype = C.c_int
        L.orc_crop.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, _ip, _ip, _fp]
        L.orc_crop.restype = C.c_int
        _LIB = L
    return _LIB


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


class OracleGraphSLAM:
    """Same call surface as ps_graph_slam::GraphSLAM (graph_slam.hpp:27-150), CPU oracle behind it."""

    def __init__(self, verbose: bool = False, threads: int = 1):
        self._L = lib()
        self._h = C.c_void_p(self._L.orc_graph_create())
        self._L.orc_graph_set_threads(self._h, threads)
        self.verbose_ = verbose
        self.history = None
        self.iterations = 0
        self.terminated = False

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_graph_destroy(self._h)
            self._h = None

    def add_se3_node(self, pose34):
        a, p = _d(pose34)
        return self._L.orc_graph_add_se3_node(self._h, p)

    def add_point_xyz_node(self, xyz):
        a, p = _d(xyz)
        return self._L.orc_graph_add_point_xyz_node(self._h, p)

    def add_se3_edge(self, v1, v2, relative_pose34, information):
        a, p = _d(relative_pose34)
        b, q = _d(information)
        r = self._L.orc_graph_add_se3_edge(self._h, v1, v2, p, q)
        if r < 0:
            raise ValueError("bad vertex ids")
        return r

    def add_se3_point_xyz_edge(self, v_se3, v_xyz, xyz, information):
        a, p = _d(xyz)
        b, q = _d(information)
        r = self._L.orc_graph_add_se3_point_xyz_edge(self._h, v_se3, v_xyz, p, q)
        if r < 0:
            raise ValueError("bad vertex ids")
        return r

    def add_point_xyz_point_xyz_edge(self, v1, v2, xyz, information):
        a, p = _d(xyz)
        b, q = _d(information)
        r = self._L.orc_graph_add_point_xyz_point_xyz_edge(self._h, v1, v2, p, q)
        if r < 0:
            raise ValueError("bad vertex ids")
        return r

    def num_vertices(self):
        return self._L.orc_graph_num_vertices(self._h)

    def num_edges(self):
        return self._L.orc_graph_num_edges(self._h)

    def optimize(self, max_iterations: int = 1024) -> bool:
        it = C.c_int(0)
        term = C.c_int(0)
        hist = np.zeros((max(max_iterations, 1), 5))
        r = self._L.orc_graph_optimize(self._h, max_iterations, C.byref(it), C.byref(term),
                                       hist.ctypes.data_as(_dp), hist.shape[0])
        self.iterations = it.value
        self.terminated = bool(term.value)
        self.history = hist[: max(it.value, 0)]
        return bool(r)

    def chi2(self) -> float:
        return self._L.orc_graph_chi2(self._h)

    # dormant plane API of the reference (graph_slam.hpp:44,74-75; include/g2o/edge_se3_plane.hpp)
    def add_plane_node(self, plane_coeffs):
        a, p = _d(plane_coeffs)
        return self._L.orc_graph_add_plane_node(self._h, p)

    def add_se3_plane_edge(self, v_se3, v_plane, plane_coeffs, information):
        a, p = _d(plane_coeffs)
        b, q = _d(information)
        r = self._L.orc_graph_add_se3_plane_edge(self._h, v_se3, v_plane, p, q)
        if r < 0:
            raise ValueError("bad vertex ids")
        return r

    def get_plane(self, vid):
        out = np.zeros(4)
        if self._L.orc_graph_get_plane(self._h, vid, out.ctypes.data_as(_dp)) != 0:
            raise ValueError("not a plane vertex")
        return out

    def set_plane(self, vid, c):
        a, p = _d(c)
        self._L.orc_graph_set_plane(self._h, vid, p)

    def get_se3(self, vid):
        out = np.zeros((3, 4))
        if self._L.orc_graph_get_se3(self._h, vid, out.ctypes.data_as(_dp)) != 0:
            raise ValueError("not an SE3 vertex")
        return out

    def get_point_xyz(self, vid):
        out = np.zeros(3)
        if self._L.orc_graph_get_point_xyz(self._h, vid, out.ctypes.data_as(_dp)) != 0:
            raise ValueError("not an XYZ vertex")
        return out

    def set_se3(self, vid, T):
        a, p = _d(T)
        self._L.orc_graph_set_se3(self._h, vid, p)

    def set_point_xyz(self, vid, x):
        a, p = _d(x)
        self._L.orc_graph_set_point_xyz(self._h, vid, p)

    def get_all(self, n_se3, n_xyz):
        a = np.zeros((n_se3, 3, 4))
        b = np.zeros((max(n_xyz, 1), 3))
        self._L.orc_graph_get_all(self._h, a.ctypes.data_as(_dp), b.ctypes.data_as(_dp))
        return a, b[:n_xyz]

    def last_stats(self):
        out = np.zeros(5)
        self._L.orc_graph_last_stats(self._h, out.ctypes.data_as(_dp))
        return dict(analyze_ms=out[0], factor_ms=out[1], linearize_ms=out[2], nnzL=int(out[3]), n=int(out[4]))

    def computeLandmarkMarginals(self, vids, relinearize=False, method="solve"):
        """3x3 blocks of H^-1.  method "solve": one pair of triangular solves per column (the checker of the GPU tests);
        "g2o": MarginalCovarianceCholesky's memoised recursion over the factor (what the reference's computeMarginals
        runs, graph_slam.cpp:225) — the timed CPU baseline of bench.py; self.marginal_map_entries = elements it computed."""
        vids = np.ascontiguousarray(vids, dtype=np.int32)
        out = np.zeros((vids.size, 3, 3))
        if method == "g2o":
            cnt = C.c_longlong(0)
            r = self._L.orc_graph_landmark_marginals_g2o(self._h, vids.ctypes.data_as(_ip), vids.size,
                                                         out.ctypes.data_as(_dp), int(relinearize), C.byref(cnt))
            self.marginal_map_entries = int(cnt.value)
        else:
            r = self._L.orc_graph_landmark_marginals(self._h, vids.ctypes.data_as(_ip), vids.size,
                                                     out.ctypes.data_as(_dp), int(relinearize))
        if r != 1:
            raise RuntimeError("marginals failed")
        return out

    # ---- test hooks ----
    def edge_linearize(self, eid, D, di, dj):
        err = np.zeros(6)
        Ji = np.zeros(36)
        Jj = np.zeros(36)
        self._L.orc_graph_edge_linearize(self._h, eid, err.ctypes.data_as(_dp), Ji.ctypes.data_as(_dp),
                                         Jj.ctypes.data_as(_dp))
        return err[:D].copy(), Ji[: D * di].reshape(D, di).copy(), Jj[: D * dj].reshape(D, dj).copy()

    def dense_system(self):
        nv = self.num_vertices()
        off = np.zeros(nv, dtype=np.int32)
        n = self._L.orc_graph_dense_system(self._h, None, None, off.ctypes.data_as(_ip))
        H = np.zeros((n, n))
        b = np.zeros(n)
        self._L.orc_graph_dense_system(self._h, H.ctypes.data_as(_dp), b.ctypes.data_as(_dp), off.ctypes.data_as(_ip))
        return H, b, off

    def sparse_system(self):
        """(H as scipy.sparse CSR full symmetric, b, scalar offset per vertex)"""
        import scipy.sparse as sp
        nv = self.num_vertices()
        off = np.zeros(nv, dtype=np.int32)
        nnz = self._L.orc_graph_sparse_system(self._h, None, None, None, None, off.ctypes.data_as(_ip))
        n = self._L.orc_graph_num_scalar(self._h)
        r = np.zeros(nnz, dtype=np.int32)
        c = np.zeros(nnz, dtype=np.int32)
        v = np.zeros(nnz)
        b = np.zeros(n)
        self._L.orc_graph_sparse_system(self._h, r.ctypes.data_as(_ip), c.ctypes.data_as(_ip), v.ctypes.data_as(_dp),
                                        b.ctypes.data_as(_dp), off.ctypes.data_as(_ip))
        U = sp.coo_matrix((v, (r, c)), shape=(n, n)).tocsr()
        U.eliminate_zeros()
        Hfull = U + sp.triu(U, k=1).T
        return Hfull.tocsr(), b, off

    def solve_once(self, lam):
        n = self._L.orc_graph_dense_system(self._h, None, None, None)
        x = np.zeros(n)
        ok = self._L.orc_graph_solve_once(self._h, float(lam), x.ctypes.data_as(_dp))
        return bool(ok), x


def to_vector_mqt(T):
    a, p = _d(T)
    v = np.zeros(6)
    lib().orc_to_vector_mqt(p, v.ctypes.data_as(_dp))
    return v


def from_vector_mqt(v):
    a, p = _d(v)
    T = np.zeros((3, 4))
    lib().orc_from_vector_mqt(p, T.ctypes.data_as(_dp))
    return T


def plane_oplus(c4, v3):
    """g2o::Plane3D::oplus"""
    c = np.array(c4, dtype=np.float64).copy()
    v = np.ascontiguousarray(v3, dtype=np.float64)
    lib().orc_plane_oplus(c.ctypes.data_as(_dp), v.ctypes.data_as(_dp))
    return c


def se3_oplus(T, v):
    a, p = _d(T)
    b, q = _d(v)
    out = np.zeros((3, 4))
    lib().orc_se3_oplus(p, q, out.ctypes.data_as(_dp))
    return out


PLANE_RESULT_DTYPE = np.dtype([("status", "i4"), ("n_points", "i4"), ("best_hyp", "i4"), ("best_count", "i4"),
                               ("iterations", "i4"), ("refined_count", "i4"), ("coef", "f4", 4), ("refined", "f4", 4),
                               ("centroid", "f4", 3), ("reserved", "i4")])


def crop(msg, width, height, point_step, row_step, offsets, box):
    """plane_segmentation::segmentPointCloudData restated. Returns (h, w, 4) float32 or None if spurious."""
    msg = np.ascontiguousarray(msg, dtype=np.uint8)
    off = np.ascontiguousarray(offsets, dtype=np.int32)
    bx = np.ascontiguousarray(box, dtype=np.int32)
    n = lib().orc_crop(msg.ctypes.data, width, height, point_step, row_step, off.ctypes.data_as(_ip),
                       bx.ctypes.data_as(_ip), None)
    if n < 0:
        return None
    out = np.zeros((int(bx[3]), int(bx[2]), 4), dtype=np.float32)
    lib().orc_crop(msg.ctypes.data, width, height, point_step, row_step, off.ctypes.data_as(_ip),
                   bx.ctypes.data_as(_ip), out.ctypes.data_as(_fp))
    return out


def ransac_plane_batch(msg, width, height, point_step, row_step, offsets, boxes, triples, threshold=0.01,
                       refine=True, mode=0, max_iterations=50, probability=0.99, want_counts=True, want_mask=True):
    """pcl::SACSegmentation(SACMODEL_PLANE, SAC_RANSAC) over every bbox crop, restated on the CPU.
    Returns (results[nb] structured array, counts[nb,K] or None, mask (concatenated uint8) or None)."""
    msg = np.ascontiguousarray(msg, dtype=np.uint8)
    off = np.ascontiguousarray(offsets, dtype=np.int32)
    boxes = np.ascontiguousarray(boxes, dtype=np.int32).reshape(-1, 4)
    triples = np.ascontiguousarray(triples, dtype=np.int32)
    nb = boxes.shape[0]
    K = triples.shape[1] if nb else 0
    res = np.zeros(nb, dtype=PLANE_RESULT_DTYPE)
    counts = np.zeros((nb, K), dtype=np.int32) if want_counts else None
    valid = (boxes[:, 2] >= 0) & (boxes[:, 3] >= 0) & (boxes[:, 0] >= 0) & (boxes[:, 1] >= 0) & \
            (boxes[:, 0] + boxes[:, 2] <= width) & (boxes[:, 1] + boxes[:, 3] <= height)
    total = int((boxes[valid, 2].astype(np.int64) * boxes[valid, 3]).sum())
    mask = np.zeros(max(total, 1), dtype=np.uint8) if want_mask else None
    r = lib().orc_ransac_batch(msg.ctypes.data, width, height, point_step, row_step, off.ctypes.data_as(_ip),
                               boxes.ctypes.data_as(_ip), nb, triples.ctypes.data_as(_ip), K, float(threshold),
                               int(refine), int(mode), int(max_iterations), float(probability), res.ctypes.data,
                               counts.ctypes.data if counts is not None else None,
                               mask.ctypes.data if mask is not None else None)
    if r != 0:
        raise RuntimeError("oracle ransac failed")
    return res, counts, (mask[:total] if mask is not None else None)


# ---- PCL's RANSAC sample stream, restated in pure Python (independent of the product's C++ and of libstdc++) ----------
class _MT19937:
    """Matsumoto & Nishimura's MT19937 with init_genrand seeding = boost::mt19937(seed) = std::mt19937(seed).
    Known answer (C++11 [rand.predef]): the 10000th output of the default-seeded (5489) engine is 4123659995."""

    def __init__(self, seed):
        mt = [0] * 624
        mt[0] = seed & 0xFFFFFFFF
        for i in range(1, 624):
            mt[i] = (1812433253 * (mt[i - 1] ^ (mt[i - 1] >> 30)) + i) & 0xFFFFFFFF
        self.mt, self.idx = mt, 624

    def __call__(self):
        mt = self.mt
        if self.idx >= 624:
            for k in range(624):
                y = (mt[k] & 0x80000000) | (mt[(k + 1) % 624] & 0x7FFFFFFF)
                mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
            self.idx = 0
        y = mt[self.idx]
        self.idx += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF


def _boost_uniform_int(engine, lo, hi):
    """boost::random::detail::generate_uniform_int for a 32-bit engine (brange = 2^32 - 1) and range = hi - lo < brange:
    bucket_size = brange / (range + 1), incremented when brange % (range + 1) == range; result = engine() / bucket_size,
    redrawn while it exceeds the range.  For uniform_int<>(0, INT_MAX) — PCL's rng_dist_ — bucket_size is 2 and nothing is
    ever redrawn."""
    rng = hi - lo
    brange = 0xFFFFFFFF
    assert 0 < rng < brange
    bucket = brange // (rng + 1)
    if brange % (rng + 1) == rng:
        bucket += 1
    while True:
        r = engine() // bucket
        if r <= rng:
            return lo + r


def pcl_sample_stream_c(n_points, n_draws, seed=12345):
    """the C++ twin of pcl_sample_stream used inside the oracle's clustering chain (oracle_ransac.cpp, own engine)"""
    out = np.zeros((n_draws, 3), dtype=np.int32)
    f = lib().orc_pcl_sample_stream
    f.argtypes = [C.c_int, C.c_int, C.c_uint, C.c_void_p]
    f.restype = None
    f(int(n_points), int(n_draws), int(seed), out.ctypes.data)
    return out


def pcl_sample_stream(n_points, n_draws, seed=12345):
    """pcl::SampleConsensusModel::drawIndexSample, n_draws times on a fresh model over n_points points
    (pcl/sample_consensus/sac_model.h): shuffled_indices_ starts as 0..n-1 and keeps its state between draws; every draw
    swaps element i (i = 0, 1, 2) with element i + rnd() % (n - i) and returns the first three.  Restated for
    tests/test_oracle_ransac.py, which compares the product's ssb_ransac_pcl_samples with it bit for bit."""
    out = np.zeros((n_draws, 3), dtype=np.int32)
    if n_points < 3:
        return out
    eng = _MT19937(seed)
    sh = list(range(n_points))
    for d in range(n_draws):
        for i in range(3):
            j = i + _boost_uniform_int(eng, 0, 2**31 - 1) % (n_points - i)
            sh[i], sh[j] = sh[j], sh[i]
        out[d] = sh[:3]
    return out


# ---- the live segmentation path (oracle_segment.cpp): integral-image normals + organised multi-plane segmentation ----
def integral_normals(cloud_hw4, max_depth_change_factor=0.03, smoothing_size=20.0):
    """pcl::IntegralImageNormalEstimation (COVARIANCE_MATRIX) on an organised (h, w, 4) float32 crop.
    Returns (normals (h, w, 4): nx ny nz curvature, distance map (h, w))."""
    c = np.ascontiguousarray(cloud_hw4, dtype=np.float32)
    h, w = c.shape[:2]
    nrm = np.zeros((h, w, 4), dtype=np.float32)
    dist = np.zeros((h, w), dtype=np.float32)
    L = lib()
    L.orc_integral_normals.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    L.orc_integral_normals(c.ctypes.data, w, h, max_depth_change_factor, smoothing_size, nrm.ctypes.data, dist.ctypes.data)
    return nrm, dist


def organized_planes(cloud_hw4, min_inliers=500, angular_threshold=0.017453 * 2.0, distance_threshold=0.02,
                     maximum_curvature=0.001, max_depth_change_factor=0.03, smoothing_size=20.0, max_regions=64):
    """IntegralImageNormalEstimation + OrganizedMultiPlaneSegmentation::segmentAndRefine + contour + polygon area."""
    c = np.ascontiguousarray(cloud_hw4, dtype=np.float32)
    h, w = c.shape[:2]
    nrm = np.zeros((h, w, 4), dtype=np.float32)
    lcc = np.zeros((h, w), dtype=np.int32)
    lref = np.zeros((h, w), dtype=np.int32)
    cen = np.zeros((max_regions, 3), dtype=np.float32)
    mod = np.zeros((max_regions, 4), dtype=np.float32)
    nin = np.zeros(max_regions, dtype=np.int32)
    ncon = np.zeros(max_regions, dtype=np.int32)
    area = np.zeros(max_regions, dtype=np.float32)
    L = lib()
    L.orc_organized_planes.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_float, C.c_float, C.c_float,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p]
    n = L.orc_organized_planes(c.ctypes.data, w, h, max_depth_change_factor, smoothing_size, int(min_inliers), angular_threshold,
                               distance_threshold, maximum_curvature, nrm.ctypes.data, lcc.ctypes.data, lref.ctypes.data, max_regions,
                               cen.ctypes.data, mod.ctypes.data, nin.ctypes.data, ncon.ctypes.data, area.ctypes.data)
    if n < 0:
        raise RuntimeError("oracle organized_planes failed")
    m = min(n, max_regions)
    return {"n": n, "normals": nrm, "labels_cc": lcc, "labels": lref, "centroid": cen[:m], "model": mod[:m], "n_inliers": nin[:m],
            "contour_points": ncon[:m], "area": area[:m]}


# ---- the dormant plane-clustering chain (oracle_cluster.cpp) ---------------------------------------------------------------
PLANE_CLUSTER_DTYPE = np.dtype([("normal", np.float32, 3), ("distance", np.float32), ("normal_label", np.int32),
                                ("distance_label", np.int32), ("n_points", np.int32), ("n_inliers", np.int32),
                                ("coef", np.float32, 4), ("row0", np.int32), ("n_rows", np.int32)])


def kmeans(data, K, rng_state=0xFFFFFFFF, attempts=10, max_count=10, eps=0.01):
    """cv::kmeans(data, K, labels, TermCriteria(EPS + COUNT, max_count, eps), attempts, KMEANS_RANDOM_CENTERS, centers) restated
    (plane_segmentation.cpp:525-535).  rng_state = cv::theRNG().state before the call.
    Returns (compactness, labels (N,), centers (K, dims), rng state after)."""
    d = np.ascontiguousarray(data, dtype=np.float32)
    if d.ndim == 1:
        d = d[:, None]
    n, dims = d.shape
    lab = np.zeros(n, dtype=np.int32)
    cen = np.zeros((K, dims), dtype=np.float32)
    st = C.c_ulonglong(rng_state if rng_state else 0xFFFFFFFF)
    L = lib()
    L.orc_kmeans.restype = C.c_double
    L.orc_kmeans.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.POINTER(C.c_ulonglong),
                             C.c_void_p, C.c_void_p]
    c = L.orc_kmeans(d.ctypes.data, n, dims, K, max_count, eps, attempts, C.byref(st), lab.ctypes.data, cen.ctypes.data)
    return c, lab, cen, st.value


def project_hull(pts4, mask, coef):
    """pcl::ProjectInliers(SACMODEL_PLANE) + pcl::ConvexHull (plane_segmentation.cpp:649-662) restated.
    Returns (hull vertices (k, 3) in PCL's output order, their indices in pts4, number of projected inliers)."""
    p = np.ascontiguousarray(pts4, dtype=np.float32).reshape(-1, 4)
    m = np.ascontiguousarray(mask, dtype=np.uint8)
    cf = np.ascontiguousarray(coef, dtype=np.float32)
    n = p.shape[0]
    rows = np.zeros((max(n, 1), 3), dtype=np.float32)
    src = np.zeros(max(n, 1), dtype=np.int32)
    nin = C.c_int(0)
    L = lib()
    L.orc_project_hull.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    k = L.orc_project_hull(p.ctypes.data, m.ctypes.data, n, cf.ctypes.data, rows.ctypes.data, src.ctypes.data, n, C.byref(nin))
    return rows[:k].copy(), src[:k].copy(), nin.value


def cluster_planes(cloud4, normals4, T, rng_state=0xFFFFFFFF, num_centroids_normals=4, num_centroids_distance=2, attempts=10,
                   max_count=10, eps=0.01, min_cluster_points=500, centroid_tolerance=0.3, ransac_hypotheses=0, ransac_seed=12345,
                   coef_override=None, max_rows=65536, max_clusters=16):
    """plane_segmentation::clusterAndSegmentAllPlanes (plane_segmentation.cpp:261-294) restated.
    Returns dict(rows (k, 8), clusters (structured), labels (n,), centers (Kn, 3), rng_state)."""
    c = np.ascontiguousarray(cloud4, dtype=np.float32).reshape(-1, 4)
    q = np.ascontiguousarray(normals4, dtype=np.float32).reshape(-1, 4)
    n = c.shape[0]
    assert q.shape[0] == n
    T16 = np.ascontiguousarray(T, dtype=np.float32).reshape(16)
    rows = np.zeros((max_rows, 8), dtype=np.float32)
    cl = np.zeros(max_clusters, dtype=PLANE_CLUSTER_DTYPE)
    lab = np.zeros(n, dtype=np.int32)
    cen = np.zeros((num_centroids_normals, 3), dtype=np.float32)
    st = C.c_ulonglong(rng_state if rng_state else 0xFFFFFFFF)
    nr, nc = C.c_int(0), C.c_int(0)
    ov = None
    if coef_override is not None:
        ov = np.ascontiguousarray(coef_override, dtype=np.float32).reshape(-1, 4)
        assert ov.shape[0] >= max_clusters or True
    L = lib()
    L.orc_cluster_planes.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                                     C.c_float, C.c_int, C.c_uint, C.POINTER(C.c_ulonglong), C.c_void_p, C.c_int, C.POINTER(C.c_int),
                                     C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_cluster_planes(c.ctypes.data, q.ctypes.data, n, T16.ctypes.data, num_centroids_normals, num_centroids_distance, attempts,
                         max_count, eps, min_cluster_points, centroid_tolerance, ransac_hypotheses, ransac_seed, C.byref(st),
                         rows.ctypes.data, max_rows, C.byref(nr), cl.ctypes.data, max_clusters, C.byref(nc), lab.ctypes.data,
                         cen.ctypes.data, ov.ctypes.data if ov is not None else None)
    return dict(rows=rows[:nr.value].copy(), clusters=cl[:nc.value].copy(), labels=lab, centers=cen, rng_state=st.value)
