"""Python mirror of the planar_segmentation RANSAC path over the C-ABI:
``plane_segmentation::segmentPointCloudData`` (bbox crop, plane_segmentation.cpp:24-82) and the
``pcl::SACSegmentation`` plane fit of ``compute2DConvexHull`` (plane_segmentation.cpp:631-647),
batched over all bounding boxes of a frame."""
from __future__ import annotations

import ctypes as C
import dataclasses
import numpy as np

from . import _lib
from ._lib import check

PLANE_RESULT_DTYPE = np.dtype([("status", "i4"), ("n_points", "i4"), ("best_hyp", "i4"), ("best_count", "i4"),
                               ("iterations", "i4"), ("refined_count", "i4"), ("coef", "f4", 4), ("refined", "f4", 4),
                               ("centroid", "f4", 3), ("reserved", "i4")])


@dataclasses.dataclass
class CloudLayout:
    width: int = 640
    height: int = 480
    point_step: int = 32
    row_step: int = 32 * 640
    offsets: tuple = (0, 4, 8, 16)

    def c(self):
        return _lib.CloudLayoutC(self.width, self.height, self.point_step, self.row_step, *self.offsets)


def pcl_sample_stream(n_points, n_draws, seed=12345):
    """The 3-point sample stream pcl::RandomSampleConsensus draws for a fresh model over `n_points` points
    (ssb_ransac_pcl_samples: boost::mt19937(seed) >> 1, drawIndexSample's running shuffle; host only).
    n_points: int or a sequence (one crop each) -> int32 [n_draws, 3] or [n_crops, n_draws, 3]."""
    L = _lib.lib()
    if np.ndim(n_points) == 0:
        out = np.zeros((n_draws, 3), dtype=np.int32)
        check(L.ssb_ransac_pcl_samples(int(n_points), int(n_draws), int(seed), out.ctypes.data), "ssb_ransac_pcl_samples")
        return out
    ns = [int(v) for v in n_points]
    out = np.zeros((len(ns), n_draws, 3), dtype=np.int32)
    for b, n in enumerate(ns):
        check(L.ssb_ransac_pcl_samples(n, int(n_draws), int(seed), out[b].ctypes.data), "ssb_ransac_pcl_samples")
    return out


class PlaneSegmentation:
    """plane_segmentation (the RANSAC part): persistent device buffers + stream."""

    def __init__(self, device: int = -1, threshold: float = 0.01, refine: bool = True, mode: int = 0,
                 max_iterations: int = 50, probability: float = 0.99):
        self._L = _lib.lib()
        h = self._L.ssb_ransac_create(device)
        if not h:
            raise _lib.SsbError("ssb_ransac_create failed: " + _lib.last_error())
        self._h = C.c_void_p(h)
        self.opts = _lib.RansacOpts()
        self._L.ssb_ransac_default_opts(C.byref(self.opts))
        self.opts.threshold = threshold
        self.opts.refine = int(refine)
        self.opts.mode = mode
        self.opts.max_iterations = max_iterations
        self.opts.probability = probability
        self._shape = None
        self._pinned = None
        self._pin_cache = {}

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.ssb_ransac_destroy(self._h)
            self._h = None

    def segmentPointCloudData(self, box, msg, layout: CloudLayout):
        """bbox crop: returns (h, w, 4) float32 [x,y,z,rgb] or None for a 'spurious' box."""
        msg = np.ascontiguousarray(msg, dtype=np.uint8)
        bx = np.ascontiguousarray(box, dtype=np.int32)
        lc = layout.c()
        n = self._L.ssb_crop_bbox(self._h, msg.ctypes.data, C.byref(lc), bx.ctypes.data, None)
        if n == -1:
            return None
        check(n, "ssb_crop_bbox")
        out = np.zeros((int(bx[3]), int(bx[2]), 4), dtype=np.float32)
        if n > 0:
            check(self._L.ssb_crop_bbox(self._h, msg.ctypes.data, C.byref(lc), bx.ctypes.data, out.ctypes.data),
                  "ssb_crop_bbox")
        return out

    def _zeros(self, shape, dtype):
        """result buffers: page-locked when torch is importable (D2H at PCIe speed), else plain numpy"""
        if self._pinned is None:
            try:
                import torch
                self._pinned = torch.cuda.is_available()
            except Exception:
                self._pinned = False
        if self._pinned:
            import torch
            n = int(np.prod(shape)) * np.dtype(dtype).itemsize
            key = (tuple(np.atleast_1d(shape)), np.dtype(dtype).str)
            buf = self._pin_cache.get(key)
            if buf is None:
                buf = torch.zeros(max(n, 1), dtype=torch.uint8).pin_memory()
                self._pin_cache[key] = buf
            return buf.numpy()[:n].view(dtype).reshape(shape)
        return np.zeros(shape, dtype=dtype)

    def _alloc(self, boxes, K, layout, want_counts, want_mask):
        nb = boxes.shape[0]
        res = self._zeros(nb, PLANE_RESULT_DTYPE)
        counts = self._zeros((nb, K), np.int32) if want_counts else None
        valid = (boxes[:, 2] >= 0) & (boxes[:, 3] >= 0) & (boxes[:, 0] >= 0) & (boxes[:, 1] >= 0) & \
                (boxes[:, 0] + boxes[:, 2] <= layout.width) & (boxes[:, 1] + boxes[:, 3] <= layout.height)
        total = int((boxes[valid, 2].astype(np.int64) * boxes[valid, 3]).sum())
        mask = self._zeros(max(total, 1), np.uint8) if want_mask else None
        return res, counts, mask, total

    def fit_planes(self, msg, layout: CloudLayout, boxes, triples, want_counts=True, want_mask=True):
        """RANSAC plane fit of every bbox crop (host buffers in, host buffers out)."""
        msg = np.ascontiguousarray(msg, dtype=np.uint8)
        boxes = np.ascontiguousarray(boxes, dtype=np.int32).reshape(-1, 4)
        triples = np.ascontiguousarray(triples, dtype=np.int32)
        nb = boxes.shape[0]
        K = triples.shape[1] if triples.ndim == 3 else 0
        res, counts, mask, total = self._alloc(boxes, K, layout, want_counts, want_mask)
        lc = layout.c()
        check(self._L.ssb_ransac_plane_batch(self._h, msg.ctypes.data, C.byref(lc), boxes.ctypes.data, nb,
                                             triples.ctypes.data, K, C.byref(self.opts), res.ctypes.data,
                                             counts.ctypes.data if counts is not None else None,
                                             mask.ctypes.data if mask is not None else None), "ssb_ransac_plane_batch")
        return res, counts, (mask[:total] if mask is not None else None)

    def fit_planes_pcl(self, msg, layout: CloudLayout, boxes, want_mask=True, n_draws=512):
        """pcl::SACSegmentation's own behaviour per crop (what compute2DConvexHull runs, plane_segmentation.cpp:637-647):
        PCL's sample stream (a fresh model per crop, seed 12345) under PCL's adaptive stopping rule — the handle must have
        been created with mode=1.  Spurious boxes get an unused stream of zeros."""
        if self.opts.mode != 1:
            raise ValueError("fit_planes_pcl needs PlaneSegmentation(mode=1) (PCL's adaptive stopping rule)")
        boxes = np.ascontiguousarray(boxes, dtype=np.int32).reshape(-1, 4)
        n = np.where((boxes[:, 2] >= 0) & (boxes[:, 3] >= 0), boxes[:, 2].astype(np.int64) * boxes[:, 3], 0)
        triples = pcl_sample_stream(n, n_draws)
        res, _, mask = self.fit_planes(msg, layout, boxes, triples, want_counts=False, want_mask=want_mask)
        return res, mask, triples

    # ---- device-resident variant (benchmark `value` leg) --------------------------------------
    def upload(self, msg, layout: CloudLayout, boxes, triples):
        msg = np.ascontiguousarray(msg, dtype=np.uint8)
        boxes = np.ascontiguousarray(boxes, dtype=np.int32).reshape(-1, 4)
        triples = np.ascontiguousarray(triples, dtype=np.int32)
        K = triples.shape[1] if triples.ndim == 3 else 0
        lc = layout.c()
        check(self._L.ssb_ransac_upload(self._h, msg.ctypes.data, C.byref(lc), boxes.ctypes.data, boxes.shape[0],
                                        triples.ctypes.data, K, C.byref(self.opts)), "ssb_ransac_upload")
        self._shape = (boxes.copy(), K, layout)

    def run_resident(self):
        check(self._L.ssb_ransac_run_resident(self._h), "ssb_ransac_run_resident")

    def fetch(self, want_counts=True, want_mask=True):
        boxes, K, layout = self._shape
        res, counts, mask, total = self._alloc(boxes, K, layout, want_counts, want_mask)
        check(self._L.ssb_ransac_fetch(self._h, res.ctypes.data, counts.ctypes.data if counts is not None else None,
                                       mask.ctypes.data if mask is not None else None), "ssb_ransac_fetch")
        return res, counts, (mask[:total] if mask is not None else None)

    def timing(self):
        """(ms of the whole device pipeline, ms of the point x hypothesis sweep) of the last run"""
        out = np.zeros(2)
        check(self._L.ssb_ransac_timing(self._h, out.ctypes.data_as(_lib.dp)), "ssb_ransac_timing")
        return float(out[0]), float(out[1])

    def stream(self):
        return self._L.ssb_ransac_stream(self._h)

    def launch_count(self):
        return self._L.ssb_ransac_launch_count(self._h)


PLANAR_REGION_DTYPE = np.dtype([("centroid", "f4", 3), ("model", "f4", 4), ("contour_points", "i4"), ("area", "f4")])


class OrganizedSegmentation(PlaneSegmentation):
    """The LIVE segmentation path of the reference on the device: ``plane_segmentation::computeNormalsFromPointCloud``
    (plane_segmentation.cpp:84-106, pcl::IntegralImageNormalEstimation) + the PCL call of ``multiPlaneSegmentation``
    (:136-156, pcl::OrganizedMultiPlaneSegmentation::segmentAndRefine) + contour / polygon area (:169,189), batched over
    all bounding boxes of a frame (include/ssb.h: ssb_organized_planes)."""

    def __init__(self, device: int = -1, num_point_seg: int = 500, norm_point_thres: int = 5000, **kw):
        super().__init__(device=device)
        self.oopts = _lib.OrganizedOpts()
        self._L.ssb_organized_default_opts(C.byref(self.oopts))
        self.oopts.min_inliers = int(num_point_seg)
        self.oopts.norm_point_thres = int(norm_point_thres)
        for k, v in kw.items():
            setattr(self.oopts, k, v)

    def segment(self, msg, layout: CloudLayout, boxes, max_regions: int = 16, want_points: bool = False):
        """Returns (regions [nb, max_regions] structured, n_regions [nb], n_inliers [nb, max_regions]) and, with
        want_points, per-point normals / labels / distance map concatenated over the non-spurious boxes."""
        msg = np.ascontiguousarray(msg, dtype=np.uint8)
        boxes = np.ascontiguousarray(boxes, dtype=np.int32).reshape(-1, 4)
        nb = boxes.shape[0]
        reg = np.zeros((nb, max_regions), dtype=PLANAR_REGION_DTYPE)
        nreg = np.zeros(nb, dtype=np.int32)
        nin = np.zeros((nb, max_regions), dtype=np.int32)
        valid = (boxes[:, 2] >= 0) & (boxes[:, 3] >= 0) & (boxes[:, 0] >= 0) & (boxes[:, 1] >= 0) & \
                (boxes[:, 0] + boxes[:, 2] <= layout.width) & (boxes[:, 1] + boxes[:, 3] <= layout.height)
        total = int((boxes[valid, 2].astype(np.int64) * boxes[valid, 3]).sum())
        nrm = lab = dist = None
        if want_points:
            nrm = np.zeros((max(total, 1), 4), dtype=np.float32)
            lab = np.zeros(max(total, 1), dtype=np.int32)
            dist = np.zeros(max(total, 1), dtype=np.float32)
        lc = layout.c()
        check(self._L.ssb_organized_planes(self._h, msg.ctypes.data, C.byref(lc), boxes.ctypes.data, nb, C.byref(self.oopts),
                                           max_regions, reg.ctypes.data, nreg.ctypes.data, nin.ctypes.data,
                                           nrm.ctypes.data if want_points else None, lab.ctypes.data if want_points else None,
                                           dist.ctypes.data if want_points else None), "ssb_organized_planes")
        self.last_ms = float(self._L.ssb_organized_last_ms(self._h))
        if want_points:
            return reg, nreg, nin, nrm[:total], lab[:total], dist[:total]
        return reg, nreg, nin


PLANE_CLUSTER_DTYPE = np.dtype([("normal", "f4", 3), ("distance", "f4"), ("normal_label", "i4"), ("distance_label", "i4"),
                                ("n_points", "i4"), ("n_inliers", "i4"), ("coef", "f4", 4), ("row0", "i4"), ("n_rows", "i4")])


class PlaneClustering(PlaneSegmentation):
    """The reference's dormant plane-clustering chain on the device (include/ssb.h: ssb_kmeans, ssb_project_hull,
    ssb_cluster_planes; csrc/ssb_cluster.cuh): ``plane_segmentation::clusterAndSegmentAllPlanes``
    (plane_segmentation.cpp:261-294) = two cv::kmeans passes (normals, then signed distances, :296-429) and, per cluster,
    ``compute2DConvexHull`` (:631-664: RANSAC plane, ProjectInliers, ConvexHull) -> one row per hull vertex (:431-477)."""

    def __init__(self, device: int = -1, **kw):
        super().__init__(device=device)
        self.copts = _lib.ClusterOpts()
        self._L.ssb_cluster_default_opts(C.byref(self.copts))
        for k, v in kw.items():
            setattr(self.copts, k, v)

    def computeKmeans(self, points, num_centroids, rng_state=0xFFFFFFFF, attempts=10, max_count=10, epsilon=0.01):
        """plane_segmentation::computeKmeans (:525-535).  Returns (compactness, labels, centroids, rng state after)."""
        d = np.ascontiguousarray(points, dtype=np.float32)
        if d.ndim == 1:
            d = d[:, None]
        n, dims = d.shape
        lab = np.zeros(n, dtype=np.int32)
        cen = np.zeros((num_centroids, dims), dtype=np.float32)
        st = C.c_ulonglong(rng_state)
        comp = C.c_double(0.0)
        check(self._L.ssb_kmeans(self._h, d.ctypes.data, n, dims, num_centroids, max_count, epsilon, attempts, C.byref(st),
                                 lab.ctypes.data, cen.ctypes.data, C.byref(comp)), "ssb_kmeans")
        return comp.value, lab, cen, st.value

    def projectAndHull(self, pts4, mask, coef):
        """ProjectInliers + ConvexHull of compute2DConvexHull (:649-662).  Returns (hull vertices (k, 3) in PCL's output
        order, their indices in pts4, number of projected inliers)."""
        p = np.ascontiguousarray(pts4, dtype=np.float32).reshape(-1, 4)
        m = np.ascontiguousarray(mask, dtype=np.uint8)
        cf = np.ascontiguousarray(coef, dtype=np.float32)
        n = p.shape[0]
        rows = np.zeros((max(n, 1), 3), dtype=np.float32)
        src = np.zeros(max(n, 1), dtype=np.int32)
        nin = C.c_int(0)
        k = check(self._L.ssb_project_hull(self._h, p.ctypes.data, m.ctypes.data, n, cf.ctypes.data, rows.ctypes.data,
                                           src.ctypes.data, n, C.byref(nin)), "ssb_project_hull")
        return rows[:k].copy(), src[:k].copy(), nin.value

    def clusterAndSegmentAllPlanes(self, cloud4, normals4, transformation_mat, rng_state=0xFFFFFFFF, max_rows=65536,
                                   max_clusters=16):
        """Returns dict(rows (k, 8) = final_pose_vec, clusters (structured), labels (n,), centers (Kn, 3), rng_state)."""
        c = np.ascontiguousarray(cloud4, dtype=np.float32).reshape(-1, 4)
        q = np.ascontiguousarray(normals4, dtype=np.float32).reshape(-1, 4)
        n = c.shape[0]
        if q.shape[0] != n:
            raise ValueError("cloud and normals must have the same number of points")
        T = np.ascontiguousarray(transformation_mat, dtype=np.float32).reshape(16)
        rows = np.zeros((max_rows, 8), dtype=np.float32)
        cl = np.zeros(max_clusters, dtype=PLANE_CLUSTER_DTYPE)
        lab = np.zeros(max(n, 1), dtype=np.int32)
        cen = np.zeros((self.copts.num_centroids_normals, 3), dtype=np.float32)
        st = C.c_ulonglong(rng_state)
        nr, nc = C.c_int(0), C.c_int(0)
        check(self._L.ssb_cluster_planes(self._h, c.ctypes.data, q.ctypes.data, n, T.ctypes.data, C.byref(self.copts), C.byref(st),
                                         rows.ctypes.data, max_rows, C.byref(nr), cl.ctypes.data, max_clusters, C.byref(nc),
                                         lab.ctypes.data, cen.ctypes.data), "ssb_cluster_planes")
        return dict(rows=rows[:nr.value].copy(), clusters=cl[:nc.value].copy(), labels=lab[:n], centers=cen, rng_state=st.value)
