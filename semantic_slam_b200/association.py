"""Python mirror of ``data_association`` (/root/reference/include/ps_graph_slam/data_association.h:20-402) over the
C-ABI (``ssb_assoc_*`` in include/ssb.h).  Same method names and argument meaning; ``type`` / ``plane_type`` are
integer ids instead of std::string, the g2o node of a landmark is replaced by its refreshed estimate."""
from __future__ import annotations

import ctypes as C
import dataclasses
import numpy as np

from . import _lib
from ._lib import check


class AssocOpts(C.Structure):
    _fields_ = [("maha_dist_thres", C.c_double), ("eq_dist_thres", C.c_double), ("land_noise_low", C.c_double),
                ("land_noise_high", C.c_double), ("use_maha_dist", C.c_int), ("use_eq_dist", C.c_int),
                ("use_rtab_map_odom", C.c_int), ("strict", C.c_int)]


class DetectionC(C.Structure):
    _fields_ = [("type", C.c_int), ("plane_type", C.c_int), ("pose", C.c_float * 3), ("normal", C.c_float * 4)]


class LandmarkObsC(C.Structure):
    _fields_ = [("is_new_landmark", C.c_int), ("id", C.c_int), ("type", C.c_int), ("plane_type", C.c_int),
                ("pose", C.c_float * 3), ("local_pose", C.c_float * 3), ("normal", C.c_float * 4),
                ("covariance", C.c_float * 9), ("information", C.c_double * 9)]


class PlanarRegionC(C.Structure):
    _fields_ = [("centroid", C.c_float * 3), ("model", C.c_float * 4), ("contour_points", C.c_int), ("area", C.c_float)]


class DetectedObjectC(C.Structure):
    _fields_ = [("type", C.c_int), ("plane_type", C.c_int), ("prob", C.c_float), ("num_points", C.c_float),
                ("pose", C.c_float * 3), ("world_pose", C.c_float * 3), ("normal_orientation", C.c_float * 4)]


def segment_planar_surfaces(regions, robot_pose, cam_angle: float, object_type: int = 0, prob: float = 1.0,
                            planar_area: float = 0.0):
    """point_cloud_segmentation::segmentPlanarSurfaces over the post-processed regions of multiPlaneSegmentation
    (point_cloud_segmentation.h:26-103, plane_segmentation.cpp:160-255).  regions: sequence of
    (centroid3, model4, contour_points, area).  Returns detections in the form DataAssociation.find_matches takes,
    plus the world pose: [(type, plane_type, pose_cam float32[3], normal float32[4], world_pose float32[3])]."""
    L = _lib.lib()
    L.ssb_segment_planar_surfaces.argtypes = [C.POINTER(PlanarRegionC), C.c_int, C.POINTER(C.c_float), C.c_float, C.c_int,
                                              C.c_float, C.c_float, C.POINTER(DetectedObjectC)]
    n = len(regions)
    reg = (PlanarRegionC * max(n, 1))()
    for k, (cen, model, cpts, area) in enumerate(regions):
        for c in range(3):
            reg[k].centroid[c] = float(np.float32(cen[c]))
        for c in range(4):
            reg[k].model[c] = float(np.float32(model[c]))
        reg[k].contour_points = int(cpts)
        reg[k].area = float(np.float32(min(area, 3.0e38)))
    rp = np.ascontiguousarray(robot_pose, dtype=np.float32)
    out = (DetectedObjectC * max(n, 1))()
    m = check(L.ssb_segment_planar_surfaces(reg, n, rp.ctypes.data_as(C.POINTER(C.c_float)), C.c_float(float(np.float32(cam_angle))),
                                            int(object_type), C.c_float(prob), C.c_float(planar_area), out),
              "segment_planar_surfaces")
    return [(o.type, o.plane_type, np.array(o.pose[:], dtype=np.float32), np.array(o.normal_orientation[:], dtype=np.float32),
             np.array(o.world_pose[:], dtype=np.float32)) for o in out[:m]]


def planar_regions_from_ransac(results):
    """Regions for segment_planar_surfaces out of PlaneSegmentation.fit_planes results: inlier centroid and refined
    model of every fitted crop; the RANSAC path has no contour, so the contour gate sees the inlier count and the
    area gate is open."""
    return [(r["centroid"], r["refined"], int(r["refined_count"]), np.inf) for r in results if int(r["status"]) == 0]


@dataclasses.dataclass
class Landmark:
    """``landmark`` (include/ps_graph_slam/landmark.h:16-35) as returned for one detection"""
    is_new_landmark: bool
    id: int
    type: int
    plane_type: int
    pose: np.ndarray          # float32 (3,), world
    local_pose: np.ndarray    # float32 (3,), robot frame
    normal_orientation: np.ndarray   # float32 (4,)
    covariance: np.ndarray    # float32 (3,3)
    information: np.ndarray   # float64 (3,3) = covariance.inverse().cast<double>()


_bound = False


def _bind(L):
    global _bound
    if _bound:
        return
    vp = C.c_void_p
    L.ssb_assoc_default_opts.argtypes = [C.POINTER(AssocOpts)]
    L.ssb_assoc_create.argtypes = [C.POINTER(AssocOpts)]
    L.ssb_assoc_create.restype = vp
    L.ssb_assoc_destroy.argtypes = [vp]
    L.ssb_assoc_find_matches.argtypes = [vp, C.POINTER(DetectionC), C.c_int, C.POINTER(C.c_float), C.c_float,
                                         C.POINTER(LandmarkObsC)]
    L.ssb_assoc_set_landmark_estimate.argtypes = [vp, C.c_int, C.POINTER(C.c_double)]
    L.ssb_assoc_set_landmark_cov.argtypes = [vp, C.c_int, C.POINTER(C.c_float)]
    L.ssb_assoc_num_landmarks.argtypes = [vp]
    L.ssb_assoc_get_landmark.argtypes = [vp, C.c_int, C.POINTER(LandmarkObsC)]
    _bound = True


class DataAssociation:
    def __init__(self, verbose: bool = False, maha_dist_thres: float = 0.5, eq_dist_thres: float = 1.21,
                 land_noise_low: float = 0.5, land_noise_high: float = 0.9, use_maha_dist: bool = True,
                 use_eq_dist: bool = False, use_rtab_map_odom: bool = False, strict: bool = False):
        self._L = _lib.lib()
        _bind(self._L)
        o = AssocOpts(maha_dist_thres, eq_dist_thres, land_noise_low, land_noise_high, int(use_maha_dist),
                      int(use_eq_dist), int(use_rtab_map_odom), int(strict))
        self._h = C.c_void_p(self._L.ssb_assoc_create(C.byref(o)))
        self.verbose_ = verbose

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.ssb_assoc_destroy(self._h)
            self._h = None

    def find_matches(self, seg_obj_info, robot_pose, cam_angle: float):
        """seg_obj_info: sequence of (type, plane_type, pose3, normal4); robot_pose: x y z roll pitch yaw."""
        n = len(seg_obj_info)
        dets = (DetectionC * max(n, 1))()
        for k, (t, pt, pose, normal) in enumerate(seg_obj_info):
            dets[k].type, dets[k].plane_type = int(t), int(pt)
            for c in range(3):
                dets[k].pose[c] = float(np.float32(pose[c]))
            for c in range(4):
                dets[k].normal[c] = float(np.float32(normal[c]))
        rp = np.ascontiguousarray(robot_pose, dtype=np.float32)
        out = (LandmarkObsC * max(n, 1))()
        check(self._L.ssb_assoc_find_matches(self._h, dets, n, rp.ctypes.data_as(C.POINTER(C.c_float)),
                                             C.c_float(float(np.float32(cam_angle))), out), "find_matches")
        res = []
        for k in range(n):
            o = out[k]
            if o.id < 0:
                continue
            res.append(Landmark(bool(o.is_new_landmark), o.id, o.type, o.plane_type,
                                np.array(o.pose[:], dtype=np.float32), np.array(o.local_pose[:], dtype=np.float32),
                                np.array(o.normal[:], dtype=np.float32),
                                np.array(o.covariance[:], dtype=np.float32).reshape(3, 3),
                                np.array(o.information[:], dtype=np.float64).reshape(3, 3)))
        return res

    def setLandmarkEstimate(self, id: int, xyz):
        a = np.ascontiguousarray(xyz, dtype=np.float64)
        check(self._L.ssb_assoc_set_landmark_estimate(self._h, id, a.ctypes.data_as(C.POINTER(C.c_double))),
              "set_landmark_estimate")

    def setLandmarkCovs(self, id: int, cov):
        a = np.ascontiguousarray(cov, dtype=np.float32)
        check(self._L.ssb_assoc_set_landmark_cov(self._h, id, a.ctypes.data_as(C.POINTER(C.c_float))), "setLandmarkCovs")

    def num_landmarks(self) -> int:
        return self._L.ssb_assoc_num_landmarks(self._h)
