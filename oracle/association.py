"""CPU oracle of the per-frame landmark association step — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Independent restatement, in Python with numpy float32 scalars (every multiply / add rounded to single precision,
sums in Eigen's order), of
  data_association::find_matches / associate_lanmarks / map_a_new_lan / inserst_a_mapped_lan
      (/root/reference/include/ps_graph_slam/data_association.h:75-318) and
  semantic_tools::transformNormalsToWorld / transformPoseFromCameraToRobot / dist
      (/root/reference/include/tools.h:18-135,293-297).
cosf / sinf / sqrtf come from the same libm the C++ product links, through ctypes.  Quirks H4, H6, H7 of SURVEY.md
appendix A are reproduced.  Parity unpinned by the reference (it has no tests); pinned by tests/test_association.py
(hand-computed frames, invariances) and compared bit-for-bit with the product implementation."""
from __future__ import annotations

import ctypes
import ctypes.util
import dataclasses
import math

import numpy as np

f32 = np.float32
_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
for _n in ("cosf", "sinf", "sqrtf"):
    getattr(_libm, _n).restype = ctypes.c_float
    getattr(_libm, _n).argtypes = [ctypes.c_float]


def cosf(x):
    return f32(_libm.cosf(float(x)))


def sinf(x):
    return f32(_libm.sinf(float(x)))


def sqrtf(x):
    return f32(_libm.sqrtf(float(x)))


def mul4(A, B):
    R = np.zeros((4, 4), dtype=f32)
    for i in range(4):
        for j in range(4):
            acc = f32(A[i, 0] * B[0, j])
            for k in range(1, 4):
                acc = f32(acc + f32(A[i, k] * B[k, j]))
            R[i, j] = acc
    return R


def mulv4(A, v):
    out = np.zeros(4, dtype=f32)
    for i in range(4):
        acc = f32(A[i, 0] * v[0])
        for k in range(1, 4):
            acc = f32(acc + f32(A[i, k] * v[k]))
        out[i] = acc
    return out


def _fixed(cam_angle):
    a = f32(-f32(cam_angle))
    rxc = np.zeros((4, 4), dtype=f32)
    rxc[0, 0] = 1
    rxc[1, 1] = cosf(a)
    rxc[1, 2] = f32(-sinf(a))
    rxc[2, 1] = sinf(a)
    rxc[2, 2] = cosf(a)
    rxc[3, 3] = 1
    c, s = math.cos(-1.5708), math.sin(-1.5708)
    rxr = np.zeros((4, 4), dtype=f32)
    rxr[0, 0] = 1
    rxr[1, 1] = f32(c)
    rxr[1, 2] = f32(-s)
    rxr[2, 1] = f32(s)
    rxr[2, 2] = f32(c)
    rxr[3, 3] = 1
    rzr = np.zeros((4, 4), dtype=f32)
    rzr[0, 0] = f32(c)
    rzr[0, 1] = f32(-s)
    rzr[1, 0] = f32(s)
    rzr[1, 1] = f32(c)
    rzr[2, 2] = 1
    rzr[3, 3] = 1
    return rxc, rxr, rzr


def transform_normals_to_world(pose6, cam_angle):
    """tools.h:18-102"""
    rxc, rxr, rzr = _fixed(cam_angle)
    roll, pitch, yaw = f32(pose6[3]), f32(pose6[4]), f32(pose6[5])
    cy, sy, cp, sp, cr, sr = cosf(yaw), sinf(yaw), cosf(pitch), sinf(pitch), cosf(roll), sinf(roll)
    T = np.zeros((4, 4), dtype=f32)
    T[0, 0] = f32(cy * cp)
    T[0, 1] = f32(f32(f32(cy * sp) * sr) - f32(sy * cr))
    T[0, 2] = f32(f32(f32(cy * sp) * cr) + f32(sy * sp))   # H7
    T[1, 0] = f32(sy * cp)
    T[1, 1] = f32(f32(f32(sy * sp) * sr) + f32(cy * cr))
    T[1, 2] = f32(f32(f32(sy * sp) * cr) - f32(cy * sr))
    T[2, 0] = f32(-sp)
    T[2, 1] = f32(cp * sr)
    T[2, 2] = f32(cp * cr)
    T[3, 3] = 1
    return mul4(mul4(mul4(T, rzr), rxr), rxc)


def transform_cam_to_robot(cam_angle):
    """tools.h:104-135"""
    rxc, rxr, rzr = _fixed(cam_angle)
    return mul4(mul4(rzr, rxr), rxc)


def dist(x1, x2, y1, y2, z1, z2):
    """tools.h:293-297"""
    dx, dy, dz = f32(x2 - x1), f32(y2 - y1), f32(z2 - z1)
    return sqrtf(f32(f32(f32(dx * dx) + f32(dy * dy)) + f32(dz * dz)))


def inv3_lu(a):
    """Eigen::MatrixXf::inverse() of the dynamic-size Q of the Mahalanobis gate (data_association.h:175-184):
    partialPivLu().inverse() = solve(Identity), float.  PartialPivLU::unblocked_lu: first largest |entry| of the column is
    the pivot, whole rows swapped, sub-column divided by the pivot, trailing a(i,j) -= l(i) u(j) (two roundings).  Then
    X = P I, unit-lower solve and upper solve in Eigen's column-oriented form (TriangularSolverMatrix.h): x_i times the
    reciprocal of the diagonal, then x_r -= x_i t(r,i) for the rows still to come."""
    A = np.array(a, dtype=f32).reshape(3, 3).copy()
    X = np.eye(3, dtype=f32)
    piv = [0, 1, 2]
    for k in range(3):
        best, big = k, abs(A[k, k])
        for i in range(k + 1, 3):
            if abs(A[i, k]) > big:
                best, big = i, abs(A[i, k])
        piv[k] = best
        if big != 0:
            if best != k:
                A[[k, best]] = A[[best, k]]
            for i in range(k + 1, 3):
                A[i, k] = f32(A[i, k] / A[k, k])
        for i in range(k + 1, 3):
            for j in range(k + 1, 3):
                A[i, j] = f32(A[i, j] - f32(A[i, k] * A[k, j]))
    for k in range(3):
        if piv[k] != k:
            X[[k, piv[k]]] = X[[piv[k], k]]
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        for j in range(3):
            for i in range(3):
                b = X[i, j]
                for q in range(i + 1, 3):
                    X[q, j] = f32(X[q, j] - f32(b * A[q, i]))
            for i in (2, 1, 0):
                X[i, j] = f32(X[i, j] * f32(f32(1.0) / A[i, i]))
                b = X[i, j]
                for q in range(i):
                    X[q, j] = f32(X[q, j] - f32(b * A[q, i]))
    return X


def inv3(a):
    """Eigen::Matrix3f::inverse(): cofactors, determinant along column 0"""
    a = np.asarray(a, dtype=f32).reshape(3, 3)

    def cof(i, j):
        i1, i2, j1, j2 = (i + 1) % 3, (i + 2) % 3, (j + 1) % 3, (j + 2) % 3
        return f32(f32(a[i1, j1] * a[i2, j2]) - f32(a[i1, j2] * a[i2, j1]))
    c0, c1, c2 = cof(0, 0), cof(1, 0), cof(2, 0)
    det = f32(f32(f32(c0 * a[0, 0]) + f32(c1 * a[1, 0])) + f32(c2 * a[2, 0]))
    invdet = f32(f32(1.0) / det)
    r = np.zeros((3, 3), dtype=f32)
    r[0] = [f32(c0 * invdet), f32(c1 * invdet), f32(c2 * invdet)]
    for i in (1, 2):
        for j in range(3):
            r[i, j] = f32(cof(j, i) * invdet)
    return r


@dataclasses.dataclass
class OLandmark:
    is_new_landmark: bool
    id: int
    type: int
    plane_type: int
    pose: np.ndarray
    local_pose: np.ndarray
    normal_orientation: np.ndarray
    covariance: np.ndarray
    information: np.ndarray
    node_estimate: np.ndarray = None


class OracleDataAssociation:
    def __init__(self, maha_dist_thres=0.5, eq_dist_thres=1.21, land_noise_low=0.5, land_noise_high=0.9,
                 use_maha_dist=True, use_eq_dist=False, use_rtab_map_odom=False, strict=False):
        self.maha_dist_thres, self.eq_dist_thres = float(maha_dist_thres), float(eq_dist_thres)
        self.use_maha_dist, self.use_eq_dist = bool(use_maha_dist), bool(use_eq_dist)
        self.use_rtab_map_odom, self.strict = bool(use_rtab_map_odom), bool(strict)
        self.first_object = True
        self.landmarks = []
        self.Q = np.zeros((3, 3), dtype=f32)
        self.Q[0, 0] = self.Q[1, 1] = self.Q[2, 2] = f32(land_noise_low)

    # -- frame conversions (data_association.h:320-373)
    def _world(self, rp, cam_angle, v4):
        t = mulv4(transform_normals_to_world(rp, cam_angle), v4)
        t[0] = f32(t[0] + rp[0])
        if not self.use_rtab_map_odom:
            t[1] = f32(t[1] + rp[1])
        else:
            t[1] = f32(float(t[1]) + (float(rp[1]) - 0.04))
        t[2] = f32(t[2] + rp[2])
        return t

    def _obs(self, det, rp, cam_angle):
        t, pt, pose, normal = det
        cam = np.array([pose[0], pose[1], pose[2], 1.0], dtype=f32)
        w = self._world(rp, cam_angle, cam)
        n = mulv4(transform_normals_to_world(rp, cam_angle), np.asarray(normal, dtype=f32))
        r = mulv4(transform_cam_to_robot(cam_angle), cam)
        return OLandmark(False, -1, int(t), int(pt), w[:3].copy(), r[:3].copy(), n, self.Q.copy(),
                         inv3(self.Q).astype(np.float64))

    def _map_new(self, det, rp, cam_angle):
        l = self._obs(det, rp, cam_angle)
        l.is_new_landmark = True
        l.id = len(self.landmarks)
        stored = dataclasses.replace(l)
        stored.node_estimate = l.pose.astype(np.float64)
        self.landmarks.append(stored)
        return l

    def setLandmarkEstimate(self, id, xyz):
        self.landmarks[id].node_estimate = np.asarray(xyz, dtype=np.float64).copy()

    def setLandmarkCovs(self, id, cov):
        self.landmarks[id].covariance = np.asarray(cov, dtype=f32).reshape(3, 3).copy()

    def num_landmarks(self):
        return len(self.landmarks)

    def find_matches(self, seg_obj_info, robot_pose, cam_angle):
        rp = np.asarray(robot_pose, dtype=f32)
        cam_angle = f32(cam_angle)
        dets = [(t, pt, np.asarray(p, dtype=f32), np.asarray(n, dtype=f32)) for (t, pt, p, n) in seg_obj_info]
        if self.first_object:
            out = [self._map_new(d, rp, cam_angle) for d in dets]
            if out:
                self.first_object = False
            return out
        out = []
        found = False
        distance = f32(0)
        distance_min = np.finfo(f32).max
        nearest = 0
        for d in dets:
            if self.strict:
                distance_min, nearest = np.finfo(f32).max, 0
            cam = np.array([d[2][0], d[2][1], d[2][2], 1.0], dtype=f32)
            actual = self._world(rp, cam_angle, cam)
            for i, l in enumerate(list(self.landmarks)):
                if d[0] != l.type or d[1] != l.plane_type:
                    continue
                found = True
                expected = l.node_estimate.astype(f32)
                if self.use_maha_dist:
                    Qi = inv3_lu(np.array([[f32(l.covariance[r, c] + self.Q[r, c]) for c in range(3)] for r in range(3)], dtype=f32))
                    z = [f32(actual[k] - expected[k]) for k in range(3)]
                    t = [f32(f32(f32(z[0] * Qi[0, c]) + f32(z[1] * Qi[1, c])) + f32(z[2] * Qi[2, c])) for c in range(3)]
                    distance = f32(f32(f32(t[0] * z[0]) + f32(t[1] * z[1])) + f32(t[2] * z[2]))
                elif self.use_eq_dist:
                    distance = dist(actual[0], expected[0], actual[1], expected[1], actual[2], expected[2])
                if distance < distance_min:
                    distance_min, nearest = distance, i
            if not found:
                out.append(self._map_new(d, rp, cam_angle))
                continue
            found = False
            if self.use_maha_dist:
                is_new = float(distance_min) > self.maha_dist_thres
            elif self.use_eq_dist:
                is_new = float(distance_min) > self.eq_dist_thres
            else:
                continue
            if is_new:
                out.append(self._map_new(d, rp, cam_angle))
            else:
                l = self._obs(d, rp, cam_angle)
                l.id = self.landmarks[nearest].id
                out.append(l)
        return out


def segment_planar_surfaces(regions, robot_pose, cam_angle, object_type=0, prob=1.0, planar_area=0.0):
    """plane_segmentation.cpp:117-132,160-255 (gates, horizontal / vertical classification, normal signs) +
    point_cloud_segmentation.h:26-103 (camera -> world), float32 like the reference"""
    rp = np.asarray(robot_pose, dtype=f32)
    M = transform_normals_to_world(rp, f32(cam_angle))
    nh = [M[2, 0], M[2, 1], M[2, 2]]          # transformation_mat^T * (0,0,1,0)
    out = []
    for cen, model, cpts, area in regions:
        cen = np.asarray(cen, dtype=f32)
        model = np.asarray(model, dtype=f32)
        if not cpts > 100:
            continue
        dot = f32(0)
        for k in range(3):
            dot = f32(dot + f32(nh[k] * model[k]))
        if not f32(min(area, 3.0e38)) >= f32(planar_area):
            continue
        if all(float(f32(abs(model[k]) - abs(nh[k]))) < 0.3 for k in range(3)):
            flag, flip = 0, bool(model[1] > 0)
        elif float(dot) < 0.5:
            flag, flip = 1, bool(model[0] > 0)
        else:
            continue
        cam = np.array([cen[0], cen[1], cen[2], 1.0], dtype=f32)
        w = mulv4(M, cam)
        world = np.array([f32(w[k] + rp[k]) for k in range(3)], dtype=f32)
        normal = (-model if flip else model).astype(f32)
        out.append((int(object_type), flag, cam[:3].copy(), normal, world))
    return out
