"""Seeded synthetic workloads for the two hot paths.

Graphs reproduce what ``semantic_graph_slam::run`` would build
(/root/reference/src/ps_graph_slam/semantic_graph_slam.cpp:104-179): one SE3 vertex per
keyframe in arrival order, an odometry edge to the previous keyframe whose measurement is the
relative pose of the two *raw odometry* poses (so odometry residuals are exactly zero at the
start, :134-135), landmarks created at first observation with the position projected through the
keyframe's initial pose (data_association.h:247-264), and one SE3-PointXYZ edge per observation
carrying the landmark position in the robot frame (:255-262).  Information matrices follow
information_matrix_calculator.cpp:28-35 with config/yolo_detector_kitti.yaml:19,23-24:
Omega_odom = diag(1/0.00667 x3, 1/0.00001 x3), Omega_lm = (1/0.1) I3.

Clouds reproduce a 640x480 organised ``sensor_msgs/PointCloud2`` (point_step 32: x,y,z @0/4/8,
rgb @16) as consumed by plane_segmentation.cpp:24-82, plus yolo-style bounding boxes
(msg/ObjectInfo.msg).
"""
from __future__ import annotations

import dataclasses
import numpy as np

SEED_BASE = 20260925


# --------------------------------------------------------------------------------------------
# SE3 helpers (numpy, float64) — used only to generate data, not by either back-end.
# --------------------------------------------------------------------------------------------
def rpy_to_R(roll, pitch, yaw):
    cr, sr = np.cos(roll), np.sin(roll)
    cp, sp = np.cos(pitch), np.sin(pitch)
    cy, sy = np.cos(yaw), np.sin(yaw)
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1.0]])
    Ry = np.array([[cp, 0, sp], [0, 1.0, 0], [-sp, 0, cp]])
    Rx = np.array([[1.0, 0, 0], [0, cr, -sr], [0, sr, cr]])
    return Rz @ Ry @ Rx


def rotvec_to_R(w):
    th = np.linalg.norm(w)
    if th < 1e-12:
        return np.eye(3)
    k = w / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def T_make(R, t):
    T = np.zeros((3, 4))
    T[:, :3] = R
    T[:, 3] = t
    return T


def T_mul(A, B):
    return T_make(A[:, :3] @ B[:, :3], A[:, :3] @ B[:, 3] + A[:, 3])


def T_inv(A):
    Rt = A[:, :3].T
    return T_make(Rt, -Rt @ A[:, 3])


# --------------------------------------------------------------------------------------------
@dataclasses.dataclass
class GraphSpec:
    """A graph as the ordered list of API calls the reference would make."""
    # vertices in creation order
    vkind: np.ndarray          # (Nv,) 0 = SE3, 1 = PointXYZ
    vpose: np.ndarray          # (Nv,3,4) initial estimate for SE3 vertices (rows of zeros otherwise)
    vxyz: np.ndarray           # (Nv,3) initial estimate for XYZ vertices
    # edges in creation order
    ekind: np.ndarray          # (Ne,) 0 = SE3-SE3, 1 = SE3-XYZ
    evi: np.ndarray            # (Ne,)
    evj: np.ndarray            # (Ne,)
    eZ: np.ndarray             # (Ne,3,4) SE3 measurement (kind 0)
    ez: np.ndarray             # (Ne,3) xyz measurement (kind 1)
    einfo6: np.ndarray         # (6,6) shared odometry information
    einfo3: np.ndarray         # (3,3) shared landmark information
    # ground truth (for diagnostics only)
    gt_pose: np.ndarray
    gt_xyz: np.ndarray
    name: str = ""

    @property
    def n_poses(self):
        return int((self.vkind == 0).sum())

    @property
    def n_landmarks(self):
        return int((self.vkind == 1).sum())

    @property
    def n_edges(self):
        return int(self.ekind.size)


def lawnmower_path(n_kf: int, step: float = 0.5, lane_kf: int | None = None, lane_gap: float = 3.0):
    """Boustrophedon ("lawn-mower") polyline sampled every `step` metres (= keyframe_delta_trans,
    keyframe_updater.hpp:23).  Returns xy positions (n,2) and heading yaw (n,)."""
    if lane_kf is None:
        lane_kf = max(10, int(round(np.sqrt(n_kf * lane_gap / step) )))  # roughly square footprint
    L = lane_kf * step
    gap_kf = max(1, int(round(lane_gap / step)))
    period = lane_kf + gap_kf
    xy = np.zeros((n_kf, 2))
    yaw = np.zeros(n_kf)
    for k in range(n_kf):
        lane = k // period
        r = k % period
        fwd = (lane % 2 == 0)
        y0 = lane * gap_kf * step
        if r < lane_kf:
            x = r * step if fwd else L - r * step
            xy[k] = (x, y0)
            yaw[k] = 0.0 if fwd else np.pi
        else:
            x = L if fwd else 0.0
            xy[k] = (x, y0 + (r - lane_kf) * step)
            yaw[k] = np.pi / 2
    return xy, yaw


def make_graph(n_kf: int, n_lm: int, obs_per_kf: int = 5, seed: int = SEED_BASE, obs_radius: float = 6.0,
               odom_sigma_t: float = 0.01, odom_sigma_r: float = 0.002, meas_sigma: float = 0.05,
               name: str = "") -> GraphSpec:
    from scipy.spatial import cKDTree

    rng = np.random.default_rng(seed)
    xy, yaw = lawnmower_path(n_kf)
    s = np.arange(n_kf) * 0.5
    z = 0.2 * np.sin(2 * np.pi * s / 40.0)
    roll = rng.normal(0, 0.01, n_kf)
    pitch = rng.normal(0, 0.01, n_kf)
    gt = np.zeros((n_kf, 3, 4))
    for k in range(n_kf):
        gt[k] = T_make(rpy_to_R(roll[k], pitch[k], yaw[k]), np.array([xy[k, 0], xy[k, 1], z[k]]))
    lo = xy.min(0) - 3.0
    hi = xy.max(0) + 3.0
    lm_gt = np.column_stack([rng.uniform(lo[0], hi[0], n_lm), rng.uniform(lo[1], hi[1], n_lm),
                             rng.uniform(0.0, 2.0, n_lm)])
    tree = cKDTree(lm_gt[:, :2])
    dists, idxs = tree.query(xy, k=min(obs_per_kf, n_lm), distance_upper_bound=obs_radius)
    if dists.ndim == 1:
        dists, idxs = dists[:, None], idxs[:, None]

    # noisy odometry chain -> initial pose estimates
    odom = np.zeros((n_kf, 3, 4))
    odom[0] = gt[0]
    Zs = np.zeros((max(n_kf - 1, 0), 3, 4))
    nt = rng.normal(0, odom_sigma_t, (n_kf, 3))
    nr = rng.normal(0, odom_sigma_r, (n_kf, 3))
    for k in range(1, n_kf):
        rel = T_mul(T_inv(gt[k - 1]), gt[k])
        noisy = T_mul(rel, T_make(rotvec_to_R(nr[k]), nt[k]))
        odom[k] = T_mul(odom[k - 1], noisy)
    for k in range(1, n_kf):
        Zs[k - 1] = T_mul(T_inv(odom[k - 1]), odom[k])   # semantic_graph_slam.cpp:134-135

    vkind, vpose, vxyz = [], [], []
    ekind, evi, evj, eZ, ez = [], [], [], [], []
    lm_vid = {}
    pose_vid = np.zeros(n_kf, dtype=np.int64)
    zero34 = np.zeros((3, 4))
    zero3 = np.zeros(3)
    mn = rng.normal(0, meas_sigma, (n_kf, dists.shape[1], 3))
    for k in range(n_kf):
        pose_vid[k] = len(vkind)
        vkind.append(0); vpose.append(odom[k]); vxyz.append(zero3)
        if k > 0:
            ekind.append(0); evi.append(pose_vid[k - 1]); evj.append(pose_vid[k]); eZ.append(Zs[k - 1]); ez.append(zero3)
        Tinv_gt = T_inv(gt[k])
        for j in range(dists.shape[1]):
            if not np.isfinite(dists[k, j]):
                continue
            l = int(idxs[k, j])
            meas = Tinv_gt[:, :3] @ lm_gt[l] + Tinv_gt[:, 3] + mn[k, j]
            # values originate as float32 in the reference (landmark.local_pose is Vector3f)
            meas = meas.astype(np.float32).astype(np.float64)
            if l not in lm_vid:
                lm_vid[l] = len(vkind)
                world = (odom[k][:, :3] @ meas + odom[k][:, 3]).astype(np.float32).astype(np.float64)
                vkind.append(1); vpose.append(zero34); vxyz.append(world)
            ekind.append(1); evi.append(pose_vid[k]); evj.append(lm_vid[l]); eZ.append(zero34); ez.append(meas)

    info6 = np.diag([1 / 0.00667] * 3 + [1 / 0.00001] * 3)
    info3 = np.eye(3) * (1 / 0.1)
    return GraphSpec(np.array(vkind, dtype=np.int32), np.array(vpose), np.array(vxyz),
                     np.array(ekind, dtype=np.int32), np.array(evi, dtype=np.int32), np.array(evj, dtype=np.int32),
                     np.array(eZ), np.array(ez), info6, info3, gt, lm_gt, name or f"kf{n_kf}_lm{n_lm}")


CONFIGS = {
    # name: (n_kf, n_lm, seed offset)   BASELINE.json configs[0], [1], [3], [4]
    "cfg1": (100, 20, 1),
    "cfg2": (10_000, 2_000, 2),
    "cfg4": (100_000, 20_000, 4),
    "cfg5": (4_000, 400, 5),
}


def make_config_graph(name: str, scale: int = 1) -> GraphSpec:
    n_kf, n_lm, off = CONFIGS[name]
    return make_graph(n_kf * scale, n_lm * scale, seed=SEED_BASE + off, name=name if scale == 1 else f"{name}x{scale}")


def load_graph(backend, spec: GraphSpec):
    """Replay `spec` through a GraphSLAM-like object (reference call surface:
    add_se3_node / add_point_xyz_node / add_se3_edge / add_se3_point_xyz_edge).  Returns vertex ids."""
    ids = np.zeros(spec.vkind.size, dtype=np.int64)
    for v in range(spec.vkind.size):
        if spec.vkind[v] == 0:
            ids[v] = backend.add_se3_node(spec.vpose[v])
        else:
            ids[v] = backend.add_point_xyz_node(spec.vxyz[v])
    for e in range(spec.ekind.size):
        if spec.ekind[e] == 0:
            backend.add_se3_edge(int(ids[spec.evi[e]]), int(ids[spec.evj[e]]), spec.eZ[e], spec.einfo6)
        else:
            backend.add_se3_point_xyz_edge(int(ids[spec.evi[e]]), int(ids[spec.evj[e]]), spec.ez[e], spec.einfo3)
    return ids


# --------------------------------------------------------------------------------------------
# Depth clouds + bounding boxes for the RANSAC path (cfg3)
# --------------------------------------------------------------------------------------------
@dataclasses.dataclass
class CloudSpec:
    msg: np.ndarray        # (480*640*32,) uint8  PointCloud2.data, point_step 32, row_step 20480
    width: int
    height: int
    point_step: int
    row_step: int
    offsets: tuple         # byte offsets of x, y, z, rgb
    boxes: np.ndarray      # (nb,4) int32: tl_x, tl_y, width, height   (msg/ObjectInfo.msg)
    triples: np.ndarray    # (nb,K,3) int32 sample indices into each crop (row-major crop index)
    name: str = "cfg3"


def make_cloud(n_boxes: int = 64, n_hyp: int = 1024, seed: int = SEED_BASE + 3, width: int = 640, height: int = 480,
               n_planes: int = 6, nan_frac: float = 0.05, box_min: int = 40, box_max: int = 200) -> CloudSpec:
    rng = np.random.default_rng(seed)
    fx = fy = 525.0
    cx, cy = (width - 1) / 2.0, (height - 1) / 2.0
    u, v = np.meshgrid(np.arange(width), np.arange(height))
    dx = (u - cx) / fx
    dy = (v - cy) / fy
    # scene: image split into vertical strips x horizontal bands, each one a random plane
    depth = np.full((height, width), np.inf)
    region = (u * 3 // width) + 3 * (v * 2 // height)
    for r in range(n_planes):
        n = rng.normal(0, 1, 3)
        n[2] = -abs(n[2]) - 1.0           # facing the camera
        n /= np.linalg.norm(n)
        d0 = rng.uniform(0.5, 6.0)
        # plane: n . p = n_z * d0 at the optical axis -> depth = n_z d0 / (n . ray)
        denom = n[0] * dx + n[1] * dy + n[2]
        zz = n[2] * d0 / denom
        m = region == (r % 6)
        depth[m] = zz[m]
    depth = np.clip(depth, 0.3, 12.0)
    depth = depth + rng.normal(0, 1, depth.shape) * 0.002 * depth * depth
    X = (dx * depth).astype(np.float32)
    Y = (dy * depth).astype(np.float32)
    Z = depth.astype(np.float32)
    drop = rng.random(depth.shape) < nan_frac
    X[drop] = np.nan; Y[drop] = np.nan; Z[drop] = np.nan
    point_step, row_step = 32, 32 * width
    msg = np.zeros((height, width, 8), dtype=np.float32)
    msg[..., 0] = X; msg[..., 1] = Y; msg[..., 2] = Z
    msg[..., 4] = rng.random(depth.shape).astype(np.float32)   # packed rgb as float (opaque payload)
    boxes = np.zeros((n_boxes, 4), dtype=np.int32)
    for b in range(n_boxes):
        w = int(rng.integers(box_min, box_max + 1))
        h = int(rng.integers(box_min, box_max + 1))
        boxes[b] = (int(rng.integers(0, width - w + 1)), int(rng.integers(0, height - h + 1)), w, h)
    triples = np.zeros((n_boxes, n_hyp, 3), dtype=np.int32)
    for b in range(n_boxes):
        r2 = np.random.default_rng(12345 + b)        # PCL seeds its RNG with 12345
        n = int(boxes[b, 2]) * int(boxes[b, 3])
        t = r2.integers(0, n, (n_hyp, 3))
        triples[b] = t
    return CloudSpec(msg.reshape(-1).view(np.uint8).copy(), width, height, point_step, row_step, (0, 4, 8, 16),
                     boxes, triples)


# --------------------------------------------------------------------------------------------
# Pose - plane graphs (the reference's dormant VertexPlane / EdgeSE3Plane path)
# --------------------------------------------------------------------------------------------
def plane_transform(T, c):
    """g2o: Plane3D operator*(Isometry3D, Plane3D) — moves plane coefficients (n, c3) by T."""
    n = T[:, :3] @ np.asarray(c[:3])
    return np.array([n[0], n[1], n[2], c[3] - T[:, 3] @ n])


def plane_perturb(c, v):
    """measurement noise on a plane: rotate the unit normal by small azimuth / elevation angles about its own chart
    (g2o Plane3D::oplus convention) and shift the distance; data generation only"""
    c = np.asarray(c, dtype=np.float64)
    n = c[:3] / np.linalg.norm(c[:3])
    az, el = np.arctan2(n[1], n[0]), np.arctan2(n[2], np.hypot(n[0], n[1]))
    Rz = np.array([[np.cos(az), -np.sin(az), 0], [np.sin(az), np.cos(az), 0], [0, 0, 1.0]])
    Ry = np.array([[np.cos(el), 0, -np.sin(el)], [0, 1.0, 0], [np.sin(el), 0, np.cos(el)]])
    d = np.array([np.cos(v[1]) * np.cos(v[0]), np.cos(v[1]) * np.sin(v[0]), np.sin(v[1])])
    n2 = Rz @ Ry @ d
    return np.array([n2[0], n2[1], n2[2], c[3] / np.linalg.norm(c[:3]) - v[2]])


@dataclasses.dataclass
class PlaneGraphSpec:
    vertices: list   # ("se3", T34) | ("xyz", p3) | ("plane", c4) in creation order
    edges: list      # ("se3", vi, vj, Z34) | ("xyz", vi, vj, z3) | ("plane", vi, vj, c4)
    info6: np.ndarray
    info3: np.ndarray
    info_plane: np.ndarray
    gt_pose: np.ndarray
    gt_planes: np.ndarray

    @property
    def n_poses(self):
        return sum(1 for v in self.vertices if v[0] == "se3")


def make_plane_graph(n_kf: int = 40, n_planes: int = 6, n_lm: int = 8, seed: int = SEED_BASE + 14,
                     plane_sigma=(0.01, 0.01, 0.02)) -> PlaneGraphSpec:
    """Keyframes on a lawn-mower path observing wall-like planes (and a few point landmarks): every keyframe
    sees 3 planes; the measurement is the plane in the robot frame, perturbed through Plane3D::oplus."""
    rng = np.random.default_rng(seed)
    base = make_graph(n_kf, max(n_lm, 1), obs_per_kf=2, seed=seed, name="plane_base")
    gt = base.gt_pose
    # planes: normals away from the z axis (azimuth/elevation chart is singular there), a few metres out
    planes = []
    for _ in range(n_planes):
        az = rng.uniform(-np.pi, np.pi)
        el = rng.uniform(-0.6, 0.6)
        n = np.array([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)])
        planes.append(np.array([n[0], n[1], n[2], -rng.uniform(2.0, 9.0)]))
    planes = np.array(planes)
    vertices, edges = [], []
    vmap = {}
    for v in range(base.vkind.size):
        vmap[v] = len(vertices)
        vertices.append(("se3", base.vpose[v]) if base.vkind[v] == 0 else ("xyz", base.vxyz[v]))
    for e in range(base.ekind.size):
        if base.ekind[e] == 0:
            edges.append(("se3", vmap[int(base.evi[e])], vmap[int(base.evj[e])], base.eZ[e]))
        else:
            edges.append(("xyz", vmap[int(base.evi[e])], vmap[int(base.evj[e])], base.ez[e]))
    pose_vids = [vmap[v] for v in range(base.vkind.size) if base.vkind[v] == 0]
    plane_vid = {}
    for k in range(n_kf):
        seen = rng.choice(n_planes, size=min(3, n_planes), replace=False)
        for j in seen:
            local = plane_transform(T_inv(gt[k]), planes[j])
            meas = plane_perturb(local, rng.normal(0, 1, 3) * np.array(plane_sigma))
            if j not in plane_vid:
                plane_vid[j] = len(vertices)
                vertices.append(("plane", plane_transform(base.vpose[np.flatnonzero(base.vkind == 0)[k]], meas)))
            edges.append(("plane", pose_vids[k], plane_vid[j], meas))
    info_plane = np.diag([1 / 0.01, 1 / 0.01, 1 / 0.02])
    return PlaneGraphSpec(vertices, edges, base.einfo6, base.einfo3, info_plane, gt, planes)


def load_plane_graph(backend, spec: PlaneGraphSpec):
    """Replay a PlaneGraphSpec through a GraphSLAM-like object; returns the vertex ids."""
    ids = []
    for v in spec.vertices:
        if v[0] == "se3":
            ids.append(backend.add_se3_node(v[1]))
        elif v[0] == "xyz":
            ids.append(backend.add_point_xyz_node(v[1]))
        else:
            ids.append(backend.add_plane_node(v[1]))
    for e in spec.edges:
        if e[0] == "se3":
            backend.add_se3_edge(ids[e[1]], ids[e[2]], e[3], spec.info6)
        elif e[0] == "xyz":
            backend.add_se3_point_xyz_edge(ids[e[1]], ids[e[2]], e[3], spec.info3)
        else:
            backend.add_se3_plane_edge(ids[e[1]], ids[e[2]], e[3], spec.info_plane)
    return ids


# --------------------------------------------------------------------------------------------
# Per-frame stream (BASELINE.json configs[4]: segment + associate + optimise per frame)
# --------------------------------------------------------------------------------------------
@dataclasses.dataclass
class FrameStream:
    odom: np.ndarray        # (n_kf,3,4) raw VIO odometry poses (chained noisy increments)
    detections: list        # per keyframe: [(type, plane_type, pose_cam float32[3], normal float32[4])]
    gt_pose: np.ndarray
    gt_landmarks: np.ndarray
    cam_angle: float
    info6: np.ndarray


def _world_from_cam_matrix(pose6, cam_angle):
    """double-precision version of semantic_tools::transformNormalsToWorld (tools.h:18-102, including the sy*sp term
    of :80-81) used only to synthesise camera-frame detections that land where the ground truth is."""
    roll, pitch, yaw = pose6[3], pose6[4], pose6[5]
    cy, sy, cp, sp, cr, sr = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch), np.cos(roll), np.sin(roll)
    T = np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sp],
                  [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                  [-sp, cp * sr, cp * cr]])
    a = -1.5708
    Rz = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]])
    Rx = np.array([[1.0, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    c = -cam_angle
    Rc = np.array([[1.0, 0, 0], [0, np.cos(c), -np.sin(c)], [0, np.sin(c), np.cos(c)]])
    return T @ Rz @ Rx @ Rc


def make_frame_stream(n_kf: int = 60, n_lm: int = 12, seed: int = SEED_BASE + 5, view_radius: float = 5.0,
                      max_det: int = 3, meas_sigma: float = 0.03, cam_angle: float = 0.1) -> FrameStream:
    """Keyframes along a lawn-mower path; every keyframe detects up to `max_det` of the landmarks within
    `view_radius` (so a landmark is seen from a run of consecutive keyframes), reported as camera-frame centroids."""
    from .semantic_graph_slam import matrix2vector
    rng = np.random.default_rng(seed)
    base = make_graph(n_kf, n_lm, obs_per_kf=1, seed=seed, name="frames")
    gt = base.gt_pose
    odom = base.vpose[base.vkind == 0]
    lm = base.gt_xyz
    dets = []
    for k in range(n_kf):
        p6 = matrix2vector(gt[k]).astype(np.float64)
        M = _world_from_cam_matrix(p6, cam_angle)
        d = np.linalg.norm(lm[:, :2] - gt[k][:2, 3], axis=1)
        near = np.argsort(d)[:max_det]
        near = [int(j) for j in near if d[j] < view_radius]
        frame = []
        for j in near:
            cam = np.linalg.solve(M, lm[j] - gt[k][:, 3]) + rng.normal(0, meas_sigma, 3)
            ptype = int(j % 2)     # "horizontal" / "vertical"
            frame.append((0, ptype, cam.astype(np.float32), np.array([0, 0, 1, 0], dtype=np.float32)))
        dets.append(frame)
    return FrameStream(odom, dets, gt, lm, cam_angle, base.einfo6)


# --------------------------------------------------------------------------------------------
# Organised crop for the dormant plane-clustering chain (row f4): two parallel "horizontal" planes at
# different distances (same normal cluster, two distance clusters) and a wall behind them
# --------------------------------------------------------------------------------------------
def make_cluster_scene(h: int = 200, w: int = 240, seed: int = 5, noise: float = 0.0015, nan_frac: float = 0.001,
                       d_a: float = 1.0, d_b: float = 1.8, wall_normal=(0.55, 0.45, -0.7), d_wall: float = 2.2):
    """Returns (cloud (h, w, 4) float32: x y z rgb, transformation_mat (4, 4) float32 whose third row is the normal of the
    horizontal planes in the camera frame — what plane_segmentation.cpp:332-346 derives from the camera pose)."""
    rng = np.random.default_rng(SEED_BASE + 40 + seed)
    fx = 525.0
    u, v = np.meshgrid(np.arange(w), np.arange(h))
    dx, dy = (u - (w - 1) / 2) / fx, (v - (h - 1) / 2) / fx
    nA = np.array([0.1, -0.5, -0.85])
    nA /= np.linalg.norm(nA)
    nC = np.asarray(wall_normal, dtype=np.float64)
    nC = nC / np.linalg.norm(nC)

    def depth(n, d0):
        return n[2] * d0 / (n[0] * dx + n[1] * dy + n[2])
    z = np.where(u < w // 2, depth(nA, d_a), depth(nA, d_b))
    z = np.where(v < h // 4, depth(nC, d_wall), z)
    z = z + rng.normal(0.0, noise, z.shape) * z * z
    c = np.zeros((h, w, 4), dtype=np.float32)
    c[..., 0], c[..., 1], c[..., 2] = dx * z, dy * z, z
    drop = rng.random((h, w)) < nan_frac
    c[drop, :3] = np.nan
    # rotation whose third row is nA
    a = np.cross(nA, [1.0, 0.0, 0.0])
    a /= np.linalg.norm(a)
    b = np.cross(nA, a)
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = np.stack([a, b, nA]).astype(np.float32)
    T[:3, 3] = [0.3, -0.2, 1.1]
    return c, T
