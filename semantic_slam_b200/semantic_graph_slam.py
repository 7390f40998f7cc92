"""Host mirror of the per-tick SLAM step ``semantic_graph_slam::run``
(/root/reference/src/ps_graph_slam/semantic_graph_slam.cpp:57-205): keyframe queue -> SE3 nodes + odometry edges
(:104-152), detections -> data association (:207-236) -> landmark nodes + SE3-XYZ edges (:154-179), optimize (:81),
landmark covariances (:181-205), dead-reckoned robot pose (:94-95, 243-247).

Back-end agnostic: it drives any object pair with the GraphSLAM / data_association call surface, so the tests run
the identical host logic once over the CUDA back-end and once over the CPU oracle.  Segmentation is outside this
class (detections arrive as camera-frame centroids, e.g. from PlaneSegmentation.fit_planes)."""
from __future__ import annotations

import time

import numpy as np


def matrix2vector(T34):
    """ros_utils.hpp:90-106: (x, y, z, roll, pitch, yaw) of a pose, single precision.  The reference goes through
    tf::Matrix3x3::getEulerYPR; here the angles are taken from the rotation matrix directly (same convention)."""
    R = np.asarray(T34, dtype=np.float64)[:, :3]
    pitch = np.arctan2(-R[2, 0], np.hypot(R[0, 0], R[1, 0]))
    yaw = np.arctan2(R[1, 0], R[0, 0])
    roll = np.arctan2(R[2, 1], R[2, 2])
    t = np.asarray(T34)[:, 3]
    return np.array([t[0], t[1], t[2], roll, pitch, yaw], dtype=np.float32)


def _mul(A, B):
    T = np.zeros((3, 4))
    T[:, :3] = A[:, :3] @ B[:, :3]
    T[:, 3] = A[:, :3] @ B[:, 3] + A[:, 3]
    return T


def _inv(A):
    T = np.zeros((3, 4))
    T[:, :3] = A[:, :3].T
    T[:, 3] = -A[:, :3].T @ A[:, 3]
    return T


def _rotvec(R):
    c = np.clip((np.trace(R) - 1.0) / 2.0, -1.0, 1.0)
    th = np.arccos(c)
    if th < 1e-12:
        return np.zeros(3)
    return th / (2.0 * np.sin(th)) * np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])


def _rot(w):
    th = np.linalg.norm(w)
    if th < 1e-12:
        return np.eye(3)
    k = w / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


class SemanticGraphSLAM:
    max_keyframes_per_update = 10          # semantic_graph_slam.cpp:18

    def __init__(self, graph_slam, data_association, odom_information, cam_angle: float = 0.0,
                 use_maha_dist: bool = False, max_iterations: int = 1024, always_marginals: bool = False,
                 marginals_kwargs=None):
        # always_marginals: getAndSetLandmarkCov after EVERY optimise whatever the gate, as the reference does
        # (semantic_graph_slam.cpp:89,181-205); False = only when the Mahalanobis gate will read the covariances
        self.always_marginals_ = always_marginals
        self.marginals_kwargs_ = dict(marginals_kwargs or {})   # e.g. method="g2o" for the oracle back-end
        self.marginals_seconds = 0.0
        self.marginals_calls = 0
        self.graph_slam_ = graph_slam
        self.data_ass_obj_ = data_association
        self.information_ = np.asarray(odom_information, dtype=np.float64)
        self.cam_angle_ = float(cam_angle)
        self.use_maha_dist_ = use_maha_dist
        self.max_iterations = max_iterations
        self.keyframe_queue_ = []
        self.keyframes_ = []
        self.landmark_nodes_ = {}            # landmark id -> graph vertex id (data_association::assignLandmarkNode)
        self.robot_pose_ = np.eye(4)[:3].copy()
        self.prev_odom_ = None
        self.first_key_added_ = False
        self.association_log = []            # per keyframe: [(landmark id, is_new)] — the parity target

    # semantic_graph_slam::VIOCallback  :234-287, the branch taken when the keyframe gate REJECTS the pose (:239-247,
    # :253-261): robot_pose_ is dead-reckoned by the odometry increment, nothing is queued
    def add_odometry(self, odom34):
        odom34 = np.asarray(odom34, dtype=np.float64)
        if self.first_key_added_ and self.prev_odom_ is not None:
            self.robot_pose_ = _mul(self.robot_pose_, _mul(_inv(self.prev_odom_), odom34))
        self.prev_odom_ = odom34

    # ... and the branch taken when the pose becomes a keyframe (:264-284): the keyframe stores robot_pose_ AS IT IS —
    # the increment of this very pose is not applied — and only prev_odom_ advances.  With a stream in
    # which every pose is a keyframe the association therefore runs on the last optimised pose (:94), one step behind.
    def add_keyframe(self, odom34, detections):
        odom34 = np.asarray(odom34, dtype=np.float64)
        self.keyframe_queue_.append({"odom": odom34, "robot_pose": self.robot_pose_.copy(), "obj_info": list(detections),
                                     "node": None})
        self.prev_odom_ = odom34

    def feed(self, odom34, detections, substeps: int = 4):
        """A VIO stream running at `substeps` x the keyframe rate: the poses between the previous keyframe and this one
        (interpolated here) are rejected by the keyframe gate and only dead-reckon robot_pose_ (add_odometry); the last
        one becomes the keyframe.  This is the call pattern of the reference's node (VIOCallback per odometry message)."""
        odom34 = np.asarray(odom34, dtype=np.float64)
        if self.prev_odom_ is not None and substeps > 1:
            A = self.prev_odom_
            rel = _mul(_inv(A), odom34)
            w = _rotvec(rel[:, :3])
            for k in range(1, substeps):
                t = k / substeps
                step = np.zeros((3, 4))
                step[:, :3] = _rot(w * t)
                step[:, 3] = rel[:, 3] * t
                self.add_odometry(_mul(A, step))
        self.add_keyframe(odom34, detections)

    # semantic_graph_slam::empty_keyframe_queue  :104-152
    def _empty_keyframe_queue(self):
        if not self.keyframe_queue_:
            return []
        n = min(len(self.keyframe_queue_), self.max_keyframes_per_update)
        new = []
        for i in range(n):
            kf = self.keyframe_queue_[i]
            new.append(kf)
            kf["node"] = self.graph_slam_.add_se3_node(kf["odom"])
            if i == 0 and not self.keyframes_:
                continue
            prev = self.keyframes_[-1] if i == 0 else self.keyframe_queue_[i - 1]
            rel = _mul(_inv(prev["odom"]), kf["odom"])
            self.graph_slam_.add_se3_edge(prev["node"], kf["node"], rel, self.information_)
        del self.keyframe_queue_[:n]
        return new

    # semantic_graph_slam::run  :57-102
    def run(self) -> bool:
        new = self._empty_keyframe_queue()
        if not new:
            return False
        for kf in new:
            if not kf["obj_info"]:
                continue
            robot_pose = matrix2vector(kf["robot_pose"])                       # :211-212 (H10: dead-reckoned pose)
            lms = self.data_ass_obj_.find_matches(kf["obj_info"], robot_pose, self.cam_angle_)
            self.association_log.append([(l.id, bool(l.is_new_landmark)) for l in lms])
            for l in lms:                                                      # empty_landmark_queue :154-179
                if l.is_new_landmark:
                    self.landmark_nodes_[l.id] = self.graph_slam_.add_point_xyz_node(np.asarray(l.pose, dtype=np.float64))
                self.graph_slam_.add_se3_point_xyz_edge(kf["node"], self.landmark_nodes_[l.id],
                                                        np.asarray(l.local_pose, dtype=np.float64), l.information)
        self.keyframes_.extend(new)
        if self.graph_slam_.optimize(self.max_iterations):
            ids = sorted(self.landmark_nodes_)
            for lid in ids:   # the reference's association reads node->estimate() live (data_association.h:378)
                self.data_ass_obj_.setLandmarkEstimate(lid, self.graph_slam_.get_point_xyz(self.landmark_nodes_[lid]))
            if (self.use_maha_dist_ or self.always_marginals_) and ids:       # getAndSetLandmarkCov :181-205
                t0 = time.perf_counter()
                covs = self.graph_slam_.computeLandmarkMarginals([self.landmark_nodes_[lid] for lid in ids], **self.marginals_kwargs_)
                self.marginals_seconds += time.perf_counter() - t0
                self.marginals_calls += 1
                if covs is not None:
                    for lid, c in zip(ids, covs):
                        self.data_ass_obj_.setLandmarkCovs(lid, np.asarray(c, dtype=np.float32))
            self.robot_pose_ = np.asarray(self.graph_slam_.get_se3(self.keyframes_[-1]["node"]), dtype=np.float64)  # :94
        self.first_key_added_ = True
        return True
