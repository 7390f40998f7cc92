"""Per-frame loop of BASELINE.json configs[4] (associate + grow the graph + optimise every keyframe): the host logic
of semantic_graph_slam::run (semantic_graph_slam.cpp:57-205) driven once over the CUDA back-end and once over the CPU
oracle.  Parity bar of north_star: association indices bit-exact, optimised parameters within 1e-5 relative."""
import numpy as np
import pytest

import oracle
from parity import point_error, pose_errors
from oracle.association import OracleDataAssociation
from semantic_slam_b200 import synth
from semantic_slam_b200.semantic_graph_slam import SemanticGraphSLAM, matrix2vector

KITTI = dict(use_maha_dist=False, use_eq_dist=True, eq_dist_thres=1.5, land_noise_low=0.1)


def _drive(graph, assoc, stream, max_iterations=30, use_maha=False):
    slam = SemanticGraphSLAM(graph, assoc, stream.info6, cam_angle=stream.cam_angle, use_maha_dist=use_maha,
                             max_iterations=max_iterations)
    for k in range(stream.odom.shape[0]):
        slam.feed(stream.odom[k], stream.detections[k])
        assert slam.run()
    return slam


def test_loop_on_the_oracle_builds_a_consistent_map():
    stream = synth.make_frame_stream(40, 10)
    # strict association: with the reference's stale distance_min frames with several detections are
    # mis-associated by construction, which bends the map; parity (below) is checked in both modes
    slam = _drive(oracle.OracleGraphSLAM(), OracleDataAssociation(strict=True, **KITTI), stream)
    n_det = sum(len(d) for d in stream.detections)
    assert sum(len(f) for f in slam.association_log) == n_det
    n_new = sum(is_new for f in slam.association_log for (_, is_new) in f)
    assert 0 < n_new < n_det / 3                    # most detections re-observe a mapped landmark
    est = np.array([slam.graph_slam_.get_se3(kf["node"])[:, 3] for kf in slam.keyframes_])
    assert np.abs(est - stream.gt_pose[:, :, 3]).max() < 0.5   # drift stays bounded with the landmark constraints
    assert matrix2vector(stream.gt_pose[3]).dtype == np.float32


@pytest.mark.gpu
@pytest.mark.parametrize("use_maha,strict", [(False, False), (False, True), (True, True)])
def test_loop_parity_gpu_vs_oracle(use_maha, strict):
    from semantic_slam_b200 import GraphSLAM, DataAssociation
    stream = synth.make_frame_stream(60, 12)
    kw = dict(use_maha_dist=True, maha_dist_thres=30.0, land_noise_low=0.1) if use_maha else dict(KITTI)
    kw["strict"] = strict
    a = _drive(GraphSLAM(preconditioner=3, pcg_tol=1e-10), DataAssociation(**kw), stream, use_maha=use_maha)
    b = _drive(oracle.OracleGraphSLAM(), OracleDataAssociation(**kw), stream, use_maha=use_maha)
    assert a.association_log == b.association_log, "association indices must be bit-exact"
    assert len(a.landmark_nodes_) == len(b.landmark_nodes_) >= 8
    for ka, kb in zip(a.keyframes_, b.keyframes_):
        Ta, Tb = a.graph_slam_.get_se3(ka["node"]), b.graph_slam_.get_se3(kb["node"])
        rot, tr = pose_errors(Ta, Tb)
        assert rot <= 1e-5 and tr <= 1e-5, (rot, tr)
    for lid in a.landmark_nodes_:
        pa = a.graph_slam_.get_point_xyz(a.landmark_nodes_[lid])
        pb = b.graph_slam_.get_point_xyz(b.landmark_nodes_[lid])
        assert point_error(pa, pb) <= 1e-5
