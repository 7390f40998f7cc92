"""Landmark association: the product's host implementation (C-ABI ssb_assoc_*) against the
independent float32 oracle restatement, bit for bit, plus hand-checkable frames.  No GPU needed: the step is host
code in the reference as well (data_association.h:75-389, tools.h:18-135)."""
import numpy as np
import pytest

from oracle.association import OracleDataAssociation, transform_normals_to_world, transform_cam_to_robot, dist
from semantic_slam_b200 import DataAssociation

KITTI = dict(use_maha_dist=False, use_eq_dist=True, eq_dist_thres=1.5, land_noise_low=0.1)   # config/yolo_detector_kitti.yaml:17-20


def _frames(seed, n_frames=25, n_det=4):
    rng = np.random.default_rng(seed)
    anchors = rng.uniform(-4, 4, (6, 3)) + np.array([0, 0, 6.0])
    for f in range(n_frames):
        rp = np.array([0.3 * f, 0.05 * f, 0.02 * f, rng.normal(0, 0.02), rng.normal(0, 0.02), 0.05 * f], dtype=np.float32)
        dets = []
        for _ in range(n_det):
            a = anchors[rng.integers(0, len(anchors))] + rng.normal(0, 0.15, 3)
            dets.append((int(rng.integers(0, 2)), int(rng.integers(0, 2)), a.astype(np.float32),
                         rng.normal(0, 1, 4).astype(np.float32)))
        yield rp, dets


def _same(A, O):
    assert [(x.id, x.is_new_landmark) for x in A] == [(y.id, y.is_new_landmark) for y in O]
    for x, y in zip(A, O):
        assert np.array_equal(x.pose, y.pose) and np.array_equal(x.local_pose, y.local_pose)
        assert np.array_equal(x.normal_orientation, y.normal_orientation)
        assert np.array_equal(x.covariance, y.covariance) and np.array_equal(x.information, y.information)


@pytest.mark.parametrize("strict", [False, True])
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_euclidean_gate_bit_exact(seed, strict):
    a, o = DataAssociation(strict=strict, **KITTI), OracleDataAssociation(strict=strict, **KITTI)
    matched = 0
    for rp, dets in _frames(seed):
        A, O = a.find_matches(dets, rp, 0.17), o.find_matches(dets, rp, 0.17)
        _same(A, O)
        matched += sum(not x.is_new_landmark for x in A)
        for lid in range(a.num_landmarks()):   # as after an optimize: refreshed node estimates
            est = o.landmarks[lid].node_estimate + 0.01
            a.setLandmarkEstimate(lid, est)
            o.setLandmarkEstimate(lid, est)
    assert a.num_landmarks() == o.num_landmarks() > 4 and matched > 10


def test_mahalanobis_gate_bit_exact():
    kw = dict(use_maha_dist=True, maha_dist_thres=3.0, land_noise_low=0.4)
    a, o = DataAssociation(**kw), OracleDataAssociation(**kw)
    rng = np.random.default_rng(7)
    for rp, dets in _frames(11):
        _same(a.find_matches(dets, rp, 0.0), o.find_matches(dets, rp, 0.0))
        for lid in range(a.num_landmarks()):
            B = rng.normal(0, 0.1, (3, 3))
            cov = (B @ B.T + 0.05 * np.eye(3)).astype(np.float32)
            a.setLandmarkCovs(lid, cov)
            o.setLandmarkCovs(lid, cov)


def test_axis_convention_of_the_camera_chain():
    """cam (x right, y down, z forward) -> robot (x forward, y left, z up): rot_z(-90) rot_x(-90), tools.h:104-135"""
    M = transform_cam_to_robot(0.0).astype(np.float64)
    v = M @ np.array([1.0, 2.0, 3.0, 1.0])
    assert np.allclose(v[:3], [3.0, -1.0, -2.0], atol=1e-4)
    a = DataAssociation(**KITTI)
    rp = np.array([10.0, -2.0, 0.5, 0, 0, np.pi / 2], dtype=np.float32)   # robot facing +y
    l = a.find_matches([(0, 0, np.array([1.0, 2.0, 3.0], dtype=np.float32), np.array([0, 0, 1, 0], dtype=np.float32))], rp, 0.0)[0]
    assert l.is_new_landmark and l.id == 0
    assert np.allclose(l.local_pose, [3.0, -1.0, -2.0], atol=1e-4)
    assert np.allclose(l.pose, [10.0 + 1.0, -2.0 + 3.0, 0.5 - 2.0], atol=1e-3)   # forward 3 m -> +y, left 1 m -> -x ... +x: -(-1)
    assert np.allclose(np.diag(l.information), 1.0 / np.float32(0.1), rtol=1e-6)


def test_quirk_h7_term_is_reproduced():
    """T_robot_world(0,2) = cy*sp*cr + sy*sp (tools.h:80-81), not the textbook sy*sr"""
    p = np.array([0, 0, 0, 0.3, 0.2, 0.4], dtype=np.float32)
    M = transform_normals_to_world(p, 0.0).astype(np.float64)
    cy, sy, sp, cr, sr = np.cos(0.4), np.sin(0.4), np.sin(0.2), np.cos(0.3), np.sin(0.3)
    rz = np.array([[0, 1, 0], [-1, 0, 0], [0, 0, 1.0]])   # rot_z(-90), up to the 1.5708 rounding
    rx = np.array([[1, 0, 0], [0, 0, 1], [0, -1, 0.0]])   # rot_x(-90)
    T02 = (M[:3, :3] @ np.linalg.inv(rz @ rx))[0, 2]
    assert abs(T02 - (cy * sp * cr + sy * sp)) < 1e-4
    assert abs(T02 - (cy * sp * cr + sy * sr)) > 1e-2


def test_quirk_h4_stale_minimum_changes_the_outcome():
    """distance_min is never reset between detections (data_association.h:100-107): a later detection that is far
    from everything inherits the previous detection's match; strict mode maps a new landmark instead."""
    det = lambda x: (0, 0, np.array([x, 0.0, 5.0], dtype=np.float32), np.array([0, 0, 1, 0], dtype=np.float32))
    rp = np.zeros(6, dtype=np.float32)
    out = {}
    for strict in (False, True):
        a = DataAssociation(strict=strict, **KITTI)
        a.find_matches([det(0.0)], rp, 0.0)                    # landmark 0
        res = a.find_matches([det(0.05), det(30.0)], rp, 0.0)  # near landmark 0, then far away
        out[strict] = [(r.id, r.is_new_landmark) for r in res]
    assert out[False] == [(0, False), (0, False)]
    assert out[True] == [(0, False), (1, True)]


def test_dist_is_single_precision():
    d = dist(np.float32(0.1), np.float32(0.4), np.float32(0.2), np.float32(0.2), np.float32(-1.0), np.float32(3.0))
    assert d.dtype == np.float32 and abs(float(d) - np.sqrt(0.09 + 16.0)) < 1e-6


def test_planar_surface_postprocessing_bit_exact():
    """multiPlaneSegmentation's region post-processing + segmentPlanarSurfaces: product vs
    the float32 oracle, bit for bit, incl. the gates, both classes, the rejected 'neither' case and the sign flips"""
    from semantic_slam_b200.association import segment_planar_surfaces
    from oracle.association import segment_planar_surfaces as oracle_sps
    rng = np.random.default_rng(3)
    n_h = n_v = n_skip = 0
    for trial in range(40):
        rp = np.array([rng.normal(), rng.normal(), 0.3, rng.normal(0, 0.03), rng.normal(0, 0.03), rng.uniform(-3, 3)], dtype=np.float32)
        regions = []
        for _ in range(6):
            nrm = rng.normal(0, 1, 3)
            nrm /= np.linalg.norm(nrm)
            regions.append((rng.uniform(-2, 2, 3) + [0, 0, 4], np.r_[nrm, rng.uniform(-5, 5)], int(rng.integers(50, 400)),
                            float(rng.uniform(0.0, 0.2))))
        a = segment_planar_surfaces(regions, rp, 0.15, object_type=2, prob=0.7, planar_area=0.05)
        o = oracle_sps(regions, rp, 0.15, object_type=2, prob=0.7, planar_area=0.05)
        assert len(a) == len(o)
        for x, y in zip(a, o):
            assert x[0] == y[0] == 2 and x[1] == y[1]
            assert np.array_equal(x[2], y[2]) and np.array_equal(x[3], y[3]) and np.array_equal(x[4], y[4])
            n_h += x[1] == 0
            n_v += x[1] == 1
        n_skip += len(regions) - len(a)
    assert n_h > 5 and n_v > 20 and n_skip > 20


def test_planar_surface_classes_on_canonical_planes():
    """camera looking forward (cam_angle 0, level robot): the floor (normal along camera -y... i.e. +-y) is horizontal with
    an upward (negative camera y) normal, a wall facing the camera is vertical with a normal pointing to camera -x / left"""
    from semantic_slam_b200.association import segment_planar_surfaces
    rp = np.zeros(6, dtype=np.float32)
    floor = ([0.0, 1.2, 3.0], [0.0, 1.0, 0.0, -1.2], 500, 1.0)      # camera y points down: floor 1.2 m below the camera
    wall = ([0.5, 0.0, 4.0], [0.6, 0.0, 0.8, -3.5], 500, 1.0)
    small = ([0.0, 0.0, 2.0], [0.0, 0.0, 1.0, -2.0], 60, 1.0)        # contour too small
    out = segment_planar_surfaces([floor, wall, small], rp, 0.0, planar_area=0.1)
    assert [o[1] for o in out] == [0, 1]
    assert np.allclose(out[0][3], [0.0, -1.0, 0.0, 1.2]) and out[0][4][2] < -1.0     # flipped upwards; 1.2 m below the robot
    assert np.allclose(out[1][3], [-0.6, 0.0, -0.8, 3.5])                             # flipped towards the left


def test_edge_cases_empty_frames_and_type_partitions():
    """Empty detection lists, the first non-empty frame (data_association.h:75-96: every detection is mapped as a new
    landmark without association, even two identical ones), a repeat of that frame (every detection matches; of two
    landmarks at the same place the first wins, `distance < distance_min`), a (type, plane_type) class without
    landmarks, and a long frame: product == oracle, and the bookkeeping is the reference's."""
    a, o = DataAssociation(strict=True, **KITTI), OracleDataAssociation(strict=True, **KITTI)
    rp = np.zeros(6, dtype=np.float32)
    assert a.find_matches([], rp, 0.0) == [] and o.find_matches([], rp, 0.0) == []
    assert a.num_landmarks() == o.num_landmarks() == 0
    p = np.array([0.5, 0.2, 4.0], dtype=np.float32)
    n = np.array([0, 0, 1, -4.0], dtype=np.float32)
    dets = [(0, 0, p, n), (0, 0, p, n), (1, 0, p, n), (0, 1, p, n)]
    A, O = a.find_matches(dets, rp, 0.0), o.find_matches(dets, rp, 0.0)
    _same(A, O)
    assert [x.is_new_landmark for x in A] == [True] * 4 and [x.id for x in A] == [0, 1, 2, 3]
    assert a.num_landmarks() == o.num_landmarks() == 4
    assert a.find_matches([], rp, 0.0) == [] and o.find_matches([], rp, 0.0) == []
    A, O = a.find_matches(dets + [(1, 1, p, n)], rp, 0.0), o.find_matches(dets + [(1, 1, p, n)], rp, 0.0)
    _same(A, O)
    assert [x.is_new_landmark for x in A] == [False, False, False, False, True]   # class (1, 1) had no landmark yet
    assert [x.id for x in A] == [0, 0, 2, 3, 4]
    assert a.num_landmarks() == o.num_landmarks() == 5
    # a long frame: 300 detections around 30 anchors
    rng = np.random.default_rng(4)
    anchors = rng.uniform(-20, 20, (30, 3)).astype(np.float32) + np.array([0, 0, 30], dtype=np.float32)
    big = [(int(k % 2), int((k // 2) % 2), (anchors[k % 30] + rng.normal(0, 0.05, 3)).astype(np.float32), n) for k in range(300)]
    _same(a.find_matches(big, rp, 0.1), o.find_matches(big, rp, 0.1))
    assert a.num_landmarks() == o.num_landmarks()


@pytest.mark.parametrize("tag,kw", [("eq", KITTI), ("maha", dict(use_maha_dist=True, maha_dist_thres=3.0, land_noise_low=0.4))])
def test_golden_association_stream(tag, kw):
    """tests/golden/assoc_stream_oracle.npz (scripts/make_golden.py): ids, new-landmark flags and float32 world poses
    over a 40-frame stream; the product's host code and the oracle must both reproduce the fixture bit for bit."""
    import os
    from semantic_slam_b200 import synth
    from semantic_slam_b200.semantic_graph_slam import matrix2vector
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "assoc_stream_oracle.npz"))
    stream = synth.make_frame_stream(40, 10)
    for cls in (DataAssociation, OracleDataAssociation):
        a = cls(**kw)
        ids, new, pose = [], [], []
        est = {}
        for k in range(40):
            rp = matrix2vector(stream.gt_pose[k]).astype(np.float32)
            for l in a.find_matches(stream.detections[k], rp, stream.cam_angle):
                ids.append(l.id); new.append(bool(l.is_new_landmark)); pose.append(np.asarray(l.pose, dtype=np.float32))
                if l.is_new_landmark:
                    est[l.id] = np.asarray(l.pose, dtype=np.float32).copy()
            for lid in range(a.num_landmarks()):
                est[lid] = (est[lid] + np.float32(0.002) * (lid % 3)).astype(np.float32)
                a.setLandmarkEstimate(lid, est[lid])
        assert np.array_equal(np.array(ids, dtype=np.int32), gold[tag + "_ids"]), cls.__name__
        assert np.array_equal(np.array(new, dtype=bool), gold[tag + "_new"]), cls.__name__
        assert np.array_equal(np.array(pose, dtype=np.float32), gold[tag + "_pose"]), cls.__name__


def test_the_two_float_inverses_product_vs_oracle():
    """Eigen::Matrix3f::inverse() (cofactors: `information = covariance.inverse()`) and Eigen::MatrixXf::inverse() (PartialPivLU +
    solve(Identity): the Mahalanobis gate's dynamic-size Q): the product's C++ against the oracle's numpy-float32 restatement bit
    for bit, both close to the float64 inverse; matrices that need no / one / two row exchanges in the LU"""
    import ctypes as C
    from semantic_slam_b200 import _lib
    from oracle.association import inv3, inv3_lu
    L = _lib.lib()
    L.ssb_assoc_inverse3.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(7)
    mats = []
    for _ in range(60):
        B = rng.normal(size=(3, 3))
        mats.append(B @ B.T + 0.05 * np.eye(3))                         # SPD, like Q = H sigma H' + Q_
    for _ in range(60):
        mats.append(rng.normal(size=(3, 3)))                             # general: pivoting happens
    mats.append(np.array([[0.0, 2.0, 1.0], [1.0, 0.0, 3.0], [4.0, 1.0, 0.0]]))   # zero leading entry: exchanges in both steps
    swaps = 0
    for M in mats:
        a = np.ascontiguousarray(M, dtype=np.float32)
        swaps += int(np.argmax(np.abs(a[:, 0])) != 0)
        for kind, ofun in ((0, inv3), (1, inv3_lu)):
            r = np.zeros((3, 3), dtype=np.float32)
            assert L.ssb_assoc_inverse3(kind, a.ctypes.data, r.ctypes.data) == 0
            o = np.asarray(ofun(a), dtype=np.float32)
            assert np.array_equal(r.view(np.uint32), o.view(np.uint32)), (kind, M)
            ref = np.linalg.inv(a.astype(np.float64))
            assert np.abs(r - ref).max() <= 2e-4 * np.abs(ref).max() * np.linalg.cond(a.astype(np.float64))
    assert swaps > 20
