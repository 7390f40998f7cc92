"""GPU parity tests of the dormant plane-clustering chain (csrc/ssb_cluster.cuh, through the C-ABI): against the committed
outputs of the real cv2.kmeans / qhull (tests/golden/cluster_*.npz) and against the CPU oracle on the same seeded inputs."""
import os

import numpy as np
import pytest

import oracle
import cluster_cases
from semantic_slam_b200 import PlaneClustering

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def pc():
    return PlaneClustering()


def test_kmeans_bit_identical_to_opencv(pc):
    g = np.load(os.path.join(GOLD, "cluster_kmeans_cv2.npz"))
    for name, data, K, seed in cluster_cases.kmeans_cases():
        comp, lab, cen, st = pc.computeKmeans(data, K, rng_state=seed)
        assert np.array_equal(lab, g[name + "_labels"].astype(np.int32)), (name, int((lab != g[name + "_labels"]).sum()))
        assert np.array_equal(cen.view(np.uint32), g[name + "_centers"].view(np.uint32)), name
        assert comp == pytest.approx(float(g[name + "_compactness"]), rel=1e-12), name
        assert st == oracle.kmeans(data, K, rng_state=seed)[3], name


def test_kmeans_rejects_what_opencv_rejects(pc):
    from semantic_slam_b200 import SsbError
    with pytest.raises(SsbError):
        pc.computeKmeans(np.zeros((3, 3), dtype=np.float32), 4)      # fewer samples than clusters: cv::kmeans throws
    comp, lab, cen, _ = pc.computeKmeans(np.arange(12, dtype=np.float32).reshape(4, 3), 1)
    assert (lab == 0).all() and np.allclose(cen[0], [4.5, 5.5, 6.5])


def test_hull_matches_qhull_and_the_oracle(pc):
    g = np.load(os.path.join(GOLD, "cluster_hull_qhull.npz"))
    for name, pts in cluster_cases.hull_cases():
        n = pts.shape[0]
        P = np.zeros((n, 4), dtype=np.float32)
        P[:, :2] = pts
        P[:, 2] = 1.5
        mask = np.ones(n, dtype=np.uint8)
        rows, src, nin = pc.projectAndHull(P, mask, [0, 0, 1, -1.5])
        assert nin == n and np.array_equal(np.sort(src), g[name + "_vertices"]), name
        ro, so, _ = oracle.project_hull(P, mask, [0, 0, 1, -1.5])
        assert np.array_equal(src, so) and np.array_equal(rows.view(np.uint32), ro.view(np.uint32)), name


def test_projection_and_hull_on_tilted_planes(pc):
    rng = np.random.default_rng(21)
    for t in range(12):
        n = int(rng.integers(1, 6000))
        P = np.zeros((n, 4), dtype=np.float32)
        P[:, :3] = rng.normal(0, 1, (n, 3)) * [1.0, 1.0, 0.02] + [0.2, -0.1, 1.4]
        nrm = rng.normal(0, 1, 3) if t % 3 else np.eye(3)[t // 3 % 3] + rng.normal(0, 0.01, 3)   # incl. planes facing x / y / z
        coef = np.array([*nrm, -float(rng.uniform(0.5, 2))], dtype=np.float32)
        mask = (rng.random(n) < 0.7).astype(np.uint8)
        rows, src, nin = pc.projectAndHull(P, mask, coef)
        ro, so, nio = oracle.project_hull(P, mask, coef)
        assert nin == nio == int(mask.sum())
        assert np.array_equal(src, so), t
        assert np.array_equal(rows.view(np.uint32), ro.view(np.uint32)), t
    # empty selection
    rows, src, nin = pc.projectAndHull(P, np.zeros(n, dtype=np.uint8), coef)
    assert nin == 0 and rows.shape[0] == 0


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_chain_matches_the_oracle(pc, variant):
    c, nrm, T = cluster_cases.scene(variant)
    r = pc.clusterAndSegmentAllPlanes(c, nrm, T)
    o = oracle.cluster_planes(c, nrm, T)
    # both k-means passes, the centroid filter and the cluster membership: exact
    assert np.array_equal(r["labels"], o["labels"])
    assert np.array_equal(r["centers"].view(np.uint32), o["centers"].view(np.uint32))
    assert r["rng_state"] == o["rng_state"]
    assert len(r["clusters"]) == len(o["clusters"]) >= 2
    for f in ("normal", "distance", "normal_label", "distance_label", "n_points"):
        assert np.array_equal(r["clusters"][f], o["clusters"][f]), f
    # the refined plane goes through an fp64 PCA whose summation order differs: 1e-5
    assert np.abs(r["clusters"]["coef"] - o["clusters"]["coef"]).max() <= 1e-5
    # ProjectInliers + ConvexHull on IDENTICAL planes (the oracle re-run with the product's coefficients): exact
    o2 = oracle.cluster_planes(c, nrm, T, coef_override=r["clusters"]["coef"])
    assert np.array_equal(r["clusters"]["n_inliers"], o2["clusters"]["n_inliers"])
    assert np.array_equal(r["clusters"]["n_rows"], o2["clusters"]["n_rows"])
    assert np.array_equal(r["rows"].view(np.uint32), o2["rows"].view(np.uint32))
    # ... and end to end the hulls agree to the plane tolerance wherever the vertex sets coincide
    if r["rows"].shape == o["rows"].shape:
        assert np.abs(r["rows"] - o["rows"]).max() <= 1e-4


def test_chain_gates(pc):
    c, nrm, T = cluster_cases.scene(0)
    few = nrm.copy()
    few[10:] = np.nan
    r = pc.clusterAndSegmentAllPlanes(c, few, T)
    assert len(r["clusters"]) == 0 and r["rows"].shape[0] == 0 and (r["labels"] == -1).all()
    T2 = T.copy()
    T2[2, :3] = [1.0, 0.0, 0.0]
    assert len(pc.clusterAndSegmentAllPlanes(c, nrm, T2)["clusters"]) == 0
    big = PlaneClustering(min_cluster_points=10**6)
    assert len(big.clusterAndSegmentAllPlanes(c, nrm, T)["clusters"]) == 0
    # fixed hypothesis count instead of PCL's stopping rule
    fx = PlaneClustering(ransac_hypotheses=256)
    r = fx.clusterAndSegmentAllPlanes(c, nrm, T)
    o = oracle.cluster_planes(c, nrm, T, ransac_hypotheses=256)
    assert len(r["clusters"]) == len(o["clusters"]) and np.abs(r["clusters"]["coef"] - o["clusters"]["coef"]).max() <= 1e-5
