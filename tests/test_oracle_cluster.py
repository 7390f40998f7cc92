"""CPU checks of the oracle restatement of the dormant plane-clustering chain (oracle/oracle_cluster.cpp) against outputs of
the REAL third-party implementations: cv2.kmeans (OpenCV 4.13) and qhull (scipy.spatial.ConvexHull), committed as fixtures by
scripts/make_cluster_golden.py (tests/golden/cluster_*.npz).  Neither library exists on the GPU box; the fixtures travel."""
import os

import numpy as np
import pytest

import oracle
import cluster_cases

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gold_kmeans():
    return np.load(os.path.join(GOLD, "cluster_kmeans_cv2.npz"))


def test_kmeans_restatement_is_bit_identical_to_opencv(gold_kmeans):
    g = gold_kmeans
    for name, data, K, seed in cluster_cases.kmeans_cases():
        assert cluster_cases.crc(data) == int(g[name + "_input_crc"]), f"{name}: the seeded input changed, regenerate the fixture"
        comp, lab, cen, _ = oracle.kmeans(data, K, rng_state=seed)
        assert np.array_equal(lab, g[name + "_labels"].astype(np.int32)), name
        assert np.array_equal(cen.view(np.uint32), g[name + "_centers"].view(np.uint32)), name
        # cv::sum adds the per-point distances in its own (vectorised) order: the sum agrees to rounding
        assert comp == pytest.approx(float(g[name + "_compactness"]), rel=1e-12), name


def test_kmeans_rng_stream_continues_across_calls():
    """the two k-means passes of the chain share cv::theRNG(): the state after one call seeds the next"""
    cases = cluster_cases.kmeans_cases()
    _, d0, K0, seed = cases[2]
    _, d1, K1, _ = cases[3]
    _, _, _, st = oracle.kmeans(d0, K0, rng_state=seed)
    assert st != seed
    a = oracle.kmeans(d1, K1, rng_state=st)
    b = oracle.kmeans(d1, K1, rng_state=seed)
    assert a[3] != b[3]
    # 10 attempts x K x dims draws per call: the state is a pure function of the draw count
    s = seed
    for _ in range(10 * K0 * d0.shape[1]):
        s = ((s & 0xFFFFFFFF) * 4164903690 + (s >> 32)) & 0xFFFFFFFFFFFFFFFF
    assert s == st


def test_hull_restatement_matches_qhull():
    g = np.load(os.path.join(GOLD, "cluster_hull_qhull.npz"))
    for name, pts in cluster_cases.hull_cases():
        assert cluster_cases.crc(pts) == int(g[name + "_input_crc"])
        n = pts.shape[0]
        P = np.zeros((n, 4), dtype=np.float32)
        P[:, :2] = pts
        P[:, 2] = 1.5
        rows, src, nin = oracle.project_hull(P, np.ones(n, dtype=np.uint8), [0, 0, 1, -1.5])
        assert nin == n
        assert np.array_equal(np.sort(src), g[name + "_vertices"]), name
        # PCL's output order: decreasing angle about the centroid of the hull vertices
        c = P[src, :2].mean(axis=0)
        ang = np.arctan2(P[src, 1] - c[1], P[src, 0] - c[0])
        assert (np.diff(ang) < 0).all()
        assert np.allclose(rows[:, :2], P[src, :2], atol=1e-6) and np.allclose(rows[:, 2], 1.5, atol=1e-6)


def test_projection_lands_on_the_plane_and_respects_the_mask():
    rng = np.random.default_rng(0)
    n = 500
    P = np.zeros((n, 4), dtype=np.float32)
    P[:, :3] = rng.normal(0, 1, (n, 3))
    coef = np.array([0.2, -0.3, 0.9, -0.7], dtype=np.float32)
    mask = (rng.random(n) < 0.5).astype(np.uint8)
    rows, src, nin = oracle.project_hull(P, mask, coef)
    assert nin == int(mask.sum()) and mask[src].all()
    nn = coef[:3] / np.linalg.norm(coef[:3])
    # hull vertices are projections of their source points along the normal ...
    d = (rows - P[src, :3]) @ nn
    assert np.allclose(rows - P[src, :3], np.outer(d, nn), atol=1e-5)
    # ... onto the plane n.x + d4 = 0 of the NORMALISED normal (ProjectInliers keeps the model's 4th coefficient)
    assert np.abs(rows @ nn + coef[3]).max() < 1e-5
    # axes: a plane facing x is hulled in (y, z)
    rows2, src2, _ = oracle.project_hull(P, mask, [1.0, 0.01, 0.02, 0.0])
    assert len(src2) >= 3 and np.abs(rows2[:, 0]).max() < 0.2


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_chain_on_the_synthetic_scene(variant):
    c, nrm, T = cluster_cases.scene(variant)
    r = oracle.cluster_planes(c, nrm, T, rng_state=0xFFFFFFFF)
    cl = r["clusters"]
    assert len(cl) >= 2
    hz = T[2, :3]
    lab = r["labels"]
    assert ((lab >= 0) == ~np.isnan(nrm[:, :3]).any(axis=1)).all()
    for k in cl:
        # filterCentroids: the kept normal is within 0.3 per component of the horizontal normal; clusters have > 500 points
        assert (np.abs(k["normal"] - hz) < 0.3).all() and k["n_points"] > 500
        assert np.array_equal(k["normal"], r["centers"][k["normal_label"]])
        rows = r["rows"][k["row0"]:k["row0"] + k["n_rows"]]
        assert k["n_rows"] >= 3 and (rows[:, 3:6] == k["normal"]).all() and (rows[:, 6] == k["distance"]).all() and (rows[:, 7] == 0).all()
        # hull vertices lie on the fitted plane, whose normal is the scene's horizontal normal (up to sign)
        nn = k["coef"][:3] / np.linalg.norm(k["coef"][:3])
        assert np.abs(rows[:, :3] @ nn + k["coef"][3]).max() < 1e-4
        assert abs(abs(float(nn @ hz)) - 1.0) < 2e-3
        # the k-means distance centroid is the plane's offset along the (approximate) normal centroid
        if k["n_points"] > 5000:
            assert abs(abs(k["distance"]) - abs(k["coef"][3])) < 0.1
    # the two parallel planes of the scene are separated by the distance k-means
    d = sorted(abs(float(k["coef"][3])) for k in cl if k["n_points"] > 5000)
    assert len(d) == 2 and d[1] - d[0] > 0.3
    # determinism + the RNG state is consumed
    r2 = oracle.cluster_planes(c, nrm, T, rng_state=0xFFFFFFFF)
    assert np.array_equal(r["rows"], r2["rows"]) and r["rng_state"] == r2["rng_state"] != 0xFFFFFFFF


def test_chain_gates():
    c, nrm, T = cluster_cases.scene(0)
    few = nrm.copy()
    few[10:] = np.nan
    r = oracle.cluster_planes(c, few, T)           # <= 10 valid normals: nothing (plane_segmentation.cpp:316-320)
    assert len(r["clusters"]) == 0 and r["rows"].shape[0] == 0
    T2 = T.copy()
    T2[2, :3] = [1.0, 0.0, 0.0]                    # no centroid near this "horizontal" normal: filterCentroids drops all
    r = oracle.cluster_planes(c, nrm, T2)
    assert len(r["clusters"]) == 0
    r = oracle.cluster_planes(c, nrm, T, min_cluster_points=10**6)
    assert len(r["clusters"]) == 0
