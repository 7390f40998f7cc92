"""CPU checks of the oracle restatement of the live segmentation path (oracle/oracle_segment.cpp: PCL's integral-image
normals + organised multi-plane segmentation).  PCL is absent from this image and the reference holds no vectors for it,
so these pin the restatement against ground truth it cannot have been fitted to: analytic planes."""
import numpy as np
import pytest

import oracle
from semantic_slam_b200 import synth


def _plane_cloud(h=120, w=160, n=(0.1, -0.2, -1.0), d0=2.0, noise=0.0, seed=0):
    n = np.asarray(n, dtype=np.float64)
    n /= np.linalg.norm(n)
    fx = 525.0
    u, v = np.meshgrid(np.arange(w), np.arange(h))
    dx, dy = (u - (w - 1) / 2) / fx, (v - (h - 1) / 2) / fx
    z = n[2] * d0 / (n[0] * dx + n[1] * dy + n[2])
    z = z + np.random.default_rng(seed).normal(0, noise, z.shape)
    c = np.zeros((h, w, 4), dtype=np.float32)
    c[..., 0], c[..., 1], c[..., 2] = dx * z, dy * z, z
    return c, n


def test_normals_of_an_analytic_plane():
    c, n = _plane_cloud()
    nrm, dist = oracle.integral_normals(c)
    inner = nrm[20:-20, 20:-20]
    assert np.isfinite(inner[..., :3]).all(), "every interior pixel of a clean plane gets a normal"
    assert np.isnan(nrm[:20]).all() and np.isnan(nrm[:, :20]).all(), "BORDER_POLICY_IGNORE leaves the border NaN"
    # flipped towards the viewpoint (origin): n . p < 0  <=>  the normal points at the camera
    assert (np.einsum("hwk,hwk->hw", inner[..., :3], c[20:-20, 20:-20, :3]) < 0).all()
    assert np.abs(np.abs(inner[..., :3] @ n) - 1.0).max() < 1e-4
    assert inner[..., 3].max() < 2e-3, "curvature of a plane (float covariance: cancellation noise only)"
    assert dist.max() > 20 and dist.min() > 0, "no depth change anywhere"


def test_distance_map_around_a_hole():
    c, _ = _plane_cloud()
    c[60, 80, :3] = np.nan
    _, dist = oracle.integral_normals(c)
    # the NaN pixel and its left / upper neighbour pairs are depth changes (distance 0); the chamfer metric grows from there
    assert dist[60, 80] == 0 and dist[60, 79] == 0 and dist[59, 80] == 0 and dist[60, 81] == 0 and dist[61, 80] == 0
    assert dist[60, 90] == pytest.approx(9.0) and dist[70, 80] == pytest.approx(9.0)
    assert dist[65, 86] == pytest.approx(7.0, rel=1e-6)     # five diagonal steps from (60, 81)


def test_two_planes_are_two_regions():
    # (close range: PCL 1.8 accumulates the region moments in single precision, and at 2 m the cancellation noise of a
    # 4 000-pixel region is already comparable to the 0.001 curvature gate — the restatement keeps that behaviour)
    a, na = _plane_cloud(n=(0.0, 0.0, -1.0), d0=1.0, noise=0.001)
    b, nb = _plane_cloud(n=(0.5, 0.0, -1.0), d0=1.3, noise=0.001, seed=1)
    c = a.copy()
    c[:, 80:] = b[:, 80:]
    r = oracle.organized_planes(c, min_inliers=500)
    assert r["n"] == 2
    got = r["model"][:, :3]
    for truth in (na, nb):
        assert np.abs(np.abs(got @ truth) - 1.0).min() < 1e-3
    # plane equation holds at the centroid; the inliers are on their plane
    for k in range(2):
        assert abs(r["model"][k, :3] @ r["centroid"][k] + r["model"][k, 3]) < 1e-4
        assert r["n_inliers"][k] > 3000 and r["contour_points"][k] >= 0
    lab = r["labels"]
    left, right = lab[30:-30, 25:70], lab[30:-30, 90:135]
    assert len(np.unique(left)) == 1 and len(np.unique(right)) == 1 and left[0, 0] != right[0, 0]


def test_refinement_grows_into_the_normal_less_border():
    c, _ = _plane_cloud(noise=0.0005)
    r = oracle.organized_planes(c, min_inliers=500)
    assert r["n"] == 1
    before = int((r["labels_cc"] == r["labels_cc"][60, 80]).sum())
    assert r["n_inliers"][0] > before, "pixels of the 20-pixel border (no normal, own tiny labels) join the plane"
    assert r["n_inliers"][0] == int((r["labels"] == r["labels"][60, 80]).sum())
    # a plane that fills the whole crop has no neighbour with another label: findLabeledRegionBoundary returns nothing
    assert r["contour_points"][0] == 0 and r["area"][0] == 0.0
    # with a second surface in view the last inlier sits on a label border and the contour closes around the region
    far, _ = _plane_cloud(n=(0.0, 0.3, -1.0), d0=4.0, noise=0.0005, seed=3)
    c[:, 110:] = far[:, 110:]
    r = oracle.organized_planes(c, min_inliers=500)
    # (the trace starts at the LAST inlier; for the region that owns the top-left corner that pixel is the corner itself —
    # no in-bounds neighbour with another label — so only the other region gets a contour)
    assert r["n"] == 2 and r["contour_points"].max() > 100 and r["area"].max() > 0.01


def test_synthetic_frame_crops_yield_planes():
    cl = synth.make_cloud(n_boxes=6, n_hyp=1, nan_frac=0.0005, box_min=150, box_max=200, seed=77)
    found = 0
    for b in range(6):
        crop = oracle.crop(cl.msg, cl.width, cl.height, cl.point_step, cl.row_step, cl.offsets, cl.boxes[b])
        r = oracle.organized_planes(crop, min_inliers=500)
        found += r["n"]
        for k in range(min(r["n"], 64)):
            assert abs(np.linalg.norm(r["model"][k, :3]) - 1.0) < 1e-4
    assert found >= 4
