"""GPU parity tests of the RANSAC plane-fit path vs the CPU oracle (integer-exact counts)."""
import numpy as np
import pytest

import oracle
from semantic_slam_b200 import PlaneSegmentation, CloudLayout, synth

pytestmark = pytest.mark.gpu


def _layout(cl):
    return CloudLayout(cl.width, cl.height, cl.point_step, cl.row_step, cl.offsets)


def _oracle(cl, **kw):
    return oracle.ransac_plane_batch(cl.msg, cl.width, cl.height, cl.point_step, cl.row_step, cl.offsets, cl.boxes,
                                     cl.triples, **kw)


def _compare(res, counts, mask, ores, ocounts, omask):
    assert np.array_equal(counts, ocounts)
    for f in ("status", "n_points", "best_hyp", "best_count", "iterations"):
        assert np.array_equal(res[f], ores[f]), f
    assert np.array_equal(res["coef"].view(np.uint32), ores["coef"].view(np.uint32)), "3-point model must be bit-exact"
    ok = res["status"] == 0
    # refined plane: fp64 PCA with a different summation order -> tolerance; sign-aligned
    a, b = res["refined"][ok], ores["refined"][ok]
    s = np.sign((a[:, :3] * b[:, :3]).sum(1))[:, None]
    assert np.abs(a * s - b).max() <= 1e-5
    # centroid of the winning model's inliers (same inlier set, fp64 accumulation in a different order)
    assert np.abs(res["centroid"][ok] - ores["centroid"][ok]).max() <= 1e-5 * max(1.0, np.abs(ores["centroid"][ok]).max())
    assert np.abs(res["refined_count"][ok].astype(int) - ores["refined_count"][ok]).max() <= np.maximum(
        2, 0.001 * ores["refined_count"][ok]).max()
    if mask is not None:
        assert mask.size == omask.size
        assert (mask != omask).mean() < 1e-3


def test_crop_matches_reference_layout():
    cl = synth.make_cloud(n_boxes=6, n_hyp=8)
    seg = PlaneSegmentation()
    lay = _layout(cl)
    for b in range(6):
        a = seg.segmentPointCloudData(cl.boxes[b], cl.msg, lay)
        o = oracle.crop(cl.msg, cl.width, cl.height, cl.point_step, cl.row_step, cl.offsets, cl.boxes[b])
        assert np.array_equal(a.view(np.uint32), o.view(np.uint32))
    # spurious boxes (plane_segmentation.cpp:34-38)
    for bad in ([600, 10, 50, 50], [10, 450, 50, 50], [10, 10, -5, 50], [-3, 10, 20, 20]):
        assert seg.segmentPointCloudData(np.array(bad), cl.msg, lay) is None
        assert oracle.crop(cl.msg, cl.width, cl.height, cl.point_step, cl.row_step, cl.offsets, np.array(bad)) is None


@pytest.mark.parametrize("nb,K", [(4, 64), (9, 300), (3, 1500)])
def test_fixed_k_counts_integer_exact(nb, K):
    cl = synth.make_cloud(n_boxes=nb, n_hyp=K, seed=100 + nb)
    seg = PlaneSegmentation()
    res, counts, mask = seg.fit_planes(cl.msg, _layout(cl), cl.boxes, cl.triples)
    _compare(res, counts, mask, *_oracle(cl))


def test_edge_cases_spurious_empty_collinear_nan():
    cl = synth.make_cloud(n_boxes=6, n_hyp=32, seed=7)
    cl.boxes[1] = (630, 470, 40, 40)       # spurious
    cl.boxes[2] = (5, 5, 0, 0)             # empty crop
    cl.boxes[3] = (100, 100, 2, 2)         # 4 points
    cl.triples[3] = cl.triples[3] % 4
    cl.triples[0, 0] = (5, 5, 5)           # degenerate sample
    cl.triples[0, 1] = (0, 1, 2)
    seg = PlaneSegmentation()
    res, counts, mask = seg.fit_planes(cl.msg, _layout(cl), cl.boxes, cl.triples)
    _compare(res, counts, mask, *_oracle(cl))
    assert res["status"][1] == 1 and res["status"][2] == 2


def test_pcl_adaptive_mode_replay():
    cl = synth.make_cloud(n_boxes=8, n_hyp=128, seed=11)
    seg = PlaneSegmentation(mode=1)
    res, counts, mask = seg.fit_planes(cl.msg, _layout(cl), cl.boxes, cl.triples)
    _compare(res, counts, mask, *_oracle(cl, mode=1))


def test_no_refine():
    cl = synth.make_cloud(n_boxes=3, n_hyp=64, seed=12)
    seg = PlaneSegmentation(refine=False)
    res, counts, mask = seg.fit_planes(cl.msg, _layout(cl), cl.boxes, cl.triples)
    ores, ocounts, omask = _oracle(cl, refine=False)
    assert np.array_equal(mask, omask)           # without the fp64 PCA everything is bit-exact
    assert np.array_equal(res["refined_count"], ores["refined_count"])


def test_cfg3_full_size():
    """BASELINE.json configs[2]: 640x480, 64 crops x 1024 hypotheses.  Checked against the oracle on a
    subset of crops and through a size-independent property on all of them: the winner's count is the
    maximum of the count table and every count is bounded by the number of finite points."""
    cl = synth.make_cloud()
    seg = PlaneSegmentation()
    res, counts, mask = seg.fit_planes(cl.msg, _layout(cl), cl.boxes, cl.triples)
    assert np.array_equal(res["best_count"], counts.max(1))
    assert np.array_equal(res["best_hyp"], counts.argmax(1))
    # every one of the 64 crops against the oracle: all 65 536 inlier counts, winners, 3-point models, refined inlier masks
    ores, ocounts, omask = _oracle(cl)
    assert np.array_equal(counts, ocounts)
    assert np.array_equal(res["best_hyp"], ores["best_hyp"]) and np.array_equal(res["best_count"], ores["best_count"])
    assert np.array_equal(res["coef"].view(np.uint32), ores["coef"].view(np.uint32))
    assert np.array_equal(res["refined_count"], ores["refined_count"]) and np.array_equal(mask, omask)
    # resident path gives the same answer
    seg.upload(cl.msg, _layout(cl), cl.boxes, cl.triples)
    seg.run_resident()
    r2, c2, m2 = seg.fetch()
    assert np.array_equal(c2, counts) and np.array_equal(m2, mask)


def test_frame_pipeline_cloud_to_association():
    """One frame end to end on the product path: PointCloud2 + boxes -> crop + RANSAC (GPU) -> planar-surface
    post-processing (multiPlaneSegmentation's classification, segmentPlanarSurfaces) -> data association, against
    the same chain over the oracle.  Classes and association indices must agree exactly, positions within 1e-5."""
    from semantic_slam_b200 import DataAssociation
    from semantic_slam_b200.association import segment_planar_surfaces, planar_regions_from_ransac
    from oracle.association import OracleDataAssociation, segment_planar_surfaces as oracle_sps
    cl = synth.make_cloud(n_boxes=12, n_hyp=256)
    lay = CloudLayout(cl.width, cl.height, cl.point_step, cl.row_step, cl.offsets)
    seg = PlaneSegmentation()
    res, counts, mask = seg.fit_planes(cl.msg, lay, cl.boxes, cl.triples)
    ores, ocounts, omask = oracle.ransac_plane_batch(cl.msg, cl.width, cl.height, cl.point_step, cl.row_step,
                                                     cl.offsets, cl.boxes, cl.triples)
    rp = np.array([1.0, -2.0, 0.4, 0.01, -0.02, 0.7], dtype=np.float32)
    kw = dict(use_maha_dist=False, use_eq_dist=True, eq_dist_thres=1.5, land_noise_low=0.1, strict=True)
    a, o = DataAssociation(**kw), OracleDataAssociation(**kw)
    for frame in range(2):   # the second pass re-observes the landmarks mapped by the first
        # the RANSAC normal's sign is arbitrary (smallest eigenvector): the reference's sign conventions remove it
        dets = segment_planar_surfaces(planar_regions_from_ransac(res), rp, 0.1, planar_area=0.0)
        odets = oracle_sps(planar_regions_from_ransac(ores), rp, 0.1, planar_area=0.0)
        assert len(dets) == len(odets) >= 6
        assert [d[1] for d in dets] == [d[1] for d in odets]
        for d, e in zip(dets, odets):
            assert np.abs(d[2] - e[2]).max() <= 1e-5 * max(1.0, np.abs(e[2]).max())
            assert np.abs(np.abs(d[3]) - np.abs(e[3])).max() <= 2e-4     # refined normals: fp64 PCA, different summation order
        A = a.find_matches([d[:4] for d in dets], rp, 0.1)
        O = o.find_matches([d[:4] for d in odets], rp, 0.1)
        assert [(x.id, x.is_new_landmark) for x in A] == [(y.id, y.is_new_landmark) for y in O]
        if frame == 1:
            assert not any(x.is_new_landmark for x in A)
