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
    sub = slice(0, 6)
    import copy
    cs = copy.copy(cl)
    cs.boxes, cs.triples = cl.boxes[sub], cl.triples[sub]
    ores, ocounts, omask = _oracle(cs)
    assert np.array_equal(counts[sub], ocounts)
    # resident path gives the same answer
    seg.upload(cl.msg, _layout(cl), cl.boxes, cl.triples)
    seg.run_resident()
    r2, c2, m2 = seg.fetch()
    assert np.array_equal(c2, counts) and np.array_equal(m2, mask)
