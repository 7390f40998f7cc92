"""CPU tests pinning the RANSAC oracle (oracle/oracle_ransac.cpp) against an independent numpy
restatement with the same float32 evaluation order, and against the committed golden fixture."""
import os

import numpy as np

import oracle
from semantic_slam_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")
f32 = np.float32


def _np_model(pts, tri):
    p0, p1, p2 = pts[tri[0]], pts[tri[1]], pts[tri[2]]
    a = (p1 - p0).astype(f32); b = (p2 - p0).astype(f32)
    with np.errstate(all="ignore"):
        r = a / b
    if r[0] == r[1] and r[2] == r[1]:
        return None
    c = np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0], 0], dtype=f32)
    n = np.sqrt(f32(f32(c[0] * c[0] + c[1] * c[1]) + f32(c[2] * c[2] + c[3] * c[3])))
    with np.errstate(all="ignore"):
        c = (c / n).astype(f32)
    dot = f32(f32(c[0] * p0[0] + c[1] * p0[1]) + f32(c[2] * p0[2] + c[3] * f32(1)))
    c[3] = f32(-1) * dot
    return c


def _np_count(pts, c, thr):
    with np.errstate(all="ignore"):
        s0 = (c[0] * pts[:, 0]).astype(f32) + (c[1] * pts[:, 1]).astype(f32)
        s1 = (c[2] * pts[:, 2]).astype(f32) + c[3]
        d = np.abs((s0.astype(f32) + s1.astype(f32)).astype(f32))
    return int((d < thr).sum())


def test_counts_match_numpy_restatement():
    cl = synth.make_cloud(n_boxes=3, n_hyp=40, seed=5)
    res, counts, mask = oracle.ransac_plane_batch(cl.msg, cl.width, cl.height, cl.point_step, cl.row_step, cl.offsets,
                                                  cl.boxes, cl.triples)
    thr = np.nextafter(f32(0.01), f32(1)) if float(f32(0.01)) < 0.01 else f32(0.01)
    for b in range(3):
        crop = oracle.crop(cl.msg, cl.width, cl.height, cl.point_step, cl.row_step, cl.offsets, cl.boxes[b])
        pts = crop.reshape(-1, 4)[:, :3]
        for k in range(40):
            c = _np_model(pts, cl.triples[b, k])
            want = 0 if c is None else _np_count(pts, c, thr)
            assert counts[b, k] == want, (b, k)
        assert res["best_count"][b] == counts[b].max() and res["best_hyp"][b] == counts[b].argmax()


def test_crop_layout_and_spurious_rule():
    cl = synth.make_cloud(n_boxes=2, n_hyp=4)
    pc = cl.msg.view(np.float32).reshape(cl.height, cl.width, 8)
    x, y, w, h = cl.boxes[0]
    c = oracle.crop(cl.msg, cl.width, cl.height, cl.point_step, cl.row_step, cl.offsets, cl.boxes[0])
    ref = pc[y:y + h, x:x + w][:, :, [0, 1, 2, 4]]
    assert np.array_equal(c.view(np.uint32), np.ascontiguousarray(ref).view(np.uint32))
    assert oracle.crop(cl.msg, cl.width, cl.height, cl.point_step, cl.row_step, cl.offsets, [600, 0, 41, 10]) is None
    assert oracle.crop(cl.msg, cl.width, cl.height, cl.point_step, cl.row_step, cl.offsets, [0, 0, 640, 480]) is not None


def test_refine_recovers_plane_and_adaptive_mode():
    rng = np.random.default_rng(0)
    cl = synth.make_cloud(n_boxes=1, n_hyp=64, seed=9)
    # overwrite the cloud with one exact plane z = 2 + 0.1 x - 0.05 y plus tiny noise
    pc = cl.msg.view(np.float32).reshape(cl.height, cl.width, 8)
    u, v = np.meshgrid(np.arange(cl.width), np.arange(cl.height))
    X = ((u - 319.5) / 525 * 2).astype(np.float32); Y = ((v - 239.5) / 525 * 2).astype(np.float32)
    Z = (2 + 0.1 * X - 0.05 * Y + rng.normal(0, 5e-4, X.shape)).astype(np.float32)
    pc[..., 0], pc[..., 1], pc[..., 2] = X, Y, Z
    res, counts, mask = oracle.ransac_plane_batch(cl.msg, cl.width, cl.height, cl.point_step, cl.row_step, cl.offsets,
                                                  cl.boxes, cl.triples)
    n = np.array([0.1, -0.05, -1.0]); n /= np.linalg.norm(n)
    r = res["refined"][0]
    s = np.sign(r[:3] @ n)
    assert np.abs(s * r[:3] - n).max() < 2e-3
    assert res["refined_count"][0] > 0.99 * res["n_points"][0]
    res2, counts2, _ = oracle.ransac_plane_batch(cl.msg, cl.width, cl.height, cl.point_step, cl.row_step, cl.offsets,
                                                 cl.boxes, cl.triples, mode=1)
    assert 1 <= res2["iterations"][0] <= 51 and (counts2[0] >= 0).sum() == res2["iterations"][0]


def test_golden_ransac_fixture():
    d = np.load(os.path.join(GOLD, "ransac_8x256_oracle.npz"))
    cl = synth.make_cloud(n_boxes=8, n_hyp=256, seed=4242)
    res, counts, mask = oracle.ransac_plane_batch(cl.msg, cl.width, cl.height, cl.point_step, cl.row_step, cl.offsets,
                                                  cl.boxes, cl.triples)
    assert np.array_equal(counts, d["counts"]) and np.array_equal(res["best_hyp"], d["best_hyp"])
    assert np.array_equal(res["coef"].view(np.uint32), d["coef"].view(np.uint32))
    assert np.array_equal(np.packbits(mask), d["mask"])


def test_mt19937_known_answer_and_boost_uniform_int():
    """the pure-Python engine behind the oracle's PCL sampler: C++11's known answer for mt19937 (10000th output of the
    default-seeded engine), and boost's bucket rule (uniform_int<>(0, INT_MAX) = output >> 1, small ranges by division)"""
    e = oracle._MT19937(5489)
    for _ in range(9999):
        e()
    assert e() == 4123659995
    a, b = oracle._MT19937(12345), oracle._MT19937(12345)
    assert [oracle._boost_uniform_int(a, 0, 2**31 - 1) for _ in range(64)] == [b() >> 1 for _ in range(64)]
    c = oracle._MT19937(7)
    assert all(0 <= oracle._boost_uniform_int(c, 0, 5) <= 5 for _ in range(1000))


def test_pcl_sample_stream_product_vs_oracle():
    """ssb_ransac_pcl_samples (the product's host code: std::mt19937, C++) against the oracle's independent pure-Python
    restatement of PCL's drawIndexSample: bit-exact, plus what the shuffle guarantees (three distinct indices per draw,
    the shuffle state carried from draw to draw)"""
    from semantic_slam_b200 import pcl_sample_stream
    for n, k, seed in ((3, 8, 12345), (4, 40, 12345), (10, 64, 12345), (14000, 512, 12345), (307200, 64, 12345), (977, 100, 99)):
        t = pcl_sample_stream(n, k, seed)
        assert t.dtype == np.int32 and t.shape == (k, 3)
        assert np.array_equal(t, oracle.pcl_sample_stream(n, k, seed)), (n, k, seed)
        assert np.array_equal(t, oracle.pcl_sample_stream_c(n, k, seed)), (n, k, seed)   # the oracle chain's C++ twin
        assert t.min() >= 0 and t.max() < n
        assert np.all((t[:, 0] != t[:, 1]) & (t[:, 1] != t[:, 2]) & (t[:, 0] != t[:, 2]))
    # first draw by hand: shuffled = 0..n-1, rnd_i = mt() >> 1, swap(i, i + rnd_i % (n - i))
    e = oracle._MT19937(12345)
    n = 1000
    sh = list(range(n))
    for i in range(3):
        j = i + (e() >> 1) % (n - i)
        sh[i], sh[j] = sh[j], sh[i]
    assert list(pcl_sample_stream(n, 1)[0]) == sh[:3]
    assert np.array_equal(pcl_sample_stream(2, 5), np.zeros((5, 3), np.int32))      # PCL selects no sample
    batch = pcl_sample_stream([10, 0, 5000], 16)
    assert batch.shape == (3, 16, 3) and np.array_equal(batch[0], pcl_sample_stream(10, 16)) and not batch[1].any()


def test_adaptive_mode_on_the_pcl_sample_stream():
    """PCL behaviour end to end on the CPU side: the oracle's adaptive RANSAC fed with PCL's own sample stream finds the
    plane of a planar crop within the 50-iteration cap"""
    rng = np.random.default_rng(1)
    cl = synth.make_cloud(n_boxes=2, n_hyp=8, seed=11)
    pc = cl.msg.view(np.float32).reshape(cl.height, cl.width, 8)
    u, v = np.meshgrid(np.arange(cl.width), np.arange(cl.height))
    X = ((u - 319.5) / 525 * 2).astype(np.float32); Y = ((v - 239.5) / 525 * 2).astype(np.float32)
    pc[..., 0], pc[..., 1], pc[..., 2] = X, Y, (2 - 0.2 * X + 0.1 * Y + rng.normal(0, 5e-4, X.shape)).astype(np.float32)
    n = cl.boxes[:, 2].astype(np.int64) * cl.boxes[:, 3]
    tri = np.stack([oracle.pcl_sample_stream(int(m), 512) for m in n])
    res, counts, _ = oracle.ransac_plane_batch(cl.msg, cl.width, cl.height, cl.point_step, cl.row_step, cl.offsets, cl.boxes, tri,
                                               mode=1)
    assert np.all(res["status"] == 0) and np.all(res["iterations"] <= 51)
    assert np.all(res["refined_count"] > 0.95 * res["n_points"])
