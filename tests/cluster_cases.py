"""Seeded inputs of the plane-clustering tests, shared by scripts/make_cluster_golden.py (which runs the real cv2.kmeans / qhull
on them in the build container) and by the tests (which check the oracle and the CUDA path against the stored outputs)."""
import zlib

import numpy as np

import oracle
from semantic_slam_b200 import synth


def crc(a) -> int:
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def scene(variant: int = 0):
    """(cloud (n, 4), normals (n, 4) from the integral-image normals, transformation_mat)"""
    kw = [dict(), dict(nan_frac=0.0003), dict(seed=9, d_b=1.5, noise=0.001)][variant]
    c, T = synth.make_cluster_scene(**kw)
    nrm, _ = oracle.integral_normals(c)
    return c.reshape(-1, 4), nrm.reshape(-1, 4), T


def kmeans_cases():
    """(name, data float32 (N, dims), K, cv RNG seed)"""
    cases = []
    c, nrm, T = scene(0)
    ok = ~np.isnan(nrm[:, :3]).any(axis=1)
    cases.append(("scene_normals", np.ascontiguousarray(nrm[ok, :3]), 4, 4242))
    d = -(c[ok, 0] * np.float32(0.1) + c[ok, 1] * np.float32(-0.5) + c[ok, 2] * np.float32(-0.85))
    cases.append(("scene_distances", np.ascontiguousarray(d[:, None].astype(np.float32)), 2, 77))
    rng = np.random.RandomState(3)
    for t in range(6):
        n = int(rng.randint(11, 4000))
        cs = rng.randn(3, 3)
        cs /= np.linalg.norm(cs, axis=1, keepdims=True)
        cases.append(("rand3d_%d" % t, (cs[rng.randint(0, 3, n)] + 0.05 * rng.randn(n, 3)).astype(np.float32), 4, 100 + t))
    for t in range(4):
        n = int(rng.randint(11, 3000))
        a = np.concatenate([rng.randn(n // 2, 1) * 0.02 + 1.0, rng.randn(n - n // 2, 1) * 0.02 + 1.5])
        cases.append(("rand1d_%d" % t, a.astype(np.float32), 2, 200 + t))
    # two tight blobs, four clusters asked for: random centres in the bounding box leave clusters empty -> the re-seeding path
    for t in range(4):
        n = 600 + 50 * t
        a = np.concatenate([rng.randn(n // 2, 3) * 0.001 + [0, 0, 1], rng.randn(n - n // 2, 3) * 0.001 + [1, 0, 0]])
        cases.append(("empty_%d" % t, a.astype(np.float32), 4, 300 + t))
    # a large case: float sums over 300 000 samples (the size of a full 640 x 480 frame)
    n = 300000
    cs = rng.randn(4, 3)
    cs /= np.linalg.norm(cs, axis=1, keepdims=True)
    cases.append(("large", (cs[rng.randint(0, 4, n)] + 0.08 * rng.randn(n, 3)).astype(np.float32), 4, 999))
    return cases


def hull_cases():
    """(name, points float32 (n, 2))"""
    rng = np.random.default_rng(11)
    cases = []
    for t in range(8):
        n = int(rng.integers(5, 20000))
        p = rng.normal(0, 1, (n, 2)) if t % 2 else rng.uniform(-1, 1, (n, 2))
        cases.append(("hull_%d" % t, p.astype(np.float32)))
    return cases
