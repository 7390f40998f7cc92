"""Time the dormant plane-clustering chain (row f4) on the GPU and on the CPU oracle: cv::kmeans of a full-frame sized sample,
the chain on the synthetic scene."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import oracle
import cluster_cases
from semantic_slam_b200 import PlaneClustering

pc = PlaneClustering()
cases = {c[0]: c for c in cluster_cases.kmeans_cases()}
for name in ("scene_normals", "large"):
    _, data, K, seed = cases[name]
    pc.computeKmeans(data, K, rng_state=seed)
    t0 = time.perf_counter()
    g = pc.computeKmeans(data, K, rng_state=seed)
    tg = time.perf_counter() - t0
    t0 = time.perf_counter()
    o = oracle.kmeans(data, K, rng_state=seed)
    to = time.perf_counter() - t0
    print(f"kmeans {name}: n = {data.shape[0]}, K = {K}: GPU {tg*1e3:.2f} ms, CPU oracle {to*1e3:.2f} ms, labels equal {np.array_equal(g[1], o[1])}")
c, nrm, T = cluster_cases.scene(0)
pc.clusterAndSegmentAllPlanes(c, nrm, T)
t0 = time.perf_counter()
r = pc.clusterAndSegmentAllPlanes(c, nrm, T)
tg = time.perf_counter() - t0
t0 = time.perf_counter()
o = oracle.cluster_planes(c, nrm, T)
to = time.perf_counter() - t0
print(f"chain on the {c.shape[0]}-point scene: GPU {tg*1e3:.2f} ms, CPU oracle {to*1e3:.2f} ms, {len(r['clusters'])} clusters, {r['rows'].shape[0]} hull rows")
