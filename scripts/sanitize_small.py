"""Small workload for compute-sanitizer (memcheck / racecheck): cfg1 LM with preconditioner 3 + planes + RANSAC."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from semantic_slam_b200 import GraphSLAM, PlaneSegmentation, CloudLayout, synth
spec = synth.make_graph(300, 40, seed=5)
g = GraphSLAM(preconditioner=3)
synth.load_graph(g, spec)
assert g.optimize(3)
print("graph ok", g.stats["chi2_final"], g.stats["total_pcg_iters"])
ps = synth.make_plane_graph(16, 4, 4)
gp = GraphSLAM(preconditioner=3)
synth.load_plane_graph(gp, ps)
assert gp.optimize(2)
print("planes ok", gp.stats["chi2_final"])
cl = synth.make_cloud(n_boxes=2, n_hyp=64)
lay = CloudLayout(cl.width, cl.height, cl.point_step, cl.row_step, cl.offsets)
seg = PlaneSegmentation()
res, counts, mask = seg.fit_planes(cl.msg, lay, cl.boxes, cl.triples)
print("ransac ok", res["best_count"].tolist())
# K5 with several columns per launch (copies of the graph, k_pcg_flow<148, false, true>) and one column per launch
ids = np.array([v for v in range(spec.vkind.size) if spec.vkind[v] == 1][:4], dtype=np.int32)
M = g.computeLandmarkMarginals(ids)
os.environ["SSB_MARG_REPLICAS"] = "1"
M1 = g.computeLandmarkMarginals(ids[:1])
print("marginals ok", float(np.abs(M[0] - M1[0]).max()))
# the dormant clustering chain: k-means (incl. the empty-cluster path), projection + hull, the chain on a small scene
from semantic_slam_b200 import PlaneClustering
pc = PlaneClustering()
rng = np.random.RandomState(3)
blobs = np.concatenate([rng.randn(150, 3) * 0.001 + [0, 0, 1], rng.randn(150, 3) * 0.001 + [1, 0, 0]]).astype(np.float32)
print("kmeans ok", pc.computeKmeans(blobs, 4, rng_state=300)[0])
P = np.zeros((700, 4), dtype=np.float32)
P[:, :3] = rng.randn(700, 3) * [1, 1, 0.01] + [0, 0, 1.5]
print("hull ok", pc.projectAndHull(P, (rng.rand(700) < 0.8).astype(np.uint8), [0.05, -0.02, 1.0, -1.5])[0].shape)
c, T = synth.make_cluster_scene(h=120, w=160)
nrm = np.full((c.shape[0] * c.shape[1], 4), np.nan, dtype=np.float32)
n0 = T[2, :3] / np.linalg.norm(T[2, :3])
ok = np.isfinite(c.reshape(-1, 4)[:, 2])
nrm[ok, :3] = n0 + rng.randn(int(ok.sum()), 3).astype(np.float32) * 0.02
pcs = PlaneClustering(min_cluster_points=200)
r = pcs.clusterAndSegmentAllPlanes(c.reshape(-1, 4), nrm, T)
print("chain ok", len(r["clusters"]), r["rows"].shape)
