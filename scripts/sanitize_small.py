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
