"""The per-frame loop with getAndSetLandmarkCov after EVERY optimise, as the reference does (semantic_graph_slam.cpp:89,181-205):
what the landmark marginals add per tick.  `gpu [n]`: the product (direct form); `cpu [n]`: the oracle with g2o's recursion."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from semantic_slam_b200 import synth
from semantic_slam_b200.semantic_graph_slam import SemanticGraphSLAM

which = sys.argv[1] if len(sys.argv) > 1 else "gpu"
n_kf = int(sys.argv[2]) if len(sys.argv) > 2 else 400
stream = synth.make_frame_stream(n_kf, max(12, n_kf // 10), seed=synth.SEED_BASE + 5, max_det=3)
kw = dict(use_maha_dist=False, use_eq_dist=True, eq_dist_thres=1.5, land_noise_low=0.1, strict=True)
if which == "gpu":
    from semantic_slam_b200 import GraphSLAM, DataAssociation
    g, a, mk = GraphSLAM(preconditioner=3, pcg_tol=1e-6), DataAssociation(**kw), {}
else:
    import oracle
    from oracle.association import OracleDataAssociation
    g, a, mk = oracle.OracleGraphSLAM(threads=1), OracleDataAssociation(**kw), {"method": "g2o"}
slam = SemanticGraphSLAM(g, a, stream.info6, cam_angle=stream.cam_angle, max_iterations=1024, always_marginals=True, marginals_kwargs=mk)
t0 = time.perf_counter()
marks = {}
for k in range(n_kf):
    slam.feed(stream.odom[k], stream.detections[k])
    slam.run()
    if (k + 1) % 100 == 0:
        marks[k + 1] = (time.perf_counter() - t0, slam.marginals_seconds, len(slam.landmark_nodes_))
for k, (t, tm, nl) in marks.items():
    print("%s loop, %4d frames: %.2f s in all, of which landmark marginals %.3f s (%d landmarks mapped)" % (which, k, t, tm, nl), flush=True)
print("%s: marginals %.2f ms per call on average over %d calls" % (which, 1e3 * slam.marginals_seconds / max(1, slam.marginals_calls),
                                                                   slam.marginals_calls))
