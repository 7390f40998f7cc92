"""Time ssb_graph_landmark_marginals (K5) on a cfg5-size graph; SSB_MARG_INKERNEL=1 re-inverts the coarse matrix
inside every solve (the behaviour before k_coarse_invert)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from semantic_slam_b200 import GraphSLAM, synth
n_kf = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
spec = synth.make_graph(n_kf, n_kf // 5, seed=77)
g = GraphSLAM(preconditioner=3, pcg_tol=1e-8)
ids = synth.load_graph(g, spec)
g.optimize(10)
lm = ids[spec.vkind == 1][:64].astype(np.int32)
for rep in range(3):
    t0 = time.perf_counter()
    M = g.computeLandmarkMarginals(lm)
    dt = time.perf_counter() - t0
    print(f"marginals of {lm.size} landmarks: {dt*1e3:.1f} ms ({dt*1e3/lm.size/3:.3f} ms per solve)  trace {np.trace(M[0]):.6e}")
