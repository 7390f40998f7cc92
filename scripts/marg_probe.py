"""Time ssb_graph_landmark_marginals (K5) on a per-frame-loop sized graph: columns per PCG launch (SSB_MARG_REPLICAS caps
the number of copies of the graph, 1 = one column per launch).  SSB_MARG_DEBUG=1 prints the PCG iterations of every launch."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from semantic_slam_b200 import GraphSLAM, synth
n_kf = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
n_lm = int(sys.argv[2]) if len(sys.argv) > 2 else 24
spec = synth.make_graph(n_kf, n_kf // 10, seed=77)
g = GraphSLAM(preconditioner=int(os.environ.get("PRECOND", "3")), pcg_tol=float(os.environ.get("TOL", "1e-8")))
ids = synth.load_graph(g, spec)
g.optimize(10)
lm = ids[spec.vkind == 1][:n_lm].astype(np.int32)
ref = None
for cap in sys.argv[3:] or ["1", "2", "4", "16"]:
    os.environ["SSB_MARG_REPLICAS"] = cap
    g.computeLandmarkMarginals(lm[:2])
    t0 = time.perf_counter()
    M = g.computeLandmarkMarginals(lm)
    dt = time.perf_counter() - t0
    ref = M if ref is None else ref
    print(f"copies <= {cap}: marginals of {lm.size} landmarks on {n_kf} keyframes: {dt*1e3:.1f} ms ({dt*1e3/lm.size/3:.3f} ms per column)  "
          f"max rel diff vs first {np.abs(M-ref).max()/np.abs(ref).max():.2e}", flush=True)
