"""cfg2: parameter difference vs the oracle after k LM iterations (k = 1, 2, 3, 5, 10) for several PCG tolerances:
the north_star bar (1e-5 relative) must hold after ANY iteration count, not only at convergence."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import oracle
from semantic_slam_b200 import GraphSLAM, synth
from parity import pose_errors, point_error
spec = synth.make_config_graph("cfg2")
for k in [1, 2, 3, 5, 10]:
    o = oracle.OracleGraphSLAM(threads=8)
    synth.load_graph(o, spec)
    o.optimize(k)
    Po, Xo = o.get_all(spec.n_poses, spec.n_landmarks)
    row = {"k": k}
    for tol in [1e-3, 1e-4, 1e-5, 1e-6]:
        g = GraphSLAM(preconditioner=3, pcg_tol=tol)
        synth.load_graph(g, spec)
        g.optimize(k)
        P, X = g.get_all(spec.n_poses, spec.n_landmarks)
        rot, tr = pose_errors(P, Po)
        row[str(tol)] = [float("%.2e" % rot), float("%.2e" % tr), float("%.2e" % point_error(X, Xo))]
    print(json.dumps(row), flush=True)
