"""cfg4 (100k keyframes): difference vs the oracle after 3 LM iterations as a function of the PCG tolerance."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import oracle
from semantic_slam_b200 import GraphSLAM, synth
from parity import pose_errors, point_error
spec = synth.make_config_graph("cfg4")
o = oracle.OracleGraphSLAM(threads=8)
synth.load_graph(o, spec)
t = time.time(); o.optimize(3); t_o = time.time() - t
Po, Xo = o.get_all(spec.n_poses, spec.n_landmarks)
for tol in [1e-8, 1e-10, 1e-12]:
    g = GraphSLAM(preconditioner=3, pcg_tol=tol, max_pcg_iters=20000)
    synth.load_graph(g, spec)
    g.optimize(3)
    P, X = g.get_all(spec.n_poses, spec.n_landmarks)
    rot, tr = pose_errors(P, Po)
    print(json.dumps({"tol": tol, "rot": rot, "trans_rel": tr, "lm_rel": point_error(X, Xo), "pcg_iters": g.stats["total_pcg_iters"],
                      "ms_device": g.stats["ms_device"], "chi2": g.history[:, 1].tolist(), "chi2_oracle": o.history[:, 1].tolist(),
                      "oracle_s": t_o}), flush=True)
