"""cfg2, 20 LM iterations: end-state difference vs the committed oracle fixture and device time as a function of the
PCG tolerance (relative M^-1-norm of the residual)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from semantic_slam_b200 import GraphSLAM, synth
from parity import pose_errors, point_error
here = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
gold = np.load(os.path.join(here, "cfg2_oracle_final.npz"))
hist = np.array(json.load(open(os.path.join(here, "cfg2_oracle_history.json")))["history"])
spec = synth.make_config_graph("cfg2")
for tol in [1e-3, 1e-4, 1e-5, 1e-6, 1e-8]:
    g = GraphSLAM(preconditioner=3, pcg_tol=tol)
    synth.load_graph(g, spec)
    g.prepare(); g.snapshot(); g.optimize_resident(20); g.restore(); g.optimize_resident(20)
    P, X = g.get_all(spec.n_poses, spec.n_landmarks)
    rot, tr = pose_errors(P, gold["poses"])
    print(json.dumps({"tol": tol, "ms_device": g.stats["ms_device"], "pcg_iters": g.stats["total_pcg_iters"], "trials": g.stats["total_trials"],
                      "rot_err": rot, "trans_rel_err": tr, "lm_rel_err": point_error(X, gold["landmarks"]),
                      "same_decisions": bool(np.array_equal(g.history[:, 4], hist[:20, 4])),
                      "chi2_final_rel": abs(g.history[-1, 1] - hist[19, 1]) / hist[19, 1],
                      "max_chi2_traj_rel": float(np.abs(g.history[:, 1] / hist[:20, 1] - 1).max())}), flush=True)
