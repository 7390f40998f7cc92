"""GPU vs committed oracle fixture on cfg2 for a few PCG tolerances (run under gpurun)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from semantic_slam_b200 import GraphSLAM, synth
gold = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "cfg2_oracle_final.npz"))
hist = np.array(json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "cfg2_oracle_history.json")))["history"])
spec = synth.make_config_graph("cfg2")
for tol in [float(x) for x in (sys.argv[1:] or ["1e-10", "1e-8", "1e-7", "1e-6"])]:
    g = GraphSLAM(preconditioner=int(os.environ.get("PRECOND", "2")), pcg_tol=tol, coarse_refresh=int(os.environ.get("REFRESH", "1")))
    synth.load_graph(g, spec)
    g.snapshot()
    g.optimize_resident(20); g.restore(); g.optimize_resident(20)
    P, X = g.get_all(spec.n_poses, spec.n_landmarks)
    dp = np.abs(P - gold["poses"]).max(); dx = np.abs(X - gold["landmarks"]).max()
    rel = max(dp / max(1, np.abs(gold["poses"]).max()), dx / max(1, np.abs(gold["landmarks"]).max()))
    same_trials = np.array_equal(g.history[:, 4], hist[:, 4])
    print(f"tol {tol:g}: ms {g.stats['ms_device']:.1f} pcg {g.stats['total_pcg_iters']} max|dP| {dp:.2e} max|dX| {dx:.2e} rel {rel:.2e} "
          f"chi2 {g.stats['chi2_final']:.8f} (oracle {hist[-1,1]:.8f}) trials_equal {same_trials}", flush=True)
