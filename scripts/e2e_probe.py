import sys, os, time
sys.path.insert(0, '/root/repo')
import numpy as np
from semantic_slam_b200 import GraphSLAM, synth
spec = synth.make_config_graph("cfg2")
g = GraphSLAM(preconditioner=3, pcg_tol=1e-6, coarse_refresh=int(os.environ.get("REFRESH", "1")))
synth.load_graph(g, spec)
P0, X0 = g.get_all(spec.n_poses, spec.n_landmarks)
for rep in range(4):
    g.set_all(P0, X0); g.invalidate()
    t0 = time.perf_counter(); g.set_all(P0, X0); t1 = time.perf_counter()
    g.optimize(20); t2 = time.perf_counter()
    P1, X1 = g.get_all(spec.n_poses, spec.n_landmarks); t3 = time.perf_counter()
    st = g.stats
    print(f"set_all {1e3*(t1-t0):.2f} optimize {1e3*(t2-t1):.2f} (prepare {st['ms_prepare']:.2f} device {st['ms_device']:.2f} pcg {st['ms_pcg']:.2f} total {st['ms_total']:.2f}) get_all {1e3*(t3-t2):.2f} launches {st['kernel_launches']} pcg_its {st['total_pcg_iters']} chi2 {st['chi2_final']:.9f}")
