"""Times computeLandmarkMarginals (K5) in its direct form (ssb_marg_direct.cuh) against the iterative form and the oracle's
g2o recursion: all landmarks of a 1 000-keyframe graph, then all landmarks of cfg2.  Run under gpurun."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from semantic_slam_b200 import GraphSLAM, synth


def run(name, spec, its, iterative_sample=None):
    g = GraphSLAM(preconditioner=3, pcg_tol=1e-8)
    ids = synth.load_graph(g, spec)
    g.optimize(its)
    lm = ids[spec.vkind == 1].astype(np.int32)
    os.environ["SSB_MARG_DIRECT"] = "1"
    g.computeLandmarkMarginals(lm[:2])
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); Md = g.computeLandmarkMarginals(lm); ts.append(time.perf_counter() - t0)
    os.environ["SSB_MARG_DIRECT"] = "0"
    sub = lm if iterative_sample is None else lm[:: max(1, lm.size // iterative_sample)][:iterative_sample]
    g.computeLandmarkMarginals(sub[:2])
    t0 = time.perf_counter(); Mi = g.computeLandmarkMarginals(sub); ti = time.perf_counter() - t0
    os.environ["SSB_MARG_DIRECT"] = "1"
    o = oracle.OracleGraphSLAM(threads=1)
    synth.load_graph(o, spec)
    o.optimize(its)
    t0 = time.perf_counter(); Mo = o.computeLandmarkMarginals(lm, method="g2o"); to = time.perf_counter() - t0
    pos = {int(v): k for k, v in enumerate(lm)}
    sel = np.array([pos[int(v)] for v in sub])
    print("%s: %d landmarks | direct %.2f ms (runs: %s) | iterative %.2f ms for %d landmarks (%.3f ms each) | oracle g2o recursion %.2f ms | "
          "rel diff direct vs oracle %.2e, iterative vs oracle %.2e" % (
              name, lm.size, 1e3 * min(ts), ", ".join("%.2f" % (1e3 * t) for t in ts), 1e3 * ti, sub.size, 1e3 * ti / sub.size, 1e3 * to,
              np.abs(Md - Mo).max() / np.abs(Mo).max(), np.abs(Mi - Mo[sel]).max() / np.abs(Mo).max()), flush=True)


run("250 keyframes", synth.make_graph(250, 40, seed=synth.SEED_BASE + 5), 10)
run("1000 keyframes", synth.make_graph(1000, 100, seed=synth.SEED_BASE + 6), 10)
if len(sys.argv) > 1 and sys.argv[1] == "cfg2":
    run("cfg2", synth.make_config_graph("cfg2"), 3, iterative_sample=16)
