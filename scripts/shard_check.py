"""One graph sharded over `world` ranks driven by host threads of this process (ssb_graph_attach_local):
   --virtual : all shards on GPU 0 (74 / 37 CTAs per rank) — the sharded protocol on a single-GPU box
   otherwise : rank r on GPU r.
Compares with the unsharded run and with the oracle; prints one JSON line."""
import argparse, json, os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from semantic_slam_b200 import GraphSLAM, synth


def run_sharded(spec, world, iters, virtual=True, preconditioner=3, pcg_tol=1e-8, key="g", timeout=300, resident_repeat=0,
                force_generic=False, marginals=None):
    cta = {1: 0, 2: 74, 4: 37}[world] if virtual else 0
    graphs = [GraphSLAM(device=0 if virtual else r, preconditioner=preconditioner, pcg_tol=pcg_tol, force_generic=force_generic)
              for r in range(world)]
    out = [None] * world
    err = [None] * world

    def work(r):
        try:
            g = graphs[r]
            synth.load_graph(g, spec)
            g.attach_local(r, world, key, cta)
            g.optimize(iters)
            P, X = g.get_all(spec.n_poses, spec.n_landmarks)
            res = {"P": P, "X": X, "history": g.history.copy(), "stats": dict(g.stats)}
            if marginals is not None:
                res["marginals"] = g.computeLandmarkMarginals(marginals)
            if resident_repeat:
                g.prepare()
                g.snapshot()
                ms = []
                for _ in range(resident_repeat):
                    g.restore()
                    g.optimize_resident(iters)
                    ms.append(g.stats["ms_device"])
                res["resident_ms"] = ms
                res["stats"] = dict(g.stats)
            out[r] = res
        except Exception as e:  # noqa: BLE001
            err[r] = repr(e)

    th = [threading.Thread(target=work, args=(r,), daemon=True) for r in range(world)]
    for t in th:
        t.start()
    t0 = time.time()
    for t in th:
        t.join(max(1.0, timeout - (time.time() - t0)))
    if any(t.is_alive() for t in th):
        raise RuntimeError(f"sharded run timed out; errors so far: {err}")
    if any(err):
        raise RuntimeError(f"sharded run failed: {err}")
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("config", nargs="?", default="cfg1")
    ap.add_argument("--world", type=int, default=2)
    ap.add_argument("--iters", type=int, default=6)
    ap.add_argument("--virtual", action="store_true")
    ap.add_argument("--precond", type=int, default=3)
    ap.add_argument("--tol", type=float, default=1e-8)
    ap.add_argument("--repeat", type=int, default=0)
    ap.add_argument("--oracle", action="store_true")
    ap.add_argument("--generic", action="store_true", help="force the streaming PCG kernel")
    a = ap.parse_args()
    spec = synth.make_config_graph(a.config)
    res = run_sharded(spec, a.world, a.iters, virtual=a.virtual, preconditioner=a.precond, pcg_tol=a.tol, resident_repeat=a.repeat,
                      force_generic=a.generic, timeout=900)
    g1 = GraphSLAM(preconditioner=a.precond, pcg_tol=a.tol, force_generic=a.generic)
    synth.load_graph(g1, spec)
    g1.optimize(a.iters)
    P1, X1 = g1.get_all(spec.n_poses, spec.n_landmarks)
    line = {"config": a.config, "world": a.world, "virtual": a.virtual, "iterations": int(res[0]["stats"]["iterations"]),
            "ranks_identical": all(np.array_equal(res[0]["P"], r["P"]) and np.array_equal(res[0]["X"], r["X"]) for r in res[1:]),
            "max_abs_vs_1gpu": float(max(np.abs(res[0]["P"] - P1).max(), np.abs(res[0]["X"] - X1).max())),
            "chi2": res[0]["history"][:, 1].tolist(), "chi2_1gpu": g1.history[:, 1].tolist(),
            "pcg_iters": int(res[0]["stats"]["total_pcg_iters"]), "pcg_iters_1gpu": int(g1.stats["total_pcg_iters"]),
            "ms_device": res[0]["stats"]["ms_device"], "ms_device_1gpu": g1.stats["ms_device"], "ms_pcg": res[0]["stats"]["ms_pcg"],
            "ms_pcg_1gpu": g1.stats["ms_pcg"]}
    if a.repeat:
        line["resident_ms"] = res[0]["resident_ms"]
    if a.oracle:
        import oracle
        o = oracle.OracleGraphSLAM()
        synth.load_graph(o, spec)
        o.optimize(a.iters)
        Po, Xo = o.get_all(spec.n_poses, spec.n_landmarks)
        line["max_abs_vs_oracle"] = float(max(np.abs(res[0]["P"] - Po).max(), np.abs(res[0]["X"] - Xo).max()))
    print(json.dumps(line), flush=True)
