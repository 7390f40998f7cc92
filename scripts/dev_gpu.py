"""Developer timing script (run under gpurun): quick numbers for both paths."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from semantic_slam_b200 import GraphSLAM, PlaneSegmentation, CloudLayout, synth

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "graph"):
    name = sys.argv[2] if len(sys.argv) > 2 else "cfg2"
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    spec = synth.make_config_graph(name)
    g = GraphSLAM(preconditioner=int(os.environ.get('PRECOND', '0')), pcg_tol=float(os.environ.get('PCGTOL', '1e-10')), force_generic=bool(int(os.environ.get('GENERIC', '0'))), coarse_refresh=int(os.environ.get('REFRESH', '1')))
    t = time.time(); synth.load_graph(g, spec); print("load", time.time() - t)
    g.snapshot()
    for rep in range(3):
        g.restore()
        t = time.time(); g.optimize_resident(iters); dt = time.time() - t
        print(f"rep {rep}: wall {dt*1e3:.1f} ms  device {g.stats['ms_device']:.1f} ms iters {g.iterations} trials {g.stats['total_trials']} "
              f"pcg {g.stats['total_pcg_iters']} ms_pcg {g.stats['ms_pcg']:.1f} launches {g.stats['kernel_launches']} chi2 {g.stats['chi2_final']:.6f}")
    np.set_printoptions(linewidth=200, precision=6)
    print(g.history)
    import ctypes as C
    tm = np.zeros(8)
    if hasattr(g._L, "ssb_graph_debug_timers"):
        g._L.ssb_graph_debug_timers.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        g._L.ssb_graph_debug_timers(g._h, tm.ctypes.data_as(C.POINTER(C.c_double)))
        names = ["A:landmark", "B:pose", "C1:blockreduce", "C2:gatherpoll(t0)", "C2:sync", "C3:fold+dots", "D:update", "-"] if os.environ.get("FLOW", "1") == "1" else ["p2:pl-loop", "gauss-jordan", "init", "phase1", "barriers", "p2:reduce", "phase3", "p2:diag+pp"]
        print("k_pcg cycles (block 0, all launches since prepare), ms @1.9GHz:", {n: round(v / 1.9e6, 2) for n, v in zip(names, tm)})
    print("us per pcg iter (upper bound):", g.stats['ms_device'] * 1e3 / max(1, g.stats['total_pcg_iters']))
if which in ("all", "ransac"):
    cl = synth.make_cloud()
    lay = CloudLayout(cl.width, cl.height, cl.point_step, cl.row_step, cl.offsets)
    seg = PlaneSegmentation()
    npts = int((cl.boxes[:, 2] * cl.boxes[:, 3]).sum())
    for rep in range(3):
        t = time.time(); res, counts, mask = seg.fit_planes(cl.msg, lay, cl.boxes, cl.triples); dt = time.time() - t
        print(f"ransac e2e rep {rep}: {dt*1e3:.2f} ms  {npts/dt/1e6:.1f} Mpts/s")
    seg.upload(cl.msg, lay, cl.boxes, cl.triples)
    import torch
    for rep in range(3):
        torch.cuda.synchronize()
        t = time.time(); seg.run_resident(); seg.fetch(False, False); dt = time.time() - t
        print(f"ransac resident rep {rep}: {dt*1e3:.2f} ms  {npts/dt/1e6:.1f} Mpts/s ({npts} pts)")
