"""Per-CTA timeline of one k_pcg_flow iteration (library built with EXTRA=-DSSB_FLOW_TRACE; run under gpurun)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from semantic_slam_b200 import GraphSLAM, synth
spec = synth.make_config_graph("cfg2")
g = GraphSLAM(preconditioner=2, pcg_tol=1e-8)
synth.load_graph(g, spec)
g.optimize(3)
buf = np.zeros(8 * 148, dtype=np.uint64)
g._L.ssb_graph_debug_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
g._L.ssb_graph_debug_trace(g._h, buf.ctypes.data_as(C.c_void_p), buf.size)
t = buf.reshape(148, 8).astype(np.int64)
# order of stamps inside an iteration: 7 start, 0 A done (v stored), 1 staged, 2 B+reduce done, 3 gather polled, 4 sync, 5 fold+dots, 6 D done
order = [7, 0, 1, 2, 3, 4, 5, 6]
names = ["start", "A done", "staged", "B+reduce", "gathered", "sync", "fold+dots", "D done"]
t0 = t[:, 7].min()
print("ns relative to the earliest CTA start: min / median / max over CTAs")
for k, nm in zip(order, names):
    c = t[:, k] - t0
    print(f"  {nm:10s} {c.min():7d} {int(np.median(c)):7d} {c.max():7d}   argmax CTA {int(c.argmax())}")
d = np.diff(t[:, order], axis=1)
print("phase durations per CTA (ns): min / median / max")
for j, nm in enumerate(names[1:]):
    print(f"  {nm:10s} {d[:, j].min():7d} {int(np.median(d[:, j])):7d} {d[:, j].max():7d}")
print("distinct timer values:", np.unique(t).size)
