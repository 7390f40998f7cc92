"""Device time of the organised segmentation pipeline (ssb_organized_planes) on a synthetic frame."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from semantic_slam_b200 import OrganizedSegmentation, CloudLayout, synth
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 3
cl = synth.make_cloud(n_boxes=nb, n_hyp=1, nan_frac=0.0005, box_min=150, box_max=200, seed=77)
lay = CloudLayout(cl.width, cl.height, cl.point_step, cl.row_step, cl.offsets)
seg = OrganizedSegmentation(num_point_seg=500)
for it in range(4):
    t = time.perf_counter()
    reg, nreg, nin = seg.segment(cl.msg, lay, cl.boxes, max_regions=8)
    print(f"call {it}: wall {1e3*(time.perf_counter()-t):.2f} ms, device {seg.last_ms:.2f} ms, regions {nreg.tolist()[:8]}", flush=True)
