import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from semantic_slam_b200 import GraphSLAM, DataAssociation, SemanticGraphSLAM, synth
n_kf = int(sys.argv[1]) if len(sys.argv) > 1 else 600
stream = synth.make_frame_stream(n_kf, max(12, n_kf // 10), seed=synth.SEED_BASE + 5, max_det=3)
g = GraphSLAM(preconditioner=int(os.environ.get("PRECOND", "3")), pcg_tol=1e-6)
a = DataAssociation(use_maha_dist=False, use_eq_dist=True, eq_dist_thres=1.5, land_noise_low=0.1, strict=True)
slam = SemanticGraphSLAM(g, a, stream.info6, cam_angle=stream.cam_angle, max_iterations=1024)
t0 = time.time()
for k in range(n_kf):
    slam.feed(stream.odom[k], stream.detections[k])
    t = time.perf_counter()
    slam.run()
    dt = time.perf_counter() - t
    if (k + 1) % (n_kf // 6) == 0 and g.stats:
        st = g.stats
        print(f"frame {k+1}: run {dt*1e3:.1f} ms  lm_its {st['iterations']} trials {st['total_trials']} pcg {st['total_pcg_iters']} "
              f"prepare {st['ms_prepare']:.2f} device {st['ms_device']:.2f} pcg_ms {st['ms_pcg']:.2f} total {st['ms_total']:.2f} launches {st['kernel_launches']}")
print("total", time.time() - t0)
# K5 on the final graph: all mapped landmarks
lm = np.array(sorted(slam.landmark_nodes_.values()), dtype=np.int32)
g.computeLandmarkMarginals(lm[:2])
t = time.perf_counter()
M = g.computeLandmarkMarginals(lm)
print(f"marginals of {lm.size} landmarks on {n_kf} keyframes: {(time.perf_counter()-t)*1e3:.1f} ms  trace0 {np.trace(M[0]):.9e}")
