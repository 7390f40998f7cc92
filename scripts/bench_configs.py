"""Secondary workloads of BASELINE.json (run under gpurun): configs[3] (cfg4: 100k keyframes, one GPU, streaming PCG
kernel) and configs[4] (cfg5: per-frame associate + grow + optimise loop).  Prints one JSON line per workload."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from semantic_slam_b200 import GraphSLAM, DataAssociation, SemanticGraphSLAM, synth

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "cfg4"):
    t = time.time()
    spec = synth.make_config_graph("cfg4")
    g = GraphSLAM(preconditioner=3, pcg_tol=1e-6)
    synth.load_graph(g, spec)
    g.snapshot()
    t_load = time.time() - t
    n_it = int(os.environ.get("CFG4_ITERS", "10"))
    g.optimize_resident(n_it)
    g.restore()
    g.optimize_resident(n_it)
    st = g.stats
    print(json.dumps({"workload": "cfg4: %d KF / %d landmarks / %d edges, %d LM iterations, 1 GPU (streaming k_pcg)" %
                      (spec.n_poses, spec.n_landmarks, spec.n_edges, n_it),
                      "lm_iters_per_s": st["iterations"] / (st["ms_device"] * 1e-3), "ms_device": st["ms_device"],
                      "ms_pcg": st["ms_pcg"], "pcg_iters": st["total_pcg_iters"], "trials": st["total_trials"],
                      "chi2": [st["chi2_initial"], st["chi2_final"]], "load_s": t_load}), flush=True)
if which in ("all", "cfg5"):
    n_kf = int(os.environ.get("CFG5_KF", "4000"))
    stream = synth.make_frame_stream(n_kf, max(12, n_kf // 10), seed=synth.SEED_BASE + 5, max_det=3)
    g = GraphSLAM(preconditioner=3, pcg_tol=1e-6)
    a = DataAssociation(use_maha_dist=False, use_eq_dist=True, eq_dist_thres=1.5, land_noise_low=0.1, strict=True)
    slam = SemanticGraphSLAM(g, a, stream.info6, cam_angle=stream.cam_angle, max_iterations=1024)
    t0 = time.time()
    t_opt = 0.0
    its = 0
    marks = {}
    for k in range(n_kf):
        slam.feed(stream.odom[k], stream.detections[k])
        slam.run()
        if g.stats is not None and g.num_edges() >= 10:
            t_opt += g.stats["ms_total"] * 1e-3
            its += g.stats["iterations"]
        if k + 1 in (n_kf // 4, n_kf // 2, n_kf):
            marks[k + 1] = time.time() - t0
    dt = time.time() - t0
    est = np.array([g.get_se3(kf["node"])[:, 3] for kf in slam.keyframes_])
    print(json.dumps({"workload": "cfg5: per-frame loop, %d KF / %d mapped landmarks (%d gt), %d edges" %
                      (n_kf, len(slam.landmark_nodes_), stream.gt_landmarks.shape[0], g.num_edges()),
                      "frames_per_s": n_kf / dt, "wall_s": dt, "optimize_s": t_opt, "lm_iterations_total": its,
                      "elapsed_at_frames": marks,
                      "max_pos_err_vs_gt_m": float(np.abs(est - stream.gt_pose[:, :, 3]).max())}), flush=True)
