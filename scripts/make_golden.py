"""Generates the committed golden fixtures under tests/golden/ from the CPU oracle.
(The reference itself cannot be run here — g2o/PCL/Eigen/ROS are absent — so the fixtures pin the
oracle, and the GPU path is compared with them at sizes the oracle needs seconds-to-minutes for.)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from semantic_slam_b200 import synth

out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
os.makedirs(out, exist_ok=True)

spec = synth.make_config_graph("cfg2")
o = oracle.OracleGraphSLAM()
synth.load_graph(o, spec)
o.optimize(20)
P, X = o.get_all(spec.n_poses, spec.n_landmarks)
json.dump({"config": "cfg2", "n_poses": spec.n_poses, "n_landmarks": spec.n_landmarks, "n_edges": spec.n_edges,
           "iterations": int(o.iterations), "terminated": bool(o.terminated),
           "history": o.history.tolist(), "columns": ["chi2_before", "chi2_after", "lambda", "rho", "trials"]},
          open(os.path.join(out, "cfg2_oracle_history.json"), "w"), indent=1)
np.savez_compressed(os.path.join(out, "cfg2_oracle_final.npz"), poses=P.astype(np.float64), landmarks=X)
print("cfg2 done", o.iterations, o.history[-1])

spec = synth.make_config_graph("cfg1")
o = oracle.OracleGraphSLAM()
synth.load_graph(o, spec)
o.optimize(8)
P, X = o.get_all(spec.n_poses, spec.n_landmarks)
json.dump({"config": "cfg1", "history": o.history.tolist(), "poses": P.tolist(), "landmarks": X.tolist()},
          open(os.path.join(out, "cfg1_oracle.json"), "w"))

cl = synth.make_cloud(n_boxes=8, n_hyp=256, seed=4242)
res, counts, mask = oracle.ransac_plane_batch(cl.msg, cl.width, cl.height, cl.point_step, cl.row_step, cl.offsets,
                                              cl.boxes, cl.triples)
np.savez_compressed(os.path.join(out, "ransac_8x256_oracle.npz"), counts=counts, best_hyp=res["best_hyp"],
                    best_count=res["best_count"], coef=res["coef"], refined=res["refined"],
                    refined_count=res["refined_count"], mask=np.packbits(mask))
print("ransac done")
