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

# plane landmarks: oracle LM on the default plane graph (numeric Jacobians like g2o)
pspec = synth.make_plane_graph()
o = oracle.OracleGraphSLAM()
pids = synth.load_plane_graph(o, pspec)
o.optimize(6)
planes = [o.get_plane(pids[k]).tolist() for k, v in enumerate(pspec.vertices) if v[0] == "plane"]
poses = [o.get_se3(pids[k]).tolist() for k, v in enumerate(pspec.vertices) if v[0] == "se3"]
json.dump({"history": o.history.tolist(), "planes": planes, "poses": poses},
          open(os.path.join(out, "plane_graph_oracle.json"), "w"))
print("planes done", o.history[-1])

# landmark association: ids, new-landmark flags and float32 poses over a 40-frame stream, both gates
from oracle.association import OracleDataAssociation
from semantic_slam_b200.semantic_graph_slam import matrix2vector
stream = synth.make_frame_stream(40, 10)
rec = {}
for tag, kw in (("eq", dict(use_maha_dist=False, use_eq_dist=True, eq_dist_thres=1.5, land_noise_low=0.1)),
                ("maha", dict(use_maha_dist=True, maha_dist_thres=3.0, land_noise_low=0.4))):
    a = OracleDataAssociation(**kw)
    ids, new, pose = [], [], []
    for k in range(40):
        rp = matrix2vector(stream.gt_pose[k]).astype(np.float32)
        for l in a.find_matches(stream.detections[k], rp, stream.cam_angle):
            ids.append(l.id); new.append(bool(l.is_new_landmark)); pose.append(np.asarray(l.pose, dtype=np.float32))
        for lid in range(a.num_landmarks()):       # stand-in for the optimiser: a deterministic nudge of the estimates
            a.setLandmarkEstimate(lid, a.landmarks[lid].node_estimate + np.float32(0.002) * (lid % 3))
    rec[tag + "_ids"] = np.array(ids, dtype=np.int32)
    rec[tag + "_new"] = np.array(new, dtype=bool)
    rec[tag + "_pose"] = np.array(pose, dtype=np.float32)
np.savez_compressed(os.path.join(out, "assoc_stream_oracle.npz"), **rec)
print("association done", {k: v.shape for k, v in rec.items()})
