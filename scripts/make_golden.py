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

# cfg1 as a g2o text file (the format GraphSLAM::save writes, graph_slam.cpp:236-239): with it a real g2o can close the
# loop off-box —   g2o -solver lm_var -i 8 -o out.g2o tests/golden/cfg1.g2o   must reproduce tests/golden/cfg1_oracle.json
# (chi2 per iteration in its verbose output, final vertices in out.g2o).
def write_g2o(spec, path):
    from scipy.spatial.transform import Rotation
    with open(path, "w") as f:
        f.write("PARAMS_SE3OFFSET 0 0 0 0 0 0 0 1\n")
        for v in range(spec.vkind.size):
            if spec.vkind[v] == 0:
                T = spec.vpose[v]
                q = Rotation.from_matrix(T[:, :3]).as_quat()
                if q[3] < 0:
                    q = -q
                f.write("VERTEX_SE3:QUAT %d %s\n" % (v, " ".join("%.17g" % x for x in list(T[:, 3]) + list(q))))
                if v == 0:
                    f.write("FIX 0\n")
            else:
                f.write("VERTEX_TRACKXYZ %d %s\n" % (v, " ".join("%.17g" % x for x in spec.vxyz[v])))
        iu6 = [(r, c) for r in range(6) for c in range(r, 6)]
        iu3 = [(r, c) for r in range(3) for c in range(r, 3)]
        for e in range(spec.ekind.size):
            if spec.ekind[e] == 0:
                Z = spec.eZ[e]
                q = Rotation.from_matrix(Z[:, :3]).as_quat()
                if q[3] < 0:
                    q = -q
                f.write("EDGE_SE3:QUAT %d %d %s %s\n" % (spec.evi[e], spec.evj[e], " ".join("%.17g" % x for x in list(Z[:, 3]) + list(q)),
                                                       " ".join("%.17g" % spec.einfo6[r, c] for r, c in iu6)))
            else:
                f.write("EDGE_SE3_TRACKXYZ %d %d 0 %s %s\n" % (spec.evi[e], spec.evj[e], " ".join("%.17g" % x for x in spec.ez[e]),
                                                             " ".join("%.17g" % spec.einfo3[r, c] for r, c in iu3)))


write_g2o(synth.make_config_graph("cfg1"), os.path.join(out, "cfg1.g2o"))
print("g2o export done")
