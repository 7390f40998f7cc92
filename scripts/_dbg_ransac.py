import sys; sys.path.insert(0,'/root/repo')
import numpy as np, oracle
from semantic_slam_b200 import PlaneSegmentation, CloudLayout, synth
cl = synth.make_cloud(n_boxes=4, n_hyp=64, seed=104)
lay = CloudLayout(cl.width, cl.height, cl.point_step, cl.row_step, cl.offsets)
seg = PlaneSegmentation()
res, counts, mask = seg.fit_planes(cl.msg, lay, cl.boxes, cl.triples)
ores, ocounts, omask = oracle.ransac_plane_batch(cl.msg, cl.width, cl.height, cl.point_step, cl.row_step, cl.offsets, cl.boxes, cl.triples)
bad = np.argwhere(counts != ocounts)
print('n mismatch', len(bad), 'of', counts.size)
for b,k in bad[:10]:
    print(b, k, counts[b,k], ocounts[b,k], 'n_points', res['n_points'][b])
print(cl.boxes)
