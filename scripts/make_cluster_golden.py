"""Golden vectors for the dormant plane-clustering chain (row f4), produced by the REAL third-party implementations that are
importable in the build container (they do not travel to the GPU box, the vectors do):

  * cv2.kmeans (OpenCV 4.13) with the reference's arguments (plane_segmentation.cpp:525-535: TermCriteria(EPS + ITER, 10,
    0.01), 10 attempts, KMEANS_RANDOM_CENTERS) after cv2.setRNGSeed(seed)  ->  labels / centres / compactness;
  * scipy.spatial.ConvexHull (qhull, the library behind pcl::ConvexHull)    ->  hull vertex sets.

Inputs are regenerated from seeds by tests/cluster_cases.py (shared with the tests); only outputs are stored.
Run:  python scripts/make_cluster_golden.py   ->  tests/golden/cluster_kmeans_cv2.npz, tests/golden/cluster_hull_qhull.npz"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cv2
import numpy as np
import scipy.spatial

import cluster_cases

out = {}
for name, data, K, seed in cluster_cases.kmeans_cases():
    cv2.setRNGSeed(seed)
    comp, lab, cen = cv2.kmeans(data, K, None, (cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_MAX_ITER, 10, 0.01), 10,
                                cv2.KMEANS_RANDOM_CENTERS)
    out[name + "_labels"] = lab.ravel().astype(np.uint8)
    out[name + "_centers"] = cen.astype(np.float32)
    out[name + "_compactness"] = np.float64(comp)
    out[name + "_input_crc"] = np.uint32(cluster_cases.crc(data))
    print(name, data.shape, K, "compactness", comp, "cluster sizes", np.bincount(lab.ravel(), minlength=K))
out["opencv_version"] = np.array(cv2.__version__)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cluster_kmeans_cv2.npz"), **out)

hull = {}
for name, pts in cluster_cases.hull_cases():
    hv = scipy.spatial.ConvexHull(pts.astype(np.float64)).vertices
    hull[name + "_vertices"] = np.sort(hv).astype(np.int32)
    hull[name + "_input_crc"] = np.uint32(cluster_cases.crc(pts))
    print(name, pts.shape, "hull vertices", hv.size)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cluster_hull_qhull.npz"), **hull)
