"""CPU look at the static work distribution of k_pcg_flow over the 148 CTAs (no GPU): per CTA the pose-landmark
incidences its poses walk, the distinct landmarks / external neighbours it stages, and the edges of the landmark parts
dealt to its warps.  Input for DESIGN.md section 9, item 3 (skew between CTAs).
usage: python scripts/balance_study.py [cfg2]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from semantic_slam_b200 import synth

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
spec = synth.make_config_graph(name)
NB = 148
vk = spec.vkind
pose_of_v = np.cumsum(vk == 0) - 1
lm_of_v = np.cumsum(vk == 1) - 1
Np, Nl = int((vk == 0).sum()), int((vk == 1).sum())
C = max(5, ((Np + NB - 1) // NB + 4) // 5 * 5)
pl = np.array([(pose_of_v[spec.evi[e]], lm_of_v[spec.evj[e]]) for e in range(spec.n_edges) if spec.ekind[e] == 1])
pp = np.array([(pose_of_v[spec.evi[e]], pose_of_v[spec.evj[e]]) for e in range(spec.n_edges) if spec.ekind[e] == 0])
cta_of_pose = np.arange(Np) // C
# pose role: incidences per CTA and the longest per-pose walk (6 lanes of a pose walk its incidences serially)
inc = np.bincount(cta_of_pose[pl[:, 0]], minlength=NB)
per_pose = np.bincount(pl[:, 0], minlength=Np)
ppdeg = np.bincount(np.concatenate([pp[:, 0], pp[:, 1]]), minlength=Np)
walk = np.array([(per_pose[b * C:(b + 1) * C] + 2 * ppdeg[b * C:(b + 1) * C]).max(initial=0) for b in range(NB)])
nuniq = np.array([np.unique(pl[cta_of_pose[pl[:, 0]] == b, 1]).size for b in range(NB)])
ext = []
for b in range(NB):
    m = (cta_of_pose[pp[:, 0]] == b) ^ (cta_of_pose[pp[:, 1]] == b)
    other = np.where(cta_of_pose[pp[m, 0]] == b, pp[m, 1], pp[m, 0])
    ext.append(np.unique(other).size)
ext = np.array(ext)
# landmark role: parts of <= 64 edges in landmark order, part q goes to warp q // NB of CTA q % NB
deg = np.bincount(pl[:, 1], minlength=Nl)
parts = []
for l in range(Nl):
    d = int(deg[l])
    while d > 0:
        parts.append(min(d, 64)); d -= 64
parts = np.array(parts)
q = np.arange(parts.size)
part_edges = np.bincount(q % NB, weights=parts, minlength=NB)
part_max = np.array([parts[q % NB == b].max(initial=0) for b in range(NB)])
# the CTAs a landmark's v goes back to (hop 2 fan-out) and the CTAs whose u it needs (hop 1 fan-in)
fan = np.array([np.unique(cta_of_pose[pl[pl[:, 1] == l, 0]]).size for l in range(Nl)])

def line(name, x):
    print(f"{name:46s} min {x.min():6.0f}  mean {x.mean():8.1f}  max {x.max():6.0f}  max/mean {x.max() / max(x.mean(), 1e-9):5.2f}")
print(f"{name}: {Np} poses, {Nl} landmarks, {pl.shape[0]} pose-landmark edges, C = {C} poses per CTA, {parts.size} landmark parts")
line("pose-landmark incidences per CTA", inc)
line("longest per-pose walk in the CTA (pl + 2 pp)", walk)
line("distinct landmarks staged per CTA (v cells x3)", nuniq)
line("external neighbour poses per CTA (u cells x6)", ext)
line("cells staged per CTA (6 ext + 3 landmarks)", 6 * ext + 3 * nuniq)
line("landmark-part edges per CTA (landmark role)", part_edges)
line("largest part of the CTA (one warp, <= 64)", part_max)
line("CTAs that see one landmark (fan-in = fan-out)", fan)
