"""CPU study (scipy, no GPU): how many PCG iterations does preconditioner 3 need when the persistent grid has fewer
CTAs than 148?  Fewer CTAs = larger per-CTA aggregates (a coarser coarse level) but larger exactly-coupled groups and
a 148/n times shorter Gauss-Jordan.  Input for DESIGN.md section 9, item 1.
usage: python scripts/grid_study.py [n_kf=2000] [lambda=1e-3]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import oracle
from semantic_slam_b200 import synth

n_kf = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
lam = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-3
spec = synth.make_graph(n_kf, n_kf // 5, seed=77)
o = oracle.OracleGraphSLAM()
ids = synth.load_graph(o, spec)
o.optimize(int(os.environ.get("WARM", "3")))
H, b, off = o.sparse_system()
vk = spec.vkind
pose_v = [v for v in range(vk.size) if vk[v] == 0 and off[v] >= 0]
lm_v = [v for v in range(vk.size) if vk[v] == 1]
ip = np.concatenate([np.arange(off[v], off[v] + 6) for v in pose_v])
il = np.concatenate([np.arange(off[v], off[v] + 3) for v in lm_v])
Hpp = H[ip][:, ip].tocsr(); Hpl = H[ip][:, il].tocsr(); Hll = (H[il][:, il] + lam * sp.eye(il.size)).tocsr()
n = ip.size; Np = n // 6
W = sp.block_diag([np.linalg.inv(Hll[3*k:3*k+3, 3*k:3*k+3].toarray()) for k in range(il.size // 3)]).tocsr()
S = (Hpp + lam * sp.eye(n) - Hpl @ W @ Hpl.T).tocsr()
g = b[ip] - Hpl @ (W @ b[il])
T = np.array([o.get_se3(ids[v]) for v in pose_v])

def level(gidx):
    rows, cols, vals = [], [], []
    for a in range(gidx.max() + 1):
        mem = np.flatnonzero(gidx == a)
        if mem.size == 0:
            continue
        cen = T[mem][:, :, 3].mean(0)
        for i in mem:
            R = T[i][:, :3]; d = T[i][:, 3] - cen
            Sx = np.array([[0, -d[2], d[1]], [d[2], 0, -d[0]], [-d[1], d[0], 0]])
            B = np.zeros((6, 6)); B[:3, :3] = R.T; B[:3, 3:] = -R.T @ Sx; B[3:, 3:] = 0.5 * R.T
            for rr in range(6):
                for cc in range(6):
                    rows.append(6 * i + rr); cols.append(6 * a + cc); vals.append(B[rr, cc])
    return sp.coo_matrix((vals, (rows, cols)), shape=(n, 6 * (gidx.max() + 1))).tocsr()

Dinv = sp.block_diag([np.linalg.inv(S[6*k:6*k+6, 6*k:6*k+6].toarray()) for k in range(Np)]).tocsr()
full = np.arange(Np) + 1                 # index in the full pose list (pose 0 is fixed)
P5 = level(full // 5 - (full // 5).min()); A5 = (P5.T @ S @ P5).tocsr()
agg_of_pose = full // 5

def pcg(M, tol=1e-6):
    x = np.zeros(n); r = g.copy(); z = M(r); p = z.copy(); rz = r @ z; rz0 = rz; it = 0
    while rz > tol * tol * rz0 and it < 5000:
        q = S @ p; a = rz / (p @ q); x += a * p; r -= a * q; z = M(r); rzn = r @ z; p = z + (rzn / rz) * p; rz = rzn; it += 1
    return it

print(f"{n_kf} keyframes, {il.size // 3} landmarks, lambda {lam}: S is {n} x {n}")
for nblk in (148, 74, 37, 19):
    C = max(5, ((n_kf + nblk - 1) // nblk + 4) // 5 * 5)
    if C > 80:
        print(f"grid {nblk:3d}: C = {C} poses per CTA exceeds the on-chip limit of 80")
        continue
    apc = C // 5
    # coarse level: one aggregate per CTA; groups: two per CTA (first half rounded up), as prepare() builds them
    Pc = level(full // C); Ac = (Pc.T @ S @ Pc).tocsc(); lu = spla.splu(Ac)
    a_idx = np.arange(agg_of_pose.min(), agg_of_pose.max() + 1)
    cta = a_idx // apc
    half = (a_idx - cta * apc) >= (apc + 1) // 2
    grp = 2 * cta + half
    mats = []
    a0 = agg_of_pose.min()
    order = []
    for q in np.unique(grp):
        mem = a_idx[grp == q] - a0
        sel = np.concatenate([np.arange(6 * m, 6 * m + 6) for m in mem])
        order.append(sel)
        mats.append(np.linalg.inv(A5[sel][:, sel].toarray()))
    perm = np.concatenate(order)
    G5 = sp.block_diag(mats).tocsr()
    def M(r, w=(0.5, 1.0, 2.0)):
        r5 = (P5.T @ r)[perm]
        z5 = np.zeros(P5.shape[1]); z5[perm] = G5 @ r5
        return w[0] * (Dinv @ r) + w[1] * (P5 @ z5) + w[2] * (Pc @ lu.solve(Pc.T @ r))
    t = time.time()
    it = pcg(M)
    print(f"grid {nblk:3d}: C = {C:2d} poses per CTA, groups of <= {(apc + 1) // 2} aggregates, coarse matrix {Ac.shape[0]:3d}^2 "
          f"(Gauss-Jordan steps {nblk:3d}): {it:4d} PCG iterations  [{time.time() - t:.1f}s]", flush=True)
