"""CPU study (scipy, no GPU): what does cutting the preconditioner at shard boundaries cost?
One graph row-sharded over W ranks by contiguous keyframe range; every rank runs 148 CTAs of C_w poses.  Compared:
  full     : 3-level preconditioner with ONE global CTA-level coarse matrix over all 148 W aggregates (not buildable:
             the rows of its inverse do not fit on chip for W >= 2)
  local    : the same levels built from S_rr only (coarse matrix block-diagonal over ranks: no cross-rank term)
  local+g  : local + an additive global level with one rigid-body aggregate per rank (6 W unknowns)
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import oracle
from semantic_slam_b200 import synth

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
lam = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-3
spec = synth.make_config_graph(name)
o = oracle.OracleGraphSLAM()
ids = synth.load_graph(o, spec)
o.optimize(int(os.environ.get("WARM", "3")))
H, b, off = o.sparse_system()
vk = spec.vkind
pose_v = [v for v in range(vk.size) if vk[v] == 0 and off[v] >= 0]
lm_v = [v for v in range(vk.size) if vk[v] == 1]
ip = np.concatenate([np.arange(off[v], off[v] + 6) for v in pose_v])
il = np.concatenate([np.arange(off[v], off[v] + 3) for v in lm_v])
Hpp = H[ip][:, ip].tocsr(); Hpl = H[ip][:, il].tocsr(); Hll = H[il][:, il].tocsr()
n = ip.size; Np = n // 6
Hll = Hll + lam * sp.eye(il.size)
W_ = sp.block_diag([np.linalg.inv(Hll[3*k:3*k+3, 3*k:3*k+3].toarray()) for k in range(il.size // 3)]).tocsr()
S = (Hpp + lam * sp.eye(n) - Hpl @ W_ @ Hpl.T).tocsr()
g = b[ip] - Hpl @ (W_ @ b[il])
T = np.array([o.get_se3(ids[v]) for v in pose_v])
print(f"S: {n} x {n}, nnz {S.nnz}", flush=True)

def basis(members, cen):
    out = []
    for i in members:
        R = T[i][:, :3]; d = T[i][:, 3] - cen
        Sx = np.array([[0, -d[2], d[1]], [d[2], 0, -d[0]], [-d[1], d[0], 0]])
        B = np.zeros((6, 6)); B[:3, :3] = R.T; B[:3, 3:] = -R.T @ Sx; B[3:, 3:] = 0.5 * R.T
        out.append(B)
    return np.vstack(out)
def level_from_gidx(gidx):
    rows, cols, vals = [], [], []
    for a in range(gidx.max() + 1):
        mem = np.flatnonzero(gidx == a)
        if mem.size == 0: continue
        cen = T[mem][:, :, 3].mean(0)
        Bm = basis(mem, cen)
        r0 = 6 * mem[0]
        rr, cc = np.meshgrid(np.arange(Bm.shape[0]), np.arange(6), indexing="ij")
        rows.append((r0 + rr).ravel()); cols.append((6 * a + cc).ravel()); vals.append(Bm.ravel())
    return sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, 6 * (gidx.max() + 1))).tocsr()
Dinv = sp.block_diag([np.linalg.inv(S[6*k:6*k+6, 6*k:6*k+6].toarray()) for k in range(Np)]).tocsr()
def pcg(M, tol=1e-6):
    x = np.zeros(n); r = g.copy(); z = M(r); p = z.copy(); rz = r @ z; rz0 = rz; it = 0
    while rz > tol * tol * rz0 and it < 5000:
        q = S @ p; a = rz / (p @ q); x += a * p; r -= a * q; z = M(r); rzn = r @ z; p = z + (rzn / rz) * p; rz = rzn; it += 1
    return it
def group_inv(A, groups):
    """exact inverse of the diagonal blocks of A given by a list of (first aggregate, one-past-last) groups"""
    mats = []
    for g0, g1 in groups:
        mats.append(np.linalg.inv(A[6*g0:6*g1, 6*g0:6*g1].toarray()))
    return sp.block_diag(mats).tocsr()

full_np = Np + 1   # pose 0 is fixed; pose index in the full list = position + 1
for Wn in [1, 2, 4, 8]:
    nb = 148
    per_rank = (full_np + Wn - 1) // Wn          # poses per rank incl. the fixed one on rank 0
    C = max(5, ((per_rank + nb - 1) // nb + 4) // 5 * 5)
    fullidx = np.arange(Np) + 1
    rank = np.minimum(Wn - 1, fullidx // (nb * C))
    local = fullidx - rank * nb * C
    cta = rank * nb + np.minimum(nb - 1, local // C)
    agg5 = fullidx // 5
    # levels
    cta_ids = np.unique(cta, return_inverse=True)[1]
    Pc = level_from_gidx(cta_ids); Ac = (Pc.T @ S @ Pc).tocsr()
    a5_ids = np.unique(agg5, return_inverse=True)[1]
    P5 = level_from_gidx(a5_ids); A5 = (P5.T @ S @ P5).tocsr()
    # groups: the 5-pose aggregates of a CTA in two halves
    groups = []
    n5 = a5_ids.max() + 1
    cta_of_a5 = np.zeros(n5, dtype=int); cta_of_a5[a5_ids] = cta_ids
    for c in range(cta_ids.max() + 1):
        mem = np.flatnonzero(cta_of_a5 == c)
        if mem.size == 0: continue
        h = (mem.size + 1) // 2
        groups.append((mem[0], mem[0] + h))
        if mem.size > h: groups.append((mem[0] + h, mem[-1] + 1))
    G5 = group_inv(A5, groups)
    lu_full = spla.splu(Ac.tocsc())
    # rank-local coarse: zero the cross-rank blocks of Ac
    rank_of_cta = np.zeros(cta_ids.max() + 1, dtype=int); rank_of_cta[cta_ids] = rank
    rc = np.repeat(rank_of_cta, 6)
    Acoo = Ac.tocoo(); keep = rc[Acoo.row] == rc[Acoo.col]
    Aloc = sp.coo_matrix((Acoo.data[keep], (Acoo.row[keep], Acoo.col[keep])), shape=Ac.shape).tocsc()
    lu_loc = spla.splu(Aloc)
    # S_rr-based variants: also cut the cross-rank landmark terms inside the group matrices (groups never straddle ranks: identical)
    Pg = level_from_gidx(rank); Ag = (Pg.T @ S @ Pg).toarray(); Agi = np.linalg.inv(Ag)
    w = (0.5, 1.0, 2.0)
    def M_full(r): return w[0] * (Dinv @ r) + w[1] * (P5 @ (G5 @ (P5.T @ r))) + w[2] * (Pc @ lu_full.solve(Pc.T @ r))
    def M_loc(r): return w[0] * (Dinv @ r) + w[1] * (P5 @ (G5 @ (P5.T @ r))) + w[2] * (Pc @ lu_loc.solve(Pc.T @ r))
    res = {"full": pcg(M_full), "local": pcg(M_loc)}
    for wg in (1.0, 2.0, 4.0):
        def M_lg(r, wg=wg): return M_loc(r) + wg * (Pg @ (Agi @ (Pg.T @ r)))
        res[f"local+g(w={wg})"] = pcg(M_lg)
    print(f"W={Wn} C={C} ctas={cta_ids.max()+1} groups={len(groups)}: " + "  ".join(f"{k}: {v}" for k, v in res.items()), flush=True)
