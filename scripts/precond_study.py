"""CPU study of additive multilevel preconditioners for the Schur complement of cfg2 (scipy; no GPU).
Counts PCG iterations for variants of the aggregate hierarchy to decide what is worth building on the device."""
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
o.optimize(int(os.environ.get("WARM", "3")))            # linearise a few iterations in
H, b, off = o.sparse_system()
vk = spec.vkind
pose_v = [v for v in range(vk.size) if vk[v] == 0 and off[v] >= 0]
lm_v = [v for v in range(vk.size) if vk[v] == 1]
ip = np.concatenate([np.arange(off[v], off[v] + 6) for v in pose_v])
il = np.concatenate([np.arange(off[v], off[v] + 3) for v in lm_v])
Hpp = H[ip][:, ip].tocsr(); Hpl = H[ip][:, il].tocsr(); Hll = H[il][:, il].tocsr()
n = ip.size; Np = n // 6
Hll = Hll + lam * sp.eye(il.size)
# block-diagonal inverse of Hll (3x3 blocks)
blocks = Hll.toarray() if il.size < 4000 else None
W = sp.block_diag([np.linalg.inv(Hll[3*k:3*k+3, 3*k:3*k+3].toarray()) for k in range(il.size // 3)]).tocsr()
t = time.time()
S = (Hpp + lam * sp.eye(n) - Hpl @ W @ Hpl.T).tocsr()
g = b[ip] - Hpl @ (W @ b[il])
print(f"S: {n} x {n}, nnz {S.nnz}, built in {time.time()-t:.1f}s", flush=True)
# poses (non-fixed) in order; pose index in the full pose list = position + 1 (pose 0 is fixed)
T = np.array([o.get_se3(ids[v]) for v in pose_v])
def basis(members, cen):
    """rows of P for the poses `members` (indices into pose list) for one aggregate with centroid cen"""
    out = []
    for i in members:
        R = T[i][:, :3]; d = T[i][:, 3] - cen
        Sx = np.array([[0, -d[2], d[1]], [d[2], 0, -d[0]], [-d[1], d[0], 0]])
        B = np.zeros((6, 6)); B[:3, :3] = R.T; B[:3, 3:] = -R.T @ Sx; B[3:, 3:] = 0.5 * R.T
        out.append(B)
    return np.vstack(out)
def level(size, shift=1):
    """aggregates of `size` consecutive poses (in the full numbering, pose 0 fixed => shift)"""
    rows, cols, vals = [], [], []
    gidx = (np.arange(Np) + shift) // size
    gidx -= gidx.min()
    for a in range(gidx.max() + 1):
        mem = np.flatnonzero(gidx == a)
        if mem.size == 0: continue
        cen = T[mem][:, :, 3].mean(0)
        Bm = basis(mem, cen)
        r0 = 6 * mem[0]
        for rr in range(Bm.shape[0]):
            for cc in range(6):
                rows.append(r0 + rr); cols.append(6 * a + cc); vals.append(Bm[rr, cc])
    return sp.coo_matrix((vals, (rows, cols)), shape=(n, 6 * (gidx.max() + 1))).tocsr()
Dinv = sp.block_diag([np.linalg.inv(S[6*k:6*k+6, 6*k:6*k+6].toarray()) for k in range(Np)]).tocsr()
def blockdiag_inv(A, bs=6):
    return sp.block_diag([np.linalg.inv(A[bs*k:bs*k+bs, bs*k:bs*k+bs].toarray()) for k in range(A.shape[0] // bs)]).tocsr()
C = ((Np + 1 + 147) // 148 + 4) // 5 * 5
levels = {}
def get(size):
    if size not in levels:
        P = level(size); A = (P.T @ S @ P).tocsr(); levels[size] = (P, A)
    return levels[size]
def run(desc, diag_levels, full_levels, tol=1e-6):
    ops = []
    for sz in diag_levels:
        P, A = get(sz); ops.append((P, blockdiag_inv(A)))
    for sz in full_levels:
        P, A = get(sz); lu = spla.splu(A.tocsc()); ops.append((P, lu))
    def M(r):
        z = Dinv @ r
        for P, Ai in ops:
            rc = P.T @ r
            z = z + P @ (Ai.solve(rc) if hasattr(Ai, "solve") else Ai @ rc)
        return z
    x = np.zeros(n); r = g.copy(); z = M(r); p = z.copy(); rz = r @ z; rz0 = rz; it = 0
    while rz > tol * tol * rz0 and it < 5000:
        q = S @ p; a = rz / (p @ q); x += a * p; r -= a * q; z = M(r); rzn = r @ z; p = z + (rzn / rz) * p; rz = rzn; it += 1
    print(f"{desc:58s} iterations {it}", flush=True)
    return it
print(f"poses per CTA aggregate C = {C}, lambda = {lam}")
def blockdiag_inv_groups(A, group):
    """exact inverse of the diagonal blocks of A formed by consecutive groups of `group` 6x6 aggregates"""
    nb = A.shape[0] // 6
    mats = []
    for g0 in range(0, nb, group):
        g1 = min(nb, g0 + group)
        mats.append(np.linalg.inv(A[6*g0:6*g1, 6*g0:6*g1].toarray()))
    return sp.block_diag(mats).tocsr()
def run2(desc, ops_spec, tol=1e-6):
    ops = []
    for kind, sz, grp in ops_spec:
        P, A = get(sz)
        if kind == "diag": ops.append((P, blockdiag_inv(A)))
        elif kind == "group": ops.append((P, blockdiag_inv_groups(A, grp)))
        else: ops.append((P, spla.splu(A.tocsc())))
    def M(r):
        z = Dinv @ r
        for P, Ai in ops:
            rc = P.T @ r
            z = z + P @ (Ai.solve(rc) if hasattr(Ai, "solve") else Ai @ rc)
        return z
    x = np.zeros(n); r = g.copy(); z = M(r); p = z.copy(); rz = r @ z; rz0 = rz; it = 0
    while rz > tol * tol * rz0 and it < 5000:
        q = S @ p; a = rz / (p @ q); x += a * p; r -= a * q; z = M(r); rzn = r @ z; p = z + (rzn / rz) * p; rz = rzn; it += 1
    print(f"{desc:70s} iterations {it}", flush=True)
def run3(desc, w, tol=1e-6):
    P5, A5 = get(5); G5 = blockdiag_inv_groups(A5, 7)
    Pc, Ac = get(C); lu = spla.splu(Ac.tocsc())
    def M(r):
        return w[0] * (Dinv @ r) + w[1] * (P5 @ (G5 @ (P5.T @ r))) + w[2] * (Pc @ lu.solve(Pc.T @ r))
    x = np.zeros(n); r = g.copy(); z = M(r); p = z.copy(); rz = r @ z; rz0 = rz; it = 0
    while rz > tol * tol * rz0 and it < 5000:
        q = S @ p; a = rz / (p @ q); x += a * p; r -= a * q; z = M(r); rzn = r @ z; p = z + (rzn / rz) * p; rz = rzn; it += 1
    print(f"{desc:50s} w={w} iterations {it}", flush=True)
for w in [(1, 1, 1), (1, 1, 2), (0.5, 1, 2), (0.5, 1, 3)]:
    run3("weighted additive", w)
