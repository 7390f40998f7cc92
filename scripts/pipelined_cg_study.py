"""CPU study (scipy, no GPU): does pipelined preconditioned CG (Ghysels & Vanroose 2014) converge like the
Chronopoulos-Gear recurrence the device kernel runs, with preconditioner 3 on the cfg2 Schur complement?  In the
pipelined form the reduction of iteration i overlaps the matvec n = S m; with the coarse level this only works when
the restricted vectors P'w, P'z are carried by recurrences fed from K'm (K = S P), which this script also checks
(DESIGN.md section 9, item 3).
usage: python scripts/pipelined_cg_study.py [cfg2] [lambda=1e-3]"""
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
Hpp = H[ip][:, ip].tocsr(); Hpl = H[ip][:, il].tocsr(); Hll = (H[il][:, il] + lam * sp.eye(il.size)).tocsr()
n = ip.size; Np = n // 6
W = sp.block_diag([np.linalg.inv(Hll[3*k:3*k+3, 3*k:3*k+3].toarray()) for k in range(il.size // 3)]).tocsr()
S = (Hpp + lam * sp.eye(n) - Hpl @ W @ Hpl.T).tocsr()
g = b[ip] - Hpl @ (W @ b[il])
T = np.array([o.get_se3(ids[v]) for v in pose_v])
n_kf = Np + 1
NB = 148
C = max(5, ((n_kf + NB - 1) // NB + 4) // 5 * 5)

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
full = np.arange(Np) + 1
P5 = level(full // 5 - (full // 5).min()); A5 = (P5.T @ S @ P5).tocsr()
Pc = level(full // C); Ac = (Pc.T @ S @ Pc).tocsc(); lu = spla.splu(Ac)
apc = C // 5
a_idx = np.arange((full // 5).min(), (full // 5).max() + 1)
cta = a_idx // apc
grp = 2 * cta + ((a_idx - cta * apc) >= (apc + 1) // 2)
order, mats = [], []
for q in np.unique(grp):
    mem = a_idx[grp == q] - a_idx[0]
    sel = np.concatenate([np.arange(6 * m, 6 * m + 6) for m in mem])
    order.append(sel); mats.append(np.linalg.inv(A5[sel][:, sel].toarray()))
perm = np.concatenate(order); G5 = sp.block_diag(mats).tocsr()
K = (S @ Pc).tocsr()          # K' m = P' S m
w3 = (0.5, 1.0, 2.0)

def M_local(r):
    r5 = (P5.T @ r)[perm]
    z5 = np.zeros(P5.shape[1]); z5[perm] = G5 @ r5
    return w3[0] * (Dinv @ r) + w3[1] * (P5 @ z5)
def M(r):
    return M_local(r) + w3[2] * (Pc @ lu.solve(Pc.T @ r))

xs = spla.spsolve(S.tocsc(), g)
def report(tag, x, it, hist):
    print(f"{tag:44s} iterations {it:4d}   final |x - x*|/|x*| {np.linalg.norm(x - xs) / np.linalg.norm(xs):.2e}   true rel. residual "
          f"{np.linalg.norm(g - S @ x) / np.linalg.norm(g):.2e}", flush=True)

def cg_gear(tol, maxit=3000):
    """single-reduction PCG as in k_pcg_flow"""
    x = np.zeros(n); r = g.copy(); u = M(r); w = S @ u
    p = np.zeros(n); s = np.zeros(n); gam_old = 1.0; alpha_old = 1.0; gam0 = None
    for it in range(maxit):
        gam = r @ u; dlt = w @ u
        if gam0 is None: gam0 = gam
        if gam <= tol * tol * gam0: return x, it
        beta = 0.0 if it == 0 else gam / gam_old
        alpha = gam / (dlt - beta * gam / alpha_old) if it else gam / dlt
        p = u + beta * p; s = w + beta * s; x = x + alpha * p; r = r - alpha * s
        u = M(r); w = S @ u
        gam_old, alpha_old = gam, alpha
    return x, maxit

def cg_pipelined(tol, maxit=3000, coarse_by_recurrence=True):
    """Ghysels-Vanroose pipelined PCG; the coarse part of m = M^-1 w uses cw = P'w carried by
    cw -= alpha * cz, cz = K'm + beta * cz (no restriction of w itself)"""
    x = np.zeros(n); r = g.copy(); u = M(r); w = S @ u
    z = np.zeros(n); q = np.zeros(n); s = np.zeros(n); p = np.zeros(n)
    cw = Pc.T @ w; cz = np.zeros(Pc.shape[1])
    gam_old = alpha_old = 1.0; gam0 = None
    for it in range(maxit):
        gam = r @ u; dlt = w @ u
        if gam0 is None: gam0 = gam
        if gam <= tol * tol * gam0: return x, it
        m = M_local(w) + w3[2] * (Pc @ lu.solve(cw if coarse_by_recurrence else Pc.T @ w))
        nn = S @ m
        km = K.T @ m                              # = P' nn, available before nn
        beta = 0.0 if it == 0 else gam / gam_old
        alpha = gam / (dlt - beta * gam / alpha_old) if it else gam / dlt
        z = nn + beta * z; q = m + beta * q; s = w + beta * s; p = u + beta * p
        cz = km + beta * cz
        x = x + alpha * p; r = r - alpha * s; u = u - alpha * q; w = w - alpha * z
        cw = cw - alpha * cz
        gam_old, alpha_old = gam, alpha
    return x, maxit

print(f"{name}: S is {n} x {n}, lambda {lam}, C = {C}")
for tol in (1e-6, 1e-8, 1e-10):
    t = time.time(); x, it = cg_gear(tol); report(f"single-reduction PCG, tol {tol:g}", x, it, None)
    x, it = cg_pipelined(tol, coarse_by_recurrence=False); report(f"pipelined PCG, tol {tol:g}", x, it, None)
    x, it = cg_pipelined(tol); report(f"pipelined PCG + restricted recurrences, tol {tol:g}", x, it, None)
