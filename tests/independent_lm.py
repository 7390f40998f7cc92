"""An INDEPENDENT Levenberg-Marquardt for pose-landmark graphs, sharing no code with oracle/oracle_graph.cpp or the CUDA
back-end: numpy + scipy only (rotations through scipy.spatial.transform, Jacobians by central differences, the damped
normal equations solved by scipy.sparse.linalg.splu).  It restates only what the reference configures and what g2o
publishes: VertexSE3 / VertexPointXYZ with right-multiplied MQT increments, EdgeSE3, EdgeSE3PointXYZ (parameter offset =
identity), first vertex fixed, solver "lm_var" = Levenberg (tau 1e-5, rho with +1e-3, scale clamp [1/3, 2/3], nu doubling,
<= 10 trials) on the FULL system (graph_slam.cpp:27,67-73,104-166,199-205).  Used by tests/test_oracle_graph.py to pin the
oracle's trajectory on cfg2."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
from scipy.spatial.transform import Rotation


def _quat_xyz_w_nonneg(R):
    q = Rotation.from_matrix(R).as_quat()            # (n, 4) scalar-last
    q = q * np.where(q[:, 3:4] < 0, -1.0, 1.0)
    return q[:, :3]


def _from_mqt_rot(v):
    """rotation part of fromVectorMQT: q = (v, sqrt(1 - |v|^2)), identity when |v|^2 > 1"""
    n2 = (v * v).sum(1)
    w = np.sqrt(np.clip(1.0 - n2, 0.0, None))
    q = np.concatenate([v, w[:, None]], 1)
    q[n2 > 1.0] = (0, 0, 0, 1)
    return Rotation.from_quat(q).as_matrix()


class State:
    def __init__(self, R, t, p):
        self.R, self.t, self.p = R.copy(), t.copy(), p.copy()

    def copy(self):
        return State(self.R, self.t, self.p)


def _oplus_pose(R, t, d):
    return R @ _from_mqt_rot(d[:, 3:]), t + np.einsum("nij,nj->ni", R, d[:, :3])


def _err_pp(Ri, ti, Rj, tj, Rz, tz):
    tB = np.einsum("nji,nj->ni", Ri, tj - ti)
    et = np.einsum("nji,nj->ni", Rz, tB - tz)
    Re = np.einsum("nji,njk->nik", Rz, np.einsum("nji,njk->nik", Ri, Rj))
    return np.concatenate([et, _quat_xyz_w_nonneg(Re)], 1)


def _err_pl(Ri, ti, p, z):
    return np.einsum("nji,nj->ni", Ri, p - ti) - z


class IndependentLM:
    def __init__(self, spec):
        self.spec = spec
        vk = spec.vkind
        self.pose_of = np.full(vk.size, -1)
        self.lm_of = np.full(vk.size, -1)
        self.pose_of[vk == 0] = np.arange((vk == 0).sum())
        self.lm_of[vk == 1] = np.arange((vk == 1).sum())
        P = spec.vpose[vk == 0]
        self.x = State(P[:, :, :3], P[:, :, 3], spec.vxyz[vk == 1])
        pp, pl = spec.ekind == 0, spec.ekind == 1
        self.pp_i, self.pp_j = self.pose_of[spec.evi[pp]], self.pose_of[spec.evj[pp]]
        self.Rz, self.tz = spec.eZ[pp][:, :, :3], spec.eZ[pp][:, :, 3]
        self.pl_i, self.pl_l = self.pose_of[spec.evi[pl]], self.lm_of[spec.evj[pl]]
        self.z = spec.ez[pl]
        self.W6, self.W3 = spec.einfo6, spec.einfo3
        # unknown layout: vertex-id order, the fixed first vertex (a pose) left out
        self.off = np.full(vk.size, -1)
        o = 0
        for v in range(vk.size):
            if v == 0:
                continue
            self.off[v] = o
            o += 6 if vk[v] == 0 else 3
        self.n = o
        self.pose_vid = np.flatnonzero(vk == 0)
        self.lm_vid = np.flatnonzero(vk == 1)

    def errors(self, x):
        return (_err_pp(x.R[self.pp_i], x.t[self.pp_i], x.R[self.pp_j], x.t[self.pp_j], self.Rz, self.tz),
                _err_pl(x.R[self.pl_i], x.t[self.pl_i], x.p[self.pl_l], self.z))

    def chi2(self, x):
        e6, e3 = self.errors(x)
        return float(np.einsum("ni,ij,nj->", e6, self.W6, e6) + np.einsum("ni,ij,nj->", e3, self.W3, e3))

    def _jac(self, x, h=1e-6):
        """central differences with respect to the right-multiplied increments of the two vertices of every edge"""
        npp, npl = self.pp_i.size, self.pl_i.size
        Ji, Jj = np.zeros((npp, 6, 6)), np.zeros((npp, 6, 6))
        Jp, Jl = np.zeros((npl, 3, 6)), np.zeros((npl, 3, 3))
        for a in range(6):
            for s in (1.0, -1.0):
                d = np.zeros((1, 6))
                d[0, a] = s * h
                Ri, ti = _oplus_pose(x.R[self.pp_i], x.t[self.pp_i], np.repeat(d, npp, 0))
                Ji[:, :, a] += s * _err_pp(Ri, ti, x.R[self.pp_j], x.t[self.pp_j], self.Rz, self.tz) / (2 * h)
                Rj, tj = _oplus_pose(x.R[self.pp_j], x.t[self.pp_j], np.repeat(d, npp, 0))
                Jj[:, :, a] += s * _err_pp(x.R[self.pp_i], x.t[self.pp_i], Rj, tj, self.Rz, self.tz) / (2 * h)
                Rp, tp = _oplus_pose(x.R[self.pl_i], x.t[self.pl_i], np.repeat(d, npl, 0))
                Jp[:, :, a] += s * _err_pl(Rp, tp, x.p[self.pl_l], self.z) / (2 * h)
        for a in range(3):
            for s in (1.0, -1.0):
                d = np.zeros(3)
                d[a] = s * h
                Jl[:, :, a] += s * _err_pl(x.R[self.pl_i], x.t[self.pl_i], x.p[self.pl_l] + d, self.z) / (2 * h)
        return Ji, Jj, Jp, Jl

    def build(self, x):
        e6, e3 = self.errors(x)
        Ji, Jj, Jp, Jl = self._jac(x)
        rows, cols, vals = [], [], []
        b = np.zeros(self.n)

        def add(off_a, Ja, off_b, Jb, W, e):
            """blocks Ja' W Jb at (off_a, off_b) for all edges with both offsets free; rhs handled by the caller"""
            ok = (off_a >= 0) & (off_b >= 0)
            blk = np.einsum("nki,kl,nlj->nij", Ja[ok], W, Jb[ok])
            da, db = Ja.shape[2], Jb.shape[2]
            r = off_a[ok][:, None, None] + np.arange(da)[None, :, None] + np.zeros((1, 1, db), dtype=int)
            c = off_b[ok][:, None, None] + np.arange(db)[None, None, :] + np.zeros((1, da, 1), dtype=int)
            rows.append(r.ravel()); cols.append(c.ravel()); vals.append(blk.ravel())

        def rhs(off_a, Ja, W, e):
            ok = off_a >= 0
            g = -np.einsum("nki,kl,nl->ni", Ja[ok], W, e[ok])
            np.add.at(b, (off_a[ok][:, None] + np.arange(Ja.shape[2])[None, :]).ravel(), g.ravel())

        oi, oj = self.off[self.pose_vid[self.pp_i]], self.off[self.pose_vid[self.pp_j]]
        for (a, Ja), (c, Jc) in (((oi, Ji), (oi, Ji)), ((oi, Ji), (oj, Jj)), ((oj, Jj), (oi, Ji)), ((oj, Jj), (oj, Jj))):
            add(a, Ja, c, Jc, self.W6, e6)
        rhs(oi, Ji, self.W6, e6); rhs(oj, Jj, self.W6, e6)
        op, ol = self.off[self.pose_vid[self.pl_i]], self.off[self.lm_vid[self.pl_l]]
        for (a, Ja), (c, Jc) in (((op, Jp), (op, Jp)), ((op, Jp), (ol, Jl)), ((ol, Jl), (op, Jp)), ((ol, Jl), (ol, Jl))):
            add(a, Ja, c, Jc, self.W3, e3)
        rhs(op, Jp, self.W3, e3); rhs(ol, Jl, self.W3, e3)
        H = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(self.n, self.n)).tocsc()
        return H, b

    def apply(self, x, dx):
        y = x.copy()
        dp = np.zeros((x.R.shape[0], 6))
        free = self.off[self.pose_vid] >= 0
        idx = self.off[self.pose_vid[free]][:, None] + np.arange(6)[None, :]
        dp[free] = dx[idx]
        y.R, y.t = _oplus_pose(x.R, x.t, dp)
        il = self.off[self.lm_vid][:, None] + np.arange(3)[None, :]
        y.p = x.p + dx[il]
        return y

    def optimize(self, iterations):
        """returns rows (chi2 before, chi2 after, lambda after, rho, trials) like the oracle's history"""
        hist = []
        cur = self.chi2(self.x)
        lam, ni = 0.0, 2.0
        for it in range(iterations):
            H, b = self.build(self.x)
            if it == 0:
                lam = 1e-5 * H.diagonal().max()
            rho, q = 0.0, 0
            before = cur
            while True:
                dx = spla.splu((H + lam * sp.eye(self.n)).tocsc()).solve(b)
                cand = self.apply(self.x, dx)
                tmp = self.chi2(cand)
                rho = (cur - tmp) / (float(dx @ (lam * dx + b)) + 1e-3)
                if rho > 0 and np.isfinite(tmp):
                    alpha = min(1.0 - (2 * rho - 1) ** 3, 2.0 / 3.0)
                    lam *= max(1.0 / 3.0, alpha)
                    ni = 2.0
                    cur = tmp
                    self.x = cand
                else:
                    lam *= ni
                    ni *= 2
                q += 1
                if not (rho < 0 and q < 10):
                    break
            hist.append((before, cur, lam, rho, q))
            if q == 10 or rho == 0:
                break
        return np.array(hist)

    def poses34(self):
        return np.concatenate([self.x.R, self.x.t[:, :, None]], 2)
