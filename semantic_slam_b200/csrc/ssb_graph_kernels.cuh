// ssb_graph_kernels.cuh — sm_100a kernels of the graph hot path.
//
// Data layout in HBM (all fp64):
//   pose[Np]   : 64 B records (t, q, pad)            lm[Nl] : 32 B records (xyz, pad)
//   pl[El]     : 80 B pose->landmark edge records, sorted landmark-major ("L-order"), CSR lm_rowptr
//   pp[Epp]    : 240 B pose->pose edge records (creation order)
//   pose_pl_rowptr/pose_pl_idx : pose-major CSR over the L-order edge positions ("P-order")
//   pose_pp_rowptr/pose_pp_idx : per pose incident pose-pose edges, (edge << 1) | role
//   Hpp[Np][36], bp[Np][6], Hoff[Epp][36] (= Ji' W Jj), Hll[Nl][6] (upper), bl[Nl][3],
//   HplL[El][18] (3x6, L-order: w_l = sum HplL_e p), HplP[El][18] (6x3, P-order), plP_lm[El]
//   per damped trial: HllInv[Nl][6], Dinv[Np][36] (block-Jacobi of the Schur complement), g[Np][6]
#pragma once
#include <cooperative_groups.h>
#include "ssb_common.cuh"
#include "ssb_math.cuh"

namespace ssb {
namespace cg = cooperative_groups;

struct DevGraph {
  int Np, Nl, El, Epp;
  Pose* pose;
  double* lm;  // 4 doubles per landmark
  const unsigned char* pose_fixed;
  const unsigned char* lm_fixed;
  const PLEdge* pl;
  const PPEdge* pp;
  const int* lm_rowptr;
  const int* pose_pl_rowptr;
  const int* pose_pl_idx;
  const int* pose_pp_rowptr;
  const int* pose_pp_idx;
  double *Hpp, *bp, *Hoff, *Hll, *bl, *HplL, *HplP;
  int* plP_lm;
  double *HllInv, *Dinv, *g;
  // PCG vectors
  double *x, *r, *z, *p0, *p1, *q, *v;
  double* dl;  // landmark increments
  // reductions
  double* part;     // 3 * PART_STRIDE partials
  double* scalars;  // [0] chi2, [1] scale, [2] maxdiag (as bits), [3] pcg rz final, [4] rz0
  int* iscalars;    // [0] pcg iters, [1] pcg status, [2] ticket chi2, [3] ticket scale
};
constexpr int PART_STRIDE = 1024;

// ---------------------------------------------------------------------------------------------
// K1a: landmark-major linearisation.  One thread per landmark walks its L-order edges:
//   Hll += R W R',  bl += -R W e,  HplL_e = R W Jp  (3x6)
// (BlockSolver::buildSystem -> EdgeSE3PointXYZ::linearizeOplus + constructQuadraticForm)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_lin_landmarks(DevGraph G) {
  int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= G.Nl) return;
  const bool lfixed = G.lm_fixed[l] != 0;
  double p[3] = {G.lm[4 * l], G.lm[4 * l + 1], G.lm[4 * l + 2]};
  double H[6] = {0, 0, 0, 0, 0, 0}, b[3] = {0, 0, 0};
  for (int e = G.lm_rowptr[l]; e < G.lm_rowptr[l + 1]; ++e) {
    const PLEdge ed = G.pl[e];
    const Pose X = G.pose[ed.p];
    PLLin L;
    pl_linearize(X, p, ed.z, L);
    double W[9];
    expand_sym3(ed.info, W);
    // RW = R * W (3x3)
    double RW[9];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) RW[3 * r + c] = L.R[3 * r] * W[c] + L.R[3 * r + 1] * W[3 + c] + L.R[3 * r + 2] * W[6 + c];
    if (!lfixed) {
      // Hll += RW * R'
      int k = 0;
      for (int r = 0; r < 3; ++r)
        for (int c = r; c < 3; ++c) {
          H[k] += RW[3 * r] * L.R[3 * c] + RW[3 * r + 1] * L.R[3 * c + 1] + RW[3 * r + 2] * L.R[3 * c + 2];
          ++k;
        }
      for (int r = 0; r < 3; ++r) b[r] -= RW[3 * r] * L.e[0] + RW[3 * r + 1] * L.e[1] + RW[3 * r + 2] * L.e[2];
    }
    double Jp[18];
    pl_jac_pose(L.pc, Jp);
    double* out = G.HplL + 18 * (size_t)e;
    const bool zero = lfixed || G.pose_fixed[ed.p];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 6; ++c)
        out[6 * r + c] = zero ? 0.0 : RW[3 * r] * Jp[c] + RW[3 * r + 1] * Jp[6 + c] + RW[3 * r + 2] * Jp[12 + c];
  }
  if (lfixed) {
    H[0] = H[3] = H[5] = 1.0;
    H[1] = H[2] = H[4] = 0.0;
  }
  for (int k = 0; k < 6; ++k) G.Hll[6 * (size_t)l + k] = H[k];
  for (int k = 0; k < 3; ++k) G.bl[3 * (size_t)l + k] = b[k];
}

// ---------------------------------------------------------------------------------------------
// K1b: pose-major linearisation.  One thread per pose accumulates its diagonal block and rhs over
// its pose-landmark edges (P-order) and its pose-pose edges; the thread that owns role i of a
// pose-pose edge writes the off-diagonal block Hoff_e = Ji' W Jj.  Deterministic (no atomics).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) k_lin_poses(DevGraph G) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= G.Np) return;
  const bool fixed = G.pose_fixed[i] != 0;
  const Pose X = G.pose[i];
  double H[36], b[6];
  for (int k = 0; k < 36; ++k) H[k] = 0.0;
  for (int k = 0; k < 6; ++k) b[k] = 0.0;
  // pose-landmark edges
  for (int kk = G.pose_pl_rowptr[i]; kk < G.pose_pl_rowptr[i + 1]; ++kk) {
    const int e = G.pose_pl_idx[kk];
    const PLEdge ed = G.pl[e];
    double p[3] = {G.lm[4 * ed.l], G.lm[4 * ed.l + 1], G.lm[4 * ed.l + 2]};
    PLLin L;
    pl_linearize(X, p, ed.z, L);
    double W[9], Jp[18];
    expand_sym3(ed.info, W);
    pl_jac_pose(L.pc, Jp);
    // JtW = Jp' W (6x3)
    double JtW[18];
    for (int r = 0; r < 6; ++r)
      for (int c = 0; c < 3; ++c) JtW[3 * r + c] = Jp[r] * W[c] + Jp[6 + r] * W[3 + c] + Jp[12 + r] * W[6 + c];
    if (!fixed) {
      for (int r = 0; r < 6; ++r) {
        for (int c = 0; c < 6; ++c) H[6 * r + c] += JtW[3 * r] * Jp[c] + JtW[3 * r + 1] * Jp[6 + c] + JtW[3 * r + 2] * Jp[12 + c];
        b[r] -= JtW[3 * r] * L.e[0] + JtW[3 * r + 1] * L.e[1] + JtW[3 * r + 2] * L.e[2];
      }
    }
    // HplP = Jp' W R'  (6x3)
    const bool zero = fixed || G.lm_fixed[ed.l];
    double* out = G.HplP + 18 * (size_t)kk;
    for (int r = 0; r < 6; ++r)
      for (int c = 0; c < 3; ++c)
        out[3 * r + c] = zero ? 0.0 : JtW[3 * r] * L.R[3 * c] + JtW[3 * r + 1] * L.R[3 * c + 1] + JtW[3 * r + 2] * L.R[3 * c + 2];
    G.plP_lm[kk] = ed.l;
  }
  // pose-pose edges
  for (int kk = G.pose_pp_rowptr[i]; kk < G.pose_pp_rowptr[i + 1]; ++kk) {
    const int code = G.pose_pp_idx[kk];
    const int e = code >> 1, role = code & 1;
    const PPEdge* ed = G.pp + e;
    const int other = role == 0 ? ed->j : ed->i;
    const Pose Y = G.pose[other];
    double err[6], Ji[36], Jj[36], W[36];
    if (role == 0)
      pp_linearize(X, Y, ed->zt, ed->zq, err, Ji, Jj);
    else
      pp_linearize(Y, X, ed->zt, ed->zq, err, Ji, Jj);
    expand_sym6(ed->info, W);
    const double* J = role == 0 ? Ji : Jj;
    double JtW[36];
    for (int r = 0; r < 6; ++r)
      for (int c = 0; c < 6; ++c) {
        double s = 0.0;
        for (int k = 0; k < 6; ++k) s += J[6 * k + r] * W[6 * k + c];
        JtW[6 * r + c] = s;
      }
    if (!fixed) {
      for (int r = 0; r < 6; ++r) {
        for (int c = 0; c < 6; ++c) {
          double s = 0.0;
          for (int k = 0; k < 6; ++k) s += JtW[6 * r + k] * J[6 * k + c];
          H[6 * r + c] += s;
        }
        double s = 0.0;
        for (int k = 0; k < 6; ++k) s += JtW[6 * r + k] * err[k];
        b[r] -= s;
      }
    }
    if (role == 0) {
      const bool zero = fixed || G.pose_fixed[other];
      double* out = G.Hoff + 36 * (size_t)e;
      for (int r = 0; r < 6; ++r)
        for (int c = 0; c < 6; ++c) {
          double s = 0.0;
          for (int k = 0; k < 6; ++k) s += JtW[6 * r + k] * Jj[6 * k + c];
          out[6 * r + c] = zero ? 0.0 : s;
        }
    }
  }
  if (fixed) {
    for (int k = 0; k < 36; ++k) H[k] = 0.0;
    for (int k = 0; k < 6; ++k) {
      H[7 * k] = 1.0;
      b[k] = 0.0;
    }
  }
  for (int k = 0; k < 36; ++k) G.Hpp[36 * (size_t)i + k] = H[k];
  for (int k = 0; k < 6; ++k) G.bp[6 * (size_t)i + k] = b[k];
}

// max |H_jj| over the non-fixed vertices (OptimizationAlgorithmLevenberg::computeLambdaInit)
__global__ void k_maxdiag(DevGraph G) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  double m = 0.0;
  if (t < G.Np) {
    if (!G.pose_fixed[t])
      for (int k = 0; k < 6; ++k) m = fmax(m, fabs(G.Hpp[36 * (size_t)t + 7 * k]));
  } else if (t < G.Np + G.Nl) {
    int l = t - G.Np;
    if (!G.lm_fixed[l]) {
      m = fmax(m, fabs(G.Hll[6 * (size_t)l + 0]));
      m = fmax(m, fabs(G.Hll[6 * (size_t)l + 3]));
      m = fmax(m, fabs(G.Hll[6 * (size_t)l + 5]));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.0)
    atomicMax((unsigned long long*)(G.scalars + 2), (unsigned long long)__double_as_longlong(m));
}

// ---------------------------------------------------------------------------------------------
// per damped trial: (Hll + lambda I)^-1, block-Jacobi preconditioner of S, reduced rhs
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_prep_landmarks(DevGraph G, double lambda) {
  int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= G.Nl) return;
  double a[6], o[6];
  for (int k = 0; k < 6; ++k) a[k] = G.Hll[6 * (size_t)l + k];
  a[0] += lambda;
  a[3] += lambda;
  a[5] += lambda;
  if (!inv_sym3(a, o))
    for (int k = 0; k < 6; ++k) o[k] = 0.0;
  for (int k = 0; k < 6; ++k) G.HllInv[6 * (size_t)l + k] = o[k];
}

__global__ void __launch_bounds__(64) k_prep_poses(DevGraph G, double lambda) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= G.Np) return;
  double D[36], g[6];
  for (int k = 0; k < 36; ++k) D[k] = G.Hpp[36 * (size_t)i + k];
  for (int k = 0; k < 6; ++k) {
    D[7 * k] += lambda;
    g[k] = G.bp[6 * (size_t)i + k];
  }
  for (int kk = G.pose_pl_rowptr[i]; kk < G.pose_pl_rowptr[i + 1]; ++kk) {
    const double* Hp = G.HplP + 18 * (size_t)kk;
    const int l = G.plP_lm[kk];
    double Wi[9];
    expand_sym3(G.HllInv + 6 * (size_t)l, Wi);
    const double* bl = G.bl + 3 * (size_t)l;
    double T[18];  // Hp * Wi (6x3)
    for (int r = 0; r < 6; ++r)
      for (int c = 0; c < 3; ++c) T[3 * r + c] = Hp[3 * r] * Wi[c] + Hp[3 * r + 1] * Wi[3 + c] + Hp[3 * r + 2] * Wi[6 + c];
    for (int r = 0; r < 6; ++r) {
      for (int c = 0; c < 6; ++c) D[6 * r + c] -= T[3 * r] * Hp[3 * c] + T[3 * r + 1] * Hp[3 * c + 1] + T[3 * r + 2] * Hp[3 * c + 2];
      g[r] -= T[3 * r] * bl[0] + T[3 * r + 1] * bl[1] + T[3 * r + 2] * bl[2];
    }
  }
  if (!inv_spd6(D)) {
    // not positive definite: fall back to the inverse of the diagonal (PCG will report breakdown)
    for (int r = 0; r < 6; ++r)
      for (int c = 0; c < 6; ++c) D[6 * r + c] = 0.0;
    for (int k = 0; k < 6; ++k) D[7 * k] = 1.0;
  }
  for (int k = 0; k < 36; ++k) G.Dinv[36 * (size_t)i + k] = D[k];
  for (int k = 0; k < 6; ++k) G.g[6 * (size_t)i + k] = g[k];
}

// ---------------------------------------------------------------------------------------------
// K3: persistent cooperative block-Jacobi PCG on the (implicit) Schur complement
//   S x = g,  S = Hpp + lambda I - Hpl (Hll + lambda I)^-1 Hlp
// 3 grid-wide syncs per iteration.  6 lanes per pose (one per row), 5 poses per warp; one warp per
// landmark in the landmark sweep.  All reductions are fixed-order => bit-reproducible.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double grid_total(const double* part, int nblk, double* sh) {
  // every block sums the same partials in the same order
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (w == 0) {
    double t = 0.0;
    for (int k = lane; k < nblk; k += 32) t += __ldcg(part + k);
    t = warp_sum(t);
    if (lane == 0) sh[32] = t;
  }
  __syncthreads();
  double r = sh[32];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(1024, 1) k_pcg(DevGraph G, double lambda, double tol2, int maxit) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double sh[33];
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int gw = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int total_warps = gridDim.x * warps_per_block;
  const int slot = lane / 6, comp = lane - 6 * slot;
  const bool lane_active = lane < 30;
  const int base_lane = 6 * slot;
  const int nblk = gridDim.x;
  double* partA = G.part;
  double* partB = G.part + PART_STRIDE;
  double* partC = G.part + 2 * PART_STRIDE;
  double* pold = G.p0;
  double* pnew = G.p1;

  // ---- init: x = 0, r = g, z = Dinv r, pold = 0
  double local = 0.0;
  for (int pbase = gw * 5; pbase < G.Np; pbase += total_warps * 5) {
    const int i = pbase + slot;
    const bool act = lane_active && i < G.Np;
    double rc = act ? G.g[6 * (size_t)i + comp] : 0.0;
    double zc = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      double rk = __shfl_sync(0xffffffffu, rc, base_lane + k);
      if (act) zc += G.Dinv[36 * (size_t)i + 6 * comp + k] * rk;
    }
    if (act) {
      G.x[6 * (size_t)i + comp] = 0.0;
      G.r[6 * (size_t)i + comp] = rc;
      G.z[6 * (size_t)i + comp] = zc;
      pold[6 * (size_t)i + comp] = 0.0;
      local += rc * zc;
    }
  }
  double bs = block_sum(local, sh);
  if (threadIdx.x == 0) partA[blockIdx.x] = bs;
  grid.sync();
  double rz = grid_total(partA, nblk, sh);
  const double rz0 = rz;
  double beta = 0.0;
  int it = 0;
  int status = 0;
  if (!(rz0 > 0.0)) {
    status = (rz0 == 0.0) ? 0 : 2;
    maxit = 0;
  }
  for (it = 0; it < maxit; ++it) {
    // ---- phase 1: v_l = (Hll+lambda)^-1 sum_e HplL_e p_e,  p = z + beta * pold (on the fly)
    for (int l = gw; l < G.Nl; l += total_warps) {
      double a0 = 0.0, a1 = 0.0, a2 = 0.0;
      const int e1 = G.lm_rowptr[l + 1];
      for (int e = G.lm_rowptr[l] + lane; e < e1; e += 32) {
        const int pi = G.pl[e].p;
        const double* Hl = G.HplL + 18 * (size_t)e;
        const double* zz = G.z + 6 * (size_t)pi;
        const double* po = pold + 6 * (size_t)pi;
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          double pc = __ldcg(zz + c) + beta * __ldcg(po + c);
          a0 += Hl[c] * pc;
          a1 += Hl[6 + c] * pc;
          a2 += Hl[12 + c] * pc;
        }
      }
      a0 = warp_sum(a0);
      a1 = warp_sum(a1);
      a2 = warp_sum(a2);
      if (lane == 0) {
        const double* Wi = G.HllInv + 6 * (size_t)l;
        G.v[3 * (size_t)l + 0] = Wi[0] * a0 + Wi[1] * a1 + Wi[2] * a2;
        G.v[3 * (size_t)l + 1] = Wi[1] * a0 + Wi[3] * a1 + Wi[4] * a2;
        G.v[3 * (size_t)l + 2] = Wi[2] * a0 + Wi[4] * a1 + Wi[5] * a2;
      }
    }
    grid.sync();
    // ---- phase 2: q = (Hpp + lambda) p + sum Hoff p_nbr - sum HplP v ; partial p.q
    local = 0.0;
    for (int pbase = gw * 5; pbase < G.Np; pbase += total_warps * 5) {
      const int i = pbase + slot;
      const bool act = lane_active && i < G.Np;
      double pc = 0.0;
      if (act) {
        pc = __ldcg(G.z + 6 * (size_t)i + comp) + beta * __ldcg(pold + 6 * (size_t)i + comp);
        pnew[6 * (size_t)i + comp] = pc;
      }
      double qc = lambda * pc;
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        double pk = __shfl_sync(0xffffffffu, pc, base_lane + k);
        if (act) qc += G.Hpp[36 * (size_t)i + 6 * comp + k] * pk;
      }
      // pose-pose neighbours (loop bounds are uniform inside a slot; slots may differ -> use ballot-free
      // formulation: every lane iterates to the warp maximum and masks)
      int k0 = 0, k1 = 0;
      if (act) {
        k0 = G.pose_pp_rowptr[i];
        k1 = G.pose_pp_rowptr[i + 1];
      }
      int nmax = k1 - k0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));
      for (int s = 0; s < nmax; ++s) {
        const bool has = act && (k0 + s < k1);
        int e = 0, role = 0, other = 0;
        double oc = 0.0;
        if (has) {
          const int code = G.pose_pp_idx[k0 + s];
          e = code >> 1;
          role = code & 1;
          other = role == 0 ? G.pp[e].j : G.pp[e].i;
          oc = __ldcg(G.z + 6 * (size_t)other + comp) + beta * __ldcg(pold + 6 * (size_t)other + comp);
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          double ok = __shfl_sync(0xffffffffu, oc, base_lane + k);
          if (has) {
            const double* Ho = G.Hoff + 36 * (size_t)e;
            qc += (role == 0 ? Ho[6 * comp + k] : Ho[6 * k + comp]) * ok;
          }
        }
      }
      if (act) {
        const int p0 = G.pose_pl_rowptr[i], p1 = G.pose_pl_rowptr[i + 1];
        for (int kk = p0; kk < p1; ++kk) {
          const double* Hp = G.HplP + 18 * (size_t)kk + 3 * comp;
          const double* vv = G.v + 3 * (size_t)G.plP_lm[kk];
          qc -= Hp[0] * __ldcg(vv) + Hp[1] * __ldcg(vv + 1) + Hp[2] * __ldcg(vv + 2);
        }
        G.q[6 * (size_t)i + comp] = qc;
        local += pc * qc;
      }
    }
    bs = block_sum(local, sh);
    if (threadIdx.x == 0) partB[blockIdx.x] = bs;
    grid.sync();
    const double pq = grid_total(partB, nblk, sh);
    if (!(pq > 0.0) || !isfinite(pq)) {  // breakdown: S not positive definite / non-finite data
      status = 1;
      break;
    }
    const double alpha = rz / pq;
    // ---- phase 3: x += alpha p ; r -= alpha q ; z = Dinv r ; partial r.z
    local = 0.0;
    for (int pbase = gw * 5; pbase < G.Np; pbase += total_warps * 5) {
      const int i = pbase + slot;
      const bool act = lane_active && i < G.Np;
      double rc = 0.0;
      if (act) {
        const size_t o = 6 * (size_t)i + comp;
        G.x[o] += alpha * pnew[o];
        rc = G.r[o] - alpha * G.q[o];
        G.r[o] = rc;
      }
      double zc = 0.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        double rk = __shfl_sync(0xffffffffu, rc, base_lane + k);
        if (act) zc += G.Dinv[36 * (size_t)i + 6 * comp + k] * rk;
      }
      if (act) {
        G.z[6 * (size_t)i + comp] = zc;
        local += rc * zc;
      }
    }
    bs = block_sum(local, sh);
    if (threadIdx.x == 0) partC[blockIdx.x] = bs;
    grid.sync();
    const double rzn = grid_total(partC, nblk, sh);
    beta = rzn / rz;
    rz = rzn;
    double* t = pold;
    pold = pnew;
    pnew = t;
    if (!(rz > tol2 * rz0)) {
      ++it;
      break;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    G.iscalars[0] = it;
    G.iscalars[1] = status;
    G.scalars[3] = rz;
    G.scalars[4] = rz0;
  }
}

// ---------------------------------------------------------------------------------------------
// K4: back-substitution + state update (+ backup for LM reject) + computeScale partials
//   dl = (Hll+lambda)^-1 (bl - sum_e HplL_e dp),  l += dl ;  X <- X * fromVectorMQT(dp)
//   scale = sum_j d_j (lambda d_j + b_j)      (OptimizationAlgorithmLevenberg::computeScale)
// Thread t < Nl handles a landmark, t >= Nl a pose.  The last block folds the partials in order.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_backsub_update(DevGraph G, double lambda, Pose* pose_bak, double* lm_bak) {
  __shared__ double sh[33];
  __shared__ int is_last;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  double sc = 0.0;
  if (t < G.Nl) {
    const int l = t;
    double a[3] = {G.bl[3 * (size_t)l], G.bl[3 * (size_t)l + 1], G.bl[3 * (size_t)l + 2]};
    for (int e = G.lm_rowptr[l]; e < G.lm_rowptr[l + 1]; ++e) {
      const double* Hl = G.HplL + 18 * (size_t)e;
      const double* dp = G.x + 6 * (size_t)G.pl[e].p;
      for (int c = 0; c < 6; ++c) {
        a[0] -= Hl[c] * dp[c];
        a[1] -= Hl[6 + c] * dp[c];
        a[2] -= Hl[12 + c] * dp[c];
      }
    }
    const double* Wi = G.HllInv + 6 * (size_t)l;
    double d[3] = {Wi[0] * a[0] + Wi[1] * a[1] + Wi[2] * a[2], Wi[1] * a[0] + Wi[3] * a[1] + Wi[4] * a[2],
                   Wi[2] * a[0] + Wi[4] * a[1] + Wi[5] * a[2]};
    if (G.lm_fixed[l]) d[0] = d[1] = d[2] = 0.0;
    for (int c = 0; c < 3; ++c) {
      const double old = G.lm[4 * (size_t)l + c];
      lm_bak[4 * (size_t)l + c] = old;
      G.lm[4 * (size_t)l + c] = old + d[c];
      G.dl[3 * (size_t)l + c] = d[c];
      sc += d[c] * (lambda * d[c] + G.bl[3 * (size_t)l + c]);
    }
  } else if (t < G.Nl + G.Np) {
    const int i = t - G.Nl;
    Pose X = G.pose[i];
    pose_bak[i] = X;
    double d[6];
    for (int c = 0; c < 6; ++c) d[c] = G.pose_fixed[i] ? 0.0 : G.x[6 * (size_t)i + c];
    if (!G.pose_fixed[i]) {
      pose_oplus(X, d);
      G.pose[i] = X;
    }
    for (int c = 0; c < 6; ++c) sc += d[c] * (lambda * d[c] + G.bp[6 * (size_t)i + c]);
  }
  double bs = block_sum(sc, sh);
  if (threadIdx.x == 0) {
    G.part[blockIdx.x] = bs;
    __threadfence();
    int ticket = atomicAdd(G.iscalars + 3, 1);
    is_last = (ticket == (int)gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    double s = 0.0;
    for (int k = threadIdx.x; k < (int)gridDim.x; k += blockDim.x) s += __ldcg(G.part + k);
    s = block_sum(s, sh);
    if (threadIdx.x == 0) {
      G.scalars[1] = s;
      G.iscalars[3] = 0;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K2: chi2 = sum_e e' W e.  Edge records are staged into shared memory tile by tile with a 1-D TMA
// bulk copy (cp.async.bulk + mbarrier); one thread per edge; fixed-order reduction.
// Tiles [0, nPL) cover pl (256 records each), tiles [nPL, nPL+nPP) cover pp (64 records each).
// ---------------------------------------------------------------------------------------------
constexpr int CHI2_THREADS = 256;
constexpr int CHI2_PL_TILE = 256;  // 20 KB
constexpr int CHI2_PP_TILE = 64;   // 15 KB

__global__ void __launch_bounds__(CHI2_THREADS) k_chi2(DevGraph G, double* part, double* out, int* ticket_ctr) {
  __shared__ __align__(128) unsigned char tile[CHI2_PL_TILE * sizeof(PLEdge)];
  __shared__ __align__(8) uint64_t bar;
  __shared__ double sh[33];
  __shared__ int is_last;
  const int nPL = (G.El + CHI2_PL_TILE - 1) / CHI2_PL_TILE;
  const int nPP = (G.Epp + CHI2_PP_TILE - 1) / CHI2_PP_TILE;
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  double acc = 0.0;
  uint32_t phase = 0;
  for (int t = blockIdx.x; t < nPL + nPP; t += gridDim.x) {
    if (t < nPL) {
      const int e0 = t * CHI2_PL_TILE;
      const int n = min(CHI2_PL_TILE, G.El - e0);
      if (tid == 0) {
        mbar_expect_tx(&bar, n * (uint32_t)sizeof(PLEdge));
        tma_load_1d(tile, G.pl + e0, n * (uint32_t)sizeof(PLEdge), &bar);
      }
      mbar_wait(&bar, phase);
      phase ^= 1;
      if (tid < n) {
        const PLEdge* ed = reinterpret_cast<const PLEdge*>(tile) + tid;
        const Pose X = G.pose[ed->p];
        double p[3] = {G.lm[4 * (size_t)ed->l], G.lm[4 * (size_t)ed->l + 1], G.lm[4 * (size_t)ed->l + 2]};
        PLLin L;
        pl_linearize(X, p, ed->z, L);
        acc += quad3(ed->info, L.e);
      }
    } else {
      const int e0 = (t - nPL) * CHI2_PP_TILE;
      const int n = min(CHI2_PP_TILE, G.Epp - e0);
      if (tid == 0) {
        mbar_expect_tx(&bar, n * (uint32_t)sizeof(PPEdge));
        tma_load_1d(tile, G.pp + e0, n * (uint32_t)sizeof(PPEdge), &bar);
      }
      mbar_wait(&bar, phase);
      phase ^= 1;
      if (tid < n) {
        const PPEdge* ed = reinterpret_cast<const PPEdge*>(tile) + tid;
        const Pose Xi = G.pose[ed->i];
        const Pose Xj = G.pose[ed->j];
        double err[6];
        pp_linearize(Xi, Xj, ed->zt, ed->zq, err, nullptr, nullptr);
        acc += quad6(ed->info, err);
      }
    }
    __syncthreads();  // tile is reused by the next bulk copy
  }
  double bs = block_sum(acc, sh);
  if (tid == 0) {
    part[blockIdx.x] = bs;
    __threadfence();
    int ticket = atomicAdd(ticket_ctr, 1);
    is_last = (ticket == (int)gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    double s = 0.0;
    for (int k = tid; k < (int)gridDim.x; k += blockDim.x) s += __ldcg(part + k);
    s = block_sum(s, sh);
    if (tid == 0) {
      *out = s;
      *ticket_ctr = 0;
    }
  }
}

// restore estimates (LM reject / benchmark restore)
__global__ void k_copy_state(Pose* dst_pose, const Pose* src_pose, int Np, double* dst_lm, const double* src_lm, int Nl) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < Np) dst_pose[t] = src_pose[t];
  if (t < Nl) {
    double4 v = reinterpret_cast<const double4*>(src_lm)[t];
    reinterpret_cast<double4*>(dst_lm)[t] = v;
  }
}

// test hook: linearise one edge (kind 0 = pp, 1 = pl)
__global__ void k_edge_linearize(DevGraph G, int kind, int e, double* out /* err[6], Ji[36], Jj[36] */) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double* err = out;
  double* Ji = out + 6;
  double* Jj = out + 42;
  if (kind == 0) {
    const PPEdge* ed = G.pp + e;
    pp_linearize(G.pose[ed->i], G.pose[ed->j], ed->zt, ed->zq, err, Ji, Jj);
  } else {
    const PLEdge ed = G.pl[e];
    double p[3] = {G.lm[4 * (size_t)ed.l], G.lm[4 * (size_t)ed.l + 1], G.lm[4 * (size_t)ed.l + 2]};
    PLLin L;
    pl_linearize(G.pose[ed.p], p, ed.z, L);
    for (int k = 0; k < 3; ++k) err[k] = L.e[k];
    pl_jac_pose(L.pc, Ji);
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) Jj[3 * r + c] = L.R[3 * c + r];
  }
}

}  // namespace ssb
