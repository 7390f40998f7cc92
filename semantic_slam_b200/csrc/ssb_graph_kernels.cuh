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
#include "ssb_peer.cuh"

namespace ssb {
namespace cg = cooperative_groups;

struct DevGraph {
  int Np, Nl, El, Epp;
  // Sharded graphs (ssb_peer.cuh): this rank's local subgraph holds its own keyframes [0, Np_own) followed by
  // "ghost" keyframes owned by other ranks (they observe a landmark one of ours observes, or are odometry
  // neighbours); landmarks with lm_owned == 0 are eliminated by another rank.  Unsharded: Np_own == Np, null.
  int Np_own;
  const unsigned char* lm_owned;
  // 1 = pseudo-keyframe: a landmark promoted into the reduced system because it carries a landmark-landmark edge
  // (ssb_math.cuh: pp_edge_linearize); null = none
  const unsigned char* pose_kind;
  Pose* pose;
  double* lm;  // 4 doubles per landmark
  const unsigned char* pose_fixed;
  const unsigned char* lm_fixed;
  const unsigned char* lm_kind;  // per landmark vertex: 0 = VertexPointXYZ, 1 = VertexPlane; null = all XYZ
  const double* pl_zd;           // per pose-landmark edge (L-order): 4th coefficient of a measured plane; null = no planes
  const PLEdge* pl;
  const PPEdge* pp;
  const int* lm_rowptr;
  const int* pose_pl_rowptr;
  const int* pose_pl_idx;
  const int* pose_pp_rowptr;
  const int* pose_pp_idx;
  const int* pose_pp_other;   // per incidence: the keyframe at the other end (saves the dependent load of the edge record)
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

// Linearisation of one pose -> landmark edge at (X, landmark record lm4): error e(3), pose Jacobian Jp (3x6)
// and the TRANSPOSED landmark Jacobian JlT (3x3), row-major.  EdgeSE3PointXYZ: Jp = [-I | 2[pc]x], Jl = R'
// (so JlT = R); EdgeSE3Plane: exact dual-number Jacobians (ssb_math.cuh).
__device__ __noinline__ void pl_plane_lin(const Pose& X, const double* pl4, const double* zn, double zd, double* err,
                                          double* Jp, double* JlT) {
  const double zm[4] = {zn[0], zn[1], zn[2], zd};
  double Jl[9];
  plane_linearize(X, pl4, zm, err, Jp, Jl);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) JlT[3 * r + c] = Jl[3 * c + r];
}
__device__ __forceinline__ bool pl_is_plane(const DevGraph& G, int l) { return G.lm_kind != nullptr && G.lm_kind[l] != 0; }
__device__ __forceinline__ void pl_edge_lin(const DevGraph& G, const PLEdge& ed, int e_idx, const Pose& X, double* err,
                                            double* Jp, double* JlT) {
  const double* lm4 = G.lm + 4 * (size_t)ed.l;
  if (pl_is_plane(G, ed.l)) {
    const double pl4[4] = {lm4[0], lm4[1], lm4[2], lm4[3]};
    pl_plane_lin(X, pl4, ed.z, G.pl_zd[e_idx], err, Jp, JlT);
  } else {
    const double p[3] = {lm4[0], lm4[1], lm4[2]};
    PLLin L;
    pl_linearize(X, p, ed.z, L);
    for (int k = 0; k < 3; ++k) err[k] = L.e[k];
    pl_jac_pose(L.pc, Jp);
    for (int k = 0; k < 9; ++k) JlT[k] = L.R[k];
  }
}
__device__ __forceinline__ void pl_edge_err(const DevGraph& G, const PLEdge& ed, int e_idx, const Pose& X, double* err) {
  const double* lm4 = G.lm + 4 * (size_t)ed.l;
  if (pl_is_plane(G, ed.l)) {
    const double pl4[4] = {lm4[0], lm4[1], lm4[2], lm4[3]};
    const double zm[4] = {ed.z[0], ed.z[1], ed.z[2], G.pl_zd[e_idx]};
    plane_error(X, pl4, zm, err);
  } else {
    const double p[3] = {lm4[0], lm4[1], lm4[2]};
    PLLin L;
    pl_linearize(X, p, ed.z, L);
    for (int k = 0; k < 3; ++k) err[k] = L.e[k];
  }
}

// ---------------------------------------------------------------------------------------------
// K1a: landmark-major linearisation.  One thread per landmark walks its L-order edges:
//   Hll += R W R',  bl += -R W e,  HplL_e = R W Jp  (3x6)
// (BlockSolver::buildSystem -> EdgeSE3PointXYZ::linearizeOplus + constructQuadraticForm)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_lin_landmarks(DevGraph G) {
  // one warp per landmark, one lane per edge (stride 32); fixed-order butterfly sums => deterministic
  const int l = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (l >= G.Nl) return;
  const bool lfixed = G.lm_fixed[l] != 0;
  double H[6] = {0, 0, 0, 0, 0, 0}, b[3] = {0, 0, 0};
  for (int e = G.lm_rowptr[l] + lane; e < G.lm_rowptr[l + 1]; e += 32) {
    const PLEdge ed = G.pl[e];
    const Pose X = G.pose[ed.p];
    double err[3], Jp[18], R[9];  // R = Jl' (= the pose rotation for a point landmark)
    pl_edge_lin(G, ed, e, X, err, Jp, R);
    double W[9];
    expand_sym3(ed.info, W);
    // RW = Jl' * W (3x3)
    double RW[9];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) RW[3 * r + c] = R[3 * r] * W[c] + R[3 * r + 1] * W[3 + c] + R[3 * r + 2] * W[6 + c];
    if (!lfixed) {
      // Hll += Jl' W Jl
      int k = 0;
      for (int r = 0; r < 3; ++r)
        for (int c = r; c < 3; ++c) {
          H[k] += RW[3 * r] * R[3 * c] + RW[3 * r + 1] * R[3 * c + 1] + RW[3 * r + 2] * R[3 * c + 2];
          ++k;
        }
      for (int r = 0; r < 3; ++r) b[r] -= RW[3 * r] * err[0] + RW[3 * r + 1] * err[1] + RW[3 * r + 2] * err[2];
    }
    double* out = G.HplL + 18 * (size_t)e;
    const bool zero = lfixed || G.pose_fixed[ed.p];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 6; ++c)
        out[6 * r + c] = zero ? 0.0 : RW[3 * r] * Jp[c] + RW[3 * r + 1] * Jp[6 + c] + RW[3 * r + 2] * Jp[12 + c];
  }
  for (int k = 0; k < 6; ++k) H[k] = warp_sum(H[k]);
  for (int k = 0; k < 3; ++k) b[k] = warp_sum(b[k]);
  if (lfixed) {
    H[0] = H[3] = H[5] = 1.0;
    H[1] = H[2] = H[4] = 0.0;
  }
  if (lane == 0) {
    for (int k = 0; k < 6; ++k) G.Hll[6 * (size_t)l + k] = H[k];
    for (int k = 0; k < 3; ++k) G.bl[3 * (size_t)l + k] = b[k];
  }
}

// ---------------------------------------------------------------------------------------------
// K1b: pose-major linearisation.  One thread per pose accumulates its diagonal block and rhs over
// its pose-landmark edges (P-order) and its pose-pose edges; the thread that owns role i of a
// pose-pose edge writes the off-diagonal block Hoff_e = Ji' W Jj.  Deterministic (no atomics).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) k_lin_poses(DevGraph G) {
  // 32 poses per block: warp 0 walks the pose-landmark edges of its poses, warp 1 their pose-pose edges (two
  // independent dependent-load chains); warp 1 hands its partial block over through shared memory
  __shared__ double acc_sh[32][43];
  const int lane = threadIdx.x & 31, role_w = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  const bool live = i < G.Np;
  const bool fixed = live && G.pose_fixed[i] != 0;
  Pose X;
  if (live) X = G.pose[i];
  double H[36], b[6];
  for (int k = 0; k < 36; ++k) H[k] = 0.0;
  for (int k = 0; k < 6; ++k) b[k] = 0.0;
  // pose-landmark edges
  if (live && role_w == 0)
  for (int kk = G.pose_pl_rowptr[i]; kk < G.pose_pl_rowptr[i + 1]; ++kk) {
    const int e = G.pose_pl_idx[kk];
    const PLEdge ed = G.pl[e];
    double W[9], Jp[18], err[3], JlT[9];
    pl_edge_lin(G, ed, e, X, err, Jp, JlT);
    expand_sym3(ed.info, W);
    // JtW = Jp' W (6x3)
    double JtW[18];
    for (int r = 0; r < 6; ++r)
      for (int c = 0; c < 3; ++c) JtW[3 * r + c] = Jp[r] * W[c] + Jp[6 + r] * W[3 + c] + Jp[12 + r] * W[6 + c];
    if (!fixed) {
      for (int r = 0; r < 6; ++r) {
        for (int c = 0; c < 6; ++c) H[6 * r + c] += JtW[3 * r] * Jp[c] + JtW[3 * r + 1] * Jp[6 + c] + JtW[3 * r + 2] * Jp[12 + c];
        b[r] -= JtW[3 * r] * err[0] + JtW[3 * r + 1] * err[1] + JtW[3 * r + 2] * err[2];
      }
    }
    // HplP = Jp' W Jl  (6x3)
    const bool zero = fixed || G.lm_fixed[ed.l];
    double* out = G.HplP + 18 * (size_t)kk;
    for (int r = 0; r < 6; ++r)
      for (int c = 0; c < 3; ++c)
        out[3 * r + c] = zero ? 0.0 : JtW[3 * r] * JlT[3 * c] + JtW[3 * r + 1] * JlT[3 * c + 1] + JtW[3 * r + 2] * JlT[3 * c + 2];
    G.plP_lm[kk] = ed.l;
  }
  // pose-pose edges
  if (live && role_w == 1)
  for (int kk = G.pose_pp_rowptr[i]; kk < G.pose_pp_rowptr[i + 1]; ++kk) {
    const int code = G.pose_pp_idx[kk];
    const int e = code >> 1, role = code & 1;
    const PPEdge* ed = G.pp + e;
    const int other = role == 0 ? ed->j : ed->i;
    const Pose Y = G.pose[other];
    double err[6], Ji[36], Jj[36], W[36];
    if (role == 0)
      pp_edge_linearize(*ed, X, Y, err, Ji, Jj);
    else
      pp_edge_linearize(*ed, Y, X, err, Ji, Jj);
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
  if (role_w == 1) {
    for (int k = 0; k < 36; ++k) acc_sh[lane][k] = H[k];
    for (int k = 0; k < 6; ++k) acc_sh[lane][36 + k] = b[k];
  }
  __syncthreads();
  if (role_w == 1 || !live) return;
  for (int k = 0; k < 36; ++k) H[k] += acc_sh[lane][k];
  for (int k = 0; k < 6; ++k) b[k] += acc_sh[lane][36 + k];
  if (fixed) {
    for (int k = 0; k < 36; ++k) H[k] = 0.0;
    for (int k = 0; k < 6; ++k) {
      H[7 * k] = 1.0;
      b[k] = 0.0;
    }
  } else if (G.pose_kind && G.pose_kind[i]) {   // promoted landmark: the three unused increments are pinned
    H[21] = H[28] = H[35] = 1.0;
  }
  for (int k = 0; k < 36; ++k) G.Hpp[36 * (size_t)i + k] = H[k];
  for (int k = 0; k < 6; ++k) G.bp[6 * (size_t)i + k] = b[k];
}

// max |H_jj| over the non-fixed vertices (OptimizationAlgorithmLevenberg::computeLambdaInit)
__global__ void k_maxdiag(DevGraph G) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  double m = 0.0;
  if (t < G.Np) {
    if (!G.pose_fixed[t] && t < G.Np_own)
      for (int k = 0; k < ((G.pose_kind && G.pose_kind[t]) ? 3 : 6); ++k) m = fmax(m, fabs(G.Hpp[36 * (size_t)t + 7 * k]));
  } else if (t < G.Np + G.Nl) {
    int l = t - G.Np;
    if (!G.lm_fixed[l] && (G.lm_owned == nullptr || G.lm_owned[l])) {
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
  if (i >= G.Np_own) return;
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
// Coarse (rigid-body) level of the preconditioner.  Every persistent CTA owns a contiguous range of
// C poses = one aggregate; its 6 coarse unknowns are a world-frame translation d and rotation w of
// the aggregate about its centroid c, prolonged to pose i (right-multiplied MQT increment) by
//   B_i = [ R_i'  -R_i' [t_i - c]x ;  0  1/2 R_i' ]      (zero for fixed poses)
// A_c = P' S P (nc = 6 * #CTAs) is assembled and inverted inside k_pcg (block Gauss-Jordan, one
// pivot aggregate per step, one grid barrier per step); each CTA keeps its 6 rows of A_c^-1 in smem.
// ---------------------------------------------------------------------------------------------
struct CoarseDev {
  int enabled;
  int C;    // poses per CTA (multiple of 5)
  int nc;   // 6 * gridDim.x
  double* Bmat;            // [Np][36]
  double* cen;             // [gridDim.x][3] centroid of every CTA's aggregate (read by the rank-level coarse level)
  double* Grun;            // [n_runs][18]  (3x6) = sum_e HplL_e B_p(e) over the run
  const int* run_lm;       // [n_runs]
  const int* run_group;    // [n_runs]
  const int* run_e0;       // [n_runs+1] L-order edge range of each run
  const int* lm_run_rowptr;   // [Nl+1] runs of each landmark (contiguous)
  const int* grp_run_rowptr;  // [ngroups+1]
  const int* grp_runs;        // runs sorted by group
  int n_runs;
  double* panel;           // [2][6*nc + 8] Gauss-Jordan pivot panels (double buffered)
  // middle level: aggregates of 5 consecutive poses (= the 5 poses one warp owns)
  int sub_enabled;
  int reuse_inverse;       // 1: skip assembly + Gauss-Jordan, reload the rows of A_c^-1 stored by an earlier solve
  double* ainv_store;      // [gridDim.x][6*nc] rows of A_c^-1 kept between solves
  int* ainv_ok;            // [1] written by k_coarse_invert: every pivot block was positive definite
  double* B1mat;           // [Np][36] prolongation blocks about the 5-pose centroid
  double* D1inv;           // [ceil(Np/5)][36] inverse of P1' S P1 diagonal blocks (zero = level off for the aggregate)
  // (landmark, 5-pose aggregate) runs of the L-order edge table, built on the host like the per-CTA runs above
  double* Grun1;           // [n_runs1][18] (3x6) = sum_e HplL_e B1_p(e) over the run
  const int* run1_lm;      // [n_runs1]
  const int* run1_e0;      // [n_runs1+1]
  const int* agg_run_rowptr;  // [n_agg+1]
  const int* agg_runs;        // runs sorted by aggregate (landmark order inside an aggregate)
  int n_runs1;
  // preconditioner 3: the 5-pose aggregates of one CTA are coupled exactly inside two groups of <= GRP_MAXA
  // aggregates each (group matrix = P1' S P1 restricted to the group, <= 48x48, inverted per damped trial)
  int grp_enabled;
  double* D1raw;              // [n_agg][36] P1' S P1 diagonal blocks before inversion
  double* GrpInv;             // [n_groups][GRP_PACK] packed lower triangle of the group inverse (zeros = level off)
  const int* grp_first_agg;   // [n_groups+1] first aggregate of every group (groups 2b, 2b+1 belong to CTA b)
  const int* grp_seg_rowptr;  // [n_groups+1] landmark segments: >= 2 consecutive runs of one landmark inside a group
  const int* grp_seg_r0;      //              first run of the segment
  const int* grp_seg_m;       //              number of runs
  const int* grp_seg_off;     //              offset of the segment's runs inside the group's flat run list
  const int* run1_agg;        // [n_runs1] aggregate of every (landmark, aggregate) run
  int n_groups;
};
constexpr int GRP_MAXA = 8;                          // aggregates per group
constexpr int GRP_N = 6 * GRP_MAXA;                  // 48
constexpr int GRP_PACK = GRP_N * (GRP_N + 1) / 2;    // 1176

struct BarSlot {  // one 64 B line per CTA and buffer; slot [2*gridDim.x] holds the arrival counter
  double v[7];
  unsigned epoch;
  unsigned pad;
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_relaxed_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u32(unsigned* p, unsigned v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_add_u32(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Grid barrier fused with a fixed-order all-reduce of one double per CTA (+ optional all-gather of 6
// doubles per CTA into `gather` (smem, 6 * gridDim.x)).  Arrival = payload store + release-add on one
// counter; ONE thread per CTA spins on the counter (many pollers per CTA were measured 2x slower: the
// L2 polling traffic delays the CTAs that are still working); then every CTA reads the gridDim.x payload
// slots and every warp folds them in the same fixed order.  Slots are double buffered on the epoch parity.
// Every thread of every CTA must call it; requires gridDim.x <= blockDim.x and co-resident CTAs.
__device__ __forceinline__ double grid_bar_sum(BarSlot* slots, unsigned& epoch, double my_partial, const double* my6,
                                               double* gather, double* part_sh) {
  ++epoch;
  BarSlot* S = slots + (size_t)(epoch & 1u) * gridDim.x;
  unsigned* counter = &slots[2 * (size_t)gridDim.x].epoch;
  __syncthreads();
  if (threadIdx.x == 0) {
    BarSlot* me = S + blockIdx.x;
    me->v[0] = my_partial;
    if (my6) {
#pragma unroll
      for (int k = 0; k < 6; ++k) me->v[1 + k] = my6[k];
    }
    red_release_add_u32(counter, 1u);  // release: orders this CTA's earlier writes (bar.sync-cumulative)
    const unsigned target = epoch * gridDim.x;
    while (ld_acquire_u32(counter) < target) {
    }
  }
  __syncthreads();
  if (threadIdx.x < gridDim.x) {
    const BarSlot* o = S + threadIdx.x;
    part_sh[threadIdx.x] = __ldcg(&o->v[0]);
    if (gather) {
#pragma unroll
      for (int k = 0; k < 6; ++k) gather[6 * threadIdx.x + k] = __ldcg(&o->v[1 + k]);
    }
  }
  __syncthreads();
  double t = 0.0;
  for (int k = (threadIdx.x & 31); k < (int)gridDim.x; k += 32) t += part_sh[k];
  return warp_sum(t);
}

// deterministic block reduction of 7 values (v0 and v6[0..5]) with a single __syncthreads: every warp
// publishes its 7 warp sums, then every warp folds the per-warp values in the same fixed order.
// Result: returned value = sum of v0; out6[0..5] (smem) = sums of v6 (valid after the next __syncthreads).
__device__ __forceinline__ double block_sum7(double v0, const double* v6, double* out6, double* scratch /* 7*32 */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  double r[7];
  r[0] = warp_sum(v0);
#pragma unroll
  for (int k = 0; k < 6; ++k) r[1 + k] = warp_sum(v6[k]);
  __syncthreads();  // scratch may still be read from the previous use
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 7; ++k) scratch[7 * w + k] = r[k];
  }
  __syncthreads();
  // lane k < 7 of every warp folds value k over the warps in order; lane 0's result is broadcast
  double t = 0.0;
  if (lane < 7)
    for (int ww = 0; ww < nw; ++ww) t += scratch[7 * ww + lane];
  if (w == 0 && lane >= 1 && lane < 7) out6[lane - 1] = t;
  return __shfl_sync(0xffffffffu, t, 0);
}

// deterministic block reduction of 6 values per thread -> out6 (smem), valid after return
__device__ __forceinline__ void block_sum6(const double* v, double* out6, double* scratch /* 6*32 */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  double r[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) r[k] = warp_sum(v[k]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 6; ++k) scratch[6 * w + k] = r[k];
  }
  __syncthreads();
  if (w == 0) {
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      double t = lane < nw ? scratch[6 * lane + k] : 0.0;
      t = warp_sum(t);
      if (lane == 0) out6[k] = t;
    }
  }
  __syncthreads();
}

// 6x6 SPD inverse by 6 lanes of one warp (lane j owns column j of [A | I]); all 32 lanes must call.
// col[6] in: column j of A (lanes >= 6: ignored); out: column j of A^-1.  Returns false if a pivot is
// not positive.
__device__ __forceinline__ bool warp_inv6(double* col) {
  const int lane = threadIdx.x & 31;
  double inv[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) inv[i] = (i == lane) ? 1.0 : 0.0;
  bool ok = true;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const double piv = __shfl_sync(0xffffffffu, col[k], k);
    if (!(piv > 0.0) || !isfinite(piv)) ok = false;
    const double ip = 1.0 / piv;
    double ck[6];  // column k of A (multipliers a_ik)
#pragma unroll
    for (int i = 0; i < 6; ++i) ck[i] = __shfl_sync(0xffffffffu, col[i], k);
    const double akj = col[k] * ip, bkj = inv[k] * ip;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      if (i == k) {
        col[i] = akj;
        inv[i] = bkj;
      } else {
        col[i] -= ck[i] * akj;
        inv[i] -= ck[i] * bkj;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) col[i] = inv[i];
  return __all_sync(0xffffffffu, ok || lane >= 6);
}

// per LM iteration (poses changed): aggregate centroids and the prolongation blocks B_i
__global__ void __launch_bounds__(256) k_coarse_basis(DevGraph G, CoarseDev Cz) {
  __shared__ double sh[33];
  __shared__ double cen[3];
  const int p0 = blockIdx.x * Cz.C, p1 = min(G.Np_own, p0 + Cz.C);
  double s[3] = {0, 0, 0};
  for (int i = p0 + threadIdx.x; i < p1; i += blockDim.x)
    for (int k = 0; k < 3; ++k) s[k] += G.pose[i].t[k];
  for (int k = 0; k < 3; ++k) {
    double t = block_sum(s[k], sh);
    if (threadIdx.x == 0) {
      cen[k] = p1 > p0 ? t / (double)(p1 - p0) : 0.0;
      if (Cz.cen) Cz.cen[3 * blockIdx.x + k] = cen[k];
    }
  }
  __syncthreads();
  for (int i = p0 + threadIdx.x; i < p1; i += blockDim.x) {
    double* B = Cz.Bmat + 36 * (size_t)i;
    if (G.pose_fixed[i] || (G.pose_kind && G.pose_kind[i])) {   // (a promoted landmark is not part of a rigid body)
      for (int k = 0; k < 36; ++k) B[k] = 0.0;
      continue;
    }
    const Pose X = G.pose[i];
    double R[9];
    quat_to_R(X.q, R);
    const double d[3] = {X.t[0] - cen[0], X.t[1] - cen[1], X.t[2] - cen[2]};
    const double Sx[9] = {0, -d[2], d[1], d[2], 0, -d[0], -d[1], d[0], 0};
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        const double rt = R[3 * c + r];  // R'
        B[6 * r + c] = rt;
        B[6 * r + 3 + c] = -(R[r] * Sx[c] + R[3 + r] * Sx[3 + c] + R[6 + r] * Sx[6 + c]);  // -(R' [d]x)
        B[6 * (3 + r) + c] = 0.0;
        B[6 * (3 + r) + 3 + c] = 0.5 * rt;
      }
  }
}

// per LM iteration: G_run = sum over the run of HplL_e (3x6) * B_p(e) (6x6)
__global__ void __launch_bounds__(128) k_coarse_runs(DevGraph G, CoarseDev Cz) {
  // four lanes per (run, entry of the 3x6 product): Grun = sum_e HplL_e B_p(e); lane `part` takes every 4th edge of
  // the run and the four partial sums are added by two butterfly steps (fixed order => deterministic)
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int part = t & 3, rq = t >> 2;
  const int r = rq / 18, q = rq - 18 * r;
  const bool live = r < Cz.n_runs;
  const int a = q / 6, c = q - 6 * a;
  double acc = 0.0;
  if (live) {
    const int e1 = Cz.run_e0[r + 1];
    for (int e = Cz.run_e0[r] + part; e < e1; e += 4) {
      const double* Hl = G.HplL + 18 * (size_t)e + 6 * a;
      const double* B = Cz.Bmat + 36 * (size_t)G.pl[e].p + c;
#pragma unroll
      for (int k = 0; k < 6; ++k) acc += Hl[k] * B[6 * k];
    }
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  if (live && part == 0) Cz.Grun[18 * (size_t)r + q] = acc;
}

// ---- middle level (5-pose aggregates) -----------------------------------------------------------
__global__ void __launch_bounds__(128) k_sub_basis(DevGraph G, CoarseDev Cz) {
  // one thread per pose; the aggregate centroid is recomputed by each of its (<= 5) threads in the same order
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= G.Np) return;
  const int i0 = 5 * (i / 5), i1 = min(G.Np_own, i0 + 5);   // ghost keyframes carry no basis (B = 0)
  double cen[3] = {0, 0, 0};
  for (int q = i0; q < i1; ++q)
    for (int k = 0; k < 3; ++k) cen[k] += G.pose[q].t[k];
  for (int k = 0; k < 3; ++k) cen[k] /= (double)max(1, i1 - i0);
  double* B = Cz.B1mat + 36 * (size_t)i;
  if (G.pose_fixed[i] || i >= G.Np_own || (G.pose_kind && G.pose_kind[i])) {
    for (int k = 0; k < 36; ++k) B[k] = 0.0;
    return;
  }
  const Pose X = G.pose[i];
  double R[9];
  quat_to_R(X.q, R);
  const double d[3] = {X.t[0] - cen[0], X.t[1] - cen[1], X.t[2] - cen[2]};
  const double Sx[9] = {0, -d[2], d[1], d[2], 0, -d[0], -d[1], d[0], 0};
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      const double rt = R[3 * c + r];
      B[6 * r + c] = rt;
      B[6 * r + 3 + c] = -(R[r] * Sx[c] + R[3 + r] * Sx[3 + c] + R[6 + r] * Sx[6 + c]);
      B[6 * (3 + r) + c] = 0.0;
      B[6 * (3 + r) + 3 + c] = 0.5 * rt;
    }
}

// per LM iteration: Grun1 of every (landmark, 5-pose aggregate) run.  18 threads per run (one per entry of the 3x6 product).
__global__ void __launch_bounds__(128) k_sub_runs(DevGraph G, CoarseDev Cz) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = t / 18, q = t - 18 * r;
  if (r >= Cz.n_runs1) return;
  const int a = q / 6, c = q - 6 * a;
  double acc = 0.0;
  for (int e = Cz.run1_e0[r]; e < Cz.run1_e0[r + 1]; ++e) {
    const double* Hl = G.HplL + 18 * (size_t)e + 6 * a;
    const double* B = Cz.B1mat + 36 * (size_t)G.pl[e].p + c;
#pragma unroll
    for (int k = 0; k < 6; ++k) acc += Hl[k] * B[6 * k];
  }
  Cz.Grun1[18 * (size_t)r + q] = acc;
}

// per damped trial: D1 = P1' S P1 restricted to the aggregate (6x6).  One block per aggregate: slice s (36 threads,
// one per entry) takes pose s of the aggregate and every 5th (landmark, aggregate) run, so the dependent-load chains
// are 5x shorter than with one slice; the slices are added in a fixed order (deterministic).  The inverse (needed by
// preconditioner 2 and by the streaming kernel only) is taken by one thread when need_inv != 0.
constexpr int SUBA_THREADS = 192;   // 5 slices x 36 entries (+ 12 idle threads)
__global__ void __launch_bounds__(SUBA_THREADS) k_sub_assemble(DevGraph G, CoarseDev Cz, double lambda, int need_inv) {
  __shared__ double part[5][36];
  __shared__ double Dsh[36];
  __shared__ int anysh[5];
  const int sl = threadIdx.x / 36, ent = threadIdx.x - 36 * sl;
  const int a = blockIdx.x;
  const int r = ent / 6, c = ent - 6 * r;
  const int i0 = 5 * a, i1 = min(G.Np_own, i0 + 5);
  if (sl < 5) {
    double d = 0.0;
    bool any = false;
    const int i = i0 + sl;
    if (i < i1 && !G.pose_fixed[i]) {
      any = true;
      const double* B = Cz.B1mat + 36 * (size_t)i;
      const double* H = G.Hpp + 36 * (size_t)i;
      const int k0 = G.pose_pp_rowptr[i], k1 = G.pose_pp_rowptr[i + 1];
      double t = 0.0;
#pragma unroll
      for (int x = 0; x < 6; ++x) {
        double hb = lambda * B[6 * x + c];
#pragma unroll
        for (int y = 0; y < 6; ++y) hb += H[6 * x + y] * B[6 * y + c];
        t += B[6 * x + r] * hb;
      }
      d += t;
      for (int kk = k0; kk < k1; ++kk) {
        const int code = G.pose_pp_idx[kk];
        if (code & 1) continue;
        const int e = code >> 1, j = G.pp[e].j;
        if (j < i0 || j >= i1) continue;
        const double* Bj = Cz.B1mat + 36 * (size_t)j;
        const double* Ho = G.Hoff + 36 * (size_t)e;
        double t1 = 0.0, t2 = 0.0;   // (B_i' Ho B_j)[r][c] and [c][r]
#pragma unroll
        for (int x = 0; x < 6; ++x) {
          double h1 = 0.0, h2 = 0.0;
#pragma unroll
          for (int y = 0; y < 6; ++y) {
            h1 += Ho[6 * x + y] * Bj[6 * y + c];
            h2 += Ho[6 * x + y] * Bj[6 * y + r];
          }
          t1 += B[6 * x + r] * h1;
          t2 += B[6 * x + c] * h2;
        }
        d += t1 + t2;
      }
    }
    // landmark terms: D -= sum over my (landmark, aggregate) runs of G' W G with G = Grun1 (3x6)
    for (int q = Cz.agg_run_rowptr[a] + sl; q < Cz.agg_run_rowptr[a + 1]; q += 5) {
      const int r1 = Cz.agg_runs[q];
      const double* Gr = Cz.Grun1 + 18 * (size_t)r1;
      const double* Wu = G.HllInv + 6 * (size_t)Cz.run1_lm[r1];
      const double w0 = Wu[0] * Gr[c] + Wu[1] * Gr[6 + c] + Wu[2] * Gr[12 + c];
      const double w1 = Wu[1] * Gr[c] + Wu[3] * Gr[6 + c] + Wu[4] * Gr[12 + c];
      const double w2 = Wu[2] * Gr[c] + Wu[4] * Gr[6 + c] + Wu[5] * Gr[12 + c];
      d -= Gr[r] * w0 + Gr[6 + r] * w1 + Gr[12 + r] * w2;
    }
    part[sl][ent] = d;
    if (ent == 0) anysh[sl] = any ? 1 : 0;
  }
  __syncthreads();
  if (threadIdx.x < 36) Dsh[ent] = (((part[0][ent] + part[1][ent]) + part[2][ent]) + part[3][ent]) + part[4][ent];
  __syncthreads();
  const bool any = (anysh[0] | anysh[1] | anysh[2] | anysh[3] | anysh[4]) != 0;
  if (threadIdx.x < 36 && Cz.grp_enabled) Cz.D1raw[36 * (size_t)a + ent] = any ? 0.5 * (Dsh[ent] + Dsh[6 * c + r]) : 0.0;
  if (threadIdx.x == 0 && need_inv) {
    double D[36];
    for (int k = 0; k < 36; ++k) D[k] = 0.5 * (Dsh[k] + Dsh[6 * (k % 6) + k / 6]);
    if (!any || !inv_spd6(D))
      for (int k = 0; k < 36; ++k) D[k] = 0.0;
    for (int k = 0; k < 36; ++k) Cz.D1inv[36 * (size_t)a + k] = D[k];
  }
}

// per damped trial (preconditioner 3): assemble P1' S P1 restricted to one group of 5-pose aggregates and invert it
// (Gauss-Jordan, SPD => no pivoting).  One CTA of 256 threads per group.  Assembly in shared memory with fixed entry
// ownership (no atomics, deterministic); the elimination keeps the matrix in registers (thread (ti, tj) owns the
// entries (ti + 16 a, tj + 16 b), a, b < 3), publishes the pivot row / column through double-buffered shared vectors
// and needs one barrier per pivot.
constexpr int GRP_THREADS = 256;
constexpr int GRP_MAXSEG = 192;   // landmark segments per group held in shared memory
constexpr int GRP_MAXRUN = 1024;  // runs of those segments
constexpr int GRP_MAXPP = 64;     // pose-pose edges between different aggregates of one group
__device__ __forceinline__ double grp_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(r, fma(-x, r, 1.0), r);
  r = fma(r, fma(-x, r, 1.0), r);
  return r;
}
__global__ void __launch_bounds__(GRP_THREADS) k_grp_invert(DevGraph G, CoarseDev Cz, int max_runs) {
  constexpr int LD = GRP_N + 1;
  extern __shared__ __align__(16) double gdyn[];   // [max_runs][18] Grun1 of every run of the group's segments
  __shared__ double A[GRP_N * LD];
  __shared__ double Wsh[GRP_MAXSEG * 6];
  __shared__ short segoff[GRP_MAXSEG + 1];
  __shared__ short segpos[GRP_MAXSEG * GRP_MAXA];   // per segment: aggregate -> staged run (or -1)
  __shared__ unsigned char segmask[GRP_MAXSEG];     // per segment: bit a = aggregate a has a run
  __shared__ int runidx[GRP_MAXRUN];                // staged run -> (landmark, aggregate) run
  __shared__ short runseg[GRP_MAXRUN];              // staged run -> segment
  __shared__ double rowk[2][GRP_N], colk[2][GRP_N];
  __shared__ int ppcnt[5 * GRP_MAXA + 1];
  __shared__ int pplist[3 * GRP_MAXPP];             // (edge, pose i, pose j)
  const int g = blockIdx.x, tid = threadIdx.x;
  const int a0 = Cz.grp_first_agg[g], a1 = Cz.grp_first_agg[g + 1];
  const int na = a1 - a0, n = 6 * na;
  double* out = Cz.GrpInv + (size_t)g * GRP_PACK;
  if (na <= 0 || na > GRP_MAXA) {
    for (int k = tid; k < GRP_PACK; k += GRP_THREADS) out[k] = 0.0;
    return;
  }
  const int sg0 = Cz.grp_seg_rowptr[g], nseg = min(Cz.grp_seg_rowptr[g + 1] - sg0, GRP_MAXSEG);
  for (int k = tid; k < GRP_N * LD; k += GRP_THREADS) A[k] = 0.0;
  // stage every landmark segment of the group: pass 1, one thread per segment (its W, the run list);
  // pass 2, flat over all (run, entry) pairs — no dependent global loads inside serial loops
  for (int sg = tid; sg < nseg; sg += GRP_THREADS) {
    const int r0 = Cz.grp_seg_r0[sg0 + sg], o = Cz.grp_seg_off[sg0 + sg], m = Cz.grp_seg_m[sg0 + sg];
    segoff[sg] = (short)o;
    if (sg == nseg - 1) segoff[nseg] = (short)(o + m);
    const double* Wu = G.HllInv + 6 * (size_t)Cz.run1_lm[r0];
    for (int k = 0; k < 6; ++k) Wsh[6 * sg + k] = Wu[k];
    for (int k = 0; k < GRP_MAXA; ++k) segpos[GRP_MAXA * sg + k] = -1;
    for (int x = 0; x < m; ++x)
      if (o + x < GRP_MAXRUN) {
        runidx[o + x] = r0 + x;
        runseg[o + x] = (short)sg;
      }
  }
  if (nseg == 0 && tid == 0) segoff[0] = 0;
  // pose-pose edges between different aggregates of the group: every pose counts its own (role 0) ...
  const int np = min(G.Np_own, 5 * a1) - 5 * a0;
  int my_n = 0;
  if (tid < np) {
    const int i = 5 * a0 + tid, ai = tid / 5;
    for (int kk = G.pose_pp_rowptr[i]; kk < G.pose_pp_rowptr[i + 1]; ++kk) {
      const int code = G.pose_pp_idx[kk];
      if (code & 1) continue;
      const int aj = G.pp[code >> 1].j / 5 - a0;
      if (aj != ai && aj >= 0 && aj < na) ++my_n;
    }
    ppcnt[tid] = my_n;
  }
  __syncthreads();
  const int nrun = min((int)segoff[nseg], min(max_runs, GRP_MAXRUN));
  for (int idx = tid; idx < 18 * nrun; idx += GRP_THREADS) {
    const int rr = idx / 18, q = idx - 18 * rr, rg = runidx[rr];
    gdyn[idx] = Cz.Grun1[18 * (size_t)rg + q];
    if (q == 0) segpos[GRP_MAXA * runseg[rr] + (Cz.run1_agg[rg] - a0)] = (short)rr;
  }
  for (int idx = tid; idx < 36 * na; idx += GRP_THREADS) {
    const int a = idx / 36, e = idx - 36 * a;
    A[(6 * a + e / 6) * LD + 6 * a + e % 6] = Cz.D1raw[36 * (size_t)(a0 + a) + e];
  }
  // ... and appends them to the list at a position fixed by the pose order (deterministic)
  if (tid < np && my_n > 0) {   // (more than GRP_MAXPP such edges per group: the rest is left out of the preconditioner)
    int o = 0;
    for (int k = 0; k < tid; ++k) o += ppcnt[k];
    const int i = 5 * a0 + tid, ai = tid / 5;
    for (int kk = G.pose_pp_rowptr[i]; kk < G.pose_pp_rowptr[i + 1]; ++kk) {
      const int code = G.pose_pp_idx[kk];
      if (code & 1) continue;
      const int e = code >> 1, j = G.pp[e].j, aj = j / 5 - a0;
      if (aj == ai || aj < 0 || aj >= na) continue;
      if (o < GRP_MAXPP) {
        pplist[3 * o] = e;
        pplist[3 * o + 1] = i;
        pplist[3 * o + 2] = j;
      }
      ++o;
    }
  }
  if (tid == 0) {
    int o = 0;
    for (int k = 0; k < np; ++k) o += ppcnt[k];
    ppcnt[5 * GRP_MAXA] = min(o, GRP_MAXPP);
  }
  __syncthreads();
  for (int sg = tid; sg < nseg; sg += GRP_THREADS) {
    unsigned m = 0;
    for (int k = 0; k < GRP_MAXA; ++k)
      if (segpos[GRP_MAXA * sg + k] >= 0) m |= 1u << k;
    segmask[sg] = (unsigned char)m;
  }
  // B_i' Hoff B_j and its transpose: 36 threads per list item; all items of one aggregate pair go to the same
  // slice in list order, so the sums are deterministic and need no atomics
  {
    const int npp = ppcnt[5 * GRP_MAXA];
    const int sl = tid / 36, ent = tid - 36 * sl, r = ent / 6, c = ent - 6 * r;
    if (sl < 7)
      for (int it = 0; it < npp; ++it) {
        const int e = pplist[3 * it], i = pplist[3 * it + 1], j = pplist[3 * it + 2];
        const int ai = i / 5 - a0, aj = j / 5 - a0;
        const int lo = min(ai, aj), hi = max(ai, aj);
        if ((lo * GRP_MAXA + hi) % 7 != sl) continue;
        // thread (r, c) owns entry (r, c) of block (lo, hi): (B_i' Ho B_j)[r][c], transposed when the edge runs hi -> lo
        const int rr = ai < aj ? r : c, cc = ai < aj ? c : r;
        const double* B = Cz.B1mat + 36 * (size_t)i;
        const double* Bj = Cz.B1mat + 36 * (size_t)j;
        const double* Ho = G.Hoff + 36 * (size_t)e;
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          double hb = 0.0;
#pragma unroll
          for (int m = 0; m < 6; ++m) hb += Ho[6 * k + m] * Bj[6 * m + cc];
          t += B[6 * k + rr] * hb;
        }
        A[(6 * lo + r) * LD + 6 * hi + c] += t;
        A[(6 * hi + c) * LD + 6 * lo + r] += t;
      }
  }
  __syncthreads();
  // landmarks seen from several aggregates of the group: every thread owns fixed entries of the blocks (ai < aj)
  // and walks the staged segments: A[i][j] -= sum_l (G_x' W_l G_y)[i%6][j%6], mirrored into (j, i)
  {
    const int npair = na * (na - 1) / 2;
    for (int idx = tid; idx < 36 * npair; idx += GRP_THREADS) {
      const int pr = idx / 36, ent = idx - 36 * pr, ri = ent / 6, cj = ent - 6 * ri;
      int aj = 1, ai = pr;
      while (ai >= aj) {
        ai -= aj;
        ++aj;
      }
      const unsigned need = (1u << ai) | (1u << aj);
      double t = 0.0;
      for (int sg = 0; sg < nseg; ++sg) {
        if ((segmask[sg] & need) != need) continue;
        const int x = segpos[GRP_MAXA * sg + ai], y = segpos[GRP_MAXA * sg + aj];
        const double* Gx = gdyn + 18 * x;
        const double* Gy = gdyn + 18 * y;
        const double* Wu = Wsh + 6 * sg;
        const double w0 = Wu[0] * Gy[cj] + Wu[1] * Gy[6 + cj] + Wu[2] * Gy[12 + cj];
        const double w1 = Wu[1] * Gy[cj] + Wu[3] * Gy[6 + cj] + Wu[4] * Gy[12 + cj];
        const double w2 = Wu[2] * Gy[cj] + Wu[4] * Gy[6 + cj] + Wu[5] * Gy[12 + cj];
        t += Gx[ri] * w0 + Gx[6 + ri] * w1 + Gx[12 + ri] * w2;
      }
      A[(6 * ai + ri) * LD + 6 * aj + cj] -= t;
      A[(6 * aj + cj) * LD + 6 * ai + ri] -= t;
    }
  }
  __syncthreads();
  // aggregates without a free pose have a zero block: decouple them with an identity
  if (tid < na) {
    bool zero = true;
    for (int k = 0; k < 6; ++k)
      if (A[(6 * tid + k) * LD + 6 * tid + k] != 0.0) zero = false;
    if (zero)
      for (int k = 0; k < 6; ++k) A[(6 * tid + k) * LD + 6 * tid + k] = 1.0;
  }
  __syncthreads();
  // in-place Gauss-Jordan inverse in registers
  const int ti = tid >> 4, tj = tid & 15;
  double v[3][3];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const int i = ti + 16 * a, j = tj + 16 * b;
      v[a][b] = (i < n && j < n) ? A[i * LD + j] : 0.0;
    }
  bool fail = false;
  for (int k = 0; k < n; ++k) {
    const int p = k & 1, ka = k >> 4, kr = k & 15;
    if (ti == kr) {
#pragma unroll
      for (int b = 0; b < 3; ++b) rowk[p][tj + 16 * b] = ka == 0 ? v[0][b] : (ka == 1 ? v[1][b] : v[2][b]);
    }
    if (tj == kr) {
#pragma unroll
      for (int a = 0; a < 3; ++a) colk[p][ti + 16 * a] = ka == 0 ? v[a][0] : (ka == 1 ? v[a][1] : v[a][2]);
    }
    __syncthreads();
    const double pk = rowk[p][k];
    if (!(pk > 0.0) || !isfinite(pk)) {   // uniform: every thread reads the same pivot
      fail = true;
      break;
    }
    const double ip = grp_rcp(pk);
    double ci[3], rj[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) ci[a] = colk[p][ti + 16 * a] * ip;
#pragma unroll
    for (int b = 0; b < 3; ++b) rj[b] = rowk[p][tj + 16 * b];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const bool ik = (ti + 16 * a) == k, jk = (tj + 16 * b) == k;
        const double upd = v[a][b] - ci[a] * rj[b];
        v[a][b] = ik ? (jk ? ip : rj[b] * ip) : (jk ? -ci[a] : upd);
      }
  }
  __syncthreads();
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) A[(ti + 16 * a) * LD + tj + 16 * b] = v[a][b];
  __syncthreads();
  // packed lower triangle, symmetrised
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const int r = ti + 16 * a, c = tj + 16 * b;
      if (c <= r) out[r * (r + 1) / 2 + c] = (fail || r >= n) ? 0.0 : 0.5 * (A[r * LD + c] + A[c * LD + r]);
    }
}

// ---------------------------------------------------------------------------------------------
// Coarse prologue shared by both PCG kernels: CTA g assembles rows [6g, 6g+6) of A_c = P' S P in
// shared memory (Arow[6][nc]) and the grid inverts A_c in place by block Gauss-Jordan (pivot
// aggregate k at step k: its owner publishes the scaled pivot panel, one barrier per step).
// Returns true when A_c^-1 is usable (every pivot block was positive definite).
// ---------------------------------------------------------------------------------------------
// assembly of my 6 rows of A_c = P' S P into Arow[6][nc] (shared memory)
template <int NT>
__device__ void coarse_assemble(const DevGraph& G, const CoarseDev& Cz, double lambda, double* Arow, double* red, int p0,
                                int p1) {
  constexpr int SLICES = NT / 36;
  const int nblk = gridDim.x, nc = 6 * nblk, myg = blockIdx.x;
  for (int k = threadIdx.x; k < 6 * nc; k += NT) Arow[k] = 0.0;
  __syncthreads();
  // (a) diagonal block: sum_i B_i' (Hpp_ii + lambda I) B_i + pose-pose edges inside the aggregate
  {
    const int ent = threadIdx.x % 36, sl = threadIdx.x / 36;
    const int r = ent / 6, c = ent - 6 * r;
    if (sl < SLICES) {
      double acc = 0.0;
      for (int i = p0 + sl; i < p1; i += SLICES) {
        const double* B = Cz.Bmat + 36 * (size_t)i;
        const double* H = G.Hpp + 36 * (size_t)i;
        double t = 0.0;
        for (int a = 0; a < 6; ++a) {
          double hb = lambda * B[6 * a + c];
          for (int b = 0; b < 6; ++b) hb += H[6 * a + b] * B[6 * b + c];
          t += B[6 * a + r] * hb;
        }
        acc += t;
        for (int kk = G.pose_pp_rowptr[i]; kk < G.pose_pp_rowptr[i + 1]; ++kk) {
          const int code = G.pose_pp_idx[kk];
          if (code & 1) continue;
          const int e = code >> 1;
          const int j = G.pp[e].j;
          if (j < p0 || j >= p1) continue;
          const double* Bj = Cz.Bmat + 36 * (size_t)j;
          const double* Ho = G.Hoff + 36 * (size_t)e;
          double t1 = 0.0, t2 = 0.0;
          for (int a = 0; a < 6; ++a) {
            double hb1 = 0.0, hb2 = 0.0;
            for (int b = 0; b < 6; ++b) {
              hb1 += Ho[6 * a + b] * Bj[6 * b + c];
              hb2 += Ho[6 * a + b] * Bj[6 * b + r];
            }
            t1 += B[6 * a + r] * hb1;  // (Bi' Hoff Bj)[r][c]
            t2 += B[6 * a + c] * hb2;  // its transpose
          }
          acc += t1 + t2;
        }
      }
      red[36 * sl + ent] = acc;
    }
    __syncthreads();
    if (threadIdx.x < 36) {
      double t = 0.0;
      for (int k = 0; k < SLICES; ++k) t += red[36 * k + threadIdx.x];
      const int rr = threadIdx.x / 6, cc = threadIdx.x - 6 * rr;
      // an aggregate without a free keyframe (empty, or only fixed keyframes / promoted landmarks: B = 0) has an exactly zero
      // block: identity keeps A_c invertible
      if (rr == cc && t == 0.0) t = 1.0;
      Arow[rr * nc + 6 * myg + cc] = t;
    }
    __syncthreads();
  }
  // (b) pose-pose edges leaving the aggregate, (c) landmark terms; fixed order, 36 lanes (one per entry)
  if (threadIdx.x < 36) {
    const int r = threadIdx.x / 6, c = threadIdx.x - 6 * r;
    for (int i = p0; i < p1; ++i) {
      const double* B = Cz.Bmat + 36 * (size_t)i;
      for (int kk = G.pose_pp_rowptr[i]; kk < G.pose_pp_rowptr[i + 1]; ++kk) {
        const int code = G.pose_pp_idx[kk];
        const int e = code >> 1, role = code & 1;
        const int j = role == 0 ? G.pp[e].j : G.pp[e].i;
        if ((j >= p0 && j < p1) || j >= G.Np_own) continue;   // ghost keyframes have no coarse unknowns here
        const int gj = j / Cz.C;
        const double* Bj = Cz.Bmat + 36 * (size_t)j;
        const double* Ho = G.Hoff + 36 * (size_t)e;
        double t = 0.0;
        for (int a = 0; a < 6; ++a) {
          double hb = 0.0;
          for (int b = 0; b < 6; ++b) hb += (role == 0 ? Ho[6 * a + b] : Ho[6 * b + a]) * Bj[6 * b + c];
          t += B[6 * a + r] * hb;
        }
        Arow[r * nc + 6 * gj + c] += t;
      }
    }
    for (int q = Cz.grp_run_rowptr[myg]; q < Cz.grp_run_rowptr[myg + 1]; ++q) {
      const int ra = Cz.grp_runs[q];
      const int l = Cz.run_lm[ra];
      const double* Ga = Cz.Grun + 18 * (size_t)ra;
      const double* Wu = G.HllInv + 6 * (size_t)l;
      const double W[9] = {Wu[0], Wu[1], Wu[2], Wu[1], Wu[3], Wu[4], Wu[2], Wu[4], Wu[5]};
      double u[3];
      for (int v = 0; v < 3; ++v) u[v] = Ga[r] * W[v] + Ga[6 + r] * W[3 + v] + Ga[12 + r] * W[6 + v];
      for (int rb = Cz.lm_run_rowptr[l]; rb < Cz.lm_run_rowptr[l + 1]; ++rb) {
        const double* Gb = Cz.Grun + 18 * (size_t)rb;
        const int gb = Cz.run_group[rb];
        Arow[r * nc + 6 * gb + c] -= u[0] * Gb[c] + u[1] * Gb[6 + c] + u[2] * Gb[12 + c];
      }
    }
  }
  __syncthreads();
}

template <int NT>
__device__ bool coarse_prologue(const DevGraph& G, const CoarseDev& Cz, BarSlot* slots, unsigned& epoch, double lambda,
                                double* Arow, double* panel_sh, double* red, double* part_sh, int p0, int p1) {
  __shared__ int s_flag;
  __shared__ double piv_sh[40];
  const int nblk = gridDim.x, nc = 6 * nblk;
  if (threadIdx.x == 0) s_flag = 0;
  coarse_assemble<NT>(G, Cz, lambda, Arow, red, p0, p1);
  // ---- block Gauss-Jordan ----
  for (int k = 0; k < nblk; ++k) {
    double* gp = Cz.panel + (size_t)(k & 1) * (6 * nc + 8);
    if ((int)blockIdx.x == k) {
      if (threadIdx.x < 32) {  // pivot inverse by warp 0
        double col[6];
        const int lane = threadIdx.x;
#pragma unroll
        for (int a = 0; a < 6; ++a) col[a] = lane < 6 ? Arow[a * nc + 6 * k + lane] : 0.0;
        const bool ok = warp_inv6(col);
        if (lane < 6) {
#pragma unroll
          for (int a = 0; a < 6; ++a) piv_sh[6 * a + lane] = ok ? col[a] : (a == lane ? 1.0 : 0.0);
        }
        if (lane == 0) piv_sh[36] = ok ? 0.0 : 1.0;
      }
      __syncthreads();
      for (int j = threadIdx.x; j < nc; j += NT) {
        double colv[6], out[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) colv[a] = Arow[a * nc + j];
        const bool inpiv = (j >= 6 * k && j < 6 * k + 6);
#pragma unroll
        for (int a = 0; a < 6; ++a) {
          double t = 0.0;
          if (inpiv) {
            t = piv_sh[6 * a + (j - 6 * k)];
          } else {
#pragma unroll
            for (int b = 0; b < 6; ++b) t += piv_sh[6 * a + b] * colv[b];
          }
          out[a] = t;
        }
#pragma unroll
        for (int a = 0; a < 6; ++a) {
          Arow[a * nc + j] = out[a];
          gp[a * nc + j] = out[a];
        }
      }
      if (threadIdx.x == 0) {
        gp[6 * nc] = piv_sh[36];
        if (piv_sh[36] != 0.0) s_flag = 1;
      }
    }
    grid_bar_sum(slots, epoch, 0.0, nullptr, nullptr, part_sh);
    if ((int)blockIdx.x != k) {
      for (int j = threadIdx.x; j < 6 * nc; j += NT) panel_sh[j] = __ldcg(gp + j);
      if (threadIdx.x == 0 && __ldcg(gp + 6 * nc) != 0.0) s_flag = 1;
      if (threadIdx.x < 36) piv_sh[threadIdx.x] = Arow[(threadIdx.x / 6) * nc + 6 * k + (threadIdx.x % 6)];  // F
      __syncthreads();
      for (int j = threadIdx.x; j < nc; j += NT) {
        const bool inpiv = (j >= 6 * k && j < 6 * k + 6);
        double pj[6];
#pragma unroll
        for (int b = 0; b < 6; ++b) pj[b] = panel_sh[b * nc + j];
#pragma unroll
        for (int a = 0; a < 6; ++a) {
          double t = 0.0;
#pragma unroll
          for (int b = 0; b < 6; ++b) t += piv_sh[6 * a + b] * pj[b];
          Arow[a * nc + j] = inpiv ? -t : Arow[a * nc + j] - t;
        }
      }
      __syncthreads();
    }
  }
  __syncthreads();
  return s_flag == 0;
}

// ---------------------------------------------------------------------------------------------
// K3: persistent cooperative PCG on the (implicit) Schur complement
//   S x = g,  S = Hpp + lambda I - Hpl (Hll + lambda I)^-1 Hlp
// preconditioned by block-Jacobi (6x6 diagonal blocks of S) plus, optionally, the rigid-body coarse
// level above (additive two-level).  One CTA per SM, each owning a contiguous pose range; 6 lanes
// per pose (one per row), 5 poses per warp; one warp per landmark in the landmark sweep.
// 3 grid-wide barriers per iteration, each fused with the all-reduce it needs; all reductions are
// fixed-order => bit-reproducible for a fixed grid.
//
// k_pcg      : generic (any size; operands streamed from L2/HBM every iteration; 3 grid barriers per iteration)
// k_pcg_flow : (ssb_pcg_flow.cuh) on-chip resident data-flow variant for graphs that fit (<= 80 poses and
//              <= 16 landmarks of degree <= 64 per CTA): Hpp/Dinv/B rows and the landmark blocks live in
//              registers, HplP / Hoff rows in shared memory for the whole solve; single-reduction CG, tagged
//              cells instead of grid barriers.
// ---------------------------------------------------------------------------------------------
#ifndef SSB_PCG_THREADS
#define SSB_PCG_THREADS 1024
#endif
constexpr int PCG_THREADS = SSB_PCG_THREADS;   // threads per CTA of the streaming k_pcg (A/B: -DSSB_PCG_THREADS=768)
constexpr int PCGF_THREADS = 512;
constexpr int PCGF_MAXPL = 448;   // pose-landmark entries cached per CTA (fast path)
constexpr int PCGF_MAXPP = 160;   // pose-pose incidences cached per CTA (fast path)
constexpr int PCGF_MAXOV = 64;    // landmark edges beyond the first 32 of a landmark, per CTA (fast path)

// middle level applied to the residual of one warp's 5 poses: returns this lane's component of
// P1 D1^-1 P1' r.  All 32 lanes must call (warp-uniform aggregate index `agg`).
__device__ __forceinline__ double sub_level_apply(const CoarseDev& Cz, int i, int agg, int comp, bool act, double rcomp) {
  const int lane = threadIdx.x & 31;
  double b1[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) b1[k] = act ? Cz.B1mat[36 * (size_t)i + 6 * comp + k] : 0.0;
  double r1[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) r1[k] = warp_sum(b1[k] * rcomp);
  double z1 = 0.0;
  if (lane < 6) {
    const double* D = Cz.D1inv + 36 * (size_t)agg + 6 * lane;
#pragma unroll
    for (int j = 0; j < 6; ++j) z1 += D[j] * r1[j];
  }
  double add = 0.0;
#pragma unroll
  for (int k = 0; k < 6; ++k) add += b1[k] * __shfl_sync(0xffffffffu, z1, k);
  return add;
}

// Sharded graphs (template parameter MR of k_pcg): what the streaming kernel writes into the other ranks' arenas.
struct StreamPeer {
  int world, rank;
  const int* zpush_rowptr;     // [Np_own + 1] ranks that hold my keyframe as a ghost ...
  double* const* zpush_z;      //   ... its 6 slots in that rank's z vector
  double* const* zpush_x;      //   ... and in its solution vector
  int n_own_lm;                // the landmarks this rank eliminates: local landmarks [0, n_own_lm)
  const int* vpush_rowptr;     // [n_own_lm + 1] ranks whose keyframes see the landmark ...
  double* const* vpush_v;      //   ... its 3 slots in that rank's v vector
  unsigned char* slots[SSB_MAX_WORLD];   // every rank's barrier slots [2][world * NB] + counter
};
constexpr int PCG_PART = PCG_THREADS > SSB_MAX_WORLD * 148 ? PCG_THREADS : SSB_MAX_WORLD * 148 + 8;

// Barrier + fixed-order all-reduce over the CTAs of EVERY rank: a CTA writes its payload slot and adds to the arrival
// counter in every rank's arena (system-scope release: everything its threads pushed to the neighbours before is visible
// there first), then waits for world * gridDim.x arrivals on its OWN counter.  The 6 values of my6 stay local (the coarse
// level is per rank).  Every CTA of every rank folds the same world * gridDim.x partials in the same order.
__device__ __forceinline__ double grid_bar_sum_mr(const StreamPeer& SP, unsigned& epoch, double my_partial, const double* my6,
                                                  double* gather, double* part_sh) {
  ++epoch;
  const int nb = gridDim.x, tot = SP.world * nb;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int r = 0; r < SP.world; ++r) {
      BarSlot* S = reinterpret_cast<BarSlot*>(SP.slots[r]) + (size_t)(epoch & 1u) * tot + SP.rank * nb + blockIdx.x;
      st_relaxed_sys_f64(&S->v[0], my_partial);
      if (my6 && r == SP.rank) {
#pragma unroll
        for (int k = 0; k < 6; ++k) S->v[1 + k] = my6[k];
      }
    }
    // ONE system-scope fence for everything this CTA pushed (cumulative through the bar.sync above), then relaxed arrivals:
    // a release per rank would drain the NVLink write queue `world` times (measured: the barrier, not the data, bound the
    // 4-GPU iteration at 127 us with red.release.sys per rank)
    __threadfence_system();
    for (int r = 0; r < SP.world; ++r)
      red_relaxed_sys_add_u32(&(reinterpret_cast<BarSlot*>(SP.slots[r]) + 2 * (size_t)tot)->epoch, 1u);
    const unsigned* counter = &(reinterpret_cast<BarSlot*>(SP.slots[SP.rank]) + 2 * (size_t)tot)->epoch;
    const unsigned target = epoch * (unsigned)tot;
    unsigned polls = 0;
    unsigned long long t0 = 0;
    while (ld_acquire_sys_u32(counter) < target) {
      if ((++polls & 1023u) == 0) {
        const unsigned long long now = peer_globaltimer();
        if (t0 == 0)
          t0 = now;
        else if (now - t0 > SSB_PEER_TIMEOUT_NS)
          __trap();
      }
    }
  }
  __syncthreads();
  const BarSlot* S = reinterpret_cast<const BarSlot*>(SP.slots[SP.rank]) + (size_t)(epoch & 1u) * tot;
  for (int k = threadIdx.x; k < tot; k += blockDim.x) part_sh[k] = __ldcg(&S[k].v[0]);
  if (gather && threadIdx.x < nb) {
    const BarSlot* o = S + SP.rank * nb + threadIdx.x;
#pragma unroll
    for (int k = 0; k < 6; ++k) gather[6 * threadIdx.x + k] = __ldcg(&o->v[1 + k]);
  }
  __syncthreads();
  double t = 0.0;
  for (int k = (threadIdx.x & 31); k < tot; k += 32) t += part_sh[k];
  return warp_sum(t);
}

// MR: this grid = the CTAs of rank SP.rank of a sharded graph.  The vectors a neighbour reads (z of boundary keyframes, v
// of shared landmarks, finally x) are also written into its arena before the barrier that publishes them; the three
// barriers of an iteration span all ranks.  The preconditioner is rank-local (ghost keyframes carry no basis).
// The streaming sweeps are chains of dependent gathers with ~640 B in flight per warp.  Requesting the operands of the NEXT
// pass of a warp into L2 one pass ahead (prefetch.global.L2; -DSSB_PCG_PREFETCH=1) was measured on cfg4: 170.1 ms against
// 163.5 ms without (1 037 PCG iterations) — the static operands are not what the sweeps wait for.  Off by default.
#ifndef SSB_PCG_PREFETCH
#define SSB_PCG_PREFETCH 0
#endif
__device__ __forceinline__ void prefetch_l2(const void* p) {
#if SSB_PCG_PREFETCH
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
}
template <bool MR>
__global__ void __launch_bounds__(PCG_THREADS, 1)
    k_pcg(DevGraph G, CoarseDev Cz, BarSlot* slots, double lambda, double tol2, int maxit, StreamPeer SP) {
  extern __shared__ __align__(16) double dsm[];
  __shared__ double sh[33];
  __shared__ double s6[8], zc6[8], red6[6 * 32];
  const int nblk = gridDim.x;
  const int nc = 6 * nblk;
  unsigned epoch_g = 0;
#define SSB_GBAR(partial_, my6_, gather_) \
  (MR ? grid_bar_sum_mr(SP, epoch_g, (partial_), (my6_), (gather_), part_sh) : grid_bar_sum(slots, epoch, (partial_), (my6_), (gather_), part_sh))
  double* part_sh = dsm;                 // [PCG_PART]
  double* Arow = part_sh + PCG_PART;     // [6][nc]   rows of A_c, then of A_c^-1
  double* panel_sh = Arow + 6 * nc;      // [6][nc]
  double* rc = panel_sh + 6 * nc;        // [nc] restricted residual (kept by recurrence)
  double* qc = rc + nc;                  // [nc] gathered restricted q
  double* red = qc + nc;                 // [28][36]
  const bool coarse = Cz.enabled != 0;

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int warps_per_block = blockDim.x >> 5;
  const int total_warps = nblk * warps_per_block;
  const int gw = warp * nblk + blockIdx.x;  // landmark sweep: round-robin over CTAs
  const int slot = lane / 6, comp = lane - 6 * slot;
  const bool lane_active = lane < 30;
  const int base_lane = 6 * slot;
  const int p0 = blockIdx.x * Cz.C, p1 = min(G.Np_own, p0 + Cz.C);
  unsigned epoch = 0;
  double* pold = G.p0;
  double* pnew = G.p1;
  int status = 0;

  const bool use_sub = Cz.sub_enabled != 0;
  bool use_coarse = false;
  if (coarse) {
    if (Cz.reuse_inverse) {
      for (int k = threadIdx.x; k < 6 * nc; k += PCG_THREADS) Arow[k] = Cz.ainv_store[(size_t)blockIdx.x * 6 * nc + k];
      use_coarse = true;
      __syncthreads();
    } else {
      use_coarse = coarse_prologue<PCG_THREADS>(G, Cz, slots, epoch, lambda, Arow, panel_sh, red, part_sh, p0, p1);
      if (use_coarse)
        for (int k = threadIdx.x; k < 6 * nc; k += PCG_THREADS) Cz.ainv_store[(size_t)blockIdx.x * 6 * nc + k] = Arow[k];
    }
  }

  // ---- init: x = 0, r = g, z = M^-1 r
  double local = 0.0;
  double l6[6] = {0, 0, 0, 0, 0, 0};
  for (int pbase = p0 + warp * 5; pbase < p1; pbase += warps_per_block * 5) {
    const int i = pbase + slot;
    const bool act = lane_active && i < p1;
    double rcomp = act ? G.g[6 * (size_t)i + comp] : 0.0;
    if (act) {
      G.x[6 * (size_t)i + comp] = 0.0;
      G.r[6 * (size_t)i + comp] = rcomp;
      pold[6 * (size_t)i + comp] = 0.0;
      if (use_coarse) {
        const double* B = Cz.Bmat + 36 * (size_t)i + 6 * comp;
#pragma unroll
        for (int k = 0; k < 6; ++k) l6[k] += B[k] * rcomp;
      }
    }
  }
  if (use_coarse) {
    block_sum6(l6, s6, red6);
    grid_bar_sum(slots, epoch, 0.0, s6, rc, part_sh);  // rc = P' r  (all aggregates)
    if (warp < 6) {
      double t = 0.0;
      for (int j = lane; j < nc; j += 32) t += Arow[warp * nc + j] * rc[j];
      t = warp_sum(t);
      if (lane == 0) zc6[warp] = t;
    }
    __syncthreads();
  }
  for (int pbase = p0 + warp * 5; pbase < p1; pbase += warps_per_block * 5) {
    const int i = pbase + slot;
    const bool act = lane_active && i < p1;
    double rcomp = act ? G.r[6 * (size_t)i + comp] : 0.0;
    double zc = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      double rk = __shfl_sync(0xffffffffu, rcomp, base_lane + k);
      if (act) zc += G.Dinv[36 * (size_t)i + 6 * comp + k] * rk;
    }
    if (use_sub) zc += sub_level_apply(Cz, i, pbase / 5, comp, act, rcomp);
    if (act) {
      if (use_coarse) {
        const double* B = Cz.Bmat + 36 * (size_t)i + 6 * comp;
#pragma unroll
        for (int k = 0; k < 6; ++k) zc += B[k] * zc6[k];
      }
      G.z[6 * (size_t)i + comp] = zc;
      if constexpr (MR)
        for (int q = SP.zpush_rowptr[i]; q < SP.zpush_rowptr[i + 1]; ++q) st_relaxed_sys_f64(SP.zpush_z[q] + comp, zc);
      local += rcomp * zc;
    }
  }
  if constexpr (MR)   // p of a ghost keyframe is rebuilt here from its z (pushed by the owner): p_0 = 0
    for (int k = 6 * G.Np_own + blockIdx.x * blockDim.x + threadIdx.x; k < 6 * G.Np; k += gridDim.x * blockDim.x) pold[k] = 0.0;
  double bs = block_sum(local, sh);
  double rz = SSB_GBAR(bs, nullptr, nullptr);
  const double rz0 = rz;
  double beta = 0.0;
  int it = 0;
  if (!(rz0 > 0.0)) {
    status = (rz0 == 0.0) ? 0 : 2;
    maxit = 0;
  }
  for (it = 0; it < maxit; ++it) {
    // ---- phase 1: v_l = (Hll+lambda)^-1 sum_e HplL_e p_e,  p = z + beta * pold (on the fly)
    for (int l = gw; l < (MR ? SP.n_own_lm : G.Nl); l += total_warps) {   // a shard's owned landmarks come first
      double a0 = 0.0, a1 = 0.0, a2 = 0.0;
      const int e1 = G.lm_rowptr[l + 1];
      const int ln = l + total_warps;
      const int en = (SSB_PCG_PREFETCH && ln < (MR ? SP.n_own_lm : G.Nl)) ? G.lm_rowptr[ln] : -1;   // consumed after the sweep below
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
      if (en >= 0 && en + lane < G.El) {   // next landmark of this warp: one edge per lane (block 144 B, record 80 B)
        prefetch_l2(G.HplL + 18 * (size_t)(en + lane));
        prefetch_l2(G.HplL + 18 * (size_t)(en + lane) + 16);
        prefetch_l2(G.pl + (en + lane));
      }
      a0 = warp_sum(a0);
      a1 = warp_sum(a1);
      a2 = warp_sum(a2);
      if (lane == 0) {
        const double* Wi = G.HllInv + 6 * (size_t)l;
        const double v0 = Wi[0] * a0 + Wi[1] * a1 + Wi[2] * a2;
        const double v1 = Wi[1] * a0 + Wi[3] * a1 + Wi[4] * a2;
        const double v2 = Wi[2] * a0 + Wi[4] * a1 + Wi[5] * a2;
        G.v[3 * (size_t)l + 0] = v0;
        G.v[3 * (size_t)l + 1] = v1;
        G.v[3 * (size_t)l + 2] = v2;
        if constexpr (MR)
          for (int q = SP.vpush_rowptr[l]; q < SP.vpush_rowptr[l + 1]; ++q) {
            st_relaxed_sys_f64(SP.vpush_v[q] + 0, v0);
            st_relaxed_sys_f64(SP.vpush_v[q] + 1, v1);
            st_relaxed_sys_f64(SP.vpush_v[q] + 2, v2);
          }
      }
    }
    SSB_GBAR(0.0, nullptr, nullptr);
    // ---- phase 2: q = (Hpp + lambda) p + sum Hoff p_nbr - sum HplP v ; partial p.q ; restricted q
    local = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) l6[k] = 0.0;
    for (int pbase = p0 + warp * 5; pbase < p1; pbase += warps_per_block * 5) {
      const int i = pbase + slot;
      const bool act = lane_active && i < p1;
      const int inx = i + warps_per_block * 5;   // my keyframe of the next pass
      const bool pfn = SSB_PCG_PREFETCH && lane_active && inx < p1;
      int qn = 0;
      if (pfn) {
        prefetch_l2(G.Hpp + 36 * (size_t)inx + 6 * comp);
        if (use_coarse) prefetch_l2(Cz.Bmat + 36 * (size_t)inx + 6 * comp);
        qn = G.pose_pl_rowptr[inx];   // consumed at the end of this pass
      }
      double pc = 0.0;
      if (act) {
        pc = __ldcg(G.z + 6 * (size_t)i + comp) + beta * __ldcg(pold + 6 * (size_t)i + comp);
        pnew[6 * (size_t)i + comp] = pc;
      }
      double qv = lambda * pc;
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        double pk = __shfl_sync(0xffffffffu, pc, base_lane + k);
        if (act) qv += G.Hpp[36 * (size_t)i + 6 * comp + k] * pk;
      }
      int k0 = 0, k1 = 0, q0 = 0, q1 = 0;
      if (act) {
        k0 = G.pose_pp_rowptr[i];
        k1 = G.pose_pp_rowptr[i + 1];
        q0 = G.pose_pl_rowptr[i];   // (requested here so that the landmark chain below overlaps the pose-pose chain)
        q1 = G.pose_pl_rowptr[i + 1];
      }
      int lm0[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) lm0[u] = (act && q0 + u < q1) ? G.plP_lm[q0 + u] : -1;
      int nmax = k1 - k0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));
      // pose-pose incidences two at a time: both neighbour vectors and both block rows are in flight together, and the
      // neighbour index comes from the incidence table instead of the edge record (one dependent load less per chain)
      for (int s = 0; s < nmax; s += 2) {
        const bool h0 = act && (k0 + s < k1), h1 = act && (k0 + s + 1 < k1);
        int c0 = 0, c1 = 0, o0 = 0, o1 = 0;
        if (h0) {
          c0 = G.pose_pp_idx[k0 + s];
          o0 = G.pose_pp_other[k0 + s];
        }
        if (h1) {
          c1 = G.pose_pp_idx[k0 + s + 1];
          o1 = G.pose_pp_other[k0 + s + 1];
        }
        double oc0 = 0.0, oc1 = 0.0, r0[6], r1[6];
        if (h0) oc0 = __ldcg(G.z + 6 * (size_t)o0 + comp) + beta * __ldcg(pold + 6 * (size_t)o0 + comp);
        if (h1) oc1 = __ldcg(G.z + 6 * (size_t)o1 + comp) + beta * __ldcg(pold + 6 * (size_t)o1 + comp);
        {
          const double* H0 = G.Hoff + 36 * (size_t)(c0 >> 1);
          const double* H1 = G.Hoff + 36 * (size_t)(c1 >> 1);
#pragma unroll
          for (int k = 0; k < 6; ++k) {
            r0[k] = h0 ? ((c0 & 1) == 0 ? H0[6 * comp + k] : H0[6 * k + comp]) : 0.0;
            r1[k] = h1 ? ((c1 & 1) == 0 ? H1[6 * comp + k] : H1[6 * k + comp]) : 0.0;
          }
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) qv += r0[k] * __shfl_sync(0xffffffffu, oc0, base_lane + k);
#pragma unroll
        for (int k = 0; k < 6; ++k) qv += r1[k] * __shfl_sync(0xffffffffu, oc1, base_lane + k);
      }
      if (act) {
        // the pose's landmark edges in batches of 4: the landmark ids, then every v triple and block row are requested
        // together (one dependent-load round trip per batch instead of two per edge)
        for (int kk = q0; kk < q1; kk += 4) {
          int lm[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) lm[u] = kk == q0 ? lm0[u] : (kk + u < q1 ? G.plP_lm[kk + u] : -1);
          double hv[4][3], vv[4][3];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const double* Hp = G.HplP + 18 * (size_t)(kk + u) + 3 * comp;
            const double* vp = G.v + 3 * (size_t)(lm[u] < 0 ? 0 : lm[u]);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              hv[u][c] = lm[u] >= 0 ? Hp[c] : 0.0;
              vv[u][c] = lm[u] >= 0 ? __ldcg(vp + c) : 0.0;
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) qv -= hv[u][0] * vv[u][0] + hv[u][1] * vv[u][1] + hv[u][2] * vv[u][2];
        }
        G.q[6 * (size_t)i + comp] = qv;
        local += pc * qv;
        if (use_coarse) {
          const double* B = Cz.Bmat + 36 * (size_t)i + 6 * comp;
#pragma unroll
          for (int k = 0; k < 6; ++k) l6[k] += B[k] * qv;
        }
      }
      if (pfn) {   // the pose-landmark blocks of the next keyframe (6 lanes x 128 B) and, for odometry chains, its two Hoff blocks
        prefetch_l2(G.HplP + 18 * (size_t)qn + 16 * comp);
        prefetch_l2(G.Hoff + 36 * (size_t)min(inx, max(G.Epp - 1, 0)) + 6 * comp);
        prefetch_l2(G.Hoff + 36 * (size_t)max(min(inx, G.Epp) - 1, 0) + 6 * comp);
      }
    }
    if constexpr (MR)   // p of the ghost keyframes (read by phase 1 / 2 of the next iteration through pold)
      for (int k = 6 * G.Np_own + blockIdx.x * blockDim.x + threadIdx.x; k < 6 * G.Np; k += gridDim.x * blockDim.x)
        pnew[k] = __ldcg(G.z + k) + beta * __ldcg(pold + k);
    bs = block_sum(local, sh);
    if (use_coarse) block_sum6(l6, s6, red6);
    const double pq = SSB_GBAR(bs, use_coarse ? s6 : nullptr, use_coarse ? qc : nullptr);
    if (!(pq > 0.0) || !isfinite(pq)) {  // breakdown: S not positive definite / non-finite data
      status = 1;
      break;
    }
    const double alpha = rz / pq;
    // ---- phase 3: x += alpha p ; r -= alpha q ; z = M^-1 r ; partial r.z
    if (use_coarse) {
      for (int j = threadIdx.x; j < nc; j += blockDim.x) rc[j] -= alpha * qc[j];
      __syncthreads();
      if (warp < 6) {
        double t = 0.0;
        for (int j = lane; j < nc; j += 32) t += Arow[warp * nc + j] * rc[j];
        t = warp_sum(t);
        if (lane == 0) zc6[warp] = t;
      }
      __syncthreads();
    }
    local = 0.0;
    for (int pbase = p0 + warp * 5; pbase < p1; pbase += warps_per_block * 5) {
      const int i = pbase + slot;
      const bool act = lane_active && i < p1;
      if (SSB_PCG_PREFETCH && lane_active && i + warps_per_block * 5 < p1) {
        const size_t on = 36 * (size_t)(i + warps_per_block * 5) + 6 * comp;
        prefetch_l2(G.Dinv + on);
        if (use_sub) prefetch_l2(Cz.B1mat + on);
        if (use_coarse) prefetch_l2(Cz.Bmat + on);
      }
      double rcomp = 0.0;
      if (act) {
        const size_t o = 6 * (size_t)i + comp;
        G.x[o] += alpha * pnew[o];
        rcomp = G.r[o] - alpha * G.q[o];
        G.r[o] = rcomp;
      }
      double zc = 0.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        double rk = __shfl_sync(0xffffffffu, rcomp, base_lane + k);
        if (act) zc += G.Dinv[36 * (size_t)i + 6 * comp + k] * rk;
      }
      if (use_sub) zc += sub_level_apply(Cz, i, pbase / 5, comp, act, rcomp);
      if (act) {
        if (use_coarse) {
          const double* B = Cz.Bmat + 36 * (size_t)i + 6 * comp;
#pragma unroll
          for (int k = 0; k < 6; ++k) zc += B[k] * zc6[k];
        }
        G.z[6 * (size_t)i + comp] = zc;
        if constexpr (MR)
          for (int q = SP.zpush_rowptr[i]; q < SP.zpush_rowptr[i + 1]; ++q) st_relaxed_sys_f64(SP.zpush_z[q] + comp, zc);
        local += rcomp * zc;
      }
    }
    bs = block_sum(local, sh);
    const double rzn = SSB_GBAR(bs, nullptr, nullptr);
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
  if constexpr (MR) {   // the solution of a boundary keyframe goes to every rank that updates it as a ghost
    for (int pbase = p0 + warp * 5; pbase < p1; pbase += warps_per_block * 5) {
      const int i = pbase + slot;
      if (lane_active && i < p1) {
        const double xv = G.x[6 * (size_t)i + comp];
        for (int q = SP.zpush_rowptr[i]; q < SP.zpush_rowptr[i + 1]; ++q) st_relaxed_sys_f64(SP.zpush_x[q] + comp, xv);
      }
    }
  }
#undef SSB_GBAR
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    G.iscalars[0] = it;
    G.iscalars[1] = status;
    G.scalars[3] = rz;
    G.scalars[4] = rz0;
  }
}

// ---------------------------------------------------------------------------------------------
// Sharded graphs: the RANK-LEVEL coarse level of the preconditioner.  Every rank is one more rigid-body aggregate (6
// unknowns: translation + rotation of all its keyframes about their centroid c_r); A_g = P_g' S P_g is 6W x 6W.  It is the
// only level that couples the ranks (the per-CTA coarse matrix and the groups see S restricted to the rank), and what
// keeps the iteration count from growing with the number of ranks (scripts/shard_precond_study.py: 196 -> 105 iterations
// per solve on 2 ranks, 409 -> 170 on 8).  P_g = P_c E with E_a = [I, -[c_a - c_r]x; 0, I] for CTA aggregate a of rank r, so
// inside the PCG kernel the level rides on the per-CTA coarse values: P_g'w = sum_a E_a' (P_c'w)_a, prolongation
// B_i (E_a z_g).  Assembly: rank r computes its 6 rows of A_g from its own keyframes (a touched landmark has all its edges
// here, ghosts included), pushes them to everybody, every rank inverts the 6W x 6W matrix redundantly.
// ---------------------------------------------------------------------------------------------
struct GlobDev {
  int world, rank;
  const int* pose_rank;   // [Np] owner of every local keyframe
  double* gcent;          // [world][4] centroids of all ranks (in my arena; slot s written by rank s)
  double* Bg;             // [Np][36] prolongation blocks about the OWNER's centroid
  double* Gg;             // [Nl][world][18] sum over the landmark's edges whose keyframe belongs to rank s of HplL_e Bg_p(e)
  double* part;           // [G_ROW_BLOCKS][world * 36]
  double* grows;          // [world][36 * world] row blocks of A_g (in my arena; slot s written by rank s)
  float* Aginv;           // [6][6 world] my rows of w_g A_g^-1 (zeros: level off)
};
constexpr int G_ROW_BLOCKS = 148;
constexpr int G_ROW_THREADS = 36 * SSB_MAX_WORLD;

// own centroid -> slot `rank` of every rank's gcent (fixed-order block reduction)
__global__ void __launch_bounds__(1024) k_g_centroid(DevGraph G, int world, int rank, double* const* gcent_all) {
  __shared__ double sh[33];
  double s[3] = {0, 0, 0};
  for (int i = threadIdx.x; i < G.Np_own; i += blockDim.x)
    for (int k = 0; k < 3; ++k) s[k] += G.pose[i].t[k];
  double c[3];
  for (int k = 0; k < 3; ++k) c[k] = block_sum(s[k], sh) / (double)max(1, G.Np_own);
  if (threadIdx.x < world)
    for (int k = 0; k < 3; ++k) st_relaxed_sys_f64(gcent_all[threadIdx.x] + 4 * rank + k, c[k]);
}
__global__ void __launch_bounds__(128) k_g_basis(DevGraph G, GlobDev Gd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= G.Np) return;
  double* B = Gd.Bg + 36 * (size_t)i;
  if (G.pose_fixed[i] || (G.pose_kind && G.pose_kind[i])) {
    for (int k = 0; k < 36; ++k) B[k] = 0.0;
    return;
  }
  const double* c = Gd.gcent + 4 * Gd.pose_rank[i];
  const Pose X = G.pose[i];
  double R[9];
  quat_to_R(X.q, R);
  const double d[3] = {X.t[0] - c[0], X.t[1] - c[1], X.t[2] - c[2]};
  const double Sx[9] = {0, -d[2], d[1], d[2], 0, -d[0], -d[1], d[0], 0};
  for (int r = 0; r < 3; ++r)
    for (int cc = 0; cc < 3; ++cc) {
      const double rt = R[3 * cc + r];
      B[6 * r + cc] = rt;
      B[6 * r + 3 + cc] = -(R[r] * Sx[cc] + R[3 + r] * Sx[3 + cc] + R[6 + r] * Sx[6 + cc]);
      B[6 * (3 + r) + cc] = 0.0;
      B[6 * (3 + r) + 3 + cc] = 0.5 * rt;
    }
}
__global__ void __launch_bounds__(128) k_g_runs(DevGraph G, GlobDev Gd) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int q = t % 18, ls = t / 18, s = ls % Gd.world, l = ls / Gd.world;
  if (l >= G.Nl) return;
  const int a = q / 6, c = q - 6 * a;
  double acc = 0.0;
  for (int e = G.lm_rowptr[l]; e < G.lm_rowptr[l + 1]; ++e) {
    const int p = G.pl[e].p;
    if (Gd.pose_rank[p] != s) continue;
    const double* Hl = G.HplL + 18 * (size_t)e + 6 * a;
    const double* B = Gd.Bg + 36 * (size_t)p + c;
#pragma unroll
    for (int k = 0; k < 6; ++k) acc += Hl[k] * B[6 * k];
  }
  Gd.Gg[(size_t)ls * 18 + q] = acc;
}
// per damped trial: my 6 rows of A_g, partial sums over chunks of own keyframes (thread = (column rank s, entry r, c))
__global__ void __launch_bounds__(G_ROW_THREADS) k_g_rows(DevGraph G, GlobDev Gd, double lambda) {
  const int s = threadIdx.x / 36, ent = threadIdx.x - 36 * s, r = ent / 6, c = ent - 6 * r;
  const int per = (G.Np_own + gridDim.x - 1) / gridDim.x;
  const int i0 = min(G.Np_own, (int)blockIdx.x * per), i1 = min(G.Np_own, i0 + per);
  double acc = 0.0;
  if (s < Gd.world)
    for (int i = i0; i < i1; ++i) {
      const double* Bi = Gd.Bg + 36 * (size_t)i;
      if (s == Gd.rank) {
        const double* H = G.Hpp + 36 * (size_t)i;
        double t = 0.0;
        for (int a = 0; a < 6; ++a) {
          double hb = lambda * Bi[6 * a + c];
          for (int b = 0; b < 6; ++b) hb += H[6 * a + b] * Bi[6 * b + c];
          t += Bi[6 * a + r] * hb;
        }
        acc += t;
      }
      for (int kk = G.pose_pp_rowptr[i]; kk < G.pose_pp_rowptr[i + 1]; ++kk) {
        const int code = G.pose_pp_idx[kk], e = code >> 1, role = code & 1;
        const int j = role == 0 ? G.pp[e].j : G.pp[e].i;
        if (Gd.pose_rank[j] != s) continue;
        const double* Bj = Gd.Bg + 36 * (size_t)j;
        const double* Ho = G.Hoff + 36 * (size_t)e;
        double t = 0.0;
        for (int a = 0; a < 6; ++a) {
          double hb = 0.0;
          for (int b = 0; b < 6; ++b) hb += (role == 0 ? Ho[6 * a + b] : Ho[6 * b + a]) * Bj[6 * b + c];
          t += Bi[6 * a + r] * hb;
        }
        acc += t;
      }
      for (int kk = G.pose_pl_rowptr[i]; kk < G.pose_pl_rowptr[i + 1]; ++kk) {
        const int l = G.plP_lm[kk];
        const double* Hp = G.HplP + 18 * (size_t)kk;
        const double* Wu = G.HllInv + 6 * (size_t)l;
        const double* Gs = Gd.Gg + ((size_t)l * Gd.world + s) * 18;
        // X = W Gs (column c), then Bi' Hp X
        const double x0 = Wu[0] * Gs[c] + Wu[1] * Gs[6 + c] + Wu[2] * Gs[12 + c];
        const double x1 = Wu[1] * Gs[c] + Wu[3] * Gs[6 + c] + Wu[4] * Gs[12 + c];
        const double x2 = Wu[2] * Gs[c] + Wu[4] * Gs[6 + c] + Wu[5] * Gs[12 + c];
        double t = 0.0;
        for (int a = 0; a < 6; ++a) t += Bi[6 * a + r] * (Hp[3 * a] * x0 + Hp[3 * a + 1] * x1 + Hp[3 * a + 2] * x2);
        acc -= t;
      }
    }
  if (s < Gd.world) Gd.part[(size_t)blockIdx.x * (36 * Gd.world) + 36 * s + ent] = acc;
}
// fold the partials in block order and hand my row block to every rank (slot `rank` of their grows)
__global__ void __launch_bounds__(G_ROW_THREADS) k_g_fold(GlobDev Gd, int nblocks, double* const* grows_all) {
  const int n = 36 * Gd.world;
  if ((int)threadIdx.x >= n) return;
  double t = 0.0;
  for (int b = 0; b < nblocks; ++b) t += Gd.part[(size_t)b * n + threadIdx.x];
  for (int q = 0; q < Gd.world; ++q) st_relaxed_sys_f64(grows_all[q] + (size_t)Gd.rank * n + threadIdx.x, t);
}
// every rank: A_g from the gathered row blocks, symmetrised, Gauss-Jordan inverse (SPD, no pivoting), my 6 rows as float
__global__ void __launch_bounds__(256) k_g_invert(GlobDev Gd, float weight) {
  constexpr int NMAX = 6 * SSB_MAX_WORLD;
  __shared__ double A[NMAX][NMAX + 1], I[NMAX][NMAX + 1];
  __shared__ int s_fail;
  const int n = 6 * Gd.world;
  if (threadIdx.x == 0) s_fail = 0;
  for (int k = threadIdx.x; k < n * n; k += blockDim.x) {
    const int i = k / n, j = k - i * n;
    // row block of rank ri: entry (r, 6 s + c) stored at [ri][36 s + 6 r + c]
    auto at = [&](int ii, int jj) { return Gd.grows[(size_t)(ii / 6) * 36 * Gd.world + 36 * (jj / 6) + 6 * (ii % 6) + (jj % 6)]; };
    A[i][j] = 0.5 * (at(i, j) + at(j, i));
    I[i][j] = i == j ? 1.0 : 0.0;
  }
  __syncthreads();
  for (int k = 0; k < n; ++k) {
    const double piv = A[k][k];
    if (!(piv > 0.0) || !isfinite(piv)) {
      if (threadIdx.x == 0) s_fail = 1;
      __syncthreads();
      break;
    }
    const double ip = 1.0 / piv;
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
      A[k][j] *= ip;
      I[k][j] *= ip;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
      const int i = idx / n, j = idx - i * n;
      if (i == k) continue;
      const double f = A[i][k];
      if (j != k) A[i][j] -= f * A[k][j];
      I[i][j] -= f * I[k][j];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x)
      if (i != k) A[i][k] = 0.0;
    __syncthreads();
  }
  __syncthreads();
  for (int k = threadIdx.x; k < 6 * n; k += blockDim.x) {
    const int r = k / n, j = k - r * n;
    const int i = 6 * Gd.rank + r;
    Gd.Aginv[k] = s_fail ? 0.0f : (float)(weight * 0.5 * (I[i][j] + I[j][i]));
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
  // threads [0, 32 Nl): one warp per landmark (lane per edge); threads [32 Nl, 32 Nl + Np): one thread per pose
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  double sc = 0.0;
  if (t < 32 * G.Nl) {
    const int l = t >> 5, lane = t & 31;
    double a[3] = {0.0, 0.0, 0.0};
    for (int e = G.lm_rowptr[l] + lane; e < G.lm_rowptr[l + 1]; e += 32) {
      const double* Hl = G.HplL + 18 * (size_t)e;
      const double* dp = G.x + 6 * (size_t)G.pl[e].p;
      for (int c = 0; c < 6; ++c) {
        a[0] -= Hl[c] * dp[c];
        a[1] -= Hl[6 + c] * dp[c];
        a[2] -= Hl[12 + c] * dp[c];
      }
    }
    for (int k = 0; k < 3; ++k) a[k] = warp_sum(a[k]) + G.bl[3 * (size_t)l + k];
    if (lane == 0) {
      const double* Wi = G.HllInv + 6 * (size_t)l;
      double d[3] = {Wi[0] * a[0] + Wi[1] * a[1] + Wi[2] * a[2], Wi[1] * a[0] + Wi[3] * a[1] + Wi[4] * a[2],
                     Wi[2] * a[0] + Wi[4] * a[1] + Wi[5] * a[2]};
      if (G.lm_fixed[l]) d[0] = d[1] = d[2] = 0.0;
      double cur[4];
      for (int c = 0; c < 4; ++c) {
        cur[c] = G.lm[4 * (size_t)l + c];
        lm_bak[4 * (size_t)l + c] = cur[c];
      }
      if (pl_is_plane(G, l)) {
        if (!G.lm_fixed[l]) plane_oplus(cur, d);  // VertexPlane::oplusImpl
      } else {
        for (int c = 0; c < 3; ++c) cur[c] += d[c];  // VertexPointXYZ::oplusImpl
      }
      for (int c = 0; c < 4; ++c) G.lm[4 * (size_t)l + c] = cur[c];
      const bool mine = G.lm_owned == nullptr || G.lm_owned[l] != 0;   // a shared landmark is counted by its owner
      for (int c = 0; c < 3; ++c) {
        G.dl[3 * (size_t)l + c] = d[c];
        if (mine) sc += d[c] * (lambda * d[c] + G.bl[3 * (size_t)l + c]);
      }
    }
  } else if (t < 32 * G.Nl + G.Np) {
    const int i = t - 32 * G.Nl;
    Pose X = G.pose[i];
    pose_bak[i] = X;
    double d[6];
    for (int c = 0; c < 6; ++c) d[c] = G.pose_fixed[i] ? 0.0 : G.x[6 * (size_t)i + c];
    if (!G.pose_fixed[i]) {
      if (G.pose_kind && G.pose_kind[i]) {   // VertexPointXYZ::oplusImpl of a promoted landmark
        for (int c = 0; c < 3; ++c) X.t[c] += d[c];
      } else
        pose_oplus(X, d);
      G.pose[i] = X;
    }
    if (i < G.Np_own)
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
        double err[3];
        pl_edge_err(G, *ed, e0 + tid, X, err);
        acc += quad3(ed->info, err);
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
        pp_edge_linearize(*ed, Xi, Xj, err, nullptr, nullptr);
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

// ---------------------------------------------------------------------------------------------
// K5: landmark marginals  (H^-1)[l,l] = W + W Hlp S^-1 Hpl W  with W = Hll_l^-1, S the undamped Schur
// complement (GraphSLAM::computeLandmarkMarginals, graph_slam.cpp:221-234).  One PCG solve per column.
// ---------------------------------------------------------------------------------------------
__global__ void k_marg_rhs(DevGraph G, int l, int c) {
  const int n = 6 * G.Np;
  for (int k = threadIdx.x; k < n; k += blockDim.x) G.g[k] = 0.0;
  __syncthreads();
  const double* Wu = G.HllInv + 6 * (size_t)l;
  const double W[9] = {Wu[0], Wu[1], Wu[2], Wu[1], Wu[3], Wu[4], Wu[2], Wu[4], Wu[5]};
  const double w[3] = {W[c], W[3 + c], W[6 + c]};  // W e_c
  for (int e = G.lm_rowptr[l] + threadIdx.x; e < G.lm_rowptr[l + 1]; e += blockDim.x) {
    const double* Hl = G.HplL + 18 * (size_t)e;  // 3x6 = Hpl_e'
    const int p = G.pl[e].p;
    for (int k = 0; k < 6; ++k) atomicAdd(G.g + 6 * (size_t)p + k, Hl[k] * w[0] + Hl[6 + k] * w[1] + Hl[12 + k] * w[2]);
  }
}
__global__ void k_marg_out(DevGraph G, int l, int c, double* out9) {
  __shared__ double sh[33];
  double a[3] = {0, 0, 0};
  for (int e = G.lm_rowptr[l] + threadIdx.x; e < G.lm_rowptr[l + 1]; e += blockDim.x) {
    const double* Hl = G.HplL + 18 * (size_t)e;
    const double* x = G.x + 6 * (size_t)G.pl[e].p;
    for (int k = 0; k < 6; ++k) {
      a[0] += Hl[k] * x[k];
      a[1] += Hl[6 + k] * x[k];
      a[2] += Hl[12 + k] * x[k];
    }
  }
  for (int k = 0; k < 3; ++k) a[k] = block_sum(a[k], sh);
  if (threadIdx.x == 0) {
    const double* Wu = G.HllInv + 6 * (size_t)l;
    const double W[9] = {Wu[0], Wu[1], Wu[2], Wu[1], Wu[3], Wu[4], Wu[2], Wu[4], Wu[5]};
    for (int r = 0; r < 3; ++r) out9[3 * r + c] = W[3 * r + c] + W[3 * r] * a[0] + W[3 * r + 1] * a[1] + W[3 * r + 2] * a[2];
  }
}

// K5 on graphs that fill only part of the chip: the graph is laid out gridDim.x times side by side (marginals_replicated in
// ssb_graph.cu; k_pcg_flow<.., REP> runs one conjugate-gradient recurrence per copy) and every copy carries its OWN
// right-hand side, so one PCG launch delivers gridDim.x columns of the marginals.  Block j works on copy j:
// keyframes [j stride_p, (j + 1) stride_p) (the copy's keyframes, then fixed padding up to a CTA boundary), landmarks
// [j Nl1, (j + 1) Nl1).
#define SSB_MARG_MAX_REP 16
struct MargCols {
  int l[SSB_MARG_MAX_REP];   // landmark (index inside one copy) of the column solved by copy j; < 0: idle copy
  int c[SSB_MARG_MAX_REP];   // column of that landmark's 3x3 block
  int o[SSB_MARG_MAX_REP];   // position of the landmark in the caller's list
};
__global__ void k_marg_rhs_rep(DevGraph G, MargCols mc, int stride_p, int Nl1) {
  const int j = blockIdx.x;
  double* gj = G.g + 6 * (size_t)j * stride_p;
  for (int k = threadIdx.x; k < 6 * stride_p; k += blockDim.x) gj[k] = 0.0;
  __syncthreads();
  if (mc.l[j] < 0) return;
  const int l = mc.l[j] + j * Nl1, c = mc.c[j];
  const double* Wu = G.HllInv + 6 * (size_t)l;
  const double W[9] = {Wu[0], Wu[1], Wu[2], Wu[1], Wu[3], Wu[4], Wu[2], Wu[4], Wu[5]};
  const double w[3] = {W[c], W[3 + c], W[6 + c]};  // W e_c
  for (int e = G.lm_rowptr[l] + threadIdx.x; e < G.lm_rowptr[l + 1]; e += blockDim.x) {
    const double* Hl = G.HplL + 18 * (size_t)e;
    const int p = G.pl[e].p;
    for (int k = 0; k < 6; ++k) atomicAdd(G.g + 6 * (size_t)p + k, Hl[k] * w[0] + Hl[6 + k] * w[1] + Hl[12 + k] * w[2]);
  }
}
// status: sticky flag, set when the solve that just finished reported a breakdown (G.iscalars[1])
__global__ void k_marg_out_rep(DevGraph G, MargCols mc, int Nl1, double* out9n, double* status) {
  __shared__ double sh[33];
  const int j = blockIdx.x;
  if (j == 0 && threadIdx.x == 0 && G.iscalars[1] != 0) status[0] = 1.0;
  if (mc.l[j] < 0) return;
  const int l = mc.l[j] + j * Nl1, c = mc.c[j];
  double a[3] = {0, 0, 0};
  for (int e = G.lm_rowptr[l] + threadIdx.x; e < G.lm_rowptr[l + 1]; e += blockDim.x) {
    const double* Hl = G.HplL + 18 * (size_t)e;
    const double* x = G.x + 6 * (size_t)G.pl[e].p;
    for (int k = 0; k < 6; ++k) {
      a[0] += Hl[k] * x[k];
      a[1] += Hl[6 + k] * x[k];
      a[2] += Hl[12 + k] * x[k];
    }
  }
  for (int k = 0; k < 3; ++k) a[k] = block_sum(a[k], sh);
  if (threadIdx.x == 0) {
    const double* Wu = G.HllInv + 6 * (size_t)l;
    const double W[9] = {Wu[0], Wu[1], Wu[2], Wu[1], Wu[3], Wu[4], Wu[2], Wu[4], Wu[5]};
    double* out9 = out9n + 9 * (size_t)mc.o[j];
    for (int r = 0; r < 3; ++r) out9[3 * r + c] = W[3 * r + c] + W[3 * r] * a[0] + W[3 * r + 1] * a[1] + W[3 * r + 2] * a[2];
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
    pp_edge_linearize(*ed, G.pose[ed->i], G.pose[ed->j], err, Ji, Jj);
  } else {
    const PLEdge ed = G.pl[e];
    double JlT[9];
    pl_edge_lin(G, ed, e, G.pose[ed.p], err, Ji, JlT);
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) Jj[3 * r + c] = JlT[3 * c + r];
  }
}

}  // namespace ssb
