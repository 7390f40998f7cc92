// K5, direct form: GraphSLAM::computeLandmarkMarginals (graph_slam.cpp:221-234) WITHOUT one iterative solve per column.
//
// The reference's graph is an odometry chain (add_se3_edge is only called for consecutive keyframes,
// semantic_graph_slam.cpp:120-135) plus pose -> landmark edges, so H_pp is block tridiagonal (6 x 6 blocks) and H_ll block
// diagonal.  Eliminating the POSES instead of the landmarks gives the landmark marginals as the diagonal blocks of the
// inverse of ONE dense matrix of size 3 Nl:
//     (H^-1)_ll = T^-1,   T = H_ll - H_lp H_pp^-1 H_pl = H_ll - Y'Y,   Y = L^-1 H_pl,   H_pp = L L'  (L block bidiagonal)
//   k_md_gather_sub : B_i = H(i, i-1) from the off-diagonal blocks of the pose-pose edges
//   k_md_factor     : block-bidiagonal Cholesky of H_pp: G_i = L_ii^-1, E_i = L(i, i-1) (a sequential chain of 6 x 6 steps, one
//                     warp, rows in registers, shuffles)
//   k_md_sweep      : Y = L^-1 H_pl, one thread per column (landmark, component), rows before the landmark's first observer
//                     stay zero
//   k_md_gemm_tn    : C = beta C + alpha A'B on 64 x 64 tiles (A, B with the contraction index leading): T -= Y'Y (lower
//                     tiles only, each tile starting at the first row where both column tiles can be non-zero), and the
//                     rank-64 updates of the inversion
//   k_md_gj_*       : in-place block Gauss-Jordan inversion of T (SPD: no pivoting), 64 columns per step
//   k_md_out        : the requested 3 x 3 diagonal blocks
// Cost O(Nl^2 Np) instead of 3 Nl latency-bound PCG solves; exact (no tolerance), deterministic.  Used when every pose-pose
// edge joins consecutive keyframes (else ssb_graph_landmark_marginals keeps solving column by column).
//
// The kernels only use blockIdx / threadIdx / __shared__ / __syncthreads (and __shfl_sync inside one single-warp CTA), and the launch sequence lives in md_run() behind a
// launcher policy, so tests/md_emulate.cpp can run THIS source on the CPU (one std::thread per CUDA thread, one barrier per
// CTA) and check it against the oracle where no GPU exists.
#pragma once

namespace ssb_md {

constexpr int TB = 64;      // tile edge of T
constexpr int KC = 16;      // contraction chunk of the tile product

#ifdef __CUDACC__
#define MD_UNROLL _Pragma("unroll")
#else
#define MD_UNROLL
#endif
__host__ __device__ inline double md_rsqrt(double d) {
#ifdef __CUDA_ARCH__
  return rsqrt(d);
#else
  return 1.0 / sqrt(d);   // the CPU emulation (tests/md_emulate.cpp)
#endif
}

struct MdGemm {
  double* C;
  int ldc;
  const double* A;
  int lda;
  const double* B;
  int ldb;
  int K;                // multiple of KC
  double alpha;
  int beta;             // 0: C = alpha A'B, 1: C += alpha A'B
  int skip_I;           // tile row left untouched (-1: none)
  int lower_only;       // 1: only tiles with I >= J
  const int* tile_k0;   // per column tile: first row that can be non-zero (multiple of KC), or null
};

// B_i = H(i, i-1): the pose-pose incidences of keyframe i whose other end is i - 1 (role 0: keyframe i is the edge's first
// vertex and Hoff = Ji' W Jj = H(i, i-1); role 1: Hoff = H(i-1, i), transposed here)
__global__ void k_md_gather_sub(const int* __restrict__ pose_pp_rowptr, const int* __restrict__ pose_pp_idx,
                                const int* __restrict__ pose_pp_other, const double* __restrict__ Hoff, int Np,
                                double* __restrict__ Bsub) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 36 * Np) return;
  const int i = t / 36, k = t - 36 * i, r = k / 6, c = k - 6 * r;
  double s = 0.0;
  for (int kk = pose_pp_rowptr[i]; kk < pose_pp_rowptr[i + 1]; ++kk) {
    if (pose_pp_other[kk] != i - 1) continue;
    const int code = pose_pp_idx[kk];
    const double* Ho = Hoff + 36 * (size_t)(code >> 1);
    s += (code & 1) == 0 ? Ho[6 * r + c] : Ho[6 * c + r];
  }
  Bsub[t] = s;
}

// Block-bidiagonal Cholesky of the block-tridiagonal H_pp:  E_i = B_i G_{i-1}',  A = H_ii - E_i E_i' = L L',  G_i = L^-1.
// The chain over the keyframes is sequential, so what counts is the latency of ONE step.  One warp, everything in registers,
// lane r < 6 owns row r of the 6 x 6 blocks; rows travel between lanes by shuffles (no shared memory, no barrier):
//   * E row r from B row r and the previous G, which every lane holds completely; all of E is then broadcast (36 shuffles);
//   * right-looking Cholesky by columns: lane cc takes ONE reciprocal square root (no sqrt, no division anywhere on the
//     chain), broadcasts it, every lane scales its entry of the column, the column is broadcast (<= 5 shuffles) and every lane
//     updates the rest of its row;
//   * after that every lane knows all of L and inverts it for itself (15 entries), which is also the G the next step needs.
// The inputs of the next step are fetched while the current one runs.  status[0] = 1 + i when A is not positive definite.
// History on a B200: one thread 7 us per keyframe; 36 threads through shared memory with 16 barrier-separated phases 3.9 us.
__global__ void __launch_bounds__(32) k_md_factor(const double* __restrict__ Hpp, const double* __restrict__ Bsub, int Np,
                                                  double* __restrict__ Ginv, double* __restrict__ Esub, int* __restrict__ status) {
  const int lane = threadIdx.x;
  const int r = lane < 6 ? lane : 5;          // lanes 6..31 shadow row 5 (they only take part in the shuffles)
  const bool own = lane < 6;
  double G[6][6];                             // previous G = L^-1 (lower triangular), complete on every lane
  MD_UNROLL
  for (int a = 0; a < 6; ++a)
    MD_UNROLL
    for (int c = 0; c < 6; ++c) G[a][c] = 0.0;
  double hn[6], bn[6];
  MD_UNROLL
  for (int c = 0; c < 6; ++c) {
    hn[c] = Hpp[6 * r + c];
    bn[c] = Bsub[6 * r + c];
  }
  for (int i = 0; i < Np; ++i) {
    double A[6], Bv[6];
    MD_UNROLL
    for (int c = 0; c < 6; ++c) {
      A[c] = hn[c];
      Bv[c] = bn[c];
    }
    if (i + 1 < Np) {
      MD_UNROLL
      for (int c = 0; c < 6; ++c) {
        hn[c] = Hpp[36 * (size_t)(i + 1) + 6 * r + c];
        bn[c] = Bsub[36 * (size_t)(i + 1) + 6 * r + c];
      }
    }
    // E row r = B row r * G'   (G lower triangular: G'(q, c) = G(c, q), q <= c); zero for the first keyframe (G = 0)
    double Er[6];
    MD_UNROLL
    for (int c = 0; c < 6; ++c) {
      double e = 0.0;
      MD_UNROLL
      for (int q = 0; q <= c; ++q) e += Bv[q] * G[c][q];
      Er[c] = e;
    }
    // A row r (lower part) -= E row r . E row c, rows c <= r fetched from their lanes
    MD_UNROLL
    for (int c = 0; c < 6; ++c) {
      double Ec[6];
      MD_UNROLL
      for (int q = 0; q < 6; ++q) Ec[q] = __shfl_sync(0xffffffffu, Er[q], c);
      double s = 0.0;
      MD_UNROLL
      for (int q = 0; q < 6; ++q) s += Er[q] * Ec[q];
      if (c <= r) A[c] -= s;
    }
    // right-looking Cholesky, column by column; L[a][c] (a > c) and il[c] = 1 / L[c][c] end up on every lane
    double L[6][6], il[6];
    MD_UNROLL
    for (int cc = 0; cc < 6; ++cc) {
      double d = A[cc];                        // on lane cc: the pivot
      if (lane == cc && !(d > 0.0)) {
        if (status[0] == 0) status[0] = 1 + i;
        d = 1.0;
      }
      const double ilc = __shfl_sync(0xffffffffu, md_rsqrt(d > 0.0 ? d : 1.0), cc);
      il[cc] = ilc;
      const double lrc = A[cc] * ilc;          // lanes r >= cc: L[r][cc]
      MD_UNROLL
      for (int a = cc + 1; a < 6; ++a) L[a][cc] = __shfl_sync(0xffffffffu, lrc, a);
      MD_UNROLL
      for (int c = cc + 1; c < 6; ++c)
        if (c <= r) A[c] -= lrc * L[c][cc];
    }
    // G = L^-1, complete on every lane: G[c][c] = il[c], G[a][c] = -il[a] sum_{q = c}^{a - 1} L[a][q] G[q][c]
    MD_UNROLL
    for (int c = 0; c < 6; ++c) {
      MD_UNROLL
      for (int a = 0; a < c; ++a) G[a][c] = 0.0;
      G[c][c] = il[c];
      MD_UNROLL
      for (int a = c + 1; a < 6; ++a) {
        double s = 0.0;
        MD_UNROLL
        for (int q = c; q < a; ++q) s += L[a][q] * G[q][c];
        G[a][c] = -s * il[a];
      }
    }
    if (own)
      MD_UNROLL
      for (int c = 0; c < 6; ++c) {
        double g = 0.0;                        // G[r][c] without indexing the register array by the lane
        MD_UNROLL
        for (int a = c; a < 6; ++a)
          if (a == r) g = G[a][c];
        Ginv[36 * (size_t)i + 6 * r + c] = g;
        Esub[36 * (size_t)i + 6 * r + c] = Er[c];
      }
  }
}

// Y = L^-1 H_pl by forward substitution along the chain: y_i = G_i (h_i - E_i y_{i-1}); h_i = column (l, c) of H_pl at
// keyframe i = sum over the landmark's edges at that keyframe of HplL_e[c][0..5] (HplL_e = H(l, p), 3 x 6).  One thread per
// column; the landmark's edges are sorted by keyframe (L-order).  Y must be zero on entry; row (6 i + r) of column col is
// Y[(6 i + r) ldY + col].  edge_pose: first int of every 80-byte PLEdge record (stride in ints passed by the caller).
__global__ void __launch_bounds__(128) k_md_sweep(const int* __restrict__ edge_pose, int edge_stride, const int* __restrict__ lm_rowptr,
                                                  const double* __restrict__ HplL, const double* __restrict__ Ginv,
                                                  const double* __restrict__ Esub, int Np, int n3, double* __restrict__ Y, int ldY) {
  // nothing that has to come from global memory sits on the chain: the landmark's NEXT edge (keyframe and its 6 entries) is
  // held in registers one edge ahead, and the G / E blocks of the next 8 keyframes are fetched into registers while the
  // current 8 are consumed from shared memory
  __shared__ double Gs[8][36];
  __shared__ double Es[8][36];
  const int col = blockIdx.x * 128 + threadIdx.x;
  const bool live = col < n3;
  const int l = live ? col / 3 : 0, c = col - 3 * l;
  int e = 0, end = 0;
  if (live) {
    e = lm_rowptr[l];
    end = lm_rowptr[l + 1];
  }
  int next_p = -1;   // keyframe of edge e, or -1 when the landmark has no edge left
  double hn[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  if (e < end) {
    next_p = edge_pose[(size_t)edge_stride * e];
    const double* Hl = HplL + 18 * (size_t)e + 6 * c;
    for (int r = 0; r < 6; ++r) hn[r] = Hl[r];
  }
  // this thread's share of a chunk of 8 x 36 G and 8 x 36 E entries: positions threadIdx.x + 128 j, j = 0..4 (576 in all)
  double pre[5];
  for (int j = 0; j < 5; ++j) {
    const int q = threadIdx.x + 128 * j;
    pre[j] = 0.0;
    if (q < 576) {
      const int which = q / 288, qq = q - 288 * which, s = qq / 36, k = qq - 36 * s;
      if (s < Np) pre[j] = (which == 0 ? Ginv : Esub)[36 * (size_t)s + k];
    }
  }
  double y[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  bool started = false;
  for (int i0 = 0; i0 < Np; i0 += 8) {
    __syncthreads();
    for (int j = 0; j < 5; ++j) {
      const int q = threadIdx.x + 128 * j;
      if (q < 576) {
        const int which = q / 288, qq = q - 288 * which, s = qq / 36, k = qq - 36 * s;
        if (which == 0)
          Gs[s][k] = pre[j];
        else
          Es[s][k] = pre[j];
        const int in = i0 + 8 + s;
        pre[j] = in < Np ? (which == 0 ? Ginv : Esub)[36 * (size_t)in + k] : 0.0;
      }
    }
    __syncthreads();
    if (!live) continue;
    for (int s = 0; s < 8 && i0 + s < Np; ++s) {
      const int i = i0 + s;
      const bool has = next_p == i;
      if (!started && !has) continue;   // rows before the first observer stay zero
      started = true;
      double t[6];
      for (int r = 0; r < 6; ++r) {
        double a = 0.0;
        for (int q = 0; q < 6; ++q) a += Es[s][6 * r + q] * y[q];
        t[r] = -a;
      }
      while (next_p == i) {
        for (int r = 0; r < 6; ++r) t[r] += hn[r];
        ++e;
        if (e < end) {
          next_p = edge_pose[(size_t)edge_stride * e];
          const double* Hl = HplL + 18 * (size_t)e + 6 * c;
          for (int r = 0; r < 6; ++r) hn[r] = Hl[r];
        } else {
          next_p = -1;
        }
      }
      for (int r = 0; r < 6; ++r) {
        double a = 0.0;
        for (int q = 0; q <= r; ++q) a += Gs[s][6 * r + q] * t[q];
        y[r] = a;
      }
      for (int r = 0; r < 6; ++r) Y[(size_t)(6 * i + r) * ldY + col] = y[r];
    }
  }
}

// per column tile of Y: the first row that can be non-zero (6 x the first observer of the tile's earliest landmark), rounded
// down to a multiple of KC; K for a tile of pure padding
__global__ void k_md_tile_k0(const int* __restrict__ edge_pose, int edge_stride, const int* __restrict__ lm_rowptr, int Nl, int nt, int K,
                             int* __restrict__ tile_k0) {
  const int J = blockIdx.x * blockDim.x + threadIdx.x;
  if (J >= nt) return;
  int first = K;
  for (int col = J * TB; col < (J + 1) * TB && col < 3 * Nl; ++col) {
    const int l = col / 3;
    if (lm_rowptr[l] < lm_rowptr[l + 1]) {
      const int f = 6 * edge_pose[(size_t)edge_stride * lm_rowptr[l]];
      first = f < first ? f : first;
    }
  }
  tile_k0[J] = (first / KC) * KC;
}

// T = blockdiag(H_ll) on the first 3 Nl rows, identity on the padding; T must be zero on entry.  Hll: 6 doubles per landmark
// (upper triangle 00 01 02 11 12 22)
__global__ void k_md_init_T(const double* __restrict__ Hll, int Nl, double* __restrict__ T, int ld) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= ld) return;
  if (r >= 3 * Nl) {
    T[(size_t)r * ld + r] = 1.0;
    return;
  }
  const int l = r / 3, a = r - 3 * l;
  const double* h = Hll + 6 * (size_t)l;
  const double S[9] = {h[0], h[1], h[2], h[1], h[3], h[4], h[2], h[4], h[5]};
  for (int b = 0; b < 3; ++b) T[(size_t)r * ld + 3 * l + b] = S[3 * a + b];
}

// C(I, J) = beta C(I, J) + alpha sum_k A[k][I 64 + i] B[k][J 64 + j]: 256 threads per 64 x 64 tile, thread (tx, ty) owns rows
// ty + 16 r and columns tx + 16 c (r, c = 0..3) so that the shared-memory reads of B are conflict-free and those of A broadcast.
__global__ void __launch_bounds__(256) k_md_gemm_tn(MdGemm a) {
  const int J = blockIdx.x, I = blockIdx.y;
  if (I == a.skip_I) return;
  if (a.lower_only && J > I) return;
  __shared__ double As[KC][TB];
  __shared__ double Bs[KC][TB];
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  double acc[4][4];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
  int k0 = 0;
  if (a.tile_k0) {
    const int ki = a.tile_k0[I], kj = a.tile_k0[J];
    k0 = ki > kj ? ki : kj;
  }
  const int lr = t >> 4, lc = (t & 15) * 4;   // this thread's 4 consecutive elements of a KC x 64 slab
  for (; k0 < a.K; k0 += KC) {
    const double* ap = a.A + (size_t)(k0 + lr) * a.lda + (size_t)I * TB + lc;
    const double* bp = a.B + (size_t)(k0 + lr) * a.ldb + (size_t)J * TB + lc;
    for (int q = 0; q < 4; ++q) {
      As[lr][lc + q] = ap[q];
      Bs[lr][lc + q] = bp[q];
    }
    __syncthreads();
    for (int kk = 0; kk < KC; ++kk) {
      double av[4], bv[4];
      for (int q = 0; q < 4; ++q) {
        av[q] = As[kk][ty + 16 * q];
        bv[q] = Bs[kk][tx + 16 * q];
      }
      for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) acc[r][c] += av[r] * bv[c];
    }
    __syncthreads();
  }
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) {
      const size_t idx = ((size_t)I * TB + ty + 16 * r) * a.ldc + (size_t)J * TB + tx + 16 * c;
      a.C[idx] = (a.beta ? a.C[idx] : 0.0) + a.alpha * acc[r][c];
    }
}

// upper triangle of T from the lower one
__global__ void k_md_mirror(double* __restrict__ T, int ld) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)ld * ld) return;
  const int r = (int)(t / ld), c = (int)(t - (size_t)r * ld);
  if (c > r) T[t] = T[(size_t)c * ld + r];
}

// Gauss-Jordan step P, part 1: Pinv = inverse of the 64 x 64 pivot tile T(P, P) (SPD: no pivoting), one CTA of 256 threads;
// thread t owns row t / 4 and 16 columns.  status[1] = 1 + P when a pivot is not positive.
__global__ void __launch_bounds__(256) k_md_gj_pivot(const double* __restrict__ T, int ld, int P, double* __restrict__ Pinv,
                                                     int* __restrict__ status) {
  __shared__ double S[TB][TB + 1];
  const int t = threadIdx.x, i = t >> 2, j0 = (t & 3) * 16;
  for (int q = 0; q < 16; ++q) S[i][j0 + q] = T[((size_t)P * TB + i) * ld + (size_t)P * TB + j0 + q];
  for (int p = 0; p < TB; ++p) {
    __syncthreads();
    const double piv = S[p][p];
    if (t == 0 && !(piv > 0.0) && status[1] == 0) status[1] = 1 + P;
    const double d = 1.0 / piv;
    const double f = S[i][p];
    double prow[16];
    for (int q = 0; q < 16; ++q) prow[q] = S[p][j0 + q] * d;
    __syncthreads();
    if (i == p) {
      for (int q = 0; q < 16; ++q) S[i][j0 + q] = (j0 + q == p) ? d : prow[q];
    } else {
      for (int q = 0; q < 16; ++q) S[i][j0 + q] = (j0 + q == p) ? -f * d : S[i][j0 + q] - f * prow[q];
    }
  }
  __syncthreads();
  for (int q = 0; q < 16; ++q) Pinv[(size_t)i * TB + j0 + q] = S[i][j0 + q];
}

// Gauss-Jordan step P, part 2 (after Row = Pinv T(P, :) has been formed by k_md_gemm_tn): save the pivot column transposed,
// ColT[k][i] = T[i][P 64 + k], clear it outside the pivot rows, and put Pinv into the pivot columns of Row — so that the
// rank-64 update T(I, J) -= ColT(:, I)' Row(:, J) over ALL column tiles also produces the new pivot column -ColT' Pinv.
__global__ void __launch_bounds__(256) k_md_gj_col(double* __restrict__ T, int ld, int P, const double* __restrict__ Pinv,
                                                   double* __restrict__ Row, double* __restrict__ ColT) {
  const int i = blockIdx.x * 256 + threadIdx.x;   // row of T
  if (i >= ld) return;
  const bool pivot_row = i / TB == P;
  for (int k = 0; k < TB; ++k) {
    const size_t idx = (size_t)i * ld + (size_t)P * TB + k;
    ColT[(size_t)k * ld + i] = T[idx];
    if (!pivot_row) T[idx] = 0.0;
  }
  if (pivot_row) {
    const int k = i - P * TB;
    for (int q = 0; q < TB; ++q) Row[(size_t)k * ld + (size_t)P * TB + q] = Pinv[(size_t)k * TB + q];
  }
}

// Gauss-Jordan step P, part 4: the pivot rows of T become Row
__global__ void k_md_gj_rowcopy(double* __restrict__ T, int ld, int P, const double* __restrict__ Row) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)TB * ld) return;
  T[(size_t)P * TB * ld + t] = Row[t];
}

// the requested diagonal blocks: out9n[k] = T[3 l .. 3 l + 2][3 l .. 3 l + 2], l = lidx[k]
__global__ void k_md_out(const double* __restrict__ T, int ld, const int* __restrict__ lidx, int n, double* __restrict__ out9n) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 9 * n) return;
  const int k = t / 9, rc = t - 9 * k, r = rc / 3, c = rc - 3 * r;
  const int l = lidx[k];
  out9n[t] = T[(size_t)(3 * l + r) * ld + 3 * l + c];
}

struct MdDims {
  int Np, Nl, n3, ld, nt, K;   // ld = n3 rounded up to 64 (leading dimension of T, Y, Row, ColT), nt = ld / 64, K = 6 Np rounded up to 16
};
inline MdDims md_dims(int Np, int Nl) {
  MdDims d;
  d.Np = Np;
  d.Nl = Nl;
  d.n3 = 3 * Nl;
  d.ld = ((d.n3 + TB - 1) / TB) * TB;
  if (d.ld == 0) d.ld = TB;
  d.nt = d.ld / TB;
  d.K = ((6 * Np + KC - 1) / KC) * KC;
  return d;
}

struct MdBuffers {   // all on the device
  // inputs (the linearised system of the last optimise, lambda = 0)
  const int *pose_pp_rowptr, *pose_pp_idx, *pose_pp_other;
  const double *Hoff, *Hpp, *Hll, *HplL;
  const int* edge_pose;   // first int of every PLEdge record (L-order)
  int edge_stride;        // ints per record (20)
  const int* lm_rowptr;
  const int* lidx;        // requested landmarks
  int n_req;
  // work
  double *Bsub, *Ginv, *Esub;   // 36 Np each
  double* Y;                    // K x ld
  double* T;                    // ld x ld
  double *Row, *ColT;           // 64 x ld each
  double* Pinv;                 // 64 x 64
  int* tile_k0;                 // nt
  int* status;                  // 2
  double* out9n;                // 9 n_req
};

// The launch sequence.  L provides  L(kernel, grid_x, grid_y, block, args...)  and  L.zero(ptr, bytes).
template <class Launcher>
inline void md_run(Launcher& L, const MdDims& d, const MdBuffers& b) {
  L.zero(b.status, 2 * sizeof(int));
  L.zero(b.Y, (size_t)d.K * d.ld * sizeof(double));
  L.zero(b.T, (size_t)d.ld * d.ld * sizeof(double));
  L(k_md_gather_sub, (36 * d.Np + 255) / 256, 1, 256, b.pose_pp_rowptr, b.pose_pp_idx, b.pose_pp_other, b.Hoff, d.Np, b.Bsub);
  L(k_md_factor, 1, 1, 32, b.Hpp, (const double*)b.Bsub, d.Np, b.Ginv, b.Esub, b.status);
  L(k_md_sweep, (d.n3 + 127) / 128, 1, 128, b.edge_pose, b.edge_stride, b.lm_rowptr, b.HplL, (const double*)b.Ginv, (const double*)b.Esub,
    d.Np, d.n3, b.Y, d.ld);
  L(k_md_tile_k0, (d.nt + 63) / 64, 1, 64, b.edge_pose, b.edge_stride, b.lm_rowptr, d.Nl, d.nt, d.K, b.tile_k0);
  L(k_md_init_T, (d.ld + 255) / 256, 1, 256, b.Hll, d.Nl, b.T, d.ld);
  {
    MdGemm g{b.T, d.ld, b.Y, d.ld, b.Y, d.ld, d.K, -1.0, 1, -1, 1, b.tile_k0};
    L(k_md_gemm_tn, d.nt, d.nt, 256, g);
  }
  L(k_md_mirror, (int)(((size_t)d.ld * d.ld + 255) / 256), 1, 256, b.T, d.ld);
  for (int P = 0; P < d.nt; ++P) {
    L(k_md_gj_pivot, 1, 1, 256, (const double*)b.T, d.ld, P, b.Pinv, b.status);
    {
      // Row = Pinv T(P, :)  (Pinv symmetric: Pinv = Pinv', so the A'B form applies with A = Pinv)
      MdGemm g{b.Row, d.ld, b.Pinv, TB, b.T + (size_t)P * TB * d.ld, d.ld, TB, 1.0, 0, -1, 0, nullptr};
      L(k_md_gemm_tn, d.nt, 1, 256, g);
    }
    L(k_md_gj_col, (d.ld + 255) / 256, 1, 256, b.T, d.ld, P, (const double*)b.Pinv, b.Row, b.ColT);
    {
      MdGemm g{b.T, d.ld, b.ColT, d.ld, b.Row, d.ld, TB, -1.0, 1, P, 0, nullptr};
      L(k_md_gemm_tn, d.nt, d.nt, 256, g);
    }
    L(k_md_gj_rowcopy, (int)(((size_t)TB * d.ld + 255) / 256), 1, 256, b.T, d.ld, P, (const double*)b.Row);
  }
  if (b.n_req > 0) L(k_md_out, (9 * b.n_req + 127) / 128, 1, 128, (const double*)b.T, d.ld, b.lidx, b.n_req, b.out9n);
}

}  // namespace ssb_md
