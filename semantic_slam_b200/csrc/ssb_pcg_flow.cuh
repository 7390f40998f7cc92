// ssb_pcg_flow.cuh — K3, the on-chip resident Schur-complement PCG (sm_100a), data-flow synchronised.
//
// One persistent CTA per SM owns a contiguous keyframe range; 6 lanes per pose, 5 poses per warp; one warp
// per landmark.  Hpp / Dinv / B rows and the landmark blocks live in registers, HplP / Hoff blocks in shared
// memory for the whole solve.  The iteration needs ONE global reduction and NO grid-wide barrier:
//
//  * Chronopoulos-Gear recurrences (single-reduction preconditioned CG):
//        w = S u,  gamma = r'u,  delta = w'u   (one fused all-reduce, + the 6 restricted values P'w per CTA)
//        beta = gamma/gamma_old,  alpha = gamma / (delta - beta*gamma/alpha_old)
//        p = u + beta p,  s = w + beta s,  x += alpha p,  r -= alpha s,  u = M^-1 r
//    identical iterates to standard PCG in exact arithmetic.
//  * Vectors that cross CTAs (u, and the landmark-space product v) travel as self-validating 16-byte
//    cells {lo32, tag, hi32, tag} (each 8-byte half is written atomically, NCCL "LL" style): a consumer
//    spins on exactly the cells it needs, so the u -> v -> w hand-offs cost one L2 hop each instead of a
//    fence + counter barrier + re-read, and a CTA only ever waits for its own producers.
//  * The all-reduce is a pull all-gather of one 128-byte line (8 cells) per CTA, folded by every CTA in
//    the same fixed order => bit-reproducible for a fixed grid.  It is also the only write-after-read
//    fence the cells need (a producer overwrites u / v only after it has seen every CTA's line of the
//    iteration in which the old values were read).
//  * The coarse level needs no restricted-residual vector: z_c = A_c^-1 P'r is carried by the recurrences
//    t = A_c^-1 P'w (6 rows per CTA, dot with the gathered P'w),  y = t + beta y,  z_c -= alpha y.
//
// Tags are tagbase + iteration; tagbase is unique per launch (host counter << 16), so no buffer has to be
// cleared between solves.
#pragma once
#include "ssb_graph_kernels.cuh"
#include "ssb_peer.cuh"

namespace ssb {

struct FlowBufs {
  uint4* ucell;   // [6 * Np]  u = M^-1 r
  uint4* vcell;   // [3 * Nl]  v = W Hlp u
  uint4* lines;   // [2][gridDim.x][8] per-CTA reduction lines, double buffered on the tag parity
  uint4* gj;      // [gridDim.x][36 gridDim.x + 8] Gauss-Jordan pivot panels of the coarse inversion, one per step
  unsigned tagbase;
  double2* hlpark;  // [gridDim.x * warps][9][32] landmark-role blocks parked in L2 between iterations
  unsigned long long* trace;  // debug (-DSSB_FLOW_TRACE): [gridDim.x][8] globaltimer stamps of one iteration
  // REP (template parameter of k_pcg_flow): the graph is rep_count copies of one graph side by side, copy j on the CTAs
  // [j rep_ctas, (j + 1) rep_ctas), each with its own right-hand side (the landmark marginals, ssb_graph.cu)
  int rep_ctas, rep_count;
};

// Sharded graphs (template parameter MR of k_pcg_flow): what a CTA writes into the arenas of the other ranks.
// All tables hold absolute pointers into the peers' mapped arenas, built by the host per structure change.
struct FlowPeer {
  int world, rank;
  const int* upush_rowptr;     // [Np_own + 1] ranks that hold my keyframe as a ghost
  uint4* const* upush_cell;    //   its 6 u cells in that rank's cell buffer
  double* const* upush_x;      //   its 6 slots in that rank's solution vector
  const int* vpush_rowptr;     // [own landmark parts + 1] ranks whose keyframes see the landmark of the part
  uint4* const* vpush_cell;    //   the 3 v cells of the part in that rank's cell buffer
  uint4* lines[SSB_MAX_WORLD]; // every rank's reduction lines [2][world][NB][8]; line (p, r, b) is written by CTA b of rank r
  // rank-level coarse level (ssb_graph_kernels.cuh: GlobDev).  glob != 0: every CTA writes (gamma, delta, E_a' P_c'w) into
  // glines of EVERY rank (its own included) and all cross-rank sums are folded from there.
  int glob;
  uint4* glines[SSB_MAX_WORLD]; // [2][world][NB][8]
  const float* Aginv;          // [6][6 world] my rank's rows of w_g A_g^-1
  const double* gcent;         // [world][4] rank centroids
};

__device__ __forceinline__ void st_cell(uint4* c, double v, unsigned tag) {
  const unsigned lo = (unsigned)__double2loint(v), hi = (unsigned)__double2hiint(v);
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(c), "r"(lo), "r"(tag), "r"(hi), "r"(tag)
               : "memory");
}
__device__ __forceinline__ uint4 ld_cell(const uint4* c) {
  uint4 u;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(c));
  return u;
}
__device__ __forceinline__ bool cell_ok(const uint4& u, unsigned tag) { return u.y == tag && u.w == tag; }
__device__ __forceinline__ double cell_val(const uint4& u) { return __hiloint2double((int)u.z, (int)u.x); }

// 6 consecutive doubles (16-byte aligned) from a shared-space byte address: 3 x LDS.128
__device__ __forceinline__ void lds_row6(uint32_t addr, double* out) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(out[0]), "=d"(out[1]) : "r"(addr));
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+16];" : "=d"(out[2]), "=d"(out[3]) : "r"(addr));
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+32];" : "=d"(out[4]), "=d"(out[5]) : "r"(addr));
}

#ifndef SSB_POLL_NS
#define SSB_POLL_NS 0    // back-off between polls of a cell that has not arrived.  Measured on cfg2 with the coalesced
                         // cell layout (A/B on one box, solve time per 20 LM iterations): 0 ns 33.1 ms, 8 and 20 ns
                         // 34.2 ms, 40 ns 39.9 ms, 100 ns +20 % -> no back-off
#endif
// Debug watchdog (build with EXTRA=-DSSB_WATCHDOG): a cell that has not arrived after this many polls (>= 10 s)
// means a producer died; trap instead of hanging (the launch then fails with a sticky error).  Off by default: the
// poll counter alone, in whichever wait it sits, was measured to cost 1.7 - 2.5 % of the solve (A/B on one box).
#ifndef SSB_SPIN_LIMIT
#define SSB_SPIN_LIMIT (1u << 24)
#endif
// spin on one cell (the first attempt was already issued by the caller)
// SYS: the cell may have been written by another rank over NVLink (system-scope load, and a globaltimer deadline
// checked every 4096 polls: a rank that died must not hang this one — trap instead)
// In-kernel watchdog of the single-GPU kernel (-DSSB_RELEASE_WATCHDOG=1): the same deadline as the cross-rank waits — a
// poll counter that only runs while a wait is actually spinning, a globaltimer read every 4096 polls against a deadline
// kept in shared memory; a cell that has not arrived after SSB_PEER_TIMEOUT_NS traps (sticky launch error).  OFF by
// default: measured on cfg2 (A/B on one box, same spills as without) 551 -> 529 LM it/s, -4 % — the spin loops ARE the
// critical path of this kernel and one extra add + test per poll shows.  The sharded kernels (MR) always carry it;
// single-GPU release builds rely on the host-side deadline of read_scalars (ssb_graph.cu) instead.
#ifndef SSB_RELEASE_WATCHDOG
#define SSB_RELEASE_WATCHDOG 0
#endif
constexpr bool kFlowWatch = SSB_RELEASE_WATCHDOG != 0;
template <bool SYS>
__device__ __forceinline__ uint4 ld_cell_t(const uint4* p) {
  if constexpr (SYS)
    return ld_cell_sys(p);
  else
    return ld_cell(p);
}
// the deadline of this launch lives in shared memory (set once by the kernel): the waits keep only a poll counter
__shared__ unsigned long long s_flow_deadline;
__device__ __forceinline__ void peer_deadline() {
  if (peer_globaltimer() > s_flow_deadline) __trap();
}
template <bool SYS = false>
__device__ __forceinline__ double cell_wait(const uint4* p, uint4 c, unsigned tag) {
#ifdef SSB_WATCHDOG
  unsigned spins = 0;
#endif
  unsigned polls = 0;
  while (!cell_ok(c, tag)) {
#if SSB_POLL_NS > 0
    __nanosleep(SSB_POLL_NS);
#endif
    c = ld_cell_t<SYS>(p);
    if constexpr (SYS || kFlowWatch) {
      if ((++polls & 4095u) == 0) peer_deadline();
    }
#ifdef SSB_WATCHDOG
    if (++spins > SSB_SPIN_LIMIT) __trap();
#endif
  }
  return cell_val(c);
}

// Batch wait: N cells base[off[m]] (off[m] < 0: not wanted) were all loaded once into c[].  Cells are consumed in
// order; when cell m had to be spun on, the copies of the later cells are stale too, so they are re-issued TOGETHER
// right after m arrived (one extra round trip for the rest of the batch instead of one per cell), and only then.
template <int N, bool SYS = false>
__device__ __forceinline__ void cells_wait(const uint4* base, const int (&off)[N], uint4 (&c)[N], unsigned tag,
                                           double (&val)[N]) {
#pragma unroll
  for (int m = 0; m < N; ++m) {
    val[m] = 0.0;
    if (off[m] >= 0) {
      if (!cell_ok(c[m], tag)) {
#ifdef SSB_WATCHDOG
        unsigned spins = 0;
#endif
        unsigned polls = 0;
        do {
#if SSB_POLL_NS > 0
          __nanosleep(SSB_POLL_NS);
#endif
          c[m] = ld_cell_t<SYS>(base + off[m]);
          if constexpr (SYS || kFlowWatch) {
            if ((++polls & 4095u) == 0) peer_deadline();
          }
#ifdef SSB_WATCHDOG
          if (++spins > SSB_SPIN_LIMIT) __trap();
#endif
        } while (!cell_ok(c[m], tag));
#pragma unroll
        for (int q = m + 1; q < N; ++q)
          if (off[q] >= 0 && !cell_ok(c[q], tag)) c[q] = ld_cell_t<SYS>(base + off[q]);
      }
      val[m] = cell_val(c[m]);
    }
  }
}

// Packed butterfly reductions: N values per lane are reduced across the warp with N-1 + (5 - log2 N)
// exchanges instead of 5 N.  The sum of value j ends up in the lanes whose upper bits spell j.
// Fixed exchange pattern => deterministic.
// 8 values: result for value j in lanes 4j .. 4j+3
__device__ __forceinline__ double warp_reduce8(const double* v) {
  const int lane = threadIdx.x & 31;
  double a[4], b[2], c;
  {
    const bool up = (lane & 16) != 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double send = up ? v[k] : v[k + 4];
      const double keep = up ? v[k + 4] : v[k];
      a[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool up = (lane & 8) != 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const double send = up ? a[k] : a[k + 2];
      const double keep = up ? a[k + 2] : a[k];
      b[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool up = (lane & 4) != 0;
    const double send = up ? b[0] : b[1];
    const double keep = up ? b[1] : b[0];
    c = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  c += __shfl_xor_sync(0xffffffffu, c, 2);
  c += __shfl_xor_sync(0xffffffffu, c, 1);
  return c;
}
// 4 values: result for value j in lanes 8j .. 8j+7
__device__ __forceinline__ double warp_reduce4(const double* v) {
  const int lane = threadIdx.x & 31;
  double a[2], b;
  {
    const bool up = (lane & 16) != 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const double send = up ? v[k] : v[k + 2];
      const double keep = up ? v[k + 2] : v[k];
      a[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool up = (lane & 8) != 0;
    const double send = up ? a[0] : a[1];
    const double keep = up ? a[1] : a[0];
    b = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  b += __shfl_xor_sync(0xffffffffu, b, 4);
  b += __shfl_xor_sync(0xffffffffu, b, 2);
  b += __shfl_xor_sync(0xffffffffu, b, 1);
  return b;
}

// per-CTA gather tables built once per structure change on the host (ssb_graph.cu: prepare)
struct FlowTabs {
  const int* ulm_rowptr;  // [nblk+1] distinct landmarks referenced by the CTA's poses
  const int* ulm;         //          their ids
  const int* pl_loc;      // [El]     P-order entry -> index into the CTA's distinct-landmark list
  const int* upp_rowptr;  // [nblk+1] distinct pose-pose edges incident to the CTA's poses
  const int* upp;         //          their ids
  const int* pp_loc;      // [2 Epp]  incidence -> index into the CTA's distinct-edge list
  const int* pp_src;      // [2 Epp]  incidence -> slot of the neighbour's u in shared memory | role << 31
  const int* ext_rowptr;  // [nblk+1] neighbour poses owned by other CTAs
  const int* ext;
  const int* gj_order;    // [nblk] pivot order of the coarse Gauss-Jordan (nested dissection, separators last)
  const unsigned* gj_mask; // [nblk][ceil(nblk/32)] structurally non-zero column blocks of the pivot panel per step
  // landmark parts: runs of <= 64 L-order edges of one landmark, one warp each (part q -> CTA q % nblk, warp q / nblk)
  const int* part_lm;     // [n_parts] landmark of the part
  const int* part_e0;     // [n_parts] first L-order edge
  const int* part_e1;     // [n_parts] one past the last edge
  const int* lm_partbase; // [Nl+1]    first part of every landmark (v cells: 3 per part, parts of a landmark adjacent)
  int n_parts;
};

constexpr int PCGW_MAXU = 320;    // distinct landmarks per CTA
constexpr int PCGW_MAXPPE = 96;   // distinct pose-pose edges per CTA
constexpr int PCGW_MAXEXT = 16;   // external neighbour poses per CTA
constexpr int PCGW_POSES = 5 * (PCGF_THREADS / 32);
constexpr int PCGW_MAXSTAGE = 3 * PCGW_MAXU + 6 * PCGW_MAXEXT;   // cells staged into shared memory per iteration
constexpr int PCGW_INTS = PCGF_MAXPL + PCGW_MAXSTAGE + 2 * PCGF_MAXPP + 6 * PCGF_MAXOV;
constexpr int PCGW_BIG = 18 * PCGF_MAXPL + 36 * PCGW_MAXPPE + 18 * PCGF_MAXOV + 2 * PCGW_POSES * 36 + 3 * PCGW_MAXU +
                         6 * (PCGW_POSES + PCGW_MAXEXT) + (PCGW_INTS + 1) / 2;
// dynamic shared memory of k_pcg_flow in doubles for a grid of nblk CTAs
__host__ __device__ constexpr size_t pcg_flow_smem_doubles(int nblk) {
  return (size_t)PCGF_THREADS + (size_t)7 * 6 * nblk + (PCGF_THREADS / 36) * 36 +
         (size_t)(PCGW_BIG > 6 * 6 * nblk ? PCGW_BIG : 6 * 6 * nblk);
}


// 1/x to full double precision without the IEEE division sequence: hardware seed (~20 bits) + 2 Newton steps
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(r, fma(-x, r, 1.0), r);
  r = fma(r, fma(-x, r, 1.0), r);
  return r;
}
// 6x6 SPD inverse by 6 lanes of one warp (lane j owns column j of [A | I]); all 32 lanes must call.
__device__ __forceinline__ bool warp_inv6_fast(double* col) {
  const int lane = threadIdx.x & 31;
  double inv[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) inv[i] = (i == lane) ? 1.0 : 0.0;
  bool ok = true;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const double piv = __shfl_sync(0xffffffffu, col[k], k);
    if (!(piv > 0.0) || !isfinite(piv)) ok = false;
    const double ip = fast_rcp(piv);
    double ck[6];
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

// Block Gauss-Jordan inversion of A_c (my 6 rows in Arow[6][nc]) without grid barriers: the owner of pivot
// aggregate k publishes its scaled panel as tagged cells into a per-step buffer, every other CTA applies the
// panels in order as they arrive, so the critical path is panel k -> CTA k+1 -> panel k+1 (one L2 hop per
// step).  CTAs whose block in pivot column k is zero skip the step without waiting.
// Returns true when every pivot block was positive definite.
template <int NT, int NB>
__device__ bool coarse_gj_flow(uint4* gj, unsigned tag, double* Arow, const int* order, const unsigned* mask) {
  constexpr int nc = 6 * NB;
  constexpr int PER = (nc + NT - 1) / NT;
  constexpr int PSTRIDE = 6 * nc + 8;
  __shared__ int s_flag;
  __shared__ double piv_sh[40];   // pivot inverse (pivot CTA) or F = my block in pivot column k
  __shared__ int order_sh[NB];
  constexpr int MW = (NB + 31) / 32;
  if (threadIdx.x == 0) s_flag = 0;
  for (int q = threadIdx.x; q < NB; q += NT) order_sh[q] = order[q];
  __syncthreads();
  for (int step = 0; step < NB; ++step) {
    const int k = order_sh[step];   // pivot aggregate of this step (nested-dissection order)
    const unsigned* mk = mask + (size_t)step * MW;   // structurally non-zero column blocks of this panel
    uint4* P = gj + (size_t)k * PSTRIDE;
    if ((int)blockIdx.x == k) {
      if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        double col[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) col[a] = lane < 6 ? Arow[a * nc + 6 * k + lane] : 0.0;
        const bool ok = warp_inv6_fast(col);
        if (lane < 6) {
#pragma unroll
          for (int a = 0; a < 6; ++a) piv_sh[6 * a + lane] = ok ? col[a] : (a == lane ? 1.0 : 0.0);
        }
        if (lane == 0) piv_sh[36] = ok ? 0.0 : 1.0;
      }
      __syncthreads();
      for (int j = threadIdx.x; j < nc; j += NT) {
        const int cb = j / 6;
        if (!((__ldg(mk + (cb >> 5)) >> (cb & 31)) & 1u)) continue;   // structurally zero block: stays zero, not published
        double colv[6];
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
          Arow[a * nc + j] = t;
          st_cell(P + a * nc + j, t, tag);
        }
      }
      if (threadIdx.x == 0) {
        st_cell(P + 6 * nc, piv_sh[36], tag);
        if (piv_sh[36] != 0.0) s_flag = 1;
      }
      __syncthreads();
    } else {
      // F = my block in pivot column k.  F == 0 (aggregates not coupled so far): the step leaves my rows
      // unchanged, so the panel is neither awaited nor read.
      double fv = 0.0;
      if (threadIdx.x < 36) {
        fv = Arow[(threadIdx.x / 6) * nc + 6 * k + (threadIdx.x % 6)];
        piv_sh[threadIdx.x] = fv;
      }
      if (__syncthreads_or(fv != 0.0) == 0) continue;
      double pj[PER][6];
#pragma unroll
      bool nz[PER];
#pragma unroll
      for (int m = 0; m < PER; ++m) {
        const int j = threadIdx.x + NT * m;
        const int cb = j / 6;
        nz[m] = j < nc && ((__ldg(mk + (cb >> 5)) >> (cb & 31)) & 1u);
        if (nz[m]) {
          uint4 c[6];
          int po[6];
#pragma unroll
          for (int b = 0; b < 6; ++b) {
            po[b] = b * nc + j;
            c[b] = ld_cell(P + po[b]);
          }
          cells_wait<6>(P, po, c, tag, pj[m]);
        }
      }
#pragma unroll
      for (int m = 0; m < PER; ++m) {
        const int j = threadIdx.x + NT * m;
        if (nz[m]) {
          const bool inpiv = (j >= 6 * k && j < 6 * k + 6);
#pragma unroll
          for (int a = 0; a < 6; ++a) {
            double t = 0.0;
#pragma unroll
            for (int b = 0; b < 6; ++b) t += piv_sh[6 * a + b] * pj[m][b];
            Arow[a * nc + j] = inpiv ? -t : Arow[a * nc + j] - t;
          }
        }
      }
      __syncthreads();
    }
  }
  // pivot failures must be seen by every CTA, including those that skipped the failing step
  if (threadIdx.x < NB) {
    const uint4* fc = gj + (size_t)threadIdx.x * PSTRIDE + 6 * nc;
    if (cell_wait(fc, ld_cell(fc), tag) != 0.0) s_flag = 1;
  }
  __syncthreads();
  return s_flag == 0;
}

// The coarse assembly + Gauss-Jordan as a kernel of its own (same device code as the prologue of k_pcg_flow): it
// needs only the linearised system and (H_ll + lambda I)^-1, so the host runs it on a third stream next to
// k_sub_assemble -> k_grp_invert and k_prep_poses; k_pcg_flow then starts from the stored rows
// (Cz.reuse_inverse = 2).  Launched with NB CTAs that wait on each other's panels: like k_pcg_flow it relies on all
// of them becoming resident, which holds because the kernels running beside it never wait on anything.
#ifndef SSB_CINV_THREADS
#define SSB_CINV_THREADS 512
#endif
#ifndef SSB_CINV_MINB
#define SSB_CINV_MINB 2   // 512 threads x <= 64 registers: must leave room for the kernels it runs beside
#endif
constexpr int CINV_THREADS = SSB_CINV_THREADS;
template <int NB>
__global__ void __launch_bounds__(CINV_THREADS, SSB_CINV_MINB)
    k_coarse_invert(DevGraph G, CoarseDev Cz, uint4* gj, unsigned tag, const int* order, const unsigned* mask, double lambda) {
  extern __shared__ __align__(16) double dsm[];
  constexpr int nc = 6 * NB;
  double* Arow = dsm;             // [6][nc]
  double* red = Arow + 6 * nc;    // [CINV_THREADS / 36][36]
  if (threadIdx.x == 0) s_flow_deadline = peer_globaltimer() + SSB_PEER_TIMEOUT_NS;
  const int p0 = blockIdx.x * Cz.C, p1 = min(G.Np_own, p0 + Cz.C);
  coarse_assemble<CINV_THREADS>(G, Cz, lambda, Arow, red, p0, p1);
  const bool ok = coarse_gj_flow<CINV_THREADS, NB>(gj, tag, Arow, order, mask);
  if (ok)
    for (int k = threadIdx.x; k < 6 * nc; k += CINV_THREADS) Cz.ainv_store[(size_t)blockIdx.x * 6 * nc + k] = Arow[k];
  if (blockIdx.x == 0 && threadIdx.x == 0) *Cz.ainv_ok = ok ? 1 : 0;
}

// NB = gridDim.x as a compile-time constant (148 = one CTA per B200 SM): every shared-memory array then has a
// constant address and the reduction / coarse loops unroll.
// MR: the graph is sharded over FP.world ranks (this grid = the CTAs of rank FP.rank); the cells a neighbour needs are
// also written into its arena, the dot products are folded over the lines of every rank.  The preconditioner stays
// rank-local (ghost keyframes carry no basis): block-Jacobi on S + exact groups + one coarse level per rank.
// REP: block-diagonal system diag(S, ..., S) with one right-hand side per block.  The blocks share nothing but the kernel:
// every block runs ITS OWN conjugate-gradient recurrence (gamma, delta, alpha, beta folded over the lines of its own CTAs
// only), so each column converges as if it were solved alone.  A converged block freezes (alpha = beta = 0: it keeps
// publishing the same cells) until every block is done — the tags advance in lock-step for the whole grid.
constexpr int PCGF_MAXREP = PCGF_THREADS / 32;   // one warp folds the lines of one block
template <int NB, bool MR = false, bool REP = false>
__global__ void __launch_bounds__(PCGF_THREADS, 1)
    k_pcg_flow(DevGraph G, CoarseDev Cz, BarSlot* slots, FlowBufs F, FlowTabs T, double lambda, double tol2, int maxit, FlowPeer FP) {
  extern __shared__ __align__(16) double dsm[];
  __shared__ double s6[8], zc6[8], red6[7 * 32];
  __shared__ double rs_sh[8 * SSB_MAX_WORLD];   // MR: per rank (gamma, delta, P_g'w [6]) folded from the lines of that rank
  __shared__ float ag_sh[MR ? 36 * SSB_MAX_WORLD : 1];   // MR: my rank's rows of w_g A_g^-1
  __shared__ double tg6_sh[8], dga_sh[4];        // MR: A_g^-1 P_g'w (my rank's 6 values); c_a - c_rank of my CTA aggregate
  __shared__ double gl8_sh[8];                   // MR: this CTA's (gamma, delta, P_c'w) of the iteration
  __shared__ int ovcnt[PCGF_THREADS / 32];
  __shared__ double rg_sh[REP ? PCGF_MAXREP : 1], rd_sh[REP ? PCGF_MAXREP : 1], g0_sh[REP ? PCGF_MAXREP : 1];   // REP: per block gamma, delta, gamma_0
  __shared__ int frz_sh[REP ? PCGF_MAXREP : 1];   // REP: 0 = iterating, 1 = converged (frozen), 2 = breakdown, 3 = bad right-hand side
  static_assert(!(MR && REP), "replicated right-hand sides are a single-GPU path");
  constexpr int nblk = NB;
  constexpr int nc = 6 * nblk;
  double* part_sh = dsm;                  // [PCGF_THREADS]  (loop: gam = [0..255], del = [256..511])
  double* Arow = part_sh + PCGF_THREADS;  // [6][nc]  my rows of A_c^-1
  double* wc = Arow + 6 * nc;             // [nc]     gathered P'w (init: P'r)
  double* red = wc + nc;                  // [14][36] prologue scratch (loop: per-warp partials)
  double* big = red + (PCGF_THREADS / 36) * 36;
  double* panel_sh = big;                              // Gauss-Jordan panel [6][nc], afterwards the resident operands:
  double* plH = big;                                   // [PCGF_MAXPL][18]  HplP blocks of my poses (6x3)
  double* ppH = plH + 18 * PCGF_MAXPL;                 // [PCGW_MAXPPE][36] Hoff blocks of the distinct edges
  double* ovH = ppH + 36 * PCGW_MAXPPE;                // [6 PCGF_MAXOV][3] HplL columns of edges 32.. of a landmark, item order
  double* b1_sh = ovH + 18 * PCGF_MAXOV;               // [80][36] P1 rows
  double* m1_sh = b1_sh + PCGW_POSES * 36;             // [80][36] P1 D1^-1 rows
  // u of my poses [80][6], then the staged cells of an iteration in one run: u of the external neighbours
  // [next][6] followed by v of the landmarks my poses see [nuniq][3]
  double* u_sh = m1_sh + PCGW_POSES * 36;
  int* pl_loc = reinterpret_cast<int*>(u_sh + 6 * PCGW_POSES + PCGW_MAXSTAGE);  // [PCGF_MAXPL]
  int* stage_src = pl_loc + PCGF_MAXPL;                // [PCGW_MAXSTAGE] cell index (u and v cells share one buffer)
  int* pp_loc = stage_src + PCGW_MAXSTAGE;             // [PCGF_MAXPP]
  int* pp_src = pp_loc + PCGF_MAXPP;                   // [PCGF_MAXPP]
  int* ov_cell = pp_src + PCGF_MAXPP;                  // [6 PCGF_MAXOV] u cell of (edge, column) items 192.. of a landmark
  double* gam = part_sh;
  double* del = part_sh + PCGF_THREADS / 2;
  double* scratch8 = red;                  // [16][8]
  double* dpart = red + 8 * (PCGF_THREADS / 32);  // [12]
  const bool use_sub = Cz.sub_enabled != 0;
  const bool coarse = Cz.enabled != 0;

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_flow_deadline = peer_globaltimer() + SSB_PEER_TIMEOUT_NS;   // read by the waits (after a barrier)
  const int slot = lane / 6, comp = lane - 6 * slot;
  const int base_lane = 6 * slot;
  const int p0 = blockIdx.x * Cz.C, p1 = min(G.Np_own, p0 + Cz.C);
  const int i = p0 + warp * 5 + slot;          // my pose (pose role)
  const bool act = lane < 30 && i < p1;
  const int part = warp * nblk + blockIdx.x;   // my landmark part (landmark role), round-robin over CTAs
  const bool lact = part < T.n_parts;
  const int l = lact ? T.part_lm[part] : 0;    // its landmark
  unsigned epoch = 0;
  int status = 0;

#ifdef SSB_PCG_TIMERS
  const long long t_start = clock64();
  long long t_asm = t_start;
#endif
  bool use_coarse = false;
  if (coarse) {
    if (Cz.reuse_inverse) {
      // 1: rows kept from an earlier solve (refresh policy); 2: rows written by k_coarse_invert for this solve
      use_coarse = Cz.reuse_inverse == 1 || *Cz.ainv_ok != 0;   // uniform over the grid
      if (use_coarse)
        for (int k = threadIdx.x; k < 6 * nc; k += PCGF_THREADS) Arow[k] = Cz.ainv_store[(size_t)blockIdx.x * 6 * nc + k];
      __syncthreads();
    } else {
      coarse_assemble<PCGF_THREADS>(G, Cz, lambda, Arow, red, p0, p1);
#ifdef SSB_PCG_TIMERS
      t_asm = clock64();
#endif
      use_coarse = coarse_gj_flow<PCGF_THREADS, NB>(F.gj, F.tagbase + 1u, Arow, T.gj_order, T.gj_mask);
      if (use_coarse)
        for (int k = threadIdx.x; k < 6 * nc; k += PCGF_THREADS) Cz.ainv_store[(size_t)blockIdx.x * 6 * nc + k] = Arow[k];
    }
  }

  // Weights of the additive levels.  With the group level (preconditioner 3) block-Jacobi is damped and the
  // CTA-level coarse correction amplified: (0.5, 1, 2) needs 6-10 % fewer iterations than (1, 1, 1) on cfg2
  // (scripts/precond_study.py); any positive weights keep M symmetric positive definite.
  const double w_jac = (Cz.sub_enabled != 0 && Cz.grp_enabled != 0) ? 0.5 : 1.0;
  const double w_coarse = (Cz.sub_enabled != 0 && Cz.grp_enabled != 0) ? 2.0 : 1.0;
  if (use_coarse && w_coarse != 1.0) {
    for (int k = threadIdx.x; k < 6 * nc; k += PCGF_THREADS) Arow[k] *= w_coarse;
    __syncthreads();
  }
  // ---- load the resident operands ------------------------------------------------------------
  double Hrow[6], Drow[6], Brow[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    Hrow[k] = act ? G.Hpp[36 * (size_t)i + 6 * comp + k] : 0.0;
    Drow[k] = act ? w_jac * G.Dinv[36 * (size_t)i + 6 * comp + k] : 0.0;
    Brow[k] = (act && use_coarse) ? Cz.Bmat[36 * (size_t)i + 6 * comp + k] : 0.0;
  }
  Hrow[comp] += act ? lambda : 0.0;
  const int plbase = G.pose_pl_rowptr[p0 < G.Np ? p0 : G.Np];
  const int ppbase = G.pose_pp_rowptr[p0 < G.Np ? p0 : G.Np];
  const int npl_blk = G.pose_pl_rowptr[p1 > p0 ? p1 : (p0 < G.Np ? p0 : G.Np)] - plbase;
  const int npp_blk = G.pose_pp_rowptr[p1 > p0 ? p1 : (p0 < G.Np ? p0 : G.Np)] - ppbase;
  const int ulm0 = T.ulm_rowptr[blockIdx.x], nuniq = T.ulm_rowptr[blockIdx.x + 1] - ulm0;
  const int upp0 = T.upp_rowptr[blockIdx.x], nupp = T.upp_rowptr[blockIdx.x + 1] - upp0;
  const int ext0 = T.ext_rowptr[blockIdx.x], next = T.ext_rowptr[blockIdx.x + 1] - ext0;
  for (int k = threadIdx.x; k < 18 * npl_blk; k += PCGF_THREADS) plH[k] = G.HplP[18 * (size_t)plbase + k];
  for (int k = threadIdx.x; k < npl_blk; k += PCGF_THREADS) pl_loc[k] = T.pl_loc[plbase + k];
  const int nstage = 6 * next + 3 * nuniq;
  const int voff = (int)(F.vcell - F.ucell);  // both live in one allocation
  for (int q = threadIdx.x; q < nstage; q += PCGF_THREADS) {
    if (q < 6 * next) {
      const int xe = q / 6;
      stage_src[q] = 6 * T.ext[ext0 + xe] + (q - 6 * xe);
    } else {
      // v of a landmark = sum over its parts: first part's cell in the low 24 bits, part count in the top 8
      const int qq = q - 6 * next, lu = qq / 3, lmid = T.ulm[ulm0 + lu];
      const int pb = T.lm_partbase[lmid], np = T.lm_partbase[lmid + 1] - pb;
      stage_src[q] = (voff + 3 * pb + (qq - 3 * lu)) | (np << 24);
    }
  }
  double* const v_sh = u_sh + 6 * (PCGW_POSES + next);
  for (int k = threadIdx.x; k < npp_blk; k += PCGF_THREADS) {
    pp_loc[k] = T.pp_loc[ppbase + k];
    pp_src[k] = T.pp_src[ppbase + k];
  }
  for (int k = threadIdx.x; k < 36 * nupp; k += PCGF_THREADS) ppH[k] = G.Hoff[36 * (size_t)T.upp[upp0 + k / 36] + (k % 36)];
  for (int k = threadIdx.x; k < 2 * 36 * PCGW_POSES; k += PCGF_THREADS) b1_sh[k] = 0.0;  // b1_sh and m1_sh (contiguous)
  __syncthreads();
  // preconditioner 3: the m1 region holds instead the packed inverses of my two aggregate groups [2][GRP_PACK],
  // the restricted residual of my aggregates r5 [96] and the group solutions z5 [96]
  const bool use_grp = use_sub && Cz.grp_enabled != 0;
  double* const GI = m1_sh;
  double* const r5_sh = m1_sh + 2 * GRP_PACK;
  double* const z5_sh = r5_sh + 6 * (PCGF_THREADS / 32);
  int gn0 = 0, gn1 = 0;   // dimensions of my two groups
  if (use_grp) {
    const int ga0 = Cz.grp_first_agg[2 * blockIdx.x], gah = Cz.grp_first_agg[2 * blockIdx.x + 1],
              ga1 = Cz.grp_first_agg[2 * blockIdx.x + 2];
    gn0 = 6 * (gah - ga0);
    gn1 = 6 * (ga1 - gah);
  }
  if (use_sub) {
    // P1 rows and M1 = P1 D1^-1 rows of my poses (zero D1^-1 = level off for that aggregate)
    for (int k = threadIdx.x; k < 36 * (p1 - p0); k += PCGF_THREADS) {
      const int ip = p0 + k / 36, rc_ = k % 36, row = rc_ / 6, j = rc_ - 6 * row;
      const double* B1 = Cz.B1mat + 36 * (size_t)ip + 6 * row;
      const double* D1 = Cz.D1inv + 36 * (size_t)(ip / 5);
      b1_sh[k] = B1[j];
      if (!use_grp) {
        double t = 0.0;
#pragma unroll
        for (int m = 0; m < 6; ++m) t += B1[m] * D1[6 * m + j];
        m1_sh[k] = t;
      }
    }
    if (use_grp)
      for (int k = threadIdx.x; k < 2 * GRP_PACK; k += PCGF_THREADS) GI[k] = Cz.GrpInv[(size_t)2 * blockIdx.x * GRP_PACK + k];
  }
  // shared-space byte address of my P1 row (lanes 30, 31 alias slot 0: their r is 0); the M1 row sits
  // 36 * PCGW_POSES doubles further
  const uint32_t b1addr = smem_u32(b1_sh + 36 * (warp * 5 + (lane < 30 ? slot : 0)) + 6 * (lane < 30 ? comp : 0));
  const int mypl0 = act ? G.pose_pl_rowptr[i] - plbase : 0, mypl1 = act ? G.pose_pl_rowptr[i + 1] - plbase : 0;
  const int mypp0 = act ? G.pose_pp_rowptr[i] - ppbase : 0, mypp1 = act ? G.pose_pp_rowptr[i + 1] - ppbase : 0;
  // landmark role.  The (edge, column) pairs of landmark l are dealt to the lanes in cell order:
  // item idx = 32 m + lane  ->  edge idx / 6, column idx % 6, so that one warp-wide load instruction reads
  // runs of 6 consecutive cells (one pose's u) instead of 32 scattered ones.
  int uoff[6];
  double Wr[3] = {0, 0, 0};  // lane k < 3: row k of W_l = (Hll + lambda I)^-1
  int deg = 0, le0 = 0;
  if (lact) {
    le0 = T.part_e0[part];
    deg = T.part_e1[part] - le0;
    if (lane < 3) {
      const double* Wu = G.HllInv + 6 * (size_t)l;
      // upper triangle storage: 0:(0,0) 1:(0,1) 2:(0,2) 3:(1,1) 4:(1,2) 5:(2,2)
      Wr[0] = lane == 0 ? Wu[0] : (lane == 1 ? Wu[1] : Wu[2]);
      Wr[1] = lane == 0 ? Wu[1] : (lane == 1 ? Wu[3] : Wu[4]);
      Wr[2] = lane == 0 ? Wu[2] : (lane == 1 ? Wu[4] : Wu[5]);
    }
  }
  // The 18 block entries of a lane are only needed during phase A: they are parked in global memory (L2
  // resident, [chunk][lane] so that the reload is coalesced) and re-read together with the u cells each
  // iteration, which leaves the register file to the pose role during phases B..D.
  double2* const park = F.hlpark + ((size_t)blockIdx.x * (PCGF_THREADS / 32) + warp) * 9 * 32 + lane;
  {
    double HLc[18];
#pragma unroll
    for (int m = 0; m < 6; ++m) {
      const int idx = 32 * m + lane, edge = idx / 6, k = idx - 6 * edge;
      if (edge < min(deg, 32)) {
        uoff[m] = 6 * G.pl[le0 + edge].p + k;
#pragma unroll
        for (int r = 0; r < 3; ++r) HLc[3 * m + r] = G.HplL[18 * (size_t)(le0 + edge) + 6 * r + k];
      } else {
        uoff[m] = -1;
#pragma unroll
        for (int r = 0; r < 3; ++r) HLc[3 * m + r] = 0.0;
      }
    }
#pragma unroll
    for (int c = 0; c < 9; ++c) park[32 * c] = make_double2(HLc[2 * c], HLc[2 * c + 1]);
  }
  // edges 32..63 of a landmark: blocks and pose ids kept in shared memory
  const int nov = max(0, deg - 32);
  if (lane == 0) ovcnt[warp] = nov;
  __syncthreads();
  int ovbase = 0;
  for (int w = 0; w < warp; ++w) ovbase += ovcnt[w];
  for (int idx = lane; idx < 6 * nov; idx += 32) {
    const int edge = idx / 6, k = idx - 6 * edge, e = le0 + 32 + edge;
    ov_cell[6 * ovbase + idx] = 6 * G.pl[e].p + k;
#pragma unroll
    for (int r = 0; r < 3; ++r) ovH[3 * (6 * ovbase + idx) + r] = G.HplL[18 * (size_t)e + 6 * r + k];
  }
  __syncthreads();

  bool use_glob = false;
  if constexpr (MR) {
    use_glob = FP.glob != 0 && use_coarse;
    if (use_glob) {
      for (int k = threadIdx.x; k < 36 * FP.world; k += PCGF_THREADS) ag_sh[k] = FP.Aginv[k];
      if (threadIdx.x < 3) dga_sh[threadIdx.x] = Cz.cen[3 * blockIdx.x + threadIdx.x] - FP.gcent[4 * FP.rank + threadIdx.x];
    }
    __syncthreads();
  }
  // MR + rank-level coarse level: publish (v0, v1, E_a' s6) of this CTA into the glines of every rank and fold the lines of
  // every rank (two warps... 16 / world warps per rank, 8 / that many values each) into rs_sh.  Called by all threads.
  auto glob_exchange = [&](double v0, double v1, const double* s6v /* smem, 6 */, unsigned tag) {
    if constexpr (MR) {
      const int W = FP.world;
      if (warp == 0) {
        // lane k < 8 holds value k: (v0, v1, s_t, d x s_t + s_r)
        double val = lane == 0 ? v0 : (lane == 1 ? v1 : (lane < 8 ? s6v[lane - 2] : 0.0));
        const double st0 = s6v[0], st1 = s6v[1], st2 = s6v[2];
        if (lane == 5) val += dga_sh[1] * st2 - dga_sh[2] * st1;
        if (lane == 6) val += dga_sh[2] * st0 - dga_sh[0] * st2;
        if (lane == 7) val += dga_sh[0] * st1 - dga_sh[1] * st0;
        const double gv = __shfl_sync(0xffffffffu, val, lane & 7);
        for (int r0 = 0; r0 < W; r0 += 4) {
          const int rr = r0 + (lane >> 3);
          if (rr < W) st_cell_sys(FP.glines[rr] + (((size_t)(tag & 1u) * W + FP.rank) * nblk + blockIdx.x) * 8 + (lane & 7), gv, tag);
        }
      }
      constexpr int RM = (NB + 31) / 32;
      const int wpr = (PCGF_THREADS / 32) / W;        // warps per rank: 8, 4, 2 for 2, 4, 8 ranks
      const int rr = warp / wpr, sub = warp - rr * wpr, nval = 8 / wpr;
      if (rr < W) {
        const uint4* Lr = FP.glines[FP.rank] + ((size_t)(tag & 1u) * W + rr) * nblk * 8;
        for (int q = 0; q < nval; ++q) {
          const int k = sub * nval + q;
          uint4 c[RM];
          int ro[RM];
          double rv[RM];
#pragma unroll
          for (int m = 0; m < RM; ++m) {
            const int cta = lane + 32 * m;
            ro[m] = cta < nblk ? 8 * cta + k : -1;
            if (ro[m] >= 0) c[m] = ld_cell_sys(Lr + ro[m]);
          }
          cells_wait<RM, true>(Lr, ro, c, tag, rv);
          double t = 0.0;
#pragma unroll
          for (int m = 0; m < RM; ++m)
            if (ro[m] >= 0) t += rv[m];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
          if (lane == 0) rs_sh[8 * rr + k] = t;
        }
      }
    }
  };
  // tg6 = (my rank's rows of w_g A_g^-1) . P_g'w; call after a barrier that follows glob_exchange, barrier afterwards
  auto glob_rows = [&]() {
    if constexpr (MR) {
      if (warp < 6) {
        const int n = 6 * FP.world;
        double t = 0.0;
        for (int j = lane; j < n; j += 32) t += (double)ag_sh[warp * n + j] * rs_sh[8 * (j / 6) + 2 + (j % 6)];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) tg6_sh[warp] = t;
      }
    }
  };
  // E_a (tau, omega) = (tau - d x omega, omega): the rank-level correction in the coordinates of my CTA aggregate
  auto glob_prolong = [&](double* eg) {
    const double t0 = tg6_sh[0], t1 = tg6_sh[1], t2 = tg6_sh[2], w0 = tg6_sh[3], w1 = tg6_sh[4], w2 = tg6_sh[5];
    eg[0] = t0 - (dga_sh[1] * w2 - dga_sh[2] * w1);
    eg[1] = t1 - (dga_sh[2] * w0 - dga_sh[0] * w2);
    eg[2] = t2 - (dga_sh[0] * w1 - dga_sh[1] * w0);
    eg[3] = w0;
    eg[4] = w1;
    eg[5] = w2;
  };
  // ---- init: x = 0, r = g, u = M^-1 r ------------------------------------------------------------
  const unsigned tb = F.tagbase;
  uint4* const my_ucell = F.ucell + 6 * (size_t)(act ? i : 0) + comp;
  double* const my_ush = u_sh + 6 * (warp * 5 + slot) + comp;
  // MR: ranks that hold my keyframe as a ghost (the first two targets in registers, the rest from the table),
  // and (landmark role, lanes 3..) the ranks whose keyframes see my part's landmark: lane 3 + 3 q + c pushes
  // component c to target q
  int upn = 0, up0i = 0;
  uint4 *upa = nullptr, *upb = nullptr, *vpush = nullptr;
  if constexpr (MR) {
    if (act) {
      up0i = FP.upush_rowptr[i];
      upn = FP.upush_rowptr[i + 1] - up0i;
      if (upn > 0) upa = FP.upush_cell[up0i] + comp;
      if (upn > 1) upb = FP.upush_cell[up0i + 1] + comp;
    }
    if (lact && lane >= 3) {
      const int q = (lane - 3) / 3, v0 = FP.vpush_rowptr[part];
      if (q < FP.vpush_rowptr[part + 1] - v0) vpush = FP.vpush_cell[v0 + q] + (lane - 3 - 3 * q);
    }
  }
#define SSB_FLOW_PUBLISH_U(tag_)                                                              \
  if (act) {                                                                                  \
    st_cell(my_ucell, uc, (tag_));                                                            \
    *my_ush = uc;                                                                             \
    if constexpr (MR) {                                                                       \
      if (upn > 0) st_cell_sys(upa, uc, (tag_));                                              \
      if (upn > 1) st_cell_sys(upb, uc, (tag_));                                              \
      for (int q_ = 2; q_ < upn; ++q_) st_cell_sys(FP.upush_cell[up0i + q_] + comp, uc, (tag_)); \
    }                                                                                         \
  }
  double xc = 0.0;
  double rcomp = act ? G.g[6 * (size_t)i + comp] : 0.0;
  double uc = 0.0, pc = 0.0, sc = 0.0, cz = 0.0, cy = 0.0, wv = 0.0;
  if (use_coarse) {
    double l6[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) l6[k] = Brow[k] * rcomp;
    block_sum6(l6, s6, red6);
    grid_bar_sum(slots, epoch, 0.0, s6, wc, part_sh);  // wc = P'r (all aggregates)
    if (warp < 6) {
      double t = 0.0;
      for (int j = lane; j < nc; j += 32) t += Arow[warp * nc + j] * wc[j];
      t = warp_sum(t);
      if (lane == 0) zc6[warp] = t;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 6; ++k) cz += Brow[k] * zc6[k];
    if constexpr (MR) {
      if (use_glob) {   // z_g = A_g^-1 P_g'r, exchanged with the tag of the launch itself
        glob_exchange(0.0, 0.0, s6, tb);
        __syncthreads();
        glob_rows();
        __syncthreads();
        double eg[6];
        glob_prolong(eg);
#pragma unroll
        for (int k = 0; k < 6; ++k) cz += Brow[k] * eg[k];
      }
    }
  }
#define SSB_FLOW_PRECOND()                                                                              \
  {                                                                                                     \
    double _z = 0.0;                                                                                    \
    _Pragma("unroll") for (int k = 0; k < 6; ++k) _z += Drow[k] * __shfl_sync(0xffffffffu, rcomp, base_lane + k); \
    if (use_sub) {                                                                                      \
      double _t[8], _b[6];                                                                              \
      lds_row6(b1addr, _b);                                                                             \
      _Pragma("unroll") for (int k = 0; k < 6; ++k) _t[k] = _b[k] * rcomp; /* r = 0 on inactive lanes */  \
      _t[6] = 0.0;                                                                                      \
      _t[7] = 0.0;                                                                                      \
      const double _r = warp_reduce8(_t); /* P1'r: value k in lanes 4k..4k+3 */                          \
      if (use_grp) {                                                                                    \
        /* exact solve inside each group of aggregates: z5 = Ginv r5 (packed symmetric, 4 threads per row) */ \
        if ((lane & 3) == 0 && lane < 24) r5_sh[6 * warp + (lane >> 2)] = _r;                             \
        __syncthreads();                                                                                \
        {                                                                                               \
          const int _row = threadIdx.x >> 2, _part = threadIdx.x & 3;                                   \
          double _s = 0.0;                                                                              \
          if (_row < gn0 + gn1) {                                                                       \
            const int _h = _row >= gn0, _i = _row - (_h ? gn0 : 0), _nh = _h ? gn1 : gn0, _base = _h ? gn0 : 0; \
            const double* _G = GI + _h * GRP_PACK;                                                       \
            const int _chunk = (_nh + 3) >> 2, _j0 = _part * _chunk, _j1 = min(_nh, _j0 + _chunk);       \
            for (int _j = _j0; _j < _j1; ++_j) {                                                         \
              const int _hi = max(_i, _j), _lo = min(_i, _j);                                            \
              _s += _G[((_hi * (_hi + 1)) >> 1) + _lo] * r5_sh[_base + _j];                               \
            }                                                                                           \
          }                                                                                             \
          _s += __shfl_xor_sync(0xffffffffu, _s, 1);                                                     \
          _s += __shfl_xor_sync(0xffffffffu, _s, 2);                                                     \
          if (_part == 0 && _row < 6 * (PCGF_THREADS / 32)) z5_sh[_row] = _s;                            \
        }                                                                                               \
        __syncthreads();                                                                                \
        _Pragma("unroll") for (int k = 0; k < 6; ++k) _z += _b[k] * z5_sh[6 * warp + k];                   \
      } else {                                                                                          \
        double _m[6];                                                                                   \
        lds_row6(b1addr + 8u * 36u * PCGW_POSES, _m);                                                   \
        _Pragma("unroll") for (int k = 0; k < 6; ++k) _z += _m[k] * __shfl_sync(0xffffffffu, _r, 4 * k);  \
      }                                                                                                 \
    }                                                                                                   \
    uc = act ? _z + cz : 0.0;                                                                           \
  }
  SSB_FLOW_PRECOND()
  SSB_FLOW_PUBLISH_U(tb + 1u)

  double gamma = 0.0, gamma0 = 0.0, inv_gamma_old = 1.0, inv_alpha = 1.0;
  bool rep_bad = false, rep_stalled = false, rep_started = false;   // REP: my block broke down / stagnated / has taken its first step
  int it = 0;
#ifdef SSB_PCG_TIMERS
  long long tmr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tlast = clock64();
  tmr[7] = tlast - t_asm;          // Gauss-Jordan + operand load + init
  tmr[6] = -(t_asm - t_start);     // (negative) coarse assembly, folded into slot 6 for display
#define SSB_FTICK(k)          \
  do {                        \
    long long _n = clock64(); \
    tmr[k] += _n - tlast;     \
    tlast = _n;               \
  } while (0)
#elif defined(SSB_FLOW_TRACE)
#define SSB_FTICK(k)                                                             \
  do {                                                                           \
    if (it == 40 && threadIdx.x == 0) {                                          \
      unsigned long long _t;                                                     \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t));                     \
      F.trace[8 * blockIdx.x + (k)] = _t;                                        \
    }                                                                            \
  } while (0)
#else
#define SSB_FTICK(k) \
  do {               \
  } while (0)
#endif
  const int ngather = 8 * nblk;
  for (it = 0;; ++it) {
    const unsigned tg = tb + (unsigned)it + 1u;
#ifdef SSB_FLOW_TRACE
    SSB_FTICK(7);
#endif
    // ---- A (landmark role): v = W_l sum_e HplL_e u_p(e) ------------------------------------------
    if (lact) {
      double a[4] = {0.0, 0.0, 0.0, 0.0};
      uint4 c[6];
#pragma unroll
      for (int m = 0; m < 6; ++m)
        if (uoff[m] >= 0) c[m] = ld_cell_t<MR>(F.ucell + uoff[m]);
      double HLc[18];
#pragma unroll
      for (int q = 0; q < 9; ++q) {
        const double2 h = __ldcg(park + 32 * q);
        HLc[2 * q] = h.x;
        HLc[2 * q + 1] = h.y;
      }
      double uv[6];
      cells_wait<6, MR>(F.ucell, uoff, c, tg, uv);
#pragma unroll
      for (int m = 0; m < 6; ++m) {   // uv = 0 and HLc = 0 for unused items
        a[0] += HLc[3 * m] * uv[m];
        a[1] += HLc[3 * m + 1] * uv[m];
        a[2] += HLc[3 * m + 2] * uv[m];
      }
      if (nov > 0) {  // warp-uniform: edges 32..63, same item order, second batch of loads
        int oc[6];
#pragma unroll
        for (int m = 0; m < 6; ++m) {
          const int idx = 32 * m + lane;
          oc[m] = -1;
          if (idx < 6 * nov) {
            oc[m] = ov_cell[6 * ovbase + idx];
            c[m] = ld_cell_t<MR>(F.ucell + oc[m]);
          }
        }
        cells_wait<6, MR>(F.ucell, oc, c, tg, uv);
#pragma unroll
        for (int m = 0; m < 6; ++m)
          if (oc[m] >= 0) {
            const double* H2 = ovH + 3 * (6 * ovbase + 32 * m + lane);
            a[0] += H2[0] * uv[m];
            a[1] += H2[1] * uv[m];
            a[2] += H2[2] * uv[m];
          }
      }
      const double rsum = warp_reduce4(a);  // value j in lanes 8j..8j+7
      const double a0 = __shfl_sync(0xffffffffu, rsum, 0);
      const double a1 = __shfl_sync(0xffffffffu, rsum, 8);
      const double a2 = __shfl_sync(0xffffffffu, rsum, 16);
      const double vval = Wr[0] * a0 + Wr[1] * a1 + Wr[2] * a2;   // valid in lanes 0..2
      if (lane < 3) st_cell(F.vcell + 3 * (size_t)part + lane, vval, tg);
      if constexpr (MR) {
        const double vv = __shfl_sync(0xffffffffu, vval, lane >= 3 ? (lane - 3) % 3 : 0);
        if (vpush) st_cell_sys(vpush, vv, tg);
      }
    }
    SSB_FTICK(0);
    // ---- stage u of external neighbours and v of my poses' landmarks into shared memory ----------------
    {
      uint4 c[3];
      int sc_[3];
#pragma unroll
      for (int m = 0; m < 3; ++m) {
        const int q = threadIdx.x + PCGF_THREADS * m;
        sc_[m] = -1;
        if (q < nstage) {
          sc_[m] = stage_src[q];
          c[m] = ld_cell_t<MR>(F.ucell + (sc_[m] & 0xffffff));
        }
      }
      int so[3];
      double sv[3];
#pragma unroll
      for (int m = 0; m < 3; ++m) so[m] = sc_[m] >= 0 ? (sc_[m] & 0xffffff) : -1;
      cells_wait<3, MR>(F.ucell, so, c, tg, sv);
#pragma unroll
      for (int m = 0; m < 3; ++m)
        if (sc_[m] >= 0) {
          const int np = sc_[m] >> 24;   // 0 for u cells, >= 1 for v cells (parts of one landmark are adjacent)
          double val = sv[m];
          for (int j = 1; j < np; ++j)
            val += cell_wait<MR>(F.ucell + so[m] + 3 * j, ld_cell_t<MR>(F.ucell + so[m] + 3 * j), tg);
          u_sh[6 * PCGW_POSES + threadIdx.x + PCGF_THREADS * m] = val;
        }
    }
    SSB_FTICK(1);
    __syncthreads();
    // ---- B (pose role): w = (Hpp + lambda I) u + sum Hoff u_nbr - sum HplP v  (all operands on chip) ---
    {
      wv = 0.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) wv += Hrow[k] * __shfl_sync(0xffffffffu, uc, base_lane + k);
      for (int s = mypp0; s < mypp1; ++s) {
        const int code = pp_src[s];
        const double* us = u_sh + 6 * (code & 0x7fffffff);
        const double* Ho = ppH + 36 * pp_loc[s];
        if (code < 0) {  // role 1: my pose is vertex j of the edge -> Hoff'
#pragma unroll
          for (int k = 0; k < 6; ++k) wv += Ho[6 * k + comp] * us[k];
        } else {
#pragma unroll
          for (int k = 0; k < 6; ++k) wv += Ho[6 * comp + k] * us[k];
        }
      }
      for (int kk = mypl0; kk < mypl1; ++kk) {
        const double* vv = v_sh + 3 * pl_loc[kk];
        const double* Hp = plH + 18 * kk + 3 * comp;
        wv -= Hp[0] * vv[0] + Hp[1] * vv[1] + Hp[2] * vv[2];
      }
      if (!act) wv = 0.0;
    }
    // ---- C: one fused reduction: gamma = r'u, delta = w'u, P'w (6 per CTA) ---------------------------
    {
      double t8[8];
      t8[0] = rcomp * uc;
      if constexpr (REP) {
        if (rep_bad) t8[0] = __longlong_as_double(0x7ff8000000000000LL);   // tells every CTA that this block broke down
        if (rep_stalled) t8[0] = 0.0;                                      // ... that it has gone as far as doubles allow
      }
      t8[1] = wv * uc;
#pragma unroll
      for (int k = 0; k < 6; ++k) t8[2 + k] = Brow[k] * wv;
      const double rs = warp_reduce8(t8);
      if ((lane & 3) == 0) scratch8[8 * warp + (lane >> 2)] = rs;
    }
    SSB_FTICK(2);
    __syncthreads();
    const int mr_world = MR ? FP.world : 1, mr_rank = MR ? FP.rank : 0;
    if (warp == 0) {
      double t = 0.0;
      if (lane < 8) {
#pragma unroll
        for (int ww = 0; ww < PCGF_THREADS / 32; ++ww) t += scratch8[8 * ww + lane];
        st_cell(F.lines + (((size_t)(tg & 1u) * mr_world + mr_rank) * nblk + blockIdx.x) * 8 + lane, t, tg);
        if constexpr (MR) gl8_sh[lane] = t;
      }
      if constexpr (MR) {
        __syncwarp();
        if (!use_glob) {
          // the two dot-product partials also go into the same line slot of every other rank: lane 8 + 2 j + k
          const double tv = __shfl_sync(0xffffffffu, t, lane >= 8 ? ((lane - 8) & 1) : 0);
          const int j = (lane - 8) >> 1;
          if (lane >= 8 && j < mr_world - 1) {
            const int peer = j < mr_rank ? j : j + 1;
            st_cell_sys(FP.lines[peer] + (((size_t)(tg & 1u) * mr_world + mr_rank) * nblk + blockIdx.x) * 8 + ((lane - 8) & 1), tv, tg);
          }
        }
      }
    }
    if constexpr (MR) {
      if (use_glob) glob_exchange(gl8_sh[0], gl8_sh[1], gl8_sh + 2, tg);   // warp 0 wrote gl8_sh itself
    }
    {
      // pull all-gather of the nblk lines: thread q reads cell q (coalesced), value k of line q/8
      const uint4* L = F.lines + ((size_t)(tg & 1u) * mr_world + mr_rank) * nblk * 8;
      uint4 c[3];
#pragma unroll
      for (int m = 0; m < 3; ++m) {
        const int q = threadIdx.x + PCGF_THREADS * m;
        if (q < ngather) c[m] = ld_cell(L + q);
      }
      int go[3];
      double gv[3];
#pragma unroll
      for (int m = 0; m < 3; ++m) go[m] = (threadIdx.x + PCGF_THREADS * m < ngather) ? threadIdx.x + PCGF_THREADS * m : -1;
      cells_wait<3>(L, go, c, tg, gv);
#pragma unroll
      for (int m = 0; m < 3; ++m) {
        const int q = threadIdx.x + PCGF_THREADS * m;
        if (q < ngather) {
          const double val = gv[m];
          const int ln = q >> 3, k = q & 7;
          if (k == 0)
            gam[ln] = val;
          else if (k == 1)
            del[ln] = val;
          else
            wc[6 * ln + k - 2] = val;
        }
      }
      for (int q = threadIdx.x + 3 * PCGF_THREADS; q < ngather; q += PCGF_THREADS) {  // grids > 192 CTAs
        const double val = cell_wait(L + q, ld_cell(L + q), tg);
        const int ln = q >> 3, k = q & 7;
        if (k == 0)
          gam[ln] = val;
        else if (k == 1)
          del[ln] = val;
        else
          wc[6 * ln + k - 2] = val;
      }
    }
    if constexpr (MR) {
      // the other ranks: two warps per rank (gamma, delta) fold its nblk partials straight from the cells, with the
      // same summation tree that rank uses for its own lines => every CTA of every rank gets identical bits
      static_assert(PCGF_THREADS / 32 >= 2 * (SSB_MAX_WORLD - 1), "two warps per remote rank");
      constexpr int RM = (NB + 31) / 32;
      const int j = warp >> 1, kq = warp & 1;
      if (!use_glob && j < mr_world - 1) {
        const int peer = j < mr_rank ? j : j + 1;
        const uint4* Lr = F.lines + ((size_t)(tg & 1u) * mr_world + peer) * nblk * 8;
        uint4 c[RM];
        int ro[RM];
        double rv[RM];
#pragma unroll
        for (int m = 0; m < RM; ++m) {
          const int cta = lane + 32 * m;
          ro[m] = cta < nblk ? 8 * cta + kq : -1;
          if (ro[m] >= 0) c[m] = ld_cell_sys(Lr + ro[m]);
        }
        cells_wait<RM, true>(Lr, ro, c, tg, rv);
        double t = 0.0;
#pragma unroll
        for (int m = 0; m < RM; ++m)
          if (ro[m] >= 0) t += rv[m];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) rs_sh[8 * peer + kq] = t;
      }
    }
    SSB_FTICK(3);
    __syncthreads();
    SSB_FTICK(4);
    double delta = 0.0;
    if constexpr (!REP) {
      double tg_ = 0.0, td_ = 0.0;
      for (int k = lane; k < nblk; k += 32) {
        tg_ += gam[k];
        td_ += del[k];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        tg_ += __shfl_xor_sync(0xffffffffu, tg_, o);
        td_ += __shfl_xor_sync(0xffffffffu, td_, o);
      }
      gamma = tg_;
      delta = td_;
      if constexpr (MR) {   // rank order, identical on every rank
        double g2 = 0.0, d2 = 0.0;
        for (int r = 0; r < mr_world; ++r) {
          const bool own = r == mr_rank && !use_glob;   // with the rank-level level every rank, mine included, comes from glines
          g2 += own ? tg_ : rs_sh[8 * r];
          d2 += own ? td_ : rs_sh[8 * r + 1];
        }
        gamma = g2;
        delta = d2;
      }
    } else {
      // warp j folds the lines of block j (the same tree in every CTA => identical bits => identical decisions everywhere)
      if (warp < F.rep_count) {
        const int c0 = warp * F.rep_ctas, c1 = c0 + F.rep_ctas;
        double tg_ = 0.0, td_ = 0.0;
        for (int k = c0 + lane; k < c1; k += 32) {
          tg_ += gam[k];
          td_ += del[k];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          tg_ += __shfl_xor_sync(0xffffffffu, tg_, o);
          td_ += __shfl_xor_sync(0xffffffffu, td_, o);
        }
        if (lane == 0) {
          rg_sh[warp] = tg_;
          rd_sh[warp] = td_;
          if (it == 0) {
            g0_sh[warp] = tg_;
            frz_sh[warp] = (tg_ > 0.0) ? 0 : (tg_ == 0.0 ? 1 : 3);   // a zero right-hand side is solved already
          } else if (frz_sh[warp] == 0) {
            if (!isfinite(tg_))
              frz_sh[warp] = 2;
            else if (!(tg_ > tol2 * g0_sh[warp]))
              frz_sh[warp] = 1;
          }
        }
      }
      if (!use_coarse) __syncthreads();   // (with a coarse level its barrier below publishes the fold)
    }
    if (use_coarse) {
      if (warp < 12) {  // two warps per row of A_c^-1
        const int row = warp >> 1, j0 = (warp & 1) * (nc >> 1), j1 = j0 + (nc >> 1);
        double t = 0.0;
        for (int j = j0 + lane; j < j1; j += 32) t += Arow[row * nc + j] * wc[j];
        t = warp_sum(t);
        if (lane == 0) dpart[warp] = t;
      }
      if constexpr (MR) {
        if (use_glob) glob_rows();
      }
      __syncthreads();
    }
    SSB_FTICK(5);
    // ---- scalars (identical in every thread of every CTA; REP: of every CTA of a block) ---------------
    double alpha = 0.0, beta = 0.0;
    if constexpr (!REP) {
      if (it == 0) {
        gamma0 = gamma;
        if (!(gamma0 > 0.0)) {
          status = (gamma0 == 0.0) ? 0 : 2;
          break;
        }
      }
      if (!(gamma > tol2 * gamma0)) break;
      if (it >= maxit) break;
      // beta = gamma/gamma_old, alpha = gamma / (delta - beta*gamma/alpha_old) with one division on the
      // critical path (1/gamma is independent of it)
      const double inv_gamma = fast_rcp(gamma);
      beta = (it == 0) ? 0.0 : gamma * inv_gamma_old;
      const double den = (it == 0) ? delta : delta - beta * gamma * inv_alpha;
      if (!(den > 0.0) || !isfinite(den)) {
        status = 1;
        break;
      }
      alpha = gamma * fast_rcp(den);
      inv_alpha = den * inv_gamma;
      inv_gamma_old = inv_gamma;
    } else {
      const int myrep = min((int)blockIdx.x / F.rep_ctas, F.rep_count - 1);
      bool all_done = true;
      for (int j = 0; j < F.rep_count; ++j) {
        const int f = frz_sh[j];
        all_done = all_done && f != 0;
        if (f == 2) status = 1;
        if (f == 3) status = 2;
      }
      gamma = rg_sh[myrep];
      delta = rd_sh[myrep];
      if (it == 0) gamma0 = g0_sh[myrep];
      if (all_done) break;
      if (it >= maxit) break;
      if (frz_sh[myrep] == 0 && !rep_bad && !rep_stalled) {
        const double inv_gamma = fast_rcp(gamma);
        beta = rep_started ? gamma * inv_gamma_old : 0.0;
        const double den = rep_started ? delta - beta * gamma * inv_alpha : delta;
        if (!(den > 0.0) || !isfinite(den)) {
          // freeze.  Once the residual is down by 1e-7 in the M^-1 norm a non-positive curvature is rounding noise at the
          // attainable accuracy (tolerances near 1e-12 on a lambda = 0 system): the block counts as converged (it publishes
          // gamma = 0).  Earlier it is a breakdown: the NaN published with the next line makes every CTA record it.
          if (isfinite(den) && gamma <= 1e-14 * gamma0)
            rep_stalled = true;
          else
            rep_bad = true;
          beta = 0.0;
        } else {
          alpha = gamma * fast_rcp(den);
          inv_alpha = den * inv_gamma;
          inv_gamma_old = inv_gamma;
          rep_started = true;
        }
      }
    }
    // ---- D: recurrences and u = M^-1 r ----------------------------------------------------------------
    pc = uc + beta * pc;
    sc = wv + beta * sc;
    xc += alpha * pc;
    rcomp -= alpha * sc;
    if (use_coarse) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) t += Brow[k] * (dpart[2 * k] + dpart[2 * k + 1]);
      if constexpr (MR) {
        if (use_glob) {
          double eg[6];
          glob_prolong(eg);
#pragma unroll
          for (int k = 0; k < 6; ++k) t += Brow[k] * eg[k];
        }
      }
      cy = t + beta * cy;
      cz -= alpha * cy;
    }
    SSB_FLOW_PRECOND()
    SSB_FLOW_PUBLISH_U(tg + 1u)
    SSB_FTICK(6);
  }
#undef SSB_FLOW_PRECOND
#undef SSB_FLOW_PUBLISH_U
  if (act) {
    G.x[6 * (size_t)i + comp] = xc;
    if constexpr (MR)   // the solution of a boundary keyframe goes to every rank that updates it as a ghost
      for (int q = 0; q < upn; ++q) st_relaxed_sys_f64(FP.upush_x[up0i + q] + comp, xc);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    G.iscalars[0] = it;
    G.iscalars[1] = status;
    G.scalars[3] = gamma;
    G.scalars[4] = gamma0;
#ifdef SSB_PCG_TIMERS
    for (int k = 0; k < 8; ++k) G.scalars[8 + k] += (double)tmr[k];
#endif
  }
}

}  // namespace ssb
