// ssb_peer.cuh — one graph sharded by contiguous keyframe range over several ranks.
//
// Every rank (one process per GPU, or one host thread per shard inside a process) maps the "peer arena" of
// every other rank (cudaIpc* between processes, plain pointers / cudaDeviceEnablePeerAccess inside one) and
// the kernels WRITE what a neighbour needs straight into that neighbour's arena over NVLink:
//   * tagged 16-byte cells of the data-flow PCG (u of boundary keyframes, v of shared landmarks, the two
//     dot-product partials of every CTA) — the consumer keeps polling its own memory, no collective call;
//   * plain vectors at kernel end (the solution of boundary keyframes), ordered by k_peer_exchange;
//   * small per-rank records (chi2 / scale / max-diagonal partials) gathered by k_peer_exchange, which is
//     also the cross-rank barrier: flag words holding monotonically increasing epochs, one writer per word.
// Nothing is ever pulled across NVLink in a spin loop.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ssb {

constexpr int SSB_MAX_WORLD = 8;
constexpr int PEER_RED_N = 8;   // doubles per rank in the small all-gather record
// record layout
enum { PR_CHI2 = 0, PR_SCALE = 1, PR_MAXDIAG = 2, PR_PCG_ITERS = 3, PR_PCG_STATUS = 4, PR_GAMMA = 5, PR_GAMMA0 = 6, PR_ERR = 7 };

struct PeerDev {
  int world, rank;
  unsigned long long* flags[SSB_MAX_WORLD];   // flags[r]: [world] epochs in rank r's arena, word s written by rank s
  double* red[SSB_MAX_WORLD];                 // red[r]:   [world][PEER_RED_N] records in rank r's arena
  uint4* lines[SSB_MAX_WORLD];                // lines[r]: [2][world][NB][8] reduction lines of the data-flow PCG
  unsigned char* slots[SSB_MAX_WORLD];        // slots[r]: [2][world * NB] BarSlot + counter (streaming PCG)
};

__device__ __forceinline__ unsigned long long peer_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys_f64(double* p, double v) {
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void st_cell_sys(uint4* c, double v, unsigned tag) {
  const unsigned lo = (unsigned)__double2loint(v), hi = (unsigned)__double2hiint(v);
  asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(c), "r"(lo), "r"(tag), "r"(hi), "r"(tag)
               : "memory");
}
__device__ __forceinline__ uint4 ld_cell_sys(const uint4* c) {
  uint4 u;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(c));
  return u;
}
__device__ __forceinline__ void red_release_sys_add_u32(unsigned* p, unsigned v) {
  asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_relaxed_sys_add_u32(unsigned* p, unsigned v) {
  asm volatile("red.relaxed.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// A peer that never arrives (crashed process, mismatched call sequence) must not hang the GPU: every cross-rank
// wait gives up after this many nanoseconds of globaltimer and the kernel traps (sticky error on this rank only).
#ifndef SSB_PEER_TIMEOUT_NS
#define SSB_PEER_TIMEOUT_NS 20000000000ull
#endif

// Small all-gather + barrier across the ranks.  One warp: lane s < world writes this rank's record into rank s's
// arena, then its arrival epoch (release, system scope); every lane then waits for rank s's epoch in OUR arena
// (acquire).  When the kernel has completed, everything the peers wrote before their own call is visible here.
// scalars / iscalars: the rank-local reduction results of the kernels before it (may be null: pure barrier).
__global__ void __launch_bounds__(32) k_peer_exchange(PeerDev P, const double* scalars, const int* iscalars, unsigned long long epoch,
                                                      int* err_flag) {
  const int lane = threadIdx.x;
  if (lane < P.world) {
    if (scalars) {
      double rec[PEER_RED_N];
      rec[PR_CHI2] = scalars[0];
      rec[PR_SCALE] = scalars[1];
      rec[PR_MAXDIAG] = scalars[2];
      rec[PR_PCG_ITERS] = (double)iscalars[0];
      rec[PR_PCG_STATUS] = (double)iscalars[1];
      rec[PR_GAMMA] = scalars[3];
      rec[PR_GAMMA0] = scalars[4];
      rec[PR_ERR] = 0.0;
      double* dst = P.red[lane] + (size_t)P.rank * PEER_RED_N;
#pragma unroll
      for (int k = 0; k < PEER_RED_N; ++k) st_relaxed_sys_f64(dst + k, rec[k]);
    }
    __threadfence_system();
    st_release_sys_u64(P.flags[lane] + P.rank, epoch);
    const unsigned long long t0 = peer_globaltimer();
    const unsigned long long* mine = P.flags[P.rank] + lane;
    while (ld_acquire_sys_u64(mine) < epoch) {
      if (peer_globaltimer() - t0 > SSB_PEER_TIMEOUT_NS) {
        if (err_flag) *err_flag = 1 + lane;
        break;
      }
    }
  }
}

// own estimates -> the full-size replica in every rank's arena (all-gather by push; k_peer_exchange orders it)
struct PeerGather {
  int world;
  void* pose_full[SSB_MAX_WORLD];   // [Np_global] Pose records
  double* lm_full[SSB_MAX_WORLD];   // [Nl_global][4]
};

}  // namespace ssb
