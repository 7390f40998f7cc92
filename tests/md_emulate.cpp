// CPU emulation of semantic_slam_b200/csrc/ssb_marg_direct.cuh (the direct form of the landmark marginals): the SAME kernel
// source and the SAME launch sequence (ssb_md::md_run), with every CUDA thread of a CTA run by a std::thread and
// __syncthreads() by a std::barrier — so that indexing, arithmetic and barrier placement are checked against the oracle on
// machines without a GPU (tests/test_marg_direct.py).  Data races between CTAs of one launch are outside what it can see.
//   usage: md_emulate <problem.bin> <out.bin>
#include <barrier>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __shared__ static
#define __restrict__
#define __launch_bounds__(x)
#define __forceinline__ inline
struct EmuDim { unsigned x = 1, y = 1, z = 1; };
static thread_local EmuDim threadIdx, blockIdx, blockDim, gridDim;
static std::barrier<>* g_cta_bar = nullptr;
static inline void __syncthreads() { g_cta_bar->arrive_and_wait(); }
// warp shuffle, for CTAs of exactly one warp (k_md_factor): every thread deposits its value, all meet, every thread reads
static double g_shfl_slots[32];
static inline double __shfl_sync(unsigned, double v, int src) {
  g_shfl_slots[threadIdx.x] = v;
  g_cta_bar->arrive_and_wait();
  const double r = g_shfl_slots[src];
  g_cta_bar->arrive_and_wait();
  return r;
}

#include "../semantic_slam_b200/csrc/ssb_marg_direct.cuh"

namespace {
constexpr int POOL = 256;
struct Pool {
  std::vector<std::thread> th;
  std::barrier<> start{POOL + 1}, done{POOL + 1};
  std::function<void()> job;
  int nthreads = 0;
  EmuDim bidx, gdim;
  bool quit = false;
  std::unique_ptr<std::barrier<>> bars[POOL + 1];
  Pool() {
    for (int n : {32, 64, 128, 256}) bars[n] = std::make_unique<std::barrier<>>(n);
    for (int id = 0; id < POOL; ++id)
      th.emplace_back([this, id] {
        for (;;) {
          start.arrive_and_wait();
          if (quit) return;
          if (id < nthreads) {
            threadIdx.x = (unsigned)id;
            blockIdx = bidx;
            gridDim = gdim;
            blockDim.x = (unsigned)nthreads;
            job();
          }
          done.arrive_and_wait();
        }
      });
  }
  ~Pool() {
    quit = true;
    start.arrive_and_wait();
    for (auto& t : th) t.join();
  }
  void cta(int bx, int by, int gx, int gy, int block, const std::function<void()>& f) {
    if (!bars[block]) {
      std::fprintf(stderr, "emulator: unsupported block size %d\n", block);
      std::exit(2);
    }
    job = f;
    nthreads = block;
    bidx.x = (unsigned)bx;
    bidx.y = (unsigned)by;
    gdim.x = (unsigned)gx;
    gdim.y = (unsigned)gy;
    g_cta_bar = bars[block].get();
    start.arrive_and_wait();
    done.arrive_and_wait();
  }
};
struct Emu {
  Pool pool;
  long launches = 0;
  template <class K, class... A>
  void operator()(K kern, int gx, int gy, int block, A... args) {
    ++launches;
    for (int by = 0; by < gy; ++by)
      for (int bx = 0; bx < gx; ++bx) pool.cta(bx, by, gx, gy, block, [&] { kern(args...); });
  }
  void zero(void* p, size_t n) { std::memset(p, 0, n); }
};
template <class T>
std::vector<T> rd(FILE* f, size_t n) {
  std::vector<T> v(n);
  if (n && std::fread(v.data(), sizeof(T), n, f) != n) {
    std::fprintf(stderr, "short read\n");
    std::exit(2);
  }
  return v;
}
}  // namespace

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 2;
  // header: Np Nl El Epp n_inc n_req
  std::vector<int> h = rd<int>(f, 6);
  const int Np = h[0], Nl = h[1], El = h[2], Epp = h[3], n_inc = h[4], n_req = h[5];
  auto pp_rowptr = rd<int>(f, Np + 1), pp_idx = rd<int>(f, n_inc), pp_other = rd<int>(f, n_inc);
  auto lm_rowptr = rd<int>(f, Nl + 1), edge_pose = rd<int>(f, El), lidx = rd<int>(f, n_req);
  auto Hoff = rd<double>(f, (size_t)36 * Epp), Hpp = rd<double>(f, (size_t)36 * Np), Hll = rd<double>(f, (size_t)6 * Nl),
       HplL = rd<double>(f, (size_t)18 * El);
  std::fclose(f);
  const ssb_md::MdDims d = ssb_md::md_dims(Np, Nl);
  std::vector<double> Bsub((size_t)36 * Np), Ginv((size_t)36 * Np), Esub((size_t)36 * Np), Y((size_t)d.K * d.ld, 7.0),
      T((size_t)d.ld * d.ld, 7.0), Row((size_t)64 * d.ld, 7.0), ColT((size_t)64 * d.ld, 7.0), Pinv(64 * 64, 7.0), out((size_t)9 * n_req);
  std::vector<int> tile_k0(d.nt), status(2, 5);
  ssb_md::MdBuffers b{};
  b.pose_pp_rowptr = pp_rowptr.data();
  b.pose_pp_idx = pp_idx.data();
  b.pose_pp_other = pp_other.data();
  b.Hoff = Hoff.data();
  b.Hpp = Hpp.data();
  b.Hll = Hll.data();
  b.HplL = HplL.data();
  b.edge_pose = edge_pose.data();
  b.edge_stride = 1;
  b.lm_rowptr = lm_rowptr.data();
  b.lidx = lidx.data();
  b.n_req = n_req;
  b.Bsub = Bsub.data();
  b.Ginv = Ginv.data();
  b.Esub = Esub.data();
  b.Y = Y.data();
  b.T = T.data();
  b.Row = Row.data();
  b.ColT = ColT.data();
  b.Pinv = Pinv.data();
  b.tile_k0 = tile_k0.data();
  b.status = status.data();
  b.out9n = out.data();
  Emu L;
  ssb_md::md_run(L, d, b);
  std::printf("emulated %ld launches: Np %d Nl %d ld %d K %d status %d %d\n", L.launches, Np, Nl, d.ld, d.K, status[0], status[1]);
  FILE* o = std::fopen(argv[2], "wb");
  if (!o) return 2;
  std::fwrite(status.data(), sizeof(int), 2, o);
  std::fwrite(out.data(), sizeof(double), out.size(), o);
  std::fclose(o);
  return 0;
}
