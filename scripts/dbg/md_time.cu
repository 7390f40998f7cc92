// Per-kernel timing of the direct landmark-marginals pipeline (semantic_slam_b200/csrc/ssb_marg_direct.cuh) on a problem file
// written by tests/test_marg_direct.py::write_problem: the same kernels and launch sequence as the product, every launch
// bracketed by CUDA events.  Developer tool (not part of the product):
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scripts/dbg/md_time scripts/dbg/md_time.cu
//   scripts/dbg/md_time problem.bin
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "../../semantic_slam_b200/csrc/ssb_marg_direct.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); std::exit(3); } } while (0)

struct TimingLauncher {
  std::map<const void*, std::string> names;
  std::map<std::string, std::pair<double, int>> acc;
  cudaEvent_t e0, e1;
  bool timing = false;
  TimingLauncher() { CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); }
  template <class K, class... A>
  void operator()(K kern, int gx, int gy, int block, A... args) {
    if (timing) CK(cudaEventRecord(e0));
    kern<<<dim3((unsigned)gx, (unsigned)gy), block>>>(args...);
    if (timing) {
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      auto it = names.find((const void*)kern);
      auto& a = acc[it == names.end() ? "?" : it->second];
      a.first += ms;
      a.second++;
    }
  }
  void zero(void* p, size_t n) {
    if (timing) CK(cudaEventRecord(e0));
    CK(cudaMemsetAsync(p, 0, n));
    if (timing) {
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      acc["memset"].first += ms;
      acc["memset"].second++;
    }
  }
};
template <class T>
static std::vector<T> rd(FILE* f, size_t n) {
  std::vector<T> v(n);
  if (n && std::fread(v.data(), sizeof(T), n, f) != n) std::exit(2);
  return v;
}
template <class T>
static T* up(const std::vector<T>& v) {
  T* d = nullptr;
  CK(cudaMalloc((void**)&d, std::max<size_t>(v.size(), 1) * sizeof(T)));
  if (!v.empty()) CK(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}
template <class T>
static T* dalloc(size_t n) {
  T* d = nullptr;
  CK(cudaMalloc((void**)&d, std::max<size_t>(n, 1) * sizeof(T)));
  return d;
}

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 2;
  std::vector<int> h = rd<int>(f, 6);
  const int Np = h[0], Nl = h[1], El = h[2], Epp = h[3], n_inc = h[4], n_req = h[5];
  auto pp_rowptr = rd<int>(f, Np + 1), pp_idx = rd<int>(f, n_inc), pp_other = rd<int>(f, n_inc);
  auto lm_rowptr = rd<int>(f, Nl + 1), edge_pose = rd<int>(f, El), lidx = rd<int>(f, n_req);
  auto Hoff = rd<double>(f, (size_t)36 * Epp), Hpp = rd<double>(f, (size_t)36 * Np), Hll = rd<double>(f, (size_t)6 * Nl),
       HplL = rd<double>(f, (size_t)18 * El);
  std::fclose(f);
  const ssb_md::MdDims d = ssb_md::md_dims(Np, Nl);
  ssb_md::MdBuffers b{};
  b.pose_pp_rowptr = up(pp_rowptr);
  b.pose_pp_idx = up(pp_idx);
  b.pose_pp_other = up(pp_other);
  b.Hoff = up(Hoff);
  b.Hpp = up(Hpp);
  b.Hll = up(Hll);
  b.HplL = up(HplL);
  b.edge_pose = up(edge_pose);
  b.edge_stride = 1;
  b.lm_rowptr = up(lm_rowptr);
  b.lidx = up(lidx);
  b.n_req = n_req;
  b.Bsub = dalloc<double>((size_t)36 * Np);
  b.Ginv = dalloc<double>((size_t)36 * Np);
  b.Esub = dalloc<double>((size_t)36 * Np);
  b.Y = dalloc<double>((size_t)d.K * d.ld);
  b.T = dalloc<double>((size_t)d.ld * d.ld);
  b.Row = dalloc<double>((size_t)64 * d.ld);
  b.ColT = dalloc<double>((size_t)64 * d.ld);
  b.Pinv = dalloc<double>(64 * 64);
  b.tile_k0 = dalloc<int>(d.nt);
  b.status = dalloc<int>(2);
  b.out9n = dalloc<double>((size_t)9 * n_req);
  TimingLauncher L;
  using namespace ssb_md;
  L.names[(const void*)k_md_gather_sub] = "k_md_gather_sub";
  L.names[(const void*)k_md_factor] = "k_md_factor";
  L.names[(const void*)k_md_sweep] = "k_md_sweep";
  L.names[(const void*)k_md_tile_k0] = "k_md_tile_k0";
  L.names[(const void*)k_md_init_T] = "k_md_init_T";
  L.names[(const void*)k_md_gemm_tn] = "k_md_gemm_tn";
  L.names[(const void*)k_md_mirror] = "k_md_mirror";
  L.names[(const void*)k_md_gj_pivot] = "k_md_gj_pivot";
  L.names[(const void*)k_md_gj_col] = "k_md_gj_col";
  L.names[(const void*)k_md_gj_rowcopy] = "k_md_gj_rowcopy";
  L.names[(const void*)k_md_out] = "k_md_out";
  md_run(L, d, b);   // warm-up
  CK(cudaDeviceSynchronize());
  cudaEvent_t t0, t1;
  CK(cudaEventCreate(&t0));
  CK(cudaEventCreate(&t1));
  CK(cudaEventRecord(t0));
  md_run(L, d, b);
  CK(cudaEventRecord(t1));
  CK(cudaEventSynchronize(t1));
  float whole = 0;
  CK(cudaEventElapsedTime(&whole, t0, t1));
  L.timing = true;
  md_run(L, d, b);
  CK(cudaDeviceSynchronize());
  int st[2];
  CK(cudaMemcpy(st, b.status, sizeof(st), cudaMemcpyDeviceToHost));
  std::printf("Np %d Nl %d: T %d x %d (%d tiles), Y %d rows; whole pipeline back to back %.3f ms; status %d %d\n", Np, Nl, d.ld, d.ld, d.nt, d.K,
              whole, st[0], st[1]);
  double sum = 0;
  for (auto& kv : L.acc) sum += kv.second.first;
  for (auto& kv : L.acc)
    std::printf("  %-18s %4d launches %9.3f ms  (%.1f %%)\n", kv.first.c_str(), kv.second.second, kv.second.first, 100.0 * kv.second.first / sum);
  std::printf("  sum of the launches timed one by one: %.3f ms\n", sum);
  return 0;
}
