// ssb_cluster.cuh — the reference's dormant plane-clustering chain on the device:
//   plane_segmentation::clusterAndSegmentAllPlanes          src/planar_segmentation/plane_segmentation.cpp:261-294
//     NormalBasedClusteringAndSegmentation                  :296-367  (removeNans :479-502, filterCentroids :504-523)
//     distanceBasedSegmentation                             :369-429
//     getFinalPoseWithNormals                               :431-477
//     computeKmeans -> cv::kmeans                           :525-535
//     compute2DConvexHull: SACSegmentation + ProjectInliers + ConvexHull   :631-664
// cv::kmeans is a chain of order-dependent single-precision sums (centre += sample in point order), so it is evaluated with
// the SAME arithmetic in an order that respects the dependencies — labels and centres are bit-identical to OpenCV's:
//   attempts          : independent once their random centres are drawn (the only use of cv::RNG) -> one CTA per attempt
//   centre sums       : float sums in point order -> the CTA stages tiles of samples into shared memory, one warp per
//                       cluster adds the samples in order (non-members masked to +0.0f, which is exact), the K chains side
//                       by side
//   assignment        : independent per point (float distances, first minimum wins)
//   empty clusters    : block-wide lexicographic arg-max (distance, index) = OpenCV's "last farthest point"
// Selections (valid normals, members of a cluster, RANSAC inliers, hull candidates) are order-preserving compactions:
// one CTA ranks the flags chunk by chunk.  The convex hull keeps the work that is parallel on the device (projection, the
// octagon of extreme points that discards interior points) and finishes the few survivors on the host.
// Compiled with -fmad=false like the rest of ssb_ransac.cu.
#pragma once
#include <cfloat>

namespace ssb_cl {

constexpr int CL_THREADS = 1024;
constexpr int CL_MAXK = 8;      // clusters per k-means (the reference uses 4 and 2)
constexpr int CL_MAXD = 4;      // dimensions per sample (3 and 1)
constexpr int CL_TILE = 4096;    // samples staged into shared memory per round of the ordered centre sums
constexpr size_t CL_KMEANS_SMEM = (size_t)(CL_TILE + 8) * (CL_MAXD + 1) * sizeof(float);   // dynamic shared memory of k_cl_kmeans

// exclusive rank of every set flag, in index order; count[0] = number of set flags.  One CTA.
__global__ void __launch_bounds__(CL_THREADS) k_cl_rank(const unsigned char* __restrict__ flags, int n, int* __restrict__ pos,
                                                        int* __restrict__ count) {
  __shared__ int wsum[32];
  __shared__ int base_s;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  for (int c0 = 0; c0 < n; c0 += CL_THREADS) {
    const int i = c0 + threadIdx.x;
    const bool f = i < n && flags[i] != 0;
    const unsigned m = __ballot_sync(0xffffffffu, f);
    const int inwarp = __popc(m & ((1u << lane) - 1u));
    if (lane == 0) wsum[w] = __popc(m);
    __syncthreads();
    int before = 0;
    for (int k = 0; k < w; ++k) before += wsum[k];
    const int base = base_s;
    if (f) pos[i] = base + before + inwarp;
    __syncthreads();
    if (threadIdx.x == CL_THREADS - 1) base_s = base + before + __popc(m);
    __syncthreads();
  }
  if (threadIdx.x == 0) count[0] = base_s;
}

// removeNans :479-502
__global__ void k_cl_flag_valid(const float4* __restrict__ nrm, int n, unsigned char* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 q = nrm[i];
  flags[i] = (!isnan(q.x) && !isnan(q.y) && !isnan(q.z)) ? 1 : 0;
}
__global__ void k_cl_scatter_valid(const float4* __restrict__ nrm, const unsigned char* __restrict__ flags, const int* __restrict__ pos,
                                   int n, int* __restrict__ keep, float* __restrict__ data3) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !flags[i]) return;
  const int p = pos[i];
  const float4 q = nrm[i];
  keep[p] = i;
  data3[3 * (size_t)p] = q.x;
  data3[3 * (size_t)p + 1] = q.y;
  data3[3 * (size_t)p + 2] = q.z;
}
__global__ void k_cl_flag_eq(const int* __restrict__ labels, int n, int value, unsigned char* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = labels[i] == value ? 1 : 0;
}
// members of one normal cluster :349-362 and their signed distances :377-392:  d = -(x c0 + y c1 + z c2), float, left to right
__global__ void k_cl_scatter_members(const float4* __restrict__ cloud, const int* __restrict__ src_idx, const unsigned char* __restrict__ flags,
                                     const int* __restrict__ pos, int n, float c0, float c1, float c2, int* __restrict__ mem,
                                     float* __restrict__ dist) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !flags[i]) return;
  const int p = pos[i], s = src_idx[i];
  const float4 q = cloud[s];
  float d = __fadd_rn(__fadd_rn(__fmul_rn(q.x, c0), __fmul_rn(q.y, c1)), __fmul_rn(q.z, c2));
  d = __fmul_rn(-1.0f, d);
  mem[p] = s;
  dist[p] = d;
}
// points of one distance cluster :407-416, straight into the RANSAC crop buffer
__global__ void k_cl_scatter_points(const float4* __restrict__ cloud, const int* __restrict__ src_idx, const unsigned char* __restrict__ flags,
                                    const int* __restrict__ pos, int n, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !flags[i]) return;
  out[pos[i]] = cloud[src_idx[i]];
}

// bounding box of the samples (cv::kmeans' `box`): min / max are order-free
__global__ void __launch_bounds__(CL_THREADS) k_cl_minmax(const float* __restrict__ data, int n, int dims, float* __restrict__ lohi) {
  __shared__ float slo[32][CL_MAXD], shi[32][CL_MAXD];
  float lo[CL_MAXD], hi[CL_MAXD];
  for (int j = 0; j < CL_MAXD; ++j) {
    lo[j] = FLT_MAX;
    hi[j] = -FLT_MAX;
  }
  for (int i = threadIdx.x; i < n; i += CL_THREADS)
    for (int j = 0; j < dims; ++j) {
      const float v = data[(size_t)i * dims + j];
      lo[j] = fminf(lo[j], v);
      hi[j] = fmaxf(hi[j], v);
    }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int j = 0; j < dims; ++j) {
    for (int o = 16; o > 0; o >>= 1) {
      lo[j] = fminf(lo[j], __shfl_xor_sync(0xffffffffu, lo[j], o));
      hi[j] = fmaxf(hi[j], __shfl_xor_sync(0xffffffffu, hi[j], o));
    }
    if (lane == 0) {
      slo[w][j] = lo[j];
      shi[w][j] = hi[j];
    }
  }
  __syncthreads();
  if (threadIdx.x < dims) {
    float a = FLT_MAX, b = -FLT_MAX;
    for (int k = 0; k < 32; ++k) {
      a = fminf(a, slo[k][threadIdx.x]);
      b = fmaxf(b, shi[k][threadIdx.x]);
    }
    lohi[threadIdx.x] = a;
    lohi[CL_MAXD + threadIdx.x] = b;
  }
}

// hal::normL2Sqr_ below the SIMD width: one float accumulator, squared differences added in index order
__device__ __forceinline__ float norm_l2_sqr(const float* a, const float* b, int dims) {
  float s = 0.f;
  int j = 0;
  for (; j <= dims - 4; j += 4) {
    const float t0 = __fsub_rn(a[j], b[j]), t1 = __fsub_rn(a[j + 1], b[j + 1]), t2 = __fsub_rn(a[j + 2], b[j + 2]), t3 = __fsub_rn(a[j + 3], b[j + 3]);
    s = __fadd_rn(s, __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t0, t0), __fmul_rn(t1, t1)), __fmul_rn(t2, t2)), __fmul_rn(t3, t3)));
  }
  for (; j < dims; ++j) {
    const float t = __fsub_rn(a[j], b[j]);
    s = __fadd_rn(s, __fmul_rn(t, t));
  }
  return s;
}

struct KmArgs {
  const float* data;        // [N][dims]
  int N, dims, K, max_count;
  double eps2;
  const float* init;        // [attempts][K][dims]  random centres (generateRandomCenter, drawn on the host from cv::RNG)
  int* labels;              // [attempts][N]
  float* centers;           // [attempts][K][dims]
  double* compactness;      // [attempts]
};

// One attempt of cv::kmeans per CTA (blockIdx.x = attempt).  DIMS = A.dims as a compile-time constant: the per-sample loops
// carry no branch on the dimension (a branch per sample cost more than the dependent add it guards).
template <int DIMS>
__global__ void __launch_bounds__(CL_THREADS, 1) k_cl_kmeans(KmArgs A) {
  __shared__ float cen[CL_MAXK * CL_MAXD], old[CL_MAXK * CL_MAXD];
  __shared__ int counters[CL_MAXK];
  __shared__ double red_d[32];
  __shared__ int red_i[32];
  __shared__ double shift_s;
  __shared__ double sh[33];
  extern __shared__ __align__(16) float tile_v[];        // [CL_TILE][dims] samples of the current tile ...
  int* tile_l = reinterpret_cast<int*>(tile_v + (size_t)(CL_TILE + 8) * CL_MAXD);   // ... and their labels
  constexpr int dims = DIMS;
  const int N = A.N, K = A.K;
  const float* data = A.data;
  int* labels = A.labels + (size_t)blockIdx.x * N;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int max_count = max(A.max_count, 2);
  double compact = 0.0;
  for (int iter = 0;;) {
    double max_center_shift = iter == 0 ? DBL_MAX : 0.0;
    if (threadIdx.x < K * dims) old[threadIdx.x] = cen[threadIdx.x];   // swap(centers, old_centers)
    __syncthreads();
    if (iter == 0) {
      if (threadIdx.x < K * dims) cen[threadIdx.x] = A.init[(size_t)blockIdx.x * K * dims + threadIdx.x];
      __syncthreads();
    } else {
      // centre sums in point order.  The whole CTA stages a tile of samples and labels into shared memory (vector loads, all
      // 32 warps: the global-memory latency is paid once per tile), then warp k walks the tile sample by sample, each value
      // masked to +0.0f when the sample is not a member of cluster k.  Adding +0.0f is exact here — a sum that starts at +0.0f can never become -0.0f, the only
      // value +0.0f would change — so the chain costs one dependent FADD per SAMPLE (4 cycles) instead of a ballot / find /
      // shuffle round trip per MEMBER (~70 cycles, profiles/README.md), and the K chains are balanced whatever the cluster
      // sizes.  Every lane of the warp carries the same sums.
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      int cnt = 0;
      for (int t0 = 0; t0 < N; t0 += CL_TILE) {
        const int nt = min(CL_TILE, N - t0);
        {
          const float* src = data + (size_t)t0 * dims;
          const int nf = nt * dims, nf4 = ((reinterpret_cast<size_t>(src) & 15) == 0) ? nf >> 2 : 0;
          const float4* src4 = reinterpret_cast<const float4*>(src);
          float4* dst4 = reinterpret_cast<float4*>(tile_v);
          for (int j = threadIdx.x; j < nf4; j += CL_THREADS) dst4[j] = src4[j];
          for (int j = 4 * nf4 + threadIdx.x; j < nf; j += CL_THREADS) tile_v[j] = src[j];
          const int* lsrc = labels + t0;
          const int nl4 = ((reinterpret_cast<size_t>(lsrc) & 15) == 0) ? nt >> 2 : 0;
          for (int j = threadIdx.x; j < nl4; j += CL_THREADS) reinterpret_cast<int4*>(tile_l)[j] = reinterpret_cast<const int4*>(lsrc)[j];
          for (int j = 4 * nl4 + threadIdx.x; j < nt; j += CL_THREADS) tile_l[j] = lsrc[j];
        }
        __syncthreads();
        if (w < K) {
          // (every lane reads the same shared-memory words — broadcasts — so the loads of the next samples are in flight
          // while the dependent adds of the current ones retire; a shuffle per sample would serialise on its latency)
          for (int j0 = 0; j0 < nt; j0 += 8) {
            int lb[8];
            float x[8][4];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int j = j0 + u;   // (the tail of the last group reads stale words: masked below)
              lb[u] = tile_l[j];
              x[u][0] = tile_v[j * dims];
              x[u][1] = dims > 1 ? tile_v[j * dims + 1] : 0.f;
              x[u][2] = dims > 2 ? tile_v[j * dims + 2] : 0.f;
              x[u][3] = dims > 3 ? tile_v[j * dims + 3] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const bool mine = j0 + u < nt && lb[u] == w;
              cnt += mine ? 1 : 0;
              s0 = __fadd_rn(s0, mine ? x[u][0] : 0.f);
              if (dims > 1) s1 = __fadd_rn(s1, mine ? x[u][1] : 0.f);
              if (dims > 2) s2 = __fadd_rn(s2, mine ? x[u][2] : 0.f);
              if (dims > 3) s3 = __fadd_rn(s3, mine ? x[u][3] : 0.f);
            }
          }
        }
        __syncthreads();   // the tile is overwritten next
      }
      if (w < K && lane == 0) {
        cen[w * dims] = s0;
        if (dims > 1) cen[w * dims + 1] = s1;
        if (dims > 2) cen[w * dims + 2] = s2;
        if (dims > 3) cen[w * dims + 3] = s3;
        counters[w] = cnt;
      }
      __syncthreads();
      // empty clusters: the farthest point of the biggest cluster becomes a one-point cluster
      for (int k = 0; k < K; ++k) {
        if (counters[k] != 0) continue;   // uniform: counters live in shared memory
        int max_k = 0;
        for (int k1 = 1; k1 < K; ++k1)
          if (counters[max_k] < counters[k1]) max_k = k1;
        float bc[CL_MAXD];
        const float scale = __fdiv_rn(1.f, (float)counters[max_k]);
        for (int j = 0; j < dims; ++j) bc[j] = __fmul_rn(cen[max_k * dims + j], scale);
        double best = -1.0;   // `max_dist <= dist` from max_dist = 0: every member qualifies, the LAST farthest one wins
        int best_i = -1;
        for (int i = threadIdx.x; i < N; i += CL_THREADS) {
          if (labels[i] != max_k) continue;
          const double d = (double)norm_l2_sqr(data + (size_t)i * dims, bc, dims);
          if (d >= best) {
            best = d;
            best_i = i;
          }
        }
        for (int o = 16; o > 0; o >>= 1) {
          const double od = __shfl_xor_sync(0xffffffffu, best, o);
          const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
          if (od > best || (od == best && oi > best_i)) {
            best = od;
            best_i = oi;
          }
        }
        if (lane == 0) {
          red_d[w] = best;
          red_i[w] = best_i;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
          for (int q = 1; q < 32; ++q)
            if (red_d[q] > best || (red_d[q] == best && red_i[q] > best_i)) {
              best = red_d[q];
              best_i = red_i[q];
            }
          counters[max_k]--;
          counters[k]++;
          labels[best_i] = k;
          const float* sp = data + (size_t)best_i * dims;
          for (int j = 0; j < dims; ++j) {
            cen[max_k * dims + j] = __fsub_rn(cen[max_k * dims + j], sp[j]);
            cen[k * dims + j] = __fadd_rn(cen[k * dims + j], sp[j]);
          }
        }
        __syncthreads();
      }
      if (threadIdx.x == 0) {
        double shift = 0.0;
        for (int k = 0; k < K; ++k) {
          const float scale = __fdiv_rn(1.f, (float)counters[k]);
          double dist = 0.0;
          for (int j = 0; j < dims; ++j) {
            cen[k * dims + j] = __fmul_rn(cen[k * dims + j], scale);
            const double t = (double)__fsub_rn(cen[k * dims + j], old[k * dims + j]);
            dist += t * t;
          }
          shift = fmax(shift, dist);
        }
        shift_s = shift;
      }
      __syncthreads();
      if (iter > 0) max_center_shift = shift_s;
    }
    const bool last = (++iter == max_count || max_center_shift <= A.eps2);
    float c[CL_MAXK * CL_MAXD];
#pragma unroll
    for (int q = 0; q < CL_MAXK * CL_MAXD; ++q) c[q] = q < K * dims ? cen[q] : 0.f;
    if (last) {
      // labels are kept; compactness = sum of the squared distances to the own centre
      double part = 0.0;
      for (int i = threadIdx.x; i < N; i += CL_THREADS) part += (double)norm_l2_sqr(data + (size_t)i * dims, c + labels[i] * dims, dims);
      compact = ssb::block_sum(part, sh);
      break;
    }
    for (int i = threadIdx.x; i < N; i += CL_THREADS) {
      const float* sp = data + (size_t)i * dims;
      int k_best = 0;
      double min_dist = DBL_MAX;
      for (int k = 0; k < K; ++k) {
        const double d = (double)norm_l2_sqr(sp, c + k * dims, dims);
        if (min_dist > d) {
          min_dist = d;
          k_best = k;
        }
      }
      labels[i] = k_best;
    }
    __syncthreads();
  }
  if (threadIdx.x < K * dims) A.centers[(size_t)blockIdx.x * K * dims + threadIdx.x] = cen[threadIdx.x];
  if (threadIdx.x == 0) A.compactness[blockIdx.x] = compact;
}

// pcl::SampleConsensusModelPlane::projectPoints as called by pcl::ProjectInliers (copy_data_fields = false):
// mc = (a, b, c, 0) normalised (host, float), d4 = the model's 4th coefficient;  distance = (mc0 x + mc1 y) + (mc2 z + d4 * 1),
// pp = p - mc * distance.  flags = the RANSAC inlier mask (refined model, k_finish).
__global__ void k_cl_project(const float4* __restrict__ pts, const unsigned char* __restrict__ mask, const int* __restrict__ pos, int n,
                             float mc0, float mc1, float mc2, float d4, float4* __restrict__ out, int* __restrict__ src) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !mask[i]) return;
  const float4 p = pts[i];
  const float d = __fadd_rn(__fadd_rn(__fmul_rn(mc0, p.x), __fmul_rn(mc1, p.y)), __fadd_rn(__fmul_rn(mc2, p.z), __fmul_rn(d4, 1.0f)));
  float4 o;
  o.x = __fsub_rn(p.x, __fmul_rn(mc0, d));
  o.y = __fsub_rn(p.y, __fmul_rn(mc1, d));
  o.z = __fsub_rn(p.z, __fmul_rn(mc2, d));
  o.w = 0.f;
  out[pos[i]] = o;
  src[pos[i]] = i;
}

// Convex hull, device part.  The 8 points extreme along +-u, +-v, +-(u+v), +-(u-v) are data points, so anything strictly inside
// their octagon cannot be a hull vertex.  k_cl_extremes finds them (ties: lowest index), k_cl_flag_outside keeps what is not
// strictly inside (orientation in double, with a margin far above its rounding error).
__device__ __forceinline__ double coord(const float4& p, int ax) { return ax == 0 ? (double)p.x : ax == 1 ? (double)p.y : (double)p.z; }
__global__ void __launch_bounds__(CL_THREADS) k_cl_extremes(const float4* __restrict__ pts, int m, int iu, int iv, int* __restrict__ ext8) {
  __shared__ double sv[8][32];
  __shared__ int si[8][32];
  double best[8];
  int bi[8];
  for (int q = 0; q < 8; ++q) {
    best[q] = -DBL_MAX;
    bi[q] = -1;
  }
  for (int i = threadIdx.x; i < m; i += CL_THREADS) {
    const float4 p = pts[i];
    const double u = coord(p, iu), v = coord(p, iv);
    const double key[8] = {u, u + v, v, v - u, -u, -(u + v), -v, u - v};   // counter-clockwise: E, NE, N, NW, W, SW, S, SE
    for (int q = 0; q < 8; ++q)
      if (key[q] > best[q]) {
        best[q] = key[q];
        bi[q] = i;
      }
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int q = 0; q < 8; ++q) {
    for (int o = 16; o > 0; o >>= 1) {
      const double od = __shfl_xor_sync(0xffffffffu, best[q], o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi[q], o);
      if (oi >= 0 && (od > best[q] || (od == best[q] && (bi[q] < 0 || oi < bi[q])))) {
        best[q] = od;
        bi[q] = oi;
      }
    }
    if (lane == 0) {
      sv[q][w] = best[q];
      si[q][w] = bi[q];
    }
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    const int q = threadIdx.x;
    double b = -DBL_MAX;
    int ix = -1;
    for (int k = 0; k < 32; ++k)
      if (si[q][k] >= 0 && (sv[q][k] > b || (sv[q][k] == b && (ix < 0 || si[q][k] < ix)))) {
        b = sv[q][k];
        ix = si[q][k];
      }
    ext8[q] = ix;
  }
}
__global__ void k_cl_flag_outside(const float4* __restrict__ pts, int m, int iu, int iv, const int* __restrict__ ext8,
                                  unsigned char* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const float4 p = pts[i];
  const double pu = coord(p, iu), pv = coord(p, iv);
  bool inside = true;
  int edges = 0;
  for (int q = 0; q < 8 && inside; ++q) {
    const int a = ext8[q], b = ext8[(q + 1) & 7];
    if (a < 0 || b < 0 || a == b) continue;
    const float4 A = pts[a], B = pts[b];
    const double au = coord(A, iu), av = coord(A, iv), bu = coord(B, iu), bv = coord(B, iv);
    if (au == bu && av == bv) continue;
    ++edges;
    const double cr = (bu - au) * (pv - av) - (bv - av) * (pu - au);
    if (!(cr > 1e-12)) inside = false;
  }
  flags[i] = (inside && edges >= 3) ? 0 : 1;
}
__global__ void k_cl_scatter_cand(const float4* __restrict__ pts, const unsigned char* __restrict__ flags, const int* __restrict__ pos, int m,
                                  float4* __restrict__ out, int* __restrict__ out_idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m || !flags[i]) return;
  out[pos[i]] = pts[i];
  out_idx[pos[i]] = i;
}

}  // namespace ssb_cl
