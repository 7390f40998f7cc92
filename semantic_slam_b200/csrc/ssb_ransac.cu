// ssb_ransac.cu — planar_segmentation RANSAC plane fit on bbox-cropped depth clouds, sm_100a.
//   K6 k_crop    : plane_segmentation::segmentPointCloudData (plane_segmentation.cpp:24-82)
//   K7 k_hyp     : SampleConsensusModelPlane::computeModelCoefficients (3-point plane)
//      k_count   : countWithinDistance for every (crop, hypothesis): float4 point tiles staged in
//                  shared memory by 1-D TMA bulk copies, hypotheses held in registers (one lane =
//                  HPT hypotheses), integer counts folded with one atomicAdd per (block, hypothesis)
//   K8 k_finish  : RandomSampleConsensus::computeModel winner selection (first best / adaptive-k
//                  replay), optimizeModelCoefficients (PCA of the inliers, fp64) and
//                  selectWithinDistance (plane_segmentation.cpp:639-647 -> pcl::SACSegmentation)
// Float evaluation order is Eigen's SSE3 4-vector dot: ((a*x + b*y) + (c*z + d)) with separate
// IEEE roundings (no FMA contraction: __fmul_rn/__fadd_rn), so inlier counts are integer-exact
// against the CPU oracle.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <random>
#include <utility>
#include <vector>

#include "../../include/ssb.h"
#include "ssb_common.cuh"
#include "ssb_organized.cuh"
#include "ssb_cluster.cuh"

namespace ssb {

template <class T>
struct RBuf {
  T* p = nullptr;
  size_t cap = 0;
  int ensure(size_t n) {
    if (n <= cap && p) return SSB_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = std::max<size_t>(n, 1);
    cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
    if (e != cudaSuccess) {
      set_error("cudaMalloc(%zu bytes) failed: %s", want * sizeof(T), cudaGetErrorString(e));
      return SSB_ERR_CUDA;
    }
    cap = want;
    return SSB_OK;
  }
  ~RBuf() {
    if (p) cudaFree(p);
  }
};

struct BoxInfo {      // per bbox, device side
  int tl_x, tl_y, w, h;
  int n;              // w*h, or -1 if spurious
  int pt_off;         // offset (in points) of the crop in the concatenated crop buffer
  int tile_off;       // first tile index of this box
  int pad;
};

__device__ __forceinline__ float plane_dist(float a, float b, float c, float d, float x, float y, float z) {
  const float s0 = __fadd_rn(__fmul_rn(a, x), __fmul_rn(b, y));
  const float s1 = __fadd_rn(__fmul_rn(c, z), d);
  return fabsf(__fadd_rn(s0, s1));
}

// ---- K6 -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_crop(const unsigned char* __restrict__ msg, ssb_cloud_layout L,
                                              const BoxInfo* __restrict__ boxes, float4* __restrict__ crop) {
  const BoxInfo B = boxes[blockIdx.y];
  if (B.n <= 0) return;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < B.n; idx += gridDim.x * blockDim.x) {
    const int pv = idx / B.w, pu = idx - pv * B.w;
    const size_t pos = (size_t)(B.tl_y + pv) * L.row_step + (size_t)(B.tl_x + pu) * L.point_step;
    float4 o;
    // 4-byte fields; PointCloud2 offsets are 4-byte aligned in every driver the reference consumes
    o.x = *reinterpret_cast<const float*>(msg + pos + L.off_x);
    o.y = *reinterpret_cast<const float*>(msg + pos + L.off_y);
    o.z = *reinterpret_cast<const float*>(msg + pos + L.off_z);
    o.w = *reinterpret_cast<const float*>(msg + pos + L.off_rgb);
    crop[(size_t)B.pt_off + idx] = o;
  }
}

// ---- K7a ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_hyp(const float4* __restrict__ crop, const BoxInfo* __restrict__ boxes,
                                             const int* __restrict__ triples, int K, float4* __restrict__ hyp,
                                             int* __restrict__ valid) {
  const int b = blockIdx.y;
  const BoxInfo B = boxes[b];
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const size_t o = (size_t)b * K + k;
  const float qnan = __int_as_float(0x7fc00000);
  if (B.n <= 0) {
    hyp[o] = make_float4(qnan, qnan, qnan, qnan);
    valid[o] = 0;
    return;
  }
  const int* t = triples + 3 * o;
  int i0 = t[0], i1 = t[1], i2 = t[2];
  if ((unsigned)i0 >= (unsigned)B.n || (unsigned)i1 >= (unsigned)B.n || (unsigned)i2 >= (unsigned)B.n) {
    hyp[o] = make_float4(qnan, qnan, qnan, qnan);
    valid[o] = 0;
    return;
  }
  const float4 p0 = crop[(size_t)B.pt_off + i0], p1 = crop[(size_t)B.pt_off + i1], p2 = crop[(size_t)B.pt_off + i2];
  const float ax = __fsub_rn(p1.x, p0.x), ay = __fsub_rn(p1.y, p0.y), az = __fsub_rn(p1.z, p0.z);
  const float bx = __fsub_rn(p2.x, p0.x), by = __fsub_rn(p2.y, p0.y), bz = __fsub_rn(p2.z, p0.z);
  const float rx = __fdiv_rn(ax, bx), ry = __fdiv_rn(ay, by), rz = __fdiv_rn(az, bz);
  if ((rx == ry) && (rz == ry)) {  // collinear sample
    hyp[o] = make_float4(qnan, qnan, qnan, qnan);
    valid[o] = 0;
    return;
  }
  float c0 = __fsub_rn(__fmul_rn(ay, bz), __fmul_rn(az, by));
  float c1 = __fsub_rn(__fmul_rn(az, bx), __fmul_rn(ax, bz));
  float c2 = __fsub_rn(__fmul_rn(ax, by), __fmul_rn(ay, bx));
  float c3 = 0.f;
  const float n2 = __fadd_rn(__fadd_rn(__fmul_rn(c0, c0), __fmul_rn(c1, c1)), __fadd_rn(__fmul_rn(c2, c2), __fmul_rn(c3, c3)));
  const float n = __fsqrt_rn(n2);
  c0 = __fdiv_rn(c0, n);
  c1 = __fdiv_rn(c1, n);
  c2 = __fdiv_rn(c2, n);
  c3 = __fdiv_rn(c3, n);
  const float dot = __fadd_rn(__fadd_rn(__fmul_rn(c0, p0.x), __fmul_rn(c1, p0.y)), __fadd_rn(__fmul_rn(c2, p0.z), __fmul_rn(c3, 1.0f)));
  hyp[o] = make_float4(c0, c1, c2, __fmul_rn(-1.f, dot));
  valid[o] = 1;
}

// ---- K7b: the point x hypothesis sweep --------------------------------------------------------
// Packed fp32x2 arithmetic (Blackwell FMUL2 / FADD2): one instruction evaluates two hypotheses against
// the broadcast point with the same per-lane IEEE roundings as the scalar form, so counts stay bit-exact.
constexpr int CNT_THREADS = 128;
constexpr int CNT_HPT = 8;                       // hypotheses per thread (4 packed pairs)
constexpr int CNT_HYP_PER_BLOCK = CNT_THREADS * CNT_HPT;
constexpr int CNT_TILE = 512;                    // points per tile (8 KB of float4)

struct TileRef {
  int box;
  int first;  // first point of the tile inside the crop
};

__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
// NOTE: ptxas 12.9 contracts `mul.rn.f32x2` + `add.rn.f32x2` into FFMA2 (unlike the scalar .rn forms),
// which changes the rounding and breaks integer-exact counts.  The `.ftz` qualifier on the multiply only
// makes the pair non-contractable (verified in SASS: FMUL2.FTZ + FADD2).  Flushing a subnormal *product*
// cannot change |n.p + d| < thr: a term below 1.2e-38 never moves a sum that is compared with ~1e-2.
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

__global__ void __launch_bounds__(CNT_THREADS, 8)
    k_count(const float4* __restrict__ crop, const BoxInfo* __restrict__ boxes, const TileRef* __restrict__ tiles,
            const float4* __restrict__ hyp, int K, float thr, int* __restrict__ counts) {
  __shared__ __align__(128) float4 pts[CNT_TILE];
  __shared__ __align__(8) uint64_t bar;
  const TileRef T = tiles[blockIdx.x];
  const BoxInfo B = boxes[T.box];
  const int n = min(CNT_TILE, B.n - T.first);
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&bar, (uint32_t)n * 16u);
    tma_load_1d(pts, crop + (size_t)B.pt_off + T.first, (uint32_t)n * 16u, &bar);
  }
  // hypotheses of this thread, packed in pairs (block-column blockIdx.y covers CNT_HYP_PER_BLOCK hypotheses)
  uint64_t ha[CNT_HPT / 2], hb[CNT_HPT / 2], hc[CNT_HPT / 2], hd[CNT_HPT / 2];
  int cnt[CNT_HPT];
  const int hbase = blockIdx.y * CNT_HYP_PER_BLOCK;
  const float qnan = __int_as_float(0x7fc00000);
#pragma unroll
  for (int j = 0; j < CNT_HPT / 2; ++j) {
    const int k0 = hbase + (2 * j) * CNT_THREADS + tid, k1 = hbase + (2 * j + 1) * CNT_THREADS + tid;
    float4 h0 = make_float4(qnan, 0.f, 0.f, 0.f), h1 = h0;
    if (k0 < K) h0 = hyp[(size_t)T.box * K + k0];
    if (k1 < K) h1 = hyp[(size_t)T.box * K + k1];
    ha[j] = pack2(h0.x, h1.x);
    hb[j] = pack2(h0.y, h1.y);
    hc[j] = pack2(h0.z, h1.z);
    hd[j] = pack2(h0.w, h1.w);
    cnt[2 * j] = 0;
    cnt[2 * j + 1] = 0;
  }
  mbar_wait(&bar, 0);
#pragma unroll 2
  for (int i = 0; i < n; ++i) {
    const float4 P = pts[i];  // warp-wide broadcast
    const uint64_t x2 = pack2(P.x, P.x), y2 = pack2(P.y, P.y), z2 = pack2(P.z, P.z);
#pragma unroll
    for (int j = 0; j < CNT_HPT / 2; ++j) {
      // ((a*x + b*y) + (c*z + d)) per lane, separate roundings
      const uint64_t s = add2(add2(mul2(ha[j], x2), mul2(hb[j], y2)), add2(mul2(hc[j], z2), hd[j]));
      float s0, s1;
      unpack2(s, s0, s1);
      cnt[2 * j] += (fabsf(s0) < thr) ? 1 : 0;
      cnt[2 * j + 1] += (fabsf(s1) < thr) ? 1 : 0;
    }
  }
#pragma unroll
  for (int j = 0; j < CNT_HPT; ++j) {
    const int k = hbase + j * CNT_THREADS + tid;
    if (k < K && cnt[j]) atomicAdd(counts + (size_t)T.box * K + k, cnt[j]);
  }
}

// ---- K8 --------------------------------------------------------------------------------------
__device__ void compute_roots2(double b, double c, double* roots) {
  roots[0] = 0;
  double d = b * b - 4.0 * c;
  if (d < 0.0) d = 0.0;
  double sd = sqrt(d);
  roots[2] = 0.5 * (b + sd);
  roots[1] = 0.5 * (b - sd);
}
__device__ void compute_roots(const double* m, double* roots) {
  double c0 = m[0] * m[4] * m[8] + 2.0 * m[1] * m[2] * m[5] - m[0] * m[5] * m[5] - m[4] * m[2] * m[2] - m[8] * m[1] * m[1];
  double c1 = m[0] * m[4] - m[1] * m[1] + m[0] * m[8] - m[2] * m[2] + m[4] * m[8] - m[5] * m[5];
  double c2 = m[0] + m[4] + m[8];
  if (fabs(c0) < 2.220446049250313e-16) {
    compute_roots2(c2, c1, roots);
    return;
  }
  const double s_inv3 = 1.0 / 3.0;
  const double s_sqrt3 = sqrt(3.0);
  double c2_over_3 = c2 * s_inv3;
  double a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
  if (a_over_3 > 0.0) a_over_3 = 0.0;
  double half_b = 0.5 * (c0 + c2_over_3 * (2.0 * c2_over_3 * c2_over_3 - c1));
  double q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
  if (q > 0.0) q = 0.0;
  double rho = sqrt(-a_over_3);
  double theta = atan2(sqrt(-q), half_b) * s_inv3;
  double cos_theta = cos(theta), sin_theta = sin(theta);
  roots[0] = c2_over_3 + 2.0 * rho * cos_theta;
  roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
  roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
  double t;
  if (roots[0] >= roots[1]) {
    t = roots[0];
    roots[0] = roots[1];
    roots[1] = t;
  }
  if (roots[1] >= roots[2]) {
    t = roots[1];
    roots[1] = roots[2];
    roots[2] = t;
    if (roots[0] >= roots[1]) {
      t = roots[0];
      roots[0] = roots[1];
      roots[1] = t;
    }
  }
  if (roots[0] <= 0) compute_roots2(c2, c1, roots);
}
__device__ void eigen33_smallest(const double* mat, double* evec) {
  double scale = 0;
  for (int i = 0; i < 9; ++i) scale = fmax(scale, fabs(mat[i]));
  if (scale <= 2.2250738585072014e-308) scale = 1.0;
  double m[9];
  for (int i = 0; i < 9; ++i) m[i] = mat[i] / scale;
  double roots[3];
  compute_roots(m, roots);
  m[0] -= roots[0];
  m[4] -= roots[0];
  m[8] -= roots[0];
  double v1[3] = {m[1] * m[5] - m[2] * m[4], m[2] * m[3] - m[0] * m[5], m[0] * m[4] - m[1] * m[3]};
  double v2[3] = {m[1] * m[8] - m[2] * m[7], m[2] * m[6] - m[0] * m[8], m[0] * m[7] - m[1] * m[6]};
  double v3[3] = {m[4] * m[8] - m[5] * m[7], m[5] * m[6] - m[3] * m[8], m[3] * m[7] - m[4] * m[6]};
  double l1 = v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2];
  double l2 = v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2];
  double l3 = v3[0] * v3[0] + v3[1] * v3[1] + v3[2] * v3[2];
  const double* v;
  double l;
  if (l1 >= l2 && l1 >= l3) {
    v = v1;
    l = l1;
  } else if (l2 >= l1 && l2 >= l3) {
    v = v2;
    l = l2;
  } else {
    v = v3;
    l = l3;
  }
  double s = sqrt(l);
  for (int i = 0; i < 3; ++i) evec[i] = v[i] / s;
}

constexpr int FIN_THREADS = 256;

__global__ void __launch_bounds__(FIN_THREADS)
    k_finish(const float4* __restrict__ crop, const BoxInfo* __restrict__ boxes, const float4* __restrict__ hyp,
             const int* __restrict__ valid, int* __restrict__ counts, int K, float thr, int refine, int mode,
             int max_iterations, double probability, ssb_plane_result* __restrict__ results,
             unsigned char* __restrict__ mask, const int* __restrict__ mask_off) {
  __shared__ double shd[33];
  __shared__ int s_best_cnt[FIN_THREADS], s_best_k[FIN_THREADS];
  __shared__ float s_coef[4], s_ref[4], s_cen[3];
  __shared__ int s_iter, s_bestk, s_bestc, s_rc;
  const int b = blockIdx.x, tid = threadIdx.x;
  const BoxInfo B = boxes[b];
  ssb_plane_result R;
  memset(&R, 0, sizeof(R));
  R.best_hyp = -1;
  if (B.n < 0) {
    R.status = 1;
    if (tid == 0) results[b] = R;
    for (int k = tid; k < K; k += FIN_THREADS) counts[(size_t)b * K + k] = -1;
    return;
  }
  R.n_points = B.n;
  if (B.n == 0)  // empty crop: nothing is evaluated
    for (int k = tid; k < K; k += FIN_THREADS) counts[(size_t)b * K + k] = -1;
  // invalid hypotheses count as 0 (fixed-K) — k_count already left them at 0 because their
  // coefficients are NaN.
  if (mode == 0) {
    int bc = 0, bk = -1;
    for (int k = tid; k < K; k += FIN_THREADS) {
      const int c = counts[(size_t)b * K + k];
      if (c > bc) {  // strictly better, ascending k => first best per thread
        bc = c;
        bk = k;
      }
    }
    s_best_cnt[tid] = bc;
    s_best_k[tid] = bk;
    __syncthreads();
    for (int o = FIN_THREADS / 2; o > 0; o >>= 1) {
      if (tid < o) {
        const int c2 = s_best_cnt[tid + o], k2 = s_best_k[tid + o];
        const int c1 = s_best_cnt[tid], k1 = s_best_k[tid];
        if (c2 > c1 || (c2 == c1 && k2 >= 0 && (k1 < 0 || k2 < k1))) {
          s_best_cnt[tid] = c2;
          s_best_k[tid] = k2;
        }
      }
      __syncthreads();
    }
    if (tid == 0) {
      s_bestc = s_best_cnt[0];
      s_bestk = s_best_cnt[0] > 0 ? s_best_k[0] : -1;
      s_iter = B.n > 0 ? K : 0;
    }
  } else if (tid == 0) {
    // RandomSampleConsensus::computeModel replayed over the precomputed counts
    double kk = 1.0;
    const double log_probability = log(1.0 - probability);
    const double one_over = B.n > 0 ? 1.0 / (double)B.n : 0.0;
    int iterations = 0, skipped = 0, s = 0, best = 0, bestk = -1;
    const int max_skip = max_iterations * 10;
    while (iterations < kk && skipped < max_skip && s < K && B.n > 0) {
      const int k = s++;
      if (!valid[(size_t)b * K + k]) {
        ++skipped;
        continue;
      }
      const int c = counts[(size_t)b * K + k];
      if (c > best) {
        best = c;
        bestk = k;
        double w = (double)best * one_over;
        double pno = 1.0 - pow(w, 3.0);
        pno = fmax(2.220446049250313e-16, pno);
        pno = fmin(1.0 - 2.220446049250313e-16, pno);
        kk = log_probability / log(pno);
      }
      ++iterations;
      if (iterations > max_iterations) break;
    }
    s_bestc = best;
    s_bestk = bestk;
    s_iter = iterations;
    // hypotheses the adaptive loop never reached are reported as not evaluated
    for (int k = s; k < K; ++k) counts[(size_t)b * K + k] = -1;
    for (int k = 0; k < s; ++k)
      if (!valid[(size_t)b * K + k]) counts[(size_t)b * K + k] = -1;
  }
  __syncthreads();
  R.best_hyp = s_bestk;
  R.best_count = s_bestc;
  R.iterations = s_iter;
  unsigned char* mk = mask ? mask + mask_off[b] : nullptr;
  if (s_bestk < 0) {
    R.status = 2;
    if (tid == 0) results[b] = R;
    if (mk)
      for (int i = tid; i < B.n; i += FIN_THREADS) mk[i] = 0;
    return;
  }
  if (tid == 0) {
    const float4 h = hyp[(size_t)b * K + s_bestk];
    s_coef[0] = h.x;
    s_coef[1] = h.y;
    s_coef[2] = h.z;
    s_coef[3] = h.w;
  }
  __syncthreads();
  const float a = s_coef[0], bb = s_coef[1], c = s_coef[2], d = s_coef[3];
  const float4* P = crop + (size_t)B.pt_off;
  if (refine) {
    double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    double cnt = 0;
    for (int i = tid; i < B.n; i += FIN_THREADS) {
      const float4 p = P[i];
      if (plane_dist(a, bb, c, d, p.x, p.y, p.z) < thr) {
        const double x = p.x, y = p.y, z = p.z;
        cnt += 1.0;
        acc[0] += x * x;
        acc[1] += x * y;
        acc[2] += x * z;
        acc[3] += y * y;
        acc[4] += y * z;
        acc[5] += z * z;
        acc[6] += x;
        acc[7] += y;
        acc[8] += z;
      }
    }
    cnt = block_sum(cnt, shd);
    for (int k = 0; k < 9; ++k) acc[k] = block_sum(acc[k], shd);
    if (tid == 0) {
      s_cen[0] = s_cen[1] = s_cen[2] = 0.f;
      if (cnt < 4.0) {
        for (int k = 0; k < 4; ++k) s_ref[k] = s_coef[k];
      } else {
        for (int k = 0; k < 9; ++k) acc[k] /= cnt;
        double cov[9];
        cov[0] = acc[0] - acc[6] * acc[6];
        cov[1] = acc[1] - acc[6] * acc[7];
        cov[2] = acc[2] - acc[6] * acc[8];
        cov[4] = acc[3] - acc[7] * acc[7];
        cov[5] = acc[4] - acc[7] * acc[8];
        cov[8] = acc[5] - acc[8] * acc[8];
        cov[3] = cov[1];
        cov[6] = cov[2];
        cov[7] = cov[5];
        double ev[3];
        eigen33_smallest(cov, ev);
        const float n0 = (float)ev[0], n1 = (float)ev[1], n2 = (float)ev[2];
        const float cx = (float)acc[6], cy = (float)acc[7], cz = (float)acc[8];
        const float dot = __fadd_rn(__fadd_rn(__fmul_rn(n0, cx), __fmul_rn(n1, cy)), __fadd_rn(__fmul_rn(n2, cz), __fmul_rn(0.f, 1.0f)));
        s_ref[0] = n0;
        s_ref[1] = n1;
        s_ref[2] = n2;
        s_ref[3] = __fmul_rn(-1.f, dot);
        s_cen[0] = cx;
        s_cen[1] = cy;
        s_cen[2] = cz;
      }
    }
  } else if (tid == 0) {
    for (int k = 0; k < 4; ++k) s_ref[k] = s_coef[k];
    s_cen[0] = s_cen[1] = s_cen[2] = 0.f;
  }
  if (tid == 0) s_rc = 0;
  __syncthreads();
  const float ra = s_ref[0], rb = s_ref[1], rc_ = s_ref[2], rd = s_ref[3];
  int local = 0;
  for (int i = tid; i < B.n; i += FIN_THREADS) {
    const float4 p = P[i];
    const bool in = plane_dist(ra, rb, rc_, rd, p.x, p.y, p.z) < thr;
    local += in;
    if (mk) mk[i] = in ? 1 : 0;
  }
  local = __reduce_add_sync(0xffffffffu, local);
  if ((tid & 31) == 0 && local) atomicAdd(&s_rc, local);
  __syncthreads();
  if (tid == 0) {
    for (int k = 0; k < 4; ++k) {
      R.coef[k] = s_coef[k];
      R.refined[k] = s_ref[k];
    }
    for (int k = 0; k < 3; ++k) R.centroid[k] = s_cen[k];
    R.refined_count = s_rc;
    results[b] = R;
  }
}

static float effective_threshold(double thr) {
  float t = (float)thr;
  if ((double)t < thr) t = std::nextafterf(t, std::numeric_limits<float>::infinity());
  return t;
}

}  // namespace ssb

using namespace ssb;

struct ssb_ransac {
  int device = 0;
  cudaStream_t stream = nullptr;
  RBuf<unsigned char> d_msg, d_mask;
  RBuf<BoxInfo> d_boxes;
  RBuf<TileRef> d_tiles;
  RBuf<int> d_triples, d_valid, d_counts, d_mask_off;
  RBuf<float4> d_crop, d_hyp;
  RBuf<ssb_plane_result> d_results;
  // current problem
  ssb_cloud_layout layout;
  ssb_ransac_opts opts;
  std::vector<BoxInfo> boxes;
  std::vector<int> mask_off;
  int nb = 0, K = 0, n_tiles = 0;
  long long total_pts = 0;
  bool uploaded = false;
  long long launches = 0;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // run begin/end, k_count begin/end
  unsigned char* h_mask_stage = nullptr;   // page-locked landing buffer of the padded inlier mask (ssb_ransac_fetch)
  size_t h_mask_cap = 0;
  // organised multi-plane segmentation (ssb_organized.cuh)
  RBuf<ssb_org::OrgBox> d_oboxes;
  RBuf<unsigned char> d_change;
  RBuf<float> d_dist, d_planed;
  RBuf<float4> d_nrm;
  RBuf<double> d_ii;
  RBuf<unsigned> d_ic;
  RBuf<int> d_parent, d_label, d_count, d_nlabels, d_ncand, d_nreg, d_l2m;
  RBuf<ssb_org::OrgRegion> d_oreg;
  RBuf<unsigned long long> d_lastev;
  cudaEvent_t oev[2] = {nullptr, nullptr};
  double org_ms = 0.0;
  // plane-clustering chain (ssb_cluster.cuh)
  RBuf<unsigned char> c_flags, c_maskh;
  RBuf<int> c_pos, c_count, c_keep, c_mem, c_labels, c_dlabels, c_src, c_cidx, c_csrc, c_ext;
  RBuf<float> c_data, c_dist, c_init, c_centers, c_lohi;
  RBuf<double> c_compact;
  RBuf<float4> c_cloud, c_nrm, c_pts, c_proj, c_cand;
  int km_smem_set = 0;   // bit d: k_cl_kmeans<d> has its dynamic shared memory opted in on this device
};

static int plan(ssb_ransac* r, const ssb_cloud_layout* L, const ssb_bbox* bx, int nb) {
  r->boxes.resize(nb);
  r->mask_off.assign(nb + 1, 0);
  std::vector<TileRef> tiles;
  long long off = 0;
  for (int b = 0; b < nb; ++b) {
    BoxInfo& B = r->boxes[b];
    B.tl_x = bx[b].tl_x;
    B.tl_y = bx[b].tl_y;
    B.w = bx[b].width;
    B.h = bx[b].height;
    B.pad = 0;
    // plane_segmentation.cpp:34-35 (+ negative corner rejection)
    const bool spurious = B.h < 0 || B.w < 0 || (B.tl_x + B.w) > L->width || (B.tl_y + B.h) > L->height || B.tl_x < 0 || B.tl_y < 0;
    B.n = spurious ? -1 : B.w * B.h;
    B.pt_off = (int)off;
    B.tile_off = (int)tiles.size();
    r->mask_off[b] = (int)off;
    if (B.n > 0) {
      for (int f = 0; f < B.n; f += CNT_TILE) tiles.push_back({b, f});
      off += B.n;
      off = (off + 3) & ~3LL;  // keep every crop 64 B aligned for the bulk copies
    }
  }
  r->mask_off[nb] = (int)off;
  r->total_pts = off;
  r->n_tiles = (int)tiles.size();
  SSB_CUDA_CHECK(cudaSetDevice(r->device));
  int rc;
  if ((rc = r->d_boxes.ensure(nb)) || (rc = r->d_tiles.ensure(tiles.size())) || (rc = r->d_crop.ensure((size_t)off + CNT_TILE)) ||
      (rc = r->d_mask.ensure((size_t)off + 16)) || (rc = r->d_mask_off.ensure(nb + 1)) || (rc = r->d_results.ensure(nb)))
    return rc;
  SSB_CUDA_CHECK(cudaMemcpyAsync(r->d_boxes.p, r->boxes.data(), nb * sizeof(BoxInfo), cudaMemcpyHostToDevice, r->stream));
  if (!tiles.empty())
    SSB_CUDA_CHECK(cudaMemcpyAsync(r->d_tiles.p, tiles.data(), tiles.size() * sizeof(TileRef), cudaMemcpyHostToDevice, r->stream));
  SSB_CUDA_CHECK(cudaMemcpyAsync(r->d_mask_off.p, r->mask_off.data(), (nb + 1) * sizeof(int), cudaMemcpyHostToDevice, r->stream));
  SSB_CUDA_CHECK(cudaStreamSynchronize(r->stream));  // `tiles` is a local
  return SSB_OK;
}

static int run_device(ssb_ransac* r, bool skip_crop = false) {
  const int nb = r->nb, K = r->K;
  cudaStream_t s = r->stream;
  const float thr = effective_threshold(r->opts.threshold);
  if (nb == 0) return SSB_OK;
  for (int k = 0; k < 4; ++k)
    if (!r->ev[k]) SSB_CUDA_CHECK(cudaEventCreate(&r->ev[k]));
  SSB_CUDA_CHECK(cudaEventRecord(r->ev[0], s));
  SSB_CUDA_CHECK(cudaEventRecord(r->ev[2], s));
  SSB_CUDA_CHECK(cudaEventRecord(r->ev[3], s));
  SSB_CUDA_CHECK(cudaMemsetAsync(r->d_counts.p, 0, (size_t)nb * std::max(K, 1) * sizeof(int), s));
  if (!skip_crop) {
    dim3 grid(8, nb);
    k_crop<<<grid, 256, 0, s>>>(r->d_msg.p, r->layout, r->d_boxes.p, r->d_crop.p);
    r->launches++;
  }
  if (K > 0) {
    dim3 grid((K + 255) / 256, nb);
    k_hyp<<<grid, 256, 0, s>>>(r->d_crop.p, r->d_boxes.p, r->d_triples.p, K, r->d_hyp.p, r->d_valid.p);
    r->launches++;
    if (r->n_tiles > 0) {
      dim3 g2(r->n_tiles, (K + CNT_HYP_PER_BLOCK - 1) / CNT_HYP_PER_BLOCK);
      SSB_CUDA_CHECK(cudaEventRecord(r->ev[2], s));
      k_count<<<g2, CNT_THREADS, 0, s>>>(r->d_crop.p, r->d_boxes.p, r->d_tiles.p, r->d_hyp.p, K, thr, r->d_counts.p);
      SSB_CUDA_CHECK(cudaEventRecord(r->ev[3], s));
      r->launches++;
    }
  }
  k_finish<<<nb, FIN_THREADS, 0, s>>>(r->d_crop.p, r->d_boxes.p, r->d_hyp.p, r->d_valid.p, r->d_counts.p, K, thr, r->opts.refine,
                                      r->opts.mode, r->opts.max_iterations, r->opts.probability, r->d_results.p, r->d_mask.p,
                                      r->d_mask_off.p);
  r->launches++;
  SSB_CUDA_CHECK(cudaEventRecord(r->ev[1], s));
  SSB_CUDA_CHECK(cudaGetLastError());
  return SSB_OK;
}

extern "C" {

// CUDA-event timing of the last run: out[0] = whole device pipeline (crop..finish) ms, out[1] = k_count ms
int ssb_ransac_timing(ssb_ransac* r, double out[2]) {
  if (!r || !out || !r->ev[0]) return SSB_ERR_INVALID;
  SSB_CUDA_CHECK(cudaSetDevice(r->device));
  SSB_CUDA_CHECK(cudaStreamSynchronize(r->stream));
  float a = 0, b = 0;
  SSB_CUDA_CHECK(cudaEventElapsedTime(&a, r->ev[0], r->ev[1]));
  SSB_CUDA_CHECK(cudaEventElapsedTime(&b, r->ev[2], r->ev[3]));
  out[0] = a;
  out[1] = b;
  return SSB_OK;
}


void ssb_ransac_default_opts(ssb_ransac_opts* o) {
  if (!o) return;
  std::memset(o, 0, sizeof(*o));
  o->threshold = 0.01;   // plane_segmentation.cpp:645
  o->refine = 1;         // :641
  o->mode = 0;
  o->max_iterations = 50;
  o->probability = 0.99;
  o->device = -1;
}

ssb_ransac* ssb_ransac_create(int device) {
  ssb_ransac* r = new ssb_ransac();
  cudaError_t e;
  if (device < 0) {
    e = cudaGetDevice(&device);
    if (e != cudaSuccess) {
      set_error("no CUDA device: %s (the CUDA back-end has no CPU fallback)", cudaGetErrorString(e));
      delete r;
      return nullptr;
    }
  }
  e = cudaSetDevice(device);
  if (e != cudaSuccess || cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking) != cudaSuccess) {
    set_error("cudaSetDevice(%d)/stream creation failed: %s (the CUDA back-end has no CPU fallback)", device,
              cudaGetErrorString(cudaGetLastError()));
    delete r;
    return nullptr;
  }
  r->device = device;
  ssb_ransac_default_opts(&r->opts);
  return r;
}

void ssb_ransac_destroy(ssb_ransac* r) {
  if (!r) return;
  cudaSetDevice(r->device);
  if (r->stream) {
    cudaStreamSynchronize(r->stream);
    cudaStreamDestroy(r->stream);
  }
  for (int k = 0; k < 4; ++k)
    if (r->ev[k]) cudaEventDestroy(r->ev[k]);
  if (r->h_mask_stage) cudaFreeHost(r->h_mask_stage);
  delete r;
}

void* ssb_ransac_stream(ssb_ransac* r) { return r ? (void*)r->stream : nullptr; }
long long ssb_ransac_launch_count(ssb_ransac* r) { return r ? r->launches : 0; }

int ssb_ransac_upload(ssb_ransac* r, const void* msg, const ssb_cloud_layout* layout, const ssb_bbox* boxes, int n_boxes,
                      const int* triples, int n_hyp, const ssb_ransac_opts* opts) {
  if (!r || !msg || !layout || (n_boxes > 0 && !boxes) || n_boxes < 0 || n_hyp < 0 || (n_hyp > 0 && n_boxes > 0 && !triples)) {
    set_error("ssb_ransac_upload: invalid argument");
    return SSB_ERR_INVALID;
  }
  if (layout->width <= 0 || layout->height <= 0 || layout->point_step < 16 || layout->row_step < layout->width * layout->point_step ||
      (layout->off_x | layout->off_y | layout->off_z | layout->off_rgb) & 3 || (layout->point_step & 3) || (layout->row_step & 3)) {
    set_error("ssb_ransac_upload: unsupported PointCloud2 layout");
    return SSB_ERR_INVALID;
  }
  SSB_CUDA_CHECK(cudaSetDevice(r->device));
  r->layout = *layout;
  if (opts)
    r->opts = *opts;
  else
    ssb_ransac_default_opts(&r->opts);
  r->nb = n_boxes;
  r->K = n_hyp;
  int rc = plan(r, layout, boxes, n_boxes);
  if (rc) return rc;
  const size_t msg_bytes = (size_t)layout->height * layout->row_step;
  const size_t nh = (size_t)n_boxes * std::max(n_hyp, 1);
  if ((rc = r->d_msg.ensure(msg_bytes + 16)) || (rc = r->d_triples.ensure(3 * nh)) || (rc = r->d_valid.ensure(nh)) ||
      (rc = r->d_counts.ensure(nh)) || (rc = r->d_hyp.ensure(nh)))
    return rc;
  SSB_CUDA_CHECK(cudaMemcpyAsync(r->d_msg.p, msg, msg_bytes, cudaMemcpyHostToDevice, r->stream));
  if (n_hyp > 0 && n_boxes > 0)
    SSB_CUDA_CHECK(cudaMemcpyAsync(r->d_triples.p, triples, 3 * (size_t)n_boxes * n_hyp * sizeof(int), cudaMemcpyHostToDevice, r->stream));
  r->uploaded = true;
  return SSB_OK;
}

int ssb_ransac_run_resident(ssb_ransac* r) {
  if (!r || !r->uploaded) {
    set_error("ssb_ransac_run_resident: nothing uploaded");
    return SSB_ERR_INVALID;
  }
  SSB_CUDA_CHECK(cudaSetDevice(r->device));
  return run_device(r);
}

int ssb_ransac_fetch(ssb_ransac* r, ssb_plane_result* results, int* counts, unsigned char* mask) {
  if (!r || !r->uploaded) return SSB_ERR_INVALID;
  SSB_CUDA_CHECK(cudaSetDevice(r->device));
  const int nb = r->nb, K = r->K;
  if (results && nb) SSB_CUDA_CHECK(cudaMemcpyAsync(results, r->d_results.p, nb * sizeof(ssb_plane_result), cudaMemcpyDeviceToHost, r->stream));
  if (counts && nb && K) SSB_CUDA_CHECK(cudaMemcpyAsync(counts, r->d_counts.p, (size_t)nb * K * sizeof(int), cudaMemcpyDeviceToHost, r->stream));
  if (mask && r->total_pts) {
    // the padded mask lands in a page-locked buffer owned by the handle (a pageable landing buffer made the copy a
    // staged, synchronous one and cost a fresh 0.9 MB allocation per frame)
    if ((size_t)r->total_pts > r->h_mask_cap) {
      if (r->h_mask_stage) cudaFreeHost(r->h_mask_stage);
      r->h_mask_stage = nullptr;
      r->h_mask_cap = 0;
      const size_t cap = (size_t)r->total_pts + (size_t)r->total_pts / 4 + 4096;
      SSB_CUDA_CHECK(cudaMallocHost((void**)&r->h_mask_stage, cap));
      r->h_mask_cap = cap;
    }
    SSB_CUDA_CHECK(cudaMemcpyAsync(r->h_mask_stage, r->d_mask.p, r->total_pts, cudaMemcpyDeviceToHost, r->stream));
  }
  SSB_CUDA_CHECK(cudaStreamSynchronize(r->stream));
  if (mask && r->total_pts) {
    // compact the 4-point alignment padding away: caller layout is the plain concatenation
    size_t o = 0;
    for (int b = 0; b < nb; ++b) {
      if (r->boxes[b].n > 0) {
        std::memcpy(mask + o, r->h_mask_stage + r->mask_off[b], r->boxes[b].n);
        o += r->boxes[b].n;
      }
    }
  }
  return SSB_OK;
}

int ssb_ransac_plane_batch(ssb_ransac* r, const void* msg, const ssb_cloud_layout* layout, const ssb_bbox* boxes, int n_boxes,
                           const int* triples, int n_hyp, const ssb_ransac_opts* opts, ssb_plane_result* results, int* counts,
                           unsigned char* mask) {
  int rc = ssb_ransac_upload(r, msg, layout, boxes, n_boxes, triples, n_hyp, opts);
  if (rc) return rc;
  rc = run_device(r);
  if (rc) return rc;
  return ssb_ransac_fetch(r, results, counts, mask);
}

// pcl::SampleConsensusModel::drawIndexSample driven by boost::variate_generator<boost::mt19937&, boost::uniform_int<>>
// (sac_model.h).  boost::mt19937 is std::mt19937 (same parameters and seeding); uniform_int<>(0, INT_MAX) over a 32-bit
// engine is generate_uniform_int's bucket division with bucket_size 2 (0xffffffff / 0x80000000 = 1, and the remainder
// equals the range, so it is incremented), i.e. the engine output shifted right by one, never rejected.
int ssb_ransac_pcl_samples(int n_indices, int n_draws, unsigned seed, int* triples) {
  if (n_draws < 0 || (n_draws > 0 && !triples)) return SSB_ERR_INVALID;
  if (n_indices < 3) {
    std::fill(triples, triples + 3 * (size_t)n_draws, 0);
    return SSB_OK;
  }
  std::mt19937 alg(seed);
  std::vector<int> shuffled(n_indices);
  for (int i = 0; i < n_indices; ++i) shuffled[i] = i;
  const size_t index_size = (size_t)n_indices;
  for (int d = 0; d < n_draws; ++d) {
    for (unsigned i = 0; i < 3; ++i) {
      const int rnd = (int)(alg() >> 1);
      std::swap(shuffled[i], shuffled[i + ((size_t)rnd % (index_size - i))]);
    }
    triples[3 * (size_t)d + 0] = shuffled[0];
    triples[3 * (size_t)d + 1] = shuffled[1];
    triples[3 * (size_t)d + 2] = shuffled[2];
  }
  return SSB_OK;
}

int ssb_crop_bbox(ssb_ransac* r, const void* msg, const ssb_cloud_layout* layout, const ssb_bbox* box, float* out) {
  if (!r || !msg || !layout || !box) return SSB_ERR_INVALID;
  ssb_ransac_opts o;
  ssb_ransac_default_opts(&o);
  int rc = ssb_ransac_upload(r, msg, layout, box, 1, nullptr, 0, &o);
  if (rc) return rc;
  if (r->boxes[0].n < 0) return -1;
  if (r->boxes[0].n == 0 || !out) return r->boxes[0].n;
  dim3 grid(8, 1);
  k_crop<<<grid, 256, 0, r->stream>>>(r->d_msg.p, r->layout, r->d_boxes.p, r->d_crop.p);
  r->launches++;
  SSB_CUDA_CHECK(cudaMemcpyAsync(out, r->d_crop.p, (size_t)r->boxes[0].n * sizeof(float4), cudaMemcpyDeviceToHost, r->stream));
  SSB_CUDA_CHECK(cudaStreamSynchronize(r->stream));
  return r->boxes[0].n;
}

// ---- the live segmentation path: integral-image normals + organised multi-plane segmentation (ssb_organized.cuh) ----
void ssb_organized_default_opts(ssb_organized_opts* o) {
  if (!o) return;
  std::memset(o, 0, sizeof(*o));
  o->max_depth_change_factor = 0.03f;      // plane_segmentation.cpp:98
  o->normal_smoothing_size = 20.0f;        // :99
  o->min_inliers = 500;                    // num_point_seg default, plane_segmentation.cpp:7
  o->angular_threshold = 0.017453f * 2.0f; // :140
  o->distance_threshold = 0.02f;           // :141
  o->maximum_curvature = 0.001f;           // pcl::OrganizedMultiPlaneSegmentation default
  o->norm_point_thres = 5000;              // plane_segmentation.cpp:8
}

int ssb_organized_planes(ssb_ransac* r, const void* msg, const ssb_cloud_layout* layout, const ssb_bbox* boxes, int n_boxes,
                         const ssb_organized_opts* opts, int max_regions, ssb_planar_region* regions, int* n_regions, int* n_inliers,
                         float* normals_out, int* labels_out, float* dist_out) {
  using namespace ssb_org;
  if (!r || !msg || !layout || (n_boxes > 0 && !boxes) || n_boxes < 0 || max_regions < 0 || (n_boxes > 0 && !n_regions)) {
    set_error("ssb_organized_planes: invalid argument");
    return SSB_ERR_INVALID;
  }
  ssb_organized_opts od;
  if (!opts) {
    ssb_organized_default_opts(&od);
    opts = &od;
  }
  ssb_ransac_opts ro;
  ssb_ransac_default_opts(&ro);
  int rc = ssb_ransac_upload(r, msg, layout, boxes, n_boxes, nullptr, 0, &ro);   // cloud + crop plan on the device
  if (rc) return rc;
  const int nb = n_boxes;
  if (nb == 0) return SSB_OK;
  if (layout->width > ORG_THREADS) {
    set_error("ssb_organized_planes: clouds wider than %d pixels are not supported", ORG_THREADS);
    return SSB_ERR_INVALID;
  }
  std::vector<OrgBox> ob(nb);
  long long ii_total = 0;
  for (int b = 0; b < nb; ++b) {
    const BoxInfo& B = r->boxes[b];
    OrgBox& o = ob[b];
    o.w = B.w;
    o.h = B.h;
    // computeNormalsFromPointCloud returns no normals for crops below norm_point_thres (:93) and the caller skips them
    o.n = (B.n > 0 && B.n >= opts->norm_point_thres) ? B.n : 0;
    o.pt_off = B.pt_off;
    o.ii_off = ii_total;
    if (o.n > 0) ii_total += (long long)(B.w + 1) * (B.h + 1);
  }
  const size_t tot = (size_t)std::max<long long>(r->total_pts, 1);
  if ((rc = r->d_oboxes.ensure(nb)) || (rc = r->d_change.ensure(tot)) || (rc = r->d_dist.ensure(tot)) || (rc = r->d_planed.ensure(tot)) ||
      (rc = r->d_nrm.ensure(tot)) || (rc = r->d_ii.ensure(9 * (size_t)std::max<long long>(ii_total, 1))) ||
      (rc = r->d_ic.ensure((size_t)std::max<long long>(ii_total, 1))) || (rc = r->d_parent.ensure(tot)) || (rc = r->d_label.ensure(tot)) ||
      (rc = r->d_count.ensure(tot)) || (rc = r->d_l2m.ensure(tot)) || (rc = r->d_nlabels.ensure(nb)) || (rc = r->d_ncand.ensure(nb)) ||
      (rc = r->d_nreg.ensure(nb)) || (rc = r->d_oreg.ensure((size_t)nb * ORG_MAXR)) || (rc = r->d_lastev.ensure((size_t)nb * ORG_MAXR)))
    return rc;
  cudaStream_t s = r->stream;
  for (int k = 0; k < 2; ++k)
    if (!r->oev[k]) SSB_CUDA_CHECK(cudaEventCreate(&r->oev[k]));
  SSB_CUDA_CHECK(cudaMemcpyAsync(r->d_oboxes.p, ob.data(), nb * sizeof(OrgBox), cudaMemcpyHostToDevice, s));
  OrgOpts O;
  O.max_depth_change_factor = opts->max_depth_change_factor;
  O.smoothing_size = opts->normal_smoothing_size;
  O.min_inliers = opts->min_inliers;
  O.cos_angular = (float)std::cos((double)opts->angular_threshold);
  O.distance_threshold = opts->distance_threshold;
  O.maximum_curvature = opts->maximum_curvature;
  O.refine_distance = 0.02f;   // PlaneRefinementComparator's default; the reference leaves it untouched
  O.norm_point_thres = opts->norm_point_thres;
  SSB_CUDA_CHECK(cudaEventRecord(r->oev[0], s));
  dim3 g2(8, nb);
  k_crop<<<g2, 256, 0, s>>>(r->d_msg.p, r->layout, r->d_boxes.p, r->d_crop.p);
  SSB_CUDA_CHECK(cudaMemsetAsync(r->d_change.p, 255, tot, s));
  k_org_change<<<g2, 256, 0, s>>>(r->d_crop.p, r->d_oboxes.p, O, r->d_change.p);
  k_org_distance<<<nb, ORG_THREADS, 0, s>>>(r->d_oboxes.p, r->d_change.p, r->d_dist.p);
  k_org_integral<<<nb, ORG_THREADS, 0, s>>>(r->d_crop.p, r->d_oboxes.p, r->d_ii.p, r->d_ic.p);
  dim3 g3(32, nb);
  k_org_normals<<<g3, 256, 0, s>>>(r->d_crop.p, r->d_oboxes.p, O, r->d_dist.p, r->d_ii.p, r->d_ic.p, r->d_nrm.p, r->d_planed.p);
  k_org_cc_init<<<g2, 256, 0, s>>>(r->d_crop.p, r->d_oboxes.p, r->d_parent.p);
  k_org_cc_merge<<<g3, 256, 0, s>>>(r->d_crop.p, r->d_oboxes.p, O, r->d_nrm.p, r->d_planed.p, r->d_parent.p);
  k_org_cc_label<<<nb, ORG_THREADS, 0, s>>>(r->d_oboxes.p, r->d_parent.p, r->d_label.p, r->d_count.p, r->d_nlabels.p);
  k_org_candidates<<<nb, ORG_THREADS, 0, s>>>(r->d_oboxes.p, O, r->d_count.p, r->d_nlabels.p, r->d_oreg.p, r->d_ncand.p, r->d_l2m.p);
  dim3 g4(nb, ORG_MAXR);
  k_org_moments<<<g4, 32, 0, s>>>(r->d_crop.p, r->d_oboxes.p, O, r->d_label.p, r->d_ncand.p, r->d_oreg.p);
  k_org_select<<<nb, 64, 0, s>>>(r->d_oboxes.p, r->d_ncand.p, r->d_oreg.p, r->d_nreg.p, r->d_l2m.p, r->d_lastev.p);
  k_org_refine<<<nb, ORG_THREADS, 0, s>>>(r->d_crop.p, r->d_oboxes.p, O, r->d_label.p, r->d_l2m.p, r->d_oreg.p, r->d_lastev.p);
  k_org_boundary<<<g4, 256, 0, s>>>(r->d_crop.p, r->d_oboxes.p, r->d_label.p, r->d_nreg.p, r->d_lastev.p, r->d_oreg.p);
  r->launches += 14;
  SSB_CUDA_CHECK(cudaEventRecord(r->oev[1], s));
  SSB_CUDA_CHECK(cudaGetLastError());
  std::vector<OrgRegion> hreg((size_t)nb * ORG_MAXR);
  std::vector<int> hn(nb);
  SSB_CUDA_CHECK(cudaMemcpyAsync(hreg.data(), r->d_oreg.p, hreg.size() * sizeof(OrgRegion), cudaMemcpyDeviceToHost, s));
  SSB_CUDA_CHECK(cudaMemcpyAsync(hn.data(), r->d_nreg.p, nb * sizeof(int), cudaMemcpyDeviceToHost, s));
  std::vector<float4> hnrm;
  std::vector<int> hlab;
  std::vector<float> hdist;
  if (normals_out) {
    hnrm.resize(tot);
    SSB_CUDA_CHECK(cudaMemcpyAsync(hnrm.data(), r->d_nrm.p, tot * sizeof(float4), cudaMemcpyDeviceToHost, s));
  }
  if (labels_out) {
    hlab.resize(tot);
    SSB_CUDA_CHECK(cudaMemcpyAsync(hlab.data(), r->d_label.p, tot * sizeof(int), cudaMemcpyDeviceToHost, s));
  }
  if (dist_out) {
    hdist.resize(tot);
    SSB_CUDA_CHECK(cudaMemcpyAsync(hdist.data(), r->d_dist.p, tot * sizeof(float), cudaMemcpyDeviceToHost, s));
  }
  SSB_CUDA_CHECK(cudaStreamSynchronize(s));
  float ms = 0.0f;
  SSB_CUDA_CHECK(cudaEventElapsedTime(&ms, r->oev[0], r->oev[1]));
  r->org_ms = ms;
  size_t o = 0;
  for (int b = 0; b < nb; ++b) {
    const BoxInfo& B = r->boxes[b];
    if (B.n < 0) {
      n_regions[b] = -1;   // "spurious" bbox (plane_segmentation.cpp:34-38)
      continue;
    }
    if (ob[b].n == 0) {
      n_regions[b] = -2;   // fewer than norm_point_thres points: no normals, crop skipped (:93, point_cloud_segmentation.h:164-165)
    } else {
      n_regions[b] = hn[b];
      for (int k = 0; k < hn[b] && k < max_regions; ++k) {
        const OrgRegion& R = hreg[(size_t)b * ORG_MAXR + k];
        if (regions) {
          ssb_planar_region& q = regions[(size_t)b * max_regions + k];
          for (int c = 0; c < 3; ++c) q.centroid[c] = R.centroid[c];
          for (int c = 0; c < 4; ++c) q.model[c] = R.model[c];
          q.contour_points = R.contour_n;
          q.area = R.area;
        }
        if (n_inliers) n_inliers[(size_t)b * max_regions + k] = R.n_inliers;
      }
    }
    if (B.n > 0) {   // per-point outputs: plain concatenation over the non-spurious boxes
      if (normals_out) std::memcpy(normals_out + 4 * o, hnrm.data() + B.pt_off, (size_t)B.n * sizeof(float4));
      if (labels_out) std::memcpy(labels_out + o, hlab.data() + B.pt_off, (size_t)B.n * sizeof(int));
      if (dist_out) std::memcpy(dist_out + o, hdist.data() + B.pt_off, (size_t)B.n * sizeof(float));
      if (ob[b].n == 0) {
        if (labels_out) std::fill(labels_out + o, labels_out + o + B.n, -1);
      }
      o += B.n;
    }
  }
  return SSB_OK;
}
// CUDA-event time of the device pipeline of the last ssb_organized_planes call (crop .. boundary), ms
double ssb_organized_last_ms(ssb_ransac* r) { return r ? r->org_ms : 0.0; }

}  // extern "C"

// =============================================================================================================
// The dormant plane-clustering chain (ssb_cluster.cuh): host side
// =============================================================================================================
namespace ssb {
// cv::RNG (multiply-with-carry): the random centres of cv::kmeans(KMEANS_RANDOM_CENTERS) are the only thing drawn from it
struct CvRng {
  unsigned long long state;
  unsigned next() {
    state = (unsigned long long)(unsigned)state * 4164903690ULL + (unsigned)(state >> 32);
    return (unsigned)state;
  }
  float uniform01() { return next() * 2.3283064365386962890625e-10f; }
};
static int cl_rank(ssb_ransac* r, const unsigned char* flags, int n, int* count) {
  int rc;
  if ((rc = r->c_pos.ensure((size_t)std::max(n, 1))) || (rc = r->c_count.ensure(16))) return rc;
  ssb_cl::k_cl_rank<<<1, ssb_cl::CL_THREADS, 0, r->stream>>>(flags, n, r->c_pos.p, r->c_count.p);
  r->launches++;
  SSB_CUDA_CHECK(cudaMemcpyAsync(count, r->c_count.p, sizeof(int), cudaMemcpyDeviceToHost, r->stream));
  SSB_CUDA_CHECK(cudaStreamSynchronize(r->stream));
  return SSB_OK;
}
// cv::kmeans on device-resident samples; the winning attempt's labels stay on the device (*d_labels), centres go to the host
static int cl_kmeans(ssb_ransac* r, const float* d_data, int N, int dims, int K, int max_count, double eps, int attempts, CvRng& rng,
                     RBuf<int>& labelbuf, int** d_labels, float* h_centers, double* compactness) {
  if (N < K || K < 1 || K > ssb_cl::CL_MAXK || dims < 1 || dims > ssb_cl::CL_MAXD) {
    set_error("kmeans: need 1 <= K <= %d, 1 <= dims <= %d and at least K samples (n = %d, K = %d, dims = %d)", ssb_cl::CL_MAXK,
              ssb_cl::CL_MAXD, N, K, dims);
    return SSB_ERR_INVALID;
  }
  attempts = std::max(attempts, 1);
  eps = std::max(eps, 0.0);
  max_count = std::min(std::max(max_count, 2), 100);
  if (K == 1) {
    attempts = 1;
    max_count = 2;
  }
  int rc;
  if ((rc = labelbuf.ensure((size_t)attempts * N)) || (rc = r->c_init.ensure((size_t)attempts * K * dims)) ||
      (rc = r->c_centers.ensure((size_t)attempts * K * dims)) || (rc = r->c_compact.ensure(attempts)) || (rc = r->c_lohi.ensure(2 * ssb_cl::CL_MAXD)))
    return rc;
  cudaStream_t s = r->stream;
  ssb_cl::k_cl_minmax<<<1, ssb_cl::CL_THREADS, 0, s>>>(d_data, N, dims, r->c_lohi.p);
  float lohi[2 * ssb_cl::CL_MAXD];
  SSB_CUDA_CHECK(cudaMemcpyAsync(lohi, r->c_lohi.p, sizeof(lohi), cudaMemcpyDeviceToHost, s));
  SSB_CUDA_CHECK(cudaStreamSynchronize(s));
  std::vector<float> init((size_t)attempts * K * dims);
  const float margin = 1.f / dims;   // generateRandomCenter
  for (int a = 0; a < attempts; ++a)
    for (int k = 0; k < K; ++k)
      for (int j = 0; j < dims; ++j) {
        const float lo = lohi[j], hi = lohi[ssb_cl::CL_MAXD + j];
        const float u = rng.uniform01();
        const float t = u * (1.f + margin * 2.f) - margin;
        init[((size_t)a * K + k) * dims + j] = t * (hi - lo) + lo;
      }
  SSB_CUDA_CHECK(cudaMemcpyAsync(r->c_init.p, init.data(), init.size() * sizeof(float), cudaMemcpyHostToDevice, s));
  ssb_cl::KmArgs A;
  A.data = d_data;
  A.N = N;
  A.dims = dims;
  A.K = K;
  A.max_count = max_count;
  A.eps2 = eps * eps;
  A.init = r->c_init.p;
  A.labels = labelbuf.p;
  A.centers = r->c_centers.p;
  A.compactness = r->c_compact.p;
  void (*kern)(ssb_cl::KmArgs) = dims == 1 ? ssb_cl::k_cl_kmeans<1> : dims == 2 ? ssb_cl::k_cl_kmeans<2> : dims == 3 ? ssb_cl::k_cl_kmeans<3> : ssb_cl::k_cl_kmeans<4>;
  if (!(r->km_smem_set & (1 << dims))) {   // opt-in once per handle / device (80 KB of dynamic shared memory: the staged tile)
    SSB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssb_cl::CL_KMEANS_SMEM));
    r->km_smem_set |= 1 << dims;
  }
  kern<<<attempts, ssb_cl::CL_THREADS, ssb_cl::CL_KMEANS_SMEM, s>>>(A);
  r->launches += 2;
  std::vector<double> comp(attempts);
  std::vector<float> cen((size_t)attempts * K * dims);
  SSB_CUDA_CHECK(cudaMemcpyAsync(comp.data(), r->c_compact.p, attempts * sizeof(double), cudaMemcpyDeviceToHost, s));
  SSB_CUDA_CHECK(cudaMemcpyAsync(cen.data(), r->c_centers.p, cen.size() * sizeof(float), cudaMemcpyDeviceToHost, s));
  SSB_CUDA_CHECK(cudaStreamSynchronize(s));
  SSB_CUDA_CHECK(cudaGetLastError());
  int best = 0;
  double bc = std::numeric_limits<double>::max();
  for (int a = 0; a < attempts; ++a)
    if (comp[a] < bc) {   // `compactness < best_compactness`: the first of equal attempts stays
      bc = comp[a];
      best = a;
    }
  *d_labels = labelbuf.p + (size_t)best * N;
  std::memcpy(h_centers, cen.data() + (size_t)best * K * dims, (size_t)K * dims * sizeof(float));
  if (compactness) *compactness = bc;
  return SSB_OK;
}
// pcl::ConvexHull::performReconstruction2D: which two coordinates (the plane normal must not lie within 10 degrees of the
// dropped axis' complement: projection_angle_thresh_ = cos(0.174532925)) -> 0: (x, y), 1: (y, z), 2: (x, z), -1: none
static int cl_hull_axes(const float* coef, int* iu, int* iv) {
  const double nn = std::sqrt((double)coef[0] * coef[0] + (double)coef[1] * coef[1] + (double)coef[2] * coef[2]);
  const float tx = std::fabs((float)(coef[0] / nn)), ty = std::fabs((float)(coef[1] / nn)), tz = std::fabs((float)(coef[2] / nn));
  const float thr = (float)std::cos(0.174532925);
  bool xy = true, yz = true, xz = true;
  if (tz > thr) xz = yz = false;
  if (tx > thr) xz = xy = false;
  if (ty > thr) xy = yz = false;
  if (xy) { *iu = 0; *iv = 1; return 0; }
  if (yz) { *iu = 1; *iv = 2; return 1; }
  if (xz) { *iu = 0; *iv = 2; return 2; }
  return -1;
}
// strictly convex vertices of the candidate points (Andrew's monotone chain, orientation in double, collinear points dropped,
// of identical points the first stands for all), then PCL's output order: decreasing atan2 about the centroid of the vertices
static void cl_hull_host(const float4* pts, int m, int iu, int iv, std::vector<int>& hull) {
  hull.clear();
  if (m <= 0) return;
  auto C = [&](int i, int ax) { return ax == 0 ? pts[i].x : ax == 1 ? pts[i].y : pts[i].z; };
  auto U = [&](int i) { return (double)C(i, iu); };
  auto V = [&](int i) { return (double)C(i, iv); };
  std::vector<int> ord(m);
  for (int i = 0; i < m; ++i) ord[i] = i;
  std::sort(ord.begin(), ord.end(), [&](int a, int b) {
    if (U(a) != U(b)) return U(a) < U(b);
    if (V(a) != V(b)) return V(a) < V(b);
    return a < b;
  });
  std::vector<int> uq;
  for (int k = 0; k < m; ++k)
    if (uq.empty() || U(ord[k]) != U(uq.back()) || V(ord[k]) != V(uq.back())) uq.push_back(ord[k]);
  const int q = (int)uq.size();
  if (q < 3) {
    hull = uq;
  } else {
    std::vector<int> st(2 * (size_t)q);
    int k = 0;
    auto turn = [&](int a, int b, int c) { return (U(b) - U(a)) * (V(c) - V(a)) - (V(b) - V(a)) * (U(c) - U(a)); };
    for (int i = 0; i < q; ++i) {
      while (k >= 2 && turn(st[k - 2], st[k - 1], uq[i]) <= 0) --k;
      st[k++] = uq[i];
    }
    for (int i = q - 2, t = k + 1; i >= 0; --i) {
      while (k >= t && turn(st[k - 2], st[k - 1], uq[i]) <= 0) --k;
      st[k++] = uq[i];
    }
    hull.assign(st.begin(), st.begin() + (k - 1));
  }
  std::vector<int> byidx = hull;
  std::sort(byidx.begin(), byidx.end());
  double cu = 0, cv = 0;
  for (int i : byidx) {
    cu += U(i);
    cv += V(i);
  }
  const float fcu = (float)(cu / (double)byidx.size()), fcv = (float)(cv / (double)byidx.size());
  std::vector<std::pair<double, int>> ang;
  for (int i : byidx) {
    const float du = C(i, iu) - fcu, dv = C(i, iv) - fcv;
    ang.push_back({std::atan2((double)dv, (double)du) + M_PI, i});
  }
  std::stable_sort(ang.begin(), ang.end(), [](const std::pair<double, int>& a, const std::pair<double, int>& b) { return a.first > b.first; });
  hull.clear();
  for (auto& a : ang) hull.push_back(a.second);
}
// ProjectInliers + ConvexHull of n device points (d_pts) selected by the device mask.  rows3 / src: host outputs.
static int cl_project_hull(ssb_ransac* r, const float4* d_pts, const unsigned char* d_mask, int n, const float* coef, float* rows3, int* src,
                           int max_rows, int* n_inliers, int* n_hull) {
  *n_hull = 0;
  if (n_inliers) *n_inliers = 0;
  if (n <= 0) return SSB_OK;
  cudaStream_t s = r->stream;
  int nin = 0, rc;
  if ((rc = cl_rank(r, d_mask, n, &nin))) return rc;
  if (n_inliers) *n_inliers = nin;
  int iu, iv;
  if (nin == 0 || cl_hull_axes(coef, &iu, &iv) < 0) return SSB_OK;
  if ((rc = r->c_proj.ensure(nin)) || (rc = r->c_src.ensure(nin)) || (rc = r->c_flags.ensure((size_t)std::max(n, nin))) || (rc = r->c_ext.ensure(8)) ||
      (rc = r->c_cand.ensure(nin)) || (rc = r->c_cidx.ensure(nin)) || (rc = r->c_csrc.ensure(nin)))
    return rc;
  // sac_model_plane.hpp projectPoints: mc = (a, b, c, 0).normalized()  [Eigen 4-float squaredNorm: (p0 + p1) + (p2 + p3)]
  float mc[4] = {coef[0], coef[1], coef[2], 0.f};
  const float sq = (mc[0] * mc[0] + mc[1] * mc[1]) + (mc[2] * mc[2] + mc[3] * mc[3]);
  const float nrm = std::sqrt(sq);
  for (int k = 0; k < 4; ++k) mc[k] = mc[k] / nrm;
  ssb_cl::k_cl_project<<<(n + 255) / 256, 256, 0, s>>>(d_pts, d_mask, r->c_pos.p, n, mc[0], mc[1], mc[2], coef[3], r->c_proj.p, r->c_src.p);
  ssb_cl::k_cl_extremes<<<1, ssb_cl::CL_THREADS, 0, s>>>(r->c_proj.p, nin, iu, iv, r->c_ext.p);
  ssb_cl::k_cl_flag_outside<<<(nin + 255) / 256, 256, 0, s>>>(r->c_proj.p, nin, iu, iv, r->c_ext.p, r->c_flags.p);
  r->launches += 3;
  int ncand = 0;
  if ((rc = cl_rank(r, r->c_flags.p, nin, &ncand))) return rc;
  if (ncand == 0) return SSB_OK;
  ssb_cl::k_cl_scatter_cand<<<(nin + 255) / 256, 256, 0, s>>>(r->c_proj.p, r->c_flags.p, r->c_pos.p, nin, r->c_cand.p, r->c_cidx.p);
  r->launches++;
  std::vector<float4> cand(ncand);
  std::vector<int> cidx(ncand), hsrc(nin);
  SSB_CUDA_CHECK(cudaMemcpyAsync(cand.data(), r->c_cand.p, (size_t)ncand * sizeof(float4), cudaMemcpyDeviceToHost, s));
  SSB_CUDA_CHECK(cudaMemcpyAsync(cidx.data(), r->c_cidx.p, (size_t)ncand * sizeof(int), cudaMemcpyDeviceToHost, s));
  SSB_CUDA_CHECK(cudaMemcpyAsync(hsrc.data(), r->c_src.p, (size_t)nin * sizeof(int), cudaMemcpyDeviceToHost, s));
  SSB_CUDA_CHECK(cudaStreamSynchronize(s));
  SSB_CUDA_CHECK(cudaGetLastError());
  std::vector<int> hull;
  cl_hull_host(cand.data(), ncand, iu, iv, hull);
  *n_hull = (int)hull.size();
  for (int k = 0; k < (int)hull.size() && k < max_rows; ++k) {
    const float4 p = cand[hull[k]];
    rows3[3 * k] = p.x;
    rows3[3 * k + 1] = p.y;
    rows3[3 * k + 2] = p.z;
    if (src) src[k] = hsrc[cidx[hull[k]]];
  }
  return SSB_OK;
}
}  // namespace ssb

extern "C" {

void ssb_cluster_default_opts(ssb_cluster_opts* o) {
  if (!o) return;
  std::memset(o, 0, sizeof(*o));
  o->num_centroids_normals = 4;
  o->num_centroids_distance = 2;
  o->kmeans_attempts = 10;
  o->kmeans_max_count = 10;
  o->kmeans_epsilon = 0.01;
  o->min_cluster_points = 500;
  o->centroid_tolerance = 0.3f;
  o->ransac_hypotheses = 0;
  o->ransac_seed = 12345u;
}

int ssb_kmeans(ssb_ransac* r, const float* data, int n, int dims, int K, int max_count, double epsilon, int attempts,
               unsigned long long* rng_state, int* labels, float* centers, double* compactness) {
  if (!r || !data || !rng_state || !labels || !centers || n < 1) {
    set_error("ssb_kmeans: invalid argument");
    return SSB_ERR_INVALID;
  }
  SSB_CUDA_CHECK(cudaSetDevice(r->device));
  int rc;
  if ((rc = r->c_data.ensure((size_t)n * std::max(dims, 1)))) return rc;
  SSB_CUDA_CHECK(cudaMemcpyAsync(r->c_data.p, data, (size_t)n * dims * sizeof(float), cudaMemcpyHostToDevice, r->stream));
  CvRng rng{*rng_state ? *rng_state : 0xffffffffULL};
  int* d_lab = nullptr;
  if ((rc = cl_kmeans(r, r->c_data.p, n, dims, K, max_count, epsilon, attempts, rng, r->c_labels, &d_lab, centers, compactness))) return rc;
  SSB_CUDA_CHECK(cudaMemcpyAsync(labels, d_lab, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, r->stream));
  SSB_CUDA_CHECK(cudaStreamSynchronize(r->stream));
  *rng_state = rng.state;
  return SSB_OK;
}

int ssb_project_hull(ssb_ransac* r, const float* pts4, const unsigned char* mask, int n, const float coef[4], float* rows3, int* src,
                     int max_rows, int* n_inliers) {
  if (!r || !pts4 || !mask || !coef || n < 0 || (max_rows > 0 && !rows3)) {
    set_error("ssb_project_hull: invalid argument");
    return SSB_ERR_INVALID;
  }
  SSB_CUDA_CHECK(cudaSetDevice(r->device));
  int rc;
  if ((rc = r->c_pts.ensure((size_t)std::max(n, 1))) || (rc = r->c_maskh.ensure((size_t)std::max(n, 1)))) return rc;
  SSB_CUDA_CHECK(cudaMemcpyAsync(r->c_pts.p, pts4, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, r->stream));
  SSB_CUDA_CHECK(cudaMemcpyAsync(r->c_maskh.p, mask, (size_t)n, cudaMemcpyHostToDevice, r->stream));
  int nh = 0;
  if ((rc = cl_project_hull(r, r->c_pts.p, r->c_maskh.p, n, coef, rows3, src, max_rows, n_inliers, &nh))) return rc;
  return nh;
}

int ssb_cluster_planes(ssb_ransac* r, const float* cloud4, const float* normals4, int n, const float T[16], const ssb_cluster_opts* opts_in,
                       unsigned long long* rng_state, float* rows8, int max_rows, int* n_rows, ssb_plane_cluster* clusters, int max_clusters,
                       int* n_clusters, int* labels_out, float* centers_out) {
  if (!r || !cloud4 || !normals4 || !T || !rng_state || !n_rows || !n_clusters || n < 0 || (max_rows > 0 && !rows8) || (max_clusters > 0 && !clusters)) {
    set_error("ssb_cluster_planes: invalid argument");
    return SSB_ERR_INVALID;
  }
  ssb_cluster_opts o;
  if (opts_in)
    o = *opts_in;
  else
    ssb_cluster_default_opts(&o);
  const int Kn = o.num_centroids_normals, Kd = o.num_centroids_distance;
  *n_rows = 0;
  *n_clusters = 0;
  if (labels_out)
    for (int i = 0; i < n; ++i) labels_out[i] = -1;
  if (n == 0) return 0;
  SSB_CUDA_CHECK(cudaSetDevice(r->device));
  cudaStream_t s = r->stream;
  int rc;
  if ((rc = r->c_cloud.ensure(n)) || (rc = r->c_nrm.ensure(n)) || (rc = r->c_flags.ensure(n)) || (rc = r->c_keep.ensure(n)) ||
      (rc = r->c_data.ensure((size_t)3 * n)) || (rc = r->c_mem.ensure(n)) || (rc = r->c_dist.ensure(n)) || (rc = r->c_pts.ensure((size_t)n + 64)))
    return rc;
  SSB_CUDA_CHECK(cudaMemcpyAsync(r->c_cloud.p, cloud4, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, s));
  SSB_CUDA_CHECK(cudaMemcpyAsync(r->c_nrm.p, normals4, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, s));
  const int nb256 = (n + 255) / 256;
  // removeNans :479-502
  ssb_cl::k_cl_flag_valid<<<nb256, 256, 0, s>>>(r->c_nrm.p, n, r->c_flags.p);
  r->launches++;
  int m = 0;
  if ((rc = cl_rank(r, r->c_flags.p, n, &m))) return rc;
  if (m <= 10 || m < Kn) return 0;   // :316-320
  ssb_cl::k_cl_scatter_valid<<<nb256, 256, 0, s>>>(r->c_nrm.p, r->c_flags.p, r->c_pos.p, n, r->c_keep.p, r->c_data.p);
  r->launches++;
  CvRng rng{*rng_state ? *rng_state : 0xffffffffULL};
  int* d_lab = nullptr;
  std::vector<float> centers((size_t)Kn * 3);
  if ((rc = cl_kmeans(r, r->c_data.p, m, 3, Kn, o.kmeans_max_count, o.kmeans_epsilon, o.kmeans_attempts, rng, r->c_labels, &d_lab, centers.data(), nullptr)))
    return rc;
  if (centers_out) std::memcpy(centers_out, centers.data(), centers.size() * sizeof(float));
  if (labels_out) {
    std::vector<int> lab(m), keep(m);
    SSB_CUDA_CHECK(cudaMemcpyAsync(lab.data(), d_lab, (size_t)m * sizeof(int), cudaMemcpyDeviceToHost, s));
    SSB_CUDA_CHECK(cudaMemcpyAsync(keep.data(), r->c_keep.p, (size_t)m * sizeof(int), cudaMemcpyDeviceToHost, s));
    SSB_CUDA_CHECK(cudaStreamSynchronize(s));
    for (int k = 0; k < m; ++k) labels_out[keep[k]] = lab[k];
  }
  // normals_of_the_horizontal_plane_in_cam = transformation_mat^T (0, 0, 1, 0) :332-346: the third row of the matrix
  const float hz[3] = {T[8], T[9], T[10]};
  struct Kept {
    int c, d, np;
    size_t off;   // in c_pts
    float dist;
  };
  std::vector<Kept> kept;
  size_t stage = 0;
  const int mb256 = (m + 255) / 256;
  for (int c = 0; c < Kn; ++c) {   // filterCentroids :504-523
    const float* cc = &centers[3 * (size_t)c];
    bool ok = true;
    for (int j = 0; j < 3; ++j)
      ok = ok && ((double)cc[j] < (double)hz[j] + (double)o.centroid_tolerance) && ((double)cc[j] > (double)hz[j] - (double)o.centroid_tolerance);
    if (!ok) continue;
    ssb_cl::k_cl_flag_eq<<<mb256, 256, 0, s>>>(d_lab, m, c, r->c_flags.p);
    r->launches++;
    int md = 0;
    if ((rc = cl_rank(r, r->c_flags.p, m, &md))) return rc;
    if (md < Kd) continue;   // cv::kmeans would throw (fewer samples than clusters)
    ssb_cl::k_cl_scatter_members<<<mb256, 256, 0, s>>>(r->c_cloud.p, r->c_keep.p, r->c_flags.p, r->c_pos.p, m, cc[0], cc[1], cc[2], r->c_mem.p, r->c_dist.p);
    r->launches++;
    int* d_dlab = nullptr;
    std::vector<float> dc(Kd);
    if ((rc = cl_kmeans(r, r->c_dist.p, md, 1, Kd, o.kmeans_max_count, o.kmeans_epsilon, o.kmeans_attempts, rng, r->c_dlabels, &d_dlab, dc.data(), nullptr)))
      return rc;
    const int db256 = (md + 255) / 256;
    for (int d = 0; d < Kd; ++d) {   // :399-425
      ssb_cl::k_cl_flag_eq<<<db256, 256, 0, s>>>(d_dlab, md, d, r->c_flags.p);
      r->launches++;
      int np = 0;
      if ((rc = cl_rank(r, r->c_flags.p, md, &np))) return rc;
      if (!(np > o.min_cluster_points)) continue;
      if ((int)kept.size() >= max_clusters) continue;
      ssb_cl::k_cl_scatter_points<<<db256, 256, 0, s>>>(r->c_cloud.p, r->c_mem.p, r->c_flags.p, r->c_pos.p, md, r->c_pts.p + stage);
      r->launches++;
      kept.push_back({c, d, np, stage, dc[d]});
      stage += (size_t)np;
    }
  }
  *rng_state = rng.state;
  const int ncl = (int)kept.size();
  *n_clusters = ncl;
  if (ncl == 0) return 1;
  // compute2DConvexHull :631-647 on every cluster at once: the clusters are the "crops" of the RANSAC kernels
  {
    ssb_cloud_layout L;
    std::memset(&L, 0, sizeof(L));
    L.width = L.height = 1 << 30;
    std::vector<ssb_bbox> bx(ncl);
    for (int b = 0; b < ncl; ++b) bx[b] = {0, 0, kept[b].np, 1};
    ssb_ransac_default_opts(&r->opts);
    const int K = o.ransac_hypotheses > 0 ? o.ransac_hypotheses : 512;
    r->opts.mode = o.ransac_hypotheses > 0 ? 0 : 1;
    r->nb = ncl;
    r->K = K;
    r->uploaded = false;
    if ((rc = plan(r, &L, bx.data(), ncl))) return rc;
    const size_t nh = (size_t)ncl * K;
    if ((rc = r->d_triples.ensure(3 * nh)) || (rc = r->d_valid.ensure(nh)) || (rc = r->d_counts.ensure(nh)) || (rc = r->d_hyp.ensure(nh))) return rc;
    std::vector<int> tri(3 * nh);
    // every compute2DConvexHull call builds its own pcl::SACSegmentation (:637): a fresh model, i.e. PCL's sample stream
    // from its seed for every cluster
    for (int b = 0; b < ncl; ++b) ssb_ransac_pcl_samples(kept[b].np, K, o.ransac_seed, &tri[(size_t)3 * K * b]);
    SSB_CUDA_CHECK(cudaMemcpyAsync(r->d_triples.p, tri.data(), tri.size() * sizeof(int), cudaMemcpyHostToDevice, s));
    for (int b = 0; b < ncl; ++b)
      SSB_CUDA_CHECK(cudaMemcpyAsync(r->d_crop.p + r->boxes[b].pt_off, r->c_pts.p + kept[b].off, (size_t)kept[b].np * sizeof(float4), cudaMemcpyDeviceToDevice, s));
    if ((rc = run_device(r, true))) return rc;
    std::vector<ssb_plane_result> res(ncl);
    SSB_CUDA_CHECK(cudaMemcpyAsync(res.data(), r->d_results.p, ncl * sizeof(ssb_plane_result), cudaMemcpyDeviceToHost, s));
    SSB_CUDA_CHECK(cudaStreamSynchronize(s));   // `tri` is a local
    int rows = 0;
    std::vector<float> h3;
    for (int b = 0; b < ncl; ++b) {
      ssb_plane_cluster& C = clusters[b];
      std::memset(&C, 0, sizeof(C));
      const float* cc = &centers[3 * (size_t)kept[b].c];
      for (int j = 0; j < 3; ++j) C.normal[j] = cc[j];
      C.distance = kept[b].dist;
      C.normal_label = kept[b].c;
      C.distance_label = kept[b].d;
      C.n_points = kept[b].np;
      std::memcpy(C.coef, res[b].refined, sizeof(C.coef));
      C.row0 = rows;
      if (res[b].status != 0 || res[b].best_hyp < 0) continue;
      // ProjectInliers + ConvexHull :649-662, then one row per hull vertex (getFinalPoseWithNormals :431-477)
      h3.resize((size_t)3 * kept[b].np);
      int nin = 0, nhull = 0;
      if ((rc = cl_project_hull(r, r->d_crop.p + r->boxes[b].pt_off, r->d_mask.p + r->mask_off[b], kept[b].np, res[b].refined, h3.data(), nullptr,
                                kept[b].np, &nin, &nhull)))
        return rc;
      C.n_inliers = nin;
      for (int k = 0; k < nhull && rows < max_rows; ++k, ++rows) {
        float* r8 = rows8 + 8 * (size_t)rows;
        r8[0] = h3[3 * k];
        r8[1] = h3[3 * k + 1];
        r8[2] = h3[3 * k + 2];
        r8[3] = cc[0];
        r8[4] = cc[1];
        r8[5] = cc[2];
        r8[6] = kept[b].dist;
        r8[7] = 0.f;
      }
      C.n_rows = rows - C.row0;
    }
    *n_rows = rows;
  }
  return 1;
}

}  // extern "C"
