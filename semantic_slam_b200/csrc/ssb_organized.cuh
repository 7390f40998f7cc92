// ssb_organized.cuh — the LIVE segmentation path of the reference on the device:
//   plane_segmentation::computeNormalsFromPointCloud  src/planar_segmentation/plane_segmentation.cpp:84-106
//       (pcl::IntegralImageNormalEstimation, COVARIANCE_MATRIX, depth-change factor 0.03, smoothing 20)
//   plane_segmentation::multiPlaneSegmentation        :136-156 (pcl::OrganizedMultiPlaneSegmentation::segmentAndRefine,
//       2 degrees / 0.02 m, findLabeledRegionBoundary, calculatePolygonArea :189)
// for every bbox crop of a frame at once.  PCL's pipeline is a chain of raster-order recurrences; each is evaluated here
// with the SAME per-pixel arithmetic in an order that respects its data dependencies, so the results are bit-identical to
// the sequential restatement (oracle/oracle_segment.cpp) instead of "equal up to borderline pixels":
//   depth-change map      : order-free (only writes zeros)                                   -> one thread per pixel
//   chamfer distance map  : (r,c) needs (r-1,c-1..c+1),(r,c-1)  -> skewed wavefront t = c + 2r, one CTA per crop
//   integral images       : (r,c+1) needs (r-1,c+1),(r,c),(r-1,c) in double -> wavefront t = r + c, one CTA per crop
//   normals               : independent per pixel (4-corner sums, float covariance, pcl::eigen33)
//   connected components  : the partition does not depend on the scan order -> lock-free union-find (root = smallest
//                           raster index), labels renumbered by first pixel like PCL's run table
//   region moments        : float sums in raster order (pcl::computeMeanAndCovarianceMatrix) -> one warp per region:
//                           coalesced loads, the matching pixels added one by one (ballot + shuffle)
//   refinement            : two raster sweeps whose label hand-offs run along rows -> rows in sequence, each row resolved
//                           in closed form with two block-wide max-scans (nearest seed / nearest failing pixel)
//   boundary + area       : Moore tracing from the last inlier, float cross-product sum in order -> one walker per region
// This translation unit is compiled with -fmad=false: no product-sum is contracted into an FMA.
#pragma once
#include <cfloat>

namespace ssb_org {

constexpr int ORG_THREADS = 1024;
constexpr int ORG_MAXR = 64;   // candidate regions per crop (labels with more than min_inliers pixels)

struct OrgBox {   // per bbox
  int w, h, n;    // n = w*h, <= 0: spurious / skipped
  int pt_off;     // offset of the crop in the per-point arrays
  long long ii_off;  // offset of the crop in the integral-image arrays ((w+1)(h+1) entries)
};
struct OrgRegion {
  float centroid[3];
  float model[4];
  int label;        // label of the region after the connected components
  int n_inliers;    // after the refinement
  int last_inlier;  // start of the boundary trace
  int contour_n;
  float area;
  int keep;         // passed the curvature gate
  int pad;
};
struct OrgOpts {
  float max_depth_change_factor, smoothing_size;
  int min_inliers;
  float cos_angular, distance_threshold, maximum_curvature, refine_distance;
  int norm_point_thres;
};

__device__ __forceinline__ bool fin(float v) { return isfinite(v); }

// ---- pcl::eigen33 (float; the transcendental functions in double, rounded — see the oracle header) ------------------
__device__ inline void compute_roots2(float b, float c, float* r) {
  r[0] = 0.0f;
  float d = (float)((double)(b * b) - 4.0 * (double)c);
  if (d < 0.0f) d = 0.0f;
  const float sd = (float)sqrt((double)d);
  r[2] = 0.5f * (b + sd);
  r[1] = 0.5f * (b - sd);
}
__device__ inline void swapf(float& a, float& b) {
  const float t = a;
  a = b;
  b = t;
}
__device__ inline void compute_roots(const float* m, float* r) {
  const float m00 = m[0], m01 = m[1], m02 = m[2], m11 = m[4], m12 = m[5], m22 = m[8];
  const float c0 = m00 * m11 * m22 + 2.0f * m01 * m02 * m12 - m00 * m12 * m12 - m11 * m02 * m02 - m22 * m01 * m01;
  const float c1 = m00 * m11 - m01 * m01 + m00 * m22 - m02 * m02 + m11 * m22 - m12 * m12;
  const float c2 = m00 + m11 + m22;
  if (fabsf(c0) < FLT_EPSILON) {
    compute_roots2(c2, c1, r);
    return;
  }
  const float s_inv3 = (float)(1.0 / 3.0);
  const float s_sqrt3 = (float)sqrt(3.0);
  const float c2_over_3 = c2 * s_inv3;
  float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
  if (a_over_3 > 0.0f) a_over_3 = 0.0f;
  const float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
  float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
  if (q > 0.0f) q = 0.0f;
  const float rho = (float)sqrt((double)(-a_over_3));
  const float theta = (float)atan2((double)(float)sqrt((double)(-q)), (double)half_b) * s_inv3;
  const float cos_theta = (float)cos((double)theta);
  const float sin_theta = (float)sin((double)theta);
  r[0] = c2_over_3 + 2.0f * rho * cos_theta;
  r[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
  r[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
  if (r[0] >= r[1]) swapf(r[0], r[1]);
  if (r[1] >= r[2]) {
    swapf(r[1], r[2]);
    if (r[0] >= r[1]) swapf(r[0], r[1]);
  }
  if (r[0] <= 0.0f) compute_roots2(c2, c1, r);
}
__device__ inline void cross3f(const float* a, const float* b, float* o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ inline void eigen33(const float* mat, float& eigenvalue, float* evec) {
  float scale = 0.0f;
  for (int k = 0; k < 9; ++k) scale = fmaxf(scale, fabsf(mat[k]));
  if (scale <= FLT_MIN) scale = 1.0f;
  float s[9];
  for (int k = 0; k < 9; ++k) s[k] = mat[k] / scale;
  float roots[3];
  compute_roots(s, roots);
  eigenvalue = roots[0] * scale;
  s[0] -= roots[0];
  s[4] -= roots[0];
  s[8] -= roots[0];
  float v1[3], v2[3], v3[3];
  cross3f(s + 0, s + 3, v1);
  cross3f(s + 0, s + 6, v2);
  cross3f(s + 3, s + 6, v3);
  const float l1 = v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2];
  const float l2 = v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2];
  const float l3 = v3[0] * v3[0] + v3[1] * v3[1] + v3[2] * v3[2];
  const float* v;
  float l;
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
  const float n = (float)sqrt((double)l);
  for (int k = 0; k < 3; ++k) evec[k] = v[k] / n;
}

// ---- depth-change map ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_org_change(const float4* __restrict__ crop, const OrgBox* __restrict__ boxes, OrgOpts O,
                                                    unsigned char* __restrict__ change) {
  const OrgBox B = boxes[blockIdx.y];
  if (B.n <= 0) return;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < B.n; idx += gridDim.x * blockDim.x) {
    const int ri = idx / B.w, ci = idx - ri * B.w;
    if (ri >= B.h - 1 || ci >= B.w - 1) continue;
    const size_t i = (size_t)B.pt_off + idx;
    const float depth = crop[i].z, depthR = crop[i + 1].z, depthD = crop[i + B.w].z;
    const float lim = (O.max_depth_change_factor * (fabsf(depth) + 1.0f) * 2.0f);
    if (fabsf(depth - depthR) > lim || !fin(depth) || !fin(depthR)) {
      change[i] = 0;
      change[i + 1] = 0;
    }
    if (fabsf(depth - depthD) > lim || !fin(depth) || !fin(depthD)) {
      change[i] = 0;
      change[i + B.w] = 0;
    }
  }
}

// ---- chamfer distance map: the two raster sweeps as skewed wavefronts, one CTA per crop ---------------------------------
__global__ void __launch_bounds__(ORG_THREADS) k_org_distance(const OrgBox* __restrict__ boxes, const unsigned char* __restrict__ change,
                                                              float* __restrict__ dist) {
  const OrgBox B = boxes[blockIdx.x];
  if (B.n <= 0) return;
  const int w = B.w, h = B.h;
  float* D = dist + B.pt_off;
  for (int i = threadIdx.x; i < B.n; i += blockDim.x) D[i] = change[B.pt_off + i] == 0 ? 0.0f : (float)(w + h);
  __syncthreads();
  // first pass: ri = 1..h-1, ci = 1..w-1; t = ci + 2 ri
  for (int t = 3; t <= (w - 1) + 2 * (h - 1); ++t) {
    // ri in [max(1, ceil((t - (w-1)) / 2)), min(h-1, (t-1)/2)]
    const int r_lo = max(1, (t - (w - 1) + 1) / 2), r_hi = min(h - 1, (t - 1) / 2);
    for (int ri = r_lo + threadIdx.x; ri <= r_hi; ri += blockDim.x) {
      const int ci = t - 2 * ri;
      float* prev = D + (size_t)(ri - 1) * w;
      float* cur = D + (size_t)ri * w;
      const float upLeft = prev[ci - 1] + 1.4f;
      const float up = prev[ci] + 1.0f;
      const float upRight = (ci + 1 < w ? prev[ci + 1] : cur[0]) + 1.4f;
      const float left = cur[ci - 1] + 1.0f;
      const float mv = fminf(fminf(upLeft, up), fminf(left, upRight));
      if (mv < cur[ci]) cur[ci] = mv;
    }
    __syncthreads();
  }
  // second pass: ri = h-2..0, ci = w-2..0; t' = (w-1-ci) + 2 (h-1-ri)
  for (int t = 3; t <= (w - 1) + 2 * (h - 1); ++t) {
    const int q_lo = max(1, (t - (w - 1) + 1) / 2), q_hi = min(h - 1, (t - 1) / 2);   // q = h-1-ri
    for (int q = q_lo + threadIdx.x; q <= q_hi; q += blockDim.x) {
      const int ri = h - 1 - q, ci = (w - 1) - (t - 2 * q);
      float* next = D + (size_t)(ri + 1) * w;
      float* cur = D + (size_t)ri * w;
      const float lowerLeft = (ci - 1 >= 0 ? next[ci - 1] : cur[w - 1]) + 1.4f;
      const float lower = next[ci] + 1.0f;
      const float lowerRight = next[ci + 1] + 1.4f;
      const float right = cur[ci + 1] + 1.0f;
      const float mv = fminf(fminf(lowerLeft, lower), fminf(right, lowerRight));
      if (mv < cur[ci]) cur[ci] = mv;
    }
    __syncthreads();
  }
}

// ---- integral images (first order 3, second order 6 doubles, finite count) as a wavefront, one CTA per crop --------------
__global__ void __launch_bounds__(ORG_THREADS) k_org_integral(const float4* __restrict__ crop, const OrgBox* __restrict__ boxes,
                                                              double* __restrict__ ii /*[entries][9]*/, unsigned* __restrict__ ic) {
  const OrgBox B = boxes[blockIdx.x];
  if (B.n <= 0) return;
  const int w = B.w, h = B.h, W1 = w + 1;
  double* I = ii + 9 * (size_t)B.ii_off;
  unsigned* C = ic + B.ii_off;
  for (long long e = threadIdx.x; e < (long long)W1 * (h + 1); e += blockDim.x) {
    const int r = (int)(e / W1), c = (int)(e - (long long)r * W1);
    if (r == 0 || c == 0) {
      for (int k = 0; k < 9; ++k) I[9 * e + k] = 0.0;
      C[e] = 0u;
    }
  }
  __syncthreads();
  for (int t = 0; t <= (w - 1) + (h - 1); ++t) {
    const int r_lo = max(0, t - (w - 1)), r_hi = min(h - 1, t);
    for (int r = r_lo + threadIdx.x; r <= r_hi; r += blockDim.x) {
      const int c = t - r;
      const size_t prev = (size_t)r * W1, cur = (size_t)(r + 1) * W1;
      const float4 p = crop[(size_t)B.pt_off + (size_t)r * w + c];
      const bool ok = fin(p.x) && fin(p.y) && fin(p.z);
      const double x = p.x, y = p.y, z = p.z;
      const double add[9] = {x, y, z, x * x, x * y, x * z, y * y, y * z, z * z};
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        double v = I[9 * (prev + c + 1) + k] + I[9 * (cur + c) + k] - I[9 * (prev + c) + k];
        if (ok) v += add[k];
        I[9 * (cur + c + 1) + k] = v;
      }
      C[cur + c + 1] = C[prev + c + 1] + C[cur + c] - C[prev + c] + (ok ? 1u : 0u);
    }
    __syncthreads();
  }
}

// ---- normals -------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_org_normals(const float4* __restrict__ crop, const OrgBox* __restrict__ boxes, OrgOpts O,
                                                     const float* __restrict__ dist, const double* __restrict__ ii,
                                                     const unsigned* __restrict__ ic, float4* __restrict__ nrm, float* __restrict__ plane_d) {
  const OrgBox B = boxes[blockIdx.y];
  if (B.n <= 0) return;
  const int w = B.w, h = B.h, W1 = w + 1;
  const double* I = ii + 9 * (size_t)B.ii_off;
  const unsigned* C = ic + B.ii_off;
  const int border = (int)O.smoothing_size;
  const float nanv = __int_as_float(0x7fc00000);
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < B.n; idx += gridDim.x * blockDim.x) {
    const int ri = idx / w, ci = idx - ri * w;
    const size_t i = (size_t)B.pt_off + idx;
    float4 out = make_float4(nanv, nanv, nanv, nanv);
    const float4 p = crop[i];
    if (ri >= border && ri < h - border && ci >= border && ci < w - border && fin(p.z)) {
      const float smoothing = fminf(dist[i], O.smoothing_size);
      if (smoothing > 2.0f) {
        const int rw = (int)smoothing, rw2 = rw / 2;
        const int sx = ci - rw2, sy = ri - rw2;
        const size_t ul = (size_t)sy * W1 + sx, ur = ul + rw, ll = (size_t)(sy + rw) * W1 + sx, lr = ll + rw;
        const unsigned count = C[lr] + C[ul] - C[ur] - C[ll];
        if (count != 0) {
          double s[9];
#pragma unroll
          for (int k = 0; k < 9; ++k) s[k] = I[9 * lr + k] + I[9 * ul + k] - I[9 * ur + k] - I[9 * ll + k];
          const float cen[3] = {(float)s[0], (float)s[1], (float)s[2]};
          float cov[9];
          cov[0] = (float)s[3];
          cov[1] = cov[3] = (float)s[4];
          cov[2] = cov[6] = (float)s[5];
          cov[4] = (float)s[6];
          cov[5] = cov[7] = (float)s[7];
          cov[8] = (float)s[8];
          const float fc = (float)count;
          for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) cov[3 * a + b] -= (cen[a] * cen[b]) / fc;
          float ev, v[3];
          eigen33(cov, ev, v);
          const float vx = 0.0f - p.x, vy = 0.0f - p.y, vz = 0.0f - p.z;
          if ((vx * v[0] + vy * v[1] + vz * v[2]) < 0.0f) {
            v[0] = -v[0];
            v[1] = -v[1];
            v[2] = -v[2];
          }
          out = make_float4(v[0], v[1], v[2], ev > 0.0f ? fabsf(ev / (cov[0] + cov[4] + cov[8])) : 0.0f);
        }
      }
    }
    nrm[i] = out;
    plane_d[i] = p.x * out.x + p.y * out.y + p.z * out.z;   // OrganizedMultiPlaneSegmentation::segment: plane_d = p . n
  }
}

// ---- connected components: union-find, root = smallest raster index of the component --------------------------------
__device__ __forceinline__ bool plane_compare(const float4* crop, const float4* nrm, const float* plane_d, const OrgOpts& O, size_t i1,
                                              size_t i2) {
  float threshold = O.distance_threshold;
  const float z = crop[i1].z;
  threshold *= z * z;
  const float4 a = nrm[i1], b = nrm[i2];
  const float dot = a.x * b.x + a.y * b.y + a.z * b.z;
  return (fabsf(plane_d[i1] - plane_d[i2]) < threshold) && (dot > O.cos_angular);
}
__device__ __forceinline__ int uf_find(int* parent, int i) {
  int p = ((volatile int*)parent)[i];
  while (p != i) {
    i = p;
    p = ((volatile int*)parent)[i];
  }
  return i;
}
__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
  while (true) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a > b) {
      const int t = a;
      a = b;
      b = t;
    }
    const int old = atomicMin(&parent[b], a);
    if (old == b) return;
    b = old;
  }
}
__global__ void __launch_bounds__(256) k_org_cc_init(const float4* __restrict__ crop, const OrgBox* __restrict__ boxes, int* __restrict__ parent) {
  const OrgBox B = boxes[blockIdx.y];
  if (B.n <= 0) return;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < B.n; idx += gridDim.x * blockDim.x)
    parent[B.pt_off + idx] = fin(crop[(size_t)B.pt_off + idx].x) ? idx : -1;
}
__global__ void __launch_bounds__(256) k_org_cc_merge(const float4* __restrict__ crop, const OrgBox* __restrict__ boxes, OrgOpts O,
                                                      const float4* __restrict__ nrm, const float* __restrict__ plane_d, int* __restrict__ parent) {
  const OrgBox B = boxes[blockIdx.y];
  if (B.n <= 0) return;
  int* P = parent + B.pt_off;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < B.n; idx += gridDim.x * blockDim.x) {
    if (P[idx] < 0) continue;
    const int ri = idx / B.w, ci = idx - ri * B.w;
    const size_t i = (size_t)B.pt_off + idx;
    if (ci > 0 && P[idx - 1] >= 0 && plane_compare(crop, nrm, plane_d, O, i, i - 1)) uf_union(P, idx, idx - 1);
    if (ri > 0 && P[idx - B.w] >= 0 && plane_compare(crop, nrm, plane_d, O, i, i - B.w)) uf_union(P, idx, idx - B.w);
  }
}
// flatten; then, one CTA per crop: number the components by their first pixel (PCL's run table order) and count them
__global__ void __launch_bounds__(ORG_THREADS) k_org_cc_label(const OrgBox* __restrict__ boxes, int* __restrict__ parent, int* __restrict__ label,
                                                              int* __restrict__ count, int* __restrict__ n_labels) {
  __shared__ int warp_tot[ORG_THREADS / 32];
  __shared__ int carry;
  const OrgBox B = boxes[blockIdx.x];
  if (B.n <= 0) {
    if (threadIdx.x == 0) n_labels[blockIdx.x] = 0;
    return;
  }
  int* P = parent + B.pt_off;
  int* L = label + B.pt_off;
  int* Cn = count + B.pt_off;
  for (int i = threadIdx.x; i < B.n; i += blockDim.x) {
    if (P[i] >= 0) P[i] = uf_find(P, i);
    Cn[i] = 0;
  }
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  // exclusive scan of the root flags in raster order -> rank of every root; stored in L[root]
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  for (int base = 0; base < B.n; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int flag = (i < B.n && P[i] == i) ? 1 : 0;
    int v = flag;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += u;
    }
    if (lane == 31) warp_tot[wp] = v;
    __syncthreads();
    int off = carry;
    for (int k = 0; k < wp; ++k) off += warp_tot[k];
    if (flag) L[i] = off + v - 1;
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += warp_tot[k];
      carry += t;
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < B.n; i += blockDim.x) {
    if (P[i] >= 0 && P[i] != i) L[i] = L[P[i]];
    else if (P[i] < 0)
      L[i] = -1;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < B.n; i += blockDim.x)
    if (L[i] >= 0) atomicAdd(&Cn[L[i]], 1);
  if (threadIdx.x == 0) n_labels[blockIdx.x] = carry;
}

// ---- regions: candidates = labels with more than min_inliers pixels (label order); moments in raster order, float ------
__global__ void __launch_bounds__(ORG_THREADS) k_org_candidates(const OrgBox* __restrict__ boxes, OrgOpts O, const int* __restrict__ count,
                                                                const int* __restrict__ n_labels, OrgRegion* __restrict__ regions,
                                                                int* __restrict__ n_cand, int* __restrict__ l2m) {
  __shared__ int warp_tot[ORG_THREADS / 32];
  __shared__ int carry;
  const OrgBox B = boxes[blockIdx.x];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  if (B.n > 0) {
    const int* Cn = count + B.pt_off;
    int* M = l2m + B.pt_off;
    const int nl = n_labels[blockIdx.x];
    for (int i = threadIdx.x; i < B.n; i += blockDim.x) M[i] = -1;
    // stream compaction in label order (block-wide exclusive scan of the "more than min_inliers pixels" flags)
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    for (int base = 0; base < nl; base += blockDim.x) {
      const int l = base + threadIdx.x;
      const int flag = (l < nl && (unsigned)Cn[l] > (unsigned)O.min_inliers) ? 1 : 0;
      int v = flag;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
      }
      if (lane == 31) warp_tot[wp] = v;
      __syncthreads();
      int off = carry;
      for (int k = 0; k < wp; ++k) off += warp_tot[k];
      const int pos = off + v - 1;
      if (flag && pos < ORG_MAXR) {
        OrgRegion& R = regions[(size_t)blockIdx.x * ORG_MAXR + pos];
        R.label = l;
        R.keep = 0;
        R.n_inliers = Cn[l];
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        int t = 0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += warp_tot[k];
        carry += t;
      }
      __syncthreads();
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) n_cand[blockIdx.x] = min(carry, ORG_MAXR);
}
// One warp per candidate region.  The warp reads 32 consecutive pixels at a time (coalesced); the pixels of the region are then
// added ONE BY ONE in raster order, every lane carrying the same nine float accumulators (pcl::computeMeanAndCovarianceMatrix
// sums in single precision in index order, and float addition does not commute with regrouping).
__global__ void __launch_bounds__(32) k_org_moments(const float4* __restrict__ crop, const OrgBox* __restrict__ boxes, OrgOpts O,
                                                    const int* __restrict__ label, const int* __restrict__ n_cand,
                                                    OrgRegion* __restrict__ regions) {
  const int b = blockIdx.x, k = blockIdx.y, lane = threadIdx.x;
  const OrgBox B = boxes[b];
  if (B.n <= 0 || k >= n_cand[b]) return;
  OrgRegion& R = regions[(size_t)b * ORG_MAXR + k];
  const int* L = label + B.pt_off;
  const float4* Pt = crop + B.pt_off;
  const int lab = R.label;
  float a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  int cnt = 0, last = -1;
  // the next 32 pixels are requested while the current ones are being added
  int l_next = lane < B.n ? L[lane] : -1;
  float4 p_next = lane < B.n ? Pt[lane] : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int base = 0; base < B.n; base += 32) {
    const bool mine = l_next == lab && base + lane < B.n;
    const float4 p = p_next;
    const int in = base + 32 + lane;
    l_next = in < B.n ? L[in] : -1;
    p_next = in < B.n ? Pt[in] : make_float4(0.f, 0.f, 0.f, 0.f);
    const bool ok = mine && fin(p.x) && fin(p.y) && fin(p.z);
    const unsigned mm = __ballot_sync(0xffffffffu, mine);
    if (mm) last = base + 31 - __clz((int)mm);
    unsigned m = __ballot_sync(0xffffffffu, ok);
    while (m) {
      const int src = __ffs((int)m) - 1;
      m &= m - 1;
      const float x = __shfl_sync(0xffffffffu, p.x, src), y = __shfl_sync(0xffffffffu, p.y, src), z = __shfl_sync(0xffffffffu, p.z, src);
      a[0] += x * x;
      a[1] += x * y;
      a[2] += x * z;
      a[3] += y * y;
      a[4] += y * z;
      a[5] += z * z;
      a[6] += x;
      a[7] += y;
      a[8] += z;
      ++cnt;
    }
  }
  if (lane != 0) return;
  const float fc = (float)cnt;
  for (int q = 0; q < 9; ++q) a[q] /= fc;
  float cov[9];
  cov[0] = a[0] - a[6] * a[6];
  cov[1] = cov[3] = a[1] - a[6] * a[7];
  cov[2] = cov[6] = a[2] - a[6] * a[8];
  cov[4] = a[3] - a[7] * a[7];
  cov[5] = cov[7] = a[4] - a[7] * a[8];
  cov[8] = a[5] - a[8] * a[8];
  float ev, v[3];
  eigen33(cov, ev, v);
  // OrganizedMultiPlaneSegmentation::segment's orientation test (rounding noise around zero, see the oracle): Eigen's
  // SSE3 4-float dot order
  float pp[4] = {v[0], v[1], v[2], 0.0f};
  const float c4[4] = {a[6], a[7], a[8], 1.0f};
  pp[3] = -1.0f * ((pp[0] * c4[0] + pp[1] * c4[1]) + (pp[2] * c4[2] + pp[3] * c4[3]));
  const float vp[4] = {0.0f - c4[0], 0.0f - c4[1], 0.0f - c4[2], 0.0f - c4[3]};
  const float cos_theta = (vp[0] * pp[0] + vp[1] * pp[1]) + (vp[2] * pp[2] + vp[3] * pp[3]);
  if (cos_theta < 0.0f) {
    for (int q = 0; q < 4; ++q) pp[q] *= -1.0f;
    pp[3] = 0.0f;
    pp[3] = -1.0f * ((pp[0] * c4[0] + pp[1] * c4[1]) + (pp[2] * c4[2] + pp[3] * c4[3]));
  }
  const float curvature = fabsf(ev / (cov[0] + cov[4] + cov[8]));
  for (int q = 0; q < 3; ++q) R.centroid[q] = a[6 + q];
  for (int q = 0; q < 4; ++q) R.model[q] = pp[q];
  R.last_inlier = last;
  R.keep = curvature < O.maximum_curvature ? 1 : 0;
  R.contour_n = 0;
  R.area = 0.0f;
}

// ---- refinement: rows in sequence, each row in closed form ------------------------------------------------------------
// block-wide inclusive max-scan along threadIdx.x (reverse = from the right); two barriers
__device__ __forceinline__ int block_scan_max(int v, bool reverse, int* warp_buf) {
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = reverse ? __shfl_down_sync(0xffffffffu, v, o) : __shfl_up_sync(0xffffffffu, v, o);
    if (reverse ? (lane + o < 32) : (lane >= o)) v = max(v, u);
  }
  __syncthreads();
  if (lane == (reverse ? 0 : 31)) warp_buf[wp] = v;
  __syncthreads();
  int m = INT_MIN;
  if (reverse) {
    for (int k = wp + 1; k < nw; ++k) m = max(m, warp_buf[k]);
  } else {
    for (int k = 0; k < wp; ++k) m = max(m, warp_buf[k]);
  }
  return max(v, m);
}
__device__ __forceinline__ bool refine_dist_ok(const float* m, const float4 p, float thr) {
  const double d = fabs((double)(m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3]));
  return d < (double)thr;
}
__device__ __forceinline__ void note_event(unsigned long long* last_ev, int region, unsigned long long seq, int target) {
  atomicMax(&last_ev[region], (seq << 24) | (unsigned long long)(unsigned)target);
}
// lab: labels after the connected components (modified in place); l2m: label -> kept region of the crop or -1
__global__ void __launch_bounds__(ORG_THREADS) k_org_refine(const float4* __restrict__ crop, const OrgBox* __restrict__ boxes, OrgOpts O,
                                                            int* __restrict__ lab, const int* __restrict__ l2m,
                                                            const OrgRegion* __restrict__ regions, unsigned long long* __restrict__ last_ev) {
  __shared__ int wbuf[ORG_THREADS / 32];
  const OrgBox B = boxes[blockIdx.x];
  if (B.n <= 0) return;
  const int w = B.w, h = B.h;
  int* L = lab + B.pt_off;
  const int* M = l2m + B.pt_off;
  const float4* Pt = crop + B.pt_off;
  const OrgRegion* R = regions + (size_t)blockIdx.x * ORG_MAXR;
  unsigned long long* EV = last_ev + (size_t)blockIdx.x * ORG_MAXR;
  const unsigned long long N2 = 2ull * (unsigned long long)B.n;
  const int c = threadIdx.x;   // w <= ORG_THREADS is checked by the host
  // ---- first sweep: current rows 0 .. h-2, the label moves right along the row, then down
  for (int r = 0; r < h - 1; ++r) {
    const int idx = r * w + c;
    const bool in = c < w;
    const int l1 = in ? L[idx] : -1;
    const int cls = !in || l1 < 0 ? 2 : (M[l1] >= 0 ? 1 : 0);   // 0 plain, 1 seed (label of a kept region), 2 barrier
    // nearest non-plain pixel on the left
    const int stop = block_scan_max(cls != 0 ? c : -1, false, wbuf);
    __syncthreads();
    // a plain pixel whose stop is a seed: does the seed's plane accept it?
    int seed_lab = -1;
    if (in && cls == 0 && stop >= 0) {
      const int sl = L[r * w + stop];
      if (sl >= 0 && M[sl] >= 0) seed_lab = sl;
    }
    const bool bad = in && cls == 0 && (seed_lab < 0 || !refine_dist_ok(R[M[seed_lab]].model, Pt[idx], O.refine_distance));
    const int first_bad = block_scan_max(bad ? c : -1, false, wbuf);   // nearest failing plain pixel at or left of c
    __syncthreads();
    int lfin = l1;
    if (in && cls == 0 && seed_lab >= 0 && first_bad <= stop) {
      lfin = seed_lab;
      L[idx] = lfin;
      note_event(EV, M[seed_lab], 2ull * (unsigned long long)(idx - 1), idx);   // right-check of the pixel on the left
    }
    __syncthreads();
    // down-checks of the current pixels (c < w-1; skipped when the right neighbour carries no label)
    if (in && c < w - 1 && lfin >= 0 && M[lfin] >= 0 && L[idx + 1] >= 0) {
      const int ll = L[idx + w];
      if (ll >= 0 && M[ll] < 0 && refine_dist_ok(R[M[lfin]].model, Pt[idx + w], O.refine_distance)) {
        L[idx + w] = lfin;
        note_event(EV, M[lfin], 2ull * (unsigned long long)idx + 1ull, idx + w);
      }
    }
    __syncthreads();
  }
  // ---- second sweep: current rows h-1 .. 1, the label moves left along the row (and from (r,0) into (r-1,w-1)), then up
  for (int r = h - 1; r >= 1; --r) {
    const int idx = r * w + c;
    const bool in = c < w;
    const int l1 = in ? L[idx] : -1;
    const int cls = !in || l1 < 0 ? 2 : (M[l1] >= 0 ? 1 : 0);
    // nearest non-plain pixel on the right (as -c so that a max-scan finds the nearest)
    const int stop_n = block_scan_max(cls != 0 ? -c : INT_MIN, true, wbuf);
    __syncthreads();
    const int stop = stop_n == INT_MIN ? -1 : -stop_n;
    int seed_lab = -1;
    if (in && cls == 0 && stop >= 0 && stop < w) {
      const int sl = L[r * w + stop];
      if (sl >= 0 && M[sl] >= 0) seed_lab = sl;
    }
    const bool bad = in && cls == 0 && (seed_lab < 0 || !refine_dist_ok(R[M[seed_lab]].model, Pt[idx], O.refine_distance));
    const int bad_n = block_scan_max(bad ? -c : INT_MIN, true, wbuf);   // nearest failing plain pixel at or right of c
    __syncthreads();
    const int first_bad = bad_n == INT_MIN ? INT_MAX : -bad_n;
    int lfin = l1;
    if (in && cls == 0 && seed_lab >= 0 && first_bad >= stop) {
      lfin = seed_lab;
      L[idx] = lfin;
      note_event(EV, M[seed_lab], N2 + 2ull * (unsigned long long)(B.n - 1 - (idx + 1)), idx);   // left-check of the pixel on the right
    }
    __syncthreads();
    // up-checks (skipped when the pixel before it in memory — (r,c-1), or (r-1,w-1) for c = 0 — carries no label)
    if (in && lfin >= 0 && M[lfin] >= 0 && L[idx - 1] >= 0) {
      const int ul = L[idx - w];
      if (ul >= 0 && M[ul] < 0 && refine_dist_ok(R[M[lfin]].model, Pt[idx - w], O.refine_distance)) {
        L[idx - w] = lfin;
        note_event(EV, M[lfin], N2 + 2ull * (unsigned long long)(B.n - 1 - idx) + 1ull, idx - w);
      }
    }
    __syncthreads();
    // the left-check of (r,0) reaches the last pixel of the row above (PCL indexes current_row + colIdx - 1 with colIdx = 0)
    if (c == 0) {
      const int cl = L[idx], tl = L[idx - 1];
      if (cl >= 0 && tl >= 0 && M[cl] >= 0 && M[tl] < 0 && refine_dist_ok(R[M[cl]].model, Pt[idx - 1], O.refine_distance)) {
        L[idx - 1] = cl;
        note_event(EV, M[cl], N2 + 2ull * (unsigned long long)(B.n - 1 - idx), idx - 1);
      }
    }
    __syncthreads();
  }
}

// ---- boundary (findLabeledRegionBoundary from the last inlier) + calculatePolygonArea, one thread per region -------------
// One CTA per (crop, region): the threads count the inliers; thread 0 then walks the contour.
__global__ void __launch_bounds__(256) k_org_boundary(const float4* __restrict__ crop, const OrgBox* __restrict__ boxes, const int* __restrict__ lab,
                                                      const int* __restrict__ n_reg, const unsigned long long* __restrict__ last_ev,
                                                      OrgRegion* __restrict__ regions) {
  __shared__ int s_cnt;
  const int b = blockIdx.x, k = blockIdx.y;
  const OrgBox B = boxes[b];
  if (B.n <= 0 || k >= n_reg[b]) return;
  OrgRegion& R = regions[(size_t)b * ORG_MAXR + k];
  const int w = B.w, h = B.h;
  const int* L = lab + B.pt_off;
  const float4* Pt = crop + B.pt_off;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  int c = 0;
  for (int i = threadIdx.x; i < B.n; i += blockDim.x) c += L[i] == R.label ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_cnt, c);
  __syncthreads();
  if (threadIdx.x != 0) return;
  R.n_inliers = s_cnt;   // inliers after the refinement
  const unsigned long long ev = last_ev[(size_t)b * ORG_MAXR + k];
  const int start = ev ? (int)(ev & 0xffffffull) : R.last_inlier;
  R.last_inlier = start;
  const int dx[8] = {-1, -1, 0, 1, 1, 1, 0, -1}, dy[8] = {0, -1, -1, -1, 0, 1, 1, 1};
  int cur = start, cx = start % w, cy = start / w;
  const int label = L[start];
  // labels of the 8 neighbours of (cx, cy), requested together; out of the image: "another label" for the first test,
  // "not the region" for the walk (PCL tests the bounds before the label in both)
  auto neigh = [&](int* nl, bool* inb) {
#pragma unroll
    for (int d = 0; d < 8; ++d) {
      const int x = cx + dx[d], y = cy + dy[d];
      inb[d] = x >= 0 && x < w && y >= 0 && y < h;
      nl[d] = inb[d] ? L[cur + dy[d] * w + dx[d]] : -2;
    }
  };
  int nl[8];
  bool inb[8];
  neigh(nl, inb);
  int direction = -1;
  for (int d = 0; d < 8; ++d)
    if (inb[d] && nl[d] != label) {
      direction = d;
      break;
    }
  if (direction == -1) {
    R.contour_n = 0;
    R.area = 0.0f;
    return;
  }
  int n = 1;
  float res[3] = {0.0f, 0.0f, 0.0f};
  float4 first = Pt[start], prev = first;
  const long long guard = 8ll * B.n + 16;
  do {
    int nd = 0;
    for (int d = 1; d <= 8; ++d) {
      nd = (direction + d) & 7;
      if (inb[nd] && nl[nd] == label) break;
    }
    direction = (nd + 4) & 7;
    cur += dy[nd] * w + dx[nd];
    cx += dx[nd];
    cy += dy[nd];
    ++n;
    const float4 p = Pt[cur];
    neigh(nl, inb);
    // polygon[i] x polygon[i+1], accumulated in order
    const float a3[3] = {prev.x, prev.y, prev.z}, b3[3] = {p.x, p.y, p.z};
    float cr[3];
    cross3f(a3, b3, cr);
    res[0] += cr[0];
    res[1] += cr[1];
    res[2] += cr[2];
    prev = p;
  } while (cur != start && n < guard);
  // the last vertex (a repeat of the first) closes with polygon[0]: (i+1) % n
  {
    const float a3[3] = {prev.x, prev.y, prev.z}, b3[3] = {first.x, first.y, first.z};
    float cr[3];
    cross3f(a3, b3, cr);
    res[0] += cr[0];
    res[1] += cr[1];
    res[2] += cr[2];
  }
  R.contour_n = n;
  R.area = (float)((double)sqrtf(res[0] * res[0] + res[1] * res[1] + res[2] * res[2]) * 0.5);
}

// keep the regions that passed the curvature gate, in label order; build label -> region
__global__ void __launch_bounds__(64) k_org_select(const OrgBox* __restrict__ boxes, const int* __restrict__ n_cand, OrgRegion* __restrict__ regions,
                                                   int* __restrict__ n_reg, int* __restrict__ l2m, unsigned long long* __restrict__ last_ev) {
  const int b = blockIdx.x;
  if (threadIdx.x != 0) return;
  const OrgBox B = boxes[b];
  int n = 0;
  if (B.n > 0) {
    OrgRegion* R = regions + (size_t)b * ORG_MAXR;
    for (int k = 0; k < n_cand[b]; ++k)
      if (R[k].keep) {
        if (n != k) R[n] = R[k];
        l2m[B.pt_off + R[n].label] = n;
        last_ev[(size_t)b * ORG_MAXR + n] = 0ull;
        ++n;
      }
  }
  n_reg[b] = n;
}

}  // namespace ssb_org
