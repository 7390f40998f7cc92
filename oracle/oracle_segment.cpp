// oracle_segment.cpp — TEST INFRASTRUCTURE ONLY (never linked into or called by the product).
//
// CPU restatement of the LIVE segmentation path of the reference:
//   plane_segmentation::computeNormalsFromPointCloud   src/planar_segmentation/plane_segmentation.cpp:84-106
//       -> pcl::IntegralImageNormalEstimation  (COVARIANCE_MATRIX, setMaxDepthChangeFactor(0.03f),
//          setNormalSmoothingSize(20.0f), BORDER_POLICY_IGNORE, no depth-dependent smoothing)
//   plane_segmentation::multiPlaneSegmentation          :108-259, the PCL call at :136-156
//       -> pcl::OrganizedMultiPlaneSegmentation::segmentAndRefine  (setMinInliers, setAngularThreshold(2 deg),
//          setDistanceThreshold(0.02), default maximum curvature 0.001, PlaneCoefficientComparator with the
//          depth-dependent distance threshold, OrganizedConnectedComponentSegmentation, PlaneRefinementComparator
//          with its default 0.02 m threshold, findLabeledRegionBoundary from the LAST inlier, the region's centroid /
//          covariance from BEFORE the refinement), pcl::calculatePolygonArea  (:189)
// PCL is not vendored under /root/reference and is absent from this image, so this follows the published PCL 1.8
// sources (features/impl/integral_image_normal.hpp, features/impl/integral_image2D.hpp, common/impl/eigen.hpp,
// common/impl/centroid.hpp, segmentation/impl/organized_multi_plane_segmentation.hpp,
// segmentation/impl/organized_connected_component_segmentation.hpp, segmentation/plane_coefficient_comparator.h,
// segmentation/plane_refinement_comparator.h, geometry/polygon_operations.h) from memory: PARITY UNPINNED — nothing
// independent of the builder confirms it.  Sequential, raster order, single precision where PCL is.
// One documented deviation: sinf / cosf / atan2f / sqrtf inside eigen33 are evaluated in double and rounded to
// float (libm's float functions differ between glibc and CUDA by an ulp; the double ones agree after rounding).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace {

struct V3 {
  float x, y, z;
};

inline bool fin(float v) { return std::isfinite(v); }

// ---- pcl::eigen33 (smallest eigenvalue and its eigenvector of a symmetric 3x3, float) -------------------------------
inline void compute_roots2(float b, float c, float* r) {
  r[0] = 0.0f;
  float d = (float)(b * b - 4.0 * c);   // PCL: Scalar d = Scalar (b * b - 4.0 * c): the product is float, the rest double
  if (d < 0.0f) d = 0.0f;
  const float sd = (float)std::sqrt((double)d);
  r[2] = 0.5f * (b + sd);
  r[1] = 0.5f * (b - sd);
}
inline void compute_roots(const float* m /*row-major 3x3*/, float* r) {
  const float m00 = m[0], m01 = m[1], m02 = m[2], m11 = m[4], m12 = m[5], m22 = m[8];
  const float c0 = m00 * m11 * m22 + 2.0f * m01 * m02 * m12 - m00 * m12 * m12 - m11 * m02 * m02 - m22 * m01 * m01;
  const float c1 = m00 * m11 - m01 * m01 + m00 * m22 - m02 * m02 + m11 * m22 - m12 * m12;
  const float c2 = m00 + m11 + m22;
  if (std::fabs(c0) < std::numeric_limits<float>::epsilon()) {
    compute_roots2(c2, c1, r);
    return;
  }
  const float s_inv3 = (float)(1.0 / 3.0);
  const float s_sqrt3 = (float)std::sqrt(3.0);
  const float c2_over_3 = c2 * s_inv3;
  float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
  if (a_over_3 > 0.0f) a_over_3 = 0.0f;
  const float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
  float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
  if (q > 0.0f) q = 0.0f;
  const float rho = (float)std::sqrt((double)(-a_over_3));
  const float theta = (float)std::atan2((double)(float)std::sqrt((double)(-q)), (double)half_b) * s_inv3;
  const float cos_theta = (float)std::cos((double)theta);
  const float sin_theta = (float)std::sin((double)theta);
  r[0] = c2_over_3 + 2.0f * rho * cos_theta;
  r[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
  r[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
  if (r[0] >= r[1]) std::swap(r[0], r[1]);
  if (r[1] >= r[2]) {
    std::swap(r[1], r[2]);
    if (r[0] >= r[1]) std::swap(r[0], r[1]);
  }
  if (r[0] <= 0.0f) compute_roots2(c2, c1, r);
}
inline void cross(const float* a, const float* b, float* o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
inline void eigen33(const float* mat, float& eigenvalue, float* evec) {
  float scale = 0.0f;
  for (int k = 0; k < 9; ++k) scale = std::max(scale, std::fabs(mat[k]));
  if (scale <= std::numeric_limits<float>::min()) scale = 1.0f;
  float s[9];
  for (int k = 0; k < 9; ++k) s[k] = mat[k] / scale;
  float roots[3];
  compute_roots(s, roots);
  eigenvalue = roots[0] * scale;
  s[0] -= roots[0];
  s[4] -= roots[0];
  s[8] -= roots[0];
  float v1[3], v2[3], v3[3];
  cross(s + 0, s + 3, v1);
  cross(s + 0, s + 6, v2);
  cross(s + 3, s + 6, v3);
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
  const float n = (float)std::sqrt((double)l);
  for (int k = 0; k < 3; ++k) evec[k] = v[k] / n;
}

// ---- IntegralImageNormalEstimation::computeFeature -------------------------------------------------------------------
void distance_map(const float* c4, int w, int h, float max_depth_change_factor, std::vector<float>& dist) {
  const size_t n = (size_t)w * h;
  std::vector<unsigned char> change(n, 255);
  for (int ri = 0; ri < h - 1; ++ri)
    for (int ci = 0; ci < w - 1; ++ci) {
      const size_t index = (size_t)ri * w + ci;
      const float depth = c4[4 * index + 2], depthR = c4[4 * (index + 1) + 2], depthD = c4[4 * (index + w) + 2];
      const float lim = (max_depth_change_factor * (std::fabs(depth) + 1.0f) * 2.0f);
      if (std::fabs(depth - depthR) > lim || !fin(depth) || !fin(depthR)) {
        change[index] = 0;
        change[index + 1] = 0;
      }
      if (std::fabs(depth - depthD) > lim || !fin(depth) || !fin(depthD)) {
        change[index] = 0;
        change[index + w] = 0;
      }
    }
  dist.resize(n);
  for (size_t i = 0; i < n; ++i) dist[i] = change[i] == 0 ? 0.0f : (float)(w + h);
  // first pass
  for (int ri = 1; ri < h; ++ri) {
    float* prev = dist.data() + (size_t)(ri - 1) * w;
    float* cur = dist.data() + (size_t)ri * w;
    for (int ci = 1; ci < w; ++ci) {
      const float upLeft = prev[ci - 1] + 1.4f;
      const float up = prev[ci] + 1.0f;
      // (at the last column PCL reads previous_row[width], one past the row: the first pixel of the current row)
      const float upRight = (ci + 1 < w ? prev[ci + 1] : cur[0]) + 1.4f;
      const float left = cur[ci - 1] + 1.0f;
      const float minValue = std::min(std::min(upLeft, up), std::min(left, upRight));
      if (minValue < cur[ci]) cur[ci] = minValue;
    }
  }
  // second pass
  for (int ri = h - 2; ri >= 0; --ri) {
    float* next = dist.data() + (size_t)(ri + 1) * w;
    float* cur = dist.data() + (size_t)ri * w;
    for (int ci = w - 2; ci >= 0; --ci) {
      const float lowerLeft = (ci - 1 >= 0 ? next[ci - 1] : cur[w - 1]) + 1.4f;   // (PCL reads one before the row start: the last pixel of the current row)
      const float lower = next[ci] + 1.0f;
      const float lowerRight = next[ci + 1] + 1.4f;
      const float right = cur[ci + 1] + 1.0f;
      const float minValue = std::min(std::min(lowerLeft, lower), std::min(right, lowerRight));
      if (minValue < cur[ci]) cur[ci] = minValue;
    }
  }
}

struct Integral {   // pcl::IntegralImage2D<float, 3> with second-order sums, double accumulators, (w+1) x (h+1)
  int w, h;
  std::vector<double> first;   // [(h+1)(w+1)][3]
  std::vector<double> second;  // [(h+1)(w+1)][6]
  std::vector<unsigned> cnt;   // [(h+1)(w+1)]
  void build(const float* c4, int w_, int h_) {
    w = w_;
    h = h_;
    const size_t n = (size_t)(w + 1) * (h + 1);
    first.assign(3 * n, 0.0);
    second.assign(6 * n, 0.0);
    cnt.assign(n, 0u);
    for (int r = 0; r < h; ++r) {
      const size_t prev = (size_t)r * (w + 1), cur = (size_t)(r + 1) * (w + 1);
      for (int c = 0; c < w; ++c) {
        for (int k = 0; k < 3; ++k) first[3 * (cur + c + 1) + k] = first[3 * (prev + c + 1) + k] + first[3 * (cur + c) + k] - first[3 * (prev + c) + k];
        for (int k = 0; k < 6; ++k) second[6 * (cur + c + 1) + k] = second[6 * (prev + c + 1) + k] + second[6 * (cur + c) + k] - second[6 * (prev + c) + k];
        cnt[cur + c + 1] = cnt[prev + c + 1] + cnt[cur + c] - cnt[prev + c];
        const float* p = c4 + 4 * ((size_t)r * w + c);
        if (fin(p[0]) && fin(p[1]) && fin(p[2])) {
          const double x = p[0], y = p[1], z = p[2];
          first[3 * (cur + c + 1) + 0] += x;
          first[3 * (cur + c + 1) + 1] += y;
          first[3 * (cur + c + 1) + 2] += z;
          double* so = &second[6 * (cur + c + 1)];
          so[0] += x * x;
          so[1] += x * y;
          so[2] += x * z;
          so[3] += y * y;
          so[4] += y * z;
          so[5] += z * z;
          ++cnt[cur + c + 1];
        }
      }
    }
  }
  template <int K>
  void sum(const std::vector<double>& img, int sx, int sy, int ww, int hh, double* out) const {
    const size_t ul = (size_t)sy * (w + 1) + sx, ur = ul + ww, ll = (size_t)(sy + hh) * (w + 1) + sx, lr = ll + ww;
    for (int k = 0; k < K; ++k) out[k] = img[K * lr + k] + img[K * ul + k] - img[K * ur + k] - img[K * ll + k];
  }
  unsigned count(int sx, int sy, int ww, int hh) const {
    const size_t ul = (size_t)sy * (w + 1) + sx, ur = ul + ww, ll = (size_t)(sy + hh) * (w + 1) + sx, lr = ll + ww;
    return cnt[lr] + cnt[ul] - cnt[ur] - cnt[ll];
  }
};

void normals(const float* c4, int w, int h, float max_depth_change_factor, float smoothing_size, float* nrm /*[n][4]*/,
             std::vector<float>& dist) {
  const float nanv = std::numeric_limits<float>::quiet_NaN();
  const size_t n = (size_t)w * h;
  for (size_t i = 0; i < 4 * n; ++i) nrm[i] = nanv;
  distance_map(c4, w, h, max_depth_change_factor, dist);
  Integral I;
  I.build(c4, w, h);
  const int border = (int)smoothing_size;
  for (int ri = border; ri < h - border; ++ri)
    for (int ci = border; ci < w - border; ++ci) {
      const size_t index = (size_t)ri * w + ci;
      const float depth = c4[4 * index + 2];
      if (!fin(depth)) continue;
      const float smoothing = std::min(dist[index], smoothing_size);
      if (!(smoothing > 2.0f)) continue;
      const int rw = (int)smoothing, rh = (int)smoothing, rw2 = rw / 2, rh2 = rh / 2;
      const unsigned count = I.count(ci - rw2, ri - rh2, rw, rh);
      if (count == 0) continue;
      double fo[3], so[6];
      I.sum<3>(I.first, ci - rw2, ri - rh2, rw, rh, fo);
      I.sum<6>(I.second, ci - rw2, ri - rh2, rw, rh, so);
      const float cx = (float)fo[0], cy = (float)fo[1], cz = (float)fo[2];
      float cov[9];
      cov[0] = (float)so[0];
      cov[1] = cov[3] = (float)so[1];
      cov[2] = cov[6] = (float)so[2];
      cov[4] = (float)so[3];
      cov[5] = cov[7] = (float)so[4];
      cov[8] = (float)so[5];
      const float fc = (float)count;
      const float cen[3] = {cx, cy, cz};
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) cov[3 * a + b] -= (cen[a] * cen[b]) / fc;
      float ev, v[3];
      eigen33(cov, ev, v);
      // flipNormalTowardsViewpoint, viewpoint = origin
      const float* p = c4 + 4 * index;
      const float vx = 0.0f - p[0], vy = 0.0f - p[1], vz = 0.0f - p[2];
      if ((vx * v[0] + vy * v[1] + vz * v[2]) < 0.0f) {
        v[0] = -v[0];
        v[1] = -v[1];
        v[2] = -v[2];
      }
      nrm[4 * index + 0] = v[0];
      nrm[4 * index + 1] = v[1];
      nrm[4 * index + 2] = v[2];
      nrm[4 * index + 3] = ev > 0.0f ? std::fabs(ev / (cov[0] + cov[4] + cov[8])) : 0.0f;
    }
}

// ---- OrganizedConnectedComponentSegmentation with PlaneCoefficientComparator ---------------------------------------
struct PlaneCmp {
  const float* c4;
  const float* nrm;
  const float* plane_d;
  float cos_ang, dist_thr;
  bool operator()(size_t i1, size_t i2) const {
    float threshold = dist_thr;
    const float z = c4[4 * i1 + 2];   // vec.dot(z_axis_), z_axis_ = (0, 0, 1)
    threshold *= z * z;
    const float dot = nrm[4 * i1] * nrm[4 * i2] + nrm[4 * i1 + 1] * nrm[4 * i2 + 1] + nrm[4 * i1 + 2] * nrm[4 * i2 + 2];
    return (std::fabs(plane_d[i1] - plane_d[i2]) < threshold) && (dot > cos_ang);
  }
};
unsigned find_root(const std::vector<unsigned>& runs, unsigned idx) {
  while (runs[idx] != idx) idx = runs[idx];
  return idx;
}
int connected_components(const float* c4, int w, int h, const PlaneCmp& cmp, std::vector<int>& labels) {
  const size_t n = (size_t)w * h;
  labels.assign(n, -1);
  std::vector<unsigned> run_ids;
  unsigned clust = 0;
  auto fresh = [&](size_t i) {
    labels[i] = (int)clust++;
    run_ids.push_back((unsigned)labels[i]);
  };
  if (fin(c4[0])) fresh(0);
  for (int c = 1; c < w; ++c) {
    if (!fin(c4[4 * (size_t)c])) continue;
    if (cmp(c, c - 1))
      labels[c] = labels[c - 1];
    else
      fresh(c);
  }
  for (int r = 1; r < h; ++r) {
    const size_t cur = (size_t)r * w, prev = cur - w;
    if (fin(c4[4 * cur])) {
      if (cmp(cur, prev))
        labels[cur] = labels[prev];
      else
        fresh(cur);
    }
    for (int c = 1; c < w; ++c) {
      const size_t i = cur + c;
      if (!fin(c4[4 * i])) continue;
      if (cmp(i, i - 1)) labels[i] = labels[i - 1];
      if (cmp(i, prev + c)) {
        if (labels[i] < 0)
          labels[i] = labels[prev + c];
        else if (labels[prev + c] >= 0) {
          const unsigned r1 = find_root(run_ids, (unsigned)labels[i]), r2 = find_root(run_ids, (unsigned)labels[prev + c]);
          if (r1 < r2)
            run_ids[r2] = r1;
          else
            run_ids[r1] = r2;
        }
      }
      if (labels[i] < 0) fresh(i);
    }
  }
  std::vector<unsigned> map(clust);
  unsigned max_id = 0;
  for (unsigned k = 0; k < run_ids.size(); ++k) {
    if (run_ids[k] == k)
      map[k] = max_id++;
    else
      map[k] = map[find_root(run_ids, k)];
  }
  for (size_t i = 0; i < n; ++i)
    if (labels[i] >= 0) labels[i] = (int)map[labels[i]];
  return (int)max_id;
}

// pcl::computeMeanAndCovarianceMatrix (indices version, float accumulators, raster order)
void mean_cov(const float* c4, const std::vector<int>& idx, float* cov, float* cen) {
  float a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  size_t cnt = 0;
  for (int i : idx) {
    const float* p = c4 + 4 * (size_t)i;
    if (!fin(p[0]) || !fin(p[1]) || !fin(p[2])) continue;
    a[0] += p[0] * p[0];
    a[1] += p[0] * p[1];
    a[2] += p[0] * p[2];
    a[3] += p[1] * p[1];
    a[4] += p[1] * p[2];
    a[5] += p[2] * p[2];
    a[6] += p[0];
    a[7] += p[1];
    a[8] += p[2];
    ++cnt;
  }
  const float fc = (float)cnt;
  for (int k = 0; k < 9; ++k) a[k] /= fc;
  cen[0] = a[6];
  cen[1] = a[7];
  cen[2] = a[8];
  cov[0] = a[0] - a[6] * a[6];
  cov[1] = cov[3] = a[1] - a[6] * a[7];
  cov[2] = cov[6] = a[2] - a[6] * a[8];
  cov[4] = a[3] - a[7] * a[7];
  cov[5] = cov[7] = a[4] - a[7] * a[8];
  cov[8] = a[5] - a[8] * a[8];
}

struct Region {
  float centroid[3], model[4];
  std::vector<int> inliers;
  int label;
};

// OrganizedMultiPlaneSegmentation::refine (PlaneRefinementComparator, threshold 0.02 m, not depth dependent)
void refine(const float* c4, int w, int h, std::vector<int>& labels, std::vector<Region>& R, int n_labels) {
  std::vector<char> grow(std::max(n_labels, 1), 0);
  std::vector<int> l2m(std::max(n_labels, 1), 0);
  for (size_t i = 0; i < R.size(); ++i) {
    const int ml = labels[R[i].inliers[0]];
    l2m[ml] = (int)i;
    grow[ml] = 1;
  }
  auto cmp = [&](size_t i1, size_t i2) {
    const int cl = labels[i1], nl = labels[i2];
    if (!(grow[cl] && !grow[nl])) return false;
    const float* m = R[l2m[cl]].model;
    const float* p = c4 + 4 * i2;
    const double d = std::fabs(m[0] * p[0] + m[1] * p[1] + m[2] * p[2] + m[3]);   // float arithmetic, as in PCL
    return d < 0.02f;
  };
  for (int r = 0; r < h - 1; ++r) {
    const size_t cur = (size_t)r * w, next = cur + w;
    for (int c = 0; c < w - 1; ++c) {
      const int cl = labels[cur + c], rl = labels[cur + c + 1];
      if (cl < 0 || rl < 0) continue;
      if (cmp(cur + c, cur + c + 1)) {
        labels[cur + c + 1] = cl;
        R[l2m[cl]].inliers.push_back((int)(cur + c + 1));
      }
      const int ll = labels[next + c];
      if (ll < 0) continue;
      if (cmp(cur + c, next + c)) {
        labels[next + c] = cl;
        R[l2m[cl]].inliers.push_back((int)(next + c));
      }
    }
  }
  for (int r = h - 1; r >= 1; --r) {
    const size_t cur = (size_t)r * w, prev = cur - w;
    for (int c = w - 1; c >= 0; --c) {
      const int cl = labels[cur + c], ll = labels[cur + c - 1];   // (c == 0: the last pixel of the previous row, as in PCL)
      if (cl < 0 || ll < 0) continue;
      if (cmp(cur + c, cur + c - 1)) {
        labels[cur + c - 1] = cl;
        R[l2m[cl]].inliers.push_back((int)(cur + c - 1));
      }
      const int ul = labels[prev + c];
      if (ul < 0) continue;
      if (cmp(cur + c, prev + c)) {
        labels[prev + c] = cl;
        R[l2m[cl]].inliers.push_back((int)(prev + c));
      }
    }
  }
}

// OrganizedConnectedComponentSegmentation::findLabeledRegionBoundary
void boundary(int start, const std::vector<int>& labels, int w, int h, std::vector<int>& out) {
  out.clear();
  const int dx[8] = {-1, -1, 0, 1, 1, 1, 0, -1}, dy[8] = {0, -1, -1, -1, 0, 1, 1, 1};
  int di[8];
  for (int k = 0; k < 8; ++k) di[k] = dy[k] * w + dx[k];
  int cur = start, cx = start % w, cy = start / w;
  const int label = labels[start];
  int direction = -1;
  for (int d = 0; d < 8; ++d) {
    const int x = cx + dx[d], y = cy + dy[d], idx = cur + di[d];
    if (x >= 0 && x < w && y >= 0 && y < h && labels[idx] != label) {
      direction = d;
      break;
    }
  }
  if (direction == -1) return;
  out.push_back(start);
  const size_t guard = (size_t)8 * w * h + 16;
  do {
    int n = 0;
    for (int d = 1; d <= 8; ++d) {
      n = (direction + d) & 7;
      const int x = cx + dx[n], y = cy + dy[n], idx = cur + di[n];
      if (x >= 0 && x < w && y >= 0 && y < h && labels[idx] == label) break;
    }
    direction = (n + 4) & 7;
    cur += di[n];
    cx += dx[n];
    cy += dy[n];
    out.push_back(cur);
  } while (cur != start && out.size() < guard);
}

float polygon_area(const float* c4, const std::vector<int>& poly) {
  float res[3] = {0, 0, 0};
  const int n = (int)poly.size();
  for (int i = 0; i < n; ++i) {
    const int j = (i + 1) % n;
    const float* a = c4 + 4 * (size_t)poly[i];
    const float* b = c4 + 4 * (size_t)poly[j];
    float c[3];
    cross(a, b, c);
    res[0] += c[0];
    res[1] += c[1];
    res[2] += c[2];
  }
  const float area = std::sqrt(res[0] * res[0] + res[1] * res[1] + res[2] * res[2]);
  return (float)(area * 0.5);
}

}  // namespace

extern "C" {

// integral-image normals only (row b5).  cloud4: [h*w][4] x,y,z,rgb organised row-major; normals_out: [h*w][4] (nx, ny,
// nz, curvature), NaN where PCL leaves NaN; dist_out (may be null): the smoothing distance map.
int orc_integral_normals(const float* cloud4, int w, int h, float max_depth_change_factor, float smoothing_size, float* normals_out,
                         float* dist_out) {
  std::vector<float> dist;
  normals(cloud4, w, h, max_depth_change_factor, smoothing_size, normals_out, dist);
  if (dist_out) std::memcpy(dist_out, dist.data(), dist.size() * sizeof(float));
  return 0;
}

// normals + segmentAndRefine (rows b4 / f1).  labels_cc / labels_ref (may be null): labels after the connected components
// and after the refinement (-1 = no label).  Regions (at most max_regions): centroid[3], model[4], inlier count (after
// refine), contour point count, polygon area of the contour.  Returns the number of regions found (may exceed
// max_regions: only the first max_regions are written), or < 0 on error.
int orc_organized_planes(const float* cloud4, int w, int h, float max_depth_change_factor, float smoothing_size, int min_inliers,
                         float angular_threshold, float distance_threshold, float maximum_curvature, float* normals_out, int* labels_cc,
                         int* labels_ref, int max_regions, float* centroid3, float* model4, int* n_inliers, int* contour_n, float* area) {
  const size_t n = (size_t)w * h;
  std::vector<float> nrm(4 * n), dist;
  normals(cloud4, w, h, max_depth_change_factor, smoothing_size, nrm.data(), dist);
  if (normals_out) std::memcpy(normals_out, nrm.data(), 4 * n * sizeof(float));
  std::vector<float> plane_d(n);
  for (size_t i = 0; i < n; ++i)
    plane_d[i] = cloud4[4 * i] * nrm[4 * i] + cloud4[4 * i + 1] * nrm[4 * i + 1] + cloud4[4 * i + 2] * nrm[4 * i + 2];
  PlaneCmp cmp{cloud4, nrm.data(), plane_d.data(), (float)std::cos((double)angular_threshold), distance_threshold};
  std::vector<int> labels;
  const int n_labels = connected_components(cloud4, w, h, cmp, labels);
  if (labels_cc) std::memcpy(labels_cc, labels.data(), n * sizeof(int));
  std::vector<std::vector<int>> label_idx(std::max(n_labels, 1));
  for (size_t i = 0; i < n; ++i)
    if (labels[i] >= 0) label_idx[labels[i]].push_back((int)i);
  std::vector<Region> R;
  for (int l = 0; l < n_labels; ++l) {
    if (!((unsigned)label_idx[l].size() > (unsigned)min_inliers)) continue;
    float cov[9], cen[3], ev, v[3];
    mean_cov(cloud4, label_idx[l], cov, cen);
    eigen33(cov, ev, v);
    // plane_params = (eigenvector, 0); plane_params[3] = -plane_params.dot(clust_centroid) with clust_centroid[3] = 1;
    // vp = 0 - clust_centroid (4 components!); cos_theta = vp.dot(plane_params) — which is -n.c - d = rounding noise
    // around 0, so the orientation PCL returns is arbitrary; reproduced with Eigen's SSE3 4-float dot order
    // ((a0 b0 + a1 b1) + (a2 b2 + a3 b3)).  The reference re-orients the normals itself (plane_segmentation.cpp:210-247).
    auto dot4 = [](const float* a, const float* b) { return (a[0] * b[0] + a[1] * b[1]) + (a[2] * b[2] + a[3] * b[3]); };
    float pp[4] = {v[0], v[1], v[2], 0.0f};
    const float c4v[4] = {cen[0], cen[1], cen[2], 1.0f};
    pp[3] = -1.0f * dot4(pp, c4v);
    const float vp[4] = {0.0f - c4v[0], 0.0f - c4v[1], 0.0f - c4v[2], 0.0f - c4v[3]};
    const float cos_theta = dot4(vp, pp);
    if (cos_theta < 0.0f) {
      for (int k = 0; k < 4; ++k) pp[k] *= -1.0f;
      pp[3] = 0.0f;
      pp[3] = -1.0f * dot4(pp, c4v);
    }
    const float curvature = std::fabs(ev / (cov[0] + cov[4] + cov[8]));
    if (curvature < maximum_curvature) {
      Region q;
      for (int k = 0; k < 3; ++k) q.centroid[k] = cen[k];
      for (int k = 0; k < 4; ++k) q.model[k] = pp[k];
      q.inliers = label_idx[l];
      q.label = l;
      R.push_back(std::move(q));
    }
  }
  refine(cloud4, w, h, labels, R, n_labels);
  if (labels_ref) std::memcpy(labels_ref, labels.data(), n * sizeof(int));
  for (size_t i = 0; i < R.size() && (int)i < max_regions; ++i) {
    std::vector<int> bnd;
    boundary(R[i].inliers.back(), labels, w, h, bnd);
    for (int k = 0; k < 3; ++k) centroid3[3 * i + k] = R[i].centroid[k];
    for (int k = 0; k < 4; ++k) model4[4 * i + k] = R[i].model[k];
    n_inliers[i] = (int)R[i].inliers.size();
    contour_n[i] = (int)bnd.size();
    area[i] = polygon_area(cloud4, bnd);
  }
  return (int)R.size();
}

}  // extern "C"
