// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked, imported or called by the product path.
//
// CPU restatement of the planar_segmentation RANSAC plane fit of hridaybavle/semantic_slam:
//   bbox crop        : plane_segmentation::segmentPointCloudData  plane_segmentation.cpp:24-82
//                      (validity rule :34-38, byte gather :48-61, organised crop :69)
//   RANSAC plane fit : plane_segmentation::compute2DConvexHull     plane_segmentation.cpp:631-647
//                      pcl::SACSegmentation, SACMODEL_PLANE, SAC_RANSAC, threshold 0.01,
//                      optimizeCoefficients = true
// The arithmetic lives in PCL (>=1.7, un-vendored; CMakeLists.txt:22-23).  Restated from its
// published algorithm:
//   sample_consensus/impl/sac_model_plane.hpp : computeModelCoefficients (3-point plane,
//        collinearity test on the component ratios, cross product, normalise, d = -n.p0),
//        countWithinDistance / selectWithinDistance ( |n.p + d| < threshold ),
//        optimizeModelCoefficients (centroid + covariance of the inliers -> smallest eigenvector)
//   sample_consensus/impl/ransac.hpp : computeModel (strictly-better keeps the first best;
//        adaptive k = log(1-p)/log(1-w^3), max_iterations 50, skipped samples)
//   segmentation/impl/sac_segmentation.hpp : segment (refine, then re-select inliers)
//   common/impl/centroid.hpp computeMeanAndCovarianceMatrix, common/impl/eigen.hpp eigen33
// Float evaluation order: Eigen's SSE3 packet dot product of 4-vectors = ((a0*b0 + a1*b1) +
// (a2*b2 + a3*b3)), separate roundings, no FMA (the reference compiles with -msse..-msse4.2 only,
// CMakeLists.txt:7-8).  The inlier covariance is accumulated in double.
// PARITY UNPINNED: the reference ships no tests or golden vectors for this path.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "oracle_ransac.h"

namespace orcr {


static bool model_from_triple(const float* pts, const int* tri, float* coef) {
  const float* p0 = pts + 4 * (size_t)tri[0];
  const float* p1 = pts + 4 * (size_t)tri[1];
  const float* p2 = pts + 4 * (size_t)tri[2];
  float a[3], b[3], r[3];
  for (int i = 0; i < 3; ++i) {
    a[i] = p1[i] - p0[i];
    b[i] = p2[i] - p0[i];
    r[i] = a[i] / b[i];
  }
  if ((r[0] == r[1]) && (r[2] == r[1])) return false;  // collinear
  coef[0] = a[1] * b[2] - a[2] * b[1];
  coef[1] = a[2] * b[0] - a[0] * b[2];
  coef[2] = a[0] * b[1] - a[1] * b[0];
  coef[3] = 0.f;
  // Eigen normalize(): squaredNorm as packet reduction ((c0^2 + c1^2) + (c2^2 + c3^2)), then /= sqrt
  float n2 = (coef[0] * coef[0] + coef[1] * coef[1]) + (coef[2] * coef[2] + coef[3] * coef[3]);
  float n = std::sqrt(n2);
  coef[0] /= n;
  coef[1] /= n;
  coef[2] /= n;
  coef[3] /= n;
  // d = -1 * coef.dot(p0) with p0.w = 1 and coef[3] = 0
  float dot = (coef[0] * p0[0] + coef[1] * p0[1]) + (coef[2] * p0[2] + coef[3] * 1.0f);
  coef[3] = -1.f * dot;
  return true;
}

// PCL's sample stream (pcl/sample_consensus/sac_model.h): the model owns a boost::mt19937 seeded with 12345u (unless
// `random`), read through boost::uniform_int<>(0, INT_MAX), and drawIndexSample shuffles the head of shuffled_indices_
// (initially 0..n-1, state kept between draws):  for i in 0..2: swap(sh[i], sh[i + rnd() % (n - i)]);  sample = sh[0..2].
// The engine is written out here (Matsumoto & Nishimura; init_genrand seeding) so that the oracle shares nothing with
// the product's std::mt19937; boost's generate_uniform_int maps a 32-bit engine onto [0, INT_MAX] by dividing by the
// bucket size 2 (0xffffffff / 0x80000000 = 1, +1 because the remainder equals the range).
namespace {
struct Mt19937 {
  uint32_t mt[624];
  int idx;
  explicit Mt19937(uint32_t seed) {
    mt[0] = seed;
    for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    idx = 624;
  }
  uint32_t next() {
    if (idx >= 624) {
      for (int k = 0; k < 624; ++k) {
        uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
        mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      idx = 0;
    }
    uint32_t y = mt[idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
  }
};
}  // namespace

void pcl_sample_stream(int n, int n_draws, unsigned seed, int* triples) {
  if (n < 3) {
    for (int k = 0; k < 3 * n_draws; ++k) triples[k] = 0;
    return;
  }
  Mt19937 eng(seed);
  std::vector<int> sh(n);
  for (int i = 0; i < n; ++i) sh[i] = i;
  for (int d = 0; d < n_draws; ++d) {
    for (int i = 0; i < 3; ++i) {
      const uint32_t rnd = eng.next() / 2u;
      std::swap(sh[i], sh[i + (int)(rnd % (uint32_t)(n - i))]);
    }
    for (int i = 0; i < 3; ++i) triples[3 * d + i] = sh[i];
  }
}

static int count_within(const float* pts, int n, const float* coef, float thr_eff) {
  int c = 0;
  for (int i = 0; i < n; ++i) {
    const float* p = pts + 4 * (size_t)i;
    if (plane_dist(coef, p[0], p[1], p[2]) < thr_eff) ++c;
  }
  return c;
}

// pcl::computeRoots / eigen33 in double
static void compute_roots2(double b, double c, double* roots) {
  roots[0] = 0;
  double d = b * b - 4.0 * c;
  if (d < 0.0) d = 0.0;
  double sd = std::sqrt(d);
  roots[2] = 0.5 * (b + sd);
  roots[1] = 0.5 * (b - sd);
}
static void compute_roots(const double* m, double* roots) {
  double c0 = m[0] * m[4] * m[8] + 2.0 * m[1] * m[2] * m[5] - m[0] * m[5] * m[5] - m[4] * m[2] * m[2] -
              m[8] * m[1] * m[1];
  double c1 = m[0] * m[4] - m[1] * m[1] + m[0] * m[8] - m[2] * m[2] + m[4] * m[8] - m[5] * m[5];
  double c2 = m[0] + m[4] + m[8];
  if (std::fabs(c0) < std::numeric_limits<double>::epsilon()) {
    compute_roots2(c2, c1, roots);
    return;
  }
  const double s_inv3 = 1.0 / 3.0;
  const double s_sqrt3 = std::sqrt(3.0);
  double c2_over_3 = c2 * s_inv3;
  double a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
  if (a_over_3 > 0.0) a_over_3 = 0.0;
  double half_b = 0.5 * (c0 + c2_over_3 * (2.0 * c2_over_3 * c2_over_3 - c1));
  double q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
  if (q > 0.0) q = 0.0;
  double rho = std::sqrt(-a_over_3);
  double theta = std::atan2(std::sqrt(-q), half_b) * s_inv3;
  double cos_theta = std::cos(theta);
  double sin_theta = std::sin(theta);
  roots[0] = c2_over_3 + 2.0 * rho * cos_theta;
  roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
  roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
  if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
  if (roots[1] >= roots[2]) {
    std::swap(roots[1], roots[2]);
    if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
  }
  if (roots[0] <= 0) compute_roots2(c2, c1, roots);
}
static void eigen33_smallest(const double* mat, double* evec) {
  double scale = 0;
  for (int i = 0; i < 9; ++i) scale = std::max(scale, std::fabs(mat[i]));
  if (scale <= std::numeric_limits<double>::min()) scale = 1.0;
  double m[9];
  for (int i = 0; i < 9; ++i) m[i] = mat[i] / scale;
  double roots[3];
  compute_roots(m, roots);
  m[0] -= roots[0];
  m[4] -= roots[0];
  m[8] -= roots[0];
  auto cross = [](const double* a, const double* b, double* o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
  };
  double v1[3], v2[3], v3[3];
  cross(m + 0, m + 3, v1);
  cross(m + 0, m + 6, v2);
  cross(m + 3, m + 6, v3);
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
  double s = std::sqrt(l);
  for (int i = 0; i < 3; ++i) evec[i] = v[i] / s;
}

// SampleConsensusModelPlane::optimizeModelCoefficients over the inliers of `coef`
static void refine_plane(const float* pts, int n, const float* coef, float thr_eff, float* out, float* cen = nullptr) {
  if (cen) cen[0] = cen[1] = cen[2] = 0.f;
  double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  long long cnt = 0;
  for (int i = 0; i < n; ++i) {
    const float* p = pts + 4 * (size_t)i;
    if (!(plane_dist(coef, p[0], p[1], p[2]) < thr_eff)) continue;
    double x = p[0], y = p[1], z = p[2];
    ++cnt;
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
  if (cnt < 4) {
    std::memcpy(out, coef, 4 * sizeof(float));
    return;
  }
  for (int i = 0; i < 9; ++i) acc[i] /= (double)cnt;
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
  out[0] = (float)ev[0];
  out[1] = (float)ev[1];
  out[2] = (float)ev[2];
  // d = -n . centroid, evaluated in float like optimized_coefficients.dot(xyz_centroid) with [3]=0
  float cx = (float)acc[6], cy = (float)acc[7], cz = (float)acc[8];
  float dot = (out[0] * cx + out[1] * cy) + (out[2] * cz + 0.f * 1.0f);
  out[3] = -1.f * dot;
  if (cen) {
    cen[0] = cx;
    cen[1] = cy;
    cen[2] = cz;
  }
}

// plane_segmentation::segmentPointCloudData :24-82.  Returns n = w*h, or -1 for a "spurious" box.
// out: n x 4 floats (x, y, z, rgb), organised row-major (index = p_v * w + p_u).
static int crop(const uint8_t* msg, int width, int height, int point_step, int row_step, const int* off,
                const int* box, float* out) {
  int tl_x = box[0], tl_y = box[1], w = box[2], h = box[3];
  // reference rule (:34-35) + rejection of negative corners (the reference wraps size_t)
  if (h < 0 || w < 0 || (tl_x + w) > width || (tl_y + h) > height || tl_x < 0 || tl_y < 0) return -1;
  if (!out) return w * h;
  for (int pu = 0; pu < w; ++pu) {
    for (int pv = 0; pv < h; ++pv) {
      size_t pos = (size_t)(tl_y + pv) * row_step + (size_t)(tl_x + pu) * point_step;
      float* o = out + 4 * ((size_t)pv * w + pu);
      std::memcpy(o + 0, msg + pos + off[0], 4);
      std::memcpy(o + 1, msg + pos + off[1], 4);
      std::memcpy(o + 2, msg + pos + off[2], 4);
      std::memcpy(o + 3, msg + pos + off[3], 4);
    }
  }
  return w * h;
}

}  // namespace orcr

using namespace orcr;

// pcl::SACSegmentation (SACMODEL_PLANE, SAC_RANSAC) on one point set: hypotheses from the index triples `tri`
// (mode 0: all K scored, first best wins; mode 1: RandomSampleConsensus::computeModel's adaptive stopping rule),
// optional refine (optimizeModelCoefficients) and re-selection of the inliers.  Used by orc_ransac_batch per crop and by
// the clustering chain (oracle_cluster.cpp) per cluster.  counts_out: K ints or null; mask_out: n bytes or null.
void orc_ransac_points(const float* pts, int n, const int* tri, int K, double threshold, int refine, int mode, int max_iterations,
                       double probability, PlaneResult& R, int* counts_out, uint8_t* mask_out) {
  const float thr = effective_threshold(threshold);
  int best = 0, best_k = -1;
  float best_coef[4] = {0, 0, 0, 0};
  int iterations = 0;
  if (counts_out)
    for (int k = 0; k < K; ++k) counts_out[k] = -1;
  if (mode == 0) {
    for (int k = 0; k < K && n > 0; ++k) {
      float coef[4];
      int c = 0;
      if (model_from_triple(pts, tri + 3 * k, coef)) c = count_within(pts, n, coef, thr);
      else c = 0;
      if (counts_out) counts_out[k] = c;
      if (c > best) {
        best = c;
        best_k = k;
        std::memcpy(best_coef, coef, sizeof(coef));
      }
      ++iterations;
    }
  } else {
    // RandomSampleConsensus::computeModel
    double kk = 1.0;
    const double log_probability = std::log(1.0 - probability);
    const double one_over_indices = n > 0 ? 1.0 / (double)n : 0.0;
    int skipped = 0;
    const int max_skip = max_iterations * 10;
    int s = 0;  // position in the sample stream
    while (iterations < kk && skipped < max_skip && s < K && n > 0) {
      float coef[4];
      int k = s++;
      if (!model_from_triple(pts, tri + 3 * k, coef)) {
        ++skipped;
        continue;
      }
      int c = count_within(pts, n, coef, thr);
      if (counts_out) counts_out[k] = c;
      if (c > best) {
        best = c;
        best_k = k;
        std::memcpy(best_coef, coef, sizeof(coef));
        double w = (double)best * one_over_indices;
        double p_no_outliers = 1.0 - std::pow(w, 3.0);
        p_no_outliers = std::max(std::numeric_limits<double>::epsilon(), p_no_outliers);
        p_no_outliers = std::min(1.0 - std::numeric_limits<double>::epsilon(), p_no_outliers);
        kk = log_probability / std::log(p_no_outliers);
      }
      ++iterations;
      if (iterations > max_iterations) break;
    }
  }
  R.iterations = iterations;
  R.best_hyp = best_k;
  R.best_count = best;
  if (best_k < 0) {
    R.status = 2;
    if (mask_out) std::memset(mask_out, 0, n);
    return;
  }
  std::memcpy(R.coef, best_coef, sizeof(best_coef));
  if (refine)
    refine_plane(pts, n, best_coef, thr, R.refined, R.centroid);
  else
    std::memcpy(R.refined, best_coef, sizeof(best_coef));
  int rc = 0;
  for (int i = 0; i < n; ++i) {
    const float* p = pts + 4 * (size_t)i;
    bool in = plane_dist(R.refined, p[0], p[1], p[2]) < thr;
    rc += in;
    if (mask_out) mask_out[i] = in ? 1 : 0;
  }
  R.refined_count = rc;
}

extern "C" {

void orc_pcl_sample_stream(int n, int n_draws, unsigned seed, int* triples) { orcr::pcl_sample_stream(n, n_draws, seed, triples); }

int orc_crop(const void* msg, int width, int height, int point_step, int row_step, const int* offsets4,
             const int* box4, float* out) {
  return crop((const uint8_t*)msg, width, height, point_step, row_step, offsets4, box4, out);
}

// mode 0: fixed-K (score all K hypotheses, first best wins)   mode 1: PCL adaptive (max_iter cap)
// counts_out: nb*K int32 (fixed-K: every hypothesis; adaptive: -1 where not evaluated), may be null
// mask_out : uint8, concatenated per crop (prefix sums of w*h over non-spurious boxes), may be null
// threshold is a double like pcl::SACSegmentation::setDistanceThreshold
int orc_ransac_batch(const void* msg, int width, int height, int point_step, int row_step, const int* offsets4,
                       const int* boxes, int nb, const int* triples, int K, double threshold, int refine, int mode,
                       int max_iterations, double probability, PlaneResult* results, int* counts_out,
                       uint8_t* mask_out) {
  size_t mask_off = 0;
  std::vector<float> pts;
  for (int b = 0; b < nb; ++b) {
    PlaneResult& R = results[b];
    std::memset(&R, 0, sizeof(R));
    R.best_hyp = -1;
    const int* box = boxes + 4 * b;
    int n = crop((const uint8_t*)msg, width, height, point_step, row_step, offsets4, box, nullptr);
    if (n < 0) {
      R.status = 1;
      if (counts_out)
        for (int k = 0; k < K; ++k) counts_out[(size_t)b * K + k] = -1;
      continue;
    }
    R.n_points = n;
    pts.resize((size_t)4 * std::max(n, 1));
    crop((const uint8_t*)msg, width, height, point_step, row_step, offsets4, box, pts.data());
    orc_ransac_points(pts.data(), n, triples + (size_t)3 * K * b, K, threshold, refine, mode, max_iterations, probability, R,
                      counts_out ? counts_out + (size_t)b * K : nullptr, mask_out ? mask_out + mask_off : nullptr);
    mask_off += (size_t)n;
  }
  return 0;
}

}  // extern "C"
