// ORACLE — test infrastructure only: what oracle_cluster.cpp needs from oracle_ransac.cpp.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>

namespace orcr {
struct PlaneResult {       // must match ssb_plane_result in include/ssb.h
  int status;              // 0 ok, 1 spurious bbox (plane_segmentation.cpp:34-38), 2 no valid model
  int n_points;            // w*h of the crop
  int best_hyp;            // index of the winning hypothesis (first best), -1 if none
  int best_count;          // its inlier count
  int iterations;          // hypotheses consumed (fixed-K: K; adaptive: PCL iterations_)
  int refined_count;       // inliers of the refined model (== best_count when refine off)
  float coef[4];           // winning 3-point model
  float refined[4];        // after optimizeModelCoefficients (== coef when refine off or < 4 inliers)
  float centroid[3];       // centroid of the winning model's inliers (zeros when not refined)
  int reserved;
};
static inline float plane_dist(const float* c, float x, float y, float z) {
  // Eigen SSE3 dot: (c0*x + c1*y) + (c2*z + c3*1)
  float a = c[0] * x;
  float b = c[1] * y;
  float cc = c[2] * z;
  float s0 = a + b;
  float s1 = cc + c[3];
  return std::fabs(s0 + s1);
}

// smallest float t with (double)t >= thr : float d satisfies (double)d < thr  <=>  d < t
static inline float effective_threshold(double thr) {
  float t = (float)thr;
  if ((double)t < thr) t = std::nextafterf(t, std::numeric_limits<float>::infinity());
  return t;
}

// pcl::SampleConsensusModel::drawIndexSample, n_draws times on a fresh model over n points (see oracle_ransac.cpp)
void pcl_sample_stream(int n, int n_draws, unsigned seed, int* triples);
}  // namespace orcr

// pcl::SACSegmentation (SACMODEL_PLANE, SAC_RANSAC) on one point set (float4 records), see oracle_ransac.cpp
void orc_ransac_points(const float* pts, int n, const int* tri, int K, double threshold, int refine, int mode, int max_iterations,
                       double probability, orcr::PlaneResult& R, int* counts_out, uint8_t* mask_out);
