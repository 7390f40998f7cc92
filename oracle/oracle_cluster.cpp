// ORACLE — test infrastructure only.  CPU restatement of the reference's dormant plane-clustering chain
//   plane_segmentation::clusterAndSegmentAllPlanes            /root/reference/src/planar_segmentation/plane_segmentation.cpp:261-294
//     NormalBasedClusteringAndSegmentation                    :296-367   (removeNans :479-502, filterCentroids :504-523)
//     distanceBasedSegmentation                               :369-429
//     getFinalPoseWithNormals                                 :431-477
//     computeKmeans -> cv::kmeans                             :525-535
//     compute2DConvexHull: ProjectInliers + ConvexHull        :649-664   (its RANSAC part, :631-647, is oracle_ransac.cpp)
//
// The arithmetic lives in third-party libraries that are not vendored in /root/reference:
//   * OpenCV cv::kmeans (modules/core/src/kmeans.cpp; the reference links the distro's OpenCV 3.x).  Restated below from
//     the published algorithm: Lloyd iterations in single precision, KMEANS_RANDOM_CENTERS drawn from cv::RNG (multiply-
//     with-carry, coefficient 4164903690), `attempts` restarts, best compactness wins, empty clusters re-seeded with the
//     farthest point of the biggest cluster.  PINNED: OpenCV itself is importable in the build container (cv2 4.13), so
//     scripts/make_cluster_golden.py runs the real cv2.kmeans on seeded inputs and tests/test_oracle_cluster.py checks this
//     restatement against those fixtures bit for bit (labels, centres) — with the 4.x loop structure (the last iteration
//     keeps the labels and only measures distances); OpenCV 3.x recomputes the labels once more before returning.
//   * PCL ProjectInliers / SampleConsensusModelPlane::projectPoints and ConvexHull (qhull) for the planar case: the hull is
//     the set of extreme points of the projected inliers in the plane's dominant 2-D coordinates, listed counter-clockwise
//     by angle about their centroid (pcl/surface/impl/convex_hull.hpp, performReconstruction2D).  Pinned against
//     scipy.spatial.ConvexHull (the same qhull) for the vertex set.
// Nothing under semantic_slam_b200/ includes or links this file.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

// cv::RNG: state = (uint32)state * 4164903690 + (state >> 32); the low 32 bits are the draw
struct CvRng {
  uint64_t state;
  unsigned next() {
    state = (uint64_t)(unsigned)state * 4164903690ULL + (unsigned)(state >> 32);
    return (unsigned)state;
  }
  float uniform01() { return next() * 2.3283064365386962890625e-10f; }   // RNG::operator float()
};

// hal::normL2Sqr_ for dims below the SIMD width: one float accumulator, terms added in index order
float norm_l2_sqr(const float* a, const float* b, int n) {
  float s = 0.f;
  int j = 0;
  for (; j <= n - 4; j += 4) {
    const float t0 = a[j] - b[j], t1 = a[j + 1] - b[j + 1], t2 = a[j + 2] - b[j + 2], t3 = a[j + 3] - b[j + 3];
    s += t0 * t0 + t1 * t1 + t2 * t2 + t3 * t3;
  }
  for (; j < n; ++j) {
    const float t = a[j] - b[j];
    s += t * t;
  }
  return s;
}

// cv::kmeans(data, K, labels, TermCriteria(EPS + COUNT, max_count, eps), attempts, KMEANS_RANDOM_CENTERS, centers)
double cv_kmeans(const float* data, int N, int dims, int K, int max_count, double eps, int attempts, CvRng& rng, int* best_labels,
                 float* best_centers) {
  attempts = std::max(attempts, 1);
  eps = std::max(eps, 0.0);
  eps *= eps;
  max_count = std::min(std::max(max_count, 2), 100);
  if (K == 1) {
    attempts = 1;
    max_count = 2;
  }
  std::vector<float> centers((size_t)K * dims), old_centers((size_t)K * dims), temp(dims);
  std::vector<int> counters(K), labels(N);
  std::vector<double> dists(N);
  std::vector<float> lo(dims), hi(dims);
  for (int j = 0; j < dims; ++j) lo[j] = hi[j] = data[j];
  for (int i = 1; i < N; ++i)
    for (int j = 0; j < dims; ++j) {
      const float v = data[(size_t)i * dims + j];
      lo[j] = std::min(lo[j], v);
      hi[j] = std::max(hi[j], v);
    }
  double best_compactness = DBL_MAX;
  for (int a = 0; a < attempts; ++a) {
    double compactness = 0;
    for (int iter = 0;;) {
      double max_center_shift = iter == 0 ? DBL_MAX : 0.0;
      std::swap(centers, old_centers);
      if (iter == 0) {
        const float margin = 1.f / dims;   // generateRandomCenter
        for (int k = 0; k < K; ++k)
          for (int j = 0; j < dims; ++j)
            centers[(size_t)k * dims + j] = (rng.uniform01() * (1.f + margin * 2.f) - margin) * (hi[j] - lo[j]) + lo[j];
      } else {
        std::fill(centers.begin(), centers.end(), 0.f);
        std::fill(counters.begin(), counters.end(), 0);
        for (int i = 0; i < N; ++i) {
          const float* sample = data + (size_t)i * dims;
          float* center = &centers[(size_t)labels[i] * dims];
          for (int j = 0; j < dims; ++j) center[j] += sample[j];
          counters[labels[i]]++;
        }
        for (int k = 0; k < K; ++k) {
          if (counters[k] != 0) continue;
          // empty cluster: the farthest point of the biggest cluster becomes a one-point cluster
          int max_k = 0;
          for (int k1 = 1; k1 < K; ++k1)
            if (counters[max_k] < counters[k1]) max_k = k1;
          double max_dist = 0;
          int farthest_i = -1;
          float* base_center = &centers[(size_t)max_k * dims];
          const float scale = 1.f / counters[max_k];
          for (int j = 0; j < dims; ++j) temp[j] = base_center[j] * scale;
          for (int i = 0; i < N; ++i) {
            if (labels[i] != max_k) continue;
            const double dist = norm_l2_sqr(data + (size_t)i * dims, temp.data(), dims);
            if (max_dist <= dist) {
              max_dist = dist;
              farthest_i = i;
            }
          }
          counters[max_k]--;
          counters[k]++;
          labels[farthest_i] = k;
          const float* sample = data + (size_t)farthest_i * dims;
          float* cur_center = &centers[(size_t)k * dims];
          for (int j = 0; j < dims; ++j) {
            base_center[j] -= sample[j];
            cur_center[j] += sample[j];
          }
        }
        for (int k = 0; k < K; ++k) {
          float* center = &centers[(size_t)k * dims];
          const float scale = 1.f / counters[k];
          for (int j = 0; j < dims; ++j) center[j] *= scale;
          if (iter > 0) {
            double dist = 0;
            const float* old_center = &old_centers[(size_t)k * dims];
            for (int j = 0; j < dims; ++j) {
              const double t = center[j] - old_center[j];   // float difference, widened
              dist += t * t;
            }
            max_center_shift = std::max(max_center_shift, dist);
          }
        }
      }
      const bool last = (++iter == std::max(max_count, 2) || max_center_shift <= eps);
      if (last) {
        // labels are kept (no new empty clusters); only the distances to the own centre are measured
        for (int i = 0; i < N; ++i) dists[i] = norm_l2_sqr(data + (size_t)i * dims, &centers[(size_t)labels[i] * dims], dims);
        compactness = 0;
        for (int i = 0; i < N; ++i) compactness += dists[i];
        break;
      }
      for (int i = 0; i < N; ++i) {
        const float* sample = data + (size_t)i * dims;
        int k_best = 0;
        double min_dist = DBL_MAX;
        for (int k = 0; k < K; ++k) {
          const double dist = norm_l2_sqr(sample, &centers[(size_t)k * dims], dims);
          if (min_dist > dist) {
            min_dist = dist;
            k_best = k;
          }
        }
        dists[i] = min_dist;
        labels[i] = k_best;
      }
    }
    if (compactness < best_compactness) {
      best_compactness = compactness;
      std::memcpy(best_centers, centers.data(), centers.size() * sizeof(float));
      std::memcpy(best_labels, labels.data(), (size_t)N * sizeof(int));
    }
  }
  return best_compactness;
}

}  // namespace

extern "C" {

// cv::kmeans with KMEANS_RANDOM_CENTERS; rng_state in/out (cv::theRNG().state; OpenCV's default seed is 0xffffffff)
double orc_kmeans(const float* data, int N, int dims, int K, int max_count, double eps, int attempts, unsigned long long* rng_state,
                  int* labels, float* centers) {
  CvRng rng{(uint64_t)*rng_state};
  const double c = cv_kmeans(data, N, dims, K, max_count, eps, attempts, rng, labels, centers);
  *rng_state = rng.state;
  return c;
}

}  // extern "C"
