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
//     the set of extreme points of the projected inliers in two of their coordinates, listed by decreasing
//     angle about their centroid (pcl/surface/impl/convex_hull.hpp, performReconstruction2D).  Pinned against
//     scipy.spatial.ConvexHull (the same qhull) for the vertex set.
// Nothing under semantic_slam_b200/ includes or links this file.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <random>
#include <utility>
#include <vector>

#include "oracle_ransac.h"

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


// ---- pcl::SampleConsensusModelPlane::projectPoints (copy_data_fields = false), as called by pcl::ProjectInliers ----------
// (sac_model_plane.hpp) mc = (a, b, c, 0) normalised, tmp_mc = (mc.xyz, d); for every inlier p:
//   distance_to_plane = tmp_mc . (x, y, z, 1)   [Eigen 4-float dot: (p0 + p1) + (p2 + p3)],   pp = p - mc * distance_to_plane
void project_inliers(const float* pts4, const uint8_t* mask, int n, const float* coef, std::vector<float>& out4, std::vector<int>& src) {
  float mc[4] = {coef[0], coef[1], coef[2], 0.f};
  const float sq = (mc[0] * mc[0] + mc[1] * mc[1]) + (mc[2] * mc[2] + mc[3] * mc[3]);
  const float nrm = std::sqrt(sq);
  for (int k = 0; k < 4; ++k) mc[k] = mc[k] / nrm;
  const float tmp[4] = {mc[0], mc[1], mc[2], coef[3]};
  out4.clear();
  src.clear();
  for (int i = 0; i < n; ++i) {
    if (!mask[i]) continue;
    const float* p = pts4 + 4 * (size_t)i;
    const float d = (tmp[0] * p[0] + tmp[1] * p[1]) + (tmp[2] * p[2] + tmp[3] * 1.0f);
    out4.push_back(p[0] - mc[0] * d);
    out4.push_back(p[1] - mc[1] * d);
    out4.push_back(p[2] - mc[2] * d);
    out4.push_back(0.f);
    src.push_back(i);
  }
}

// ---- pcl::ConvexHull::performReconstruction2D (convex_hull.hpp) ------------------------------------------------------------
// Coordinates: (x, y) unless the plane normal is within 10 degrees of the x or y axis (projection_angle_thresh_ =
// cos(0.174532925)), then (y, z), then (x, z).  qhull on those two coordinates (as doubles) = the strictly convex vertices;
// restated as Andrew's monotone chain with the orientation test in double, collinear points dropped.  Output order: sorted by
// DEcreasing atan2(v - cv, u - cu) about the centroid of the hull vertices (comparePoints2D).  PCL takes the normal for the
// axis choice from three of the projected points; here it is the model's normal (the same plane).
int hull_axes(const float* coef, int* iu, int* iv) {
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
double orient(const double* a, const double* b, const double* c) { return (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0]); }
// points: m records of 4 floats; returns the hull vertices (indices into the m points) in PCL's output order
void convex_hull_2d(const float* pts4, int m, int iu, int iv, std::vector<int>& hull) {
  hull.clear();
  if (m <= 0) return;
  std::vector<int> ord(m);
  for (int i = 0; i < m; ++i) ord[i] = i;
  auto U = [&](int i) { return (double)pts4[4 * (size_t)i + iu]; };
  auto V = [&](int i) { return (double)pts4[4 * (size_t)i + iv]; };
  std::sort(ord.begin(), ord.end(), [&](int a, int b) {
    if (U(a) != U(b)) return U(a) < U(b);
    if (V(a) != V(b)) return V(a) < V(b);
    return a < b;
  });
  // identical coordinates: the lowest index stands for all of them
  std::vector<int> uq;
  for (int k = 0; k < m; ++k)
    if (uq.empty() || U(ord[k]) != U(uq.back()) || V(ord[k]) != V(uq.back())) uq.push_back(ord[k]);
  const int q = (int)uq.size();
  if (q < 3) {
    hull = uq;
  } else {
    std::vector<int> st(2 * q);
    int k = 0;
    auto turn = [&](int a, int b, int c) {
      const double A[2] = {U(a), V(a)}, B[2] = {U(b), V(b)}, C[2] = {U(c), V(c)};
      return orient(A, B, C);
    };
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
  // PCL's output order: decreasing angle about the centroid of the hull vertices
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
    const float du = pts4[4 * (size_t)i + iu] - fcu, dv = pts4[4 * (size_t)i + iv] - fcv;
    ang.push_back({std::atan2((double)dv, (double)du) + M_PI, i});
  }
  std::stable_sort(ang.begin(), ang.end(), [](const std::pair<double, int>& a, const std::pair<double, int>& b) { return a.first > b.first; });
  hull.clear();
  for (auto& a : ang) hull.push_back(a.second);
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


// pcl::ProjectInliers (SACMODEL_PLANE) + pcl::ConvexHull of the points selected by `mask` (plane_segmentation.cpp:649-662).
// rows3: the hull vertices (projected points) in PCL's output order; src_out: index of each vertex in pts4.  Returns the count.
int orc_project_hull(const float* pts4, const unsigned char* mask, int n, const float* coef, float* rows3, int* src_out, int max_rows,
                     int* n_inliers) {
  std::vector<float> proj;
  std::vector<int> src, hull;
  project_inliers(pts4, mask, n, coef, proj, src);
  if (n_inliers) *n_inliers = (int)src.size();
  int iu, iv;
  if (hull_axes(coef, &iu, &iv) < 0) return 0;
  convex_hull_2d(proj.data(), (int)src.size(), iu, iv, hull);
  int k = 0;
  for (int h : hull) {
    if (k >= max_rows) break;
    for (int c = 0; c < 3; ++c) rows3[3 * k + c] = proj[4 * (size_t)h + c];
    if (src_out) src_out[k] = src[h];
    ++k;
  }
  return (int)hull.size();
}

struct OrcPlaneCluster {   // must match ssb_plane_cluster in include/ssb.h
  float normal[3];
  float distance;
  int normal_label, distance_label;
  int n_points, n_inliers;
  float coef[4];
  int row0, n_rows;
};

// plane_segmentation::clusterAndSegmentAllPlanes (:261-294).  cloud4 / normals4: n records of 4 floats (x y z rgb / nx ny nz
// curvature); T16: transformation_mat, row-major.  RANSAC sample stream per cluster: PCL's own (orcr::pcl_sample_stream, a fresh
// model per compute2DConvexHull call); n_hyp = 0: PCL's adaptive stopping rule on a 512-draw stream.
// coef_override (4 floats per cluster, may be null): use these planes for ProjectInliers + ConvexHull instead of the oracle's own
// refined ones (the GPU's refine differs in summation order, compared at 1e-5: the hull stage is checked on identical planes).
// labels_out: n ints (first k-means, -1 where the normal is NaN) or null; centers_out: Kn x 3 or null.
int orc_cluster_planes(const float* cloud4, const float* normals4, int n, const float* T16, int Kn, int Kd, int attempts, int max_count,
                       double eps, int min_pts, float tol, int n_hyp, unsigned seed, unsigned long long* rng_state, float* rows8,
                       int max_rows, int* n_rows, OrcPlaneCluster* clusters, int max_clusters, int* n_clusters, int* labels_out,
                       float* centers_out, const float* coef_override) {
  *n_rows = 0;
  *n_clusters = 0;
  if (labels_out)
    for (int i = 0; i < n; ++i) labels_out[i] = -1;
  // removeNans :479-502
  std::vector<int> keep;
  std::vector<float> nrm3;
  for (int i = 0; i < n; ++i) {
    const float* q = normals4 + 4 * (size_t)i;
    if (!std::isnan(q[0]) && !std::isnan(q[1]) && !std::isnan(q[2])) {
      keep.push_back(i);
      nrm3.push_back(q[0]);
      nrm3.push_back(q[1]);
      nrm3.push_back(q[2]);
    }
  }
  const int m = (int)keep.size();
  if (m <= 10 || m < Kn) return 0;   // :316-320
  CvRng rng{(uint64_t)*rng_state};
  std::vector<int> labels(m);
  std::vector<float> centers((size_t)Kn * 3);
  cv_kmeans(nrm3.data(), m, 3, Kn, max_count, eps, attempts, rng, labels.data(), centers.data());
  if (labels_out)
    for (int k = 0; k < m; ++k) labels_out[keep[k]] = labels[k];
  if (centers_out) std::memcpy(centers_out, centers.data(), centers.size() * sizeof(float));
  // normals_of_the_horizontal_plane_in_cam = transformation_mat^T * (0, 0, 1, 0) :332-346 = third row of the matrix
  const float hz[3] = {T16[8], T16[9], T16[10]};
  int rows = 0, ncl = 0;
  for (int c = 0; c < Kn; ++c) {   // filterCentroids :504-523 (float against float + double 0.3)
    const float* cc = &centers[3 * (size_t)c];
    bool ok = true;
    for (int j = 0; j < 3; ++j) ok = ok && ((double)cc[j] < (double)hz[j] + (double)tol) && ((double)cc[j] > (double)hz[j] - (double)tol);
    if (!ok) continue;
    // the points of this normal cluster :349-362, their signed distances :377-392
    std::vector<int> mem;
    std::vector<float> dist;
    for (int k = 0; k < m; ++k) {
      if (labels[k] != c) continue;
      const float* p = cloud4 + 4 * (size_t)keep[k];
      float d = p[0] * cc[0] + p[1] * cc[1] + p[2] * cc[2];
      d = -1 * d;
      mem.push_back(keep[k]);
      dist.push_back(d);
    }
    const int md = (int)mem.size();
    if (md < Kd) continue;   // cv::kmeans would throw (N < K); the reference never guards it
    std::vector<int> dl(md);
    std::vector<float> dc(Kd);
    cv_kmeans(dist.data(), md, 1, Kd, max_count, eps, attempts, rng, dl.data(), dc.data());
    for (int d = 0; d < Kd; ++d) {   // :399-425
      std::vector<float> pts;
      for (int k = 0; k < md; ++k)
        if (dl[k] == d) {
          const float* p = cloud4 + 4 * (size_t)mem[k];
          pts.insert(pts.end(), p, p + 4);
        }
      const int np = (int)(pts.size() / 4);
      if (!(np > min_pts)) continue;
      if (ncl >= max_clusters) continue;
      OrcPlaneCluster& C = clusters[ncl];
      std::memset(&C, 0, sizeof(C));
      for (int j = 0; j < 3; ++j) C.normal[j] = cc[j];
      C.distance = dc[d];
      C.normal_label = c;
      C.distance_label = d;
      C.n_points = np;
      // compute2DConvexHull :631-664
      const int K = n_hyp > 0 ? n_hyp : 512;
      std::vector<int> tri((size_t)3 * K);
      orcr::pcl_sample_stream(np, K, seed, tri.data());
      orcr::PlaneResult R;
      std::memset(&R, 0, sizeof(R));
      R.n_points = np;
      std::vector<uint8_t> mask(np);
      orc_ransac_points(pts.data(), np, tri.data(), K, 0.01, 1, n_hyp > 0 ? 0 : 1, 50, 0.99, R, nullptr, mask.data());
      float coef[4];
      std::memcpy(coef, R.refined, sizeof(coef));
      if (coef_override) {
        std::memcpy(coef, coef_override + 4 * (size_t)ncl, sizeof(coef));
        for (int i = 0; i < np; ++i) {
          const float* p = pts.data() + 4 * (size_t)i;
          mask[i] = orcr::plane_dist(coef, p[0], p[1], p[2]) < orcr::effective_threshold(0.01) ? 1 : 0;
        }
      }
      std::memcpy(C.coef, coef, sizeof(coef));
      C.row0 = rows;
      if (R.best_hyp >= 0) {
        std::vector<float> h3((size_t)3 * np);
        int nin = 0;
        const int nh = orc_project_hull(pts.data(), mask.data(), np, coef, h3.data(), nullptr, np, &nin);
        C.n_inliers = nin;
        for (int k = 0; k < nh && rows < max_rows; ++k, ++rows) {   // getFinalPoseWithNormals :431-477
          float* r8 = rows8 + 8 * (size_t)rows;
          r8[0] = h3[3 * k];
          r8[1] = h3[3 * k + 1];
          r8[2] = h3[3 * k + 2];
          r8[3] = cc[0];
          r8[4] = cc[1];
          r8[5] = cc[2];
          r8[6] = dc[d];
          r8[7] = 0.f;
        }
        C.n_rows = rows - C.row0;
      }
      ++ncl;
    }
  }
  *rng_state = rng.state;
  *n_rows = rows;
  *n_clusters = ncl;
  return 1;
}

}  // extern "C"
