// Host-side mirror of the RANSAC part of /root/reference/include/planar_segmentation/plane_segmentation.h
// over the C-ABI (include/ssb.h):
//   segmentPointCloudData (plane_segmentation.cpp:24-82)      -> crop of one bbox
//   compute2DConvexHull's pcl::SACSegmentation (:631-647)     -> fitPlanes over all bboxes of a frame
//   the dormant clustering chain (:261-477, 525-535, 649-664)  -> computeKmeans, compute2DConvexHull, clusterAndSegmentAllPlanes
// sensor_msgs::PointCloud2 / semantic_SLAM::ObjectInfo are reduced to the fields those functions read
// (data pointer + layout, tl_x/tl_y/width/height), so the header needs neither ROS nor PCL.
#ifndef SSB_PLANE_SEGMENTATION_B200_H
#define SSB_PLANE_SEGMENTATION_B200_H

#include <string>
#include <vector>

#include "ssb.h"

class plane_segmentation_b200 {
 public:
  explicit plane_segmentation_b200(bool verbose) : verbose_(verbose) { h_ = ssb_ransac_create(-1); }
  ~plane_segmentation_b200() { ssb_ransac_destroy(h_); }
  bool ok() const { return h_ != nullptr; }

  // plane_segmentation::segmentPointCloudData: returns false for a "spurious" box (:34-38)
  bool segmentPointCloudData(const ssb_bbox& object_info, const void* cloud_data, const ssb_cloud_layout& layout,
                             std::vector<float>& segmented_xyzrgb) {
    int n = ssb_crop_bbox(h_, cloud_data, &layout, &object_info, nullptr);
    if (n < 0) return false;
    segmented_xyzrgb.resize((size_t)4 * n);
    if (n > 0) ssb_crop_bbox(h_, cloud_data, &layout, &object_info, segmented_xyzrgb.data());
    return true;
  }

  // pcl::SACSegmentation(SACMODEL_PLANE, SAC_RANSAC, 0.01, optimize) over every bbox.
  // n_hyp = 0 -> PCL's behaviour: its stopping rule (adaptive k, <= 50 iterations) on ITS sample stream — every crop gets
  // a fresh model like the `seg` object of compute2DConvexHull (:637), i.e. boost::mt19937(12345u) through
  // uniform_int<>(0, INT_MAX) and drawIndexSample's running shuffle (ssb_ransac_pcl_samples).  n_hyp > 0 -> score exactly
  // n_hyp hypotheses of that same stream, first best wins.
  std::vector<ssb_plane_result> fitPlanes(const void* cloud_data, const ssb_cloud_layout& layout,
                                          const std::vector<ssb_bbox>& boxes, int n_hyp = 0,
                                          std::vector<unsigned char>* inlier_mask = nullptr) {
    ssb_ransac_opts o;
    ssb_ransac_default_opts(&o);
    int K = n_hyp;
    if (K <= 0) {
      o.mode = 1;
      K = 512;  // 51 evaluated hypotheses at most + the draws PCL would reject and repeat (collinear samples)
    }
    std::vector<int> triples((size_t)3 * K * boxes.size());
    for (size_t b = 0; b < boxes.size(); ++b) {
      const bool spurious = boxes[b].width < 0 || boxes[b].height < 0;
      const long long n = spurious ? 0 : (long long)boxes[b].width * boxes[b].height;
      ssb_ransac_pcl_samples((int)n, K, 12345u, &triples[(size_t)3 * K * b]);
    }
    std::vector<ssb_plane_result> res(boxes.size());
    if (inlier_mask) {   // concatenated over the non-spurious boxes
      size_t total = 0;
      for (auto& b : boxes)
        if (b.width >= 0 && b.height >= 0) total += (size_t)b.width * b.height;
      inlier_mask->assign(total + 1, 0);
    }
    if (ssb_ransac_plane_batch(h_, cloud_data, &layout, boxes.data(), (int)boxes.size(), triples.data(), K, &o, res.data(),
                               nullptr, inlier_mask ? inlier_mask->data() : nullptr) != SSB_OK)
      res.clear();
    return res;
  }

  // plane_segmentation::computeKmeans (:525-535): cv::kmeans(points, K, labels, (EPS + ITER, 10, 0.01), 10, RANDOM_CENTERS,
  // centroids).  points: n x dims floats; rng_state stands for cv::theRNG() (initial state 0xffffffff) and is advanced.
  double computeKmeans(const std::vector<float>& points, int dims, int num_centroids, std::vector<int>& labels,
                       std::vector<float>& centroids) {
    const int n = dims > 0 ? (int)(points.size() / dims) : 0;
    labels.assign(n, 0);
    centroids.assign((size_t)num_centroids * dims, 0.f);
    double compactness = 0.0;
    if (ssb_kmeans(h_, points.data(), n, dims, num_centroids, 10, 0.01, 10, &rng_state, labels.data(), centroids.data(), &compactness) != SSB_OK)
      return -1.0;
    return compactness;
  }

  // plane_segmentation::compute2DConvexHull (:631-664) of one point set (n x 4 floats): RANSAC plane, ProjectInliers,
  // ConvexHull -> the hull vertices (k x 3 floats, PCL's output order)
  std::vector<float> compute2DConvexHull(const std::vector<float>& filtered_point_cloud_xyzrgb) {
    const int n = (int)(filtered_point_cloud_xyzrgb.size() / 4);
    std::vector<float> hull;
    if (n < 3) return hull;
    ssb_cloud_layout lay{n, 1, 16, 16 * n, 0, 4, 8, 12};   // the point set as a 1-row organised cloud of float4 records
    std::vector<ssb_bbox> box{{0, 0, n, 1}};
    std::vector<ssb_plane_result> res = fitPlanes(filtered_point_cloud_xyzrgb.data(), lay, box, 0, &mask_);
    if (res.empty() || res[0].status != 0) return hull;
    hull.resize((size_t)3 * n);
    int nin = 0;
    const int k = ssb_project_hull(h_, filtered_point_cloud_xyzrgb.data(), mask_.data(), n, res[0].refined, hull.data(), nullptr, n, &nin);
    hull.resize(k > 0 ? (size_t)3 * k : 0);
    return hull;
  }

  // plane_segmentation::clusterAndSegmentAllPlanes (:261-294): rows of 8 floats (x, y, z, nx, ny, nz, d, 0), one per hull vertex
  std::vector<float> clusterAndSegmentAllPlanes(const std::vector<float>& point_cloud_xyzrgb, const std::vector<float>& point_normal,
                                                const float transformation_mat[16], std::vector<ssb_plane_cluster>* info = nullptr) {
    const int n = (int)(point_cloud_xyzrgb.size() / 4);
    std::vector<float> rows((size_t)8 * 4096);
    std::vector<ssb_plane_cluster> cl(16);
    int n_rows = 0, n_cl = 0;
    if (ssb_cluster_planes(h_, point_cloud_xyzrgb.data(), point_normal.data(), n, transformation_mat, nullptr, &rng_state, rows.data(), 4096,
                           &n_rows, cl.data(), 16, &n_cl, nullptr, nullptr) < 0)
      n_rows = n_cl = 0;
    rows.resize((size_t)8 * n_rows);
    if (info) info->assign(cl.begin(), cl.begin() + n_cl);
    return rows;
  }

  unsigned long long rng_state = 0xffffffffULL;   // cv::theRNG().state

 private:
  std::vector<unsigned char> mask_;
  ssb_ransac* h_ = nullptr;
  bool verbose_;
};
#endif
