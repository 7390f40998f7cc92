// Host-side mirror of the RANSAC part of /root/reference/include/planar_segmentation/plane_segmentation.h
// over the C-ABI (include/ssb.h):
//   segmentPointCloudData (plane_segmentation.cpp:24-82)      -> crop of one bbox
//   compute2DConvexHull's pcl::SACSegmentation (:631-647)     -> fitPlanes over all bboxes of a frame
// sensor_msgs::PointCloud2 / semantic_SLAM::ObjectInfo are reduced to the fields those functions read
// (data pointer + layout, tl_x/tl_y/width/height), so the header needs neither ROS nor PCL.
#ifndef SSB_PLANE_SEGMENTATION_B200_H
#define SSB_PLANE_SEGMENTATION_B200_H

#include <random>
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
  // n_hyp = 0 -> PCL's STOPPING RULE (adaptive k, <= 50 iterations).  The 3-point samples come from std::mt19937(12345)
  // modulo n, not from pcl::RandomSampleConsensus' boost::mt19937 + drawIndexSample shuffle (boost is absent here), so
  // PCL's hypothesis sequence itself is not reproduced — pass your own index triples to ssb_ransac_plane_batch for that.
  std::vector<ssb_plane_result> fitPlanes(const void* cloud_data, const ssb_cloud_layout& layout,
                                          const std::vector<ssb_bbox>& boxes, int n_hyp = 0) {
    ssb_ransac_opts o;
    ssb_ransac_default_opts(&o);
    int K = n_hyp;
    if (K <= 0) {
      o.mode = 1;
      K = 512;  // sample stream long enough for 50 valid iterations + skipped samples
    }
    std::vector<int> triples((size_t)3 * K * boxes.size());
    for (size_t b = 0; b < boxes.size(); ++b) {
      std::mt19937 rng(12345u);
      const long long n = (long long)boxes[b].width * boxes[b].height;
      for (int k = 0; k < 3 * K; ++k) triples[3 * K * b + k] = n > 0 ? (int)(rng() % (unsigned long long)n) : 0;
    }
    std::vector<ssb_plane_result> res(boxes.size());
    if (ssb_ransac_plane_batch(h_, cloud_data, &layout, boxes.data(), (int)boxes.size(), triples.data(), K, &o, res.data(),
                               nullptr, nullptr) != SSB_OK)
      res.clear();
    return res;
  }

 private:
  ssb_ransac* h_ = nullptr;
  bool verbose_;
};
#endif
