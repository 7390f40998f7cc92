// Host-side mirror of /root/reference/include/planar_segmentation/point_cloud_segmentation.h (header-only there too) over
// the C-ABI (include/ssb.h): same class and method names, same call order —
//   point_cloud_segmentation::segmentallPointCloudData   :105-181  class filter (:126-130), bbox crop, normals, planes
//   point_cloud_segmentation::segmentPlanarSurfaces      :26-103   regions -> detected_object (camera -> world)
// sensor_msgs::PointCloud2, semantic_SLAM::ObjectInfo, Eigen and PCL types are reduced to the fields those functions
// read, so the header needs neither ROS nor PCL nor Eigen.  The crop / integral-image normals / organised multi-plane
// segmentation of ALL accepted detections of a frame run in one device pass (ssb_organized_planes).
#ifndef SSB_POINT_CLOUD_SEGMENTATION_H
#define SSB_POINT_CLOUD_SEGMENTATION_H

#include <string>
#include <vector>

#include "ssb.h"

namespace ssb_host {

struct ObjectInfo {            // msg/ObjectInfo.msg:1-6
  std::string type;
  float prob = 0.0f;
  int tl_x = 0, tl_y = 0, width = 0, height = 0;
};
struct PointCloud2View {       // the fields of sensor_msgs::PointCloud2 read at plane_segmentation.cpp:44-61
  const void* data = nullptr;
  ssb_cloud_layout layout{};
};
struct segmented_object {      // include/planar_segmentation/detected_object.h:14-24, all fields
  int id = 0;
  float prob = 0.0f;
  float num_points = 0.0f;
  std::string type;
  std::string plane_type;
  float pose[3] = {0, 0, 0};
  float world_pose[3] = {0, 0, 0};
  float normal_orientation[4] = {0, 0, 0, 0};
};

}  // namespace ssb_host

class point_cloud_segmentation {
 public:
  explicit point_cloud_segmentation(bool verbose, int num_point_seg = 500, int norm_point_thres = 5000, float planar_area = 0.1f)
      : verbose_(verbose), planar_area_(planar_area) {
    h_ = ssb_ransac_create(-1);
    ssb_organized_default_opts(&opts_);
    opts_.min_inliers = num_point_seg;         // ros param num_point_seg, plane_segmentation.cpp:7
    opts_.norm_point_thres = norm_point_thres; // norm_point_thres, :8
  }
  ~point_cloud_segmentation() { ssb_ransac_destroy(h_); }
  point_cloud_segmentation(const point_cloud_segmentation&) = delete;
  point_cloud_segmentation& operator=(const point_cloud_segmentation&) = delete;
  bool ok() const { return h_ != nullptr; }

  // the class list of :126-130
  static bool accepted_class(const std::string& t) {
    return t == "chair" || t == "tvmonitor" || t == "book" || t == "keyboard" || t == "laptop" || t == "bucket" || t == "car";
  }

  // point_cloud_segmentation::segmentPlanarSurfaces: the planar regions of ONE crop -> detected objects
  std::vector<ssb_host::segmented_object> segmentPlanarSurfaces(const std::vector<ssb_planar_region>& regions, const float robot_pose[6],
                                                                float cam_angle, const std::string& object_type, float prob) {
    std::vector<ssb_host::segmented_object> out;
    if (regions.empty()) return out;
    std::vector<ssb_detected_object> det(regions.size());
    const int n = ssb_segment_planar_surfaces(regions.data(), (int)regions.size(), robot_pose, cam_angle, 0, prob, planar_area_, det.data());
    for (int k = 0; k < n; ++k) {
      ssb_host::segmented_object o;
      o.prob = det[k].prob;
      o.num_points = det[k].num_points;
      o.type = object_type;
      o.plane_type = det[k].plane_type == 0 ? "horizontal" : "vertical";
      for (int c = 0; c < 3; ++c) {
        o.pose[c] = det[k].pose[c];
        o.world_pose[c] = det[k].world_pose[c];
      }
      for (int c = 0; c < 4; ++c) o.normal_orientation[c] = det[k].normal_orientation[c];
      out.push_back(o);
    }
    return out;
  }

  // point_cloud_segmentation::segmentallPointCloudData
  std::vector<ssb_host::segmented_object> segmentallPointCloudData(const float robot_pose[6], float cam_angle,
                                                                   const std::vector<ssb_host::ObjectInfo>& object_info,
                                                                   const ssb_host::PointCloud2View& point_cloud) {
    std::vector<ssb_host::segmented_object> complete_obj_info_vec;
    std::vector<ssb_bbox> boxes;
    std::vector<size_t> src;
    for (size_t i = 0; i < object_info.size(); ++i)
      if (accepted_class(object_info[i].type)) {
        boxes.push_back({object_info[i].tl_x, object_info[i].tl_y, object_info[i].width, object_info[i].height});
        src.push_back(i);
      }
    if (boxes.empty() || !h_) return complete_obj_info_vec;
    const int max_regions = 16;
    std::vector<ssb_planar_region> regions(boxes.size() * max_regions);
    std::vector<int> n_regions(boxes.size(), 0);
    if (ssb_organized_planes(h_, point_cloud.data, &point_cloud.layout, boxes.data(), (int)boxes.size(), &opts_, max_regions,
                             regions.data(), n_regions.data(), nullptr, nullptr, nullptr, nullptr) != SSB_OK)
      return complete_obj_info_vec;
    for (size_t b = 0; b < boxes.size(); ++b) {
      if (n_regions[b] <= 0) continue;   // spurious box (:141), crop without normals (:164-165) or no plane
      const int m = n_regions[b] < max_regions ? n_regions[b] : max_regions;
      std::vector<ssb_planar_region> one(regions.begin() + b * max_regions, regions.begin() + b * max_regions + m);
      for (auto& o : segmentPlanarSurfaces(one, robot_pose, cam_angle, object_info[src[b]].type, object_info[src[b]].prob))
        complete_obj_info_vec.push_back(o);
    }
    return complete_obj_info_vec;
  }

 private:
  ssb_ransac* h_ = nullptr;
  ssb_organized_opts opts_;
  bool verbose_;
  float planar_area_;
};
#endif
