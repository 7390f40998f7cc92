// data_association_b200.h — drop-in for `class data_association`
// (/root/reference/include/ps_graph_slam/data_association.h:20-402) over the C-ABI of include/ssb.h.
//
// Same method names and meaning: find_matches / setLandmarkCovs / getMappedLandmarks / assignLandmarkNode.  The
// reference keeps a g2o::VertexPointXYZ* per landmark and reads node->estimate() while associating (:378); here
// assignLandmarkNode binds a callable that returns that estimate (the facade's VertexPointXYZ shim works as is).
// Types: `type` / `plane_type` are integer ids (the reference compares std::string for equality only, :123-124).
#pragma once
#include <functional>
#include <string>
#include <vector>

#include "ssb.h"

namespace ssb_host {

struct detected_object {   // include/planar_segmentation/detected_object.h:14-24 (fields used by the association)
  int type = 0, plane_type = 0;
  float pose[3] = {0, 0, 0};
  float normal_orientation[4] = {0, 0, 0, 0};
};

class data_association {
 public:
  explicit data_association(bool verbose, const ssb_assoc_opts* opts = nullptr) : verbose_(verbose) {
    h_ = ssb_assoc_create(opts);
  }
  ~data_association() { ssb_assoc_destroy(h_); }
  data_association(const data_association&) = delete;
  data_association& operator=(const data_association&) = delete;

  // data_association::find_matches  :75-95  (robot_pose = x y z roll pitch yaw)
  std::vector<ssb_landmark_obs> find_matches(const std::vector<detected_object>& seg_obj_info, const float robot_pose[6],
                                             float cam_angle) {
    refresh_estimates();
    std::vector<ssb_detection> d(seg_obj_info.size());
    for (size_t k = 0; k < d.size(); ++k) {
      d[k].type = seg_obj_info[k].type;
      d[k].plane_type = seg_obj_info[k].plane_type;
      for (int c = 0; c < 3; ++c) d[k].pose[c] = seg_obj_info[k].pose[c];
      for (int c = 0; c < 4; ++c) d[k].normal[c] = seg_obj_info[k].normal_orientation[c];
    }
    std::vector<ssb_landmark_obs> out(d.size());
    if (ssb_assoc_find_matches(h_, d.data(), (int)d.size(), robot_pose, cam_angle, out.data()) < 0) out.clear();
    std::vector<ssb_landmark_obs> res;
    for (auto& o : out)
      if (o.id >= 0) res.push_back(o);
    return res;
  }
  // data_association::assignLandmarkNode  :391-393
  void assignLandmarkNode(int id, std::function<void(double[3])> estimate) {
    if ((int)nodes_.size() <= id) nodes_.resize(id + 1);
    nodes_[id] = std::move(estimate);
  }
  // data_association::setLandmarkCovs  :395-397
  void setLandmarkCovs(int id, const float cov[9]) { ssb_assoc_set_landmark_cov(h_, id, cov); }
  // data_association::getMappedLandmarks  :399
  void getMappedLandmarks(std::vector<ssb_landmark_obs>& l) const {
    l.resize(ssb_assoc_num_landmarks(h_));
    for (size_t k = 0; k < l.size(); ++k) ssb_assoc_get_landmark(h_, (int)k, &l[k]);
  }

 private:
  void refresh_estimates() {
    for (size_t id = 0; id < nodes_.size(); ++id)
      if (nodes_[id]) {
        double p[3];
        nodes_[id](p);
        ssb_assoc_set_landmark_estimate(h_, (int)id, p);
      }
  }
  bool verbose_;
  ssb_assoc* h_ = nullptr;
  std::vector<std::function<void(double[3])>> nodes_;
};

}  // namespace ssb_host
