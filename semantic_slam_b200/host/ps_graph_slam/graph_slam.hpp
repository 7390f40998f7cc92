// Drop-in replacement for /root/reference/include/ps_graph_slam/graph_slam.hpp (class
// ps_graph_slam::GraphSLAM, :27-150) backed by the B200 C-ABI (include/ssb.h) instead of g2o.
//
// Same class name, method names, argument meaning and return conventions as the reference:
//   add_se3_node / add_point_xyz_node / add_se3_edge / add_se3_point_xyz_edge /
//   add_point_xyz_point_xyz_edge / optimize / computeLandmarkMarginals / save.
// The g2o handle types the callers touch are replaced by light shims in namespace g2o that expose
// exactly what the reference's callers use:
//   node->estimate()        semantic_graph_slam.cpp:94-95, data_association.h:378
//   node->hessianIndex()    semantic_graph_slam.cpp:188-190      node->id()   keyframe.cpp:36
//   node->unlockQuadraticForm()                                  semantic_graph_slam.cpp:188
//   spinv.block(i,i)->eval()                                     semantic_graph_slam.cpp:199-201
// With Eigen present (the reference's build) the signatures take Eigen::Isometry3d / Vector3d /
// MatrixXd exactly like the reference; without Eigen (this repo's CI image) a POD twin with the same
// layout is used so the facade is compiled and tested here too.
#ifndef SSB_GRAPH_SLAM_HPP
#define SSB_GRAPH_SLAM_HPP

#include <array>
#include <cstdio>
#include <iostream>
#include <memory>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "ssb.h"

#if defined(SSB_USE_EIGEN) || __has_include(<Eigen/Core>)
#include <Eigen/Core>
#include <Eigen/Geometry>
#define SSB_HAVE_EIGEN 1
namespace ssb_host {
using Isometry3d = Eigen::Isometry3d;
using Vector3d = Eigen::Vector3d;
using Vector4d = Eigen::Vector4d;
using MatrixXd = Eigen::MatrixXd;
inline void to34(const Isometry3d& T, double* o) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) o[4 * r + c] = T.matrix()(r, c);
}
inline Isometry3d from34(const double* o) {
  Isometry3d T = Isometry3d::Identity();
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) T.matrix()(r, c) = o[4 * r + c];
  return T;
}
inline void toRowMajor(const MatrixXd& M, int n, double* o) {
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c) o[n * r + c] = M(r, c);
}
}  // namespace ssb_host
#else
namespace ssb_host {
// POD twins (row-major), only what the facade needs
struct Isometry3d {
  double m[12];  // 3x4 [R|t]
  static Isometry3d Identity() { return Isometry3d{{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0}}; }
};
struct Vector3d {
  double v[3];
  double& operator()(int i) { return v[i]; }
  double operator()(int i) const { return v[i]; }
};
struct Vector4d {
  double v[4];
  double& operator()(int i) { return v[i]; }
  double operator()(int i) const { return v[i]; }
};
struct MatrixXd {
  int n = 0;
  std::vector<double> a;
  MatrixXd() = default;
  explicit MatrixXd(int n_) : n(n_), a((size_t)n_ * n_, 0.0) {}
  MatrixXd(int r_, int c_) : n(r_), a((size_t)r_ * c_, 0.0) {}   // (rows, cols) like Eigen::MatrixXd; square only
  static MatrixXd Identity(int n_) {
    MatrixXd M(n_);
    for (int i = 0; i < n_; ++i) M(i, i) = 1.0;
    return M;
  }
  double& operator()(int r, int c) { return a[(size_t)r * n + c]; }
  double operator()(int r, int c) const { return a[(size_t)r * n + c]; }
  int rows() const { return n; }
};
inline void to34(const Isometry3d& T, double* o) {
  for (int k = 0; k < 12; ++k) o[k] = T.m[k];
}
inline Isometry3d from34(const double* o) {
  Isometry3d T;
  for (int k = 0; k < 12; ++k) T.m[k] = o[k];
  return T;
}
inline void toRowMajor(const MatrixXd& M, int n, double* o) {
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c) o[n * r + c] = M(r, c);
}
}  // namespace ssb_host
#endif

namespace g2o {
// Shims for the vertex / edge handle types the reference's callers hold as raw pointers.
class VertexBase {
 public:
  VertexBase(ssb_graph* g, int id) : g_(g), id_(id) {}
  int id() const { return id_; }
  int hessianIndex() const { return ssb_graph_hessian_index(g_, id_); }
  void unlockQuadraticForm() {}
  void setFixed(bool f) { ssb_graph_set_fixed(g_, id_, f ? 1 : 0); }

 protected:
  ssb_graph* g_;
  int id_;
};
class VertexSE3 : public VertexBase {
 public:
  using VertexBase::VertexBase;
  ssb_host::Isometry3d estimate() const {
    double T[12];
    ssb_graph_get_se3(g_, id_, T);
    return ssb_host::from34(T);
  }
  void setEstimate(const ssb_host::Isometry3d& T) {
    double o[12];
    ssb_host::to34(T, o);
    ssb_graph_set_se3(g_, id_, o);
  }
};
class VertexPointXYZ : public VertexBase {
 public:
  using VertexBase::VertexBase;
  ssb_host::Vector3d estimate() const {
    double p[3];
    ssb_graph_get_point_xyz(g_, id_, p);
    ssb_host::Vector3d v;
    v(0) = p[0];
    v(1) = p[1];
    v(2) = p[2];
    return v;
  }
};
// g2o::VertexPlane stand-in (estimate() = the 4 normalised plane coefficients, g2o::Plane3D::toVector())
class VertexPlane : public VertexBase {
 public:
  using VertexBase::VertexBase;
  ssb_host::Vector4d estimate() const {
    double c[4];
    ssb_graph_get_plane(g_, id_, c);
    ssb_host::Vector4d v;
    for (int k = 0; k < 4; ++k) v(k) = c[k];
    return v;
  }
};
struct EdgeHandle {
  int id;
};
using EdgeSE3Plane = EdgeHandle;
using EdgeSE3 = EdgeHandle;
using EdgeSE3PointXYZ = EdgeHandle;
using EdgePointXYZ = EdgeHandle;

// g2o::SparseBlockMatrix<MatrixXd> stand-in for the marginals: block(i,i)->eval()
template <class M>
class SparseBlockMatrix {
 public:
  struct Block {
    M m;
    const M& eval() const { return m; }
  };
  const Block* block(int r, int c) const {
    for (auto& e : blocks_)
      if (e.first.first == r && e.first.second == c) return &e.second;
    return nullptr;
  }
  void set(int r, int c, const M& m) { blocks_.push_back({{r, c}, Block{m}}); }
  void clear() { blocks_.clear(); }

 private:
  std::vector<std::pair<std::pair<int, int>, Block>> blocks_;
};
}  // namespace g2o

namespace ps_graph_slam {

class GraphSLAM {
 public:
  // graph_slam.cpp:40-97
  explicit GraphSLAM(bool verbose) : verbose_(verbose) {
    std::cout << "construct solver... " << std::endl;
    ssb_graph_opts o;
    ssb_graph_default_opts(&o);
    o.verbose = verbose ? 1 : 0;
    o.preconditioner = 2;
    graph = ssb_graph_create(&o);
    if (!graph) {
      std::cerr << std::endl << "error : failed to allocate solver!! " << ssb_last_error() << std::endl;
      return;
    }
    std::cout << "done" << std::endl;
  }
  ~GraphSLAM() { ssb_graph_destroy(graph); }  // graph_slam.cpp:102
  GraphSLAM(const GraphSLAM&) = delete;
  GraphSLAM& operator=(const GraphSLAM&) = delete;

  // graph_slam.cpp:104-115 (first vertex fixed)
  g2o::VertexSE3* add_se3_node(const ssb_host::Isometry3d& pose) {
    double T[12];
    ssb_host::to34(pose, T);
    int id = ssb_graph_add_se3_node(graph, T);
    se3_.emplace_back(new g2o::VertexSE3(graph, id));
    return se3_.back().get();
  }
  // graph_slam.cpp:127-134
  g2o::VertexPointXYZ* add_point_xyz_node(const ssb_host::Vector3d& xyz) {
    double p[3] = {xyz(0), xyz(1), xyz(2)};
    int id = ssb_graph_add_point_xyz_node(graph, p);
    xyz_.emplace_back(new g2o::VertexPointXYZ(graph, id));
    return xyz_.back().get();
  }
  // graph_slam.cpp:136-148
  g2o::EdgeSE3* add_se3_edge(g2o::VertexSE3* v1, g2o::VertexSE3* v2, const ssb_host::Isometry3d& relative_pose,
                             const ssb_host::MatrixXd& information_matrix) {
    double Z[12], I[36];
    ssb_host::to34(relative_pose, Z);
    ssb_host::toRowMajor(information_matrix, 6, I);
    int id = ssb_graph_add_se3_edge(graph, v1->id(), v2->id(), Z, I);
    edges_.emplace_back(new g2o::EdgeHandle{id});
    return edges_.back().get();
  }
  // graph_slam.cpp:150-166
  g2o::EdgeSE3PointXYZ* add_se3_point_xyz_edge(g2o::VertexSE3* v_se3, g2o::VertexPointXYZ* v_xyz,
                                               const ssb_host::Vector3d& xyz,
                                               const ssb_host::MatrixXd& information_matrix) {
    double z[3] = {xyz(0), xyz(1), xyz(2)}, I[9];
    ssb_host::toRowMajor(information_matrix, 3, I);
    int id = ssb_graph_add_se3_point_xyz_edge(graph, v_se3->id(), v_xyz->id(), z, I);
    edges_.emplace_back(new g2o::EdgeHandle{id});
    return edges_.back().get();
  }
  // graph_slam.cpp:168-180
  // graph_slam.hpp:44 / graph_slam.cpp:117-125 (commented out in the reference): plane landmark
  g2o::VertexPlane* add_plane_node(const ssb_host::Vector4d& plane_coeffs) {
    double c[4] = {plane_coeffs(0), plane_coeffs(1), plane_coeffs(2), plane_coeffs(3)};
    int id = ssb_graph_add_plane_node(graph, c);
    if (id < 0) return nullptr;
    planes_.emplace_back(new g2o::VertexPlane(graph, id));
    return planes_.back().get();
  }
  // graph_slam.hpp:74-75 (commented out in the reference): g2o::EdgeSE3Plane (include/g2o/edge_se3_plane.hpp)
  g2o::EdgeSE3Plane* add_se3_plane_edge(g2o::VertexSE3* v_se3, g2o::VertexPlane* v_plane,
                                        const ssb_host::Vector4d& plane_coeffs,
                                        const ssb_host::MatrixXd& information_matrix) {
    double c[4] = {plane_coeffs(0), plane_coeffs(1), plane_coeffs(2), plane_coeffs(3)}, I[9];
    ssb_host::toRowMajor(information_matrix, 3, I);
    int id = ssb_graph_add_se3_plane_edge(graph, v_se3->id(), v_plane->id(), c, I);
    if (id < 0) return nullptr;
    edges_.emplace_back(new g2o::EdgeHandle{id});
    return edges_.back().get();
  }
  g2o::EdgePointXYZ* add_point_xyz_point_xyz_edge(g2o::VertexPointXYZ* v1_xyz, g2o::VertexPointXYZ* v2_xyz,
                                                  const ssb_host::Vector3d& xyz,
                                                  const ssb_host::MatrixXd& information_matrix) {
    double z[3] = {xyz(0), xyz(1), xyz(2)}, I[9];
    ssb_host::toRowMajor(information_matrix, 3, I);
    int id = ssb_graph_add_point_xyz_point_xyz_edge(graph, v1_xyz->id(), v2_xyz->id(), z, I);
    edges_.emplace_back(new g2o::EdgeHandle{id});
    return edges_.back().get();
  }

  // graph_slam.cpp:182-219
  bool optimize() {
    if (ssb_graph_num_edges(graph) < 10) return false;
    if (verbose_) {
      std::cout << std::endl << "--- pose graph optimization ---" << std::endl;
      std::cout << "nodes: " << ssb_graph_num_vertices(graph) << "   edges: " << ssb_graph_num_edges(graph) << std::endl;
      std::cout << "optimizing... " << std::flush;
    }
    ssb_lm_stats st;
    int r = ssb_graph_optimize(graph, 1024, &st);
    if (r < 0) {
      std::cerr << "optimize failed: " << ssb_last_error() << std::endl;
      return false;
    }
    if (verbose_) {
      std::cout << "done" << std::endl;
      std::cout << "iterations: " << st.iterations << std::endl;
      std::cout << "chi2: (before)" << st.chi2_initial << " -> (after)" << st.chi2_final << std::endl;
      char buf[64];
      std::snprintf(buf, sizeof(buf), "%.3f", st.ms_total * 1e-3);
      std::cout << "time: " << buf << "[sec]" << std::endl;
    }
    return r == 1;
  }

  // graph_slam.cpp:221-234: vert_pairs are (hessianIndex, hessianIndex) of landmark vertices
  bool computeLandmarkMarginals(g2o::SparseBlockMatrix<ssb_host::MatrixXd>& spinv,
                                std::vector<std::pair<int, int>> vert_pairs_vec) {
    std::unordered_map<int, int> vid_of_hidx;   // one pass over the landmark vertices instead of a scan per pair
    vid_of_hidx.reserve(xyz_.size());
    for (auto& v : xyz_) vid_of_hidx.emplace(v->hessianIndex(), v->id());
    std::vector<int> vids;
    vids.reserve(vert_pairs_vec.size());
    for (auto& pr : vert_pairs_vec) {
      auto it = vid_of_hidx.find(pr.first);
      if (it == vid_of_hidx.end()) return false;
      vids.push_back(it->second);
    }
    std::vector<double> out(9 * vids.size());
    if (ssb_graph_landmark_marginals(graph, vids.data(), (int)vids.size(), out.data()) != 1) {
      if (verbose_) std::cout << "not computing marginals " << std::endl;
      return false;
    }
    spinv.clear();
    for (size_t k = 0; k < vids.size(); ++k) {
      ssb_host::MatrixXd M(3, 3);   // the (rows, cols) form also compiles when MatrixXd is Eigen::MatrixXd
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) M(r, c) = out[9 * k + 3 * r + c];
      spinv.set(vert_pairs_vec[k].first, vert_pairs_vec[k].second, M);
    }
    if (verbose_) std::cout << "computed marginals " << std::endl;
    return true;
  }

  // graph_slam.cpp:236-239
  void save(const std::string& filename) { ssb_graph_save_g2o(graph, filename.c_str()); }

 public:
  ssb_graph* graph = nullptr;  // stands in for std::shared_ptr<g2o::SparseOptimizer> (graph_slam.hpp:147)
  bool verbose_;

 private:
  std::vector<std::unique_ptr<g2o::VertexSE3>> se3_;
  std::vector<std::unique_ptr<g2o::VertexPointXYZ>> xyz_;
  std::vector<std::unique_ptr<g2o::VertexPlane>> planes_;
  std::vector<std::unique_ptr<g2o::EdgeHandle>> edges_;
};

}  // namespace ps_graph_slam
#endif  // SSB_GRAPH_SLAM_HPP
