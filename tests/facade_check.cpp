// Compile-and-run check of the C++ facade.  Built by tests/test_facade.py with g++ against include/ssb.h and libssb.so,
// twice: with the POD twins (Eigen is absent from this image) and with -DSSB_USE_EIGEN -I tests/mock_eigen, a mock that
// has Eigen's call syntax and compile-time restrictions, so the branch a maintainer of the reference would compile is
// exercised too.
#include <cmath>
#include <cstdio>
#include "ps_graph_slam/graph_slam.hpp"
#include "planar_segmentation/plane_segmentation_b200.h"
#include "planar_segmentation/point_cloud_segmentation.h"
#include "ps_graph_slam/data_association_b200.h"

// the only places where the two value-type families differ: construction and the translation accessor
#ifdef SSB_HAVE_EIGEN
static ssb_host::Isometry3d iso_x(double x) {
  ssb_host::Isometry3d T = ssb_host::Isometry3d::Identity();
  T.matrix()(0, 3) = x;
  return T;
}
static double x_of(const ssb_host::Isometry3d& T) { return T.matrix()(0, 3); }
static ssb_host::Vector3d v3(double x, double y, double z) { return ssb_host::Vector3d(x, y, z); }
static ssb_host::Vector4d v4(double a, double b, double c, double d) { return ssb_host::Vector4d(a, b, c, d); }
static ssb_host::MatrixXd eye(int n) { return ssb_host::MatrixXd::Identity(n, n); }
#else
static ssb_host::Isometry3d iso_x(double x) {
  ssb_host::Isometry3d T = ssb_host::Isometry3d::Identity();
  T.m[3] = x;
  return T;
}
static double x_of(const ssb_host::Isometry3d& T) { return T.m[3]; }
static ssb_host::Vector3d v3(double x, double y, double z) { return ssb_host::Vector3d{{x, y, z}}; }
static ssb_host::Vector4d v4(double a, double b, double c, double d) { return ssb_host::Vector4d{{a, b, c, d}}; }
static ssb_host::MatrixXd eye(int n) { return ssb_host::MatrixXd::Identity(n); }
#endif

int main(int argc, char** argv) {
  const bool run = argc > 1;  // without arguments: construct nothing (no GPU needed), just prove it links
  {
    // association facade (host only): first frame maps, second frame matches
    ssb_assoc_opts ao;
    ssb_assoc_default_opts(&ao);
    ao.use_maha_dist = 0;
    ao.use_eq_dist = 1;
    ao.eq_dist_thres = 1.5;
    ssb_host::data_association da(false, &ao);
    ssb_host::detected_object d;
    d.pose[2] = 4.0f;
    const float rp[6] = {0, 0, 0, 0, 0, 0};
    auto first = da.find_matches({d}, rp, 0.0f);
    if (first.size() != 1 || !first[0].is_new_landmark) return 7;
    da.assignLandmarkNode(0, [&](double p[3]) { for (int k = 0; k < 3; ++k) p[k] = first[0].pose[k]; });
    d.pose[0] = 0.1f;
    auto second = da.find_matches({d}, rp, 0.0f);
    if (second.size() != 1 || second[0].is_new_landmark || second[0].id != 0) return 8;
  }
  if (!run) {
    std::printf("facade linked: %s\n", ssb_build_info());
    return 0;
  }
  {
    // the live segmentation facade under the reference's names: a synthetic 640x480 frame with a tilted wall in front of a
    // far background, one accepted and one ignored detection class
    const int W = 640, H = 480;
    std::vector<float> msg((size_t)W * H * 8, 0.0f);
    for (int v = 0; v < H; ++v)
      for (int u = 0; u < W; ++u) {
        const float dx = (u - 319.5f) / 525.0f, dy = (v - 239.5f) / 525.0f;
        const bool wall = u >= 200 && u < 420 && v >= 120 && v < 360;
        const float z = wall ? 1.2f / (1.0f + 0.3f * dx) : 4.0f;
        float* p = &msg[((size_t)v * W + u) * 8];
        p[0] = dx * z;
        p[1] = dy * z;
        p[2] = z;
      }
    point_cloud_segmentation seg(false, 500, 5000, 0.0f);
    if (!seg.ok()) return 9;
    ssb_host::PointCloud2View pc;
    pc.data = msg.data();
    pc.layout = ssb_cloud_layout{W, H, 32, 32 * W, 0, 4, 8, 16};
    ssb_host::ObjectInfo a, b;
    a.type = "chair";
    a.prob = 0.9f;
    a.tl_x = 150; a.tl_y = 80; a.width = 320; a.height = 320;
    b = a;
    b.type = "person";   // not in the class list of point_cloud_segmentation.h:126-130
    const float rp[6] = {1.0f, 2.0f, 0.5f, 0.0f, 0.0f, 0.0f};
    auto objs = seg.segmentallPointCloudData(rp, 0.0f, {a, b}, pc);
    if (objs.empty()) return 10;
    for (auto& o : objs)
      if (o.type != "chair" || (o.plane_type != "horizontal" && o.plane_type != "vertical")) return 11;
    std::printf("segmentallPointCloudData: %zu planar surface(s), first at z = %.3f (%s)\n", objs.size(), objs[0].pose[2],
                objs[0].plane_type.c_str());
  }
  if (run) {
    // the dormant clustering chain under the reference's names: k-means of two blobs, hull of a noisy square patch
    plane_segmentation_b200 ps(false);
    if (!ps.ok()) return 12;
    std::vector<float> pts;
    for (int i = 0; i < 400; ++i) {
      pts.push_back(i % 2 ? 1.0f + 0.001f * (i % 7) : -1.0f - 0.001f * (i % 5));
      pts.push_back(0.002f * (i % 11));
      pts.push_back(0.5f);
    }
    std::vector<int> labels;
    std::vector<float> cen;
    const double comp = ps.computeKmeans(pts, 3, 2, labels, cen);
    if (comp < 0 || labels.size() != 400 || labels[0] == labels[1] || labels[0] != labels[2]) return 13;
    std::vector<float> patch;
    for (int r = 0; r < 40; ++r)
      for (int c = 0; c < 40; ++c) {
        patch.push_back(0.01f * c);
        patch.push_back(0.01f * r);
        patch.push_back(1.0f + 0.0005f * ((r * 7 + c * 3) % 5));
        patch.push_back(0.f);
      }
    std::vector<float> hull = ps.compute2DConvexHull(patch);
    if (hull.size() < 12) return 14;
    std::printf("computeKmeans: compactness %.6f; compute2DConvexHull: %zu hull vertices\n", comp, hull.size() / 3);
  }
  ps_graph_slam::GraphSLAM gs(false);
  if (!gs.graph) return 2;
  using namespace ssb_host;
  std::vector<g2o::VertexSE3*> kf;
  for (int k = 0; k < 12; ++k) {
    kf.push_back(gs.add_se3_node(iso_x(0.5 * k + (k ? 0.03 : 0.0))));
    if (k) gs.add_se3_edge(kf[k - 1], kf[k], iso_x(0.5), eye(6));
  }
  g2o::VertexPointXYZ* lm = gs.add_point_xyz_node(v3(2.0, 1.0, 0.5));
  for (int k = 0; k < 12; ++k) gs.add_se3_point_xyz_edge(kf[k], lm, v3(2.0 - 0.5 * k, 1.0, 0.5), eye(3));
  // plane landmark (the reference's dormant VertexPlane / EdgeSE3Plane API): the floor z = 0 seen from every keyframe
  Vector4d floor_w = v4(0.0, 0.0, 1.0, 0.0);
  g2o::VertexPlane* fl = gs.add_plane_node(floor_w);
  if (!fl) return 5;
  for (int k = 0; k < 12; ++k) gs.add_se3_plane_edge(kf[k], fl, floor_w, eye(3));
  if (!gs.optimize()) return 3;
  Vector4d fe = fl->estimate();
  if (std::fabs(fe(2) - 1.0) > 1e-6 || std::fabs(fe(3)) > 1e-6) return 6;
  // semantic_graph_slam.cpp:181-205: marginals asked by (hessianIndex, hessianIndex), read through block(i,i)->eval()
  lm->unlockQuadraticForm();
  g2o::SparseBlockMatrix<MatrixXd> spinv;
  const int hi = lm->hessianIndex();
  if (!gs.computeLandmarkMarginals(spinv, {{hi, hi}})) return 7;
  if (!spinv.block(hi, hi)) return 8;
  const MatrixXd cov = spinv.block(hi, hi)->eval();
  // every column is its own PCG solve (pcg_tol 1e-8): symmetric up to that tolerance, not bit for bit
  const double dmax = std::fmax(cov(0, 0), std::fmax(cov(1, 1), cov(2, 2)));
  if (!(cov(0, 0) > 0 && cov(1, 1) > 0 && cov(2, 2) > 0) || std::fabs(cov(0, 1) - cov(1, 0)) > 1e-5 * dmax) {
    std::printf("landmark covariance: [%g %g %g; %g %g %g; %g %g %g]\n", cov(0, 0), cov(0, 1), cov(0, 2), cov(1, 0), cov(1, 1), cov(1, 2),
                cov(2, 0), cov(2, 1), cov(2, 2));
    return 9;
  }
  Isometry3d last = kf.back()->estimate();
  std::printf("last keyframe x = %.6f (expect 5.5), hessianIndex(first)=%d id(last)=%d, landmark covariance xx = %.3e\n", x_of(last),
              kf[0]->hessianIndex(), kf.back()->id(), cov(0, 0));
#ifdef SSB_HAVE_EIGEN
  std::printf("value types: Eigen call syntax\n");
#else
  std::printf("value types: POD twins\n");
#endif
  return std::fabs(x_of(last) - 5.5) < 1e-6 ? 0 : 4;
}
