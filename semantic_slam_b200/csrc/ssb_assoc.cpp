// ssb_assoc.cpp — host side of the per-frame landmark association step behind the C-ABI of include/ssb.h.
//
// Restates (does not copy) the float arithmetic of the reference's
//   data_association::find_matches / associate_lanmarks / map_a_new_lan / inserst_a_mapped_lan
//     (/root/reference/include/ps_graph_slam/data_association.h:75-389) and
//   semantic_tools::transformNormalsToWorld / transformPoseFromCameraToRobot / dist
//     (/root/reference/include/tools.h:18-135,293-297)
// including the quirks a faithful replacement has to reproduce:
//   H4  distance_min / distance / nearest id live outside the per-detection loop and are never reset
//       (opts.strict = 1 selects the sane per-detection reset instead),
//   H6  4-vectors are (x, y, z, 1),
//   H7  T_robot_world(0,2) = cy*sp*cr + sy*sp  (the textbook term is sy*sr).
// The step is host code in the reference as well (a handful of detections against a few hundred
// landmarks per frame); everything is single precision, evaluated in Eigen's order with separately
// rounded multiplies and adds (this file is compiled with -ffp-contract=off), so that the association
// indices are bit-exact against the oracle restatement.
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

#include "../../include/ssb.h"

namespace ssb {
void set_error(const char* fmt, ...);
}

namespace {

struct Mat4 {
  float m[16];  // row-major
};
Mat4 zero4() {
  Mat4 r;
  for (float& x : r.m) x = 0.0f;
  return r;
}
// Eigen 4x4 float product: every entry is a0*b0, then (+ a1*b1), (+ a2*b2), (+ a3*b3), each op rounded
Mat4 mul4(const Mat4& A, const Mat4& B) {
  Mat4 R;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      float acc = A.m[4 * i] * B.m[j];
      for (int k = 1; k < 4; ++k) acc = acc + A.m[4 * i + k] * B.m[4 * k + j];
      R.m[4 * i + j] = acc;
    }
  return R;
}
void mulv4(const Mat4& A, const float* v, float* out) {
  for (int i = 0; i < 4; ++i) {
    float acc = A.m[4 * i] * v[0];
    for (int k = 1; k < 4; ++k) acc = acc + A.m[4 * i + k] * v[k];
    out[i] = acc;
  }
}
// tools.h:104-135 (rot_z_robot * rot_x_robot * rot_x_cam)
void fixed_rotations(float cam_angle, Mat4& rot_x_cam, Mat4& rot_x_robot, Mat4& rot_z_robot) {
  rot_x_cam = zero4();
  rot_x_robot = zero4();
  rot_z_robot = zero4();
  rot_x_cam.m[0] = 1;
  rot_x_cam.m[5] = std::cos(-cam_angle);       // float overloads: cosf / sinf
  rot_x_cam.m[6] = -std::sin(-cam_angle);
  rot_x_cam.m[9] = std::sin(-cam_angle);
  rot_x_cam.m[10] = std::cos(-cam_angle);
  rot_x_cam.m[15] = 1;
  rot_x_robot.m[0] = 1;                         // rotation of -90 deg, double literals rounded to float
  rot_x_robot.m[5] = (float)std::cos(-1.5708);
  rot_x_robot.m[6] = (float)(-std::sin(-1.5708));
  rot_x_robot.m[9] = (float)std::sin(-1.5708);
  rot_x_robot.m[10] = (float)std::cos(-1.5708);
  rot_x_robot.m[15] = 1;
  rot_z_robot.m[0] = (float)std::cos(-1.5708);
  rot_z_robot.m[1] = (float)(-std::sin(-1.5708));
  rot_z_robot.m[4] = (float)std::sin(-1.5708);
  rot_z_robot.m[5] = (float)std::cos(-1.5708);
  rot_z_robot.m[10] = 1;
  rot_z_robot.m[15] = 1;
}
// semantic_tools::transformNormalsToWorld  tools.h:18-102
Mat4 transform_normals_to_world(const float* pose6, float cam_angle) {
  Mat4 rxc, rxr, rzr, T = zero4();
  fixed_rotations(cam_angle, rxc, rxr, rzr);
  const float roll = pose6[3], pitch = pose6[4], yaw = pose6[5];
  const float cy = std::cos(yaw), sy = std::sin(yaw), cp = std::cos(pitch), sp = std::sin(pitch), cr = std::cos(roll),
              sr = std::sin(roll);
  T.m[0] = cy * cp;
  T.m[1] = cy * sp * sr - sy * cr;
  T.m[2] = cy * sp * cr + sy * sp;  // H7: as written in the reference (tools.h:80-81)
  T.m[4] = sy * cp;
  T.m[5] = sy * sp * sr + cy * cr;
  T.m[6] = sy * sp * cr - cy * sr;
  T.m[8] = -sp;
  T.m[9] = cp * sr;
  T.m[10] = cp * cr;
  T.m[15] = 1;
  return mul4(mul4(mul4(T, rzr), rxr), rxc);
}
// semantic_tools::transformPoseFromCameraToRobot  tools.h:104-135
Mat4 transform_cam_to_robot(float cam_angle) {
  Mat4 rxc, rxr, rzr;
  fixed_rotations(cam_angle, rxc, rxr, rzr);
  return mul4(mul4(rzr, rxr), rxc);
}
// semantic_tools::dist  tools.h:293-297
float dist3(float x1, float x2, float y1, float y2, float z1, float z2) {
  return std::sqrt((x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1) + (z2 - z1) * (z2 - z1));
}
// Eigen::Matrix3f::inverse() (cofactor expansion along column 0), float
void inv3_cofactor(const float* a, float* r) {
  auto cof = [&](int i, int j) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return a[3 * i1 + j1] * a[3 * i2 + j2] - a[3 * i1 + j2] * a[3 * i2 + j1];
  };
  const float c0 = cof(0, 0), c1 = cof(1, 0), c2 = cof(2, 0);
  const float det = (c0 * a[0] + c1 * a[3]) + c2 * a[6];
  const float invdet = 1.0f / det;
  r[0] = c0 * invdet;
  r[1] = c1 * invdet;
  r[2] = c2 * invdet;
  r[3] = cof(0, 1) * invdet;
  r[4] = cof(1, 1) * invdet;
  r[5] = cof(2, 1) * invdet;
  r[6] = cof(0, 2) * invdet;
  r[7] = cof(1, 2) * invdet;
  r[8] = cof(2, 2) * invdet;
}

// Eigen::MatrixXf::inverse() — what `Q.inverse()` is for the DYNAMIC-size Q of the Mahalanobis gate
// (data_association.h:175-184): compute_inverse<.., Dynamic> = partialPivLu().inverse() = solve(Identity), in float:
//   PartialPivLU::unblocked_lu (sizes <= 16): per column the FIRST largest |entry| is the pivot, whole rows are swapped,
//     the column below the pivot is DIVIDED by it, the trailing block gets  a(i,j) -= l(i) * u(j)  (product rounded, then
//     the difference: no FMA, the reference builds with SSE flags only);
//   X = P * Identity; unit-lower solve, then upper solve, both column-oriented (TriangularSolverMatrix.h, column-major
//     triangle): x_i is MULTIPLIED by the reciprocal of the diagonal (1 for the unit triangle), then  x_r -= x_i * t(r,i)
//     for the rows still to come (below for the lower solve, above for the upper one).
// a, r: 3 x 3 row-major.  A zero pivot column is left undivided like Eigen does (the reciprocal is then inf).
void inv3_partial_piv_lu(const float* a, float* r) {
  float A[3][3], X[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      A[i][j] = a[3 * i + j];
      X[i][j] = i == j ? 1.0f : 0.0f;
    }
  int piv[3];
  for (int k = 0; k < 3; ++k) {
    int best = k;
    float big = std::fabs(A[k][k]);
    for (int i = k + 1; i < 3; ++i)
      if (std::fabs(A[i][k]) > big) {
        big = std::fabs(A[i][k]);
        best = i;
      }
    piv[k] = best;
    if (big != 0.0f) {
      if (best != k)
        for (int j = 0; j < 3; ++j) std::swap(A[k][j], A[best][j]);
      for (int i = k + 1; i < 3; ++i) A[i][k] = A[i][k] / A[k][k];
    }
    for (int i = k + 1; i < 3; ++i)
      for (int j = k + 1; j < 3; ++j) {
        const float prod = A[i][k] * A[k][j];
        A[i][j] = A[i][j] - prod;
      }
  }
  for (int k = 0; k < 3; ++k)
    if (piv[k] != k)
      for (int j = 0; j < 3; ++j) std::swap(X[k][j], X[piv[k]][j]);
  for (int j = 0; j < 3; ++j) {
    // unit lower
    for (int i = 0; i < 3; ++i) {
      const float b = X[i][j] * 1.0f;
      for (int q = i + 1; q < 3; ++q) {
        const float prod = b * A[q][i];
        X[q][j] = X[q][j] - prod;
      }
    }
    // upper, from the last row up
    for (int i = 2; i >= 0; --i) {
      const float inv = 1.0f / A[i][i];
      X[i][j] = X[i][j] * inv;
      const float b = X[i][j];
      for (int q = 0; q < i; ++q) {
        const float prod = b * A[q][i];
        X[q][j] = X[q][j] - prod;
      }
    }
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r[3 * i + j] = X[i][j];
}

struct Landmark {
  int id, type, plane_type;
  float pose[3], local_pose[3], cov[9], normal[4];
  double node_estimate[3];  // l.node->estimate(): refreshed by the caller after every optimize
};

}  // namespace

struct ssb_assoc {
  ssb_assoc_opts o;
  bool first_object = true;
  std::vector<Landmark> landmarks;
  float Q[9];
};

static void world_pose(const ssb_assoc* a, const float* robot_pose, float cam_angle, const float* v4, float* out4) {
  // data_association::convertPoseToWorld  :320-343
  mulv4(transform_normals_to_world(robot_pose, cam_angle), v4, out4);
  out4[0] += robot_pose[0];
  if (!a->o.use_rtab_map_odom)
    out4[1] += robot_pose[1];
  else
    out4[1] = (float)((double)out4[1] + ((double)robot_pose[1] - 0.04));
  out4[2] += robot_pose[2];
}

static void fill_observation(const ssb_assoc* a, const ssb_detection& d, const float* robot_pose, float cam_angle,
                             ssb_landmark_obs* out) {
  const float cam[4] = {d.pose[0], d.pose[1], d.pose[2], 1.0f};  // H6
  float w[4], n[4], r[4];
  world_pose(a, robot_pose, cam_angle, cam, w);
  mulv4(transform_normals_to_world(robot_pose, cam_angle), d.normal, n);  // convertNormalsToWorld :345-359
  mulv4(transform_cam_to_robot(cam_angle), cam, r);                       // convertCamToRobot :361-373
  for (int k = 0; k < 3; ++k) {
    out->local_pose[k] = r[k];
    out->pose[k] = w[k];
  }
  for (int k = 0; k < 4; ++k) out->normal[k] = n[k];
  std::memcpy(out->covariance, a->Q, sizeof(a->Q));
  float inf[9];
  inv3_cofactor(a->Q, inf);  // semantic_graph_slam.cpp:170: information = covariance.inverse()
  for (int k = 0; k < 9; ++k) out->information[k] = (double)inf[k];
  out->type = d.type;
  out->plane_type = d.plane_type;
}

static void map_new(ssb_assoc* a, const ssb_detection& d, const float* robot_pose, float cam_angle, ssb_landmark_obs* out) {
  // data_association::map_a_new_lan  :237-276
  fill_observation(a, d, robot_pose, cam_angle, out);
  out->is_new_landmark = 1;
  out->id = (int)a->landmarks.size();
  Landmark l;
  l.id = out->id;
  l.type = d.type;
  l.plane_type = d.plane_type;
  std::memcpy(l.pose, out->pose, sizeof(l.pose));
  std::memcpy(l.local_pose, out->local_pose, sizeof(l.local_pose));
  std::memcpy(l.cov, a->Q, sizeof(l.cov));
  std::memcpy(l.normal, out->normal, sizeof(l.normal));
  for (int k = 0; k < 3; ++k) l.node_estimate[k] = (double)out->pose[k];  // the node is created from l.pose
  a->landmarks.push_back(l);
}

extern "C" {

void ssb_assoc_default_opts(ssb_assoc_opts* o) {
  if (!o) return;
  std::memset(o, 0, sizeof(*o));
  o->maha_dist_thres = 0.5;   // data_association.h:47-53 defaults
  o->eq_dist_thres = 1.21;
  o->land_noise_low = 0.5;
  o->land_noise_high = 0.9;
  o->use_maha_dist = 1;
  o->use_eq_dist = 0;
  o->use_rtab_map_odom = 0;
  o->strict = 0;
}

ssb_assoc* ssb_assoc_create(const ssb_assoc_opts* opts) {
  ssb_assoc* a = new ssb_assoc();
  if (opts)
    a->o = *opts;
  else
    ssb_assoc_default_opts(&a->o);
  for (float& q : a->Q) q = 0.0f;
  a->Q[0] = a->Q[4] = a->Q[8] = (float)a->o.land_noise_low;  // :64-66
  return a;
}
void ssb_assoc_destroy(ssb_assoc* a) { delete a; }
int ssb_assoc_num_landmarks(const ssb_assoc* a) { return a ? (int)a->landmarks.size() : SSB_ERR_INVALID; }

int ssb_assoc_set_landmark_estimate(ssb_assoc* a, int id, const double xyz[3]) {
  if (!a || !xyz || id < 0 || id >= (int)a->landmarks.size()) return SSB_ERR_INVALID;
  std::memcpy(a->landmarks[id].node_estimate, xyz, 3 * sizeof(double));
  return SSB_OK;
}
int ssb_assoc_set_landmark_cov(ssb_assoc* a, int id, const float cov[9]) {
  if (!a || !cov || id < 0 || id >= (int)a->landmarks.size()) return SSB_ERR_INVALID;
  std::memcpy(a->landmarks[id].cov, cov, 9 * sizeof(float));
  return SSB_OK;
}
int ssb_assoc_inverse3(int kind, const float a[9], float r[9]) {
  if (!a || !r || (kind != 0 && kind != 1)) return SSB_ERR_INVALID;
  if (kind == 0)
    inv3_cofactor(a, r);
  else
    inv3_partial_piv_lu(a, r);
  return SSB_OK;
}

int ssb_assoc_get_landmark(const ssb_assoc* a, int id, ssb_landmark_obs* out) {
  if (!a || !out || id < 0 || id >= (int)a->landmarks.size()) return SSB_ERR_INVALID;
  const Landmark& l = a->landmarks[id];
  std::memset(out, 0, sizeof(*out));
  out->id = l.id;
  out->type = l.type;
  out->plane_type = l.plane_type;
  std::memcpy(out->pose, l.pose, sizeof(l.pose));
  std::memcpy(out->local_pose, l.local_pose, sizeof(l.local_pose));
  std::memcpy(out->covariance, l.cov, sizeof(l.cov));
  std::memcpy(out->normal, l.normal, sizeof(l.normal));
  return SSB_OK;
}

// plane_segmentation::multiPlaneSegmentation's per-region post-processing (plane_segmentation.cpp:117-132,160-255:
// gates, horizontal / vertical classification against gravity in the camera frame, normal sign conventions) followed
// by point_cloud_segmentation::segmentPlanarSurfaces (point_cloud_segmentation.h:26-103: camera -> world, the
// detected_object record).  The regions themselves come from a plane extractor (PCL's organised multi-plane
// segmentation in the reference; ssb_ransac_plane_batch here).
int ssb_segment_planar_surfaces(const ssb_planar_region* regions, int n, const float robot_pose[6], float cam_angle,
                                int object_type, float prob, float planar_area, ssb_detected_object* out) {
  if ((n > 0 && (!regions || !out)) || !robot_pose || n < 0) {
    ssb::set_error("ssb_segment_planar_surfaces: invalid argument");
    return SSB_ERR_INVALID;
  }
  const Mat4 M = transform_normals_to_world(robot_pose, cam_angle);
  // normals_of_the_horizontal_plane_in_cam = transformation_mat^T * (0, 0, 1, 0)  (:126-132)
  const float nh[3] = {M.m[8], M.m[9], M.m[10]};
  int m = 0;
  for (int i = 0; i < n; ++i) {
    const ssb_planar_region& R = regions[i];
    if (!(R.contour_points > 100)) continue;                 // :169
    float dot = 0;                                           // computeDotProduct :547-555
    for (int k = 0; k < 3; ++k) dot = dot + nh[k] * R.model[k];
    if (!(R.area >= planar_area)) continue;                  // :195
    int flag;
    bool flip;
    if ((double)(std::fabs(R.model[0]) - std::fabs(nh[0])) < 0.3 && (double)(std::fabs(R.model[1]) - std::fabs(nh[1])) < 0.3 &&
        (double)(std::fabs(R.model[2]) - std::fabs(nh[2])) < 0.3) {
      flag = 0;                 // horizontal (:197-224); normals upwards
      flip = R.model[1] > 0;
    } else if ((double)dot < 0.5) {
      flag = 1;                 // vertical (:226-250); normals to the left
      flip = R.model[0] > 0;
    } else {
      continue;
    }
    ssb_detected_object& o = out[m++];
    std::memset(&o, 0, sizeof(o));
    const float cam[4] = {R.centroid[0], R.centroid[1], R.centroid[2], 1.0f};
    float w[4];
    mulv4(M, cam, w);           // point_cloud_segmentation.h:56-57
    o.type = object_type;
    o.plane_type = flag;        // 0 "horizontal", 1 "vertical" (:79-82)
    o.prob = prob;
    o.num_points = (float)R.contour_points;
    for (int k = 0; k < 3; ++k) {
      o.pose[k] = cam[k];
      o.world_pose[k] = w[k] + robot_pose[k];   // :91-94
    }
    for (int k = 0; k < 4; ++k) o.normal_orientation[k] = flip ? -R.model[k] : R.model[k];
  }
  return m;
}

int ssb_assoc_find_matches(ssb_assoc* a, const ssb_detection* dets, int n, const float robot_pose[6], float cam_angle,
                           ssb_landmark_obs* out) {
  if (!a || (n > 0 && (!dets || !out)) || !robot_pose || n < 0) {
    ssb::set_error("ssb_assoc_find_matches: invalid argument");
    return SSB_ERR_INVALID;
  }
  // data_association::find_matches  :75-95
  if (a->first_object) {
    for (int j = 0; j < n; ++j) map_new(a, dets[j], robot_pose, cam_angle, out + j);
    if (n > 0) a->first_object = false;
    return n;
  }
  // data_association::associate_lanmarks  :97-235
  bool found = false;
  float distance = 0;
  float distance_min = std::numeric_limits<float>::max();
  int nearest = 0;  // declared (uninitialised) inside the loop in the reference; its stack slot persists (H4)
  for (int j = 0; j < n; ++j) {
    if (a->o.strict) {
      distance_min = std::numeric_limits<float>::max();
      nearest = 0;
    }
    const float cam[4] = {dets[j].pose[0], dets[j].pose[1], dets[j].pose[2], 1.0f};
    float actual[4];
    world_pose(a, robot_pose, cam_angle, cam, actual);
    const size_t nl = a->landmarks.size();
    for (size_t i = 0; i < nl; ++i) {
      const Landmark& l = a->landmarks[i];
      if (dets[j].type != l.type || dets[j].plane_type != l.plane_type) continue;
      found = true;
      const float expected[3] = {(float)l.node_estimate[0], (float)l.node_estimate[1], (float)l.node_estimate[2]};
      if (a->o.use_maha_dist) {
        // Q = H sigma H' + Q_ with H = I; distance = z' Q^-1 z on the xyz components (H5)
        float Q[9], Qi[9], z[3];
        for (int k = 0; k < 9; ++k) Q[k] = l.cov[k] + a->Q[k];
        inv3_partial_piv_lu(Q, Qi);   // Q is a dynamic MatrixXf in the reference: LU, not the fixed-size cofactor formula
        for (int k = 0; k < 3; ++k) z[k] = actual[k] - expected[k];
        float t[3];
        for (int c = 0; c < 3; ++c) t[c] = (z[0] * Qi[c] + z[1] * Qi[3 + c]) + z[2] * Qi[6 + c];
        distance = (t[0] * z[0] + t[1] * z[1]) + t[2] * z[2];
      } else if (a->o.use_eq_dist) {
        distance = dist3(actual[0], expected[0], actual[1], expected[1], actual[2], expected[2]);
      }
      if (distance < distance_min) {
        distance_min = distance;
        nearest = (int)i;
      }
    }
    if (!found) {
      map_new(a, dets[j], robot_pose, cam_angle, out + j);
    } else {
      found = false;
      bool is_new = false, emit = false;
      if (a->o.use_maha_dist) {
        is_new = (double)distance_min > a->o.maha_dist_thres;
        emit = true;
      } else if (a->o.use_eq_dist) {
        is_new = (double)distance_min > a->o.eq_dist_thres;
        emit = true;
      }
      if (!emit) {
        // neither gate enabled: the reference pushes nothing for this detection
        std::memset(out + j, 0, sizeof(out[j]));
        out[j].id = -1;
        continue;
      }
      if (is_new) {
        map_new(a, dets[j], robot_pose, cam_angle, out + j);
      } else {
        // data_association::inserst_a_mapped_lan  :278-318
        fill_observation(a, dets[j], robot_pose, cam_angle, out + j);
        out[j].is_new_landmark = 0;
        out[j].id = a->landmarks[nearest].id;
      }
    }
  }
  return n;
}

}  // extern "C"
