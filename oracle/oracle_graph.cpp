// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked, imported or called by the product path
// (semantic_slam_b200/); only tests/, __graft_entry__.smoke() and bench.py's CPU legs use it.
//
// CPU restatement (double precision, single thread by default) of the graph hot path of
// hridaybavle/semantic_slam:  ps_graph_slam::GraphSLAM over g2o "lm_var"
//   reference call sites:  src/ps_graph_slam/graph_slam.cpp:40-239
//     add_se3_node :104-115 (first vertex ever added is fixed), add_point_xyz_node :127-134,
//     add_se3_edge :136-148, add_se3_point_xyz_edge :150-166 (robust kernel = uninitialised
//     pointer => none), add_point_xyz_point_xyz_edge :168-180,
//     optimize :182-219 (|E|<10 => false; initializeOptimization; optimize(1024)),
//     computeLandmarkMarginals :221-234.
// The arithmetic lives in g2o, which is NOT vendored under /root/reference (distro package
// ros-kinetic/melodic-libg2o, see README.md:39-44).  Its published algorithm is restated here:
//   core/optimization_algorithm_levenberg.cpp  (solve, computeLambdaInit tau=1e-5, computeScale,
//        goodStep scale clamp [1/3,2/3], ni doubling, <=10 trials, Terminate on rho==0)
//   core/block_solver.hpp (buildSystem, setLambda/restoreDiagonal, no Schur: nothing marginalised)
//   core/base_binary_edge.hpp (constructQuadraticForm)
//   core/sparse_optimizer.cpp (buildIndexMapping: hessian index = id order, fixed = -1;
//        push/pop/discardTop; update)
//   solvers/csparse/linear_solver_csparse.h + CSparse cs_chol (up-looking sparse Cholesky)
//   types/slam3d/{edge_se3,edge_se3_pointxyz,edge_pointxyz,vertex_se3,vertex_pointxyz,
//        isometry3d_mappings,isometry3d_gradients,dquat2mat}.{h,cpp}
// PARITY UNPINNED: the reference ships no tests / golden vectors for this path; the
// oracle is pinned instead by finite differences, a dense numpy LM and scipy sparse solves
// (tests/test_oracle_graph.py).
//
// Build: see oracle/Makefile (g++ -O3 -march=native -ffp-contract=off -shared -fPIC).

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <queue>
#include <string>
#include <unordered_map>
#include <vector>
#include <chrono>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace orc {

// ------------------------------------------------------------------------------------------
// Small fixed-size algebra (row-major).
// ------------------------------------------------------------------------------------------
struct Iso {
  double R[9];
  double t[3];
};

static inline void mat3_mul(const double* A, const double* B, double* C) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
static inline void mat3_T(const double* A, double* At) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) At[3 * i + j] = A[3 * j + i];
}
static inline void mat3_vec(const double* A, const double* v, double* o) {
  for (int i = 0; i < 3; ++i) o[i] = A[3 * i] * v[0] + A[3 * i + 1] * v[1] + A[3 * i + 2] * v[2];
}
static inline Iso iso_mul(const Iso& a, const Iso& b) {
  Iso c;
  mat3_mul(a.R, b.R, c.R);
  double rt[3];
  mat3_vec(a.R, b.t, rt);
  for (int i = 0; i < 3; ++i) c.t[i] = rt[i] + a.t[i];
  return c;
}
static inline Iso iso_inv(const Iso& a) {
  Iso c;
  mat3_T(a.R, c.R);
  double rt[3];
  mat3_vec(c.R, a.t, rt);
  for (int i = 0; i < 3; ++i) c.t[i] = -rt[i];
  return c;
}
static inline Iso iso_identity() {
  Iso c;
  std::memset(&c, 0, sizeof(c));
  c.R[0] = c.R[4] = c.R[8] = 1.0;
  return c;
}

// Eigen::Quaternion(Matrix3) restated (Shepperd branches), coefficient order (x,y,z,w).
// `branch` returns 3 for the trace>0 case, else the index i of the dominant diagonal.
static inline int quat_from_R(const double* m, double q[4]) {
  double t = m[0] + m[4] + m[8];
  if (t > 0.0) {
    t = std::sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (m[7] - m[5]) * t;
    q[1] = (m[2] - m[6]) * t;
    q[2] = (m[3] - m[1]) * t;
    return 3;
  }
  int i = 0;
  if (m[4] > m[0]) i = 1;
  if (m[8] > m[4 * i]) i = 2;
  int j = (i + 1) % 3, k = (j + 1) % 3;
  t = std::sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1.0);
  q[i] = 0.5 * t;
  t = 0.5 / t;
  q[3] = (m[3 * k + j] - m[3 * j + k]) * t;
  q[j] = (m[3 * j + i] + m[3 * i + j]) * t;
  q[k] = (m[3 * k + i] + m[3 * i + k]) * t;
  return i;
}

// g2o internal::toCompactQuaternion: Quaterniond(R), normalize, make w >= 0, return (x,y,z).
static inline void to_compact_quat(const double* R, double v[3], double* w_out = nullptr) {
  double q[4];
  quat_from_R(R, q);
  double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; ++i) q[i] /= n;
  if (q[3] < 0)
    for (int i = 0; i < 4; ++i) q[i] = -q[i];
  v[0] = q[0];
  v[1] = q[1];
  v[2] = q[2];
  if (w_out) *w_out = q[3];
}

// Eigen Quaternion::toRotationMatrix restated.
static inline void quat_to_R(double x, double y, double z, double w, double* R) {
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz);
  R[1] = txy - twz;
  R[2] = txz + twy;
  R[3] = txy + twz;
  R[4] = 1 - (txx + tzz);
  R[5] = tyz - twx;
  R[6] = txz - twy;
  R[7] = tyz + twx;
  R[8] = 1 - (txx + tyy);
}

// g2o internal::fromVectorMQT: t = v[0:3], R = fromCompactQuaternion(v[3:6]) (identity if |q|^2>1).
static inline Iso from_vector_mqt(const double* v) {
  Iso T = iso_identity();
  double w = 1.0 - (v[3] * v[3] + v[4] * v[4] + v[5] * v[5]);
  if (w >= 0) {
    w = std::sqrt(w);
    quat_to_R(v[3], v[4], v[5], w, T.R);
  }
  T.t[0] = v[0];
  T.t[1] = v[1];
  T.t[2] = v[2];
  return T;
}
static inline void to_vector_mqt(const Iso& T, double* v) {
  v[0] = T.t[0];
  v[1] = T.t[1];
  v[2] = T.t[2];
  to_compact_quat(T.R, v + 3);
}

// g2o internal::approximateNearestOrthogonalMatrix:  R -= 0.5 * R * (R^T R - I)
static inline void approx_nearest_orthogonal(double* R) {
  double Rt[9], E[9], RE[9];
  mat3_T(R, Rt);
  mat3_mul(Rt, R, E);
  E[0] -= 1;
  E[4] -= 1;
  E[8] -= 1;
  mat3_mul(R, E, RE);
  for (int i = 0; i < 9; ++i) R[i] -= 0.5 * RE[i];
}

// d(qx,qy,qz)/d(R entries) for the (sign-normalised) compact quaternion, g2o dquat2mat.cpp
// (compute_dq_dR and its four branch helpers) restated as the analytic derivative of the
// branch formula Eigen selects.  dq is 3x9 row-major; column index c = 3*row + col of R.
static inline void compute_dq_dR(const double* m, double dq[27]) {
  std::memset(dq, 0, 27 * sizeof(double));
  double q[4];
  int br = quat_from_R(m, q);
  auto D = [&](int comp, int r, int c) -> double& { return dq[9 * comp + 3 * r + c]; };
  if (br == 3) {
    // qw = S/4 with S = 2 sqrt(1+tr);  q_x = (m21-m12)/S ...; dS/dm_kk = 2/S
    double S = 4.0 * q[3];
    double iS = 1.0 / S;
    double num[3] = {m[7] - m[5], m[2] - m[6], m[3] - m[1]};
    // off-diagonal terms
    D(0, 2, 1) = iS;
    D(0, 1, 2) = -iS;
    D(1, 0, 2) = iS;
    D(1, 2, 0) = -iS;
    D(2, 1, 0) = iS;
    D(2, 0, 1) = -iS;
    for (int c = 0; c < 3; ++c)
      for (int k = 0; k < 3; ++k) D(c, k, k) = -num[c] * iS * iS * (2.0 * iS);
  } else {
    int i = br, j = (i + 1) % 3, k = (j + 1) % 3;
    // q_i = S/4, S = 2 sqrt(1 + m_ii - m_jj - m_kk); q_j = (m_ji+m_ij)/S; q_k = (m_ki+m_ik)/S
    double S = 4.0 * q[i];
    double iS = 1.0 / S;
    double dS[3];  // dS/dm_ii, dm_jj, dm_kk
    dS[0] = 2.0 * iS;
    dS[1] = -2.0 * iS;
    dS[2] = -2.0 * iS;
    int diag[3] = {i, j, k};
    for (int a = 0; a < 3; ++a) D(i, diag[a], diag[a]) = 0.25 * dS[a];
    double nj = m[3 * j + i] + m[3 * i + j], nk = m[3 * k + i] + m[3 * i + k];
    D(j, j, i) = iS;
    D(j, i, j) = iS;
    D(k, k, i) = iS;
    D(k, i, k) = iS;
    for (int a = 0; a < 3; ++a) {
      D(j, diag[a], diag[a]) += -nj * iS * iS * dS[a];
      D(k, diag[a], diag[a]) += -nk * iS * iS * dS[a];
    }
  }
  // sign normalisation (w >= 0) of toCompactQuaternion
  if (q[3] < 0)
    for (int a = 0; a < 27; ++a) dq[a] = -dq[a];
}

// g2o internal::skew / skewT helpers.  skew(S,t): S = 2*[t]x ; skewT(S,t): S = (2*[t]x)^T.
static inline void skew2(const double* t, double* S) {
  S[0] = 0;
  S[1] = -2 * t[2];
  S[2] = 2 * t[1];
  S[3] = 2 * t[2];
  S[4] = 0;
  S[5] = -2 * t[0];
  S[6] = -2 * t[1];
  S[7] = 2 * t[0];
  S[8] = 0;
}

// ------------------------------------------------------------------------------------------
// Graph containers
// ------------------------------------------------------------------------------------------
enum VKind { V_SE3 = 0, V_XYZ = 1, V_PLANE = 2 };
enum EKind { E_SE3 = 0, E_SE3_XYZ = 1, E_XYZ_XYZ = 2, E_SE3_PLANE = 3 };

struct Vertex {
  int kind;
  bool fixed;
  int hidx;        // hessian (block) index, -1 if fixed
  int num_oplus;   // VertexSE3::_numOplusCalls
  Iso T;           // SE3 estimate
  double p[4];     // XYZ estimate, or the 4 normalised plane coefficients (V_PLANE)
  std::vector<Iso> stackT;
  std::vector<double> stackP;
  int dim() const { return kind == V_SE3 ? 6 : 3; }
};

struct Edge {
  int kind;
  int vi, vj;
  Iso Z, Zinv;       // E_SE3 measurement
  double z[4];       // E_SE3_XYZ / E_XYZ_XYZ measurement, or the measured plane (E_SE3_PLANE)
  double info[36];   // row-major DxD
  int D() const { return kind == E_SE3 ? 6 : 3; }
};

struct IterRecord {
  double chi2_before, chi2_after, lambda, rho;
  int trials;
};

struct SparseChol;  // fwd

struct Graph {
  std::vector<Vertex> V;
  std::vector<Edge> E;
  // optimisation state
  std::vector<int> ivmap;          // block index -> vertex id
  std::vector<int> boff;           // scalar offset of each block
  int nscalar = 0;
  // block hessian (upper triangular in block index), built by build_structure
  std::map<std::pair<int, int>, int> blk_index;   // (bi,bj) bi<=bj -> slot
  std::vector<int> blk_bi, blk_bj, blk_ptr;       // slot -> ids and value offset
  std::vector<double> Hval;                       // block values, row-major (dim_i x dim_j)
  std::vector<double> b, x;
  std::vector<IterRecord> history;
  SparseChol* chol = nullptr;
  double last_analyze_ms = 0, last_factor_ms = 0, last_linearize_ms = 0;
  int num_threads = 1;
  ~Graph();
};

// ------------------------------------------------------------------------------------------
// Edge error / Jacobians
// ------------------------------------------------------------------------------------------
// EdgeSE3::computeError: e = toVectorMQT(Z^-1 * Xi^-1 * Xj)
static void edge_se3_error(const Edge& e, const Iso& Xi, const Iso& Xj, double* err) {
  Iso delta = iso_mul(e.Zinv, iso_mul(iso_inv(Xi), Xj));
  to_vector_mqt(delta, err);
}

// internal::computeEdgeSE3Gradient with Pi = Pj = identity. Ji, Jj are 6x6 row-major.
static void edge_se3_jac(const Edge& e, const Iso& Xi, const Iso& Xj, double* Ji, double* Jj) {
  Iso A = e.Zinv;  // Z^-1 * Pi^-1
  Iso B = iso_mul(iso_inv(Xi), Xj);
  Iso AB = iso_mul(A, B);
  // C = identity => BC = B, E = AB
  const double* Re = AB.R;
  const double* Ra = A.R;
  const double* Rab = AB.R;
  const double* Rbc = B.R;
  const double* tbc = B.t;
  double dq[27];
  compute_dq_dR(Re, dq);
  std::memset(Ji, 0, 36 * sizeof(double));
  std::memset(Jj, 0, 36 * sizeof(double));
  // dte/dti = -Ra ; dte/dtj = Rab
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      Ji[6 * r + c] = -Ra[3 * r + c];
      Jj[6 * r + c] = Rab[3 * r + c];
    }
  // dte/dqi = Ra * skewT(tbc)   with skewT(t) = (2[t]x)^T = -2[t]x ... g2o: skewT gives S^T of skew.
  // Derivation (checked against finite differences in tests): d(dR^T (tB - dt))/dv = 2 [tB]x.
  {
    double S[9], RS[9];
    skew2(tbc, S);
    mat3_mul(Ra, S, RS);
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) Ji[6 * r + 3 + c] = RS[3 * r + c];
  }
  // dte/dqj = Rab * skewT(tc) = 0 (tc = 0)
  // dre/dqi = dq_dR * M,  M(9x3): column a = vec of d(Re)/d(v_a) with Re' = Ra * dR^T * Rbc
  //   dR^T ~ I - 2[v]x  => dRe/dv_a = Ra * (-2 [e_a]x) * Rbc
  // dre/dqj: Re' = Rab * dR => dRe/dv_a = Rab * (2 [e_a]x)
  for (int a = 0; a < 3; ++a) {
    double ea[3] = {0, 0, 0};
    ea[a] = 1.0;
    double G[9];
    skew2(ea, G);  // 2[e_a]x
    double T1[9], dRi[9], dRj[9];
    mat3_mul(Ra, G, T1);
    mat3_mul(T1, Rbc, dRi);
    for (int q = 0; q < 9; ++q) dRi[q] = -dRi[q];
    mat3_mul(Rab, G, dRj);
    for (int comp = 0; comp < 3; ++comp) {
      double si = 0, sj = 0;
      for (int q = 0; q < 9; ++q) {
        si += dq[9 * comp + q] * dRi[q];
        sj += dq[9 * comp + q] * dRj[q];
      }
      Ji[6 * (3 + comp) + 3 + a] = si;
      Jj[6 * (3 + comp) + 3 + a] = sj;
    }
  }
}

// EdgeSE3PointXYZ (offset parameter id 0 = identity, graph_slam.cpp:75-83,162)
static void edge_se3_xyz_error(const Edge& e, const Iso& X, const double* p, double* err) {
  Iso w2n = iso_inv(X);
  double pc[3];
  mat3_vec(w2n.R, p, pc);
  for (int i = 0; i < 3; ++i) err[i] = pc[i] + w2n.t[i] - e.z[i];
}
// Ji 3x6, Jj 3x3 row-major
static void edge_se3_xyz_jac(const Edge&, const Iso& X, const double* p, double* Ji, double* Jj) {
  Iso w2l = iso_inv(X);
  double Zc[3];
  mat3_vec(w2l.R, p, Zc);
  for (int i = 0; i < 3; ++i) Zc[i] += w2l.t[i];
  std::memset(Ji, 0, 18 * sizeof(double));
  Ji[0] = Ji[7] = Ji[14] = -1.0;
  Ji[6 * 0 + 4] = -2 * Zc[2];
  Ji[6 * 0 + 5] = 2 * Zc[1];
  Ji[6 * 1 + 3] = 2 * Zc[2];
  Ji[6 * 1 + 5] = -2 * Zc[0];
  Ji[6 * 2 + 3] = -2 * Zc[1];
  Ji[6 * 2 + 4] = 2 * Zc[0];
  for (int i = 0; i < 9; ++i) Jj[i] = w2l.R[i];
}
// EdgePointXYZ: e = (xj - xi) - z ; Ji = -I, Jj = I
static void edge_xyz_xyz_error(const Edge& e, const double* pi, const double* pj, double* err) {
  for (int i = 0; i < 3; ++i) err[i] = (pj[i] - pi[i]) - e.z[i];
}

// ---- g2o::Plane3D (types/slam3d_addons/plane3d.h) and EdgeSE3Plane
// (/root/reference/include/g2o/edge_se3_plane.hpp:8-48, dormant in the reference) ----
static inline void plane_normalize(double* c) {
  const double n = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
  for (int i = 0; i < 4; ++i) c[i] /= n;
}
// Plane3D::rotation(v) = AngleAxis(azimuth, Z) * AngleAxis(-elevation, Y)
static inline void plane_rotation(const double* v, double* R) {
  const double az = std::atan2(v[1], v[0]);
  const double el = std::atan2(v[2], std::sqrt(v[0] * v[0] + v[1] * v[1]));
  const double ca = std::cos(az), sa = std::sin(az), ce = std::cos(el), se = std::sin(el);
  R[0] = ca * ce; R[1] = -sa; R[2] = -ca * se;
  R[3] = sa * ce; R[4] = ca;  R[5] = -sa * se;
  R[6] = se;      R[7] = 0.0; R[8] = ce;
}
// Plane3D::oplus
static inline void plane_oplus(double* c, const double* v) {
  const double s = std::sin(v[1]), co = std::cos(v[1]);
  const double n[3] = {co * std::cos(v[0]), co * std::sin(v[0]), s};
  double R[9], rn[3];
  plane_rotation(c, R);
  const double d = -c[3] + v[2];
  mat3_vec(R, n, rn);
  c[0] = rn[0]; c[1] = rn[1]; c[2] = rn[2];
  c[3] = -d;
  plane_normalize(c);
}
// EdgeSE3Plane::computeError: local = X^-1 * plane ; error = local.ominus(measurement)
static void edge_se3_plane_error(const Edge& e, const Iso& X, const double* pl, double* err) {
  Iso w2n = iso_inv(X);
  double v2[4];
  mat3_vec(w2n.R, pl, v2);
  v2[3] = pl[3] - (w2n.t[0] * v2[0] + w2n.t[1] * v2[1] + w2n.t[2] * v2[2]);
  plane_normalize(v2);  // Plane3D(v2)
  double R[9], Rt[9], n[3];
  plane_rotation(v2, R);
  mat3_T(R, Rt);
  mat3_vec(Rt, e.z, n);
  err[0] = std::atan2(n[1], n[0]);
  err[1] = std::atan2(n[2], std::sqrt(n[0] * n[0] + n[1] * n[1]));
  err[2] = (-v2[3]) - (-e.z[3]);
}
// BaseBinaryEdge::linearizeOplus (numeric, central differences, delta = 1e-9) as g2o does for an edge
// that does not override it
static void edge_se3_plane_jac(const Edge& e, const Iso& X, const double* pl, double* Ji, double* Jj) {
  const double delta = 1e-9, scalar = 1.0 / (2 * delta);
  for (int d = 0; d < 6; ++d) {
    double add[6] = {0, 0, 0, 0, 0, 0}, e1[3], e2[3];
    add[d] = delta;
    edge_se3_plane_error(e, iso_mul(X, from_vector_mqt(add)), pl, e1);
    add[d] = -delta;
    edge_se3_plane_error(e, iso_mul(X, from_vector_mqt(add)), pl, e2);
    for (int r = 0; r < 3; ++r) Ji[6 * r + d] = scalar * (e1[r] - e2[r]);
  }
  for (int d = 0; d < 3; ++d) {
    double add[3] = {0, 0, 0}, e1[3], e2[3], q[4];
    add[d] = delta;
    std::memcpy(q, pl, sizeof(q));
    plane_oplus(q, add);
    edge_se3_plane_error(e, X, q, e1);
    add[d] = -delta;
    std::memcpy(q, pl, sizeof(q));
    plane_oplus(q, add);
    edge_se3_plane_error(e, X, q, e2);
    for (int r = 0; r < 3; ++r) Jj[3 * r + d] = scalar * (e1[r] - e2[r]);
  }
}

static void edge_error(const Graph& g, const Edge& e, double* err) {
  const Vertex& a = g.V[e.vi];
  const Vertex& b = g.V[e.vj];
  if (e.kind == E_SE3)
    edge_se3_error(e, a.T, b.T, err);
  else if (e.kind == E_SE3_XYZ)
    edge_se3_xyz_error(e, a.T, b.p, err);
  else if (e.kind == E_SE3_PLANE)
    edge_se3_plane_error(e, a.T, b.p, err);
  else
    edge_xyz_xyz_error(e, a.p, b.p, err);
}
static void edge_jac(const Graph& g, const Edge& e, double* Ji, double* Jj) {
  const Vertex& a = g.V[e.vi];
  const Vertex& b = g.V[e.vj];
  if (e.kind == E_SE3)
    edge_se3_jac(e, a.T, b.T, Ji, Jj);
  else if (e.kind == E_SE3_XYZ)
    edge_se3_xyz_jac(e, a.T, b.p, Ji, Jj);
  else if (e.kind == E_SE3_PLANE)
    edge_se3_plane_jac(e, a.T, b.p, Ji, Jj);
  else {
    std::memset(Ji, 0, 9 * sizeof(double));
    std::memset(Jj, 0, 9 * sizeof(double));
    Ji[0] = Ji[4] = Ji[8] = -1;
    Jj[0] = Jj[4] = Jj[8] = 1;
  }
}

static double edge_chi2(const Edge& e, const double* err) {
  int D = e.D();
  double s = 0;
  for (int r = 0; r < D; ++r) {
    double t = 0;
    for (int c = 0; c < D; ++c) t += e.info[D * r + c] * err[c];
    s += err[r] * t;
  }
  return s;
}

// SparseOptimizer::computeActiveErrors + activeRobustChi2 (no robust kernels)
static double active_chi2(const Graph& g) {
  double chi = 0;
  const int ne = (int)g.E.size();
#ifdef _OPENMP
#pragma omp parallel for reduction(+ : chi) num_threads(g.num_threads) if (g.num_threads > 1)
#endif
  for (int k = 0; k < ne; ++k) {
    double err[6];
    edge_error(g, g.E[k], err);
    chi += edge_chi2(g.E[k], err);
  }
  return chi;
}

// ------------------------------------------------------------------------------------------
// Sparse Cholesky (CSparse cs_schol/cs_chol restated: up-looking, scalar, with a block
// minimum-degree ordering standing in for cs_amd)
// ------------------------------------------------------------------------------------------
struct SparseChol {
  int n = 0;
  std::vector<int> perm_blk;   // new block position -> old block
  std::vector<int> pinv_s;     // old scalar -> new scalar
  // permuted upper-triangular CSC of A
  std::vector<int> Cp, Ci;
  std::vector<double> Cx;
  std::vector<int> parent;
  std::vector<int> Lp, Li;
  std::vector<double> Lx;
  std::vector<int> lnz_next;  // fill pointer per column
  // per block-slot mapping into Cx: for slot s, column c (0..dimj-1) start offset and transposed flag
  std::vector<int> slot_colstart;  // slot -> index into slot_cols
  std::vector<int> slot_cols;      // for each (slot, local col) start position in Cx of local row 0
  std::vector<char> slot_transposed;
  std::vector<int> diag_pos;       // scalar (new index) -> position of diagonal entry in Cx
  long long lnz = 0;
  double flops = 0;
};

Graph::~Graph() { delete chol; }

// exact minimum degree on the block graph (explicit elimination graph)
static std::vector<int> min_degree_order(int nb, const std::vector<std::vector<int>>& adj0) {
  std::vector<std::vector<int>> adj = adj0;
  for (auto& a : adj) {
    std::sort(a.begin(), a.end());
    a.erase(std::unique(a.begin(), a.end()), a.end());
  }
  std::vector<char> done(nb, 0);
  typedef std::pair<int, int> PI;
  std::priority_queue<PI, std::vector<PI>, std::greater<PI>> pq;
  for (int i = 0; i < nb; ++i) pq.push(PI((int)adj[i].size(), i));
  std::vector<int> order;
  order.reserve(nb);
  std::vector<int> merged;
  while (!pq.empty()) {
    PI top = pq.top();
    pq.pop();
    int v = top.second;
    if (done[v] || top.first != (int)adj[v].size()) continue;
    done[v] = 1;
    order.push_back(v);
    const std::vector<int>& nv = adj[v];
    for (int u : nv) {
      // adj[u] = (adj[u] U nv) \ {u, v}
      merged.clear();
      merged.reserve(adj[u].size() + nv.size());
      std::set_union(adj[u].begin(), adj[u].end(), nv.begin(), nv.end(), std::back_inserter(merged));
      std::vector<int>& au = adj[u];
      au.clear();
      for (int w : merged)
        if (w != u && w != v) au.push_back(w);
      pq.push(PI((int)au.size(), u));
    }
    std::vector<int>().swap(adj[v]);
  }
  return order;
}

static void chol_analyze(Graph& g) {
  delete g.chol;
  g.chol = new SparseChol();
  SparseChol& S = *g.chol;
  const int nb = (int)g.ivmap.size();
  S.n = g.nscalar;
  // block adjacency
  std::vector<std::vector<int>> adj(nb);
  for (size_t s = 0; s < g.blk_bi.size(); ++s) {
    int a = g.blk_bi[s], b = g.blk_bj[s];
    if (a != b) {
      adj[a].push_back(b);
      adj[b].push_back(a);
    }
  }
  S.perm_blk = min_degree_order(nb, adj);
  std::vector<int> pinv_blk(nb);
  for (int k = 0; k < nb; ++k) pinv_blk[S.perm_blk[k]] = k;
  // new scalar offsets
  std::vector<int> noff(nb + 1, 0);
  for (int k = 0; k < nb; ++k) noff[k + 1] = noff[k] + g.V[g.ivmap[S.perm_blk[k]]].dim();
  S.pinv_s.resize(S.n);
  for (int ob = 0; ob < nb; ++ob) {
    int d = g.V[g.ivmap[ob]].dim();
    for (int r = 0; r < d; ++r) S.pinv_s[g.boff[ob] + r] = noff[pinv_blk[ob]] + r;
  }
  // per new block column: list of (new block row, slot, transposed)
  struct Ent {
    int brow, slot;
    char tr;
  };
  std::vector<std::vector<Ent>> cols(nb);
  for (size_t s = 0; s < g.blk_bi.size(); ++s) {
    int pa = pinv_blk[g.blk_bi[s]], pb = pinv_blk[g.blk_bj[s]];
    if (pa <= pb)
      cols[pb].push_back({pa, (int)s, 0});
    else
      cols[pa].push_back({pb, (int)s, 1});
  }
  S.Cp.assign(S.n + 1, 0);
  S.slot_colstart.assign(g.blk_bi.size(), 0);
  S.slot_transposed.assign(g.blk_bi.size(), 0);
  S.slot_cols.clear();
  S.diag_pos.assign(S.n, 0);
  // first pass: sizes
  std::vector<int> dimnew(nb);
  for (int k = 0; k < nb; ++k) dimnew[k] = noff[k + 1] - noff[k];
  for (int pb = 0; pb < nb; ++pb) {
    std::sort(cols[pb].begin(), cols[pb].end(), [](const Ent& x, const Ent& y) { return x.brow < y.brow; });
    int above = 0;
    for (auto& e : cols[pb])
      if (e.brow != pb) above += dimnew[e.brow];
    for (int c = 0; c < dimnew[pb]; ++c) S.Cp[noff[pb] + c + 1] = above + (c + 1);
  }
  for (int i = 0; i < S.n; ++i) S.Cp[i + 1] += S.Cp[i];
  S.Ci.assign(S.Cp[S.n], 0);
  S.Cx.assign(S.Cp[S.n], 0.0);
  for (int pb = 0; pb < nb; ++pb) {
    int rowpos = 0;
    for (auto& e : cols[pb]) {
      S.slot_colstart[e.slot] = (int)S.slot_cols.size();
      S.slot_transposed[e.slot] = e.tr;
      int dr = dimnew[e.brow];
      for (int c = 0; c < dimnew[pb]; ++c) {
        int base = S.Cp[noff[pb] + c] + rowpos;
        S.slot_cols.push_back(base);
        int nr = (e.brow == pb) ? (c + 1) : dr;
        for (int r = 0; r < nr; ++r) S.Ci[base + r] = noff[e.brow] + r;
        if (e.brow == pb) S.diag_pos[noff[pb] + c] = base + c;
      }
      rowpos += dr;
    }
  }
  // elimination tree + column counts (cs_etree + ereach-based counting)
  const int n = S.n;
  S.parent.assign(n, -1);
  {
    std::vector<int> anc(n, -1);
    for (int k = 0; k < n; ++k) {
      for (int p = S.Cp[k]; p < S.Cp[k + 1]; ++p) {
        int i = S.Ci[p];
        while (i != -1 && i < k) {
          int inext = anc[i];
          anc[i] = k;
          if (inext == -1) S.parent[i] = k;
          i = inext;
        }
      }
    }
  }
  std::vector<int> cnt(n, 1), mark(n, -1);
  for (int k = 0; k < n; ++k) {
    mark[k] = k;
    for (int p = S.Cp[k]; p < S.Cp[k + 1]; ++p) {
      int i = S.Ci[p];
      while (i < k && mark[i] != k) {
        mark[i] = k;
        cnt[i]++;
        i = S.parent[i];
      }
    }
  }
  S.Lp.assign(n + 1, 0);
  for (int i = 0; i < n; ++i) S.Lp[i + 1] = S.Lp[i] + cnt[i];
  S.lnz = S.Lp[n];
  S.Li.assign(S.lnz, 0);
  S.Lx.assign(S.lnz, 0.0);
  S.lnz_next.assign(n, 0);
  S.flops = 0;
  for (int i = 0; i < n; ++i) S.flops += (double)cnt[i] * cnt[i];
}

// copy block hessian (+lambda on the diagonal) into the permuted CSC
static void chol_fill(Graph& g, double lambda) {
  SparseChol& S = *g.chol;
  std::fill(S.Cx.begin(), S.Cx.end(), 0.0);
  for (size_t s = 0; s < g.blk_bi.size(); ++s) {
    int bi = g.blk_bi[s], bj = g.blk_bj[s];
    int di = g.V[g.ivmap[bi]].dim(), dj = g.V[g.ivmap[bj]].dim();
    const double* H = &g.Hval[g.blk_ptr[s]];
    const int* colbase = &S.slot_cols[S.slot_colstart[s]];
    if (bi == bj) {
      for (int c = 0; c < dj; ++c)
        for (int r = 0; r <= c; ++r) S.Cx[colbase[c] + r] = H[di * r + c];  // symmetric; upper part
    } else if (!S.slot_transposed[s]) {
      for (int c = 0; c < dj; ++c)
        for (int r = 0; r < di; ++r) S.Cx[colbase[c] + r] = H[dj * r + c];
    } else {
      // stored block is (bi,bj) but in permuted order bj comes first: write H^T, columns = di
      for (int c = 0; c < di; ++c)
        for (int r = 0; r < dj; ++r) S.Cx[colbase[c] + r] = H[dj * c + r];
    }
  }
  for (int i = 0; i < S.n; ++i) S.Cx[S.diag_pos[i]] += lambda;
}

// cs_chol (up-looking).  Returns false if not positive definite.
static bool chol_factor(Graph& g) {
  SparseChol& S = *g.chol;
  const int n = S.n;
  std::vector<double> x(n, 0.0);
  std::vector<int> stack(n), flag(n, -1);
  for (int k = 0; k < n; ++k) S.lnz_next[k] = S.Lp[k];
  for (int k = 0; k < n; ++k) {
    // ereach: nonzero pattern of row k of L, in topological order in stack[top..n-1]
    int top = n;
    flag[k] = k;
    double d = 0;
    for (int p = S.Cp[k]; p < S.Cp[k + 1]; ++p) {
      int i = S.Ci[p];
      if (i > k) continue;
      if (i == k) {
        d = S.Cx[p];
        continue;
      }
      x[i] = S.Cx[p];
      int len = 0;
      // walk up etree
      int ii = i;
      // temp path stored at the bottom of stack
      while (flag[ii] != k) {
        stack[len++] = ii;
        flag[ii] = k;
        ii = S.parent[ii];
      }
      while (len > 0) stack[--top] = stack[--len];
    }
    for (; top < n; ++top) {
      int i = stack[top];
      double lki = x[i] / S.Lx[S.Lp[i]];
      x[i] = 0;
      int pend = S.lnz_next[i];
      for (int p = S.Lp[i] + 1; p < pend; ++p) x[S.Li[p]] -= S.Lx[p] * lki;
      d -= lki * lki;
      int p = S.lnz_next[i]++;
      S.Li[p] = k;
      S.Lx[p] = lki;
    }
    if (!(d > 0)) return false;
    int p = S.lnz_next[k]++;
    S.Li[p] = k;
    S.Lx[p] = std::sqrt(d);
  }
  return true;
}

// solve (P A P^T) y = P b using L L^T; xout in original ordering
static void chol_solve(const Graph& g, const double* bvec, double* xout) {
  const SparseChol& S = *g.chol;
  const int n = S.n;
  std::vector<double> y(n);
  for (int i = 0; i < n; ++i) y[S.pinv_s[i]] = bvec[i];
  for (int j = 0; j < n; ++j) {
    y[j] /= S.Lx[S.Lp[j]];
    for (int p = S.Lp[j] + 1; p < S.Lp[j + 1]; ++p) y[S.Li[p]] -= S.Lx[p] * y[j];
  }
  for (int j = n - 1; j >= 0; --j) {
    for (int p = S.Lp[j] + 1; p < S.Lp[j + 1]; ++p) y[j] -= S.Lx[p] * y[S.Li[p]];
    y[j] /= S.Lx[S.Lp[j]];
  }
  for (int i = 0; i < n; ++i) xout[i] = y[S.pinv_s[i]];
}

// ------------------------------------------------------------------------------------------
// initializeOptimization / buildStructure / buildSystem
// ------------------------------------------------------------------------------------------
static void initialize_optimization(Graph& g) {
  g.ivmap.clear();
  g.boff.clear();
  int off = 0;
  for (size_t id = 0; id < g.V.size(); ++id) {
    Vertex& v = g.V[id];
    if (v.fixed) {
      v.hidx = -1;
      continue;
    }
    v.hidx = (int)g.ivmap.size();
    g.ivmap.push_back((int)id);
    g.boff.push_back(off);
    off += v.dim();
  }
  g.nscalar = off;
}

static int get_block(Graph& g, int bi, int bj) {
  auto key = std::make_pair(bi, bj);
  auto it = g.blk_index.find(key);
  if (it != g.blk_index.end()) return it->second;
  int slot = (int)g.blk_bi.size();
  g.blk_index[key] = slot;
  g.blk_bi.push_back(bi);
  g.blk_bj.push_back(bj);
  return slot;
}

static void build_structure(Graph& g) {
  g.blk_index.clear();
  g.blk_bi.clear();
  g.blk_bj.clear();
  for (size_t b = 0; b < g.ivmap.size(); ++b) get_block(g, (int)b, (int)b);
  for (const Edge& e : g.E) {
    int hi = g.V[e.vi].hidx, hj = g.V[e.vj].hidx;
    if (hi >= 0 && hj >= 0) get_block(g, std::min(hi, hj), std::max(hi, hj));
  }
  g.blk_ptr.resize(g.blk_bi.size());
  int ptr = 0;
  for (size_t s = 0; s < g.blk_bi.size(); ++s) {
    g.blk_ptr[s] = ptr;
    ptr += g.V[g.ivmap[g.blk_bi[s]]].dim() * g.V[g.ivmap[g.blk_bj[s]]].dim();
  }
  g.Hval.assign(ptr, 0.0);
  g.b.assign(g.nscalar, 0.0);
  g.x.assign(g.nscalar, 0.0);
}

// BlockSolver::buildSystem: per edge linearizeOplus + constructQuadraticForm
static void build_system(Graph& g) {
  std::fill(g.Hval.begin(), g.Hval.end(), 0.0);
  std::fill(g.b.begin(), g.b.end(), 0.0);
  const int ne = (int)g.E.size();
  // Jacobians/errors are computed edge-parallel when threads > 1; accumulation is kept in edge
  // order (deterministic, identical to single-thread g2o order).
  struct Lin {
    double err[6], Ji[36], Jj[36];
  };
  const int CH = 4096;
  std::vector<Lin> buf(std::min(ne, CH));
  for (int base = 0; base < ne; base += CH) {
    int m = std::min(CH, ne - base);
#ifdef _OPENMP
#pragma omp parallel for num_threads(g.num_threads) if (g.num_threads > 1)
#endif
    for (int k = 0; k < m; ++k) {
      const Edge& e = g.E[base + k];
      edge_error(g, e, buf[k].err);
      edge_jac(g, e, buf[k].Ji, buf[k].Jj);
    }
    for (int k = 0; k < m; ++k) {
      const Edge& e = g.E[base + k];
      const Lin& L = buf[k];
      const int D = e.D();
      const Vertex& va = g.V[e.vi];
      const Vertex& vb = g.V[e.vj];
      const int di = va.dim(), dj = vb.dim();
      const bool fromNF = !va.fixed, toNF = !vb.fixed;
      if (!fromNF && !toNF) continue;
      double omega_r[6];
      for (int r = 0; r < D; ++r) {
        double t = 0;
        for (int c = 0; c < D; ++c) t += e.info[D * r + c] * L.err[c];
        omega_r[r] = -t;
      }
      // AtO = A^T * omega (di x D), BtO
      double AtO[36], BtO[36];
      for (int r = 0; r < di; ++r)
        for (int c = 0; c < D; ++c) {
          double t = 0;
          for (int q = 0; q < D; ++q) t += L.Ji[di * q + r] * e.info[D * q + c];
          AtO[D * r + c] = t;
        }
      for (int r = 0; r < dj; ++r)
        for (int c = 0; c < D; ++c) {
          double t = 0;
          for (int q = 0; q < D; ++q) t += L.Jj[dj * q + r] * e.info[D * q + c];
          BtO[D * r + c] = t;
        }
      if (fromNF) {
        double* bi = &g.b[g.boff[va.hidx]];
        for (int r = 0; r < di; ++r) {
          double t = 0;
          for (int q = 0; q < D; ++q) t += L.Ji[di * q + r] * omega_r[q];
          bi[r] += t;
        }
        double* Hii = &g.Hval[g.blk_ptr[g.blk_index[std::make_pair(va.hidx, va.hidx)]]];
        for (int r = 0; r < di; ++r)
          for (int c = 0; c < di; ++c) {
            double t = 0;
            for (int q = 0; q < D; ++q) t += AtO[D * r + q] * L.Ji[di * q + c];
            Hii[di * r + c] += t;
          }
        if (toNF) {
          int hi = va.hidx, hj = vb.hidx;
          if (hi < hj) {
            double* Hij = &g.Hval[g.blk_ptr[g.blk_index[std::make_pair(hi, hj)]]];
            for (int r = 0; r < di; ++r)
              for (int c = 0; c < dj; ++c) {
                double t = 0;
                for (int q = 0; q < D; ++q) t += AtO[D * r + q] * L.Jj[dj * q + c];
                Hij[dj * r + c] += t;
              }
          } else {
            double* Hji = &g.Hval[g.blk_ptr[g.blk_index[std::make_pair(hj, hi)]]];
            for (int r = 0; r < dj; ++r)
              for (int c = 0; c < di; ++c) {
                double t = 0;
                for (int q = 0; q < D; ++q) t += BtO[D * r + q] * L.Ji[di * q + c];
                Hji[di * r + c] += t;
              }
          }
        }
      }
      if (toNF) {
        double* bj = &g.b[g.boff[vb.hidx]];
        for (int r = 0; r < dj; ++r) {
          double t = 0;
          for (int q = 0; q < D; ++q) t += L.Jj[dj * q + r] * omega_r[q];
          bj[r] += t;
        }
        double* Hjj = &g.Hval[g.blk_ptr[g.blk_index[std::make_pair(vb.hidx, vb.hidx)]]];
        for (int r = 0; r < dj; ++r)
          for (int c = 0; c < dj; ++c) {
            double t = 0;
            for (int q = 0; q < D; ++q) t += BtO[D * r + q] * L.Jj[dj * q + c];
            Hjj[dj * r + c] += t;
          }
      }
    }
  }
}

// vertex oplus / stack
static void vertex_oplus(Vertex& v, const double* upd) {
  if (v.kind == V_SE3) {
    Iso inc = from_vector_mqt(upd);
    v.T = iso_mul(v.T, inc);
    if (++v.num_oplus > 1000) {  // VertexSE3::orthogonalizeAfter
      v.num_oplus = 0;
      approx_nearest_orthogonal(v.T.R);
    }
  } else if (v.kind == V_PLANE) {
    plane_oplus(v.p, upd);   // VertexPlane::oplusImpl
  } else {
    for (int i = 0; i < 3; ++i) v.p[i] += upd[i];
  }
}
static void graph_push(Graph& g) {
  for (int id : g.ivmap) {
    Vertex& v = g.V[id];
    if (v.kind == V_SE3)
      v.stackT.push_back(v.T);
    else {
      for (int i = 0; i < 4; ++i) v.stackP.push_back(v.p[i]);
    }
  }
}
static void graph_pop(Graph& g) {
  for (int id : g.ivmap) {
    Vertex& v = g.V[id];
    if (v.kind == V_SE3) {
      v.T = v.stackT.back();
      v.stackT.pop_back();
    } else {
      size_t n = v.stackP.size();
      for (int i = 0; i < 4; ++i) v.p[i] = v.stackP[n - 4 + i];
      v.stackP.resize(n - 4);
    }
  }
}
static void graph_discard_top(Graph& g) {
  for (int id : g.ivmap) {
    Vertex& v = g.V[id];
    if (v.kind == V_SE3)
      v.stackT.pop_back();
    else
      v.stackP.resize(v.stackP.size() - 4);
  }
}

static inline double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// SparseOptimizer::optimize(max_iters) with OptimizationAlgorithmLevenberg::solve.
// Returns number of iterations performed (g2o returns 0 if the last solve reported !OK — we return
// the count of iterations entered and the terminate flag separately).
static int optimize(Graph& g, int max_iters, int* terminated) {
  g.history.clear();
  *terminated = 0;
  initialize_optimization(g);
  if (g.ivmap.empty()) return -1;
  double currentLambda = 0, ni = 2;
  int it = 0;
  g.last_analyze_ms = g.last_factor_ms = g.last_linearize_ms = 0;
  bool ok = true;
  for (it = 0; it < max_iters && ok; ++it) {
    if (it == 0) {
      build_structure(g);
      double t0 = now_ms();
      chol_analyze(g);  // LinearSolverCSparse: symbolic decomposition on first solve
      g.last_analyze_ms += now_ms() - t0;
    }
    double currentChi = active_chi2(g);
    double tempChi = currentChi;
    double t0 = now_ms();
    build_system(g);
    g.last_linearize_ms += now_ms() - t0;
    if (it == 0) {
      // computeLambdaInit: tau * max |H_jj|
      double maxDiag = 0;
      for (size_t bidx = 0; bidx < g.ivmap.size(); ++bidx) {
        int d = g.V[g.ivmap[bidx]].dim();
        const double* H = &g.Hval[g.blk_ptr[g.blk_index[std::make_pair((int)bidx, (int)bidx)]]];
        for (int j = 0; j < d; ++j) maxDiag = std::max(std::fabs(H[d * j + j]), maxDiag);
      }
      currentLambda = 1e-5 * maxDiag;
      ni = 2;
    }
    double rho = 0;
    int qmax = 0;
    IterRecord rec;
    rec.chi2_before = currentChi;
    do {
      graph_push(g);
      double t1 = now_ms();
      chol_fill(g, currentLambda);  // setLambda(lambda, backup) + fillCCS
      bool ok2 = chol_factor(g);
      if (ok2)
        chol_solve(g, g.b.data(), g.x.data());
      else
        std::fill(g.x.begin(), g.x.end(), 0.0);
      g.last_factor_ms += now_ms() - t1;
      // update
      for (size_t bidx = 0; bidx < g.ivmap.size(); ++bidx) vertex_oplus(g.V[g.ivmap[bidx]], &g.x[g.boff[bidx]]);
      // restoreDiagonal is implicit (Hval never modified)
      tempChi = active_chi2(g);
      if (!ok2) tempChi = std::numeric_limits<double>::max();
      rho = currentChi - tempChi;
      double scale = 0;
      for (int j = 0; j < g.nscalar; ++j) scale += g.x[j] * (currentLambda * g.x[j] + g.b[j]);
      scale += 1e-3;
      rho /= scale;
      if (rho > 0 && std::isfinite(tempChi)) {
        double alpha = 1.0 - std::pow(2 * rho - 1, 3);
        alpha = std::min(alpha, 2.0 / 3.0);
        double scaleFactor = std::max(1.0 / 3.0, alpha);
        currentLambda *= scaleFactor;
        ni = 2;
        currentChi = tempChi;
        graph_discard_top(g);
      } else {
        currentLambda *= ni;
        ni *= 2;
        graph_pop(g);
      }
      qmax++;
    } while (rho < 0 && qmax < 10);
    rec.chi2_after = currentChi;
    rec.lambda = currentLambda;
    rec.rho = rho;
    rec.trials = qmax;
    g.history.push_back(rec);
    if (qmax == 10 || rho == 0) {
      *terminated = 1;
      ok = false;
    }
  }
  return it;
}

}  // namespace orc

// ------------------------------------------------------------------------------------------
// C API (ctypes).  T/Z are row-major 3x4 [R|t].
// ------------------------------------------------------------------------------------------
using namespace orc;

static Iso iso_from_3x4(const double* T) {
  Iso r;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) r.R[3 * i + j] = T[4 * i + j];
    r.t[i] = T[4 * i + 3];
  }
  return r;
}
static void iso_to_3x4(const Iso& r, double* T) {
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) T[4 * i + j] = r.R[3 * i + j];
    T[4 * i + 3] = r.t[i];
  }
}

extern "C" {

void* orc_graph_create() { return new Graph(); }
void orc_graph_destroy(void* h) { delete (Graph*)h; }
void orc_graph_set_threads(void* h, int n) { ((Graph*)h)->num_threads = n < 1 ? 1 : n; }

int orc_graph_add_se3_node(void* h, const double* T) {
  Graph& g = *(Graph*)h;
  Vertex v;
  v.kind = V_SE3;
  v.fixed = g.V.empty();  // graph_slam.cpp:109-111
  v.hidx = -1;
  v.num_oplus = 0;
  v.T = iso_from_3x4(T);
  v.p[0] = v.p[1] = v.p[2] = v.p[3] = 0;
  g.V.push_back(v);
  return (int)g.V.size() - 1;
}
int orc_graph_add_point_xyz_node(void* h, const double* p) {
  Graph& g = *(Graph*)h;
  Vertex v;
  v.kind = V_XYZ;
  v.fixed = false;  // graph_slam.cpp:127-134 never fixes a landmark
  v.hidx = -1;
  v.num_oplus = 0;
  v.T = iso_identity();
  v.p[0] = p[0];
  v.p[1] = p[1];
  v.p[2] = p[2];
  v.p[3] = 0;
  g.V.push_back(v);
  return (int)g.V.size() - 1;
}
static int check_v(Graph& g, int id, int kind) { return id >= 0 && id < (int)g.V.size() && g.V[id].kind == kind; }

int orc_graph_add_se3_edge(void* h, int vi, int vj, const double* Z, const double* info36) {
  Graph& g = *(Graph*)h;
  if (!check_v(g, vi, V_SE3) || !check_v(g, vj, V_SE3)) return -1;
  Edge e;
  e.kind = E_SE3;
  e.vi = vi;
  e.vj = vj;
  e.Z = iso_from_3x4(Z);
  e.Zinv = iso_inv(e.Z);
  std::memcpy(e.info, info36, 36 * sizeof(double));
  g.E.push_back(e);
  return (int)g.E.size() - 1;
}
int orc_graph_add_se3_point_xyz_edge(void* h, int vp, int vl, const double* z, const double* info9) {
  Graph& g = *(Graph*)h;
  if (!check_v(g, vp, V_SE3) || !check_v(g, vl, V_XYZ)) return -1;
  Edge e;
  e.kind = E_SE3_XYZ;
  e.vi = vp;
  e.vj = vl;
  e.Z = e.Zinv = iso_identity();
  std::memcpy(e.z, z, 3 * sizeof(double));
  std::memset(e.info, 0, sizeof(e.info));
  std::memcpy(e.info, info9, 9 * sizeof(double));
  g.E.push_back(e);
  return (int)g.E.size() - 1;
}
int orc_graph_add_point_xyz_point_xyz_edge(void* h, int v1, int v2, const double* z, const double* info9) {
  Graph& g = *(Graph*)h;
  if (!check_v(g, v1, V_XYZ) || !check_v(g, v2, V_XYZ)) return -1;
  Edge e;
  e.kind = E_XYZ_XYZ;
  e.vi = v1;
  e.vj = v2;
  e.Z = e.Zinv = iso_identity();
  std::memcpy(e.z, z, 3 * sizeof(double));
  std::memset(e.info, 0, sizeof(e.info));
  std::memcpy(e.info, info9, 9 * sizeof(double));
  g.E.push_back(e);
  return (int)g.E.size() - 1;
}
int orc_graph_num_vertices(void* h) { return (int)((Graph*)h)->V.size(); }
int orc_graph_num_edges(void* h) { return (int)((Graph*)h)->E.size(); }

int orc_graph_get_se3(void* h, int id, double* T) {
  Graph& g = *(Graph*)h;
  if (!check_v(g, id, V_SE3)) return -1;
  iso_to_3x4(g.V[id].T, T);
  return 0;
}
int orc_graph_set_se3(void* h, int id, const double* T) {
  Graph& g = *(Graph*)h;
  if (!check_v(g, id, V_SE3)) return -1;
  g.V[id].T = iso_from_3x4(T);
  return 0;
}
// commented-out API of the reference (graph_slam.hpp:44,74-75): VertexPlane / EdgeSE3Plane
int orc_graph_add_plane_node(void* h, const double* c4) {
  Graph& g = *(Graph*)h;
  Vertex v;
  v.kind = V_PLANE;
  v.fixed = false;
  v.hidx = -1;
  v.num_oplus = 0;
  v.T = iso_identity();
  std::memcpy(v.p, c4, 4 * sizeof(double));
  plane_normalize(v.p);  // Plane3D(const Vector4D&)
  g.V.push_back(v);
  return (int)g.V.size() - 1;
}
int orc_graph_add_se3_plane_edge(void* h, int vp, int vl, const double* plane4, const double* info9) {
  Graph& g = *(Graph*)h;
  if (!check_v(g, vp, V_SE3) || !check_v(g, vl, V_PLANE)) return -1;
  Edge e;
  e.kind = E_SE3_PLANE;
  e.vi = vp;
  e.vj = vl;
  e.Z = e.Zinv = iso_identity();
  std::memcpy(e.z, plane4, 4 * sizeof(double));
  plane_normalize(e.z);  // setMeasurement(Plane3D(v))
  std::memset(e.info, 0, sizeof(e.info));
  std::memcpy(e.info, info9, 9 * sizeof(double));
  g.E.push_back(e);
  return (int)g.E.size() - 1;
}
int orc_graph_get_plane(void* h, int id, double* c4) {
  Graph& g = *(Graph*)h;
  if (!check_v(g, id, V_PLANE)) return -1;
  std::memcpy(c4, g.V[id].p, 4 * sizeof(double));
  return 0;
}
int orc_graph_set_plane(void* h, int id, const double* c4) {
  Graph& g = *(Graph*)h;
  if (!check_v(g, id, V_PLANE)) return -1;
  std::memcpy(g.V[id].p, c4, 4 * sizeof(double));
  plane_normalize(g.V[id].p);
  return 0;
}
void orc_plane_oplus(double* c4, const double* v3) { plane_oplus(c4, v3); }

int orc_graph_get_point_xyz(void* h, int id, double* p) {
  Graph& g = *(Graph*)h;
  if (!check_v(g, id, V_XYZ)) return -1;
  std::memcpy(p, g.V[id].p, 3 * sizeof(double));
  return 0;
}
int orc_graph_set_point_xyz(void* h, int id, const double* p) {
  Graph& g = *(Graph*)h;
  if (!check_v(g, id, V_XYZ)) return -1;
  std::memcpy(g.V[id].p, p, 3 * sizeof(double));
  return 0;
}
int orc_graph_set_fixed(void* h, int id, int fixed) {
  Graph& g = *(Graph*)h;
  if (id < 0 || id >= (int)g.V.size()) return -1;
  g.V[id].fixed = fixed != 0;
  return 0;
}
// bulk getters: poses as 3x4 row-major (12 doubles each) for all SE3 vertices in id order, etc.
int orc_graph_get_all(void* h, double* se3_out, double* xyz_out) {
  Graph& g = *(Graph*)h;
  size_t a = 0, b = 0;
  for (auto& v : g.V) {
    if (v.kind == V_SE3) {
      iso_to_3x4(v.T, se3_out + 12 * a);
      ++a;
    } else {
      std::memcpy(xyz_out + 3 * b, v.p, 24);
      ++b;
    }
  }
  return 0;
}

double orc_graph_chi2(void* h) { return active_chi2(*(Graph*)h); }

// optimize: returns 0 if skipped (|E|<10, graph_slam.cpp:184-186), else 1. stats: [iters, terminated]
// history (up to hist_cap records of 5 doubles: chi2_before, chi2_after, lambda, rho, trials)
int orc_graph_optimize(void* h, int max_iters, int* iters_out, int* terminated_out, double* hist, int hist_cap) {
  Graph& g = *(Graph*)h;
  if (g.E.size() < 10) {
    if (iters_out) *iters_out = 0;
    return 0;
  }
  int term = 0;
  int it = optimize(g, max_iters, &term);
  if (iters_out) *iters_out = it;
  if (terminated_out) *terminated_out = term;
  if (hist) {
    int n = std::min((int)g.history.size(), hist_cap);
    for (int k = 0; k < n; ++k) {
      hist[5 * k + 0] = g.history[k].chi2_before;
      hist[5 * k + 1] = g.history[k].chi2_after;
      hist[5 * k + 2] = g.history[k].lambda;
      hist[5 * k + 3] = g.history[k].rho;
      hist[5 * k + 4] = g.history[k].trials;
    }
  }
  return 1;
}
// timing / factor stats of the last optimize: [analyze_ms, factor_ms(total), linearize_ms(total), nnz(L), n]
void orc_graph_last_stats(void* h, double* out5) {
  Graph& g = *(Graph*)h;
  out5[0] = g.last_analyze_ms;
  out5[1] = g.last_factor_ms;
  out5[2] = g.last_linearize_ms;
  out5[3] = g.chol ? (double)g.chol->lnz : 0;
  out5[4] = g.nscalar;
}

// ---- test hooks ----
// per-edge error and Jacobians at the current estimates (Ji: D x di, Jj: D x dj, row-major)
int orc_graph_edge_linearize(void* h, int eid, double* err, double* Ji, double* Jj) {
  Graph& g = *(Graph*)h;
  if (eid < 0 || eid >= (int)g.E.size()) return -1;
  edge_error(g, g.E[eid], err);
  edge_jac(g, g.E[eid], Ji, Jj);
  return 0;
}
// dense H (n x n row-major, full symmetric) and b at the current estimates; returns n. If H is
// null only returns n.  hidx_out (size |V|) gets the scalar offset of each vertex or -1.
int orc_graph_dense_system(void* h, double* H, double* b, int* off_out) {
  Graph& g = *(Graph*)h;
  initialize_optimization(g);
  build_structure(g);
  int n = g.nscalar;
  if (off_out)
    for (size_t id = 0; id < g.V.size(); ++id) off_out[id] = g.V[id].hidx < 0 ? -1 : g.boff[g.V[id].hidx];
  if (!H) return n;
  build_system(g);
  std::fill(H, H + (size_t)n * n, 0.0);
  for (size_t s = 0; s < g.blk_bi.size(); ++s) {
    int bi = g.blk_bi[s], bj = g.blk_bj[s];
    int di = g.V[g.ivmap[bi]].dim(), dj = g.V[g.ivmap[bj]].dim();
    const double* B = &g.Hval[g.blk_ptr[s]];
    for (int r = 0; r < di; ++r)
      for (int c = 0; c < dj; ++c) {
        H[(size_t)(g.boff[bi] + r) * n + g.boff[bj] + c] = B[dj * r + c];
        H[(size_t)(g.boff[bj] + c) * n + g.boff[bi] + r] = B[dj * r + c];
      }
  }
  std::memcpy(b, g.b.data(), n * sizeof(double));
  return n;
}

// sparse export of the block system at the current estimates (test hook): COO triplets of the
// upper-triangular blocks (row, col, val) in scalar indices; returns nnz (call with null to size).
long long orc_graph_sparse_system(void* h, int* rows, int* cols, double* vals, double* b, int* off_out) {
  Graph& g = *(Graph*)h;
  initialize_optimization(g);
  build_structure(g);
  if (off_out)
    for (size_t id = 0; id < g.V.size(); ++id) off_out[id] = g.V[id].hidx < 0 ? -1 : g.boff[g.V[id].hidx];
  long long nnz = (long long)g.Hval.size();
  if (!rows) return nnz;
  build_system(g);
  long long k = 0;
  for (size_t s = 0; s < g.blk_bi.size(); ++s) {
    int bi = g.blk_bi[s], bj = g.blk_bj[s];
    int di = g.V[g.ivmap[bi]].dim(), dj = g.V[g.ivmap[bj]].dim();
    const double* B = &g.Hval[g.blk_ptr[s]];
    for (int r = 0; r < di; ++r)
      for (int c = 0; c < dj; ++c) {
        rows[k] = g.boff[bi] + r;
        cols[k] = g.boff[bj] + c;
        vals[k] = (bi == bj && r > c) ? 0.0 : B[dj * r + c];  // strictly-upper + diagonal only
        ++k;
      }
  }
  std::memcpy(b, g.b.data(), g.nscalar * sizeof(double));
  return nnz;
}
int orc_graph_num_scalar(void* h) {
  Graph& g = *(Graph*)h;
  initialize_optimization(g);
  return g.nscalar;
}
// solve (H + lambda I) x = b with the sparse Cholesky at the current linearisation. returns 1 ok.
int orc_graph_solve_once(void* h, double lambda, double* x) {
  Graph& g = *(Graph*)h;
  initialize_optimization(g);
  build_structure(g);
  chol_analyze(g);
  build_system(g);
  chol_fill(g, lambda);
  if (!chol_factor(g)) return 0;
  chol_solve(g, g.b.data(), x);
  return 1;
}

// Landmark marginals: 3x3 diagonal blocks of H^-1 (H = last built system, undamped), as
// SparseOptimizer::computeMarginals / MarginalCovarianceCholesky would return for the pairs
// (hessianIndex, hessianIndex) (semantic_graph_slam.cpp:181-205).  Re-linearises at the current
// estimate when `relinearize` != 0, else uses the H of the last optimize().
int orc_graph_landmark_marginals(void* h, const int* vids, int n, double* out9n, int relinearize) {
  Graph& g = *(Graph*)h;
  if (relinearize || !g.chol) {
    initialize_optimization(g);
    build_structure(g);
    chol_analyze(g);
    build_system(g);
  }
  chol_fill(g, 0.0);
  if (!chol_factor(g)) return 0;
  std::vector<double> rhs(g.nscalar), sol(g.nscalar);
  for (int k = 0; k < n; ++k) {
    int id = vids[k];
    if (id < 0 || id >= (int)g.V.size() || g.V[id].hidx < 0) return -1;
    int off = g.boff[g.V[id].hidx], d = g.V[id].dim();
    for (int c = 0; c < d; ++c) {
      std::fill(rhs.begin(), rhs.end(), 0.0);
      rhs[off + c] = 1.0;
      chol_solve(g, rhs.data(), sol.data());
      for (int r = 0; r < d; ++r) out9n[(size_t)k * d * d + d * r + c] = sol[off + r];
    }
  }
  return 1;
}

// The same blocks the way g2o computes them: MarginalCovarianceCholesky::computeCovariance
// (core/marginal_covariance_cholesky.cpp), called by LinearSolverCCS::solvePattern through
// BlockSolver::computeMarginals (graph_slam.cpp:225).  No solve per column: every requested element of
// (L L^T)^-1 follows from the entries of its own column of L and elements further down (Takahashi's
// recursion), memoised in a hash map —
//   diag[r] = 1 / L(r,r)
//   s(r,c)  = sum over the off-diagonal entries (rr, L(rr,r)) of column r of  inv(min(rr,c), max(rr,c)) * L(rr,r)
//   inv(r,r) = diag[r] * (diag[r] - s),   inv(r,c) = -s * diag[r]  (r < c)
// with the requested elements sorted like g2o's MatrixElem::operator< (column descending, then row
// descending) so that later requests find most of their dependencies in the map.  g2o recurses; the
// restatement keeps its own stack (the recursion is as deep as the elimination tree is high) and adds
// the terms in the same order.
int orc_graph_landmark_marginals_g2o(void* h, const int* vids, int n, double* out9n, int relinearize, long long* map_entries) {
  Graph& g = *(Graph*)h;
  if (relinearize || !g.chol) {
    initialize_optimization(g);
    build_structure(g);
    chol_analyze(g);
    build_system(g);
  }
  chol_fill(g, 0.0);
  if (!chol_factor(g)) return 0;
  const SparseChol& S = *g.chol;
  const int N = S.n;
  std::vector<double> diag(N);
  for (int r = 0; r < N; ++r) diag[r] = 1.0 / S.Lx[S.Lp[r]];
  struct Elem { int r, c; };
  std::vector<Elem> elems;
  elems.reserve((size_t)n * 9);
  for (int k = 0; k < n; ++k) {
    int id = vids[k];
    if (id < 0 || id >= (int)g.V.size() || g.V[id].hidx < 0) return -1;
    int off = g.boff[g.V[id].hidx], d = g.V[id].dim();
    for (int r = 0; r < d; ++r)
      for (int c = 0; c < d; ++c) {
        int rr = S.pinv_s[off + r], cc = S.pinv_s[off + c];
        if (rr > cc) std::swap(rr, cc);
        elems.push_back({rr, cc});
      }
  }
  std::sort(elems.begin(), elems.end(), [](const Elem& a, const Elem& b) { return a.c > b.c || (a.c == b.c && a.r > b.r); });
  std::unordered_map<unsigned long long, double> inv;
  inv.reserve((size_t)n * 64 + 1024);
  auto key = [N](int r, int c) { return (unsigned long long)r * (unsigned long long)N + (unsigned long long)c; };
  struct Frame { int r, c, p; double s; };
  std::vector<Frame> st;
  for (const Elem& e : elems) {
    if (inv.find(key(e.r, e.c)) != inv.end()) continue;
    st.push_back({e.r, e.c, S.Lp[e.r] + 1, 0.0});
    while (!st.empty()) {
      Frame& f = st.back();
      if (f.p < S.Lp[f.r + 1]) {
        const int rr = S.Li[f.p];
        const int a = rr < f.c ? rr : f.c, b = rr < f.c ? f.c : rr;
        auto it = inv.find(key(a, b));
        if (it != inv.end()) {
          f.s += it->second * S.Lx[f.p];
          ++f.p;
        } else {
          st.push_back({a, b, S.Lp[a] + 1, 0.0});   // f is dangling from here on: re-read st.back() next round
        }
      } else {
        const double res = f.r == f.c ? diag[f.r] * (diag[f.r] - f.s) : -f.s * diag[f.r];
        inv[key(f.r, f.c)] = res;
        st.pop_back();
      }
    }
  }
  for (int k = 0; k < n; ++k) {
    int id = vids[k];
    int off = g.boff[g.V[id].hidx], d = g.V[id].dim();
    for (int r = 0; r < d; ++r)
      for (int c = 0; c < d; ++c) {
        int rr = S.pinv_s[off + r], cc = S.pinv_s[off + c];
        if (rr > cc) std::swap(rr, cc);
        out9n[(size_t)k * d * d + d * r + c] = inv[key(rr, cc)];
      }
  }
  if (map_entries) *map_entries = (long long)inv.size();
  return 1;
}

// SE3 helper hooks for unit tests
void orc_to_vector_mqt(const double* T, double* v6) { to_vector_mqt(iso_from_3x4(T), v6); }
void orc_from_vector_mqt(const double* v6, double* T) { iso_to_3x4(from_vector_mqt(v6), T); }
void orc_se3_oplus(const double* T, const double* v6, double* Tout) {
  iso_to_3x4(iso_mul(iso_from_3x4(T), from_vector_mqt(v6)), Tout);
}

}  // extern "C"
