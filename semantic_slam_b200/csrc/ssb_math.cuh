// ssb_math.cuh — SE3 / quaternion algebra and per-edge residual + Jacobian evaluation for the
// graph hot path (device + host).  Poses are stored as (t, q) with q = (x,y,z,w) unit quaternion;
// increments are g2o's "MQT" minimal vectors (translation, vector part of a unit quaternion), applied
// on the right:  X <- X * fromVectorMQT(d)   (g2o VertexSE3::oplusImpl).
// Closed forms derived here are the exact derivatives g2o obtains through dq/dR
// (types/slam3d/isometry3d_gradients.h); tests compare them with the oracle and finite differences.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#define SSB_HD __host__ __device__ __forceinline__

namespace ssb {

struct Pose {  // 64 B record in HBM: t(3) q(4) pad
  double t[3];
  double q[4];  // x y z w
  double pad;
};

// ---- edge records (AoS, sizes multiples of 16 B so that 1-D TMA bulk copies can stage tiles) ----
struct PLEdge {  // pose -> landmark observation, 80 B  
  int p, l;
  double z[3];
  double info[6];  // upper triangle: 00 01 02 11 12 22
};
struct PPEdge {  // pose -> pose (odometry / loop closure), 240 B  (padded from 232)
  int i, j;
  double zt[3];
  double zq[4];
  double info[21];  // upper triangle row-major
  double pad;
};
static_assert(sizeof(PLEdge) == 80, "PLEdge must be 80 bytes");
static_assert(sizeof(PPEdge) == 240, "PPEdge must be 240 bytes");

SSB_HD void cross3(const double* a, const double* b, double* o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}

// rotation matrix (row-major) of unit quaternion (x,y,z,w) — Eigen::Quaternion::toRotationMatrix
SSB_HD void quat_to_R(const double* q, double* R) {
  const double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
  const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  const double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  const double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
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

// quaternion product a (x) b, (x,y,z,w) storage
SSB_HD void quat_mul(const double* a, const double* b, double* o) {
  o[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  o[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  o[1] = a[3] * b[1] - a[0] * b[2] + a[1] * b[3] + a[2] * b[0];
  o[2] = a[3] * b[2] + a[0] * b[1] - a[1] * b[0] + a[2] * b[3];
}
SSB_HD void quat_conj(const double* a, double* o) {
  o[0] = -a[0];
  o[1] = -a[1];
  o[2] = -a[2];
  o[3] = a[3];
}
SSB_HD void quat_normalize(double* q) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  double s = 1.0 / n;
  q[0] *= s;
  q[1] *= s;
  q[2] *= s;
  q[3] *= s;
}
// R(q)^T v
SSB_HD void rotT(const double* R, const double* v, double* o) {
  o[0] = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
  o[1] = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
  o[2] = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
}
SSB_HD void rot(const double* R, const double* v, double* o) {
  o[0] = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
  o[1] = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
  o[2] = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
}

// Rotation matrix -> quaternion (Eigen's branches), used only when importing 3x4 poses (host).
SSB_HD void R_to_quat(const double* m, double* q) {
  double t = m[0] + m[4] + m[8];
  if (t > 0.0) {
    t = sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (m[7] - m[5]) * t;
    q[1] = (m[2] - m[6]) * t;
    q[2] = (m[3] - m[1]) * t;
  } else {
    int i = 0;
    if (m[4] > m[0]) i = 1;
    if (m[8] > m[4 * i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (m[3 * k + j] - m[3 * j + k]) * t;
    q[j] = (m[3 * j + i] + m[3 * i + j]) * t;
    q[k] = (m[3 * k + i] + m[3 * i + k]) * t;
  }
  quat_normalize(q);
}

// X <- X * fromVectorMQT(d)   (VertexSE3::oplusImpl; rotation = identity when |dq|^2 > 1)
SSB_HD void pose_oplus(Pose& X, const double* d) {
  double R[9], rt[3];
  quat_to_R(X.q, R);
  rot(R, d, rt);
  X.t[0] += rt[0];
  X.t[1] += rt[1];
  X.t[2] += rt[2];
  double w2 = 1.0 - (d[3] * d[3] + d[4] * d[4] + d[5] * d[5]);
  if (w2 >= 0.0) {
    double dq[4] = {d[3], d[4], d[5], sqrt(w2)};
    double o[4];
    quat_mul(X.q, dq, o);
    quat_normalize(o);
    X.q[0] = o[0];
    X.q[1] = o[1];
    X.q[2] = o[2];
    X.q[3] = o[3];
  }
}

SSB_HD void expand_sym3(const double* u, double* M) {  // upper 6 -> full 3x3
  M[0] = u[0];
  M[1] = u[1];
  M[2] = u[2];
  M[3] = u[1];
  M[4] = u[3];
  M[5] = u[4];
  M[6] = u[2];
  M[7] = u[4];
  M[8] = u[5];
}
SSB_HD void expand_sym6(const double* u, double* M) {  // upper 21 -> full 6x6
  int k = 0;
  for (int r = 0; r < 6; ++r)
    for (int c = r; c < 6; ++c) {
      M[6 * r + c] = u[k];
      M[6 * c + r] = u[k];
      ++k;
    }
}

// -------- EdgeSE3PointXYZ (offset = identity):  e = X^-1 p - z ;  Jp = [-I | 2[pc]x] ; Jl = R^T
// (g2o types/slam3d/edge_se3_pointxyz.cpp computeError/linearizeOplus)
struct PLLin {
  double e[3];
  double pc[3];
  double R[9];  // pose rotation (world <- robot)
};
SSB_HD void pl_linearize(const Pose& X, const double* p, const double* z, PLLin& L) {
  quat_to_R(X.q, L.R);
  double d[3] = {p[0] - X.t[0], p[1] - X.t[1], p[2] - X.t[2]};
  rotT(L.R, d, L.pc);
  L.e[0] = L.pc[0] - z[0];
  L.e[1] = L.pc[1] - z[1];
  L.e[2] = L.pc[2] - z[2];
}
// Jp (3x6 row-major) from pc
SSB_HD void pl_jac_pose(const double* pc, double* J) {
  for (int i = 0; i < 18; ++i) J[i] = 0.0;
  J[0] = J[7] = J[14] = -1.0;
  J[6 * 0 + 4] = -2 * pc[2];
  J[6 * 0 + 5] = 2 * pc[1];
  J[6 * 1 + 3] = 2 * pc[2];
  J[6 * 1 + 5] = -2 * pc[0];
  J[6 * 2 + 3] = -2 * pc[1];
  J[6 * 2 + 4] = 2 * pc[0];
}

// -------- EdgeSE3Plane / VertexPlane (dormant in the reference: include/g2o/edge_se3_plane.hpp:8-48,
// graph_slam.hpp:44,74-75; upstream g2o types/slam3d_addons/plane3d.h) ------------------
// A plane is 4 normalised coefficients (n, c3) with |n| = 1 and distance() = -c3; its minimal increment is
// (d azimuth, d elevation, d distance) applied by Plane3D::oplus.  The edge error is
//   e = (X^-1 * plane).ominus(measurement) = (azimuth(n'), elevation(n'), distance - distance_m),
//   n' = rotation(n_local)^T n_m.
// g2o differentiates this edge numerically (central differences, delta 1e-9); here the Jacobians are exact:
// the error is evaluated on forward-mode dual numbers seeded with the 6 pose + 3 plane increments.
template <int N>
struct Dual {
  double v;
  double d[N];
};
template <int N>
SSB_HD Dual<N> dconst(double v) {
  Dual<N> r;
  r.v = v;
  for (int k = 0; k < N; ++k) r.d[k] = 0.0;
  return r;
}
template <int N>
SSB_HD Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r;
  r.v = a.v + b.v;
  for (int k = 0; k < N; ++k) r.d[k] = a.d[k] + b.d[k];
  return r;
}
template <int N>
SSB_HD Dual<N> operator+(const Dual<N>& a, double b) {
  Dual<N> r = a;
  r.v += b;
  return r;
}
template <int N>
SSB_HD Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r;
  r.v = a.v - b.v;
  for (int k = 0; k < N; ++k) r.d[k] = a.d[k] - b.d[k];
  return r;
}
template <int N>
SSB_HD Dual<N> operator-(const Dual<N>& a) {
  Dual<N> r;
  r.v = -a.v;
  for (int k = 0; k < N; ++k) r.d[k] = -a.d[k];
  return r;
}
template <int N>
SSB_HD Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r;
  r.v = a.v * b.v;
  for (int k = 0; k < N; ++k) r.d[k] = a.d[k] * b.v + a.v * b.d[k];
  return r;
}
template <int N>
SSB_HD Dual<N> operator*(const Dual<N>& a, double b) {
  Dual<N> r;
  r.v = a.v * b;
  for (int k = 0; k < N; ++k) r.d[k] = a.d[k] * b;
  return r;
}
template <int N>
SSB_HD Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r;
  const double ib = 1.0 / b.v;
  r.v = a.v * ib;
  for (int k = 0; k < N; ++k) r.d[k] = (a.d[k] - r.v * b.d[k]) * ib;
  return r;
}
template <int N>
SSB_HD Dual<N> dsqrt(const Dual<N>& a) {
  Dual<N> r;
  r.v = sqrt(a.v);
  const double g = r.v > 0.0 ? 0.5 / r.v : 0.0;
  for (int k = 0; k < N; ++k) r.d[k] = a.d[k] * g;
  return r;
}
template <int N>
SSB_HD Dual<N> dsin(const Dual<N>& a) {
  Dual<N> r;
  r.v = sin(a.v);
  const double g = cos(a.v);
  for (int k = 0; k < N; ++k) r.d[k] = a.d[k] * g;
  return r;
}
template <int N>
SSB_HD Dual<N> dcos(const Dual<N>& a) {
  Dual<N> r;
  r.v = cos(a.v);
  const double g = -sin(a.v);
  for (int k = 0; k < N; ++k) r.d[k] = a.d[k] * g;
  return r;
}
template <int N>
SSB_HD Dual<N> datan2(const Dual<N>& y, const Dual<N>& x) {
  Dual<N> r;
  r.v = atan2(y.v, x.v);
  const double n2 = x.v * x.v + y.v * y.v;
  const double gx = n2 > 0.0 ? -y.v / n2 : 0.0, gy = n2 > 0.0 ? x.v / n2 : 0.0;
  for (int k = 0; k < N; ++k) r.d[k] = gy * y.d[k] + gx * x.d[k];
  return r;
}
SSB_HD double dconst0(double v) { return v; }
SSB_HD double dsqrt(double a) { return sqrt(a); }
SSB_HD double dsin(double a) { return sin(a); }
SSB_HD double dcos(double a) { return cos(a); }
SSB_HD double datan2(double y, double x) { return atan2(y, x); }

// Plane3D::rotation(v) = AngleAxis(azimuth(v), Z) * AngleAxis(-elevation(v), Y), row-major
template <class T>
SSB_HD void plane_rotation(const T* v, T* R, const T& zero) {
  const T az = datan2(v[1], v[0]);
  const T el = datan2(v[2], dsqrt(v[0] * v[0] + v[1] * v[1]));
  const T ca = dcos(az), sa = dsin(az), ce = dcos(el), se = dsin(el);
  R[0] = ca * ce;
  R[1] = zero - sa;
  R[2] = zero - ca * se;
  R[3] = sa * ce;
  R[4] = ca;
  R[5] = zero - sa * se;
  R[6] = se;
  R[7] = zero;
  R[8] = ce;
}
// EdgeSE3Plane::computeError on scalar type T (double or Dual): pose rotation R (world <- robot) and
// translation t, plane coefficients pl (world), measurement zm (normalised, robot frame)
template <class T>
SSB_HD void plane_error_t(const T* R, const T* t, const T* pl, const double* zm, T* e, const T& zero) {
  // local = X^-1 * plane : n_l = R^T n, c3_l = c3 + t . n ; Plane3D(v) normalises
  T nl[3];
  for (int r = 0; r < 3; ++r) nl[r] = R[r] * pl[0] + R[3 + r] * pl[1] + R[6 + r] * pl[2];
  T c3 = pl[3] + t[0] * pl[0] + t[1] * pl[1] + t[2] * pl[2];
  const T nn = dsqrt(nl[0] * nl[0] + nl[1] * nl[1] + nl[2] * nl[2]);
  for (int r = 0; r < 3; ++r) nl[r] = nl[r] / nn;
  c3 = c3 / nn;
  // ominus: n' = rotation(n_l)^T n_m
  T Rn[9];
  plane_rotation(nl, Rn, zero);
  T n[3];
  for (int r = 0; r < 3; ++r) n[r] = Rn[r] * zm[0] + Rn[3 + r] * zm[1] + Rn[6 + r] * zm[2];
  e[0] = datan2(n[1], n[0]);
  e[1] = datan2(n[2], dsqrt(n[0] * n[0] + n[1] * n[1]));
  e[2] = (zero - c3) + zm[3];
}
SSB_HD void plane_normalize(double* c) {
  const double n = sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
  for (int k = 0; k < 4; ++k) c[k] /= n;
}
// Plane3D::oplus  (VertexPlane::oplusImpl)
SSB_HD void plane_oplus(double* c, const double* v) {
  const double s = sin(v[1]), co = cos(v[1]);
  const double n[3] = {co * cos(v[0]), co * sin(v[0]), s};
  double R[9];
  plane_rotation(c, R, 0.0);
  const double d = -c[3] + v[2];
  const double r0 = R[0] * n[0] + R[1] * n[1] + R[2] * n[2];
  const double r1 = R[3] * n[0] + R[4] * n[1] + R[5] * n[2];
  const double r2 = R[6] * n[0] + R[7] * n[1] + R[8] * n[2];
  c[0] = r0;
  c[1] = r1;
  c[2] = r2;
  c[3] = -d;
  plane_normalize(c);
}
SSB_HD void plane_error(const Pose& X, const double* pl, const double* zm, double* e) {
  double R[9];
  quat_to_R(X.q, R);
  plane_error_t<double>(R, X.t, pl, zm, e, 0.0);
}
// error + exact Jacobians: Jp 3x6 wrt the pose increment (X <- X * fromVectorMQT(d): t' = t + R dt,
// R' = R (I + 2 [dq]x) to first order), Jl 3x3 wrt the plane increment (n' = rotation(n) n(daz, del):
// d n / d az = column 1, d n / d el = column 2 of rotation(n); c3' = c3 - dd).  Row-major.
SSB_HD void plane_linearize(const Pose& X, const double* pl, const double* zm, double* e, double* Jp, double* Jl) {
  typedef Dual<9> D;
  double R0[9], Rp[9];
  quat_to_R(X.q, R0);
  plane_rotation(pl, Rp, 0.0);
  const D zero = dconst<9>(0.0);
  D R[9], t[3], p[4];
  for (int i = 0; i < 3; ++i) {
    t[i] = dconst<9>(X.t[i]);
    for (int k = 0; k < 3; ++k) t[i].d[k] = R0[3 * i + k];
    for (int j = 0; j < 3; ++j) R[3 * i + j] = dconst<9>(R0[3 * i + j]);
  }
  // d(R [e_k]x)/... : [e_0]x = (0 0 0; 0 0 -1; 0 1 0), [e_1]x = (0 0 1; 0 0 0; -1 0 0), [e_2]x = (0 -1 0; 1 0 0; 0 0 0)
  for (int i = 0; i < 3; ++i) {
    const double r0 = R0[3 * i], r1 = R0[3 * i + 1], r2 = R0[3 * i + 2];
    R[3 * i + 1].d[3] = 2 * r2;   // k = x: column 1 += r2, column 2 -= r1
    R[3 * i + 2].d[3] = -2 * r1;
    R[3 * i + 0].d[4] = -2 * r2;  // k = y: column 0 -= r2, column 2 += r0
    R[3 * i + 2].d[4] = 2 * r0;
    R[3 * i + 0].d[5] = 2 * r1;   // k = z: column 0 += r1, column 1 -= r0
    R[3 * i + 1].d[5] = -2 * r0;
  }
  for (int i = 0; i < 3; ++i) {
    p[i] = dconst<9>(pl[i]);
    p[i].d[6] = Rp[3 * i + 1];
    p[i].d[7] = Rp[3 * i + 2];
  }
  p[3] = dconst<9>(pl[3]);
  p[3].d[8] = -1.0;
  D ed[3];
  plane_error_t<D>(R, t, p, zm, ed, zero);
  for (int r = 0; r < 3; ++r) {
    e[r] = ed[r].v;
    for (int c = 0; c < 6; ++c) Jp[6 * r + c] = ed[r].d[c];
    for (int c = 0; c < 3; ++c) Jl[3 * r + c] = ed[r].d[6 + c];
  }
}

// -------- EdgeSE3:  e = toVectorMQT(Z^-1 Xi^-1 Xj)  (g2o types/slam3d/edge_se3.cpp)
// Ji, Jj 6x6 row-major (may be null when only the error is needed).
SSB_HD void pp_linearize(const Pose& Xi, const Pose& Xj, const double* zt, const double* zq, double* e, double* Ji,
                         double* Jj) {
  double Ri[9], Rz[9];
  quat_to_R(Xi.q, Ri);
  quat_to_R(zq, Rz);
  double d[3] = {Xj.t[0] - Xi.t[0], Xj.t[1] - Xi.t[1], Xj.t[2] - Xi.t[2]};
  double tB[3];
  rotT(Ri, d, tB);
  double dz[3] = {tB[0] - zt[0], tB[1] - zt[1], tB[2] - zt[2]};
  rotT(Rz, dz, e);
  double qic[4], qB[4], qA[4], qE[4];
  quat_conj(Xi.q, qic);
  quat_mul(qic, Xj.q, qB);
  quat_conj(zq, qA);
  quat_mul(qA, qB, qE);
  quat_normalize(qE);
  const double sgn = qE[3] < 0 ? -1.0 : 1.0;
  e[3] = sgn * qE[0];
  e[4] = sgn * qE[1];
  e[5] = sgn * qE[2];
  if (!Ji) return;
  for (int k = 0; k < 36; ++k) {
    Ji[k] = 0.0;
    Jj[k] = 0.0;
  }
  // translation rows
  double Rj[9];
  quat_to_R(Xj.q, Rj);
  // RzT (=Ra), RE = Rz^T Ri^T Rj
  double RiTRj[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) RiTRj[3 * r + c] = Ri[r] * Rj[c] + Ri[3 + r] * Rj[3 + c] + Ri[6 + r] * Rj[6 + c];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      Ji[6 * r + c] = -Rz[3 * c + r];  // -Rz^T
      Jj[6 * r + c] = Rz[r] * RiTRj[c] + Rz[3 + r] * RiTRj[3 + c] + Rz[6 + r] * RiTRj[6 + c];
    }
  // d e_t / d v_i = Rz^T * 2[tB]x
  {
    double S[9] = {0, -2 * tB[2], 2 * tB[1], 2 * tB[2], 0, -2 * tB[0], -2 * tB[1], 2 * tB[0], 0};
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) Ji[6 * r + 3 + c] = Rz[r] * S[c] + Rz[3 + r] * S[3 + c] + Rz[6 + r] * S[6 + c];
  }
  // d e_r / d v_j = sgn * (wE I + [vE]x)
  {
    const double w = qE[3], x = qE[0], y = qE[1], z = qE[2];
    double M[9] = {w, -z, y, z, w, -x, -y, x, w};
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) Jj[6 * (3 + r) + 3 + c] = sgn * M[3 * r + c];
  }
  // d e_r / d v_i = -sgn * M(qA, qB),  M = -bv av^T + bw (aw I + [av]x) - [bv]x (aw I + [av]x)
  {
    const double aw = qA[3], bw = qB[3];
    const double* av = qA;
    const double* bv = qB;
    double K[9] = {aw, -av[2], av[1], av[2], aw, -av[0], -av[1], av[0], aw};  // aw I + [av]x
    double Bx[9] = {0, -bv[2], bv[1], bv[2], 0, -bv[0], -bv[1], bv[0], 0};   // [bv]x
    // the product is for un-normalised qE = qA*qB; qE was normalised above but |qA*qB| = 1 up to rounding
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        double BK = Bx[3 * r] * K[c] + Bx[3 * r + 1] * K[3 + c] + Bx[3 * r + 2] * K[6 + c];
        double m = -bv[r] * av[c] + bw * K[3 * r + c] - BK;
        Ji[6 * (3 + r) + 3 + c] = -sgn * m;
      }
  }
}

// -------- pose-pose table entries of other kinds (PPEdge::pad = kind).  A landmark that carries a landmark-landmark edge
// (g2o::EdgePointXYZ, GraphSLAM::add_point_xyz_point_xyz_edge graph_slam.cpp:168-180) cannot be eliminated point-wise by
// the Schur step: it is PROMOTED into the reduced system as a pseudo-keyframe (t = xyz, identity rotation, the last three
// increments pinned by an identity block).  Its edges become pose-pose entries with the 3x3 information in the upper-left
// corner of the 6x6 one and errors / Jacobians padded with zeros, so every accumulation kernel runs unchanged.
//   kind 1: EdgeSE3PointXYZ, vertex i = keyframe, vertex j = promoted landmark   e = Ri'(pj - ti) - z
//   kind 2: EdgePointXYZ,    both promoted landmarks                             e = (pj - pi) - z
SSB_HD int pp_kind(const PPEdge& ed) { return (int)ed.pad; }
SSB_HD void pp_edge_linearize(const PPEdge& ed, const Pose& Xi, const Pose& Xj, double* e, double* Ji, double* Jj) {
  const int kind = pp_kind(ed);
  if (kind == 0) {
    pp_linearize(Xi, Xj, ed.zt, ed.zq, e, Ji, Jj);
    return;
  }
  e[3] = e[4] = e[5] = 0.0;
  if (Ji)
    for (int k = 0; k < 36; ++k) {
      Ji[k] = 0.0;
      Jj[k] = 0.0;
    }
  if (kind == 1) {
    PLLin L;
    pl_linearize(Xi, Xj.t, ed.zt, L);
    for (int k = 0; k < 3; ++k) e[k] = L.e[k];
    if (!Ji) return;
    double Jp[18];
    pl_jac_pose(L.pc, Jp);
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 6; ++c) Ji[6 * r + c] = Jp[6 * r + c];
      for (int c = 0; c < 3; ++c) Jj[6 * r + c] = L.R[3 * c + r];   // Ri'
    }
  } else {
    for (int k = 0; k < 3; ++k) e[k] = (Xj.t[k] - Xi.t[k]) - ed.zt[k];
    if (!Ji) return;
    for (int k = 0; k < 3; ++k) {
      Ji[7 * k] = -1.0;
      Jj[7 * k] = 1.0;
    }
  }
}

// chi2 helpers
SSB_HD double quad3(const double* u, const double* e) {
  return u[0] * e[0] * e[0] + u[3] * e[1] * e[1] + u[5] * e[2] * e[2] +
         2.0 * (u[1] * e[0] * e[1] + u[2] * e[0] * e[2] + u[4] * e[1] * e[2]);
}
SSB_HD double quad6(const double* u, const double* e) {
  double s = 0.0;
  int k = 0;
  for (int r = 0; r < 6; ++r) {
    s += u[k] * e[r] * e[r];
    ++k;
    for (int c = r + 1; c < 6; ++c) {
      s += 2.0 * u[k] * e[r] * e[c];
      ++k;
    }
  }
  return s;
}

// inverse of a symmetric positive definite 3x3 given as upper 6; out upper 6. returns false if singular
SSB_HD bool inv_sym3(const double* a, double* o) {
  double c00 = a[3] * a[5] - a[4] * a[4];
  double c01 = a[2] * a[4] - a[1] * a[5];
  double c02 = a[1] * a[4] - a[2] * a[3];
  double det = a[0] * c00 + a[1] * c01 + a[2] * c02;
  if (!(fabs(det) > 0.0)) return false;
  double id = 1.0 / det;
  o[0] = c00 * id;
  o[1] = c01 * id;
  o[2] = c02 * id;
  o[3] = (a[0] * a[5] - a[2] * a[2]) * id;
  o[4] = (a[1] * a[2] - a[0] * a[4]) * id;
  o[5] = (a[0] * a[3] - a[1] * a[1]) * id;
  return true;
}

// in-place inverse of SPD 6x6 (full row-major) via Cholesky; returns false if not PD
SSB_HD bool inv_spd6(double* A) {
  double L[36];
  for (int i = 0; i < 36; ++i) L[i] = 0.0;
  for (int j = 0; j < 6; ++j) {
    double d = A[6 * j + j];
    for (int k = 0; k < j; ++k) d -= L[6 * j + k] * L[6 * j + k];
    if (!(d > 0.0)) return false;
    double ljj = sqrt(d);
    L[6 * j + j] = ljj;
    double inv = 1.0 / ljj;
    for (int i = j + 1; i < 6; ++i) {
      double s = A[6 * i + j];
      for (int k = 0; k < j; ++k) s -= L[6 * i + k] * L[6 * j + k];
      L[6 * i + j] = s * inv;
    }
  }
  // invert L (lower) in place into Li
  double Li[36];
  for (int i = 0; i < 36; ++i) Li[i] = 0.0;
  for (int j = 0; j < 6; ++j) {
    Li[6 * j + j] = 1.0 / L[6 * j + j];
    for (int i = j + 1; i < 6; ++i) {
      double s = 0.0;
      for (int k = j; k < i; ++k) s -= L[6 * i + k] * Li[6 * k + j];
      Li[6 * i + j] = s / L[6 * i + i];
    }
  }
  // A^-1 = Li^T Li
  for (int r = 0; r < 6; ++r)
    for (int c = r; c < 6; ++c) {
      double s = 0.0;
      for (int k = c; k < 6; ++k) s += Li[6 * k + r] * Li[6 * k + c];
      A[6 * r + c] = s;
      A[6 * c + r] = s;
    }
  return true;
}

}  // namespace ssb
