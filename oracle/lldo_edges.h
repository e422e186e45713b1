// ORACLE — TEST INFRASTRUCTURE ONLY (see lldo_math.h header).  Parity unpinned by reference tests.
//
// Residuals and analytic Jacobians of the g2o edge types the three Optimizer entry points use,
// restated from Thirdparty/g2o/g2o/types/types_six_dof_expmap.{h,cpp}.
#pragma once
#include "lldo_math.h"

namespace lldo {

struct Intr {
  double fx, fy, cx, cy, bf;
};

// chi2 = e . (Omega e) with Omega = s*I  (Thirdparty/g2o/g2o/core/base_edge.h:58-61)
inline double chi2_of(const double* e, int dim, double s) {
  double c = 0;
  for (int i = 0; i < dim; i++) c += e[i] * (s * e[i]);
  return c;
}

// RobustKernelHuber::robustify  Thirdparty/g2o/g2o/core/robust_kernel_impl.cpp:65-91
inline void huber(double e, double delta, double rho[3]) {
  const double dsqr = delta * delta;
  if (e <= dsqr) {
    rho[0] = e; rho[1] = 1.; rho[2] = 0.;
  } else {
    const double sqrte = std::sqrt(e);
    rho[0] = 2 * sqrte * delta - dsqr;
    rho[1] = delta / sqrte;
    rho[2] = -0.5 * rho[1] / e;
  }
}

// ---- point edges ---------------------------------------------------------------------------
// EdgeSE3ProjectXYZ::computeError / cam_project  types_six_dof_expmap.h:93-98, .cpp:45-50,149-155
inline void pt_err_mono(const Pose& T, const double X[3], const double obs[2], const Intr& k, double err[2]) {
  double xc[3];
  pose_map(T, X, xc);
  const double p0 = xc[0] / xc[2], p1 = xc[1] / xc[2];
  err[0] = obs[0] - (p0 * k.fx + k.cx);
  err[1] = obs[1] - (p1 * k.fy + k.cy);
}
// EdgeStereoSE3ProjectXYZ::cam_project(trans_xyz, const float& bf)  .cpp:158-165 : invz AND bf are float there
inline void pt_err_stereo_binary(const Pose& T, const double X[3], const double obs[3], const Intr& k, double err[3]) {
  double xc[3];
  pose_map(T, X, xc);
  const float invz = 1.0f / xc[2];  // (float)(1.0 / z): 1.0f promotes to double, quotient narrows to float
  const float bff = (float)k.bf;
  const double r0 = xc[0] * invz * k.fx + k.cx;
  const double r1 = xc[1] * invz * k.fy + k.cy;
  const double r2 = r0 - bff * invz;  // float * float product
  err[0] = obs[0] - r0; err[1] = obs[1] - r1; err[2] = obs[2] - r2;
}
// EdgeStereoSE3ProjectXYZOnlyPose::cam_project  .cpp:307-314 : invz float, bf the double member
inline void pt_err_stereo_unary(const Pose& T, const double X[3], const double obs[3], const Intr& k, double err[3]) {
  double xc[3];
  pose_map(T, X, xc);
  const float invz = 1.0f / xc[2];
  const double r0 = xc[0] * invz * k.fx + k.cx;
  const double r1 = xc[1] * invz * k.fy + k.cy;
  const double r2 = r0 - k.bf * invz;
  err[0] = obs[0] - r0; err[1] = obs[1] - r1; err[2] = obs[2] - r2;
}
inline bool pt_depth_positive(const Pose& T, const double X[3]) {  // .h:100-104
  double xc[3];
  pose_map(T, X, xc);
  return xc[2] > 0.0;
}
// linearizeOplus of the binary edges (.cpp:111-147 mono, :196-242 stereo).  Jl: dim x 3, Jp: dim x 6 row-major.
inline void pt_jac_binary(const Pose& T, const double X[3], const Intr& k, bool stereo, double* Jl, double* Jp) {
  double xc[3], R[9];
  pose_map(T, X, xc);
  quat_to_R(T.q, R);
  const double x = xc[0], y = xc[1], z = xc[2], z_2 = z * z;
  const double fx = k.fx, fy = k.fy, bf = k.bf;
  if (!stereo) {
    // _jacobianOplusXi = -1/z * tmp * R
    double tmp[6] = {fx, 0, -x / z * fx, 0, fy, -y / z * fy};
    double tR[6];
    matmul(tmp, R, tR, 2, 3, 3);
    for (int i = 0; i < 6; i++) Jl[i] = -1. / z * tR[i];
  } else {
    for (int c = 0; c < 3; c++) {
      Jl[0 * 3 + c] = -fx * R[0 * 3 + c] / z + fx * x * R[2 * 3 + c] / z_2;
      Jl[1 * 3 + c] = -fy * R[1 * 3 + c] / z + fy * y * R[2 * 3 + c] / z_2;
      Jl[2 * 3 + c] = Jl[0 * 3 + c] - bf * R[2 * 3 + c] / z_2;
    }
  }
  Jp[0] = x * y / z_2 * fx;
  Jp[1] = -(1 + (x * x / z_2)) * fx;
  Jp[2] = y / z * fx;
  Jp[3] = -1. / z * fx;
  Jp[4] = 0;
  Jp[5] = x / z_2 * fx;
  Jp[6] = (1 + y * y / z_2) * fy;
  Jp[7] = -x * y / z_2 * fy;
  Jp[8] = -x / z * fy;
  Jp[9] = 0;
  Jp[10] = -1. / z * fy;
  Jp[11] = y / z_2 * fy;
  if (stereo) {
    Jp[12] = Jp[0] - bf * y / z_2;
    Jp[13] = Jp[1] + bf * x / z_2;
    Jp[14] = Jp[2];
    Jp[15] = Jp[3];
    Jp[16] = 0;
    Jp[17] = Jp[5] - bf / z_2;
  }
}
// linearizeOplus of the unary (OnlyPose) edges (.cpp:274-296 mono, :343-372 stereo): uses invz products
inline void pt_jac_unary(const Pose& T, const double X[3], const Intr& k, bool stereo, double* Jp) {
  double xc[3];
  pose_map(T, X, xc);
  const double x = xc[0], y = xc[1];
  const double invz = 1.0 / xc[2], invz_2 = invz * invz;
  const double fx = k.fx, fy = k.fy, bf = k.bf;
  Jp[0] = x * y * invz_2 * fx;
  Jp[1] = -(1 + (x * x * invz_2)) * fx;
  Jp[2] = y * invz * fx;
  Jp[3] = -invz * fx;
  Jp[4] = 0;
  Jp[5] = x * invz_2 * fx;
  Jp[6] = (1 + y * y * invz_2) * fy;
  Jp[7] = -x * y * invz_2 * fy;
  Jp[8] = -x * invz * fy;
  Jp[9] = 0;
  Jp[10] = -invz * fy;
  Jp[11] = y * invz_2 * fy;
  if (stereo) {
    Jp[12] = Jp[0] - bf * y * invz_2;
    Jp[13] = Jp[1] + bf * x * invz_2;
    Jp[14] = Jp[2];
    Jp[15] = Jp[3];
    Jp[16] = 0;
    Jp[17] = Jp[5] - bf * invz_2;
  }
}

// ---- line edges ----------------------------------------------------------------------------
struct LineCam {
  double f, cx, cy;
  double bx;  // b = (bx, 0, 0): 0 for the left image, -baseline for the right
};
inline void line_K(const LineCam& c, double K[9]) {
  K[0] = c.f; K[1] = 0; K[2] = c.cx;
  K[3] = 0; K[4] = c.f; K[5] = c.cy;
  K[6] = 0; K[7] = 0; K[8] = 1;
}
// common tail of EdgeSE3ProjectLine{,OnlyPose}::computeError  types_six_dof_expmap.h:365-372,412-417
inline void line_err_from_points(const Pose& T, const double X1[3], const double X2[3], const LineCam& c,
                                 const double x1[3], const double x2[3], double err[2]) {
  double K[9], a[3], A1[3], A2[3], lt[3];
  line_K(c, K);
  pose_map(T, X1, a); a[0] += c.bx; matvec3(K, a, A1);
  pose_map(T, X2, a); a[0] += c.bx; matvec3(K, a, A2);
  cross3(A1, A2, lt);
  const double n = std::sqrt(lt[0] * lt[0] + lt[1] * lt[1]);
  const double l[3] = {lt[0] / n, lt[1] / n, lt[2] / n};
  err[0] = dot3(x1, l);
  err[1] = dot3(x2, l);
}
inline void line_points(const LineState& L, double X1[3], double X2[3], double R[9]) {  // .h:352-354
  line_R(L, R);
  for (int i = 0; i < 3; i++) {
    X1[i] = R[3 * i + 1] * L.alpha;
    X2[i] = X1[i] + R[3 * i + 0];
  }
}
inline void line_err(const Pose& T, const LineState& L, const LineCam& c, const double x1[3], const double x2[3],
                     double err[2]) {
  double X1[3], X2[3], R[9];
  line_points(L, X1, X2, R);
  line_err_from_points(T, X1, X2, c, x1, x2, err);
}
// FormJacobianLineWRTCam  types_six_dof_expmap.cpp:472-499.  X1m/X2m = T.map(Xi) WITHOUT b.
inline void form_jac_line_wrt_cam(const double X1m[3], const double X2m[3], const LineCam& c, double J_l[18],
                                  double D[9], double A1[3], double A2[3], double K[9]) {
  line_K(c, K);
  double a[3];
  a[0] = X1m[0] + c.bx; a[1] = X1m[1]; a[2] = X1m[2]; matvec3(K, a, A1);
  a[0] = X2m[0] + c.bx; a[1] = X2m[1]; a[2] = X2m[2]; matvec3(K, a, A2);
  double lt[3];
  cross3(A1, A2, lt);
  const double n = std::sqrt(lt[0] * lt[0] + lt[1] * lt[1]);
  const double dn[3] = {-lt[0] / (n * n * n), -lt[1] / (n * n * n), 0};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) D[3 * i + j] = lt[i] * dn[j] + (i == j ? 1.0 / n : 0.0);
  double J2[18], J1[18], cp[9], Kc[9];
  cpmat(X2m, cp); matmul(K, cp, Kc, 3, 3, 3);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) { J2[6 * i + j] = -Kc[3 * i + j]; J2[6 * i + 3 + j] = K[3 * i + j]; }
  cpmat(X1m, cp); matmul(K, cp, Kc, 3, 3, 3);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) { J1[6 * i + j] = -Kc[3 * i + j]; J1[6 * i + 3 + j] = K[3 * i + j]; }
  double cA1[9], cA2[9], t1[18], t2[18], Jt[18];
  cpmat(A1, cA1); cpmat(A2, cA2);
  matmul(cA1, J2, t1, 3, 3, 6);
  matmul(cA2, J1, t2, 3, 3, 6);
  for (int i = 0; i < 18; i++) Jt[i] = t1[i] - t2[i];
  matmul(D, Jt, J_l, 3, 3, 6);
}
// EdgeSE3ProjectLine::linearize  types_six_dof_expmap.cpp:507-553.  Jl: 2x4 (line vertex), Jp: 2x6 (pose)
inline void line_jac_binary(const Pose& T, const LineState& L, const LineCam& c, const double x1[3],
                            const double x2[3], double Jl[8], double Jp[12]) {
  double X1[3], X2[3], R[9], X1m[3], X2m[3];
  line_points(L, X1, X2, R);
  pose_map(T, X1, X1m);
  pose_map(T, X2, X2m);
  double J_l[18], D[9], A1[3], A2[3], K[9];
  form_jac_line_wrt_cam(X1m, X2m, c, J_l, D, A1, A2, K);
  for (int j = 0; j < 6; j++) {
    Jp[j] = x1[0] * J_l[j] + x1[1] * J_l[6 + j] + x1[2] * J_l[12 + j];
    Jp[6 + j] = x2[0] * J_l[j] + x2[1] * J_l[6 + j] + x2[2] * J_l[12 + j];
  }
  const double r1[3] = {R[0], R[3], R[6]}, r2[3] = {R[1], R[4], R[7]};
  const double ar2[3] = {L.alpha * r2[0], L.alpha * r2[1], L.alpha * r2[2]};
  double c1[9], c2[9];
  cpmat(ar2, c1);
  cpmat(r1, c2);
  double dX1[12], dX2[12];
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) {
      dX1[4 * i + j] = 2 * (-c1[3 * i + j]);
      dX2[4 * i + j] = dX1[4 * i + j] - 2 * c2[3 * i + j];
    }
    dX1[4 * i + 3] = r2[i];
    dX2[4 * i + 3] = r2[i];
  }
  double Rc[9], KR[9], cA1[9], cA2[9], m1[9], m2[9], t1[12], t2[12], dlt[12], Dl[12];
  quat_to_R(T.q, Rc);
  matmul(K, Rc, KR, 3, 3, 3);
  cpmat(A1, cA1); cpmat(A2, cA2);
  matmul(cA1, KR, m1, 3, 3, 3);
  matmul(cA2, KR, m2, 3, 3, 3);
  matmul(m1, dX2, t1, 3, 3, 4);
  matmul(m2, dX1, t2, 3, 3, 4);
  for (int i = 0; i < 12; i++) dlt[i] = t1[i] - t2[i];
  matmul(D, dlt, Dl, 3, 3, 4);
  for (int j = 0; j < 4; j++) {
    Jl[j] = x1[0] * Dl[j] + x1[1] * Dl[4 + j] + x1[2] * Dl[8 + j];
    Jl[4 + j] = x2[0] * Dl[j] + x2[1] * Dl[4 + j] + x2[2] * Dl[8 + j];
  }
}
// EdgeSE3ProjectLineOnlyPose::linearize  types_six_dof_expmap.cpp:583-613
inline void line_jac_unary(const Pose& T, const double X1[3], const double X2[3], const LineCam& c,
                           const double x1[3], const double x2[3], double Jp[12]) {
  double X1m[3], X2m[3];
  pose_map(T, X1, X1m);
  pose_map(T, X2, X2m);
  double J_l[18], D[9], A1[3], A2[3], K[9];
  form_jac_line_wrt_cam(X1m, X2m, c, J_l, D, A1, A2, K);
  for (int j = 0; j < 6; j++) {
    Jp[j] = x1[0] * J_l[j] + x1[1] * J_l[6 + j] + x1[2] * J_l[12 + j];
    Jp[6 + j] = x2[0] * J_l[j] + x2[1] * J_l[6 + j] + x2[2] * J_l[12 + j];
  }
}
// EdgeSE3ProjectLine::IsDepthPositive  types_six_dof_expmap.h:312-342
inline bool line_depth_positive(const Pose& T, const LineState& L, const LineCam& c, const double x1[3],
                                const double x2[3]) {
  double R[9];
  line_R(L, R);
  double X0[3], X0d[3];
  for (int i = 0; i < 3; i++) {
    X0[i] = R[3 * i + 1] * L.alpha;
    X0d[i] = X0[i] + R[3 * i + 0];
  }
  double X0l[3], t[3], ldl[3];
  pose_map(T, X0, X0l); X0l[0] += c.bx;
  pose_map(T, X0d, t); t[0] += c.bx;
  for (int i = 0; i < 3; i++) ldl[i] = t[i] - X0l[i];
  double K[9];
  line_K(c, K);
  double d1, d2, p;
  reproject_line_point(X0l, ldl, x1, K, &d1, &p);
  reproject_line_point(X0l, ldl, x2, K, &d2, &p);
  return !(d1 < 0 || d2 < 0);
}

}  // namespace lldo
