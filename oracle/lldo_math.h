// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product; only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may build, load or call anything under oracle/.
//
// CPU restatement of the LLD-SLAM point+line BA / matching arithmetic (dependency-free C++17).
// PARITY UNPINNED by reference tests or reference outputs: the reference ships no tests or golden vectors
//  and cannot be compiled here (Eigen / OpenCV / LBDMOD absent).  What pins the oracle instead:
// the in-tree source it follows (cited per function, paths relative to the reference root) and, for every
// entry point, an independent second transcription of the same reference code in numpy / plain Python that it
// must reproduce (tests/g2o_numpy.py + tests/test_oracle_pin.py for LocalBundleAdjustment, BundleAdjustment and
// PoseOptimization; tests/test_cpu_oracle.py for the matchers, ComputeStereoMatches, AddLinesFrom and the medoid
// rule), plus numeric Jacobians, numpy dense solves and cv2 Hamming.  DESIGN.md §2.
//
// Small fixed-size linear algebra restating the Eigen 3.x conventions the reference relies on
// (Eigen itself is not vendored in the reference).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace lldo {

// ---- 3-vectors / 3x3 row-major -------------------------------------------------------------
inline void cross3(const double a[3], const double b[3], double c[3]) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
inline double dot3(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline double norm3(const double a[3]) { return std::sqrt(dot3(a, a)); }
inline void matvec3(const double M[9], const double v[3], double o[3]) {
  for (int i = 0; i < 3; i++) o[i] = M[3 * i] * v[0] + M[3 * i + 1] * v[1] + M[3 * i + 2] * v[2];
}
// C(r x c) = A(r x k) * B(k x c), row-major, generic small sizes
inline void matmul(const double* A, const double* B, double* C, int r, int k, int c) {
  for (int i = 0; i < r; i++)
    for (int j = 0; j < c; j++) {
      double s = 0;
      for (int t = 0; t < k; t++) s += A[i * k + t] * B[t * c + j];
      C[i * c + j] = s;
    }
}
// g2o cpmat / skew (Thirdparty/g2o/g2o/types/types_six_dof_expmap.cpp:36-43, se3_ops.hpp:27-38)
inline void cpmat(const double t[3], double M[9]) {
  M[0] = 0; M[1] = -t[2]; M[2] = t[1];
  M[3] = t[2]; M[4] = 0; M[5] = -t[0];
  M[6] = -t[1]; M[7] = t[0]; M[8] = 0;
}

// ---- quaternions, stored x y z w as Eigen::Quaterniond::coeffs() -----------------------------
// Eigen Quaterniond(Matrix3d) (quaternionbase_assign_impl<.,3,3>)
inline void quat_from_R(const double m[9], double q[4]) {
  double t = m[0] + m[4] + m[8];
  if (t > 0.0) {
    t = std::sqrt(t + 1.0);
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
    t = std::sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (m[3 * k + j] - m[3 * j + k]) * t;
    q[j] = (m[3 * j + i] + m[3 * i + j]) * t;
    q[k] = (m[3 * k + i] + m[3 * i + k]) * t;
  }
}
// Eigen QuaternionBase::toRotationMatrix
inline void quat_to_R(const double q[4], double R[9]) {
  const double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
  const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  const double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  const double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
// Eigen QuaternionBase::_transformVector
inline void quat_rot(const double q[4], const double v[3], double o[3]) {
  double uv[3], c2[3];
  cross3(q, v, uv);
  uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
  cross3(q, uv, c2);
  for (int i = 0; i < 3; i++) o[i] = v[i] + q[3] * uv[i] + c2[i];
}
// Hamilton product a*b
inline void quat_mul(const double a[4], const double b[4], double o[4]) {
  const double w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  const double x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  const double y = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
  const double z = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z; o[3] = w;
}
inline void quat_normalize(double q[4]) {
  const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; i++) q[i] /= n;
}

// ---- SE3Quat (Thirdparty/g2o/g2o/types/se3quat.h) -------------------------------------------
struct Pose {
  double q[4];  // x y z w
  double t[3];
};
// se3quat.h:280-285
inline void pose_normalize(Pose& p) {
  if (p.q[3] < 0)
    for (int i = 0; i < 4; i++) p.q[i] *= -1;
  quat_normalize(p.q);
}
// SE3Quat(R,t)  se3quat.h:57-59 ; Converter::toSE3Quat src/Converter.cc:37-47
inline Pose pose_from_Rt(const double Rt[12]) {
  Pose p;
  quat_from_R(Rt, p.q);
  p.t[0] = Rt[9]; p.t[1] = Rt[10]; p.t[2] = Rt[11];
  pose_normalize(p);
  return p;
}
// to_homogeneous_matrix se3quat.h:269-277 -> [R | t]
inline void pose_to_Rt(const Pose& p, double Rt[12]) {
  quat_to_R(p.q, Rt);
  Rt[9] = p.t[0]; Rt[10] = p.t[1]; Rt[11] = p.t[2];
}
// map  se3quat.h:217-220
inline void pose_map(const Pose& p, const double X[3], double o[3]) {
  quat_rot(p.q, X, o);
  o[0] += p.t[0]; o[1] += p.t[1]; o[2] += p.t[2];
}
// operator*  se3quat.h:104-110
inline Pose pose_mul(const Pose& a, const Pose& b) {
  Pose r = a;
  double rt[3];
  quat_rot(a.q, b.t, rt);
  r.t[0] += rt[0]; r.t[1] += rt[1]; r.t[2] += rt[2];
  quat_mul(a.q, b.q, r.q);
  pose_normalize(r);
  return r;
}
// exp  se3quat.h:223-257  (update = omega, upsilon)
inline Pose pose_exp(const double u[6]) {
  const double* om = u;
  const double* up = u + 3;
  const double theta = norm3(om);
  double Om[9], Om2[9], R[9], V[9];
  cpmat(om, Om);  // skew == cpmat
  matmul(Om, Om, Om2, 3, 3, 3);
  static const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  if (theta < 0.00001) {
    for (int i = 0; i < 9; i++) R[i] = I[i] + Om[i] + Om2[i];
    for (int i = 0; i < 9; i++) V[i] = R[i];
  } else {
    const double a = std::sin(theta) / theta;
    const double b = (1 - std::cos(theta)) / (theta * theta);
    const double c = (theta - std::sin(theta)) / (std::pow(theta, 3));
    for (int i = 0; i < 9; i++) R[i] = I[i] + a * Om[i] + b * Om2[i];
    for (int i = 0; i < 9; i++) V[i] = I[i] + b * Om[i] + c * Om2[i];
  }
  Pose p;
  quat_from_R(R, p.q);
  matvec3(V, up, p.t);
  pose_normalize(p);
  return p;
}

// ---- LineParams / VertexSBALine (Thirdparty/g2o/g2o/types/types_sba.h:62-110, types_sba.cpp:58-92) ----
struct LineState {
  double q[4];  // stored un-normalised; GetQ() normalises
  double alpha;
};
inline void line_getq(const LineState& l, double q[4]) {
  for (int i = 0; i < 4; i++) q[i] = l.q[i];
  quat_normalize(q);
}
inline void line_R(const LineState& l, double R[9]) {
  double q[4];
  line_getq(l, q);
  quat_to_R(q, R);
}
// LineOptimizer::AddLineMinimal src/LineOptimizer.cc:44-50
inline LineState line_from_x0_dir(const double x0[3], const double dir[3]) {
  const double n = norm3(x0);
  double c[3];
  cross3(dir, x0, c);
  double R[9];
  for (int i = 0; i < 3; i++) {
    R[3 * i + 0] = dir[i];
    R[3 * i + 1] = x0[i] / n;
    R[3 * i + 2] = c[i] / n;
  }
  LineState l;
  quat_from_R(R, l.q);
  l.alpha = n;
  return l;
}
// read-back src/LineOptimizer.cc:180-182
inline void line_to_x0_dir(const LineState& l, double x0[3], double dir[3]) {
  double R[9];
  line_R(l, R);
  for (int i = 0; i < 3; i++) {
    dir[i] = R[3 * i + 0];
    x0[i] = l.alpha * R[3 * i + 1];
  }
}
// oplusImpl types_sba.h:95-108
inline void line_oplus(LineState& l, const double u[4]) {
  double qr[4] = {u[0], u[1], u[2], 0};
  qr[3] = std::sqrt(1.0 - (u[0] * u[0] + u[1] * u[1] + u[2] * u[2]));
  double qn[4], o[4];
  line_getq(l, qn);
  quat_mul(qr, qn, o);
  for (int i = 0; i < 4; i++) l.q[i] = o[i];
  l.alpha += u[3];
}

// ---- colPivHouseholderQr restatement for the tiny systems in src/vgl.cc ----------------------
// Least-squares solve of an (m x n) system, m>=n<=3, with column pivoting; returns rank
// (Eigen default threshold: |pivot| > maxpivot * eps * min(m,n)).
inline int colpiv_qr_solve(int m, int n, const double* A_in, const double* b_in, double* x) {
  double A[9], b[3];
  int perm[3] = {0, 1, 2};
  for (int i = 0; i < m * n; i++) A[i] = A_in[i];
  for (int i = 0; i < m; i++) b[i] = b_in[i];
  double maxpiv = 0;
  double diag[3] = {0, 0, 0};
  int rank = n;
  for (int k = 0; k < n; k++) {
    // pick column with largest remaining norm
    int best = k;
    double bestn = -1;
    for (int j = k; j < n; j++) {
      double s = 0;
      for (int i = k; i < m; i++) s += A[i * n + j] * A[i * n + j];
      if (s > bestn) { bestn = s; best = j; }
    }
    if (best != k) {
      for (int i = 0; i < m; i++) std::swap(A[i * n + k], A[i * n + best]);
      std::swap(perm[k], perm[best]);
    }
    // Householder on column k, rows k..m-1
    double nrm = std::sqrt(bestn);
    if (nrm == 0) { diag[k] = 0; continue; }
    double alpha = A[k * n + k] > 0 ? -nrm : nrm;
    double v[3] = {0, 0, 0};
    for (int i = k; i < m; i++) v[i] = A[i * n + k];
    v[k] -= alpha;
    double vn2 = 0;
    for (int i = k; i < m; i++) vn2 += v[i] * v[i];
    if (vn2 > 0) {
      for (int j = k; j < n; j++) {
        double s = 0;
        for (int i = k; i < m; i++) s += v[i] * A[i * n + j];
        s = 2 * s / vn2;
        for (int i = k; i < m; i++) A[i * n + j] -= s * v[i];
      }
      double s = 0;
      for (int i = k; i < m; i++) s += v[i] * b[i];
      s = 2 * s / vn2;
      for (int i = k; i < m; i++) b[i] -= s * v[i];
    }
    diag[k] = A[k * n + k];
    if (std::fabs(diag[k]) > maxpiv) maxpiv = std::fabs(diag[k]);
  }
  const double thr = maxpiv * 2.220446049250313e-16 * (double)(m < n ? m : n);
  rank = 0;
  for (int k = 0; k < n; k++)
    if (std::fabs(diag[k]) > thr) rank++;
  // back substitution on the leading rank x rank block
  double y[3] = {0, 0, 0};
  for (int k = rank - 1; k >= 0; k--) {
    double s = b[k];
    for (int j = k + 1; j < rank; j++) s -= A[k * n + j] * y[j];
    y[k] = s / A[k * n + k];
  }
  for (int k = 0; k < n; k++) x[perm[k]] = (k < rank) ? y[k] : 0.0;
  return rank;
}

// vgl::ReprojectLinePointTo3D  src/vgl.cc:336-346
inline void reproject_line_point(const double X0[3], const double ld[3], const double pp[2], const double K[9],
                                 double* depth, double* param) {
  double Kd[3], KX[3];
  matvec3(K, ld, Kd);
  matvec3(K, X0, KX);
  double M[6] = {pp[0], -Kd[0], pp[1], -Kd[1], 1.0, -Kd[2]};
  double sol[2];
  colpiv_qr_solve(3, 2, M, KX, sol);
  *depth = sol[0];
  *param = sol[1];
}

// ---- envelope (skyline) LDL^T: stands in for Eigen::SimplicialLDLT / Eigen::LDLT ---------------
// (Thirdparty/g2o/g2o/solvers/linear_solver_eigen.h:94-124, linear_solver_dense.h:65-113).  Any exact
// factorisation reproduces them to rounding; no fill outside the envelope.
struct Skyline {
  int n = 0;
  std::vector<int> first;       // first stored column of row i
  std::vector<size_t> rowptr;   // offset of row i (entries first[i]..i)
  std::vector<double> a;
  void init(int n_, const std::vector<int>& first_) {
    n = n_;
    first = first_;
    rowptr.assign(n + 1, 0);
    for (int i = 0; i < n; i++) rowptr[i + 1] = rowptr[i] + (size_t)(i - first[i] + 1);
    a.assign(rowptr[n], 0.0);
  }
  inline double& at(int i, int j) { return a[rowptr[i] + (j - first[i])]; }  // j in [first[i], i]
  void zero() { std::fill(a.begin(), a.end(), 0.0); }
};
// in-place LDL^T; strict lower part becomes L, diagonal becomes D.  require_positive mirrors
// Eigen::LDLT::isPositive() (dense pose solver); otherwise only a zero / non-finite pivot fails
// (SimplicialLDLT NumericalIssue).
inline bool skyline_ldlt(Skyline& S, bool require_positive) {
  const int n = S.n;
  std::vector<double> y(n);
  for (int i = 0; i < n; i++) {
    const int fi = S.first[i];
    for (int j = fi; j < i; j++) {
      double s = S.at(i, j);
      const int k0 = fi > S.first[j] ? fi : S.first[j];
      for (int k = k0; k < j; k++) s -= y[k] * S.at(j, k);
      y[j] = s;
    }
    double d = S.at(i, i);
    for (int j = fi; j < i; j++) {
      const double l = y[j] / S.at(j, j);
      d -= y[j] * l;
      S.at(i, j) = l;
    }
    S.at(i, i) = d;
    if (!std::isfinite(d) || d == 0.0) return false;
    if (require_positive && d <= 0.0) return false;
  }
  return true;
}
inline void skyline_solve(Skyline& S, const double* b, double* x) {
  const int n = S.n;
  for (int i = 0; i < n; i++) {
    double s = b[i];
    for (int j = S.first[i]; j < i; j++) s -= S.at(i, j) * x[j];
    x[i] = s;
  }
  for (int i = 0; i < n; i++) x[i] /= S.at(i, i);
  for (int i = n - 1; i >= 0; i--) {
    const double xi = x[i];
    for (int j = S.first[i]; j < i; j++) x[j] -= S.at(i, j) * xi;
  }
}

// general inverse of a d x d block (d = 3 or 4) by Gauss-Jordan with partial pivoting
// (stands in for MatrixXd::inverse(), Thirdparty/g2o/g2o/core/block_solver.hpp:389)
inline void inverse_small(const double* A, double* Ainv, int d) {
  double M[4][8];
  for (int i = 0; i < d; i++)
    for (int j = 0; j < d; j++) {
      M[i][j] = A[i * d + j];
      M[i][j + d] = (i == j) ? 1.0 : 0.0;
    }
  for (int c = 0; c < d; c++) {
    int p = c;
    for (int r = c + 1; r < d; r++)
      if (std::fabs(M[r][c]) > std::fabs(M[p][c])) p = r;
    if (p != c)
      for (int j = 0; j < 2 * d; j++) std::swap(M[c][j], M[p][j]);
    const double inv = 1.0 / M[c][c];
    for (int j = 0; j < 2 * d; j++) M[c][j] *= inv;
    for (int r = 0; r < d; r++)
      if (r != c) {
        const double f = M[r][c];
        if (f != 0.0)
          for (int j = 0; j < 2 * d; j++) M[r][j] -= f * M[c][j];
      }
  }
  for (int i = 0; i < d; i++)
    for (int j = 0; j < d; j++) Ainv[i * d + j] = M[i][j + d];
}

}  // namespace lldo
