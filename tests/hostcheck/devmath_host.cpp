// Test-only: compiles the product's device math header (lld_slam_b200/csrc/lld_math.cuh) for the host so that
// tests/test_device_math_host.py can compare it with the CPU oracle without a GPU.
#include "../../lld_slam_b200/csrc/lld_math.cuh"
using namespace lld;
extern "C" {
// kind 0 mono / 1 stereo binary / 3 stereo unary point; lm = xyz.  Jl 3x3, Jp 3x6 (rows beyond dim are zero)
int dm_point(int kind, const double* Tcw12, const double* X, const double* intr, const float* obs, double* err,
             double* Jl, double* Jp) {
  double qt[7], Rt[12], xc[3];
  pose_from_Rt(Tcw12, qt);
  pose_to_Rt(qt, Rt);
  map_Rt(Rt, X, xc);
  const bool st = kind != 0;
  if (kind == 3) pt_residual<false>(xc, intr, obs, st, err);
  else pt_residual<true>(xc, intr, obs, st, err);
  pt_jac_point(xc, Rt, intr, st, Jl);
  pt_jac_pose(xc, intr, st, Jp);
  return st ? 3 : 2;
}
// line edge: lm = x0,dir ; lcam = f cx cy bx ; obs = x1[3], x2[3]
int dm_line(const double* Tcw12, const double* x0dir, const double* lcam, const double* obs, double* err, double* Jl,
            double* Jp, double* err_only, int* depth_pos) {
  double qt[7], Rt[12];
  pose_from_Rt(Tcw12, qt);
  pose_to_Rt(qt, Rt);
  double st[5], r1[3], r2[3], X1[3], X2[3], P1[3], P2[3];
  line_from_x0_dir(x0dir, x0dir + 3, st);
  line_axes(st, r1, r2);
  for (int i = 0; i < 3; i++) { X1[i] = st[4] * r2[i]; X2[i] = X1[i] + r1[i]; }
  map_Rt(Rt, X1, P1);
  map_Rt(Rt, X2, P2);
  LineObs o;
  for (int i = 0; i < 3; i++) { o.x1[i] = obs[i]; o.x2[i] = obs[3 + i]; }
  line_linearize<true>(P1, P2, lcam[0], lcam[1], lcam[2], lcam[3], o, Rt, X1, X2, r2, err, Jp, Jl);
  line_residual(P1, P2, lcam[0], lcam[1], lcam[2], lcam[3], o, err_only);
  double X0c[3] = {P1[0] + lcam[3], P1[1], P1[2]};
  double ldc[3] = {P2[0] - P1[0], P2[1] - P1[1], P2[2] - P1[2]};
  *depth_pos = line_depth_positive(X0c, ldc, lcam[0], lcam[1], lcam[2], o.x1, o.x2) ? 1 : 0;
  return 2;
}
void dm_pose_oplus(const double* Tcw12, const double* u, double* out12) {
  double qt[7], q2[7];
  pose_from_Rt(Tcw12, qt);
  pose_oplus(qt, u, q2);
  pose_to_Rt(q2, out12);
}
void dm_line_oplus(const double* x0dir, const double* u, double* out6) {
  double st[5], s2[5], r1[3], r2[3];
  line_from_x0_dir(x0dir, x0dir + 3, st);
  line_oplus(st, u, s2);
  line_axes(s2, r1, r2);
  for (int i = 0; i < 3; i++) { out6[i] = s2[4] * r2[i]; out6[3 + i] = r1[i]; }
}
void dm_inv(int d, const double* A, double* I) {
  if (d == 3) inv3_sym(A, I);
  else inv4(A, I);
}
double dm_huber(double e, double delta, double* w) { return huber(e, delta, w); }
}
