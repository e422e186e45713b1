// ORACLE — TEST INFRASTRUCTURE ONLY (see lldo_math.h header).  Parity unpinned by reference tests.
//
// CPU restatement of Optimizer::PoseOptimization (src/Optimizer.cc:562-932): one 6-DoF vertex, unary
// point / line edges, BlockSolver_6_3 + LinearSolverDense + Levenberg, 4 rounds x 10 iterations with
// inlier re-classification after every round.
#include <algorithm>
#include <limits>

#include "../include/lldba.h"
#include "lldo_edges.h"

namespace lldo {

struct UEdge {
  int kind;  // 0 mono point, 1 stereo point, 2 line
  int idx;   // point index / line index (frame-local)
  int side;  // line: 0 left 1 right
  double X1[3], X2[3];  // point: X1 = Xw ; line: X1 = X0, X2 = X0 + dir
  double obs[3];
  double x1[3], x2[3];
  LineCam cam;
  double info;
  double delta;
  bool robust = true;
  int level = 0;
  bool gate_stereo = false;
  double err[3] = {0, 0, 0};
  int dim() const { return kind == 1 ? 3 : 2; }
  double chi2() const { return chi2_of(err, dim(), info); }
};

struct FrameOpt {
  Pose T;
  Intr k;
  std::vector<UEdge> ed;
  double lambda = -1, ni = 2;
  int n_bad = 0;
  double last_chi = 0;

  void compute_error(UEdge& e) const {
    if (e.kind == 0) pt_err_mono(T, e.X1, e.obs, k, e.err);
    else if (e.kind == 1) pt_err_stereo_unary(T, e.X1, e.obs, k, e.err);
    else line_err_from_points(T, e.X1, e.X2, e.cam, e.x1, e.x2, e.err);
  }
  void compute_active_errors() {
    for (auto& e : ed)
      if (e.level == 0) compute_error(e);
  }
  double active_robust_chi2() const {
    double chi = 0, rho[3];
    for (const auto& e : ed) {
      if (e.level != 0) continue;
      if (e.robust) { huber(e.chi2(), e.delta, rho); chi += rho[0]; }
      else chi += e.chi2();
    }
    return chi;
  }
  // BaseUnaryEdge::constructQuadraticForm  Thirdparty/g2o/g2o/core/base_unary_edge.hpp:43-72
  void build_system(double H[36], double b[6]) const {
    for (int i = 0; i < 36; i++) H[i] = 0;
    for (int i = 0; i < 6; i++) b[i] = 0;
    double Jp[18], rho[3];
    for (const auto& e : ed) {
      if (e.level != 0) continue;
      const int D = e.dim();
      if (e.kind == 2) line_jac_unary(T, e.X1, e.X2, e.cam, e.x1, e.x2, Jp);
      else pt_jac_unary(T, e.X1, k, e.kind == 1, Jp);
      double w = 1.0;
      if (e.robust) { huber(e.chi2(), e.delta, rho); w = rho[1]; }
      for (int r = 0; r < 6; r++) {
        for (int i = 0; i < D; i++) b[r] -= w * Jp[i * 6 + r] * e.info * e.err[i];
        for (int c = 0; c < 6; c++)
          for (int i = 0; i < D; i++) H[r * 6 + c] += Jp[i * 6 + r] * (w * e.info) * Jp[i * 6 + c];
      }
    }
  }
  // LinearSolverDense (Eigen::LDLT + isPositive)  Thirdparty/g2o/g2o/solvers/linear_solver_dense.h:65-113
  static bool solve6(const double H[36], const double b[6], double x[6]) {
    Skyline S;
    S.init(6, std::vector<int>(6, 0));
    for (int r = 0; r < 6; r++)
      for (int c = 0; c <= r; c++) S.at(r, c) = H[r * 6 + c];
    if (!skyline_ldlt(S, true)) return false;
    skyline_solve(S, b, x);
    return true;
  }
  int lm_solve(int iteration, double x[6]) {
    compute_active_errors();
    double currentChi = active_robust_chi2();
    double tempChi = currentChi;
    const double iniChi = currentChi;
    double H[36], b[6];
    build_system(H, b);
    if (iteration == 0) {
      double mx = 0;
      for (int j = 0; j < 6; j++) mx = std::max(std::fabs(H[j * 7]), mx);
      lambda = 1e-5 * mx;
      ni = 2;
      n_bad = 0;
    }
    double rho = 0;
    int qmax = 0;
    do {
      const Pose bk = T;
      double Hl[36];
      for (int i = 0; i < 36; i++) Hl[i] = H[i];
      for (int j = 0; j < 6; j++) Hl[j * 7] += lambda;
      const bool ok2 = solve6(Hl, b, x);
      T = pose_mul(pose_exp(x), T);
      compute_active_errors();
      tempChi = active_robust_chi2();
      if (!ok2) tempChi = std::numeric_limits<double>::max();
      rho = currentChi - tempChi;
      double scale = 0;
      for (int j = 0; j < 6; j++) scale += x[j] * (lambda * x[j] + b[j]);
      scale += 1e-3;
      rho /= scale;
      if (rho > 0 && std::isfinite(tempChi)) {
        double alpha = 1. - std::pow((2 * rho - 1), 3);
        alpha = (std::min)(alpha, 2. / 3.);
        lambda *= (std::max)(1. / 3., alpha);
        ni = 2;
        currentChi = tempChi;
      } else {
        lambda *= ni;
        ni *= 2;
        T = bk;
      }
      qmax++;
    } while (rho < 0 && qmax < 10);
    last_chi = currentChi;
    if (qmax == 10 || rho == 0) return 1;
    if ((iniChi - currentChi) * 1e3 < iniChi) n_bad++;
    else n_bad = 0;
    if (n_bad >= 3) return 1;
    return 0;
  }
  void optimize(int its) {
    bool any = false;
    for (auto& e : ed) any |= (e.level == 0);
    if (!any) return;  // empty index mapping: optimize() returns -1 untouched
    double x[6] = {0, 0, 0, 0, 0, 0};
    bool ok = true;
    for (int i = 0; i < its && ok; i++) ok = (lm_solve(i, x) == 0);
  }
};

}  // namespace lldo

using namespace lldo;

extern "C" int lldo_pose_opt(void*, const lld_pose_problem* p, lld_pose_result* out) {
  for (int f = 0; f < p->n_frames; f++) {
    FrameOpt F;
    const double* T0 = &p->Tcw[(size_t)f * 12];
    const double* in = &p->intr[(size_t)f * 5];
    const double* lc = &p->line_cam[(size_t)f * 4];
    F.k = Intr{in[0], in[1], in[2], in[3], in[4]};
    const int p0 = p->pt_off[f], p1 = p->pt_off[f + 1];
    const int l0 = p->ln_off[f], l1 = p->ln_off[f + 1];
    int n_init = 0;
    for (int i = p0; i < p1; i++) {
      UEdge e;
      const float* o = &p->pt_uvr[(size_t)i * 3];
      e.kind = (o[2] < 0) ? 0 : 1;
      e.idx = i - p0;
      e.side = 0;
      for (int c = 0; c < 3; c++) { e.X1[c] = p->pt_xw[(size_t)i * 3 + c]; e.obs[c] = o[c]; }
      e.info = p->pt_info[i];
      e.delta = e.kind == 1 ? p->delta_stereo : p->delta_mono;
      F.ed.push_back(e);
      n_init++;
      out->pt_outlier[i] = 0;
    }
    const size_t n_pt_edges = F.ed.size();
    for (int i = l0; i < l1; i++) {
      out->ln_outlier[i] = 0;
      const double* xd = &p->ln_x0_dir[(size_t)i * 6];
      for (int si = 0; si < 2; si++) {
        const float* seg = (si == 0 ? p->ln_left : p->ln_right) + (size_t)i * 4;
        if (si == 1 && seg[0] < 0) continue;
        UEdge e;
        e.kind = 2; e.idx = i - l0; e.side = si;
        for (int c = 0; c < 3; c++) { e.X1[c] = xd[c]; e.X2[c] = xd[c] + xd[3 + c]; }
        e.x1[0] = seg[0]; e.x1[1] = seg[1]; e.x1[2] = 1.0;
        e.x2[0] = seg[2]; e.x2[1] = seg[3]; e.x2[2] = 1.0;
        e.cam = LineCam{lc[0], lc[1], lc[2], si == 1 ? -lc[3] : 0.0};
        e.info = p->ln_info[(size_t)i * 2 + si];
        e.delta = p->ln_stereo[i] ? p->delta_ln_stereo : p->delta_ln_mono;
        e.gate_stereo = p->ln_gate_stereo[(size_t)i * 2 + si] != 0;
        F.ed.push_back(e);
      }
    }
    const Pose Tinit = pose_from_Rt(T0);
    F.T = Tinit;
    if (n_init < 3) {  // src/Optimizer.cc:809-810
      for (int c = 0; c < 12; c++) out->Tcw[(size_t)f * 12 + c] = T0[c];
      out->n_inliers[f] = 0;
      if (out->chi2_final) out->chi2_final[f] = 0;
      continue;
    }
    int nBad = 0;
    for (int it = 0; it < p->n_rounds; it++) {
      F.T = Tinit;  // :823
      F.optimize(p->its);
      nBad = 0;
      for (size_t ei = 0; ei < n_pt_edges; ei++) {
        UEdge& e = F.ed[ei];
        if (out->pt_outlier[p0 + e.idx]) F.compute_error(e);  // :834-837
        const float chi2 = (float)e.chi2();
        const float th = e.kind == 1 ? p->chi2_stereo : p->chi2_mono;
        if (chi2 > th) { out->pt_outlier[p0 + e.idx] = 1; e.level = 1; nBad++; }
        else { out->pt_outlier[p0 + e.idx] = 0; e.level = 0; }
        if (it == 2) e.robust = false;
      }
      if (F.ed.size() < 10) break;  // :886
      for (size_t ei = n_pt_edges; ei < F.ed.size(); ei++) {
        UEdge& e = F.ed[ei];
        F.compute_error(e);  // :895
        const float chi2 = (float)e.chi2();
        const double thr = e.gate_stereo ? p->gate_ln_stereo : p->gate_ln_mono;
        if (chi2 > thr) { out->ln_outlier[l0 + e.idx] = 1; e.level = 1; }
        else { out->ln_outlier[l0 + e.idx] = 0; e.level = 0; }
        if (it == 2) e.robust = false;
      }
    }
    pose_to_Rt(F.T, &out->Tcw[(size_t)f * 12]);
    out->n_inliers[f] = n_init - nBad;
    if (out->chi2_final) out->chi2_final[f] = F.last_chi;
  }
  return 0;
}
