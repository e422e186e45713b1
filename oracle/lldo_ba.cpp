// ORACLE — TEST INFRASTRUCTURE ONLY (see lldo_math.h header).  Parity unpinned by reference tests.
//
// CPU restatement of the g2o path behind Optimizer::LocalBundleAdjustment / BundleAdjustment /
// PoseOptimization:  SparseOptimizer active-set handling, OptimizationAlgorithmLevenberg, BlockSolver
// (Schur complement), LineOptimizer gating.  Single-threaded like the reference (G2O_USE_OPENMP OFF,
// Thirdparty/g2o/CMakeLists.txt:48).  Exposes the same C-ABI as include/lldba.h with prefix lldo_.
#include <algorithm>
#include <cfloat>
#include <cstdio>
#include <limits>
#include <map>

#include "../include/lldba.h"
#include "lldo_edges.h"

namespace lldo {

struct PEdge {
  int pt, kf;
  bool stereo;
  double obs[3];
  double info;
  int level = 0;
  bool robust;
  double delta;
  double err[3] = {0, 0, 0};
  int dim() const { return stereo ? 3 : 2; }
  double chi2() const { return chi2_of(err, dim(), info); }
};
struct LEdge {
  int ln, kf, cell, side;
  double x1[3], x2[3];
  double info;
  LineCam cam;
  bool robust = true;
  double delta;
  bool stereo_sel;
  int level = 0;
  bool removed = false;  // edge removed from the graph together with its line vertex
  double err[2] = {0, 0};
  double chi2() const { return chi2_of(err, 2, info); }
};

// One independent BA problem.
struct Window {
  // vertices
  std::vector<Pose> poses;
  std::vector<uint8_t> fixed;
  std::vector<Intr> intr;
  std::vector<double> pts;        // 3 per point
  std::vector<LineState> lines;
  std::vector<uint8_t> ln_removed;
  // edges in insertion order: all point edges, then all line edges
  std::vector<PEdge> pe;
  std::vector<LEdge> le;
  std::vector<int> pt_e0, ln_e0;  // CSR: edges of landmark
  const volatile uint8_t* stop = nullptr;

  // ---- active set (SparseOptimizer::initializeOptimization, sparse_optimizer.cpp:199-267,166-190) ----
  std::vector<int> pose_h, pt_h, ln_h;  // hessian index or -1
  int n_ph = 0, n_pts_act = 0, n_lns_act = 0;
  std::vector<int> act_pts, act_lns;
  int size_p = 0, size_l = 0;
  std::vector<int> pt_col, ln_col;  // column of landmark in the landmark part

  void initialize_optimization() {
    const int nk = (int)poses.size(), np = (int)pt_e0.size() - 1, nl = (int)ln_e0.size() - 1;
    std::vector<uint8_t> kf_act(nk, 0);
    pt_h.assign(np, -1); ln_h.assign(nl, -1); pose_h.assign(nk, -1);
    act_pts.clear(); act_lns.clear();
    for (int p = 0; p < np; p++) {
      bool any = false;
      for (int e = pt_e0[p]; e < pt_e0[p + 1]; e++)
        if (pe[e].level == 0) { any = true; kf_act[pe[e].kf] = 1; }
      if (any) act_pts.push_back(p);
    }
    for (int l = 0; l < nl; l++) {
      if (ln_removed[l]) continue;
      bool any = false;
      for (int e = ln_e0[l]; e < ln_e0[l + 1]; e++)
        if (le[e].level == 0 && !le[e].removed) { any = true; kf_act[le[e].kf] = 1; }
      if (any) act_lns.push_back(l);
    }
    n_ph = 0;
    for (int k = 0; k < nk; k++)
      if (kf_act[k] && !fixed[k]) pose_h[k] = n_ph++;
    size_p = 6 * n_ph;
    size_l = 0;
    pt_col.assign(np, -1); ln_col.assign(nl, -1);
    for (int p : act_pts) { pt_h[p] = 1; pt_col[p] = size_l; size_l += 3; }
    for (int l : act_lns) { ln_h[l] = 1; ln_col[l] = size_l; size_l += 4; }
    x.assign(size_p + size_l, 0.0);
    b.assign(size_p + size_l, 0.0);
    structure_built = false;
  }
  bool pe_active(const PEdge& e) const { return e.level == 0; }
  bool le_active(const LEdge& e) const { return e.level == 0 && !e.removed; }

  // ---- errors (sparse_optimizer.cpp:61-114) ----
  void compute_active_errors() {
    for (auto& e : pe) {
      if (!pe_active(e)) continue;
      if (e.stereo) pt_err_stereo_binary(poses[e.kf], &pts[3 * e.pt], e.obs, intr[e.kf], e.err);
      else pt_err_mono(poses[e.kf], &pts[3 * e.pt], e.obs, intr[e.kf], e.err);
    }
    for (auto& e : le) {
      if (!le_active(e)) continue;
      line_err(poses[e.kf], lines[e.ln], e.cam, e.x1, e.x2, e.err);
    }
  }
  double active_robust_chi2() const {
    double chi = 0.0;
    double rho[3];
    for (const auto& e : pe) {
      if (!pe_active(e)) continue;
      if (e.robust) { huber(e.chi2(), e.delta, rho); chi += rho[0]; }
      else chi += e.chi2();
    }
    for (const auto& e : le) {
      if (!le_active(e)) continue;
      if (e.robust) { huber(e.chi2(), e.delta, rho); chi += rho[0]; }
      else chi += e.chi2();
    }
    return chi;
  }

  // ---- linear system (block_solver.hpp:502-560 buildSystem, base_binary_edge.hpp:55-120) ----
  std::vector<double> Hpp;           // n_ph x 36 (diagonal blocks)
  std::vector<double> Hll_p, Hll_l;  // 9 per active point / 16 per active line (indexed by landmark id)
  std::vector<double> Wp, Wl;        // Hpl block of each edge: 6x3 / 6x4 row-major (pose rows)
  std::vector<double> b, x;          // [poses | landmarks]
  std::vector<double> Dinv_p, Dinv_l;
  Skyline S;
  bool structure_built = false;
  std::vector<double> bschur, coeff;

  void build_system() {
    const int np = (int)pt_e0.size() - 1, nl = (int)ln_e0.size() - 1;
    Hpp.assign((size_t)n_ph * 36, 0.0);
    Hll_p.assign((size_t)np * 9, 0.0);
    Hll_l.assign((size_t)nl * 16, 0.0);
    Wp.assign(pe.size() * 18, 0.0);
    Wl.assign(le.size() * 24, 0.0);
    std::fill(b.begin(), b.end(), 0.0);
    double Jl[12], Jp[18], rho[3];
    for (size_t ei = 0; ei < pe.size(); ei++) {
      const PEdge& e = pe[ei];
      if (!pe_active(e)) continue;
      const int D = e.dim();
      pt_jac_binary(poses[e.kf], &pts[3 * e.pt], intr[e.kf], e.stereo, Jl, Jp);
      double w = 1.0;
      if (e.robust) { huber(e.chi2(), e.delta, rho); w = rho[1]; }
      const double wo = w * e.info;                    // weightedOmega = rho[1]*Omega
      double omega_r[3];
      for (int i = 0; i < D; i++) omega_r[i] = -(e.info * e.err[i]) * (e.robust ? rho[1] : 1.0);
      // landmark (vertex 0): never fixed
      double* bl = &b[size_p + pt_col[e.pt]];
      double* H = &Hll_p[(size_t)e.pt * 9];
      for (int r = 0; r < 3; r++) {
        for (int i = 0; i < D; i++) bl[r] += Jl[i * 3 + r] * omega_r[i];
        for (int c = 0; c < 3; c++)
          for (int i = 0; i < D; i++) H[r * 3 + c] += Jl[i * 3 + r] * wo * Jl[i * 3 + c];
      }
      const int h = pose_h[e.kf];
      if (h >= 0) {
        double* bp = &b[6 * h];
        double* P = &Hpp[(size_t)h * 36];
        double* W = &Wp[ei * 18];
        for (int r = 0; r < 6; r++) {
          for (int i = 0; i < D; i++) bp[r] += Jp[i * 6 + r] * omega_r[i];
          for (int c = 0; c < 6; c++)
            for (int i = 0; i < D; i++) P[r * 6 + c] += Jp[i * 6 + r] * wo * Jp[i * 6 + c];
          for (int c = 0; c < 3; c++)
            for (int i = 0; i < D; i++) W[r * 3 + c] += Jp[i * 6 + r] * wo * Jl[i * 3 + c];
        }
      }
    }
    for (size_t ei = 0; ei < le.size(); ei++) {
      const LEdge& e = le[ei];
      if (!le_active(e)) continue;
      line_jac_binary(poses[e.kf], lines[e.ln], e.cam, e.x1, e.x2, Jl, Jp);
      double w = 1.0;
      if (e.robust) { huber(e.chi2(), e.delta, rho); w = rho[1]; }
      const double wo = w * e.info;
      double omega_r[2];
      for (int i = 0; i < 2; i++) omega_r[i] = -(e.info * e.err[i]) * (e.robust ? rho[1] : 1.0);
      double* bl = &b[size_p + ln_col[e.ln]];
      double* H = &Hll_l[(size_t)e.ln * 16];
      for (int r = 0; r < 4; r++) {
        for (int i = 0; i < 2; i++) bl[r] += Jl[i * 4 + r] * omega_r[i];
        for (int c = 0; c < 4; c++)
          for (int i = 0; i < 2; i++) H[r * 4 + c] += Jl[i * 4 + r] * wo * Jl[i * 4 + c];
      }
      const int h = pose_h[e.kf];
      if (h >= 0) {
        double* bp = &b[6 * h];
        double* P = &Hpp[(size_t)h * 36];
        double* W = &Wl[ei * 24];
        for (int r = 0; r < 6; r++) {
          for (int i = 0; i < 2; i++) bp[r] += Jp[i * 6 + r] * omega_r[i];
          for (int c = 0; c < 6; c++)
            for (int i = 0; i < 2; i++) P[r * 6 + c] += Jp[i * 6 + r] * wo * Jp[i * 6 + c];
          for (int c = 0; c < 4; c++)
            for (int i = 0; i < 2; i++) W[r * 4 + c] += Jp[i * 6 + r] * wo * Jl[i * 4 + c];
        }
      }
    }
  }

  // envelope of the Schur complement: block row j starts at the smallest block i sharing a landmark
  void build_structure() {
    std::vector<int> first_blk(n_ph);
    for (int i = 0; i < n_ph; i++) first_blk[i] = i;
    auto touch = [&](const std::vector<int>& hs) {
      if (hs.empty()) return;
      int mn = *std::min_element(hs.begin(), hs.end());
      for (int h : hs) first_blk[h] = std::min(first_blk[h], mn);
    };
    std::vector<int> hs;
    for (int p : act_pts) {
      hs.clear();
      for (int e = pt_e0[p]; e < pt_e0[p + 1]; e++)
        if (pe_active(pe[e]) && pose_h[pe[e].kf] >= 0) hs.push_back(pose_h[pe[e].kf]);
      touch(hs);
    }
    for (int l : act_lns) {
      hs.clear();
      for (int e = ln_e0[l]; e < ln_e0[l + 1]; e++)
        if (le_active(le[e]) && pose_h[le[e].kf] >= 0) hs.push_back(pose_h[le[e].kf]);
      touch(hs);
    }
    std::vector<int> first(size_p);
    for (int i = 0; i < n_ph; i++)
      for (int r = 0; r < 6; r++) first[6 * i + r] = 6 * first_blk[i];
    S.init(size_p, first);
    structure_built = true;
  }

  // S(block hj row, block hi col) -= M (6x6, representing upper block (hi,hj)); store transposed in the lower skyline
  inline void schur_sub(int hi, int hj, const double M[36]) {
    if (hi == hj) {
      for (int r = 0; r < 6; r++)
        for (int c = 0; c <= r; c++) S.at(6 * hi + r, 6 * hi + c) -= M[r * 6 + c];
    } else {  // hi < hj : upper block (hi,hj) == lower block (hj,hi)^T
      for (int r = 0; r < 6; r++)
        for (int c = 0; c < 6; c++) S.at(6 * hj + c, 6 * hi + r) -= M[r * 6 + c];
    }
  }

  // BlockSolver::solve with lambda already "set" (block_solver.hpp:354-486, setLambda :564-589)
  template <int D>
  void schur_landmark(const double* Hll, const double* bl, double lambda, double* Dinv,
                      const std::vector<std::pair<int, const double*>>& obs /* (h, W 6xD) sorted by h, merged */) {
    double Dm[16];
    for (int i = 0; i < D * D; i++) Dm[i] = Hll[i];
    for (int i = 0; i < D; i++) Dm[i * D + i] += lambda;
    inverse_small(Dm, Dinv, D);
    double db[4];
    for (int i = 0; i < D; i++) {
      double s = 0;
      for (int j = 0; j < D; j++) s += Dinv[i * D + j] * bl[j];
      db[i] = s;
    }
    for (size_t a = 0; a < obs.size(); a++) {
      const int hi = obs[a].first;
      const double* Bi = obs[a].second;
      double BD[24];
      matmul(Bi, Dinv, BD, 6, D, D);
      for (int r = 0; r < 6; r++) {
        double s = 0;
        for (int j = 0; j < D; j++) s += Bi[r * D + j] * db[j];
        coeff[6 * hi + r] += s;
      }
      for (size_t c = a; c < obs.size(); c++) {
        const int hj = obs[c].first;
        const double* Bj = obs[c].second;
        double M[36];
        for (int r = 0; r < 6; r++)
          for (int q = 0; q < 6; q++) {
            double s = 0;
            for (int j = 0; j < D; j++) s += BD[r * D + j] * Bj[q * D + j];
            M[r * 6 + q] = s;
          }
        schur_sub(hi, hj, M);
      }
    }
  }

  // gathers per-landmark merged Hpl blocks (left+right line edges of one KF share a block in g2o)
  std::vector<double> merged;
  template <int D, class EdgeVec>
  void gather_obs(const EdgeVec& ev, int e0, int e1, const std::vector<double>& W,
                  std::vector<std::pair<int, const double*>>& obs, bool (Window::*act)(const typename EdgeVec::value_type&) const) {
    obs.clear();
    merged.clear();
    std::map<int, std::vector<int>> by_h;
    for (int e = e0; e < e1; e++) {
      if (!(this->*act)(ev[e])) continue;
      const int h = pose_h[ev[e].kf];
      if (h < 0) continue;
      by_h[h].push_back(e);
    }
    merged.reserve(by_h.size() * 6 * D);
    std::vector<size_t> offs;
    for (auto& kv : by_h) {
      offs.push_back(merged.size());
      merged.resize(merged.size() + 6 * D, 0.0);
      double* m = &merged[offs.back()];
      for (int e : kv.second)
        for (int i = 0; i < 6 * D; i++) m[i] += W[(size_t)e * 6 * D + i];
    }
    size_t k = 0;
    for (auto& kv : by_h) obs.push_back({kv.first, &merged[offs[k++]]});
  }

  bool solve_system(double lambda) {
    if (!structure_built) build_structure();
    S.zero();
    for (int h = 0; h < n_ph; h++)
      for (int r = 0; r < 6; r++) {
        for (int c = 0; c <= r; c++) S.at(6 * h + r, 6 * h + c) = Hpp[(size_t)h * 36 + r * 6 + c];
        S.at(6 * h + r, 6 * h + r) += lambda;
      }
    coeff.assign(size_p, 0.0);
    const int np = (int)pt_e0.size() - 1, nl = (int)ln_e0.size() - 1;
    Dinv_p.assign((size_t)np * 9, 0.0);
    Dinv_l.assign((size_t)nl * 16, 0.0);
    std::vector<std::pair<int, const double*>> obs;
    for (int p : act_pts) {
      gather_obs<3>(pe, pt_e0[p], pt_e0[p + 1], Wp, obs, &Window::pe_active);
      schur_landmark<3>(&Hll_p[(size_t)p * 9], &b[size_p + pt_col[p]], lambda, &Dinv_p[(size_t)p * 9], obs);
    }
    for (int l : act_lns) {
      gather_obs<4>(le, ln_e0[l], ln_e0[l + 1], Wl, obs, &Window::le_active);
      schur_landmark<4>(&Hll_l[(size_t)l * 16], &b[size_p + ln_col[l]], lambda, &Dinv_l[(size_t)l * 16], obs);
    }
    bschur.resize(size_p);
    for (int i = 0; i < size_p; i++) bschur[i] = b[i] - coeff[i];
    if (size_p > 0) {
      if (!skyline_ldlt(S, false)) return false;
      skyline_solve(S, bschur.data(), x.data());
    }
    // landmarks: xl = Dinv (bl - Hpl^T xp)
    for (int p : act_pts) {
      double cl[3] = {b[size_p + pt_col[p]], b[size_p + pt_col[p] + 1], b[size_p + pt_col[p] + 2]};
      for (int e = pt_e0[p]; e < pt_e0[p + 1]; e++) {
        if (!pe_active(pe[e])) continue;
        const int h = pose_h[pe[e].kf];
        if (h < 0) continue;
        const double* W = &Wp[(size_t)e * 18];
        for (int c = 0; c < 3; c++)
          for (int r = 0; r < 6; r++) cl[c] -= W[r * 3 + c] * x[6 * h + r];
      }
      const double* Di = &Dinv_p[(size_t)p * 9];
      for (int i = 0; i < 3; i++) x[size_p + pt_col[p] + i] = Di[i * 3] * cl[0] + Di[i * 3 + 1] * cl[1] + Di[i * 3 + 2] * cl[2];
    }
    for (int l : act_lns) {
      double cl[4];
      for (int i = 0; i < 4; i++) cl[i] = b[size_p + ln_col[l] + i];
      for (int e = ln_e0[l]; e < ln_e0[l + 1]; e++) {
        if (!le_active(le[e])) continue;
        const int h = pose_h[le[e].kf];
        if (h < 0) continue;
        const double* W = &Wl[(size_t)e * 24];
        for (int c = 0; c < 4; c++)
          for (int r = 0; r < 6; r++) cl[c] -= W[r * 4 + c] * x[6 * h + r];
      }
      const double* Di = &Dinv_l[(size_t)l * 16];
      for (int i = 0; i < 4; i++) {
        double s = 0;
        for (int j = 0; j < 4; j++) s += Di[i * 4 + j] * cl[j];
        x[size_p + ln_col[l] + i] = s;
      }
    }
    return true;
  }

  // SparseOptimizer::update (sparse_optimizer.cpp:422-435) + oplusImpl of the three vertex types
  void update() {
    for (size_t k = 0; k < poses.size(); k++) {
      const int h = pose_h[k];
      if (h < 0) continue;
      poses[k] = pose_mul(pose_exp(&x[6 * h]), poses[k]);
    }
    for (int p : act_pts)
      for (int i = 0; i < 3; i++) pts[3 * p + i] += x[size_p + pt_col[p] + i];
    for (int l : act_lns) line_oplus(lines[l], &x[size_p + ln_col[l]]);
  }

  // ---- Levenberg-Marquardt (optimization_algorithm_levenberg.cpp:61-189) ----
  double lambda = -1, ni = 2;
  int n_bad = 0;
  std::vector<double> chi2_log, lambda_log;
  std::vector<int> trials_log;

  bool terminate() const { return stop && *stop; }

  double compute_lambda_init() const {
    double mx = 0;
    for (int h = 0; h < n_ph; h++)
      for (int j = 0; j < 6; j++) mx = std::max(std::fabs(Hpp[(size_t)h * 36 + j * 7]), mx);
    for (int p : act_pts)
      for (int j = 0; j < 3; j++) mx = std::max(std::fabs(Hll_p[(size_t)p * 9 + j * 4]), mx);
    for (int l : act_lns)
      for (int j = 0; j < 4; j++) mx = std::max(std::fabs(Hll_l[(size_t)l * 16 + j * 5]), mx);
    return 1e-5 * mx;
  }

  // returns 0 = OK, 1 = Terminate
  int lm_solve(int iteration) {
    compute_active_errors();
    double currentChi = active_robust_chi2();
    double tempChi = currentChi;
    const double iniChi = currentChi;
    if (iteration == 0) chi2_log.push_back(currentChi);
    build_system();
    if (iteration == 0) {
      lambda = compute_lambda_init();
      ni = 2;
      n_bad = 0;
    }
    double rho = 0;
    int qmax = 0;
    do {
      // push
      std::vector<Pose> bk_poses = poses;
      std::vector<double> bk_pts = pts;
      std::vector<LineState> bk_lines = lines;
      for (auto& l : bk_lines) line_getq(l, l.q);  // LineParams copy-ctor normalises (types_sba.cpp:64-68)
      bool ok2 = solve_system(lambda);
      update();
      compute_active_errors();
      tempChi = active_robust_chi2();
      if (!ok2) tempChi = std::numeric_limits<double>::max();
      rho = (currentChi - tempChi);
      double scale = 0.;
      for (size_t j = 0; j < x.size(); j++) scale += x[j] * (lambda * x[j] + b[j]);
      scale += 1e-3;
      rho /= scale;
      if (rho > 0 && std::isfinite(tempChi)) {
        double alpha = 1. - std::pow((2 * rho - 1), 3);
        alpha = (std::min)(alpha, 2. / 3.);
        double scaleFactor = (std::max)(1. / 3., alpha);
        lambda *= scaleFactor;
        ni = 2;
        currentChi = tempChi;
      } else {
        lambda *= ni;
        ni *= 2;
        poses = bk_poses;
        pts = bk_pts;
        lines = bk_lines;
      }
      qmax++;
    } while (rho < 0 && qmax < 10 && !terminate());
    chi2_log.push_back(currentChi);
    lambda_log.push_back(lambda);
    trials_log.push_back(qmax);
    if (qmax == 10 || rho == 0) return 1;
    if ((iniChi - currentChi) * 1e3 < iniChi) n_bad++;
    else n_bad = 0;
    if (n_bad >= 3) return 1;
    return 0;
  }

  // SparseOptimizer::optimize (sparse_optimizer.cpp:354-419); returns iterations executed
  int optimize(int iterations) {
    if (size_p + size_l == 0) return -1;
    int done = 0;
    bool ok = true;
    for (int i = 0; i < iterations && !terminate() && ok; i++) {
      ok = (lm_solve(i) == 0);
      done++;
    }
    return done;
  }
};

static void load_window(const lld_ba_problem* p, int w, Window& W) {
  const int k0 = p->kf_off[w], k1 = p->kf_off[w + 1];
  const int p0 = p->pt_off[w], p1 = p->pt_off[w + 1];
  const int l0 = p->ln_off[w], l1 = p->ln_off[w + 1];
  const int nk = k1 - k0, np = p1 - p0, nl = l1 - l0;
  W.poses.resize(nk); W.fixed.resize(nk); W.intr.resize(nk);
  for (int k = 0; k < nk; k++) {
    W.poses[k] = pose_from_Rt(&p->kf_Tcw[(size_t)(k0 + k) * 12]);
    W.fixed[k] = p->kf_fixed[k0 + k];
    const double* in = &p->kf_intr[(size_t)(k0 + k) * 5];
    W.intr[k] = Intr{in[0], in[1], in[2], in[3], in[4]};
  }
  W.pts.assign(p->pt_xyz + (size_t)p0 * 3, p->pt_xyz + (size_t)p1 * 3);
  W.pt_e0.assign(np + 1, 0);
  const int eb = np > 0 ? p->pt_obs_off[p0] : 0;
  for (int i = 0; i <= np; i++) W.pt_e0[i] = (np > 0 ? p->pt_obs_off[p0 + i] : eb) - eb;
  W.pe.clear();
  for (int i = 0; i < np; i++)
    for (int e = p->pt_obs_off[p0 + i]; e < p->pt_obs_off[p0 + i + 1]; e++) {
      PEdge E;
      E.pt = i; E.kf = p->pt_obs_kf[e];
      const float* o = &p->pt_obs_uvr[(size_t)e * 3];
      E.stereo = !(o[2] < 0);
      E.obs[0] = o[0]; E.obs[1] = o[1]; E.obs[2] = o[2];
      E.info = p->pt_obs_info[e];
      E.robust = p->robust_points != 0;
      E.delta = E.stereo ? p->delta_pt_stereo : p->delta_pt_mono;
      W.pe.push_back(E);
    }
  W.lines.resize(nl); W.ln_removed.assign(nl, 0);
  W.ln_e0.assign(nl + 1, 0);
  W.le.clear();
  for (int i = 0; i < nl; i++) {
    const double* xd = &p->ln_x0_dir[(size_t)(l0 + i) * 6];
    W.lines[i] = line_from_x0_dir(xd, xd + 3);
    W.ln_e0[i] = (int)W.le.size();
    for (int c = p->ln_obs_off[l0 + i]; c < p->ln_obs_off[l0 + i + 1]; c++) {
      const int kf = p->ln_obs_kf[c];
      const double* lc = &p->kf_line_cam[(size_t)(k0 + kf) * 4];
      const double* in = &p->kf_intr[(size_t)(k0 + kf) * 5];
      for (int si = 0; si < 2; si++) {
        const float* seg = (si == 0 ? p->ln_obs_left : p->ln_obs_right) + (size_t)c * 4;
        if (si == 1 && seg[0] < 0) continue;
        LEdge E;
        E.ln = i; E.kf = kf; E.cell = c; E.side = si;
        E.cam = LineCam{lc[0], lc[1], lc[2], si == 1 ? -lc[3] : 0.0};
        if (p->ln_endpoints_normalized) {  // K^-1 * (x,y,1)   src/Optimizer.cc:234-235
          E.x1[0] = ((double)seg[0] - in[2]) / in[0]; E.x1[1] = ((double)seg[1] - in[3]) / in[1]; E.x1[2] = 1.0;
          E.x2[0] = ((double)seg[2] - in[2]) / in[0]; E.x2[1] = ((double)seg[3] - in[3]) / in[1]; E.x2[2] = 1.0;
        } else {
          E.x1[0] = seg[0]; E.x1[1] = seg[1]; E.x1[2] = 1.0;
          E.x2[0] = seg[2]; E.x2[1] = seg[3]; E.x2[2] = 1.0;
        }
        E.info = p->ln_obs_info[(size_t)c * 2 + si];
        E.stereo_sel = p->ln_obs_stereo[c] != 0;
        E.delta = E.stereo_sel ? p->delta_ln_stereo : p->delta_ln_mono;
        // AddLineMinimal calls e->computeError() on creation (src/LineOptimizer.cc:114)
        line_err(W.poses[kf], W.lines[i], E.cam, E.x1, E.x2, E.err);
        W.le.push_back(E);
      }
    }
  }
  W.ln_e0[nl] = (int)W.le.size();
}

static void store_state(const lld_ba_problem* p, int w, const Window& W, lld_ba_result* out) {
  const int k0 = p->kf_off[w], p0 = p->pt_off[w], l0 = p->ln_off[w];
  for (size_t k = 0; k < W.poses.size(); k++) pose_to_Rt(W.poses[k], &out->kf_Tcw[(size_t)(k0 + k) * 12]);
  std::copy(W.pts.begin(), W.pts.end(), out->pt_xyz + (size_t)p0 * 3);
  for (size_t l = 0; l < W.lines.size(); l++) {
    double* o = &out->ln_x0_dir[(size_t)(l0 + l) * 6];
    if (W.ln_removed[l]) {  // GetLineData returns false: the caller keeps its own value
      const double* in = &p->ln_x0_dir[(size_t)(l0 + l) * 6];
      for (int i = 0; i < 6; i++) o[i] = in[i];
    } else {
      line_to_x0_dir(W.lines[l], o, o + 3);
    }
  }
}

static void store_logs(int w, const Window& W, int it_r1, int it_r2, lld_ba_result* out) {
  const int st = out->log_stride;
  if (out->chi2_log)
    for (int i = 0; i < st; i++) out->chi2_log[(size_t)w * st + i] = i < (int)W.chi2_log.size() ? W.chi2_log[i] : 0.0;
  if (out->lambda_log)
    for (int i = 0; i < st; i++) out->lambda_log[(size_t)w * st + i] = i < (int)W.lambda_log.size() ? W.lambda_log[i] : 0.0;
  if (out->trials_log)
    for (int i = 0; i < st; i++) out->trials_log[(size_t)w * st + i] = i < (int)W.trials_log.size() ? W.trials_log[i] : 0;
  if (out->n_iter_done) {
    out->n_iter_done[2 * w] = it_r1;
    out->n_iter_done[2 * w + 1] = it_r2;
  }
}

}  // namespace lldo

using namespace lldo;

extern "C" {

// Optimizer::LocalBundleAdjustment  src/Optimizer.cc:1220-1329 (+ LineOptimizer.cc:129-201)
int lldo_ba_local(void*, const lld_ba_problem* p, int its1, int its2, const volatile uint8_t* stop,
                  lld_ba_result* out) {
  for (int w = 0; w < p->n_win; w++) {
    Window W;
    load_window(p, w, W);
    W.stop = stop;
    const int pe0 = p->pt_off[w + 1] > p->pt_off[w] ? p->pt_obs_off[p->pt_off[w]] : 0;
    const int l0 = p->ln_off[w];
    int it1 = 0, it2 = 0;
    if (stop && *stop) {  // :1220-1222 early return: nothing is written back
      store_state(p, w, W, out);
      for (size_t e = 0; e < W.pe.size(); e++) out->pt_obs_bad[pe0 + e] = 0;
      for (auto& e : W.le) out->ln_obs_bad[(size_t)e.cell * 2 + e.side] = 0;
      store_logs(w, W, 0, 0, out);
      continue;
    }
    W.initialize_optimization();
    it1 = std::max(0, W.optimize(its1));
    bool more = !(stop && *stop);
    if (more) {
      for (auto& e : W.pe) {
        const double th = e.stereo ? p->chi2_pt_stereo : p->chi2_pt_mono;
        if (e.chi2() > th || !pt_depth_positive(W.poses[e.kf], &W.pts[3 * e.pt])) e.level = 1;
        e.robust = false;
      }
      // LineOptimizer::DisableOutliers
      std::vector<int> cnt(W.lines.size(), 0);
      for (auto& e : W.le) {
        const double d = e.stereo_sel ? p->delta_ln_stereo : p->delta_ln_mono;
        const double thr = d * d;
        if (e.chi2() > thr || !line_depth_positive(W.poses[e.kf], W.lines[e.ln], e.cam, e.x1, e.x2)) e.level = 1;
        else cnt[e.ln] += 2;
        e.robust = false;
      }
      for (size_t l = 0; l < W.lines.size(); l++)
        if (W.ln_e0[l + 1] > W.ln_e0[l] && cnt[l] <= p->ln_filter) {
          W.ln_removed[l] = 1;
          for (int e = W.ln_e0[l]; e < W.ln_e0[l + 1]; e++) W.le[e].removed = true;
        }
      W.initialize_optimization();
      it2 = std::max(0, W.optimize(its2));
    }
    // final classification (:1281-1311): stale per-edge errors, current vertices
    for (size_t e = 0; e < W.pe.size(); e++) {
      const PEdge& E = W.pe[e];
      const double th = E.stereo ? p->chi2_pt_stereo : p->chi2_pt_mono;
      out->pt_obs_bad[pe0 + e] = (E.chi2() > th || !pt_depth_positive(W.poses[E.kf], &W.pts[3 * E.pt])) ? 1 : 0;
    }
    for (int c = p->ln_obs_off[l0]; c < p->ln_obs_off[p->ln_off[w + 1]]; c++) out->ln_obs_bad[2 * c] = out->ln_obs_bad[2 * c + 1] = 0;
    for (auto& e : W.le) {  // GetLineData: depth first, then recompute the error at the final state
      if (W.ln_removed[e.ln]) continue;
      const bool dp = line_depth_positive(W.poses[e.kf], W.lines[e.ln], e.cam, e.x1, e.x2);
      line_err(W.poses[e.kf], W.lines[e.ln], e.cam, e.x1, e.x2, e.err);
      const double d = e.stereo_sel ? p->delta_ln_stereo : p->delta_ln_mono;
      out->ln_obs_bad[(size_t)e.cell * 2 + e.side] = (e.chi2() > d * d || !dp) ? 1 : 0;
    }
    for (size_t l = 0; l < W.lines.size(); l++) out->ln_removed[l0 + l] = W.ln_removed[l];
    store_state(p, w, W, out);
    store_logs(w, W, it1, it2, out);
  }
  return 0;
}

// Optimizer::BundleAdjustment  src/Optimizer.cc:491-557
int lldo_ba_global(void*, const lld_ba_problem* p, int n_iter, const volatile uint8_t* stop, lld_ba_result* out) {
  for (int w = 0; w < p->n_win; w++) {
    Window W;
    load_window(p, w, W);
    W.stop = stop;
    W.initialize_optimization();
    int it = std::max(0, W.optimize(n_iter));
    const int pe0 = p->pt_off[w + 1] > p->pt_off[w] ? p->pt_obs_off[p->pt_off[w]] : 0;
    for (size_t e = 0; e < W.pe.size(); e++) out->pt_obs_bad[pe0 + e] = 0;
    for (auto& e : W.le) out->ln_obs_bad[(size_t)e.cell * 2 + e.side] = 0;
    for (size_t l = 0; l < W.lines.size(); l++) out->ln_removed[p->ln_off[w] + l] = 0;
    store_state(p, w, W, out);
    store_logs(w, W, it, 0, out);
  }
  return 0;
}

// ---- test hooks: single-edge evaluators so tests can check Jacobians numerically -------------
// kind: 0 mono point, 1 stereo point (binary), 2 line (binary), 3 stereo point (unary residual)
// state: Tcw[12], landmark (3 xyz | 6 x0,dir), intr[5], line_cam[4] (f cx cy bx), obs (2|3 | x1[3],x2[3])
int lldo_edge_eval(int kind, const double* Tcw, const double* lm, const double* intr, const double* lcam,
                   const double* obs, double* err, double* Jl, double* Jp) {
  Pose T = pose_from_Rt(Tcw);
  Intr k{intr[0], intr[1], intr[2], intr[3], intr[4]};
  if (kind == 0) { pt_err_mono(T, lm, obs, k, err); if (Jl) pt_jac_binary(T, lm, k, false, Jl, Jp); return 2; }
  if (kind == 1) { pt_err_stereo_binary(T, lm, obs, k, err); if (Jl) pt_jac_binary(T, lm, k, true, Jl, Jp); return 3; }
  if (kind == 3) { pt_err_stereo_unary(T, lm, obs, k, err); if (Jp) pt_jac_unary(T, lm, k, true, Jp); return 3; }
  if (kind == 2) {
    LineState L = line_from_x0_dir(lm, lm + 3);
    LineCam c{lcam[0], lcam[1], lcam[2], lcam[3]};
    line_err(T, L, c, obs, obs + 3, err);
    if (Jl) line_jac_binary(T, L, c, obs, obs + 3, Jl, Jp);
    return 2;
  }
  return -1;
}
// apply the vertex updates: pose <- exp(u6)*pose ; line <- oplus(u4)
void lldo_pose_oplus(const double* Tcw, const double* u6, double* Tcw_out) {
  Pose T = pose_from_Rt(Tcw);
  T = pose_mul(pose_exp(u6), T);
  pose_to_Rt(T, Tcw_out);
}
void lldo_line_oplus(const double* x0dir, const double* u4, double* x0dir_out) {
  LineState L = line_from_x0_dir(x0dir, x0dir + 3);
  line_oplus(L, u4);
  line_to_x0_dir(L, x0dir_out, x0dir_out + 3);
}
int lldo_line_depth_positive(const double* Tcw, const double* x0dir, const double* lcam, const double* x1x2) {
  Pose T = pose_from_Rt(Tcw);
  LineState L = line_from_x0_dir(x0dir, x0dir + 3);
  LineCam c{lcam[0], lcam[1], lcam[2], lcam[3]};
  return line_depth_positive(T, L, c, x1x2, x1x2 + 3) ? 1 : 0;
}

}  // extern "C"
