// ba_fused.cuh — one-pass linearise -> Schur kernel of the dense-mode (local BA) LM step.
//
// Replaces, for every LM step after the first of a round, the sequence
//   k_lin_points / k_lin_lines (r, J, Huber, H_ll, b_l, W -> HBM)  +  k_lin_poses (recomputes every residual for H_pp, b_p)
//   + k_schur_points / k_schur_lines ((H_ll + lambda I)^-1)  +  k_schur_tile (re-reads every W block)
// by ONE kernel per landmark class.  A persistent CTA takes whole pieces (runs of landmarks with the same co-visibility
// signature, W slots contiguous); per chunk of landmarks
//   P1  thread = one free (landmark, keyframe) edge, found through the W-slot -> edge map (flat, no search): residual,
//       analytic Jacobians, Huber weight; W goes to SHARED memory (and once to HBM for the back-substitution), the edge's
//       H_ll / b_l part to shared memory, its H_pp / b_p part into 28 registers of the thread — the keyframe slot of a
//       thread is the same for every chunk of the piece, so the per-keyframe blocks are summed in registers and reduced
//       once per piece (block_solver.hpp:531-545 / base_binary_edge.hpp:55-120 without atomics);
//   P1b edges to FIXED keyframes (they only feed H_ll, b_l) in passes of FU_FXB threads;
//   P2  thread = landmark: fixed-order sum of its edges, (H_ll + lambda I)^-1 and D^-1 b_l into shared memory and the
//       back-substitution record;
//   P3/P4  Z = W D^-1 and the tile contraction S_ab += W_a Z_b^T from shared memory exactly as k_schur_tile does.
// W never travels HBM -> SM again for the Schur complement and no residual is computed twice.  lambda of a step is known
// when the step starts (it only depends on the previous decision), except in the first step of a round, which still runs
// the separate kernels because lambda_0 = 1e-5 max diag(H) (optimization_algorithm_levenberg.cpp:166-180) needs a full
// linearisation first.
#pragma once
#include "ba_kernels.cuh"

namespace lld {

struct __align__(16) FusedPiece {
  int l0, nl, n, w;      // first landmark (sorted position), landmarks, free keyframes per landmark, window
  long long w0;          // first W slot
  long long out;         // dpart offset of the piece: [npair][36] S blocks, then per keyframe slot FU_KS doubles
  int lc, pad0, pad1, pad2;
};

constexpr int FU_TPB = 128;
constexpr int FU_CAP_W = 2432;     // doubles: W blocks + inverse records of a chunk
constexpr int FU_CAP_Z = 2688;     // doubles: Z blocks (aliased by the per-edge H_ll / b_l parts before Z is formed)
constexpr int FU_FXB = 64;         // fixed-keyframe edges per pass
constexpr int FU_ES = 16;          // doubles per edge part (points use 11, lines 16)
constexpr int FU_SMEM_DOUBLES = FU_CAP_W + FU_CAP_Z + FU_FXB * FU_ES;
constexpr int FU_SMEM_BYTES = FU_SMEM_DOUBLES * 8;
constexpr int FU_KS = SCHUR_KS;    // per keyframe slot of a piece in dpart: b_schur part (6) + H_pp upper (21) + b_p (6) + #edges (1)

// landmarks per chunk of a piece with n free keyframes (host and device agree through FusedPiece::lc)
inline int fused_chunk_len(int D, int n, int nl) {
  const int WS = 6 * D, DS = D == 3 ? 10 : 14;
  if (n == 0) return nl < FU_TPB ? nl : FU_TPB;   // landmarks seen by fixed keyframes only: thread = landmark, no Schur part
  int L = FU_CAP_W / (n * WS + DS);
  const int Lz = FU_CAP_Z / ((n + 1) * WS);
  if (Lz < L) L = Lz;
  if (FU_TPB / n < L) L = FU_TPB / n;
  if (nl < L) L = nl;
  if (L < 1) L = 1;
  return L;
}

// ---- per-edge linearisation, shared by the free and the fixed pass ----------------------------------------------------
// points: part = {H00 H01 H02 H11 H12 H22, b0 b1 b2, rho, active}
template <bool FREE>
__device__ __forceinline__ void fused_point_edge(const BaView& v, int sel, int e, const double* X, double* part /*11*/,
                                                 double* Wv /*18, FREE only*/, double* pacc /*28, FREE only*/, bool acc_pose) {
#pragma unroll
  for (int k = 0; k < 11; k++) part[k] = 0.0;
  if (FREE) {
#pragma unroll
    for (int k = 0; k < 18; k++) Wv[k] = 0.0;
  }
  if (v.pe_level[e] != 0) return;
  const int kf = v.pe_kf[e];
  double Rt[12];
  {
    const double2* rp = reinterpret_cast<const double2*>(v.pose_Rt[sel] + 12 * (size_t)kf);
#pragma unroll
    for (int k = 0; k < 6; k++) {
      const double2 t2 = rp[k];
      Rt[2 * k] = t2.x; Rt[2 * k + 1] = t2.y;
    }
  }
  const double* intr = v.kf_intr + 5 * (size_t)kf;
  const float obs[3] = {v.pe_uvr[3 * (size_t)e], v.pe_uvr[3 * (size_t)e + 1], v.pe_uvr[3 * (size_t)e + 2]};
  const bool stereo = !(obs[2] < 0.f);
  double xc[3], err[3], Jl[9];
  map_Rt(Rt, X, xc);
  pt_residual<true>(xc, intr, obs, stereo, err);
  const double info = (double)v.pe_info[e];
  const double c2 = err[0] * (info * err[0]) + err[1] * (info * err[1]) + err[2] * (info * err[2]);
  double wgt = 1.0, rho = c2;
  if (v.prm.robust_pt) rho = huber(c2, stereo ? v.prm.delta_pt_stereo : v.prm.delta_pt_mono, &wgt);
  const double wo = wgt * info;
  pt_jac_point(xc, Rt, intr, stereo, Jl);
  double JW[9];
#pragma unroll
  for (int k = 0; k < 9; k++) JW[k] = wo * Jl[k];
  part[0] = JW[0] * Jl[0] + JW[3] * Jl[3] + JW[6] * Jl[6];
  part[1] = JW[0] * Jl[1] + JW[3] * Jl[4] + JW[6] * Jl[7];
  part[2] = JW[0] * Jl[2] + JW[3] * Jl[5] + JW[6] * Jl[8];
  part[3] = JW[1] * Jl[1] + JW[4] * Jl[4] + JW[7] * Jl[7];
  part[4] = JW[1] * Jl[2] + JW[4] * Jl[5] + JW[7] * Jl[8];
  part[5] = JW[2] * Jl[2] + JW[5] * Jl[5] + JW[8] * Jl[8];
#pragma unroll
  for (int c = 0; c < 3; c++) part[6 + c] = -(JW[c] * err[0] + JW[3 + c] * err[1] + JW[6 + c] * err[2]);
  part[9] = rho;
  part[10] = 1.0;
  if (FREE) {
    double Jp[18];
    pt_jac_pose(xc, intr, stereo, Jp);
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
      for (int c = 0; c < 3; c++) Wv[r * 3 + c] = Jp[r] * JW[c] + Jp[6 + r] * JW[3 + c] + Jp[12 + r] * JW[6 + c];
    if (acc_pose) {
      int k = 0;
#pragma unroll
      for (int r = 0; r < 6; r++) {
#pragma unroll
        for (int c = r; c < 6; c++, k++) pacc[k] += wo * (Jp[r] * Jp[c] + Jp[6 + r] * Jp[6 + c] + Jp[12 + r] * Jp[12 + c]);
        pacc[21 + r] -= wo * (Jp[r] * err[0] + Jp[6 + r] * err[1] + Jp[12 + r] * err[2]);
      }
      pacc[27] += 1.0;
    }
  }
}

// lines: a "cell" = the left + right edge of one keyframe; part = {H (10 upper), b (4), rho, active}
template <bool FREE>
__device__ __forceinline__ void fused_line_cell(const BaView& v, int sel, int c, const double* st, bool removed, double* part /*16*/,
                                                double* Wv /*24*/, double* pacc /*28*/, bool acc_pose) {
#pragma unroll
  for (int k = 0; k < 16; k++) part[k] = 0.0;
  if (FREE) {
#pragma unroll
    for (int k = 0; k < 24; k++) Wv[k] = 0.0;
  }
  const unsigned lv = *reinterpret_cast<const unsigned short*>(v.lc_level + 2 * (size_t)c);
  const unsigned lv0 = lv & 0xffu, lv1 = lv >> 8;
  if (removed || (lv0 != 0 && lv1 != 0)) return;
  const int kf = v.lc_kf[c];
  double r1[3], r2[3], X1[3], X2[3];
  line_axes(st, r1, r2);
#pragma unroll
  for (int i = 0; i < 3; i++) {
    X1[i] = st[4] * r2[i];
    X2[i] = X1[i] + r1[i];
  }
  double Rt[12];
  {
    const double2* rp = reinterpret_cast<const double2*>(v.pose_Rt[sel] + 12 * (size_t)kf);
#pragma unroll
    for (int k = 0; k < 6; k++) {
      const double2 t2 = rp[k];
      Rt[2 * k] = t2.x; Rt[2 * k + 1] = t2.y;
    }
  }
  const double* cam = v.kf_lcam + 4 * (size_t)kf;
  double P1[3], P2[3];
  map_Rt(Rt, X1, P1);
  map_Rt(Rt, X2, P2);
  const double delta = v.lc_stereo[c] ? v.prm.delta_ln_stereo : v.prm.delta_ln_mono;
#pragma unroll 1
  for (int side = 0; side < 2; side++) {
    if ((side == 0 ? lv0 : lv1) != 0) continue;
    part[15] += 1.0;
    LineObs o;
    make_line_obs(v, kf, (side == 0 ? v.lc_left : v.lc_right) + 4 * (size_t)c, o);
    double err[2], Jp[12], Jl[8];
    line_linearize<true>(P1, P2, cam[0], cam[1], cam[2], side ? -cam[3] : 0.0, o, Rt, X1, X2, r2, err, Jp, Jl);
    const double info = v.lc_info[2 * (size_t)c + side];
    const double c2 = err[0] * (info * err[0]) + err[1] * (info * err[1]);
    double wgt = 1.0, rho = c2;
    if (v.prm.robust_ln) rho = huber(c2, delta, &wgt);
    part[14] += rho;
    const double wo = wgt * info;
    double JW[8];
#pragma unroll
    for (int k = 0; k < 8; k++) JW[k] = wo * Jl[k];
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int cc = r; cc < 4; cc++) part[u4(r, cc)] += JW[r] * Jl[cc] + JW[4 + r] * Jl[4 + cc];
      part[10 + r] -= JW[r] * err[0] + JW[4 + r] * err[1];
    }
    if (FREE) {
#pragma unroll
      for (int r = 0; r < 6; r++)
#pragma unroll
        for (int cc = 0; cc < 4; cc++) Wv[r * 4 + cc] += Jp[r] * JW[cc] + Jp[6 + r] * JW[4 + cc];
      if (acc_pose) {
        int k = 0;
#pragma unroll
        for (int r = 0; r < 6; r++) {
#pragma unroll
          for (int cc = r; cc < 6; cc++, k++) pacc[k] += wo * (Jp[r] * Jp[cc] + Jp[6 + r] * Jp[6 + cc]);
          pacc[21 + r] -= wo * (Jp[r] * err[0] + Jp[6 + r] * err[1]);
        }
        pacc[27] += 1.0;
      }
    }
  }
}

// (H_ll + lambda I)^-1 and D^-1 b_l of one landmark from its summed edge parts Hs = {H upper, b, rho, #active}: into shared
// memory (sd, may be null) and, when `store`, into the arrays the back-substitution kernels read
template <int D>
__device__ __forceinline__ void fused_store_record(const BaView& v, const double* Hs, double lam, int pos, int lm, int lm_base,
                                                   double* Dglob, double* sd, bool store) {
  constexpr int DS = D == 3 ? 10 : 14;
  constexpr int ES = D == 3 ? 11 : 16;
  double rec[DS];
  if (D == 3) {
    const double A[9] = {Hs[0] + lam, Hs[1], Hs[2], Hs[1], Hs[3] + lam, Hs[4], Hs[2], Hs[4], Hs[5] + lam};
    double Di[9];
    inv3_sym(A, Di);
    rec[0] = Di[0]; rec[1] = Di[1]; rec[2] = Di[2]; rec[3] = Di[4]; rec[4] = Di[5]; rec[5] = Di[8];
#pragma unroll
    for (int i = 0; i < 3; i++) rec[6 + i] = Di[3 * i] * Hs[6] + Di[3 * i + 1] * Hs[7] + Di[3 * i + 2] * Hs[8];
    rec[9] = 0.0;
  } else {
    double A[16], Di[16];
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int cc = 0; cc < 4; cc++) A[4 * r + cc] = Hs[r <= cc ? u4(r, cc) : u4(cc, r)] + (r == cc ? lam : 0.0);
    inv4(A, Di);
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int cc = r; cc < 4; cc++) rec[u4(r, cc)] = Di[4 * r + cc];
#pragma unroll
    for (int i = 0; i < 4; i++)
      rec[10 + i] = Di[4 * i] * Hs[10] + Di[4 * i + 1] * Hs[11] + Di[4 * i + 2] * Hs[12] + Di[4 * i + 3] * Hs[13];
  }
  if (sd) {
#pragma unroll
    for (int k = 0; k < DS; k++) sd[k] = rec[k];
  }
  if (store) {
    double* dg = Dglob + (size_t)pos * DS;
#pragma unroll
    for (int k = 0; k < DS; k++) dg[k] = rec[k];
    if (D == 3) {
      double* Ho = v.pt_H + 9 * (size_t)lm;
#pragma unroll
      for (int k = 0; k < 9; k++) Ho[k] = Hs[k];
    } else {
      double* Ho = v.ln_H + 14 * (size_t)lm;
#pragma unroll
      for (int k = 0; k < 14; k++) Ho[k] = Hs[k];
    }
    v.lm_chi2lin[lm_base + lm] = Hs[ES - 2];
    v.lm_active[lm_base + lm] = Hs[ES - 1] > 0.0;
  }
}

// D = 3: map points, D = 4: map lines.  Persistent CTAs; CTA b takes pieces b, b + gridDim.x, ... (cost-sorted on the host).
template <int D>
__global__ void __launch_bounds__(FU_TPB, D == 3 ? 4 : 2) k_fused(BaView v, const FusedPiece* __restrict__ pieces, int n_pieces,
                                                                    const int* __restrict__ ws_edge, const int* __restrict__ fx_off,
                                                                    const int* __restrict__ fx_edge, const int* __restrict__ fx_lm) {
  constexpr int WS = 6 * D;
  constexpr int DS = D == 3 ? 10 : 14;
  constexpr int NH = D == 3 ? 6 : 10;          // packed upper triangle of H_ll
  constexpr int ES = D == 3 ? 11 : 16;         // edge part: H (NH), b (D), rho, active
  constexpr int OFF_C = NH;
  extern __shared__ __align__(16) double fu_smem[];
  double* const sW = fu_smem;                  // [Lc][n][WS] then [Lc][DS]
  double* const sZ = fu_smem + FU_CAP_W;       // [Lc][n + 1][WS]; before P3: edge parts [Lc * n][ES]
  double* const sF = sZ + FU_CAP_Z;            // [FU_FXB][ES]
  const int tid = threadIdx.x;
  const int* lm_sorted = D == 3 ? v.pt_sorted : v.ln_sorted;
  const int lm_base = D == 3 ? 0 : v.n_pt;     // offset into the lm_* arrays
  double* Wglob = D == 3 ? v.pe_Wl : v.lc_Wl;
  double* Dglob = D == 3 ? v.pts_D : v.lns_D;

  for (int pi = blockIdx.x; pi < n_pieces; pi += gridDim.x) {
    const FusedPiece R = pieces[pi];
    if (v.w_phase[R.w] == PH_DONE) continue;     // uniform over the CTA
    const int sel = v.w_sel[R.w];
    const double lam = v.w_lambda[R.w];
    const int n = R.n, nl = R.nl, Lc = R.lc;
    const int npair = n * (n + 1) / 2, ntask = 2 * npair + n;
    const int rowW = n * WS, rowZ = (n + 1) * WS;
    const int nchunk = (nl + Lc - 1) / Lc;
    double* sD = sW + Lc * rowW;
    if (n == 0) {
      // landmarks without a free keyframe: H_ll / b_l from their fixed-keyframe edges, inverse record for the
      // back-substitution (x_l = D^-1 b_l); they contribute nothing to the reduced camera system
      for (int lc0 = 0; lc0 < nl; lc0 += Lc) {
        const int m = min(Lc, nl - lc0), pos0 = R.l0 + lc0;
        double Hs[ES];
#pragma unroll
        for (int k = 0; k < ES; k++) Hs[k] = 0.0;
        const int f0 = fx_off[pos0], f1 = fx_off[pos0 + m];
        for (int fb = f0; fb < f1; fb += FU_FXB) {
          const int cnt = min(FU_FXB, f1 - fb);
          if (tid < cnt) {
            const int e = fx_edge[fb + tid];
            const int lm = lm_sorted[fx_lm[fb + tid]];
            double part[ES];
            if (D == 3) {
              const double* Xp = v.pt_xyz[sel] + 3 * (size_t)lm;
              const double X[3] = {Xp[0], Xp[1], Xp[2]};
              fused_point_edge<false>(v, sel, e, X, part, nullptr, nullptr, false);
            } else {
              const double* stp = v.ln_st[sel] + 5 * (size_t)lm;
              const double st[5] = {stp[0], stp[1], stp[2], stp[3], stp[4]};
              fused_line_cell<false>(v, sel, e, st, v.ln_removed[lm] != 0, part, nullptr, nullptr, false);
            }
            double* pf = sF + (size_t)tid * ES;
#pragma unroll
            for (int k = 0; k < ES; k++) pf[k] = part[k];
          }
          __syncthreads();
          if (tid < m) {
            const int a0 = max(fb, fx_off[pos0 + tid]), a1 = min(fb + cnt, fx_off[pos0 + tid + 1]);
            for (int f = a0; f < a1; f++) {
              const double* pf = sF + (size_t)(f - fb) * ES;
#pragma unroll
              for (int k = 0; k < ES; k++) Hs[k] += pf[k];
            }
          }
          __syncthreads();
        }
        if (tid < m) fused_store_record<D>(v, Hs, lam, pos0 + tid, lm_sorted[pos0 + tid], lm_base, Dglob, nullptr, true);
      }
      continue;
    }
    double pacc[28];
#pragma unroll
    for (int k = 0; k < 28; k++) pacc[k] = 0.0;
    double* out = v.dpart + R.out;

    for (int t0 = 0; t0 < ntask; t0 += FU_TPB) {          // one pass per block of FU_TPB tasks (one pass when n <= 10)
      const bool first = t0 == 0;
      const int ntb = min(FU_TPB, ntask - t0);
      int S = FU_TPB / ntb;
      if (S > SP_MAX_S) S = SP_MAX_S;
      if (S > Lc / 2) S = Lc / 2;
      if (S < 1) S = 1;
      const int s = tid / ntb, tl = tid - s * ntb;
      const bool act = s < S;
      int ia = 0, zb = n, h = 0;
      {
        const int task = t0 + tl;
        if (task < 2 * npair) {
          int pr = task >> 1;
          h = task & 1;
          while (pr >= n - ia) { pr -= n - ia; ia++; }
          zb = ia + pr;
        } else {
          ia = min(task - 2 * npair, n - 1);
        }
      }
      const int zl0 = tid / (n + 1), zb0 = tid - zl0 * (n + 1);
      const int zdl = FU_TPB / (n + 1), zdb = FU_TPB - zdl * (n + 1);
      double acc[18];
#pragma unroll
      for (int q = 0; q < 18; q++) acc[q] = 0.0;

      for (int c = 0; c < nchunk; c++) {
        const int lc0 = c * Lc, m = min(Lc, nl - lc0);
        const int pos0 = R.l0 + lc0;
        // ---- P1: free edges, thread = W slot of the chunk ----
        if (tid < m * n) {
          const int li = tid / n;
          const long long slot = R.w0 + (long long)lc0 * n + tid;
          const int e = ws_edge[slot];
          const int lm = lm_sorted[pos0 + li];
          double part[ES], Wv[WS];
          if (D == 3) {
            const double* Xp = v.pt_xyz[sel] + 3 * (size_t)lm;
            const double X[3] = {Xp[0], Xp[1], Xp[2]};
            fused_point_edge<true>(v, sel, e, X, part, Wv, pacc, first);
          } else {
            const double* stp = v.ln_st[sel] + 5 * (size_t)lm;
            const double st[5] = {stp[0], stp[1], stp[2], stp[3], stp[4]};
            fused_line_cell<true>(v, sel, e, st, v.ln_removed[lm] != 0, part, Wv, pacc, first);
          }
          double2* wd = reinterpret_cast<double2*>(sW + (size_t)tid * WS);
#pragma unroll
          for (int k = 0; k < WS / 2; k++) wd[k] = make_double2(Wv[2 * k], Wv[2 * k + 1]);
          if (first) {   // the back-substitution reads W from HBM (contiguous slots of the chunk)
            double2* wg = reinterpret_cast<double2*>(Wglob + (size_t)slot * WS);
#pragma unroll
            for (int k = 0; k < WS / 2; k++) wg[k] = make_double2(Wv[2 * k], Wv[2 * k + 1]);
          }
          double* pe = sZ + (size_t)tid * ES;
#pragma unroll
          for (int k = 0; k < ES; k++) pe[k] = part[k];
        }
        __syncthreads();
        // ---- P2a: thread = landmark: sum of its free edges in slot order ----
        double Hs[ES];
#pragma unroll
        for (int k = 0; k < ES; k++) Hs[k] = 0.0;
        if (tid < m) {
          for (int a = 0; a < n; a++) {
            const double* pe = sZ + (size_t)(tid * n + a) * ES;
#pragma unroll
            for (int k = 0; k < ES; k++) Hs[k] += pe[k];
          }
        }
        // ---- P1b: edges to fixed keyframes, FU_FXB per pass ----
        const int f0 = fx_off[pos0], f1 = fx_off[pos0 + m];
        for (int fb = f0; fb < f1; fb += FU_FXB) {
          const int cnt = min(FU_FXB, f1 - fb);
          if (tid < cnt) {
            const int e = fx_edge[fb + tid];
            const int lm = lm_sorted[fx_lm[fb + tid]];
            double part[ES];
            if (D == 3) {
              const double* Xp = v.pt_xyz[sel] + 3 * (size_t)lm;
              const double X[3] = {Xp[0], Xp[1], Xp[2]};
              fused_point_edge<false>(v, sel, e, X, part, nullptr, nullptr, false);
            } else {
              const double* stp = v.ln_st[sel] + 5 * (size_t)lm;
              const double st[5] = {stp[0], stp[1], stp[2], stp[3], stp[4]};
              fused_line_cell<false>(v, sel, e, st, v.ln_removed[lm] != 0, part, nullptr, nullptr, false);
            }
            double* pf = sF + (size_t)tid * ES;
#pragma unroll
            for (int k = 0; k < ES; k++) pf[k] = part[k];
          }
          __syncthreads();
          if (tid < m) {
            const int a0 = max(fb, fx_off[pos0 + tid]), a1 = min(fb + cnt, fx_off[pos0 + tid + 1]);
            for (int f = a0; f < a1; f++) {
              const double* pf = sF + (size_t)(f - fb) * ES;
#pragma unroll
              for (int k = 0; k < ES; k++) Hs[k] += pf[k];
            }
          }
          __syncthreads();
        }
        // ---- P2b: (H_ll + lambda I)^-1, D^-1 b_l -> shared memory + back-substitution record ----
        if (tid < m) fused_store_record<D>(v, Hs, lam, pos0 + tid, lm_sorted[pos0 + tid], lm_base, Dglob, sD + (size_t)tid * DS, first);
        __syncthreads();   // edge parts consumed, inverse records visible: sZ may now be overwritten
        // ---- P3: Z = W D^-1 per (landmark, keyframe) + the pseudo block [D^-1 b_l, 0 ...] ----
        for (int i = tid, l = zl0, b = zb0; i < m * (n + 1); i += FU_TPB) {
          const double* Dv = sD + l * DS;
          double* z = sZ + l * rowZ + b * WS;
          if (b < n) {
            const double* wv = sW + l * rowW + b * WS;
            double dm[D][D];
#pragma unroll
            for (int k = 0; k < D; k++)
#pragma unroll
              for (int j = 0; j < D; j++) {
                const int r = k < j ? k : j, c2 = k < j ? j : k;
                dm[k][j] = Dv[r * D - (r * (r - 1)) / 2 + (c2 - r)];
              }
#pragma unroll
            for (int cc = 0; cc < 6; cc++) {
              double wb[D];
#pragma unroll
              for (int j = 0; j < D; j++) wb[j] = wv[cc * D + j];
#pragma unroll
              for (int k = 0; k < D; k++) {
                double zz = 0;
#pragma unroll
                for (int j = 0; j < D; j++) zz += dm[k][j] * wb[j];
                z[cc * D + k] = zz;
              }
            }
          } else {
#pragma unroll
            for (int k = 0; k < WS; k++) z[k] = k < D ? Dv[OFF_C + k] : 0.0;
          }
          l += zdl; b += zdb;
          if (b > n) { b -= n + 1; l++; }
        }
        __syncthreads();
        // ---- P4: tile contraction from shared memory ----
        if (act) {
          const double* wa_p = sW + ia * WS;
          const double* z_p = sZ + zb * WS + h * 3 * D;
#pragma unroll 2
          for (int l = s; l < m; l += S) {
            double wa[WS];
            const double* wp = wa_p + l * rowW;
#pragma unroll
            for (int k = 0; k < WS; k += 2) {
              const double2 t2 = *reinterpret_cast<const double2*>(wp + k);
              wa[k] = t2.x; wa[k + 1] = t2.y;
            }
            const double* zp = z_p + l * rowZ;
#pragma unroll
            for (int cc = 0; cc < 3; cc++) {
              double z[D];
#pragma unroll
              for (int k = 0; k < D; k++) z[k] = zp[cc * D + k];
#pragma unroll
              for (int r = 0; r < 6; r++) {
                double a2 = acc[cc * 6 + r];
#pragma unroll
                for (int k = 0; k < D; k++) a2 += wa[D * r + k] * z[k];
                acc[cc * 6 + r] = a2;
              }
            }
          }
        }
        __syncthreads();
      }
      // ---- task outputs of this pass: slices summed in fixed order through shared memory ----
      {
        double* red = sZ;     // S * ntb * 18 <= FU_TPB * 18 <= FU_CAP_Z
        if (act)
#pragma unroll
          for (int q = 0; q < 18; q++) red[(s * ntb + tl) * 18 + q] = acc[q];
        __syncthreads();
        for (int e = tid; e < ntb * 18; e += FU_TPB) {
          const int tl2 = e / 18, q = e - 18 * tl2;
          const int task = t0 + tl2;
          size_t o;
          if (task < 2 * npair) o = (size_t)(task >> 1) * 36 + (task & 1) * 18 + q;
          else if (q < 6) o = (size_t)36 * npair + (size_t)(task - 2 * npair) * FU_KS + q;
          else continue;
          double sum = 0;
          for (int s2 = 0; s2 < S; s2++) sum += red[(s2 * ntb + tl2) * 18 + q];
          out[o] = sum;
        }
        __syncthreads();
      }
      if (first) {
        // ---- per-keyframe H_pp / b_p of the piece: threads with the same slot (tid mod n), summed in fixed order ----
        double* red = fu_smem;   // [28][FU_TPB + 1]: 3612 doubles <= FU_CAP_W + FU_CAP_Z
        const int P = (FU_TPB / n) * n;
#pragma unroll
        for (int k = 0; k < 28; k++) red[k * (FU_TPB + 1) + tid] = pacc[k];
        __syncthreads();
        for (int e = tid; e < n * 28; e += FU_TPB) {
          const int a = e / 28, k = e - 28 * a;
          double sum = 0;
          for (int t = a; t < P; t += n) sum += red[k * (FU_TPB + 1) + t];
          out[(size_t)36 * npair + (size_t)a * FU_KS + 6 + k] = sum;
        }
        __syncthreads();
      }
    }
  }
}

// S(a,b) = [a==b] (H_pp,a + lambda I) - sum over contributing (piece, pair); b_schur,a = b_p,a - sum of the b parts, with
// H_pp,a / b_p,a / the active-edge count of keyframe a gathered from the pieces' per-slot sums (fixed gather order).
__global__ void __launch_bounds__(256) k_reduce_fused(BaView v, int n_blocks) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n_blocks * 36) {
    const int blk = t / 36, e = t - 36 * blk, c = e / 6, r = e - 6 * c;
    int lo = 0, hi = v.n_free_total;   // owning free block row g: largest g with nb_off[g] <= blk
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (v.nb_off[mid] <= blk) lo = mid;
      else hi = mid;
    }
    const int g = lo, j = blk - v.nb_off[g];
    const int w = v.kf_win[v.g_kf[g]];
    if (v.w_phase[w] == PH_DONE) return;
    double s = 0;
    const int q1 = v.gb_off[blk + 1];
    int q = v.gb_off[blk];
    for (; q + 4 <= q1; q += 4) {
      const long long o0 = v.gb_src[q], o1 = v.gb_src[q + 1], o2 = v.gb_src[q + 2], o3 = v.gb_src[q + 3];
      const double a0 = v.dpart[o0 + e], a1 = v.dpart[o1 + e], a2 = v.dpart[o2 + e], a3 = v.dpart[o3 + e];
      s += a0; s += a1; s += a2; s += a3;
    }
    for (; q < q1; q++) s += v.dpart[v.gb_src[q] + e];
    double d = 0.0;
    if (j == 0) {
      const int rr = r < c ? r : c, cc = r < c ? c : r;
      const int hidx = 6 + (rr * 6 - (rr * (rr - 1)) / 2 + (cc - rr));
      for (int q2 = v.gv_off[g]; q2 < v.gv_off[g + 1]; q2++) d += v.dpart[v.gv_src[q2] + hidx];
      if (r == c) d += v.w_lambda[w];
    }
    v.S_blk[36 * (size_t)blk + 6 * r + c] = d - s;
  } else {
    const int u = t - n_blocks * 36;
    const int g = u / 7, r = u - 7 * g;
    if (g >= v.n_free_total) return;
    const int w = v.kf_win[v.g_kf[g]];
    if (v.w_phase[w] == PH_DONE) return;
    if (r < 6) {
      double s = 0, bp = 0;
      for (int q = v.gv_off[g]; q < v.gv_off[g + 1]; q++) {
        const double* src = v.dpart + v.gv_src[q];
        s += src[r];
        bp += src[6 + 21 + r];
      }
      v.g_bp[6 * (size_t)g + r] = bp;
      v.g_bs[6 * (size_t)g + r] = bp - s;
    } else {
      double cnt = 0;
      for (int q = v.gv_off[g]; q < v.gv_off[g + 1]; q++) cnt += v.dpart[v.gv_src[q] + 6 + 27];
      v.g_nact[g] = (int)(cnt + 0.5);
    }
  }
}

// trial reduction + decision of the fused step: a window that starts a new outer iteration (PH_LIN) carries its chi2 over
// from the accepted trial of the previous step — the same state, so activeRobustChi2() of
// optimization_algorithm_levenberg.cpp:75 would return the same sum — instead of reducing it again.
__global__ void __launch_bounds__(FUSED_RED_TPB) k_decide_carry(BaView v, int round, int stop_now) {
  const int w = blockIdx.x;
  if (v.w_phase[w] == PH_DONE) return;
  __shared__ double sm[32];
  const int tid = threadIdx.x;
  double chi = 0, sc = 0;
  for (int p = v.pt_off[w] + tid; p < v.pt_off[w + 1]; p += FUSED_RED_TPB) {
    chi += v.lm_chi2[p];
    sc += v.lm_scale[p];
  }
  for (int l = v.ln_off[w] + tid; l < v.ln_off[w + 1]; l += FUSED_RED_TPB) {
    chi += v.lm_chi2[v.n_pt + l];
    sc += v.lm_scale[v.n_pt + l];
  }
  chi = block_sum(chi, sm);
  sc = block_sum(sc, sm);
  if (tid == 0) {
    if (v.w_phase[w] == PH_LIN) {   // solve() entry of a new outer iteration
      v.w_inichi[w] = v.w_curchi[w];
      v.w_trials[w] = 0;
    }
    v.w_red_sum[4 * w + 0] = chi;
    v.w_red_sum[4 * w + 1] = sc;
    decide_window(v, w, round, stop_now);
  }
}

// W slot -> edge map of dense mode (inverse of k_dense_wpos)
__global__ void k_ws_edge(int n_e, const int* __restrict__ wpos, int* __restrict__ ws_edge) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_e) return;
  const int s = wpos[e];
  if (s >= 0) ws_edge[s] = e;
}

}  // namespace lld
