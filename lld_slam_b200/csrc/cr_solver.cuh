// cr_solver.cuh — block cyclic reduction of the global-BA reduced camera system (parallel in space).
//
// The reduced camera system of a long trajectory is block banded: keyframe g only shares landmarks with keyframes
// g-B .. g+B (B = longest track).  Grouping B consecutive 6x6 keyframe blocks into one super-block of m = 6B scalars makes
// it block TRIDIAGONAL with N = ceil(n_free / B) dense m x m diagonal blocks D_I and couplings U_I = S(I, I+1).  The
// reference factors this matrix with Eigen::SimplicialLDLT after an AMD ordering
// (Thirdparty/g2o/g2o/solvers/linear_solver_eigen.h:94-124,147-151); the ordering that is parallel in space for a chain is
// odd/even nested dissection = cyclic reduction: at level l (stride s = 2^l) every node i = s (2k+1) is eliminated
// independently of the others,
//     Z_i = D_i^-1 [ U_a^T | U_i | b_i ]                a = i - s,  c = i + s
//     D_a -= U_a Za_i      b_a -= U_a zc_i              D_c -= U_i^T Zb_i     b_c -= U_i^T zc_i      U'_a = - U_a Zb_i
// and the surviving nodes 0, 2s, 4s, ... form a block tridiagonal system again.  log2(N) levels of three kernels
//   k_cr_factor : CTA per eliminated node : dense LDL^T of D_i in shared memory (ldlt_solve_cta), zc_i
//   k_cr_solve  : CTA per (node, 64 right-hand-side columns) : L^-1, D^-1, L^-T on register-resident column slices
//   k_cr_update : CTA per 64x64 output tile of a surviving node : both Schur contributions in fixed order (no atomics)
// then the root node and the back-substitution x_i = zc_i - Za_i x_a - Zb_i x_c level by level.
// A zero / non-finite pivot anywhere clears *ok: the LM trial is rejected like a failed Eigen factorisation
// (optimization_algorithm_levenberg.cpp:126-127).
#pragma once
#include "ba.cuh"

namespace lld {

struct CrView {
  int N;        // super-blocks
  int m;        // scalars per super-block (6 * mb)
  int mb;       // keyframe blocks per super-block
  int zs;       // row stride of Z (2 m)
  double* D;    // [N][m][m] diagonal super-blocks (full storage; the factorisation reads the lower triangle)
  double* U[2]; // [N][m][m] coupling (j, j + stride) stored at its left node; double-buffered across levels
  double* b;    // [N][m]
  double* L;    // [N][m][m] factor of each eliminated node: strict lower = L, diagonal = D
  double* Z;    // [N][m][2m]: Za = D^-1 U_a^T | Zb = D^-1 U_i
  double* zc;   // [N][m]     D^-1 b
  double* x;    // [N][m]     solution
  int* ok;      // [1]
};

constexpr int CR_TC = 64;       // right-hand-side columns per k_cr_solve CTA
constexpr int CR_RG = 8;        // row groups (a thread owns RPT consecutive rows of two columns; RPT = 15: m <= 120, 20: m <= 160)
constexpr int CR_RPT = 20;
constexpr int CR_MAX_M = 156;   // shared-memory capacity of k_cr_factor: (m^2 + 8 m + 40) doubles
constexpr int CR_TILE = 64;     // k_cr_update output tile
constexpr int CR_KC = 16;       // k-chunk of the tile products

// scatter the block rows of S (upper blocks, nb lists) and b_schur into the super-block arrays (zeroed beforehand);
// padding rows of the last super-block get a unit diagonal
__global__ void __launch_bounds__(256) k_cr_assemble(BaView v, CrView cr, int n_blocks) {
  const int w = 0;
  if (v.w_phase[w] == PH_DONE) return;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int g0 = v.w_g0[w], nf = v.w_g0[w + 1] - g0;
  const int m = cr.m, mb = cr.mb;
  if (t == 0) *cr.ok = 1;
  if (t < n_blocks * 36) {
    const int q = t / 36, e = t - 36 * q, r = e / 6, c = e - 6 * r;
    int lo = 0, hi = v.n_free_total;   // owning block row: largest g with nb_off[g] <= q
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (v.nb_off[mid] <= q) lo = mid;
      else hi = mid;
    }
    const int a = lo - g0, h = v.nb_g[q] - g0;   // block (a, h), h >= a
    const double val = v.S_blk[36 * (size_t)q + e];
    const int I = a / mb, J = h / mb;
    const int ra = (a - I * mb) * 6 + r, ch = (h - J * mb) * 6 + c;
    if (I == J) {
      double* Dm = cr.D + (size_t)I * m * m;
      if (a == h) {
        if (c >= r) { Dm[(size_t)ra * m + ch] = val; Dm[(size_t)ch * m + ra] = val; }   // one triangle of the diagonal block, mirrored
      } else {
        Dm[(size_t)ra * m + ch] = val;
        Dm[(size_t)ch * m + ra] = val;
      }
    } else {
      cr.U[0][(size_t)I * m * m + (size_t)ra * m + ch] = val;
    }
  } else {
    const int i = t - n_blocks * 36;
    if (i < cr.N * m) {
      double bv = 0.0;
      if (i < 6 * nf) bv = v.g_bs[6 * (size_t)g0 + i];
      else cr.D[(size_t)(i / m) * m * m + (size_t)(i % m) * m + (i % m)] = 1.0;
      cr.b[i] = bv;
    }
  }
}

// eliminated node i = s (2 blockIdx.x + 1)  (s = 0: the root node 0): LDL^T of D_i, zc_i = D_i^-1 b_i
__global__ void __launch_bounds__(512) k_cr_factor(BaView v, CrView cr, int s) {
  if (v.w_phase[0] == PH_DONE) return;
  extern __shared__ double crsm[];
  __shared__ int flag;
  const int m = cr.m, tid = threadIdx.x, nt = blockDim.x;
  const int i = s == 0 ? 0 : s * (2 * (int)blockIdx.x + 1);
  if (!*cr.ok) return;
  double* A = crsm;
  double* rhs = A + (size_t)m * m;
  double* tmp = rhs + m;
  const double* Dg = cr.D + (size_t)i * m * m;
  for (int k = tid; k < m * m; k += nt) A[k] = Dg[k];
  for (int k = tid; k < m; k += nt) rhs[k] = cr.b[(size_t)i * m + k];
  __syncthreads();
  const bool ok = ldlt_solve_cta(A, m, m, rhs, tmp, &flag);
  __syncthreads();
  if (!ok) {
    if (tid == 0) *cr.ok = 0;
    return;
  }
  double* Lg = cr.L + (size_t)i * m * m;
  for (int k = tid; k < m * m; k += nt) Lg[k] = A[k];
  for (int k = tid; k < m; k += nt) {
    cr.zc[(size_t)i * m + k] = rhs[k];
    if (s == 0) cr.x[k] = rhs[k];
  }
}

// Z_i = D_i^-1 [U_a^T | U_i] for the eliminated nodes of stride s; CTA = (node, tile of CR_TC columns).
// Thread (column pair c, row group rg) keeps rows [rg RPT, (rg+1) RPT) of its two columns in registers; block forward /
// backward substitution with the L block row / column read from shared memory (all lanes of a warp read the same L
// element: broadcast; one load feeds two FMAs), two barriers per block step.
template <int RPT>
__global__ void __launch_bounds__(CR_TC / 2 * CR_RG) k_cr_solve(BaView v, CrView cr, int s, int cur) {
  if (v.w_phase[0] == PH_DONE) return;
  extern __shared__ double crsm[];
  const int m = cr.m, tid = threadIdx.x;
  const int ntile = (2 * m + CR_TC - 1) / CR_TC;
  const int node = blockIdx.x / ntile, tile = blockIdx.x - node * ntile;
  const int i = s * (2 * node + 1);
  if (!*cr.ok) return;
  constexpr int HC = CR_TC / 2;          // column pairs per CTA: thread c owns columns c and c + HC of the tile
  double* Ls = crsm;                     // [m][m]
  double* yp = Ls + (size_t)m * m;       // [RPT][CR_TC] published block
  {
    const double2* Lg = reinterpret_cast<const double2*>(cr.L + (size_t)i * m * m);
    double2* L2 = reinterpret_cast<double2*>(Ls);
    for (int k = tid; k < m * m / 2; k += blockDim.x) L2[k] = Lg[k];   // m is even
  }
  const int c = tid % HC, rg = tid / HC;
  const int r0 = rg * RPT;
  double y[2][RPT];
  {
    const int a = i - s;
    const bool right = i + s < cr.N;
    const double* Ua = cr.U[cur] + (size_t)a * m * m;   // coupling (a, i): rows a, cols i
    const double* Ui = cr.U[cur] + (size_t)i * m * m;   // coupling (i, i + s)
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int col = tile * CR_TC + c + h * HC;
#pragma unroll
      for (int k = 0; k < RPT; k++) {
        const int r = r0 + k;
        double val = 0.0;
        if (col < 2 * m && r < m) {
          if (col < m) val = Ua[(size_t)col * m + r];                   // (U_a^T)[r][col]
          else if (right) val = Ui[(size_t)r * m + (col - m)];
        }
        y[h][k] = val;
      }
    }
  }
  __syncthreads();
  const int ngrp = (m + RPT - 1) / RPT;
  // forward: L y = r
  for (int g = 0; g < ngrp; g++) {
    if (rg == g) {
#pragma unroll
      for (int k = 0; k < RPT; k++) {
        if (r0 + k < m) {
          double a0 = y[0][k], a1 = y[1][k];
#pragma unroll
          for (int q = 0; q < RPT; q++)
            if (q < k) {
              const double l = Ls[(size_t)(r0 + k) * m + r0 + q];
              a0 -= l * y[0][q]; a1 -= l * y[1][q];
            }
          y[0][k] = a0; y[1][k] = a1;
          yp[k * CR_TC + c] = a0; yp[k * CR_TC + c + HC] = a1;
        }
      }
    }
    __syncthreads();
    if (rg > g && rg < ngrp) {
      const int q0 = g * RPT;
#pragma unroll
      for (int q = 0; q < RPT; q++) {
        if (q0 + q < m) {
          const double y0 = yp[q * CR_TC + c], y1 = yp[q * CR_TC + c + HC];
#pragma unroll
          for (int k = 0; k < RPT; k++)
            if (r0 + k < m) {
              const double l = Ls[(size_t)(r0 + k) * m + q0 + q];
              y[0][k] -= l * y0; y[1][k] -= l * y1;
            }
        }
      }
    }
    __syncthreads();
  }
  // diagonal
#pragma unroll
  for (int k = 0; k < RPT; k++)
    if (r0 + k < m) {
      const double id = 1.0 / Ls[(size_t)(r0 + k) * m + r0 + k];
      y[0][k] *= id; y[1][k] *= id;
    }
  // backward: L^T x = y
  for (int g = ngrp - 1; g >= 0; g--) {
    if (rg == g) {
#pragma unroll
      for (int k = RPT - 1; k >= 0; k--) {
        if (r0 + k < m) {
          double a0 = y[0][k], a1 = y[1][k];
#pragma unroll
          for (int q = 0; q < RPT; q++)
            if (q > k && r0 + q < m) {
              const double l = Ls[(size_t)(r0 + q) * m + r0 + k];
              a0 -= l * y[0][q]; a1 -= l * y[1][q];
            }
          y[0][k] = a0; y[1][k] = a1;
          yp[k * CR_TC + c] = a0; yp[k * CR_TC + c + HC] = a1;
        }
      }
    }
    __syncthreads();
    if (rg < g) {
      const int q0 = g * RPT;
#pragma unroll
      for (int q = 0; q < RPT; q++) {
        if (q0 + q < m) {
          const double x0 = yp[q * CR_TC + c], x1 = yp[q * CR_TC + c + HC];
#pragma unroll
          for (int k = 0; k < RPT; k++)
            if (r0 + k < m) {
              const double l = Ls[(size_t)(q0 + q) * m + r0 + k];
              y[0][k] -= l * x0; y[1][k] -= l * x1;
            }
        }
      }
    }
    __syncthreads();
  }
  double* Zg = cr.Z + (size_t)i * m * cr.zs;
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const int col = tile * CR_TC + c + h * HC;
    if (col < 2 * m)
#pragma unroll
      for (int k = 0; k < RPT; k++)
        if (r0 + k < m) Zg[(size_t)(r0 + k) * cr.zs + col] = y[h][k];
  }
}

// acc[u][w] += sum_k A(k, r0 + ty + 16 u) B(k, c0 + tx + 16 w) over one 64x64 tile (256 threads = 16 x 16);
// TA: A is stored [r][k] (row-major m x m), else [k][r].  The k-chunk after the one being multiplied is already in
// registers (global loads issued before the FMAs of the current chunk), so the L2 latency is paid once per tile.
template <bool TA>
__device__ __forceinline__ void cr_tile_product(const double* __restrict__ Ag, int lda, const double* __restrict__ Bg, int ldb,
                                                int m, int r0, int c0, double (*As)[CR_TILE + 4], double (*Bs)[CR_TILE + 4],
                                                double acc[4][4]) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  constexpr int PER = CR_KC * CR_TILE / 256;   // elements of each operand chunk per thread
  double pa[PER], pb[PER];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int q = 0; q < PER; q++) {
      const int e = tid + 256 * q;
      int kk, rr;
      if (TA) { kk = e % CR_KC; rr = e / CR_KC; }      // consecutive threads walk k (contiguous in memory)
      else { rr = e % CR_TILE; kk = e / CR_TILE; }
      const int k = k0 + kk, r = r0 + rr;
      pa[q] = (k < m && r < m) ? (TA ? Ag[(size_t)r * lda + k] : Ag[(size_t)k * lda + r]) : 0.0;
      const int cc = e % CR_TILE, kb = e / CR_TILE;
      const int k2 = k0 + kb, c = c0 + cc;
      pb[q] = (k2 < m && c < m) ? Bg[(size_t)k2 * ldb + c] : 0.0;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < m; k0 += CR_KC) {
    __syncthreads();
#pragma unroll
    for (int q = 0; q < PER; q++) {
      const int e = tid + 256 * q;
      int kk, rr;
      if (TA) { kk = e % CR_KC; rr = e / CR_KC; }
      else { rr = e % CR_TILE; kk = e / CR_TILE; }
      As[kk][rr] = pa[q];
      Bs[e / CR_TILE][e % CR_TILE] = pb[q];
    }
    __syncthreads();
    if (k0 + CR_KC < m) fetch(k0 + CR_KC);
#pragma unroll
    for (int kk = 0; kk < CR_KC; kk++) {
      double a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; u++) { a[u] = As[kk][ty + 16 * u]; b[u] = Bs[kk][tx + 16 * u]; }
#pragma unroll
      for (int u = 0; u < 4; u++)
#pragma unroll
        for (int w2 = 0; w2 < 4; w2++) acc[u][w2] += a[u] * b[w2];
    }
  }
}

// surviving node j = 2 s node: D_j, b_j and the new coupling (j, j + 2s).  blockIdx = node * per + piece;
// pieces [0, T^2) = D tiles, [T^2, 2 T^2) = U tiles, 2 T^2 = rhs.
__global__ void __launch_bounds__(256) k_cr_update(BaView v, CrView cr, int s, int cur) {
  if (v.w_phase[0] == PH_DONE) return;
  if (!*cr.ok) return;
  __shared__ double As[CR_KC][CR_TILE + 4];
  __shared__ double Bs[CR_KC][CR_TILE + 4];
  const int m = cr.m, zs = cr.zs;
  const int T = (m + CR_TILE - 1) / CR_TILE, per = 2 * T * T + 1;
  const int node = blockIdx.x / per, piece = blockIdx.x - node * per;
  const int j = 2 * s * node;
  const int il = j - s, ir = j + s;
  const bool has_l = il >= 0, has_r = ir < cr.N;
  const size_t mm = (size_t)m * m;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  if (piece == 2 * T * T) {  // rhs: b_j -= U_il^T zc_il + U_j zc_ir
    for (int r = tid; r < m; r += 256) {
      double acc = 0.0;
      if (has_l) {
        const double* Ul = cr.U[cur] + (size_t)il * mm;
        const double* z = cr.zc + (size_t)il * m;
        for (int k = 0; k < m; k++) acc += Ul[(size_t)k * m + r] * z[k];
      }
      double acc2 = 0.0;
      if (has_r) {
        const double* Uj = cr.U[cur] + (size_t)j * mm;
        const double* z = cr.zc + (size_t)ir * m;
        for (int k = 0; k < m; k++) acc2 += Uj[(size_t)r * m + k] * z[k];
      }
      cr.b[(size_t)j * m + r] -= acc + acc2;
    }
    return;
  }
  const bool isU = piece >= T * T;
  const int tl = isU ? piece - T * T : piece;
  const int r0 = (tl / T) * CR_TILE, c0 = (tl % T) * CR_TILE;
  double acc[4][4];
#pragma unroll
  for (int u = 0; u < 4; u++)
#pragma unroll
    for (int w2 = 0; w2 < 4; w2++) acc[u][w2] = 0.0;
  if (!isU) {
    if (has_l)   // U_il^T Zb_il : A(k, r) = U_il[k][r], B(k, c) = Z_il[k][m + c]
      cr_tile_product<false>(cr.U[cur] + (size_t)il * mm, m, cr.Z + (size_t)il * m * zs + m, zs, m, r0, c0, As, Bs, acc);
    if (has_r)   // U_j Za_ir : A(k, r) = U_j[r][k], B(k, c) = Z_ir[k][c]
      cr_tile_product<true>(cr.U[cur] + (size_t)j * mm, m, cr.Z + (size_t)ir * m * zs, zs, m, r0, c0, As, Bs, acc);
    double* Dj = cr.D + (size_t)j * mm;
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int w2 = 0; w2 < 4; w2++) {
        const int r = r0 + ty + 16 * u, c = c0 + tx + 16 * w2;
        if (r < m && c < m) Dj[(size_t)r * m + c] -= acc[u][w2];
      }
  } else {
    const bool has_rr = j + 2 * s < cr.N;
    if (has_r && has_rr)  // - U_j Zb_ir
      cr_tile_product<true>(cr.U[cur] + (size_t)j * mm, m, cr.Z + (size_t)ir * m * zs + m, zs, m, r0, c0, As, Bs, acc);
    double* Un = cr.U[cur ^ 1] + (size_t)j * mm;
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int w2 = 0; w2 < 4; w2++) {
        const int r = r0 + ty + 16 * u, c = c0 + tx + 16 * w2;
        if (r < m && c < m) Un[(size_t)r * m + c] = -acc[u][w2];
      }
  }
}

// back-substitution of the nodes of stride s: x_i = zc_i - Za_i x_(i-s) - Zb_i x_(i+s); one warp per row
__global__ void __launch_bounds__(256) k_cr_backsub(BaView v, CrView cr, int s) {
  if (v.w_phase[0] == PH_DONE) return;
  if (!*cr.ok) return;
  const int m = cr.m, zs = cr.zs;
  const int rows_per_cta = 8;
  const int per = (m + rows_per_cta - 1) / rows_per_cta;
  const int node = blockIdx.x / per, chunk = blockIdx.x - node * per;
  const int i = s * (2 * node + 1);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int r = chunk * rows_per_cta + wid;
  if (r >= m) return;
  const double* Zr = cr.Z + (size_t)i * m * zs + (size_t)r * zs;
  const double* xa = cr.x + (size_t)(i - s) * m;
  const bool right = i + s < cr.N;
  const double* xc = cr.x + (size_t)(right ? i + s : 0) * m;
  double acc = 0.0;
  for (int k = lane; k < m; k += 32) {
    acc += Zr[k] * xa[k];
    if (right) acc += Zr[m + k] * xc[k];
  }
  acc = warp_sum(acc);
  if (lane == 0) cr.x[(size_t)i * m + r] = cr.zc[(size_t)i * m + r] - acc;
}

// solution -> g_x (kept on failure), pose update into the trial buffer, pose part of computeScale, ok flag
__global__ void __launch_bounds__(1024) k_cr_finish(BaView v, CrView cr) {
  const int w = 0;
  if (v.w_phase[w] == PH_DONE) return;
  __shared__ double red[32];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int g0 = v.w_g0[w], nf = v.w_g0[w + 1] - g0;
  const int sel = v.w_sel[w];
  const bool ok = *cr.ok != 0;
  if (ok)
    for (int i = tid; i < 6 * nf; i += nt) v.g_x[6 * (size_t)g0 + i] = cr.x[i];
  __syncthreads();
  for (int k = v.kf_off[w] + tid; k < v.kf_off[w + 1]; k += nt) {
    const int g = v.kf_g[k];
    double qt[7];
    const double* src = v.pose_qt[sel] + 7 * (size_t)k;
    if (g >= 0 && v.g_nact[g] > 0) {
      pose_oplus(src, v.g_x + 6 * (size_t)g, qt);
    } else {
#pragma unroll
      for (int q = 0; q < 7; q++) qt[q] = src[q];
    }
    double Rt[12];
    pose_to_Rt(qt, Rt);
    double* dq = v.pose_qt[sel ^ 1] + 7 * (size_t)k;
    double* dr = v.pose_Rt[sel ^ 1] + 12 * (size_t)k;
#pragma unroll
    for (int q = 0; q < 7; q++) dq[q] = qt[q];
#pragma unroll
    for (int q = 0; q < 12; q++) dr[q] = Rt[q];
  }
  const double lam = v.w_lambda[w];
  double sc = 0;
  for (int i = tid; i < 6 * nf; i += nt) {
    const int g = g0 + i / 6;
    if (v.g_nact[g] == 0) continue;
    const double x = v.g_x[6 * (size_t)g0 + i];
    sc += x * (lam * x + v.g_bp[6 * (size_t)g0 + i]);
  }
  sc = block_sum(sc, red);
  if (tid == 0) {
    v.w_scale_p[w] = sc;
    v.w_ok[w] = ok ? 1 : 0;
  }
}

}  // namespace lld
