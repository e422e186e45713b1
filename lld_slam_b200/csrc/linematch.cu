// linematch.cu — stereo line matching with float line descriptors (TwoFrameLineMatcher::MatchLines).
//
// Reference: src/TwoFrameLineMatcher.cc:26-124 with vgl::TriangulateLine (src/vgl.cc:78-108),
// ReprojectKeyLineTo3D (src/LineMatching.cc:277-291), NormalizedLineEquation (src/vgl.cc:578-585).
// The descriptor distance LineMatcher::MatchLineDescriptors lives in the un-vendored LBDMOD library
// (un-vendored and unpinned, so parity is unpinned): defined here as the L2 norm of the float rows.
//
// Device plan per stereo pair:
//   k_line_prep   : per line, K^T-normalised image line equation and pixel length
//   k_line_dist   : 32x32 tiles of the (left x right) pair matrix; descriptors staged in shared memory, dense
//                   ||a-b||^2 contraction in FP32, fused epilogue = the geometric gates of CheckLinePair
//                   (octave, length, triangulation angle, |X0|, endpoint depths) and the tau threshold;
//                   writes the masked distance matrix
//   k_line_greedy : one warp per pair replays the reference's sequential greedy assignment (left lines in index
//                   order, first minimum wins, matched right lines are removed)
// Tolerance (stated, because LBDMOD is unpinned): |d_gpu - d_oracle| <= 1e-5 * max(1, d); identical matches unless
// the two best candidates of a row are closer than that.
#include <cfloat>
#include <vector>
#include <algorithm>

#include "lld_ctx.h"
#include "lld_math.cuh"

using namespace lld;

namespace {

struct LineMatchView {
  int n_pairs, D;
  const int* left_off;
  const int* right_off;
  const float* left_seg;
  const int* left_oct;
  const float* right_seg;
  const int* right_oct;
  const float* left_desc;
  const float* right_desc;
  double K[9];
  double baseline, tau;
  int min_len;
  double* left_leq;   // [n_left][3]
  double* right_leq;
  double* left_len;
  double* right_len;
  const long long* mat_off;  // [n_pairs] offset of the pair's nl x nr matrix
  float* dist;        // masked distance matrix, +inf where a gate fails
  int* match;
  float* mdist;
};

__global__ void k_line_prep(LineMatchView v, int n_left, int n_right) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  for (int side = 0; side < 2; side++) {
    const int n = side ? n_right : n_left;
    if (i >= n) continue;
    const float* s = (side ? v.right_seg : v.left_seg) + 4 * (size_t)i;
    const double Xs[3] = {s[0], s[1], 1.0}, Xe[3] = {s[2], s[3], 1.0};
    double li[3], leq[3];
    cross3(Xs, Xe, li);
#pragma unroll
    for (int k = 0; k < 3; k++) leq[k] = v.K[k] * li[0] + v.K[3 + k] * li[1] + v.K[6 + k] * li[2];  // K^T l
    const double n2 = sqrt(leq[0] * leq[0] + leq[1] * leq[1]);
    double* o = (side ? v.right_leq : v.left_leq) + 3 * (size_t)i;
    o[0] = leq[0] / n2; o[1] = leq[1] / n2; o[2] = leq[2] / n2;
    const double dx = (double)s[0] - (double)s[2], dy = (double)s[1] - (double)s[3];
    (side ? v.right_len : v.left_len)[i] = sqrt(dx * dx + dy * dy);
  }
}

// least squares [ (px,py,1) | -K d ] (depth, s) = K X0 -> s   (vgl::ReprojectLinePointTo3D)
__device__ __forceinline__ double reproject_param(const double* y, const double* c, double px, double py) {
  const double a[3] = {px, py, 1.0};
  const double aa = dot3(a, a), ac = dot3(a, c), cc = dot3(c, c), ay = dot3(a, y), cy = dot3(c, y);
  const double det = aa * cc - ac * ac;
  return (aa * cy - ac * ay) / det;
}

__device__ __forceinline__ bool line_pair_gate(const LineMatchView& v, const float* s1, const double* l1, const double* l2) {
  // vgl::TriangulateLine with T1 = [I|0], T2 = [I|(b,0,0)]
  const double n1n = sqrt(dot3(l1, l1)), n2n = sqrt(dot3(l2, l2));
  if (fabs(dot3(l1, l2)) / n1n / n2n > 0.975) return false;
  double dir[3];
  cross3(l1, l2, dir);
  const double dn = sqrt(dot3(dir, dir));
  dir[0] /= dn; dir[1] /= dn; dir[2] /= dn;
  // rows (n1, n2, dir), rhs (0, n2.t2, 0):  X0 = beta (dir x n1) / (n1 . (n2 x dir))
  const double beta = l2[0] * v.baseline;
  double c1[3], c2[3];
  cross3(dir, l1, c1);
  cross3(l2, dir, c2);
  const double det = dot3(l1, c2);
  if (!(fabs(det) > 0.0)) return false;
  const double X0[3] = {beta * c1[0] / det, beta * c1[1] / det, beta * c1[2] / det};
  if (sqrt(dot3(X0, X0)) < 0.5) return false;
  // ReprojectKeyLineTo3D with T = I: both endpoints must land at z >= 0
  const double* K = v.K;
  const double y[3] = {K[0] * X0[0] + K[1] * X0[1] + K[2] * X0[2], K[3] * X0[0] + K[4] * X0[1] + K[5] * X0[2],
                       K[6] * X0[0] + K[7] * X0[1] + K[8] * X0[2]};
  const double c[3] = {-(K[0] * dir[0] + K[1] * dir[1] + K[2] * dir[2]), -(K[3] * dir[0] + K[4] * dir[1] + K[5] * dir[2]),
                       -(K[6] * dir[0] + K[7] * dir[1] + K[8] * dir[2])};
  const double pa = reproject_param(y, c, s1[0], s1[1]);
  const double pb = reproject_param(y, c, s1[2], s1[3]);
  if (X0[2] + pa * dir[2] < 0 || X0[2] + pb * dir[2] < 0) return false;
  return true;
}

constexpr int LT = 32;  // tile edge

// grid: (tiles_x * tiles_y summed over pairs) flattened through tile_pair / tile_row / tile_col tables
__global__ void __launch_bounds__(256) k_line_dist(LineMatchView v, const int* __restrict__ tile_pair,
                                                   const int* __restrict__ tile_r, const int* __restrict__ tile_c) {
  extern __shared__ float sm[];
  const int p = tile_pair[blockIdx.x];
  const int a0 = v.left_off[p], na = v.left_off[p + 1] - a0;
  const int b0 = v.right_off[p], nb = v.right_off[p + 1] - b0;
  const int r0 = tile_r[blockIdx.x] * LT, c0 = tile_c[blockIdx.x] * LT;
  const int D = v.D, ld = D + 1;
  float* sa = sm;
  float* sb = sm + LT * ld;
  for (int i = threadIdx.x; i < LT * D; i += blockDim.x) {
    const int r = i / D, k = i - r * D;
    sa[r * ld + k] = (r0 + r < na) ? v.left_desc[(size_t)(a0 + r0 + r) * D + k] : 0.f;
    sb[r * ld + k] = (c0 + r < nb) ? v.right_desc[(size_t)(b0 + c0 + r) * D + k] : 0.f;
  }
  __syncthreads();
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, 2 x 2 micro tile
  float acc[2][2] = {{0, 0}, {0, 0}};
  for (int k = 0; k < D; k++) {
    const float x0 = sa[(2 * ty) * ld + k], x1 = sa[(2 * ty + 1) * ld + k];
    const float y0 = sb[(2 * tx) * ld + k], y1 = sb[(2 * tx + 1) * ld + k];
    float d;
    d = x0 - y0; acc[0][0] = fmaf(d, d, acc[0][0]);
    d = x0 - y1; acc[0][1] = fmaf(d, d, acc[0][1]);
    d = x1 - y0; acc[1][0] = fmaf(d, d, acc[1][0]);
    d = x1 - y1; acc[1][1] = fmaf(d, d, acc[1][1]);
  }
  float* M = v.dist + v.mat_off[p];
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const int r = r0 + 2 * ty + i, c = c0 + 2 * tx + j;
      if (r >= na || c >= nb) continue;
      float out = INFINITY;
      const float d = sqrtf(acc[i][j]);
      if (v.left_oct[a0 + r] == v.right_oct[b0 + c] && !(v.left_len[a0 + r] < (double)v.min_len) &&
          !(v.right_len[b0 + c] < (double)v.min_len) && (double)d < v.tau) {
        if (line_pair_gate(v, v.left_seg + 4 * (size_t)(a0 + r), v.left_leq + 3 * (size_t)(a0 + r),
                           v.right_leq + 3 * (size_t)(b0 + c)))
          out = d;
      }
      M[(size_t)r * nb + c] = out;
    }
}

// one warp per pair: sequential greedy over the left lines
__global__ void __launch_bounds__(32) k_line_greedy(LineMatchView v, int* taken_all, const long long* taken_off) {
  const int p = blockIdx.x;
  const int lane = threadIdx.x;
  const int a0 = v.left_off[p], na = v.left_off[p + 1] - a0;
  const int b0 = v.right_off[p], nb = v.right_off[p + 1] - b0;
  const float* M = v.dist + v.mat_off[p];
  int* taken = taken_all + taken_off[p];
  for (int c = lane; c < nb; c += 32) taken[c] = 0;
  __syncwarp();
  for (int j = 0; j < na; j++) {
    float best = INFINITY;
    int bi = -1;
    for (int c = lane; c < nb; c += 32) {
      const float d = M[(size_t)j * nb + c];
      if (d < best && !taken[c]) { best = d; bi = c; }  // strict <: first minimum within the lane's stride
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oi >= 0 && (od < best || (od == best && (bi < 0 || oi < bi)))) { best = od; bi = oi; }
    }
    if (lane == 0) {
      v.match[a0 + j] = bi;
      v.mdist[a0 + j] = bi >= 0 ? best : INFINITY;
      if (bi >= 0) taken[bi] = 1;
    }
    __syncwarp();
  }
}

template <typename T>
int upm(LldCtx* c, T** dst, const T* src, size_t n) {
  cudaError_t e = cudaSuccess;
  T* d = c->alloc<T>(n ? n : 1, &e);
  if (e != cudaSuccess) {
    snprintf(c->err, sizeof(c->err), "cudaMalloc: %s", cudaGetErrorString(e));
    return LLD_ERR_CUDA;
  }
  if (n && src) {
    e = cudaMemcpyAsync(d, src, n * sizeof(T), cudaMemcpyHostToDevice, c->stream);
    if (e != cudaSuccess) {
      snprintf(c->err, sizeof(c->err), "cudaMemcpyAsync H2D: %s", cudaGetErrorString(e));
      return LLD_ERR_CUDA;
    }
  }
  *dst = d;
  return LLD_OK;
}
#define UPM(dst, T, src, n)                                \
  do {                                                     \
    T* _p = nullptr;                                       \
    int _r = upm<T>(c, &_p, (const T*)(src), (size_t)(n)); \
    if (_r) return _r;                                     \
    (dst) = _p;                                            \
  } while (0)

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

}  // namespace

extern "C" int lld_line_match(void* ctx, const lld_line_match_problem* p, lld_line_match_result* out) {
  LldCtx* c = lld_ctx_cast(ctx);
  if (!c || !p || !out) return LLD_ERR_ARG;
  LLD_CUDA(c, cudaSetDevice(c->device));
  c->launches = 0;
  c->pool_reset();
  LLD_ARG(c, p->n_pairs >= 1 && p->desc_dim >= 1 && p->desc_dim <= 512);
  LineMatchView v;
  const int P = p->n_pairs;
  const int n_left = p->left_off[P], n_right = p->right_off[P];
  v.n_pairs = P; v.D = p->desc_dim;
  for (int i = 0; i < 9; i++) v.K[i] = p->K[i];
  v.baseline = p->baseline; v.tau = p->tau; v.min_len = p->min_line_length;
  // tiles + matrix offsets
  std::vector<long long> mat_off(P), taken_off(P);
  std::vector<int> tp, tr, tc;
  long long tot = 0, ttot = 0;
  for (int i = 0; i < P; i++) {
    const int na = p->left_off[i + 1] - p->left_off[i], nb = p->right_off[i + 1] - p->right_off[i];
    mat_off[i] = tot; tot += (long long)na * nb;
    taken_off[i] = ttot; ttot += nb;
    for (int r = 0; r < cdiv(na, LT); r++)
      for (int cc = 0; cc < cdiv(nb, LT); cc++) { tp.push_back(i); tr.push_back(r); tc.push_back(cc); }
  }
  LLD_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
  UPM(v.left_off, int, p->left_off, P + 1);
  UPM(v.right_off, int, p->right_off, P + 1);
  UPM(v.left_seg, float, p->left_seg, 4 * (size_t)n_left);
  UPM(v.left_oct, int, p->left_octave, n_left);
  UPM(v.right_seg, float, p->right_seg, 4 * (size_t)n_right);
  UPM(v.right_oct, int, p->right_octave, n_right);
  UPM(v.left_desc, float, p->left_desc, (size_t)n_left * v.D);
  UPM(v.right_desc, float, p->right_desc, (size_t)n_right * v.D);
  UPM(v.mat_off, long long, mat_off.data(), P);
  long long* d_taken_off;
  UPM(d_taken_off, long long, taken_off.data(), P);
  int *d_tp, *d_tr, *d_tc;
  UPM(d_tp, int, tp.data(), tp.size());
  UPM(d_tr, int, tr.data(), tr.size());
  UPM(d_tc, int, tc.data(), tc.size());
  UPM(v.left_leq, double, nullptr, 3 * (size_t)n_left);
  UPM(v.right_leq, double, nullptr, 3 * (size_t)n_right);
  UPM(v.left_len, double, nullptr, n_left);
  UPM(v.right_len, double, nullptr, n_right);
  UPM(v.dist, float, nullptr, (size_t)tot);
  UPM(v.match, int, nullptr, n_left);
  UPM(v.mdist, float, nullptr, n_left);
  int* d_taken;
  UPM(d_taken, int, nullptr, (size_t)ttot);
  LLD_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
  const int nmax = std::max(std::max(n_left, n_right), 1);
  LLD_LAUNCH(c, k_line_prep, cdiv(nmax, 128), 128, 0, v, n_left, n_right);
  if (!tp.empty()) {
    const size_t smem = sizeof(float) * 2 * LT * (v.D + 1);
    LLD_CUDA(c, cudaFuncSetAttribute(k_line_dist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LLD_LAUNCH(c, k_line_dist, (int)tp.size(), 256, smem, v, d_tp, d_tr, d_tc);
  }
  LLD_LAUNCH(c, k_line_greedy, P, 32, 0, v, d_taken, d_taken_off);
  LLD_CUDA(c, cudaGetLastError());
  LLD_CUDA(c, cudaEventRecord(c->ev[2], c->stream));
  if (n_left) {
    LLD_CUDA(c, cudaMemcpyAsync(out->match, v.match, sizeof(int) * (size_t)n_left, cudaMemcpyDeviceToHost, c->stream));
    if (out->dist) LLD_CUDA(c, cudaMemcpyAsync(out->dist, v.mdist, sizeof(float) * (size_t)n_left, cudaMemcpyDeviceToHost, c->stream));
  }
  LLD_CUDA(c, cudaEventRecord(c->ev[3], c->stream));
  LLD_CUDA(c, cudaStreamSynchronize(c->stream));
  cudaEventElapsedTime(&c->ms_h2d, c->ev[0], c->ev[1]);
  cudaEventElapsedTime(&c->ms_compute, c->ev[1], c->ev[2]);
  cudaEventElapsedTime(&c->ms_d2h, c->ev[2], c->ev[3]);
  return LLD_OK;
}
