// match.cu — 256-bit ORB Hamming matching: both Frame-level ORBmatcher::SearchByProjection variants.
//
// Reference: src/ORBmatcher.cc:45-129 (map points -> frame), :1328-1470 (last frame -> current frame),
// Frame::GetFeaturesInArea / AssignFeaturesToGrid (src/Frame.cc:294-309,391-456), DescriptorDistance (:1647-1663).
//
// Device plan per frame pair:
//   k_cell_count  : PosInGrid cell of every current keypoint, per-cell counts (integer atomics)
//   k_cell_scan   : exclusive scan of the 64x48 cell counts (one CTA per pair)
//   k_cell_fill   : keypoints re-ordered by cell into a packed, coalesced copy (xy, octave, uRight, descriptor)
//   k_match       : 8 lanes per query; the window's cell columns are contiguous ranges of the packed copy;
//                   uint4 descriptor loads + __popc; lexicographic (distance, grid order) top-2 so that ties
//                   resolve exactly like the reference's sequential scan (cell x outer, y inner, insertion order)
//   k_claim / k_match again: the reference lets an accepted match "claim" its keypoint for later queries
//                   (src/ORBmatcher.cc:87-89,1403-1405); we iterate the parallel matcher to the fixed point of that
//                   sequential rule (query i ignores keypoints owned by an accepted query j < i).
//   k_finalize    : rotation histogram + ComputeThreeMaxima (src/ORBmatcher.cc:1431-1466,1601-1642), match table.
// Float arithmetic is kept in float with explicit round-to-nearest intrinsics (no FMA contraction) wherever the
// reference computes in float; compile this unit with --fmad=false.
#include <algorithm>
#include <climits>

#include "lld_ctx.h"

namespace {

constexpr int GRID_COLS = 64, GRID_ROWS = 48, N_CELLS = GRID_COLS * GRID_ROWS;
constexpr int TH_HIGH = 100, HISTO_LENGTH = 30;

struct MatchView {
  int n_pairs, n_cur, n_q;
  int variant;  // 0: frame-frame (best only), 1: map points (best / second best + ratio)
  // geometry
  float fx, fy, cx, cy, bf, b;
  float min_x, max_x, min_y, max_y, winv, hinv;
  int n_levels;
  const float* scale;
  float th, nn_ratio;
  int mono, check_ori;
  // current keypoints (input order)
  const int* cur_off;
  const float* cur_xy;
  const uint8_t* cur_octave;
  const float* cur_angle;
  const float* cur_uright;
  const uint8_t* cur_desc;
  const uint8_t* cur_claimed;
  // packed by cell
  int* cell_count;  // [n_pairs][N_CELLS+1] -> start offsets after the scan
  int* cell_fill;   // [n_pairs][N_CELLS]
  int* kp_cell;     // [n_cur] cell of each keypoint or -1
  uint4* s_rec;     // packed by cell, 48 B per keypoint: {x, y, uRight, meta} + 32 B descriptor, 16-byte aligned
                    // meta = idx (16) | cell (12) << 16 | octave (3) << 28 | claimed-on-entry << 31
  // queries
  const int* q_off;
  const uint8_t* q_valid;
  const float* q_xw;       // v0: world point ; v1: proj (u, v, uR)
  const uint8_t* q_octave; // v0
  const int* q_level;      // v1
  const float* q_viewcos;  // v1
  const float* q_angle;    // v0
  const uint8_t* q_desc;
  const uint8_t* q_has_obs;
  const float* cur_Tcw;    // v0 [n_pairs][12]
  const float* last_Tcw;
  // results / iteration state
  int* owner;      // [n_cur] smallest accepted query (with observations) that matched the keypoint; -1 = claimed on entry
  int* q_best;     // [n_q] matched keypoint (pair-local, original index) or -1
  int* q_dist;
  int* changed;    // [1]
  int* q_pair;     // [n_q] pair of the query
  int* pair_mode;  // [n_pairs] v0: 0 = levels [oct-1, oct+1], 1 = forward, 2 = backward
  unsigned long long* q_top;  // [n_q][4] four best candidates, packed (dist | cell | idx | octave), ascending
  int* q_ncand;    // [n_q] number of admissible candidates seen by the scan
  int* match;      // [n_cur]
  int* n_matches;  // [n_pairs]
};

__device__ __forceinline__ int popc256(const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1) {
  return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) + __popc(a1.x ^ b1.x) +
         __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

__device__ __forceinline__ int find_pair(const int* off, int n_pairs, int i) {
  int lo = 0, hi = n_pairs;  // largest p with off[p] <= i
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (off[mid] <= i) lo = mid;
    else hi = mid;
  }
  return lo;
}

__global__ void k_cell_count(MatchView v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= v.n_cur) return;
  const int p = find_pair(v.cur_off, v.n_pairs, i);
  // Frame::PosInGrid  src/Frame.cc:446-456
  const float x = v.cur_xy[2 * (size_t)i], y = v.cur_xy[2 * (size_t)i + 1];
  const int px = (int)roundf(__fmul_rn(__fsub_rn(x, v.min_x), v.winv));
  const int py = (int)roundf(__fmul_rn(__fsub_rn(y, v.min_y), v.hinv));
  int cell = -1;
  if (!(px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS)) {
    cell = px * GRID_ROWS + py;
    atomicAdd(&v.cell_count[(size_t)p * (N_CELLS + 1) + cell], 1);
  }
  v.kp_cell[i] = cell;
  v.owner[i] = v.cur_claimed[i] ? -1 : INT_MAX;
  v.match[i] = -1;
}

__global__ void __launch_bounds__(1024) k_cell_scan(MatchView v) {
  const int p = blockIdx.x;
  int* cnt = v.cell_count + (size_t)p * (N_CELLS + 1);
  __shared__ int part[1024];
  const int t = threadIdx.x;
  // 3 cells per thread
  const int c0 = cnt[3 * t], c1 = cnt[3 * t + 1], c2 = cnt[3 * t + 2];
  part[t] = c0 + c1 + c2;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const int x = (t >= o) ? part[t - o] : 0;
    __syncthreads();
    part[t] += x;
    __syncthreads();
  }
  const int base = part[t] - (c0 + c1 + c2);
  cnt[3 * t] = base;
  cnt[3 * t + 1] = base + c0;
  cnt[3 * t + 2] = base + c0 + c1;
  if (t == 1023) cnt[N_CELLS] = part[t];
  v.cell_fill[(size_t)p * N_CELLS + 3 * t] = 0;
  v.cell_fill[(size_t)p * N_CELLS + 3 * t + 1] = 0;
  v.cell_fill[(size_t)p * N_CELLS + 3 * t + 2] = 0;
}

__global__ void k_cell_fill(MatchView v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= v.n_cur) return;
  const int cell = v.kp_cell[i];
  if (cell < 0) return;
  const int p = find_pair(v.cur_off, v.n_pairs, i);
  const int slot = atomicAdd(&v.cell_fill[(size_t)p * N_CELLS + cell], 1);
  const size_t d = (size_t)v.cur_off[p] + v.cell_count[(size_t)p * (N_CELLS + 1) + cell] + slot;
  const unsigned meta = (unsigned)(i - v.cur_off[p]) | ((unsigned)cell << 16) | ((unsigned)(v.cur_octave[i] & 7) << 28) |
                        (v.cur_claimed[i] ? 0x80000000u : 0u);
  uint4 h;
  h.x = __float_as_uint(v.cur_xy[2 * (size_t)i]);
  h.y = __float_as_uint(v.cur_xy[2 * (size_t)i + 1]);
  h.z = __float_as_uint(v.cur_uright[i]);
  h.w = meta;
  const uint4* src = reinterpret_cast<const uint4*>(v.cur_desc + 32 * (size_t)i);
  v.s_rec[3 * d] = h;
  v.s_rec[3 * d + 1] = src[0];
  v.s_rec[3 * d + 2] = src[1];
}

// packed candidate key: lexicographic (distance, grid traversal order) == integer order
//   [dist:9 | cell:12 | idx:16 | octave:3]   (cell = ix*48+iy : x outer, y inner; idx = insertion order inside a cell)
constexpr unsigned long long EMPTY_KEY = ~0ull;
__device__ __forceinline__ unsigned long long pack_key(int d, int cell, int idx, int oct) {
  return ((unsigned long long)d << 31) | ((unsigned long long)cell << 19) | ((unsigned long long)idx << 3) | (unsigned long long)oct;
}
__device__ __forceinline__ int key_dist(unsigned long long k) { return (int)(k >> 31); }
__device__ __forceinline__ int key_idx(unsigned long long k) { return (int)((k >> 3) & 0xFFFFull); }
__device__ __forceinline__ int key_oct(unsigned long long k) { return (int)(k & 7ull); }

// cv::Mat (CV_32F) row of R*x + t : double accumulation, one rounding
__device__ __forceinline__ float gemm_row(const float* R, const float* x, float t) {
  const double s = __dadd_rn(__dadd_rn(__dmul_rn((double)R[0], (double)x[0]), __dmul_rn((double)R[1], (double)x[1])),
                             __dmul_rn((double)R[2], (double)x[2]));
  return (float)__dadd_rn(s, (double)t);
}

// per pair: forward / backward decision of the frame-to-frame variant (src/ORBmatcher.cc:1340-1350)
__global__ void k_pair_prep(MatchView v) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= v.n_pairs) return;
  int mode = 0;
  if (v.variant == 0) {
    const float* Tc = v.cur_Tcw + 12 * (size_t)p;
    const float* Tl = v.last_Tcw + 12 * (size_t)p;
    float twc[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const double s = __dadd_rn(__dadd_rn(__dmul_rn((double)Tc[i], (double)Tc[9]), __dmul_rn((double)Tc[3 + i], (double)Tc[10])),
                                 __dmul_rn((double)Tc[6 + i], (double)Tc[11]));
      twc[i] = (float)(-s);
    }
    const float tlc2 = gemm_row(Tl + 6, twc, Tl[11]);
    const bool fwd = tlc2 > v.b && !v.mono;
    const bool bwd = -tlc2 > v.b && !v.mono;
    mode = fwd ? 1 : (bwd ? 2 : 0);
  }
  v.pair_mode[p] = mode;
}
__global__ void k_query_pair(MatchView v) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= v.n_q) return;
  v.q_pair[q] = find_pair(v.q_off, v.n_pairs, q);
}

struct QueryWin {
  bool valid;
  float x, y, r, urq;
  int minLevel, maxLevel;
};
__device__ __forceinline__ QueryWin query_window(const MatchView& v, int q, int p) {
  QueryWin w;
  w.valid = v.q_valid[q] != 0;
  w.x = w.y = w.r = w.urq = 0.f;
  w.minLevel = w.maxLevel = -1;
  if (!w.valid) return w;
  if (v.variant == 0) {
    const float* Tc = v.cur_Tcw + 12 * (size_t)p;
    const float* Xw = v.q_xw + 3 * (size_t)q;
    const float xc = gemm_row(Tc, Xw, Tc[9]);
    const float yc = gemm_row(Tc + 3, Xw, Tc[10]);
    const float zc = gemm_row(Tc + 6, Xw, Tc[11]);
    const float invzc = (float)(1.0 / (double)zc);
    if (invzc < 0) w.valid = false;
    w.x = __fadd_rn(__fmul_rn(__fmul_rn(v.fx, xc), invzc), v.cx);
    w.y = __fadd_rn(__fmul_rn(__fmul_rn(v.fy, yc), invzc), v.cy);
    if (w.x < v.min_x || w.x > v.max_x) w.valid = false;
    if (w.y < v.min_y || w.y > v.max_y) w.valid = false;
    const int oct = v.q_octave[q];
    w.r = __fmul_rn(v.th, v.scale[oct]);
    w.urq = __fsub_rn(w.x, __fmul_rn(v.bf, invzc));
    const int mode = v.pair_mode[p];
    if (mode == 1) { w.minLevel = oct; w.maxLevel = -1; }
    else if (mode == 2) { w.minLevel = 0; w.maxLevel = oct; }
    else { w.minLevel = oct - 1; w.maxLevel = oct + 1; }
  } else {
    const float* pj = v.q_xw + 3 * (size_t)q;
    w.x = pj[0]; w.y = pj[1]; w.urq = pj[2];
    const int lvl = v.q_level[q];
    float rr = ((double)v.q_viewcos[q] > 0.998) ? 2.5f : 4.0f;  // RadiusByViewingCos :131-137
    if (v.th != 1.0f) rr = __fmul_rn(rr, v.th);
    w.r = __fmul_rn(rr, v.scale[lvl]);
    w.minLevel = lvl - 1; w.maxLevel = lvl;
  }
  return w;
}

// Scan the window of one query (one thread).  excl_below >= 0: skip keypoints owned by a query < excl_below.
// Keeps the four smallest keys in t[0..3]; returns the number of admissible candidates.
__device__ __forceinline__ int scan_window(const MatchView& v, int q, int p, const QueryWin& w, int excl_below,
                                           unsigned long long* t) {
  t[0] = t[1] = t[2] = t[3] = EMPTY_KEY;
  if (!w.valid) return 0;
  // Frame::GetFeaturesInArea window  src/Frame.cc:396-410
  const int x0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(w.x, v.min_x), w.r), v.winv)));
  const int x1 = min(GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(w.x, v.min_x), w.r), v.winv)));
  const int y0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(w.y, v.min_y), w.r), v.hinv)));
  const int y1 = min(GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(w.y, v.min_y), w.r), v.hinv)));
  if (!(x0 < GRID_COLS && x1 >= 0 && y0 < GRID_ROWS && y1 >= 0)) return 0;
  const bool check_levels = (w.minLevel > 0) || (w.maxLevel >= 0);
  const uint4* qd = reinterpret_cast<const uint4*>(v.q_desc + 32 * (size_t)q);
  const uint4 a0 = qd[0], a1 = qd[1];
  const int* cstart = v.cell_count + (size_t)p * (N_CELLS + 1);
  const size_t base = (size_t)v.cur_off[p];
  const int* own = v.owner + base;
  int ncand = 0;
  const uint4* rec = v.s_rec + 3 * base;
  for (int ix = x0; ix <= x1; ix++) {
    const int s0 = cstart[ix * GRID_ROWS + y0], s1 = cstart[ix * GRID_ROWS + y1 + 1];
#pragma unroll 2
    for (int s = s0; s < s1; s++) {
      const uint4 h = rec[3 * (size_t)s];
      const uint4 b0 = rec[3 * (size_t)s + 1], b1 = rec[3 * (size_t)s + 2];
      const unsigned meta = h.w;
      if (meta & 0x80000000u) continue;  // claimed on entry
      const int oct = (meta >> 28) & 7;
      if (check_levels) {
        if (oct < w.minLevel) continue;
        if (w.maxLevel >= 0 && oct > w.maxLevel) continue;
      }
      const float dx = __fsub_rn(__uint_as_float(h.x), w.x), dy = __fsub_rn(__uint_as_float(h.y), w.y);
      if (!(fabsf(dx) < w.r && fabsf(dy) < w.r)) continue;
      const int idx = meta & 0xFFFF;
      if (excl_below >= 0 && own[idx] < excl_below) continue;  // owned by an earlier accepted query
      const float ur = __uint_as_float(h.z);
      if (ur > 0) {
        const float er = fabsf(__fsub_rn(w.urq, ur));
        if (er > w.r) continue;
      }
      const int d = popc256(a0, a1, b0, b1);
      ncand++;
      unsigned long long k = pack_key(d, (meta >> 16) & 0xFFF, idx, oct);
      if (k < t[3]) {
        t[3] = k;
        if (t[3] < t[2]) { const unsigned long long u = t[2]; t[2] = t[3]; t[3] = u; }
        if (t[2] < t[1]) { const unsigned long long u = t[1]; t[1] = t[2]; t[2] = u; }
        if (t[1] < t[0]) { const unsigned long long u = t[0]; t[0] = t[1]; t[1] = u; }
      }
    }
  }
  return ncand;
}

// pass 0: one thread per query scans its window once and caches the four best candidates
__global__ void __launch_bounds__(128) k_match_scan(MatchView v) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= v.n_q) return;
  const int p = v.q_pair[q];
  const QueryWin w = query_window(v, q, p);
  unsigned long long t[4];
  const int n = scan_window(v, q, p, w, -1, t);
  ulonglong2* dst = reinterpret_cast<ulonglong2*>(v.q_top + 4 * (size_t)q);
  dst[0] = make_ulonglong2(t[0], t[1]);
  dst[1] = make_ulonglong2(t[2], t[3]);
  v.q_ncand[q] = n;
}

// resolution pass: best (and second best) candidate not owned by an earlier accepted query, from the cached top-4;
// a rescan is only needed when the cache is exhausted although more candidates exist
__global__ void __launch_bounds__(128) k_match_resolve(MatchView v, int pass) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= v.n_q) return;
  const int p = v.q_pair[q];
  const int qi = q - v.q_off[p];
  const int* own = v.owner + (size_t)v.cur_off[p];
  const ulonglong2* src = reinterpret_cast<const ulonglong2*>(v.q_top + 4 * (size_t)q);
  const ulonglong2 ta = src[0], tb = src[1];
  unsigned long long t[4] = {ta.x, ta.y, tb.x, tb.y};
  const int need = v.variant == 1 ? 2 : 1;
  unsigned long long k1 = EMPTY_KEY, k2 = EMPTY_KEY;
  int found = 0;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    if (t[j] == EMPTY_KEY || found >= need) continue;
    if (own[key_idx(t[j])] < qi) continue;
    if (found == 0) k1 = t[j];
    else k2 = t[j];
    found++;
  }
  if (found < need && v.q_ncand[q] > 4) {  // cache exhausted: scan again with the exclusion applied
    const QueryWin w = query_window(v, q, p);
    scan_window(v, q, p, w, qi, t);
    k1 = t[0];
    k2 = t[1];
  }
  int best = -1, bdist = 256;
  if (k1 != EMPTY_KEY && key_dist(k1) <= TH_HIGH) {
    bool ok = true;
    if (v.variant == 1) {
      // ratio test only when best and second best share the level  (src/ORBmatcher.cc:118-121)
      const int d2 = k2 != EMPTY_KEY ? key_dist(k2) : 256;
      const int l2 = k2 != EMPTY_KEY ? key_oct(k2) : -1;
      if (key_oct(k1) == l2 && (float)key_dist(k1) > __fmul_rn(v.nn_ratio, (float)d2)) ok = false;
    }
    if (ok) { best = key_idx(k1); bdist = key_dist(k1); }
  }
  if (pass > 0 && v.q_best[q] != best) atomicOr(v.changed, 1);
  v.q_best[q] = best;
  v.q_dist[q] = bdist;
}

__global__ void k_owner_reset(MatchView v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= v.n_cur) return;
  v.owner[i] = v.cur_claimed[i] ? -1 : INT_MAX;
  if (i == 0) *v.changed = 0;
}
__global__ void k_claim(MatchView v) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= v.n_q) return;
  const int b = v.q_best[q];
  if (b < 0 || !v.q_has_obs[q]) return;
  const int p = v.q_pair[q];
  atomicMin(&v.owner[(size_t)v.cur_off[p] + b], q - v.q_off[p]);
}

// per pair: match table, rotation-consistency filter, counts
__global__ void __launch_bounds__(256) k_finalize(MatchView v) {
  const int p = blockIdx.x;
  __shared__ int hist[HISTO_LENGTH];
  __shared__ int keep[3];
  __shared__ int cnt_acc, cnt_rej;
  const int q0 = v.q_off[p], q1 = v.q_off[p + 1];
  const size_t base = (size_t)v.cur_off[p];
  if (threadIdx.x < HISTO_LENGTH) hist[threadIdx.x] = 0;
  if (threadIdx.x == 0) { cnt_acc = 0; cnt_rej = 0; keep[0] = keep[1] = keep[2] = -1; }
  __syncthreads();
  const bool ori = v.variant == 0 && v.check_ori;
  const float factor = 1.0f / HISTO_LENGTH;
  for (int q = q0 + threadIdx.x; q < q1; q += blockDim.x) {
    const int b = v.q_best[q];
    if (b < 0) continue;
    atomicAdd(&cnt_acc, 1);
    atomicMax(&v.match[base + b], q - q0);
    if (ori) {
      float rot = __fsub_rn(v.q_angle[q], v.cur_angle[base + b]);
      if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
      int bin = (int)roundf(__fmul_rn(rot, factor));
      if (bin == HISTO_LENGTH) bin = 0;
      atomicAdd(&hist[bin], 1);
    }
  }
  __syncthreads();
  if (ori) {
    if (threadIdx.x == 0) {  // ORBmatcher::ComputeThreeMaxima
      int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
      for (int i = 0; i < HISTO_LENGTH; i++) {
        const int s = hist[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
      }
      if ((float)max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
      else if ((float)max3 < 0.1f * (float)max1) { ind3 = -1; }
      keep[0] = ind1; keep[1] = ind2; keep[2] = ind3;
    }
    __syncthreads();
    for (int q = q0 + threadIdx.x; q < q1; q += blockDim.x) {
      const int b = v.q_best[q];
      if (b < 0) continue;
      float rot = __fsub_rn(v.q_angle[q], v.cur_angle[base + b]);
      if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
      int bin = (int)roundf(__fmul_rn(rot, factor));
      if (bin == HISTO_LENGTH) bin = 0;
      if (bin != keep[0] && bin != keep[1] && bin != keep[2]) {
        v.match[base + b] = -1;  // runs after every atomicMax above (barrier) -> NULL wins, as in the reference
        atomicAdd(&cnt_rej, 1);
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) v.n_matches[p] = cnt_acc - cnt_rej;
}

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

template <typename T>
int up(LldCtx* c, T** dst, const T* src, size_t n) {
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
#define UPC(dst, T, src, n)                              \
  do {                                                   \
    T* _p = nullptr;                                     \
    int _r = up<T>(c, &_p, (const T*)(src), (size_t)(n)); \
    if (_r) return _r;                                   \
    (dst) = _p;                                          \
  } while (0)

struct MatchState {
  MatchView v{};
};

}  // namespace

// device pipeline shared by both variants; everything already uploaded into v
static int match_run(LldCtx* c, MatchView& v, int* passes_out) {
  LLD_CUDA(c, cudaMemsetAsync(v.cell_count, 0, sizeof(int) * (size_t)v.n_pairs * (N_CELLS + 1), c->stream));
  if (v.n_cur) LLD_LAUNCH(c, k_cell_count, cdiv(v.n_cur, 256), 256, 0, v);
  LLD_LAUNCH(c, k_cell_scan, v.n_pairs, 1024, 0, v);
  if (v.n_cur) LLD_LAUNCH(c, k_cell_fill, cdiv(v.n_cur, 256), 256, 0, v);
  int* h_changed = reinterpret_cast<int*>(c->pinned);
  int pass = 0;
  const int max_pass = 64;
  if (v.n_q) {
    LLD_LAUNCH(c, k_pair_prep, cdiv(v.n_pairs, 128), 128, 0, v);
    LLD_LAUNCH(c, k_query_pair, cdiv(v.n_q, 256), 256, 0, v);
    LLD_LAUNCH(c, k_match_scan, cdiv(v.n_q, 128), 128, 0, v);
    while (true) {
      LLD_LAUNCH(c, k_match_resolve, cdiv(v.n_q, 128), 128, 0, v, pass);
      if (pass > 0) {
        LLD_CUDA(c, cudaMemcpyAsync(h_changed, v.changed, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        LLD_CUDA(c, cudaStreamSynchronize(c->stream));
        if (!*h_changed) break;
      }
      if (pass >= max_pass) {
        snprintf(c->err, sizeof(c->err), "claim resolution did not converge in %d passes", max_pass);
        return LLD_ERR_CUDA;
      }
      if (v.n_cur) LLD_LAUNCH(c, k_owner_reset, cdiv(v.n_cur, 256), 256, 0, v);
      LLD_LAUNCH(c, k_claim, cdiv(v.n_q, 256), 256, 0, v);
      pass++;
    }
  }
  LLD_LAUNCH(c, k_finalize, v.n_pairs, 256, 0, v);
  LLD_CUDA(c, cudaGetLastError());
  if (passes_out) *passes_out = pass + 1;
  return LLD_OK;
}

static int match_alloc_common(LldCtx* c, MatchView& v) {
  UPC(v.cell_count, int, nullptr, (size_t)v.n_pairs * (N_CELLS + 1));
  UPC(v.cell_fill, int, nullptr, (size_t)v.n_pairs * N_CELLS);
  UPC(v.kp_cell, int, nullptr, v.n_cur);
  UPC(v.s_rec, uint4, nullptr, 3 * (size_t)v.n_cur);
  UPC(v.owner, int, nullptr, v.n_cur);
  UPC(v.q_best, int, nullptr, v.n_q);
  UPC(v.q_dist, int, nullptr, v.n_q);
  UPC(v.changed, int, nullptr, 1);
  UPC(v.q_pair, int, nullptr, v.n_q);
  UPC(v.pair_mode, int, nullptr, v.n_pairs);
  UPC(v.q_top, unsigned long long, nullptr, 4 * (size_t)v.n_q);
  UPC(v.q_ncand, int, nullptr, v.n_q);
  UPC(v.match, int, nullptr, v.n_cur);
  UPC(v.n_matches, int, nullptr, v.n_pairs);
  LLD_CUDA(c, cudaMemsetAsync(v.changed, 0, sizeof(int), c->stream));
  return LLD_OK;
}

static void set_geom(MatchView& v, const lld_frame_geom& g) {
  v.fx = g.fx; v.fy = g.fy; v.cx = g.cx; v.cy = g.cy; v.bf = g.bf; v.b = g.b;
  v.min_x = g.min_x; v.max_x = g.max_x; v.min_y = g.min_y; v.max_y = g.max_y;
  v.winv = static_cast<float>(GRID_COLS) / (g.max_x - g.min_x);   // src/Frame.cc:143-144
  v.hinv = static_cast<float>(GRID_ROWS) / (g.max_y - g.min_y);
  v.n_levels = g.n_levels;
}

static int match_download(LldCtx* c, MatchView& v, lld_sbp_result* out) {
  LLD_CUDA(c, cudaEventRecord(c->ev[2], c->stream));
  if (out->match && v.n_cur) LLD_CUDA(c, cudaMemcpyAsync(out->match, v.match, sizeof(int) * (size_t)v.n_cur, cudaMemcpyDeviceToHost, c->stream));
  if (out->n_matches) LLD_CUDA(c, cudaMemcpyAsync(out->n_matches, v.n_matches, sizeof(int) * (size_t)v.n_pairs, cudaMemcpyDeviceToHost, c->stream));
  if (out->best_idx && v.n_q) LLD_CUDA(c, cudaMemcpyAsync(out->best_idx, v.q_best, sizeof(int) * (size_t)v.n_q, cudaMemcpyDeviceToHost, c->stream));
  if (out->best_dist && v.n_q) LLD_CUDA(c, cudaMemcpyAsync(out->best_dist, v.q_dist, sizeof(int) * (size_t)v.n_q, cudaMemcpyDeviceToHost, c->stream));
  LLD_CUDA(c, cudaEventRecord(c->ev[3], c->stream));
  LLD_CUDA(c, cudaStreamSynchronize(c->stream));
  cudaEventElapsedTime(&c->ms_h2d, c->ev[0], c->ev[1]);
  cudaEventElapsedTime(&c->ms_compute, c->ev[1], c->ev[2]);
  cudaEventElapsedTime(&c->ms_d2h, c->ev[2], c->ev[3]);
  return LLD_OK;
}

static MatchView g_resident;  // resident-mode view (bench): uploaded once, run many times
static bool g_resident_valid = false;

static int sbp_frame_upload(LldCtx* c, const lld_sbp_frame_problem* p, MatchView& v) {
  c->pool_reset();
  v = MatchView();
  v.n_pairs = p->n_pairs;
  LLD_ARG(c, p->n_pairs >= 1);
  v.n_cur = p->cur_off[p->n_pairs];
  v.n_q = p->last_off[p->n_pairs];
  LLD_ARG(c, p->geom.n_levels >= 1 && p->geom.n_levels <= 8);  // octave is packed into 3 bits of the candidate key
  for (int i = 0; i < p->n_pairs; i++) LLD_ARG(c, p->cur_off[i + 1] - p->cur_off[i] <= 65535);
  v.variant = 0;
  set_geom(v, p->geom);
  v.th = p->th; v.nn_ratio = 0; v.mono = p->mono; v.check_ori = p->check_orientation;
  UPC(v.scale, float, p->geom.scale_factors, p->geom.n_levels);
  UPC(v.cur_off, int, p->cur_off, p->n_pairs + 1);
  UPC(v.cur_xy, float, p->cur_xy, 2 * (size_t)v.n_cur);
  UPC(v.cur_octave, uint8_t, p->cur_octave, v.n_cur);
  UPC(v.cur_angle, float, p->cur_angle, v.n_cur);
  UPC(v.cur_uright, float, p->cur_uright, v.n_cur);
  UPC(v.cur_desc, uint8_t, p->cur_desc, 32 * (size_t)v.n_cur);
  UPC(v.cur_claimed, uint8_t, p->cur_claimed, v.n_cur);
  UPC(v.cur_Tcw, float, p->cur_Tcw, 12 * (size_t)p->n_pairs);
  UPC(v.last_Tcw, float, p->last_Tcw, 12 * (size_t)p->n_pairs);
  UPC(v.q_off, int, p->last_off, p->n_pairs + 1);
  UPC(v.q_valid, uint8_t, p->last_valid, v.n_q);
  UPC(v.q_xw, float, p->last_xw, 3 * (size_t)v.n_q);
  UPC(v.q_octave, uint8_t, p->last_octave, v.n_q);
  UPC(v.q_angle, float, p->last_angle, v.n_q);
  UPC(v.q_desc, uint8_t, p->last_desc, 32 * (size_t)v.n_q);
  UPC(v.q_has_obs, uint8_t, p->last_has_obs, v.n_q);
  return match_alloc_common(c, v);
}

static int sbp_mp_upload(LldCtx* c, const lld_sbp_mp_problem* p, MatchView& v) {
  c->pool_reset();
  v = MatchView();
  v.n_pairs = p->n_pairs;
  LLD_ARG(c, p->n_pairs >= 1);
  v.n_cur = p->cur_off[p->n_pairs];
  v.n_q = p->mp_off[p->n_pairs];
  LLD_ARG(c, p->geom.n_levels >= 1 && p->geom.n_levels <= 8);
  for (int i = 0; i < p->n_pairs; i++) LLD_ARG(c, p->cur_off[i + 1] - p->cur_off[i] <= 65535);
  v.variant = 1;
  set_geom(v, p->geom);
  v.th = p->th; v.nn_ratio = p->nn_ratio; v.mono = 0; v.check_ori = 0;
  UPC(v.scale, float, p->geom.scale_factors, p->geom.n_levels);
  UPC(v.cur_off, int, p->cur_off, p->n_pairs + 1);
  UPC(v.cur_xy, float, p->cur_xy, 2 * (size_t)v.n_cur);
  UPC(v.cur_octave, uint8_t, p->cur_octave, v.n_cur);
  UPC(v.cur_uright, float, p->cur_uright, v.n_cur);
  UPC(v.cur_desc, uint8_t, p->cur_desc, 32 * (size_t)v.n_cur);
  UPC(v.cur_claimed, uint8_t, p->cur_claimed, v.n_cur);
  v.cur_angle = nullptr;
  UPC(v.q_off, int, p->mp_off, p->n_pairs + 1);
  UPC(v.q_valid, uint8_t, p->mp_valid, v.n_q);
  UPC(v.q_xw, float, p->mp_proj, 3 * (size_t)v.n_q);
  UPC(v.q_level, int, p->mp_level, v.n_q);
  UPC(v.q_viewcos, float, p->mp_viewcos, v.n_q);
  UPC(v.q_desc, uint8_t, p->mp_desc, 32 * (size_t)v.n_q);
  UPC(v.q_has_obs, uint8_t, p->mp_has_obs, v.n_q);
  return match_alloc_common(c, v);
}

extern "C" int lld_sbp_frame(void* ctx, const lld_sbp_frame_problem* p, lld_sbp_result* out) {
  LldCtx* c = lld_ctx_cast(ctx);
  if (!c || !p || !out) return LLD_ERR_ARG;
  LLD_CUDA(c, cudaSetDevice(c->device));
  c->launches = 0;
  MatchView v;
  LLD_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
  int r = sbp_frame_upload(c, p, v);
  if (r) return r;
  LLD_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
  r = match_run(c, v, nullptr);
  if (r) return r;
  return match_download(c, v, out);
}

extern "C" int lld_sbp_mappoints(void* ctx, const lld_sbp_mp_problem* p, lld_sbp_result* out) {
  LldCtx* c = lld_ctx_cast(ctx);
  if (!c || !p || !out) return LLD_ERR_ARG;
  LLD_CUDA(c, cudaSetDevice(c->device));
  c->launches = 0;
  MatchView v;
  LLD_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
  int r = sbp_mp_upload(c, p, v);
  if (r) return r;
  LLD_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
  r = match_run(c, v, nullptr);
  if (r) return r;
  return match_download(c, v, out);
}

// resident mode for the bench: upload once, run repeatedly with inputs in HBM
extern "C" int lld_sbp_frame_upload(void* ctx, const lld_sbp_frame_problem* p) {
  LldCtx* c = lld_ctx_cast(ctx);
  if (!c || !p) return LLD_ERR_ARG;
  LLD_CUDA(c, cudaSetDevice(c->device));
  int r = sbp_frame_upload(c, p, g_resident);
  if (r) return r;
  g_resident_valid = true;
  LLD_CUDA(c, cudaStreamSynchronize(c->stream));
  return LLD_OK;
}
extern "C" int lld_sbp_run(void* ctx, int* passes) {
  LldCtx* c = lld_ctx_cast(ctx);
  if (!c || !g_resident_valid) return LLD_ERR_ARG;
  LLD_CUDA(c, cudaSetDevice(c->device));
  return match_run(c, g_resident, passes);
}
extern "C" int lld_sbp_download(void* ctx, lld_sbp_result* out) {
  LldCtx* c = lld_ctx_cast(ctx);
  if (!c || !g_resident_valid || !out) return LLD_ERR_ARG;
  LLD_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
  LLD_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
  return match_download(c, g_resident, out);
}

// host inline popcount distance (ORBmatcher::DescriptorDistance): API completeness, no device involved
extern "C" int lld_descriptor_distance(const uint8_t a[32], const uint8_t b[32]) {
  int dist = 0;
  for (int i = 0; i < 8; i++) {
    uint32_t pa, pb;
    memcpy(&pa, a + 4 * i, 4);
    memcpy(&pb, b + 4 * i, 4);
    dist += __builtin_popcount(pa ^ pb);
  }
  return dist;
}
