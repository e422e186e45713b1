// ORACLE — TEST INFRASTRUCTURE ONLY (see lldo_math.h header).  Parity unpinned by reference tests.
//
// CPU restatement of the descriptor-matching path:
//   ORBmatcher::DescriptorDistance, both Frame-level ORBmatcher::SearchByProjection variants,
//   Frame::GetFeaturesInArea / AssignFeaturesToGrid, ORBmatcher::ComputeThreeMaxima,
//   TwoFrameLineMatcher::MatchLines / CheckLinePair with the vgl / LineMatching helpers it calls.
// Float arithmetic is kept in float exactly where the reference uses float; this file must be compiled
// with -ffp-contract=off (see oracle/Makefile) so that no fused multiply-add changes those roundings.
//
// LineMatcher::MatchLineDescriptors lives in the un-vendored, unpinned LBDMOD library
// (alexandervakhitov/lbdmod; call sites src/TwoFrameLineMatcher.cc:112, src/Tracking.cc:1092,1532).
// PARITY UNPINNED: we define it as the float32-input L2 norm ||a-b||_2 evaluated in double
// (in-tree evidence: cv::norm(a-b) in src/MapLine.cc:175, threshold scale mdThr: 2.0).
#include <climits>
#include <algorithm>
#include <cmath>
#include <limits>
#include <vector>

#include "../include/lldba.h"
#include "lldo_math.h"

namespace lldo {

// src/ORBmatcher.cc:1647-1663
static int descriptor_distance(const uint8_t* a, const uint8_t* b) {
  int dist = 0;
  for (int i = 0; i < 8; i++) {
    uint32_t pa, pb;
    std::memcpy(&pa, a + 4 * i, 4);
    std::memcpy(&pb, b + 4 * i, 4);
    uint32_t v = pa ^ pb;
    v = v - ((v >> 1) & 0x55555555);
    v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
    dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
  }
  return dist;
}

static const int GRID_COLS = 64, GRID_ROWS = 48;  // include/Frame.h:43-44

struct Grid {
  float min_x, min_y, winv, hinv;
  std::vector<std::vector<int>> cell;  // [col*ROWS + row], insertion order = keypoint index order
  // Frame::AssignFeaturesToGrid / PosInGrid  src/Frame.cc:294-309,446-456
  void build(const lld_frame_geom& g, const float* xy, int n) {
    min_x = g.min_x; min_y = g.min_y;
    winv = static_cast<float>(GRID_COLS) / (g.max_x - g.min_x);
    hinv = static_cast<float>(GRID_ROWS) / (g.max_y - g.min_y);
    cell.assign(GRID_COLS * GRID_ROWS, {});
    for (int i = 0; i < n; i++) {
      const int px = (int)std::round((xy[2 * i] - min_x) * winv);
      const int py = (int)std::round((xy[2 * i + 1] - min_y) * hinv);
      if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) continue;
      cell[px * GRID_ROWS + py].push_back(i);
    }
  }
  // Frame::GetFeaturesInArea  src/Frame.cc:391-444
  void features_in_area(float x, float y, float r, int minLevel, int maxLevel, const float* xy,
                        const uint8_t* octave, std::vector<int>& out) const {
    out.clear();
    const int nMinCellX = std::max(0, (int)std::floor((x - min_x - r) * winv));
    if (nMinCellX >= GRID_COLS) return;
    const int nMaxCellX = std::min(GRID_COLS - 1, (int)std::ceil((x - min_x + r) * winv));
    if (nMaxCellX < 0) return;
    const int nMinCellY = std::max(0, (int)std::floor((y - min_y - r) * hinv));
    if (nMinCellY >= GRID_ROWS) return;
    const int nMaxCellY = std::min(GRID_ROWS - 1, (int)std::ceil((y - min_y + r) * hinv));
    if (nMaxCellY < 0) return;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
      for (int iy = nMinCellY; iy <= nMaxCellY; iy++)
        for (int idx : cell[ix * GRID_ROWS + iy]) {
          if (bCheckLevels) {
            if ((int)octave[idx] < minLevel) continue;
            if (maxLevel >= 0 && (int)octave[idx] > maxLevel) continue;
          }
          const float distx = xy[2 * idx] - x;
          const float disty = xy[2 * idx + 1] - y;
          if (std::fabs(distx) < r && std::fabs(disty) < r) out.push_back(idx);
        }
  }
};

// ORBmatcher::ComputeThreeMaxima  src/ORBmatcher.cc:1601-1642
static void three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {
  int max1 = 0, max2 = 0, max3 = 0;
  for (int i = 0; i < L; i++) {
    const int s = (int)histo[i].size();
    if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
    else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
    else if (s > max3) { max3 = s; ind3 = i; }
  }
  if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
  else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

// cv::Mat (CV_32F) expression A*x + t: one gemm with double accumulation, rounded to float once
static inline float gemm_row(const float* Rrow, const float* x, float t) {
  const double s = (double)Rrow[0] * (double)x[0] + (double)Rrow[1] * (double)x[1] + (double)Rrow[2] * (double)x[2];
  return (float)(s + (double)t);
}

}  // namespace lldo

using namespace lldo;

extern "C" {

int lldo_descriptor_distance(const uint8_t* a, const uint8_t* b) { return descriptor_distance(a, b); }

// ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th, bMono)  src/ORBmatcher.cc:1328-1470
int lldo_sbp_frame(void*, const lld_sbp_frame_problem* p, lld_sbp_result* out) {
  const int HISTO_LENGTH = 30, TH_HIGH = 100;
  const lld_frame_geom& g = p->geom;
  std::vector<int> cand;
  for (int pr = 0; pr < p->n_pairs; pr++) {
    const int c0 = p->cur_off[pr], nc = p->cur_off[pr + 1] - c0;
    const int l0 = p->last_off[pr], nl = p->last_off[pr + 1] - l0;
    const float* xy = p->cur_xy + 2 * (size_t)c0;
    const uint8_t* oct = p->cur_octave + c0;
    Grid G;
    G.build(g, xy, nc);
    std::vector<int> match(nc, -1);
    std::vector<uint8_t> claimed(p->cur_claimed + c0, p->cur_claimed + c0 + nc);
    std::vector<int> rotHist[30];
    const float factor = 1.0f / HISTO_LENGTH;
    const float* Tc = p->cur_Tcw + 12 * (size_t)pr;   // R row-major (9), t (3)
    const float* Tl = p->last_Tcw + 12 * (size_t)pr;
    // twc = -Rcw^T * tcw ; tlc = Rlw*twc + tlw   (:1340-1347), float matrices, double-accumulating gemm
    float twc[3], tlc[3];
    for (int i = 0; i < 3; i++) {
      const double s = (double)Tc[0 * 3 + i] * (double)Tc[9] + (double)Tc[1 * 3 + i] * (double)Tc[10] + (double)Tc[2 * 3 + i] * (double)Tc[11];
      twc[i] = (float)(-s);
    }
    for (int i = 0; i < 3; i++) tlc[i] = gemm_row(Tl + 3 * i, twc, Tl[9 + i]);
    const bool bForward = tlc[2] > g.b && !p->mono;
    const bool bBackward = -tlc[2] > g.b && !p->mono;
    int nmatches = 0;
    for (int i = 0; i < nl; i++) {
      if (out->best_idx) out->best_idx[l0 + i] = -1;
      if (out->best_dist) out->best_dist[l0 + i] = 256;
      if (!p->last_valid[l0 + i]) continue;
      const float* Xw = p->last_xw + 3 * (size_t)(l0 + i);
      const float xc = gemm_row(Tc + 0, Xw, Tc[9]);
      const float yc = gemm_row(Tc + 3, Xw, Tc[10]);
      const float zc = gemm_row(Tc + 6, Xw, Tc[11]);
      const float invzc = 1.0 / zc;
      if (invzc < 0 && !p->allow_negative_depth) continue;   // the relocalisation variant (:1497-1502) has no such test
      float u = g.fx * xc * invzc + g.cx;
      float v = g.fy * yc * invzc + g.cy;
      if (u < g.min_x || u > g.max_x) continue;
      if (v < g.min_y || v > g.max_y) continue;
      const int nLastOctave = p->last_octave[l0 + i];
      const float radius = p->th * g.scale_factors[nLastOctave];
      if (bForward) G.features_in_area(u, v, radius, nLastOctave, -1, xy, oct, cand);
      else if (bBackward) G.features_in_area(u, v, radius, 0, nLastOctave, xy, oct, cand);
      else G.features_in_area(u, v, radius, nLastOctave - 1, nLastOctave + 1, xy, oct, cand);
      if (cand.empty()) continue;
      const uint8_t* dMP = p->last_desc + 32 * (size_t)(l0 + i);
      int bestDist = 256, bestIdx2 = -1;
      for (int i2 : cand) {
        if (claimed[i2]) continue;
        if (p->cur_uright[c0 + i2] > 0) {
          const float ur = u - g.bf * invzc;
          const float er = std::fabs(ur - p->cur_uright[c0 + i2]);
          if (er > radius) continue;
        }
        const int dist = descriptor_distance(dMP, p->cur_desc + 32 * (size_t)(c0 + i2));
        if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
      }
      if (bestDist <= (p->th_high > 0 ? p->th_high : TH_HIGH)) {   // ORBdist in the relocalisation variant (:1547)
        match[bestIdx2] = i;
        if (p->last_has_obs[l0 + i]) claimed[bestIdx2] = 1;
        nmatches++;
        if (out->best_idx) out->best_idx[l0 + i] = bestIdx2;
        if (out->best_dist) out->best_dist[l0 + i] = bestDist;
        if (p->check_orientation) {
          float rot = p->last_angle[l0 + i] - p->cur_angle[c0 + bestIdx2];
          if (rot < 0.0) rot += 360.0f;
          int bin = (int)std::round(rot * factor);
          if (bin == HISTO_LENGTH) bin = 0;
          rotHist[bin].push_back(bestIdx2);
        }
      }
    }
    if (p->check_orientation) {
      int ind1 = -1, ind2 = -1, ind3 = -1;
      three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
      for (int i = 0; i < HISTO_LENGTH; i++)
        if (i != ind1 && i != ind2 && i != ind3)
          for (int j : rotHist[i]) { match[j] = -1; nmatches--; }
    }
    for (int i = 0; i < nc; i++) out->match[c0 + i] = match[i];
    out->n_matches[pr] = nmatches;
  }
  return 0;
}

// ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, th)  src/ORBmatcher.cc:45-129
int lldo_sbp_mappoints(void*, const lld_sbp_mp_problem* p, lld_sbp_result* out) {
  const int TH_HIGH = 100;
  const lld_frame_geom& g = p->geom;
  std::vector<int> cand;
  const bool bFactor = p->th != 1.0;
  for (int pr = 0; pr < p->n_pairs; pr++) {
    const int c0 = p->cur_off[pr], nc = p->cur_off[pr + 1] - c0;
    const int m0 = p->mp_off[pr], nm = p->mp_off[pr + 1] - m0;
    const float* xy = p->cur_xy + 2 * (size_t)c0;
    const uint8_t* oct = p->cur_octave + c0;
    Grid G;
    G.build(g, xy, nc);
    std::vector<int> match(nc, -1);
    std::vector<uint8_t> claimed(p->cur_claimed + c0, p->cur_claimed + c0 + nc);
    int nmatches = 0;
    for (int iMP = 0; iMP < nm; iMP++) {
      if (out->best_idx) out->best_idx[m0 + iMP] = -1;
      if (out->best_dist) out->best_dist[m0 + iMP] = 256;
      if (!p->mp_valid[m0 + iMP]) continue;
      const int nPredictedLevel = p->mp_level[m0 + iMP];
      float r = p->mp_viewcos[m0 + iMP] > 0.998 ? 2.5 : 4.0;  // RadiusByViewingCos :131-137
      if (bFactor) r *= p->th;
      const float* pj = p->mp_proj + 3 * (size_t)(m0 + iMP);
      const float rs = r * g.scale_factors[nPredictedLevel];
      G.features_in_area(pj[0], pj[1], rs, nPredictedLevel - 1, nPredictedLevel, xy, oct, cand);
      if (cand.empty()) continue;
      const uint8_t* d0 = p->mp_desc + 32 * (size_t)(m0 + iMP);
      int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
      for (int idx : cand) {
        if (claimed[idx]) continue;
        if (p->cur_uright[c0 + idx] > 0) {
          const float er = std::fabs(pj[2] - p->cur_uright[c0 + idx]);
          if (er > r * g.scale_factors[nPredictedLevel]) continue;
        }
        const int dist = descriptor_distance(d0, p->cur_desc + 32 * (size_t)(c0 + idx));
        if (dist < bestDist) {
          bestDist2 = bestDist; bestDist = dist;
          bestLevel2 = bestLevel; bestLevel = oct[idx];
          bestIdx = idx;
        } else if (dist < bestDist2) {
          bestLevel2 = oct[idx];
          bestDist2 = dist;
        }
      }
      if (bestDist <= TH_HIGH) {
        if (bestLevel == bestLevel2 && bestDist > p->nn_ratio * bestDist2) continue;
        match[bestIdx] = iMP;
        if (p->mp_has_obs[m0 + iMP]) claimed[bestIdx] = 1;
        nmatches++;
        if (out->best_idx) out->best_idx[m0 + iMP] = bestIdx;
        if (out->best_dist) out->best_dist[m0 + iMP] = bestDist;
      }
    }
    for (int i = 0; i < nc; i++) out->match[c0 + i] = match[i];
    out->n_matches[pr] = nmatches;
  }
  return 0;
}

// ORBmatcher::Fuse(KeyFrame*, vpMapPoints, th) src/ORBmatcher.cc:825-975, Fuse(KeyFrame*, Scw, ...) :977-1100 and
// SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th) :290-403: the part from KeyFrame::GetFeaturesInArea to the acceptance
int lldo_kf_search(void*, const lld_kf_search_problem* p, lld_sbp_result* out) {
  const lld_frame_geom& g = p->geom;
  std::vector<int> cand;
  for (int pr = 0; pr < p->n_pairs; pr++) {
    const int c0 = p->kp_off[pr], nc = p->kp_off[pr + 1] - c0;
    const int m0 = p->mp_off[pr], nm = p->mp_off[pr + 1] - m0;
    const float* xy = p->kp_xy + 2 * (size_t)c0;
    const uint8_t* oct = p->kp_octave + c0;
    Grid G;
    G.build(g, xy, nc);
    std::vector<int> match(nc, -1);
    std::vector<uint8_t> matched(p->kp_claimed + c0, p->kp_claimed + c0 + nc);   // vpMatched[idx] != NULL
    int nmatches = 0;
    for (int iMP = 0; iMP < nm; iMP++) {
      if (out->best_idx) out->best_idx[m0 + iMP] = -1;
      if (out->best_dist) out->best_dist[m0 + iMP] = 256;
      if (!p->mp_valid[m0 + iMP]) continue;
      const int nPredictedLevel = p->mp_level[m0 + iMP];
      const float* pj = p->mp_proj + 3 * (size_t)(m0 + iMP);
      const float u = pj[0], v = pj[1], ur = pj[2];
      const float radius = p->th * g.scale_factors[nPredictedLevel];
      G.features_in_area(u, v, radius, -1, -1, xy, oct, cand);   // KeyFrame::GetFeaturesInArea has no level filter
      if (cand.empty()) continue;
      const uint8_t* dMP = p->mp_desc + 32 * (size_t)(m0 + iMP);
      int bestDist = 256, bestIdx = -1;
      for (int idx : cand) {
        if (matched[idx]) continue;
        const int kpLevel = oct[idx];
        if (kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel) continue;
        if (p->chi2_gate) {
          const float kpx = xy[2 * idx], kpy = xy[2 * idx + 1];
          const float ex = u - kpx, ey = v - kpy;
          if (p->kp_uright[c0 + idx] >= 0) {
            const float er = ur - p->kp_uright[c0 + idx];
            const float e2 = ex * ex + ey * ey + er * er;
            if (e2 * p->inv_level_sigma2[kpLevel] > 7.8) continue;
          } else {
            const float e2 = ex * ex + ey * ey;
            if (e2 * p->inv_level_sigma2[kpLevel] > 5.99) continue;
          }
        }
        const int dist = descriptor_distance(dMP, p->kp_desc + 32 * (size_t)(c0 + idx));
        if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
      }
      if (bestDist <= p->th_low) {
        match[bestIdx] = iMP;
        if (p->sequential_claims) matched[bestIdx] = 1;
        nmatches++;
        if (out->best_idx) out->best_idx[m0 + iMP] = bestIdx;
        if (out->best_dist) out->best_dist[m0 + iMP] = bestDist;
      }
    }
    for (int i = 0; i < nc; i++) out->match[c0 + i] = match[i];
    out->n_matches[pr] = nmatches;
  }
  return 0;
}

// ORBmatcher::SearchForTriangulation  src/ORBmatcher.cc:657-823, CheckDistEpipolarLine :140-157
int lldo_tri_search(void*, const lld_tri_search_problem* p, lld_tri_search_result* out) {
  const int TH_LOW = 50, HISTO_LENGTH = 30;
  for (int pr = 0; pr < p->n_pairs; pr++) {
    const int a0 = p->kp1_off[pr], n1 = p->kp1_off[pr + 1] - a0;
    const int b0 = p->kp2_off[pr];
    const float* F12 = p->F12 + 9 * (size_t)pr;
    const float ex = p->epipole[2 * pr], ey = p->epipole[2 * pr + 1];
    int nmatches = 0;
    std::vector<int> vMatches12(n1, -1);
    std::vector<int> rotHist[30];
    const float factor = 1.0f / HISTO_LENGTH;
    int f1 = p->fv1_node_off[pr], f2 = p->fv2_node_off[pr];
    const int f1end = p->fv1_node_off[pr + 1], f2end = p->fv2_node_off[pr + 1];
    while (f1 != f1end && f2 != f2end) {
      if (p->fv1_node[f1] == p->fv2_node[f2]) {
        for (int i1 = p->fv1_idx_off[f1]; i1 < p->fv1_idx_off[f1 + 1]; i1++) {
          const int idx1 = p->fv1_idx[i1];
          if (p->kp1_has_mp[a0 + idx1]) continue;
          const bool bStereo1 = p->kp1_uright[a0 + idx1] >= 0;
          if (p->only_stereo && !bStereo1) continue;
          const float k1x = p->kp1_xy[2 * (size_t)(a0 + idx1)], k1y = p->kp1_xy[2 * (size_t)(a0 + idx1) + 1];
          const uint8_t* d1 = p->kp1_desc + 32 * (size_t)(a0 + idx1);
          int bestDist = TH_LOW, bestIdx2 = -1;
          for (int i2 = p->fv2_idx_off[f2]; i2 < p->fv2_idx_off[f2 + 1]; i2++) {
            const int idx2 = p->fv2_idx[i2];
            if (p->kp2_has_mp[b0 + idx2]) continue;      // (vbMatched2 is never set)
            const bool bStereo2 = p->kp2_uright[b0 + idx2] >= 0;
            if (p->only_stereo && !bStereo2) continue;
            const int dist = descriptor_distance(d1, p->kp2_desc + 32 * (size_t)(b0 + idx2));
            if (dist > TH_LOW || dist > bestDist) continue;
            const float k2x = p->kp2_xy[2 * (size_t)(b0 + idx2)], k2y = p->kp2_xy[2 * (size_t)(b0 + idx2) + 1];
            const int oct2 = p->kp2_octave[b0 + idx2];
            if (!bStereo1 && !bStereo2) {
              const float distex = ex - k2x, distey = ey - k2y;
              if (distex * distex + distey * distey < 100 * p->scale_factors[oct2]) continue;
            }
            // CheckDistEpipolarLine
            const float a = k1x * F12[0] + k1y * F12[3] + F12[6];
            const float b = k1x * F12[1] + k1y * F12[4] + F12[7];
            const float c = k1x * F12[2] + k1y * F12[5] + F12[8];
            const float num = a * k2x + b * k2y + c;
            const float den = a * a + b * b;
            if (den == 0) continue;
            const float dsqr = num * num / den;
            if (dsqr < 3.84 * p->level_sigma2[oct2]) { bestIdx2 = idx2; bestDist = dist; }
          }
          if (bestIdx2 >= 0) {
            vMatches12[idx1] = bestIdx2;
            nmatches++;
            if (p->check_orientation) {
              float rot = p->kp1_angle[a0 + idx1] - p->kp2_angle[b0 + bestIdx2];
              if (rot < 0.0) rot += 360.0f;
              int bin = (int)std::round(rot * factor);
              if (bin == HISTO_LENGTH) bin = 0;
              rotHist[bin].push_back(idx1);
            }
          }
        }
        f1++; f2++;
      } else if (p->fv1_node[f1] < p->fv2_node[f2]) {
        while (f1 != f1end && p->fv1_node[f1] < p->fv2_node[f2]) f1++;   // lower_bound
      } else {
        while (f2 != f2end && p->fv2_node[f2] < p->fv1_node[f1]) f2++;
      }
    }
    if (p->check_orientation) {
      int ind1 = -1, ind2 = -1, ind3 = -1;
      three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
      for (int i = 0; i < HISTO_LENGTH; i++) {
        if (i == ind1 || i == ind2 || i == ind3) continue;
        for (int j : rotHist[i]) { vMatches12[j] = -1; nmatches--; }
      }
    }
    for (int i = 0; i < n1; i++) out->match12[a0 + i] = vMatches12[i];
    out->n_matches[pr] = nmatches;
  }
  return 0;
}

// ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...) src/ORBmatcher.cc:159-288 and SearchByBoW(KeyFrame*, KeyFrame*, ...) :522-655
int lldo_bow_search(void*, const lld_bow_search_problem* p, lld_tri_search_result* out) {
  const int TH_LOW = 50, HISTO_LENGTH = 30;
  for (int pr = 0; pr < p->n_pairs; pr++) {
    const int a0 = p->kp1_off[pr], n1 = p->kp1_off[pr + 1] - a0;
    const int b0 = p->kp2_off[pr], n2 = p->kp2_off[pr + 1] - b0;
    std::vector<int> m12(n1, -1);
    std::vector<uint8_t> matched2(n2, 0);
    std::vector<int> rotHist[30];
    const float factor = 1.0f / HISTO_LENGTH;
    int nmatches = 0;
    int f1 = p->fv1_node_off[pr], f2 = p->fv2_node_off[pr];
    const int f1end = p->fv1_node_off[pr + 1], f2end = p->fv2_node_off[pr + 1];
    while (f1 != f1end && f2 != f2end) {
      if (p->fv1_node[f1] == p->fv2_node[f2]) {
        for (int i1 = p->fv1_idx_off[f1]; i1 < p->fv1_idx_off[f1 + 1]; i1++) {
          const int idx1 = p->fv1_idx[i1];
          if (!p->kp1_valid[a0 + idx1]) continue;
          const uint8_t* d1 = p->kp1_desc + 32 * (size_t)(a0 + idx1);
          int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
          for (int i2 = p->fv2_idx_off[f2]; i2 < p->fv2_idx_off[f2 + 1]; i2++) {
            const int idx2 = p->fv2_idx[i2];
            if (matched2[idx2] || !p->kp2_valid[b0 + idx2]) continue;
            const int dist = descriptor_distance(d1, p->kp2_desc + 32 * (size_t)(b0 + idx2));
            if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = idx2; }
            else if (dist < bestDist2) bestDist2 = dist;
          }
          if (p->strict_th ? bestDist1 < TH_LOW : bestDist1 <= TH_LOW) {
            if (static_cast<float>(bestDist1) < p->nn_ratio * static_cast<float>(bestDist2)) {
              m12[idx1] = bestIdx2;
              matched2[bestIdx2] = 1;
              if (p->check_orientation) {
                float rot = p->kp1_angle[a0 + idx1] - p->kp2_angle[b0 + bestIdx2];
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)std::round(rot * factor);
                if (bin == HISTO_LENGTH) bin = 0;
                rotHist[bin].push_back(idx1);
              }
              nmatches++;
            }
          }
        }
        f1++; f2++;
      } else if (p->fv1_node[f1] < p->fv2_node[f2]) {
        while (f1 != f1end && p->fv1_node[f1] < p->fv2_node[f2]) f1++;
      } else {
        while (f2 != f2end && p->fv2_node[f2] < p->fv1_node[f1]) f2++;
      }
    }
    if (p->check_orientation) {
      int ind1 = -1, ind2 = -1, ind3 = -1;
      three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
      for (int i = 0; i < HISTO_LENGTH; i++) {
        if (i == ind1 || i == ind2 || i == ind3) continue;
        for (int j : rotHist[i]) { m12[j] = -1; nmatches--; }
      }
    }
    for (int i = 0; i < n1; i++) out->match12[a0 + i] = m12[i];
    out->n_matches[pr] = nmatches;
  }
  return 0;
}

// ---- TwoFrameLineMatcher ----------------------------------------------------------------------
static double line_length(const float* s) {  // LineLength src/LineMatching.cc:50-59
  const double dx = (double)s[0] - (double)s[2], dy = (double)s[1] - (double)s[3];
  return std::sqrt(dx * dx + dy * dy);
}
// vgl::NormalizedLineEquation src/vgl.cc:578-585
static void normalized_line_eq(const float* s, const double K[9], double leq[3]) {
  const double Xs[3] = {s[0], s[1], 1.0}, Xe[3] = {s[2], s[3], 1.0};
  double li[3];
  cross3(Xs, Xe, li);
  for (int i = 0; i < 3; i++) leq[i] = K[0 * 3 + i] * li[0] + K[1 * 3 + i] * li[1] + K[2 * 3 + i] * li[2];  // K^T * l
  const double n = std::sqrt(leq[0] * leq[0] + leq[1] * leq[1]);
  for (int i = 0; i < 3; i++) leq[i] /= n;
}
// vgl::TriangulateLine src/vgl.cc:78-108 ; T1 = [I|0], T2 = [I | (b,0,0)] (GetTForRight, src/LineMatching.cc:228-237)
static bool triangulate_line(const double t2[3], const double l1[3], const double l2[3], double X0[3], double dir[3]) {
  const double* n1 = l1;  // R = I
  const double* n2 = l2;
  if (std::fabs(dot3(n1, n2)) / norm3(n1) / norm3(n2) > 0.975) return false;
  cross3(n1, n2, dir);
  const double dn = norm3(dir);
  for (int i = 0; i < 3; i++) dir[i] /= dn;
  double M[9] = {n1[0], n1[1], n1[2], n2[0], n2[1], n2[2], dir[0], dir[1], dir[2]};
  double b[3] = {0.0, dot3(n2, t2), 0.0};  // n1 . t1 = 0 (t1 = 0)
  if (colpiv_qr_solve(3, 3, M, b, X0) < 3) return false;
  return true;
}

// TwoFrameLineMatcher::MatchLines src/TwoFrameLineMatcher.cc:26-77 and CheckLinePair :79-124
int lldo_line_match(void*, const lld_line_match_problem* p, lld_line_match_result* out) {
  const int D = p->desc_dim;
  const double t2[3] = {p->baseline, 0, 0};
  for (int pr = 0; pr < p->n_pairs; pr++) {
    const int a0 = p->left_off[pr], na = p->left_off[pr + 1] - a0;
    const int b0 = p->right_off[pr], nb = p->right_off[pr + 1] - b0;
    std::vector<uint8_t> other_matched(nb, 0);
    std::vector<double> leqL(3 * (size_t)na), leqR(3 * (size_t)nb);
    for (int j = 0; j < na; j++) normalized_line_eq(p->left_seg + 4 * (size_t)(a0 + j), p->K, &leqL[3 * j]);
    for (int j = 0; j < nb; j++) normalized_line_eq(p->right_seg + 4 * (size_t)(b0 + j), p->K, &leqR[3 * j]);
    for (int j = 0; j < na; j++) {
      double min_d = std::numeric_limits<double>::max();
      int min_j = -1;
      const float* s1 = p->left_seg + 4 * (size_t)(a0 + j);
      for (int oi = 0; oi < nb; oi++) {
        if (other_matched[oi]) continue;
        const float* s2 = p->right_seg + 4 * (size_t)(b0 + oi);
        if (p->left_octave[a0 + j] != p->right_octave[b0 + oi]) continue;
        const double len_thr = p->min_line_length;
        if (line_length(s1) < len_thr || line_length(s2) < len_thr) continue;
        double X0[3], dir[3];
        if (!triangulate_line(t2, &leqL[3 * j], &leqR[3 * oi], X0, dir) || norm3(X0) < 0.5) continue;
        // ReprojectKeyLineTo3D with T = I  src/LineMatching.cc:277-291
        double d, pa, pb;
        const double xs[2] = {s1[0], s1[1]}, xe[2] = {s1[2], s1[3]};
        reproject_line_point(X0, dir, xs, p->K, &d, &pa);
        reproject_line_point(X0, dir, xe, p->K, &d, &pb);
        const double p1z = X0[2] + pa * dir[2], p2z = X0[2] + pb * dir[2];
        if (p1z < 0 || p2z < 0) continue;
        // MatchLineDescriptors (LBDMOD): L2 norm, double accumulation
        const float* da = p->left_desc + (size_t)D * (a0 + j);
        const float* db = p->right_desc + (size_t)D * (b0 + oi);
        double ss = 0;
        for (int k = 0; k < D; k++) {
          const double df = (double)da[k] - (double)db[k];
          ss += df * df;
        }
        const double dist = std::sqrt(ss);
        if (dist < min_d && dist < p->tau) { min_d = dist; min_j = oi; }
      }
      if (min_j >= 0) other_matched[min_j] = 1;
      out->match[a0 + j] = min_j;
      if (out->dist) out->dist[a0 + j] = min_j >= 0 ? (float)min_d : std::numeric_limits<float>::infinity();
    }
  }
  return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------------
// Medoid descriptor of a landmark: MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:242-307) and
// MapLine::ComputeDistinctiveDescriptors (src/MapLine.cc:133-201).  SURVEY §8(f) row 4.
// Landmark l owns descriptors [off[l], off[l+1]) (the observations of non-bad keyframes, in std::map order); best[l] is
// the index (landmark-local) of the descriptor with the least median distance to the others, -1 for an empty landmark.
// ------------------------------------------------------------------------------------------------------------------
extern "C" {

int lldo_medoid_orb(void*, int n_lm, const int32_t* off, const uint8_t* desc, int32_t* best) {
  for (int l = 0; l < n_lm; l++) {
    const int N = off[l + 1] - off[l];
    best[l] = -1;
    if (N <= 0) continue;
    const uint8_t* d = desc + 32 * (size_t)off[l];
    std::vector<float> Dm((size_t)N * N, 0.f);   // float Distances[N][N]  (:277)
    for (int i = 0; i < N; i++)
      for (int j = i + 1; j < N; j++) {
        const int dij = lldo_descriptor_distance(d + 32 * (size_t)i, d + 32 * (size_t)j);
        Dm[(size_t)i * N + j] = (float)dij;
        Dm[(size_t)j * N + i] = (float)dij;
      }
    int BestMedian = INT_MAX, BestIdx = 0;
    for (int i = 0; i < N; i++) {
      std::vector<int> v(Dm.begin() + (size_t)i * N, Dm.begin() + (size_t)(i + 1) * N);   // vector<int> from floats (:293)
      std::sort(v.begin(), v.end());
      const int median = v[(size_t)(0.5 * (N - 1))];
      if (median < BestMedian) { BestMedian = median; BestIdx = i; }
    }
    best[l] = BestIdx;
  }
  return 0;
}

int lldo_medoid_float(void*, int n_lm, const int32_t* off, int D, const float* desc, int32_t* best) {
  for (int l = 0; l < n_lm; l++) {
    const int N = off[l + 1] - off[l];
    best[l] = -1;
    if (N <= 0) continue;
    const float* d = desc + (size_t)D * off[l];
    std::vector<float> Dm((size_t)N * N, 0.f);
    for (int i = 0; i < N; i++)
      for (int j = i + 1; j < N; j++) {
        double s = 0;   // cv::norm(a - b): float difference, squares accumulated in double (:175)
        for (int k = 0; k < D; k++) {
          const float df = d[(size_t)i * D + k] - d[(size_t)j * D + k];
          s += (double)df * (double)df;
        }
        const double dij = std::sqrt(s);
        Dm[(size_t)i * N + j] = (float)dij;
        Dm[(size_t)j * N + i] = (float)dij;
      }
    int BestMedian = INT_MAX, BestIdx = 0;
    for (int i = 0; i < N; i++) {
      std::vector<float> v(Dm.begin() + (size_t)i * N, Dm.begin() + (size_t)(i + 1) * N);
      std::sort(v.begin(), v.end());
      const int median = (int)v[(size_t)(0.5 * (N - 1))];   // `int median = vDists[...]`: the float is truncated (:190)
      if (median < BestMedian) { BestMedian = median; BestIdx = i; }
    }
    best[l] = BestIdx;
  }
  return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------------
// Temporal line association: Tracking::AddLinesFrom (src/Tracking.cc:996-1124) with the reprojection gate
// GetReprojErrPixelsL1 -> vgl::LineReprojErrorL1 (src/LineMatching.cc:270-275, src/vgl.cc:548-559), vgl::MapPoint
// (src/vgl.cc:587-590) and GetReprojThrPyramid (src/LineMatching.cc:239-247).  SURVEY §8(f) row 3.
// The candidate lists (sub_inds of SubselectWithGrid over Frame::lines_grid) are an INPUT: the release never populates
// lines_grid, so the caller decides which current lines a map line is compared with.
// ------------------------------------------------------------------------------------------------------------------
namespace {
inline void map_point_c2w(const double* T, const double* X, double* o) {   // T row-major 4x4: R^T (X - c)
  const double d[3] = {X[0] - T[3], X[1] - T[7], X[2] - T[11]};
  for (int i = 0; i < 3; i++) o[i] = T[0 * 4 + i] * d[0] + T[1 * 4 + i] * d[1] + T[2 * 4 + i] * d[2];
}
inline double line_reproj_err_l1(const float* seg, const double* T, const double* X0, const double* dir, const double* K) {
  double Xa[3], Xb[3], P2[3] = {X0[0] + dir[0], X0[1] + dir[1], X0[2] + dir[2]};
  map_point_c2w(T, X0, Xa);
  map_point_c2w(T, P2, Xb);
  double c1[3], c2[3];
  for (int i = 0; i < 3; i++) {
    c1[i] = K[3 * i] * Xa[0] + K[3 * i + 1] * Xa[1] + K[3 * i + 2] * Xa[2];
    c2[i] = K[3 * i] * Xb[0] + K[3 * i + 1] * Xb[1] + K[3 * i + 2] * Xb[2];
  }
  double l[3] = {c1[1] * c2[2] - c1[2] * c2[1], c1[2] * c2[0] - c1[0] * c2[2], c1[0] * c2[1] - c1[1] * c2[0]};
  const double n = std::sqrt(l[0] * l[0] + l[1] * l[1]);
  l[0] /= n; l[1] /= n; l[2] /= n;
  const double e1 = std::fabs((double)seg[0] * l[0] + (double)seg[1] * l[1] + l[2]);
  const double e2 = std::fabs((double)seg[2] * l[0] + (double)seg[3] * l[1] + l[2]);
  return e1 + e2;
}
}  // namespace

extern "C" int lldo_line_associate(void*, const lld_line_assoc_problem* p, lld_line_assoc_result* out) {
  const int D = p->desc_dim;
  for (int f = 0; f < p->n_frames; f++) {
    const int m0 = p->ml_off[f], m1 = p->ml_off[f + 1];
    const int c0 = p->cur_off[f], nc = p->cur_off[f + 1] - c0;
    const int r0 = p->right_off[f];
    const double* Tc = p->T_curr + 16 * (size_t)f;
    const double* Tr = p->T_right + 16 * (size_t)f;
    std::vector<uint8_t> taken(p->cur_taken + c0, p->cur_taken + c0 + nc);
    for (int i = 0; i < nc; i++) out->cur_assoc[c0 + i] = -1;
    int added = 0;
    for (int i = m0; i < m1; i++) {
      if (!p->ml_valid[i]) continue;
      const double* X0 = p->ml_x0_dir + 6 * (size_t)i;
      const double* dir = X0 + 3;
      double X1c[3], X2c[3];
      map_point_c2w(Tc, p->ml_x1x2 + 6 * (size_t)i, X1c);
      map_point_c2w(Tc, p->ml_x1x2 + 6 * (size_t)i + 3, X2c);
      int match_id = -1;
      double md = 1e10;
      for (int q = p->cand_off[i]; q < p->cand_off[i + 1]; q++) {
        const int si = p->cand_idx[q];
        if (taken[si]) continue;
        const int ri = p->cur_line_match[c0 + si];
        if (ri < 0 && !p->monocular) continue;
        if (X1c[2] < 0 || X2c[2] < 0) continue;
        double thr = p->thr_reproj_base;
        for (int l = 0; l < p->cur_octave[c0 + si]; l++) thr *= 1.44;
        const double se = line_reproj_err_l1(p->cur_left + 4 * (size_t)(c0 + si), Tc, X0, dir, p->K);
        double se2 = 0;
        if (!p->monocular) se2 = line_reproj_err_l1(p->cur_right + 4 * (size_t)(r0 + ri), Tr, X0, dir, p->K);
        if (se > thr || se2 > thr) continue;
        const float* da = p->ml_desc + (size_t)D * i;
        const float* db = p->cur_desc + (size_t)D * (c0 + si);
        double ss = 0;
        for (int k = 0; k < D; k++) {
          const double df = (double)da[k] - (double)db[k];
          ss += df * df;
        }
        const double cd = std::sqrt(ss);
        if (cd < md) { md = cd; match_id = si; }
      }
      if (md > p->md_thr) continue;
      if (match_id >= 0 && !taken[match_id]) {
        taken[match_id] = 1;
        out->cur_assoc[c0 + match_id] = i - m0;
        added++;
      }
    }
    if (out->n_added) out->n_added[f] = added;
  }
  return 0;
}
