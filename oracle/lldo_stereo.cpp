// lldo_stereo.cpp — ORACLE (test infrastructure only; never linked into or called by the product).
//
// CPU restatement of Frame::ComputeStereoMatches (src/Frame.cc:530-704), SURVEY.md §8(f) row 1: row-banded Hamming
// stereo matching of ORB keypoints + 11x11 SAD sliding refinement + parabola fit + median distance gate.
// It checks lld_slam_b200/csrc/stereo.cu (bit-exact mvuRight / mvDepth) and is itself pinned: tests/test_cpu_oracle.py compares
// it with an independent numpy / cv2 transcription (cv2.norm NORM_L1 for the patch distance, cv2.NORM_HAMMING for the
// descriptor distance).
//
// Reference undefined behaviour that the oracle resolves (each is a `continue`, i.e. "no stereo match for this point"):
//  * vRowIndices[yi] is indexed without a range check (src/Frame.cc:555): rows outside the image are skipped here;
//  * the left patch and the right patches at incR != 0 are taken with rowRange / colRange without a range check
//    (:629, :647; only the incR = +L end is tested, :641-643): a patch that leaves the pyramid image rejects the point;
//  * vDistIdx[size/2] on an empty list (:693): nothing to gate.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <utility>
#include <vector>

namespace {

inline int hamming256(const uint8_t* a, const uint8_t* b) {  // ORBmatcher::DescriptorDistance  src/ORBmatcher.cc:1647-1663
  const uint32_t* pa = reinterpret_cast<const uint32_t*>(a);
  const uint32_t* pb = reinterpret_cast<const uint32_t*>(b);
  int dist = 0;
  for (int i = 0; i < 8; i++, pa++, pb++) {
    unsigned int v = *pa ^ *pb;
    v = v - ((v >> 1) & 0x55555555);
    v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
    dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
  }
  return dist;
}

}  // namespace

extern "C" {

// kp*: [n][2] (x, y) pixel coordinates at level 0; oct*: pyramid level; desc*: [n][32] bytes.
// pyr_L / pyr_R: n_levels pointers to 8-bit images (rows x cols, row stride in bytes) = ORBextractor::mvImagePyramid.
// Outputs mvuRight / mvDepth [N] (-1 = no match); returns the number of matched points.
int lldo_stereo_matches(int N, const float* kpL, const int32_t* octL, const uint8_t* descL,
                        int Nr, const float* kpR, const int32_t* octR, const uint8_t* descR,
                        int n_levels, const float* scale_factors, const float* inv_scale_factors,
                        const uint8_t* const* pyr_L, const uint8_t* const* pyr_R, const int32_t* pyr_rows,
                        const int32_t* pyr_cols, const int32_t* pyr_stride, float mb, float mbf,
                        float* mvuRight, float* mvDepth) {
  const int TH_HIGH = 100, TH_LOW = 50;
  for (int i = 0; i < N; i++) { mvuRight[i] = -1.0f; mvDepth[i] = -1.0f; }
  const int thOrbDist = (TH_HIGH + TH_LOW) / 2;
  const int nRows = pyr_rows[0];
  // keypoints of the right image by row band  (:539-556)
  std::vector<std::vector<size_t>> vRowIndices((size_t)nRows);
  for (int iR = 0; iR < Nr; iR++) {
    const float kpY = kpR[2 * iR + 1];
    const float r = 2.0f * scale_factors[octR[iR]];
    const int maxr = (int)std::ceil(kpY + r);
    const int minr = (int)std::floor(kpY - r);
    for (int yi = minr; yi <= maxr; yi++)
      if (yi >= 0 && yi < nRows) vRowIndices[(size_t)yi].push_back((size_t)iR);
  }
  const float minZ = mb;
  const float minD = 0;
  const float maxD = mbf / minZ;
  std::vector<std::pair<int, int>> vDistIdx;
  vDistIdx.reserve((size_t)N);
  for (int iL = 0; iL < N; iL++) {
    const int levelL = octL[iL];
    const float vL = kpL[2 * iL + 1];
    const float uL = kpL[2 * iL];
    if (!(vL >= 0.0f) || (int)vL >= nRows) continue;
    const std::vector<size_t>& vCandidates = vRowIndices[(size_t)vL];
    if (vCandidates.empty()) continue;
    const float minU = uL - maxD;
    const float maxU = uL - minD;
    if (maxU < 0) continue;
    int bestDist = TH_HIGH;
    size_t bestIdxR = 0;
    const uint8_t* dL = descL + 32 * (size_t)iL;
    for (size_t iC = 0; iC < vCandidates.size(); iC++) {
      const size_t iR = vCandidates[iC];
      if (octR[iR] < levelL - 1 || octR[iR] > levelL + 1) continue;
      const float uR = kpR[2 * iR];
      if (uR >= minU && uR <= maxU) {
        const int dist = hamming256(dL, descR + 32 * iR);
        if (dist < bestDist) {
          bestDist = dist;
          bestIdxR = iR;
        }
      }
    }
    if (bestDist < thOrbDist) {  // subpixel match by correlation  (:615-688)
      const float uR0 = kpR[2 * bestIdxR];
      const float scaleFactor = inv_scale_factors[levelL];
      const float scaleduL = std::round(uL * scaleFactor);
      const float scaledvL = std::round(vL * scaleFactor);
      const float scaleduR0 = std::round(uR0 * scaleFactor);
      const int w = 5;
      const int rows = pyr_rows[levelL], cols = pyr_cols[levelL], stride = pyr_stride[levelL];
      const uint8_t* imL = pyr_L[levelL];
      const uint8_t* imR = pyr_R[levelL];
      const int r0 = (int)(scaledvL - w), c0L = (int)(scaleduL - w);
      if (r0 < 0 || r0 + 2 * w + 1 > rows || c0L < 0 || c0L + 2 * w + 1 > cols) continue;   // carve-out (see header)
      float IL[11][11];
      {
        const float centre = (float)imL[(size_t)(r0 + w) * stride + c0L + w];
        for (int y = 0; y < 11; y++)
          for (int x = 0; x < 11; x++) IL[y][x] = (float)imL[(size_t)(r0 + y) * stride + c0L + x] - centre;
      }
      int bestDistS = INT_MAX;
      int bestincR = 0;
      const int L = 5;
      float vDists[2 * 5 + 1];
      const float iniu = scaleduR0 + L - w;
      const float endu = scaleduR0 + L + w + 1;
      if (iniu < 0 || endu >= (float)cols) continue;
      bool inside = true;
      for (int incR = -L; incR <= +L; incR++) {
        const int c0R = (int)(scaleduR0 + incR - w);
        if (c0R < 0 || c0R + 2 * w + 1 > cols) { inside = false; break; }   // carve-out (see header)
        const float centre = (float)imR[(size_t)(r0 + w) * stride + c0R + w];
        double acc = 0.0;   // cv::norm(IL, IR, NORM_L1) on CV_32F accumulates in double
        for (int y = 0; y < 11; y++)
          for (int x = 0; x < 11; x++) {
            const float ir = (float)imR[(size_t)(r0 + y) * stride + c0R + x] - centre;
            acc += (double)std::fabs(IL[y][x] - ir);
          }
        const float dist = (float)acc;
        if (dist < (float)bestDistS) {   // int promoted to float in the comparison, float truncated on assignment  (:653-657)
          bestDistS = (int)dist;
          bestincR = incR;
        }
        vDists[L + incR] = dist;
      }
      if (!inside) continue;
      if (bestincR == -L || bestincR == L) continue;
      const float dist1 = vDists[L + bestincR - 1];
      const float dist2 = vDists[L + bestincR];
      const float dist3 = vDists[L + bestincR + 1];
      const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
      if (deltaR < -1 || deltaR > 1) continue;
      float bestuR = scale_factors[levelL] * ((float)scaleduR0 + (float)bestincR + deltaR);
      float disparity = (uL - bestuR);
      if (disparity >= minD && disparity < maxD) {
        if (disparity <= 0) {
          disparity = 0.01;
          bestuR = uL - 0.01;
        }
        mvDepth[iL] = mbf / disparity;
        mvuRight[iL] = bestuR;
        vDistIdx.push_back(std::pair<int, int>(bestDistS, iL));
      }
    }
  }
  if (vDistIdx.empty()) return 0;
  std::sort(vDistIdx.begin(), vDistIdx.end());
  const float median = vDistIdx[vDistIdx.size() / 2].first;
  const float thDist = 1.5f * 1.4f * median;
  int n_matched = (int)vDistIdx.size();
  for (int i = (int)vDistIdx.size() - 1; i >= 0; i--) {
    if (vDistIdx[i].first < thDist) break;
    mvuRight[vDistIdx[i].second] = -1;
    mvDepth[vDistIdx[i].second] = -1;
    n_matched--;
  }
  return n_matched;
}

}  // extern "C"
