// lld_shim.h — C++ host shim that keeps the reference's class / method names over the C-ABI (include/lldba.h).
//
// The reference's KeyFrame / MapPoint / MapLine / Frame classes need OpenCV and the rest of the SLAM system; the shim
// works on POD mirrors that carry exactly the fields the hot path reads and writes, under the same names
// (include/KeyFrame.h, include/MapPoint.h, include/MapLine.h, include/Frame.h).  A maintainer dropping this into
// LLD-SLAM replaces the bodies of the corresponding members with the flatten -> lld_* -> write-back code below
// (INTEGRATION.md).  Header-only, C++14, no dependency besides include/lldba.h.
//
//   lld::Optimizer::LocalBundleAdjustment   <- src/Optimizer.cc:936-1388   (graph construction :1037-1218, write-back :1334-1386)
//   lld::Optimizer::BundleAdjustment        <- src/Optimizer.cc:321-559
//   lld::Optimizer::PoseOptimization        <- src/Optimizer.cc:653-932
//   lld::ORBmatcher::SearchByProjection     <- src/ORBmatcher.cc:1328-1470
//   lld::TwoFrameLineMatcher::MatchLines    <- src/TwoFrameLineMatcher.cc:26-77
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <utility>
#include <vector>

#include "../../include/lldba.h"

namespace lld {

struct KeyPoint { float x, y; int octave; float angle; };
struct KeyLine { float startPointX, startPointY, endPointX, endPointY; int octave; };

struct KeyFrame {
  unsigned long mnId = 0;
  float Tcw[12];                       // R row-major + t  (GetPose(), CV_32F)
  float fx, fy, cx, cy, mbf;
  std::vector<float> mvInvLevelSigma2;
  std::vector<KeyPoint> mvKeysUn;
  std::vector<float> mvuRight;
  std::vector<KeyLine> mvLinesLeft, mvLinesRight;
  std::vector<int> line_matches;
  bool bad = false;
  bool isBad() const { return bad; }
  // GBA shadow fields (src/Optimizer.cc:509-512)
  float mTcwGBA[12];
  unsigned long mnBAGlobalForKF = 0;
};
struct MapPoint {
  unsigned long mnId = 0;
  float pos[3];                                      // GetWorldPos(), CV_32F
  std::map<KeyFrame*, size_t> observations;          // GetObservations()
  bool bad = false;
  float mPosGBA[3];
  unsigned long mnBAGlobalForKF = 0;
};
struct MapLine {
  unsigned long mnId = 0;
  double X0[3], line_dir[3];                         // GetMinimalPos(): doubles (include/MapLine.h:120)
  std::map<KeyFrame*, size_t> observations;
  bool bad = false;
  int Observations() const { return (int)observations.size(); }
};
struct Frame {
  float mTcw[12];
  float fx, fy, cx, cy, mbf, mb;
  float mnMinX, mnMaxX, mnMinY, mnMaxY;
  std::vector<float> mvScaleFactors, mvInvLevelSigma2;
  int N = 0;
  std::vector<KeyPoint> mvKeys, mvKeysUn;
  std::vector<float> mvuRight;
  std::vector<uint8_t> mDescriptors;                  // N x 32
  std::vector<MapPoint*> mvpMapPoints;
  std::vector<uint8_t> mvpMapPointDesc;               // descriptor of mvpMapPoints[i] (pMP->GetDescriptor()), N x 32
  std::vector<uint8_t> mvpMapPointHasObs;             // pMP->Observations() > 0
  std::vector<bool> mvbOutlier;
  std::vector<KeyLine> mvLinesLeft, mvLinesRight;
  std::vector<int> line_matches;
  std::vector<MapLine*> mvpMapLines;
  std::vector<bool> mvbOutlierLines;
};

// LinePyrFactor = 1.44, GetReprojThrPyramid  (src/LineMatching.cc:27,239-247)
inline double GetReprojThrPyramid(double base, int lev) {
  double t = base;
  for (int i = 0; i < lev; i++) t *= 1.44;
  return t;
}

namespace detail {
inline void widen12(const float* T, std::vector<double>& out) { for (int i = 0; i < 12; i++) out.push_back((double)T[i]); }
}

// The sets LocalBundleAdjustment gathers before building the graph (src/Optimizer.cc:938-1018).
struct LocalWindow {
  KeyFrame* pKF = nullptr;                    // current keyframe (its K and baseline drive every line edge)
  std::vector<KeyFrame*> lLocalKeyFrames;     // optimised (fixed only when mnId==0)
  std::vector<KeyFrame*> lFixedCameras;
  std::vector<MapPoint*> lLocalMapPoints;
  std::vector<MapLine*> lLocalMapLines;       // already filtered by Observations()>=4 (:974)
};
struct LocalBAResult {
  std::vector<std::pair<KeyFrame*, MapPoint*>> vToErase;       // src/Optimizer.cc:1278-1311
  std::vector<std::pair<KeyFrame*, MapLine*>> vToEraseLines;   // :1313-1329
};

class Optimizer {
 public:
  // void static LocalBundleAdjustment(KeyFrame* pKF, bool* pbStopFlag, Map* pMap, double gamma = 1.0)  include/Optimizer.h:49
  static int LocalBundleAdjustment(void* ctx, const LocalWindow& w, bool* pbStopFlag, double gamma, LocalBAResult* res) {
    std::vector<KeyFrame*> kfs(w.lLocalKeyFrames);
    kfs.insert(kfs.end(), w.lFixedCameras.begin(), w.lFixedCameras.end());
    std::map<KeyFrame*, int> kfi;
    std::vector<double> T, intr, lcam, pxyz, lxd, linfo;
    std::vector<uint8_t> fixed, lstereo;
    std::vector<int32_t> poff{0}, pkf, loff{0}, lkf;
    std::vector<float> puvr, pinfo, lleft, lright;
    const float baseline = w.pKF->mbf / w.pKF->fx;   // pKF->mbf / pKF->mK.at<float>(0,0): float division (:1215)
    for (size_t i = 0; i < kfs.size(); i++) {
      KeyFrame* k = kfs[i];
      kfi[k] = (int)i;
      detail::widen12(k->Tcw, T);
      fixed.push_back(i >= w.lLocalKeyFrames.size() || k->mnId == 0);   // :1043,1057
      const double in[5] = {k->fx, k->fy, k->cx, k->cy, k->mbf};
      intr.insert(intr.end(), in, in + 5);
      const double lc[4] = {w.pKF->fx, w.pKF->cx, w.pKF->cy, baseline};  // current KF's K for every line edge (:1211-1215)
      lcam.insert(lcam.end(), lc, lc + 4);
    }
    std::vector<std::pair<KeyFrame*, MapPoint*>> edge_owner;
    for (MapPoint* mp : w.lLocalMapPoints) {
      for (int c = 0; c < 3; c++) pxyz.push_back((double)mp->pos[c]);
      for (auto& ob : mp->observations) {
        KeyFrame* k = ob.first;
        if (k->isBad() || !kfi.count(k)) continue;
        const KeyPoint& kp = k->mvKeysUn[ob.second];
        pkf.push_back(kfi[k]);
        puvr.push_back(kp.x); puvr.push_back(kp.y); puvr.push_back(k->mvuRight[ob.second]);
        pinfo.push_back(k->mvInvLevelSigma2[kp.octave]);
        edge_owner.push_back({k, mp});
      }
      poff.push_back((int32_t)pkf.size());
    }
    std::vector<std::pair<KeyFrame*, MapLine*>> cell_owner;
    for (MapLine* ml : w.lLocalMapLines) {
      for (int c = 0; c < 3; c++) lxd.push_back(ml->X0[c]);
      for (int c = 0; c < 3; c++) lxd.push_back(ml->line_dir[c]);
      std::map<int, std::pair<KeyFrame*, size_t>> by_id;     // proj_map is keyed by mnId (:1189-1209)
      for (auto& ob : ml->observations)
        if (!ob.first->isBad() && kfi.count(ob.first)) by_id[(int)ob.first->mnId] = {ob.first, ob.second};
      for (auto& kv : by_id) {
        KeyFrame* k = kv.second.first;
        const size_t li = kv.second.second;
        const KeyLine& kl = k->mvLinesLeft[li];
        lkf.push_back(kfi[k]);
        const float l4[4] = {kl.startPointX, kl.startPointY, kl.endPointX, kl.endPointY};
        lleft.insert(lleft.end(), l4, l4 + 4);
        double thrL = GetReprojThrPyramid(1.0, kl.octave), thrR = thrL;
        if (k->line_matches[li] >= 0) {
          const KeyLine& kr = k->mvLinesRight[k->line_matches[li]];
          const float r4[4] = {kr.startPointX, kr.startPointY, kr.endPointX, kr.endPointY};
          lright.insert(lright.end(), r4, r4 + 4);
          thrR = GetReprojThrPyramid(1.0, kr.octave);
          lstereo.push_back(1);
        } else {
          const float r4[4] = {-1, -1, -1, -1};
          lright.insert(lright.end(), r4, r4 + 4);
          lstereo.push_back(0);
        }
        linfo.push_back(gamma * gamma / (thrL * thrL));   // src/LineOptimizer.cc:33-36,97-101
        linfo.push_back(gamma * gamma / (thrR * thrR));
        cell_owner.push_back({k, ml});
      }
      loff.push_back((int32_t)lkf.size());
    }
    const int32_t one_kf[2] = {0, (int32_t)kfs.size()}, one_pt[2] = {0, (int32_t)w.lLocalMapPoints.size()},
                  one_ln[2] = {0, (int32_t)w.lLocalMapLines.size()};
    lld_ba_problem p;
    std::memset(&p, 0, sizeof(p));
    p.n_win = 1; p.kf_off = one_kf; p.pt_off = one_pt; p.ln_off = one_ln;
    p.kf_Tcw = T.data(); p.kf_fixed = fixed.data(); p.kf_intr = intr.data(); p.kf_line_cam = lcam.data();
    p.pt_xyz = pxyz.data(); p.pt_obs_off = poff.data(); p.pt_obs_kf = pkf.data(); p.pt_obs_uvr = puvr.data(); p.pt_obs_info = pinfo.data();
    p.ln_x0_dir = lxd.data(); p.ln_obs_off = loff.data(); p.ln_obs_kf = lkf.data(); p.ln_obs_left = lleft.data();
    p.ln_obs_right = lright.data(); p.ln_obs_info = linfo.data(); p.ln_obs_stereo = lstereo.data();
    const float thHuberMono = std::sqrt(5.991f), thHuberStereo = std::sqrt(7.815f);   // const float = sqrt(double) (:1088-1089)
    p.robust_points = 1;
    p.delta_pt_mono = (float)std::sqrt(5.991); p.delta_pt_stereo = (float)std::sqrt(7.815);
    (void)thHuberMono; (void)thHuberStereo;
    p.delta_ln_mono = p.delta_pt_mono * gamma; p.delta_ln_stereo = p.delta_pt_stereo * gamma;
    p.chi2_pt_mono = 5.991; p.chi2_pt_stereo = 7.815; p.ln_endpoints_normalized = 0; p.ln_filter = 4;
    std::vector<double> oT(T.size()), oP(pxyz.size()), oL(lxd.size());
    std::vector<uint8_t> pbad(pkf.size()), lbad(2 * lkf.size()), lrem(w.lLocalMapLines.size());
    lld_ba_result r;
    std::memset(&r, 0, sizeof(r));
    r.kf_Tcw = oT.data(); r.pt_xyz = oP.data(); r.ln_x0_dir = oL.data();
    r.pt_obs_bad = pbad.data(); r.ln_obs_bad = lbad.data(); r.ln_removed = lrem.data();
    volatile uint8_t stop = (pbStopFlag && *pbStopFlag) ? 1 : 0;
    // pbStopFlag is a bool written by another thread; bool and uint8_t share their object representation here
    const volatile uint8_t* sp = pbStopFlag ? reinterpret_cast<const volatile uint8_t*>(pbStopFlag) : &stop;
    if (pbStopFlag && *pbStopFlag) return 0;   // :1220-1222 nothing is written back
    const int rc = lld_ba_local(ctx, &p, 5, 15, sp, &r);
    if (rc) return rc;
    // write-back under the map mutex in the reference (:1334-1386); float narrowing as Converter::toCvMat
    if (res) {
      for (size_t e = 0; e < pbad.size(); e++)
        if (pbad[e]) res->vToErase.push_back(edge_owner[e]);
      for (size_t c = 0; c < lkf.size(); c++) {
        const size_t line = std::upper_bound(loff.begin(), loff.end(), (int32_t)c) - loff.begin() - 1;
        if (lrem[line]) continue;
        for (int s = 0; s < 2; s++)
          if (lbad[2 * c + s]) res->vToEraseLines.push_back(cell_owner[c]);   // one entry per bad edge, as GetLineData
      }
    }
    for (size_t i = 0; i < w.lLocalKeyFrames.size(); i++)
      for (int c = 0; c < 12; c++) kfs[i]->Tcw[c] = (float)oT[12 * i + c];
    for (size_t i = 0; i < w.lLocalMapPoints.size(); i++)
      for (int c = 0; c < 3; c++) w.lLocalMapPoints[i]->pos[c] = (float)oP[3 * i + c];
    for (size_t i = 0; i < w.lLocalMapLines.size(); i++) {
      if (lrem[i]) continue;   // GetLineData returned false: SetMinimalPos is not called
      for (int c = 0; c < 3; c++) { w.lLocalMapLines[i]->X0[c] = oL[6 * i + c]; w.lLocalMapLines[i]->line_dir[c] = oL[6 * i + 3 + c]; }
    }
    return 0;
  }

  // int static PoseOptimization(Frame* pFrame, double gamma = 1.0)   include/Optimizer.h:50
  static int PoseOptimization(void* ctx, Frame* F, double gamma = 1.0) {
    std::vector<float> xw, uvr, info, left, right;
    std::vector<double> x0d, linfo;
    std::vector<uint8_t> lst, lgate;
    std::vector<int> pidx, lidx;
    for (int i = 0; i < F->N; i++) {
      MapPoint* mp = F->mvpMapPoints[i];
      if (!mp) continue;
      F->mvbOutlier[i] = false;
      pidx.push_back(i);
      for (int c = 0; c < 3; c++) xw.push_back(mp->pos[c]);
      uvr.push_back(F->mvKeysUn[i].x); uvr.push_back(F->mvKeysUn[i].y); uvr.push_back(F->mvuRight[i]);
      info.push_back(F->mvInvLevelSigma2[F->mvKeysUn[i].octave]);
    }
    std::vector<bool> vnStereoLines;   // one entry per EDGE, indexed by LINE id in the reference (:894-898)
    for (size_t i = 0; i < F->mvpMapLines.size(); i++) {
      MapLine* ml = F->mvpMapLines[i];
      if (!ml) continue;
      lidx.push_back((int)i);
      for (int c = 0; c < 3; c++) x0d.push_back(ml->X0[c]);
      for (int c = 0; c < 3; c++) x0d.push_back(ml->line_dir[c]);
      const KeyLine& kl = F->mvLinesLeft[i];
      const float l4[4] = {kl.startPointX, kl.startPointY, kl.endPointX, kl.endPointY};
      left.insert(left.end(), l4, l4 + 4);
      const bool st = F->line_matches[i] >= 0;
      double tl = GetReprojThrPyramid(1.0, kl.octave), tr = tl;
      if (st) {
        const KeyLine& kr = F->mvLinesRight[F->line_matches[i]];
        const float r4[4] = {kr.startPointX, kr.startPointY, kr.endPointX, kr.endPointY};
        right.insert(right.end(), r4, r4 + 4);
        tr = GetReprojThrPyramid(1.0, kr.octave);
      } else {
        const float r4[4] = {-1, -1, -1, -1};
        right.insert(right.end(), r4, r4 + 4);
      }
      linfo.push_back(gamma * gamma / (tl * tl)); linfo.push_back(gamma * gamma / (tr * tr));
      lst.push_back(st);
      vnStereoLines.push_back(st);
      if (st) vnStereoLines.push_back(st);
    }
    for (size_t j = 0; j < lidx.size(); j++)
      for (int s = 0; s < 2; s++) {
        const size_t idx = (size_t)lidx[j];   // the reference reads vnStereoLines[idx]; out of range -> own flag (carve-out)
        lgate.push_back(idx < vnStereoLines.size() ? (uint8_t)vnStereoLines[idx] : lst[j]);
      }
    std::vector<double> T;
    detail::widen12(F->mTcw, T);
    const double in[5] = {F->fx, F->fy, F->cx, F->cy, F->mbf};
    const float bl = F->mbf / F->fx;
    const double lc[4] = {F->fx, F->cx, F->cy, bl};
    const int32_t po[2] = {0, (int32_t)pidx.size()}, lo[2] = {0, (int32_t)lidx.size()};
    lld_pose_problem p;
    std::memset(&p, 0, sizeof(p));
    p.n_frames = 1; p.Tcw = T.data(); p.intr = in; p.line_cam = lc;
    p.pt_off = po; p.pt_xw = xw.data(); p.pt_uvr = uvr.data(); p.pt_info = info.data();
    p.ln_off = lo; p.ln_x0_dir = x0d.data(); p.ln_left = left.data(); p.ln_right = right.data(); p.ln_info = linfo.data();
    p.ln_stereo = lst.data(); p.ln_gate_stereo = lgate.data();
    const float dM = (float)std::sqrt(5.991), dS = (float)std::sqrt(7.815);
    float dLS = dS, dLM = dM;
    dLS *= gamma; dLM *= gamma;                      // float *= double (:702-703)
    p.delta_mono = dM; p.delta_stereo = dS; p.delta_ln_mono = dLM; p.delta_ln_stereo = dLS;
    p.chi2_mono = 5.991f; p.chi2_stereo = 7.815f;
    p.gate_ln_mono = (double)(dLM * dLM); p.gate_ln_stereo = (double)(dLS * dLS);
    p.n_rounds = 4; p.its = 10;
    std::vector<double> oT(12);
    std::vector<uint8_t> po_(pidx.size() + 1), lo_(lidx.size() + 1);
    int n_inl = 0;
    lld_pose_result r;
    std::memset(&r, 0, sizeof(r));
    r.Tcw = oT.data(); r.pt_outlier = po_.data(); r.ln_outlier = lo_.data(); r.n_inliers = &n_inl;
    const int rc = lld_pose_opt(ctx, &p, &r);
    if (rc) return rc;
    if (pidx.size() < 3) return 0;
    for (size_t j = 0; j < pidx.size(); j++) F->mvbOutlier[pidx[j]] = po_[j] != 0;
    for (size_t j = 0; j < lidx.size(); j++) F->mvbOutlierLines[lidx[j]] = lo_[j] != 0;
    for (int c = 0; c < 12; c++) F->mTcw[c] = (float)oT[c];
    return n_inl;
  }
};

class ORBmatcher {
 public:
  ORBmatcher(float nnratio = 0.6f, bool checkOri = true) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}
  static int DescriptorDistance(const uint8_t* a, const uint8_t* b) { return lld_descriptor_distance(a, b); }
  // int SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono)  include/ORBmatcher.h:52
  int SearchByProjection(void* ctx, Frame& Cur, const Frame& Last, float th, bool bMono, std::vector<int>* match_out) {
    lld_sbp_frame_problem p;
    std::memset(&p, 0, sizeof(p));
    p.n_pairs = 1;
    p.geom.fx = Cur.fx; p.geom.fy = Cur.fy; p.geom.cx = Cur.cx; p.geom.cy = Cur.cy; p.geom.bf = Cur.mbf; p.geom.b = Cur.mb;
    p.geom.min_x = Cur.mnMinX; p.geom.max_x = Cur.mnMaxX; p.geom.min_y = Cur.mnMinY; p.geom.max_y = Cur.mnMaxY;
    p.geom.n_levels = (int)Cur.mvScaleFactors.size(); p.geom.scale_factors = Cur.mvScaleFactors.data();
    p.th = th; p.mono = bMono; p.check_orientation = mbCheckOrientation;
    std::vector<float> cxy, cang, lxw, lang;
    std::vector<uint8_t> coct, cclaimed, lvalid, loct, lhas, ldesc(32 * (size_t)Last.N, 0);
    for (int i = 0; i < Cur.N; i++) {
      cxy.push_back(Cur.mvKeysUn[i].x); cxy.push_back(Cur.mvKeysUn[i].y);
      coct.push_back((uint8_t)Cur.mvKeysUn[i].octave); cang.push_back(Cur.mvKeysUn[i].angle);
      cclaimed.push_back(Cur.mvpMapPoints[i] && Cur.mvpMapPointHasObs[i]);
    }
    for (int i = 0; i < Last.N; i++) {
      MapPoint* mp = Last.mvpMapPoints[i];
      const bool ok = mp && !Last.mvbOutlier[i];
      lvalid.push_back(ok);
      for (int c = 0; c < 3; c++) lxw.push_back(ok ? mp->pos[c] : 0.f);
      loct.push_back((uint8_t)Last.mvKeys[i].octave); lang.push_back(Last.mvKeysUn[i].angle);
      lhas.push_back(ok && Last.mvpMapPointHasObs[i]);
      if (ok) std::memcpy(&ldesc[32 * (size_t)i], &Last.mvpMapPointDesc[32 * (size_t)i], 32);
    }
    const int32_t co[2] = {0, Cur.N}, lo[2] = {0, Last.N};
    p.cur_off = co; p.cur_xy = cxy.data(); p.cur_octave = coct.data(); p.cur_angle = cang.data(); p.cur_uright = Cur.mvuRight.data();
    p.cur_desc = Cur.mDescriptors.data(); p.cur_claimed = cclaimed.data(); p.cur_Tcw = Cur.mTcw; p.last_Tcw = Last.mTcw;
    p.last_off = lo; p.last_valid = lvalid.data(); p.last_xw = lxw.data(); p.last_octave = loct.data(); p.last_angle = lang.data();
    p.last_desc = ldesc.data(); p.last_has_obs = lhas.data();
    std::vector<int32_t> match(Cur.N, -1);
    int32_t nm = 0;
    lld_sbp_result r;
    std::memset(&r, 0, sizeof(r));
    r.match = match.data(); r.n_matches = &nm;
    const int rc = lld_sbp_frame(ctx, &p, &r);
    if (rc) return rc;
    for (int i = 0; i < Cur.N; i++)
      if (match[i] >= 0) Cur.mvpMapPoints[i] = Last.mvpMapPoints[match[i]];
    if (match_out) match_out->assign(match.begin(), match.end());
    return nm;
  }
 private:
  float mfNNratio;
  bool mbCheckOrientation;
};

class TwoFrameLineMatcher {
 public:
  // TwoFrameLineMatcher(const Eigen::Matrix3d& K, double b, double tau, int minLineLength, LineMatcher*)  include/TwoFrameLineMatcher.h:31-37
  TwoFrameLineMatcher(const double K_[9], double b, double tau_, int minLineLength_) : b_(b), tau(tau_), minLineLength(minLineLength_) {
    for (int i = 0; i < 9; i++) K[i] = K_[i];
  }
  // void MatchLines(lines, other_lines, descsLeft, descsRight, std::vector<int>* desc_matches)  :39-40
  int MatchLines(void* ctx, const std::vector<KeyLine>& L, const std::vector<KeyLine>& R, const float* descL, const float* descR,
                 int desc_dim, std::vector<int>* desc_matches) {
    std::vector<float> ls, rs;
    std::vector<int32_t> lo_, ro_;
    for (auto& k : L) { ls.push_back(k.startPointX); ls.push_back(k.startPointY); ls.push_back(k.endPointX); ls.push_back(k.endPointY); lo_.push_back(k.octave); }
    for (auto& k : R) { rs.push_back(k.startPointX); rs.push_back(k.startPointY); rs.push_back(k.endPointX); rs.push_back(k.endPointY); ro_.push_back(k.octave); }
    const int32_t lof[2] = {0, (int32_t)L.size()}, rof[2] = {0, (int32_t)R.size()};
    lld_line_match_problem p;
    std::memset(&p, 0, sizeof(p));
    p.n_pairs = 1; p.desc_dim = desc_dim; p.left_off = lof; p.right_off = rof;
    p.left_seg = ls.data(); p.left_octave = lo_.data(); p.right_seg = rs.data(); p.right_octave = ro_.data();
    p.left_desc = descL; p.right_desc = descR;
    for (int i = 0; i < 9; i++) p.K[i] = K[i];
    p.baseline = b_; p.tau = tau; p.min_line_length = minLineLength;
    std::vector<int32_t> m(L.size() + 1, -1);
    lld_line_match_result r;
    std::memset(&r, 0, sizeof(r));
    r.match = m.data();
    const int rc = lld_line_match(ctx, &p, &r);
    if (rc) return rc;
    desc_matches->assign(m.begin(), m.begin() + L.size());
    return 0;
  }
 private:
  double K[9], b_, tau;
  int minLineLength;
};

}  // namespace lld
