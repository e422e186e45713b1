// Test-only driver of the C++ shim (lld_slam_b200/host/lld_shim.h): builds POD KeyFrame / MapPoint / MapLine / Frame graphs
// from a flattened synthetic problem (written by tests/shim_io.py), runs the shim's reference-named entry points and dumps
// what they wrote back into the objects.  It also compares the shim's own flattening with the input arrays element by
// element, which catches ordering bugs (std::map<KeyFrame*> iteration, proj_map by mnId, local-then-fixed keyframes).
//
//   shim_run <mode> <in.bin> <out.bin>      mode = local | global | pose | fuse | tri | bow
// Built twice by the tests: against liblldba.so (GPU) and, with -DLLD_SHIM_ORACLE, against oracle/liblld_oracle.so (the
// same C-ABI under the lldo_ prefix) so that the host logic is covered on a CPU-only box.
#ifdef LLD_SHIM_ORACLE
#define lld_ba_local lldo_ba_local
#define lld_ba_global lldo_ba_global
#define lld_pose_opt lldo_pose_opt
#define lld_sbp_frame lldo_sbp_frame
#define lld_sbp_mappoints lldo_sbp_mappoints
#define lld_kf_search lldo_kf_search
#define lld_tri_search lldo_tri_search
#define lld_bow_search lldo_bow_search
#define lld_line_match lldo_line_match
#define lld_descriptor_distance lldo_descriptor_distance
#endif
#include <cstdio>
#include <cstdlib>
#include <string>

#include "../../lld_slam_b200/host/lld_shim.h"

#ifdef LLD_SHIM_ORACLE
extern "C" {
int lldo_ba_local(void*, const lld_ba_problem*, int, int, const volatile uint8_t*, lld_ba_result*);
int lldo_ba_global(void*, const lld_ba_problem*, int, const volatile uint8_t*, lld_ba_result*);
int lldo_pose_opt(void*, const lld_pose_problem*, lld_pose_result*);
}
#endif

struct Arr { char type; std::vector<uint8_t> bytes; size_t n() const { return bytes.size() / (type == 'd' ? 8 : type == 'b' ? 1 : 4); } };
static std::map<std::string, Arr> g_in;
static std::map<std::string, Arr> g_out;

static bool read_all(const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) return false;
  for (;;) {
    uint32_t nl;
    if (fread(&nl, 4, 1, f) != 1) break;
    std::string name(nl, ' ');
    if (fread(&name[0], 1, nl, f) != nl) return false;
    Arr a;
    uint64_t nb;
    if (fread(&a.type, 1, 1, f) != 1 || fread(&nb, 8, 1, f) != 1) return false;
    a.bytes.resize(nb);
    if (nb && fread(a.bytes.data(), 1, nb, f) != nb) return false;
    g_in[name] = a;
  }
  fclose(f);
  return true;
}
static bool write_all(const char* path) {
  FILE* f = fopen(path, "wb");
  if (!f) return false;
  for (auto& kv : g_out) {
    const uint32_t nl = (uint32_t)kv.first.size();
    const uint64_t nb = kv.second.bytes.size();
    fwrite(&nl, 4, 1, f); fwrite(kv.first.data(), 1, nl, f); fwrite(&kv.second.type, 1, 1, f); fwrite(&nb, 8, 1, f);
    if (nb) fwrite(kv.second.bytes.data(), 1, nb, f);
  }
  fclose(f);
  return true;
}
template <typename T> static const T* in(const char* k) { return reinterpret_cast<const T*>(g_in.at(k).bytes.data()); }
static size_t in_n(const char* k) { return g_in.at(k).n(); }
template <typename T> static void out(const char* k, char type, const T* p, size_t n) {
  Arr a; a.type = type; a.bytes.assign(reinterpret_cast<const uint8_t*>(p), reinterpret_cast<const uint8_t*>(p) + n * sizeof(T));
  g_out[k] = a;
}
static double scalar(const char* k) { return in<double>(k)[0]; }

template <typename T> static int cmp(const char* what, const std::vector<T>& a, const T* b, size_t n) {
  if (a.size() < n) { fprintf(stderr, "flatten mismatch: %s has %zu entries, expected %zu\n", what, a.size(), n); return 1; }
  for (size_t i = 0; i < n; i++)
    if (!(a[i] == b[i])) { fprintf(stderr, "flatten mismatch: %s[%zu]\n", what, i); return 1; }
  return 0;
}

static int octave_of(const std::vector<float>& table, float info) {
  for (size_t l = 0; l < table.size(); l++)
    if (table[l] == info) return (int)l;
  fprintf(stderr, "no pyramid level with invSigma2 %.9g\n", info);
  exit(3);
}
static int line_octave_of(double info, double gamma) {
  for (int l = 0; l < 8; l++) {
    const double t = lld::GetReprojThrPyramid(1.0, l);
    if (gamma * gamma / (t * t) == info) return l;
  }
  fprintf(stderr, "no line octave with information %.17g\n", info);
  exit(3);
}

// POD graph of one flattened single-window BA problem.  Keyframe i of the problem becomes kf[i] with mnId = id_of[i];
// the objects live in vectors that are never resized afterwards, so pointer order == index order.
struct Graph {
  std::vector<lld::KeyFrame> kf;
  std::vector<lld::MapPoint> mp;
  std::vector<lld::MapLine> ml;
};

static void build_graph(Graph& G, double gamma, bool global_mode) {
  const int nk = (int)in_n("kf_fixed"), np = (int)in_n("pt_xyz") / 3, nl = (int)in_n("ln_x0_dir") / 6;
  const double* T = in<double>("kf_Tcw"); const double* intr = in<double>("kf_intr");
  G.kf.resize(nk); G.mp.resize(np); G.ml.resize(nl);
  for (int i = 0; i < nk; i++) {
    lld::KeyFrame& k = G.kf[i];
    k.mnId = (unsigned long)i;                   // keyframe 0 is the one the reference fixes by id
    for (int c = 0; c < 12; c++) k.Tcw[c] = (float)T[12 * i + c];
    k.fx = (float)intr[5 * i]; k.fy = (float)intr[5 * i + 1]; k.cx = (float)intr[5 * i + 2]; k.cy = (float)intr[5 * i + 3]; k.mbf = (float)intr[5 * i + 4];
    k.mvInvLevelSigma2.resize(8);
    for (int l = 0; l < 8; l++) k.mvInvLevelSigma2[l] = in<float>("inv_level_sigma2")[l];
  }
  const int32_t* poff = in<int32_t>("pt_obs_off"); const int32_t* pkf = in<int32_t>("pt_obs_kf");
  const float* uvr = in<float>("pt_obs_uvr"); const float* pinfo = in<float>("pt_obs_info");
  for (int i = 0; i < np; i++) {
    lld::MapPoint& m = G.mp[i];
    m.mnId = (unsigned long)i;
    for (int c = 0; c < 3; c++) m.pos[c] = (float)in<double>("pt_xyz")[3 * i + c];
    for (int e = poff[i]; e < poff[i + 1]; e++) {
      lld::KeyFrame& k = G.kf[pkf[e]];
      lld::KeyPoint kp;
      kp.x = uvr[3 * e]; kp.y = uvr[3 * e + 1]; kp.angle = 0;
      const int oct = octave_of(k.mvInvLevelSigma2, pinfo[e]);
      kp.octave = global_mode ? oct / 2 : oct;   // the GBA reads mvInvLevelSigma2[octave*2] (src/Optimizer.cc:405)
      m.observations[&k] = k.mvKeysUn.size();
      k.mvKeysUn.push_back(kp);
      k.mvuRight.push_back(uvr[3 * e + 2]);
    }
  }
  const int32_t* loff = in<int32_t>("ln_obs_off"); const int32_t* lkf = in<int32_t>("ln_obs_kf");
  const float* left = in<float>("ln_obs_left"); const float* right = in<float>("ln_obs_right"); const double* linfo = in<double>("ln_obs_info");
  for (int i = 0; i < nl; i++) {
    lld::MapLine& m = G.ml[i];
    m.mnId = (unsigned long)i;
    for (int c = 0; c < 3; c++) { m.X0[c] = in<double>("ln_x0_dir")[6 * i + c]; m.line_dir[c] = in<double>("ln_x0_dir")[6 * i + 3 + c]; }
    for (int c = loff[i]; c < loff[i + 1]; c++) {
      lld::KeyFrame& k = G.kf[lkf[c]];
      lld::KeyLine kl;
      kl.startPointX = left[4 * c]; kl.startPointY = left[4 * c + 1]; kl.endPointX = left[4 * c + 2]; kl.endPointY = left[4 * c + 3];
      kl.octave = global_mode ? 0 : line_octave_of(linfo[2 * c], gamma);
      m.observations[&k] = k.mvLinesLeft.size();
      k.mvLinesLeft.push_back(kl);
      if (right[4 * c] >= 0) {
        lld::KeyLine kr;
        kr.startPointX = right[4 * c]; kr.startPointY = right[4 * c + 1]; kr.endPointX = right[4 * c + 2]; kr.endPointY = right[4 * c + 3];
        kr.octave = global_mode ? 0 : line_octave_of(linfo[2 * c + 1], gamma);
        k.line_matches.push_back((int)k.mvLinesRight.size());
        k.mvLinesRight.push_back(kr);
      } else {
        k.line_matches.push_back(-1);
      }
    }
  }
}

static int check_flat(const lld::FlatBA& F) {
  int bad = 0;
  bad += cmp("kf_Tcw", F.T, in<double>("kf_Tcw"), in_n("kf_Tcw"));
  bad += cmp("kf_fixed", F.fixed, in<uint8_t>("kf_fixed"), in_n("kf_fixed"));
  bad += cmp("kf_intr", F.intr, in<double>("kf_intr"), in_n("kf_intr"));
  bad += cmp("kf_line_cam", F.lcam, in<double>("kf_line_cam"), in_n("kf_line_cam"));
  bad += cmp("pt_xyz", F.pxyz, in<double>("pt_xyz"), in_n("pt_xyz"));
  bad += cmp("pt_obs_off", F.poff, in<int32_t>("pt_obs_off"), in_n("pt_obs_off"));
  bad += cmp("pt_obs_kf", F.pkf, in<int32_t>("pt_obs_kf"), in_n("pt_obs_kf"));
  bad += cmp("pt_obs_uvr", F.puvr, in<float>("pt_obs_uvr"), in_n("pt_obs_uvr"));
  bad += cmp("pt_obs_info", F.pinfo, in<float>("pt_obs_info"), in_n("pt_obs_info"));
  bad += cmp("ln_x0_dir", F.lxd, in<double>("ln_x0_dir"), in_n("ln_x0_dir"));
  bad += cmp("ln_obs_off", F.loff, in<int32_t>("ln_obs_off"), in_n("ln_obs_off"));
  bad += cmp("ln_obs_kf", F.lkf, in<int32_t>("ln_obs_kf"), in_n("ln_obs_kf"));
  bad += cmp("ln_obs_left", F.lleft, in<float>("ln_obs_left"), in_n("ln_obs_left"));
  bad += cmp("ln_obs_right", F.lright, in<float>("ln_obs_right"), in_n("ln_obs_right"));
  bad += cmp("ln_obs_info", F.linfo, in<double>("ln_obs_info"), in_n("ln_obs_info"));
  bad += cmp("ln_obs_stereo", F.lstereo, in<uint8_t>("ln_obs_stereo"), in_n("ln_obs_stereo"));
  const lld_ba_problem& p = F.p;
  if (p.delta_pt_mono != scalar("delta_pt_mono") || p.delta_pt_stereo != scalar("delta_pt_stereo") || p.delta_ln_mono != scalar("delta_ln_mono") ||
      p.delta_ln_stereo != scalar("delta_ln_stereo") || p.robust_points != (int)scalar("robust_points") ||
      p.ln_endpoints_normalized != (int)scalar("ln_endpoints_normalized")) {
    fprintf(stderr, "flatten mismatch: entry-point constants\n");
    bad++;
  }
  return bad;
}

static void dump_state(const Graph& G) {
  std::vector<float> T, P, TG, PG;
  std::vector<double> L;
  for (auto& k : G.kf) { T.insert(T.end(), k.Tcw, k.Tcw + 12); TG.insert(TG.end(), k.mTcwGBA, k.mTcwGBA + 12); }
  for (auto& m : G.mp) { P.insert(P.end(), m.pos, m.pos + 3); PG.insert(PG.end(), m.mPosGBA, m.mPosGBA + 3); }
  for (auto& m : G.ml) { L.insert(L.end(), m.X0, m.X0 + 3); L.insert(L.end(), m.line_dir, m.line_dir + 3); }
  out("kf_Tcw", 'f', T.data(), T.size()); out("pt_xyz", 'f', P.data(), P.size()); out("ln_x0_dir", 'd', L.data(), L.size());
  out("kf_TcwGBA", 'f', TG.data(), TG.size()); out("pt_xyzGBA", 'f', PG.data(), PG.size());
}

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  const std::string mode = argv[1];
  if (!read_all(argv[2])) { fprintf(stderr, "cannot read %s\n", argv[2]); return 2; }
  void* ctx = nullptr;
#ifndef LLD_SHIM_ORACLE
  if (lld_ctx_create(0, &ctx) != LLD_OK) { fprintf(stderr, "no CUDA device\n"); return 4; }
#endif
  int rc = 0;
  if (mode == "local" || mode == "global") {
    const double gamma = scalar("gamma");
    Graph G;
    build_graph(G, gamma, mode == "global");
    const uint8_t* fixed = in<uint8_t>("kf_fixed");
    if (mode == "local") {
      lld::LocalWindow w;
      w.pKF = &G.kf[0];
      for (size_t i = 0; i < G.kf.size(); i++) (fixed[i] && i > 0 ? w.lFixedCameras : w.lLocalKeyFrames).push_back(&G.kf[i]);
      for (auto& m : G.mp) w.lLocalMapPoints.push_back(&m);
      for (auto& m : G.ml) w.lLocalMapLines.push_back(&m);
      {
        lld::FlatBA F;
        lld::Optimizer::FlattenLocal(w, gamma, &F);
        if (check_flat(F)) return 5;
      }
      lld::LocalBAResult res;
      bool stop = false;
      rc = lld::Optimizer::LocalBundleAdjustment(ctx, w, &stop, gamma, &res);
      std::vector<int32_t> er, el;
      for (auto& x : res.vToErase) { er.push_back((int32_t)x.first->mnId); er.push_back((int32_t)x.second->mnId); }
      for (auto& x : res.vToEraseLines) { el.push_back((int32_t)x.first->mnId); el.push_back((int32_t)x.second->mnId); }
      out("vToErase", 'i', er.data(), er.size()); out("vToEraseLines", 'i', el.data(), el.size());
    } else {
      std::vector<lld::KeyFrame*> kfs; std::vector<lld::MapPoint*> mps; std::vector<lld::MapLine*> mls;
      for (auto& k : G.kf) kfs.push_back(&k);
      for (auto& m : G.mp) mps.push_back(&m);
      for (auto& m : G.ml) mls.push_back(&m);
      const bool robust = scalar("robust_points") != 0;
      {
        lld::FlatBA F;
        std::vector<bool> a, b;
        lld::Optimizer::FlattenGlobal(kfs, mps, mls, robust, &F, &a, &b);
        if (check_flat(F)) return 5;
      }
      lld::Map map;
      map.keyframes = kfs; map.points = mps; map.lines = mls;
      rc = lld::Optimizer::GlobalBundleAdjustemnt(ctx, &map, (int)scalar("n_iter"), nullptr, (unsigned long)scalar("nLoopKF"), robust);
    }
    dump_state(G);
  } else if (mode == "pose") {
    const int F = (int)in_n("n_inliers_slot");
    const int32_t* poff = in<int32_t>("pt_off"); const int32_t* loff = in<int32_t>("ln_off");
    const double gamma = scalar("gamma");
    std::vector<float> Tout;
    std::vector<int32_t> ninl;
    std::vector<uint8_t> pout, lout;
    for (int f = 0; f < F; f++) {
      lld::Frame fr;
      for (int c = 0; c < 12; c++) fr.mTcw[c] = (float)in<double>("Tcw")[12 * f + c];
      const double* it = in<double>("intr") + 5 * f;
      fr.fx = (float)it[0]; fr.fy = (float)it[1]; fr.cx = (float)it[2]; fr.cy = (float)it[3]; fr.mbf = (float)it[4];
      fr.mvInvLevelSigma2.resize(8);
      for (int l = 0; l < 8; l++) fr.mvInvLevelSigma2[l] = in<float>("inv_level_sigma2")[l];
      const int n = poff[f + 1] - poff[f], m = loff[f + 1] - loff[f];
      std::vector<lld::MapPoint> mps(n);
      std::vector<lld::MapLine> mls(m);
      fr.N = n;
      for (int i = 0; i < n; i++) {
        const int e = poff[f] + i;
        for (int c = 0; c < 3; c++) mps[i].pos[c] = in<float>("pt_xw")[3 * e + c];
        lld::KeyPoint kp;
        kp.x = in<float>("pt_uvr")[3 * e]; kp.y = in<float>("pt_uvr")[3 * e + 1]; kp.angle = 0;
        kp.octave = octave_of(fr.mvInvLevelSigma2, in<float>("pt_info")[e]);
        fr.mvKeysUn.push_back(kp); fr.mvuRight.push_back(in<float>("pt_uvr")[3 * e + 2]);
        fr.mvpMapPoints.push_back(&mps[i]); fr.mvbOutlier.push_back(false);
      }
      for (int i = 0; i < m; i++) {
        const int c = loff[f] + i;
        for (int q = 0; q < 3; q++) { mls[i].X0[q] = in<double>("ln_x0_dir")[6 * c + q]; mls[i].line_dir[q] = in<double>("ln_x0_dir")[6 * c + 3 + q]; }
        lld::KeyLine kl;
        const float* l4 = in<float>("ln_left") + 4 * c; const float* r4 = in<float>("ln_right") + 4 * c;
        kl.startPointX = l4[0]; kl.startPointY = l4[1]; kl.endPointX = l4[2]; kl.endPointY = l4[3];
        kl.octave = line_octave_of(in<double>("ln_info")[2 * c], gamma);
        fr.mvLinesLeft.push_back(kl);
        if (r4[0] >= 0) {
          lld::KeyLine kr;
          kr.startPointX = r4[0]; kr.startPointY = r4[1]; kr.endPointX = r4[2]; kr.endPointY = r4[3];
          kr.octave = line_octave_of(in<double>("ln_info")[2 * c + 1], gamma);
          fr.line_matches.push_back((int)fr.mvLinesRight.size());
          fr.mvLinesRight.push_back(kr);
        } else {
          fr.line_matches.push_back(-1);
        }
        fr.mvpMapLines.push_back(&mls[i]); fr.mvbOutlierLines.push_back(false);
      }
      const int r = lld::Optimizer::PoseOptimization(ctx, &fr, gamma);
      if (r < 0) { rc = r; break; }
      ninl.push_back(r);
      Tout.insert(Tout.end(), fr.mTcw, fr.mTcw + 12);
      for (int i = 0; i < n; i++) pout.push_back(fr.mvbOutlier[i]);
      for (int i = 0; i < m; i++) lout.push_back(fr.mvbOutlierLines[i]);
    }
    out("Tcw", 'f', Tout.data(), Tout.size()); out("n_inliers", 'i', ninl.data(), ninl.size());
    out("pt_outlier", 'b', pout.data(), pout.size()); out("ln_outlier", 'b', lout.data(), lout.size());
  } else if (mode == "fuse") {
    // one keyframe, its keypoints (some already carrying a map point), a list of candidate map points: ORBmatcher::Fuse
    lld::KeyFrame kf;
    for (int c = 0; c < 12; c++) kf.Tcw[c] = in<float>("Tcw")[c];
    const float* it = in<float>("intr");
    kf.fx = it[0]; kf.fy = it[1]; kf.cx = it[2]; kf.cy = it[3]; kf.mbf = it[4];
    const float* bd = in<float>("bounds");
    kf.mnMinX = bd[0]; kf.mnMaxX = bd[1]; kf.mnMinY = bd[2]; kf.mnMaxY = bd[3];
    const int nl = (int)in_n("scale_factors");
    kf.mvScaleFactors.assign(in<float>("scale_factors"), in<float>("scale_factors") + nl);
    kf.mvInvLevelSigma2.assign(in<float>("inv_level_sigma2"), in<float>("inv_level_sigma2") + nl);
    kf.mfLogScaleFactor = in<float>("log_scale_factor")[0];
    kf.mnScaleLevels = nl;
    const int N = (int)in_n("kp_octave");
    for (int i = 0; i < N; i++) {
      lld::KeyPoint kp;
      kp.x = in<float>("kp_xy")[2 * i]; kp.y = in<float>("kp_xy")[2 * i + 1]; kp.octave = in<int32_t>("kp_octave")[i]; kp.angle = 0;
      kf.mvKeysUn.push_back(kp);
    }
    kf.mvuRight.assign(in<float>("kp_uright"), in<float>("kp_uright") + N);
    kf.mDescriptors.assign(in<uint8_t>("kp_desc"), in<uint8_t>("kp_desc") + 32 * (size_t)N);
    const int M = (int)in_n("mp_nobs");
    std::vector<lld::MapPoint> mps(M), existing(N);
    std::vector<lld::KeyFrame> others(8);             // observers that only make Observations() count
    kf.mvpMapPoints.assign(N, nullptr);
    for (int i = 0; i < N; i++) {
      const int nobs = in<int32_t>("kp_mp_nobs")[i];  // < 0: the keypoint has no map point
      if (nobs < 0) continue;
      existing[i].mnId = 100000 + i;
      for (int o = 0; o < nobs && o < 8; o++) existing[i].observations[&others[o]] = 0;
      kf.mvpMapPoints[i] = &existing[i];
    }
    std::vector<lld::MapPoint*> vp;
    for (int i = 0; i < M; i++) {
      lld::MapPoint& m = mps[i];
      m.mnId = i;
      for (int c = 0; c < 3; c++) { m.pos[c] = in<float>("mp_pos")[3 * i + c]; m.mNormalVector[c] = in<float>("mp_normal")[3 * i + c]; }
      m.mfMinDistance = in<float>("mp_minmax")[2 * i]; m.mfMaxDistance = in<float>("mp_minmax")[2 * i + 1];
      std::memcpy(m.mDescriptor, in<uint8_t>("mp_desc") + 32 * (size_t)i, 32);
      m.bad = in<uint8_t>("mp_bad")[i] != 0;
      for (int o = 0; o < in<int32_t>("mp_nobs")[i] && o < 8; o++) m.observations[&others[o]] = 0;
      if (in<uint8_t>("mp_in_kf")[i]) m.observations[&kf] = 0;
      vp.push_back(in<uint8_t>("mp_null")[i] ? nullptr : &m);
    }
    lld::ORBmatcher matcher(0.6f, true);
    std::vector<lld::ORBmatcher::FuseAction> acts;
    const int nFused = matcher.Fuse(ctx, &kf, vp, in<float>("th")[0], &acts);
    if (nFused < 0) rc = nFused;
    std::vector<int32_t> a_mp, a_idx, a_kind, kf_mp(N, -1);
    for (auto& a : acts) {
      a_mp.push_back((int32_t)a.pMP->mnId); a_idx.push_back(a.bestIdx);
      a_kind.push_back(a.survivor == nullptr ? 0 : (a.survivor == a.pMP ? 1 : 2));   // 0 added, 1 the new point survives, 2 the keyframe's point survives
    }
    for (int i = 0; i < N; i++)
      if (kf.mvpMapPoints[i]) kf_mp[i] = (int32_t)kf.mvpMapPoints[i]->mnId;
    const int32_t nf = nFused;
    out("n_fused", 'i', &nf, 1);
    out("act_mp", 'i', a_mp.data(), a_mp.size()); out("act_idx", 'i', a_idx.data(), a_idx.size()); out("act_kind", 'i', a_kind.data(), a_kind.size());
    out("kf_mp", 'i', kf_mp.data(), kf_mp.size());
  } else if (mode == "tri" || mode == "bow") {
    // two keyframes with feature vectors: ORBmatcher::SearchForTriangulation
    lld::KeyFrame kf[2];
    for (int k = 0; k < 2; k++) {
      const std::string sfx = k ? "2" : "1";
      auto key = [&](const char* base) { return std::string(base) + sfx; };
      lld::KeyFrame& K = kf[k];
      for (int c = 0; c < 12; c++) K.Tcw[c] = in<float>(key("Tcw").c_str())[c];
      const float* it = in<float>("intr");
      K.fx = it[0]; K.fy = it[1]; K.cx = it[2]; K.cy = it[3]; K.mbf = it[4];
      const int nl = (int)in_n("scale_factors");
      K.mvScaleFactors.assign(in<float>("scale_factors"), in<float>("scale_factors") + nl);
      K.mvLevelSigma2.assign(in<float>("level_sigma2"), in<float>("level_sigma2") + nl);
      const int N = (int)in_n(key("kp_angle").c_str());
      for (int i = 0; i < N; i++) {
        lld::KeyPoint kp;
        kp.x = in<float>(key("kp_xy").c_str())[2 * i]; kp.y = in<float>(key("kp_xy").c_str())[2 * i + 1];
        kp.octave = in<uint8_t>(key("kp_octave").c_str())[i]; kp.angle = in<float>(key("kp_angle").c_str())[i];
        K.mvKeysUn.push_back(kp);
      }
      K.mvuRight.assign(in<float>(key("kp_uright").c_str()), in<float>(key("kp_uright").c_str()) + N);
      K.mDescriptors.assign(in<uint8_t>(key("kp_desc").c_str()), in<uint8_t>(key("kp_desc").c_str()) + 32 * (size_t)N);
      static std::vector<lld::MapPoint> pool[2];      // one map point object per keypoint that has one (mnId = keypoint index)
      pool[k].assign((size_t)N, lld::MapPoint());
      K.mvpMapPoints.assign(N, nullptr);
      for (int i = 0; i < N; i++)
        if (in<uint8_t>(key("kp_has_mp").c_str())[i]) { pool[k][(size_t)i].mnId = (unsigned long)i; K.mvpMapPoints[i] = &pool[k][(size_t)i]; }
      const int nn = (int)in_n(key("fv_node").c_str());
      for (int n = 0; n < nn; n++) {
        auto& lst = K.mFeatVec[(unsigned)in<int32_t>(key("fv_node").c_str())[n]];
        for (int e = in<int32_t>(key("fv_idx_off").c_str())[n]; e < in<int32_t>(key("fv_idx_off").c_str())[n + 1]; e++)
          lst.push_back((unsigned)in<int32_t>(key("fv_idx").c_str())[e]);
      }
    }
    if (mode == "bow") {
      lld::ORBmatcher bm(in<float>("nn_ratio")[0], in<uint8_t>("check_orientation")[0] != 0);
      std::vector<lld::MapPoint*> m12;
      const int nb = bm.SearchByBoW(ctx, &kf[0], &kf[1], m12);
      if (nb < 0) rc = nb;
      std::vector<int32_t> flat(m12.size(), -1);
      for (size_t i = 0; i < m12.size(); i++)
        if (m12[i]) flat[i] = (int32_t)m12[i]->mnId;
      const int32_t n32 = nb;
      out("n_matches", 'i', &n32, 1); out("match12", 'i', flat.data(), flat.size());
#ifndef LLD_SHIM_ORACLE
      lld_ctx_destroy(ctx);
#endif
      if (rc) { fprintf(stderr, "shim entry point failed with %d\n", rc); return 6; }
      return write_all(argv[3]) ? 0 : 7;
    }
    lld::ORBmatcher matcher(0.6f, in<uint8_t>("check_orientation")[0] != 0);
    std::vector<std::pair<size_t, size_t>> pairs;
    float ep[2] = {0, 0};
    const int nm = matcher.SearchForTriangulation(ctx, &kf[0], &kf[1], in<float>("F12"), pairs, in<uint8_t>("only_stereo")[0] != 0, ep);
    if (nm < 0) rc = nm;
    std::vector<int32_t> flat;
    for (auto& pr : pairs) { flat.push_back((int32_t)pr.first); flat.push_back((int32_t)pr.second); }
    const int32_t n32 = nm;
    out("n_matches", 'i', &n32, 1); out("pairs", 'i', flat.data(), flat.size()); out("epipole", 'f', ep, 2);
  } else {
    return 2;
  }
#ifndef LLD_SHIM_ORACLE
  lld_ctx_destroy(ctx);
#endif
  if (rc) { fprintf(stderr, "shim entry point failed with %d\n", rc); return 6; }
  return write_all(argv[3]) ? 0 : 7;
}
