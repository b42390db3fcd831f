// adapter_vs_reference.cpp — include/dslam_b200_adapter.hpp instantiated with THE REFERENCE'S OWN TYPES (dso::FrameHessian,
// SE3, AffLight, Vec5, Vec3, CalibHessian as the reference's headers name them; the stand-ins of oracle/shim provide them in
// this image) and driven through the same FrontEnd-style call sequence as the reference's own TrackerAndScaler.cpp, which is
// compiled into this very program (oracle/ref_build.py extracts it in place; nothing is copied into the repository).
// TEST INFRASTRUCTURE ONLY: built into oracle/_ref/adapter_vs_reference, run by tests/test_adapter_reference_types.py.
//
// Sequence (src/FrontEnd.cpp): makeImages of the keyframe / new frames / right frame (:605, :680) -> makeK +
// setCoarseTrackingRef (:57-58, :797-798) -> trackNewestCoarse (:204-206) -> optimizeScale (:992) -> scaleCoarseDepthL0
// (:1032) -> trackNewestCoarse of the next frame against the rescaled template; plus the hypothesis loop of
// FrontEnd::trackNewCoarse (:192-247) through the adapter's trackNewCoarse wrapper against the sequential loop on the
// reference object.  The reference accumulates H, b in fp32 (4 SSE lanes x 3 tiers), the GPU in fp64: results agree to the
// reference's own summation noise, which is what the tolerances below are (measured distances are printed).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

#define private public
#define protected public
#include "tracker_extract.inc"
#undef private
#undef protected

namespace dso {
int pyrLevelsUsed = 1;
int wG[PYR_LEVELS], hG[PYR_LEVELS];
float setting_huberTH = 9;
float setting_coarseCutoffTH = 20;
float setting_affineOptModeA = 0;
float setting_affineOptModeB = 0;
bool setting_debugout_runquiet = true;
int setting_gammaWeightsPixelSelect = 1;
float freeDebugParam3 = 1;
#include "makeimages_extract.inc"
}  // namespace dso

#include "dslam_b200_adapter.hpp"

using namespace dso;
typedef dslam_b200::TrackerAndScaler<FrameHessian, SE3, AffLight, Vec5, Vec3> GpuTracker;

namespace {
struct Scene {
  int w, h, levels, npts;
  float K[4];
  double T[16];
  std::vector<float> img_ref, img_new, img_new2, img_right, pid, hdif;
  std::vector<int> pu, pv;
  double init[2][7];
};
bool load(const char *path, Scene &s) {
  FILE *f = fopen(path, "rb");
  if (!f) return false;
  int hdr[4];
  bool ok = fread(hdr, sizeof(int), 4, f) == 4;
  s.w = hdr[0]; s.h = hdr[1]; s.levels = hdr[2]; s.npts = hdr[3];
  ok = ok && fread(s.K, sizeof(float), 4, f) == 4 && fread(s.T, sizeof(double), 16, f) == 16 && fread(s.init, sizeof(double), 14, f) == 14;
  const size_t px = (size_t)s.w * s.h;
  for (std::vector<float> *v : {&s.img_ref, &s.img_new, &s.img_new2, &s.img_right}) {
    v->resize(px);
    ok = ok && fread(v->data(), sizeof(float), px, f) == px;
  }
  s.pu.resize(s.npts); s.pv.resize(s.npts); s.pid.resize(s.npts); s.hdif.resize(s.npts);
  ok = ok && fread(s.pu.data(), sizeof(int), s.npts, f) == (size_t)s.npts && fread(s.pv.data(), sizeof(int), s.npts, f) == (size_t)s.npts &&
       fread(s.pid.data(), sizeof(float), s.npts, f) == (size_t)s.npts && fread(s.hdif.data(), sizeof(float), s.npts, f) == (size_t)s.npts;
  fclose(f);
  return ok;
}
struct Frame {
  FrameHessian fh;
  FrameShell shell;
  Frame() { fh.shell = &shell; fh.ab_exposure = 1.0f; for (int l = 0; l < PYR_LEVELS; l++) { fh.dIp[l] = nullptr; fh.absSquaredGrad[l] = nullptr; } }
  ~Frame() { for (int l = 0; l < PYR_LEVELS; l++) { delete[] fh.dIp[l]; delete[] fh.absSquaredGrad[l]; } }
};
double rel7(const double *a, const double *b) {
  double d = 0, n = 0;
  for (int i = 0; i < 7; i++) { d += (a[i] - b[i]) * (a[i] - b[i]); n += b[i] * b[i]; }
  return std::sqrt(d / n);
}
int fails = 0;
void expect(bool c, const char *what) {
  if (!c) { std::printf("FAIL: %s\n", what); fails++; }
}
}  // namespace

int main(int argc, char **argv) {
  if (argc < 2) { std::printf("usage: adapter_vs_reference scene.bin\n"); return 2; }
  int ndev = 0;
  if (dslam_device_count(&ndev) != DSLAM_OK || ndev < 1) { std::printf("adapter_vs_reference: no CUDA device (%s)\n", dslam_last_error()); return 3; }
  Scene sc;
  if (!load(argv[1], sc)) { std::printf("cannot read %s\n", argv[1]); return 2; }
  const int w = sc.w, h = sc.h, levels = sc.levels;
  pyrLevelsUsed = levels;
  for (int l = 0; l < levels; l++) { wG[l] = w >> l; hG[l] = h >> l; }
  CalibHessian calib{sc.K[0], sc.K[1], sc.K[2], sc.K[3]};
  for (int i = 0; i < 256; i++) calib.B[i] = (float)i;
  const std::vector<double> tfm(sc.T, sc.T + 16);

  // ---------------- the reference's own code ----------------
  Frame r_ref, r_new, r_new2, r_right;
  r_ref.fh.makeImages(sc.img_ref.data(), &calib);
  r_new.fh.makeImages(sc.img_new.data(), &calib);
  r_new2.fh.makeImages(sc.img_new2.data(), &calib);
  r_right.fh.makeImages(sc.img_right.data(), nullptr);
  Mat33f K1m;
  K1m << sc.K[0], 0.0f, sc.K[2], 0.0f, sc.K[1], sc.K[3], 0.0f, 0.0f, 1.0f;
  TrackerAndScaler ref(w, h, tfm, K1m);
  ref.makeK(&calib);
  std::vector<PointHessian> ph((size_t)sc.npts);
  std::vector<PointFrameResidual> res((size_t)sc.npts);
  std::vector<EFPoint> efp((size_t)sc.npts);
  EFResidual efr;
  auto bind_points = [&](FrameHessian *host) {
    host->pointHessians.clear();
    for (int i = 0; i < sc.npts; i++) {
      efp[i].HdiF = sc.hdif[i];
      res[i].efResidual = &efr;
      res[i].target = host;
      res[i].centerProjectedTo = Vec3f((float)sc.pu[i], (float)sc.pv[i], sc.pid[i]);
      ph[i].lastResiduals[0] = std::make_pair(&res[i], ResState::IN);
      ph[i].lastResiduals[1] = std::make_pair((PointFrameResidual *)nullptr, ResState::OOB);
      ph[i].efPoint = &efp[i];
      host->pointHessians.push_back(&ph[i]);
    }
  };
  bind_points(&r_ref.fh);
  ref.setCoarseTrackingRef({&r_ref.fh});
  Vec5 nanres, r_last, r_last2;
  for (int i = 0; i < 5; i++) nanres[i] = NAN;
  SE3 r_pose = SE3::from7(sc.init[0]);
  AffLight r_aff(0, 0);
  ref.new_frame_ = &r_new.fh;
  const bool r_ok = ref.trackNewestCoarse(&r_new.fh, r_pose, r_aff, levels - 1, nanres, r_last);
  const Vec3 r_flow = ref.lastFlowIndicators;
  float r_scale = 1.0f;
  const float r_rmse = ref.optimizeScale(&r_right.fh, r_scale, levels - 1);
  ref.scaleCoarseDepthL0(r_scale);
  SE3 r_pose2 = SE3::from7(sc.init[1]);
  AffLight r_aff2(0, 0);
  ref.new_frame_ = &r_new2.fh;
  const bool r_ok2 = ref.trackNewestCoarse(&r_new2.fh, r_pose2, r_aff2, levels - 1, nanres, r_last2);

  // ---------------- the adapter, same types, same calls ----------------
  dslam_b200::Session session(0);
  dslam_b200::FramePyramids<FrameHessian> pyr(session, w, h, levels);
  Frame g_ref, g_new, g_new2, g_right;
  pyr.makeImages(&g_ref.fh, sc.img_ref.data(), &calib);  // allocates dIp / absSquaredGrad like HessianBlocks.cpp:131-136
  pyr.makeImages(&g_new.fh, sc.img_new.data(), &calib);
  pyr.makeImages(&g_new2.fh, sc.img_new2.data(), &calib);
  pyr.makeImages(&g_right.fh, sc.img_right.data(), (CalibHessian *)nullptr);
  // host mirrors == the reference's arrays, bit for bit (rows 0 and h-1 of dx, dy, absSquaredGrad are uninitialised in the reference)
  long mism = 0;
  Frame *gf[4] = {&g_ref, &g_new, &g_new2, &g_right}, *rf[4] = {&r_ref, &r_new, &r_new2, &r_right};
  for (int k = 0; k < 4; k++)
    for (int l = 0; l < levels; l++) {
      const int wl = w >> l, hl = h >> l;
      for (int i = 0; i < wl * hl; i++) {
        const bool edge = i < wl || i >= wl * (hl - 1);
        const float *a = gf[k]->fh.dIp[l][i].d, *b = rf[k]->fh.dIp[l][i].d;
        if (std::memcmp(a, b, 4) != 0) mism++;
        if (!edge && std::memcmp(a + 1, b + 1, 8) != 0) mism++;
        if (!edge && std::memcmp(&gf[k]->fh.absSquaredGrad[l][i], &rf[k]->fh.absSquaredGrad[l][i], 4) != 0) mism++;
      }
    }
  expect(mism == 0, "makeImages host mirrors equal the reference's arrays bit for bit");
  expect(g_ref.fh.dI == g_ref.fh.dIp[0], "dI = dIp[0]");
  const float K1[4] = {sc.K[0], sc.K[1], sc.K[2], sc.K[3]};
  GpuTracker trk(session, pyr, w, h, levels, tfm, K1);
  trk.makeK(&calib);
  bind_points(&g_ref.fh);
  dslam_b200::ActivePoints pts;  // the export the maintainer writes at the call site (TrackerAndScaler.cpp:149-166)
  for (PointHessian *p : g_ref.fh.pointHessians) {
    PointFrameResidual *r = p->lastResiduals[0].first;
    pts.push((int)(r->centerProjectedTo[0] + 0.5f), (int)(r->centerProjectedTo[1] + 0.5f), r->centerProjectedTo[2],
             sqrtf(1e-3 / (p->efPoint->HdiF + 1e-12)));
  }
  trk.setCoarseTrackingRef({&g_ref.fh}, pts);
  for (int l = 0; l < levels; l++) expect(trk.pc_n()[l] == ref.pc_n_[l], "template size per level equals the reference's pc_n_");
  SE3 g_pose = SE3::from7(sc.init[0]);
  AffLight g_aff(0, 0);
  Vec5 g_last, g_last2;
  const bool g_ok = trk.trackNewestCoarse(&g_new.fh, g_pose, g_aff, levels - 1, nanres, g_last);
  const Vec3 g_flow = trk.lastFlowIndicators;
  float g_scale = 1.0f;
  const float g_rmse = trk.optimizeScale(&g_right.fh, g_scale, levels - 1);
  trk.scaleCoarseDepthL0(g_scale);
  SE3 g_pose2 = SE3::from7(sc.init[1]);
  AffLight g_aff2(0, 0);
  const bool g_ok2 = trk.trackNewestCoarse(&g_new2.fh, g_pose2, g_aff2, levels - 1, nanres, g_last2);

  const double e1 = rel7(g_pose.data(), r_pose.data()), e2 = rel7(g_pose2.data(), r_pose2.data());
  std::printf("track 1: ok %d/%d pose rel %.3e aff (%.6f %.4f | %.6f %.4f) rmse0 %.5f | %.5f\n", (int)g_ok, (int)r_ok, e1, g_aff.a, g_aff.b, r_aff.a, r_aff.b,
              g_last[0], r_last[0]);
  std::printf("scale  : %.6f | %.6f  rmse %.5f | %.5f\n", g_scale, r_scale, g_rmse, r_rmse);
  std::printf("track 2: ok %d/%d pose rel %.3e rmse0 %.5f | %.5f\n", (int)g_ok2, (int)r_ok2, e2, g_last2[0], r_last2[0]);
  expect(g_ok == r_ok && g_ok2 == r_ok2, "trackNewestCoarse return values");
  expect(e1 < 2e-5 && e2 < 2e-5, "poses within the reference's fp32 summation noise (2e-5 relative)");
  expect(std::fabs(g_aff.a - r_aff.a) < 1e-4 && std::fabs(g_aff.b - r_aff.b) < 1e-2, "affine parameters");
  for (int i = 0; i < levels; i++) expect(std::fabs(g_last[i] - r_last[i]) <= 1e-3 * std::fabs(r_last[i]), "lastResiduals");
  for (int i = 0; i < 3; i++) expect(std::fabs(g_flow[i] - r_flow[i]) <= 1e-3 * std::fabs(r_flow[i]) + 1e-9, "lastFlowIndicators");
  expect(std::fabs(g_scale - r_scale) <= 1e-4f * std::fabs(r_scale), "optimizeScale scale");
  expect(std::fabs(g_rmse - r_rmse) <= 1e-3f * std::fabs(r_rmse), "optimizeScale return value");
  expect(trk.refFrameID == ref.refFrameID && trk.lastRef == &g_ref.fh, "pure-output members");

  // ---------------- the hypothesis loop of FrontEnd::trackNewCoarse (:192-247) ----------------
  std::vector<SE3> tries;
  const double off[7] = {0.0, 0.03, 0.0, 0.99955, 0.4, 0.0, 0.0};  // a wrong start first, then the good ones: the loop must move on
  tries.push_back(SE3::from7(off));
  tries.push_back(SE3::from7(sc.init[1]));
  tries.push_back(SE3());
  Vec5 last_rmse;
  for (int i = 0; i < 5; i++) last_rmse[i] = r_last2[i];
  const double reTrack = 1.5;
  // sequential loop on the reference object, as written in FrontEnd.cpp
  Vec5 achievedRes;
  for (int i = 0; i < 5; i++) achievedRes[i] = NAN;
  bool haveOneGood = false;
  int tryIterations = 0;
  SE3 lastF_2_fh;
  AffLight aff_g2l(0, 0);
  for (size_t i = 0; i < tries.size(); i++) {
    AffLight aff_this(0, 0);
    SE3 pose_this = tries[i];
    Vec5 cur;
    const bool good = ref.trackNewestCoarse(&r_new2.fh, pose_this, aff_this, levels - 1, achievedRes, cur);
    tryIterations++;
    if (good && std::isfinite((float)cur[0]) && !(cur[0] >= achievedRes[0])) { aff_g2l = aff_this; lastF_2_fh = pose_this; haveOneGood = true; }
    if (haveOneGood)
      for (int k = 0; k < 5; k++)
        if (!std::isfinite((float)achievedRes[k]) || achievedRes[k] > cur[k]) achievedRes[k] = cur[k];
    if (haveOneGood && achievedRes[0] < last_rmse[0] * reTrack) break;
  }
  SE3 g_l2f;
  AffLight g_affl(0, 0);
  Vec5 g_ach;
  Vec3 g_fv;
  bool g_have = false;
  const int g_tries = trk.trackNewCoarse(&g_new2.fh, tries, AffLight(0, 0), levels - 1, last_rmse, reTrack, g_l2f, g_affl, g_ach, g_fv, g_have);
  std::printf("trackNewCoarse: tries %d | %d good %d | %d pose rel %.3e achieved0 %.5f | %.5f\n", g_tries, tryIterations, (int)g_have, (int)haveOneGood,
              rel7(g_l2f.data(), lastF_2_fh.data()), g_ach[0], achievedRes[0]);
  expect(g_tries == tryIterations && g_have == haveOneGood, "trackNewCoarse: tryIterations / haveOneGood");
  expect(rel7(g_l2f.data(), lastF_2_fh.data()) < 2e-5, "trackNewCoarse: winning pose");
  expect(std::fabs(g_ach[0] - achievedRes[0]) <= 1e-3 * std::fabs(achievedRes[0]), "trackNewCoarse: achievedRes");

  pyr.release(&g_new.fh);
  std::printf("adapter_vs_reference: %s (%d failed checks)\n", fails == 0 ? "PASS" : "FAIL", fails);
  return fails == 0 ? 0 : 1;
}
