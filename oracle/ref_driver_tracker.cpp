// ref_driver_tracker.cpp — C entry points around the REFERENCE'S OWN TrackerAndScaler source, compiled in place.
//
// oracle/ref_build.py writes `tracker_extract.inc` into a temporary directory: lines 1-336 and 451-1172 of
// /root/reference/src/scale_optimization/TrackerAndScaler.cpp (constructor, makeK, makeCoarseDepthL0,
// setCoarseTrackingRef, scaleCoarseDepthL0, trackNewestCoarse, calcGSSSEPose, calcResPose, optimizeScale,
// calcGSSSEScale, calcResScale — the debug plots :338-449 and CoarseDistanceMap :1174-1362 are left out), and compiles
// this file against the reference's real TrackerAndScaler.h / ScaleAccumulator.h, DSO's real MatrixAccumulators.h and
// util/globalFuncs.h (getInterpolatedElement33), and the stand-ins under oracle/shim for Eigen, Sophus, OpenCV and the
// DSO structs.  Nothing of the reference is copied into the repository.  TEST INFRASTRUCTURE ONLY.
//
// What this pins: every formula and the whole control flow of the hot path as written in the reference's source text.
// What it cannot pin: the evaluation order real Eigen / Sophus give those expressions (restated in oracle/shim).
#include <chrono>
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
float setting_huberTH = 9;          // deps:dso/src/util/settings.cpp:127
float setting_coarseCutoffTH = 20;  // :138
float setting_affineOptModeA = 0;   // mode 1 of src/main.cpp:117-122
float setting_affineOptModeB = 0;
bool setting_debugout_runquiet = true;
int setting_gammaWeightsPixelSelect = 1;  // deps:dso/src/util/settings.cpp
float freeDebugParam3 = 1;
}  // namespace dso

// FrameHessian::makeImages, deps:dso/src/FullSystem/HessianBlocks.cpp:128-191 (extracted by oracle/ref_build.py)
namespace dso {
#include "makeimages_extract.inc"
}  // namespace dso

using namespace dso;

namespace {
struct Frame {
  FrameHessian fh;
  FrameShell shell;
};
struct RefTracker {
  std::unique_ptr<TrackerAndScaler> trk;
  CalibHessian calib;
  int w, h, levels;
  Frame ref, cur, right;
  std::vector<PointHessian> ph;
  std::vector<PointFrameResidual> res;
  std::vector<EFPoint> efp;
  EFResidual efr;
};
void bind(Frame &f, const float *dIp_all, int w, int h, int levels, float exposure) {
  size_t off = 0;
  for (int l = 0; l < PYR_LEVELS; l++) f.fh.dIp[l] = nullptr;
  for (int l = 0; l < levels; l++) {
    f.fh.dIp[l] = (Eigen::Vector3f *)(dIp_all + 3 * off);
    off += (size_t)(w >> l) * (h >> l);
  }
  f.fh.ab_exposure = exposure;
  f.fh.shell = &f.shell;
}
}  // namespace

extern "C" {

void *reft_create(int w, int h, int levels, const float K0[4], const float K1[4], const double T[16]) {
  pyrLevelsUsed = levels;
  for (int l = 0; l < levels; l++) { wG[l] = w >> l; hG[l] = h >> l; }
  RefTracker *R = new RefTracker();
  R->w = w; R->h = h; R->levels = levels;
  Mat33f K1m;
  K1m << K1[0], 0.0f, K1[2], 0.0f, K1[1], K1[3], 0.0f, 0.0f, 1.0f;
  R->trk.reset(new TrackerAndScaler(w, h, std::vector<double>(T, T + 16), K1m));
  R->calib = CalibHessian{K0[0], K0[1], K0[2], K0[3]};
  R->trk->makeK(&R->calib);
  return R;
}
void reft_destroy(void *p) { delete (RefTracker *)p; }
void reft_set_aff_mode(float a, float b) { setting_affineOptModeA = a; setting_affineOptModeB = b; }

// setCoarseTrackingRef({ref}) with the active points given as (integer pixel, idepth, HdiF); weight = sqrtf(1e-3/(HdiF+1e-12))
void reft_set_ref(void *p, const float *dIp_ref_all, int npts, const int *pu, const int *pv, const float *pid, const float *hdif, float exposure,
                  double a, double b) {
  RefTracker &R = *(RefTracker *)p;
  bind(R.ref, dIp_ref_all, R.w, R.h, R.levels, exposure);
  R.ref.fh.aff = AffLight(a, b);
  R.ph.assign((size_t)npts, PointHessian());
  R.res.assign((size_t)npts, PointFrameResidual());
  R.efp.assign((size_t)npts, EFPoint());
  R.ref.fh.pointHessians.clear();
  for (int i = 0; i < npts; i++) {
    R.efp[i].HdiF = hdif[i];
    R.res[i].efResidual = &R.efr;
    R.res[i].target = &R.ref.fh;
    R.res[i].centerProjectedTo = Vec3f((float)pu[i], (float)pv[i], pid[i]);
    R.ph[i].lastResiduals[0] = std::make_pair(&R.res[i], ResState::IN);
    R.ph[i].lastResiduals[1] = std::make_pair((PointFrameResidual *)nullptr, ResState::OOB);
    R.ph[i].efPoint = &R.efp[i];
    R.ref.fh.pointHessians.push_back(&R.ph[i]);
  }
  std::vector<FrameHessian *> fhs{&R.ref.fh};
  R.trk->setCoarseTrackingRef(fhs);
}
int reft_get_ref_level(void *p, int lvl, float *u, float *v, float *id, float *c) {
  TrackerAndScaler &t = *((RefTracker *)p)->trk;
  const int n = t.pc_n_[lvl];
  if (u)
    for (int i = 0; i < n; i++) { u[i] = t.pc_u_[lvl][i]; v[i] = t.pc_v_[lvl][i]; id[i] = t.pc_idepth_[lvl][i]; c[i] = t.pc_color_[lvl][i]; }
  return n;
}
void reft_scale_idepth(void *p, float s) { ((RefTracker *)p)->trk->scaleCoarseDepthL0(s); }
void reft_set_new_frame(void *p, const float *dIp_all, float exposure) {
  RefTracker &R = *(RefTracker *)p;
  bind(R.cur, dIp_all, R.w, R.h, R.levels, exposure);
  R.trk->new_frame_ = &R.cur.fh;
}
void reft_set_right_frame(void *p, const float *dIp_all) {
  RefTracker &R = *(RefTracker *)p;
  bind(R.right, dIp_all, R.w, R.h, R.levels, 1.0f);
  R.trk->fh1_ = &R.right.fh;
}
int reft_calc_res_pose(void *p, int lvl, const double pose7[7], double a, double b, float cutoff, double res6[6]) {
  TrackerAndScaler &t = *((RefTracker *)p)->trk;
  const Vec6 r = t.calcResPose(lvl, SE3::from7(pose7), AffLight(a, b), cutoff);
  for (int i = 0; i < 6; i++) res6[i] = r[i];
  return t.pose_buf_warped_n_;
}
void reft_calc_gs_pose(void *p, int lvl, const double pose7[7], double a, double b, double H64[64], double b8[8]) {
  TrackerAndScaler &t = *((RefTracker *)p)->trk;
  Mat88 H;
  Vec8 bb;
  t.calcGSSSEPose(lvl, H, bb, SE3::from7(pose7), AffLight(a, b));
  for (int i = 0; i < 64; i++) H64[i] = H.d[i];
  for (int i = 0; i < 8; i++) b8[i] = bb[i];
}
// which 0: pose buffers (idepth,u,v,dx,dy,residual,weight,refColor), 1: scale buffers (rx1,rx2,rx3,dx,dy,residual,weight,refColor)
int reft_get_warped(void *p, int which, float *out8n) {
  TrackerAndScaler &t = *((RefTracker *)p)->trk;
  const int n = which ? t.scale_buf_warped_n_ : t.pose_buf_warped_n_;
  if (!out8n) return n;
  float *src[8];
  if (which == 0) {
    float *s[8] = {t.pose_buf_warped_idepth_, t.pose_buf_warped_u_, t.pose_buf_warped_v_, t.pose_buf_warped_dx_, t.pose_buf_warped_dy_,
                   t.pose_buf_warped_residual_, t.pose_buf_warped_weight_, t.pose_buf_warped_refColor_};
    for (int i = 0; i < 8; i++) src[i] = s[i];
  } else {
    float *s[8] = {t.scale_buf_warped_rx1_, t.scale_buf_warped_rx2_, t.scale_buf_warped_rx3_, t.scale_buf_warped_dx_, t.scale_buf_warped_dy_,
                   t.scale_buf_warped_residual_, t.scale_buf_warped_weight_, t.scale_buf_warped_ref_color_};
    for (int i = 0; i < 8; i++) src[i] = s[i];
  }
  for (int b = 0; b < 8; b++)
    for (int i = 0; i < n; i++) out8n[(size_t)b * n + i] = src[b][i];
  return n;
}
int reft_track(void *p, double pose7_io[7], double aff_io[2], int coarsestLvl, const double minRes[5], double last[5], double flow[3]) {
  RefTracker &R = *(RefTracker *)p;
  SE3 pose = SE3::from7(pose7_io);
  AffLight aff(aff_io[0], aff_io[1]);
  Vec5 mr, lr;
  for (int i = 0; i < 5; i++) mr[i] = minRes[i];
  const bool ok = R.trk->trackNewestCoarse(&R.cur.fh, pose, aff, coarsestLvl, mr, lr);
  pose.to7(pose7_io);
  aff_io[0] = aff.a; aff_io[1] = aff.b;
  for (int i = 0; i < 5; i++) last[i] = lr[i];
  for (int i = 0; i < 3; i++) flow[i] = R.trk->lastFlowIndicators[i];
  return ok ? 1 : 0;
}
int reft_calc_res_scale(void *p, int lvl, float scale, float cutoff, double res6[6]) {
  TrackerAndScaler &t = *((RefTracker *)p)->trk;
  const Vec6 r = t.calcResScale(lvl, scale, cutoff);
  for (int i = 0; i < 6; i++) res6[i] = r[i];
  return t.scale_buf_warped_n_;
}
void reft_calc_gs_scale(void *p, int lvl, float scale, float *H, float *b) { ((RefTracker *)p)->trk->calcGSSSEScale(lvl, *H, *b, scale); }
float reft_optimize_scale(void *p, float *scale_io, int coarsestLvl) {
  RefTracker &R = *(RefTracker *)p;
  return R.trk->optimizeScale(&R.right.fh, *scale_io, coarsestLvl);
}
// The reference's makeImages on a w x h image with `levels` pyramid levels.  Outputs in the oracle's layout (all levels
// concatenated); rows 0 and h-1 of dx, dy, absSquaredGrad are whatever `new[]` returned in the reference — zeroed here
// before the call is not possible (the function allocates), so they are zeroed afterwards.
void refimg_make_images(const float *color, int w, int h, int levels, const float *B256, float *dIp_all, float *absg_all) {
  if (pyrLevelsUsed != levels || wG[0] != w || hG[0] != h) {  // globals of the reference; only touched when the geometry changes
    pyrLevelsUsed = levels;
    for (int l = 0; l < levels; l++) { wG[l] = w >> l; hG[l] = h >> l; }
  }
  FrameHessian fh;
  CalibHessian calib{0, 0, 0, 0};
  if (B256) for (int i = 0; i < 256; i++) calib.B[i] = B256[i];
  fh.makeImages(const_cast<float *>(color), B256 ? &calib : nullptr);
  size_t off = 0;
  for (int l = 0; l < levels; l++) {
    const int wl = w >> l, hl = h >> l;
    for (int i = 0; i < wl * hl; i++) {
      const bool edge = i < wl || i >= wl * (hl - 1);
      dIp_all[3 * (off + i) + 0] = fh.dIp[l][i][0];
      dIp_all[3 * (off + i) + 1] = edge ? 0.f : fh.dIp[l][i][1];
      dIp_all[3 * (off + i) + 2] = edge ? 0.f : fh.dIp[l][i][2];
      absg_all[off + i] = edge ? 0.f : fh.absSquaredGrad[l][i];
    }
    off += (size_t)wl * hl;
    delete[] fh.dIp[l];
    delete[] fh.absSquaredGrad[l];
  }
}

// One stereo stream driven the way FrontEnd drives it, entirely on the C side (timed CPU baseline of bench.py): per frame
// `new FrameHessian; makeImages(left, &HCalib)` + trackNewestCoarse (src/FrontEnd.cpp:585-606, 204-206); on keyframes
// `makeImages(right, 0)` + optimizeScale from seed 1.0 (:676-680, 992).  The frames own their pyramids exactly as in the
// reference (new[] inside makeImages, delete[] when the frame goes) — no copies, no Python objects, no globals written per
// frame.  Frame k uses image k & 1 and pose_init k & 1; it is a keyframe when (k + phase) % kf_every == 0.
// Returns the number of frames whose tracking succeeded; pose7_out / scale_out receive the last results.
int reft_run_frames(void *p, int k0, int k1, int phase, int kf_every, const float *img_new0, const float *img_new1, const float *img_right,
                    const double *pose_init_2x7, int coarsestLvl, double *pose7_out, float *scale_out, double *phase_s /* [3] or null */) {
  RefTracker &R = *(RefTracker *)p;
  typedef std::chrono::steady_clock Clk;
  double t_img = 0, t_trk = 0, t_kf = 0;
  for (int i = 0; i < 256; i++) R.calib.B[i] = (float)i;  // identity response (HessianBlocks.h:329-330)
  int good = 0;
  Vec5 mr, lr;
  for (int i = 0; i < 5; i++) mr[i] = NAN;
  for (int k = k0; k < k1; k++) {
    const int v = k & 1;
    Frame *f = new Frame();
    f->fh.shell = &f->shell;
    f->fh.ab_exposure = 1.0f;
    const Clk::time_point c0 = Clk::now();
    f->fh.makeImages(const_cast<float *>(v ? img_new1 : img_new0), &R.calib);
    const Clk::time_point c1 = Clk::now();
    R.trk->new_frame_ = &f->fh;
    SE3 pose = SE3::from7(pose_init_2x7 + 7 * v);
    AffLight aff(0, 0);
    good += R.trk->trackNewestCoarse(&f->fh, pose, aff, coarsestLvl, mr, lr) ? 1 : 0;
    if (pose7_out) pose.to7(pose7_out);
    const Clk::time_point c2 = Clk::now();
    t_img += std::chrono::duration<double>(c1 - c0).count();
    t_trk += std::chrono::duration<double>(c2 - c1).count();
    if ((k + phase) % kf_every == 0) {
      Frame *f1 = new Frame();
      f1->fh.shell = &f1->shell;
      f1->fh.makeImages(const_cast<float *>(img_right), nullptr);
      float scale = 1.0f;
      R.trk->optimizeScale(&f1->fh, scale, coarsestLvl);
      if (scale_out) *scale_out = scale;
      for (int l = 0; l < R.levels; l++) { delete[] f1->fh.dIp[l]; delete[] f1->fh.absSquaredGrad[l]; }
      delete f1;
      t_kf += std::chrono::duration<double>(Clk::now() - c2).count();
    }
    for (int l = 0; l < R.levels; l++) { delete[] f->fh.dIp[l]; delete[] f->fh.absSquaredGrad[l]; }
    delete f;
  }
  if (phase_s) { phase_s[0] = t_img; phase_s[1] = t_trk; phase_s[2] = t_kf; }
  return good;
}

void reft_get_K(void *p, int lvl, float out[17]) {
  TrackerAndScaler &t = *((RefTracker *)p)->trk;
  out[0] = t.fx_[lvl]; out[1] = t.fy_[lvl]; out[2] = t.cx_[lvl]; out[3] = t.cy_[lvl];
  for (int i = 0; i < 9; i++) out[4 + i] = t.Ki_[lvl].d[i];
  out[13] = t.fx1_[lvl]; out[14] = t.fy1_[lvl]; out[15] = t.cx1_[lvl]; out[16] = t.cy1_[lvl];
}

}  // extern "C"
