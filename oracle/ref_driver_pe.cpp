// ref_driver_pe.cpp — C entry points around the REFERENCE'S OWN PoseEstimator.cpp, compiled in place as a whole
// (src/loop_closure/pose_estimation/PoseEstimator.{h,cpp}) against the stand-ins of oracle/shim, DSO's real
// MatrixAccumulators.h and util/globalFuncs.h.  Built by oracle/ref_build.py; TEST INFRASTRUCTURE ONLY.
#include <cstdint>
#include <stdint.h>
#include <vector>
using std::uintptr_t;

#define private public
#include "loop_closure/pose_estimation/PoseEstimator.cpp"
#undef private

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
}  // namespace dso

using namespace dso;

namespace {
struct RefPE {
  PoseEstimator *pe;
  int w, h, levels;
  std::vector<float> cam;
  std::vector<std::pair<Eigen::Vector3d, float *>> pts;
  std::vector<std::vector<float>> colors;  // per point: levels floats
  float ref_exposure = 1.f;
  FrameHessian fh;
  FrameShell shell;
};
}  // namespace

extern "C" {

void *refpe_create(int w, int h, int levels, const float cam[4]) {
  pyrLevelsUsed = levels;
  for (int l = 0; l < levels; l++) { wG[l] = w >> l; hG[l] = h >> l; }
  RefPE *R = new RefPE();
  R->w = w; R->h = h; R->levels = levels;
  R->cam.assign(cam, cam + 4);
  R->pe = new PoseEstimator(w, h);
  return R;
}
void refpe_destroy(void *p) { delete ((RefPE *)p)->pe; delete (RefPE *)p; }
void refpe_set_aff_mode(float a, float b) { setting_affineOptModeA = a; setting_affineOptModeB = b; }
void refpe_set_points(void *p, int n, const double *pts, const float *colors, float ref_exposure) {
  RefPE &R = *(RefPE *)p;
  R.colors.assign((size_t)n, std::vector<float>((size_t)R.levels));
  R.pts.clear();
  for (int i = 0; i < n; i++) {
    for (int l = 0; l < R.levels; l++) R.colors[i][l] = colors[(size_t)l * n + i];
    R.pts.emplace_back(Eigen::Vector3d(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]), R.colors[i].data());
  }
  R.ref_exposure = ref_exposure;
}
void refpe_set_new_frame(void *p, const float *dIp_all, float exposure) {
  RefPE &R = *(RefPE *)p;
  size_t off = 0;
  for (int l = 0; l < R.levels; l++) {
    R.fh.dIp[l] = (Eigen::Vector3f *)(dIp_all + 3 * off);
    off += (size_t)(R.w >> l) * (R.h >> l);
  }
  R.fh.ab_exposure = exposure;
  R.fh.shell = &R.shell;
}
int refpe_estimate(void *p, double T_io[16], int coarsest_lvl, float *pose_error) {
  RefPE &R = *(RefPE *)p;
  Eigen::Matrix4d T;
  for (int i = 0; i < 16; i++) T.d[i] = T_io[i];
  const bool ok = R.pe->estimate(R.pts, R.ref_exposure, &R.fh, R.cam, coarsest_lvl, T, *pose_error);
  for (int i = 0; i < 16; i++) T_io[i] = T.d[i];
  return ok ? 1 : 0;
}
// one calcRes + calcGSSSE at a given pose (private members reached through the "#define private public" above)
int refpe_calc_res(void *p, int lvl, const double T16[16], double aff_a, double aff_b, float cutoff, double res6[6], double H64[64], double b8[8]) {
  RefPE &R = *(RefPE *)p;
  R.pe->makeK(R.cam);
  R.pe->pts_ = R.pts;
  R.pe->new_frame_ = &R.fh;
  R.pe->ref_aff_g2l_ = AffLight();
  R.pe->ref_ab_exposure_ = R.ref_exposure;
  Eigen::Matrix4d T;
  for (int i = 0; i < 16; i++) T.d[i] = T16[i];
  const SE3 pose(T.block<3, 3>(0, 0), T.block<3, 1>(0, 3));
  const Vec6 r = R.pe->calcRes(lvl, pose, AffLight(aff_a, aff_b), cutoff);
  for (int i = 0; i < 6; i++) res6[i] = r[i];
  Mat88 H;
  Vec8 b;
  R.pe->calcGSSSE(lvl, H, b, pose, AffLight(aff_a, aff_b));
  for (int i = 0; i < 64; i++) H64[i] = H.d[i];
  for (int i = 0; i < 8; i++) b8[i] = b[i];
  return R.pe->buf_warped_n_;
}

}  // extern "C"
