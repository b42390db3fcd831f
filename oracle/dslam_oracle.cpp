// dslam_oracle.cpp — CPU restatement of the direct_stereo_slam photometric hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under direct_stereo_slam_b200/ may include, link or call this
// file.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// load the shared library built from it (oracle/_build/libdslam_oracle.so).
//
// PARITY STATUS.  The reference (IRVLab/direct_stereo_slam @ fd12853c, DSO @ aca17755 inside dependencies.zip) ships no
// tests, golden vectors or fixtures for this path and cannot be built as a whole in this image (no Eigen,
// Sophus' dependencies, OpenCV, ROS).  The restatement below is pinned as follows:
//   * BIT-EXACT against the reference's OWN SOURCE TEXT compiled in place (oracle/ref_build.py -> oracle/_ref/, g++ on the
//     reference's files, nothing copied): TrackerAndScaler.cpp hot path (:1-336, :451-1172 — constructor, makeK,
//     makeCoarseDepthL0, trackNewestCoarse, calcResPose, calcGSSSEPose, optimizeScale, calcResScale, calcGSSSEScale),
//     FrameHessian::makeImages (deps:dso HessianBlocks.cpp:128-191), Accumulator9 (deps:dso MatrixAccumulators.h),
//     ScaleAccumulator.h, getInterpolatedElement33 (deps:dso util/globalFuncs.h), search_place.h (search_ringkey,
//     search_sc).  tests/test_oracle_ref.py, tests/test_oracle_ref_tracker.py.
//   * Those sources are compiled against stand-ins (oracle/shim) for Eigen, Sophus::SE3d, OpenCV and the DSO structs, so
//     what remains UNPINNED is only what the stand-ins restate: the evaluation order Eigen 3.3 gives the small fixed-size
//     expressions (3-term coefficient products reduce as e0 + (e1 + e2); 3x3 inverse by cofactors; `scale * M * v`
//     keeps the scalar in the lhs coefficients), Eigen's LDLT (any backward-stable 8x8 double solve agrees to ~1e-15)
//     and Sophus' quaternion SE3 exp / product.  FLANN (un-vendored, unpinned; Ubuntu 20.04 libflann-dev 1.9.1) is
//     replaced by exact brute force in flann::L2's summation order.  ScanContext::generate (sc_generate.cpp) is unpinned
//     (it depends on Eigen's SelfAdjointEigenSolver sign conventions).
// Every function restates the arithmetic of the cited reference lines in the same operation order, with floating-point
// contraction OFF (-ffp-contract=off).
//
// "src/..."  = /root/reference/src/...          "deps:dso/..." = dso/ inside dependencies.zip
//
// Build: see oracle/Makefile  (g++ -O2 -ffp-contract=off for goldens; -O3 -march=native
// -ffp-contract=off for the timed CPU baseline — contraction stays off so both builds agree bitwise).

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include <utility>

namespace {

constexpr int kMaxLevels = 6;                 // PYR_LEVELS, deps:dso/src/util/settings.h:50
constexpr float kHuberTH = 9.0f;              // setting_huberTH, deps:dso/src/util/settings.cpp:127
constexpr float kCoarseCutoffTH = 20.0f;      // setting_coarseCutoffTH, settings.cpp:138
constexpr float kScaleXiRot = 1.0f;           // SCALE_XI_ROT   deps:dso/src/FullSystem/HessianBlocks.h:59
constexpr float kScaleXiTrans = 0.5f;         // SCALE_XI_TRANS HessianBlocks.h:60
constexpr float kScaleA = 10.0f;              // SCALE_A        HessianBlocks.h:64
constexpr float kScaleB = 1000.0f;            // SCALE_B        HessianBlocks.h:65

// ------------------------------------------------------------------------------------------------
// Sophus SE3d restated: unit quaternion (x,y,z,w) + translation, all double.
// deps:dso/thirdparty/Sophus/sophus/se3.hpp:160-163 (fastMultiply), :239-243, :268-271 (operator*=),
// :407-428 (exp); so3.hpp:165-167, :196-202 (normalize), :343-369 (expAndTheta), :631-633 (ctor).
// ------------------------------------------------------------------------------------------------
struct SE3 {
  double q[4];  // x y z w  (Eigen::Quaterniond::coeffs() order == SE3d::data() order)
  double t[3];
};

inline void quat_normalize(double q[4]) {
  double len = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; i++) q[i] /= len;
}

// Eigen quaternion product a*b (Eigen/src/Geometry/Quaternion.h quat_product generic form).
inline void quat_mul(const double a[4], const double b[4], double r[4]) {
  const double ax = a[0], ay = a[1], az = a[2], aw = a[3];
  const double bx = b[0], by = b[1], bz = b[2], bw = b[3];
  r[3] = aw * bw - ax * bx - ay * by - az * bz;
  r[0] = aw * bx + ax * bw + ay * bz - az * by;
  r[1] = aw * by + ay * bw + az * bx - ax * bz;
  r[2] = aw * bz + az * bw + ax * by - ay * bx;
}

// Eigen QuaternionBase::_transformVector: uv = 2*(qv x v); v + w*uv + qv x uv.
inline void quat_rotate(const double q[4], const double v[3], double r[3]) {
  double uv[3] = {q[1] * v[2] - q[2] * v[1], q[2] * v[0] - q[0] * v[2], q[0] * v[1] - q[1] * v[0]};
  uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
  double c[3] = {q[1] * uv[2] - q[2] * uv[1], q[2] * uv[0] - q[0] * uv[2], q[0] * uv[1] - q[1] * uv[0]};
  for (int i = 0; i < 3; i++) r[i] = v[i] + q[3] * uv[i] + c[i];
}

// Eigen QuaternionBase::toRotationMatrix (row-major output).
inline void quat_to_R(const double q[4], double R[9]) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

inline SE3 se3_identity() { SE3 s; s.q[0] = s.q[1] = s.q[2] = 0; s.q[3] = 1; s.t[0] = s.t[1] = s.t[2] = 0; return s; }

// a * b   (se3.hpp:239-243 -> operator*= -> fastMultiply + normalize)
inline SE3 se3_mul(const SE3 &a, const SE3 &b) {
  SE3 r = a;
  double rt[3];
  quat_rotate(a.q, b.t, rt);
  for (int i = 0; i < 3; i++) r.t[i] = a.t[i] + rt[i];
  quat_mul(a.q, b.q, r.q);
  quat_normalize(r.q);
  return r;
}

// se3.hpp:407-428 ; tangent = (upsilon[3], omega[3])
inline SE3 se3_exp(const double a[6]) {
  const double *ups = a, *om = a + 3;
  const double theta_sq = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
  const double theta = std::sqrt(theta_sq);
  const double half_theta = 0.5 * theta;
  double imag, real;
  const double eps = 1e-10;  // SophusConstants<double>::epsilon()
  if (theta < eps) {
    const double theta_po4 = theta_sq * theta_sq;
    imag = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * theta_po4;
    real = 1.0 - 0.5 * theta_sq + (1.0 / 384.0) * theta_po4;
  } else {
    const double s = std::sin(half_theta);
    imag = s / theta;
    real = std::cos(half_theta);
  }
  SE3 r;
  r.q[3] = real; r.q[0] = imag * om[0]; r.q[1] = imag * om[1]; r.q[2] = imag * om[2];
  quat_normalize(r.q);  // SO3Group(Quaternion) ctor normalises, so3.hpp:631-633
  double Om[9] = {0, -om[2], om[1], om[2], 0, -om[0], -om[1], om[0], 0};
  double Om2[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += Om[i * 3 + k] * Om[k * 3 + j];
      Om2[i * 3 + j] = s;
    }
  double V[9];
  if (theta < eps) {
    quat_to_R(r.q, V);
  } else {
    const double c1 = (1.0 - std::cos(theta)) / theta_sq;
    const double c2 = (theta - std::sin(theta)) / (theta_sq * theta);
    for (int i = 0; i < 9; i++) V[i] = ((i % 4 == 0) ? 1.0 : 0.0) + c1 * Om[i] + c2 * Om2[i];
  }
  for (int i = 0; i < 3; i++) r.t[i] = V[i * 3 + 0] * ups[0] + V[i * 3 + 1] * ups[1] + V[i * 3 + 2] * ups[2];
  return r;
}

// ------------------------------------------------------------------------------------------------
// Pivoted LDL^T solve for n<=8 (stands in for Eigen's Hl.ldlt().solve(-b), TrackerAndScaler.cpp:509;
// any backward-stable double solve agrees to O(1e-15)).
// ------------------------------------------------------------------------------------------------
void ldlt_solve(int n, const double *Ain, int lda, const double *rhs, double *x) {
  double A[64];
  int perm[8];
  for (int i = 0; i < n; i++) {
    perm[i] = i;
    for (int j = 0; j < n; j++) A[i * 8 + j] = Ain[i * lda + j];
  }
  for (int k = 0; k < n; k++) {
    int p = k;
    double best = std::fabs(A[k * 8 + k]);
    for (int i = k + 1; i < n; i++)
      if (std::fabs(A[i * 8 + i]) > best) { best = std::fabs(A[i * 8 + i]); p = i; }
    if (p != k) {
      for (int j = 0; j < n; j++) std::swap(A[k * 8 + j], A[p * 8 + j]);
      for (int i = 0; i < n; i++) std::swap(A[i * 8 + k], A[i * 8 + p]);
      std::swap(perm[k], perm[p]);
    }
    const double d = A[k * 8 + k];
    if (d == 0.0) continue;
    for (int i = k + 1; i < n; i++) {
      const double l = A[i * 8 + k] / d;
      for (int j = k + 1; j < n; j++) A[i * 8 + j] -= l * A[k * 8 + j];
      A[i * 8 + k] = l;
    }
  }
  double y[8];
  for (int i = 0; i < n; i++) {
    double s = rhs[perm[i]];
    for (int j = 0; j < i; j++) s -= A[i * 8 + j] * y[j];
    y[i] = s;
  }
  for (int i = 0; i < n; i++) {
    const double d = A[i * 8 + i];
    y[i] = (std::fabs(d) > 1e-300) ? y[i] / d : 0.0;
  }
  double z[8];
  for (int i = n - 1; i >= 0; i--) {
    double s = y[i];
    for (int j = i + 1; j < n; j++) s -= A[j * 8 + i] * z[j];
    z[i] = s;
  }
  for (int i = 0; i < n; i++) x[perm[i]] = z[i];
}

// Eigen 3.3.7 coefficient-based product of 3 terms: (a.cwiseProduct(b)).sum() with complete
// unrolling reduces as e0 + (e1 + e2)  (Eigen/src/Core/Redux.h redux_novec_unroller).
inline float dot3f(float a0, float b0, float a1, float b1, float a2, float b2) {
  return a0 * b0 + (a1 * b1 + a2 * b2);
}

// Mat33f product (row-major storage here).
inline void mat33f_mul(const float A[9], const float B[9], float C[9]) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      C[i * 3 + j] = dot3f(A[i * 3 + 0], B[0 * 3 + j], A[i * 3 + 1], B[1 * 3 + j], A[i * 3 + 2], B[2 * 3 + j]);
}

// Eigen 3x3 inverse by cofactors (Eigen/src/LU/InverseImpl.h compute_inverse<Matrix3f>), float.
inline float cof3(const float m[9], int i, int j) {
  const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
  return m[i1 * 3 + j1] * m[i2 * 3 + j2] - m[i1 * 3 + j2] * m[i2 * 3 + j1];
}
inline void mat33f_inverse(const float m[9], float r[9]) {
  const float c0 = cof3(m, 0, 0), c1 = cof3(m, 1, 0), c2 = cof3(m, 2, 0);
  const float det = dot3f(c0, m[0], c1, m[3], c2, m[6]);
  const float invdet = 1.0f / det;
  r[0] = c0 * invdet; r[1] = c1 * invdet; r[2] = c2 * invdet;
  r[3] = cof3(m, 0, 1) * invdet; r[4] = cof3(m, 1, 1) * invdet; r[5] = cof3(m, 2, 1) * invdet;
  r[6] = cof3(m, 0, 2) * invdet; r[7] = cof3(m, 1, 2) * invdet; r[8] = cof3(m, 2, 2) * invdet;
}

// AffLight::fromToVecExposure  deps:dso/src/util/NumType.h:173-185
inline void aff_from_to(float exposureF, float exposureT, double aF, double bF, double aT, double bT, double out[2]) {
  if (exposureF == 0 || exposureT == 0) exposureT = exposureF = 1;
  const double a = std::exp(aT - aF) * exposureT / exposureF;
  const double b = bT - a * bF;
  out[0] = a; out[1] = b;
}

// getInterpolatedElement33  deps:dso/src/util/globalFuncs.h:75-89  (mat = AoS float[3] per pixel)
inline void interp33(const float *mat, float x, float y, int width, float out[3]) {
  const int ix = (int)x;
  const int iy = (int)y;
  const float dx = x - ix;
  const float dy = y - iy;
  const float dxdy = dx * dy;
  const float *bp = mat + 3 * (ix + iy * width);
  const float w11 = dxdy, w01 = dy - dxdy, w10 = dx - dxdy, w00 = 1 - dx - dy + dxdy;
  for (int c = 0; c < 3; c++)
    out[c] = w11 * bp[3 * (1 + width) + c] + w01 * bp[3 * width + c] + w10 * bp[3 + c] + w00 * bp[c];
}

// ------------------------------------------------------------------------------------------------
// SSE accumulators restated lane-by-lane (each __m128 op = 4 independent IEEE fp32 ops).
// Accumulator9: deps:dso/src/OptimizationBackend/MatrixAccumulators.h:982-1345
// ScaleAccumulator: src/scale_optimization/ScaleAccumulator.h:27-106
// ------------------------------------------------------------------------------------------------
template <int NJ>  // NJ = 9 (pose: J0..J7,r) or 2 (scale: J,r)
struct TieredAcc {
  static constexpr int NE = NJ * (NJ + 1) / 2;
  float d1[NE][4], d1k[NE][4], d1m[NE][4];
  float numIn1, numIn1k, numIn1m;
  void initialize() {
    std::memset(d1, 0, sizeof(d1)); std::memset(d1k, 0, sizeof(d1k)); std::memset(d1m, 0, sizeof(d1m));
    numIn1 = numIn1k = numIn1m = 0;
  }
  // updateSSE_eighted (:1091-1166) / updateSSE_oneed (ScaleAccumulator.h:60-77): J[j][lane], w[lane]
  void update(const float J[NJ][4], const float w[4]) {
    int e = 0;
    for (int r = 0; r < NJ; r++) {
      float Jw[4];
      for (int l = 0; l < 4; l++) Jw[l] = J[r][l] * w[l];
      for (int c = r; c < NJ; c++, e++)
        for (int l = 0; l < 4; l++) d1[e][l] = d1[e][l] + Jw[l] * J[c][l];
    }
    numIn1++;
    shiftUp(false);
  }
  void shiftUp(bool force) {  // :1325-1344
    if (numIn1 > 1000 || force) {
      for (int e = 0; e < NE; e++) for (int l = 0; l < 4; l++) d1k[e][l] = d1[e][l] + d1k[e][l];
      numIn1k += numIn1; numIn1 = 0; std::memset(d1, 0, sizeof(d1));
    }
    if (numIn1k > 1000 || force) {
      for (int e = 0; e < NE; e++) for (int l = 0; l < 4; l++) d1m[e][l] = d1k[e][l] + d1m[e][l];
      numIn1m += numIn1k; numIn1k = 0; std::memset(d1k, 0, sizeof(d1k));
    }
  }
  // finish (:1001-1017): out[e] = ((l0+l1)+l2)+l3 of the 1m tier, upper triangle row-major
  void finish(float out[NE]) {
    shiftUp(true);
    for (int e = 0; e < NE; e++) out[e] = d1m[e][0] + d1m[e][1] + d1m[e][2] + d1m[e][3];
  }
};

struct TraceRec {  // one LM iteration (or level start with iteration=-1)
  int lvl, iteration, accept, n;
  double lambda, e_old, e_new;  // E/nTerms before / after
  double inc[8];
};

struct Tracker {
  int levels = 0;
  int w[kMaxLevels], h[kMaxLevels];
  float fx[kMaxLevels], fy[kMaxLevels], cx[kMaxLevels], cy[kMaxLevels];
  float Ki[kMaxLevels][9];
  float fx1[kMaxLevels], fy1[kMaxLevels], cx1[kMaxLevels], cy1[kMaxLevels];
  SE3 tfm_f1_f0;
  // template (pc_* buffers, TrackerAndScaler.h:90-94)
  std::vector<float> pc_u[kMaxLevels], pc_v[kMaxLevels], pc_idepth[kMaxLevels], pc_color[kMaxLevels];
  int pc_n[kMaxLevels];
  // warped buffers (TrackerAndScaler.h:97-105, 125-133): 0 idepth/rx1, 1 u/rx2, 2 v/rx3, 3 dx, 4 dy, 5 residual, 6 weight, 7 refColor
  std::vector<float> pbuf[8], sbuf[8];
  int pose_n = 0, scale_n = 0;
  // frames
  const float *dIp_new[kMaxLevels];    // new left frame, AoS (I,dx,dy)
  const float *dIp_right[kMaxLevels];  // right frame (fh1_)
  float new_exposure = 1.f, ref_exposure = 1.f;
  double ref_a = 0, ref_b = 0;  // lastRef_aff_g2l
  int affModeA = 0, affModeB = 0;  // setting_affineOptModeA/B (mode=1 default: src/main.cpp:117-122)
  double lastFlow[3];
  std::vector<TraceRec> trace;
  long n_res_evals = 0, n_gs_evals = 0;  // bookkeeping for GN-iterations/s
  // 0: E and the flow-indicator sums accumulate in fp32 in point order exactly like the reference (:797-809, :776-783)
  // 1: the same fp32 terms are summed in fp64 — the accumulation the CUDA path implements (see calc_gs_pose mode 1)
  int res_acc_mode = 0;
};

void make_K(Tracker &T, int w0, int h0, float fx0, float fy0, float cx0, float cy0) {
  // TrackerAndScaler::makeK  src/scale_optimization/TrackerAndScaler.cpp:117-141
  T.w[0] = w0; T.h[0] = h0; T.fx[0] = fx0; T.fy[0] = fy0; T.cx[0] = cx0; T.cy[0] = cy0;
  for (int l = 1; l < T.levels; l++) {
    T.w[l] = T.w[0] >> l; T.h[l] = T.h[0] >> l;
    T.fx[l] = T.fx[l - 1] * 0.5; T.fy[l] = T.fy[l - 1] * 0.5;
    T.cx[l] = (T.cx[0] + 0.5) / ((int)1 << l) - 0.5;
    T.cy[l] = (T.cy[0] + 0.5) / ((int)1 << l) - 0.5;
  }
  for (int l = 0; l < T.levels; l++) {
    const float K[9] = {T.fx[l], 0.0f, T.cx[l], 0.0f, T.fy[l], T.cy[l], 0.0f, 0.0f, 1.0f};
    mat33f_inverse(K, T.Ki[l]);
  }
}

void make_K1(Tracker &T, float fx0, float fy0, float cx0, float cy0) {
  // TrackerAndScaler ctor  src/scale_optimization/TrackerAndScaler.cpp:88-98
  T.fx1[0] = fx0; T.fy1[0] = fy0; T.cx1[0] = cx0; T.cy1[0] = cy0;
  for (int l = 1; l < T.levels; l++) {
    T.fx1[l] = T.fx1[l - 1] * 0.5; T.fy1[l] = T.fy1[l - 1] * 0.5;
    T.cx1[l] = (T.cx1[0] + 0.5) / ((int)1 << l) - 0.5;
    T.cy1[l] = (T.cy1[0] + 0.5) / ((int)1 << l) - 0.5;
  }
}

// ------------------------------------------------------------------------------------------------
// calcResPose  src/scale_optimization/TrackerAndScaler.cpp:699-852
// ------------------------------------------------------------------------------------------------
void calc_res_pose(Tracker &T, int lvl, const SE3 &refToNew, double aff_a, double aff_b, float cutoffTH, double rs[6]) {
  T.n_res_evals++;
  float E = 0;
  double Ed = 0, sTd = 0, sRTd = 0;  // res_acc_mode 1
  int numTermsInE = 0, numTermsInWarped = 0, numSaturated = 0;
  const int wl = T.w[lvl], hl = T.h[lvl];
  const float *dINewl = T.dIp_new[lvl];
  const float fxl = T.fx[lvl], fyl = T.fy[lvl], cxl = T.cx[lvl], cyl = T.cy[lvl];
  double Rd[9]; quat_to_R(refToNew.q, Rd);
  float Rf[9]; for (int i = 0; i < 9; i++) Rf[i] = (float)Rd[i];
  float RKi[9]; mat33f_mul(Rf, T.Ki[lvl], RKi);
  const float t[3] = {(float)refToNew.t[0], (float)refToNew.t[1], (float)refToNew.t[2]};
  double affd[2]; aff_from_to(T.ref_exposure, T.new_exposure, T.ref_a, T.ref_b, aff_a, aff_b, affd);
  const float affLL[2] = {(float)affd[0], (float)affd[1]};
  float sumSquaredShiftT = 0, sumSquaredShiftRT = 0, sumSquaredShiftNum = 0;
  const float maxEnergy = 2 * kHuberTH * cutoffTH - kHuberTH * kHuberTH;
  const int nl = T.pc_n[lvl];
  const float *lpc_u = T.pc_u[lvl].data(), *lpc_v = T.pc_v[lvl].data();
  const float *lpc_idepth = T.pc_idepth[lvl].data(), *lpc_color = T.pc_color[lvl].data();
  const float *Ki = T.Ki[lvl];
  for (int b = 0; b < 8; b++) if ((int)T.pbuf[b].size() < nl + 4) T.pbuf[b].resize(nl + 4);

  for (int i = 0; i < nl; i++) {
    const float id = lpc_idepth[i], x = lpc_u[i], y = lpc_v[i];
    float pt[3];
    for (int r = 0; r < 3; r++) pt[r] = dot3f(RKi[r * 3], x, RKi[r * 3 + 1], y, RKi[r * 3 + 2], 1.0f) + t[r] * id;
    const float u = pt[0] / pt[2], v = pt[1] / pt[2];
    const float Ku = fxl * u + cxl, Kv = fyl * v + cyl;
    const float new_idepth = id / pt[2];

    if (lvl == 0 && i % 32 == 0) {  // :754-784
      float ptT[3], ptT2[3], pt3[3];
      for (int r = 0; r < 3; r++) {
        const float kx = dot3f(Ki[r * 3], x, Ki[r * 3 + 1], y, Ki[r * 3 + 2], 1.0f);
        const float rx = dot3f(RKi[r * 3], x, RKi[r * 3 + 1], y, RKi[r * 3 + 2], 1.0f);
        ptT[r] = kx + t[r] * id; ptT2[r] = kx - t[r] * id; pt3[r] = rx - t[r] * id;
      }
      const float uT = ptT[0] / ptT[2], vT = ptT[1] / ptT[2];
      const float KuT = fxl * uT + cxl, KvT = fyl * vT + cyl;
      const float uT2 = ptT2[0] / ptT2[2], vT2 = ptT2[1] / ptT2[2];
      const float KuT2 = fxl * uT2 + cxl, KvT2 = fyl * vT2 + cyl;
      const float u3 = pt3[0] / pt3[2], v3 = pt3[1] / pt3[2];
      const float Ku3 = fxl * u3 + cxl, Kv3 = fyl * v3 + cyl;
      const float sT1 = (KuT - x) * (KuT - x) + (KvT - y) * (KvT - y), sT2 = (KuT2 - x) * (KuT2 - x) + (KvT2 - y) * (KvT2 - y);
      const float sRT1 = (Ku - x) * (Ku - x) + (Kv - y) * (Kv - y), sRT2 = (Ku3 - x) * (Ku3 - x) + (Kv3 - y) * (Kv3 - y);
      sumSquaredShiftT += sT1;
      sumSquaredShiftT += sT2;
      sumSquaredShiftRT += sRT1;
      sumSquaredShiftRT += sRT2;
      sTd += (double)sT1; sTd += (double)sT2; sRTd += (double)sRT1; sRTd += (double)sRT2;
      sumSquaredShiftNum += 2;
    }

    if (!(Ku > 2 && Kv > 2 && Ku < wl - 3 && Kv < hl - 3 && new_idepth > 0)) continue;
    const float refColor = lpc_color[i];
    float hit[3]; interp33(dINewl, Ku, Kv, wl, hit);
    if (!std::isfinite(hit[0])) continue;
    const float residual = hit[0] - (float)(affLL[0] * refColor + affLL[1]);
    const float hw = std::fabs(residual) < kHuberTH ? 1 : kHuberTH / std::fabs(residual);
    if (std::fabs(residual) > cutoffTH) {
      E += maxEnergy; Ed += (double)maxEnergy; numTermsInE++; numSaturated++;
    } else {
      E += hw * residual * residual * (2 - hw);
      Ed += (double)(hw * residual * residual * (2 - hw));
      numTermsInE++;
      T.pbuf[0][numTermsInWarped] = new_idepth; T.pbuf[1][numTermsInWarped] = u; T.pbuf[2][numTermsInWarped] = v;
      T.pbuf[3][numTermsInWarped] = hit[1]; T.pbuf[4][numTermsInWarped] = hit[2];
      T.pbuf[5][numTermsInWarped] = residual; T.pbuf[6][numTermsInWarped] = hw; T.pbuf[7][numTermsInWarped] = lpc_color[i];
      numTermsInWarped++;
    }
  }
  while (numTermsInWarped % 4 != 0) {
    for (int b = 0; b < 8; b++) T.pbuf[b][numTermsInWarped] = 0;
    numTermsInWarped++;
  }
  T.pose_n = numTermsInWarped;
  rs[0] = E; rs[1] = numTermsInE; rs[2] = sumSquaredShiftT / (sumSquaredShiftNum + 0.1); rs[3] = 0;
  rs[4] = sumSquaredShiftRT / (sumSquaredShiftNum + 0.1); rs[5] = numSaturated / (float)numTermsInE;
  if (T.res_acc_mode == 1) { rs[0] = Ed; rs[2] = sTd / (sumSquaredShiftNum + 0.1); rs[4] = sRTd / (sumSquaredShiftNum + 0.1); }
}

// Per-point Jacobian row of calcGSSSEPose (:658-678), one lane.
inline void pose_jacobian(float id, float u, float v, float dxr, float dyr, float refColor, float fxl, float fyl, float a, float b0, float J[8]) {
  const float dx = dxr * fxl, dy = dyr * fyl;
  J[0] = id * dx;
  J[1] = id * dy;
  J[2] = 0.0f - (id * ((u * dx) + (v * dy)));
  J[3] = 0.0f - (((u * v) * dx) + (dy * (1.0f + (v * v))));
  J[4] = ((u * v) * dy) + (dx * (1.0f + (u * u)));
  J[5] = (u * dy) - (v * dx);
  J[6] = a * (b0 - refColor);
  J[7] = -1.0f;
}

// calcGSSSEPose  src/scale_optimization/TrackerAndScaler.cpp:640-697
// mode 0: SSE-faithful 4-lane / 3-tier fp32 accumulation (Accumulator9)
// mode 1: same fp32 Jacobian rows and fp32 (J_r*w), but acc += (double)(J_r*w) * (double)J_c in fp64
//         (each product exact in double) — the accumulation the CUDA path implements.
// acc45_out (optional): the 45 raw sums (upper triangle row-major, index 8 = residual) before /n and scaling.
void calc_gs_pose(Tracker &T, int lvl, int mode, double aff_a, double aff_b, double H[64], double b[8], double *acc45_out) {
  T.n_gs_evals++;
  const float fxl = T.fx[lvl], fyl = T.fy[lvl];
  const float b0 = (float)T.ref_b;
  double affd[2]; aff_from_to(T.ref_exposure, T.new_exposure, T.ref_a, T.ref_b, aff_a, aff_b, affd);
  const float a = (float)affd[0];
  const int n = T.pose_n;
  double acc[45];
  if (mode == 0) {
    static thread_local TieredAcc<9> A;
    A.initialize();
    for (int i = 0; i < n; i += 4) {
      float J[9][4], w[4];
      for (int l = 0; l < 4; l++) {
        float Jl[8];
        pose_jacobian(T.pbuf[0][i + l], T.pbuf[1][i + l], T.pbuf[2][i + l], T.pbuf[3][i + l], T.pbuf[4][i + l], T.pbuf[7][i + l], fxl, fyl, a, b0, Jl);
        for (int j = 0; j < 8; j++) J[j][l] = Jl[j];
        J[8][l] = T.pbuf[5][i + l];
        w[l] = T.pbuf[6][i + l];
      }
      A.update(J, w);
    }
    float o[45]; A.finish(o);
    for (int e = 0; e < 45; e++) acc[e] = o[e];
  } else {
    for (int e = 0; e < 45; e++) acc[e] = 0;
    for (int i = 0; i < n; i++) {
      float J[9];
      pose_jacobian(T.pbuf[0][i], T.pbuf[1][i], T.pbuf[2][i], T.pbuf[3][i], T.pbuf[4][i], T.pbuf[7][i], fxl, fyl, a, b0, J);
      J[8] = T.pbuf[5][i];
      const float w = T.pbuf[6][i];
      int e = 0;
      for (int r = 0; r < 9; r++) {
        const float Jw = J[r] * w;
        for (int c = r; c < 9; c++, e++) acc[e] += (double)Jw * (double)J[c];
      }
    }
  }
  if (acc45_out) for (int e = 0; e < 45; e++) acc45_out[e] = acc[e];
  // :682-683  H = acc.H.topLeftCorner<8,8>().cast<double>() * (1.0f/n)
  const float invn = 1.0f / n;
  double Hf[81];
  { int e = 0; for (int r = 0; r < 9; r++) for (int c = r; c < 9; c++, e++) Hf[r * 9 + c] = Hf[c * 9 + r] = acc[e]; }
  // mode 0 passes through float H (Mat99f) exactly; mode 1 keeps the fp64 sums.
  for (int r = 0; r < 8; r++) {
    for (int c = 0; c < 8; c++) H[r * 8 + c] = Hf[r * 9 + c] * invn;
    b[r] = Hf[r * 9 + 8] * invn;
  }
  // :685-696 column / row scaling
  const double sc[8] = {kScaleXiRot, kScaleXiRot, kScaleXiRot, kScaleXiTrans, kScaleXiTrans, kScaleXiTrans, kScaleA, kScaleB};
  for (int r = 0; r < 8; r++) for (int c = 0; c < 8; c++) H[r * 8 + c] *= sc[c];
  for (int r = 0; r < 8; r++) for (int c = 0; c < 8; c++) H[r * 8 + c] *= sc[r];
  for (int r = 0; r < 8; r++) b[r] *= sc[r];
}

// ------------------------------------------------------------------------------------------------
// trackNewestCoarse  src/scale_optimization/TrackerAndScaler.cpp:451-638
// ------------------------------------------------------------------------------------------------
int track_newest_coarse(Tracker &T, int mode, SE3 &lastToNew_out, double aff_io[2], int coarsestLvl, const double minResForAbort[5], double lastResiduals[5]) {
  for (int i = 0; i < 5; i++) lastResiduals[i] = NAN;
  T.lastFlow[0] = T.lastFlow[1] = T.lastFlow[2] = 1000;
  const int maxIterations[] = {10, 20, 50, 50, 50};
  const float lambdaExtrapolationLimit = 0.001;
  T.res_acc_mode = mode;
  SE3 refToNew_current = lastToNew_out;
  double aff_cur[2] = {aff_io[0], aff_io[1]};
  bool haveRepeated = false;
  T.trace.clear();

  for (int lvl = coarsestLvl; lvl >= 0; lvl--) {
    double H[64], b[8];
    float levelCutoffRepeat = 1;
    double resOld[6];
    calc_res_pose(T, lvl, refToNew_current, aff_cur[0], aff_cur[1], kCoarseCutoffTH * levelCutoffRepeat, resOld);
    while (resOld[5] > 0.6 && levelCutoffRepeat < 50) {
      levelCutoffRepeat *= 2;
      calc_res_pose(T, lvl, refToNew_current, aff_cur[0], aff_cur[1], kCoarseCutoffTH * levelCutoffRepeat, resOld);
    }
    calc_gs_pose(T, lvl, mode, aff_cur[0], aff_cur[1], H, b, nullptr);
    float lambda = 0.01;
    { TraceRec tr{}; tr.lvl = lvl; tr.iteration = -1; tr.accept = 1; tr.n = T.pose_n; tr.lambda = lambda; tr.e_old = 0; tr.e_new = resOld[0] / resOld[1]; T.trace.push_back(tr); }

    for (int iteration = 0; iteration < maxIterations[lvl]; iteration++) {
      double Hl[64]; std::memcpy(Hl, H, sizeof(Hl));
      for (int i = 0; i < 8; i++) Hl[i * 8 + i] *= (1 + lambda);
      double nb[8]; for (int i = 0; i < 8; i++) nb[i] = -b[i];
      double inc[8];
      ldlt_solve(8, Hl, 8, nb, inc);
      if (T.affModeA < 0 && T.affModeB < 0) {  // fix a, b  (:511-515)
        ldlt_solve(6, Hl, 8, nb, inc); inc[6] = inc[7] = 0;
      }
      if (!(T.affModeA < 0) && T.affModeB < 0) {  // fix b  (:516-520)
        ldlt_solve(7, Hl, 8, nb, inc); inc[7] = 0;
      }
      if (T.affModeA < 0 && !(T.affModeB < 0)) {  // fix a  (:521-534)
        double Hs[64]; std::memcpy(Hs, Hl, sizeof(Hs));
        double bs[8]; std::memcpy(bs, nb, sizeof(bs));
        for (int i = 0; i < 8; i++) Hs[i * 8 + 6] = Hs[i * 8 + 7];
        for (int j = 0; j < 8; j++) Hs[6 * 8 + j] = Hs[7 * 8 + j];
        bs[6] = bs[7];
        double is[8]; ldlt_solve(7, Hs, 8, bs, is);
        for (int i = 0; i < 6; i++) inc[i] = is[i];
        inc[6] = 0; inc[7] = is[6];
      }
      float extrapFac = 1;
      if (lambda < lambdaExtrapolationLimit) extrapFac = std::sqrt(std::sqrt(lambdaExtrapolationLimit / lambda));
      for (int i = 0; i < 8; i++) inc[i] *= extrapFac;
      double incScaled[8];
      for (int i = 0; i < 3; i++) incScaled[i] = inc[i] * kScaleXiRot;
      for (int i = 3; i < 6; i++) incScaled[i] = inc[i] * kScaleXiTrans;
      incScaled[6] = inc[6] * kScaleA; incScaled[7] = inc[7] * kScaleB;
      { double s = 0; for (int i = 0; i < 8; i++) s += incScaled[i]; if (!std::isfinite(s)) for (int i = 0; i < 8; i++) incScaled[i] = 0; }
      SE3 refToNew_new = se3_mul(se3_exp(incScaled), refToNew_current);
      double aff_new[2] = {aff_cur[0] + incScaled[6], aff_cur[1] + incScaled[7]};
      double resNew[6];
      calc_res_pose(T, lvl, refToNew_new, aff_new[0], aff_new[1], kCoarseCutoffTH * levelCutoffRepeat, resNew);
      const bool accept = (resNew[0] / resNew[1]) < (resOld[0] / resOld[1]);
      { TraceRec tr{}; tr.lvl = lvl; tr.iteration = iteration; tr.accept = accept; tr.n = T.pose_n; tr.lambda = lambda;
        tr.e_old = resOld[0] / resOld[1]; tr.e_new = resNew[0] / resNew[1]; for (int i = 0; i < 8; i++) tr.inc[i] = inc[i]; T.trace.push_back(tr); }
      if (accept) {
        calc_gs_pose(T, lvl, mode, aff_new[0], aff_new[1], H, b, nullptr);
        std::memcpy(resOld, resNew, sizeof(resOld));
        aff_cur[0] = aff_new[0]; aff_cur[1] = aff_new[1];
        refToNew_current = refToNew_new;
        lambda *= 0.5;
      } else {
        lambda *= 4;
        if (lambda < lambdaExtrapolationLimit) lambda = lambdaExtrapolationLimit;
      }
      double nrm = 0; for (int i = 0; i < 8; i++) nrm += inc[i] * inc[i];
      nrm = std::sqrt(nrm);
      if (!(nrm > 1e-3)) break;
    }
    lastResiduals[lvl] = sqrtf((float)(resOld[0] / resOld[1]));
    T.lastFlow[0] = resOld[2]; T.lastFlow[1] = resOld[3]; T.lastFlow[2] = resOld[4];
    if (lastResiduals[lvl] > 1.5 * minResForAbort[lvl]) return 0;
    if (levelCutoffRepeat > 1 && !haveRepeated) { lvl++; haveRepeated = true; }
  }
  lastToNew_out = refToNew_current;
  aff_io[0] = aff_cur[0]; aff_io[1] = aff_cur[1];
  if ((T.affModeA != 0 && (fabsf((float)aff_io[0]) > 1.2)) || (T.affModeB != 0 && (fabsf((float)aff_io[1]) > 200))) return 0;
  double rel[2]; aff_from_to(T.ref_exposure, T.new_exposure, T.ref_a, T.ref_b, aff_io[0], aff_io[1], rel);
  const float relA = (float)rel[0], relB = (float)rel[1];
  if ((T.affModeA == 0 && (fabsf(logf(relA)) > 1.5)) || (T.affModeB == 0 && (fabsf(relB) > 200))) return 0;
  if (T.affModeA < 0) aff_io[0] = 0;
  if (T.affModeB < 0) aff_io[1] = 0;
  return 1;
}

// ------------------------------------------------------------------------------------------------
// calcResScale  src/scale_optimization/TrackerAndScaler.cpp:1007-1172
// ------------------------------------------------------------------------------------------------
void calc_res_scale(Tracker &T, int lvl, float scale, float cutoffTH, double rs[6]) {
  T.n_res_evals++;
  float E = 0;
  double Ed = 0, sTd = 0, sRTd = 0;  // res_acc_mode 1
  int numTermsInE = 0, numTermsInWarped = 0, numSaturated = 0;
  const int wl = T.w[lvl], hl = T.h[lvl];
  const float *dINewl = T.dIp_right[lvl];
  const float fx1l = T.fx1[lvl], fy1l = T.fy1[lvl], cx1l = T.cx1[lvl], cy1l = T.cy1[lvl];
  double Rd[9]; quat_to_R(T.tfm_f1_f0.q, Rd);
  float Rf[9]; for (int i = 0; i < 9; i++) Rf[i] = (float)Rd[i];
  float M[9]; mat33f_mul(Rf, T.Ki[lvl], M);  // rot_f1_f0_K0_i
  const float tsl[3] = {(float)T.tfm_f1_f0.t[0], (float)T.tfm_f1_f0.t[1], (float)T.tfm_f1_f0.t[2]};
  const float *Ki = T.Ki[lvl];
  float sumSquaredShiftT = 0, sumSquaredShiftRT = 0, sumSquaredShiftNum = 0;
  const float maxEnergy = 2 * kHuberTH * cutoffTH - kHuberTH * kHuberTH;
  const int nl = T.pc_n[lvl];
  const float *lpc_u = T.pc_u[lvl].data(), *lpc_v = T.pc_v[lvl].data();
  const float *lpc_idepth = T.pc_idepth[lvl].data(), *lpc_color = T.pc_color[lvl].data();
  for (int b = 0; b < 8; b++) if ((int)T.sbuf[b].size() < nl + 4) T.sbuf[b].resize(nl + 4);

  for (int i = 0; i < nl; i++) {
    const float id = lpc_idepth[i], x = lpc_u[i], y = lpc_v[i];
    // pt = (scale * M) * (x,y,1) + tsl * id   (Eigen 3.3.7: scalar stays inside the lhs coefficients)
    float pt[3], rx[3];
    for (int r = 0; r < 3; r++) {
      pt[r] = dot3f(scale * M[r * 3], x, scale * M[r * 3 + 1], y, scale * M[r * 3 + 2], 1.0f) + tsl[r] * id;
      rx[r] = dot3f(M[r * 3], x, M[r * 3 + 1], y, M[r * 3 + 2], 1.0f) / id;  // :1068
    }
    const float u = pt[0] / pt[2], v = pt[1] / pt[2];
    const float Ku = fx1l * u + cx1l, Kv = fy1l * v + cy1l;
    const float new_idepth = id / pt[2];

    if (lvl == 0 && i % 32 == 0) {  // :1070-1100 (computed, unused by the caller)
      float ptT[3], ptT2[3], pt3[3];
      for (int r = 0; r < 3; r++) {
        const float kx = dot3f(scale * Ki[r * 3], x, scale * Ki[r * 3 + 1], y, scale * Ki[r * 3 + 2], 1.0f);
        const float mx = dot3f(scale * M[r * 3], x, scale * M[r * 3 + 1], y, scale * M[r * 3 + 2], 1.0f);
        ptT[r] = kx + tsl[r] * id; ptT2[r] = kx - tsl[r] * id; pt3[r] = mx - tsl[r] * id;
      }
      const float KuT = fx1l * (ptT[0] / ptT[2]) + cx1l, KvT = fy1l * (ptT[1] / ptT[2]) + cy1l;
      const float KuT2 = fx1l * (ptT2[0] / ptT2[2]) + cx1l, KvT2 = fy1l * (ptT2[1] / ptT2[2]) + cy1l;
      const float Ku3 = fx1l * (pt3[0] / pt3[2]) + cx1l, Kv3 = fy1l * (pt3[1] / pt3[2]) + cy1l;
      const float sT1 = (KuT - x) * (KuT - x) + (KvT - y) * (KvT - y), sT2 = (KuT2 - x) * (KuT2 - x) + (KvT2 - y) * (KvT2 - y);
      const float sRT1 = (Ku - x) * (Ku - x) + (Kv - y) * (Kv - y), sRT2 = (Ku3 - x) * (Ku3 - x) + (Kv3 - y) * (Kv3 - y);
      sumSquaredShiftT += sT1;
      sumSquaredShiftT += sT2;
      sumSquaredShiftRT += sRT1;
      sumSquaredShiftRT += sRT2;
      sTd += (double)sT1; sTd += (double)sT2; sRTd += (double)sRT1; sRTd += (double)sRT2;
      sumSquaredShiftNum += 2;
    }

    if (!(Ku > 2 && Kv > 2 && Ku < wl - 3 && Kv < hl - 3 && new_idepth > 0)) continue;
    const float refColor = lpc_color[i];
    float hit[3]; interp33(dINewl, Ku, Kv, wl, hit);
    if (!std::isfinite(hit[0])) continue;
    const float residual = hit[0] - refColor;  // :1109 no affine
    const float hw = std::fabs(residual) < kHuberTH ? 1 : kHuberTH / std::fabs(residual);
    if (std::fabs(residual) > cutoffTH) {
      E += maxEnergy; Ed += (double)maxEnergy; numTermsInE++; numSaturated++;
    } else {
      E += hw * residual * residual * (2 - hw);
      Ed += (double)(hw * residual * residual * (2 - hw));
      numTermsInE++;
      T.sbuf[0][numTermsInWarped] = rx[0]; T.sbuf[1][numTermsInWarped] = rx[1]; T.sbuf[2][numTermsInWarped] = rx[2];
      T.sbuf[3][numTermsInWarped] = hit[1]; T.sbuf[4][numTermsInWarped] = hit[2];
      T.sbuf[5][numTermsInWarped] = residual; T.sbuf[6][numTermsInWarped] = hw; T.sbuf[7][numTermsInWarped] = lpc_color[i];
      numTermsInWarped++;
    }
  }
  while (numTermsInWarped % 4 != 0) {
    for (int b = 0; b < 8; b++) T.sbuf[b][numTermsInWarped] = 0;
    numTermsInWarped++;
  }
  T.scale_n = numTermsInWarped;
  rs[0] = E; rs[1] = numTermsInE; rs[2] = sumSquaredShiftT / (sumSquaredShiftNum + 0.1); rs[3] = 0;
  rs[4] = sumSquaredShiftRT / (sumSquaredShiftNum + 0.1); rs[5] = numSaturated / (float)numTermsInE;
  if (T.res_acc_mode == 1) { rs[0] = Ed; rs[2] = sTd / (sumSquaredShiftNum + 0.1); rs[4] = sRTd / (sumSquaredShiftNum + 0.1); }
}

// Per-point scale Jacobian (calcGSSSEScale :983-997), one lane.
inline float scale_jacobian(float rx1, float rx2, float rx3, float dxr, float dyr, float fx1l, float fy1l, float s, float tx, float ty, float tz) {
  const float dxfx = dxr * fx1l, dyfy = dyr * fy1l;
  const float deno_sqrt = (s * rx3) + tz;
  const float deno = 1.0f / (deno_sqrt * deno_sqrt);
  const float xno = (rx1 * tz) - (rx3 * tx);
  const float yno = (rx2 * tz) - (rx3 * ty);
  return (dxfx * (deno * xno)) + (dyfy * (deno * yno));
}

// calcGSSSEScale  src/scale_optimization/TrackerAndScaler.cpp:966-1005 ; acc3 = (JwJ, Jwr, rwr) raw sums
void calc_gs_scale(Tracker &T, int lvl, int mode, float scale, float *H_out, float *b_out, double *acc3_out) {
  T.n_gs_evals++;
  const float fx1l = T.fx1[lvl], fy1l = T.fy1[lvl];
  const float tx = (float)T.tfm_f1_f0.t[0], ty = (float)T.tfm_f1_f0.t[1], tz = (float)T.tfm_f1_f0.t[2];
  const int n = T.scale_n;
  double acc[3];
  if (mode == 0) {
    static thread_local TieredAcc<2> A;
    A.initialize();
    for (int i = 0; i < n; i += 4) {
      float J[2][4], w[4];
      for (int l = 0; l < 4; l++) {
        J[0][l] = scale_jacobian(T.sbuf[0][i + l], T.sbuf[1][i + l], T.sbuf[2][i + l], T.sbuf[3][i + l], T.sbuf[4][i + l], fx1l, fy1l, scale, tx, ty, tz);
        J[1][l] = T.sbuf[5][i + l];
        w[l] = T.sbuf[6][i + l];
      }
      A.update(J, w);
    }
    float o[3]; A.finish(o);
    for (int e = 0; e < 3; e++) acc[e] = o[e];
  } else {
    acc[0] = acc[1] = acc[2] = 0;
    for (int i = 0; i < n; i++) {
      const float J = scale_jacobian(T.sbuf[0][i], T.sbuf[1][i], T.sbuf[2][i], T.sbuf[3][i], T.sbuf[4][i], fx1l, fy1l, scale, tx, ty, tz);
      const float r = T.sbuf[5][i], w = T.sbuf[6][i];
      const float Jw = J * w, rw = r * w;
      acc[0] += (double)Jw * (double)J; acc[1] += (double)Jw * (double)r; acc[2] += (double)rw * (double)r;
    }
  }
  if (acc3_out) { acc3_out[0] = acc[0]; acc3_out[1] = acc[1]; acc3_out[2] = acc[2]; }
  // :1003-1004  hessian_ is Mat22f: the sums pass through float
  *H_out = (float)acc[0] * (1.0f / n);
  *b_out = (float)acc[1] * (1.0f / n);
}

// optimizeScale  src/scale_optimization/TrackerAndScaler.cpp:854-964
float optimize_scale(Tracker &T, int mode, float &scale, int coarsestLvl) {
  float last_residuals[5]; for (int i = 0; i < 5; i++) last_residuals[i] = NAN;
  const int maxIterations[] = {10, 20, 50, 50, 50};
  const float lambdaExtrapolationLimit = 0.001;
  T.res_acc_mode = mode;
  float scale_current = scale;
  bool haveRepeated = false;
  T.trace.clear();
  for (int lvl = coarsestLvl; lvl >= 0; lvl--) {
    float H, b;
    float levelCutoffRepeat = 1;
    double resOld[6];
    calc_res_scale(T, lvl, scale_current, kCoarseCutoffTH * levelCutoffRepeat, resOld);
    while (resOld[5] > 0.6 && levelCutoffRepeat < 50) {
      levelCutoffRepeat *= 2;
      calc_res_scale(T, lvl, scale_current, kCoarseCutoffTH * levelCutoffRepeat, resOld);
    }
    calc_gs_scale(T, lvl, mode, scale_current, &H, &b, nullptr);
    float lambda = 0.01;
    { TraceRec tr{}; tr.lvl = lvl; tr.iteration = -1; tr.accept = 1; tr.n = T.scale_n; tr.lambda = lambda; tr.e_new = resOld[0] / resOld[1]; tr.inc[1] = scale_current; T.trace.push_back(tr); }
    for (int iteration = 0; iteration < maxIterations[lvl]; iteration++) {
      float Hl = H;
      Hl *= (1 + lambda);
      float inc = -b / Hl;
      float extrapFac = 1;
      if (lambda < lambdaExtrapolationLimit) extrapFac = std::sqrt(std::sqrt(lambdaExtrapolationLimit / lambda));
      inc *= extrapFac;
      if (!std::isfinite(inc) || std::fabs(inc) > scale_current) inc = 0.0;
      const float scale_new = scale_current + inc;
      double resNew[6];
      calc_res_scale(T, lvl, scale_new, kCoarseCutoffTH * levelCutoffRepeat, resNew);
      const bool accept = (resNew[0] / resNew[1]) < (resOld[0] / resOld[1]);
      { TraceRec tr{}; tr.lvl = lvl; tr.iteration = iteration; tr.accept = accept; tr.n = T.scale_n; tr.lambda = lambda;
        tr.e_old = resOld[0] / resOld[1]; tr.e_new = resNew[0] / resNew[1]; tr.inc[0] = inc; tr.inc[1] = scale_new; T.trace.push_back(tr); }
      if (accept) {
        calc_gs_scale(T, lvl, mode, scale_new, &H, &b, nullptr);
        std::memcpy(resOld, resNew, sizeof(resOld));
        scale_current = scale_new;
        lambda *= 0.5;
      } else {
        lambda *= 4;
        if (lambda < lambdaExtrapolationLimit) lambda = lambdaExtrapolationLimit;
      }
      if (!(inc > 1e-3)) break;
    }
    last_residuals[lvl] = sqrtf((float)(resOld[0] / resOld[1]));
    if (levelCutoffRepeat > 1 && !haveRepeated) { lvl++; haveRepeated = true; }
  }
  scale = scale_current;
  return last_residuals[0];
}

// ------------------------------------------------------------------------------------------------
// PoseEstimator (loop-closure photometric alignment)  src/loop_closure/pose_estimation/PoseEstimator.cpp
// The third copy of the 8-DoF Gauss-Newton: 3-D reference points with one colour per pyramid level.
// ------------------------------------------------------------------------------------------------
struct PoseEst {
  int levels = 0;
  int w[kMaxLevels], h[kMaxLevels];
  float fx[kMaxLevels], fy[kMaxLevels], cx[kMaxLevels], cy[kMaxLevels];
  std::vector<double> pts;                  // n x 3 (Eigen::Vector3d first of each pair)
  std::vector<float> colors[kMaxLevels];    // pair.second[lvl]
  int n = 0;
  std::vector<float> buf[8];
  int buf_n = 0;
  const float *dIp_new[kMaxLevels];
  float new_exposure = 1.f, ref_exposure = 1.f;
  double ref_a = 0, ref_b = 0;  // ref_aff_g2l_ = AffLight()  (:316)
  int affModeA = 0, affModeB = 0;
  int res_acc_mode = 0;
  std::vector<TraceRec> trace;
};

// PoseEstimator::makeK  :66-83
void pe_make_K(PoseEst &P, int w0, int h0, const float cam[4]) {
  P.w[0] = w0; P.h[0] = h0; P.fx[0] = cam[0]; P.fy[0] = cam[1]; P.cx[0] = cam[2]; P.cy[0] = cam[3];
  for (int l = 1; l < P.levels; l++) {
    P.w[l] = P.w[0] >> l; P.h[l] = P.h[0] >> l;
    P.fx[l] = P.fx[l - 1] * 0.5; P.fy[l] = P.fy[l - 1] * 0.5;
    P.cx[l] = (P.cx[0] + 0.5) / ((int)1 << l) - 0.5;
    P.cy[l] = (P.cy[0] + 0.5) / ((int)1 << l) - 0.5;
  }
}

// PoseEstimator::calcRes  :141-296
void pe_calc_res(PoseEst &P, int lvl, const SE3 &refToNew, double aff_a, double aff_b, float cutoffTH, double rs[6]) {
  float E = 0;
  double Ed = 0, sTd = 0, sRTd = 0;
  int numTermsInE = 0, numTermsInWarped = 0, numSaturated = 0;
  const int wl = P.w[lvl], hl = P.h[lvl];
  const float *dINewl = P.dIp_new[lvl];
  const float fxl = P.fx[lvl], fyl = P.fy[lvl], cxl = P.cx[lvl], cyl = P.cy[lvl];
  double Rd[9]; quat_to_R(refToNew.q, Rd);
  float R[9]; for (int i = 0; i < 9; i++) R[i] = (float)Rd[i];
  const float t[3] = {(float)refToNew.t[0], (float)refToNew.t[1], (float)refToNew.t[2]};
  double affd[2]; aff_from_to(P.ref_exposure, P.new_exposure, P.ref_a, P.ref_b, aff_a, aff_b, affd);
  const float affLL[2] = {(float)affd[0], (float)affd[1]};
  float sumSquaredShiftT = 0, sumSquaredShiftRT = 0, sumSquaredShiftNum = 0;
  const float maxEnergy = 2 * kHuberTH * cutoffTH - kHuberTH * kHuberTH;
  for (int b = 0; b < 8; b++) if ((int)P.buf[b].size() < P.n + 4) P.buf[b].resize(P.n + 4);
  for (int i = 0; i < P.n; i++) {
    const float x = (float)P.pts[3 * i], y = (float)P.pts[3 * i + 1], z = (float)P.pts[3 * i + 2];
    const float u0 = x / z, v0 = y / z;
    const float Ku0 = fxl * u0 + cxl, Kv0 = fyl * v0 + cyl;
    float pt[3];
    for (int r = 0; r < 3; r++) pt[r] = dot3f(R[r * 3], x, R[r * 3 + 1], y, R[r * 3 + 2], z) + t[r];
    const float u = pt[0] / pt[2], v = pt[1] / pt[2];
    const float Ku = fxl * u + cxl, Kv = fyl * v + cyl;
    const float new_idepth = 1 / pt[2];
    if (lvl == 0 && i % 32 == 0) {  // :191-226
      const float ptT[3] = {x + t[0], y + t[1], 1 + t[2]}, ptT2[3] = {x - t[0], y - t[1], 1 - t[2]};
      float pt3[3];
      for (int r = 0; r < 3; r++) pt3[r] = dot3f(R[r * 3], x, R[r * 3 + 1], y, R[r * 3 + 2], 1.0f) - t[r];
      const float KuT = fxl * (ptT[0] / ptT[2]) + cxl, KvT = fyl * (ptT[1] / ptT[2]) + cyl;
      const float KuT2 = fxl * (ptT2[0] / ptT2[2]) + cxl, KvT2 = fyl * (ptT2[1] / ptT2[2]) + cyl;
      const float Ku3 = fxl * (pt3[0] / pt3[2]) + cxl, Kv3 = fyl * (pt3[1] / pt3[2]) + cyl;
      const float sT1 = (KuT - Ku0) * (KuT - Ku0) + (KvT - Kv0) * (KvT - Kv0), sT2 = (KuT2 - Ku0) * (KuT2 - Ku0) + (KvT2 - Kv0) * (KvT2 - Kv0);
      const float sRT1 = (Ku - Ku0) * (Ku - Ku0) + (Kv - Kv0) * (Kv - Kv0), sRT2 = (Ku3 - Ku0) * (Ku3 - Ku0) + (Kv3 - Kv0) * (Kv3 - Kv0);
      sumSquaredShiftT += sT1; sumSquaredShiftT += sT2; sumSquaredShiftRT += sRT1; sumSquaredShiftRT += sRT2;
      sTd += (double)sT1; sTd += (double)sT2; sRTd += (double)sRT1; sRTd += (double)sRT2;
      sumSquaredShiftNum += 2;
    }
    if (!(Ku > 2 && Kv > 2 && Ku < wl - 3 && Kv < hl - 3 && new_idepth > 0)) continue;
    const float refColor = P.colors[lvl][i];
    float hit[3]; interp33(dINewl, Ku, Kv, wl, hit);
    if (!std::isfinite(hit[0])) continue;
    const float residual = hit[0] - (float)(affLL[0] * refColor + affLL[1]);
    const float hw = std::fabs(residual) < kHuberTH ? 1 : kHuberTH / std::fabs(residual);
    if (std::fabs(residual) > cutoffTH) {
      E += maxEnergy; Ed += (double)maxEnergy; numTermsInE++; numSaturated++;
    } else {
      E += hw * residual * residual * (2 - hw);
      Ed += (double)(hw * residual * residual * (2 - hw));
      numTermsInE++;
      P.buf[0][numTermsInWarped] = new_idepth; P.buf[1][numTermsInWarped] = u; P.buf[2][numTermsInWarped] = v;
      P.buf[3][numTermsInWarped] = hit[1]; P.buf[4][numTermsInWarped] = hit[2];
      P.buf[5][numTermsInWarped] = residual; P.buf[6][numTermsInWarped] = hw; P.buf[7][numTermsInWarped] = refColor;
      numTermsInWarped++;
    }
  }
  while (numTermsInWarped % 4 != 0) {
    for (int b = 0; b < 8; b++) P.buf[b][numTermsInWarped] = 0;
    numTermsInWarped++;
  }
  P.buf_n = numTermsInWarped;
  rs[0] = E; rs[1] = numTermsInE; rs[2] = sumSquaredShiftT / (sumSquaredShiftNum + 0.1); rs[3] = 0;
  rs[4] = sumSquaredShiftRT / (sumSquaredShiftNum + 0.1); rs[5] = numSaturated / (float)numTermsInE;
  if (P.res_acc_mode == 1) { rs[0] = Ed; rs[2] = sTd / (sumSquaredShiftNum + 0.1); rs[4] = sRTd / (sumSquaredShiftNum + 0.1); }
}

// PoseEstimator::calcGSSSE  :84-139  (identical to calcGSSSEPose with b0 = ref_aff_g2l_.b)
void pe_calc_gs(PoseEst &P, int lvl, int mode, double aff_a, double aff_b, double H[64], double b[8], double *acc45_out) {
  const float fxl = P.fx[lvl], fyl = P.fy[lvl];
  const float b0 = (float)P.ref_b;
  double affd[2]; aff_from_to(P.ref_exposure, P.new_exposure, P.ref_a, P.ref_b, aff_a, aff_b, affd);
  const float a = (float)affd[0];
  const int n = P.buf_n;
  double acc[45];
  if (mode == 0) {
    static thread_local TieredAcc<9> A;
    A.initialize();
    for (int i = 0; i < n; i += 4) {
      float J[9][4], w[4];
      for (int l = 0; l < 4; l++) {
        float Jl[8];
        pose_jacobian(P.buf[0][i + l], P.buf[1][i + l], P.buf[2][i + l], P.buf[3][i + l], P.buf[4][i + l], P.buf[7][i + l], fxl, fyl, a, b0, Jl);
        for (int j = 0; j < 8; j++) J[j][l] = Jl[j];
        J[8][l] = P.buf[5][i + l];
        w[l] = P.buf[6][i + l];
      }
      A.update(J, w);
    }
    float o[45]; A.finish(o);
    for (int e = 0; e < 45; e++) acc[e] = o[e];
  } else {
    for (int e = 0; e < 45; e++) acc[e] = 0;
    for (int i = 0; i < n; i++) {
      float J[9];
      pose_jacobian(P.buf[0][i], P.buf[1][i], P.buf[2][i], P.buf[3][i], P.buf[4][i], P.buf[7][i], fxl, fyl, a, b0, J);
      J[8] = P.buf[5][i];
      const float w = P.buf[6][i];
      int e = 0;
      for (int r = 0; r < 9; r++) {
        const float Jw = J[r] * w;
        for (int c = r; c < 9; c++, e++) acc[e] += (double)Jw * (double)J[c];
      }
    }
  }
  if (acc45_out) for (int e = 0; e < 45; e++) acc45_out[e] = acc[e];
  const float invn = 1.0f / n;
  double Hf[81];
  { int e = 0; for (int r = 0; r < 9; r++) for (int c = r; c < 9; c++, e++) Hf[r * 9 + c] = Hf[c * 9 + r] = acc[e]; }
  for (int r = 0; r < 8; r++) {
    for (int c = 0; c < 8; c++) H[r * 8 + c] = Hf[r * 9 + c] * invn;
    b[r] = Hf[r * 9 + 8] * invn;
  }
  const double sc[8] = {kScaleXiRot, kScaleXiRot, kScaleXiRot, kScaleXiTrans, kScaleXiTrans, kScaleXiTrans, kScaleA, kScaleB};
  for (int r = 0; r < 8; r++) for (int c = 0; c < 8; c++) H[r * 8 + c] *= sc[c];
  for (int r = 0; r < 8; r++) for (int c = 0; c < 8; c++) H[r * 8 + c] *= sc[r];
  for (int r = 0; r < 8; r++) b[r] *= sc[r];
}

// SE3(Matrix3d, Vector3d) as used at :321-322 (quaternion from the rotation block) and SE3::matrix() (:463)
SE3 se3_from_matrix4(const double *m) {
  SE3 s;
  double q[4];
  const double tr = m[0] + m[5] + m[10];
  if (tr > 0) {
    double t = std::sqrt(tr + 1.0); q[3] = 0.5 * t; t = 0.5 / t;
    q[0] = (m[9] - m[6]) * t; q[1] = (m[2] - m[8]) * t; q[2] = (m[4] - m[1]) * t;
  } else {
    int i = 0; if (m[5] > m[0]) i = 1; if (m[10] > m[i * 4 + i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    double t = std::sqrt(m[i * 4 + i] - m[j * 4 + j] - m[k * 4 + k] + 1.0);
    q[i] = 0.5 * t; t = 0.5 / t;
    q[3] = (m[k * 4 + j] - m[j * 4 + k]) * t; q[j] = (m[j * 4 + i] + m[i * 4 + j]) * t; q[k] = (m[k * 4 + i] + m[i * 4 + k]) * t;
  }
  std::memcpy(s.q, q, sizeof(q));
  quat_normalize(s.q);
  s.t[0] = m[3]; s.t[1] = m[7]; s.t[2] = m[11];
  return s;
}
void se3_to_matrix4(const SE3 &s, double *m) {
  double R[9]; quat_to_R(s.q, R);
  for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) m[r * 4 + c] = R[r * 3 + c]; m[r * 4 + 3] = s.t[r]; }
  m[12] = m[13] = m[14] = 0; m[15] = 1;
}

// PoseEstimator::estimate  :298-506.  Returns aff_good && low_res && enough_inlier; outputs pose_error, inlier_percent.
int pe_estimate(PoseEst &P, int mode, double T_io[16], int coarsest_lvl, float *pose_error, int *inlier_percent_out) {
  const int maxIterations[] = {10, 20, 50, 50, 50};
  const float lambdaExtrapolationLimit = 0.001;
  P.res_acc_mode = mode;
  P.trace.clear();
  int lastInners[kMaxLevels] = {0};
  double lastResiduals[5]; for (int i = 0; i < 5; i++) lastResiduals[i] = NAN;
  double aff_cur[2] = {0, 0};
  SE3 refToNew_current = se3_from_matrix4(T_io);
  bool haveRepeated = false;
  for (int lvl = coarsest_lvl; lvl >= 0; lvl--) {
    double H[64], b[8];
    float levelCutoffRepeat = 1;
    double resOld[6];
    pe_calc_res(P, lvl, refToNew_current, aff_cur[0], aff_cur[1], kCoarseCutoffTH * levelCutoffRepeat, resOld);
    while (resOld[5] > 0.6 && levelCutoffRepeat < 50) {
      levelCutoffRepeat *= 2;
      pe_calc_res(P, lvl, refToNew_current, aff_cur[0], aff_cur[1], kCoarseCutoffTH * levelCutoffRepeat, resOld);
    }
    pe_calc_gs(P, lvl, mode, aff_cur[0], aff_cur[1], H, b, nullptr);
    float lambda = 0.01;
    { TraceRec tr{}; tr.lvl = lvl; tr.iteration = -1; tr.accept = 1; tr.n = P.buf_n; tr.lambda = lambda; tr.e_new = resOld[0] / resOld[1]; P.trace.push_back(tr); }
    for (int iteration = 0; iteration < maxIterations[lvl]; iteration++) {
      double Hl[64]; std::memcpy(Hl, H, sizeof(Hl));
      for (int i = 0; i < 8; i++) Hl[i * 8 + i] *= (1 + lambda);
      double nb[8]; for (int i = 0; i < 8; i++) nb[i] = -b[i];
      double inc[8];
      ldlt_solve(8, Hl, 8, nb, inc);
      if (P.affModeA < 0 && P.affModeB < 0) { ldlt_solve(6, Hl, 8, nb, inc); inc[6] = inc[7] = 0; }
      if (!(P.affModeA < 0) && P.affModeB < 0) { ldlt_solve(7, Hl, 8, nb, inc); inc[7] = 0; }
      if (P.affModeA < 0 && !(P.affModeB < 0)) {
        double Hs[64]; std::memcpy(Hs, Hl, sizeof(Hs));
        double bs[8]; std::memcpy(bs, nb, sizeof(bs));
        for (int i = 0; i < 8; i++) Hs[i * 8 + 6] = Hs[i * 8 + 7];
        for (int j = 0; j < 8; j++) Hs[6 * 8 + j] = Hs[7 * 8 + j];
        bs[6] = bs[7];
        double is[8]; ldlt_solve(7, Hs, 8, bs, is);
        for (int i = 0; i < 6; i++) inc[i] = is[i];
        inc[6] = 0; inc[7] = is[6];
      }
      float extrapFac = 1;
      if (lambda < lambdaExtrapolationLimit) extrapFac = std::sqrt(std::sqrt(lambdaExtrapolationLimit / lambda));
      for (int i = 0; i < 8; i++) inc[i] *= extrapFac;
      double incScaled[8];
      for (int i = 0; i < 3; i++) incScaled[i] = inc[i] * kScaleXiRot;
      for (int i = 3; i < 6; i++) incScaled[i] = inc[i] * kScaleXiTrans;
      incScaled[6] = inc[6] * kScaleA; incScaled[7] = inc[7] * kScaleB;
      { double s = 0; for (int i = 0; i < 8; i++) s += incScaled[i]; if (!std::isfinite(s)) for (int i = 0; i < 8; i++) incScaled[i] = 0; }
      SE3 refToNew_new = se3_mul(se3_exp(incScaled), refToNew_current);
      double aff_new[2] = {aff_cur[0] + incScaled[6], aff_cur[1] + incScaled[7]};
      double resNew[6];
      pe_calc_res(P, lvl, refToNew_new, aff_new[0], aff_new[1], kCoarseCutoffTH * levelCutoffRepeat, resNew);
      const bool accept = (resNew[0] / resNew[1]) < (resOld[0] / resOld[1]);
      { TraceRec tr{}; tr.lvl = lvl; tr.iteration = iteration; tr.accept = accept; tr.n = P.buf_n; tr.lambda = lambda;
        tr.e_old = resOld[0] / resOld[1]; tr.e_new = resNew[0] / resNew[1]; for (int i = 0; i < 8; i++) tr.inc[i] = inc[i]; P.trace.push_back(tr); }
      if (accept) {
        pe_calc_gs(P, lvl, mode, aff_new[0], aff_new[1], H, b, nullptr);
        std::memcpy(resOld, resNew, sizeof(resOld));
        aff_cur[0] = aff_new[0]; aff_cur[1] = aff_new[1];
        refToNew_current = refToNew_new;
        lambda *= 0.5;
      } else {
        lambda *= 4;
        if (lambda < lambdaExtrapolationLimit) lambda = lambdaExtrapolationLimit;
      }
      double nrm = 0; for (int i = 0; i < 8; i++) nrm += inc[i] * inc[i];
      if (!(std::sqrt(nrm) > 1e-3)) break;
    }
    lastResiduals[lvl] = sqrtf((float)(resOld[0] / resOld[1]));
    lastInners[lvl] = (int)resOld[1];
    if (levelCutoffRepeat > 1 && !haveRepeated) { lvl++; haveRepeated = true; }
  }
  se3_to_matrix4(refToNew_current, T_io);
  *pose_error = (float)lastResiduals[0];
  bool aff_good = true;
  if ((P.affModeA != 0 && (fabsf((float)aff_cur[0]) > 1.2)) || (P.affModeB != 0 && (fabsf((float)aff_cur[1]) > 200))) aff_good = false;
  double rel[2]; aff_from_to(P.ref_exposure, P.new_exposure, P.ref_a, P.ref_b, aff_cur[0], aff_cur[1], rel);
  const float relA = (float)rel[0], relB = (float)rel[1];
  if ((P.affModeA == 0 && (fabsf(logf(relA)) > 1.5)) || (P.affModeB == 0 && (fabsf(relB) > 200))) aff_good = false;
  const bool low_res = *pose_error < 10.0;                                     // RES_THRES  PoseEstimator.h:25
  const int inlier_percent = 100 * float(lastInners[0]) / P.n;                  // :480
  const bool enough_inlier = inlier_percent > 90;                              // INNER_PERCENT  :26
  *inlier_percent_out = inlier_percent;
  return (aff_good && low_res && enough_inlier) ? 1 : 0;
}

}  // namespace

// =================================================================================================
// C interface (ctypes)
// =================================================================================================
extern "C" {

// ---- FrameHessian::makeImages  deps:dso/src/FullSystem/HessianBlocks.cpp:128-191 -----------------
// dIp_all: 3*sum(w_l*h_l) floats, level l at pixel offset sum_{k<l} w_k*h_k ; absg_all: sum(w_l*h_l).
// B256: CalibHessian::B (non-null == "HCalib!=0 && setting_gammaWeightsPixelSelect==1"), else null.
// First / last row dx, dy, absSquaredGrad are uninitialised in the reference (new[] at :133-134);
// here they are written as 0 so outputs are deterministic — compare only rows [1, h-2].
void orc_make_images(const float *color, int w, int h, int levels, const float *B256, float *dIp_all, float *absg_all) {
  int off[kMaxLevels + 1]; off[0] = 0;
  for (int l = 0; l < levels; l++) off[l + 1] = off[l] + (w >> l) * (h >> l);
  std::memset(dIp_all, 0, sizeof(float) * 3 * off[levels]);
  std::memset(absg_all, 0, sizeof(float) * off[levels]);
  float *dI = dIp_all;
  for (int i = 0; i < w * h; i++) dI[3 * i] = color[i];
  for (int lvl = 0; lvl < levels; lvl++) {
    const int wl = w >> lvl, hl = h >> lvl;
    float *dI_l = dIp_all + 3 * off[lvl];
    float *dabs_l = absg_all + off[lvl];
    if (lvl > 0) {
      const int wlm1 = w >> (lvl - 1);
      const float *dI_lm = dIp_all + 3 * off[lvl - 1];
      for (int y = 0; y < hl; y++)
        for (int x = 0; x < wl; x++)
          dI_l[3 * (x + y * wl)] = 0.25f * (dI_lm[3 * (2 * x + 2 * y * wlm1)] + dI_lm[3 * (2 * x + 1 + 2 * y * wlm1)] +
                                            dI_lm[3 * (2 * x + 2 * y * wlm1 + wlm1)] + dI_lm[3 * (2 * x + 1 + 2 * y * wlm1 + wlm1)]);
    }
    for (int idx = wl; idx < wl * (hl - 1); idx++) {
      float dx = 0.5f * (dI_l[3 * (idx + 1)] - dI_l[3 * (idx - 1)]);
      float dy = 0.5f * (dI_l[3 * (idx + wl)] - dI_l[3 * (idx - wl)]);
      if (!std::isfinite(dx)) dx = 0;
      if (!std::isfinite(dy)) dy = 0;
      dI_l[3 * idx + 1] = dx;
      dI_l[3 * idx + 2] = dy;
      dabs_l[idx] = dx * dx + dy * dy;
      if (B256) {  // getBGradOnly  deps:dso/src/FullSystem/HessianBlocks.h:384-390
        int c = dI_l[3 * idx] + 0.5f;
        if (c < 5) c = 5;
        if (c > 250) c = 250;
        const float gw = B256[c + 1] - B256[c];
        dabs_l[idx] *= gw * gw;
      }
    }
  }
}

// ---- tracker object -----------------------------------------------------------------------------
// K0 = (fx,fy,cx,cy) of camera 0 (HCalib->fxl() etc. are floats), K1 likewise; T_stereo row-major 4x4
// (tfm_vec, TrackerAndScaler.cpp:82-86: SE3(Matrix4d) -> quaternion from the rotation block).
void *orc_tracker_create(int w, int h, int levels, const float K0[4], const float K1[4], const double T_stereo[16]) {
  Tracker *T = new Tracker();
  T->levels = levels;
  make_K(*T, w, h, K0[0], K0[1], K0[2], K0[3]);
  make_K1(*T, K1[0], K1[1], K1[2], K1[3]);
  // rotation matrix -> quaternion (Eigen's Quaternion(Matrix3) — Shepperd's method); identity in all shipped calibrations.
  const double *m = T_stereo;
  double q[4];
  const double tr = m[0] + m[5] + m[10];
  if (tr > 0) {
    double t = std::sqrt(tr + 1.0); q[3] = 0.5 * t; t = 0.5 / t;
    q[0] = (m[9] - m[6]) * t; q[1] = (m[2] - m[8]) * t; q[2] = (m[4] - m[1]) * t;
  } else {
    int i = 0; if (m[5] > m[0]) i = 1; if (m[10] > m[i * 4 + i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    double t = std::sqrt(m[i * 4 + i] - m[j * 4 + j] - m[k * 4 + k] + 1.0);
    q[i] = 0.5 * t; t = 0.5 / t;
    q[3] = (m[k * 4 + j] - m[j * 4 + k]) * t; q[j] = (m[j * 4 + i] + m[i * 4 + j]) * t; q[k] = (m[k * 4 + i] + m[i * 4 + k]) * t;
  }
  std::memcpy(T->tfm_f1_f0.q, q, sizeof(q));
  quat_normalize(T->tfm_f1_f0.q);
  T->tfm_f1_f0.t[0] = m[3]; T->tfm_f1_f0.t[1] = m[7]; T->tfm_f1_f0.t[2] = m[11];
  for (int l = 0; l < kMaxLevels; l++) { T->pc_n[l] = 0; T->dIp_new[l] = nullptr; T->dIp_right[l] = nullptr; }
  return T;
}
void orc_tracker_destroy(void *p) { delete (Tracker *)p; }

void orc_tracker_get_K(void *p, int lvl, float out[4 + 9 + 4]) {
  Tracker &T = *(Tracker *)p;
  out[0] = T.fx[lvl]; out[1] = T.fy[lvl]; out[2] = T.cx[lvl]; out[3] = T.cy[lvl];
  for (int i = 0; i < 9; i++) out[4 + i] = T.Ki[lvl][i];
  out[13] = T.fx1[lvl]; out[14] = T.fy1[lvl]; out[15] = T.cx1[lvl]; out[16] = T.cy1[lvl];
}

void orc_tracker_set_res_acc_mode(void *p, int mode) { ((Tracker *)p)->res_acc_mode = mode; }
void orc_tracker_set_aff_mode(void *p, int modeA, int modeB) { ((Tracker *)p)->affModeA = modeA; ((Tracker *)p)->affModeB = modeB; }

// template of one level (pc_u, pc_v, pc_idepth, pc_color; copied)
void orc_tracker_set_ref_level(void *p, int lvl, int n, const float *u, const float *v, const float *idepth, const float *color) {
  Tracker &T = *(Tracker *)p;
  T.pc_u[lvl].assign(u, u + n); T.pc_v[lvl].assign(v, v + n);
  T.pc_idepth[lvl].assign(idepth, idepth + n); T.pc_color[lvl].assign(color, color + n);
  T.pc_n[lvl] = n;
}
int orc_tracker_get_ref_level(void *p, int lvl, float *u, float *v, float *idepth, float *color) {
  Tracker &T = *(Tracker *)p;
  const int n = T.pc_n[lvl];
  if (u) {
    std::memcpy(u, T.pc_u[lvl].data(), n * 4); std::memcpy(v, T.pc_v[lvl].data(), n * 4);
    std::memcpy(idepth, T.pc_idepth[lvl].data(), n * 4); std::memcpy(color, T.pc_color[lvl].data(), n * 4);
  }
  return n;
}
// lastRef->ab_exposure, lastRef_aff_g2l  (setCoarseTrackingRef :317-327)
void orc_tracker_set_ref_aff(void *p, float exposure, double a, double b) {
  Tracker &T = *(Tracker *)p; T.ref_exposure = exposure; T.ref_a = a; T.ref_b = b;
}
// scaleCoarseDepthL0 :329-336
void orc_tracker_scale_idepth(void *p, float scale) {
  Tracker &T = *(Tracker *)p;
  for (int l = 0; l < T.levels; l++) for (int i = 0; i < T.pc_n[l]; i++) T.pc_idepth[l][i] /= scale;
}
// borrowed pointers into a dIp_all buffer produced by orc_make_images (must outlive the calls)
void orc_tracker_set_new_frame(void *p, const float *dIp_all, float exposure) {
  Tracker &T = *(Tracker *)p; int off = 0;
  for (int l = 0; l < T.levels; l++) { T.dIp_new[l] = dIp_all + 3 * off; off += T.w[l] * T.h[l]; }
  T.new_exposure = exposure;
}
void orc_tracker_set_right_frame(void *p, const float *dIp_all) {
  Tracker &T = *(Tracker *)p; int off = 0;
  for (int l = 0; l < T.levels; l++) { T.dIp_right[l] = dIp_all + 3 * off; off += T.w[l] * T.h[l]; }
}

// makeCoarseDepthL0  src/scale_optimization/TrackerAndScaler.cpp:143-315
// points: integer pixel (u,v) = centerProjectedTo+0.5 truncated, idepth, weight = sqrtf(1e-3/(HdiF+1e-12));
// dIp_ref_all = reference keyframe pyramid (lastRef->dIp) for pc_color.
void orc_tracker_make_coarse_depth(void *p, int npts, const int *pu, const int *pv, const float *pid, const float *pweight, const float *dIp_ref_all) {
  Tracker &T = *(Tracker *)p;
  std::vector<float> idepth[kMaxLevels], wsum[kMaxLevels], wbak;
  for (int l = 0; l < T.levels; l++) { idepth[l].assign(T.w[l] * T.h[l], 0.f); wsum[l].assign(T.w[l] * T.h[l], 0.f); }
  for (int i = 0; i < npts; i++) {
    idepth[0][pu[i] + T.w[0] * pv[i]] += pid[i] * pweight[i];
    wsum[0][pu[i] + T.w[0] * pv[i]] += pweight[i];
  }
  for (int lvl = 1; lvl < T.levels; lvl++) {
    const int wl = T.w[lvl], hl = T.h[lvl], wlm1 = T.w[lvl - 1];
    for (int y = 0; y < hl; y++)
      for (int x = 0; x < wl; x++) {
        const int bidx = 2 * x + 2 * y * wlm1;
        idepth[lvl][x + y * wl] = idepth[lvl - 1][bidx] + idepth[lvl - 1][bidx + 1] + idepth[lvl - 1][bidx + wlm1] + idepth[lvl - 1][bidx + wlm1 + 1];
        wsum[lvl][x + y * wl] = wsum[lvl - 1][bidx] + wsum[lvl - 1][bidx + 1] + wsum[lvl - 1][bidx + wlm1] + wsum[lvl - 1][bidx + wlm1 + 1];
      }
  }
  for (int lvl = 0; lvl < T.levels; lvl++) {  // :190-275 dilation: diagonal on lvl 0,1; 4-neighbourhood on lvl >= 2
    const int wl = T.w[lvl], wh = T.w[lvl] * T.h[lvl] - T.w[lvl];
    wbak = wsum[lvl];
    float *idl = idepth[lvl].data(); float *wsl = wsum[lvl].data();
    const int o[2][4] = {{1 + wl, -1 - wl, wl - 1, -wl + 1}, {1, -1, wl, -wl}};
    const int *oo = o[lvl < 2 ? 0 : 1];
    for (int i = wl; i < wh; i++) {
      if (wbak[i] <= 0) {
        float sum = 0, num = 0, numn = 0;
        for (int k = 0; k < 4; k++) {
          // the reference indexes one element outside the grid at (x=0, y=1) and (x=w-1, y=h-2); those pixels are outside
          // the interior [2, w-2) x [2, h-2) the template is read from, so an out-of-range neighbour counts as empty
          const int j = i + oo[k];
          if (j < 0 || j >= wl * T.h[lvl]) continue;
          if (wbak[j] > 0) { sum += idl[j]; num += wbak[j]; numn++; }
        }
        if (numn > 0) { idl[i] = sum / numn; wsl[i] = num / numn; }
      }
    }
  }
  int off = 0;
  for (int lvl = 0; lvl < T.levels; lvl++) {  // :277-314
    const int wl = T.w[lvl], hl = T.h[lvl];
    const float *dIRefl = dIp_ref_all + 3 * off; off += wl * hl;
    T.pc_u[lvl].assign(wl * hl, 0.f); T.pc_v[lvl].assign(wl * hl, 0.f); T.pc_idepth[lvl].assign(wl * hl, 0.f); T.pc_color[lvl].assign(wl * hl, 0.f);
    int lpc_n = 0;
    for (int y = 2; y < hl - 2; y++)
      for (int x = 2; x < wl - 2; x++) {
        const int i = x + y * wl;
        if (wsum[lvl][i] > 0) {
          idepth[lvl][i] /= wsum[lvl][i];
          T.pc_u[lvl][lpc_n] = x; T.pc_v[lvl][lpc_n] = y; T.pc_idepth[lvl][lpc_n] = idepth[lvl][i]; T.pc_color[lvl][lpc_n] = dIRefl[3 * i];
          if (!std::isfinite(T.pc_color[lvl][lpc_n]) || !(idepth[lvl][i] > 0)) { idepth[lvl][i] = -1; continue; }
          lpc_n++;
        } else idepth[lvl][i] = -1;
        wsum[lvl][i] = 1;
      }
    T.pc_n[lvl] = lpc_n;
    T.pc_u[lvl].resize(lpc_n); T.pc_v[lvl].resize(lpc_n); T.pc_idepth[lvl].resize(lpc_n); T.pc_color[lvl].resize(lpc_n);
  }
}

// pose7 = (qx,qy,qz,qw,tx,ty,tz) == Sophus::SE3d::data()
static inline SE3 se3_from7(const double *p) { SE3 s; std::memcpy(s.q, p, 32); std::memcpy(s.t, p + 4, 24); return s; }
static inline void se3_to7(const SE3 &s, double *p) { std::memcpy(p, s.q, 32); std::memcpy(p + 4, s.t, 24); }

void orc_se3_exp(const double a[6], double out7[7]) { se3_to7(se3_exp(a), out7); }
void orc_se3_mul(const double a7[7], const double b7[7], double out7[7]) { se3_to7(se3_mul(se3_from7(a7), se3_from7(b7)), out7); }
void orc_se3_R(const double a7[7], double R[9]) { quat_to_R(a7, R); }
void orc_ldlt_solve(int n, const double *A, const double *rhs, double *x) { ldlt_solve(n, A, 8, rhs, x); }

int orc_calc_res_pose(void *p, int lvl, const double pose7[7], double aff_a, double aff_b, float cutoffTH, double res6[6]) {
  Tracker &T = *(Tracker *)p;
  calc_res_pose(T, lvl, se3_from7(pose7), aff_a, aff_b, cutoffTH, res6);
  return T.pose_n;
}
// warped buffers of the last calcResPose / calcResScale (8 arrays of n floats, order in Tracker::pbuf)
int orc_get_warped(void *p, int which /*0 pose, 1 scale*/, float *out8n) {
  Tracker &T = *(Tracker *)p;
  const int n = which ? T.scale_n : T.pose_n;
  if (out8n) for (int b = 0; b < 8; b++) std::memcpy(out8n + (size_t)b * n, (which ? T.sbuf[b] : T.pbuf[b]).data(), n * 4);
  return n;
}
void orc_calc_gs_pose(void *p, int lvl, int mode, double aff_a, double aff_b, double H64[64], double b8[8], double acc45[45]) {
  calc_gs_pose(*(Tracker *)p, lvl, mode, aff_a, aff_b, H64, b8, acc45);
}
// Jacobian rows of the last calcResPose as calcGSSSEPose forms them (:658-678): J9n = 9 arrays of n floats (J0..J7, r), w[n].
// Lets tests feed the very same rows to the reference's own Accumulator9 (oracle/_ref).
int orc_pose_rows(void *p, int lvl, double aff_a, double aff_b, float *J9n, float *w) {
  Tracker &T = *(Tracker *)p;
  const int n = T.pose_n;
  if (!J9n) return n;
  double affd[2]; aff_from_to(T.ref_exposure, T.new_exposure, T.ref_a, T.ref_b, aff_a, aff_b, affd);
  const float a = (float)affd[0], b0 = (float)T.ref_b;
  for (int i = 0; i < n; i++) {
    float J[8];
    pose_jacobian(T.pbuf[0][i], T.pbuf[1][i], T.pbuf[2][i], T.pbuf[3][i], T.pbuf[4][i], T.pbuf[7][i], T.fx[lvl], T.fy[lvl], a, b0, J);
    for (int j = 0; j < 8; j++) J9n[(size_t)j * n + i] = J[j];
    J9n[(size_t)8 * n + i] = T.pbuf[5][i];
    w[i] = T.pbuf[6][i];
  }
  return n;
}
int orc_scale_rows(void *p, int lvl, float scale, float *J, float *r, float *w) {
  Tracker &T = *(Tracker *)p;
  const int n = T.scale_n;
  if (!J) return n;
  const float tx = (float)T.tfm_f1_f0.t[0], ty = (float)T.tfm_f1_f0.t[1], tz = (float)T.tfm_f1_f0.t[2];
  for (int i = 0; i < n; i++) {
    J[i] = scale_jacobian(T.sbuf[0][i], T.sbuf[1][i], T.sbuf[2][i], T.sbuf[3][i], T.sbuf[4][i], T.fx1[lvl], T.fy1[lvl], scale, tx, ty, tz);
    r[i] = T.sbuf[5][i];
    w[i] = T.sbuf[6][i];
  }
  return n;
}
// The tiered 4-lane accumulators on caller-provided rows (same layout as oracle/_ref's ref_accumulator9 / ref_scale_accumulator)
void orc_accumulator9(const float *J9n, const float *w, int n, float *out45) {
  static thread_local TieredAcc<9> A;
  A.initialize();
  for (int i = 0; i < n; i += 4) {
    float J[9][4], ww[4];
    for (int l = 0; l < 4; l++) {
      for (int j = 0; j < 9; j++) J[j][l] = J9n[(size_t)j * n + i + l];
      ww[l] = w[i + l];
    }
    A.update(J, ww);
  }
  A.finish(out45);
}
void orc_scale_accumulator(const float *Jin, const float *r, const float *w, int n, float *out3) {
  static thread_local TieredAcc<2> A;
  A.initialize();
  for (int i = 0; i < n; i += 4) {
    float J[2][4], ww[4];
    for (int l = 0; l < 4; l++) { J[0][l] = Jin[i + l]; J[1][l] = r[i + l]; ww[l] = w[i + l]; }
    A.update(J, ww);
  }
  A.finish(out3);
}
int orc_track_newest_coarse(void *p, int mode, double pose7_io[7], double aff_io[2], int coarsestLvl, const double minResForAbort[5], double lastResiduals[5], double flow3[3]) {
  Tracker &T = *(Tracker *)p;
  SE3 s = se3_from7(pose7_io);
  const int ok = track_newest_coarse(T, mode, s, aff_io, coarsestLvl, minResForAbort, lastResiduals);
  se3_to7(s, pose7_io);
  for (int i = 0; i < 3; i++) flow3[i] = T.lastFlow[i];
  return ok;
}
// The hypothesis loop of FrontEnd::trackNewCoarse  src/FrontEnd.cpp:192-252 (restated; FrontEnd.cpp itself needs the whole
// DSO system and cannot be compiled in place).  tries7 = ntries poses, every trial starts from aff_init.
// Returns haveOneGood; *tries_out = tryIterations.
int orc_track_new_coarse(void *p, int mode, int ntries, const double *tries7, const double aff_init[2], int coarsestLvl, const double last_coarse_rmse[5],
                         double reTrackThreshold, double pose7_out[7], double aff_out[2], double achievedRes[5], double flow3[3], int *tries_out) {
  Tracker &T = *(Tracker *)p;
  double flowVecs[3] = {100, 100, 100};
  SE3 lastF_2_fh = se3_identity();
  double aff_g2l[2] = {0, 0};
  for (int i = 0; i < 5; i++) achievedRes[i] = NAN;
  bool haveOneGood = false;
  int tryIterations = 0;
  for (int i = 0; i < ntries; i++) {
    double aff_this[2] = {aff_init[0], aff_init[1]};
    SE3 lastF_2_fh_this = se3_from7(tries7 + 7 * i);
    double currentRes[5];
    const bool trackingIsGood = track_newest_coarse(T, mode, lastF_2_fh_this, aff_this, coarsestLvl, achievedRes, currentRes) != 0;
    tryIterations++;
    if (trackingIsGood && std::isfinite((float)currentRes[0]) && !(currentRes[0] >= achievedRes[0])) {
      for (int k = 0; k < 3; k++) flowVecs[k] = T.lastFlow[k];
      aff_g2l[0] = aff_this[0]; aff_g2l[1] = aff_this[1];
      lastF_2_fh = lastF_2_fh_this;
      haveOneGood = true;
    }
    if (haveOneGood)
      for (int k = 0; k < 5; k++)
        if (!std::isfinite((float)achievedRes[k]) || achievedRes[k] > currentRes[k]) achievedRes[k] = currentRes[k];
    if (haveOneGood && achievedRes[0] < last_coarse_rmse[0] * reTrackThreshold) break;
  }
  if (!haveOneGood) {
    flowVecs[0] = flowVecs[1] = flowVecs[2] = 0;
    aff_g2l[0] = aff_init[0]; aff_g2l[1] = aff_init[1];
    lastF_2_fh = se3_from7(tries7);
  }
  se3_to7(lastF_2_fh, pose7_out);
  aff_out[0] = aff_g2l[0]; aff_out[1] = aff_g2l[1];
  for (int k = 0; k < 3; k++) flow3[k] = flowVecs[k];
  *tries_out = tryIterations;
  return haveOneGood ? 1 : 0;
}
int orc_calc_res_scale(void *p, int lvl, float scale, float cutoffTH, double res6[6]) {
  Tracker &T = *(Tracker *)p;
  calc_res_scale(T, lvl, scale, cutoffTH, res6);
  return T.scale_n;
}
void orc_calc_gs_scale(void *p, int lvl, int mode, float scale, float *H, float *b, double acc3[3]) {
  calc_gs_scale(*(Tracker *)p, lvl, mode, scale, H, b, acc3);
}
float orc_optimize_scale(void *p, int mode, float *scale_io, int coarsestLvl) {
  return optimize_scale(*(Tracker *)p, mode, *scale_io, coarsestLvl);
}
// trace of the last track / optimizeScale: rows of 15 doubles (lvl, it, accept, n, lambda, e_old, e_new, inc[8])
int orc_get_trace(void *p, double *out, int max_rows) {
  Tracker &T = *(Tracker *)p;
  const int n = (int)T.trace.size();
  if (out)
    for (int i = 0; i < n && i < max_rows; i++) {
      const TraceRec &r = T.trace[i]; double *o = out + 15 * i;
      o[0] = r.lvl; o[1] = r.iteration; o[2] = r.accept; o[3] = r.n; o[4] = r.lambda; o[5] = r.e_old; o[6] = r.e_new;
      for (int k = 0; k < 8; k++) o[7 + k] = r.inc[k];
    }
  return n;
}
void orc_get_counters(void *p, long out[2]) { Tracker &T = *(Tracker *)p; out[0] = T.n_res_evals; out[1] = T.n_gs_evals; }

// ---- PoseEstimator -------------------------------------------------------------------------------
void *orc_pe_create(int w, int h, int levels, const float cam[4]) {
  PoseEst *P = new PoseEst();
  P->levels = levels;
  pe_make_K(*P, w, h, cam);
  for (int l = 0; l < kMaxLevels; l++) P->dIp_new[l] = nullptr;
  return P;
}
void orc_pe_destroy(void *p) { delete (PoseEst *)p; }
// pts: n x 3 doubles; colors: levels arrays of n floats, level-major
void orc_pe_set_points(void *p, int n, const double *pts, const float *colors, float ref_exposure) {
  PoseEst &P = *(PoseEst *)p;
  P.n = n;
  P.pts.assign(pts, pts + 3 * (size_t)n);
  for (int l = 0; l < P.levels; l++) P.colors[l].assign(colors + (size_t)l * n, colors + (size_t)(l + 1) * n);
  P.ref_exposure = ref_exposure;
}
void orc_pe_set_new_frame(void *p, const float *dIp_all, float exposure) {
  PoseEst &P = *(PoseEst *)p; int off = 0;
  for (int l = 0; l < P.levels; l++) { P.dIp_new[l] = dIp_all + 3 * off; off += P.w[l] * P.h[l]; }
  P.new_exposure = exposure;
}
void orc_pe_set_aff_mode(void *p, int a, int b) { ((PoseEst *)p)->affModeA = a; ((PoseEst *)p)->affModeB = b; }
int orc_pe_calc_res(void *p, int lvl, int mode, const double T16[16], double aff_a, double aff_b, float cutoff, double res6[6], double H64[64], double b8[8],
                    double acc45[45]) {
  PoseEst &P = *(PoseEst *)p;
  P.res_acc_mode = mode;
  pe_calc_res(P, lvl, se3_from_matrix4(T16), aff_a, aff_b, cutoff, res6);
  if (H64) pe_calc_gs(P, lvl, mode, aff_a, aff_b, H64, b8, acc45);
  return P.buf_n;
}
int orc_pe_estimate(void *p, int mode, double T_io[16], int coarsest_lvl, float *pose_error, int *inlier_percent) {
  return pe_estimate(*(PoseEst *)p, mode, T_io, coarsest_lvl, pose_error, inlier_percent);
}
int orc_pe_get_trace(void *p, double *out, int max_rows) {
  PoseEst &P = *(PoseEst *)p;
  const int n = (int)P.trace.size();
  if (out)
    for (int i = 0; i < n && i < max_rows; i++) {
      const TraceRec &r = P.trace[i]; double *o = out + 15 * i;
      o[0] = r.lvl; o[1] = r.iteration; o[2] = r.accept; o[3] = r.n; o[4] = r.lambda; o[5] = r.e_old; o[6] = r.e_new;
      for (int k = 0; k < 8; k++) o[7 + k] = r.inc[k];
    }
  return n;
}

// ---- Scan Context --------------------------------------------------------------------------------
// search_sc  src/loop_closure/loop_detection/search_place.h:59-85 on sparse index-sorted signatures.
// DB signatures are CSR: sig_ptr[N+1], sig_idx[], sig_val[] (double); query likewise (q_idx,q_val,q_nnz).
void orc_search_sc(const int *q_idx, const double *q_val, int q_nnz, const int *sig_ptr, const int *sig_idx, const double *sig_val,
                   const int *candidates, int n_cand, int sc_width, int *res_idx, float *res_diff) {
  *res_idx = candidates[0];
  *res_diff = 1.1;
  for (int ci = 0; ci < n_cand; ci++) {
    const int cand = candidates[ci];
    float cur_prod = 0;
    int m = 0, n = sig_ptr[cand];
    const int nend = sig_ptr[cand + 1];
    while (m < q_nnz && n < nend) {
      if (q_idx[m] == sig_idx[n]) {
        cur_prod += q_val[m++] * sig_val[n++];  // float += double*double (:73)
      } else {
        q_idx[m] < sig_idx[n] ? m++ : n++;
      }
    }
    const float cur_diff = (1 - cur_prod / sc_width) / 2.0;
    if (*res_diff > cur_diff) { *res_idx = cand; *res_diff = cur_diff; }
  }
}

// Same arithmetic on dense fp32 descriptors (n_cells = sectors*rings, 0 == empty cell): equals search_sc on the
// sparse form of the same data as long as stored values are exactly representable (zeros contribute +0.0).
// cand == nullptr -> all rows [0, n_rows) in ascending order (brute force; ties -> lowest id by the strict '>').
void orc_search_sc_dense(const float *q, const float *db, int n_cells, const int *candidates, int n_cand, int sc_width, int *res_idx, float *res_diff) {
  *res_idx = candidates ? candidates[0] : 0;
  *res_diff = 1.1;
  for (int ci = 0; ci < n_cand; ci++) {
    const int cand = candidates ? candidates[ci] : ci;
    const float *d = db + (size_t)cand * n_cells;
    float cur_prod = 0;
    for (int k = 0; k < n_cells; k++)
      if (q[k] != 0.0f && d[k] != 0.0f) cur_prod += (double)q[k] * (double)d[k];
    const float cur_diff = (1 - cur_prod / sc_width) / 2.0;
    if (*res_diff > cur_diff) { *res_idx = cand; *res_diff = cur_diff; }
  }
}

// flann::L2<float>::operator() (FLANN 1.9.1 flann/algorithms/dist.h, un-vendored dependency): squared
// Euclidean distance, 4 elements per step: result += d0*d0 + d1*d1 + d2*d2 + d3*d3, then a scalar tail.
static inline float flann_l2(const float *a, const float *b, int size) {
  float result = 0;
  int i = 0;
  for (; i + 3 < size; i += 4) {
    const float d0 = a[i] - b[i], d1 = a[i + 1] - b[i + 1], d2 = a[i + 2] - b[i + 2], d3 = a[i + 3] - b[i + 3];
    result += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
  }
  for (; i < size; i++) { const float d0 = a[i] - b[i]; result += d0 * d0; }
  return result;
}

// Exact brute-force replacement of search_ringkey's knnSearch (search_place.h:25-40): k nearest rows of
// keys[0..n_rows) by squared L2 (ties -> lowest row), then keep dist < thres. Row ids are returned as-is
// (the reference's "idx>0 ... idx-1" dummy-row bookkeeping is the caller's: LoopHandler.cpp:35-39).
int orc_search_ringkey(const float *q, const float *keys, int n_rows, int dim, int k, float thres, int *cand_out, float *dist_out) {
  std::vector<std::pair<float, int>> best;
  for (int r = 0; r < n_rows; r++) {
    const float d = flann_l2(q, keys + (size_t)r * dim, dim);
    if ((int)best.size() < k) { best.emplace_back(d, r); std::sort(best.begin(), best.end()); }
    else if (std::make_pair(d, r) < best.back()) { best.back() = {d, r}; std::sort(best.begin(), best.end()); }
  }
  int n = 0;
  for (auto &b : best) if (b.first < thres) { cand_out[n] = b.second; if (dist_out) dist_out[n] = b.first; n++; }
  return n;
}

}  // extern "C"
