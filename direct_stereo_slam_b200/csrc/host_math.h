// host_math.h — the O(1) host-side algebra that stays on the CPU next to the kernels (north star: "the tiny
// 8x8 solve stays on the host"): Sophus-compatible SE3 (unit quaternion + translation, double), the 8x8
// damped normal-equation solve, camera pyramids and the affine brightness transfer.
//
// Mirrors, in arithmetic and operation order:
//   Sophus  deps:dso/thirdparty/Sophus/sophus/se3.hpp:160-163,239-243,268-271,407-428 ; so3.hpp:196-202,343-369,631-633
//   TrackerAndScaler::makeK / ctor  src/scale_optimization/TrackerAndScaler.cpp:88-98, 117-141
//   AffLight::fromToVecExposure     deps:dso/src/util/NumType.h:173-185
// Compiled with -ffp-contract=off (host) so float expressions round like the reference's scalar code.
#pragma once
#include <cmath>
#include <cstring>

namespace dslam {
namespace hm {

struct Quat {
  double x, y, z, w;
};

struct Se3 {
  Quat q{0, 0, 0, 1};
  double t[3]{0, 0, 0};

  static Se3 from7(const double *p) {
    Se3 s;
    s.q = Quat{p[0], p[1], p[2], p[3]};
    s.t[0] = p[4]; s.t[1] = p[5]; s.t[2] = p[6];
    return s;
  }
  void to7(double *p) const {
    p[0] = q.x; p[1] = q.y; p[2] = q.z; p[3] = q.w;
    p[4] = t[0]; p[5] = t[1]; p[6] = t[2];
  }
};

inline Quat qnormalized(Quat q) {
  const double n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  return Quat{q.x / n, q.y / n, q.z / n, q.w / n};
}

inline Quat qmul(const Quat &a, const Quat &b) {
  Quat r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}

// v' = v + w*(2 qv x v) + qv x (2 qv x v)
inline void qrotate(const Quat &q, const double v[3], double out[3]) {
  double uv[3] = {q.y * v[2] - q.z * v[1], q.z * v[0] - q.x * v[2], q.x * v[1] - q.y * v[0]};
  uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
  const double c0 = q.y * uv[2] - q.z * uv[1], c1 = q.z * uv[0] - q.x * uv[2], c2 = q.x * uv[1] - q.y * uv[0];
  out[0] = v[0] + q.w * uv[0] + c0;
  out[1] = v[1] + q.w * uv[1] + c1;
  out[2] = v[2] + q.w * uv[2] + c2;
}

inline void qmatrix(const Quat &q, double R[9]) {
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}

// rotation matrix (row-major 3x3 inside a row-major 4x4 with stride ld) -> quaternion
inline Quat qfrom_matrix(const double *m, int ld) {
  auto M = [&](int r, int c) { return m[r * ld + c]; };
  double q[4];
  const double tr = M(0, 0) + M(1, 1) + M(2, 2);
  if (tr > 0) {
    double t = std::sqrt(tr + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (M(2, 1) - M(1, 2)) * t; q[1] = (M(0, 2) - M(2, 0)) * t; q[2] = (M(1, 0) - M(0, 1)) * t;
  } else {
    int i = 0;
    if (M(1, 1) > M(0, 0)) i = 1;
    if (M(2, 2) > M(i, i)) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    double t = std::sqrt(M(i, i) - M(j, j) - M(k, k) + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (M(k, j) - M(j, k)) * t; q[j] = (M(j, i) + M(i, j)) * t; q[k] = (M(k, i) + M(i, k)) * t;
  }
  return qnormalized(Quat{q[0], q[1], q[2], q[3]});
}

inline Se3 operator*(const Se3 &a, const Se3 &b) {
  Se3 r;
  double rt[3];
  qrotate(a.q, b.t, rt);
  for (int i = 0; i < 3; i++) r.t[i] = a.t[i] + rt[i];
  r.q = qnormalized(qmul(a.q, b.q));
  return r;
}

inline Se3 se3_exp(const double a[6]) {
  const double *ups = a, *om = a + 3;
  const double th2 = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
  const double th = std::sqrt(th2), half = 0.5 * th;
  constexpr double kEps = 1e-10;
  double im, re;
  if (th < kEps) {
    const double th4 = th2 * th2;
    im = 0.5 - (1.0 / 48.0) * th2 + (1.0 / 3840.0) * th4;
    re = 1.0 - 0.5 * th2 + (1.0 / 384.0) * th4;
  } else {
    im = std::sin(half) / th;
    re = std::cos(half);
  }
  Se3 r;
  r.q = qnormalized(Quat{im * om[0], im * om[1], im * om[2], re});
  const double Om[9] = {0, -om[2], om[1], om[2], 0, -om[0], -om[1], om[0], 0};
  double Om2[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += Om[i * 3 + k] * Om[k * 3 + j];
      Om2[i * 3 + j] = s;
    }
  double V[9];
  if (th < kEps) {
    qmatrix(r.q, V);
  } else {
    const double c1 = (1.0 - std::cos(th)) / th2, c2 = (th - std::sin(th)) / (th2 * th);
    for (int i = 0; i < 9; i++) V[i] = ((i % 4 == 0) ? 1.0 : 0.0) + c1 * Om[i] + c2 * Om2[i];
  }
  for (int i = 0; i < 3; i++) r.t[i] = V[i * 3] * ups[0] + V[i * 3 + 1] * ups[1] + V[i * 3 + 2] * ups[2];
  return r;
}

// Symmetric indefinite-safe LDL^T with diagonal pivoting for n <= 8 (the role of Eigen's ldlt().solve()).
inline void ldlt_solve(int n, const double *Ain, int lda, const double *rhs, double *x) {
  double A[8][8];
  int piv[8];
  for (int i = 0; i < n; i++) {
    piv[i] = i;
    for (int j = 0; j < n; j++) A[i][j] = Ain[i * lda + j];
  }
  for (int k = 0; k < n; k++) {
    int p = k;
    for (int i = k + 1; i < n; i++)
      if (std::fabs(A[i][i]) > std::fabs(A[p][p])) p = i;
    if (p != k) {
      for (int j = 0; j < n; j++) { const double tmp = A[k][j]; A[k][j] = A[p][j]; A[p][j] = tmp; }
      for (int i = 0; i < n; i++) { const double tmp = A[i][k]; A[i][k] = A[i][p]; A[i][p] = tmp; }
      const int ti = piv[k]; piv[k] = piv[p]; piv[p] = ti;
    }
    const double d = A[k][k];
    if (d == 0.0) continue;
    for (int i = k + 1; i < n; i++) {
      const double l = A[i][k] / d;
      for (int j = k + 1; j < n; j++) A[i][j] -= l * A[k][j];
      A[i][k] = l;
    }
  }
  double y[8], z[8];
  for (int i = 0; i < n; i++) {
    double s = rhs[piv[i]];
    for (int j = 0; j < i; j++) s -= A[i][j] * y[j];
    y[i] = s;
  }
  for (int i = 0; i < n; i++) y[i] = (std::fabs(A[i][i]) > 1e-300) ? y[i] / A[i][i] : 0.0;
  for (int i = n - 1; i >= 0; i--) {
    double s = y[i];
    for (int j = i + 1; j < n; j++) s -= A[j][i] * z[j];
    z[i] = s;
  }
  for (int i = 0; i < n; i++) x[piv[i]] = z[i];
}

// e0 + (e1 + e2): the reduction order of Eigen 3.3's unrolled 3-term coefficient product
inline float dot3(float a0, float b0, float a1, float b1, float a2, float b2) { return a0 * b0 + (a1 * b1 + a2 * b2); }

inline void mat33f_mul(const float A[9], const float B[9], float C[9]) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) C[i * 3 + j] = dot3(A[i * 3], B[j], A[i * 3 + 1], B[3 + j], A[i * 3 + 2], B[6 + j]);
}

inline float cofactor3(const float m[9], int i, int j) {
  const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
  return m[i1 * 3 + j1] * m[i2 * 3 + j2] - m[i1 * 3 + j2] * m[i2 * 3 + j1];
}
// Eigen's 3x3 inverse by cofactors (Ki_[level] = K.inverse(), TrackerAndScaler.cpp:139)
inline void mat33f_inverse(const float m[9], float r[9]) {
  const float c0 = cofactor3(m, 0, 0), c1 = cofactor3(m, 1, 0), c2 = cofactor3(m, 2, 0);
  const float invdet = 1.0f / dot3(c0, m[0], c1, m[3], c2, m[6]);
  r[0] = c0 * invdet; r[1] = c1 * invdet; r[2] = c2 * invdet;
  for (int i = 1; i < 3; i++)
    for (int j = 0; j < 3; j++) r[i * 3 + j] = cofactor3(m, j, i) * invdet;
}

inline void aff_from_to(float exposureF, float exposureT, double aF, double bF, double aT, double bT, double out[2]) {
  if (exposureF == 0 || exposureT == 0) exposureT = exposureF = 1;
  const double a = std::exp(aT - aF) * exposureT / exposureF;
  out[0] = a;
  out[1] = bT - a * bF;
}

// Symmetric 3x3 eigen-decomposition by cyclic Jacobi rotations: eigenvalues ascending (the order Eigen's
// SelfAdjointEigenSolver returns them in, ScanContext.cpp:42-46), eigenvectors as the COLUMNS of evecs (row-major 3x3),
// each normalised and sign-fixed so that its largest-magnitude component is positive (Eigen's signs are an artefact of
// its QR iteration; this library documents its own convention instead).
inline void sym_eig3(const double Ain[9], double evals[3], double evecs[9]) {
  double A[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int i = 0; i < 9; i++) A[i] = Ain[i];
  for (int sweep = 0; sweep < 64; sweep++) {
    const double off = A[1] * A[1] + A[2] * A[2] + A[5] * A[5];
    if (off < 1e-300) break;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        const double apq = A[p * 3 + q];
        if (std::fabs(apq) < 1e-300) continue;
        const double theta = (A[q * 3 + q] - A[p * 3 + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), sn = t * c;
        for (int k = 0; k < 3; k++) {
          const double akp = A[k * 3 + p], akq = A[k * 3 + q];
          A[k * 3 + p] = c * akp - sn * akq; A[k * 3 + q] = sn * akp + c * akq;
        }
        for (int k = 0; k < 3; k++) {
          const double apk = A[p * 3 + k], aqk = A[q * 3 + k];
          A[p * 3 + k] = c * apk - sn * aqk; A[q * 3 + k] = sn * apk + c * aqk;
        }
        for (int k = 0; k < 3; k++) {
          const double vkp = V[k * 3 + p], vkq = V[k * 3 + q];
          V[k * 3 + p] = c * vkp - sn * vkq; V[k * 3 + q] = sn * vkp + c * vkq;
        }
      }
  }
  int order[3] = {0, 1, 2};
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 2 - i; j++)
      if (A[order[j + 1] * 4] < A[order[j] * 4]) { const int tmp = order[j]; order[j] = order[j + 1]; order[j + 1] = tmp; }
  for (int j = 0; j < 3; j++) {
    const int src = order[j];
    evals[j] = A[src * 3 + src];
    const double v[3] = {V[src], V[3 + src], V[6 + src]};
    const double nrm = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    int big = 0;
    for (int k = 1; k < 3; k++)
      if (std::fabs(v[k]) > std::fabs(v[big])) big = k;
    const double sgn = (v[big] < 0 ? -1.0 : 1.0) / nrm;
    for (int k = 0; k < 3; k++) evecs[k * 3 + j] = v[k] * sgn;
  }
}

struct CamPyramid {
  float fx[8], fy[8], cx[8], cy[8];
  void set(int levels, float fx0, float fy0, float cx0, float cy0) {
    fx[0] = fx0; fy[0] = fy0; cx[0] = cx0; cy[0] = cy0;
    for (int l = 1; l < levels; l++) {
      fx[l] = fx[l - 1] * 0.5;
      fy[l] = fy[l - 1] * 0.5;
      cx[l] = (cx[0] + 0.5) / ((int)1 << l) - 0.5;
      cy[l] = (cy[0] + 0.5) / ((int)1 << l) - 0.5;
    }
  }
};

}  // namespace hm
}  // namespace dslam
