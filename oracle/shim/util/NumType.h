// Stand-in for deps:dso/src/util/NumType.h (TEST INFRASTRUCTURE ONLY): the typedef names and the two small value types
// the hot path uses.  SE3 restates Sophus::SE3d (deps:dso/thirdparty/Sophus/sophus/se3.hpp:160-163, 239-243, 268-271,
// 407-428; so3.hpp:196-202, 343-369, 631-633) exactly like oracle/dslam_oracle.cpp does; AffLight restates
// deps:dso/src/util/NumType.h:166-192.
#pragma once
#include <Eigen/Core>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>
#include <xmmintrin.h>
#include <emmintrin.h>

namespace dso {

typedef Eigen::Matrix<double, 2, 1> Vec2;
typedef Eigen::Matrix<double, 3, 1> Vec3;
typedef Eigen::Matrix<double, 5, 1> Vec5;
typedef Eigen::Matrix<double, 6, 1> Vec6;
typedef Eigen::Matrix<double, 7, 1> Vec7;
typedef Eigen::Matrix<double, 8, 1> Vec8;
typedef Eigen::Matrix<double, 3, 3> Mat33;
typedef Eigen::Matrix<double, 8, 8> Mat88;
typedef Eigen::Matrix<float, 2, 1> Vec2f;
typedef Eigen::Matrix<float, 3, 1> Vec3f;
typedef Eigen::Matrix<float, 9, 1> Vec9f;
typedef Eigen::Matrix<float, 14, 1> Vec14f;
typedef Eigen::Matrix<float, 2, 2> Mat22f;
typedef Eigen::Matrix<float, 3, 3> Mat33f;
typedef Eigen::Matrix<float, 9, 9> Mat99f;
typedef Eigen::Matrix<float, 13, 13> Mat1313f;
typedef Eigen::Matrix<float, 14, 14> Mat1414f;
typedef Eigen::Matrix<unsigned char, 3, 1> Vec3b;

class SE3 {
 public:
  SE3() { q_[0] = q_[1] = q_[2] = 0; q_[3] = 1; t_.setZero(); }
  // SE3(Matrix4d): quaternion from the rotation block (Eigen's Quaternion(Matrix3): Shepperd), translation from column 3
  explicit SE3(const Eigen::Matrix4d &T) {
    auto M = [&](int r, int c) { return T(r, c); };
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
    std::memcpy(q_, q, sizeof(q));
    normalize();
    t_[0] = T(0, 3); t_[1] = T(1, 3); t_[2] = T(2, 3);
  }
  // SE3(rotation matrix, translation) — both may be matrices or block views
  template <class RotT, class TransT>
  SE3(const RotT &R, const TransT &t) {
    Eigen::Matrix4d T;
    T.setZero();
    for (int r = 0; r < 3; r++) {
      for (int c = 0; c < 3; c++) T(r, c) = R(r, c);
      T(r, 3) = t[r];
    }
    T(3, 3) = 1;
    *this = SE3(T);
  }
  Eigen::Matrix4d matrix() const {
    Eigen::Matrix4d T;
    T.setZero();
    const Mat33 R = rotationMatrix();
    for (int r = 0; r < 3; r++) {
      for (int c = 0; c < 3; c++) T(r, c) = R(r, c);
      T(r, 3) = t_[r];
    }
    T(3, 3) = 1;
    return T;
  }
  static SE3 from7(const double *p) {
    SE3 s;
    std::memcpy(s.q_, p, 32);
    s.t_[0] = p[4]; s.t_[1] = p[5]; s.t_[2] = p[6];
    return s;
  }
  void to7(double *p) const { std::memcpy(p, q_, 32); p[4] = t_[0]; p[5] = t_[1]; p[6] = t_[2]; }

  Mat33 rotationMatrix() const {  // Eigen QuaternionBase::toRotationMatrix
    const double x = q_[0], y = q_[1], z = q_[2], w = q_[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    Mat33 R;
    R(0, 0) = 1 - (tyy + tzz); R(0, 1) = txy - twz; R(0, 2) = txz + twy;
    R(1, 0) = txy + twz; R(1, 1) = 1 - (txx + tzz); R(1, 2) = tyz - twx;
    R(2, 0) = txz - twy; R(2, 1) = tyz + twx; R(2, 2) = 1 - (txx + tyy);
    return R;
  }
  const Vec3 &translation() const { return t_; }
  Vec3 &translation() { return t_; }

  SE3 operator*(const SE3 &b) const {  // se3.hpp:239-243 -> operator*= -> fastMultiply + normalize
    SE3 r;
    const double v[3] = {b.t_[0], b.t_[1], b.t_[2]};
    double uv[3] = {q_[1] * v[2] - q_[2] * v[1], q_[2] * v[0] - q_[0] * v[2], q_[0] * v[1] - q_[1] * v[0]};
    uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
    const double c[3] = {q_[1] * uv[2] - q_[2] * uv[1], q_[2] * uv[0] - q_[0] * uv[2], q_[0] * uv[1] - q_[1] * uv[0]};
    for (int i = 0; i < 3; i++) r.t_[i] = t_[i] + (v[i] + q_[3] * uv[i] + c[i]);
    const double ax = q_[0], ay = q_[1], az = q_[2], aw = q_[3], bx = b.q_[0], by = b.q_[1], bz = b.q_[2], bw = b.q_[3];
    r.q_[3] = aw * bw - ax * bx - ay * by - az * bz;
    r.q_[0] = aw * bx + ax * bw + ay * bz - az * by;
    r.q_[1] = aw * by + ay * bw + az * bx - ax * bz;
    r.q_[2] = aw * bz + az * bw + ax * by - ay * bx;
    r.normalize();
    return r;
  }
  static SE3 exp(const Vec6 &a) {  // se3.hpp:407-428
    const double *ups = a.d, *om = a.d + 3;
    const double theta_sq = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
    const double theta = std::sqrt(theta_sq), half_theta = 0.5 * theta;
    const double eps = 1e-10;
    double imag, real;
    if (theta < eps) {
      const double theta_po4 = theta_sq * theta_sq;
      imag = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * theta_po4;
      real = 1.0 - 0.5 * theta_sq + (1.0 / 384.0) * theta_po4;
    } else {
      imag = std::sin(half_theta) / theta;
      real = std::cos(half_theta);
    }
    SE3 r;
    r.q_[3] = real; r.q_[0] = imag * om[0]; r.q_[1] = imag * om[1]; r.q_[2] = imag * om[2];
    r.normalize();
    const double Om[9] = {0, -om[2], om[1], om[2], 0, -om[0], -om[1], om[0], 0};
    double Om2[9];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += Om[i * 3 + k] * Om[k * 3 + j];
        Om2[i * 3 + j] = s;
      }
    double V[9];
    if (theta < eps) {
      const Mat33 R = r.rotationMatrix();
      for (int i = 0; i < 9; i++) V[i] = R.d[i];
    } else {
      const double c1 = (1.0 - std::cos(theta)) / theta_sq, c2 = (theta - std::sin(theta)) / (theta_sq * theta);
      for (int i = 0; i < 9; i++) V[i] = ((i % 4 == 0) ? 1.0 : 0.0) + c1 * Om[i] + c2 * Om2[i];
    }
    for (int i = 0; i < 3; i++) r.t_[i] = V[i * 3] * ups[0] + V[i * 3 + 1] * ups[1] + V[i * 3 + 2] * ups[2];
    return r;
  }
  Vec6 log() const { Vec6 v; v.setZero(); return v; }  // only reached from DEBUG_PRINT code
  double *data() { return q_; }

 private:
  void normalize() {
    const double len = std::sqrt(q_[0] * q_[0] + q_[1] * q_[1] + q_[2] * q_[2] + q_[3] * q_[3]);
    for (int i = 0; i < 4; i++) q_[i] /= len;
  }
  double q_[4];
  Vec3 t_;
};

struct AffLight {
  AffLight(double a_, double b_) : a(a_), b(b_) {}
  AffLight() : a(0), b(0) {}
  double a, b;
  static Vec2 fromToVecExposure(float exposureF, float exposureT, AffLight g2F, AffLight g2T) {
    if (exposureF == 0 || exposureT == 0) exposureT = exposureF = 1;
    const double a = exp(g2T.a - g2F.a) * exposureT / exposureF;
    const double b = g2T.b - a * g2F.b;
    return Vec2(a, b);
  }
  Vec2 vec() { return Vec2(a, b); }
};

}  // namespace dso
