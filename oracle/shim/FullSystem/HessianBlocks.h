// Stand-in for deps:dso/src/FullSystem/HessianBlocks.h: the fields of FrameHessian / PointHessian / CalibHessian the
// tracker reads, and the SCALE_* constants (HessianBlocks.h:58-65).
#pragma once
#include <utility>
#include <vector>
#include "util/NumType.h"
#include "util/settings.h"
#define SCALE_IDEPTH 1.0f
#define SCALE_XI_ROT 1.0f
#define SCALE_XI_TRANS 0.5f
#define SCALE_F 50.0f
#define SCALE_C 50.0f
#define SCALE_W 1.0f
#define SCALE_A 10.0f
#define SCALE_B 1000.0f
namespace dso {
struct FrameHessian;
struct FrameShell { int id = 0; };
struct EFPoint { float HdiF = 0; };
struct EFResidual { bool isActive() const { return true; } };
enum ResState { IN = 0, OOB, OUTLIER };
struct PointFrameResidual {
  EFResidual *efResidual = nullptr;
  FrameHessian *target = nullptr;
  Vec3f centerProjectedTo;
};
struct PointHessian {
  std::pair<PointFrameResidual *, ResState> lastResiduals[2];
  EFPoint *efPoint = nullptr;
};
struct CalibHessian;
struct FrameHessian {
  Eigen::Vector3f *dI = nullptr;  // = dIp[0]
  Eigen::Vector3f *dIp[PYR_LEVELS];
  float *absSquaredGrad[PYR_LEVELS];
  void makeImages(float *color, CalibHessian *HCalib);  // defined by the reference's HessianBlocks.cpp:128-191
  float ab_exposure = 1.f;
  FrameShell *shell = nullptr;
  std::vector<PointHessian *> pointHessians;
  AffLight aff;
  AffLight aff_g2l() const { return aff; }
};
struct CalibHessian {
  float fx, fy, cx, cy;
  float B[256];  // inverse response; identity unless a gamma file is loaded (HessianBlocks.h:329-330)
  float getBGradOnly(float color) {  // HessianBlocks.h:384-390
    int c = color + 0.5f;
    if (c < 5) c = 5;
    if (c > 250) c = 250;
    return B[c + 1] - B[c];
  }
  float fxl() const { return fx; }
  float fyl() const { return fy; }
  float cxl() const { return cx; }
  float cyl() const { return cy; }
};
}  // namespace dso
