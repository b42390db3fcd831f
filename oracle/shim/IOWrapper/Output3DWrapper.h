#pragma once
#include <vector>
#include "util/NumType.h"
namespace dso {
struct MinimalImageB3 {  // only what the plot_img branches touch (never executed: plot_img is false)
  int w, h;
  std::vector<Vec3b> data;
  MinimalImageB3(int w_, int h_) : w(w_), h(h_), data((size_t)w_ * h_) {}
  void setBlack() {}
  void setConst(Vec3b) {}
  void setPixel4(float, float, Vec3b) {}
  void setPixel1(float, float, Vec3b) {}
  void setPixel9(int, int, Vec3b) {}
  Vec3b &at(int i) { return data[(size_t)i]; }
};
namespace IOWrap {
class Output3DWrapper {};
inline void displayImage(const char *, MinimalImageB3 *, bool = false) {}
inline int waitKey(int) { return 0; }
}  // namespace IOWrap
}  // namespace dso
