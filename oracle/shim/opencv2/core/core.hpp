#pragma once
#include <vector>
namespace cv {
struct Mat {  // just enough for "cv::Mat(tfm_vec).reshape(0, 4)" + cv2eigen (TrackerAndScaler.cpp:82-86)
  std::vector<double> v;
  Mat() {}
  explicit Mat(const std::vector<double> &x) : v(x) {}
  Mat reshape(int, int) const { return *this; }
};
}  // namespace cv
