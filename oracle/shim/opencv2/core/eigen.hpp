#pragma once
#include <Eigen/Core>
#include <opencv2/core/core.hpp>
namespace cv {
inline void cv2eigen(const Mat &m, Eigen::Matrix4d &out) {
  for (int i = 0; i < 16; i++) out.d[i] = m.v[(size_t)i];  // row-major 4x4
}
}  // namespace cv
