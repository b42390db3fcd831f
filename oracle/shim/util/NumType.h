// Stand-in for deps:dso/src/util/NumType.h: only the typedef names used by
//   deps:dso/src/OptimizationBackend/MatrixAccumulators.h and src/scale_optimization/ScaleAccumulator.h
#pragma once
#include <Eigen/Core>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <xmmintrin.h>
#include <emmintrin.h>
namespace dso {
typedef Eigen::Matrix<float, 2, 2> Mat22f;
typedef Eigen::Matrix<float, 2, 1> Vec2f;
typedef Eigen::Matrix<float, 9, 9> Mat99f;
typedef Eigen::Matrix<float, 9, 1> Vec9f;
typedef Eigen::Matrix<float, 13, 13> Mat1313f;
typedef Eigen::Matrix<float, 14, 14> Mat1414f;
typedef Eigen::Matrix<float, 14, 1> Vec14f;
}  // namespace dso
