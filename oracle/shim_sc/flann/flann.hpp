// Stand-in for <flann/flann.hpp>: only the row-major matrix view ScanContext::generate fills (ScanContext.cpp:87-90).
#pragma once
#include <cstddef>
namespace flann {
template <typename T>
struct Matrix {
  T *data = nullptr;
  size_t rows = 0, cols = 0;
  Matrix() {}
  Matrix(T *d, size_t r, size_t c) : data(d), rows(r), cols(c) {}
  T *operator[](size_t r) const { return data + r * cols; }
};
}  // namespace flann
