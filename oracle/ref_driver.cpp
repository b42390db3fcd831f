// ref_driver.cpp — thin C entry points around pieces of the REFERENCE's own source, compiled in place from
// /root/reference (never copied into this repository) by oracle/ref_build.py:
//   deps:dso/src/OptimizationBackend/MatrixAccumulators.h   (Accumulator9 — extracted from dependencies.zip to a temp dir)
//   src/scale_optimization/ScaleAccumulator.h               (ScaleAccumulator)
//   src/loop_closure/loop_detection/search_place.h          (search_ringkey, search_sc)
// Eigen / FLANN / LoopFrame are replaced by the tiny shims under oracle/shim and below.  FLANN's kd-tree is replaced
// by an exact brute-force index (flann::L2 arithmetic restated from FLANN 1.9.1 dist.h) because the library is not
// vendored; everything else — accumulation order, tier shifts, the LOOP_MARGIN queue, the merge-join and its float /
// double mixing — is the reference's own code.
// TEST INFRASTRUCTURE ONLY.
#include <algorithm>
#include <utility>
#include <vector>

#include "OptimizationBackend/MatrixAccumulators.h"
#include "scale_optimization/ScaleAccumulator.h"

namespace flann {
template <typename T>
struct Matrix {
  T *data;
  size_t rows, cols;
  Matrix() : data(nullptr), rows(0), cols(0) {}
  Matrix(T *d, size_t r, size_t c) : data(d), rows(r), cols(c) {}
  T *operator[](size_t r) const { return data + r * cols; }
};
template <typename T>
struct L2 {
  typedef T ElementType;
  T operator()(const T *a, const T *b, size_t size) const {  // FLANN 1.9.1 flann/algorithms/dist.h L2::operator()
    T result = T();
    T diff0, diff1, diff2, diff3;
    const T *last = a + size;
    const T *lastgroup = last - 3;
    while (a < lastgroup) {
      diff0 = (T)(a[0] - b[0]);
      diff1 = (T)(a[1] - b[1]);
      diff2 = (T)(a[2] - b[2]);
      diff3 = (T)(a[3] - b[3]);
      result += diff0 * diff0 + diff1 * diff1 + diff2 * diff2 + diff3 * diff3;
      a += 4;
      b += 4;
    }
    while (a < last) {
      diff0 = (T)(*a++ - *b++);
      result += diff0 * diff0;
    }
    return result;
  }
};
struct SearchParams {
  explicit SearchParams(int) {}
};
template <typename Distance>
struct Index {
  typedef typename Distance::ElementType T;
  std::vector<std::vector<T>> pts;
  size_t size() const { return pts.size(); }
  void addPoints(const Matrix<T> &m) {
    for (size_t r = 0; r < m.rows; r++) pts.emplace_back(m[r], m[r] + m.cols);
  }
  void knnSearch(const Matrix<T> &q, Matrix<int> &idces, Matrix<T> &dists, int k, const SearchParams &) {
    std::vector<std::pair<T, int>> all;
    Distance d;
    for (size_t i = 0; i < pts.size(); i++) all.emplace_back(d(q[0], pts[i].data(), q.cols), (int)i);
    std::sort(all.begin(), all.end());
    for (int i = 0; i < k; i++) {
      idces[0][i] = i < (int)all.size() ? all[i].second : -1;
      dists[0][i] = i < (int)all.size() ? all[i].first : (T)1e30;
    }
  }
};
}  // namespace flann

typedef std::vector<std::pair<int, double>> SigType;  // src/loop_closure/loop_detection/ScanContext.h:24
namespace dso {
struct LoopFrame {  // only the member search_sc reads (src/loop_closure/LoopHandler.h:75)
  SigType signature;
};
}  // namespace dso

#include "loop_closure/loop_detection/search_place.h"

extern "C" {

// J: 9 arrays of n floats (J0..J7, r) ; w: n floats ; n % 4 == 0.  out45 = upper triangle of acc.H row-major.
void ref_accumulator9(const float *J, const float *w, int n, float *out45) {
  static dso::Accumulator9 acc;
  acc.initialize();
  for (int i = 0; i < n; i += 4)
    acc.updateSSE_eighted(_mm_loadu_ps(J + 0 * n + i), _mm_loadu_ps(J + 1 * n + i), _mm_loadu_ps(J + 2 * n + i), _mm_loadu_ps(J + 3 * n + i),
                          _mm_loadu_ps(J + 4 * n + i), _mm_loadu_ps(J + 5 * n + i), _mm_loadu_ps(J + 6 * n + i), _mm_loadu_ps(J + 7 * n + i),
                          _mm_loadu_ps(J + 8 * n + i), _mm_loadu_ps(w + i));
  acc.finish();
  int e = 0;
  for (int r = 0; r < 9; r++)
    for (int c = r; c < 9; c++) out45[e++] = acc.H(r, c);
}

void ref_scale_accumulator(const float *J, const float *r, const float *w, int n, float *out3) {
  static dso::ScaleAccumulator acc;
  acc.initialize();
  for (int i = 0; i < n; i += 4) acc.updateSSE_oneed(_mm_loadu_ps(J + i), _mm_loadu_ps(r + i), _mm_loadu_ps(w + i));
  acc.finish();
  out3[0] = acc.hessian_(0, 0);
  out3[1] = acc.hessian_(0, 1);
  out3[2] = acc.hessian_(1, 1);
}

// search_sc on CSR signatures (same layout as orc_search_sc)
void ref_search_sc(const int *q_idx, const double *q_val, int q_nnz, const int *sig_ptr, const int *sig_idx, const double *sig_val, int n_frames,
                   const int *candidates, int n_cand, int sc_width, int *res_idx, float *res_diff) {
  std::vector<dso::LoopFrame> frames(n_frames);
  std::vector<dso::LoopFrame *> ptrs(n_frames);
  for (int f = 0; f < n_frames; f++) {
    for (int k = sig_ptr[f]; k < sig_ptr[f + 1]; k++) frames[f].signature.emplace_back(sig_idx[k], sig_val[k]);
    ptrs[f] = &frames[f];
  }
  SigType q;
  for (int k = 0; k < q_nnz; k++) q.emplace_back(q_idx[k], q_val[k]);
  std::vector<int> cand(candidates, candidates + n_cand);
  search_sc(q, ptrs, cand, sc_width, *res_idx, *res_diff);
}

// Feeds n_keys ring keys through search_ringkey in order against an index that starts with the dummy row 0
// (src/loop_closure/LoopHandler.cpp:35-39).  cand_out[n_keys][3] (-1 padded), n_cand_out[n_keys].
// NB search_ringkey keeps its LOOP_MARGIN queue in function-local statics: call this ONCE per process.
void ref_search_ringkey_sequence(const float *keys, int n_keys, int dim, int *cand_out, int *n_cand_out) {
  flann::Index<flann::L2<float>> index;
  std::vector<float> dummy(dim, 0.0f);
  index.addPoints(flann::Matrix<float>(dummy.data(), 1, dim));
  for (int i = 0; i < n_keys; i++) {
    std::vector<float> k(keys + (size_t)i * dim, keys + (size_t)(i + 1) * dim);
    flann::Matrix<float> km(k.data(), 1, dim);
    std::vector<int> cands;
    search_ringkey(km, &index, cands);
    n_cand_out[i] = (int)cands.size();
    for (int j = 0; j < 3; j++) cand_out[3 * i + j] = j < (int)cands.size() ? cands[j] : -1;
  }
}

}  // extern "C"
