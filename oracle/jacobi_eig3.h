// jacobi_eig3.h — the 3x3 symmetric eigen-solver shared by the oracle's ScanContext::generate restatement (sc_generate.cpp) and
// by the SelfAdjointEigenSolver stand-in the reference's own ScanContext.cpp is compiled against (shim_sc/Eigen/Eigenvalues).
// TEST INFRASTRUCTURE ONLY.  Cyclic Jacobi, eigenvalues ascending (the order Eigen returns), every eigenvector normalised and
// sign-fixed so that its largest-magnitude component is positive.  Eigen's own tridiagonal-QR iteration (absent from this image)
// returns the same vectors up to sign and ~1e-16: that difference is what stays unpinned in row f-2.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>

inline void dslam_jacobi_eig3(const double Ain[9], double evals[3], double evecs[9] /* columns = eigenvectors, row-major 3x3 */) {
  double A[9];
  std::memcpy(A, Ain, sizeof(A));
  double V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int sweep = 0; sweep < 64; sweep++) {
    const double off = A[1] * A[1] + A[2] * A[2] + A[5] * A[5];
    if (off < 1e-300) break;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        const double apq = A[p * 3 + q];
        if (std::fabs(apq) < 1e-300) continue;
        const double theta = (A[q * 3 + q] - A[p * 3 + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; k++) {  // A <- A * J
          const double akp = A[k * 3 + p], akq = A[k * 3 + q];
          A[k * 3 + p] = c * akp - s * akq;
          A[k * 3 + q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; k++) {  // A <- J^T * A
          const double apk = A[p * 3 + k], aqk = A[q * 3 + k];
          A[p * 3 + k] = c * apk - s * aqk;
          A[q * 3 + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; k++) {
          const double vkp = V[k * 3 + p], vkq = V[k * 3 + q];
          V[k * 3 + p] = c * vkp - s * vkq;
          V[k * 3 + q] = s * vkp + c * vkq;
        }
      }
  }
  int order[3] = {0, 1, 2};
  std::sort(order, order + 3, [&](int a, int b) { return A[a * 3 + a] < A[b * 3 + b]; });
  for (int j = 0; j < 3; j++) {
    const int src = order[j];
    evals[j] = A[src * 3 + src];
    double v[3] = {V[0 * 3 + src], V[1 * 3 + src], V[2 * 3 + src]};
    const double nrm = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    int big = 0;
    for (int k = 1; k < 3; k++)
      if (std::fabs(v[k]) > std::fabs(v[big])) big = k;
    const double sgn = (v[big] < 0 ? -1.0 : 1.0) / nrm;
    for (int k = 0; k < 3; k++) evecs[k * 3 + j] = v[k] * sgn;
  }
}
