// sc_generate.cpp — CPU restatement of ScanContext::generate (descriptor generation).
//
// TEST INFRASTRUCTURE ONLY (same rules as dslam_oracle.cpp).
//
// Follows src/loop_closure/loop_detection/ScanContext.cpp:19-66 (align_points_PCA) and :78-142 (generate).
// PARITY STATUS: parity unpinned.  The reference uses Eigen::SelfAdjointEigenSolver<MatrixXd> (:42-46), whose
// eigenvector SIGNS are an artefact of its tridiagonal-QR iteration; Eigen is absent from this image.  Here the
// 3x3 symmetric eigenproblem is solved by cyclic Jacobi, eigenvalues ascending (as Eigen orders them), and each
// eigenvector is sign-normalised so that its largest-magnitude component is positive.  Everything after the
// eigenvectors (polar binning, max-height, ring key, per-sector L2 normalisation) is restated op for op.

#include <cmath>
#include <cstring>
#include <vector>
#include <algorithm>

namespace {

void jacobi_eig3(const double Ain[9], double evals[3], double evecs[9] /* columns = eigenvectors, row-major 3x3 */) {
  double A[9]; std::memcpy(A, Ain, sizeof(A));
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
          A[k * 3 + p] = c * akp - s * akq; A[k * 3 + q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; k++) {  // A <- J^T * A
          const double apk = A[p * 3 + k], aqk = A[q * 3 + k];
          A[p * 3 + k] = c * apk - s * aqk; A[q * 3 + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; k++) {
          const double vkp = V[k * 3 + p], vkq = V[k * 3 + q];
          V[k * 3 + p] = c * vkp - s * vkq; V[k * 3 + q] = s * vkp + c * vkq;
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
    int big = 0; for (int k = 1; k < 3; k++) if (std::fabs(v[k]) > std::fabs(v[big])) big = k;
    const double sgn = (v[big] < 0 ? -1.0 : 1.0) / nrm;
    for (int k = 0; k < 3; k++) evecs[k * 3 + j] = v[k] * sgn;
  }
}

}  // namespace

extern "C" {

// pts: n x 3 doubles (pts_spherical).  Outputs: ringkey[num_r] (float), sparse signature (sig_idx ascending,
// sig_val double, capacity num_s*num_r), tfm_pca_rig row-major 4x4.  Returns nnz.
int orc_sc_generate(const double *pts, int n, double lidar_range, int num_s, int num_r, float *ringkey, int *sig_idx, double *sig_val, double tfm_pca_rig[16]) {
  // align_points_PCA :19-66
  double mx = 0, my = 0, mz = 0;
  for (int i = 0; i < n; i++) { mx += pts[3 * i]; my += pts[3 * i + 1]; mz += pts[3 * i + 2]; }
  mx /= n; my /= n; mz /= n;
  std::vector<double> pm(3 * (size_t)n);
  for (int i = 0; i < n; i++) { pm[3 * i] = pts[3 * i] - mx; pm[3 * i + 1] = pts[3 * i + 1] - my; pm[3 * i + 2] = pts[3 * i + 2] - mz; }
  double cov[9] = {0};
  for (int i = 0; i < n; i++)
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) cov[a * 3 + b] += pm[3 * i + a] * pm[3 * i + b];
  double evals[3], ev[9];
  jacobi_eig3(cov, evals, ev);
  std::vector<double> al(3 * (size_t)n);  // nx, ny, nz = pts_mat * v0, v1, v2
  for (int i = 0; i < n; i++)
    for (int j = 0; j < 3; j++) al[3 * i + j] = pm[3 * i] * ev[0 * 3 + j] + pm[3 * i + 1] * ev[1 * 3 + j] + pm[3 * i + 2] * ev[2 * 3 + j];
  for (int i = 0; i < 16; i++) tfm_pca_rig[i] = (i % 5 == 0) ? 1.0 : 0.0;
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) tfm_pca_rig[r * 4 + c] = ev[c * 3 + r];  // rows = v_r^T
  for (int r = 0; r < 3; r++) tfm_pca_rig[r * 4 + 3] = -(tfm_pca_rig[r * 4] * mx + tfm_pca_rig[r * 4 + 1] * my + tfm_pca_rig[r * 4 + 2] * mz);

  // generate :78-142
  for (int i = 0; i < num_r; i++) ringkey[i] = 0.0;
  std::vector<double> max_height((size_t)num_s * num_r, -lidar_range - 1.0);
  for (int i = 0; i < n; i++) {
    const double yp = al[3 * i + 1], zp = al[3 * i + 2];
    const double rho = std::sqrt(yp * yp + zp * zp);
    double theta = std::atan2(zp, yp);
    while (theta < 0) theta += 2.0 * M_PI;
    while (theta >= 2.0 * M_PI) theta -= 2.0 * M_PI;
    const int si = theta / (2.0 * M_PI) * num_s;
    const int ri = rho / lidar_range * num_r;
    if (ri >= num_r) continue;
    if (si >= num_s) continue;  // reference asserts; cannot happen for theta < 2pi
    max_height[si * num_r + ri] = std::max(max_height[si * num_r + ri], al[3 * i]);
  }
  std::vector<double> sig_norm_si(num_s, 0.0);
  int nnz = 0;
  for (int i = 0; i < num_s * num_r; i++) {
    if (max_height[i] >= (-lidar_range)) {
      ringkey[i % num_r]++;
      sig_idx[nnz] = i; sig_val[nnz] = max_height[i]; nnz++;
      sig_norm_si[i / num_r] += max_height[i] * max_height[i];
    }
  }
  for (int i = 0; i < num_r; i++) ringkey[i] /= num_s;
  for (int i = 0; i < num_s; i++) sig_norm_si[i] = std::sqrt(sig_norm_si[i]);
  for (int i = 0; i < nnz; i++) sig_val[i] /= sig_norm_si[sig_idx[i] / num_r];
  return nnz;
}

}  // extern "C"
