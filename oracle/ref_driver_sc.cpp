// ref_driver_sc.cpp — C entry point around the REFERENCE'S OWN ScanContext.cpp (align_points_PCA :19-66, ScanContext::generate
// :78-142), compiled in place by oracle/ref_build.py against the stand-ins of oracle/shim_sc (dynamic-size Eigen pieces,
// SelfAdjointEigenSolver = oracle/jacobi_eig3.h, flann::Matrix).  Nothing of the reference is copied into the repository.
// TEST INFRASTRUCTURE ONLY.  Pins: rotation into the PCA frame, polar binning, max-height, ring key, per-sector L2
// normalisation, tfm_pca_rig — as written in the reference's source text.  Cannot pin: the eigen-solver and Eigen's GEMM order.
#include "loop_closure/loop_detection/ScanContext.cpp"

extern "C" int refsc_generate(const double *pts, int n, double lidar_range, int num_s, int num_r, float *ringkey, int *sig_idx, double *sig_val,
                              double tfm_pca_rig[16]) {
  std::vector<Eigen::Vector3d> cloud;
  cloud.reserve((size_t)n);
  for (int i = 0; i < n; i++) cloud.emplace_back(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
  ScanContext sc(num_s, num_r);
  flann::Matrix<float> rk;
  SigType sig;
  Eigen::Matrix4d tfm;
  sc.generate(cloud, rk, sig, lidar_range, tfm);
  for (int i = 0; i < num_r; i++) ringkey[i] = rk[0][i];
  delete[] rk.data;
  for (size_t i = 0; i < sig.size(); i++) {
    sig_idx[i] = sig[i].first;
    sig_val[i] = sig[i].second;
  }
  for (int i = 0; i < 16; i++) tfm_pca_rig[i] = tfm.d[i];
  return (int)sig.size();
}
