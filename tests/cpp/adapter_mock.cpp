// Compiles include/dslam_b200_adapter.hpp against mock stand-ins of the reference's types (this image has no Eigen /
// Sophus / DSO headers) and, when a GPU is present, runs one tracking + scale call through it.
// Build: g++ -std=c++14 -I include tests/cpp/adapter_mock.cpp -L direct_stereo_slam_b200 -ldslam_b200 -o adapter_mock
#include <cmath>
#include <cstdio>
#include <vector>

#include "dslam_b200_adapter.hpp"

namespace mock {
struct Vector3f { float v[3]; };
struct AffLight { double a = 0, b = 0; };
struct SE3 {  // Sophus::SE3d::data(): quaternion (x,y,z,w) + translation
  double d[7] = {0, 0, 0, 1, 0, 0, 0};
  double *data() { return d; }
};
struct Vec5 { double v[5]; double &operator[](int i) { return v[i]; } const double &operator[](int i) const { return v[i]; } };
struct Vec3 { double v[3]; double &operator[](int i) { return v[i]; } };
struct Vec3d { double v[3]; double operator[](int i) const { return v[i]; } double operator()(int i) const { return v[i]; } };
struct FlannMatrix {  // flann::Matrix<float>
  float *data; size_t rows, cols;
  float *operator[](size_t r) const { return data + r * cols; }
};
typedef std::vector<std::pair<int, double>> SigType;
struct Mat44 {  // Eigen::Matrix4d (column-major storage, (row, col) access)
  double m[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  double &operator()(int r, int c) { return m[c * 4 + r]; }
};
struct FrameShell { int id = 0; };
struct FrameHessian {
  Vector3f *dI = nullptr;
  Vector3f *dIp[6] = {nullptr};
  float *absSquaredGrad[6] = {nullptr};
  float ab_exposure = 1.f;
  FrameShell *shell = nullptr;
  AffLight aff;
  AffLight aff_g2l() const { return aff; }
};
struct CalibHessian {
  float fx, fy, cx, cy;
  float B[256];
  float fxl() const { return fx; } float fyl() const { return fy; } float cxl() const { return cx; } float cyl() const { return cy; }
};
}  // namespace mock

using Tracker = dslam_b200::TrackerAndScaler<mock::FrameHessian, mock::SE3, mock::AffLight, mock::Vec5, mock::Vec3>;

int main() {
  int ndev = 0;
  if (dslam_device_count(&ndev) != DSLAM_OK || ndev < 1) {
    std::printf("adapter_mock: no CUDA device, compile/link check only (%s)\n", dslam_last_error());
    return 0;
  }
  const int w = 320, h = 192, levels = 3;
  dslam_b200::Session session(0);
  dslam_b200::FramePyramids<mock::FrameHessian> frames(session, w, h, levels);
  mock::CalibHessian calib{200.f, 200.f, 159.5f, 95.5f, {0}};
  const float K1[4] = {200.f, 200.f, 159.5f, 95.5f};
  std::vector<double> tfm = {1, 0, 0, -0.3, 0, 1, 0, 0, 0, 0, 1, 1e-9, 0, 0, 0, 1};
  Tracker tracker(session, frames, w, h, levels, tfm, K1);
  tracker.makeK(&calib);
  // a textured plane at 10 m seen by the keyframe, the new frame (shifted 2 px) and the right camera
  auto render = [&](float shift, std::vector<float> &img) {
    img.resize((size_t)w * h);
    for (int y = 0; y < h; y++)
      for (int x = 0; x < w; x++) img[(size_t)y * w + x] = 128.f + 60.f * std::sin(0.11f * (x + shift)) * std::cos(0.07f * y) + 20.f * std::sin(0.31f * (x + shift) + 0.2f * y);
  };
  std::vector<float> img;
  mock::FrameShell shell[3];
  mock::FrameHessian fh[3];
  std::vector<std::vector<float>> store;
  for (int k = 0; k < 3; k++) {
    fh[k].shell = &shell[k];
    for (int l = 0; l < levels; l++) {
      store.emplace_back((size_t)3 * (w >> l) * (h >> l));
      fh[k].dIp[l] = reinterpret_cast<mock::Vector3f *>(store.back().data());
      store.emplace_back((size_t)(w >> l) * (h >> l));
      fh[k].absSquaredGrad[l] = store.back().data();
    }
    render(k == 0 ? 0.f : (k == 1 ? 2.f : 6.f), img);  // disparity f*b/z = 200*0.3/10 = 6 px for the right camera
    if (k == 1) frames.makeImagesOverlapped(&fh[k], img.data(), &calib, false);  // the new frame: built on the pyramid stream
    else frames.makeImages(&fh[k], img.data(), &calib, false);
    frames.wait_host(&fh[k]);
  }
  dslam_b200::ActivePoints pts;
  for (int y = 8; y < h - 8; y += 6)
    for (int x = 8; x < w - 8; x += 6) pts.push(x, y, 0.1f, 1.0f);
  tracker.setCoarseTrackingRef({&fh[0]}, pts);
  mock::SE3 pose;
  mock::AffLight aff;
  mock::Vec5 minres, last;
  for (int i = 0; i < 5; i++) minres[i] = NAN;
  const bool ok = tracker.trackNewestCoarse(&fh[1], pose, aff, levels - 1, minres, last);
  float scale = 1.f;
  const float rmse = tracker.optimizeScale(&fh[2], scale, levels - 1);
  // loop-closure alignment of the keyframe's 3-D points (LoopHandler.cpp:166-178, 274-277) against the same new frame,
  // whose device pyramid was released and is rebuilt from the host mirror
  frames.release(&fh[1]);
  dslam_b200::PoseEstimator<mock::FrameHessian> pose_estimator(session, frames, w, h, levels);
  std::vector<std::pair<mock::Vec3d, float *>> pts_dso;
  std::vector<std::vector<float>> colors;
  for (size_t i = 0; i < pts.u.size(); i++) {
    const double z = 10.0;
    colors.emplace_back(levels);
    for (int l = 0; l < levels; l++) {
      const int ul = (int)((pts.u[i] + 0.5) / (1 << l) - 0.5 + 0.5), vl = (int)((pts.v[i] + 0.5) / (1 << l) - 0.5 + 0.5);
      colors.back()[l] = fh[0].dIp[l][ul + vl * (w >> l)].v[0];
    }
    pts_dso.push_back({mock::Vec3d{{(pts.u[i] - calib.cx) / calib.fx * z, (pts.v[i] - calib.cy) / calib.fy * z, z}}, colors.back().data()});
  }
  mock::Mat44 ref_to_new;
  float pose_error = 0;
  const bool pe_ok = pose_estimator.estimate(pts_dso, 1.f, &fh[1], {calib.fx, calib.fy, calib.cx, calib.cy}, levels - 1, ref_to_new, pose_error);
  std::printf("adapter_mock: pe_ok=%d pe_tx=%.4f (expect about -0.1) pose_error=%.3f inliers=%d%%\n", (int)pe_ok, ref_to_new(0, 3), pose_error,
              pose_estimator.inlier_percent);
  // a 2 px shift of a plane at 10 m with f = 200 is a translation of -0.1 m (the scene moved +x, so the camera moved -x)
  std::printf("adapter_mock: ok=%d tx=%.4f (expect about -0.1) rmse0=%.3f scale=%.3f (expect about 1) scale_rmse=%.3f pc_n0=%d\n", (int)ok,
              pose.data()[4], last[0], scale, rmse, tracker.pc_n()[0]);
  bool pass = ok && std::fabs(pose.data()[4] + 0.1) < 0.01 && std::fabs(scale - 1.f) < 0.05f && std::fabs(ref_to_new(0, 3) + 0.1) < 0.01;

  // makeImages with the reference's allocation behaviour (new[] inside, dI = dIp[0]) + the hypothesis loop in one call
  {
    mock::FrameHessian fa;
    mock::FrameShell sa;
    fa.shell = &sa;
    render(2.f, img);
    frames.makeImages(&fa, img.data(), &calib);
    pass = pass && fa.dI == fa.dIp[0] && fa.dIp[0][5 * w + 7].v[0] == img[5 * w + 7];
    std::vector<mock::SE3> tries(3);
    tries[0].d[4] = 0.5;  // a bad start, then the identity twice
    mock::SE3 best;
    mock::AffLight aff0, aff_best;
    mock::Vec5 last_rmse, achieved;
    for (int i = 0; i < 5; i++) last_rmse[i] = 100.0;
    mock::Vec3 flowv;
    bool have = false;
    const int ntried = tracker.trackNewCoarse(&fa, tries, aff0, levels - 1, last_rmse, 1.5, best, aff_best, achieved, flowv, have);
    std::printf("adapter_mock: trackNewCoarse tried %d of 3, haveOneGood=%d tx=%.4f\n", ntried, (int)have, best.data()[4]);
    pass = pass && have && ntried >= 1 && std::fabs(best.data()[4] + 0.1) < 0.02;
    frames.release(&fa);
    for (int l = 0; l < levels; l++) { delete[] fa.dIp[l]; delete[] fa.absSquaredGrad[l]; }
  }
  // the loop-closure calls one by one, by the reference's names and argument order (LoopHandler.cpp:239-259)
  {
    dslam_b200::Session loop_session(0);  // the LoopHandler thread owns its session
    dslam_b200::LoopDatabase db(loop_session, 8, 60, 20, true);  // grows; keeps the double signature values
    unsigned rs = 12345u;
    auto rnd = [&]() { rs = rs * 1664525u + 1013904223u; return (double)(rs >> 8) / 16777216.0 - 0.5; };
    std::vector<std::vector<mock::Vec3d>> clouds(130);
    int found = -2;
    float diff = 9.f;
    for (int k = 0; k < 130; k++) {
      for (int i = 0; i < 1500; i++) clouds[k].push_back(mock::Vec3d{{60.0 * rnd() * (1 + 0.3 * (k % 7)), 40.0 * rnd(), 6.0 * rnd() * (1 + 0.2 * (k % 5))}});
      if (k == 129) clouds[k] = clouds[10];  // a revisit of keyframe 10, 119 keyframes later (> LOOP_MARGIN)
      float ringkey[20];
      mock::SigType signature;
      mock::Mat44 tfm;
      db.generate(clouds[k], ringkey, signature, 40.0, tfm);
      mock::FlannMatrix rk{ringkey, 1, 20};
      std::vector<int> candidates;
      dslam_b200::search_ringkey(rk, &db, candidates);
      if (k == 129 && !candidates.empty()) {
        int idx = -1;
        dslam_b200::search_sc(signature, db, candidates, 60, idx, diff);
        found = idx;
        float qdiff = 9.f;
        const int qidx = db.query(signature, qdiff);
        pass = pass && qidx == 10;
      } else if (k < 103) {
        pass = pass && candidates.empty();  // nothing is old enough yet
      }
    }
    std::printf("adapter_mock: loop search found keyframe %d (expect 10) diff %.5f, database rows %d\n", found, diff, db.size());
    pass = pass && found == 10 && diff < 1e-5f && db.size() == 130;
  }
  return pass ? 0 : 1;
}
