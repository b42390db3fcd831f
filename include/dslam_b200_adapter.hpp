// dslam_b200_adapter.hpp — header-only C++ glue between the reference's host code and the C ABI of dslam_b200.h.
//
// `dslam_b200::TrackerAndScaler` has the member surface FrontEnd uses on `dso::TrackerAndScaler`
// (src/scale_optimization/TrackerAndScaler.h:34-64): makeK, setCoarseTrackingRef, scaleCoarseDepthL0,
// trackNewestCoarse, optimizeScale and the "pure output" members refFrameID, lastRef, lastRef_aff_g2l,
// lastFlowIndicators, firstCoarseRMSE.  `dslam_b200::FramePyramids` replaces FrameHessian::makeImages
// (deps:dso/src/FullSystem/HessianBlocks.cpp:128-191) and keeps the device twin of every live FrameHessian.
//
// The classes are templates over the reference's own types so that this header compiles against real DSO / Sophus /
// Eigen headers in the reference tree AND against the tiny mock types of tests/cpp/adapter_mock.cpp in this
// repository (which has no Eigen).  Requirements on the types:
//   SE3       : `double* data()` — 7 doubles (qx,qy,qz,qw,tx,ty,tz) like Sophus::SE3d
//   AffLight  : public doubles `a`, `b`                       (deps:dso/src/util/NumType.h:166-192)
//   Vec5/Vec3 : `operator[]`                                   (Eigen fixed-size vectors)
//   FrameHessian : `Eigen::Vector3f* dIp[PYR_LEVELS]`, `float* absSquaredGrad[PYR_LEVELS]`, `float ab_exposure`,
//                  `AffLight aff_g2l()`, `shell->id`
//   CalibHessian : `fxl() fyl() cxl() cyl()`, `float* B` (256-entry inverse response, HessianBlocks.h:329-330)
#pragma once
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "dslam_b200.h"

namespace dslam_b200 {

inline void check(int rc, const char *what) {
  if (rc != DSLAM_OK) throw std::runtime_error(std::string(what) + ": " + dslam_last_error());
}

// One per process (or per tracking thread): owns the CUDA stream all objects below are ordered on.
class Session {
 public:
  explicit Session(int device = 0) { check(dslam_session_create(device, &s_), "dslam_session_create"); }
  ~Session() { dslam_session_destroy(s_); }
  Session(const Session &) = delete;
  Session &operator=(const Session &) = delete;
  dslam_session *get() const { return s_; }

 private:
  dslam_session *s_ = nullptr;
};

// Device twins of the image pyramids of live FrameHessians, keyed by the FrameHessian pointer.
template <class FrameHessian>
class FramePyramids {
 public:
  FramePyramids(Session &s, int w, int h, int levels) : s_(s), w_(w), h_(h), levels_(levels) {}
  ~FramePyramids() {
    for (auto &kv : live_) dslam_frame_destroy(kv.second);
    for (dslam_frame *f : free_) dslam_frame_destroy(f);
  }
  // Drop-in for `fh->makeImages(color, HCalib)`: the caller has already allocated fh->dIp[l] / fh->absSquaredGrad[l]
  // (ideally with dslam_host_alloc so the mirror copy is a true DMA).  The copies land asynchronously; call
  // wait_host(fh) before untouched host code (traceOn, pixel selector, BA) reads them.
  template <class CalibHessian>
  void makeImages(FrameHessian *fh, const float *color, CalibHessian *HCalib, bool gamma_weights) {
    dslam_frame *f = acquire(fh);
    float *dIp[DSLAM_MAX_LEVELS] = {nullptr}, *ag[DSLAM_MAX_LEVELS] = {nullptr};
    for (int l = 0; l < levels_; l++) {
      dIp[l] = reinterpret_cast<float *>(fh->dIp[l]);
      ag[l] = fh->absSquaredGrad[l];
    }
    const float *B = (HCalib != nullptr && gamma_weights) ? HCalib->B : nullptr;  // HessianBlocks.cpp:182-188
    check(dslam_frame_make_images(f, color, B, dIp, ag), "dslam_frame_make_images");
  }
  // The same, but the pyramid is built on the session's pyramid stream (dslam_frame_build_batch, stage_host bit 2): call it
  // for the image that has just arrived BEFORE tracking the previous frame — the build then runs beside those LM rounds —
  // and every later call that takes `fh` is ordered behind it automatically.
  template <class CalibHessian>
  void makeImagesOverlapped(FrameHessian *fh, const float *color, CalibHessian *HCalib, bool gamma_weights) {
    dslam_frame *f = acquire(fh);
    float *dIp[DSLAM_MAX_LEVELS] = {nullptr}, *ag[DSLAM_MAX_LEVELS] = {nullptr};
    for (int l = 0; l < levels_; l++) {
      dIp[l] = reinterpret_cast<float *>(fh->dIp[l]);
      ag[l] = fh->absSquaredGrad[l];
    }
    const float *B = (HCalib != nullptr && gamma_weights) ? HCalib->B : nullptr;
    check(dslam_frame_upload(f, color), "dslam_frame_upload");
    check(dslam_frame_build_batch(1, &f, B, 1 | 2 | 4), "dslam_frame_build_batch");
    check(dslam_frame_download(f, dIp, ag), "dslam_frame_download");
  }
  void wait_host(FrameHessian *fh) { check(dslam_frame_wait_host(at(fh)), "dslam_frame_wait_host"); }
  // call from FrameHessian::~FrameHessian / FrameHessian::release
  void release(FrameHessian *fh) {
    auto it = live_.find(fh);
    if (it == live_.end()) return;
    free_.push_back(it->second);
    live_.erase(it);
  }
  dslam_frame *at(FrameHessian *fh) const {
    auto it = live_.find(fh);
    if (it == live_.end()) throw std::runtime_error("FrameHessian has no device pyramid (makeImages not called)");
    return it->second;
  }
  // Device pyramid of a frame whose twin may already have been released (a keyframe the LoopHandler thread kept after the
  // front end marginalised it): rebuilt from the intensity channel of the host mirror fh->dIp[0] when it is gone.
  dslam_frame *ensure(FrameHessian *fh) {
    auto it = live_.find(fh);
    if (it != live_.end()) return it->second;
    dslam_frame *f = acquire(fh);
    std::vector<float> color((size_t)w_ * h_);
    const float *src = reinterpret_cast<const float *>(fh->dIp[0]);
    for (size_t i = 0; i < color.size(); i++) color[i] = src[3 * i];
    check(dslam_frame_make_images(f, color.data(), nullptr, nullptr, nullptr), "dslam_frame_make_images");
    return f;
  }

 private:
  dslam_frame *acquire(FrameHessian *fh) {
    auto it = live_.find(fh);
    if (it != live_.end()) return it->second;
    dslam_frame *f = nullptr;
    if (!free_.empty()) {
      f = free_.back();
      free_.pop_back();
    } else {
      check(dslam_frame_create(s_.get(), w_, h_, levels_, &f), "dslam_frame_create");
    }
    live_[fh] = f;
    return f;
  }
  Session &s_;
  int w_, h_, levels_;
  std::unordered_map<FrameHessian *, dslam_frame *> live_;
  std::vector<dslam_frame *> free_;
};

// Flat export of the active points that makeCoarseDepthL0 iterates over (TrackerAndScaler.cpp:149-166): the maintainer
// fills it from frameHessians[*]->pointHessians (u = int(centerProjectedTo[0] + 0.5f), v likewise,
// idepth = centerProjectedTo[2], weight = sqrtf(1e-3 / (efPoint->HdiF + 1e-12))).
struct ActivePoints {
  std::vector<int> u, v;
  std::vector<float> idepth, weight;
  void clear() { u.clear(); v.clear(); idepth.clear(); weight.clear(); }
  void push(int uu, int vv, float id, float w) { u.push_back(uu); v.push_back(vv); idepth.push_back(id); weight.push_back(w); }
};

template <class FrameHessian, class SE3, class AffLight, class Vec5, class Vec3>
class TrackerAndScaler {
 public:
  // TrackerAndScaler(int w, int h, const std::vector<double>& tfm_vec, const Mat33f& K1)  (:47-109); K1 as (fx,fy,cx,cy)
  TrackerAndScaler(Session &s, FramePyramids<FrameHessian> &frames, int w, int h, int levels, const std::vector<double> &tfm_vec,
                   const float K1[4])
      : lastRef_aff_g2l(), frames_(frames) {
    if (tfm_vec.size() != 16) throw std::invalid_argument("tfm_vec must hold a row-major 4x4");
    check(dslam_ctx_create(s.get(), w, h, levels, K1, K1, tfm_vec.data(), &c_), "dslam_ctx_create");
  }
  ~TrackerAndScaler() { dslam_ctx_destroy(c_); }
  TrackerAndScaler(const TrackerAndScaler &) = delete;
  TrackerAndScaler &operator=(const TrackerAndScaler &) = delete;

  // makeK(CalibHessian*)  (:117-141)
  template <class CalibHessian>
  void makeK(CalibHessian *HCalib) {
    const float K0[4] = {HCalib->fxl(), HCalib->fyl(), HCalib->cxl(), HCalib->cyl()};
    check(dslam_ctx_make_K(c_, K0), "dslam_ctx_make_K");
  }
  void setAffineOptModes(int modeA, int modeB) { check(dslam_ctx_set_affine_mode(c_, modeA, modeB), "dslam_ctx_set_affine_mode"); }

  // setCoarseTrackingRef(std::vector<FrameHessian*>)  (:317-327): lastRef = frameHessians.back(); template built on the device
  void setCoarseTrackingRef(const std::vector<FrameHessian *> &frameHessians, const ActivePoints &pts) {
    lastRef = frameHessians.back();
    check(dslam_ref_build(c_, frames_.at(lastRef), (int)pts.u.size(), pts.u.data(), pts.v.data(), pts.idepth.data(), pts.weight.data(), pc_n_),
          "dslam_ref_build");
    refFrameID = lastRef->shell->id;
    lastRef_aff_g2l = lastRef->aff_g2l();
    check(dslam_ref_set_affine(c_, lastRef->ab_exposure, lastRef_aff_g2l.a, lastRef_aff_g2l.b), "dslam_ref_set_affine");
    firstCoarseRMSE = -1;
  }
  // scaleCoarseDepthL0(float)  (:329-336)
  void scaleCoarseDepthL0(float scale) { check(dslam_ref_scale_idepth(c_, scale), "dslam_ref_scale_idepth"); }

  // bool trackNewestCoarse(FrameHessian*, SE3&, AffLight&, int, Vec5, Vec5&, Output3DWrapper* = 0)  (:451-638)
  bool trackNewestCoarse(FrameHessian *newFrameHessian, SE3 &lastToNew_out, AffLight &aff_g2l_out, int coarsestLvl, Vec5 minResForAbort,
                         Vec5 &lastResiduals, void * /*wrap*/ = nullptr) {
    double pose[7], aff[2] = {aff_g2l_out.a, aff_g2l_out.b}, minres[5], last[5], flow[3];
    for (int i = 0; i < 7; i++) pose[i] = lastToNew_out.data()[i];
    for (int i = 0; i < 5; i++) minres[i] = minResForAbort[i];
    int ok = 0;
    check(dslam_track_newest_coarse(c_, frames_.at(newFrameHessian), newFrameHessian->ab_exposure, pose, aff, coarsestLvl, minres, last, flow, &ok),
          "dslam_track_newest_coarse");
    for (int i = 0; i < 7; i++) lastToNew_out.data()[i] = pose[i];  // untouched by the library when the level loop aborted
    aff_g2l_out.a = aff[0];
    aff_g2l_out.b = aff[1];
    for (int i = 0; i < 5; i++) lastResiduals[i] = last[i];
    for (int i = 0; i < 3; i++) lastFlowIndicators[i] = flow[i];
    return ok != 0;
  }

  // The retry loop of FrontEnd::trackNewCoarse (src/FrontEnd.cpp:147-247) in one call: all hypotheses advance in lock step.
  // Returns per-hypothesis results; the caller applies the reference's acceptance rule in order.
  void trackNewestCoarseMulti(FrameHessian *newFrameHessian, std::vector<SE3> &poses, std::vector<AffLight> &affs, int coarsestLvl,
                              Vec5 minResForAbort, std::vector<double> &lastResiduals5, std::vector<double> &flow3, std::vector<int> &ok) {
    const int n = (int)poses.size();
    std::vector<double> p(7 * n), a(2 * n);
    double minres[5];
    for (int i = 0; i < 5; i++) minres[i] = minResForAbort[i];
    for (int k = 0; k < n; k++) {
      for (int i = 0; i < 7; i++) p[7 * k + i] = poses[k].data()[i];
      a[2 * k] = affs[k].a;
      a[2 * k + 1] = affs[k].b;
    }
    lastResiduals5.assign(5 * n, 0.0);
    flow3.assign(3 * n, 0.0);
    ok.assign(n, 0);
    check(dslam_track_newest_coarse_multi(c_, frames_.at(newFrameHessian), newFrameHessian->ab_exposure, n, p.data(), a.data(), coarsestLvl, minres,
                                          lastResiduals5.data(), flow3.data(), ok.data()),
          "dslam_track_newest_coarse_multi");
    for (int k = 0; k < n; k++) {
      for (int i = 0; i < 7; i++) poses[k].data()[i] = p[7 * k + i];
      affs[k].a = a[2 * k];
      affs[k].b = a[2 * k + 1];
    }
  }

  // float optimizeScale(FrameHessian* fh1, float& scale, int coarsestLvl)  (:854-964)
  float optimizeScale(FrameHessian *fh1, float &scale, int coarsestLvl) {
    float rmse = 0;
    check(dslam_optimize_scale(c_, frames_.at(fh1), &scale, coarsestLvl, &rmse), "dslam_optimize_scale");
    return rmse;
  }
  // the seed loop of FrontEnd::optimizeScale (src/FrontEnd.cpp:995-1003) in one lock step
  void optimizeScaleSeeds(FrameHessian *fh1, std::vector<float> &scales, int coarsestLvl, std::vector<float> &rmse) {
    rmse.assign(scales.size(), 0.f);
    check(dslam_optimize_scale_multi(c_, frames_.at(fh1), (int)scales.size(), scales.data(), coarsestLvl, rmse.data()), "dslam_optimize_scale_multi");
  }

  // "act as pure output" (TrackerAndScaler.h:59-64)
  int refFrameID = -1;
  FrameHessian *lastRef = nullptr;
  AffLight lastRef_aff_g2l;
  Vec3 lastFlowIndicators;
  double firstCoarseRMSE = -1;

  const int *pc_n() const { return pc_n_; }
  dslam_ctx *handle() const { return c_; }

 private:
  FramePyramids<FrameHessian> &frames_;
  dslam_ctx *c_ = nullptr;
  int pc_n_[DSLAM_MAX_LEVELS] = {0};
};

// dso::PoseEstimator (src/loop_closure/pose_estimation/PoseEstimator.h:34-49) with the member surface LoopHandler uses
// (src/loop_closure/LoopHandler.cpp:41, 274-277).  Vector3d: `operator[]`; Matrix4d: `operator()(row, col)`.
template <class FrameHessian>
class PoseEstimator {
 public:
  PoseEstimator(Session &s, FramePyramids<FrameHessian> &frames, int w, int h, int levels) : frames_(frames), levels_(levels) {
    check(dslam_pe_create(s.get(), w, h, levels, &p_), "dslam_pe_create");
  }
  ~PoseEstimator() { dslam_pe_destroy(p_); }
  PoseEstimator(const PoseEstimator &) = delete;
  PoseEstimator &operator=(const PoseEstimator &) = delete;
  void setAffineOptMode(int modeA, int modeB) { check(dslam_pe_set_affine_mode(p_, modeA, modeB), "dslam_pe_set_affine_mode"); }

  template <class Vector3d, class Matrix4d>
  bool estimate(const std::vector<std::pair<Vector3d, float *>> &pts, float ref_ab_exposure, FrameHessian *new_fh, const std::vector<float> &new_cam,
                int coarsest_lvl, Matrix4d &ref_to_new, float &pose_error) {
    xyz_.resize(3 * pts.size());
    colors_.resize((size_t)levels_ * pts.size());
    for (size_t i = 0; i < pts.size(); i++) {
      for (int k = 0; k < 3; k++) xyz_[3 * i + k] = pts[i].first[k];
      for (int l = 0; l < levels_; l++) colors_[i * levels_ + l] = pts[i].second[l];
    }
    check(dslam_pe_set_points(p_, (int)pts.size(), xyz_.data(), colors_.data(), ref_ab_exposure), "dslam_pe_set_points");
    double T[16];
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) T[r * 4 + c] = ref_to_new(r, c);
    int ok = 0;
    check(dslam_pe_estimate(p_, frames_.ensure(new_fh), new_fh->ab_exposure, new_cam.data(), coarsest_lvl, T, &pose_error, &inlier_percent, &ok),
          "dslam_pe_estimate");
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) ref_to_new(r, c) = T[r * 4 + c];
    return ok != 0;
  }
  int inlier_percent = 0;  // 100 * lastInners[0] / pts.size() of the last call (:480)

 private:
  FramePyramids<FrameHessian> &frames_;
  int levels_;
  dslam_pe *p_ = nullptr;
  std::vector<double> xyz_;
  std::vector<float> colors_;
};

// search_ringkey + search_sc of src/loop_closure/loop_detection/search_place.h behind one database object.
class LoopDatabase {
 public:
  static constexpr int kLoopMargin = 100;   // LOOP_MARGIN  search_place.h:22
  static constexpr int kFlannNN = 3;        // FLANN_NN     :21
  static constexpr float kRingkeyThres() { return 0.1f; }  // RINGKEY_THRES :23
  LoopDatabase(Session &s, int capacity, int n_sectors = 60, int n_rings = 20) : n_rings_(n_rings), n_cells_(n_sectors * n_rings) {
    check(dslam_sc_create(s.get(), n_sectors, n_rings, capacity, &db_), "dslam_sc_create");
  }
  ~LoopDatabase() { dslam_sc_destroy(db_); }
  // One LoopHandler::run iteration (src/loop_closure/LoopHandler.cpp:236-264): the new keyframe's descriptor is appended
  // at once under id = number of descriptors so far; only ids < id - LOOP_MARGIN compete (the reference's delay queue).
  // signature: sparse (index, value) pairs as ScanContext::generate produces them (SigType).
  // Returns the matched id or -1; res_diff like search_sc.
  int addAndSearch(const float *ringkey, const std::vector<std::pair<int, double>> &signature, float &res_diff) {
    std::vector<int> idx(signature.size());
    std::vector<double> val(signature.size());
    std::vector<float> dense((size_t)n_cells_, 0.f);
    for (size_t i = 0; i < signature.size(); i++) {
      idx[i] = signature[i].first;
      val[i] = signature[i].second;
      dense[(size_t)idx[i]] = (float)val[i];
    }
    const int id = count_;
    check(dslam_sc_add_sparse(db_, ringkey, idx.data(), val.data(), (int)idx.size(), id), "dslam_sc_add_sparse");
    count_++;
    res_diff = 1.1f;
    const int max_id = id - kLoopMargin;
    if (max_id < kFlannNN) return -1;  // "ringkeys->size() > FLANN_NN" with the dummy row (:28)
    int cand[kFlannNN];
    float dist[kFlannNN];
    check(dslam_sc_search_ringkey(db_, 1, ringkey, kFlannNN, kRingkeyThres(), max_id, cand, dist), "dslam_sc_search_ringkey");
    if (cand[0] < 0) return -1;  // no candidate: LoopHandler skips search_sc (:249-252)
    int res_idx = -1;
    check(dslam_sc_search_sc(db_, 1, dense.data(), cand, kFlannNN, &res_idx, &res_diff), "dslam_sc_search_sc");
    return res_idx;
  }
  dslam_scdb *handle() const { return db_; }

 private:
  dslam_scdb *db_ = nullptr;
  int n_rings_, n_cells_, count_ = 0;
};

}  // namespace dslam_b200
