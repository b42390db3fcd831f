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
//
// Threading contract.  A dslam_session (and everything created from it) is used by ONE host thread at a time.  The reference
// runs two threads that touch this path: the tracking thread (FrontEnd: makeImages, trackNewCoarse, optimizeScale,
// setCoarseTrackingRef) and the LoopHandler thread (ScanContext::generate, search_ringkey, search_sc, PoseEstimator::estimate).
// Give each thread its OWN Session and its own FramePyramids; objects of different sessions share nothing.  The loop thread's
// PoseEstimator never keeps device twins of frames: estimate() rebuilds the pyramid of `cur_frame->fh` from its host mirror
// (fh->dIp[0], which that thread owns) and releases the twin before it returns, so a FrameHessian address that is reused after
// `delete cur_frame->fh` can never hit a stale device pyramid.  On the tracking thread call FramePyramids::release(fh) from
// FrameHessian::~FrameHessian (or wherever the front end deletes a frame) — that is the only hook needed there.
#pragma once
#include <stdexcept>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <utility>
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
  // `fh->makeImages(color, HCalib)` with the reference's own allocation behaviour (HessianBlocks.cpp:131-136): dIp[l] and
  // absSquaredGrad[l] are new[]-ed here, dI = dIp[0], and FrameHessian::~FrameHessian delete[]s them as it always did.  The
  // mirrors are complete when this returns (pageable memory: the copy is staged by the driver; allocate the arrays with
  // dslam_host_alloc and use the overload above to get true DMA + overlap).
  template <class CalibHessian>
  void makeImages(FrameHessian *fh, const float *color, CalibHessian *HCalib) {
    typedef typename std::remove_pointer<typename std::decay<decltype(fh->dIp[0])>::type>::type Texel;  // Eigen::Vector3f
    static_assert(sizeof(Texel) == 3 * sizeof(float), "dIp texels must be 3 packed floats");
    for (int l = 0; l < levels_; l++) {
      const size_t n = (size_t)(w_ >> l) * (h_ >> l);
      fh->dIp[l] = new Texel[n];
      fh->absSquaredGrad[l] = new float[n];
    }
    fh->dI = fh->dIp[0];
    makeImages(fh, color, HCalib, HCalib != nullptr);  // gamma weights whenever a calibration is passed (setting_gammaWeightsPixelSelect == 1)
    wait_host(fh);
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
  // Device pyramid of a frame this object has no (trustworthy) twin of — a keyframe the LoopHandler thread kept after the front
  // end released it, seen from the loop thread's own FramePyramids: ALWAYS rebuilt from the intensity channel of the host
  // mirror fh->dIp[0] (never a cached twin: the address of a deleted FrameHessian may have been reused).  Pair with release().
  dslam_frame *rebuild(FrameHessian *fh) {
    dslam_frame *f = acquire(fh);
    color_.resize((size_t)w_ * h_);
    const float *src = reinterpret_cast<const float *>(fh->dIp[0]);
    for (size_t i = 0; i < color_.size(); i++) color_[i] = src[3 * i];
    check(dslam_frame_make_images(f, color_.data(), nullptr, nullptr, nullptr), "dslam_frame_make_images");
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
  std::vector<float> color_;
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

  // The whole hypothesis loop of FrontEnd::trackNewCoarse (src/FrontEnd.cpp:192-247) in one call: `tries` in the reference's
  // order, `aff_last_2_l` the start affine of every try, `last_coarse_rmse` / `reTrackThreshold` of the break test (:244-246).
  // Outputs exactly what the sequential loop leaves behind: lastF_2_fh, aff_g2l, achievedRes, flowVecs (via lastFlowIndicators
  // semantics: returned in flowVecs), haveOneGood; returns tryIterations.  The hypotheses are evaluated speculatively in
  // lock-step batches (1, then 4, then the rest) and the acceptance rule is replayed in order (dslam_track_new_coarse).
  template <class SE3Vector>
  int trackNewCoarse(FrameHessian *fh, const SE3Vector &tries, const AffLight &aff_last_2_l, int coarsestLvl, const Vec5 &last_coarse_rmse,
                     double reTrackThreshold, SE3 &lastF_2_fh, AffLight &aff_g2l, Vec5 &achievedRes, Vec3 &flowVecs, bool &haveOneGood) {
    const int n = (int)tries.size();
    std::vector<double> p(7 * (size_t)n);
    for (int k = 0; k < n; k++) {
      SE3 t = tries[k];
      for (int i = 0; i < 7; i++) p[7 * (size_t)k + i] = t.data()[i];
    }
    const double aff0[2] = {aff_last_2_l.a, aff_last_2_l.b};
    double last[5], pose[7], aff[2], ach[5], flow[3];
    for (int i = 0; i < 5; i++) last[i] = last_coarse_rmse[i];
    int good = 0, ntry = 0;
    check(dslam_track_new_coarse(c_, frames_.at(fh), fh->ab_exposure, n, p.data(), aff0, coarsestLvl, last, reTrackThreshold, pose, aff, ach, flow, &good,
                                 &ntry),
          "dslam_track_new_coarse");
    for (int i = 0; i < 7; i++) lastF_2_fh.data()[i] = pose[i];
    aff_g2l.a = aff[0];
    aff_g2l.b = aff[1];
    for (int i = 0; i < 5; i++) achievedRes[i] = ach[i];
    for (int i = 0; i < 3; i++) flowVecs[i] = flow[i];
    haveOneGood = good != 0;
    return ntry;
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
  // `s` and `frames` are the LOOP THREAD's own Session / FramePyramids (not the tracking thread's)
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
    dslam_frame *twin = frames_.rebuild(new_fh);  // from the host mirror; see the threading contract at the top of this header
    const int rc = dslam_pe_estimate(p_, twin, new_fh->ab_exposure, new_cam.data(), coarsest_lvl, T, &pose_error, &inlier_percent, &ok);
    frames_.release(new_fh);
    check(rc, "dslam_pe_estimate");
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
  // capacity = initial allocation (the tables grow like the reference's index does); fp64 keeps the double signature values
  LoopDatabase(Session &s, int capacity, int n_sectors = 60, int n_rings = 20, bool fp64 = false)
      : n_rings_(n_rings), n_cells_(n_sectors * n_rings), fp64_(fp64) {
    check(dslam_sc_create_ex(s.get(), n_sectors, n_rings, capacity, fp64 ? DSLAM_SC_FP64 : 0, &db_), "dslam_sc_create_ex");
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
  // ---- the three calls of LoopHandler::run one by one (src/loop_closure/LoopHandler.cpp:239-259) ---------------------------
  // ScanContext::generate (ScanContext.cpp:78-142) on the device; the descriptor is appended to the database under
  // id = number of descriptors so far (device to device) and returned in the reference's forms: ringkey[n_rings] floats, sparse
  // signature (SigType), tfm_pca_rig.  Vector3d: operator()(i); Matrix4d: operator()(r, c).
  template <class Vector3d, class SigType, class Matrix4d>
  void generate(const std::vector<Vector3d> &pts_spherical, float *ringkey, SigType &signature, double lidar_range, Matrix4d &tfm_pca_rig) {
    xyz_.resize(3 * pts_spherical.size());
    for (size_t i = 0; i < pts_spherical.size(); i++)
      for (int k = 0; k < 3; k++) xyz_[3 * i + k] = pts_spherical[i](k);
    sig64_.resize((size_t)n_cells_);
    double T[16];
    check(dslam_sc_generate(db_, xyz_.data(), (int)pts_spherical.size(), lidar_range, ringkey, nullptr, sig64_.data(), T, 1, count_), "dslam_sc_generate");
    count_++;
    signature.clear();
    for (int i = 0; i < n_cells_; i++)
      if (sig64_[(size_t)i] != 0.0) signature.push_back({i, sig64_[(size_t)i]});
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) tfm_pca_rig(r, c) = T[r * 4 + c];
  }
  // search_ringkey (search_place.h:25-57) for the descriptor appended LAST (by generate / addAndSearch): exact 3-NN among the
  // ids older than LOOP_MARGIN keyframes (the reference's delay queue), kept when dist < RINGKEY_THRES.  Differences from the
  // FLANN call it replaces: exact instead of a randomised kd-tree with 128 checks, and the three neighbours are always real rows
  // (FLANN's three may include the uninitialised dummy row 0 the reference's index starts with, which it then discards).
  void search_ringkey(const float *ringkey, std::vector<int> &candidates) {
    const int max_id = (count_ - 1) - kLoopMargin;
    if (max_id < kFlannNN) return;  // "ringkeys->size() > FLANN_NN" with the dummy row (:28)
    int cand[kFlannNN];
    float dist[kFlannNN];
    check(dslam_sc_search_ringkey(db_, 1, ringkey, kFlannNN, kRingkeyThres(), max_id, cand, dist), "dslam_sc_search_ringkey");
    for (int i = 0; i < kFlannNN; i++)
      if (cand[i] >= 0) candidates.emplace_back(cand[i]);
  }
  // search_sc (search_place.h:59-85): the candidates' signatures are the database rows (the reference reads
  // loop_frames[candidate]->signature); float += double * double in cell order, strict '>' running minimum from 1.1.
  // On a database created with fp64 = true the doubles of `signature` are used unrounded.
  template <class SigType>
  void search_sc(const SigType &signature, const std::vector<int> &candidates, int /*sc_width*/, int &res_idx, float &res_diff) {
    if (fp64_) {
      sig64_.assign((size_t)n_cells_, 0.0);
      for (size_t i = 0; i < signature.size(); i++) sig64_[(size_t)signature[i].first] = signature[i].second;
      check(dslam_sc_search_sc64(db_, 1, sig64_.data(), candidates.data(), (int)candidates.size(), &res_idx, &res_diff), "dslam_sc_search_sc64");
    } else {
      dense_.assign((size_t)n_cells_, 0.f);
      for (size_t i = 0; i < signature.size(); i++) dense_[(size_t)signature[i].first] = (float)signature[i].second;
      check(dslam_sc_search_sc(db_, 1, dense_.data(), candidates.data(), (int)candidates.size(), &res_idx, &res_diff), "dslam_sc_search_sc");
    }
  }
  // exhaustive alternative to the two-stage search: sector-cosine scan of every row older than LOOP_MARGIN keyframes
  // (dslam_sc_query; sharded over the GPUs of the box when a communicator is attached)
  template <class SigType>
  int query(const SigType &signature, float &res_diff) {
    dense_.assign((size_t)n_cells_, 0.f);
    for (size_t i = 0; i < signature.size(); i++) dense_[(size_t)signature[i].first] = (float)signature[i].second;
    int res_idx = -1;
    check(dslam_sc_query(db_, 1, nullptr, dense_.data(), -1.0f, (count_ - 1) - kLoopMargin, &res_idx, &res_diff), "dslam_sc_query");
    return res_idx;
  }
  int size() const { return count_; }
  dslam_scdb *handle() const { return db_; }

 private:
  dslam_scdb *db_ = nullptr;
  int n_rings_, n_cells_, count_ = 0;
  bool fp64_ = false;
  std::vector<double> xyz_, sig64_;
  std::vector<float> dense_;
};

// The reference's free functions by their own names and argument order (search_place.h:25-27, 59-63), with the FLANN index /
// the loop-frame vector replaced by the LoopDatabase that holds the descriptors:
//   search_ringkey(ringkey, ringkeys_, candidates)                      ->  search_ringkey(ringkey, &loop_db, candidates)
//   search_sc(signature, loop_frames_, candidates, width, idx, diff)    ->  search_sc(signature, loop_db, candidates, width, idx, diff)
// FlannMatrix: `operator[](row)` -> float* (flann::Matrix<float>).
template <class FlannMatrix>
inline void search_ringkey(const FlannMatrix &ringkey, LoopDatabase *ringkeys, std::vector<int> &candidates) {
  ringkeys->search_ringkey(ringkey[0], candidates);
}
template <class SigType>
inline void search_sc(SigType &signature, LoopDatabase &loop_frames, const std::vector<int> &candidates, int sc_width, int &res_idx, float &res_diff) {
  loop_frames.search_sc(signature, candidates, sc_width, res_idx, res_diff);
}

}  // namespace dslam_b200
