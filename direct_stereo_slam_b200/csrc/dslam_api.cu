// dslam_api.cu — sessions, frame pyramids and the tracker object behind the C ABI of include/dslam_b200.h.
//
// Host side of the photometric Gauss-Newton path.  What stays on the host is exactly what the north star
// leaves there: the Levenberg-Marquardt control flow of TrackerAndScaler::trackNewestCoarse
// (src/scale_optimization/TrackerAndScaler.cpp:451-638) and ::optimizeScale (:854-964), the damped 8x8 solve
// and the SE3 update.  Everything per-point runs in the fused kernels of kernels_residual.cu; a Levenberg-
// Marquardt round is ONE kernel launch whose result lands in mapped pinned memory (no memcpy, no stream
// synchronise), and several independent starts (pose hypotheses, scale seeds) advance in lock step inside the
// same launch.
//
// There is no CPU fallback in this file: without a CUDA device every entry point fails with DSLAM_ENODEVICE.

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <new>

#include "dslam_internal.h"
#include <limits>

namespace dslam {

// ---------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
int cuda_fail(cudaError_t e, const char *what) {
  snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInitializationError) return DSLAM_ENODEVICE;
  if (e == cudaErrorMemoryAllocation) return DSLAM_ENOMEM;
  return DSLAM_ECUDA;
}

namespace {

constexpr float kCoarseCutoffTH = 20.0f;  // setting_coarseCutoffTH  deps:dso/src/util/settings.cpp:138
constexpr float kHuberTH = 9.0f;          // setting_huberTH         deps:dso/src/util/settings.cpp:127
constexpr double kScaleXiRot = 1.0, kScaleXiTrans = 0.5, kScaleA = 10.0, kScaleB = 1000.0;  // HessianBlocks.h:59-65
const int kMaxIterations[dslam::kMaxLevels] = {10, 20, 50, 50, 50, 50};  // :463 (+50 for an optional 6th level)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---------------------------------------------------------------------------------------------------
// driver entry point for tensor maps (no link-time dependency on libcuda)
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// ---------------------------------------------------------------------------------------------------
// evaluation plumbing
// ---------------------------------------------------------------------------------------------------
struct EvalOut {           // one calcRes* + calcGSSSE* evaluation, host view
  double acc[kPoseVals];   // raw sums
  int nE, nSat, nInl;
  int n_padded;            // pose_buf_warped_n_ / scale_buf_warped_n_
  double res6[6];          // the Vec6 of calcResPose / calcResScale
};

// CTAs per item of one launch.  Every item would like one template point per thread (shortest critical path); when the
// launch as a whole would exceed one resident wave (96 registers x 128 threads -> 5 CTAs per SM) the wave is dealt to the
// items in proportion to their size, so that the level-0 items of a mixed round are not held to the share of the tiny
// coarse-level items next to them.  Fills nblocks / ppt_stride / cta_begin; returns the size of the flat grid.
// `lanes` = launches of this kind in flight at the same time (groups x lanes of a lock step): a launch that shares the GPU
// takes a smaller slice of the resident wave — measured on B200 with 8 concurrent lanes: 2-3 CTAs per SM per launch beat 5
// by 8 % and 8 by 17 % (the CTAs of concurrent launches queue behind each other).
int assign_blocks(EvalItem *items, int cnt, int num_sms, int lanes) {
  long want_total = 0;
  for (int i = 0; i < cnt; i++) {
    int per = (items[i].n + kEvalThreads - 1) / kEvalThreads;
    if (per < 1) per = 1;
    if (per > kMaxBlocksPerItem) per = kMaxBlocksPerItem;
    items[i].nblocks = per;
    want_total += per;
  }
  static const int per_sm_env = [] { const char *e = getenv("DSLAM_EVAL_CTAS_PER_SM"); const int v = e ? atoi(e) : 0; return v >= 1 && v <= 16 ? v : 0; }();
  int per_sm = lanes <= 2 ? 5 : (12 + lanes - 1) / lanes;
  if (per_sm < 2) per_sm = 2;
  if (per_sm_env) per_sm = per_sm_env;
  const long budget = (long)num_sms * per_sm;
  int total = 0;
  for (int i = 0; i < cnt; i++) {
    EvalItem &it = items[i];
    if (want_total > budget) {
      int per = (int)((long)it.nblocks * budget / want_total);
      it.nblocks = per < 1 ? 1 : per;
    }
    it.ppt_stride = it.nblocks * kEvalThreads;
    it.cta_begin = total;
    total += it.nblocks;
  }
  return total;
}

// Launch `n` prepared items (result slots [slot0, slot0 + n)) on `stream`; returns the sequence number through *seq_out.
int launch_items(dslam_session *s, std::vector<EvalItem> &items, int slot0, cudaStream_t stream, EvalScratch scratch, unsigned *seq_out,
                 long long *launch_counter, int lanes = 1) {
  const int n = (int)items.size();
  // kernel flavour: 0 = all pose, 1 = all scale, 2 = mixed (bit 1 of EvalItem::flags marks a scale item)
  int n_scale = 0;
  for (const EvalItem &it : items) n_scale += (it.flags & 2) ? 1 : 0;
  const int mode = n_scale == 0 ? 0 : (n_scale == n ? 1 : 2);
  const unsigned seq = ++s->seq;  // atomic: group threads launch concurrently
  *seq_out = seq;
  EvalBatch batch;
  for (int base = 0; base < n; base += kMaxItemsPerLaunch) {
    const int cnt = n - base < kMaxItemsPerLaunch ? n - base : kMaxItemsPerLaunch;
    long long pts = 0;
    const int gx = assign_blocks(&items[base], cnt, s->num_sms, lanes);
    for (int i = 0; i < cnt; i++) {
      batch.item[i] = items[base + i];
      pts += items[base + i].n;
    }
    const bool prof = s->prof_on;
    size_t slot = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (prof) {
      std::lock_guard<std::mutex> lock(s->prof_mutex);
      slot = s->prof_used++;
      while (s->prof_ev.size() < 2 * (slot + 1)) {
        cudaEvent_t e;
        DSLAM_CUDA(cudaEventCreate(&e));
        s->prof_ev.push_back(e);
      }
      if (s->prof_mode.size() <= slot) { s->prof_mode.resize(slot + 1); s->prof_points.resize(slot + 1); }
      s->prof_mode[slot] = mode;
      s->prof_points[slot] = pts;
      ev0 = s->prof_ev[2 * slot];
      ev1 = s->prof_ev[2 * slot + 1];
    }
    if (prof) DSLAM_CUDA(cudaEventRecord(ev0, stream));
    DSLAM_CUDA(launch_eval(mode, batch, cnt, gx, scratch, s->results_dev + slot0 + base, seq, stream));
    if (prof) DSLAM_CUDA(cudaEventRecord(ev1, stream));
    s->launches++;
    if (launch_counter) (*launch_counter)++;
  }
  return DSLAM_OK;
}

// Wait for the result records of a launch group.  The last CTA of every item writes its record into mapped pinned
// memory as self-validating words (payload | sequence number).
int collect_items(dslam_session *s, const std::vector<EvalItem> &items, int slot0, cudaStream_t stream, unsigned seq, std::vector<EvalOut> &outs) {
  const int n = (int)items.size();
  outs.resize(n);
  const auto t0 = std::chrono::steady_clock::now();
  unsigned long spins = 0;
  for (int i = 0; i < n; i++) {
    const bool is_scale = (items[i].flags & 2) != 0;
    const int nv = is_scale ? kScaleVals : kPoseVals;
    const volatile unsigned long long *w = s->results_host[slot0 + i].w;
    EvalOut &o = outs[i];
    auto wait_word = [&](int k, unsigned long long *out) -> int {
      unsigned long long v;
      while ((unsigned)((v = w[k]) & 0xffffffffull) != seq) {
        if ((++spins & 0x3fff) == 0) {
          const cudaError_t q = cudaStreamQuery(stream);
          if (q != cudaSuccess && q != cudaErrorNotReady) return cuda_fail(q, "evaluation kernel");
          const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
          if (dt > s->timeout_s) return fail(DSLAM_ETIMEOUT, "evaluation kernel did not publish its result within %.1f s", s->timeout_s);
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
      }
      *out = v;
      return DSLAM_OK;
    };
    int cnts[3];
    for (int k = 2; k >= 0; k--) {  // the counters are written last by the kernel: wait for them first
      unsigned long long v = 0;
      const int rc = wait_word(kResultCountBase + k, &v);
      if (rc != DSLAM_OK) return rc;
      cnts[k] = (int)(unsigned)(v >> 32);
    }
    for (int k = 0; k < nv; k++) {
      unsigned long long hi = 0, lo = 0;
      int rc = wait_word(2 * k, &hi);
      if (rc != DSLAM_OK) return rc;
      rc = wait_word(2 * k + 1, &lo);
      if (rc != DSLAM_OK) return rc;
      const unsigned long long bits = (hi & 0xffffffff00000000ull) | (lo >> 32);
      double d;
      std::memcpy(&d, &bits, 8);
      o.acc[k] = d;
    }
    o.nE = cnts[0];
    o.nSat = cnts[1];
    o.nInl = cnts[2];
    o.n_padded = (o.nInl + 3) & ~3;  // zero padding to a multiple of 4  (:824-835, :1147-1158)
    const EvalItem &it = items[i];
    const int iE = is_scale ? 3 : 45, iT = is_scale ? 4 : 46, iRT = is_scale ? 5 : 47;
    const float shiftNum = (it.flags & 1) ? 2.0f * (float)((it.n + 31) / 32) : 0.0f;  // sumSquaredShiftNum
    o.res6[0] = o.acc[iE];
    o.res6[1] = o.nE;
    o.res6[2] = o.acc[iT] / (shiftNum + 0.1);
    o.res6[3] = 0;
    o.res6[4] = o.acc[iRT] / (shiftNum + 0.1);
    o.res6[5] = o.nSat / (float)o.nE;  // :851 float division
  }
  return DSLAM_OK;
}

int run_items(dslam_session *s, std::vector<EvalItem> &items, std::vector<EvalOut> &outs, long long *launch_counter) {
  const int n = (int)items.size();
  if (n < 1) return DSLAM_OK;
  if (n > kResultSlots) return fail(DSLAM_EINVAL, "too many evaluation items in one round (%d > %d)", n, kResultSlots);
  unsigned seq = 0;
  const int rc = launch_items(s, items, 0, s->stream, s->scratch, &seq, launch_counter);
  if (rc != DSLAM_OK) return rc;
  return collect_items(s, items, 0, s->stream, seq, outs);
}

void fill_common(EvalItem &it, const dslam_ctx *c, const dslam_frame *f, int lvl, float cutoff) {
  it.tex = f->L.tex[lvl];
  it.pts = c->pts[lvl];
  it.n = c->pc_n[lvl];
  it.w = c->w[lvl];
  it.h = c->h[lvl];
  it.flags = lvl == 0 ? 1 : 0;
  for (int k = 0; k < 9; k++) it.Ki[k] = c->Ki[lvl][k];
  it.cutoff = cutoff;
  it.maxEnergy = 2 * kHuberTH * cutoff - kHuberTH * kHuberTH;  // :726-728
  it.nblocks = 1;
  it.ppt_stride = kEvalThreads;
  it.cta_begin = 0;
}

// calcResPose prologue :711-720
void fill_pose_item(EvalItem &it, const dslam_ctx *c, const dslam_frame *f, float new_exposure, int lvl, const hm::Se3 &refToNew,
                    const double aff[2], float cutoff) {
  fill_common(it, c, f, lvl, cutoff);
  it.fx = c->cam0.fx[lvl]; it.fy = c->cam0.fy[lvl]; it.cx = c->cam0.cx[lvl]; it.cy = c->cam0.cy[lvl];
  double Rd[9];
  hm::qmatrix(refToNew.q, Rd);
  float Rf[9];
  for (int k = 0; k < 9; k++) Rf[k] = (float)Rd[k];
  if (c->point3d) {  // PoseEstimator::calcRes :160-161: R and t as floats, no K^-1 (the records are 3-D points)
    for (int k = 0; k < 9; k++) it.M[k] = Rf[k];
    it.flags |= 4;
  } else {
    hm::mat33f_mul(Rf, c->Ki[lvl], it.M);
  }
  for (int k = 0; k < 3; k++) it.t[k] = (float)refToNew.t[k];
  double affd[2];
  hm::aff_from_to(c->ref_exposure, new_exposure, c->ref_a, c->ref_b, aff[0], aff[1], affd);
  it.p0 = (float)affd[0];
  it.p1 = (float)affd[1];
  it.p2 = (float)c->ref_b;  // b0 of calcGSSSEPose :649
}

// calcResScale prologue :1019-1027
void fill_scale_item(EvalItem &it, const dslam_ctx *c, const dslam_frame *f, int lvl, float scale, float cutoff) {
  fill_common(it, c, f, lvl, cutoff);
  it.fx = c->cam1.fx[lvl]; it.fy = c->cam1.fy[lvl]; it.cx = c->cam1.cx[lvl]; it.cy = c->cam1.cy[lvl];
  for (int k = 0; k < 9; k++) it.M[k] = c->M_stereo[lvl][k];
  for (int k = 0; k < 3; k++) it.t[k] = (float)c->T_f1_f0.t[k];
  it.p0 = scale;
  it.p1 = it.p2 = 0.f;
  it.flags |= 2;  // scale item
}

// calcGSSSEPose epilogue :682-696 from the raw sums
void pose_normal_equations(const EvalOut &o, double H[64], double b[8]) {
  const float invn = 1.0f / o.n_padded;
  double Hf[81];
  int e = 0;
  for (int r = 0; r < 9; r++)
    for (int cc = r; cc < 9; cc++, e++) Hf[r * 9 + cc] = Hf[cc * 9 + r] = o.acc[e];
  for (int r = 0; r < 8; r++) {
    for (int cc = 0; cc < 8; cc++) H[r * 8 + cc] = Hf[r * 9 + cc] * invn;
    b[r] = Hf[r * 9 + 8] * invn;
  }
  const double sc[8] = {kScaleXiRot, kScaleXiRot, kScaleXiRot, kScaleXiTrans, kScaleXiTrans, kScaleXiTrans, kScaleA, kScaleB};
  for (int r = 0; r < 8; r++)
    for (int cc = 0; cc < 8; cc++) H[r * 8 + cc] *= sc[cc];
  for (int r = 0; r < 8; r++)
    for (int cc = 0; cc < 8; cc++) H[r * 8 + cc] *= sc[r];
  for (int r = 0; r < 8; r++) b[r] *= sc[r];
}

// calcGSSSEScale epilogue :1003-1004 (hessian_ is a float matrix)
void scale_normal_equations(const EvalOut &o, float *H, float *b) {
  const float invn = 1.0f / o.n_padded;
  *H = (float)o.acc[0] * invn;
  *b = (float)o.acc[1] * invn;
}

void trace_row(std::vector<double> *tr, int lvl, int it, int accept, int n, double lambda, double e_old, double e_new, const double *inc,
               int ninc) {
  if (!tr) return;
  const size_t o = tr->size();
  tr->resize(o + 15, 0.0);
  double *r = tr->data() + o;
  r[0] = lvl; r[1] = it; r[2] = accept; r[3] = n; r[4] = lambda; r[5] = e_old; r[6] = e_new;
  for (int k = 0; k < ninc && k < 8; k++) r[7 + k] = inc[k];
}

// ---------------------------------------------------------------------------------------------------
// trackNewestCoarse as a resumable state machine: request() says which evaluation it needs next, consume()
// advances with the result.  Several machines run in lock step (one launch per round for all of them).
// ---------------------------------------------------------------------------------------------------
struct PoseLM {
  // configuration
  const dslam_ctx *c = nullptr;
  const dslam_frame *f = nullptr;
  float new_exposure = 1.f;
  double minResForAbort[5]{};
  std::vector<double> *trace = nullptr;
  // state
  enum Phase { START, ITER, DONE } phase = START;
  hm::Se3 cur, cand;
  double aff_cur[2]{}, aff_cand[2]{};
  int lvl = 0, iteration = 0;
  bool haveRepeated = false, last_rejected = false;
  float levelCutoffRepeat = 1, lambda = 0.01f;
  double H[64]{}, b[8]{}, inc[8]{};
  EvalOut resOld{};
  // outputs
  bool ok = false;
  double lastResiduals[5]{};
  double flow[3]{};
  long long iters = 0;
  struct LevelEvent {  // what the reference's level epilogue (:595-599) saw, in order (a level may appear twice: :601-604)
    int lvl;
    double residual;
    double flow[3];
    int numTermsInE;  // (int)resOld[1]: lastInners[lvl] of PoseEstimator::estimate :455
  };
  std::vector<LevelEvent> events;

  void begin(const dslam_ctx *c_, const dslam_frame *f_, float exposure, const double pose7[7], const double aff[2], int coarsestLvl,
             const double minRes[5], std::vector<double> *tr) {
    c = c_; f = f_; new_exposure = exposure; trace = tr;
    cur = hm::Se3::from7(pose7);
    aff_cur[0] = aff[0]; aff_cur[1] = aff[1];
    for (int i = 0; i < 5; i++) { lastResiduals[i] = NAN; minResForAbort[i] = minRes[i]; }
    flow[0] = flow[1] = flow[2] = 1000;  // lastFlowIndicators.setConstant(1000)  :460
    lvl = coarsestLvl;
    haveRepeated = false;
    events.clear();
    start_level();
  }
  void start_level() {
    levelCutoffRepeat = 1;
    phase = START;
    last_rejected = false;
  }
  // Speculation over a run of rejected steps.  After a rejection the next candidate depends only on (H, b, cur, lambda, iteration)
  // — not on the result that was just rejected — so the driver can ask for the candidates of the next rejections in advance and
  // evaluate them in the SAME round; consume() then walks the results in order and stops at the first one whose request no
  // longer matches.  simulate_reject() is consume()'s rejection branch without a result (:583-590); snap / restore undo it.
  struct Snap {
    float lambda;
    int iteration;
    double inc[8], aff_cand[2];
    hm::Se3 cand;
  };
  Snap snap() const {
    Snap sn;
    sn.lambda = lambda; sn.iteration = iteration; sn.cand = cand;
    std::memcpy(sn.inc, inc, sizeof(inc));
    sn.aff_cand[0] = aff_cand[0]; sn.aff_cand[1] = aff_cand[1];
    return sn;
  }
  void restore(const Snap &sn) {
    lambda = sn.lambda; iteration = sn.iteration; cand = sn.cand;
    std::memcpy(inc, sn.inc, sizeof(inc));
    aff_cand[0] = sn.aff_cand[0]; aff_cand[1] = sn.aff_cand[1];
  }
  bool simulate_reject() {  // true: a further candidate of the same level exists (request() returns it)
    if (phase != ITER) return false;
    lambda *= 4;
    if (lambda < 0.001f) lambda = 0.001f;
    double nrm = 0;
    for (int i = 0; i < 8; i++) nrm += inc[i] * inc[i];
    nrm = std::sqrt(nrm);
    if (!(nrm > 1e-3)) return false;  // the level would end here (:588)
    if (iteration + 1 >= kMaxIterations[lvl]) return false;
    iteration++;
    prepare_iteration();
    return true;
  }
  void request(EvalItem &it) const {
    const float cutoff = kCoarseCutoffTH * levelCutoffRepeat;
    if (phase == START) fill_pose_item(it, c, f, new_exposure, lvl, cur, aff_cur, cutoff);
    else fill_pose_item(it, c, f, new_exposure, lvl, cand, aff_cand, cutoff);
  }
  void consume(const EvalOut &o) {
    if (phase == START) {
      resOld = o;
      if (resOld.res6[5] > 0.6 && levelCutoffRepeat < 50) {  // :475-485
        levelCutoffRepeat *= 2;
        return;
      }
      pose_normal_equations(o, H, b);  // :487
      lambda = 0.01f;
      iteration = 0;
      trace_row(trace, lvl, -1, 1, o.n_padded, lambda, 0.0, resOld.res6[0] / resOld.res6[1], nullptr, 0);
      prepare_iteration();
      return;
    }
    // ITER: resNew = o
    iters++;
    const bool accept = (o.res6[0] / o.res6[1]) < (resOld.res6[0] / resOld.res6[1]);  // :559
    last_rejected = !accept;
    trace_row(trace, lvl, iteration, accept, o.n_padded, lambda, resOld.res6[0] / resOld.res6[1], o.res6[0] / o.res6[1], inc, 8);
    if (accept) {  // :576-582
      pose_normal_equations(o, H, b);
      resOld = o;
      aff_cur[0] = aff_cand[0]; aff_cur[1] = aff_cand[1];
      cur = cand;
      lambda *= 0.5;
    } else {
      lambda *= 4;
      if (lambda < 0.001f) lambda = 0.001f;
    }
    double nrm = 0;
    for (int i = 0; i < 8; i++) nrm += inc[i] * inc[i];
    nrm = std::sqrt(nrm);
    if (!(nrm > 1e-3)) {  // :588
      finish_level();
      return;
    }
    iteration++;
    prepare_iteration();
  }
  void prepare_iteration() {
    if (iteration >= kMaxIterations[lvl]) {
      finish_level();
      return;
    }
    const float lambdaExtrapolationLimit = 0.001f;
    double Hl[64];
    std::memcpy(Hl, H, sizeof(Hl));
    for (int i = 0; i < 8; i++) Hl[i * 8 + i] *= (1 + lambda);
    double nb[8];
    for (int i = 0; i < 8; i++) nb[i] = -b[i];
    hm::ldlt_solve(8, Hl, 8, nb, inc);  // :509
    const int mA = c->affModeA, mB = c->affModeB;
    if (mA < 0 && mB < 0) {  // fix a, b  :511-515
      hm::ldlt_solve(6, Hl, 8, nb, inc);
      inc[6] = inc[7] = 0;
    }
    if (!(mA < 0) && mB < 0) {  // fix b  :516-520
      hm::ldlt_solve(7, Hl, 8, nb, inc);
      inc[7] = 0;
    }
    if (mA < 0 && !(mB < 0)) {  // fix a  :521-534
      double Hs[64], bs[8], is[8];
      std::memcpy(Hs, Hl, sizeof(Hs));
      std::memcpy(bs, nb, sizeof(bs));
      for (int i = 0; i < 8; i++) Hs[i * 8 + 6] = Hs[i * 8 + 7];
      for (int j = 0; j < 8; j++) Hs[6 * 8 + j] = Hs[7 * 8 + j];
      bs[6] = bs[7];
      hm::ldlt_solve(7, Hs, 8, bs, is);
      for (int i = 0; i < 6; i++) inc[i] = is[i];
      inc[6] = 0;
      inc[7] = is[6];
    }
    float extrapFac = 1;
    if (lambda < lambdaExtrapolationLimit) extrapFac = std::sqrt(std::sqrt(lambdaExtrapolationLimit / lambda));  // :536-539
    for (int i = 0; i < 8; i++) inc[i] *= extrapFac;
    double incScaled[8];
    for (int i = 0; i < 3; i++) incScaled[i] = inc[i] * kScaleXiRot;
    for (int i = 3; i < 6; i++) incScaled[i] = inc[i] * kScaleXiTrans;
    incScaled[6] = inc[6] * kScaleA;
    incScaled[7] = inc[7] * kScaleB;
    double sum = 0;
    for (int i = 0; i < 8; i++) sum += incScaled[i];
    if (!std::isfinite(sum))
      for (int i = 0; i < 8; i++) incScaled[i] = 0;  // :547-548
    cand = hm::se3_exp(incScaled) * cur;             // :550-551
    aff_cand[0] = aff_cur[0] + incScaled[6];
    aff_cand[1] = aff_cur[1] + incScaled[7];
    phase = ITER;
  }
  void finish_level() {
    lastResiduals[lvl] = sqrtf((float)(resOld.res6[0] / resOld.res6[1]));  // :595
    flow[0] = resOld.res6[2]; flow[1] = resOld.res6[3]; flow[2] = resOld.res6[4];
    events.push_back(LevelEvent{lvl, lastResiduals[lvl], {flow[0], flow[1], flow[2]}, (int)resOld.res6[1]});
    if (lastResiduals[lvl] > 1.5 * minResForAbort[lvl]) {  // :597-598
      ok = false;
      phase = DONE;
      return;
    }
    if (levelCutoffRepeat > 1 && !haveRepeated) {  // :601-604
      lvl++;
      haveRepeated = true;
    }
    lvl--;
    if (lvl >= 0) {
      start_level();
      return;
    }
    phase = DONE;
    ok = true;
  }
};

struct ScaleLM {
  const dslam_ctx *c = nullptr;
  const dslam_frame *f = nullptr;
  std::vector<double> *trace = nullptr;
  enum Phase { START, ITER, DONE } phase = START;
  float scale_current = 1.f, scale_new = 1.f, inc = 0.f;
  int lvl = 0, iteration = 0;
  bool haveRepeated = false, last_rejected = false;
  float levelCutoffRepeat = 1, lambda = 0.01f, H = 0.f, b = 0.f;
  EvalOut resOld{};
  float last_residuals[5]{};
  long long iters = 0;

  void begin(const dslam_ctx *c_, const dslam_frame *f_, float scale, int coarsestLvl, std::vector<double> *tr) {
    c = c_; f = f_; trace = tr;
    scale_current = scale;
    for (int i = 0; i < 5; i++) last_residuals[i] = NAN;
    lvl = coarsestLvl;
    haveRepeated = false;
    levelCutoffRepeat = 1;
    phase = START;
  }
  void request(EvalItem &it) const {
    fill_scale_item(it, c, f, lvl, phase == START ? scale_current : scale_new, kCoarseCutoffTH * levelCutoffRepeat);
  }
  // speculation over a run of rejected steps, as in PoseLM (the rejection branch of :926-945 without a result)
  struct Snap {
    float lambda, inc, scale_new;
    int iteration;
  };
  Snap snap() const { return Snap{lambda, inc, scale_new, iteration}; }
  void restore(const Snap &sn) { lambda = sn.lambda; inc = sn.inc; scale_new = sn.scale_new; iteration = sn.iteration; }
  bool simulate_reject() {
    if (phase != ITER) return false;
    lambda *= 4;
    if (lambda < 0.001f) lambda = 0.001f;
    if (!(inc > 1e-3)) return false;  // :937: the level would end here
    if (iteration + 1 >= kMaxIterations[lvl]) return false;
    iteration++;
    prepare_iteration();
    return true;
  }
  void consume(const EvalOut &o) {
    if (phase == START) {
      resOld = o;
      if (resOld.res6[5] > 0.6 && levelCutoffRepeat < 50) {  // :873-881
        levelCutoffRepeat *= 2;
        return;
      }
      scale_normal_equations(o, &H, &b);  // :883
      lambda = 0.01f;
      iteration = 0;
      const double t[2] = {0.0, scale_current};
      trace_row(trace, lvl, -1, 1, o.n_padded, lambda, 0.0, resOld.res6[0] / resOld.res6[1], t, 2);
      prepare_iteration();
      return;
    }
    iters++;
    const bool accept = (o.res6[0] / o.res6[1]) < (resOld.res6[0] / resOld.res6[1]);  // :915
    last_rejected = !accept;
    const double t[2] = {inc, scale_new};
    trace_row(trace, lvl, iteration, accept, o.n_padded, lambda, resOld.res6[0] / resOld.res6[1], o.res6[0] / o.res6[1], t, 2);
    if (accept) {  // :926-936
      scale_normal_equations(o, &H, &b);
      resOld = o;
      scale_current = scale_new;
      lambda *= 0.5;
    } else {
      lambda *= 4;
      if (lambda < 0.001f) lambda = 0.001f;
    }
    if (!(inc > 1e-3)) {  // :937 (signed: any non-positive step ends the level)
      finish_level();
      return;
    }
    iteration++;
    prepare_iteration();
  }
  void prepare_iteration() {
    if (iteration >= kMaxIterations[lvl]) {
      finish_level();
      return;
    }
    const float lambdaExtrapolationLimit = 0.001f;
    float Hl = H;
    Hl *= (1 + lambda);
    inc = -b / Hl;  // :898-900
    float extrapFac = 1;
    if (lambda < lambdaExtrapolationLimit) extrapFac = std::sqrt(std::sqrt(lambdaExtrapolationLimit / lambda));
    inc *= extrapFac;
    if (!std::isfinite(inc) || std::fabs(inc) > scale_current) inc = 0.0;  // :907
    scale_new = scale_current + inc;
    phase = ITER;
  }
  void finish_level() {
    last_residuals[lvl] = sqrtf((float)(resOld.res6[0] / resOld.res6[1]));  // :946
    if (levelCutoffRepeat > 1 && !haveRepeated) {
      lvl++;
      haveRepeated = true;
    }
    lvl--;
    if (lvl >= 0) {
      levelCutoffRepeat = 1;
      phase = START;
      last_rejected = false;
      return;
    }
    phase = DONE;
  }
};

// ---- resident evaluation server, host side -------------------------------------------------------------------------------
int server_ensure_buffers(dslam_session *s) {
  if (s->srv_doors) return DSLAM_OK;
  DSLAM_CUDA(cudaHostAlloc((void **)&s->srv_doors, sizeof(ServerDoor) * kSrvLanes, cudaHostAllocMapped));
  std::memset(s->srv_doors, 0, sizeof(ServerDoor) * kSrvLanes);
  DSLAM_CUDA(cudaHostGetDevicePointer((void **)&s->srv_doors_dev, s->srv_doors, 0));
  DSLAM_CUDA(cudaHostAlloc((void **)&s->srv_items, sizeof(EvalItem) * kSrvLanes * kMaxItemsPerLaunch, cudaHostAllocMapped));
  DSLAM_CUDA(cudaHostGetDevicePointer((void **)&s->srv_items_alias, s->srv_items, 0));
  DSLAM_CUDA(cudaHostAlloc((void **)&s->srv_units, sizeof(unsigned) * kSrvLanes * kSrvMaxUnits, cudaHostAllocMapped));
  DSLAM_CUDA(cudaHostGetDevicePointer((void **)&s->srv_units_alias, s->srv_units, 0));
  DSLAM_CUDA(cudaMalloc((void **)&s->srv_queue, sizeof(unsigned long long) * kSrvQueueSlots));
  DSLAM_CUDA(cudaMemsetAsync(s->srv_queue, 0, sizeof(unsigned long long) * kSrvQueueSlots, s->stream));
  DSLAM_CUDA(cudaMalloc((void **)&s->srv_qctl, 256));
  DSLAM_CUDA(cudaMalloc((void **)&s->srv_lane_seq, sizeof(unsigned) * kSrvLanes));
  DSLAM_CUDA(cudaMalloc((void **)&s->srv_items_dev, sizeof(EvalItem) * kSrvLanes * kMaxItemsPerLaunch));
  return DSLAM_OK;
}

struct ServerLaneHost {  // where lane `id` of a lock step keeps its scratch and its slice of the result ring
  EvalScratch scratch;
  int slot0;
};

// Launch the server on the session stream (behind everything queued there: pyramids, template uploads).
int server_start(dslam_session *s, int nlanes, const ServerLaneHost *lanes) {
  int rc = server_ensure_buffers(s);
  if (rc != DSLAM_OK) return rc;
  ServerParams P;
  std::memset(&P, 0, sizeof(P));
  P.nlanes = nlanes;
  P.gen = ++s->srv_gen;
  if ((P.gen & 0xffu) == 0) P.gen = ++s->srv_gen;  // the low byte tags queue slots; zero is what an untouched slot holds
  P.doors = s->srv_doors_dev;
  for (int l = 0; l < nlanes; l++) {
    P.lane[l].partials = lanes[l].scratch.partials;
    P.lane[l].counters = lanes[l].scratch.counters;
    P.lane[l].results = s->results_dev + lanes[l].slot0;
    P.lane[l].items_host = s->srv_items_alias + (size_t)l * kMaxItemsPerLaunch;
    P.lane[l].units_host = s->srv_units_alias + (size_t)l * kSrvMaxUnits;
  }
  for (int l = 0; l < kSrvLanes; l++) P.last_round[l] = s->srv_round[l];
  P.queue = s->srv_queue;
  P.qctl = s->srv_qctl;
  P.items_dev = s->srv_items_dev;
  P.lane_seq = s->srv_lane_seq;
  P.idle_ns = 2ull * 1000 * 1000 * 1000;   // no doorbell for 2 s: the host is gone
  P.life_ns = 30ull * 1000 * 1000 * 1000;  // hard cap on the life of one server (an LM call times out on the host after timeout_s = 20 s)
  DSLAM_CUDA(cudaMemsetAsync(s->srv_qctl, 0, 256, s->stream));
  DSLAM_CUDA(cudaMemsetAsync(s->srv_queue, 0, sizeof(unsigned long long) * kSrvQueueSlots, s->stream));  // no stale units, whatever their tag
  DSLAM_CUDA(launch_eval_server(P, s->num_sms * s->srv_ctas_per_sm, s->stream));
  s->launches++;
  return DSLAM_OK;
}

// One LM round of a lane: items + work units into the lane's mapped buffers, then the doorbell.
int server_post(dslam_session *s, int lane, std::vector<EvalItem> &items, unsigned seq, int lanes_in_flight) {
  const int n = (int)items.size();
  if (n < 1 || n > kMaxItemsPerLaunch) return fail(DSLAM_EINVAL, "server round with %d items", n);
  int total = assign_blocks(items.data(), n, s->num_sms, lanes_in_flight);
  if (total > kSrvMaxUnits) {  // shrink proportionally: a lane round is at most kSrvMaxUnits work units
    int t2 = 0;
    for (int i = 0; i < n; i++) {
      int per = (int)((long)items[i].nblocks * kSrvMaxUnits / total);
      items[i].nblocks = per < 1 ? 1 : per;
      t2 += items[i].nblocks;
    }
    while (t2 > kSrvMaxUnits) {  // rounding up of the small items: take the excess from the largest ones
      int big = 0;
      for (int i = 1; i < n; i++)
        if (items[i].nblocks > items[big].nblocks) big = i;
      items[big].nblocks--;
      t2--;
    }
    total = 0;
    for (int i = 0; i < n; i++) {
      items[i].ppt_stride = items[i].nblocks * kEvalThreads;
      items[i].cta_begin = total;
      total += items[i].nblocks;
    }
  }
  EvalItem *dst = s->srv_items + (size_t)lane * kMaxItemsPerLaunch;
  std::memcpy(dst, items.data(), sizeof(EvalItem) * (size_t)n);
  unsigned *u = s->srv_units + (size_t)lane * kSrvMaxUnits;
  int k = 0;
  for (int i = 0; i < n; i++)
    for (int b = 0; b < items[i].nblocks; b++) u[k++] = ((unsigned)i << 16) | (unsigned)b;
  ServerDoor *d = s->srv_doors + lane;
  d->n_items = (unsigned)n;
  d->n_units = (unsigned)total;
  d->seq = seq;
  const unsigned r = ++s->srv_round[lane];
  std::atomic_thread_fence(std::memory_order_release);
  reinterpret_cast<std::atomic<unsigned> *>(&d->round)->store(r, std::memory_order_release);
  s->srv_rounds++;
  return DSLAM_OK;
}

void server_stop(dslam_session *s) {
  std::atomic_thread_fence(std::memory_order_release);
  reinterpret_cast<std::atomic<unsigned> *>(&s->srv_doors[0].stop)->store(s->srv_gen, std::memory_order_release);
}

// All machines (pose and scale alike) advance one evaluation per round.  The machines are dealt into groups that run
// concurrently: each group has its own host thread, CUDA stream, scratch and result slots and loops
// "prepare -> launch (one kernel for the whole group) -> wait -> consume" until its machines are done.  Every machine
// sees exactly the evaluations it would see alone, so results do not depend on the grouping.
void worker_main(dslam_session *s, int g) {
  cudaSetDevice(s->device);
  dslam_session::Worker *w = s->workers[g];
  std::unique_lock<std::mutex> lock(w->m);
  for (;;) {
    w->cv.wait(lock, [&] { return w->has_job || w->quit; });
    if (w->quit) return;
    lock.unlock();
    w->job();
    lock.lock();
    w->has_job = false;
    w->done = true;
    w->cv.notify_all();
  }
}

int run_lock_step(dslam_session *s, std::vector<PoseLM> &pose, std::vector<ScaleLM> &scale, dslam_ctx *counters) {
  constexpr int G = dslam_session::kLmGroups;
  const int total = (int)(pose.size() + scale.size());
  int ngroups = s->lm_groups;
  while (ngroups > 1 && total < 4 * ngroups) ngroups--;  // a group needs a few machines to be worth a launch of its own
  const int slots_per_group = (kResultSlots / ngroups) & ~1;  // result-ring slice of a group (two lanes share it half and half)
  struct Group {
    std::vector<int> members;  // >= 0: pose machine index, < 0: ~index of a scale machine
    int rc = DSLAM_OK;
    std::string err;
    long long evals = 0, launches = 0, spec_used = 0;
  } grp[G];
  {
    int k = 0;
    for (size_t i = 0; i < pose.size(); i++) grp[k++ % ngroups].members.push_back((int)i);
    for (size_t i = 0; i < scale.size(); i++) grp[k++ % ngroups].members.push_back(~(int)i);
  }
  for (int g = 0; g < ngroups; g++)
    if ((int)grp[g].members.size() > slots_per_group) return fail(DSLAM_EINVAL, "too many machines in one lock step");
  // Every group runs two launch lanes: its machines are dealt into two halves and while one half is being evaluated on the
  // GPU the host consumes the results of the other half and prepares its next round — the LM algebra of a group (the larger
  // part of a round at >= 32 machines) hides behind the kernel of the other lane.  A machine still sees exactly the
  // evaluations it would see alone.
  const int nhalves = s->lm_halves;
  const int slots_per_half = slots_per_group / 2;
  for (int g = 0; g < ngroups; g++)
    if (nhalves > 1 && ((int)grp[g].members.size() + 1) / 2 > slots_per_half) return fail(DSLAM_EINVAL, "too many machines in one lock step");
  auto two_lanes = [&](int g) { return nhalves > 1 && grp[g].members.size() >= 8; };
  int total_lanes = 0;
  for (int g = 0; g < ngroups; g++) total_lanes += two_lanes(g) ? 2 : 1;
  bool any_side = ngroups > 1;
  for (int g = 0; g < ngroups; g++) any_side |= two_lanes(g);
  if (any_side) {  // the side streams start after everything queued on the session stream (pyramids, template uploads)
    DSLAM_CUDA(cudaEventRecord(s->lm_fork, s->stream));
    for (int g = 1; g < ngroups; g++) DSLAM_CUDA(cudaStreamWaitEvent(s->lm_stream[g], s->lm_fork, 0));
    for (int g = 0; g < ngroups; g++)
      if (two_lanes(g)) DSLAM_CUDA(cudaStreamWaitEvent(s->lm_stream2[g], s->lm_fork, 0));
  }
  // Resident server: one kernel for the whole call; every lane (group x half) rings its doorbell per round.  Not used while
  // per-launch profiling is on (there are no launches to time) or when a lane could exceed one parameter block of items.
  bool use_server = s->srv_enabled && !s->prof_on;
  for (int g = 0; g < ngroups && use_server; g++) use_server = (int)grp[g].members.size() <= (two_lanes(g) ? 2 : 1) * kMaxItemsPerLaunch;
  if (use_server) {
    ServerLaneHost lanes[kSrvLanes];
    for (int g = 0; g < ngroups; g++) {
      lanes[2 * g] = ServerLaneHost{s->lm_scratch[g], g * slots_per_group};
      lanes[2 * g + 1] = ServerLaneHost{s->lm_scratch2[g], g * slots_per_group + slots_per_half};
    }
    // the side streams' earlier kernels (previous launch-path calls) must have drained before the server reuses their scratch
    for (int g = 1; g < ngroups; g++) {
      DSLAM_CUDA(cudaEventRecord(s->lm_done[g], s->lm_stream[g]));
      DSLAM_CUDA(cudaStreamWaitEvent(s->stream, s->lm_done[g], 0));
    }
    for (int g = 0; g < ngroups; g++) {
      DSLAM_CUDA(cudaEventRecord(s->lm_done2[g], s->lm_stream2[g]));
      DSLAM_CUDA(cudaStreamWaitEvent(s->stream, s->lm_done2[g], 0));
    }
    const int rc = server_start(s, 2 * ngroups, lanes);
    if (rc != DSLAM_OK) return rc;
  }
  auto now_ns = []() { return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  auto drive = [&](int g) {
    Group &gr = grp[g];
    struct Lane {
      std::vector<int> members, who;
      std::vector<unsigned char> spec;  // per item: 0 = the machine's next evaluation, d > 0 = its candidate after d further rejections
      std::vector<EvalItem> items;
      std::vector<EvalOut> outs;
      unsigned seq = 0;
      bool inflight = false;
      int slot0 = 0;
      cudaStream_t stream = nullptr;
      EvalScratch scratch{nullptr, nullptr};
    } lane[2];
    const int nl = two_lanes(g) ? 2 : 1;
    for (size_t k = 0; k < gr.members.size(); k++) lane[nl == 2 ? (k & 1) : 0].members.push_back(gr.members[k]);
    lane[0].slot0 = g * slots_per_group;
    lane[0].stream = s->lm_stream[g];
    lane[0].scratch = s->lm_scratch[g];
    lane[1].slot0 = g * slots_per_group + slots_per_half;
    lane[1].stream = s->lm_stream2[g];
    lane[1].scratch = s->lm_scratch2[g];
    long long t_prep = 0, t_launch = 0, t_wait = 0, rounds = 0;
    // prepare the next round of a lane and launch it; leaves inflight = false when all its machines are done
    auto issue = [&](Lane &L) {
      const long long t0 = now_ns();
      L.items.clear();
      L.who.clear();
      L.spec.clear();
      size_t cap = (size_t)(nl == 2 ? slots_per_half : slots_per_group);  // result slots of this lane
      if (use_server && cap > (size_t)kMaxItemsPerLaunch) cap = kMaxItemsPerLaunch;
      const int depth = s->lm_spec_depth;
      size_t todo = L.members.size();  // members that still need their (mandatory) next evaluation in this round
      for (int m : L.members) {
        todo--;
        if (m >= 0) {
          PoseLM &M = pose[m];
          if (M.phase == PoseLM::DONE) continue;
          L.items.emplace_back();
          M.request(L.items.back());
          L.who.push_back(m);
          L.spec.push_back(0);
          const int dm = (M.phase == PoseLM::ITER && M.last_rejected) ? depth : 1;
          if (dm > 1) {  // inside a run of rejections: evaluate its next candidates too
            const PoseLM::Snap sn = M.snap();
            for (int d = 1; d < dm && L.items.size() + todo < cap && M.simulate_reject(); d++) {
              L.items.emplace_back();
              M.request(L.items.back());
              L.who.push_back(m);
              L.spec.push_back((unsigned char)d);
            }
            M.restore(sn);
          }
        } else {
          ScaleLM &M = scale[~m];
          if (M.phase == ScaleLM::DONE) continue;
          L.items.emplace_back();
          M.request(L.items.back());
          L.who.push_back(m);
          L.spec.push_back(0);
          const int dm = (M.phase == ScaleLM::ITER && M.last_rejected) ? depth : 1;
          if (dm > 1) {
            const ScaleLM::Snap sn = M.snap();
            for (int d = 1; d < dm && L.items.size() + todo < cap && M.simulate_reject(); d++) {
              L.items.emplace_back();
              M.request(L.items.back());
              L.who.push_back(m);
              L.spec.push_back((unsigned char)d);
            }
            M.restore(sn);
          }
        }
      }
      const long long t1 = now_ns();
      t_prep += t1 - t0;
      L.inflight = false;
      if (L.items.empty()) return;
      if (use_server) {
        L.seq = ++s->seq;
        gr.rc = server_post(s, 2 * g + (int)(&L - lane), L.items, L.seq, total_lanes);
      } else {
        gr.rc = launch_items(s, L.items, L.slot0, L.stream, L.scratch, &L.seq, &gr.launches, total_lanes);
      }
      t_launch += now_ns() - t1;
      rounds++;
      L.inflight = gr.rc == DSLAM_OK;
    };
    // wait for the results of a lane's launch and let its machines consume them
    auto retire = [&](Lane &L) {
      const long long t2 = now_ns();
      gr.rc = collect_items(s, L.items, L.slot0, L.stream, L.seq, L.outs);
      L.inflight = false;
      if (gr.rc != DSLAM_OK) return;
      const long long t3 = now_ns();
      t_wait += t3 - t2;
      gr.evals += (long long)L.items.size();
      // A speculative result is consumed only if, after the results before it, the machine asks for exactly that evaluation
      // (same item up to the launch-layout fields): the sequence of evaluations a machine SEES is the sequential one.
      auto same_request = [](const EvalItem &a, const EvalItem &b) { return std::memcmp(&a, &b, offsetof(EvalItem, nblocks)) == 0; };
      bool chain_ok = true;
      for (size_t k = 0; k < L.who.size(); k++) {
        const int m = L.who[k];
        if (L.spec[k] == 0) {
          chain_ok = true;
        } else {
          if (!chain_ok) continue;
          EvalItem want;
          std::memset(&want, 0, sizeof(want));
          bool live;
          if (m >= 0) {
            live = pose[m].phase == PoseLM::ITER && pose[m].last_rejected;
            if (live) pose[m].request(want);
          } else {
            live = scale[~m].phase == ScaleLM::ITER && scale[~m].last_rejected;
            if (live) scale[~m].request(want);
          }
          if (!live || !same_request(want, L.items[k])) {
            chain_ok = false;
            continue;
          }
          gr.spec_used++;
        }
        if (m >= 0) pose[m].consume(L.outs[k]);
        else scale[~m].consume(L.outs[k]);
      }
      t_prep += now_ns() - t3;
    };
    for (int l = 0; l < nl && gr.rc == DSLAM_OK; l++) issue(lane[l]);
    while (gr.rc == DSLAM_OK && (lane[0].inflight || lane[1].inflight)) {
      for (int l = 0; l < nl && gr.rc == DSLAM_OK; l++) {
        if (!lane[l].inflight) continue;
        retire(lane[l]);
        if (gr.rc == DSLAM_OK) issue(lane[l]);
      }
    }
    if (gr.rc != DSLAM_OK) gr.err = g_err;
    s->t_prep_ns += t_prep;
    s->t_launch_ns += t_launch;
    s->t_wait_ns += t_wait;
    s->n_rounds += rounds;
  };
  for (int g = 1; g < ngroups; g++) {
    if (!s->workers[g]) {
      s->workers[g] = new dslam_session::Worker();
      s->workers[g]->th = std::thread(worker_main, s, g);
    }
    dslam_session::Worker *w = s->workers[g];
    std::lock_guard<std::mutex> lock(w->m);
    w->job = [&drive, g]() { drive(g); };
    w->done = false;
    w->has_job = true;
    w->cv.notify_all();
  }
  drive(0);
  for (int g = 1; g < ngroups; g++) {
    dslam_session::Worker *w = s->workers[g];
    std::unique_lock<std::mutex> lock(w->m);
    w->cv.wait(lock, [&] { return w->done; });
  }
  if (use_server) server_stop(s);
  int rc = DSLAM_OK;
  for (int g = 0; g < ngroups; g++) {
    if (counters) {
      counters->n_evals += grp[g].evals;
      counters->n_launches += grp[g].launches;
    }
    if (grp[g].rc != DSLAM_OK && rc == DSLAM_OK) {
      rc = grp[g].rc;
      set_error("%s", grp[g].err.c_str());
    }
  }
  // later work on the session stream (next pyramids, template changes) is ordered after the side streams
  for (int k = 1; k < ngroups; k++) {
    DSLAM_CUDA(cudaEventRecord(s->lm_done[k], s->lm_stream[k]));
    DSLAM_CUDA(cudaStreamWaitEvent(s->stream, s->lm_done[k], 0));
  }
  for (int k = 0; k < ngroups; k++)
    if (two_lanes(k)) {
      DSLAM_CUDA(cudaEventRecord(s->lm_done2[k], s->lm_stream2[k]));
      DSLAM_CUDA(cudaStreamWaitEvent(s->stream, s->lm_done2[k], 0));
    }
  return rc;
}

// trackNewestCoarse epilogue :612-637 for one finished machine
bool finish_pose(const dslam_ctx *c, PoseLM &m, float new_exposure, double *pose7_io, double *aff) {
  bool good = m.ok;
  if (!good) return false;
  m.cur.to7(pose7_io);  // outputs are only written when the level loop ran to completion
  aff[0] = m.aff_cur[0];
  aff[1] = m.aff_cur[1];
  if ((c->affModeA != 0 && (fabsf((float)aff[0]) > 1.2)) || (c->affModeB != 0 && (fabsf((float)aff[1]) > 200))) return false;
  double rel[2];
  hm::aff_from_to(c->ref_exposure, new_exposure, c->ref_a, c->ref_b, aff[0], aff[1], rel);
  const float relA = (float)rel[0], relB = (float)rel[1];
  if ((c->affModeA == 0 && (fabsf(logf(relA)) > 1.5)) || (c->affModeB == 0 && (fabsf(relB) > 200))) return false;
  if (c->affModeA < 0) aff[0] = 0;
  if (c->affModeB < 0) aff[1] = 0;
  return true;
}

// Order the session stream behind the asynchronous build that last wrote `f` (no-op for synchronous builds and for
// generations the stream already waited for: the builds of a session complete in order).
int frame_acquire(const dslam_frame *f) {
  dslam_session *s = f->s;
  if (f->async_gen > s->pyr_waited) {
    DSLAM_CUDA(cudaStreamWaitEvent(s->stream, s->pyr_ev[f->async_gen % dslam_session::kPyrEvents], 0));
    s->pyr_waited = f->async_gen;
  }
  return DSLAM_OK;
}

int check_ctx_frame(const dslam_ctx *c, const dslam_frame *f, int coarsestLvl) {
  if (!c || !f) return fail(DSLAM_EINVAL, "null context or frame");
  if (c->s != f->s) return fail(DSLAM_EINVAL, "context and frame belong to different sessions");
  if (!c->have_ref) return fail(DSLAM_ESTATE, "no tracking reference uploaded (dslam_ref_upload / dslam_ref_build)");
  if (!f->built) return fail(DSLAM_ESTATE, "frame pyramid has not been built");
  if (f->w != c->w[0] || f->h != c->h[0] || f->levels < c->levels) return fail(DSLAM_EINVAL, "frame geometry does not match the context");
  if (coarsestLvl < 0 || coarsestLvl >= c->levels || coarsestLvl >= 5) return fail(DSLAM_EINVAL, "coarsestLvl %d out of range", coarsestLvl);
  return frame_acquire(f);
}

}  // namespace
}  // namespace dslam

#ifdef DSLAM_KERNEL_TIMING
namespace dslam { cudaError_t debug_times(unsigned long long *out8, int reset); }
#endif
using namespace dslam;

// ===================================================================================================
// C ABI
// ===================================================================================================
extern "C" {

int dslam_version(void) { return 100; }
const char *dslam_last_error(void) { return dslam::g_err; }

int dslam_device_count(int *count) {
  if (!count) return fail(DSLAM_EINVAL, "null argument");
  int n = 0;
  const cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    *count = 0;
    return cuda_fail(e, "cudaGetDeviceCount");
  }
  *count = n;
  return DSLAM_OK;
}

// ---- sessions ------------------------------------------------------------------------------------
int dslam_session_create(int device, dslam_session **out) {
  if (!out) return fail(DSLAM_EINVAL, "null argument");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount");
  if (n < 1) return fail(DSLAM_ENODEVICE, "no CUDA device");
  if (device < 0 || device >= n) return fail(DSLAM_EINVAL, "device %d out of range (%d devices)", device, n);
  DSLAM_CUDA(cudaSetDevice(device));
  dslam_session *s = new (std::nothrow) dslam_session();
  if (!s) return fail(DSLAM_ENOMEM, "out of host memory");
  s->device = device;
  cudaDeviceGetAttribute(&s->num_sms, cudaDevAttrMultiProcessorCount, device);
  if (s->num_sms <= 0) s->num_sms = 148;
  int prio_lo = 0;
  DSLAM_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &s->prio_hi));
  DSLAM_CUDA(cudaStreamCreateWithPriority(&s->stream, cudaStreamNonBlocking, s->prio_hi));
  DSLAM_CUDA(cudaHostAlloc((void **)&s->results_host, sizeof(EvalResult) * kResultSlots, cudaHostAllocMapped));
  std::memset(s->results_host, 0, sizeof(EvalResult) * kResultSlots);  // sequence numbers start at 1
  DSLAM_CUDA(cudaHostGetDevicePointer((void **)&s->results_dev, s->results_host, 0));
  DSLAM_CUDA(cudaMalloc((void **)&s->scratch.partials, sizeof(double) * kMaxItemsPerLaunch * kMaxBlocksPerItem * kPoseVals));
  DSLAM_CUDA(cudaMalloc((void **)&s->scratch.counters, sizeof(int) * kMaxItemsPerLaunch * 4));
  DSLAM_CUDA(cudaMemsetAsync(s->scratch.counters, 0, sizeof(int) * kMaxItemsPerLaunch * 4, s->stream));
  DSLAM_CUDA(cudaEventCreate(&s->mark[0]));
  DSLAM_CUDA(cudaEventCreate(&s->mark[1]));
  s->lm_stream[0] = s->stream;
  s->lm_scratch[0] = s->scratch;
  {  // one host thread per group: as many as the cores this process can count on (torchrun exports LOCAL_WORLD_SIZE)
    int ranks = 1;
    if (const char *e = getenv("LOCAL_WORLD_SIZE")) ranks = atoi(e) > 0 ? atoi(e) : 1;
    int g = (int)(std::thread::hardware_concurrency() / (unsigned)ranks);
    g = g > dslam_session::kLmGroupsDefault ? dslam_session::kLmGroupsDefault : (g < 1 ? 1 : g);
    s->lm_groups = g;
  }
  if (const char *e = getenv("DSLAM_LM_GROUPS")) {
    const int v = atoi(e);
    if (v >= 1 && v <= dslam_session::kLmGroups) s->lm_groups = v;
  }
  if (const char *e = getenv("DSLAM_LM_SPEC")) {
    const int v = atoi(e);
    if (v >= 1 && v <= 8) s->lm_spec_depth = v;
  }
  if (const char *e = getenv("DSLAM_LM_SERVER")) s->srv_enabled = atoi(e) != 0;
  if (const char *e = getenv("DSLAM_LM_SERVER_CTAS")) {
    const int v = atoi(e);
    if (v >= 1 && v <= 5) s->srv_ctas_per_sm = v;
  }
  DSLAM_CUDA(cudaEventCreateWithFlags(&s->lm_fork, cudaEventDisableTiming));
  {  // LM rounds are latency-critical, the overlapped pyramid builds are not: highest / lowest stream priority
    DSLAM_CUDA(cudaStreamCreateWithPriority(&s->pyr_stream, cudaStreamNonBlocking, prio_lo));
  }
  if (const char *e = getenv("DSLAM_PYR_ASYNC_CTAS")) {
    const int v = atoi(e);
    if (v >= 1 && v <= 6) s->pyr_async_ctas = v;
  }
  DSLAM_CUDA(cudaEventCreateWithFlags(&s->pyr_in, cudaEventDisableTiming));
  for (int k = 0; k < dslam_session::kPyrEvents; k++) DSLAM_CUDA(cudaEventCreateWithFlags(&s->pyr_ev[k], cudaEventDisableTiming));
  // two launch lanes per group pay off when host threads are scarce (measured on B200 / 16 cores, 512 streams: +20 % at 4
  // groups, -15 % at 8: with enough threads the extra launches cost more than the hidden host algebra gains)
  s->lm_halves = s->lm_groups <= 4 ? 2 : 1;
  if (const char *e = getenv("DSLAM_LM_HALVES")) s->lm_halves = atoi(e) == 1 ? 1 : 2;
  for (int g = 0; g < dslam_session::kLmGroups; g++) {
    DSLAM_CUDA(cudaStreamCreateWithPriority(&s->lm_stream2[g], cudaStreamNonBlocking, s->prio_hi));
    DSLAM_CUDA(cudaMalloc((void **)&s->lm_scratch2[g].partials, sizeof(double) * kMaxItemsPerLaunch * kMaxBlocksPerItem * kPoseVals));
    DSLAM_CUDA(cudaMalloc((void **)&s->lm_scratch2[g].counters, sizeof(int) * kMaxItemsPerLaunch * 4));
    DSLAM_CUDA(cudaMemsetAsync(s->lm_scratch2[g].counters, 0, sizeof(int) * kMaxItemsPerLaunch * 4, s->lm_stream2[g]));
    DSLAM_CUDA(cudaEventCreateWithFlags(&s->lm_done2[g], cudaEventDisableTiming));
    DSLAM_CUDA(cudaStreamSynchronize(s->lm_stream2[g]));
  }
  for (int g = 1; g < dslam_session::kLmGroups; g++) {
    DSLAM_CUDA(cudaStreamCreateWithPriority(&s->lm_stream[g], cudaStreamNonBlocking, s->prio_hi));
    DSLAM_CUDA(cudaMalloc((void **)&s->lm_scratch[g].partials, sizeof(double) * kMaxItemsPerLaunch * kMaxBlocksPerItem * kPoseVals));
    DSLAM_CUDA(cudaMalloc((void **)&s->lm_scratch[g].counters, sizeof(int) * kMaxItemsPerLaunch * 4));
    DSLAM_CUDA(cudaMemsetAsync(s->lm_scratch[g].counters, 0, sizeof(int) * kMaxItemsPerLaunch * 4, s->lm_stream[g]));
    DSLAM_CUDA(cudaEventCreateWithFlags(&s->lm_done[g], cudaEventDisableTiming));
    DSLAM_CUDA(cudaStreamSynchronize(s->lm_stream[g]));
  }
  DSLAM_CUDA(cudaStreamSynchronize(s->stream));
  *out = s;
  return DSLAM_OK;
}

int dslam_session_destroy(dslam_session *s) {
  if (!s) return DSLAM_OK;
  cudaSetDevice(s->device);
  cudaStreamSynchronize(s->stream);
  for (int g = 1; g < dslam_session::kLmGroups; g++) {
    if (s->workers[g]) {
      {
        std::lock_guard<std::mutex> lock(s->workers[g]->m);
        s->workers[g]->quit = true;
        s->workers[g]->cv.notify_all();
      }
      s->workers[g]->th.join();
      delete s->workers[g];
      s->workers[g] = nullptr;
    }
    if (!s->lm_stream[g]) continue;
    cudaStreamSynchronize(s->lm_stream[g]);
    cudaFree(s->lm_scratch[g].partials);
    cudaFree(s->lm_scratch[g].counters);
    cudaEventDestroy(s->lm_done[g]);
    cudaStreamDestroy(s->lm_stream[g]);
  }
  for (int g = 0; g < dslam_session::kLmGroups; g++) {
    if (!s->lm_stream2[g]) continue;
    cudaStreamSynchronize(s->lm_stream2[g]);
    cudaFree(s->lm_scratch2[g].partials);
    cudaFree(s->lm_scratch2[g].counters);
    cudaEventDestroy(s->lm_done2[g]);
    cudaStreamDestroy(s->lm_stream2[g]);
  }
  if (s->lm_fork) cudaEventDestroy(s->lm_fork);
  if (s->pyr_stream) {
    cudaStreamSynchronize(s->pyr_stream);
    cudaStreamDestroy(s->pyr_stream);
  }
  if (s->pyr_in) cudaEventDestroy(s->pyr_in);
  for (int k = 0; k < dslam_session::kPyrEvents; k++)
    if (s->pyr_ev[k]) cudaEventDestroy(s->pyr_ev[k]);
  if (s->srv_doors) cudaFreeHost(s->srv_doors);
  if (s->srv_items) cudaFreeHost(s->srv_items);
  if (s->srv_units) cudaFreeHost(s->srv_units);
  cudaFree(s->srv_queue); cudaFree(s->srv_qctl); cudaFree(s->srv_lane_seq); cudaFree(s->srv_items_dev);
  cudaFree(s->upload_arena);
  cudaFree(s->scratch.partials);
  cudaFree(s->scratch.counters);
  cudaFreeHost(s->results_host);
  cudaEventDestroy(s->mark[0]);
  cudaEventDestroy(s->mark[1]);
  for (cudaEvent_t e : s->prof_ev) cudaEventDestroy(e);
  cudaStreamDestroy(s->stream);
  delete s;
  return DSLAM_OK;
}

int dslam_session_sync(dslam_session *s) {
  if (!s) return fail(DSLAM_EINVAL, "null session");
  for (int g = 1; g < dslam_session::kLmGroups; g++) DSLAM_CUDA(cudaStreamSynchronize(s->lm_stream[g]));
  for (int g = 0; g < dslam_session::kLmGroups; g++)
    if (s->lm_stream2[g]) DSLAM_CUDA(cudaStreamSynchronize(s->lm_stream2[g]));
  DSLAM_CUDA(cudaStreamSynchronize(s->stream));
  if (s->pyr_stream) DSLAM_CUDA(cudaStreamSynchronize(s->pyr_stream));
  return DSLAM_OK;
}

int dslam_session_stream(dslam_session *s, void **cuda_stream_out) {
  if (!s || !cuda_stream_out) return fail(DSLAM_EINVAL, "null argument");
  *cuda_stream_out = (void *)s->stream;
  return DSLAM_OK;
}

int dslam_session_launch_count(dslam_session *s, long long *count) {
  if (!s || !count) return fail(DSLAM_EINVAL, "null argument");
  *count = s->launches.load();
  return DSLAM_OK;
}

int dslam_session_mark(dslam_session *s, int which) {
  if (!s || which < 0 || which > 1) return fail(DSLAM_EINVAL, "bad argument");
  DSLAM_CUDA(cudaEventRecord(s->mark[which], s->stream));
  return DSLAM_OK;
}

int dslam_session_elapsed_ms(dslam_session *s, float *ms) {
  if (!s || !ms) return fail(DSLAM_EINVAL, "null argument");
  DSLAM_CUDA(cudaEventSynchronize(s->mark[1]));
  DSLAM_CUDA(cudaEventElapsedTime(ms, s->mark[0], s->mark[1]));
  return DSLAM_OK;
}

int dslam_session_profile(dslam_session *s, int enable) {
  if (!s) return fail(DSLAM_EINVAL, "null session");
  DSLAM_CUDA(cudaStreamSynchronize(s->stream));
  s->prof_on = enable != 0;
  s->prof_used = 0;
  return DSLAM_OK;
}

int dslam_session_profile_read(dslam_session *s, double out[12]) {
  if (!s || !out) return fail(DSLAM_EINVAL, "null argument");
  DSLAM_CUDA(cudaStreamSynchronize(s->stream));
  for (int i = 0; i < 12; i++) out[i] = 0;
  for (size_t k = 0; k < s->prof_used; k++) {
    float ms = 0;
    DSLAM_CUDA(cudaEventElapsedTime(&ms, s->prof_ev[2 * k], s->prof_ev[2 * k + 1]));
    const int m = 4 * s->prof_mode[k];
    out[m + 0] += 1;
    out[m + 1] += ms;
    out[m + 2] += (double)s->prof_points[k];
    if (ms > out[m + 3]) out[m + 3] = ms;
  }
  s->prof_used = 0;
  return DSLAM_OK;
}

int dslam_session_host_times(dslam_session *s, double out[4]) {
  if (!s || !out) return fail(DSLAM_EINVAL, "null argument");
  out[0] = s->t_prep_ns.exchange(0) * 1e-6;
  out[1] = s->t_launch_ns.exchange(0) * 1e-6;
  out[2] = s->t_wait_ns.exchange(0) * 1e-6;
  out[3] = (double)s->n_rounds.exchange(0);
  return DSLAM_OK;
}

#ifdef DSLAM_KERNEL_TIMING
int dslam_debug_kernel_times(dslam_session *s, unsigned long long *out8, int reset) {
  DSLAM_CUDA(cudaStreamSynchronize(s->stream));
  DSLAM_CUDA(dslam::debug_times(out8, reset));
  return DSLAM_OK;
}
#endif

int dslam_plan_eval_launch(int n_items, const int *n_points, int num_sms, int lanes, int *nblocks, int *cta_begin, int *total_ctas) {
  if (n_items < 1 || n_items > kMaxItemsPerLaunch || !n_points || !nblocks || !cta_begin || !total_ctas || num_sms < 1 || lanes < 1)
    return fail(DSLAM_EINVAL, "bad argument");
  std::vector<EvalItem> items((size_t)n_items);
  for (int i = 0; i < n_items; i++) {
    if (n_points[i] < 0) return fail(DSLAM_EINVAL, "negative point count");
    items[i].n = n_points[i];
  }
  *total_ctas = assign_blocks(items.data(), n_items, num_sms, lanes);
  for (int i = 0; i < n_items; i++) {
    nblocks[i] = items[i].nblocks;
    cta_begin[i] = items[i].cta_begin;
  }
  return DSLAM_OK;
}

int dslam_host_alloc(unsigned long long bytes, void **out) {
  if (!out) return fail(DSLAM_EINVAL, "null argument");
  DSLAM_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
  return DSLAM_OK;
}
int dslam_host_free(void *p) {
  if (p) DSLAM_CUDA(cudaFreeHost(p));
  return DSLAM_OK;
}

// ---- frames --------------------------------------------------------------------------------------
int dslam_frame_create(dslam_session *s, int w, int h, int levels, dslam_frame **out) {
  if (!s || !out) return fail(DSLAM_EINVAL, "null argument");
  *out = nullptr;
  if (w < 8 || h < 8 || levels < 1 || levels > DSLAM_MAX_LEVELS || (w >> (levels - 1)) < 1 || (h >> (levels - 1)) < 1)
    return fail(DSLAM_EINVAL, "bad pyramid geometry %dx%d, %d levels", w, h, levels);
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return fail(DSLAM_ENODEVICE, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  DSLAM_CUDA(cudaSetDevice(s->device));
  dslam_frame *f = new (std::nothrow) dslam_frame();
  if (!f) return fail(DSLAM_ENOMEM, "out of host memory");
  f->s = s; f->w = w; f->h = h; f->levels = levels;
  PyramidLevels &L = f->L;
  L.levels = levels;
  size_t bytes = 0, plane_off[kMaxLevels], tex_off[kMaxLevels];
  f->px_off[0] = 0;
  int tiles = 0;
  for (int l = 0; l < levels; l++) {
    L.w[l] = w >> l;
    L.h[l] = h >> l;
    L.pitch[l] = (int)align_up((size_t)L.w[l], 4);
    plane_off[l] = bytes;
    bytes = align_up(bytes + (size_t)L.pitch[l] * L.h[l] * sizeof(float), 256);
    tex_off[l] = bytes;
    bytes = align_up(bytes + (size_t)L.w[l] * L.h[l] * sizeof(float4), 256);
    f->px_off[l + 1] = f->px_off[l] + (size_t)L.w[l] * L.h[l];
    L.tiles_x[l] = (L.w[l] + kGradTileW - 1) / kGradTileW;
    L.tile_begin[l] = tiles;
    tiles += L.tiles_x[l] * ((L.h[l] + kGradTileH - 1) / kGradTileH);
    L.host_dIp[l] = nullptr;
    L.host_abs[l] = nullptr;
  }
  for (int l = levels; l <= kMaxLevels; l++) L.tile_begin[l] = tiles;
  cudaError_t e = cudaMalloc(&f->block, bytes);
  if (e != cudaSuccess) {
    delete f;
    return cuda_fail(e, "cudaMalloc(frame)");
  }
  for (int l = 0; l < levels; l++) {
    L.plane[l] = (float *)((char *)f->block + plane_off[l]);
    L.tex[l] = (float4 *)((char *)f->block + tex_off[l]);
    const cuuint64_t gdim[2] = {(cuuint64_t)L.w[l], (cuuint64_t)L.h[l]};
    const cuuint64_t gstride[1] = {(cuuint64_t)L.pitch[l] * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)kGradBoxW, (cuuint32_t)kGradBoxH};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(&f->devh.map[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, L.plane[l], gdim, gstride, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      cudaFree(f->block);
      delete f;
      return fail(DSLAM_ECUDA, "cuTensorMapEncodeTiled failed (%d) for level %d", (int)r, l);
    }
  }
  for (int l = levels; l < kMaxLevels; l++) f->devh.map[l] = f->devh.map[0];
  f->geom.levels = levels;
  for (int l = 0; l < kMaxLevels; l++) {
    f->geom.w[l] = L.w[l]; f->geom.h[l] = L.h[l]; f->geom.pitch[l] = L.pitch[l]; f->geom.tiles_x[l] = L.tiles_x[l];
    f->devh.plane[l] = l < levels ? L.plane[l] : nullptr;
    f->devh.tex[l] = l < levels ? L.tex[l] : nullptr;
    f->devh.host_dIp[l] = nullptr;
    f->devh.host_abs[l] = nullptr;
  }
  for (int l = 0; l <= kMaxLevels; l++) f->geom.tile_begin[l] = L.tile_begin[l];
  f->devh.B256 = nullptr;
  e = cudaMalloc((void **)&f->dev, sizeof(FrameDev));
  if (e != cudaSuccess) {
    cudaFree(f->block);
    delete f;
    return cuda_fail(e, "cudaMalloc(frame descriptor)");
  }
  f->dev_dirty = true;
  e = cudaEventCreateWithFlags(&f->host_ready, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&f->built_ev, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&f->copy_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    cudaFree(f->block);
    delete f;
    return cuda_fail(e, "cudaEventCreate");
  }
  *out = f;
  return DSLAM_OK;
}

int dslam_frame_destroy(dslam_frame *f) {
  if (!f) return DSLAM_OK;
  cudaSetDevice(f->s->device);
  cudaStreamSynchronize(f->s->stream);
  if (f->async_gen && f->s->pyr_stream) cudaStreamSynchronize(f->s->pyr_stream);
  cudaFree(f->block);
  cudaFree(f->dev);
  cudaFree(f->stage_dIp);
  cudaFree(f->stage_abs);
  cudaFree(f->B_dev);
  if (f->copy_stream) {
    cudaStreamSynchronize(f->copy_stream);
    cudaStreamDestroy(f->copy_stream);
  }
  cudaEventDestroy(f->host_ready);
  cudaEventDestroy(f->built_ev);
  delete f;
  return DSLAM_OK;
}

int dslam_frame_upload(dslam_frame *f, const float *color) {
  if (!f || !color) return fail(DSLAM_EINVAL, "null argument");
  const PyramidLevels &L = f->L;
  const int ra = frame_acquire(f);  // an asynchronous build may still be reading the old image
  if (ra != DSLAM_OK) return ra;
  if (L.pitch[0] == f->w)  // dense plane: a plain 1-D copy
    DSLAM_CUDA(cudaMemcpyAsync(L.plane[0], color, (size_t)f->w * f->h * sizeof(float), cudaMemcpyHostToDevice, f->s->stream));
  else
    DSLAM_CUDA(cudaMemcpy2DAsync(L.plane[0], (size_t)L.pitch[0] * sizeof(float), color, (size_t)f->w * sizeof(float), (size_t)f->w * sizeof(float),
                                 f->h, cudaMemcpyHostToDevice, f->s->stream));
  f->uploaded = true;
  f->built = false;
  return DSLAM_OK;
}

static bool same_geometry(const dslam_frame *a, const dslam_frame *b);

// The images of a step in as few DMA transfers as possible: runs of images that are contiguous in host memory (a capture
// ring / pinned arena, image i+1 directly behind image i) travel as ONE host-to-device copy into a session arena and are
// dealt to the frames' level-0 planes by one small kernel per 64 frames.  PCIe moves large transfers markedly better than
// many 1.8 MB ones when the opposite direction is busy with the host mirrors (tools/pcie_bw.py).
int dslam_frame_upload_batch(int n, dslam_frame *const *frames, const float *const *colors) {
  if (n < 1 || !frames || !colors) return fail(DSLAM_EINVAL, "bad argument");
  for (int i = 0; i < n; i++) {
    if (!frames[i] || !colors[i]) return fail(DSLAM_EINVAL, "null frame or image");
    if (frames[i]->s != frames[0]->s) return fail(DSLAM_EINVAL, "all frames of a batch must live in one session");
    const int ra = frame_acquire(frames[i]);
    if (ra != DSLAM_OK) return ra;
  }
  dslam_session *s = frames[0]->s;
  DSLAM_CUDA(cudaSetDevice(s->device));
  int i = 0;
  while (i < n) {
    dslam_frame *f0 = frames[i];
    const size_t px = (size_t)f0->w * f0->h;
    int j = i + 1;
    const bool dense = f0->L.pitch[0] == f0->w && px % 4 == 0 && ((uintptr_t)colors[i] & 15) == 0;
    while (dense && j < n && j - i < kMaxFramesPerLaunch * 8 && same_geometry(frames[j], f0) && colors[j] == colors[j - 1] + px) j++;
    const int run = j - i;
    if (run < 2) {
      const int rc = dslam_frame_upload(f0, colors[i]);
      if (rc != DSLAM_OK) return rc;
      i = j;
      continue;
    }
    if (s->upload_arena_floats < px * run) {
      // the arena may still feed a scatter kernel queued earlier: the free is stream-ordered
      if (s->upload_arena) DSLAM_CUDA(cudaFreeAsync(s->upload_arena, s->stream));
      s->upload_arena = nullptr;
      s->upload_arena_floats = 0;
      DSLAM_CUDA(cudaMallocAsync((void **)&s->upload_arena, px * run * sizeof(float), s->stream));
      s->upload_arena_floats = px * run;
    }
    DSLAM_CUDA(cudaMemcpyAsync(s->upload_arena, colors[i], px * run * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    for (int b = 0; b < run; b += kMaxFramesPerLaunch) {
      const int cnt = run - b < kMaxFramesPerLaunch ? run - b : kMaxFramesPerLaunch;
      PlaneBatch B;
      for (int k = 0; k < cnt; k++) B.plane[k] = frames[i + b + k]->L.plane[0];
      DSLAM_CUDA(launch_scatter_planes(B, cnt, (int)(px / 4), s->upload_arena + (size_t)b * px, s->stream));
      s->launches++;
    }
    for (int k = i; k < j; k++) {
      frames[k]->uploaded = true;
      frames[k]->built = false;
    }
    i = j;
  }
  return DSLAM_OK;
}

static int frame_ensure_staging(dslam_frame *f, bool want_dIp, bool want_abs) {
  const size_t tot = f->px_off[f->levels];
  if (want_dIp && !f->stage_dIp) DSLAM_CUDA(cudaMalloc((void **)&f->stage_dIp, tot * 3 * sizeof(float)));
  if (want_abs && !f->stage_abs) DSLAM_CUDA(cudaMalloc((void **)&f->stage_abs, tot * sizeof(float)));
  return DSLAM_OK;
}

static bool same_geometry(const dslam_frame *a, const dslam_frame *b) {
  return a->w == b->w && a->h == b->h && a->levels == b->levels; }

// point the device descriptor of a frame at its staging copies / gamma table and upload it if anything changed
static int frame_prepare(dslam_frame *f, const float *B256, bool stage_dIp, bool stage_abs) {
  dslam_session *s = f->s;
  const int ra = frame_acquire(f);
  if (ra != DSLAM_OK) return ra;
  const float *Bd = nullptr;
  if (B256) {
    if (!f->B_dev) DSLAM_CUDA(cudaMalloc((void **)&f->B_dev, 256 * sizeof(float)));
    DSLAM_CUDA(cudaMemcpyAsync(f->B_dev, B256, 256 * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    Bd = f->B_dev;
  }
  const int rc = frame_ensure_staging(f, stage_dIp, stage_abs);
  if (rc != DSLAM_OK) return rc;
  if (f->host_pending) {  // the previous frame's mirrors may still be draining from the staging copies
    DSLAM_CUDA(cudaStreamWaitEvent(s->stream, f->host_ready, 0));
    f->host_pending = false;
  }
  FrameDev &d = f->devh;
  for (int l = 0; l < f->levels; l++) {
    float *hd = stage_dIp ? f->stage_dIp + 3 * f->px_off[l] : nullptr;
    float *ha = stage_abs ? f->stage_abs + f->px_off[l] : nullptr;
    if (d.host_dIp[l] != hd || d.host_abs[l] != ha) f->dev_dirty = true;
    d.host_dIp[l] = hd;
    d.host_abs[l] = ha;
  }
  if (d.B256 != Bd) f->dev_dirty = true;
  d.B256 = Bd;
  if (f->dev_dirty) {
    DSLAM_CUDA(cudaMemcpyAsync(f->dev, &f->devh, sizeof(FrameDev), cudaMemcpyHostToDevice, s->stream));
    f->dev_dirty = false;
  }
  return DSLAM_OK;
}

// two launches for every run of <= kMaxFramesPerLaunch frames with the same geometry
static int frames_build(int n, dslam_frame *const *frames, const float *B256, bool stage_dIp, bool stage_abs, bool async = false) {
  for (int i = 0; i < n; i++) {
    if (!frames[i]) return fail(DSLAM_EINVAL, "null frame");
    if (!frames[i]->uploaded) return fail(DSLAM_ESTATE, "dslam_frame_build before dslam_frame_upload");
    if (frames[i]->s != frames[0]->s) return fail(DSLAM_EINVAL, "all frames of a batch must live in one session");
    const int rc = frame_prepare(frames[i], B256, stage_dIp, stage_abs);
    if (rc != DSLAM_OK) return rc;
  }
  dslam_session *s = frames[0]->s;
  cudaStream_t st = s->stream;
  if (async) {  // behind everything queued on the session stream so far (uploads, the LM rounds that read these frame objects)
    st = s->pyr_stream;
    DSLAM_CUDA(cudaEventRecord(s->pyr_in, s->stream));
    DSLAM_CUDA(cudaStreamWaitEvent(st, s->pyr_in, 0));
  }
  int i = 0;
  while (i < n) {
    FrameBatch B;
    B.G = frames[i]->geom;
    int cnt = 0;
    while (i + cnt < n && cnt < kMaxFramesPerLaunch && same_geometry(frames[i], frames[i + cnt])) {
      B.f[cnt] = frames[i + cnt]->dev;
      cnt++;
    }
    if (B.G.levels > 1) {
      DSLAM_CUDA(launch_downsample(B, cnt, st));
      s->launches++;
    }
    DSLAM_CUDA(launch_gradients(B, cnt, st, async ? s->pyr_async_ctas : 6));
    s->launches++;
    for (int k = 0; k < cnt; k++) {
      frames[i + k]->built = true;
      frames[i + k]->staged = stage_dIp || stage_abs;
      frames[i + k]->staged_dIp = stage_dIp;  // which host-layout copies THIS build filled (a buffer left over from an earlier
      frames[i + k]->staged_abs = stage_abs;  // build holds another image's data)
      frames[i + k]->async_gen = async ? s->pyr_gen + 1 : 0;
    }
    i += cnt;
  }
  if (async) {
    s->pyr_gen++;
    DSLAM_CUDA(cudaEventRecord(s->pyr_ev[s->pyr_gen % dslam_session::kPyrEvents], st));
  }
  return DSLAM_OK;
}

int dslam_frame_build(dslam_frame *f, const float *B256) {
  if (!f) return fail(DSLAM_EINVAL, "null frame");
  return frames_build(1, &f, B256, false, false);
}

int dslam_frame_build_batch(int n, dslam_frame *const *frames, const float *B256, int stage_host) {
  if (n < 1 || !frames) return fail(DSLAM_EINVAL, "bad argument");
  return frames_build(n, frames, B256, (stage_host & 1) != 0, (stage_host & 2) != 0, (stage_host & 4) != 0);
}

static int frame_copy_out(dslam_frame *f, float *const *host_dIp, float *const *host_absgrad) {
  dslam_session *s = f->s;
  // the copies run on the frame's own stream, behind the kernels that filled the staging buffers
  if (f->async_gen > s->pyr_waited) {  // built on the pyramid stream: wait for that build, not for the session stream
    DSLAM_CUDA(cudaStreamWaitEvent(f->copy_stream, s->pyr_ev[f->async_gen % dslam_session::kPyrEvents], 0));
  } else {
    DSLAM_CUDA(cudaEventRecord(f->built_ev, s->stream));
    DSLAM_CUDA(cudaStreamWaitEvent(f->copy_stream, f->built_ev, 0));
  }
  // contiguous destination (one block for all levels) -> one DMA per array instead of one per level
  for (int pass = 0; pass < 2; pass++) {
    float *const *dst = pass == 0 ? host_dIp : host_absgrad;
    if (!dst) continue;
    const size_t mul = pass == 0 ? 3 : 1;
    const float *src = pass == 0 ? f->stage_dIp : f->stage_abs;
    bool contiguous = true;
    for (int l = 0; l < f->levels; l++)
      if (!dst[l] || dst[l] != dst[0] + mul * f->px_off[l]) contiguous = false;
    if (contiguous) {
      DSLAM_CUDA(cudaMemcpyAsync(dst[0], src, mul * f->px_off[f->levels] * sizeof(float), cudaMemcpyDeviceToHost, f->copy_stream));
    } else {
      for (int l = 0; l < f->levels; l++)
        if (dst[l])
          DSLAM_CUDA(cudaMemcpyAsync(dst[l], src + mul * f->px_off[l], mul * (f->px_off[l + 1] - f->px_off[l]) * sizeof(float),
                                     cudaMemcpyDeviceToHost, f->copy_stream));
    }
  }
  DSLAM_CUDA(cudaEventRecord(f->host_ready, f->copy_stream));
  f->host_pending = true;
  return DSLAM_OK;
}

int dslam_frame_download(dslam_frame *f, float *const *host_dIp, float *const *host_absgrad) {
  if (!f) return fail(DSLAM_EINVAL, "null frame");
  if (!f->built) return fail(DSLAM_ESTATE, "frame pyramid has not been built");
  if (!host_dIp && !host_absgrad) return DSLAM_OK;
  const bool need_d = host_dIp != nullptr, need_a = host_absgrad != nullptr;
  const bool have = (!need_d || f->staged_dIp) && (!need_a || f->staged_abs);
  if (!have) {  // the requested array was not staged by the current build: unpack it from the texels
    const int ra = frame_acquire(f);
    if (ra != DSLAM_OK) return ra;
    const int rc = frame_ensure_staging(f, need_d, need_a);
    if (rc != DSLAM_OK) return rc;
    const bool un_d = need_d && !f->staged_dIp, un_a = need_a && !f->staged_abs;
    const int rp = frame_prepare(f, nullptr, un_d, un_a);  // the unpack kernel writes only the arrays that are missing
    if (rp != DSLAM_OK) return rp;
    FrameBatch B;
    B.G = f->geom;
    B.f[0] = f->dev;
    DSLAM_CUDA(launch_unpack(B, 1, f->s->stream));
    f->s->launches++;
    f->staged = true;
    f->staged_dIp = f->staged_dIp || un_d;
    f->staged_abs = f->staged_abs || un_a;
  }
  return frame_copy_out(f, host_dIp, host_absgrad);
}

int dslam_frame_wait_host(dslam_frame *f) {
  if (!f) return fail(DSLAM_EINVAL, "null frame");
  if (f->host_pending) DSLAM_CUDA(cudaEventSynchronize(f->host_ready));
  return DSLAM_OK;
}

int dslam_frame_make_images(dslam_frame *f, const float *color, const float *B256, float *const *host_dIp, float *const *host_absgrad) {
  int rc = dslam_frame_upload(f, color);
  if (rc != DSLAM_OK) return rc;
  rc = frames_build(1, &f, B256, host_dIp != nullptr, host_absgrad != nullptr);
  if (rc != DSLAM_OK) return rc;
  if (host_dIp || host_absgrad) return frame_copy_out(f, host_dIp, host_absgrad);
  return DSLAM_OK;
}

// ---- tracker object ----------------------------------------------------------------------------------
static void ctx_make_K(dslam_ctx *c, const float K0[4]) {
  c->cam0.set(c->levels, K0[0], K0[1], K0[2], K0[3]);  // makeK :117-141
  double Rd[9];
  hm::qmatrix(c->T_f1_f0.q, Rd);
  float Rf[9];
  for (int k = 0; k < 9; k++) Rf[k] = (float)Rd[k];
  for (int l = 0; l < c->levels; l++) {
    const float K[9] = {c->cam0.fx[l], 0.0f, c->cam0.cx[l], 0.0f, c->cam0.fy[l], c->cam0.cy[l], 0.0f, 0.0f, 1.0f};
    hm::mat33f_inverse(K, c->Ki[l]);
    hm::mat33f_mul(Rf, c->Ki[l], c->M_stereo[l]);  // rot_f1_f0_K0_i :1022-1024
  }
}

int dslam_ctx_create(dslam_session *s, int w, int h, int levels, const float K0[4], const float K1[4], const double T_f1_f0[16],
                     dslam_ctx **out) {
  if (!s || !out || !K0 || !K1 || !T_f1_f0) return fail(DSLAM_EINVAL, "null argument");
  *out = nullptr;
  if (w < 8 || h < 8 || levels < 1 || levels > DSLAM_MAX_LEVELS) return fail(DSLAM_EINVAL, "bad geometry");
  DSLAM_CUDA(cudaSetDevice(s->device));
  dslam_ctx *c = new (std::nothrow) dslam_ctx();
  if (!c) return fail(DSLAM_ENOMEM, "out of host memory");
  c->s = s;
  c->levels = levels;
  c->px_off[0] = 0;
  for (int l = 0; l < levels; l++) {
    c->w[l] = w >> l;
    c->h[l] = h >> l;
    c->px_off[l + 1] = c->px_off[l] + (size_t)c->w[l] * c->h[l];
  }
  // SE3(Matrix4d): quaternion from the rotation block, translation from the last column (:82-86)
  c->T_f1_f0.q = hm::qfrom_matrix(T_f1_f0, 4);
  c->T_f1_f0.t[0] = T_f1_f0[3]; c->T_f1_f0.t[1] = T_f1_f0[7]; c->T_f1_f0.t[2] = T_f1_f0[11];
  c->cam1.set(levels, K1[0], K1[1], K1[2], K1[3]);  // :88-98
  ctx_make_K(c, K0);
  float4 *block = nullptr;
  cudaError_t e = cudaMalloc((void **)&block, c->px_off[levels] * sizeof(float4));
  if (e != cudaSuccess) {
    delete c;
    return cuda_fail(e, "cudaMalloc(template)");
  }
  for (int l = 0; l < levels; l++) c->pts[l] = block + c->px_off[l];
  e = cudaHostAlloc((void **)&c->pts_stage, c->px_off[levels] * sizeof(float4), cudaHostAllocDefault);
  if (e != cudaSuccess) {
    cudaFree(block);
    delete c;
    return cuda_fail(e, "cudaHostAlloc(template staging)");
  }
  *out = c;
  return DSLAM_OK;
}

int dslam_ctx_destroy(dslam_ctx *c) {
  if (!c) return DSLAM_OK;
  cudaSetDevice(c->s->device);
  cudaStreamSynchronize(c->s->stream);
  cudaFree(c->pts[0]);
  cudaFreeHost(c->pts_stage);
  cudaFree(c->grid_block);
  cudaFree(c->scan_block);
  cudaFree(c->pt_stage_dev);
  delete c;
  return DSLAM_OK;
}

int dslam_ctx_make_K(dslam_ctx *c, const float K0[4]) {
  if (!c || !K0) return fail(DSLAM_EINVAL, "null argument");
  ctx_make_K(c, K0);
  return DSLAM_OK;
}

int dslam_ctx_set_affine_mode(dslam_ctx *c, int modeA, int modeB) {
  if (!c) return fail(DSLAM_EINVAL, "null context");
  c->affModeA = modeA;
  c->affModeB = modeB;
  return DSLAM_OK;
}

int dslam_ref_upload(dslam_ctx *c, int lvl, int n, const float *u, const float *v, const float *idepth, const float *color) {
  if (!c) return fail(DSLAM_EINVAL, "null context");
  if (lvl < 0 || lvl >= c->levels) return fail(DSLAM_EINVAL, "level %d out of range", lvl);
  if (n < 0 || (size_t)n > c->px_off[lvl + 1] - c->px_off[lvl]) return fail(DSLAM_EINVAL, "pc_n %d exceeds w*h of level %d", n, lvl);
  if (n > 0 && (!u || !v || !idepth || !color)) return fail(DSLAM_EINVAL, "null template array");
  // the staging slice of this level may still be in flight from the previous upload
  DSLAM_CUDA(cudaStreamSynchronize(c->s->stream));
  float4 *st = c->pts_stage + c->px_off[lvl];
  for (int i = 0; i < n; i++) st[i] = make_float4(u[i], v[i], idepth[i], color[i]);
  if (n > 0) DSLAM_CUDA(cudaMemcpyAsync(c->pts[lvl], st, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, c->s->stream));
  c->pc_n[lvl] = n;
  c->have_ref = true;
  return DSLAM_OK;
}

int dslam_ref_set_affine(dslam_ctx *c, float ref_exposure, double ref_a, double ref_b) {
  if (!c) return fail(DSLAM_EINVAL, "null context");
  c->ref_exposure = ref_exposure;
  c->ref_a = ref_a;
  c->ref_b = ref_b;
  return DSLAM_OK;
}

int dslam_ref_scale_idepth(dslam_ctx *c, float scale) {
  if (!c) return fail(DSLAM_EINVAL, "null context");
  if (!c->have_ref) return fail(DSLAM_ESTATE, "no tracking reference uploaded");
  for (int l = 0; l < c->levels; l++)
    if (c->pc_n[l] > 0) {
      DSLAM_CUDA(launch_scale_idepth(c->pts[l], c->pc_n[l], scale, c->s->stream));
      c->s->launches++;
    }
  return DSLAM_OK;
}

int dslam_ref_build(dslam_ctx *c, dslam_frame *ref_frame, int npts, const int *pu, const int *pv, const float *pidepth, const float *pweight,
                    int *pc_n_out) {
  if (!c || !ref_frame) return fail(DSLAM_EINVAL, "null argument");
  if (npts < 0 || (npts > 0 && (!pu || !pv || !pidepth || !pweight))) return fail(DSLAM_EINVAL, "bad point export");
  if (c->s != ref_frame->s) return fail(DSLAM_EINVAL, "context and frame belong to different sessions");
  if (!ref_frame->built) return fail(DSLAM_ESTATE, "reference frame pyramid has not been built");
  if (ref_frame->w != c->w[0] || ref_frame->h != c->h[0] || ref_frame->levels < c->levels) return fail(DSLAM_EINVAL, "frame geometry mismatch");
  dslam_session *s = c->s;
  DSLAM_CUDA(cudaSetDevice(s->device));
  {
    const int ra = frame_acquire(ref_frame);
    if (ra != DSLAM_OK) return ra;
  }
  const size_t tot = c->px_off[c->levels];
  TemplateGrids G{};
  G.levels = c->levels;
  int nblocks = 0;
  for (int l = 0; l < c->levels; l++) {
    G.w[l] = c->w[l];
    G.h[l] = c->h[l];
    G.cblock_begin[l] = nblocks;
    nblocks += template_compact_blocks(c->w[l], c->h[l]);
  }
  for (int l = c->levels; l <= kMaxLevels; l++) G.cblock_begin[l] = nblocks;
  if (!c->grid_block) DSLAM_CUDA(cudaMalloc((void **)&c->grid_block, 3 * tot * sizeof(float)));
  if (!c->scan_block) DSLAM_CUDA(cudaMalloc((void **)&c->scan_block, (size_t)(2 * nblocks + kMaxLevels + 8) * sizeof(int)));
  for (int l = 0; l < c->levels; l++) {
    G.idepth[l] = c->grid_block + c->px_off[l];
    G.wsum[l] = c->grid_block + tot + c->px_off[l];
    G.wsum2[l] = c->grid_block + 2 * tot + c->px_off[l];
    G.tex[l] = ref_frame->L.tex[l];
    G.out[l] = c->pts[l];
  }
  if (npts > c->pt_stage_cap) {
    cudaFree(c->pt_stage_dev);
    c->pt_stage_dev = nullptr;
    const int cap = npts + npts / 2 + 1024;
    DSLAM_CUDA(cudaMalloc(&c->pt_stage_dev, (size_t)cap * 16));
    c->pt_stage_cap = cap;
  }
  int *d_pu = (int *)c->pt_stage_dev, *d_pv = d_pu + c->pt_stage_cap;
  float *d_id = (float *)(d_pv + c->pt_stage_cap), *d_w = d_id + c->pt_stage_cap;
  if (npts > 0) {
    DSLAM_CUDA(cudaMemcpyAsync(d_pu, pu, (size_t)npts * 4, cudaMemcpyHostToDevice, s->stream));
    DSLAM_CUDA(cudaMemcpyAsync(d_pv, pv, (size_t)npts * 4, cudaMemcpyHostToDevice, s->stream));
    DSLAM_CUDA(cudaMemcpyAsync(d_id, pidepth, (size_t)npts * 4, cudaMemcpyHostToDevice, s->stream));
    DSLAM_CUDA(cudaMemcpyAsync(d_w, pweight, (size_t)npts * 4, cudaMemcpyHostToDevice, s->stream));
  }
  int *block_counts = c->scan_block, *block_offsets = c->scan_block + nblocks, *pc_n_dev = c->scan_block + 2 * nblocks;
  int nl = 0;
  DSLAM_CUDA(launch_template_build(G, d_pu, d_pv, d_id, d_w, npts, block_counts, block_offsets, pc_n_dev, &nl, s->stream));
  s->launches += nl;
  int pcn[kMaxLevels] = {0};
  DSLAM_CUDA(cudaMemcpyAsync(pcn, pc_n_dev, sizeof(int) * c->levels, cudaMemcpyDeviceToHost, s->stream));
  DSLAM_CUDA(cudaStreamSynchronize(s->stream));
  for (int l = 0; l < c->levels; l++) {
    c->pc_n[l] = pcn[l];
    if (pc_n_out) pc_n_out[l] = pcn[l];
  }
  c->have_ref = true;
  return DSLAM_OK;
}

int dslam_ref_download(dslam_ctx *c, int lvl, int *n_out, float *u, float *v, float *idepth, float *color) {
  if (!c || !n_out) return fail(DSLAM_EINVAL, "null argument");
  if (lvl < 0 || lvl >= c->levels) return fail(DSLAM_EINVAL, "level %d out of range", lvl);
  const int n = c->pc_n[lvl];
  *n_out = n;
  if (!u && !v && !idepth && !color) return DSLAM_OK;
  std::vector<float4> tmp((size_t)n);
  if (n > 0) {
    DSLAM_CUDA(cudaMemcpyAsync(tmp.data(), c->pts[lvl], (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost, c->s->stream));
    DSLAM_CUDA(cudaStreamSynchronize(c->s->stream));
  }
  for (int i = 0; i < n; i++) {
    if (u) u[i] = tmp[i].x;
    if (v) v[i] = tmp[i].y;
    if (idepth) idepth[i] = tmp[i].z;
    if (color) color[i] = tmp[i].w;
  }
  return DSLAM_OK;
}

// ---- batched evaluation ----------------------------------------------------------------------------
int dslam_pose_eval(dslam_ctx *c, dslam_frame *f, float new_exposure, int lvl, int nb, const double *pose7, const double *aff_ab,
                    float cutoffTH, double *H64, double *b8, double *res6, int *n_padded, double *acc48) {
  int rc = check_ctx_frame(c, f, 0);
  if (rc != DSLAM_OK) return rc;
  if (lvl < 0 || lvl >= c->levels) return fail(DSLAM_EINVAL, "level %d out of range", lvl);
  if (nb < 1 || !pose7 || !aff_ab) return fail(DSLAM_EINVAL, "bad batch");
  std::vector<EvalItem> items;
  std::vector<EvalOut> outs;
  for (int base = 0; base < nb; base += kResultSlots) {
    const int cnt = nb - base < kResultSlots ? nb - base : kResultSlots;
    items.assign((size_t)cnt, EvalItem());
    for (int i = 0; i < cnt; i++)
      fill_pose_item(items[i], c, f, new_exposure, lvl, hm::Se3::from7(pose7 + 7 * (size_t)(base + i)), aff_ab + 2 * (size_t)(base + i), cutoffTH);
    rc = run_items(c->s, items, outs, &c->n_launches);
    if (rc != DSLAM_OK) return rc;
    c->n_evals += cnt;
    for (int i = 0; i < cnt; i++) {
      const size_t g = (size_t)base + i;
      double H[64], b[8];
      pose_normal_equations(outs[i], H, b);
      if (H64) std::memcpy(H64 + 64 * g, H, sizeof(H));
      if (b8) std::memcpy(b8 + 8 * g, b, sizeof(b));
      if (res6) std::memcpy(res6 + 6 * g, outs[i].res6, sizeof(double) * 6);
      if (n_padded) n_padded[g] = outs[i].n_padded;
      if (acc48) std::memcpy(acc48 + 48 * g, outs[i].acc, sizeof(double) * 48);
    }
  }
  return DSLAM_OK;
}

int dslam_scale_eval(dslam_ctx *c, dslam_frame *f_right, int lvl, int nb, const float *scales, float cutoffTH, float *Hb2, double *res6,
                     int *n_padded, double *acc8) {
  int rc = check_ctx_frame(c, f_right, 0);
  if (rc != DSLAM_OK) return rc;
  if (lvl < 0 || lvl >= c->levels) return fail(DSLAM_EINVAL, "level %d out of range", lvl);
  if (nb < 1 || !scales) return fail(DSLAM_EINVAL, "bad batch");
  std::vector<EvalItem> items;
  std::vector<EvalOut> outs;
  for (int base = 0; base < nb; base += kResultSlots) {
    const int cnt = nb - base < kResultSlots ? nb - base : kResultSlots;
    items.assign((size_t)cnt, EvalItem());
    for (int i = 0; i < cnt; i++) fill_scale_item(items[i], c, f_right, lvl, scales[base + i], cutoffTH);
    rc = run_items(c->s, items, outs, &c->n_launches);
    if (rc != DSLAM_OK) return rc;
    c->n_evals += cnt;
    for (int i = 0; i < cnt; i++) {
      const size_t g = (size_t)base + i;
      if (Hb2) scale_normal_equations(outs[i], Hb2 + 2 * g, Hb2 + 2 * g + 1);
      if (res6) std::memcpy(res6 + 6 * g, outs[i].res6, sizeof(double) * 6);
      if (n_padded) n_padded[g] = outs[i].n_padded;
      if (acc8) std::memcpy(acc8 + 8 * g, outs[i].acc, sizeof(double) * 8);
    }
  }
  return DSLAM_OK;
}

// ---- Levenberg-Marquardt drivers ---------------------------------------------------------------------
// n_pose tracking jobs and n_scale scale-optimisation jobs (independent of each other: different tracker objects,
// frames or streams) advanced in lock step; every round of ALL of them is one kernel launch.
int dslam_lm_batch(int n_pose, dslam_ctx *const *pose_ctxs, dslam_frame *const *pose_frames, const float *new_exposure, double *pose7_io,
                   double *aff_io, int coarsestLvl, const double minResForAbort[5], double *lastResiduals, double *flow3, int *ok, int n_scale,
                   dslam_ctx *const *scale_ctxs, dslam_frame *const *scale_frames, float *scales_io, int scale_coarsestLvl, float *rmse_out) {
  if (n_pose < 0 || n_scale < 0 || n_pose + n_scale < 1 || n_pose + n_scale > kResultSlots) return fail(DSLAM_EINVAL, "bad job counts");
  if (n_pose > 0 && (!pose_ctxs || !pose_frames || !pose7_io || !aff_io || !minResForAbort)) return fail(DSLAM_EINVAL, "null pose argument");
  if (n_scale > 0 && (!scale_ctxs || !scale_frames || !scales_io)) return fail(DSLAM_EINVAL, "null scale argument");
  dslam_session *s = n_pose > 0 ? (pose_ctxs[0] ? pose_ctxs[0]->s : nullptr) : (scale_ctxs[0] ? scale_ctxs[0]->s : nullptr);
  for (int i = 0; i < n_pose; i++) {
    const int rc = check_ctx_frame(pose_ctxs[i], pose_frames[i], coarsestLvl);
    if (rc != DSLAM_OK) return rc;
    if (pose_ctxs[i]->s != s) return fail(DSLAM_EINVAL, "all jobs of a batch must live in one session");
  }
  for (int i = 0; i < n_scale; i++) {
    const int rc = check_ctx_frame(scale_ctxs[i], scale_frames[i], scale_coarsestLvl);
    if (rc != DSLAM_OK) return rc;
    if (scale_ctxs[i]->s != s) return fail(DSLAM_EINVAL, "all jobs of a batch must live in one session");
  }
  std::vector<PoseLM> pose((size_t)n_pose);
  std::vector<ScaleLM> scale((size_t)n_scale);
  for (int i = 0; i < n_pose; i++) pose_ctxs[i]->trace.clear();
  for (int i = 0; i < n_scale; i++) scale_ctxs[i]->trace.clear();
  // the machines of a batch run on several host threads: a tracker object that appears more than once (hypotheses of one
  // frame passed as separate jobs) records the trace of its FIRST job only — one writer per trace vector
  std::vector<const dslam_ctx *> traced;
  for (int i = 0; i < n_pose; i++) {
    const bool first = std::find(traced.begin(), traced.end(), pose_ctxs[i]) == traced.end();
    if (first) traced.push_back(pose_ctxs[i]);
    pose[i].begin(pose_ctxs[i], pose_frames[i], new_exposure ? new_exposure[i] : 1.0f, pose7_io + 7 * i, aff_io + 2 * i, coarsestLvl, minResForAbort,
                  first ? &pose_ctxs[i]->trace : nullptr);
  }
  for (int i = 0; i < n_scale; i++) {
    // a tracker object that also tracks in this batch keeps the pose trace; its scale trace is not recorded
    const bool first = std::find(traced.begin(), traced.end(), scale_ctxs[i]) == traced.end();
    if (first) traced.push_back(scale_ctxs[i]);
    scale[i].begin(scale_ctxs[i], scale_frames[i], scales_io[i], scale_coarsestLvl, first ? &scale_ctxs[i]->trace : nullptr);
  }
  dslam_ctx *counters = n_pose > 0 ? pose_ctxs[0] : scale_ctxs[0];
  const int rc = run_lock_step(s, pose, scale, counters);
  if (rc != DSLAM_OK) return rc;
  for (int i = 0; i < n_pose; i++) {
    pose_ctxs[i]->n_iters += pose[i].iters;
    if (lastResiduals) std::memcpy(lastResiduals + 5 * i, pose[i].lastResiduals, sizeof(double) * 5);
    if (flow3) std::memcpy(flow3 + 3 * i, pose[i].flow, sizeof(double) * 3);
    const bool good = finish_pose(pose_ctxs[i], pose[i], new_exposure ? new_exposure[i] : 1.0f, pose7_io + 7 * i, aff_io + 2 * i);
    if (ok) ok[i] = good ? 1 : 0;
  }
  for (int i = 0; i < n_scale; i++) {
    scale_ctxs[i]->n_iters += scale[i].iters;
    scales_io[i] = scale[i].scale_current;                 // :954
    if (rmse_out) rmse_out[i] = scale[i].last_residuals[0];  // :963
  }
  return DSLAM_OK;
}

int dslam_track_newest_coarse_batch(int n, dslam_ctx *const *ctxs, dslam_frame *const *frames, const float *new_exposure, double *pose7_io,
                                    double *aff_io, int coarsestLvl, const double minResForAbort[5], double *lastResiduals, double *flow3,
                                    int *ok) {
  if (n < 1) return fail(DSLAM_EINVAL, "bad argument");
  return dslam_lm_batch(n, ctxs, frames, new_exposure, pose7_io, aff_io, coarsestLvl, minResForAbort, lastResiduals, flow3, ok, 0, nullptr, nullptr,
                        nullptr, 0, nullptr);
}

int dslam_optimize_scale_batch(int n, dslam_ctx *const *ctxs, dslam_frame *const *frames_right, float *scales_io, int coarsestLvl,
                               float *rmse_out) {
  if (n < 1) return fail(DSLAM_EINVAL, "bad argument");
  return dslam_lm_batch(0, nullptr, nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr, n, ctxs, frames_right, scales_io,
                        coarsestLvl, rmse_out);
}

// several pose hypotheses of ONE tracker object / frame (the retry loop of FrontEnd::trackNewCoarse, src/FrontEnd.cpp:147-247)
int dslam_track_newest_coarse_multi(dslam_ctx *c, dslam_frame *f, float new_exposure, int nhyp, double *pose7_io, double *aff_io,
                                    int coarsestLvl, const double minResForAbort[5], double *lastResiduals, double *flow3, int *ok) {
  int rc = check_ctx_frame(c, f, coarsestLvl);
  if (rc != DSLAM_OK) return rc;
  if (nhyp < 1 || nhyp > kResultSlots || !pose7_io || !aff_io || !minResForAbort) return fail(DSLAM_EINVAL, "bad argument");
  c->trace.clear();
  std::vector<PoseLM> lms((size_t)nhyp);
  std::vector<ScaleLM> none;
  for (int i = 0; i < nhyp; i++) lms[i].begin(c, f, new_exposure, pose7_io + 7 * i, aff_io + 2 * i, coarsestLvl, minResForAbort, i == 0 ? &c->trace : nullptr);
  rc = run_lock_step(c->s, lms, none, c);
  if (rc != DSLAM_OK) return rc;
  for (int i = 0; i < nhyp; i++) {
    c->n_iters += lms[i].iters;
    if (lastResiduals) std::memcpy(lastResiduals + 5 * i, lms[i].lastResiduals, sizeof(double) * 5);
    if (flow3) std::memcpy(flow3 + 3 * i, lms[i].flow, sizeof(double) * 3);
    const bool good = finish_pose(c, lms[i], new_exposure, pose7_io + 7 * i, aff_io + 2 * i);
    if (ok) ok[i] = good ? 1 : 0;
  }
  return DSLAM_OK;
}

int dslam_track_newest_coarse(dslam_ctx *c, dslam_frame *f, float new_exposure, double pose7_io[7], double aff_io[2], int coarsestLvl,
                              const double minResForAbort[5], double lastResiduals[5], double flow3[3], int *ok) {
  return dslam_track_newest_coarse_multi(c, f, new_exposure, 1, pose7_io, aff_io, coarsestLvl, minResForAbort, lastResiduals, flow3, ok);
}

// several scale seeds of ONE tracker object / right frame (the seed loop of FrontEnd::optimizeScale, src/FrontEnd.cpp:995-1003)
int dslam_optimize_scale_multi(dslam_ctx *c, dslam_frame *f_right, int nseeds, float *scales_io, int coarsestLvl, float *rmse_out) {
  int rc = check_ctx_frame(c, f_right, coarsestLvl);
  if (rc != DSLAM_OK) return rc;
  if (nseeds < 1 || nseeds > kResultSlots || !scales_io) return fail(DSLAM_EINVAL, "bad argument");
  c->trace.clear();
  std::vector<PoseLM> none;
  std::vector<ScaleLM> lms((size_t)nseeds);
  for (int i = 0; i < nseeds; i++) lms[i].begin(c, f_right, scales_io[i], coarsestLvl, i == 0 ? &c->trace : nullptr);
  rc = run_lock_step(c->s, none, lms, c);
  if (rc != DSLAM_OK) return rc;
  for (int i = 0; i < nseeds; i++) {
    c->n_iters += lms[i].iters;
    scales_io[i] = lms[i].scale_current;                   // :954
    if (rmse_out) rmse_out[i] = lms[i].last_residuals[0];  // :963
  }
  return DSLAM_OK;
}

int dslam_optimize_scale(dslam_ctx *c, dslam_frame *f_right, float *scale_io, int coarsestLvl, float *rmse_out) {
  return dslam_optimize_scale_multi(c, f_right, 1, scale_io, coarsestLvl, rmse_out);
}

// The hypothesis loop of FrontEnd::trackNewCoarse (src/FrontEnd.cpp:192-247) with the retries evaluated speculatively in
// lock-step batches.  A trial does not depend on `achievedRes` except for the abort test at the end of every level
// (TrackerAndScaler.cpp:597-598), so each trial is run without that test, its per-level (residual, flow) events are
// recorded, and the sequential loop is then REPLAYED on the recorded events with the evolving achievedRes: an aborted
// trial is cut at the level where the reference would have returned false.  The outputs are exactly those of the
// sequential loop.  Trials are evaluated in stages (1, then 4, then all the rest) and a stage only runs if the replay
// reaches it — in the common case the first hypothesis ends the loop and nothing else is evaluated.
int dslam_track_new_coarse(dslam_ctx *c, dslam_frame *f, float new_exposure, int ntries, const double *pose7_tries, const double aff_init[2],
                           int coarsestLvl, const double last_coarse_rmse[5], double reTrackThreshold, double pose7_out[7], double aff_out[2],
                           double achievedRes_out[5], double flow3_out[3], int *haveOneGood_out, int *tryIterations_out) {
  int rc = check_ctx_frame(c, f, coarsestLvl);
  if (rc != DSLAM_OK) return rc;
  if (ntries < 1 || ntries > kResultSlots || !pose7_tries || !aff_init || !last_coarse_rmse || !pose7_out || !aff_out || !achievedRes_out)
    return fail(DSLAM_EINVAL, "bad argument");
  c->trace.clear();
  const double never[5] = {NAN, NAN, NAN, NAN, NAN};  // "x > 1.5 * NaN" is false: no abort while speculating
  std::vector<PoseLM> lms((size_t)ntries);
  std::vector<ScaleLM> none;
  std::vector<double> p0((size_t)ntries * 7), a0((size_t)ntries * 2);
  int computed = 0;
  auto compute_until = [&](int upto) -> int {  // evaluate trials [computed, upto) in one lock step
    if (upto > ntries) upto = ntries;
    if (upto <= computed) return DSLAM_OK;
    std::vector<PoseLM> batch((size_t)(upto - computed));
    for (int i = computed; i < upto; i++) {
      std::memcpy(&p0[7 * (size_t)i], pose7_tries + 7 * (size_t)i, sizeof(double) * 7);
      a0[2 * (size_t)i] = aff_init[0];
      a0[2 * (size_t)i + 1] = aff_init[1];
      batch[i - computed].begin(c, f, new_exposure, &p0[7 * (size_t)i], &a0[2 * (size_t)i], coarsestLvl, never, i == 0 ? &c->trace : nullptr);
    }
    const int r = run_lock_step(c->s, batch, none, c);
    if (r != DSLAM_OK) return r;
    for (int i = computed; i < upto; i++) {
      lms[i] = std::move(batch[i - computed]);
      c->n_iters += lms[i].iters;
    }
    computed = upto;
    return DSLAM_OK;
  };
  double achieved[5] = {NAN, NAN, NAN, NAN, NAN};
  double flow[3] = {100, 100, 100};
  double pose[7] = {0, 0, 0, 1, 0, 0, 0}, aff[2] = {0, 0};
  bool haveOneGood = false;
  int tries = 0;
  const int stage_end[3] = {1, 5, ntries};
  for (int i = 0; i < ntries; i++) {
    if (i >= computed) {
      int upto = ntries;
      for (int st = 0; st < 3; st++)
        if (i < stage_end[st]) { upto = stage_end[st]; break; }
      rc = compute_until(upto);
      if (rc != DSLAM_OK) return rc;
    }
    PoseLM &m = lms[i];
    // replay trackNewestCoarse(..., achievedRes, currentRes) on the recorded level events
    double cur[5] = {NAN, NAN, NAN, NAN, NAN}, fl[3] = {1000, 1000, 1000};
    bool aborted = false;
    for (const PoseLM::LevelEvent &e : m.events) {
      cur[e.lvl] = e.residual;
      fl[0] = e.flow[0]; fl[1] = e.flow[1]; fl[2] = e.flow[2];
      if (e.residual > 1.5 * achieved[e.lvl]) {  // TrackerAndScaler.cpp:597-598
        aborted = true;
        break;
      }
    }
    double pose_i[7], aff_i[2] = {aff_init[0], aff_init[1]};
    std::memcpy(pose_i, pose7_tries + 7 * (size_t)i, sizeof(pose_i));
    bool good = false;
    if (!aborted) good = finish_pose(c, m, new_exposure, pose_i, aff_i);
    tries++;
    if (good && std::isfinite((float)cur[0]) && !(cur[0] >= achieved[0])) {  // FrontEnd.cpp:222-229
      flow[0] = fl[0]; flow[1] = fl[1]; flow[2] = fl[2];
      aff[0] = aff_i[0]; aff[1] = aff_i[1];
      std::memcpy(pose, pose_i, sizeof(pose));
      haveOneGood = true;
    }
    if (haveOneGood)  // :232-239
      for (int k = 0; k < 5; k++)
        if (!std::isfinite((float)achieved[k]) || achieved[k] > cur[k]) achieved[k] = cur[k];
    if (haveOneGood && achieved[0] < last_coarse_rmse[0] * reTrackThreshold) break;  // :241-243
  }
  if (!haveOneGood) {  // :246-252
    flow[0] = flow[1] = flow[2] = 0;
    aff[0] = aff_init[0]; aff[1] = aff_init[1];
    std::memcpy(pose, pose7_tries, sizeof(pose));
  }
  std::memcpy(pose7_out, pose, sizeof(pose));
  aff_out[0] = aff[0]; aff_out[1] = aff[1];
  std::memcpy(achievedRes_out, achieved, sizeof(achieved));
  if (flow3_out) std::memcpy(flow3_out, flow, sizeof(flow));
  if (haveOneGood_out) *haveOneGood_out = haveOneGood ? 1 : 0;
  if (tryIterations_out) *tryIterations_out = tries;
  return DSLAM_OK;
}

// ---------------------------------------------------------------------------------------------------
// dso::PoseEstimator  src/loop_closure/pose_estimation/PoseEstimator.cpp  (loop-closure direct alignment)
// ---------------------------------------------------------------------------------------------------
int dslam_pe_create(dslam_session *s, int w, int h, int levels, dslam_pe **out) {
  if (!s || !out) return fail(DSLAM_EINVAL, "null argument");
  *out = nullptr;
  if (w < 8 || h < 8 || levels < 1 || levels > DSLAM_MAX_LEVELS) return fail(DSLAM_EINVAL, "bad geometry");
  dslam_pe *p = new (std::nothrow) dslam_pe();
  if (!p) return fail(DSLAM_ENOMEM, "out of host memory");
  dslam_ctx &c = p->ctx;
  c.s = s;
  c.levels = levels;
  c.point3d = true;
  for (int l = 0; l < levels; l++) {  // PoseEstimator::PoseEstimator :36-52
    c.w[l] = w >> l;
    c.h[l] = h >> l;
  }
  *out = p;
  return DSLAM_OK;
}

int dslam_pe_destroy(dslam_pe *p) {
  if (!p) return DSLAM_OK;
  cudaSetDevice(p->ctx.s->device);
  cudaStreamSynchronize(p->ctx.s->stream);
  cudaFree(p->block);
  cudaFreeHost(p->stage);
  delete p;
  return DSLAM_OK;
}

int dslam_pe_set_affine_mode(dslam_pe *p, int modeA, int modeB) {
  if (!p) return fail(DSLAM_EINVAL, "null estimator");
  p->ctx.affModeA = modeA;
  p->ctx.affModeB = modeB;
  return DSLAM_OK;
}

// pts_ = pts; ref_ab_exposure_ = ref_ab_exposure; ref_aff_g2l_ = AffLight()  :307-317
int dslam_pe_set_points(dslam_pe *p, int n, const double *pts_xyz, const float *colors, float ref_ab_exposure) {
  if (!p || !pts_xyz || !colors) return fail(DSLAM_EINVAL, "null argument");
  if (n < 1) return fail(DSLAM_EINVAL, "no points");
  dslam_ctx &c = p->ctx;
  dslam_session *s = c.s;
  DSLAM_CUDA(cudaSetDevice(s->device));
  const int levels = c.levels;
  if (n > p->cap) {
    DSLAM_CUDA(cudaStreamSynchronize(s->stream));
    cudaFree(p->block);
    cudaFreeHost(p->stage);
    p->block = nullptr;
    p->stage = nullptr;
    p->cap = 0;
    const int cap = (n + 1023) & ~1023;
    DSLAM_CUDA(cudaMalloc((void **)&p->block, (size_t)levels * cap * sizeof(float4)));
    DSLAM_CUDA(cudaHostAlloc((void **)&p->stage, (size_t)levels * cap * sizeof(float4), cudaHostAllocDefault));
    p->cap = cap;
  } else {
    DSLAM_CUDA(cudaStreamSynchronize(s->stream));  // the staging buffer may still feed the previous upload
  }
  for (int l = 0; l < levels; l++) {
    float4 *dst = p->stage + (size_t)l * n;
    for (int i = 0; i < n; i++)  // :168-171  the point as floats, pts_[i].second[lvl]
      dst[i] = make_float4((float)pts_xyz[3 * (size_t)i], (float)pts_xyz[3 * (size_t)i + 1], (float)pts_xyz[3 * (size_t)i + 2],
                           colors[(size_t)i * levels + l]);
    c.pts[l] = p->block + (size_t)l * n;
    c.pc_n[l] = n;
  }
  DSLAM_CUDA(cudaMemcpyAsync(p->block, p->stage, (size_t)levels * n * sizeof(float4), cudaMemcpyHostToDevice, s->stream));
  c.ref_exposure = ref_ab_exposure;
  c.ref_a = c.ref_b = 0;
  c.have_ref = true;
  return DSLAM_OK;
}

static int pe_prepare(dslam_pe *p, dslam_frame *f, const float new_cam[4], int lvl) {
  if (!p || !new_cam) return fail(DSLAM_EINVAL, "null argument");
  p->ctx.cam0.set(p->ctx.levels, new_cam[0], new_cam[1], new_cam[2], new_cam[3]);  // makeK(new_cam) :66-83
  return check_ctx_frame(&p->ctx, f, lvl);
}

int dslam_pe_eval(dslam_pe *p, dslam_frame *new_fh, float new_exposure, const float new_cam[4], int lvl, const double T_ref_to_new[16],
                  const double aff_ab[2], float cutoffTH, double *H64, double *b8, double *res6, int *n_padded, double *acc48) {
  int rc = pe_prepare(p, new_fh, new_cam, 0);
  if (rc != DSLAM_OK) return rc;
  if (!T_ref_to_new || !aff_ab) return fail(DSLAM_EINVAL, "null argument");
  hm::Se3 T;
  T.q = hm::qfrom_matrix(T_ref_to_new, 4);
  T.t[0] = T_ref_to_new[3]; T.t[1] = T_ref_to_new[7]; T.t[2] = T_ref_to_new[11];
  double p7[7];
  T.to7(p7);
  return dslam_pose_eval(&p->ctx, new_fh, new_exposure, lvl, 1, p7, aff_ab, cutoffTH, H64, b8, res6, n_padded, acc48);
}

int dslam_pe_estimate(dslam_pe *p, dslam_frame *new_fh, float new_exposure, const float new_cam[4], int coarsest_lvl, double ref_to_new_io[16],
                      float *pose_error, int *inlier_percent, int *ok) {
  int rc = pe_prepare(p, new_fh, new_cam, coarsest_lvl);
  if (rc != DSLAM_OK) return rc;
  if (!ref_to_new_io || !pose_error || !ok) return fail(DSLAM_EINVAL, "null argument");
  dslam_ctx *c = &p->ctx;
  c->trace.clear();
  // SE3(ref_to_new.block<3,3>(0,0), ref_to_new.block<3,1>(0,3))  :321-322
  hm::Se3 T;
  T.q = hm::qfrom_matrix(ref_to_new_io, 4);
  T.t[0] = ref_to_new_io[3]; T.t[1] = ref_to_new_io[7]; T.t[2] = ref_to_new_io[11];
  double p7[7];
  T.to7(p7);
  const double aff0[2] = {0, 0};  // aff_g2l_current = AffLight()  :319
  const double inf = std::numeric_limits<double>::infinity();
  const double no_abort[5] = {inf, inf, inf, inf, inf};  // estimate() has no minResForAbort test
  std::vector<PoseLM> lms(1);
  std::vector<ScaleLM> none;
  lms[0].begin(c, new_fh, new_exposure, p7, aff0, coarsest_lvl, no_abort, &c->trace);
  rc = run_lock_step(c->s, lms, none, c);
  if (rc != DSLAM_OK) return rc;
  PoseLM &m = lms[0];
  c->n_iters += m.iters;
  // ref_to_new = refToNew_current.matrix()  :463
  double R[9];
  hm::qmatrix(m.cur.q, R);
  for (int r = 0; r < 3; r++) {
    for (int k = 0; k < 3; k++) ref_to_new_io[r * 4 + k] = R[r * 3 + k];
    ref_to_new_io[r * 4 + 3] = m.cur.t[r];
  }
  ref_to_new_io[12] = ref_to_new_io[13] = ref_to_new_io[14] = 0;
  ref_to_new_io[15] = 1;
  *pose_error = (float)m.lastResiduals[0];  // :464
  // :466-478 affine sanity, the same test as the tracker's epilogue
  bool aff_good = true;
  const double *aff = m.aff_cur;
  if ((c->affModeA != 0 && (fabsf((float)aff[0]) > 1.2)) || (c->affModeB != 0 && (fabsf((float)aff[1]) > 200))) aff_good = false;
  double rel[2];
  hm::aff_from_to(c->ref_exposure, new_exposure, c->ref_a, c->ref_b, aff[0], aff[1], rel);
  const float relA = (float)rel[0], relB = (float)rel[1];
  if ((c->affModeA == 0 && (fabsf(logf(relA)) > 1.5)) || (c->affModeB == 0 && (fabsf(relB) > 200))) aff_good = false;
  int lastInners0 = 0;
  for (const PoseLM::LevelEvent &e : m.events)
    if (e.lvl == 0) lastInners0 = e.numTermsInE;
  const bool low_res = *pose_error < 10.0;                                  // RES_THRES  PoseEstimator.h:25
  const int percent = 100 * float(lastInners0) / c->pc_n[0];                // :480
  const bool enough_inlier = percent > 90;                                  // INNER_PERCENT  PoseEstimator.h:26
  if (inlier_percent) *inlier_percent = percent;
  *ok = (aff_good && low_res && enough_inlier) ? 1 : 0;
  return DSLAM_OK;
}

int dslam_pe_get_trace(dslam_pe *p, double *rows, int max_rows, int *rows_out) {
  if (!p) return fail(DSLAM_EINVAL, "null estimator");
  return dslam_get_trace(&p->ctx, rows, max_rows, rows_out);
}

int dslam_get_trace(dslam_ctx *c, double *rows, int max_rows, int *rows_out) {
  if (!c || !rows_out) return fail(DSLAM_EINVAL, "null argument");
  const int n = (int)(c->trace.size() / 15);
  *rows_out = n;
  if (rows)
    for (int i = 0; i < n && i < max_rows; i++) std::memcpy(rows + 15 * i, c->trace.data() + 15 * (size_t)i, sizeof(double) * 15);
  return DSLAM_OK;
}

int dslam_ctx_counters(dslam_ctx *c, long long out[3]) {
  if (!c || !out) return fail(DSLAM_EINVAL, "null argument");
  out[0] = c->n_evals;
  out[1] = c->n_launches;
  out[2] = c->n_iters;
  return DSLAM_OK;
}

}  // extern "C"
