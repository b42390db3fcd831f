// dslam_internal.h — host-side object layouts shared by the translation units of libdslam_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/dslam_b200.h"
#include "dslam_kernels.h"
#include "host_math.h"

namespace dslam {

// ---- error plumbing -------------------------------------------------------------------------------
void set_error(const char *fmt, ...);
int fail(int code, const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);

#define DSLAM_CUDA(call)                                        \
  do {                                                          \
    cudaError_t e__ = (call);                                   \
    if (e__ != cudaSuccess) return ::dslam::cuda_fail(e__, #call); \
  } while (0)

constexpr int kResultSlots = 2048;  // pinned, mapped EvalResult ring of a session (16 launches of 128 items)

}  // namespace dslam

// One CUDA stream plus the pinned result ring the evaluation kernels publish into.
struct dslam_session {
  int device = 0;
  cudaStream_t stream = nullptr;
  // Asynchronous pyramid builds (dslam_frame_build_batch, stage_host bit 2) run on their own stream so that the pyramids
  // of the NEXT frames are built while the LM rounds of the current ones run.  Every build records one event of a ring;
  // a frame remembers the generation of its build and any later use of it orders the session stream behind that event.
  static constexpr int kPyrEvents = 8;
  cudaStream_t pyr_stream = nullptr;
  int prio_hi = 0;
  int pyr_async_ctas = 3;  // CTAs per SM of the gradient kernel in asynchronous builds (DSLAM_PYR_ASYNC_CTAS)
  cudaEvent_t pyr_in = nullptr, pyr_ev[kPyrEvents] = {};
  unsigned long long pyr_gen = 0, pyr_waited = 0;
  float *upload_arena = nullptr;  // device staging of dslam_frame_upload_batch (one H2D for all images of a step)
  size_t upload_arena_floats = 0;
  dslam::EvalResult *results_host = nullptr;  // cudaHostAllocMapped
  dslam::EvalResult *results_dev = nullptr;   // device alias of results_host
  dslam::EvalScratch scratch{nullptr, nullptr};
  std::atomic<unsigned> seq{0};        // sequence number of the last evaluation launch group
  std::atomic<long long> launches{0};  // kernels of this library launched on the session's streams
  cudaEvent_t mark[2] = {nullptr, nullptr};
  int num_sms = 148;
  double timeout_s = 20.0;
  // Lock-step LM rounds: the machines of a call are dealt into `lm_groups` groups; every group has its own CUDA stream,
  // reduction scratch, slice of the result ring and HOST THREAD, and runs "prepare -> launch -> wait -> consume" on its
  // own.  The kernels of different groups overlap on the GPU (each is latency-bound and fills a fraction of the SMs) and
  // the host-side LM algebra + launch overhead (which bounds a single thread at ~15 us per round) runs in parallel.
  // Group 0 uses `stream` / `scratch` and the calling thread.
  static constexpr int kLmGroups = 16;         // hard cap (DSLAM_LM_GROUPS)
  static constexpr int kLmGroupsDefault = 8;   // cap of the automatic choice
  int lm_groups = 4;  // DSLAM_LM_GROUPS (1..kLmGroups)
  cudaStream_t lm_stream[kLmGroups] = {};
  dslam::EvalScratch lm_scratch[kLmGroups] = {};
  cudaEvent_t lm_done[kLmGroups] = {};
  // second launch lane of every group: a group deals its machines into two halves and keeps one half on the GPU while the
  // host consumes / prepares the other (DSLAM_LM_HALVES=1 disables)
  cudaStream_t lm_stream2[kLmGroups] = {};
  dslam::EvalScratch lm_scratch2[kLmGroups] = {};
  cudaEvent_t lm_done2[kLmGroups] = {};
  int lm_halves = 2;
  // Speculation depth over runs of rejected LM steps (DSLAM_LM_SPEC, 1 = off): after a rejection the next candidates depend only on
  // (H, b, lambda), so up to lm_spec_depth - 1 of them are evaluated in the same round; every machine still consumes exactly the
  // sequential sequence of evaluations (exact replay), a round trip is saved per rejection of a run.
  int lm_spec_depth = 4;
  // (Anticipating the FIRST rejection of a run as well — one extra candidate next to every ordinary step — was measured and does
  // not reduce the round count further: a run of r rejections needs 1 + ceil((r-1)/depth) rounds either way.)
  cudaEvent_t lm_fork = nullptr;
  struct Worker {
    std::thread th;
    std::mutex m;
    std::condition_variable cv;
    std::function<void()> job;
    bool has_job = false, done = false, quit = false;
  };
  Worker *workers[kLmGroups] = {};  // [0] unused (the caller's thread)
  // Resident evaluation server (kernels_residual.cu: eval_server_kernel): for the duration of a lock-step LM call one kernel
  // stays on the GPU; the lanes ring doorbells in mapped pinned memory instead of launching a kernel per LM round.
  // OFF by default (DSLAM_LM_SERVER=1 turns it on): measured on B200 / PCIe Gen5 the doorbell protocol costs more than the
  // launch it replaces — the SMs must poll host memory (>= 2 dependent PCIe read round trips of ~1.5 us per round: doorbell,
  // then items + units) while cudaLaunchKernel is a posted write plus one front-end fetch: S = 1 tracking 0.585 ms vs 0.498 ms,
  // 512 streams 76-83 k vs 95 k frames/s (profiles/r02_resident_server.md).  Results are bit-identical either way
  // (tests/test_gpu_batch.py); the launch path is also used while per-launch profiling is on.
  bool srv_enabled = false;
  int srv_ctas_per_sm = 3;                     // resident worker CTAs per SM (DSLAM_LM_SERVER_CTAS); the rest of the SM stays free for the pyramid stream
  dslam::ServerDoor *srv_doors = nullptr;      // pinned mapped [kSrvLanes]
  dslam::ServerDoor *srv_doors_dev = nullptr;  // device alias
  dslam::EvalItem *srv_items = nullptr, *srv_items_alias = nullptr;    // pinned mapped [kSrvLanes][kMaxItemsPerLaunch] + device alias
  unsigned *srv_units = nullptr, *srv_units_alias = nullptr;            // pinned mapped [kSrvLanes][kSrvMaxUnits] + device alias
  unsigned long long *srv_queue = nullptr;     // device
  unsigned *srv_qctl = nullptr, *srv_lane_seq = nullptr;
  dslam::EvalItem *srv_items_dev = nullptr;
  unsigned srv_round[dslam::kSrvLanes] = {};   // last doorbell value written per lane
  unsigned srv_gen = 0;
  std::atomic<long long> srv_rounds{0};        // LM rounds served through doorbells (diagnostics)
  std::mutex prof_mutex;
  // host-side time split of the lock-step driver (ns, summed over the group threads)
  std::atomic<long long> t_prep_ns{0}, t_launch_ns{0}, t_wait_ns{0}, n_rounds{0};
  // optional per-launch profiling of the evaluation kernels (CUDA events on this stream around every launch)
  bool prof_on = false;
  std::vector<cudaEvent_t> prof_ev;     // pairs
  std::vector<int> prof_mode;           // per pair: 0 pose, 1 scale
  std::vector<long long> prof_points;   // per pair: template points evaluated by the launch (sum over items)
  size_t prof_used = 0;                 // pairs recorded since the last read
};

struct dslam_frame {
  dslam_session *s = nullptr;
  int w = 0, h = 0, levels = 0;
  dslam::PyramidLevels L{};    // host view: sizes, pitches, device pointers
  dslam::PyramidGeom geom{};   // the part of it the kernels take by value
  dslam::FrameDev devh{};      // host copy of the device-resident descriptor (pointers + tensor maps)
  dslam::FrameDev *dev = nullptr;
  bool dev_dirty = true;
  void *block = nullptr;       // one allocation: intensity planes + texels
  float *stage_dIp = nullptr;  // device staging in the reference host layout, all levels contiguous (lazy)
  float *stage_abs = nullptr;
  float *B_dev = nullptr;      // 256-float gamma table (lazy)
  size_t px_off[dslam::kMaxLevels + 1]{};
  bool uploaded = false, built = false, staged = false;
  bool staged_dIp = false, staged_abs = false;  // host-layout copies filled by the CURRENT build (or unpacked since)
  cudaEvent_t host_ready = nullptr;  // recorded on copy_stream after the D2H of the host mirrors
  cudaEvent_t built_ev = nullptr;    // recorded on the session stream after the kernels that fill the staging copies
  cudaStream_t copy_stream = nullptr;  // D2H of the host mirrors overlaps the tracking kernels of the session stream
  bool host_pending = false;
  unsigned long long async_gen = 0;  // generation of the asynchronous build that last wrote this frame (0 = none)
};

struct dslam_ctx {
  dslam_session *s = nullptr;
  int w[dslam::kMaxLevels]{}, h[dslam::kMaxLevels]{}, levels = 0;
  dslam::hm::CamPyramid cam0{}, cam1{};
  float Ki[dslam::kMaxLevels][9]{};
  dslam::hm::Se3 T_f1_f0{};
  float M_stereo[dslam::kMaxLevels][9]{};  // R_f1_f0 * Ki[lvl]
  // template (pc_u, pc_v, pc_idepth, pc_color packed as float4), one device array per level
  float4 *pts[dslam::kMaxLevels]{};
  int pc_n[dslam::kMaxLevels]{};
  float4 *pts_stage = nullptr;  // pinned host staging for uploads, level l at px_off[l]
  size_t px_off[dslam::kMaxLevels + 1]{};
  bool have_ref = false;
  bool point3d = false;  // PoseEstimator flavour: pts[lvl] hold (x, y, z, color[lvl]) of the matched keyframe's points
  float ref_exposure = 1.f;
  double ref_a = 0, ref_b = 0;
  int affModeA = 0, affModeB = 0;
  // device scratch of dslam_ref_build
  float *grid_block = nullptr;  // idepth / weight sums / backup grids, all levels
  int *scan_block = nullptr;
  void *pt_stage_dev = nullptr;
  int pt_stage_cap = 0;
  std::vector<double> trace;  // rows of 15 doubles
  long long n_evals = 0, n_launches = 0, n_iters = 0;
};

// dso::PoseEstimator (src/loop_closure/pose_estimation/PoseEstimator.h): the 8-DoF machinery of dslam_ctx over 3-D points
struct dslam_pe {
  dslam_ctx ctx;
  int cap = 0;              // points per level the device / staging buffers hold
  float4 *block = nullptr;  // levels * cap records
  float4 *stage = nullptr;  // pinned
};
