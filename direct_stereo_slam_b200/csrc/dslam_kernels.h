// dslam_kernels.h — internal launch interface between the host runtime (dslam_api.cu) and the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>

namespace dslam {

constexpr int kMaxLevels = 6;

// ---------------------------------------------------------------------------------------------------
// residual / Jacobian / normal-equation kernels (kernels_residual.cu)
// ---------------------------------------------------------------------------------------------------
constexpr int kEvalThreads = 128;        // threads per CTA of the evaluation kernels
constexpr int kMaxBlocksPerItem = 96;    // partial-sum slots per work item
constexpr int kMaxItemsPerLaunch = 128;  // work items carried in kernel-parameter space per launch
constexpr int kPoseVals = 48;            // 45 upper-triangle sums + E + shiftT + shiftRT
constexpr int kScaleVals = 8;            // JwJ, Jwr, rwr, E, shiftT, shiftRT, 0, 0

// One evaluation of calcRes*+calcGSSSE* = one work item. Lives in kernel-parameter (constant) space.
struct alignas(16) EvalItem {
  const float4 *tex;  // level texels (I, dx, dy, absgrad) of the frame sampled
  const float4 *pts;  // template records (u, v, idepth, color); 3-D point records (x, y, z, color[lvl]) when flags bit2
  int n;              // pc_n[lvl]
  int w, h;           // level size
  int flags;          // bit0: accumulate flow indicators (lvl == 0); bit1: scale item; bit2: PoseEstimator 3-D point item
  float fx, fy, cx, cy;  // intrinsics of the camera sampled (cam0 for pose, cam1 for scale)
  float M[9];         // pose: R*Ki ; scale: R_f1_f0*Ki ; 3-D point item: R
  float t[3];         // pose: translation ; scale: t_f1_f0
  float Ki[9];        // K^-1 of the level (flow indicators)
  float p0, p1, p2;   // pose: affLL a, affLL b, b0 ; scale: scale, unused, unused
  float cutoff, maxEnergy;
  int nblocks;        // CTAs working on this item
  int ppt_stride;     // = nblocks * kEvalThreads (grid stride)
  int cta_begin;      // first CTA of this item in the flat 1-D grid of the launch (items in order, no idle CTAs)
};

struct EvalBatch {
  EvalItem item[kMaxItemsPerLaunch];
};

// Result record written by the last CTA of an item straight into mapped pinned host memory.  Every 8-byte word
// carries the launch sequence number in its low half and 32 payload bits in its high half, so the host needs no
// flag and the kernel needs no system-scope fence (which costs ~10 us of idle GPU per launch): the host simply
// waits until every word of the record shows the current sequence number.  8-byte aligned stores are not torn.
//   w[2*i], w[2*i+1] : high / low 32 bits of acc[i]  (i < kPoseVals)
//   w[96..98]        : numTermsInE, numSaturated, numTermsInWarped (unpadded)
constexpr int kResultWords = 128;
constexpr int kResultCountBase = 2 * kPoseVals;
struct alignas(128) EvalResult {
  unsigned long long w[kResultWords];
};
static_assert(sizeof(EvalResult) == 1024, "EvalResult layout");

// device scratch of a session: per item-slot partial sums, counters and tickets
struct EvalScratch {
  double *partials;   // [kMaxItemsPerLaunch][kMaxBlocksPerItem][kPoseVals]
  int *counters;      // [kMaxItemsPerLaunch][4]: nE, nSat, nInl, ticket
};

// ---- resident evaluation server (kernels_residual.cu: eval_server_kernel) -----------------------------------------------------
constexpr int kSrvLanes = 32;              // doorbells: one per LM lane (group x half)
constexpr unsigned kSrvQueueSlots = 16384;  // work-unit queue in HBM (a lane round is <= 512 units, <= kSrvLanes rounds in flight)
constexpr int kSrvMaxUnits = 512;          // CTA-sized work units per lane round
// doorbell of one lane in mapped pinned memory (its own 128-byte line).  The host fills items_host / units_host of the lane
// and n_items / n_units / seq, then stores `round` (release).  doors[0].stop = generation number asks the server to leave.
struct alignas(128) ServerDoor {
  unsigned round, n_items, n_units, seq;
  unsigned stop;
  unsigned pad[27];
};
static_assert(sizeof(ServerDoor) == 128, "ServerDoor layout");
struct ServerLane {
  double *partials;                // reduction scratch of the lane (EvalScratch)
  int *counters;
  EvalResult *results;             // device alias of the lane's slice of the mapped result ring
  const EvalItem *items_host;      // mapped pinned: items of the lane's current round
  const unsigned *units_host;      // mapped pinned: (item << 16 | CTA index) per work unit
};
struct ServerParams {
  int nlanes;
  unsigned gen;                    // generation of this server (tags queue slots; doors[0].stop == gen stops it)
  const volatile ServerDoor *doors;  // mapped pinned
  ServerLane lane[kSrvLanes];
  unsigned last_round[kSrvLanes];  // doorbell values at launch
  unsigned long long *queue;       // [kSrvQueueSlots]
  unsigned *qctl;                  // [0] next ticket, [1] next free index (zeroed before every launch)
  EvalItem *items_dev;             // [kSrvLanes][kMaxItemsPerLaunch]
  unsigned *lane_seq;              // [kSrvLanes]
  unsigned long long idle_ns, life_ns;
};
cudaError_t launch_eval_server(const ServerParams &P, int workers, cudaStream_t stream);

// mode 0 = pose (8-DoF), 1 = scale (1-DoF). results_dev = device alias of the mapped EvalResult array.
cudaError_t launch_eval(int mode, const EvalBatch &batch, int nitems, int total_ctas, EvalScratch scratch, EvalResult *results_dev,
                        unsigned seq, cudaStream_t stream);

// ---------------------------------------------------------------------------------------------------
// pyramid kernels (kernels_pyramid.cu)
// ---------------------------------------------------------------------------------------------------
struct PyramidLevels {
  int levels;
  int w[kMaxLevels], h[kMaxLevels];
  int pitch[kMaxLevels];        // floats per row of the intensity planes (multiple of 4)
  float *plane[kMaxLevels];     // intensity planes; plane[0] = uploaded image
  float4 *tex[kMaxLevels];      // texels (I, dx, dy, absgrad), dense pitch = w
  float *host_dIp[kMaxLevels];  // optional staging in the reference host layout (3 floats AoS); may be null
  float *host_abs[kMaxLevels];  // optional staging float plane; may be null
  int tile_begin[kMaxLevels + 1];  // prefix of gradient tiles per level (kernel B)
  int tiles_x[kMaxLevels];
};
constexpr int kGradTileW = 64, kGradTileH = 16;
constexpr int kGradBoxW = kGradTileW + 8;  // 72 floats = 288 B; the box starts at x0-4 because TMA needs a 16-B aligned source address
constexpr int kGradBoxH = kGradTileH + 2;
constexpr int kDownTileW = 64, kDownTileH = 32;  // level-0 tile of the box-mean chain

// Device-resident description of one frame (pointers + tensor maps); the pyramid kernels take a batch of these so
// that all frames of a step are built by two launches in total.
struct alignas(128) FrameDev {
  CUtensorMap map[kMaxLevels];  // first: keeps the 64-B descriptor alignment
  float *plane[kMaxLevels];
  float4 *tex[kMaxLevels];
  float *host_dIp[kMaxLevels];  // staging copies in the reference host layouts, or null
  float *host_abs[kMaxLevels];
  const float *B256;            // gamma table or null
};
constexpr int kMaxFramesPerLaunch = 64;
struct PyramidGeom {
  int levels;
  int w[kMaxLevels], h[kMaxLevels], pitch[kMaxLevels];
  int tile_begin[kMaxLevels + 1];
  int tiles_x[kMaxLevels];
};
struct FrameBatch {
  PyramidGeom G;
  const FrameDev *f[kMaxFramesPerLaunch];
};
// all frames of the batch share the geometry G
cudaError_t launch_downsample(const FrameBatch &B, int nframes, cudaStream_t stream);
cudaError_t launch_gradients(const FrameBatch &B, int nframes, cudaStream_t stream, int ctas_per_sm = 6);
// texels -> the reference's host layouts (Vector3f AoS + float plane) for a frame built without staging
cudaError_t launch_unpack(const FrameBatch &B, int nframes, cudaStream_t stream);
// contiguous arena of nframes raw images (4 * quads floats each) -> dense level-0 planes
struct PlaneBatch {
  float *plane[kMaxFramesPerLaunch];
};
cudaError_t launch_scatter_planes(const PlaneBatch &B, int nframes, int quads, const float *arena, cudaStream_t stream);
cudaError_t launch_scale_idepth(float4 *pts, int n, float scale, cudaStream_t stream);

// ---------------------------------------------------------------------------------------------------
// template kernels (kernels_template.cu): makeCoarseDepthL0 / scaleCoarseDepthL0 on the device
// ---------------------------------------------------------------------------------------------------
struct TemplateGrids {
  int levels;
  int w[kMaxLevels], h[kMaxLevels];
  float *idepth[kMaxLevels];      // idepth_[lvl]
  float *wsum[kMaxLevels];        // weight_sums_[lvl] before dilation (plays weight_sums_bak_)
  float *wsum2[kMaxLevels];       // weight_sums_[lvl] after dilation
  const float4 *tex[kMaxLevels];  // texels of the reference keyframe (colour = .x)
  float4 *out[kMaxLevels];        // compacted template records (u, v, idepth, color)
  int cblock_begin[kMaxLevels + 1];  // prefix of 1024-pixel compaction blocks per level
};
int template_compact_blocks(int w, int h);
// makeCoarseDepthL0 from a flat device export of the active points; pc_n_dev[levels] receives the counts
cudaError_t launch_template_build(const TemplateGrids &G, const int *pu, const int *pv, const float *pid, const float *pw, int npts,
                                  int *block_counts, int *block_offsets, int *pc_n_dev, int *launches, cudaStream_t stream);
cudaError_t launch_scale_idepth(float4 *pts, int n, float scale, cudaStream_t stream);

// ---------------------------------------------------------------------------------------------------
// Scan-Context kernels (kernels_sc.cu)
// ---------------------------------------------------------------------------------------------------
constexpr int kScTopK = 8;
// ring-key scan: per query the k (<=8) nearest rows by squared L2 in flann::L2 arithmetic, as packed
// (float_bits(dist) << 32 | global_id) keys sorted ascending; out[nq][kScTopK], unused = ~0ull
cudaError_t launch_sc_ringkey(const float *keys, const int *ids, int n_rows, int dim, const float *queries, int nq, int max_id,
                              unsigned long long *out, unsigned long long *scratch, cudaStream_t stream);
// sector-cosine scan: per query and CTA the top-kScTopK packed (approximate dist, LOCAL ROW) keys over rows with id < max_id
// and (thres < 0 or ring dist < thres), written to scratch as lists[q][sc_list_stride()][K] (*nlists_out lists are valid per
// query); the re-score kernel merges them.  ids must ascend with the row so that ties still resolve to the lowest id.
cudaError_t launch_sc_scan(const float *sigs, const float *keys, const int *ids, int n_rows, int n_cells, int key_dim,
                           const float *q_sigs, const float *q_keys, int nq, float ringkey_thres, int max_id, float sc_width,
                           unsigned long long *scratch, int *nlists_out, float *q_split /* sc_qsplit_floats(nq, n_cells) floats, or null */, cudaStream_t stream);
int sc_list_stride();
size_t sc_scratch_bytes(int nq);
size_t sc_qsplit_floats(int cap, int n_cells);  // scratch of the tensor-core scan for batches of up to cap queries
// scan kernel selection: 0 = by batch size (default; DSLAM_SC_SCAN=stream|tile|umma overrides), 1 = streaming, 2 = tiled, 3 = tcgen05 (3xTF32)
void sc_set_scan_flavour(int flavour);
// ScanContext::generate on the device: moments (mean[3], cov[6]) of an n x 3 fp64 cloud; then binning + normalisation
// with the PCA axes the host derived from them.  sig = dense fp32 [num_s * num_r], sig64 (optional) the fp64 values.
cudaError_t launch_sc_moments(const double *pts, int n, double *out9, cudaStream_t stream);
cudaError_t launch_sc_bin_finalize(const double *pts, int n, const double mean[3], const double v9[9], double lidar_range, int num_s, int num_r,
                                   unsigned long long *cells, float *ringkey, float *sig, double *sig64, cudaStream_t stream);
// NVLink mailbox exchange of the per-query best keys (kernels_sc.cu: sc_rescore_topk_kernel).  keys[r] / arrived[r] are THIS
// rank's mappings of rank r's mailbox (own allocation for r == rank, CUDA IPC mappings otherwise).
constexpr int kScXchgMaxQ = 1024;
constexpr unsigned long long kScXchgErrorKey = ~0ull - 1;  // published instead of a key when a peer did not arrive in time
struct ScExchange {
  int world = 1, rank = 0;
  unsigned long long *keys[8] = {};
  unsigned *arrived[8] = {};
};
size_t sc_exchange_bytes();  // size of one rank's mailbox allocation; keys first, counters at sc_exchange_arrived_offset()
size_t sc_exchange_arrived_offset();
// exact re-score of the scan's survivors in search_sc's arithmetic -> per-query best packed (dist, GLOBAL id) key, published
// to host_words (mapped pinned, 2 self-validating words per query: high / low 32 key bits | seq) — after the mailbox
// exchange when xchg->world > 1.  fp64: sigs / q_sigs are double tables (the reference's SigType values).
cudaError_t launch_sc_rescore_topk(const unsigned long long *lists, int nlists, const void *sigs, int fp64, const int *ids, const void *q_sigs, int nq, int n_cells,
                                   int sc_width, unsigned long long *exact_keys, unsigned long long *best, unsigned long long *host_words,
                                   unsigned seq, const ScExchange *xchg, unsigned xchg_seq, int q0, unsigned *ticket, cudaStream_t stream);
// exact distances of explicit (query, local row) pairs; row < 0 = skip (+inf)
cudaError_t launch_sc_rescore_pairs(const int *pair_q, const int *pair_row, int npairs, const void *sigs, int fp64, const void *q_sigs, int n_cells,
                                    int sc_width, float *diff_out, cudaStream_t stream);
cudaError_t launch_sc_widen(const float *src, double *dst, size_t n, cudaStream_t stream);
cudaError_t launch_sc_narrow(const double *src, float *dst, size_t n, cudaStream_t stream);

}  // namespace dslam
