/* dslam_b200.h — C ABI of libdslam_b200.so: the B200-native (sm_100a) implementation of the photometric
 * Gauss-Newton hot path of IRVLab/direct_stereo_slam.
 *
 * Each entry point names the reference interface it replaces ("src/..." = /root/reference/src/...,
 * "deps:dso/..." = dso/ inside the reference's dependencies.zip).  INTEGRATION.md shows the C++ glue a
 * maintainer adds at the reference's call sites (src/FrontEnd.cpp:57-58, 204-206, 605, 680, 797-798, 992,
 * 997, 1032; src/loop_closure/LoopHandler.cpp:247, 256).
 *
 * Conventions
 *  - every function returns DSLAM_OK (0) or a negative DSLAM_E* code; nothing throws; dslam_last_error()
 *    gives the text of the most recent failure on the calling thread.
 *  - all pointers are caller-owned HOST memory unless the parameter is an opaque handle.
 *  - a dslam_session owns one CUDA stream; every object created in a session is stream-ordered on it and the
 *    session is not re-entrant (one host thread at a time), exactly like the reference's single tracking
 *    thread (src/FrontEnd.cpp:589 track_mutex_).  Different sessions are independent and may run
 *    concurrently from different host threads.
 *  - poses are 7 doubles (qx,qy,qz,qw,tx,ty,tz) == Sophus::SE3d::data(); affine brightness is (a,b) doubles
 *    == dso::AffLight.
 *  - there is NO CPU fallback: without a CUDA device every compute entry point returns DSLAM_ENODEVICE.
 */
#ifndef DSLAM_B200_H_
#define DSLAM_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define DSLAM_MAX_LEVELS 6 /* PYR_LEVELS, deps:dso/src/util/settings.h:50 */

enum {
  DSLAM_OK = 0,
  DSLAM_EINVAL = -1,    /* bad argument */
  DSLAM_ENODEVICE = -2, /* no CUDA device / driver */
  DSLAM_ECUDA = -3,     /* CUDA runtime error (see dslam_last_error) */
  DSLAM_ENOMEM = -4,
  DSLAM_ESTATE = -5,    /* call sequence error (e.g. tracking before a reference was uploaded) */
  DSLAM_ENCCL = -6,     /* NCCL unavailable or failed */
  DSLAM_ETIMEOUT = -7   /* device did not answer (kernel fault or hang) */
};

typedef struct dslam_session dslam_session; /* one CUDA stream + pinned result ring */
typedef struct dslam_frame dslam_frame;     /* device twin of one dso::FrameHessian image pyramid */
typedef struct dslam_ctx dslam_ctx;         /* device twin of one dso::TrackerAndScaler */
typedef struct dslam_scdb dslam_scdb;       /* Scan-Context descriptor database (one shard per rank) */

int dslam_version(void);
const char *dslam_last_error(void);
int dslam_device_count(int *count);

/* ---- sessions ------------------------------------------------------------------------------------ */
int dslam_session_create(int device, dslam_session **out);
int dslam_session_destroy(dslam_session *s);
int dslam_session_sync(dslam_session *s);
/* cudaStream_t of the session, for callers that time with CUDA events on the launching stream */
int dslam_session_stream(dslam_session *s, void **cuda_stream_out);
/* number of kernels of this library launched on the session so far */
int dslam_session_launch_count(dslam_session *s, long long *count);
/* record a CUDA event on the session stream now / elapsed ms between marks 0 and 1 (after both completed) */
int dslam_session_mark(dslam_session *s, int which);
int dslam_session_elapsed_ms(dslam_session *s, float *ms);
/* per-launch profiling of the residual kernels: CUDA events on the session stream around every launch.
 * read -> out[0..3] pose-only launches (launches, total ms, template points evaluated, max ms), out[4..7] scale-only
 * launches, out[8..11] mixed pose+scale launches; reading resets the counters. */
int dslam_session_profile(dslam_session *s, int enable);
int dslam_session_profile_read(dslam_session *s, double out[12]);
/* host-side time split of the lock-step LM driver since the last call, in ms: out[0] LM algebra + item preparation,
 * out[1] kernel launches, out[2] waiting for results, out[3] number of launches */
int dslam_session_host_times(dslam_session *s, double out[4]);
/* Diagnostics (pure host logic, needs no device): how one evaluation launch would be laid out.  n_points[i] = template points of
 * work item i (n_items <= 128), num_sms = SMs of the device, lanes = launches of this kind in flight at the same time.  Outputs:
 * nblocks[i] = CTAs of item i, cta_begin[i] = its first CTA in the flat grid, *total_ctas = grid size. */
int dslam_plan_eval_launch(int n_items, const int *n_points, int num_sms, int lanes, int *nblocks, int *cta_begin, int *total_ctas);
/* pinned host memory helpers (cudaHostAlloc) so callers can make H2D/D2H truly asynchronous */
int dslam_host_alloc(unsigned long long bytes, void **out);
int dslam_host_free(void *p);

/* ---- image pyramid: replaces FrameHessian::makeImages(float* color, CalibHessian* HCalib) ------------
 * deps:dso/src/FullSystem/HessianBlocks.cpp:128-191, called at src/FrontEnd.cpp:605 (left) and :680 (right).
 * Level l has size (w>>l) x (h>>l).  Device layout: one float4 texel (I, dx, dy, absSquaredGrad) per pixel.
 * Rows 0 and h_l-1 of dx, dy, absSquaredGrad are written as 0 (the reference leaves them uninitialised). */
int dslam_frame_create(dslam_session *s, int w, int h, int levels, dslam_frame **out);
int dslam_frame_destroy(dslam_frame *f);
/* H2D of the level-0 image (w*h floats, 0..255), asynchronous on the session stream */
int dslam_frame_upload(dslam_frame *f, const float *color);
/* the same for n frames of one session.  Runs of images that lie back to back in host memory (colors[i+1] == colors[i] + w*h:
 * a capture ring / pinned arena) travel as ONE host-to-device transfer and are dealt to the frames on the device; other
 * images are uploaded one by one.  Asynchronous on the session stream; the host images must stay valid until it has drained. */
int dslam_frame_upload_batch(int n, dslam_frame *const *frames, const float *const *colors);
/* build all levels on the device from the uploaded image.  B256 = CalibHessian::B (256 floats) applies the
 * gamma weight gw^2 to absSquaredGrad ("HCalib != 0 && setting_gammaWeightsPixelSelect == 1"); NULL = none. */
int dslam_frame_build(dslam_frame *f, const float *B256);
/* the same for n uploaded frames of one session in two kernel launches in total (per run of 64 frames of equal geometry);
 * stage_host bit 0 / bit 1 also fill the device-side staging copies of dIp / absSquaredGrad so that a following
 * dslam_frame_download is a plain D2H copy.
 * stage_host bit 2 = asynchronous build: the kernels run on the session's pyramid stream, ordered behind everything queued
 * on the session stream so far (the uploads of these frames, earlier tracking calls that read these frame objects), and every
 * later call that takes one of these frames is ordered behind the build — but tracking calls on OTHER frames queued
 * afterwards run concurrently with it.  This is how the pyramids of frame k+1 are built while frame k is being tracked. */
int dslam_frame_build_batch(int n, dslam_frame *const *frames, const float *B256, int stage_host);
/* asynchronous D2H into the reference's host layouts: host_dIp[l] = Eigen::Vector3f[w_l*h_l] (I,dx,dy AoS),
 * host_absgrad[l] = float[w_l*h_l]; either array (or single entries) may be NULL. */
int dslam_frame_download(dslam_frame *f, float *const *host_dIp, float *const *host_absgrad);
/* block until the downloads of this frame have landed */
int dslam_frame_wait_host(dslam_frame *f);
/* the makeImages drop-in: upload + build + download (download is asynchronous; call dslam_frame_wait_host
 * before host code reads dIp/absSquaredGrad) */
int dslam_frame_make_images(dslam_frame *f, const float *color, const float *B256, float *const *host_dIp,
                            float *const *host_absgrad);

/* ---- tracker object: replaces dso::TrackerAndScaler -----------------------------------------------
 * src/scale_optimization/TrackerAndScaler.h:34-137.  K0/K1 = (fx, fy, cx, cy) of camera 0 / camera 1 at level
 * 0; T_f1_f0 = row-major 4x4 (the tfm_vec of the constructor, TrackerAndScaler.cpp:47-109). */
int dslam_ctx_create(dslam_session *s, int w, int h, int levels, const float K0[4], const float K1[4],
                     const double T_f1_f0[16], dslam_ctx **out);
int dslam_ctx_destroy(dslam_ctx *c);
/* makeK(CalibHessian*)  TrackerAndScaler.cpp:117-141 (camera-0 intrinsics change when DSO re-optimises them) */
int dslam_ctx_make_K(dslam_ctx *c, const float K0[4]);
/* setting_affineOptModeA / B (deps:dso/src/util/settings.cpp; 0/0 with the shipped mode=1, src/main.cpp:117-122) */
int dslam_ctx_set_affine_mode(dslam_ctx *c, int modeA, int modeB);
/* template upload after setCoarseTrackingRef (TrackerAndScaler.cpp:317-327): pc_u/pc_v/pc_idepth/pc_color of one
 * level, n = pc_n[lvl] */
int dslam_ref_upload(dslam_ctx *c, int lvl, int n, const float *u, const float *v, const float *idepth,
                     const float *color);
/* lastRef->ab_exposure and lastRef_aff_g2l (a, b) */
int dslam_ref_set_affine(dslam_ctx *c, float ref_exposure, double ref_a, double ref_b);
/* scaleCoarseDepthL0(float)  TrackerAndScaler.cpp:329-336 */
int dslam_ref_scale_idepth(dslam_ctx *c, float scale);
/* makeCoarseDepthL0 on the device (TrackerAndScaler.cpp:143-315) from a flat export of the active points:
 * integer pixel (u,v), idepth, weight; colours are read from ref_frame's level textures.  pc_n_out[levels]. */
int dslam_ref_build(dslam_ctx *c, dslam_frame *ref_frame, int npts, const int *pu, const int *pv, const float *pidepth,
                    const float *pweight, int *pc_n_out);
/* read the device template of one level back (tests); pass NULL arrays to get n only */
int dslam_ref_download(dslam_ctx *c, int lvl, int *n_out, float *u, float *v, float *idepth, float *color);

/* batched evaluation of calcResPose + calcGSSSEPose (TrackerAndScaler.cpp:699-852, 640-697) for nb pose
 * hypotheses on one level of one frame.  Outputs per hypothesis: H (8x8 row-major, scaled like :685-696),
 * b (8), res6 = Vec6 of calcResPose, n_padded = pose_buf_warped_n_; acc48 (optional) = the 45 raw sums
 * [J0..J7 r]^T w [J0..J7 r] upper triangle + (E, sumSquaredShiftT, sumSquaredShiftRT) in fp64. */
int dslam_pose_eval(dslam_ctx *c, dslam_frame *f, float new_exposure, int lvl, int nb, const double *pose7,
                    const double *aff_ab, float cutoffTH, double *H64, double *b8, double *res6, int *n_padded,
                    double *acc48);
/* batched calcResScale + calcGSSSEScale (TrackerAndScaler.cpp:1007-1172, 966-1005) for nb scales against the
 * right frame.  Hb = (H, b) floats per scale; acc8 (optional) = (JwJ, Jwr, rwr, E, shiftT, shiftRT, 0, 0). */
int dslam_scale_eval(dslam_ctx *c, dslam_frame *f_right, int lvl, int nb, const float *scales, float cutoffTH,
                     float *Hb2, double *res6, int *n_padded, double *acc8);

/* trackNewestCoarse (TrackerAndScaler.cpp:451-638; call site src/FrontEnd.cpp:204-206).  pose7_io / aff_io are
 * written only when the level loop completes (like :612-613); lastResiduals[5] is NaN-filled first; flow3 =
 * lastFlowIndicators; *ok = the bool result. */
int dslam_track_newest_coarse(dslam_ctx *c, dslam_frame *f, float new_exposure, double pose7_io[7], double aff_io[2],
                              int coarsestLvl, const double minResForAbort[5], double lastResiduals[5], double flow3[3],
                              int *ok);
/* optimizeScale (TrackerAndScaler.cpp:854-964; call sites src/FrontEnd.cpp:992, 997) */
int dslam_optimize_scale(dslam_ctx *c, dslam_frame *f_right, float *scale_io, int coarsestLvl, float *rmse_out);
/* the same LM loops for several independent starts run in lock step, one kernel launch per round for all of
 * them; each result is identical to the corresponding sequential call (SURVEY.md §8 f-4 lever).
 * scale seeds: the {0.1,1,5,10,15,25,30,50} loop of src/FrontEnd.cpp:995-1003. */
int dslam_optimize_scale_multi(dslam_ctx *c, dslam_frame *f_right, int nseeds, float *scales_io, int coarsestLvl,
                               float *rmse_out);
/* pose hypotheses with a common minResForAbort (no early exit between hypotheses) */
int dslam_track_newest_coarse_multi(dslam_ctx *c, dslam_frame *f, float new_exposure, int nhyp, double *pose7_io,
                                    double *aff_io, int coarsestLvl, const double minResForAbort[5], double *lastResiduals,
                                    double *flow3, int *ok);
/* n independent stereo streams (one tracker object and one frame each, all in one session) in lock step: every LM
 * round of all streams is one kernel launch; results are those of n sequential calls.  new_exposure may be NULL (1). */
int dslam_track_newest_coarse_batch(int n, dslam_ctx *const *ctxs, dslam_frame *const *frames, const float *new_exposure,
                                    double *pose7_io, double *aff_io, int coarsestLvl, const double minResForAbort[5],
                                    double *lastResiduals, double *flow3, int *ok);
int dslam_optimize_scale_batch(int n, dslam_ctx *const *ctxs, dslam_frame *const *frames_right, float *scales_io, int coarsestLvl,
                               float *rmse_out);
/* n_pose tracking jobs AND n_scale scale-optimisation jobs in the same lock step: the two LM loops are independent
 * (optimizeScale reads only its template and the right image), so pose and scale items share every launch. */
int dslam_lm_batch(int n_pose, dslam_ctx *const *pose_ctxs, dslam_frame *const *pose_frames, const float *new_exposure, double *pose7_io,
                   double *aff_io, int coarsestLvl, const double minResForAbort[5], double *lastResiduals, double *flow3, int *ok,
                   int n_scale, dslam_ctx *const *scale_ctxs, dslam_frame *const *scale_frames, float *scales_io, int scale_coarsestLvl,
                   float *rmse_out);
/* The retry loop of FrontEnd::trackNewCoarse (src/FrontEnd.cpp:192-247): ntries pose hypotheses (the 5 motion models +
 * 78 small rotations of :147-180), all starting from aff_init, tried in order with the evolving achievedRes as
 * minResForAbort until achievedRes[0] < last_coarse_rmse[0] * reTrackThreshold (setting_reTrackThreshold = 1.5).
 * Results are exactly those of the sequential loop; hypotheses are evaluated speculatively in lock-step batches
 * (1, then 4, then the rest) and the loop is replayed on their recorded per-level residuals.  When nothing was good
 * the outputs are (tries[0], aff_init, flow = 0) like :246-252. */
int dslam_track_new_coarse(dslam_ctx *c, dslam_frame *f, float new_exposure, int ntries, const double *pose7_tries, const double aff_init[2],
                           int coarsestLvl, const double last_coarse_rmse[5], double reTrackThreshold, double pose7_out[7], double aff_out[2],
                           double achievedRes_out[5], double flow3_out[3], int *haveOneGood_out, int *tryIterations_out);
/* per-iteration trace of the last track / optimizeScale call on this ctx (first start only for *_multi):
 * rows of 15 doubles (lvl, iteration (-1 = level start), accept, n_padded, lambda, E/n old, E/n new, inc[8]).
 * Returns the number of rows through *rows_out. */
int dslam_get_trace(dslam_ctx *c, double *rows, int max_rows, int *rows_out);
/* counters since creation: [0] residual/Jacobian evaluations (items), [1] kernel launches, [2] LM iterations */
int dslam_ctx_counters(dslam_ctx *c, long long out[3]);

/* ---- PoseEstimator: the direct alignment of a loop-closure candidate -------------------------------
 * src/loop_closure/pose_estimation/PoseEstimator.h:34-49, PoseEstimator.cpp:36-506 (call site
 * src/loop_closure/LoopHandler.cpp:274-277).  The same 8-DoF Gauss-Newton as the tracker, over the matched keyframe's
 * 3-D points (camera frame) with one stored colour per pyramid level; no abort test; the return value of estimate()
 * is aff_good && pose_error < RES_THRES(10) && inlier_percent > INNER_PERCENT(90). */
typedef struct dslam_pe dslam_pe;
int dslam_pe_create(dslam_session *s, int w, int h, int levels, dslam_pe **out);
int dslam_pe_destroy(dslam_pe *p);
/* setting_affineOptModeA / B (deps:dso/src/util/settings.cpp:83-84), as dslam_ctx_set_affine_mode */
int dslam_pe_set_affine_mode(dslam_pe *p, int modeA, int modeB);
/* pts (first argument of estimate): pts_xyz[n*3] = pair.first, colors[n*levels] = pair.second[0..levels) per point;
 * ref_ab_exposure = matched_frame->ab_exposure */
int dslam_pe_set_points(dslam_pe *p, int n, const double *pts_xyz, const float *colors, float ref_ab_exposure);
/* one calcRes + calcGSSSE (:84-296) at T_ref_to_new (row-major 4x4) / aff_ab on level lvl; outputs as dslam_pose_eval */
int dslam_pe_eval(dslam_pe *p, dslam_frame *new_fh, float new_exposure, const float new_cam[4], int lvl, const double T_ref_to_new[16],
                  const double aff_ab[2], float cutoffTH, double *H64, double *b8, double *res6, int *n_padded, double *acc48);
/* PoseEstimator::estimate :298-506 on the points of dslam_pe_set_points.  new_cam = (fx, fy, cx, cy) of level 0 of the
 * new frame; ref_to_new_io = row-major Matrix4d, updated in place; *ok = estimate()'s return value. */
int dslam_pe_estimate(dslam_pe *p, dslam_frame *new_fh, float new_exposure, const float new_cam[4], int coarsest_lvl,
                      double ref_to_new_io[16], float *pose_error, int *inlier_percent, int *ok);
int dslam_pe_get_trace(dslam_pe *p, double *rows, int max_rows, int *rows_out);

/* ---- Scan-Context database: replaces search_ringkey / search_sc ------------------------------------
 * src/loop_closure/loop_detection/search_place.h:25-57, 59-85 (call sites src/loop_closure/LoopHandler.cpp:247,
 * 256).  Descriptors: ring key = n_rings floats; signature = dense n_sectors*n_rings floats, cell index =
 * sector*n_rings + ring, 0 = empty (ScanContext.cpp:119-141).  ids are insertion order; with world_size > 1
 * row i of the global DB lives on rank i % world_size. */
int dslam_sc_create(dslam_session *s, int n_sectors, int n_rings, int capacity, dslam_scdb **out);
/* flags: DSLAM_SC_FP64 keeps, next to the fp32 table the scan streams, the signature values as DOUBLES — the reference's
 * SigType is vector<pair<int,double>> (ScanContext.h:24) and search_sc multiplies doubles (search_place.h:71-77) — and
 * the exact re-score of the K survivors reads those, so res_diff equals the reference's on unrounded signatures
 * (+9,600 B per row).  `capacity` is the initial allocation; the tables grow geometrically when it is exceeded. */
#define DSLAM_SC_FP64 1
int dslam_sc_create_ex(dslam_session *s, int n_sectors, int n_rings, int capacity, int flags, dslam_scdb **out);
int dslam_sc_destroy(dslam_scdb *db);
/* append n descriptors that belong to THIS shard; global_ids[n] strictly ascending (NULL = local running count) */
int dslam_sc_add(dslam_scdb *db, int n, const float *ringkeys, const float *sigs_dense, const int *global_ids);
/* append one descriptor in the reference's sparse form (SigType = vector<pair<int,double>>, ScanContext.h:24);
 * values are rounded to the database's fp32 storage format; global_id < 0 = previous id + 1 */
int dslam_sc_add_sparse(dslam_scdb *db, const float *ringkey, const int *idx, const double *val, int nnz, int global_id);
/* dslam_sc_add with double signatures (DSLAM_SC_FP64 databases): the doubles are stored as they are, the fp32 scan copy is
 * derived from them on the device; dslam_sc_add_sparse keeps the doubles too on such a database */
int dslam_sc_add64(dslam_scdb *db, int n, const float *ringkeys, const double *sigs_dense64, const int *global_ids);
/* On-disk dump of the shard, ".scdb": 64-byte header {"SCDB", u32 version = 1, n_sectors, n_rings, rows,
 * flags (bit 0: an fp64 table follows), 40 reserved bytes}, int32 ids[rows], fp32 ringkeys[rows][n_rings],
 * fp32 sigs[rows][n_sectors*n_rings], then fp64 sigs64[rows][...] when flagged; little endian.  dslam_sc_load APPENDS the
 * rows of a file to a database of the same geometry (ids must ascend past those present); *rows_out = rows read. */
int dslam_sc_save(dslam_scdb *db, const char *path);
int dslam_sc_load(dslam_scdb *db, const char *path, int *rows_out);
/* ScanContext::generate on the device (src/loop_closure/loop_detection/ScanContext.cpp:19-142): PCA alignment of the
 * n x 3 fp64 cloud (the 3x3 eigen-decomposition runs on the host; eigenvectors are sign-normalised so that their largest
 * component is positive), polar max-height binning, ring key, per-sector L2 normalisation.  Outputs (any may be NULL):
 * ringkey[n_rings], dense signature as fp32 / fp64 [n_sectors*n_rings] (0 = empty cell), tfm_pca_rig row-major 4x4.
 * append != 0 also appends the descriptor to the database (device to device) under global_id (< 0: previous + 1). */
int dslam_sc_generate(dslam_scdb *db, const double *pts_xyz, int n, double lidar_range, float *ringkey_out, float *sig_dense_out,
                      double *sig_dense64_out, double tfm_pca_rig[16], int append, int global_id);
int dslam_sc_size(dslam_scdb *db, int *n_local);
/* brute-force exact replacement of search_ringkey's kNN: for each of nq queries the k nearest ring keys
 * (squared L2, flann::L2 arithmetic) among ids < max_id with dist < thres; ties -> lowest id.
 * cand_out[nq*k] (-1 padded), dist_out[nq*k]. */
int dslam_sc_search_ringkey(dslam_scdb *db, int nq, const float *ringkeys, int k, float thres, int max_id, int *cand_out,
                            float *dist_out);
/* search_sc over explicit candidates (n_cand per query, -1 = skip): res_idx / res_diff per query in the
 * reference's arithmetic (float += double*double; (1 - prod/sc_width)/2; strict '>' running min from 1.1) */
int dslam_sc_search_sc(dslam_scdb *db, int nq, const float *sigs_dense, const int *candidates, int n_cand, int *res_idx,
                       float *res_diff);
/* the same with the query signatures as doubles (DSLAM_SC_FP64 databases): products are double * double rounded to double,
 * exactly search_place.h:71-77 on the reference's SigType values */
int dslam_sc_search_sc64(dslam_scdb *db, int nq, const double *sigs_dense64, const int *candidates, int n_cand, int *res_idx,
                         float *res_diff);
/* full sector-cosine scan of the whole (sharded) DB for a query batch: argmin id + distance per query among
 * ids < max_id, optionally gated by ring-key distance < ringkey_thres (pass a negative thres to disable the
 * gate).  The device scan keeps the top-K per query and re-scores those K in the reference's exact arithmetic
 * (float += double*double in cell order); the re-score kernel publishes the result straight into mapped host memory.
 * With a communicator attached (collective call: every rank passes the same batch) the per-rank winners are min-combined
 * as packed (dist,id) keys in every rank's NVLink mailbox by that same kernel (system-scope atomics over CUDA IPC peer
 * mappings; dslam_sc_exchange_mode reports 1), or — without peer access / with DSLAM_SC_EXCHANGE=nccl — by one NCCL
 * all-reduce(min).  res_idx = -1 (res_diff = 1.1) when nothing qualifies. */
int dslam_sc_query(dslam_scdb *db, int nq, const float *ringkeys, const float *sigs_dense, float ringkey_thres, int max_id,
                   int *res_idx, float *res_diff);
int dslam_sc_query64(dslam_scdb *db, int nq, const float *ringkeys, const double *sigs_dense64, float ringkey_thres, int max_id,
                     int *res_idx, float *res_diff);
int dslam_sc_exchange_mode(dslam_scdb *db, int *p2p_out);
/* the same scan without the collective: the local shard's best per query as a packed key
 * (order-preserving bits of the exact distance << 32 | global id; all-ones = nothing qualified) — what the
 * all-reduce(min) combines; dslam_sc_decode_key unpacks one (id = -1, dist = 1.1 for the empty key). */
int dslam_sc_query_keys(dslam_scdb *db, int nq, const float *ringkeys, const float *sigs_dense, float ringkey_thres, int max_id,
                        unsigned long long *keys_out);
int dslam_sc_decode_key(unsigned long long key, int *id, float *dist);
/* multi-GPU: rank 0 calls dslam_sc_unique_id (128 bytes), broadcasts it by any means (torch.distributed), then
 * every rank calls dslam_sc_comm_init. */
int dslam_sc_unique_id(unsigned char id128[128]);
int dslam_sc_comm_init(dslam_scdb *db, const unsigned char id128[128], int world_size, int rank);
/* device time of the last dslam_sc_query scan kernel(s) in ms (CUDA events on the session stream) */
int dslam_sc_last_scan_ms(dslam_scdb *db, float *ms);
/* Tuning / test knob (process-wide): which scan kernel dslam_sc_query uses.  0 = chosen per query batch (default: the
 * HBM-streaming kernel for batches of <= 8 queries or small shards; above that the tcgen05 tensor-core kernel (3xTF32, TMEM
 * accumulators) for shards of >= 2 x 128 rows per SM and the register-blocked TMA / FFMA2 tile kernel otherwise; the environment
 * variable DSLAM_SC_SCAN=stream|tile|umma|umma_masked sets the initial value), 1 = always streaming, 2 = always the FFMA tile
 * kernel, 3 = tcgen05 wherever the shard is large enough, 4 = the same with the hi half of the 3xTF32 split stored explicitly
 * (3 feeds the raw fp32 operand and relies on the tensor cores not reading the low 13 mantissa bits).  All produce the same top-K survivors up to fp32-level rounding of the
 * approximate distances; the final (index, distance) is re-scored exactly. */
int dslam_sc_set_scan_kernel(int flavour);

#ifdef __cplusplus
}
#endif
#endif /* DSLAM_B200_H_ */
