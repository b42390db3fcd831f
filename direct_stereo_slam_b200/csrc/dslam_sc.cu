// dslam_sc.cu — Scan-Context descriptor database behind the C ABI (dslam_sc_* of include/dslam_b200.h).
//
// Replaces search_ringkey / search_sc (src/loop_closure/loop_detection/search_place.h:25-57, 59-85; call sites
// src/loop_closure/LoopHandler.cpp:247, 256).  The database is a device-resident dense fp32 table (4,800 B per
// descriptor + 80 B ring key); a query batch is answered by ONE streaming pass over the local shard
// (sc_scan_kernel), an exact re-score of the K survivors in the reference's arithmetic on the device
// (sc_rescore_topk_kernel) and — when the database is sharded over several GPUs — one NCCL all-reduce(min) of
// the packed (distance, id) keys over NVLink.  Rows are appended in id order, so "lowest id wins ties" is the
// same rule on every shard and across shards.
//
// NCCL is loaded with dlopen at dslam_sc_comm_init so that the library has no link-time dependency on it (the
// process usually already holds torch's bundled libnccl.so.2, which dlopen then returns).

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <new>
#include <unordered_map>

#include "dslam_internal.h"

namespace {

typedef unsigned long long u64;
constexpr u64 kKeyMax = ~0ull;

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi *nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.handle ? &api : nullptr;
  tried = true;
  const char *env = getenv("DSLAM_NCCL_LIB");
  const char *names[] = {env, "libnccl.so.2", "libnccl.so"};
  void *h = nullptr;
  for (const char *n : names) {
    if (!n || !*n) continue;
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) return nullptr;
  api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
  api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
  api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
  api.AllGather = (decltype(api.AllGather))dlsym(h, "ncclAllGather");
  api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
  if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce || !api.AllGather || !api.GetErrorString) {
    dlclose(h);
    return nullptr;
  }
  api.handle = h;
  return &api;
}

inline float key_dist(u64 key) {
  unsigned b = (unsigned)(key >> 32);
  b ^= (b >> 31) ? 0x80000000u : 0xffffffffu;
  float f;
  std::memcpy(&f, &b, 4);
  return f;
}
inline int key_id(u64 key) { return (int)(unsigned)(key & 0xffffffffull); }

}  // namespace

struct dslam_scdb {
  dslam_session *s = nullptr;
  int n_sectors = 0, n_rings = 0, n_cells = 0, capacity = 0, n = 0;
  float *d_sigs = nullptr, *d_keys = nullptr;
  int *d_ids = nullptr;
  std::vector<int> ids;                  // global id of every local row (ascending)
  std::unordered_map<int, int> row_of;   // global id -> local row
  // query-side buffers (grown on demand)
  int qcap = 0;
  float *d_qsigs = nullptr, *d_qkeys = nullptr;
  u64 *d_topk = nullptr, *d_exact = nullptr, *d_best = nullptr, *d_gather = nullptr, *d_scratch = nullptr;
  size_t scratch_bytes = 0;
  u64 *h_keys = nullptr;  // pinned: max(qcap * K * world, ...)
  size_t h_keys_cap = 0;
  int paircap = 0;
  int *d_pair_q = nullptr, *d_pair_row = nullptr;
  float *d_pair_diff = nullptr;
  // descriptor generation scratch
  double *d_pts = nullptr, *d_mom = nullptr, *d_gen_sig64 = nullptr;
  int pts_cap = 0;
  u64 *d_cells = nullptr;
  float *d_gen_sig = nullptr, *d_gen_key = nullptr;
  // staging for appends
  float *h_stage = nullptr;  // pinned
  size_t h_stage_floats = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool have_scan_time = false;
  ncclComm_t comm = nullptr;
  int world = 1, rank = 0;
};

using namespace dslam;

namespace {

int ensure_query_buffers(dslam_scdb *db, int nq) {
  const int world = db->world;
  if (nq > db->qcap) {
    cudaFree(db->d_qsigs); cudaFree(db->d_qkeys); cudaFree(db->d_topk); cudaFree(db->d_exact); cudaFree(db->d_best); cudaFree(db->d_gather);
    db->d_qsigs = db->d_qkeys = nullptr;
    db->d_topk = db->d_exact = db->d_best = db->d_gather = nullptr;
    const int cap = std::max(32, nq);
    DSLAM_CUDA(cudaMalloc((void **)&db->d_qsigs, (size_t)cap * db->n_cells * sizeof(float)));
    DSLAM_CUDA(cudaMalloc((void **)&db->d_qkeys, (size_t)cap * db->n_rings * sizeof(float)));
    DSLAM_CUDA(cudaMalloc((void **)&db->d_topk, (size_t)cap * kScTopK * sizeof(u64)));
    DSLAM_CUDA(cudaMalloc((void **)&db->d_exact, (size_t)cap * kScTopK * sizeof(u64)));
    DSLAM_CUDA(cudaMalloc((void **)&db->d_best, (size_t)cap * sizeof(u64)));
    DSLAM_CUDA(cudaMalloc((void **)&db->d_gather, (size_t)cap * kScTopK * sizeof(u64) * world));
    db->qcap = cap;
  }
  const size_t sb = sc_scratch_bytes(nq);
  if (sb > db->scratch_bytes) {
    cudaFree(db->d_scratch);
    db->d_scratch = nullptr;
    DSLAM_CUDA(cudaMalloc((void **)&db->d_scratch, sb));
    db->scratch_bytes = sb;
  }
  const size_t hk = (size_t)std::max(32, nq) * kScTopK * world;
  if (hk > db->h_keys_cap) {
    if (db->h_keys) cudaFreeHost(db->h_keys);
    db->h_keys = nullptr;
    DSLAM_CUDA(cudaHostAlloc((void **)&db->h_keys, hk * sizeof(u64), cudaHostAllocDefault));
    db->h_keys_cap = hk;
  }
  return DSLAM_OK;
}

int ensure_stage(dslam_scdb *db, size_t floats) {
  if (floats > db->h_stage_floats) {
    if (db->h_stage) cudaFreeHost(db->h_stage);
    db->h_stage = nullptr;
    DSLAM_CUDA(cudaHostAlloc((void **)&db->h_stage, floats * sizeof(float), cudaHostAllocDefault));
    db->h_stage_floats = floats;
  }
  return DSLAM_OK;
}

int upload_queries(dslam_scdb *db, int nq, const float *ringkeys, const float *sigs) {
  const int rc = ensure_query_buffers(db, nq);
  if (rc != DSLAM_OK) return rc;
  if (sigs) DSLAM_CUDA(cudaMemcpyAsync(db->d_qsigs, sigs, (size_t)nq * db->n_cells * sizeof(float), cudaMemcpyHostToDevice, db->s->stream));
  if (ringkeys)
    DSLAM_CUDA(cudaMemcpyAsync(db->d_qkeys, ringkeys, (size_t)nq * db->n_rings * sizeof(float), cudaMemcpyHostToDevice, db->s->stream));
  return DSLAM_OK;
}

int nccl_fail(ncclResult_t r, const char *what) {
  NcclApi *api = nccl_api();
  return fail(DSLAM_ENCCL, "NCCL error %d (%s) in %s", (int)r, api ? api->GetErrorString(r) : "?", what);
}

// scan + exact re-score of the local shard: d_best[q] = packed (exact dist, global id) or ~0
int local_query(dslam_scdb *db, int nq, const float *ringkeys, const float *sigs, float ringkey_thres, int max_id) {
  if (ringkey_thres >= 0.f && !ringkeys) return fail(DSLAM_EINVAL, "ring-key gate requested without query ring keys");
  int rc = upload_queries(db, nq, ringkeys, sigs);
  if (rc != DSLAM_OK) return rc;
  dslam_session *s = db->s;
  DSLAM_CUDA(cudaEventRecord(db->ev0, s->stream));
  DSLAM_CUDA(launch_sc_scan(db->d_sigs, db->d_keys, db->d_ids, db->n, db->n_cells, db->n_rings, db->d_qsigs, db->d_qkeys, nq, ringkey_thres,
                            max_id, (float)db->n_sectors, db->d_topk, db->d_scratch, s->stream));
  DSLAM_CUDA(cudaEventRecord(db->ev1, s->stream));
  db->have_scan_time = true;
  s->launches += 2 * ((nq + 31) / 32);
  DSLAM_CUDA(launch_sc_rescore_topk(db->d_topk, db->d_sigs, db->d_ids, db->d_qsigs, nq, db->n_cells, db->n_sectors, db->d_exact, db->d_best,
                                    s->stream));
  s->launches++;
  return DSLAM_OK;
}

}  // namespace

extern "C" {

int dslam_sc_create(dslam_session *s, int n_sectors, int n_rings, int capacity, dslam_scdb **out) {
  if (!s || !out) return fail(DSLAM_EINVAL, "null argument");
  *out = nullptr;
  if (n_sectors < 1 || n_rings < 1 || capacity < 1) return fail(DSLAM_EINVAL, "bad database geometry");
  const int n_cells = n_sectors * n_rings;
  if (n_cells % 4 != 0 || n_cells > 1280 || n_rings > 64)
    return fail(DSLAM_EINVAL, "unsupported descriptor shape %dx%d (cells must be a multiple of 4 and <= 1280, rings <= 64)", n_sectors, n_rings);
  DSLAM_CUDA(cudaSetDevice(s->device));
  dslam_scdb *db = new (std::nothrow) dslam_scdb();
  if (!db) return fail(DSLAM_ENOMEM, "out of host memory");
  db->s = s; db->n_sectors = n_sectors; db->n_rings = n_rings; db->n_cells = n_cells; db->capacity = capacity;
  cudaError_t e = cudaMalloc((void **)&db->d_sigs, (size_t)capacity * n_cells * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void **)&db->d_keys, (size_t)capacity * n_rings * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void **)&db->d_ids, (size_t)capacity * sizeof(int));
  if (e == cudaSuccess) e = cudaEventCreate(&db->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&db->ev1);
  if (e != cudaSuccess) {
    cudaFree(db->d_sigs); cudaFree(db->d_keys); cudaFree(db->d_ids);
    delete db;
    return cuda_fail(e, "dslam_sc_create");
  }
  *out = db;
  return DSLAM_OK;
}

int dslam_sc_destroy(dslam_scdb *db) {
  if (!db) return DSLAM_OK;
  cudaSetDevice(db->s->device);
  cudaStreamSynchronize(db->s->stream);
  if (db->comm) {
    NcclApi *api = nccl_api();
    if (api) api->CommDestroy(db->comm);
  }
  cudaFree(db->d_sigs); cudaFree(db->d_keys); cudaFree(db->d_ids);
  cudaFree(db->d_qsigs); cudaFree(db->d_qkeys); cudaFree(db->d_topk); cudaFree(db->d_exact); cudaFree(db->d_best); cudaFree(db->d_gather);
  cudaFree(db->d_pts); cudaFree(db->d_mom); cudaFree(db->d_gen_sig64); cudaFree(db->d_cells); cudaFree(db->d_gen_sig); cudaFree(db->d_gen_key);
  cudaFree(db->d_scratch); cudaFree(db->d_pair_q); cudaFree(db->d_pair_row); cudaFree(db->d_pair_diff);
  if (db->h_keys) cudaFreeHost(db->h_keys);
  if (db->h_stage) cudaFreeHost(db->h_stage);
  cudaEventDestroy(db->ev0);
  cudaEventDestroy(db->ev1);
  delete db;
  return DSLAM_OK;
}

int dslam_sc_add(dslam_scdb *db, int n, const float *ringkeys, const float *sigs_dense, const int *global_ids) {
  if (!db || n < 0) return fail(DSLAM_EINVAL, "bad argument");
  if (n == 0) return DSLAM_OK;
  if (!ringkeys || !sigs_dense) return fail(DSLAM_EINVAL, "null descriptor array");
  if (db->n + n > db->capacity) return fail(DSLAM_ENOMEM, "database capacity %d exceeded (%d + %d)", db->capacity, db->n, n);
  int last = db->ids.empty() ? -1 : db->ids.back();
  for (int i = 0; i < n; i++) {
    const int id = global_ids ? global_ids[i] : db->n + i;
    if (id <= last) return fail(DSLAM_EINVAL, "global ids must be strictly ascending within a shard (%d after %d)", id, last);
    last = id;
  }
  dslam_session *s = db->s;
  DSLAM_CUDA(cudaMemcpyAsync(db->d_sigs + (size_t)db->n * db->n_cells, sigs_dense, (size_t)n * db->n_cells * sizeof(float), cudaMemcpyHostToDevice,
                             s->stream));
  DSLAM_CUDA(cudaMemcpyAsync(db->d_keys + (size_t)db->n * db->n_rings, ringkeys, (size_t)n * db->n_rings * sizeof(float), cudaMemcpyHostToDevice,
                             s->stream));
  const size_t base = db->ids.size();
  for (int i = 0; i < n; i++) {
    const int id = global_ids ? global_ids[i] : db->n + i;
    db->row_of[id] = db->n + i;
    db->ids.push_back(id);
  }
  DSLAM_CUDA(cudaMemcpyAsync(db->d_ids + db->n, db->ids.data() + base, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, s->stream));
  // ids vector may reallocate on a later append: make sure the copy has consumed it
  DSLAM_CUDA(cudaStreamSynchronize(s->stream));
  db->n += n;
  return DSLAM_OK;
}

int dslam_sc_add_sparse(dslam_scdb *db, const float *ringkey, const int *idx, const double *val, int nnz, int global_id) {
  if (!db || !ringkey || nnz < 0 || (nnz > 0 && (!idx || !val))) return fail(DSLAM_EINVAL, "bad argument");
  const int rc = ensure_stage(db, (size_t)db->n_cells);
  if (rc != DSLAM_OK) return rc;
  DSLAM_CUDA(cudaStreamSynchronize(db->s->stream));
  std::memset(db->h_stage, 0, sizeof(float) * db->n_cells);
  for (int i = 0; i < nnz; i++) {
    if (idx[i] < 0 || idx[i] >= db->n_cells) return fail(DSLAM_EINVAL, "cell index %d out of range", idx[i]);
    db->h_stage[idx[i]] = (float)val[i];  // the device format is fp32
  }
  const int id = global_id < 0 ? (db->ids.empty() ? 0 : db->ids.back() + 1) : global_id;
  return dslam_sc_add(db, 1, ringkey, db->h_stage, &id);
}

// ScanContext::generate (src/loop_closure/loop_detection/ScanContext.cpp:78-142 with align_points_PCA :19-66).
int dslam_sc_generate(dslam_scdb *db, const double *pts_xyz, int n, double lidar_range, float *ringkey_out, float *sig_dense_out,
                      double *sig_dense64_out, double tfm_pca_rig[16], int append, int global_id) {
  if (!db || !pts_xyz || n < 1 || !(lidar_range > 0)) return fail(DSLAM_EINVAL, "bad argument");
  if (append && db->n + 1 > db->capacity) return fail(DSLAM_ENOMEM, "database capacity %d exceeded", db->capacity);
  const int last = db->ids.empty() ? -1 : db->ids.back();
  const int id = global_id < 0 ? last + 1 : global_id;
  if (append && id <= last) return fail(DSLAM_EINVAL, "global ids must be strictly ascending within a shard (%d after %d)", id, last);
  dslam_session *s = db->s;
  DSLAM_CUDA(cudaSetDevice(s->device));
  if (n > db->pts_cap) {
    cudaFree(db->d_pts);
    db->d_pts = nullptr;
    const int cap = n + n / 2 + 1024;
    DSLAM_CUDA(cudaMalloc((void **)&db->d_pts, (size_t)cap * 3 * sizeof(double)));
    db->pts_cap = cap;
  }
  if (!db->d_mom) {
    DSLAM_CUDA(cudaMalloc((void **)&db->d_mom, 16 * sizeof(double)));
    DSLAM_CUDA(cudaMalloc((void **)&db->d_cells, (size_t)db->n_cells * sizeof(u64)));
    DSLAM_CUDA(cudaMalloc((void **)&db->d_gen_sig, (size_t)db->n_cells * sizeof(float)));
    DSLAM_CUDA(cudaMalloc((void **)&db->d_gen_sig64, (size_t)db->n_cells * sizeof(double)));
    DSLAM_CUDA(cudaMalloc((void **)&db->d_gen_key, (size_t)db->n_rings * sizeof(float)));
  }
  DSLAM_CUDA(cudaMemcpyAsync(db->d_pts, pts_xyz, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  DSLAM_CUDA(launch_sc_moments(db->d_pts, n, db->d_mom, s->stream));
  s->launches++;
  double mom[9];
  DSLAM_CUDA(cudaMemcpyAsync(mom, db->d_mom, sizeof(mom), cudaMemcpyDeviceToHost, s->stream));
  DSLAM_CUDA(cudaStreamSynchronize(s->stream));
  // the 3x3 eigen-decomposition stays on the host (like the 8x8 solve of the trackers)
  const double cov[9] = {mom[3], mom[4], mom[5], mom[4], mom[6], mom[7], mom[5], mom[7], mom[8]};
  double evals[3], ev[9], v9[9];
  hm::sym_eig3(cov, evals, ev);
  for (int j = 0; j < 3; j++)
    for (int c = 0; c < 3; c++) v9[3 * j + c] = ev[c * 3 + j];
  if (tfm_pca_rig) {  // :54-65
    for (int i = 0; i < 16; i++) tfm_pca_rig[i] = (i % 5 == 0) ? 1.0 : 0.0;
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) tfm_pca_rig[r * 4 + c] = v9[3 * r + c];
    for (int r = 0; r < 3; r++) tfm_pca_rig[r * 4 + 3] = -(tfm_pca_rig[r * 4] * mom[0] + tfm_pca_rig[r * 4 + 1] * mom[1] + tfm_pca_rig[r * 4 + 2] * mom[2]);
  }
  DSLAM_CUDA(launch_sc_bin_finalize(db->d_pts, n, mom, v9, lidar_range, db->n_sectors, db->n_rings, db->d_cells, db->d_gen_key, db->d_gen_sig,
                                    db->d_gen_sig64, s->stream));
  s->launches += 3;
  if (append) {  // the new descriptor goes into the database without touching the host
    DSLAM_CUDA(cudaMemcpyAsync(db->d_sigs + (size_t)db->n * db->n_cells, db->d_gen_sig, (size_t)db->n_cells * sizeof(float), cudaMemcpyDeviceToDevice, s->stream));
    DSLAM_CUDA(cudaMemcpyAsync(db->d_keys + (size_t)db->n * db->n_rings, db->d_gen_key, (size_t)db->n_rings * sizeof(float), cudaMemcpyDeviceToDevice, s->stream));
    DSLAM_CUDA(cudaMemcpyAsync(db->d_ids + db->n, &id, sizeof(int), cudaMemcpyHostToDevice, s->stream));
  }
  if (ringkey_out) DSLAM_CUDA(cudaMemcpyAsync(ringkey_out, db->d_gen_key, (size_t)db->n_rings * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
  if (sig_dense_out) DSLAM_CUDA(cudaMemcpyAsync(sig_dense_out, db->d_gen_sig, (size_t)db->n_cells * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
  if (sig_dense64_out) DSLAM_CUDA(cudaMemcpyAsync(sig_dense64_out, db->d_gen_sig64, (size_t)db->n_cells * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  DSLAM_CUDA(cudaStreamSynchronize(s->stream));
  if (append) {
    db->row_of[id] = db->n;
    db->ids.push_back(id);
    db->n += 1;
  }
  return DSLAM_OK;
}

int dslam_sc_size(dslam_scdb *db, int *n_local) {
  if (!db || !n_local) return fail(DSLAM_EINVAL, "null argument");
  *n_local = db->n;
  return DSLAM_OK;
}

int dslam_sc_search_ringkey(dslam_scdb *db, int nq, const float *ringkeys, int k, float thres, int max_id, int *cand_out, float *dist_out) {
  if (!db || nq < 1 || !ringkeys || !cand_out) return fail(DSLAM_EINVAL, "bad argument");
  if (k < 1 || k > kScTopK) return fail(DSLAM_EINVAL, "k must be in [1, %d]", kScTopK);
  int rc = upload_queries(db, nq, ringkeys, nullptr);
  if (rc != DSLAM_OK) return rc;
  dslam_session *s = db->s;
  DSLAM_CUDA(launch_sc_ringkey(db->d_keys, db->d_ids, db->n, db->n_rings, db->d_qkeys, nq, max_id, db->d_topk, db->d_scratch, s->stream));
  s->launches += 2;
  const u64 *src = db->d_topk;
  int lists = 1;
  if (db->comm) {
    NcclApi *api = nccl_api();
    const ncclResult_t r = api->AllGather(db->d_topk, db->d_gather, (size_t)nq * kScTopK, ncclUint64, db->comm, s->stream);
    if (r != ncclSuccess) return nccl_fail(r, "ncclAllGather");
    src = db->d_gather;
    lists = db->world;
  }
  DSLAM_CUDA(cudaMemcpyAsync(db->h_keys, src, (size_t)nq * kScTopK * lists * sizeof(u64), cudaMemcpyDeviceToHost, s->stream));
  DSLAM_CUDA(cudaStreamSynchronize(s->stream));
  std::vector<u64> merged;
  for (int q = 0; q < nq; q++) {
    merged.clear();
    for (int l = 0; l < lists; l++)
      for (int i = 0; i < kScTopK; i++) merged.push_back(db->h_keys[((size_t)l * nq + q) * kScTopK + i]);
    std::sort(merged.begin(), merged.end());
    int m = 0;
    for (int i = 0; i < k; i++) {
      cand_out[(size_t)q * k + i] = -1;
      if (dist_out) dist_out[(size_t)q * k + i] = INFINITY;
    }
    for (int i = 0; i < k && i < (int)merged.size(); i++) {
      if (merged[i] == kKeyMax) break;
      const float d = key_dist(merged[i]);
      if (!(d < thres)) continue;  // "dists[0][i] < RINGKEY_THRES"  search_place.h:35
      cand_out[(size_t)q * k + m] = key_id(merged[i]);
      if (dist_out) dist_out[(size_t)q * k + m] = d;
      m++;
    }
  }
  return DSLAM_OK;
}

int dslam_sc_search_sc(dslam_scdb *db, int nq, const float *sigs_dense, const int *candidates, int n_cand, int *res_idx, float *res_diff) {
  if (!db || nq < 1 || !sigs_dense || !candidates || n_cand < 1 || !res_idx || !res_diff) return fail(DSLAM_EINVAL, "bad argument");
  if (db->world > 1) return fail(DSLAM_ESTATE, "dslam_sc_search_sc needs the candidates' rows on this rank; use dslam_sc_query on a sharded database");
  int rc = upload_queries(db, nq, nullptr, sigs_dense);
  if (rc != DSLAM_OK) return rc;
  const int np = nq * n_cand;
  if (np > db->paircap) {
    cudaFree(db->d_pair_q); cudaFree(db->d_pair_row); cudaFree(db->d_pair_diff);
    db->d_pair_q = db->d_pair_row = nullptr;
    db->d_pair_diff = nullptr;
    DSLAM_CUDA(cudaMalloc((void **)&db->d_pair_q, (size_t)np * sizeof(int)));
    DSLAM_CUDA(cudaMalloc((void **)&db->d_pair_row, (size_t)np * sizeof(int)));
    DSLAM_CUDA(cudaMalloc((void **)&db->d_pair_diff, (size_t)np * sizeof(float)));
    db->paircap = np;
  }
  std::vector<int> pq((size_t)np), prow((size_t)np);
  for (int q = 0; q < nq; q++)
    for (int j = 0; j < n_cand; j++) {
      const int cand = candidates[(size_t)q * n_cand + j];
      int row = -1;
      if (cand >= 0) {
        auto itr = db->row_of.find(cand);
        if (itr == db->row_of.end()) return fail(DSLAM_EINVAL, "candidate id %d is not in the database", cand);
        row = itr->second;
      }
      pq[(size_t)q * n_cand + j] = q;
      prow[(size_t)q * n_cand + j] = row;
    }
  dslam_session *s = db->s;
  DSLAM_CUDA(cudaMemcpyAsync(db->d_pair_q, pq.data(), (size_t)np * sizeof(int), cudaMemcpyHostToDevice, s->stream));
  DSLAM_CUDA(cudaMemcpyAsync(db->d_pair_row, prow.data(), (size_t)np * sizeof(int), cudaMemcpyHostToDevice, s->stream));
  DSLAM_CUDA(launch_sc_rescore_pairs(db->d_pair_q, db->d_pair_row, np, db->d_sigs, db->d_qsigs, db->n_cells, db->n_sectors, db->d_pair_diff,
                                     s->stream));
  s->launches++;
  std::vector<float> diff((size_t)np);
  DSLAM_CUDA(cudaMemcpyAsync(diff.data(), db->d_pair_diff, (size_t)np * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
  DSLAM_CUDA(cudaStreamSynchronize(s->stream));
  for (int q = 0; q < nq; q++) {
    // search_place.h:63-64, 80-83: running minimum in candidate order with a strict '>' from 1.1
    int best = candidates[(size_t)q * n_cand];
    float bd = 1.1f;
    for (int j = 0; j < n_cand; j++) {
      if (candidates[(size_t)q * n_cand + j] < 0) continue;
      const float d = diff[(size_t)q * n_cand + j];
      if (bd > d) {
        best = candidates[(size_t)q * n_cand + j];
        bd = d;
      }
    }
    res_idx[q] = best;
    res_diff[q] = bd;
  }
  return DSLAM_OK;
}

int dslam_sc_query_keys(dslam_scdb *db, int nq, const float *ringkeys, const float *sigs_dense, float ringkey_thres, int max_id,
                        unsigned long long *keys_out) {
  if (!db || nq < 1 || !sigs_dense || !keys_out) return fail(DSLAM_EINVAL, "bad argument");
  const int rc = local_query(db, nq, ringkeys, sigs_dense, ringkey_thres, max_id);
  if (rc != DSLAM_OK) return rc;
  DSLAM_CUDA(cudaMemcpyAsync(db->h_keys, db->d_best, (size_t)nq * sizeof(u64), cudaMemcpyDeviceToHost, db->s->stream));
  DSLAM_CUDA(cudaStreamSynchronize(db->s->stream));
  std::memcpy(keys_out, db->h_keys, (size_t)nq * sizeof(u64));
  return DSLAM_OK;
}

int dslam_sc_decode_key(unsigned long long key, int *id, float *dist) {
  if (key == kKeyMax) {
    if (id) *id = -1;
    if (dist) *dist = 1.1f;
    return DSLAM_OK;
  }
  if (id) *id = key_id(key);
  if (dist) *dist = key_dist(key);
  return DSLAM_OK;
}

int dslam_sc_query(dslam_scdb *db, int nq, const float *ringkeys, const float *sigs_dense, float ringkey_thres, int max_id, int *res_idx,
                   float *res_diff) {
  if (!db || nq < 1 || !sigs_dense || !res_idx) return fail(DSLAM_EINVAL, "bad argument");
  const int rc = local_query(db, nq, ringkeys, sigs_dense, ringkey_thres, max_id);
  if (rc != DSLAM_OK) return rc;
  dslam_session *s = db->s;
  if (db->comm) {
    NcclApi *api = nccl_api();
    const ncclResult_t r = api->AllReduce(db->d_best, db->d_best, (size_t)nq, ncclUint64, ncclMin, db->comm, s->stream);
    if (r != ncclSuccess) return nccl_fail(r, "ncclAllReduce");
  }
  DSLAM_CUDA(cudaMemcpyAsync(db->h_keys, db->d_best, (size_t)nq * sizeof(u64), cudaMemcpyDeviceToHost, s->stream));
  DSLAM_CUDA(cudaStreamSynchronize(s->stream));
  for (int q = 0; q < nq; q++) {
    int id;
    float d;
    dslam_sc_decode_key(db->h_keys[q], &id, &d);
    res_idx[q] = id;
    if (res_diff) res_diff[q] = d;
  }
  return DSLAM_OK;
}

int dslam_sc_unique_id(unsigned char id128[128]) {
  if (!id128) return fail(DSLAM_EINVAL, "null argument");
  NcclApi *api = nccl_api();
  if (!api) return fail(DSLAM_ENCCL, "libnccl.so.2 could not be loaded (%s)", dlerror() ? dlerror() : "not found");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId id;
  const ncclResult_t r = api->GetUniqueId(&id);
  if (r != ncclSuccess) return nccl_fail(r, "ncclGetUniqueId");
  std::memcpy(id128, &id, 128);
  return DSLAM_OK;
}

int dslam_sc_comm_init(dslam_scdb *db, const unsigned char id128[128], int world_size, int rank) {
  if (!db || !id128 || world_size < 1 || rank < 0 || rank >= world_size) return fail(DSLAM_EINVAL, "bad argument");
  if (db->comm) return fail(DSLAM_ESTATE, "communicator already attached");
  NcclApi *api = nccl_api();
  if (!api) return fail(DSLAM_ENCCL, "libnccl.so.2 could not be loaded");
  DSLAM_CUDA(cudaSetDevice(db->s->device));
  ncclUniqueId id;
  std::memcpy(&id, id128, 128);
  ncclComm_t comm = nullptr;
  const ncclResult_t r = api->CommInitRank(&comm, world_size, id, rank);
  if (r != ncclSuccess) return nccl_fail(r, "ncclCommInitRank");
  db->comm = comm;
  db->world = world_size;
  db->rank = rank;
  db->qcap = 0;  // gather buffers depend on the world size
  return DSLAM_OK;
}

int dslam_sc_last_scan_ms(dslam_scdb *db, float *ms) {
  if (!db || !ms) return fail(DSLAM_EINVAL, "null argument");
  if (!db->have_scan_time) return fail(DSLAM_ESTATE, "no scan has run yet");
  DSLAM_CUDA(cudaEventSynchronize(db->ev1));
  DSLAM_CUDA(cudaEventElapsedTime(ms, db->ev0, db->ev1));
  return DSLAM_OK;
}

int dslam_sc_set_scan_kernel(int flavour) {
  if (flavour < 0 || flavour > 2) return fail(DSLAM_EINVAL, "scan kernel flavour must be 0 (auto), 1 (stream) or 2 (tile)");
  sc_set_scan_flavour(flavour);
  return DSLAM_OK;
}

}  // extern "C"
