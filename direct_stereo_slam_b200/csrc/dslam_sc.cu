// dslam_sc.cu — Scan-Context descriptor database behind the C ABI (dslam_sc_* of include/dslam_b200.h).
//
// Replaces search_ringkey / search_sc (src/loop_closure/loop_detection/search_place.h:25-57, 59-85; call sites
// src/loop_closure/LoopHandler.cpp:247, 256).  The database is a device-resident dense fp32 table (4,800 B per
// descriptor + 80 B ring key; optionally an fp64 side table with the reference's double values for the exact
// re-score); a query batch is answered by ONE pass over the local shard (sc_scan_kernel / sc_scan_tile_kernel, the
// merge of the per-CTA lists fused into the last CTA), an exact re-score of the K survivors in the reference's
// arithmetic (sc_rescore_topk_kernel) which also PUBLISHES the result: straight into mapped pinned host memory as
// self-validating words on one GPU, and — when the database is sharded over several GPUs — after an NVLink mailbox
// exchange of the packed (distance, id) keys (system-scope atomic-min into every peer's HBM through CUDA IPC mappings;
// no library collective on the query path; NCCL bootstraps the mappings and remains as the fallback exchange).
// Rows are appended in id order, so "lowest id wins ties" is the same rule on every shard and across shards.
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
  bool fp64 = false;                     // keep the reference's double signature values for the exact re-score
  float *d_sigs = nullptr, *d_keys = nullptr;
  double *d_sigs64 = nullptr;
  int *d_ids = nullptr;
  std::vector<int> ids;                  // global id of every local row (ascending)
  std::unordered_map<int, int> row_of;   // global id -> local row
  // query-side buffers (grown on demand)
  int qcap = 0;
  float *d_qsigs = nullptr, *d_qkeys = nullptr, *d_qsplit = nullptr;  // d_qsplit: hi / lo halves of the batch for the tensor-core scan
  double *d_qsigs64 = nullptr;
  u64 *d_topk = nullptr, *d_exact = nullptr, *d_best = nullptr, *d_gather = nullptr, *d_scratch = nullptr;
  size_t scratch_bytes = 0;
  u64 *h_keys = nullptr;  // pinned: max(qcap * K * world, ...)
  size_t h_keys_cap = 0;
  u64 *h_words = nullptr, *d_words = nullptr;  // mapped pinned result words (2 per query) the re-score kernel publishes into
  unsigned seq = 0;                            // sequence number carried by those words
  float *h_qstage = nullptr;                   // pinned staging of small query batches (the caller's arrays are pageable)
  size_t h_qstage_bytes = 0;
  unsigned *d_ticket = nullptr;                // ticket of the re-score grid
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
  // NVLink mailbox exchange (kernels_sc.cu): own mailbox + IPC mappings of the peers'
  void *d_mail = nullptr;
  void *peer_mail[8] = {};
  dslam::ScExchange xchg;
  bool p2p = false;
  unsigned xseq = 0;  // exchange sequence number: advances identically on every rank (collective calls)
};

using namespace dslam;

namespace {

int ensure_query_buffers(dslam_scdb *db, int nq) {
  const int world = db->world;
  if (nq > db->qcap) {
    cudaFree(db->d_qsigs); cudaFree(db->d_qkeys); cudaFree(db->d_qsigs64); cudaFree(db->d_topk); cudaFree(db->d_exact); cudaFree(db->d_best);
    cudaFree(db->d_gather); cudaFree(db->d_qsplit);
    if (db->h_words) cudaFreeHost(db->h_words);
    db->d_qsigs = db->d_qkeys = db->d_qsplit = nullptr;
    db->d_qsigs64 = nullptr;
    db->d_topk = db->d_exact = db->d_best = db->d_gather = nullptr;
    db->h_words = db->d_words = nullptr;
    db->qcap = 0;
    const int cap = std::max(32, nq);
    DSLAM_CUDA(cudaMalloc((void **)&db->d_qsigs, (size_t)cap * db->n_cells * sizeof(float)));
    DSLAM_CUDA(cudaMalloc((void **)&db->d_qkeys, (size_t)cap * db->n_rings * sizeof(float)));
    DSLAM_CUDA(cudaMalloc((void **)&db->d_qsplit, dslam::sc_qsplit_floats(cap, db->n_cells) * sizeof(float)));
    if (db->fp64) DSLAM_CUDA(cudaMalloc((void **)&db->d_qsigs64, (size_t)cap * db->n_cells * sizeof(double)));
    DSLAM_CUDA(cudaMalloc((void **)&db->d_topk, (size_t)cap * kScTopK * sizeof(u64)));
    DSLAM_CUDA(cudaMalloc((void **)&db->d_exact, (size_t)cap * kScTopK * sizeof(u64)));
    DSLAM_CUDA(cudaMalloc((void **)&db->d_best, (size_t)cap * sizeof(u64)));
    DSLAM_CUDA(cudaMalloc((void **)&db->d_gather, (size_t)cap * kScTopK * sizeof(u64) * world));
    DSLAM_CUDA(cudaHostAlloc((void **)&db->h_words, (size_t)cap * 2 * sizeof(u64), cudaHostAllocMapped));
    std::memset(db->h_words, 0, (size_t)cap * 2 * sizeof(u64));
    DSLAM_CUDA(cudaHostGetDevicePointer((void **)&db->d_words, db->h_words, 0));
    db->qcap = cap;
  }
  if (!db->d_ticket) {
    DSLAM_CUDA(cudaMalloc((void **)&db->d_ticket, 256));
    DSLAM_CUDA(cudaMemsetAsync(db->d_ticket, 0, 256, db->s->stream));
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

// Room for `need` rows: the device tables grow geometrically (new allocation + device-to-device copy on the session
// stream) — the reference's FLANN index and loop_frames_ grow without bound (src/loop_closure/LoopHandler.cpp:225-229).
int ensure_capacity(dslam_scdb *db, int need) {
  if (need <= db->capacity) return DSLAM_OK;
  long cap = db->capacity;
  while (cap < need) cap = cap + cap / 2 + 64;
  if (cap > 0x7fffff00L) return fail(DSLAM_ENOMEM, "database too large");
  dslam_session *s = db->s;
  float *sigs = nullptr, *keys = nullptr;
  double *sigs64 = nullptr;
  int *ids = nullptr;
  cudaError_t e = cudaMalloc((void **)&sigs, (size_t)cap * db->n_cells * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void **)&keys, (size_t)cap * db->n_rings * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void **)&ids, (size_t)cap * sizeof(int));
  if (e == cudaSuccess && db->fp64) e = cudaMalloc((void **)&sigs64, (size_t)cap * db->n_cells * sizeof(double));
  if (e == cudaSuccess && db->n > 0) {
    e = cudaMemcpyAsync(sigs, db->d_sigs, (size_t)db->n * db->n_cells * sizeof(float), cudaMemcpyDeviceToDevice, s->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(keys, db->d_keys, (size_t)db->n * db->n_rings * sizeof(float), cudaMemcpyDeviceToDevice, s->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ids, db->d_ids, (size_t)db->n * sizeof(int), cudaMemcpyDeviceToDevice, s->stream);
    if (e == cudaSuccess && db->fp64)
      e = cudaMemcpyAsync(sigs64, db->d_sigs64, (size_t)db->n * db->n_cells * sizeof(double), cudaMemcpyDeviceToDevice, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
  }
  if (e != cudaSuccess) {
    cudaFree(sigs); cudaFree(keys); cudaFree(ids); cudaFree(sigs64);
    return cuda_fail(e, "growing the Scan-Context tables");
  }
  cudaFree(db->d_sigs); cudaFree(db->d_keys); cudaFree(db->d_ids); cudaFree(db->d_sigs64);
  db->d_sigs = sigs; db->d_keys = keys; db->d_ids = ids; db->d_sigs64 = sigs64;
  db->capacity = (int)cap;
  return DSLAM_OK;
}

// Query descriptors onto the device.  Small batches go through a pinned staging buffer (the caller's arrays are pageable:
// a pageable cudaMemcpyAsync is a synchronous staged copy inside the driver); fp64 databases keep both precisions of the
// queries: fp32 for the scan, fp64 for the exact re-score.
int upload_queries(dslam_scdb *db, int nq, const float *ringkeys, const float *sigs, const double *sigs64 = nullptr) {
  const int rc = ensure_query_buffers(db, nq);
  if (rc != DSLAM_OK) return rc;
  cudaStream_t st = db->s->stream;
  const size_t sig_b = (sigs || sigs64) ? (size_t)nq * db->n_cells * (sigs64 ? sizeof(double) : sizeof(float)) : 0;
  const size_t key_b = ringkeys ? (size_t)nq * db->n_rings * sizeof(float) : 0;
  const void *sig_src = sigs64 ? (const void *)sigs64 : (const void *)sigs;
  const void *key_src = ringkeys;
  if (sig_b + key_b > 0 && sig_b + key_b <= (1u << 20)) {
    if (sig_b + key_b > db->h_qstage_bytes) {
      DSLAM_CUDA(cudaStreamSynchronize(st));
      if (db->h_qstage) cudaFreeHost(db->h_qstage);
      db->h_qstage = nullptr;
      const size_t cap = std::max<size_t>(sig_b + key_b, 256u << 10);
      DSLAM_CUDA(cudaHostAlloc((void **)&db->h_qstage, cap, cudaHostAllocDefault));
      db->h_qstage_bytes = cap;
    }
    // the previous query's copy out of the staging buffer completed before its result was read (every query waits)
    unsigned char *st8 = reinterpret_cast<unsigned char *>(db->h_qstage);
    if (sig_src) { std::memcpy(st8, sig_src, sig_b); sig_src = st8; }
    if (key_src) { std::memcpy(st8 + sig_b, key_src, key_b); key_src = st8 + sig_b; }
  }
  if (sigs64) {
    if (!db->fp64) return fail(DSLAM_ESTATE, "fp64 queries need a database created with DSLAM_SC_FP64");
    DSLAM_CUDA(cudaMemcpyAsync(db->d_qsigs64, sig_src, sig_b, cudaMemcpyHostToDevice, st));
    DSLAM_CUDA(launch_sc_narrow(db->d_qsigs64, db->d_qsigs, (size_t)nq * db->n_cells, st));
    db->s->launches++;
  } else if (sigs) {
    DSLAM_CUDA(cudaMemcpyAsync(db->d_qsigs, sig_src, sig_b, cudaMemcpyHostToDevice, st));
    if (db->fp64) {
      DSLAM_CUDA(launch_sc_widen(db->d_qsigs, db->d_qsigs64, (size_t)nq * db->n_cells, st));
      db->s->launches++;
    }
  }
  if (ringkeys) DSLAM_CUDA(cudaMemcpyAsync(db->d_qkeys, key_src, key_b, cudaMemcpyHostToDevice, st));
  return DSLAM_OK;
}

int nccl_fail(ncclResult_t r, const char *what) {
  NcclApi *api = nccl_api();
  return fail(DSLAM_ENCCL, "NCCL error %d (%s) in %s", (int)r, api ? api->GetErrorString(r) : "?", what);
}

// how the re-score kernel hands its result on
enum Publish { kToDevice = 0 /* d_best only */, kToHost = 1 /* local keys -> host words */, kExchange = 2 /* NVLink mailboxes -> host words */ };

// scan + exact re-score of the local shard: d_best[q] = packed (exact dist, global id) or ~0; published as requested
int local_query(dslam_scdb *db, int nq, const float *ringkeys, const float *sigs, const double *sigs64, float ringkey_thres, int max_id, Publish pub,
                unsigned *seq_out) {
  if (ringkey_thres >= 0.f && !ringkeys) return fail(DSLAM_EINVAL, "ring-key gate requested without query ring keys");
  if (pub == kExchange && nq > kScXchgMaxQ) return fail(DSLAM_EINVAL, "at most %d queries per sharded batch", kScXchgMaxQ);
  int rc = upload_queries(db, nq, ringkeys, sigs, sigs64);
  if (rc != DSLAM_OK) return rc;
  dslam_session *s = db->s;
  DSLAM_CUDA(cudaEventRecord(db->ev0, s->stream));
  int nlists = 0;
  DSLAM_CUDA(launch_sc_scan(db->d_sigs, db->d_keys, db->d_ids, db->n, db->n_cells, db->n_rings, db->d_qsigs, db->d_qkeys, nq, ringkey_thres,
                            max_id, (float)db->n_sectors, db->d_scratch, &nlists, db->d_qsplit, s->stream));
  DSLAM_CUDA(cudaEventRecord(db->ev1, s->stream));
  db->have_scan_time = true;
  s->launches += (nq + 31) / 32;
  if (++db->seq == 0) ++db->seq;  // never 0: the words start out as 0
  const unsigned seq = db->seq;
  if (seq_out) *seq_out = seq;
  const void *tab = db->fp64 ? (const void *)db->d_sigs64 : (const void *)db->d_sigs;
  const void *qtab = db->fp64 ? (const void *)db->d_qsigs64 : (const void *)db->d_qsigs;
  DSLAM_CUDA(launch_sc_rescore_topk(db->d_scratch, nlists, tab, db->fp64 ? 1 : 0, db->d_ids, qtab, nq, db->n_cells, db->n_sectors, db->d_exact, db->d_best,
                                    pub == kToDevice ? nullptr : db->d_words, seq, pub == kExchange ? &db->xchg : nullptr,
                                    pub == kExchange ? db->xseq++ : 0u, 0, db->d_ticket, s->stream));
  s->launches++;
  return DSLAM_OK;
}

// wait for the 2 * nq self-validating words of query `seq` (mapped pinned memory the re-score kernel writes)
int wait_words(dslam_scdb *db, int nq, unsigned seq, u64 *keys_out) {
  const volatile u64 *w = db->h_words;
  const auto t0 = std::chrono::steady_clock::now();
  unsigned long spins = 0;
  for (int k = 0; k < 2 * nq; k++) {
    while ((unsigned)(w[k] & 0xffffffffull) != seq) {
      if ((++spins & 0x3fff) == 0) {
        const cudaError_t q = cudaStreamQuery(db->s->stream);
        if (q != cudaSuccess && q != cudaErrorNotReady) return cuda_fail(q, "Scan-Context query");
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > db->s->timeout_s)
          return fail(DSLAM_ETIMEOUT, "Scan-Context query did not publish its result within %.1f s", db->s->timeout_s);
      }
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    }
  }
  for (int q = 0; q < nq; q++) {
    keys_out[q] = (w[2 * q] & 0xffffffff00000000ull) | (w[2 * q + 1] >> 32);
    if (keys_out[q] == kScXchgErrorKey) return fail(DSLAM_ETIMEOUT, "a peer rank did not join the Scan-Context query exchange in time");
  }
  return DSLAM_OK;
}

}  // namespace

extern "C" {

int dslam_sc_create_ex(dslam_session *s, int n_sectors, int n_rings, int capacity, int flags, dslam_scdb **out) {
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
  db->fp64 = (flags & DSLAM_SC_FP64) != 0;
  cudaError_t e = cudaMalloc((void **)&db->d_sigs, (size_t)capacity * n_cells * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void **)&db->d_keys, (size_t)capacity * n_rings * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void **)&db->d_ids, (size_t)capacity * sizeof(int));
  if (e == cudaSuccess && db->fp64) e = cudaMalloc((void **)&db->d_sigs64, (size_t)capacity * n_cells * sizeof(double));
  if (e == cudaSuccess) e = cudaEventCreate(&db->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&db->ev1);
  if (e != cudaSuccess) {
    cudaFree(db->d_sigs); cudaFree(db->d_keys); cudaFree(db->d_ids); cudaFree(db->d_sigs64);
    delete db;
    return cuda_fail(e, "dslam_sc_create");
  }
  *out = db;
  return DSLAM_OK;
}

int dslam_sc_create(dslam_session *s, int n_sectors, int n_rings, int capacity, dslam_scdb **out) {
  return dslam_sc_create_ex(s, n_sectors, n_rings, capacity, 0, out);
}

int dslam_sc_destroy(dslam_scdb *db) {
  if (!db) return DSLAM_OK;
  cudaSetDevice(db->s->device);
  cudaStreamSynchronize(db->s->stream);
  for (int r = 0; r < 8; r++)
    if (db->peer_mail[r]) cudaIpcCloseMemHandle(db->peer_mail[r]);
  if (db->comm) {
    NcclApi *api = nccl_api();
    if (api) api->CommDestroy(db->comm);
  }
  cudaFree(db->d_mail);
  cudaFree(db->d_sigs); cudaFree(db->d_keys); cudaFree(db->d_ids); cudaFree(db->d_sigs64);
  cudaFree(db->d_qsigs); cudaFree(db->d_qkeys); cudaFree(db->d_qsigs64); cudaFree(db->d_topk); cudaFree(db->d_exact); cudaFree(db->d_best);
  cudaFree(db->d_gather); cudaFree(db->d_ticket); cudaFree(db->d_qsplit);
  cudaFree(db->d_pts); cudaFree(db->d_mom); cudaFree(db->d_gen_sig64); cudaFree(db->d_cells); cudaFree(db->d_gen_sig); cudaFree(db->d_gen_key);
  cudaFree(db->d_scratch); cudaFree(db->d_pair_q); cudaFree(db->d_pair_row); cudaFree(db->d_pair_diff);
  if (db->h_keys) cudaFreeHost(db->h_keys);
  if (db->h_words) cudaFreeHost(db->h_words);
  if (db->h_qstage) cudaFreeHost(db->h_qstage);
  if (db->h_stage) cudaFreeHost(db->h_stage);
  cudaEventDestroy(db->ev0);
  cudaEventDestroy(db->ev1);
  delete db;
  return DSLAM_OK;
}

// common tail of the append entry points: ids, bookkeeping
static int sc_append_ids(dslam_scdb *db, int n, const int *global_ids) {
  dslam_session *s = db->s;
  const size_t base = db->ids.size();
  for (int i = 0; i < n; i++) {
    const int id = global_ids ? global_ids[i] : db->n + i;
    db->row_of[id] = db->n + i;
    db->ids.push_back(id);
  }
  DSLAM_CUDA(cudaMemcpyAsync(db->d_ids + db->n, db->ids.data() + base, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, s->stream));
  // the caller's arrays (and the ids vector, which may reallocate on a later append) must have been consumed on return
  DSLAM_CUDA(cudaStreamSynchronize(s->stream));
  db->n += n;
  return DSLAM_OK;
}
static int sc_check_append(dslam_scdb *db, int n, const int *global_ids) {
  int last = db->ids.empty() ? -1 : db->ids.back();
  for (int i = 0; i < n; i++) {
    const int id = global_ids ? global_ids[i] : db->n + i;
    if (id <= last) return fail(DSLAM_EINVAL, "global ids must be strictly ascending within a shard (%d after %d)", id, last);
    last = id;
  }
  return ensure_capacity(db, db->n + n);
}

int dslam_sc_add(dslam_scdb *db, int n, const float *ringkeys, const float *sigs_dense, const int *global_ids) {
  if (!db || n < 0) return fail(DSLAM_EINVAL, "bad argument");
  if (n == 0) return DSLAM_OK;
  if (!ringkeys || !sigs_dense) return fail(DSLAM_EINVAL, "null descriptor array");
  DSLAM_CUDA(cudaSetDevice(db->s->device));
  const int rc = sc_check_append(db, n, global_ids);
  if (rc != DSLAM_OK) return rc;
  dslam_session *s = db->s;
  DSLAM_CUDA(cudaMemcpyAsync(db->d_sigs + (size_t)db->n * db->n_cells, sigs_dense, (size_t)n * db->n_cells * sizeof(float), cudaMemcpyHostToDevice,
                             s->stream));
  if (db->fp64)
    DSLAM_CUDA(launch_sc_widen(db->d_sigs + (size_t)db->n * db->n_cells, db->d_sigs64 + (size_t)db->n * db->n_cells, (size_t)n * db->n_cells, s->stream));
  DSLAM_CUDA(cudaMemcpyAsync(db->d_keys + (size_t)db->n * db->n_rings, ringkeys, (size_t)n * db->n_rings * sizeof(float), cudaMemcpyHostToDevice,
                             s->stream));
  return sc_append_ids(db, n, global_ids);
}

// the reference's values (SigType = vector<pair<int, double>>, ScanContext.h:24) kept as doubles for the exact re-score
int dslam_sc_add64(dslam_scdb *db, int n, const float *ringkeys, const double *sigs_dense64, const int *global_ids) {
  if (!db || n < 0) return fail(DSLAM_EINVAL, "bad argument");
  if (!db->fp64) return fail(DSLAM_ESTATE, "dslam_sc_add64 needs a database created with DSLAM_SC_FP64");
  if (n == 0) return DSLAM_OK;
  if (!ringkeys || !sigs_dense64) return fail(DSLAM_EINVAL, "null descriptor array");
  DSLAM_CUDA(cudaSetDevice(db->s->device));
  const int rc = sc_check_append(db, n, global_ids);
  if (rc != DSLAM_OK) return rc;
  dslam_session *s = db->s;
  DSLAM_CUDA(cudaMemcpyAsync(db->d_sigs64 + (size_t)db->n * db->n_cells, sigs_dense64, (size_t)n * db->n_cells * sizeof(double), cudaMemcpyHostToDevice,
                             s->stream));
  DSLAM_CUDA(launch_sc_narrow(db->d_sigs64 + (size_t)db->n * db->n_cells, db->d_sigs + (size_t)db->n * db->n_cells, (size_t)n * db->n_cells, s->stream));
  DSLAM_CUDA(cudaMemcpyAsync(db->d_keys + (size_t)db->n * db->n_rings, ringkeys, (size_t)n * db->n_rings * sizeof(float), cudaMemcpyHostToDevice,
                             s->stream));
  return sc_append_ids(db, n, global_ids);
}

int dslam_sc_add_sparse(dslam_scdb *db, const float *ringkey, const int *idx, const double *val, int nnz, int global_id) {
  if (!db || !ringkey || nnz < 0 || (nnz > 0 && (!idx || !val))) return fail(DSLAM_EINVAL, "bad argument");
  const int rc = ensure_stage(db, (size_t)db->n_cells * 2);
  if (rc != DSLAM_OK) return rc;
  DSLAM_CUDA(cudaStreamSynchronize(db->s->stream));
  const int id = global_id < 0 ? (db->ids.empty() ? 0 : db->ids.back() + 1) : global_id;
  if (db->fp64) {  // the double values travel unrounded; the fp32 scan copy is derived on the device
    double *st = reinterpret_cast<double *>(db->h_stage);
    std::memset(st, 0, sizeof(double) * db->n_cells);
    for (int i = 0; i < nnz; i++) {
      if (idx[i] < 0 || idx[i] >= db->n_cells) return fail(DSLAM_EINVAL, "cell index %d out of range", idx[i]);
      st[idx[i]] = val[i];
    }
    return dslam_sc_add64(db, 1, ringkey, st, &id);
  }
  std::memset(db->h_stage, 0, sizeof(float) * db->n_cells);
  for (int i = 0; i < nnz; i++) {
    if (idx[i] < 0 || idx[i] >= db->n_cells) return fail(DSLAM_EINVAL, "cell index %d out of range", idx[i]);
    db->h_stage[idx[i]] = (float)val[i];  // fp32 database: the device format is fp32
  }
  return dslam_sc_add(db, 1, ringkey, db->h_stage, &id);
}

// ScanContext::generate (src/loop_closure/loop_detection/ScanContext.cpp:78-142 with align_points_PCA :19-66).
int dslam_sc_generate(dslam_scdb *db, const double *pts_xyz, int n, double lidar_range, float *ringkey_out, float *sig_dense_out,
                      double *sig_dense64_out, double tfm_pca_rig[16], int append, int global_id) {
  if (!db || !pts_xyz || n < 1 || !(lidar_range > 0)) return fail(DSLAM_EINVAL, "bad argument");
  const int last = db->ids.empty() ? -1 : db->ids.back();
  const int id = global_id < 0 ? last + 1 : global_id;
  if (append && id <= last) return fail(DSLAM_EINVAL, "global ids must be strictly ascending within a shard (%d after %d)", id, last);
  dslam_session *s = db->s;
  DSLAM_CUDA(cudaSetDevice(s->device));
  if (append) {
    const int rc = ensure_capacity(db, db->n + 1);
    if (rc != DSLAM_OK) return rc;
  }
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
    for (int r = 0; r < 3; r++)  // Eigen's coefficient-based 3-term product: e0 + (e1 + e2), negation inside the coefficients
      tfm_pca_rig[r * 4 + 3] = (-tfm_pca_rig[r * 4]) * mom[0] + ((-tfm_pca_rig[r * 4 + 1]) * mom[1] + (-tfm_pca_rig[r * 4 + 2]) * mom[2]);
  }
  DSLAM_CUDA(launch_sc_bin_finalize(db->d_pts, n, mom, v9, lidar_range, db->n_sectors, db->n_rings, db->d_cells, db->d_gen_key, db->d_gen_sig,
                                    db->d_gen_sig64, s->stream));
  s->launches += 3;
  if (append) {  // the new descriptor goes into the database without touching the host
    DSLAM_CUDA(cudaMemcpyAsync(db->d_sigs + (size_t)db->n * db->n_cells, db->d_gen_sig, (size_t)db->n_cells * sizeof(float), cudaMemcpyDeviceToDevice, s->stream));
    if (db->fp64)
      DSLAM_CUDA(cudaMemcpyAsync(db->d_sigs64 + (size_t)db->n * db->n_cells, db->d_gen_sig64, (size_t)db->n_cells * sizeof(double),
                                 cudaMemcpyDeviceToDevice, s->stream));
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

// ---- on-disk format (.scdb): the database shard as it lives in HBM ---------------------------------------------------
//   header (64 bytes): magic "SCDB", u32 version = 1, u32 n_sectors, u32 n_rings, u32 rows, u32 flags (bit 0: fp64 table
//   follows), 40 reserved bytes; then int32 ids[rows], fp32 ringkeys[rows][n_rings], fp32 sigs[rows][n_cells] and, with
//   flag bit 0, fp64 sigs64[rows][n_cells].  Little endian, no padding.
struct ScdbHeader {
  char magic[4];
  unsigned version, n_sectors, n_rings, rows, flags;
  unsigned char reserved[40];
};
static_assert(sizeof(ScdbHeader) == 64, "ScdbHeader layout");

int dslam_sc_save(dslam_scdb *db, const char *path) {
  if (!db || !path) return fail(DSLAM_EINVAL, "null argument");
  DSLAM_CUDA(cudaSetDevice(db->s->device));
  FILE *f = fopen(path, "wb");
  if (!f) return fail(DSLAM_EINVAL, "cannot open %s for writing", path);
  ScdbHeader h;
  std::memset(&h, 0, sizeof(h));
  std::memcpy(h.magic, "SCDB", 4);
  h.version = 1; h.n_sectors = (unsigned)db->n_sectors; h.n_rings = (unsigned)db->n_rings; h.rows = (unsigned)db->n; h.flags = db->fp64 ? 1u : 0u;
  bool ok = fwrite(&h, sizeof(h), 1, f) == 1;
  if (db->n > 0) ok = ok && fwrite(db->ids.data(), sizeof(int), (size_t)db->n, f) == (size_t)db->n;
  std::vector<unsigned char> buf;
  cudaError_t ce = cudaSuccess;
  auto dump = [&](const void *dev, size_t bytes) {
    if (bytes == 0 || ce != cudaSuccess) return;
    buf.resize(bytes);
    ce = cudaMemcpyAsync(buf.data(), dev, bytes, cudaMemcpyDeviceToHost, db->s->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(db->s->stream);
    if (ce == cudaSuccess) ok = ok && fwrite(buf.data(), 1, bytes, f) == bytes;
  };
  dump(db->d_keys, (size_t)db->n * db->n_rings * sizeof(float));
  dump(db->d_sigs, (size_t)db->n * db->n_cells * sizeof(float));
  if (db->fp64) dump(db->d_sigs64, (size_t)db->n * db->n_cells * sizeof(double));
  ok = (fclose(f) == 0) && ok;
  if (ce != cudaSuccess) return cuda_fail(ce, "dslam_sc_save");
  return ok ? DSLAM_OK : fail(DSLAM_EINVAL, "short write to %s", path);
}

// appends the rows of the file (its ids must ascend past the rows already present; geometry must match)
int dslam_sc_load(dslam_scdb *db, const char *path, int *rows_out) {
  if (!db || !path) return fail(DSLAM_EINVAL, "null argument");
  DSLAM_CUDA(cudaSetDevice(db->s->device));
  FILE *f = fopen(path, "rb");
  if (!f) return fail(DSLAM_EINVAL, "cannot open %s", path);
  ScdbHeader h;
  int rc = DSLAM_OK;
  std::vector<int> ids;
  std::vector<float> keys, sigs;
  std::vector<double> sigs64;
  if (fread(&h, sizeof(h), 1, f) != 1 || std::memcmp(h.magic, "SCDB", 4) != 0 || h.version != 1)
    rc = fail(DSLAM_EINVAL, "%s is not a version-1 .scdb file", path);
  else if ((int)h.n_sectors != db->n_sectors || (int)h.n_rings != db->n_rings)
    rc = fail(DSLAM_EINVAL, "%s holds %ux%u descriptors, the database %dx%d", path, h.n_sectors, h.n_rings, db->n_sectors, db->n_rings);
  else {
    const size_t n = h.rows;
    ids.resize(n); keys.resize(n * db->n_rings); sigs.resize(n * db->n_cells);
    bool ok = n == 0 || (fread(ids.data(), sizeof(int), n, f) == n && fread(keys.data(), sizeof(float), keys.size(), f) == keys.size() &&
                         fread(sigs.data(), sizeof(float), sigs.size(), f) == sigs.size());
    if (ok && (h.flags & 1u) && db->fp64 && n > 0) {
      sigs64.resize(n * db->n_cells);
      ok = fread(sigs64.data(), sizeof(double), sigs64.size(), f) == sigs64.size();
    }
    if (!ok) rc = fail(DSLAM_EINVAL, "%s is truncated", path);
  }
  fclose(f);
  if (rc != DSLAM_OK) return rc;
  if (rows_out) *rows_out = (int)h.rows;
  if (h.rows == 0) return DSLAM_OK;
  if (!sigs64.empty()) return dslam_sc_add64(db, (int)h.rows, keys.data(), sigs64.data(), ids.data());
  return dslam_sc_add(db, (int)h.rows, keys.data(), sigs.data(), ids.data());
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

static int sc_search_sc_impl(dslam_scdb *db, int nq, const float *sigs_dense, const double *sigs_dense64, const int *candidates, int n_cand,
                             int *res_idx, float *res_diff) {
  if (!db || nq < 1 || (!sigs_dense && !sigs_dense64) || !candidates || n_cand < 1 || !res_idx || !res_diff) return fail(DSLAM_EINVAL, "bad argument");
  if (db->world > 1) return fail(DSLAM_ESTATE, "dslam_sc_search_sc needs the candidates' rows on this rank; use dslam_sc_query on a sharded database");
  DSLAM_CUDA(cudaSetDevice(db->s->device));
  int rc = upload_queries(db, nq, nullptr, sigs_dense, sigs_dense64);
  if (rc != DSLAM_OK) return rc;
  const int np = nq * n_cand;
  if (np > db->paircap) {
    cudaFree(db->d_pair_q); cudaFree(db->d_pair_row); cudaFree(db->d_pair_diff);
    db->d_pair_q = db->d_pair_row = nullptr;
    db->d_pair_diff = nullptr;
    db->paircap = 0;
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
  const void *tab = db->fp64 ? (const void *)db->d_sigs64 : (const void *)db->d_sigs;
  const void *qtab = db->fp64 ? (const void *)db->d_qsigs64 : (const void *)db->d_qsigs;
  DSLAM_CUDA(launch_sc_rescore_pairs(db->d_pair_q, db->d_pair_row, np, tab, db->fp64 ? 1 : 0, qtab, db->n_cells, db->n_sectors, db->d_pair_diff,
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

int dslam_sc_search_sc(dslam_scdb *db, int nq, const float *sigs_dense, const int *candidates, int n_cand, int *res_idx, float *res_diff) {
  return sc_search_sc_impl(db, nq, sigs_dense, nullptr, candidates, n_cand, res_idx, res_diff);
}
int dslam_sc_search_sc64(dslam_scdb *db, int nq, const double *sigs_dense64, const int *candidates, int n_cand, int *res_idx, float *res_diff) {
  return sc_search_sc_impl(db, nq, nullptr, sigs_dense64, candidates, n_cand, res_idx, res_diff);
}

int dslam_sc_query_keys(dslam_scdb *db, int nq, const float *ringkeys, const float *sigs_dense, float ringkey_thres, int max_id,
                        unsigned long long *keys_out) {
  if (!db || nq < 1 || !sigs_dense || !keys_out) return fail(DSLAM_EINVAL, "bad argument");
  DSLAM_CUDA(cudaSetDevice(db->s->device));
  unsigned seq = 0;
  const int rc = local_query(db, nq, ringkeys, sigs_dense, nullptr, ringkey_thres, max_id, kToHost, &seq);
  if (rc != DSLAM_OK) return rc;
  return wait_words(db, nq, seq, keys_out);
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

// One query batch against the whole (possibly sharded) database.  Single GPU: scan -> re-score -> the kernel writes the
// result words the host spins on.  Sharded: the re-score kernel additionally min-combines the keys in every rank's NVLink
// mailbox and publishes the combined keys (collective call: every rank passes the same batch).  Fallback exchange
// (no peer access, or DSLAM_SC_EXCHANGE=nccl): one ncclAllReduce(min, uint64 x Q) + a D2H copy.
static int sc_query_impl(dslam_scdb *db, int nq, const float *ringkeys, const float *sigs_dense, const double *sigs_dense64, float ringkey_thres,
                         int max_id, int *res_idx, float *res_diff) {
  if (!db || nq < 1 || (!sigs_dense && !sigs_dense64) || !res_idx) return fail(DSLAM_EINVAL, "bad argument");
  dslam_session *s = db->s;
  DSLAM_CUDA(cudaSetDevice(s->device));
  const int chunk_max = db->comm && db->p2p ? kScXchgMaxQ : nq;
  std::vector<u64> keys((size_t)nq);
  for (int q0 = 0; q0 < nq; q0 += chunk_max) {
    const int nqc = std::min(chunk_max, nq - q0);
    const float *rk = ringkeys ? ringkeys + (size_t)q0 * db->n_rings : nullptr;
    const float *sg = sigs_dense ? sigs_dense + (size_t)q0 * db->n_cells : nullptr;
    const double *sg64 = sigs_dense64 ? sigs_dense64 + (size_t)q0 * db->n_cells : nullptr;
    unsigned seq = 0;
    if (db->comm && !db->p2p) {
      int rc = local_query(db, nqc, rk, sg, sg64, ringkey_thres, max_id, kToDevice, &seq);
      if (rc != DSLAM_OK) return rc;
      NcclApi *api = nccl_api();
      const ncclResult_t r = api->AllReduce(db->d_best, db->d_best, (size_t)nqc, ncclUint64, ncclMin, db->comm, s->stream);
      if (r != ncclSuccess) return nccl_fail(r, "ncclAllReduce");
      DSLAM_CUDA(cudaMemcpyAsync(db->h_keys, db->d_best, (size_t)nqc * sizeof(u64), cudaMemcpyDeviceToHost, s->stream));
      DSLAM_CUDA(cudaStreamSynchronize(s->stream));
      std::memcpy(keys.data() + q0, db->h_keys, (size_t)nqc * sizeof(u64));
    } else {
      int rc = local_query(db, nqc, rk, sg, sg64, ringkey_thres, max_id, db->comm ? kExchange : kToHost, &seq);
      if (rc != DSLAM_OK) return rc;
      rc = wait_words(db, nqc, seq, keys.data() + q0);
      if (rc != DSLAM_OK) return rc;
    }
  }
  for (int q = 0; q < nq; q++) {
    int id;
    float d;
    dslam_sc_decode_key(keys[q], &id, &d);
    res_idx[q] = id;
    if (res_diff) res_diff[q] = d;
  }
  return DSLAM_OK;
}

int dslam_sc_query(dslam_scdb *db, int nq, const float *ringkeys, const float *sigs_dense, float ringkey_thres, int max_id, int *res_idx,
                   float *res_diff) {
  return sc_query_impl(db, nq, ringkeys, sigs_dense, nullptr, ringkey_thres, max_id, res_idx, res_diff);
}
int dslam_sc_query64(dslam_scdb *db, int nq, const float *ringkeys, const double *sigs_dense64, float ringkey_thres, int max_id, int *res_idx,
                     float *res_diff) {
  return sc_query_impl(db, nq, ringkeys, nullptr, sigs_dense64, ringkey_thres, max_id, res_idx, res_diff);
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

// NVLink mailboxes: every rank allocates one, exports it as a CUDA IPC handle, the handles travel through one ncclAllGather,
// and every rank maps its peers' mailboxes (peer access enabled lazily by cudaIpcOpenMemHandle).  All ranks agree on the
// outcome through one ncclAllReduce(min) of a success flag, so either all use the mailboxes or all use the NCCL exchange.
static int sc_setup_mailboxes(dslam_scdb *db) {
  NcclApi *api = nccl_api();
  dslam_session *s = db->s;
  const int world = db->world, rank = db->rank;
  int ok = world <= 8 ? 1 : 0;
  const char *mode = getenv("DSLAM_SC_EXCHANGE");
  if (mode && !strcmp(mode, "nccl")) ok = 0;
  cudaIpcMemHandle_t mine;
  std::memset(&mine, 0, sizeof(mine));
  if (ok) {
    if (cudaMalloc(&db->d_mail, sc_exchange_bytes()) != cudaSuccess) ok = 0;
    if (ok && cudaMemsetAsync(db->d_mail, 0xff, sc_exchange_arrived_offset(), s->stream) != cudaSuccess) ok = 0;
    if (ok && cudaMemsetAsync((char *)db->d_mail + sc_exchange_arrived_offset(), 0, 256, s->stream) != cudaSuccess) ok = 0;
    if (ok && cudaStreamSynchronize(s->stream) != cudaSuccess) ok = 0;
    if (ok && cudaIpcGetMemHandle(&mine, db->d_mail) != cudaSuccess) ok = 0;
    cudaGetLastError();
  }
  // gather the handles (device buffers: NCCL moves device memory)
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
  unsigned char *d_h = nullptr;
  std::vector<cudaIpcMemHandle_t> all((size_t)world);
  DSLAM_CUDA(cudaMalloc((void **)&d_h, (size_t)64 * (world + 1)));
  DSLAM_CUDA(cudaMemcpyAsync(d_h + (size_t)64 * world, &mine, 64, cudaMemcpyHostToDevice, s->stream));
  ncclResult_t r = api->AllGather(d_h + (size_t)64 * world, d_h, 64, ncclChar, db->comm, s->stream);
  if (r != ncclSuccess) { cudaFree(d_h); return nccl_fail(r, "ncclAllGather (mailbox handles)"); }
  DSLAM_CUDA(cudaMemcpyAsync(all.data(), d_h, (size_t)64 * world, cudaMemcpyDeviceToHost, s->stream));
  DSLAM_CUDA(cudaStreamSynchronize(s->stream));
  if (ok) {
    for (int p = 0; p < world && ok; p++) {
      void *ptr = db->d_mail;
      if (p != rank) {
        ptr = nullptr;
        if (cudaIpcOpenMemHandle(&ptr, all[p], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); break; }
        db->peer_mail[p] = ptr;
      }
      db->xchg.keys[p] = reinterpret_cast<unsigned long long *>(ptr);
      db->xchg.arrived[p] = reinterpret_cast<unsigned *>(reinterpret_cast<char *>(ptr) + sc_exchange_arrived_offset());
    }
  }
  // agreement: min over the ranks of the success flag
  unsigned long long flag = ok ? 1ull : 0ull;
  DSLAM_CUDA(cudaMemcpyAsync(d_h, &flag, 8, cudaMemcpyHostToDevice, s->stream));
  r = api->AllReduce(d_h, d_h, 1, ncclUint64, ncclMin, db->comm, s->stream);
  if (r != ncclSuccess) { cudaFree(d_h); return nccl_fail(r, "ncclAllReduce (mailbox agreement)"); }
  DSLAM_CUDA(cudaMemcpyAsync(&flag, d_h, 8, cudaMemcpyDeviceToHost, s->stream));
  DSLAM_CUDA(cudaStreamSynchronize(s->stream));
  cudaFree(d_h);
  db->p2p = flag == 1ull;
  db->xchg.world = world;
  db->xchg.rank = rank;
  db->xseq = 0;
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
  return sc_setup_mailboxes(db);
}

// 1 = the sharded query exchanges its keys through the NVLink mailboxes, 0 = through ncclAllReduce (or not sharded)
int dslam_sc_exchange_mode(dslam_scdb *db, int *p2p_out) {
  if (!db || !p2p_out) return fail(DSLAM_EINVAL, "null argument");
  *p2p_out = db->comm && db->p2p ? 1 : 0;
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
  if (flavour < 0 || flavour > 4) return fail(DSLAM_EINVAL, "scan kernel flavour must be 0 (auto), 1 (stream), 2 (tile), 3 (tcgen05) or 4 (tcgen05, masked hi operand)");
  sc_set_scan_flavour(flavour);
  return DSLAM_OK;
}

}  // extern "C"
