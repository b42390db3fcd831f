// kernels_sc.cu — Scan-Context loop-candidate search on the device.  sm_100a.
//
// Replaces the two search stages of src/loop_closure/loop_detection/search_place.h:
//   search_ringkey :25-57  (FLANN kd-tree kNN on the 20-float ring keys)  -> sc_ringkey_kernel: EXACT brute-force
//       k-nearest by squared L2 in flann::L2's arithmetic (4 squared differences summed left to right per step,
//       FLANN 1.9.1 flann/algorithms/dist.h), so the "< RINGKEY_THRES" gate (:35) is bit-exact;
//   search_sc      :59-85  (sector-cosine distance over the candidates)   -> sc_scan_kernel: the same distance
//       (1 - sum_sectors cos / sc_width) / 2 evaluated against EVERY database row (dense 60x20 fp32 descriptors,
//       columns L2-normalised at generation, ScanContext.cpp:137-141, so the sum of cosines is one 1200-long dot
//       product), keeping the kScTopK best per query.  The host re-scores those K in the reference's exact
//       arithmetic (float += double*double) so the final argmin is bit-exact (dslam_api.cu).
//
// The scan is HBM-bound pointwise work (4,800 B per row, ~0.5 flop/B per query): one warp streams whole rows with
// 16-B coalesced loads (10 x LDG.128 per lane per row, two rows in flight per warp), the query batch (<= 32
// queries per pass) sits in shared memory, lane q owns the running top-K of query q.  No tensor cores: the
// products must stay fp32-exact enough for the host re-rank to see the true winner, and at Q <= 5 the kernel is
// bandwidth-bound anyway.  Keys are packed (ordered_float_bits(dist) << 32 | global_id), so "min" is the
// argmin with ties going to the lowest id — the same key a multi-GPU all-reduce(min) combines.

#include "dslam_kernels.h"
#include <cstdlib>
#include <cstring>

namespace dslam {

namespace {

typedef unsigned long long u64;
constexpr u64 kKeyMax = ~0ull;
constexpr int kScanThreads = 512;
constexpr int kScanWarps = kScanThreads / 32;
constexpr int kQChunk = 32;

__device__ __forceinline__ u64 make_key(float dist, int id) {
  unsigned b = __float_as_uint(dist);
  b ^= (b >> 31) ? 0xffffffffu : 0x80000000u;  // total order for signed floats
  return ((u64)b << 32) | (unsigned)id;
}

struct TopK {
  u64 k[kScTopK];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int i = 0; i < kScTopK; i++) k[i] = kKeyMax;
  }
  __device__ __forceinline__ void insert(u64 key) {
    if (key < k[kScTopK - 1]) {
      k[kScTopK - 1] = key;
#pragma unroll
      for (int i = kScTopK - 1; i > 0; i--) {
        const u64 a = k[i - 1], b = k[i];
        const bool sw = b < a;
        k[i - 1] = sw ? b : a;
        k[i] = sw ? a : b;
      }
    }
  }
  __device__ __forceinline__ void pop_front() {
#pragma unroll
    for (int i = 0; i < kScTopK - 1; i++) k[i] = k[i + 1];
    k[kScTopK - 1] = kKeyMax;
  }
};

__device__ __forceinline__ u64 warp_min_u64(u64 v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
    const u64 o = __shfl_xor_sync(0xffffffffu, v, m);
    v = o < v ? o : v;
  }
  return v;
}

// union of the 32 per-lane sorted lists -> the kScTopK smallest, returned in every lane
__device__ __forceinline__ void warp_merge(TopK &t, u64 (&out)[kScTopK]) {
#pragma unroll
  for (int r = 0; r < kScTopK; r++) {
    const u64 m = warp_min_u64(t.k[0]);
    out[r] = m;
    if (t.k[0] == m && m != kKeyMax) t.pop_front();
  }
}

// flann::L2<float>: result += d0*d0 + d1*d1 + d2*d2 + d3*d3 per group of 4 (compiled -fmad=false)
__device__ __forceinline__ float flann_l2(const float *__restrict__ a, const float *__restrict__ b, int dim) {
  float result = 0.f;
  int i = 0;
  for (; i + 3 < dim; i += 4) {
    const float d0 = a[i] - b[i], d1 = a[i + 1] - b[i + 1], d2 = a[i + 2] - b[i + 2], d3 = a[i + 3] - b[i + 3];
    result += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
  }
  for (; i < dim; i++) {
    const float d0 = a[i] - b[i];
    result += d0 * d0;
  }
  return result;
}

// ---- ring-key kNN: grid (X, nq), block 256; each thread scans rows x*256+tid, stride X*256 -------------------
__global__ void __launch_bounds__(256) sc_ringkey_kernel(const float *__restrict__ keys, const int *__restrict__ ids, int n_rows, int dim,
                                                        const float *__restrict__ queries, int max_id, u64 *__restrict__ scratch) {
  __shared__ float qk[64];
  __shared__ u64 wl[8][kScTopK];
  const int q = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < dim) qk[tid] = queries[(size_t)q * dim + tid];
  __syncthreads();
  TopK t;
  t.init();
  const bool vec4 = (dim & 3) == 0;
  for (int r = blockIdx.x * 256 + tid; r < n_rows; r += gridDim.x * 256) {
    const int id = ids[r];
    if (id >= max_id) continue;
    const float *kr = keys + (size_t)r * dim;
    float result = 0.f;
    int i = 0;
    for (; vec4 && i + 3 < dim; i += 4) {
      const float4 kv = __ldg(reinterpret_cast<const float4 *>(kr + i));
      const float d0 = qk[i] - kv.x, d1 = qk[i + 1] - kv.y, d2 = qk[i + 2] - kv.z, d3 = qk[i + 3] - kv.w;
      result += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
    for (; i + 3 < dim; i += 4) {  // rows are only 16-B aligned when dim % 4 == 0: scalar loads otherwise, same arithmetic
      const float d0 = qk[i] - __ldg(kr + i), d1 = qk[i + 1] - __ldg(kr + i + 1), d2 = qk[i + 2] - __ldg(kr + i + 2), d3 = qk[i + 3] - __ldg(kr + i + 3);
      result += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
    for (; i < dim; i++) {
      const float d0 = qk[i] - __ldg(kr + i);
      result += d0 * d0;
    }
    t.insert(make_key(result, id));
  }
  u64 out[kScTopK];
  warp_merge(t, out);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < kScTopK; i++) wl[warp][i] = out[i];
  }
  __syncthreads();
  if (warp == 0) {
    TopK m;
    m.init();
    if (lane < 8) {
#pragma unroll
      for (int i = 0; i < kScTopK; i++) m.insert(wl[lane][i]);
    }
    warp_merge(m, out);
    if (lane == 0) {
      u64 *dst = scratch + ((size_t)q * gridDim.x + blockIdx.x) * kScTopK;
#pragma unroll
      for (int i = 0; i < kScTopK; i++) dst[i] = out[i];
    }
  }
}

// ---- final merge: grid nq, block 32: scratch[q][nlists][K] -> out[q][K] ---------------------------------------
__global__ void __launch_bounds__(32) sc_merge_kernel(const u64 *__restrict__ scratch, int nlists, u64 *__restrict__ out) {
  const int q = blockIdx.x, lane = threadIdx.x;
  TopK t;
  t.init();
  const u64 *src = scratch + (size_t)q * nlists * kScTopK;
  for (int i = lane; i < nlists * kScTopK; i += 32) t.insert(src[i]);
  u64 o[kScTopK];
  warp_merge(t, o);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < kScTopK; i++) out[(size_t)q * kScTopK + i] = o[i];
  }
}

// ---- sector-cosine scan: persistent grid, 16 warps per CTA, warp streams rows, lane q owns query q ------------
// dynamic smem: qs[nqc][n_cells] floats followed by qkeys[nqc][key_dim]
__global__ void __launch_bounds__(kScanThreads) sc_scan_kernel(const float *__restrict__ sigs, const float *__restrict__ keys,
                                                              const int *__restrict__ ids, int n_rows, int n_cells, int key_dim,
                                                              const float *__restrict__ q_sigs, const float *__restrict__ q_keys, int nqc,
                                                              float ringkey_thres, int max_id, float sc_width, u64 *__restrict__ scratch,
                                                              int list_stride) {
  extern __shared__ __align__(16) float smem[];
  float *qs = smem;
  float *qk = smem + (size_t)nqc * n_cells;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < nqc * n_cells / 4; i += kScanThreads)
    reinterpret_cast<float4 *>(qs)[i] = __ldg(reinterpret_cast<const float4 *>(q_sigs) + i);
  if (ringkey_thres >= 0.f)  // without the ring-key gate the caller need not have uploaded query keys
    for (int i = tid; i < nqc * key_dim; i += kScanThreads) qk[i] = q_keys[i];
  __syncthreads();

  const int n4 = n_cells / 4;  // float4 per row (300)
  constexpr int R4 = 10;       // float4 per lane per row: supports n_cells <= 1280
  TopK top;
  top.init();
  const int gw = blockIdx.x * kScanWarps + warp, nw = gridDim.x * kScanWarps;
  for (int r0 = gw * 2; r0 < n_rows; r0 += nw * 2) {
    const bool has1 = r0 + 1 < n_rows;
    const float4 *ra = reinterpret_cast<const float4 *>(sigs + (size_t)r0 * n_cells);
    const float4 *rb = reinterpret_cast<const float4 *>(sigs + (size_t)(has1 ? r0 + 1 : r0) * n_cells);
    float4 a[R4], b[R4];
#pragma unroll
    for (int j = 0; j < R4; j++) {
      const int c = lane + 32 * j;
      a[j] = c < n4 ? __ldcs(ra + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      b[j] = c < n4 ? __ldcs(rb + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const int ida = ids[r0], idb = has1 ? ids[r0 + 1] : 0x7fffffff;
    float my_da = 0.f, my_db = 0.f;  // lane q keeps the distances of query q
    for (int q = 0; q < nqc; q++) {
      const float4 *qv = reinterpret_cast<const float4 *>(qs + (size_t)q * n_cells);
      float sa = 0.f, sb = 0.f;
#pragma unroll
      for (int j = 0; j < R4; j++) {
        const int c = lane + 32 * j;
        if (c < n4) {
          const float4 x = qv[c];
          sa = fmaf(a[j].x, x.x, sa); sa = fmaf(a[j].y, x.y, sa); sa = fmaf(a[j].z, x.z, sa); sa = fmaf(a[j].w, x.w, sa);
          sb = fmaf(b[j].x, x.x, sb); sb = fmaf(b[j].y, x.y, sb); sb = fmaf(b[j].z, x.z, sb); sb = fmaf(b[j].w, x.w, sb);
        }
      }
#pragma unroll
      for (int m = 16; m > 0; m >>= 1) {
        sa += __shfl_xor_sync(0xffffffffu, sa, m);
        sb += __shfl_xor_sync(0xffffffffu, sb, m);
      }
      if (lane == q) {
        my_da = (1.0f - sa / sc_width) / 2.0f;
        my_db = (1.0f - sb / sc_width) / 2.0f;
      }
    }
    if (lane < nqc) {
      bool oka = ida < max_id, okb = has1 && idb < max_id;
      if (ringkey_thres >= 0.f) {
        if (oka) oka = flann_l2(qk + lane * key_dim, keys + (size_t)r0 * key_dim, key_dim) < ringkey_thres;
        if (okb) okb = flann_l2(qk + lane * key_dim, keys + (size_t)(r0 + 1) * key_dim, key_dim) < ringkey_thres;
      }
      if (oka) top.insert(make_key(my_da, r0));
      if (okb) top.insert(make_key(my_db, r0 + 1));
    }
  }
  // CTA merge: reuse the query smem for the per-warp lists [warp][q][K]
  __syncthreads();
  u64 *lists = reinterpret_cast<u64 *>(smem);
  if (lane < nqc) {
#pragma unroll
    for (int i = 0; i < kScTopK; i++) lists[((size_t)warp * kQChunk + lane) * kScTopK + i] = top.k[i];
  }
  __syncthreads();
  if (tid < nqc) {
    TopK m;
    m.init();
    for (int wv = 0; wv < kScanWarps; wv++) {
#pragma unroll
      for (int i = 0; i < kScTopK; i++) m.insert(lists[((size_t)wv * kQChunk + tid) * kScTopK + i]);
    }
    u64 *dst = scratch + ((size_t)tid * list_stride + blockIdx.x) * kScTopK;  // lists[q][cta][K]: merged by the re-score kernel
#pragma unroll
    for (int i = 0; i < kScTopK; i++) dst[i] = m.k[i];
  }
}

// ---- sector-cosine scan, batched flavour (query batches > 8): register-blocked tiles fed by TMA ----------------
// With a batch of queries the scan is a tall-skinny product D[rows x 32] = DB[rows x n_cells] * Q^T: 16 flop per DB
// byte at 32 queries, which is ABOVE the fp32 ridge of the machine (~72 TFLOP/s / 6.5 TB/s = 11 flop/B), so the
// batched scan is bound by the FFMA pipe, not by HBM, and the streaming kernel above (one LDS.128 per 8 FFMA, a
// 5-step shuffle reduction per row and query) leaves it 5x short of that.  Here a CTA works on tiles of 256 DB rows:
// a producer warp streams [4 x 64 rows x 32 cells] boxes of the DB (TMA, 128-byte swizzle) and the matching
// [32 queries x 32 cells] box of the query batch through a 4-stage mbarrier pipeline; each of the 8 consumer warps
// owns 128 rows x 8 queries, a lane 4 rows x 8 queries = 32 accumulators.  Per 4 cells a lane issues 4 LDS.128 of
// its rows (conflict-free through the swizzle) + 8 broadcast LDS.128 of the queries for 128 FFMA: 0.19 shared-memory
// wavefronts per FFMA (the SM sustains 0.25) and no shuffles.  The finished tile goes through shared memory once
// ([row][query], padded) so that thread (query = tid % 32, row group = tid / 32) feeds its running top-K — the same
// packed keys and the same merge as the streaming kernel.  fp32 CUDA-core FFMA on purpose (no TF32 tensor cores):
// the approximate ranking must keep the true winner inside the top-K that the exact re-score sees.
constexpr int kTileRows = 256;   // DB rows per tile
constexpr int kTileK = 32;       // cells per pipeline stage (128 B per row: one swizzle atom)
constexpr int kTileStages = 4;
constexpr int kTileConsumers = 256;
constexpr int kTileThreads = kTileConsumers + 32;  // + the producer warp
constexpr int kDistStride = 36;  // floats per row of the distance tile (conflict-free STS.128 / LDS.32)
constexpr int kStageABytes = kTileRows * kTileK * 4;  // 32 KB
constexpr int kStageBBytes = kQChunk * kTileK * 4;    // 4 KB
constexpr size_t kTileSmemBytes = 1024 /*alignment slack*/ + (size_t)kTileStages * (kStageABytes + kStageBBytes) +
                                  (size_t)kTileRows * kDistStride * 4 + 2 * kTileStages * sizeof(uint64_t);

__device__ __forceinline__ uint32_t sc_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sc_mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sc_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void sc_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sc_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sc_mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool sc_mbar_try_wait(uint64_t *bar, uint32_t phase) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(sc_smem_u32(bar)), "r"(phase)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void sc_mbar_wait(uint64_t *bar, uint32_t phase) {
  while (!sc_mbar_try_wait(bar, phase)) {
  }
}
__device__ __forceinline__ void sc_tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   sc_smem_u32(dst)),
               "l"(map), "r"(c0), "r"(c1), "r"(sc_smem_u32(bar))
               : "memory");
}
// plain (1-D) bulk copy global -> shared, completion on an mbarrier: 16-byte aligned addresses, size a multiple of 16
__device__ __forceinline__ void sc_bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sc_smem_u32(dst)), "l"(src), "r"(bytes),
               "r"(sc_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void sc_consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kTileConsumers) : "memory"); }

// Work split: the database is cut into HALF-GROUPS of 32 rows (one TMA box, one consumer warp's rows); CTA c owns the
// contiguous run of half-groups [c * hpc, (c+1) * hpc) and walks it in tiles of up to 8 half-groups (256 rows), so all CTAs
// carry the same load to within 32 rows: a 100k-row database fills 148 CTAs to 99 %, a 12.5k-row shard (8 GPUs) still keeps
// 131 CTAs busy in one wave (a 64-row granularity left 50 SMs idle there, a 256-row one idles 12 % at 100k rows).  Half-group t
// of a tile lands at rows [32 t, 32 t + 32) of the stage — two consecutive boxes lie exactly like one 64-row group — and belongs
// to the consumer warps of row half rh = t & 1.  A partial last tile loads and multiplies only its half-groups.
constexpr int kHalfRows = 32;
constexpr int kHalvesPerTile = kTileRows / kHalfRows;  // 8
constexpr int kHalfBytes = kHalfRows * kTileK * 4;     // 4 KB per TMA box
constexpr int kGroupRows = 64;                         // rows between consecutive groups of one lane

// One pipeline stage (32 cells) of a tile with NG row groups: per 4 cells a lane reads its NG rows (LDS.128, conflict-free
// through the swizzle) and the 8 queries of its warp (broadcast LDS.128).  Packed FP32 FMAs (FFMA2, sm_100: two FMAs on a
// 64-bit register pair per instruction); the pairs run along the cell index — (a.x, a.y) * (b.x, b.y) are adjacent registers
// of the LDS.128 results — into an even-cell and an odd-cell partial sum per (row, query), added at the end of the tile.
// Measured: 10 % faster than scalar FFMA (same FMA pipe — a 3-register FFMA issues every other cycle per sub-partition,
// 64 lanes / clock / SM — but half the issue slots).
template <int NG>
__device__ __forceinline__ void sc_tile_stage(const float4 *__restrict__ A, const float4 *__restrict__ B, int sw, float2 (&acc)[4][8]) {
#pragma unroll
  for (int k4 = 0; k4 < kTileK / 4; k4++) {
    float4 a[NG];
#pragma unroll
    for (int i = 0; i < NG; i++) a[i] = A[(size_t)i * kGroupRows * (kTileK / 4) + (k4 ^ sw)];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const float4 b = B[j * (kTileK / 4) + k4];
#pragma unroll
      for (int i = 0; i < NG; i++) {
        acc[i][j] = __ffma2_rn(make_float2(a[i].x, a[i].y), make_float2(b.x, b.y), acc[i][j]);
        acc[i][j] = __ffma2_rn(make_float2(a[i].z, a[i].w), make_float2(b.z, b.w), acc[i][j]);
      }
    }
  }
}

__global__ void __launch_bounds__(kTileThreads, 1)
    sc_scan_tile_kernel(const __grid_constant__ CUtensorMap map_db, const __grid_constant__ CUtensorMap map_q, const float *__restrict__ keys,
                        const int *__restrict__ ids, int n_rows, int n_cells, int key_dim, const float *__restrict__ q_keys, int nqc,
                        float ringkey_thres, int max_id, float sc_width, int halves_per_cta, u64 *__restrict__ scratch, int list_stride) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char *base = smem_raw + ((1024u - (sc_smem_u32(smem_raw) & 1023u)) & 1023u);  // swizzle atoms need 1024-B alignment
  float *sA = reinterpret_cast<float *>(base);
  float *sB = reinterpret_cast<float *>(base + (size_t)kTileStages * kStageABytes);
  float *sD = reinterpret_cast<float *>(base + (size_t)kTileStages * (kStageABytes + kStageBBytes));
  uint64_t *full = reinterpret_cast<uint64_t *>(sD + kTileRows * kDistStride);
  uint64_t *empty = full + kTileStages;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_halves = (n_rows + kHalfRows - 1) / kHalfRows;
  const int h_begin = blockIdx.x * halves_per_cta;
  const int h_end = min(h_begin + halves_per_cta, n_halves);
  const int row_end = min(h_end * kHalfRows, n_rows);
  const int n_chunks = (n_cells + kTileK - 1) / kTileK;
  if (tid == 0) {
    for (int s = 0; s < kTileStages; s++) {
      sc_mbar_init(&full[s], 1);
      sc_mbar_init(&empty[s], kTileConsumers / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == kTileConsumers / 32) {
    // ---- producer warp: one lane walks (tile, chunk) and keeps kTileStages stages in flight ----
    if (lane == 0) {
      int it = 0;
      for (int h0 = h_begin; h0 < h_end; h0 += kHalvesPerTile) {
        const int nh = min(kHalvesPerTile, h_end - h0);
        for (int c = 0; c < n_chunks; c++, it++) {
          const int s = it % kTileStages;
          sc_mbar_wait(&empty[s], ((it / kTileStages) & 1) ^ 1);  // passes immediately on the first lap
          sc_mbar_expect_tx(&full[s], nh * kHalfBytes + kStageBBytes);
          unsigned char *dst = reinterpret_cast<unsigned char *>(sA) + (size_t)s * kStageABytes;
          for (int t = 0; t < nh; t++) sc_tma_load_2d(dst + (size_t)t * kHalfBytes, &map_db, c * kTileK, (h0 + t) * kHalfRows, &full[s]);
          sc_tma_load_2d(reinterpret_cast<unsigned char *>(sB) + (size_t)s * kStageBBytes, &map_q, c * kTileK, 0, &full[s]);
        }
      }
    }
    return;
  }

  // ---- consumers ----
  // warp = (query octet qt, row half rh); lane owns rows g * 64 + rh * 32 + lane of the tile's groups g = 0..3
  const int qt = warp & 3, rh = warp >> 2;
  const int q_own = lane, sub = warp;  // top-K ownership: query q_own, tile rows sub*32 .. sub*32+31
  const int sw = lane & 7;             // 128-B swizzle: 16-B chunk index ^ (row & 7); row & 7 == lane & 7 for all rows of a lane
  TopK top;
  top.init();
  int it = 0;
  for (int h0 = h_begin; h0 < h_end; h0 += kHalvesPerTile) {
    const int nh = min(kHalvesPerTile, h_end - h0);
    const int ng = (nh - rh + 1) >> 1;  // half-groups of this tile that belong to this warp's row half: t = rh, rh + 2, ...
    float2 acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 8; j++) acc[i][j] = make_float2(0.f, 0.f);
    for (int c = 0; c < n_chunks; c++, it++) {
      const int s = it % kTileStages;
      sc_mbar_wait(&full[s], (it / kTileStages) & 1);
      const float4 *A = reinterpret_cast<const float4 *>(reinterpret_cast<const unsigned char *>(sA) + (size_t)s * kStageABytes) +
                        (size_t)(rh * 32 + lane) * (kTileK / 4);
      const float4 *B = reinterpret_cast<const float4 *>(reinterpret_cast<const unsigned char *>(sB) + (size_t)s * kStageBBytes) +
                        (size_t)(qt * 8) * (kTileK / 4);
      if (ng == 4) sc_tile_stage<4>(A, B, sw, acc);
      else if (ng == 3) sc_tile_stage<3>(A, B, sw, acc);  // partial last tile of the CTA's run: only its half-groups were loaded
      else if (ng == 2) sc_tile_stage<2>(A, B, sw, acc);
      else if (ng == 1) sc_tile_stage<1>(A, B, sw, acc);
      __syncwarp();
      if (lane == 0) sc_mbar_arrive(&empty[s]);
    }
    // distances of the tile -> shared [row][query]
#pragma unroll
    for (int i = 0; i < 4; i++) {
      float d[8];
#pragma unroll
      for (int j = 0; j < 8; j++) d[j] = (1.0f - (acc[i][j].x + acc[i][j].y) / sc_width) / 2.0f;
      float4 *dst = reinterpret_cast<float4 *>(sD + (size_t)(i * kGroupRows + rh * 32 + lane) * kDistStride + qt * 8);
      dst[0] = make_float4(d[0], d[1], d[2], d[3]);
      dst[1] = make_float4(d[4], d[5], d[6], d[7]);
    }
    sc_consumer_sync();
    if (q_own < nqc) {
      const int row0 = h0 * kHalfRows + sub * 32;
#pragma unroll 4
      for (int r = 0; r < 32; r++) {
        const int row = row0 + r;
        if (row >= row_end) break;
        bool ok = __ldg(ids + row) < max_id;
        if (ok && ringkey_thres >= 0.f) ok = flann_l2(q_keys + (size_t)q_own * key_dim, keys + (size_t)row * key_dim, key_dim) < ringkey_thres;
        if (ok) top.insert(make_key(sD[(size_t)(sub * 32 + r) * kDistStride + q_own], row));
      }
    }
    sc_consumer_sync();
  }
  // CTA merge through the distance tile's storage: lists[sub][q][K]
  u64 *lists = reinterpret_cast<u64 *>(sD);
#pragma unroll
  for (int i = 0; i < kScTopK; i++) lists[((size_t)sub * kQChunk + q_own) * kScTopK + i] = top.k[i];
  sc_consumer_sync();
  if (tid < nqc) {
    TopK m;
    m.init();
    for (int wv = 0; wv < kTileConsumers / 32; wv++) {
#pragma unroll
      for (int i = 0; i < kScTopK; i++) m.insert(lists[((size_t)wv * kQChunk + tid) * kScTopK + i]);
    }
    u64 *dst = scratch + ((size_t)tid * list_stride + blockIdx.x) * kScTopK;
#pragma unroll
    for (int i = 0; i < kScTopK; i++) dst[i] = m.k[i];
  }
}

// ---- sector-cosine scan on the 5th-generation tensor cores (tcgen05.mma + TMEM): query batches --------------------
// With a batch of queries the scan is D[rows x Q] = DB[rows x cells] * Q^T: 16 flop per DB byte at Q = 32, above the fp32
// CUDA-core ridge — the FFMA tile kernel above tops out at a third of the HBM rate.  The tensor cores have the headroom, but
// the top-K must still contain the exact winner, so a plain TF32 product (10-bit mantissas) is not good enough: every fp32
// operand is split into hi (the 11 significant bits TF32 keeps, exactly) + lo (the remainder, exact in fp32) and the product is
// formed as hi*hi + hi*lo + lo*hi with fp32 accumulation in TMEM ("3xTF32"; the dropped lo*lo term is below 2^-22 relative).
//   warp 0      : producer — TMA [128 rows x 32 cells] boxes of the DB (128-byte swizzle = the canonical K-major UMMA layout) and one
//                 bulk copy of the matching pre-split, pre-swizzled [b_hi ; b_lo] tile of the query batch, UmCfg::kStages deep
//   warps 1-4   : splitters — write lo = x - hi of the landed DB box next to it (elementwise: the swizzle does not matter), then fence
//                 the generic-proxy writes towards the async proxy the tensor cores read through.  The box itself serves as the hi
//                 operand as it is: a tf32 operand's low 13 mantissa bits are not read by the tensor cores (truncation), so
//                 raw x and the masked hi are the same operand (flavour "umma_masked" stores the masked hi in place instead and
//                 does not rely on that; tests/test_gpu_scan_context.py::test_tensor_core_split_is_exact pins the behaviour)
//   warp 5      : one lane issues 8 x tcgen05.mma.kind::tf32 per stage (per 8 cells: a_hi x [b_hi ; b_lo] with N = 2 NQ, a_lo x b_hi with N = NQ; M 128)
//                 into one of two TMEM accumulators,
//                 tcgen05.commit hands the stage back to the producer and, after the last stage of a tile, the tile to the epilogue
//   warps 6-9   : epilogue — tcgen05.ld the accumulator in chunks of 32 queries (thread = row, 32 columns = queries), distances -> shared
//                 [row][query], then thread (query, row group) feeds its running top-K of that chunk: same packed keys, same merge,
//                 same exact re-score.
// The DB streams through shared memory exactly once per pass (4,800 B per row per NQ queries).  Measured bound: the instruction rate of
// the K = 8 tf32 MMA on a 128-byte-swizzled operand (~140 clocks per MMA, N <= 128), not HBM — profiles/r02_sc_umma_kernel.md.
constexpr int kUmRows = 128;                 // UMMA M: DB rows per tile
constexpr int kUmK = 32;                     // cells per stage = one 128-byte swizzle atom row
constexpr int kUmABytes = kUmRows * kUmK * 4;              // 16 KB
constexpr int kUmThreads = 320;
// Queries per pass NQ = UMMA N of the lo*hi product (the [hi*hi | hi*lo] product runs at N = 2 NQ <= 256).  A pass streams the DB
// once, so a larger NQ divides the HBM traffic and the MMA count per query: 32 (5 stages of 40 KB), 64 (4 x 48 KB) or 128 (3 x 64 KB).
template <int NQ>
struct UmCfg {
  static_assert(NQ == 32 || NQ == 64 || NQ == 128, "queries per pass");
  static constexpr int kStages = NQ == 32 ? 5 : (NQ == 64 ? 4 : 3);
  static constexpr int kBBytes = NQ * kUmK * 4;                    // 4 / 8 / 16 KB
  static constexpr int kStageBytes = 2 * kUmABytes + 2 * kBBytes;  // A hi | A lo | B hi | B lo
  static constexpr int kAccCols = 2 * NQ;       // one accumulator stage: columns [0, NQ) = hi*hi + lo*hi, [NQ, 2 NQ) = hi*lo
  static constexpr int kTmemCols = 2 * kAccCols;  // two accumulator stages: 128 / 256 / 512 columns (a power of two >= 32)
  static constexpr size_t kSmemBytes = 1024 + (size_t)kStages * kStageBytes + (size_t)kUmRows * kDistStride * 4 + 32 * sizeof(uint64_t);
};

// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row atoms 1024 B apart (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ u64 um_desc(const void *smem) {
  const uint32_t a = sc_smem_u32(smem);
  return (u64)((a >> 4) & 0x3fffu) | ((u64)1 << 16) /* LBO: unused for swizzled K-major */ | ((u64)(1024 >> 4) << 32) /* SBO */ |
         ((u64)1 << 46) /* descriptor version 1 (sm_100) */ | ((u64)2 << 61) /* SWIZZLE_128B */;
}
// instruction descriptor of kind::tf32: D fp32, A / B tf32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24 (cute::UMMA::InstrDescriptor)
__host__ __device__ constexpr uint32_t um_idesc(int n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kUmRows >> 4) << 24); }
__device__ __forceinline__ void um_mma(uint32_t tmem_d, u64 adesc, u64 bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void um_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(sc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void um_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void um_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void um_epilogue_sync() { asm volatile("bar.sync 2, 128;" ::: "memory"); }

// hi / lo split of the query batch of one pass, written as the tensor cores want to read it: hi keeps the sign, exponent and the
// 10 mantissa bits TF32 has, lo = x - hi is exact.  out[chunk][hi | lo][NQ rows][32 cells] — every [NQ][32] block is one K-major
// SWIZZLE_128B operand tile (row = 128 bytes, the 16-byte unit index XOR-ed with row % 8, exactly what a TMA box load with
// CU_TENSOR_MAP_SWIZZLE_128B would leave in shared memory), hi and lo of a chunk back to back: ONE contiguous bulk copy per
// pipeline stage instead of 2 NQ separate 128-byte row segments (the TMA unit moves a box row by row; at ~7 clocks per
// segment the query boxes, not HBM, bounded the first version of this kernel).  Queries >= nqc and cells >= n_cells are zero.
__global__ void sc_split_tf32_tiled_kernel(const float *__restrict__ q, int nqc, int n_cells, int nq_pass, int n_chunks, float *__restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n_chunks * nq_pass * kUmK) return;
  const int k = (int)(i % kUmK), row = (int)((i / kUmK) % nq_pass), c = (int)(i / ((size_t)kUmK * nq_pass));
  const int cell = c * kUmK + k;
  const float v = (row < nqc && cell < n_cells) ? q[(size_t)row * n_cells + cell] : 0.f;
  const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
  const size_t tile = (size_t)nq_pass * kUmK;
  const size_t at = (size_t)c * 2 * tile + (size_t)row * kUmK + (size_t)((((k >> 2) ^ (row & 7)) << 2) | (k & 3));
  out[at] = h;
  out[at + tile] = v - h;
}
template <int NQ>
__global__ void __launch_bounds__(kUmThreads, 1)
    sc_scan_umma_kernel(const __grid_constant__ CUtensorMap map_db, const float *__restrict__ q_tiles /* sc_split_tf32_tiled_kernel's output */,
                        const float *__restrict__ keys, const int *__restrict__ ids, int n_rows, int n_cells, int key_dim,
                        const float *__restrict__ q_keys, int nqc, float ringkey_thres, int max_id, float sc_width, int tiles_per_cta,
                        u64 *__restrict__ scratch, int list_stride, int raw_hi) {
  using C = UmCfg<NQ>;
  constexpr int kUmStages = C::kStages, kUmBBytes = C::kBBytes, kUmStageBytes = C::kStageBytes, kUmAccCols = C::kAccCols, kUmTmemCols = C::kTmemCols;
  extern __shared__ unsigned char smem_raw[];
  unsigned char *base = smem_raw + ((1024u - (sc_smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char *stages = base;
  float *sD = reinterpret_cast<float *>(base + (size_t)kUmStages * kUmStageBytes);
  uint64_t *bars = reinterpret_cast<uint64_t *>(sD + kUmRows * kDistStride);
  uint64_t *full = bars, *split = bars + kUmStages, *empty = bars + 2 * kUmStages, *tfull = bars + 3 * kUmStages, *tempty = tfull + 2;
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(tempty + 2);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_tiles = (n_rows + kUmRows - 1) / kUmRows;
  const int t_begin = blockIdx.x * tiles_per_cta;
  const int t_end = min(t_begin + tiles_per_cta, n_tiles);
  const int n_chunks = (n_cells + kUmK - 1) / kUmK;

  if (tid == 0) {
    for (int s = 0; s < kUmStages; s++) {
      sc_mbar_init(&full[s], 1);
      sc_mbar_init(&split[s], 4);
      sc_mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; a++) {
      sc_mbar_init(&tfull[a], 1);
      sc_mbar_init(&tempty[a], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {  // one warp allocates (and later frees) the tensor memory of this CTA
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sc_smem_u32(s_tmem)), "r"(kUmTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  um_fence_before();
  __syncthreads();
  um_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 0) {
    // ---- TMA producer ----
    if (lane == 0) {
      int it = 0;
      for (int t = t_begin; t < t_end; t++) {
        for (int c = 0; c < n_chunks; c++, it++) {
          const int s = it % kUmStages;
          sc_mbar_wait(&empty[s], ((it / kUmStages) & 1) ^ 1);  // passes immediately on the first lap
          sc_mbar_expect_tx(&full[s], kUmABytes + 2 * kUmBBytes);
          unsigned char *st = stages + (size_t)s * kUmStageBytes;
          sc_tma_load_2d(st, &map_db, c * kUmK, t * kUmRows, &full[s]);
          sc_bulk_load(st + 2 * kUmABytes, q_tiles + (size_t)c * (2 * kUmBBytes / 4), 2 * kUmBBytes, &full[s]);  // [b_hi ; b_lo] of this chunk
        }
      }
    }
  } else if (warp >= 1 && warp <= 4) {
    // ---- splitters: A -> (hi in place, lo next to it) ----
    const int st_tid = tid - 32;
    int it = 0;
    for (int t = t_begin; t < t_end; t++) {
      for (int c = 0; c < n_chunks; c++, it++) {
        const int s = it % kUmStages;
        sc_mbar_wait(&full[s], (it / kUmStages) & 1);
        uint4 *A = reinterpret_cast<uint4 *>(stages + (size_t)s * kUmStageBytes);
        float4 *Alo = reinterpret_cast<float4 *>(stages + (size_t)s * kUmStageBytes + kUmABytes);
        // all loads of the thread first (8 x 16 B in flight: one shared-memory latency per stage, not eight), then the stores
        constexpr int kPer = kUmABytes / 16 / 128;
        uint4 v[kPer];
#pragma unroll
        for (int j = 0; j < kPer; j++) v[j] = A[j * 128 + st_tid];
#pragma unroll
        for (int j = 0; j < kPer; j++) {
          const int i = j * 128 + st_tid;
          uint4 h;
          h.x = v[j].x & 0xffffe000u; h.y = v[j].y & 0xffffe000u; h.z = v[j].z & 0xffffe000u; h.w = v[j].w & 0xffffe000u;
          if (!raw_hi) A[i] = h;
          Alo[i] = make_float4(__uint_as_float(v[j].x) - __uint_as_float(h.x), __uint_as_float(v[j].y) - __uint_as_float(h.y),
                               __uint_as_float(v[j].z) - __uint_as_float(h.z), __uint_as_float(v[j].w) - __uint_as_float(h.w));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor cores' async proxy
        __syncwarp();
        if (lane == 0) sc_mbar_arrive(&split[s]);
      }
    }
  } else if (warp == 5) {
    // ---- MMA issuer (one lane) ----
    if (lane == 0) {
      int it = 0, ti = 0;
      for (int t = t_begin; t < t_end; t++, ti++) {
        const int a = ti & 1;
        sc_mbar_wait(&tempty[a], ((ti >> 1) & 1) ^ 1);  // the epilogue has drained this accumulator (passes at once the first two times)
        um_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(a * kUmAccCols);
        for (int c = 0; c < n_chunks; c++, it++) {
          const int s = it % kUmStages;
          sc_mbar_wait(&split[s], (it / kUmStages) & 1);
          um_fence_after();
          unsigned char *st = stages + (size_t)s * kUmStageBytes;
          // B hi and B lo lie back to back: one N = 2 * NQ operand [b_hi ; b_lo], so a_hi is read once for both of its products
          const u64 a_hi = um_desc(st), a_lo = um_desc(st + kUmABytes), b_both = um_desc(st + 2 * kUmABytes);
#pragma unroll
          for (int k = 0; k < kUmK / 8; k++) {  // UMMA K of tf32 = 8 elements = 32 bytes: the start address advances inside the swizzle atom
            const u64 off = (u64)(k * 32 >> 4);
            um_mma(tmem_d, a_hi + off, b_both + off, um_idesc(2 * NQ), (c | k) != 0 ? 1u : 0u);  // [hi*hi | hi*lo]
            um_mma(tmem_d, a_lo + off, b_both + off, um_idesc(NQ), 1u);                              // + lo*hi onto the first NQ columns
          }
          um_commit(&empty[s]);  // arrives when the MMAs above have read the stage (implies fence::before_thread_sync)
        }
        um_commit(&tfull[a]);
      }
    }
  } else {
    // ---- epilogue + running top-K (warps 6-9; a warp may only touch the TMEM lanes of its quadrant = warp % 4) ----
    // The accumulator is drained in chunks of 32 queries through the [128 rows][32 queries] distance tile; thread (query, row group)
    // keeps one running top-K per chunk.
    constexpr int kChunks = NQ / kQChunk;
    const int quad = warp & 3;
    const int e = tid - 192, q_own = e & 31, sub = e >> 5;
    TopK top[kChunks];
#pragma unroll
    for (int cq = 0; cq < kChunks; cq++) top[cq].init();
    // The approximate distance only ranks candidates for the exact re-score: one multiplication by the reciprocal instead of the
    // IEEE division of the reference formula (the divisions alone kept the XU pipe 36 % busy and this warp-starved role was the
    // critical path of the kernel: profiles/r02_sc_umma_kernel.md).
    const float inv_w = 1.0f / sc_width;
    int ti = 0;
    for (int t = t_begin; t < t_end; t++, ti++) {
      const int a = ti & 1;
      // which of this warp's 32 rows exist and pass the id limit: one coalesced load per tile instead of one dependent load per row and chunk
      const int row0 = t * kUmRows + sub * 32;
      const bool row_ok = row0 + lane < n_rows && __ldg(ids + row0 + lane) < max_id;
      const unsigned vmask = __ballot_sync(0xffffffffu, row_ok);
      sc_mbar_wait(&tfull[a], (ti >> 1) & 1);
      um_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t)(a * kUmAccCols) + ((uint32_t)(quad * 32) << 16);
#define DSLAM_TMEM_LD32(R, ADDR)                                                                                                                      \
  asm volatile(                                                                                                                                       \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, " \
      "%22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                                                                                    \
      : "=r"(R[0]), "=r"(R[1]), "=r"(R[2]), "=r"(R[3]), "=r"(R[4]), "=r"(R[5]), "=r"(R[6]), "=r"(R[7]), "=r"(R[8]), "=r"(R[9]), "=r"(R[10]),        \
        "=r"(R[11]), "=r"(R[12]), "=r"(R[13]), "=r"(R[14]), "=r"(R[15]), "=r"(R[16]), "=r"(R[17]), "=r"(R[18]), "=r"(R[19]), "=r"(R[20]),          \
        "=r"(R[21]), "=r"(R[22]), "=r"(R[23]), "=r"(R[24]), "=r"(R[25]), "=r"(R[26]), "=r"(R[27]), "=r"(R[28]), "=r"(R[29]), "=r"(R[30]),          \
        "=r"(R[31])                                                                                                                                 \
      : "r"(ADDR)                                                                                                                                   \
      : "memory")
#pragma unroll
      for (int cq = 0; cq < kChunks; cq++) {
        uint32_t r[32], r2[32];
        DSLAM_TMEM_LD32(r, taddr + (uint32_t)(cq * kQChunk));
        DSLAM_TMEM_LD32(r2, taddr + (uint32_t)(NQ + cq * kQChunk));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (cq == kChunks - 1) {
          um_fence_before();
          __syncwarp();
          if (lane == 0) sc_mbar_arrive(&tempty[a]);  // the accumulator may be overwritten by the tile after next
        }
        float *drow = sD + (size_t)(quad * 32 + lane) * kDistStride;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 d;
          d.x = (1.0f - (__uint_as_float(r[j]) + __uint_as_float(r2[j])) * inv_w) * 0.5f;
          d.y = (1.0f - (__uint_as_float(r[j + 1]) + __uint_as_float(r2[j + 1])) * inv_w) * 0.5f;
          d.z = (1.0f - (__uint_as_float(r[j + 2]) + __uint_as_float(r2[j + 2])) * inv_w) * 0.5f;
          d.w = (1.0f - (__uint_as_float(r[j + 3]) + __uint_as_float(r2[j + 3])) * inv_w) * 0.5f;
          *reinterpret_cast<float4 *>(drow + j) = d;
        }
        um_epilogue_sync();
        const int q = cq * kQChunk + q_own;
        if (q < nqc) {
          for (int r8 = 0; r8 < 32; r8 += 8) {
            float dv[8];  // independent shared-memory loads first, the serial top-K chain after them
#pragma unroll
            for (int j = 0; j < 8; j++) dv[j] = sD[(size_t)(sub * 32 + r8 + j) * kDistStride + q_own];
#pragma unroll
            for (int j = 0; j < 8; j++) {
              if ((vmask >> (r8 + j)) & 1u) {
                const int row = row0 + r8 + j;
                bool ok = true;
                if (ringkey_thres >= 0.f) ok = flann_l2(q_keys + (size_t)q * key_dim, keys + (size_t)row * key_dim, key_dim) < ringkey_thres;
                if (ok) top[cq].insert(make_key(dv[j], row));
              }
            }
          }
        }
        um_epilogue_sync();
      }
#undef DSLAM_TMEM_LD32
    }
    // CTA merge of the four row groups through the distance tile's storage, one chunk of queries at a time: lists[sub][q][K]
    u64 *lists = reinterpret_cast<u64 *>(sD);
#pragma unroll
    for (int cq = 0; cq < kChunks; cq++) {
#pragma unroll
      for (int i = 0; i < kScTopK; i++) lists[((size_t)sub * kQChunk + q_own) * kScTopK + i] = top[cq].k[i];
      um_epilogue_sync();
      const int q = cq * kQChunk + e;
      if (e < kQChunk && q < nqc) {
        TopK m;
        m.init();
        for (int wv = 0; wv < 4; wv++) {
#pragma unroll
          for (int i = 0; i < kScTopK; i++) m.insert(lists[((size_t)wv * kQChunk + e) * kScTopK + i]);
        }
        u64 *dst = scratch + ((size_t)q * list_stride + blockIdx.x) * kScTopK;
#pragma unroll
        for (int i = 0; i < kScTopK; i++) dst[i] = m.k[i];
      }
      um_epilogue_sync();
    }
  }
  // ---- teardown: every role is done with the tensor memory ----
  um_fence_before();
  __syncthreads();
  if (warp == 5) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kUmTmemCols) : "memory");
}

// ---- exact re-score -------------------------------------------------------------------------------------------
// search_sc's arithmetic (search_place.h:71-79) on dense descriptors: float cur_prod += double(q)*double(d) over the
// cells where both are occupied, in ascending cell order; diff = (1 - cur_prod / sc_width) / 2.0.  The chain of float
// roundings is inherently serial (F2F -> DADD -> F2F per term), so everything else is taken off it: a warp forms the
// double products of 128 cells in parallel (coalesced loads, the next 128 cells already in flight), compacts the
// non-zero ones in cell order into its shared-memory strip (ballot + popc, adding an exact zero is a no-op in the
// reference's chain), and then every lane walks the strip with broadcast LDS.64 — no ballot / ffs / shuffle in the
// dependent loop (that loop was 59 us per query with them).  T = float: fp32 signatures (the product of two floats is
// exact in fp64); T = double: the reference's SigType values (ScanContext.h:24), product rounded to double as in C++.
constexpr int kRescoreChunk = 128;
template <typename T>
__device__ __forceinline__ float exact_sc_diff(const T *__restrict__ q, const T *__restrict__ d, int n_cells, int sc_width, int lane,
                                               double *__restrict__ strip /* [kRescoreChunk], this warp's */) {
  constexpr int NJ = kRescoreChunk / 32;
  float cur = 0.f;
  T qv[NJ], dv[NJ];
#pragma unroll
  for (int j = 0; j < NJ; j++) {
    const int k = 32 * j + lane;
    qv[j] = k < n_cells ? q[k] : (T)0;
    dv[j] = k < n_cells ? __ldg(d + k) : (T)0;
  }
  for (int base = 0; base < n_cells; base += kRescoreChunk) {
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < NJ; j++) {
      const bool nz = qv[j] != (T)0 && dv[j] != (T)0;
      const unsigned m = __ballot_sync(0xffffffffu, nz);
      if (nz) strip[cnt + __popc(m & ((1u << lane) - 1u))] = __dmul_rn((double)qv[j], (double)dv[j]);
      cnt += __popc(m);
    }
#pragma unroll
    for (int j = 0; j < NJ; j++) {  // next chunk: in flight while the chain below runs
      const int k = base + kRescoreChunk + 32 * j + lane;
      qv[j] = k < n_cells ? q[k] : (T)0;
      dv[j] = k < n_cells ? __ldg(d + k) : (T)0;
    }
    __syncwarp();
#pragma unroll 4
    for (int i = 0; i < cnt; i++) cur = __double2float_rn(__dadd_rn((double)cur, strip[i]));
    __syncwarp();
  }
  const float t = 1.0f - cur / (float)sc_width;
  return __double2float_rn((double)t / 2.0);
}

// Cross-GPU exchange of the per-query best keys without a library collective: every rank holds, in its own HBM, a small
// mailbox (kScXchgSlots slots x [keys[kScXchgMaxQ], arrived]) that all peers can reach over NVLink (CUDA IPC mappings).
// A rank min-combines its keys into EVERY rank's mailbox (atom.sys.min.u64 — the packed (distance, id) keys make "min" the
// argmin with ties to the lowest id), fences, and bumps every mailbox's arrival counter; the last CTA of the local grid
// then waits until its own mailbox has heard from all `world` ranks and publishes the combined keys to the host.
// Slot = query sequence number mod kScXchgSlots: a peer can run at most one query ahead of this rank (its next query
// needs this rank's contribution), so a slot is long consumed and reset when it comes round again.
constexpr int kScXchgSlots = 4;
struct ScXchg {
  int world, rank;
  u64 *keys[8];       // keys[r]    : mailbox keys of rank r   [kScXchgSlots][kScXchgMaxQ]   (this rank's mapping)
  unsigned *arrived[8];  // arrived[r]: arrival counters of rank r [kScXchgSlots]
};

// one CTA per query, one warp per top-K entry: exact distance of each survivor, the per-query best as a packed
// (ordered dist bits << 32 | GLOBAL id) key, and — fused — the publication: straight to the host as self-validating
// words (32 payload bits | query sequence number; the host spins on them, no D2H copy, no stream synchronise), after
// the NVLink mailbox exchange above when the database is sharded.
template <typename T>
__global__ void __launch_bounds__(32 * kScTopK) sc_rescore_topk_kernel(const u64 *__restrict__ lists, int nlists, int list_stride,
                                                                       const T *__restrict__ sigs,
                                                                       const int *__restrict__ ids, const T *__restrict__ q_sigs,
                                                                       int n_cells, int sc_width, u64 *__restrict__ exact_keys,
                                                                       u64 *__restrict__ best, u64 *__restrict__ host_words, unsigned seq,
                                                                       const __grid_constant__ ScXchg X, int slot, int q0,
                                                                       unsigned *__restrict__ ticket, unsigned long long timeout_ns) {
  __shared__ u64 sk[kScTopK];
  __shared__ u64 swl[kScTopK][kScTopK];
  __shared__ double strips[kScTopK][kRescoreChunk];
  __shared__ int s_last;
  const int q = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // merge of the scan's per-CTA lists of this query (every query has its own CTA here, so the merge is as parallel as the
  // batch is wide; a merge kernel of its own cost 11 us, a merge in the scan's last CTA serialised the queries)
  {
    TopK t;
    t.init();
    const u64 *src = lists + (size_t)q * list_stride * kScTopK;
    for (int i = threadIdx.x; i < nlists * kScTopK; i += 32 * kScTopK) t.insert(__ldcg(src + i));
    u64 o[kScTopK];
    warp_merge(t, o);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < kScTopK; i++) swl[warp][i] = o[i];
    }
    __syncthreads();
    if (warp == 0) {
      TopK m;
      m.init();
      if (lane < kScTopK) {
#pragma unroll
        for (int i = 0; i < kScTopK; i++) m.insert(swl[lane][i]);
      }
      warp_merge(m, o);
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < kScTopK; i++) sk[i] = o[i];
      }
    }
    __syncthreads();
  }
  const u64 key = sk[warp];
  __syncthreads();  // sk is reused for the exact keys below
  u64 out = kKeyMax;
  if (key != kKeyMax) {
    const int row = (int)(unsigned)(key & 0xffffffffull);
    const float diff = exact_sc_diff<T>(q_sigs + (size_t)q * n_cells, sigs + (size_t)row * n_cells, n_cells, sc_width, lane, strips[warp]);
    out = make_key(diff, ids[row]);
  }
  if (lane == 0) {
    sk[warp] = out;
    if (exact_keys) exact_keys[(size_t)q * kScTopK + warp] = out;
  }
  __syncthreads();
  if (threadIdx.x != 0 && X.world <= 1) return;
  u64 b = kKeyMax;
  if (threadIdx.x == 0) {
    b = sk[0];
#pragma unroll
    for (int i = 1; i < kScTopK; i++) b = sk[i] < b ? sk[i] : b;
    best[q] = b;
  }
  if (X.world <= 1) {
    if (host_words) {
      host_words[2 * (size_t)(q0 + q)] = (b & 0xffffffff00000000ull) | seq;
      host_words[2 * (size_t)(q0 + q) + 1] = (b << 32) | seq;
    }
    return;
  }
  // ---- sharded: NVLink mailbox exchange (slot = exchange sequence number mod kScXchgSlots, the same on every rank) ----
  if (threadIdx.x == 0) {
    for (int r = 0; r < X.world; r++) {
      u64 *dst = X.keys[r] + (size_t)slot * kScXchgMaxQ + q;
      asm volatile("red.relaxed.sys.global.min.u64 [%0], %1;" ::"l"(dst), "l"(b) : "memory");
    }
    __threadfence_system();
    unsigned t;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(t) : "l"(ticket) : "memory");
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  // last CTA of this rank: all local contributions are out -> announce to every mailbox, wait for all ranks, publish
  unsigned *mine = X.arrived[X.rank] + slot;
  if (threadIdx.x == 0) {
    *ticket = 0;
    __threadfence_system();
    for (int r = 0; r < X.world; r++) asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(X.arrived[r] + slot) : "memory");
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    unsigned seen = 0;
    for (;;) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
      if (seen >= (unsigned)X.world) break;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
      if (t1 - t0 > timeout_ns) break;  // a peer never arrived: the host sees the error word instead of a hung GPU
      __nanosleep(200);
    }
    s_last = seen >= (unsigned)X.world ? 1 : 2;
  }
  __syncthreads();
  const bool ok = s_last == 1;
  u64 *mykeys = X.keys[X.rank] + (size_t)slot * kScXchgMaxQ;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
    u64 v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mykeys + i) : "memory");
    if (!ok) v = kScXchgErrorKey;
    best[i] = v;
    if (host_words) {
      host_words[2 * (size_t)(q0 + i)] = (v & 0xffffffff00000000ull) | seq;
      host_words[2 * (size_t)(q0 + i) + 1] = (v << 32) | seq;
    }
    mykeys[i] = kKeyMax;  // reset for the query that reuses this slot (kScXchgSlots queries later)
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    *mine = 0;
  }
}

// explicit (query, row) pairs: diff per pair (row < 0 -> skipped, diff = +inf)
template <typename T>
__global__ void __launch_bounds__(256) sc_rescore_pairs_kernel(const int *__restrict__ pair_q, const int *__restrict__ pair_row, int npairs,
                                                               const T *__restrict__ sigs, const T *__restrict__ q_sigs, int n_cells,
                                                               int sc_width, float *__restrict__ diff_out) {
  __shared__ double strips[8][kRescoreChunk];
  const int lane = threadIdx.x & 31;
  const int p = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (p >= npairs) return;
  const int row = pair_row[p];
  float diff = __int_as_float(0x7f800000);
  if (row >= 0)
    diff = exact_sc_diff<T>(q_sigs + (size_t)pair_q[p] * n_cells, sigs + (size_t)row * n_cells, n_cells, sc_width, lane, strips[threadIdx.x >> 5]);
  if (lane == 0) diff_out[p] = diff;
}

// ---- descriptor generation (ScanContext::generate, src/loop_closure/loop_detection/ScanContext.cpp:19-142) ------
// Stage 1 (one CTA): mean of the cloud, then the 3x3 scatter matrix of the centred points (align_points_PCA :22-41), fp64.
// The sums are block reductions in a fixed order (deterministic; they differ from the reference's sequential sums only
// by the order of the fp64 additions).  out[0..2] = mean, out[3..8] = cov (xx, xy, xz, yy, yz, zz).
__device__ __forceinline__ double block_sum(double v, double *sh) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += sh[w];
  return s;
}

__global__ void __launch_bounds__(1024) sc_moments_kernel(const double *__restrict__ pts, int n, double *__restrict__ out) {
  __shared__ double sh[32];
  double sx = 0, sy = 0, sz = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    sx += pts[3 * i];
    sy += pts[3 * i + 1];
    sz += pts[3 * i + 2];
  }
  const double mx = block_sum(sx, sh) / n, my = block_sum(sy, sh) / n, mz = block_sum(sz, sh) / n;
  double c[6] = {0, 0, 0, 0, 0, 0};
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double x = pts[3 * i] - mx, y = pts[3 * i + 1] - my, z = pts[3 * i + 2] - mz;
    c[0] += x * x; c[1] += x * y; c[2] += x * z; c[3] += y * y; c[4] += y * z; c[5] += z * z;
  }
  double r[6];
#pragma unroll
  for (int k = 0; k < 6; k++) r[k] = block_sum(c[k], sh);
  if (threadIdx.x == 0) {
    out[0] = mx; out[1] = my; out[2] = mz;
#pragma unroll
    for (int k = 0; k < 6; k++) out[3 + k] = r[k];
  }
}

__device__ __forceinline__ u64 order_double(double d) {
  const u64 b = (u64)__double_as_longlong(d);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double unorder_double(u64 k) {
  const u64 b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

struct ScGenParams {
  double mean[3];
  double v[9];  // v[3*j + c]: component c of eigenvector j (ascending eigenvalues): x' = pm . v0 (up), y' = pm . v1, z' = pm . v2
  double lidar_range;
  int num_s, num_r;
};

// Stage 2: rotate into the PCA frame and take the per-cell maximum height (:98-117); cells hold order-preserving keys.
__global__ void __launch_bounds__(256) sc_bin_kernel(const double *__restrict__ pts, int n, const __grid_constant__ ScGenParams P, u64 *__restrict__ cells) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = pts[3 * i] - P.mean[0], y = pts[3 * i + 1] - P.mean[1], z = pts[3 * i + 2] - P.mean[2];
  const double hx = __dadd_rn(__dadd_rn(__dmul_rn(x, P.v[0]), __dmul_rn(y, P.v[1])), __dmul_rn(z, P.v[2]));
  const double yp = __dadd_rn(__dadd_rn(__dmul_rn(x, P.v[3]), __dmul_rn(y, P.v[4])), __dmul_rn(z, P.v[5]));
  const double zp = __dadd_rn(__dadd_rn(__dmul_rn(x, P.v[6]), __dmul_rn(y, P.v[7])), __dmul_rn(z, P.v[8]));
  const double rho = sqrt(__dadd_rn(__dmul_rn(yp, yp), __dmul_rn(zp, zp)));
  double theta = atan2(zp, yp);
  const double two_pi = 2.0 * 3.14159265358979323846;
  while (theta < 0) theta += two_pi;
  while (theta >= two_pi) theta -= two_pi;
  const int si = (int)(theta / two_pi * P.num_s);
  const int ri = (int)(rho / P.lidar_range * P.num_r);
  if (ri >= P.num_r || si >= P.num_s) return;
  atomicMax(cells + si * P.num_r + ri, order_double(hx));
}

// Stage 3 (one CTA): ring key = occupied sectors per ring / num_s, per-sector L2 normalisation (:119-141); dense fp32 out.
__global__ void __launch_bounds__(256) sc_finalize_kernel(const u64 *__restrict__ cells, const __grid_constant__ ScGenParams P, float *__restrict__ ringkey,
                                                         float *__restrict__ sig, double *__restrict__ sig64) {
  __shared__ double norm[256];
  const int tid = threadIdx.x, ns = P.num_s, nr = P.num_r;
  const double empty_below = -P.lidar_range;
  for (int sct = tid; sct < ns; sct += blockDim.x) {
    double acc = 0.0;
    for (int r = 0; r < nr; r++) {  // ascending cell index inside the sector, like the reference's loop over i
      const double h = unorder_double(cells[sct * nr + r]);
      if (h >= empty_below) acc += h * h;
    }
    norm[sct] = sqrt(acc);
  }
  for (int r = tid; r < nr; r += blockDim.x) {
    float cnt = 0.f;
    for (int sct = 0; sct < ns; sct++)
      if (unorder_double(cells[sct * nr + r]) >= empty_below) cnt++;
    ringkey[r] = cnt / ns;
  }
  __syncthreads();
  for (int i = tid; i < ns * nr; i += blockDim.x) {
    const double h = unorder_double(cells[i]);
    const double v = h >= empty_below ? h / norm[i / nr] : 0.0;
    sig[i] = (float)v;
    if (sig64) sig64[i] = v;
  }
}

__global__ void sc_fill_cells_kernel(u64 *cells, int n, double v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) cells[i] = order_double(v);
}

// SM count of the CURRENT device (a process may hold sessions on several devices)
int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cached[dev] = n > 0 ? n : 148;
  }
  return cached[dev];
}

}  // namespace

constexpr int kRingGridX = 64;

// per-CTA top-K lists of every query of a batch: lists[q][sc_list_stride()][K]
int sc_list_stride() { return num_sms() > kRingGridX ? num_sms() : kRingGridX; }
// floats of the split-query scratch for batches of up to `cap` queries: every pass holds at most max(2 nqc, 128) rows of
// ceil(n_cells / 32) * 32 cells, hi and lo
size_t sc_qsplit_floats(int cap, int n_cells) { return (size_t)2 * ((size_t)2 * cap + 128) * (size_t)((n_cells + kUmK - 1) / kUmK) * kUmK; }
size_t sc_scratch_bytes(int nq) { return (size_t)(nq > kQChunk ? nq : kQChunk) * sc_list_stride() * kScTopK * sizeof(u64); }

cudaError_t launch_sc_ringkey(const float *keys, const int *ids, int n_rows, int dim, const float *queries, int nq, int max_id,
                              unsigned long long *out, unsigned long long *scratch, cudaStream_t stream) {
  if (dim > 64 || nq < 1) return cudaErrorInvalidValue;
  int gx = (n_rows + 255) / 256;
  if (gx > kRingGridX) gx = kRingGridX;
  if (gx < 1) gx = 1;
  sc_ringkey_kernel<<<dim3(gx, nq), 256, 0, stream>>>(keys, ids, n_rows, dim, queries, max_id, scratch);
  sc_merge_kernel<<<nq, 32, 0, stream>>>(scratch, gx, out);
  return cudaGetLastError();
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*ScEncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static ScEncodeTiledFn sc_encode_tiled() {
  static ScEncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = (ScEncodeTiledFn)p;
  }
  return fn;
}

// 0 = choose by batch size, 1 = always the streaming kernel, 2 = always the tiled kernel, 3 / 4 = tensor-core kernel (tests / sweeps)
static int g_sc_scan_flavour = -1;
static int sc_scan_flavour() {
  if (g_sc_scan_flavour < 0) {
    const char *e = getenv("DSLAM_SC_SCAN");
    g_sc_scan_flavour = 0;
    if (e && !strcmp(e, "stream")) g_sc_scan_flavour = 1;
    if (e && !strcmp(e, "tile")) g_sc_scan_flavour = 2;
    if (e && !strcmp(e, "umma")) g_sc_scan_flavour = 3;
    if (e && !strcmp(e, "umma_masked")) g_sc_scan_flavour = 4;
  }
  return g_sc_scan_flavour;
}
void sc_set_scan_flavour(int f) { g_sc_scan_flavour = f; }

static cudaError_t launch_sc_scan_tiles(const float *sigs, const float *keys, const int *ids, int n_rows, int n_cells, int key_dim, const float *q_sigs,
                                        const float *q_keys, int nqc, float ringkey_thres, int max_id, float sc_width, unsigned long long *scratch,
                                        int list_stride, int grid, cudaStream_t stream) {
  ScEncodeTiledFn enc = sc_encode_tiled();
  if (!enc) return cudaErrorNotSupported;
  {  // per device and cheap: set on every launch (a process may drive several devices)
    cudaError_t e = cudaFuncSetAttribute(sc_scan_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTileSmemBytes);
    if (e != cudaSuccess) return e;
  }
  CUtensorMap map_db, map_q;
  const cuuint32_t estr[2] = {1, 1};
  const cuuint64_t gstride[1] = {(cuuint64_t)n_cells * sizeof(float)};
  {
    const cuuint64_t gdim[2] = {(cuuint64_t)n_cells, (cuuint64_t)n_rows};
    const cuuint32_t box[2] = {(cuuint32_t)kTileK, (cuuint32_t)kHalfRows};
    if (enc(&map_db, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(sigs), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return cudaErrorInvalidValue;
  }
  {
    const cuuint64_t gdim[2] = {(cuuint64_t)n_cells, (cuuint64_t)nqc};
    const cuuint32_t box[2] = {(cuuint32_t)kTileK, (cuuint32_t)kQChunk};
    if (enc(&map_q, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(q_sigs), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return cudaErrorInvalidValue;
  }
  const int n_halves = (n_rows + kHalfRows - 1) / kHalfRows;
  const int halves_per_cta = (n_halves + grid - 1) / grid;
  sc_scan_tile_kernel<<<grid, kTileThreads, kTileSmemBytes, stream>>>(map_db, map_q, keys, ids, n_rows, n_cells, key_dim, q_keys, nqc, ringkey_thres,
                                                                      max_id, sc_width, halves_per_cta, scratch, list_stride);
  return cudaGetLastError();
}

// tensor-core flavour: q_tiles = this pass's slice of the split-query scratch (sc_qsplit_floats)
template <int NQ>
static cudaError_t launch_sc_scan_umma(const float *sigs, const float *keys, const int *ids, int n_rows, int n_cells, int key_dim, const float *q_sigs,
                                       const float *q_keys, int nqc, float ringkey_thres, int max_id, float sc_width, unsigned long long *scratch,
                                       int list_stride, int grid, float *q_tiles, int raw_hi, cudaStream_t stream) {
  ScEncodeTiledFn enc = sc_encode_tiled();
  if (!enc) return cudaErrorNotSupported;
  cudaError_t e = cudaFuncSetAttribute(sc_scan_umma_kernel<NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UmCfg<NQ>::kSmemBytes);
  if (e != cudaSuccess) return e;
  const int n_chunks = (n_cells + kUmK - 1) / kUmK;
  const size_t nq_elems = (size_t)n_chunks * NQ * kUmK;
  sc_split_tf32_tiled_kernel<<<(unsigned)((nq_elems + 255) / 256), 256, 0, stream>>>(q_sigs, nqc, n_cells, NQ, n_chunks, q_tiles);
  CUtensorMap map_db;
  const cuuint32_t estr[2] = {1, 1};
  const cuuint64_t gstride[1] = {(cuuint64_t)n_cells * sizeof(float)};
  {
    const cuuint64_t gdim[2] = {(cuuint64_t)n_cells, (cuuint64_t)n_rows};
    const cuuint32_t box[2] = {(cuuint32_t)kUmK, (cuuint32_t)kUmRows};
    if (enc(&map_db, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(sigs), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return cudaErrorInvalidValue;
  }
  const int n_tiles = (n_rows + kUmRows - 1) / kUmRows;
  const int tiles_per_cta = (n_tiles + grid - 1) / grid;
  sc_scan_umma_kernel<NQ><<<grid, kUmThreads, UmCfg<NQ>::kSmemBytes, stream>>>(map_db, q_tiles, keys, ids, n_rows, n_cells, key_dim, q_keys, nqc,
                                                                               ringkey_thres, max_id, sc_width, tiles_per_cta, scratch, list_stride, raw_hi);
  return cudaGetLastError();
}

// queries per pass of the tensor-core scan: 0 = by batch size (DSLAM_SC_UMMA_NQ=32|64|128 pins it: tests / sweeps)
static int sc_umma_nq_override() {
  static const int v = [] { const char *e = getenv("DSLAM_SC_UMMA_NQ"); const int x = e ? atoi(e) : 0; return (x == 32 || x == 64 || x == 128) ? x : 0; }();
  return v;
}

cudaError_t launch_sc_scan(const float *sigs, const float *keys, const int *ids, int n_rows, int n_cells, int key_dim, const float *q_sigs,
                           const float *q_keys, int nq, float ringkey_thres, int max_id, float sc_width, unsigned long long *scratch, int *nlists_out,
                           float *q_split, cudaStream_t stream) {
  if (n_cells % 4 != 0 || n_cells > 1280 || key_dim > 64) return cudaErrorInvalidValue;
  {  // worst case of any descriptor shape this library accepts; per device and cheap, so set on every launch
    cudaError_t e = cudaFuncSetAttribute(sc_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kQChunk * 1280 * 4 + kQChunk * 64 * 4);
    if (e != cudaSuccess) return e;
  }
  const int stride = sc_list_stride();
  const int n_tiles = (n_rows + kTileRows - 1) / kTileRows;
  // one grid size for all chunks of the batch (the re-score kernel merges `grid` lists per query)
  const int flavour = sc_scan_flavour();
  const int nq_first = nq < kQChunk ? nq : kQChunk;
  // The streaming kernel is HBM-bound up to ~4 queries and costs ~18 us per further query and 100k rows; the tiled
  // kernel costs the same for 1..32 queries (FFMA-bound) and wants at least one tile per two SMs.
  // (a shard smaller than one 64-row TMA box always takes the streaming kernel)
  // tensor-core flavour (tcgen05, 3xTF32): query batches over shards of >= 4096 rows (DSLAM_SC_UMMA_MIN_ROWS).  Measured against the
  // FFMA tile kernel at Q = 32: 100k rows 0.128 vs 0.220 ms, 50k 0.077 vs 0.134, 25k 0.061 vs 0.076, 12.5k 0.042 vs 0.059, 6.25k 0.042 vs 0.054
  static const int umma_min_rows = [] { const char *e = getenv("DSLAM_SC_UMMA_MIN_ROWS"); const int v = e ? atoi(e) : 0; return v >= 128 ? v : 4096; }();
  const bool umma = q_split != nullptr && n_rows >= umma_min_rows && (flavour == 3 || flavour == 4 || (flavour == 0 && nq_first > 8));
  const int raw_hi = flavour != 4;  // 4: the splitters store the masked hi operand explicitly
  const bool tiles = !umma && n_rows >= 64 && (flavour == 2 || flavour == 3 || flavour == 4 || (flavour == 0 && nq_first > 8 && n_tiles * 2 >= num_sms()));
  int grid = num_sms();
  if (umma) {
    const int n_t = (n_rows + kUmRows - 1) / kUmRows;
    const int tpc = (n_t + grid - 1) / grid;
    grid = (n_t + tpc - 1) / tpc;
  } else if (tiles) {
    const int n_halves = (n_rows + 31) / 32;
    if (grid > n_halves) grid = n_halves;
    const int hpc = (n_halves + grid - 1) / grid;
    grid = (n_halves + hpc - 1) / hpc;  // no CTA without rows
  }
  if (nlists_out) *nlists_out = grid;
  size_t split_at = 0;  // floats of q_split used by the passes so far
  for (int q0 = 0; q0 < nq;) {
    int pass = kQChunk;  // queries of this pass
    if (umma) {
      pass = sc_umma_nq_override();
      if (pass == 0) pass = nq - q0 > 64 ? 128 : (nq - q0 > 32 ? 64 : 32);
    }
    const int nqc = nq - q0 < pass ? nq - q0 : pass;
    unsigned long long *lists = scratch + (size_t)q0 * stride * kScTopK;
    if (umma) {
      float *q_tiles = q_split + split_at;
      split_at += (size_t)2 * pass * ((n_cells + kUmK - 1) / kUmK) * kUmK;
      const float *qs = q_sigs + (size_t)q0 * n_cells, *qk = q_keys + (size_t)q0 * key_dim;
      cudaError_t e;
      if (pass == 128)
        e = launch_sc_scan_umma<128>(sigs, keys, ids, n_rows, n_cells, key_dim, qs, qk, nqc, ringkey_thres, max_id, sc_width, lists, stride, grid, q_tiles, raw_hi, stream);
      else if (pass == 64)
        e = launch_sc_scan_umma<64>(sigs, keys, ids, n_rows, n_cells, key_dim, qs, qk, nqc, ringkey_thres, max_id, sc_width, lists, stride, grid, q_tiles, raw_hi, stream);
      else
        e = launch_sc_scan_umma<32>(sigs, keys, ids, n_rows, n_cells, key_dim, qs, qk, nqc, ringkey_thres, max_id, sc_width, lists, stride, grid, q_tiles, raw_hi, stream);
      if (e != cudaSuccess) return e;
    } else if (tiles) {
      cudaError_t e = launch_sc_scan_tiles(sigs, keys, ids, n_rows, n_cells, key_dim, q_sigs + (size_t)q0 * n_cells, q_keys + (size_t)q0 * key_dim, nqc,
                                           ringkey_thres, max_id, sc_width, lists, stride, grid, stream);
      if (e != cudaSuccess) return e;
    } else {
      size_t smem = (size_t)nqc * n_cells * 4 + (size_t)nqc * key_dim * 4;
      const size_t lists_bytes = (size_t)kScanWarps * kQChunk * kScTopK * sizeof(u64);
      if (smem < lists_bytes) smem = lists_bytes;
      sc_scan_kernel<<<grid, kScanThreads, smem, stream>>>(sigs, keys, ids, n_rows, n_cells, key_dim, q_sigs + (size_t)q0 * n_cells,
                                                           q_keys + (size_t)q0 * key_dim, nqc, ringkey_thres, max_id, sc_width, lists, stride);
    }
    q0 += nqc;
  }
  return cudaGetLastError();
}

cudaError_t launch_sc_moments(const double *pts, int n, double *out9, cudaStream_t stream) {
  sc_moments_kernel<<<1, 1024, 0, stream>>>(pts, n, out9);
  return cudaGetLastError();
}

cudaError_t launch_sc_bin_finalize(const double *pts, int n, const double mean[3], const double v9[9], double lidar_range, int num_s, int num_r,
                                   unsigned long long *cells, float *ringkey, float *sig, double *sig64, cudaStream_t stream) {
  if (num_s > 256) return cudaErrorInvalidValue;
  ScGenParams P;
  for (int i = 0; i < 3; i++) P.mean[i] = mean[i];
  for (int i = 0; i < 9; i++) P.v[i] = v9[i];
  P.lidar_range = lidar_range;
  P.num_s = num_s;
  P.num_r = num_r;
  const int nc = num_s * num_r;
  sc_fill_cells_kernel<<<(nc + 255) / 256, 256, 0, stream>>>(cells, nc, -lidar_range - 1.0);
  if (n > 0) sc_bin_kernel<<<(n + 255) / 256, 256, 0, stream>>>(pts, n, P, cells);
  sc_finalize_kernel<<<1, 256, 0, stream>>>(cells, P, ringkey, sig, sig64);
  return cudaGetLastError();
}

cudaError_t launch_sc_rescore_topk(const unsigned long long *lists, int nlists, const void *sigs, int fp64, const int *ids, const void *q_sigs, int nq,
                                   int n_cells,
                                   int sc_width, unsigned long long *exact_keys, unsigned long long *best, unsigned long long *host_words,
                                   unsigned seq, const ScExchange *xchg, unsigned xchg_seq, int q0, unsigned *ticket, cudaStream_t stream) {
  if (nq < 1) return cudaSuccess;
  const int slot = (int)(xchg_seq % kScXchgSlots);
  ScXchg X;
  std::memset(&X, 0, sizeof(X));
  X.world = 1;
  if (xchg && xchg->world > 1) {
    if (xchg->world > 8 || nq > kScXchgMaxQ) return cudaErrorInvalidValue;
    X.world = xchg->world;
    X.rank = xchg->rank;
    for (int r = 0; r < xchg->world; r++) {
      X.keys[r] = xchg->keys[r];
      X.arrived[r] = xchg->arrived[r];
    }
  }
  const unsigned long long timeout_ns = 2000000000ull;
  if (fp64)
    sc_rescore_topk_kernel<double><<<nq, 32 * kScTopK, 0, stream>>>(lists, nlists, sc_list_stride(), (const double *)sigs, ids, (const double *)q_sigs, n_cells, sc_width,
                                                                    exact_keys, best, host_words, seq, X, slot, q0, ticket, timeout_ns);
  else
    sc_rescore_topk_kernel<float><<<nq, 32 * kScTopK, 0, stream>>>(lists, nlists, sc_list_stride(), (const float *)sigs, ids, (const float *)q_sigs, n_cells, sc_width, exact_keys,
                                                                   best, host_words, seq, X, slot, q0, ticket, timeout_ns);
  return cudaGetLastError();
}

cudaError_t launch_sc_rescore_pairs(const int *pair_q, const int *pair_row, int npairs, const void *sigs, int fp64, const void *q_sigs, int n_cells,
                                    int sc_width, float *diff_out, cudaStream_t stream) {
  if (npairs < 1) return cudaSuccess;
  if (fp64)
    sc_rescore_pairs_kernel<double><<<(npairs + 7) / 8, 256, 0, stream>>>(pair_q, pair_row, npairs, (const double *)sigs, (const double *)q_sigs, n_cells,
                                                                          sc_width, diff_out);
  else
    sc_rescore_pairs_kernel<float><<<(npairs + 7) / 8, 256, 0, stream>>>(pair_q, pair_row, npairs, (const float *)sigs, (const float *)q_sigs, n_cells,
                                                                         sc_width, diff_out);
  return cudaGetLastError();
}

size_t sc_exchange_arrived_offset() { return (size_t)kScXchgSlots * kScXchgMaxQ * sizeof(u64); }
size_t sc_exchange_bytes() { return sc_exchange_arrived_offset() + 256; }

// widen fp32 signatures to the fp64 side table (rows appended through the fp32 entry points of an fp64 database)
__global__ void sc_widen_kernel(const float *__restrict__ src, double *__restrict__ dst, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (double)src[i];
}
__global__ void sc_narrow_kernel(const double *__restrict__ src, float *__restrict__ dst, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (float)src[i];
}
cudaError_t launch_sc_widen(const float *src, double *dst, size_t n, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  sc_widen_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(src, dst, n);
  return cudaGetLastError();
}
cudaError_t launch_sc_narrow(const double *src, float *dst, size_t n, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  sc_narrow_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(src, dst, n);
  return cudaGetLastError();
}

}  // namespace dslam
