// kernels_template.cu — the tracking template of a keyframe built on the device.  sm_100a.
//
// Replaces TrackerAndScaler::makeCoarseDepthL0  src/scale_optimization/TrackerAndScaler.cpp:143-315:
//   1. scatter   idepth[0][u + w*v] += idepth*weight, weightSums[0][..] += weight over the active points (:149-166)
//   2. pool      level l = sum of the 2x2 block of level l-1, ((a + b) + c) + d  (:168-189)
//   3. dilate    empty cells take the mean of their occupied diagonal neighbours (levels 0, 1: :191-236) or
//                4-neighbours (levels >= 2: :238-279); reads only cells that were occupied before the pass
//   4. compact   raster scan of the interior [2, w-2) x [2, h-2): idepth /= weightSum, keep (x, y, idepth, color)
//                when the colour is finite and idepth > 0 (:281-314) — written in raster order so that the
//                point order (and with it the "every 32nd point" flow-indicator sampling) matches the host loop.
// Arithmetic is the reference's fp32 in the reference's order (-fmad=false).  The only order freedom is step 1
// when three or more points fall on one pixel (float atomics commute for two).

#include "dslam_kernels.h"

namespace dslam {

namespace {

__global__ void tmpl_scatter_kernel(const int *__restrict__ pu, const int *__restrict__ pv, const float *__restrict__ pid,
                                    const float *__restrict__ pw, int npts, int w, int h, float *__restrict__ idepth, float *__restrict__ wsum) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npts) return;
  const int u = pu[i], v = pv[i];
  if (u < 0 || v < 0 || u >= w || v >= h) return;
  atomicAdd(idepth + u + w * v, pid[i] * pw[i]);
  atomicAdd(wsum + u + w * v, pw[i]);
}

__global__ void tmpl_pool_kernel(const float *__restrict__ id_lm, const float *__restrict__ ws_lm, int wlm1, float *__restrict__ id_l,
                                 float *__restrict__ ws_l, int wl, int hl) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= wl || y >= hl) return;
  const int bidx = 2 * x + 2 * y * wlm1;
  id_l[x + y * wl] = id_lm[bidx] + id_lm[bidx + 1] + id_lm[bidx + wlm1] + id_lm[bidx + wlm1 + 1];
  ws_l[x + y * wl] = ws_lm[bidx] + ws_lm[bidx + 1] + ws_lm[bidx + wlm1] + ws_lm[bidx + wlm1 + 1];
}

// all levels in one grid; writes the dilated weight sums into ws_out (the input plays weightSumsl_bak)
__global__ void tmpl_dilate_kernel(const __grid_constant__ TemplateGrids G) {
  const int lvl = blockIdx.y;
  if (lvl >= G.levels) return;
  const int wl = G.w[lvl], hl = G.h[lvl], n = wl * hl;
  float *__restrict__ idl = G.idepth[lvl];
  const float *__restrict__ bak = G.wsum[lvl];
  float *__restrict__ wout = G.wsum2[lvl];
  const int o0 = lvl < 2 ? 1 + wl : 1, o1 = lvl < 2 ? -1 - wl : -1, o2 = lvl < 2 ? wl - 1 : wl, o3 = lvl < 2 ? -wl + 1 : -wl;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float wv = bak[i];
    if (i >= wl && i < n - wl && wv <= 0) {
      float sum = 0, num = 0, numn = 0;
      // the reference reads one element before / after the grid at the two corner pixels (x = 0, y = 1) and
      // (x = w-1, y = h-2) (:207, :203); those pixels lie outside the interior the template is taken from, so an
      // out-of-range neighbour simply counts as empty here
      const int j0 = i + o0, j1 = i + o1, j2 = i + o2, j3 = i + o3;
      if (j0 < n && bak[j0] > 0) { sum += idl[j0]; num += bak[j0]; numn++; }
      if (j1 >= 0 && bak[j1] > 0) { sum += idl[j1]; num += bak[j1]; numn++; }
      if (j2 < n && bak[j2] > 0) { sum += idl[j2]; num += bak[j2]; numn++; }
      if (j3 >= 0 && bak[j3] > 0) { sum += idl[j3]; num += bak[j3]; numn++; }
      if (numn > 0) {
        idl[i] = sum / numn;
        wv = num / numn;
      }
    }
    wout[i] = wv;
  }
}

constexpr int kCompactBlock = 1024;

__device__ __forceinline__ bool tmpl_valid(const TemplateGrids &G, int lvl, int j, float4 *rec) {
  // j indexes the interior raster: (wl-4) columns x (hl-4) rows
  const int wl = G.w[lvl], iw = wl - 4;
  const int x = 2 + j % iw, y = 2 + j / iw;
  const int i = x + y * wl;
  const float ws = G.wsum2[lvl][i];
  if (!(ws > 0)) return false;
  const float id = G.idepth[lvl][i] / ws;
  const float color = G.tex[lvl][i].x;
  if (!isfinite(color) || !(id > 0)) return false;
  *rec = make_float4((float)x, (float)y, id, color);
  return true;
}

__device__ __forceinline__ int block_level(const TemplateGrids &G, int b, int *first) {
  int lvl = 0;
#pragma unroll
  for (int l = 1; l < kMaxLevels; l++)
    if (l < G.levels && b >= G.cblock_begin[l]) lvl = l;
  *first = G.cblock_begin[lvl];
  return lvl;
}

// phase a: count of kept pixels per block of 1024 interior pixels
__global__ void __launch_bounds__(kCompactBlock) tmpl_count_kernel(const __grid_constant__ TemplateGrids G, int *__restrict__ block_counts) {
  int first;
  const int lvl = block_level(G, blockIdx.x, &first);
  const int interior = (G.w[lvl] - 4) * (G.h[lvl] - 4);
  const int j = (blockIdx.x - first) * kCompactBlock + threadIdx.x;
  float4 rec;
  const bool keep = j < interior && tmpl_valid(G, lvl, j, &rec);
  const int c = __syncthreads_count(keep);
  if (threadIdx.x == 0) block_counts[blockIdx.x] = c;
}

// phase b: exclusive scan of the block counts inside every level (one CTA; <= a few thousand blocks)
__global__ void __launch_bounds__(1024) tmpl_scan_kernel(const __grid_constant__ TemplateGrids G, int *__restrict__ block_counts,
                                                        int *__restrict__ block_offsets, int *__restrict__ pc_n) {
  __shared__ int warp_sums[32];
  __shared__ int carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int lvl = 0; lvl < G.levels; lvl++) {
    const int b0 = G.cblock_begin[lvl], b1 = G.cblock_begin[lvl + 1];
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = b0; base < b1; base += 1024) {
      const int b = base + tid;
      const int v = b < b1 ? block_counts[b] : 0;
      int inc = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
      }
      if (lane == 31) warp_sums[warp] = inc;
      __syncthreads();
      if (warp == 0) {
        int ws = warp_sums[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, ws, d);
          if (lane >= d) ws += t;
        }
        warp_sums[lane] = ws;
      }
      __syncthreads();
      const int prefix = carry + (warp > 0 ? warp_sums[warp - 1] : 0) + inc - v;
      if (b < b1) block_offsets[b] = prefix;
      __syncthreads();
      if (tid == 1023) carry = prefix + v;
      __syncthreads();
    }
    if (tid == 0) pc_n[lvl] = carry;
    __syncthreads();
  }
}

// phase c: ordered write
__global__ void __launch_bounds__(kCompactBlock) tmpl_write_kernel(const __grid_constant__ TemplateGrids G, const int *__restrict__ block_offsets) {
  __shared__ int warp_sums[32];
  int first;
  const int lvl = block_level(G, blockIdx.x, &first);
  const int interior = (G.w[lvl] - 4) * (G.h[lvl] - 4);
  const int j = (blockIdx.x - first) * kCompactBlock + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 rec = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool keep = j < interior && tmpl_valid(G, lvl, j, &rec);
  const unsigned m = __ballot_sync(0xffffffffu, keep);
  if (lane == 0) warp_sums[warp] = __popc(m);
  __syncthreads();
  if (warp == 0) {
    int ws = warp_sums[lane];
    const int v = ws;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, ws, d);
      if (lane >= d) ws += t;
    }
    warp_sums[lane] = ws - v;  // exclusive
  }
  __syncthreads();
  if (keep) {
    const int pos = block_offsets[blockIdx.x] + warp_sums[warp] + __popc(m & ((1u << lane) - 1u));
    G.out[lvl][pos] = rec;
  }
}

}  // namespace

cudaError_t launch_template_build(const TemplateGrids &G, const int *pu, const int *pv, const float *pid, const float *pw, int npts,
                                  int *block_counts, int *block_offsets, int *pc_n_dev, int *launches, cudaStream_t stream) {
  cudaError_t e;
  const size_t n0 = (size_t)G.w[0] * G.h[0];
  if ((e = cudaMemsetAsync(G.idepth[0], 0, n0 * sizeof(float), stream)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(G.wsum[0], 0, n0 * sizeof(float), stream)) != cudaSuccess) return e;
  int nl = 0;
  if (npts > 0) {
    tmpl_scatter_kernel<<<(npts + 255) / 256, 256, 0, stream>>>(pu, pv, pid, pw, npts, G.w[0], G.h[0], G.idepth[0], G.wsum[0]);
    nl++;
  }
  for (int l = 1; l < G.levels; l++) {
    dim3 grid((G.w[l] + 127) / 128, G.h[l]);
    tmpl_pool_kernel<<<grid, 128, 0, stream>>>(G.idepth[l - 1], G.wsum[l - 1], G.w[l - 1], G.idepth[l], G.wsum[l], G.w[l], G.h[l]);
    nl++;
  }
  {
    dim3 grid(148 * 2, G.levels);
    tmpl_dilate_kernel<<<grid, 256, 0, stream>>>(G);
    nl++;
  }
  const int nblocks = G.cblock_begin[G.levels];
  if (nblocks > 0) {
    tmpl_count_kernel<<<nblocks, kCompactBlock, 0, stream>>>(G, block_counts);
    tmpl_scan_kernel<<<1, 1024, 0, stream>>>(G, block_counts, block_offsets, pc_n_dev);
    tmpl_write_kernel<<<nblocks, kCompactBlock, 0, stream>>>(G, block_offsets);
    nl += 3;
  } else {
    if ((e = cudaMemsetAsync(pc_n_dev, 0, sizeof(int) * kMaxLevels, stream)) != cudaSuccess) return e;
  }
  if (launches) *launches = nl;
  return cudaGetLastError();
}

int template_compact_blocks(int w, int h) {
  if (w <= 4 || h <= 4) return 0;
  return ((w - 4) * (h - 4) + kCompactBlock - 1) / kCompactBlock;
}

}  // namespace dslam
