// kernels_pyramid.cu — image pyramids of a batch of frames on the device.  sm_100a.
//
// Replaces FrameHessian::makeImages  deps:dso/src/FullSystem/HessianBlocks.cpp:128-191:
//   level 0 channel 0 = input; level k channel 0 = 0.25f * (((a + b) + c) + d) of the 2x2 block of level k-1
//   (:162-165, same operand order); for linear idx in [w, w*(h-1)): dx = 0.5f*(I[idx+1]-I[idx-1]),
//   dy = 0.5f*(I[idx+w]-I[idx-w]), non-finite -> 0 (:169-179) — note the LINEAR index: at x = 0 / x = w-1 the
//   horizontal neighbour wraps to the previous / next row, and that is reproduced here;
//   absSquaredGrad = dx*dx + dy*dy, times gw*gw with gw = B[c+1]-B[c], c = clamp((int)(I+0.5f), 5, 250)
//   (:182-188, getBGradOnly deps:dso/src/FullSystem/HessianBlocks.h:384-390) when a B table is given.
//   Rows 0 and h-1 (uninitialised in the reference) are written as dx = dy = absSquaredGrad = 0.
//
// Two launches for ALL frames of a batch (blockIdx.z / blockIdx.y = frame), both bit-exact (compiled -fmad=false;
// only adds, multiplies by 0.25f/0.5f and the squared-gradient sum are involved):
//   A  downsample_chain_kernel : one CTA per 64x32 level-0 tile builds the intensity of levels 1..L-1 through
//      shared memory (float2 coalesced loads, every level written once);
//   B  gradient_kernel         : one CTA per 64x16 tile of ANY level (all levels in one grid); the (64+8)x(16+2)
//      halo box of the intensity plane is staged into shared memory by TMA (cp.async.bulk.tensor.2d, zero
//      fill outside the image, completion on an mbarrier; the box starts at x0-4 because the TMA source address
//      must be 16-byte aligned); each thread emits one float4 texel (I, dx, dy, absSquaredGrad) per pixel —
//      16-B coalesced stores — plus, when the host wants the reference's layouts, the Vector3f AoS / float plane
//      staging copies that are then DMA'd to the host.
// Per-frame pointers and tensor maps live in a device-resident FrameDev record; the launch carries only pointers to them.
// Algorithmic traffic: read 4*P0, write 16*sum(P_l) (+16*sum(P_l) staging when host copies are requested).

#include "dslam_kernels.h"

namespace dslam {

namespace {

// ---------------------------------------------------------------------------------------------------
// A: box-mean chain
// ---------------------------------------------------------------------------------------------------
template <int TW, int TH>  // tile of the level being produced; source in shared memory has size 2TW x 2TH
__device__ __forceinline__ void down_from_smem(const float *src, float *dst_s, float *dst_g, int pitch, int x0, int y0, int w, int h,
                                               int tid) {
  for (int p = tid; p < TW * TH; p += 256) {
    const int lx = p % TW, ly = p / TW;
    const float a = src[(2 * ly) * (2 * TW) + 2 * lx], b = src[(2 * ly) * (2 * TW) + 2 * lx + 1];
    const float c = src[(2 * ly + 1) * (2 * TW) + 2 * lx], d = src[(2 * ly + 1) * (2 * TW) + 2 * lx + 1];
    const float v = 0.25f * (a + b + c + d);
    dst_s[ly * TW + lx] = v;
    const int x = x0 + lx, y = y0 + ly;
    if (x < w && y < h) dst_g[(size_t)y * pitch + x] = v;
  }
}

__global__ void __launch_bounds__(256) downsample_chain_kernel(const __grid_constant__ FrameBatch B) {
  __shared__ float s1[32 * 16], s2[16 * 8], s3[8 * 4], s4[4 * 2], s5[2];
  __shared__ float *s_plane[kMaxLevels];
  const int tid = threadIdx.x;
  const PyramidGeom &P = B.G;
  if (tid < kMaxLevels) s_plane[tid] = B.f[blockIdx.z]->plane[tid];
  __syncthreads();
  const int X0 = blockIdx.x * kDownTileW, Y0 = blockIdx.y * kDownTileH;  // level-0 origin of the tile
  // level 1 from global level 0 (float2 loads; pitch is a multiple of 4 floats, x even -> 8-B aligned)
  {
    const float *src = s_plane[0];
    float *dst = s_plane[1];
    const int pitch0 = P.pitch[0], w0 = P.w[0], h0 = P.h[0];
    const int w1 = P.w[1], h1 = P.h[1], x10 = X0 >> 1, y10 = Y0 >> 1;
    for (int p = tid; p < 32 * 16; p += 256) {
      const int lx = p % 32, ly = p / 32;
      const int x = x10 + lx, y = y10 + ly;
      float v = 0.f;
      if (2 * x + 1 < w0 && 2 * y + 1 < h0) {
        const float2 r0 = __ldg(reinterpret_cast<const float2 *>(src + (size_t)(2 * y) * pitch0 + 2 * x));
        const float2 r1 = __ldg(reinterpret_cast<const float2 *>(src + (size_t)(2 * y + 1) * pitch0 + 2 * x));
        v = 0.25f * (r0.x + r0.y + r1.x + r1.y);
        if (x < w1 && y < h1) dst[(size_t)y * P.pitch[1] + x] = v;
      }
      s1[p] = v;
    }
  }
  if (P.levels <= 2) return;
  __syncthreads();
  down_from_smem<16, 8>(s1, s2, s_plane[2], P.pitch[2], X0 >> 2, Y0 >> 2, P.w[2], P.h[2], tid);
  if (P.levels <= 3) return;
  __syncthreads();
  down_from_smem<8, 4>(s2, s3, s_plane[3], P.pitch[3], X0 >> 3, Y0 >> 3, P.w[3], P.h[3], tid);
  if (P.levels <= 4) return;
  __syncthreads();
  down_from_smem<4, 2>(s3, s4, s_plane[4], P.pitch[4], X0 >> 4, Y0 >> 4, P.w[4], P.h[4], tid);
  if (P.levels <= 5) return;
  __syncthreads();
  down_from_smem<2, 1>(s4, s5, s_plane[5], P.pitch[5], X0 >> 5, Y0 >> 5, P.w[5], P.h[5], tid);
}

// ---------------------------------------------------------------------------------------------------
// B: gradients + texel packing, TMA-staged halo tiles
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(phase)
      : "memory");
}
// The tensor map lives in global memory (written by the host before the launch): make it visible to the
// tensor-map proxy of this SM before the first use.
__device__ __forceinline__ void tensormap_acquire(const CUtensorMap *map) {
  asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(dst)),
               "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}

// Persistent CTAs with a two-stage TMA pipeline: while the threads turn the halo box of tile i into texels, the box of
// tile i+1 (possibly of another level / frame) is already in flight.  A tile that lives alone in its CTA pays the
// descriptor fetch, the TMA round trip and the pointer loads serially (~2 us for 16 KB of output) — that, not
// bandwidth, bounded the one-tile-per-CTA version at 2.8 TB/s.
struct TileInfo {
  float4 *tex;
  float *hd, *ha;
  const float *plane, *Bt;
  int x0, y0, w, h, pitch;
};

__global__ void __launch_bounds__(256) gradient_kernel(const __grid_constant__ FrameBatch B, int nframes) {
  constexpr int kBoxStride = (kGradBoxH * kGradBoxW + 31) / 32 * 32;  // every stage 128-B aligned (TMA destination)
  __shared__ __align__(128) float box[2][kBoxStride];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ TileInfo info[2];
  // host-layout staging of one tile (Vector3f AoS: 12 B per pixel, 768 B per 64-pixel row): collected here and written out as
  // 16-B vectors — three scalar stores per pixel at a 12-B stride cost three times the store instructions
  __shared__ __align__(16) float sh3[kGradTileH][3 * kGradTileW];
  const int tid = threadIdx.x;
  const PyramidGeom &G = B.G;
  const int tiles_per_frame = G.tile_begin[G.levels];
  const int total = tiles_per_frame * nframes;

  auto issue = [&](int g, int stage) {  // thread 0 only
    const int frame = g / tiles_per_frame, tix = g - frame * tiles_per_frame;
    int lvl = 0;
#pragma unroll
    for (int l = 1; l < kMaxLevels; l++)
      if (l < G.levels && tix >= G.tile_begin[l]) lvl = l;
    const int t = tix - G.tile_begin[lvl];
    const int tx = t % G.tiles_x[lvl], ty = t / G.tiles_x[lvl];
    const FrameDev *F = B.f[frame];
    TileInfo ti;
    ti.x0 = tx * kGradTileW; ti.y0 = ty * kGradTileH;
    ti.w = G.w[lvl]; ti.h = G.h[lvl]; ti.pitch = G.pitch[lvl];
    ti.tex = F->tex[lvl]; ti.hd = F->host_dIp[lvl]; ti.ha = F->host_abs[lvl]; ti.plane = F->plane[lvl]; ti.Bt = F->B256;
    info[stage] = ti;
    const CUtensorMap *map = &F->map[lvl];
    tensormap_acquire(map);
    mbar_expect_tx(&bar[stage], kGradBoxH * kGradBoxW * sizeof(float));
    tma_load_2d(box[stage], map, ti.x0 - 4, ti.y0 - 1, &bar[stage]);
  };

  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if ((int)blockIdx.x < total) issue(blockIdx.x, 0);
  }
  __syncthreads();

  int it = 0;
  for (int g = blockIdx.x; g < total; g += gridDim.x, it++) {
    const int stage = it & 1;
    if (tid == 0 && g + (int)gridDim.x < total) issue(g + gridDim.x, stage ^ 1);  // stage^1 was released by the barrier below
    mbar_wait(&bar[stage], (it >> 1) & 1);
    const TileInfo ti = info[stage];
    const float *__restrict__ bx = box[stage];
    // 16-B stores need 16-B aligned rows: 12 * w bytes per row, i.e. w % 4 == 0 (levels 0-2 of every shipped calibration, 98 % of
    // the bytes), and a full-width tile; other tiles keep the scalar stores
    const bool vec3 = ti.hd != nullptr && (ti.w & 3) == 0 && ti.x0 + kGradTileW <= ti.w && (reinterpret_cast<size_t>(ti.hd) & 15) == 0;
#pragma unroll
    for (int k = 0; k < kGradTileH / 4; k++) {
      const int lx = tid % kGradTileW, ly = tid / kGradTileW + 4 * k;
      const int x = ti.x0 + lx, y = ti.y0 + ly;
      if (x >= ti.w || y >= ti.h) continue;
      const float *s = bx + (ly + 1) * kGradBoxW + (lx + 4);
      const float c = s[0];
      float dx = 0.f, dy = 0.f, ag = 0.f;
      if (y >= 1 && y <= ti.h - 2) {
        // linear-index neighbours: wrap at the row ends exactly like dI_l[idx-1] / dI_l[idx+1]
        const float left = (x == 0) ? __ldg(ti.plane + (size_t)(y - 1) * ti.pitch + (ti.w - 1)) : s[-1];
        const float right = (x == ti.w - 1) ? __ldg(ti.plane + (size_t)(y + 1) * ti.pitch) : s[1];
        dx = 0.5f * (right - left);
        dy = 0.5f * (s[kGradBoxW] - s[-kGradBoxW]);
        if (!isfinite(dx)) dx = 0.f;
        if (!isfinite(dy)) dy = 0.f;
        ag = dx * dx + dy * dy;
        if (ti.Bt != nullptr) {
          int ci = (int)(c + 0.5f);
          if (ci < 5) ci = 5;
          if (ci > 250) ci = 250;
          const float gw = __ldg(ti.Bt + ci + 1) - __ldg(ti.Bt + ci);
          ag *= gw * gw;
        }
      }
      const size_t idx = (size_t)y * ti.w + x;
      __stcs(ti.tex + idx, make_float4(c, dx, dy, ag));
      if (vec3) {
        sh3[ly][3 * lx + 0] = c;
        sh3[ly][3 * lx + 1] = dx;
        sh3[ly][3 * lx + 2] = dy;
      } else if (ti.hd != nullptr) {
        ti.hd[3 * idx + 0] = c;
        ti.hd[3 * idx + 1] = dx;
        ti.hd[3 * idx + 2] = dy;
      }
      if (ti.ha != nullptr) ti.ha[idx] = ag;
    }
    if (vec3) {
      __syncthreads();
      constexpr int kRowVec = 3 * kGradTileW / 4;  // 48 x 16 B per tile row
      for (int v = tid; v < kGradTileH * kRowVec; v += 256) {
        const int ly = v / kRowVec, c4 = v - ly * kRowVec;
        const int y = ti.y0 + ly;
        if (y < ti.h) *reinterpret_cast<float4 *>(ti.hd + 3 * ((size_t)y * ti.w + ti.x0) + 4 * c4) = *reinterpret_cast<const float4 *>(&sh3[ly][4 * c4]);
      }
    }
    __syncthreads();  // everyone is done with box[stage] / info[stage] / sh3: thread 0 may refill them in the next iteration
  }
}

// texels -> staging copies in the reference's host layouts, all levels in one grid (for frames that were
// built before the host asked for its copies)
__global__ void __launch_bounds__(256) unpack_kernel(const __grid_constant__ FrameBatch B, int total) {
  const PyramidGeom &G = B.G;
  const FrameDev *__restrict__ F = B.f[blockIdx.y];
  for (int i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) {
    int lvl = 0, base = 0, acc = 0;
#pragma unroll
    for (int l = 0; l < kMaxLevels; l++) {
      if (l < G.levels) {
        if (i >= acc) { lvl = l; base = acc; }
        acc += G.w[l] * G.h[l];
      }
    }
    const int idx = i - base;
    const float4 t = F->tex[lvl][idx];
    float *hd = F->host_dIp[lvl], *ha = F->host_abs[lvl];
    if (hd != nullptr) {
      hd[3 * (size_t)idx + 0] = t.x;
      hd[3 * (size_t)idx + 1] = t.y;
      hd[3 * (size_t)idx + 2] = t.z;
    }
    if (ha != nullptr) ha[idx] = t.w;
  }
}

// scaleCoarseDepthL0  src/scale_optimization/TrackerAndScaler.cpp:329-336  (IEEE division, like the host loop)
__global__ void scale_idepth_kernel(float4 *__restrict__ pts, int n, float scale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) pts[i].z = pts[i].z / scale;
}

}  // namespace

cudaError_t launch_downsample(const FrameBatch &B, int nframes, cudaStream_t stream) {
  if (B.G.levels < 2 || nframes < 1) return cudaSuccess;
  dim3 grid((B.G.w[0] + kDownTileW - 1) / kDownTileW, (B.G.h[0] + kDownTileH - 1) / kDownTileH, nframes);
  downsample_chain_kernel<<<grid, 256, 0, stream>>>(B);
  return cudaGetLastError();
}

cudaError_t launch_gradients(const FrameBatch &B, int nframes, cudaStream_t stream, int ctas_per_sm) {
  if (nframes < 1) return cudaSuccess;
  const int total = B.G.tile_begin[B.G.levels] * nframes;
  if (ctas_per_sm < 1 || ctas_per_sm > 6) ctas_per_sm = 6;
  int grid = 148 * ctas_per_sm;  // persistent: up to 6 CTAs (6 x 2 boxes of 5 KB) per SM; fewer when the build shares the GPU with LM rounds
  if (grid > total) grid = total;
  gradient_kernel<<<grid, 256, 0, stream>>>(B, nframes);
  return cudaGetLastError();
}

cudaError_t launch_unpack(const FrameBatch &B, int nframes, cudaStream_t stream) {
  if (nframes < 1) return cudaSuccess;
  int total = 0;
  for (int l = 0; l < B.G.levels; l++) total += B.G.w[l] * B.G.h[l];
  int gx = (total + 255) / 256;
  if (gx > 148 * 8) gx = 148 * 8;
  unpack_kernel<<<dim3(gx, nframes), 256, 0, stream>>>(B, total);
  return cudaGetLastError();
}

// arena of n raw images (contiguous, w*h floats each) -> the level-0 planes of n frames (pitch == w), one float4 per thread
__global__ void __launch_bounds__(256) scatter_planes_kernel(const __grid_constant__ PlaneBatch B, const float4 *__restrict__ arena, int quads) {
  float4 *__restrict__ dst = reinterpret_cast<float4 *>(B.plane[blockIdx.y]);
  const float4 *__restrict__ src = arena + (size_t)blockIdx.y * quads;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < quads; i += gridDim.x * 256) dst[i] = __ldcs(src + i);
}

cudaError_t launch_scatter_planes(const PlaneBatch &B, int nframes, int quads, const float *arena, cudaStream_t stream) {
  if (nframes < 1) return cudaSuccess;
  int gx = (quads + 255) / 256;
  if (gx > 64) gx = 64;
  scatter_planes_kernel<<<dim3(gx, nframes), 256, 0, stream>>>(B, reinterpret_cast<const float4 *>(arena), quads);
  return cudaGetLastError();
}

cudaError_t launch_scale_idepth(float4 *pts, int n, float scale, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  scale_idepth_kernel<<<(n + 255) / 256, 256, 0, stream>>>(pts, n, scale);
  return cudaGetLastError();
}

}  // namespace dslam
