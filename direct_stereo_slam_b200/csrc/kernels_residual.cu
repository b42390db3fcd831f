// kernels_residual.cu — fused warp + bilinear sample + Huber residual + Jacobian + normal-equation
// accumulation for the 8-DoF pose tracker and the 1-DoF stereo scale optimiser.  sm_100a.
//
// Replaces, per launch and for every work item (hypothesis) at once:
//   pose : TrackerAndScaler::calcResPose   src/scale_optimization/TrackerAndScaler.cpp:699-852
//          TrackerAndScaler::calcGSSSEPose                                          :640-697
//          Accumulator9::updateSSE_eighted deps:dso/src/OptimizationBackend/MatrixAccumulators.h:1091-1166
//   scale: TrackerAndScaler::calcResScale                                           :1007-1172
//          TrackerAndScaler::calcGSSSEScale                                         :966-1005
//          ScaleAccumulator::updateSSE_oneed src/scale_optimization/ScaleAccumulator.h:60-77
//          getInterpolatedElement33        deps:dso/src/util/globalFuncs.h:75-89
//   (PoseEstimator::calcRes / calcGSSSE, src/loop_closure/pose_estimation/PoseEstimator.cpp:84-296, is the pose flavour
//    with 3-D point records.)
//
// Numerics contract (checked by tests/test_gpu_tracker.py, test_gpu_full_size.py, test_gpu_golden.py against
// oracle/dslam_oracle.cpp):
//   * this translation unit is compiled with -fmad=false and every per-point expression is written in the
//     reference's operation order, so warp, bilinear weights, residual, Huber weight, energy term and the
//     Jacobian row are bit-identical to the fp32 CPU arithmetic;
//   * the weighted outer product is accumulated as acc += (double)(J_r*w) * (double)J_c — the product of two fp32 values
//     is exact in fp64, so the only rounding is the fp64 add.  The reference's 4-lane / 3-tier fp32 SSE accumulation is
//     strictly noisier; the oracle's "fp64" mode is this arithmetic, its "sse" mode is the reference's, and their
//     distance is reported as the noise floor;
//   * reduction order is fixed (thread-sequential, warp reduce-scatter tree, warps in order, CTAs in order),
//     so results are bit-reproducible run to run for a given launch geometry.
//
// Data movement: one 16-B template record + four 16-B texel taps per point — a data-dependent gather through L1/L2 (the
// warp target is only known after the projection, so the taps cannot be a TMA tile; a cp.async-staged software pipeline
// of the taps was built and measured in round 2: 15-35 % SLOWER, the kernel is not bound by the latency of one gather
// but sits at 40-50 % of three units at once — L1 data-pipe wavefronts, the XU pipe (fp32<->fp64 conversions, IEEE
// divisions) and instruction issue; profiles/r02_eval_kernel.md).  Each thread keeps its partial sums in registers; a
// warp's 8x8 block goes through the FP64 tensor cores; a warp folds the remaining fp64 values with a reduce-scatter
// butterfly, warps meet in shared memory, and each CTA publishes one partial record; the last CTA of an item (ticket
// atomic) sums the partials in CTA order and writes the result record directly into mapped pinned host memory as
// self-validating words (payload | sequence number) the host spins on — no memcpy, no stream synchronise and no
// system-scope fence on the LM critical path.

#include "dslam_kernels.h"

namespace dslam {

namespace {

__device__ __forceinline__ double shfl_xor_d(double v, int mask) { return __shfl_xor_sync(0xffffffffu, v, mask); }

// Warp reduce-scatter over NV register-resident doubles. After run(): lanes with writer(lane) hold KEEP
// fully reduced values a[0..KEEP) that belong at output indices base(lane)+0..KEEP-1.
template <int NV>
struct WarpRS;

template <int N>
__device__ __forceinline__ void rs_halve(double *a, int mask, bool upper) {
#pragma unroll
  for (int i = 0; i < N / 2; i++) {
    const double send = upper ? a[i] : a[i + N / 2];
    const double keep = upper ? a[i + N / 2] : a[i];
    a[i] = keep + shfl_xor_d(send, mask);
  }
}

template <>
struct WarpRS<8> {
  __device__ static __forceinline__ void run(double (&a)[8], int lane) {
    rs_halve<8>(a, 16, lane & 16);
    rs_halve<4>(a, 8, lane & 8);
    rs_halve<2>(a, 4, lane & 4);
    a[0] = a[0] + shfl_xor_d(a[0], 2);
    a[0] = a[0] + shfl_xor_d(a[0], 1);
  }
  __device__ static __forceinline__ bool writer(int lane) { return (lane & 3) == 0; }
  __device__ static __forceinline__ int base(int lane) { return ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1); }
};

template <>
struct WarpRS<16> {
  __device__ static __forceinline__ void run(double (&a)[16], int lane) {
    rs_halve<16>(a, 16, lane & 16);
    rs_halve<8>(a, 8, lane & 8);
    rs_halve<4>(a, 4, lane & 4);
    rs_halve<2>(a, 2, lane & 2);
    a[0] = a[0] + shfl_xor_d(a[0], 1);
  }
  __device__ static __forceinline__ bool writer(int lane) { return (lane & 1) == 0; }
  __device__ static __forceinline__ int base(int lane) {
    return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
  }
};

// Eigen 3.3 coefficient product of three terms: e0 + (e1 + e2), third factor is the literal 1.
__device__ __forceinline__ float dot3_xy1(float m0, float m1, float m2, float x, float y) { return m0 * x + (m1 * y + m2); }
__device__ __forceinline__ float dot3_xyz(float m0, float m1, float m2, float x, float y, float z) { return m0 * x + (m1 * y + m2 * z); }

// getInterpolatedElement33 on float4 texels; .w of the texel (absSquaredGrad) is ignored.
__device__ __forceinline__ float3 interp33(const float4 *__restrict__ tex, float x, float y, int width) {
  const int ix = (int)x;
  const int iy = (int)y;
  const float dx = x - ix;
  const float dy = y - iy;
  const float dxdy = dx * dy;
  const float4 *bp = tex + ix + iy * width;
  const float4 p00 = __ldg(bp), p10 = __ldg(bp + 1), p01 = __ldg(bp + width), p11 = __ldg(bp + 1 + width);
  const float w11 = dxdy, w01 = dy - dxdy, w10 = dx - dxdy, w00 = 1 - dx - dy + dxdy;
  float3 r;
  r.x = w11 * p11.x + w01 * p01.x + w10 * p10.x + w00 * p00.x;
  r.y = w11 * p11.y + w01 * p01.y + w10 * p10.y + w00 * p00.y;
  r.z = w11 * p11.z + w01 * p01.z + w10 * p10.z + w00 * p00.z;
  return r;
}

#ifdef DSLAM_KERNEL_TIMING
// diagnostic build only: phase timestamps (globaltimer, ns) of the last launch: [0] first CTA entry, [1] last CTA leaves
// the point loop, [2] last ticket taken, [3] last result record written, [4] first CTA leaves the loop
__device__ unsigned long long g_dbg_times[8];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define DBG_MIN(i) do { if (threadIdx.x == 0) atomicMin(&g_dbg_times[i], gtime()); } while (0)
#define DBG_MAX(i) do { if (threadIdx.x == 0) atomicMax(&g_dbg_times[i], gtime()); } while (0)
#else
#define DBG_MIN(i)
#define DBG_MAX(i)
#endif

constexpr float kHuberTH = 9.0f;  // setting_huberTH  deps:dso/src/util/settings.cpp:127

template <int CAP>
struct BatchT {
  EvalItem item[CAP];
};

enum { kPose = 0, kScale = 1, kPose3d = 2 };


// Flow indicators (:754-784 / :1070-1100 / PoseEstimator.cpp:191-226): every 32nd record of level 0, whether or not it
// projects into the image.  A separate dense pass — inside the main loop one lane per warp would drag the whole warp through
// ~200 extra instructions per point.  Adds 2 x sumSquaredShiftT terms to accT and 2 x sumSquaredShiftRT terms to accRT.
template <int KIND>
__device__ __forceinline__ void flow_pass(const EvalItem &it, int bx, double &accT, double &accRT) {
  const float4 *__restrict__ pts = it.pts;
  const float fxl = it.fx, fyl = it.fy, cxl = it.cx, cyl = it.cy;
  const int nflow = (it.n + 31) >> 5;
  for (int kf = bx * kEvalThreads + threadIdx.x; kf < nflow; kf += it.ppt_stride) {
    const float4 p = __ldg(pts + 32 * kf);
    float sT1, sT2, sRT1, sRT2;
    if (KIND == kPose3d) {
      // same four probes, but on the raw (x, y) of the 3-D point with z replaced by 1 and measured against the projection
      // (Ku0, Kv0) of the untransformed point
      const float x = p.x, y = p.y, z = p.z;
      const float Ku0 = fxl * (x / z) + cxl, Kv0 = fyl * (y / z) + cyl;
      const float pt0 = dot3_xyz(it.M[0], it.M[1], it.M[2], x, y, z) + it.t[0];
      const float pt1 = dot3_xyz(it.M[3], it.M[4], it.M[5], x, y, z) + it.t[1];
      const float pt2 = dot3_xyz(it.M[6], it.M[7], it.M[8], x, y, z) + it.t[2];
      const float Ku = fxl * (pt0 / pt2) + cxl, Kv = fyl * (pt1 / pt2) + cyl;
      const float ptTz = 1.0f + it.t[2], ptT2z = 1.0f - it.t[2];
      const float pt3z = dot3_xy1(it.M[6], it.M[7], it.M[8], x, y) - it.t[2];
      const float KuT = fxl * ((x + it.t[0]) / ptTz) + cxl, KvT = fyl * ((y + it.t[1]) / ptTz) + cyl;
      const float KuT2 = fxl * ((x - it.t[0]) / ptT2z) + cxl, KvT2 = fyl * ((y - it.t[1]) / ptT2z) + cyl;
      const float Ku3 = fxl * ((dot3_xy1(it.M[0], it.M[1], it.M[2], x, y) - it.t[0]) / pt3z) + cxl;
      const float Kv3 = fyl * ((dot3_xy1(it.M[3], it.M[4], it.M[5], x, y) - it.t[1]) / pt3z) + cyl;
      sT1 = (KuT - Ku0) * (KuT - Ku0) + (KvT - Kv0) * (KvT - Kv0);
      sT2 = (KuT2 - Ku0) * (KuT2 - Ku0) + (KvT2 - Kv0) * (KvT2 - Kv0);
      sRT1 = (Ku - Ku0) * (Ku - Ku0) + (Kv - Kv0) * (Kv - Kv0);
      sRT2 = (Ku3 - Ku0) * (Ku3 - Ku0) + (Kv3 - Kv0) * (Kv3 - Kv0);
    } else {
      const float x = p.x, y = p.y, id = p.z;
      const float s = KIND == kScale ? it.p0 : 1.0f;  // (multiplying by the literal 1 is exact)
      const float kx0 = dot3_xy1(s * it.Ki[0], s * it.Ki[1], s * it.Ki[2], x, y);
      const float kx1 = dot3_xy1(s * it.Ki[3], s * it.Ki[4], s * it.Ki[5], x, y);
      const float kx2 = dot3_xy1(s * it.Ki[6], s * it.Ki[7], s * it.Ki[8], x, y);
      const float mx0 = dot3_xy1(s * it.M[0], s * it.M[1], s * it.M[2], x, y);
      const float mx1 = dot3_xy1(s * it.M[3], s * it.M[4], s * it.M[5], x, y);
      const float mx2 = dot3_xy1(s * it.M[6], s * it.M[7], s * it.M[8], x, y);
      const float tT0 = it.t[0] * id, tT1 = it.t[1] * id, tT2 = it.t[2] * id;
      // the point itself (pt = M*(x,y,1) + t*id), as in the main loop
      const float ptz = mx2 + tT2;
      const float Ku = fxl * ((mx0 + tT0) / ptz) + cxl, Kv = fyl * ((mx1 + tT1) / ptz) + cyl;
      const float ptT2z = kx2 + tT2, ptT2nz = kx2 - tT2, pt3z = mx2 - tT2;
      const float KuT = fxl * ((kx0 + tT0) / ptT2z) + cxl, KvT = fyl * ((kx1 + tT1) / ptT2z) + cyl;
      const float KuT2 = fxl * ((kx0 - tT0) / ptT2nz) + cxl, KvT2 = fyl * ((kx1 - tT1) / ptT2nz) + cyl;
      const float Ku3 = fxl * ((mx0 - tT0) / pt3z) + cxl, Kv3 = fyl * ((mx1 - tT1) / pt3z) + cyl;
      sT1 = (KuT - x) * (KuT - x) + (KvT - y) * (KvT - y);
      sT2 = (KuT2 - x) * (KuT2 - x) + (KvT2 - y) * (KvT2 - y);
      sRT1 = (Ku - x) * (Ku - x) + (Kv - y) * (Kv - y);
      sRT2 = (Ku3 - x) * (Ku3 - x) + (Kv3 - y) * (Kv3 - y);
    }
    accT += (double)sT1;
    accT += (double)sT2;
    accRT += (double)sRT1;
    accRT += (double)sRT2;
  }
}

// Scale items: calcResScale + calcGSSSEScale with per-thread fp64 FMAs.
// acc layout: [0] JwJ, [1] Jwr, [2] rwr, [3] E, [4] shiftT, [5] shiftRT
__device__ __forceinline__ void eval_scale_points(const EvalItem &it, int bx, double (&acc)[kScaleVals], int &nE, int &nSat, int &nInl) {
  const int tid = threadIdx.x;
  const float4 *__restrict__ tex = it.tex;
  const float4 *__restrict__ pts = it.pts;
  const int wl = it.w, hl = it.h, n = it.n, stride = it.ppt_stride;
  const float fxl = it.fx, fyl = it.fy, cxl = it.cx, cyl = it.cy;
  const float cutoff = it.cutoff, maxEnergy = it.maxEnergy;
  const float wlm3 = (float)(wl - 3), hlm3 = (float)(hl - 3);

  int i = bx * kEvalThreads + tid;
  float4 p_next = i < n ? __ldg(pts + i) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (; i < n; i += stride) {
    const float4 p = p_next;
    if (i + stride < n) p_next = __ldg(pts + i + stride);  // the next record is in flight while this point is processed
    const float x = p.x, y = p.y, id = p.z, refColor = p.w;
    // :1061  pt = (scale*M) * (x,y,1) + t*id
    const float s = it.p0;
    const float pt0 = dot3_xy1(s * it.M[0], s * it.M[1], s * it.M[2], x, y) + it.t[0] * id;
    const float pt1 = dot3_xy1(s * it.M[3], s * it.M[4], s * it.M[5], x, y) + it.t[1] * id;
    const float pt2 = dot3_xy1(s * it.M[6], s * it.M[7], s * it.M[8], x, y) + it.t[2] * id;
    const float u = pt0 / pt2;
    const float v = pt1 / pt2;
    const float Ku = fxl * u + cxl;
    const float Kv = fyl * v + cyl;
    const float new_idepth = id / pt2;

    if (!(Ku > 2 && Kv > 2 && Ku < wlm3 && Kv < hlm3 && new_idepth > 0)) continue;
    const float3 hit = interp33(tex, Ku, Kv, wl);
    if (!isfinite(hit.x)) continue;
    const float residual = hit.x - refColor;  // :1109 (no affine)
    const float absr = fabsf(residual);
    const float hw = absr < kHuberTH ? 1.0f : kHuberTH / absr;
    nE++;
    if (absr > cutoff) {
      acc[3] += (double)maxEnergy;
      nSat++;
      continue;
    }
    acc[3] += (double)(hw * residual * residual * (2 - hw));
    nInl++;
    // calcGSSSEScale :983-997 (rx = M*(x,y,1) / id, :1068)
    const float rx0 = dot3_xy1(it.M[0], it.M[1], it.M[2], x, y) / id;
    const float rx1 = dot3_xy1(it.M[3], it.M[4], it.M[5], x, y) / id;
    const float rx2 = dot3_xy1(it.M[6], it.M[7], it.M[8], x, y) / id;
    const float tx = it.t[0], ty = it.t[1], tz = it.t[2];
    const float dxfx = hit.y * fxl, dyfy = hit.z * fyl;
    const float deno_sqrt = (it.p0 * rx2) + tz;
    const float deno = 1.0f / (deno_sqrt * deno_sqrt);
    const float xno = (rx0 * tz) - (rx2 * tx);
    const float yno = (rx1 * tz) - (rx2 * ty);
    const float J = (dxfx * (deno * xno)) + (dyfy * (deno * yno));
    const double Jw = (double)(J * hw), rw = (double)(residual * hw);
    acc[0] = fma(Jw, (double)J, acc[0]);
    acc[1] = fma(Jw, (double)residual, acc[1]);
    acc[2] = fma(rw, (double)residual, acc[2]);
  }
  if (it.flags & 1) flow_pass<kScale>(it, bx, acc[4], acc[5]);
}

// D(8x8) += A(8x4) * B(4x8) in fp64 on the tensor cores (DMMA).  Fragments: A[row = lane/4][col = lane%4],
// B[row = lane%4][col = lane/4], C/D[row = lane/4][col = 2*(lane%4) + {0,1}].
__device__ __forceinline__ void dmma_8x8x4(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// index of H(g, n), g <= n, in the 45-entry upper triangle of the 9x9 [J0..J7 r] system
__device__ __forceinline__ int tri9(int g, int n) { return 9 * g - (g * (g - 1)) / 2 + (n - g); }

// Pose items: calcResPose + calcGSSSEPose with the 8x8 block H = sum_p (J_p w_p) J_p^T accumulated by DMMA.
// A warp takes 32 template points per iteration; every lane warps / samples its point and forms its Jacobian row in the
// reference's fp32 arithmetic, the rows go through shared memory (the transposition the MMA fragments need: lane (g, k)
// of MMA m reads element g of point 4m + k; the two 16-B halves of a row are swapped for points 4..7 mod 8, which makes the
// row stores bank-conflict free as well as the fragment loads — unswizzled, the stores were 2-way conflicted and cost 13 %
// of the L1 data-pipe wavefronts) and eight m8n8k4 DMMAs add the 32 outer products to the warp's 8x8
// accumulator — 2 registers per lane instead of 72 per thread, no warp reduction for H at all.  Products of fp32 values
// are exact in fp64, so the result differs from the per-thread FMA form only in the order of the fp64 additions.
// b = sum J w r, sum w r^2, E and the flow sums stay per lane (12 doubles) and are folded by a 16-wide reduce-scatter.
// Results land in sred_w[48] (this warp's slot, pre-zeroed) in the oracle's acc layout.
template <int KIND>  // kPose: template records (u, v, idepth, color); kPose3d: 3-D point records (PoseEstimator)
__device__ __forceinline__ void eval_pose_mma(const EvalItem &it, int bx, double *sred_w, float *sJ, float *sJw, int &nE, int &nSat, int &nInl) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float4 *__restrict__ tex = it.tex;
  const float4 *__restrict__ pts = it.pts;
  const int wl = it.w, hl = it.h, n = it.n, stride = it.ppt_stride;
  const float fxl = it.fx, fyl = it.fy, cxl = it.cx, cyl = it.cy;
  const float cutoff = it.cutoff, maxEnergy = it.maxEnergy;
  const float wlm3 = (float)(wl - 3), hlm3 = (float)(hl - 3);
  const int g = lane >> 2, k = lane & 3;
  constexpr bool point3d = KIND == kPose3d;  // PoseEstimator flavour: records are 3-D points of the matched keyframe
  const int h0 = (lane >> 2) & 1;  // staging swizzle: rows 4..7 (mod 8) keep their 16-B halves swapped

  double c0 = 0.0, c1 = 0.0;  // H(g, 2k), H(g, 2k+1)
  double ext[16];             // [0..7] b, [8] sum w r^2, [9] E, [10] shiftT, [11] shiftRT
#pragma unroll
  for (int i = 0; i < 16; i++) ext[i] = 0.0;

  int base = (bx * (kEvalThreads / 32) + warp) * 32;
  float4 p_next = base + lane < n ? __ldg(pts + base + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (; base < n; base += stride) {  // warp-uniform trip count: mma.sync needs all 32 lanes
    const int i = base + lane;
    const float4 p = p_next;
    if (base + stride + lane < n) p_next = __ldg(pts + base + stride + lane);
    float J[8], Jw[8];
#pragma unroll
    for (int c = 0; c < 8; c++) J[c] = Jw[c] = 0.f;
    if (i < n) {
      const float x = p.x, y = p.y, refColor = p.w;
      // template point : record (u, v, idepth, color)  :747  pt = RKi * (x,y,1) + t*id ; new_idepth = id / pt[2]
      // 3-D point      : record (x, y, z, color)       PoseEstimator.cpp:172-181  pt = R * (x,y,z) + t ; new_idepth = 1 / pt[2]
      // One expression serves both: multiplying by the literal 1 is exact, so zz = 1 / idm = 1 reproduce either form bit for bit.
      const float zz = point3d ? p.z : 1.0f, idm = point3d ? 1.0f : p.z;
      const float pt0 = dot3_xyz(it.M[0], it.M[1], it.M[2], x, y, zz) + it.t[0] * idm;
      const float pt1 = dot3_xyz(it.M[3], it.M[4], it.M[5], x, y, zz) + it.t[1] * idm;
      const float pt2 = dot3_xyz(it.M[6], it.M[7], it.M[8], x, y, zz) + it.t[2] * idm;
      const float u = pt0 / pt2;
      const float v = pt1 / pt2;
      const float Ku = fxl * u + cxl;
      const float Kv = fyl * v + cyl;
      const float new_idepth = idm / pt2;
      if (Ku > 2 && Kv > 2 && Ku < wlm3 && Kv < hlm3 && new_idepth > 0) {
        const float3 hit = interp33(tex, Ku, Kv, wl);
        if (isfinite(hit.x)) {
          const float residual = hit.x - (it.p0 * refColor + it.p1);
          const float absr = fabsf(residual);
          const float hw = absr < kHuberTH ? 1.0f : kHuberTH / absr;
          nE++;
          if (absr > cutoff) {
            ext[9] += (double)maxEnergy;
            nSat++;
          } else {
            ext[9] += (double)(hw * residual * residual * (2 - hw));
            nInl++;
            // calcGSSSEPose :658-678, lane arithmetic of the SSE code
            const float dx = hit.y * fxl, dy = hit.z * fyl;
            J[0] = new_idepth * dx;
            J[1] = new_idepth * dy;
            J[2] = 0.0f - (new_idepth * ((u * dx) + (v * dy)));
            J[3] = 0.0f - (((u * v) * dx) + (dy * (1.0f + (v * v))));
            J[4] = ((u * v) * dy) + (dx * (1.0f + (u * u)));
            J[5] = (u * dy) - (v * dx);
            J[6] = it.p0 * (it.p2 - refColor);
            J[7] = -1.0f;
            const double rd = (double)residual;
#pragma unroll
            for (int c = 0; c < 8; c++) {
              Jw[c] = J[c] * hw;
              ext[c] = fma((double)Jw[c], rd, ext[c]);
            }
            ext[8] = fma((double)(residual * hw), rd, ext[8]);
          }
        }
      }
    }
    // rows -> shared memory (32 B per lane, halves swizzled: every quarter-warp of the STS.128 covers all 32 banks), then
    // the fragment gathers (bank = 8k + (g ^ swizzle): all 32 distinct)
    float4 *dJ = reinterpret_cast<float4 *>(sJ + lane * 8), *dW = reinterpret_cast<float4 *>(sJw + lane * 8);
    dJ[h0] = make_float4(J[0], J[1], J[2], J[3]);
    dJ[h0 ^ 1] = make_float4(J[4], J[5], J[6], J[7]);
    dW[h0] = make_float4(Jw[0], Jw[1], Jw[2], Jw[3]);
    dW[h0 ^ 1] = make_float4(Jw[4], Jw[5], Jw[6], Jw[7]);
    __syncwarp();
#pragma unroll
    for (int m = 0; m < 8; m++) {
      const int e = (4 * m + k) * 8 + (g ^ ((m & 1) << 2));
      const double a = (double)sJw[e];
      const double b = (double)sJ[e];
      dmma_8x8x4(c0, c1, a, b);
    }
    __syncwarp();
  }

  if (it.flags & 1) flow_pass<KIND>(it, bx, ext[10], ext[11]);

  // warp results -> this warp's slot of the CTA reduction buffer (oracle layout: upper triangle, E, shiftT, shiftRT)
  if (g <= 2 * k) sred_w[tri9(g, 2 * k)] = c0;
  if (g <= 2 * k + 1) sred_w[tri9(g, 2 * k + 1)] = c1;
  WarpRS<16>::run(ext, lane);
  if (WarpRS<16>::writer(lane)) {
    const int j = WarpRS<16>::base(lane);
    if (j < 8) sred_w[tri9(j, 8)] = ext[0];
    else if (j == 8) sred_w[44] = ext[0];
    else if (j < 12) sred_w[45 + (j - 9)] = ext[0];
  }
}

// MODE 0 = pose items only, 1 = scale items only, 2 = mixed (bit 1 of EvalItem::flags selects scale; used when the pose
// tracker and the scale optimiser of a stereo frame advance in the same launch).  The grid is flat: the CTAs of item 0,
// then those of item 1, ... (EvalItem::cta_begin); a CTA finds its item by bisection over the <= 128 prefix entries in
// constant memory, so items of very different sizes share a launch without idle CTAs.
// The work of ONE CTA on ONE item (item slot iy of its launch / lane round, CTA bx of the item's it.nblocks): walk the CTA's
// share of the records, reduce to one partial record, take the item's ticket and — in the CTA that draws the last one — sum
// the partials in CTA order and publish the result record to the host.  Shared by the launch kernel below (items in
// kernel-parameter space) and the resident server kernel (items in shared memory).
template <int MODE>
__device__ __forceinline__ void eval_cta(const EvalItem &it, int iy, int bx, EvalScratch scratch, EvalResult *__restrict__ results, unsigned seq) {
  constexpr int NV = MODE == 1 ? kScaleVals : kPoseVals;
  constexpr int NW = kEvalThreads / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  DBG_MIN(0);
  __shared__ double sred[NW][NV];
  __shared__ __align__(16) float sJ[MODE == 1 ? 1 : NW][MODE == 1 ? 4 : 256], sJw[MODE == 1 ? 1 : NW][MODE == 1 ? 4 : 256];
  __shared__ int scnt[3];
  __shared__ int s_last;
  if (tid < 3) scnt[tid] = 0;
  for (int i = lane; i < NV; i += 32) sred[warp][i] = 0.0;
  __syncwarp();
  int nE = 0, nSat = 0, nInl = 0;
  const bool scale_item = MODE == 1 || (MODE == 2 && (it.flags & 2));
  if (!scale_item) {
    if (MODE != 1) {
      if (it.flags & 4) eval_pose_mma<kPose3d>(it, bx, sred[warp], sJ[MODE == 1 ? 0 : warp], sJw[MODE == 1 ? 0 : warp], nE, nSat, nInl);
      else eval_pose_mma<kPose>(it, bx, sred[warp], sJ[MODE == 1 ? 0 : warp], sJw[MODE == 1 ? 0 : warp], nE, nSat, nInl);
    }
  } else {
    double acc[kScaleVals];
#pragma unroll
    for (int i = 0; i < kScaleVals; i++) acc[i] = 0.0;
    eval_scale_points(it, bx, acc, nE, nSat, nInl);
    WarpRS<kScaleVals>::run(acc, lane);
    if (WarpRS<kScaleVals>::writer(lane)) sred[warp][WarpRS<kScaleVals>::base(lane)] = acc[0];
  }

  DBG_MAX(1);
  DBG_MIN(4);
  // ---- CTA reduction ---------------------------------------------------------------------------------
  nE = __reduce_add_sync(0xffffffffu, nE);
  nSat = __reduce_add_sync(0xffffffffu, nSat);
  nInl = __reduce_add_sync(0xffffffffu, nInl);
  __syncthreads();
  if (lane == 0) {
    atomicAdd(&scnt[0], nE);
    atomicAdd(&scnt[1], nSat);
    atomicAdd(&scnt[2], nInl);
  }
  double *part = scratch.partials + ((size_t)iy * kMaxBlocksPerItem + bx) * kPoseVals;
  if (tid < NV) {
    double s = sred[0][tid];
#pragma unroll
    for (int wv = 1; wv < NW; wv++) s += sred[wv][tid];
    __stcg(part + tid, s);
  }
  // Publication protocol: the CTA's writes are ordered before the barrier; ONE thread then issues the gpu-scope
  // fence (cumulative over everything it observed through the barrier) and takes the ticket.  Fencing from every
  // thread costs a MEMBAR per warp and was 1/3 of the kernel's stall samples.
  __syncthreads();
  int *cnt = scratch.counters + iy * 4;
  if (tid == 0) {
    atomicAdd(cnt + 0, scnt[0]);
    atomicAdd(cnt + 1, scnt[1]);
    atomicAdd(cnt + 2, scnt[2]);
    // one acq_rel RMW instead of fence + relaxed atomic + fence: releases this CTA's partial record and counters
    // (ordered before it through the barrier) and acquires those of the CTAs that took earlier tickets
    int ticket;
    asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], 1;" : "=r"(ticket) : "l"(cnt + 3) : "memory");
    s_last = (ticket == it.nblocks - 1);
  }
  DBG_MAX(2);
  __syncthreads();
  if (!s_last) return;

  // ---- last CTA of the item: ordered sum of the partials, publish to the host ---------------------------
  EvalResult *res = results + iy;
  const double *pbase = scratch.partials + (size_t)iy * kMaxBlocksPerItem * kPoseVals;
  // 128 threads: PARTS interleaved groups per value; every thread first issues ALL its loads (independent, one L2
  // latency in total instead of one per partial) and then adds them in CTA order — the order is fixed, so the result
  // is bit-reproducible for a given launch geometry.
  constexpr int PARTS = kEvalThreads / NV >= 2 ? 2 : 1;
  constexpr int CHUNK = 12;  // independent loads in flight per thread (one L2 latency per chunk instead of one per partial)
  __shared__ double sfin[PARTS][NV];
  if (tid < NV * PARTS) {
    const int vi = tid % NV, part_i = tid / NV;
    double s = 0.0;
    for (int b0 = part_i; b0 < it.nblocks; b0 += CHUNK * PARTS) {
      double v[CHUNK];
#pragma unroll
      for (int k = 0; k < CHUNK; k++) {
        const int b = b0 + k * PARTS;
        v[k] = b < it.nblocks ? __ldcg(pbase + (size_t)b * kPoseVals + vi) : 0.0;
      }
#pragma unroll
      for (int k = 0; k < CHUNK; k++) s += v[k];
    }
    sfin[part_i][vi] = s;
  }
  __syncthreads();
  unsigned long long *w = res->w;
  if (tid < NV) {
    double s = sfin[0][tid];
    if (PARTS == 2) s += sfin[1][tid];
    const unsigned long long bits = (unsigned long long)__double_as_longlong(s);
    w[2 * tid] = (bits & 0xffffffff00000000ull) | seq;
    w[2 * tid + 1] = (bits << 32) | seq;
  }
  if (tid >= 64 && tid < 67) {
    const int k = tid - 64;
    const int v = __ldcg(cnt + k);
    w[kResultCountBase + k] = ((unsigned long long)(unsigned)v << 32) | seq;
    __stcg(cnt + k, 0);
  }
  DBG_MAX(3);
  if (tid == 67) __stcg(cnt + 3, 0);  // ticket: the next evaluation of this slot starts after this item's result was consumed
}

#ifndef DSLAM_EVAL_MIN_CTAS
#define DSLAM_EVAL_MIN_CTAS 5
#endif
template <int MODE, int CAP>
__global__ void __launch_bounds__(kEvalThreads, DSLAM_EVAL_MIN_CTAS) eval_kernel(const __grid_constant__ BatchT<CAP> batch, EvalScratch scratch,
                                                           EvalResult *__restrict__ results, unsigned seq, int nitems) {
  int iy = 0;
  if (CAP > 1) {  // last item whose cta_begin <= blockIdx.x
    int lo = 0, hi = nitems - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (batch.item[mid].cta_begin <= (int)blockIdx.x) lo = mid;
      else hi = mid - 1;
    }
    iy = lo;
  }
  const EvalItem &it = batch.item[iy];
  eval_cta<MODE>(it, iy, (int)blockIdx.x - it.cta_begin, scratch, results, seq);
}

// ---- resident evaluation server -------------------------------------------------------------------------------------
// One kernel that stays on the GPU for the duration of a lock-step LM call and takes the launch out of every LM round:
//   * the host threads ("lanes") write the items of their next round and a list of CTA-sized work units (item, CTA index)
//     into mapped pinned memory and ring their doorbell (a round counter in a 128-byte line of their own);
//   * CTA 0 is the DISPATCHER: one warp polls all doorbells at once (one PCIe read per lane, issued together), then the CTA
//     pulls the rung lanes' items into HBM and their units into a bounded multi-producer / multi-consumer queue in HBM
//     (a slot = one 64-bit word: generation | lap tag | lane | item | CTA index, so publishing a unit is one store);
//   * every other CTA is a WORKER: it draws a ticket, waits for that queue slot to carry its tag, copies the unit's item into
//     shared memory and runs eval_cta — the very code of the launch kernel — publishing results exactly as launches do.
// Nothing here can hang the GPU: the dispatcher leaves when the host says stop, when no doorbell rang for idle_ns, or after
// life_ns; it then poisons the queue so that every worker leaves too, and a worker that waits longer than life_ns leaves by
// itself.  The 8x8 solve stays on the host (north star); the doorbell replaces cudaLaunchKernel, whose cost under eight
// contending host threads (16-19 us) was a fifth of an LM round.
constexpr unsigned kSrvPoison = 0xffu;
__device__ __forceinline__ unsigned ld_volatile_u32(const volatile unsigned *p) {
  unsigned v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ld_volatile_v4(const void *p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long srv_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned long long srv_slot_word(unsigned gen, unsigned idx, unsigned lane, unsigned item, unsigned bx) {
  const unsigned long long tag = ((unsigned long long)(gen & 0xffu) << 16) | ((idx / kSrvQueueSlots + 1u) & 0xffffu);
  return (tag << 40) | ((unsigned long long)(lane & 0xffu) << 32) | ((unsigned long long)(item & 0xffu) << 24) | ((unsigned long long)(bx & 0xffffu) << 8);
}

__global__ void __launch_bounds__(kEvalThreads, DSLAM_EVAL_MIN_CTAS) eval_server_kernel(const __grid_constant__ ServerParams P) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned long long t0 = srv_now();
  if (blockIdx.x == 0) {
    // ---------------- dispatcher ----------------
    __shared__ unsigned s_round[kSrvLanes], s_seen[kSrvLanes], s_mask, s_hdr[4], s_base;
    __shared__ int s_stop;
    if (tid < kSrvLanes) s_seen[tid] = P.last_round[tid];
    __syncthreads();
    unsigned long long t_act = t0;  // (lane 0 of warp 0 keeps the clock: every decision below is taken by one thread and read by all)
    for (;;) {
      if (warp == 0) {
        unsigned r = 0;
        bool rung = false;
        if (lane < P.nlanes) {
          r = ld_volatile_u32(&P.doors[lane].round);
          rung = r != s_seen[lane];
          s_round[lane] = r;
        }
        const unsigned mask = __ballot_sync(0xffffffffu, rung);
        if (lane == 0) {
          const unsigned long long now = srv_now();
          if (mask) t_act = now;
          s_mask = mask;
          s_stop = mask == 0 && (ld_volatile_u32(&P.doors[0].stop) == P.gen || now - t_act > P.idle_ns || now - t0 > P.life_ns);
        }
      }
      __syncthreads();
      unsigned mask = s_mask;
      const bool leave = s_stop != 0;
      __syncthreads();  // s_mask / s_stop are rewritten by the next poll
      if (leave) break;
      if (mask == 0) continue;
      while (mask) {
        const int L = __ffs(mask) - 1;
        mask &= mask - 1;
        if (tid == 0) {
          const uint4 h = ld_volatile_v4((const void *)&P.doors[L]);  // round, n_items, n_units, seq
          s_hdr[0] = h.y;
          s_hdr[1] = h.z;
          s_hdr[2] = h.w;
          s_base = atomicAdd(P.qctl + 1, h.z);
          P.lane_seq[L] = h.w;
        }
        __syncthreads();
        const int n_items = (int)s_hdr[0], n_units = (int)s_hdr[1];
        const unsigned base = s_base;
        // items: host -> HBM (all loads of a thread are issued before its stores: one PCIe round trip, not one per record)
        const uint4 *src = reinterpret_cast<const uint4 *>(P.lane[L].items_host);
        uint4 *dst = reinterpret_cast<uint4 *>(P.items_dev + (size_t)L * kMaxItemsPerLaunch);
        constexpr int IT4 = (int)(sizeof(EvalItem) / 16);
        for (int k0 = 0; k0 < n_items * IT4; k0 += kEvalThreads * 8) {
          uint4 v[8];
#pragma unroll
          for (int j = 0; j < 8; j++) {
            const int k = k0 + j * kEvalThreads + tid;
            if (k < n_items * IT4) v[j] = ld_volatile_v4(src + k);
          }
#pragma unroll
          for (int j = 0; j < 8; j++) {
            const int k = k0 + j * kEvalThreads + tid;
            if (k < n_items * IT4) dst[k] = v[j];
          }
        }
        unsigned u[4];
        const unsigned *usrc = P.lane[L].units_host;
#pragma unroll
        for (int j = 0; j < 4; j++) {  // <= 512 units per round and lane
          const int k = j * kEvalThreads + tid;
          u[j] = k < n_units ? ld_volatile_u32(usrc + k) : 0u;
        }
        __threadfence();   // items and lane_seq before the units that point at them
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int k = j * kEvalThreads + tid;
          if (k < n_units) {
            const unsigned idx = base + (unsigned)k;
            st_release_u64(P.queue + idx % kSrvQueueSlots, srv_slot_word(P.gen, idx, (unsigned)L, u[j] >> 16, u[j] & 0xffffu));
          }
        }
        if (tid == 0) s_seen[L] = s_round[L];
        __syncthreads();
      }
    }
    // leave: poison every worker
    if (tid == 0) s_base = atomicAdd(P.qctl + 1, gridDim.x - 1);
    __syncthreads();
    for (unsigned k = tid; k < gridDim.x - 1; k += kEvalThreads) {
      const unsigned idx = s_base + k;
      st_release_u64(P.queue + idx % kSrvQueueSlots, srv_slot_word(P.gen, idx, kSrvPoison, kSrvPoison, 0));
    }
    return;
  }
  // ---------------- worker ----------------
  __shared__ unsigned long long s_desc;
  __shared__ unsigned s_seq;
  __shared__ __align__(16) unsigned char s_item[sizeof(EvalItem)];
  for (;;) {
    __syncthreads();
    if (tid == 0) {
      const unsigned idx = atomicAdd(P.qctl + 0, 1u);
      const unsigned long long want = (((unsigned long long)(P.gen & 0xffu) << 16) | ((idx / kSrvQueueSlots + 1u) & 0xffffu));
      const unsigned long long *slot = P.queue + idx % kSrvQueueSlots;
      unsigned long long d;
      unsigned spins = 0;
      for (;;) {
        d = ld_acquire_u64(slot);
        if ((d >> 40) == want) break;
        if ((++spins & 63u) == 0 && srv_now() - t0 > P.life_ns) {  // the dispatcher is gone or the host never came back
          d = srv_slot_word(P.gen, idx, kSrvPoison, kSrvPoison, 0);
          break;
        }
        __nanosleep(100);
      }
      s_desc = d;
    }
    __syncthreads();
    const unsigned long long d = s_desc;
    const unsigned L = (unsigned)(d >> 32) & 0xffu, item = (unsigned)(d >> 24) & 0xffu, bx = (unsigned)(d >> 8) & 0xffffu;
    if (L == kSrvPoison) return;
    constexpr int IT4 = (int)(sizeof(EvalItem) / 16);
    if (tid < IT4) reinterpret_cast<uint4 *>(s_item)[tid] = __ldcg(reinterpret_cast<const uint4 *>(P.items_dev + (size_t)L * kMaxItemsPerLaunch + item) + tid);
    if (tid == 32) s_seq = __ldcg(P.lane_seq + L);
    __syncthreads();
    EvalScratch sc;
    sc.partials = P.lane[L].partials;
    sc.counters = P.lane[L].counters;
    eval_cta<2>(*reinterpret_cast<const EvalItem *>(s_item), (int)item, (int)bx, sc, P.lane[L].results, s_seq);
  }
}

template <int MODE, int CAP>
cudaError_t launch_cap(const EvalBatch &batch, int nitems, int total_ctas, EvalScratch scratch, EvalResult *results, unsigned seq,
                       cudaStream_t stream) {
  BatchT<CAP> b;
  for (int i = 0; i < nitems; i++) b.item[i] = batch.item[i];
  eval_kernel<MODE, CAP><<<total_ctas, kEvalThreads, 0, stream>>>(b, scratch, results, seq, nitems);
  return cudaGetLastError();
}

// The items travel in kernel-parameter space; the cost of a launch (host side and front end) grows with the size of the
// parameter block (176 B per item), so the capacity is chosen close to the item count.
template <int MODE>
cudaError_t launch_mode(const EvalBatch &batch, int nitems, int total_ctas, EvalScratch scratch, EvalResult *results, unsigned seq,
                        cudaStream_t stream) {
  if (nitems <= 1) return launch_cap<MODE, 1>(batch, nitems, total_ctas, scratch, results, seq, stream);
  if (nitems <= 4) return launch_cap<MODE, 4>(batch, nitems, total_ctas, scratch, results, seq, stream);
  if (nitems <= 8) return launch_cap<MODE, 8>(batch, nitems, total_ctas, scratch, results, seq, stream);
  if (nitems <= 16) return launch_cap<MODE, 16>(batch, nitems, total_ctas, scratch, results, seq, stream);
  if (nitems <= 24) return launch_cap<MODE, 24>(batch, nitems, total_ctas, scratch, results, seq, stream);
  if (nitems <= 32) return launch_cap<MODE, 32>(batch, nitems, total_ctas, scratch, results, seq, stream);
  if (nitems <= 48) return launch_cap<MODE, 48>(batch, nitems, total_ctas, scratch, results, seq, stream);
  if (nitems <= 64) return launch_cap<MODE, 64>(batch, nitems, total_ctas, scratch, results, seq, stream);
  if (nitems <= 96) return launch_cap<MODE, 96>(batch, nitems, total_ctas, scratch, results, seq, stream);
  return launch_cap<MODE, kMaxItemsPerLaunch>(batch, nitems, total_ctas, scratch, results, seq, stream);
}

}  // namespace

#ifdef DSLAM_KERNEL_TIMING
cudaError_t debug_times(unsigned long long *out8, int reset) {
  unsigned long long init[8] = {~0ull, 0, 0, 0, ~0ull, 0, 0, 0};
  cudaError_t e = cudaMemcpyFromSymbol(out8, g_dbg_times, sizeof(init));
  if (e == cudaSuccess && reset) e = cudaMemcpyToSymbol(g_dbg_times, init, sizeof(init));
  return e;
}
#endif

cudaError_t launch_eval_server(const ServerParams &P, int workers, cudaStream_t stream) {
  if (workers < 1 || P.nlanes < 1 || P.nlanes > kSrvLanes) return cudaErrorInvalidValue;
  eval_server_kernel<<<workers + 1, kEvalThreads, 0, stream>>>(P);
  return cudaGetLastError();
}

cudaError_t launch_eval(int mode, const EvalBatch &batch, int nitems, int total_ctas, EvalScratch scratch, EvalResult *results_dev,
                        unsigned seq, cudaStream_t stream) {
  if (nitems < 1 || nitems > kMaxItemsPerLaunch || total_ctas < nitems || total_ctas > nitems * kMaxBlocksPerItem) return cudaErrorInvalidValue;
  if (mode == 0) return launch_mode<0>(batch, nitems, total_ctas, scratch, results_dev, seq, stream);
  if (mode == 1) return launch_mode<1>(batch, nitems, total_ctas, scratch, results_dev, seq, stream);
  return launch_mode<2>(batch, nitems, total_ctas, scratch, results_dev, seq, stream);
}

}  // namespace dslam
