// fp64_pipes.cu — micro-benchmarks of the units the fused residual kernel leans on (B200, sm_100a): DMMA m8n8k4 throughput and
// dependent-chain latency, F2F.F64.F32 / F2F.F32.F64 conversion rate, DFMA rate, IEEE fp32 division rate.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_pipes fp64_pipes.cu ; run: ./fp64_pipes
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int CHAINS>
__global__ void k_dmma(double *out, int iters, double a, double b) {
  double c[CHAINS][2];
#pragma unroll
  for (int i = 0; i < CHAINS; i++) c[i][0] = c[i][1] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) dmma(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS>
__global__ void k_dfma(double *out, int iters, double a, double b) {
  double c[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; i++) c[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS>
__global__ void k_f2f(double *out, int iters, float seed) {
  float f[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; i++) f[i] = seed + threadIdx.x + i;
  double acc = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
      double d;
      asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(f[i]));
      float g;
      asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(g) : "d"(d));
      f[i] = g;  // two conversions per step, dependent
    }
  }
#pragma unroll
  for (int i = 0; i < CHAINS; i++) acc += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int CHAINS>
__global__ void k_div(float *out, int iters, float d) {
  float f[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; i++) f[i] = 1e20f + threadIdx.x + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) f[i] = __fdiv_rn(f[i], d);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  launch();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  launch();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const int sms = p.multiProcessorCount;
  const double ghz = clk_khz * 1e-6;
  printf("%s, %d SMs, %.3f GHz (attribute)\n", p.name, sms, ghz);
  double *out;
  cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  const int iters = 20000;
  // ---- DMMA: throughput (many warps x independent chains) and latency (1 warp, 1 chain)
  for (int warps : {1, 4, 8, 16, 32}) {
    float ms1 = time_ms([&] { k_dmma<1><<<sms, 32 * warps>>>(out, iters, 1.0000001, 0.9999999); });
    float ms4 = time_ms([&] { k_dmma<4><<<sms, 32 * warps>>>(out, iters, 1.0000001, 0.9999999); });
    const double cyc1 = ms1 * 1e-3 * ghz * 1e9 / iters, cyc4 = ms4 * 1e-3 * ghz * 1e9 / (iters * 4.0);
    printf("DMMA m8n8k4  %2d warps/SM: 1 chain %.1f cyc per DMMA per warp ; 4 chains %.1f cyc per DMMA per warp -> %.2f DMMA/clk/SM = %.1f TFLOP/s chip\n", warps,
           cyc1, cyc4, warps / cyc4, warps / cyc4 * 512 * ghz * 1e9 * sms * 1e-12);
  }
  for (int warps : {1, 8, 32}) {
    float ms1 = time_ms([&] { k_dfma<1><<<sms, 32 * warps>>>(out, iters, 1.0000001, 1e-9); });
    float ms4 = time_ms([&] { k_dfma<8><<<sms, 32 * warps>>>(out, iters, 1.0000001, 1e-9); });
    const double cyc1 = ms1 * 1e-3 * ghz * 1e9 / iters, cyc4 = ms4 * 1e-3 * ghz * 1e9 / (iters * 8.0);
    printf("DFMA         %2d warps/SM: latency %.1f cyc ; 8 chains %.2f cyc per DFMA per warp -> %.1f lanes/clk/SM = %.1f TFLOP/s chip\n", warps, cyc1, cyc4,
           32 * warps / cyc4, 32 * warps / cyc4 * 2 * ghz * 1e9 * sms * 1e-12);
  }
  for (int warps : {1, 8, 32}) {
    float ms1 = time_ms([&] { k_f2f<1><<<sms, 32 * warps>>>(out, iters, 1.5f); });
    float ms4 = time_ms([&] { k_f2f<8><<<sms, 32 * warps>>>(out, iters, 1.5f); });
    const double cyc1 = ms1 * 1e-3 * ghz * 1e9 / (iters * 2.0), cyc4 = ms4 * 1e-3 * ghz * 1e9 / (iters * 16.0);
    printf("F2F 64<->32  %2d warps/SM: latency %.1f cyc per conversion ; 8 chains %.2f cyc per conversion per warp -> %.1f lanes/clk/SM\n", warps, cyc1, cyc4,
           32 * warps / cyc4);
  }
  for (int warps : {1, 8, 32}) {
    float ms1 = time_ms([&] { k_div<1><<<sms, 32 * warps>>>((float *)out, iters, 1.0000001f); });
    float ms4 = time_ms([&] { k_div<8><<<sms, 32 * warps>>>((float *)out, iters, 1.0000001f); });
    const double cyc1 = ms1 * 1e-3 * ghz * 1e9 / iters, cyc4 = ms4 * 1e-3 * ghz * 1e9 / (iters * 8.0);
    printf("FDIV.rn f32  %2d warps/SM: latency %.1f cyc ; 8 chains %.2f cyc per div per warp -> %.1f lanes/clk/SM\n", warps, cyc1, cyc4, 32 * warps / cyc4);
  }
  return 0;
}
