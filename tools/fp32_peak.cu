// fp32_peak.cu -- FP32 issue-rate microbenchmark for B200 (sm_100a): scalar FFMA vs packed FFMA2,
// and FFMA co-issued with ALU-pipe FMNMX.  Gives the measured FP32 denominator of the roofline
// (MEASURED_PEAKS.json has no FP32 entry).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cuda_runtime.h>
#include <stdio.h>

constexpr int ITERS = 2048;
constexpr int ACC = 8;

__global__ void k_ffma(float* out, float a, float b) {
  float x[ACC];
#pragma unroll
  for (int i = 0; i < ACC; ++i) x[i] = threadIdx.x * 1e-3f + i;
  float aa = a + threadIdx.x * 1e-9f, bb = b;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < ACC; ++i) x[i] = fmaf(x[i], aa, bb);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ACC; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ffma2(float* out, float a, float b) {
  float2 x[ACC];
#pragma unroll
  for (int i = 0; i < ACC; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
  float2 aa = make_float2(a + threadIdx.x * 1e-9f, a), bb = make_float2(b, b * 0.5f);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < ACC; ++i) x[i] = __ffma2_rn(x[i], aa, bb);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ACC; ++i) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 2 packed FFMA2 + 1 FMNMX + 1 SHF per group: does the ALU pipe co-issue for free?
__global__ void k_mix(float* out, float a, float b) {
  float2 x[ACC];
  float m[ACC];
  unsigned sh[ACC];
#pragma unroll
  for (int i = 0; i < ACC; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i), m[i] = 1e30f, sh[i] = 0;
  float2 aa = make_float2(a + threadIdx.x * 1e-9f, a), bb = make_float2(b, b * 0.5f);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int i = 0; i < ACC; ++i) {
        x[i] = __ffma2_rn(x[i], aa, bb);
        x[i] = __ffma2_rn(x[i], bb, aa);
        m[i] = fminf(m[i], fabsf(x[i].x));
        sh[i] = __funnelshift_l(__float_as_uint(x[i].y), sh[i], 1);
      }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ACC; ++i) s += x[i].x + x[i].y + m[i] + (float)sh[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float time_ms(F f) {
  cudaEvent_t a, b;
  cudaEventCreate(&a), cudaEventCreate(&b);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < 5; ++i) f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms / 5;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int blocks = p.multiProcessorCount * 8, threads = 256;
  float* out;
  cudaMalloc(&out, (size_t)blocks * threads * 4);
  const double nthreads = (double)blocks * threads;
  float ms1 = time_ms([&] { k_ffma<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
  float ms2 = time_ms([&] { k_ffma2<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
  float ms3 = time_ms([&] { k_mix<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
  const double fma1 = nthreads * ITERS * 4 * ACC, fma2 = fma1 * 2, fma3 = nthreads * ITERS * 2 * ACC * 2 * 2;
  printf("{\"sm_count\": %d, \"clock_mhz_max\": %d, ", p.multiProcessorCount, p.clockRate / 1000);
  printf("\"ffma_scalar_tflops\": %.2f, \"ffma2_packed_tflops\": %.2f, \"ffma2_plus_alu_tflops\": %.2f, ", 2 * fma1 / ms1 / 1e9,
         2 * fma2 / ms2 / 1e9, 2 * fma3 / ms3 / 1e9);
  printf("\"ffma_scalar_lanes_per_clk_sm\": %.1f, \"ffma2_lanes_per_clk_sm\": %.1f}\n",
         fma1 / (ms1 * 1e-3) / p.multiProcessorCount / (p.clockRate * 1e3), fma2 / (ms2 * 1e-3) / p.multiProcessorCount / (p.clockRate * 1e3));
  return 0;
}
