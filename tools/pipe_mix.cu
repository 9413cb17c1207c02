// pipe_mix.cu -- which non-FMA instructions cost FP32-pipe throughput on sm_100a?
// Every variant runs NF packed FFMA2 (or 2*NF scalar FFMA) plus NA instructions of one other kind per
// group, on independent accumulators, and reports SM cycles per group per SMSP-warp-slot.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_mix pipe_mix.cu
#include <cuda_runtime.h>
#include <stdio.h>

constexpr int ITERS = 1024;
constexpr int ACC = 8;

enum Op { NONE, FMNMX, FMNMX3, SHF, LOP3, IADD, MUFU, FSETP_SEL, IMAD, FADD2OP, LDS128, POPC };

template <int OP>
__device__ __forceinline__ void other(float2& x, float& m, unsigned& u, const float4* sm, int i) {
  if constexpr (OP == FMNMX) {
    asm volatile("min.NaN.f32 %0, %0, %1;" : "+f"(m) : "f"(fabsf(x.x)));
  } else if constexpr (OP == FMNMX3) {
    asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m) : "f"(fabsf(x.x)), "f"(fabsf(x.y)));
  } else if constexpr (OP == SHF) {
    u = __funnelshift_l(__float_as_uint(x.y), u, 1);
  } else if constexpr (OP == LOP3) {
    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u) : "r"(__float_as_uint(x.x)), "r"(__float_as_uint(x.y)));
  } else if constexpr (OP == IADD) {
    asm volatile("add.u32 %0, %0, %1;" : "+r"(u) : "r"(__float_as_uint(x.x)));
  } else if constexpr (OP == MUFU) {
    float r;
    asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x.x));
    m += r;  // costs an FADD too; see NONE+FADD for the reference
  } else if constexpr (OP == FSETP_SEL) {
    asm volatile("{.reg .pred p; setp.lt.f32 p, %1, 0f00000000; @p add.u32 %0, %0, 1;}" : "+r"(u) : "f"(x.x));
  } else if constexpr (OP == IMAD) {
    asm volatile("mad.lo.u32 %0, %0, 3, %1;" : "+r"(u) : "r"(__float_as_uint(x.x)));
  } else if constexpr (OP == FADD2OP) {
    float2 t = __fadd2_rn(x, make_float2(m, m));
    m = t.x;
  } else if constexpr (OP == LDS128) {
    float4 v = sm[(i + (int)u) & 63];
    m += v.x;
    u += 1;
  } else if constexpr (OP == POPC) {
    u += __popc(__float_as_uint(x.x));
  }
}

template <int OP, int NF, int NA, bool SCALAR>
__global__ void __launch_bounds__(128) k(float* out, float a, float b, long long* cyc) {
  __shared__ float4 sm[64];
  if (threadIdx.x < 64) sm[threadIdx.x] = make_float4(a, b, a, b);
  __syncthreads();
  float2 x[ACC];
  float m[ACC];
  unsigned u[ACC];
#pragma unroll
  for (int i = 0; i < ACC; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i), m[i] = 1e30f, u[i] = 0;
  float2 aa = make_float2(a + threadIdx.x * 1e-9f, a), bb = make_float2(b, b * 0.5f);
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ACC; ++i) {
#pragma unroll
      for (int f = 0; f < NF; ++f) {
        if constexpr (SCALAR) {
          x[i].x = fmaf(x[i].x, aa.x, bb.x);
          x[i].y = fmaf(x[i].y, aa.y, bb.y);
        } else {
          x[i] = __ffma2_rn(x[i], aa, bb);
        }
      }
#pragma unroll
      for (int q = 0; q < NA; ++q) other<OP>(x[i], m[i], u[i], sm, i);
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < ACC; ++i) s += x[i].x + x[i].y + m[i] + (float)u[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP, int NF, int NA, bool SCALAR>
void run(const char* name, int ctas_per_sm, float* out, long long* dcyc, int sms) {
  const int blocks = sms * ctas_per_sm, threads = 128;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  k<OP, NF, NA, SCALAR><<<blocks, threads>>>(out, 1.0001f, 0.5f, dcyc);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<OP, NF, NA, SCALAR><<<blocks, threads>>>(out, 1.0001f, 0.5f, dcyc);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  long long cyc;
  cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost);
  // warps per SMSP = ctas_per_sm (128 threads = 4 warps = 1 per SMSP); groups per warp = ITERS*ACC
  const double per_group = (double)cyc / ((double)ITERS * ACC * ctas_per_sm);
  const double tflops = 2.0 * 2 * NF * (double)blocks * threads * ITERS * ACC / (ms * 1e9);
  printf("{\"variant\": \"%s\", \"ffma2_per_group\": %d, \"other_per_group\": %d, \"scalar\": %s, \"warps_per_smsp\": %d, "
         "\"smsp_cycles_per_group\": %.3f, \"fp32_tflops\": %.2f}\n",
         name, NF, NA, SCALAR ? "true" : "false", ctas_per_sm, per_group, tflops);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  float* out;
  long long* dcyc;
  cudaMalloc(&out, (size_t)p.multiProcessorCount * 16 * 128 * 4);
  cudaMalloc(&dcyc, 8);
  const int sms = p.multiProcessorCount;
  for (int w : {4, 8}) {
    run<NONE, 4, 0, false>("ffma2_only", w, out, dcyc, sms);
    run<NONE, 4, 0, true>("ffma_scalar_only", w, out, dcyc, sms);
    run<FMNMX, 4, 1, false>("fmnmx", w, out, dcyc, sms);
    run<FMNMX, 4, 2, false>("fmnmx", w, out, dcyc, sms);
    run<FMNMX3, 4, 1, false>("fmnmx3", w, out, dcyc, sms);
    run<SHF, 4, 1, false>("shf", w, out, dcyc, sms);
    run<SHF, 4, 2, false>("shf", w, out, dcyc, sms);
    run<LOP3, 4, 1, false>("lop3", w, out, dcyc, sms);
    run<LOP3, 4, 2, false>("lop3", w, out, dcyc, sms);
    run<IADD, 4, 1, false>("iadd", w, out, dcyc, sms);
    run<IADD, 4, 2, false>("iadd", w, out, dcyc, sms);
    run<POPC, 4, 1, false>("popc", w, out, dcyc, sms);
    run<MUFU, 4, 1, false>("mufu_rsq+fadd", w, out, dcyc, sms);
    run<FSETP_SEL, 4, 1, false>("fsetp+pred_iadd", w, out, dcyc, sms);
    run<IMAD, 4, 1, false>("imad", w, out, dcyc, sms);
    run<FADD2OP, 4, 1, false>("fadd2", w, out, dcyc, sms);
    run<LDS128, 4, 1, false>("lds128+fadd+iadd", w, out, dcyc, sms);
    run<FMNMX, 4, 1, true>("fmnmx", w, out, dcyc, sms);
    run<SHF, 4, 1, true>("shf", w, out, dcyc, sms);
    run<SHF, 4, 2, true>("shf", w, out, dcyc, sms);
  }
  return 0;
}
