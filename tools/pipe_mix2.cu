// pipe_mix2.cu -- second pipe microbenchmark for sm_100a: what does one extra instruction of a given
// form cost a stream of packed FFMA2?  Separates register-read count, register write and pipe.
// Output: one JSON line per variant with "slot_cycles" = SMSP cycles per group of 4 FFMA2 (+ extras),
// derived from the event-timed duration of a full-chip launch (8 warps per SMSP).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_mix2 pipe_mix2.cu
#include <cuda_runtime.h>
#include <stdio.h>

constexpr int ITERS = 4096;
constexpr int ACC = 8;

enum Form { F_RRR, F_RSR, F_RSU, F_RRU, F_SCALAR_RRR, F_SCALAR_RRU };
enum Op {
  NONE, FSETP_RR, FSETP_RI, IADD_IMM, IADD_RR, FMNMX_RR, FMNMX_RU, LOP3_RRI, LOP3_RRR, SHF_RR, PRED_LOP3_T, PRED_LOP3_F,
  FSETP2_AND, MUFU_ONLY, FMNMX3_RRR, FADD2_RR, FADD2_RU, LDS128_ONLY, FMNMX_ABS, SHF_X2, MAXSHFMIN, SETP2_PLOP_MIN
};

template <int OP>
__device__ __forceinline__ void other(float2& x, float& m, unsigned& u, float cu, unsigned ubit, const float4* sm, int i,
                                      float4& sink) {
  if constexpr (OP == FSETP_RR) {
    asm volatile("{.reg .pred p; setp.lt.f32 p, %1, %2; @p add.u32 %0, %0, 1;}" : "+r"(u) : "f"(x.x), "f"(m));  // m=-1e30: never
  } else if constexpr (OP == FSETP_RI) {
    asm volatile("{.reg .pred p; setp.lt.f32 p, %1, 0fF0000000; @p add.u32 %0, %0, 1;}" : "+r"(u) : "f"(x.x));
  } else if constexpr (OP == IADD_IMM) {
    asm volatile("add.u32 %0, %0, 3;" : "+r"(u));
  } else if constexpr (OP == IADD_RR) {
    asm volatile("add.u32 %0, %0, %1;" : "+r"(u) : "r"(__float_as_uint(x.x)));
  } else if constexpr (OP == FMNMX_RR) {
    asm volatile("min.NaN.f32 %0, %0, %1;" : "+f"(m) : "f"(x.x));
  } else if constexpr (OP == FMNMX_ABS) {
    asm volatile("{.reg .f32 t; abs.f32 t, %1; min.NaN.f32 %0, %0, t;}" : "+f"(m) : "f"(x.x));
  } else if constexpr (OP == FMNMX_RU) {
    asm volatile("min.NaN.f32 %0, %1, %2;" : "=f"(m) : "f"(x.x), "f"(cu));
  } else if constexpr (OP == LOP3_RRI) {
    asm volatile("lop3.b32 %0, %0, %1, 0x55aa55aa, 0x96;" : "+r"(u) : "r"(__float_as_uint(x.x)));
  } else if constexpr (OP == LOP3_RRR) {
    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u) : "r"(__float_as_uint(x.x)), "r"(__float_as_uint(x.y)));
  } else if constexpr (OP == SHF_RR) {
    u = __funnelshift_l(__float_as_uint(x.y), u, 1);
  } else if constexpr (OP == SHF_X2) {
    u = __funnelshift_l(__float_as_uint(x.y), u, 1);
    u = __funnelshift_l(__float_as_uint(x.x), u, 1);
  } else if constexpr (OP == PRED_LOP3_T) {  // predicate always true
    asm volatile("{.reg .pred p; setp.gt.f32 p, %1, 0fF0000000; @p or.b32 %0, %0, %2;}" : "+r"(u) : "f"(x.x), "r"(ubit));
  } else if constexpr (OP == PRED_LOP3_F) {  // predicate never true
    asm volatile("{.reg .pred p; setp.lt.f32 p, %1, 0fF0000000; @p or.b32 %0, %0, %2;}" : "+r"(u) : "f"(x.x), "r"(ubit));
  } else if constexpr (OP == FSETP2_AND) {
    asm volatile(
        "{.reg .pred p, q; setp.gt.f32 p, %1, 0fF0000000; setp.gt.and.f32 q, %2, 0fF0000000, p; @q or.b32 %0, %0, %3;}"
        : "+r"(u)
        : "f"(x.x), "f"(x.y), "r"(ubit));
  } else if constexpr (OP == MUFU_ONLY) {
    asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(m) : "f"(x.x));
  } else if constexpr (OP == FMNMX3_RRR) {
    asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m) : "f"(x.x), "f"(x.y));
  } else if constexpr (OP == FADD2_RR) {
    float2 t = __fadd2_rn(x, make_float2(m, cu));
    m = t.x;
    sink.y += t.y;
  } else if constexpr (OP == LDS128_ONLY) {
    sink = sm[i];
  } else if constexpr (OP == MAXSHFMIN) {  // what the score kernel does per evaluation today
    float t;
    asm volatile("max.NaN.f32 %0, %1, %2;" : "=f"(t) : "f"(x.x), "f"(x.y));
    u = __funnelshift_l(__float_as_uint(t), u, 1);
    asm volatile("{.reg .f32 a; abs.f32 a, %1; min.NaN.f32 %0, %0, a;}" : "+f"(m) : "f"(t));
  } else if constexpr (OP == SETP2_PLOP_MIN) {  // candidate replacement: 2 FSETP + predicated OR + min3(|e|,|nt|)
    asm volatile(
        "{.reg .pred p, q; .reg .f32 a, b; setp.lt.f32 p, %2, 0f00000000; setp.lt.and.f32 q, %3, 0f00000000, p; "
        "@q or.b32 %0, %0, %4; abs.f32 a, %2; abs.f32 b, %3; min.f32 %1, %1, a, b;}"
        : "+r"(u), "+f"(m)
        : "f"(x.x), "f"(x.y), "r"(ubit));
  }
}

template <int FORM, int OP, int NA>
__global__ void __launch_bounds__(128) k(float* out, const float a, const float b, const float cu, const unsigned ubit) {
  __shared__ float4 sm[64];
  if (threadIdx.x < 64) sm[threadIdx.x] = make_float4(a, b, a, b);
  __syncthreads();
  float2 x[ACC];
  float m[ACC], s[ACC];
  unsigned u[ACC];
  float4 sink = make_float4(0, 0, 0, 0);
#pragma unroll
  for (int i = 0; i < ACC; ++i) {
    x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
    m[i] = -1e30f;
    u[i] = 0;
    s[i] = a + i * 1e-7f + threadIdx.x * 1e-9f;
  }
  float2 aa = make_float2(a + threadIdx.x * 1e-9f, a), bb = make_float2(b + threadIdx.x * 1e-9f, b * 0.5f + threadIdx.x * 1e-9f);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ACC; ++i) {
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        if constexpr (FORM == F_RRR) x[i] = __ffma2_rn(x[i], aa, bb);
        if constexpr (FORM == F_RSR) x[i] = __ffma2_rn(x[i], make_float2(s[(i + f) & (ACC - 1)], s[(i + f) & (ACC - 1)]), bb);
        if constexpr (FORM == F_RSU) x[i] = __ffma2_rn(x[i], make_float2(s[(i + f) & (ACC - 1)], s[(i + f) & (ACC - 1)]), make_float2(cu, cu));
        if constexpr (FORM == F_RRU) x[i] = __ffma2_rn(x[i], aa, make_float2(cu, cu));
        if constexpr (FORM == F_SCALAR_RRR) {
          x[i].x = fmaf(x[i].x, s[(i + f) & (ACC - 1)], bb.x);
          x[i].y = fmaf(x[i].y, s[(i + f) & (ACC - 1)], bb.y);
        }
        if constexpr (FORM == F_SCALAR_RRU) {
          x[i].x = fmaf(x[i].x, s[(i + f) & (ACC - 1)], cu);
          x[i].y = fmaf(x[i].y, s[(i + f) & (ACC - 1)], cu);
        }
      }
#pragma unroll
      for (int q = 0; q < NA; ++q) other<OP>(x[i], m[i], u[i], cu, ubit, sm, i, sink);
    }
  }
  float r = sink.x + sink.y + sink.z + sink.w;
#pragma unroll
  for (int i = 0; i < ACC; ++i) r += x[i].x + x[i].y + m[i] + (float)u[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

static int g_sms = 0;
static float g_mhz = 0;
static float* g_out = nullptr;

template <int FORM, int OP, int NA>
void run(const char* form, const char* name) {
  const int warps_per_smsp = 8;
  const int blocks = g_sms * warps_per_smsp, threads = 128;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  k<FORM, OP, NA><<<blocks, threads>>>(g_out, 1.0001f, 0.5f, 0.25f, 4u);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 3; ++r) k<FORM, OP, NA><<<blocks, threads>>>(g_out, 1.0001f, 0.5f, 0.25f, 4u);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= 3;
  const double groups_per_smsp = (double)ITERS * ACC * warps_per_smsp;
  const double cyc = ms * 1e-3 * g_mhz * 1e6 / groups_per_smsp;
  printf("{\"ffma_form\": \"%s\", \"extra\": \"%s\", \"extra_per_4ffma2\": %d, \"slot_cycles_per_group\": %.3f, \"ms\": %.4f}\n", form,
         name, NA, cyc, ms);
  fflush(stdout);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  g_sms = p.multiProcessorCount;
  g_mhz = p.clockRate / 1000.f;
  cudaMalloc(&g_out, (size_t)g_sms * 16 * 128 * 4);
  run<F_RRR, NONE, 0>("R64,R64,R64", "none");
  run<F_RSR, NONE, 0>("R64,R32bc,R64", "none");
  run<F_RSU, NONE, 0>("R64,R32bc,UR", "none");
  run<F_RRU, NONE, 0>("R64,R64,UR", "none");
  run<F_SCALAR_RRR, NONE, 0>("scalar R,R,R (x2)", "none");
  run<F_SCALAR_RRU, NONE, 0>("scalar R,R,UR (x2)", "none");
#define BOTH(OP, NA, NAME)                      \
  run<F_RSR, OP, NA>("R64,R32bc,R64", NAME);    \
  run<F_RSU, OP, NA>("R64,R32bc,UR", NAME);
  BOTH(FSETP_RR, 1, "fsetp r,r")
  BOTH(FSETP_RI, 1, "fsetp r,imm")
  BOTH(FSETP_RI, 2, "fsetp r,imm")
  BOTH(IADD_IMM, 1, "iadd r,imm")
  BOTH(IADD_RR, 1, "iadd r,r")
  BOTH(FMNMX_RR, 1, "fmnmx r,r")
  BOTH(FMNMX_ABS, 1, "fmnmx r,|r|")
  BOTH(FMNMX_RR, 2, "fmnmx r,r")
  BOTH(FMNMX_RU, 1, "fmnmx r,ur (no chain)")
  BOTH(LOP3_RRI, 1, "lop3 r,r,imm")
  BOTH(LOP3_RRR, 1, "lop3 r,r,r")
  BOTH(SHF_RR, 1, "shf r,r")
  BOTH(SHF_X2, 1, "shf r,r x2 chained")
  BOTH(PRED_LOP3_T, 1, "fsetp + @p(true) or r,r")
  BOTH(PRED_LOP3_F, 1, "fsetp + @p(false) or r,r")
  BOTH(FSETP2_AND, 1, "fsetp, fsetp.and, @q(true) or")
  BOTH(MUFU_ONLY, 1, "mufu.rsq")
  BOTH(FMNMX3_RRR, 1, "fmnmx3 r,r,r")
  BOTH(FADD2_RR, 1, "fadd2 r64,r64 (+fadd)")
  BOTH(LDS128_ONLY, 1, "lds.128")
  BOTH(MAXSHFMIN, 1, "fmnmx(max) + shf + fmnmx(min |.|)")
  BOTH(MAXSHFMIN, 2, "fmnmx(max) + shf + fmnmx(min |.|)")
  BOTH(SETP2_PLOP_MIN, 1, "fsetp, fsetp.and, @q or, fmnmx3(|e|,|nt|)")
  BOTH(SETP2_PLOP_MIN, 2, "fsetp, fsetp.and, @q or, fmnmx3(|e|,|nt|)")
  return 0;
}
