// loop_bench.cu -- the inner loop of the K2 score kernel in isolation (sm_100a): K=4 candidates per
// thread as two packed pairs, 128 points resident in shared memory read as broadcast LDS.128,
// the real eval2<T>() of rsc_eval.cuh.  Variants strip the ALU-pipe work (max / funnel shift / min)
// one by one, so that the cost of the packed FP32 stream and of every integer op beside it can be
// read off as SMSP cycles per (pair of candidates x point).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../include -I../ransac.jl_b200/csrc -o loop_bench loop_bench.cu
#include <stdio.h>

#include "rsc_eval.cuh"

using namespace rsc;


// ---- explicit-order packed ops (volatile: ptxas keeps their relative order) ----------------------
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float2 v) { return *reinterpret_cast<u64*>(&v); }
__device__ __forceinline__ float2 upk(u64 v) { return *reinterpret_cast<float2*>(&v); }
__device__ __forceinline__ u64 bcast(float s) {
  u64 r;
  asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(s));
  return r;
}
__device__ __forceinline__ u64 vfma(u64 a, u64 b, u64 c) {
  u64 d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
// cell order: the point operand (slot B) or the record operand (slot A) repeats between neighbours
__device__ constexpr int kSnakeJ[8] = {0, 1, 1, 0, 0, 1, 1, 0};
__device__ constexpr int kSnakeQ[8] = {0, 0, 1, 1, 2, 2, 3, 3};

// plane, 2 pairs x 4 points, level by level in snake order; returns e (=|d|-eps) and nt per cell
__device__ __forceinline__ void plane_lockstep(const float2 (*r2)[RecN<RSC_PLANE>::n], const float* px, const float* py,
                                               const float* pz, const float* nx, const float* ny, const float* nz, float eps,
                                               float cosa, float2 (*e)[4], float2 (*nt)[4]) {
  u64 R[2][7];
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int f = 0; f < 7; ++f) R[j][f] = pk(r2[j][f]);
  u64 d[2][4], n[2][4];
  const u64 CA = bcast(cosa);
#pragma unroll
  for (int c = 0; c < 8; ++c) { const int j = kSnakeJ[c], q = kSnakeQ[c]; d[j][q] = vfma(R[j][2], bcast(pz[q]), R[j][3]); }
#pragma unroll
  for (int c = 0; c < 8; ++c) { const int j = kSnakeJ[c], q = kSnakeQ[c]; n[j][q] = vfma(R[j][6], bcast(nz[q]), CA); }
#pragma unroll
  for (int c = 0; c < 8; ++c) { const int j = kSnakeJ[c], q = kSnakeQ[c]; d[j][q] = vfma(R[j][1], bcast(py[q]), d[j][q]); }
#pragma unroll
  for (int c = 0; c < 8; ++c) { const int j = kSnakeJ[c], q = kSnakeQ[c]; n[j][q] = vfma(R[j][5], bcast(ny[q]), n[j][q]); }
#pragma unroll
  for (int c = 0; c < 8; ++c) { const int j = kSnakeJ[c], q = kSnakeQ[c]; d[j][q] = vfma(R[j][0], bcast(px[q]), d[j][q]); }
#pragma unroll
  for (int c = 0; c < 8; ++c) { const int j = kSnakeJ[c], q = kSnakeQ[c]; n[j][q] = vfma(R[j][4], bcast(nx[q]), n[j][q]); }
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      e[j][q] = add2(abs2(upk(d[j][q])), bc2(-eps));
      nt[j][q] = upk(n[j][q]);
    }
}

constexpr int kPts = 128;
constexpr int ITERS = 512;  // passes over the 128 points

// VAR: 0 = FP32 ops only (margins kept alive by an empty asm), 1 = + max, 2 = + max + shf, 3 = + max + shf + min|.|
template <int T, int VAR, int MINB>
__global__ void __launch_bounds__(128, MINB) loop_kernel(const float* __restrict__ rec, float* __restrict__ out, float eps, float cosa) {
  constexpr int NR = RecN<T>::n;
  constexpr int KP = 2;
  __shared__ __align__(16) float pts[6 * kPts];
  for (int i = threadIdx.x; i < 6 * kPts; i += blockDim.x) pts[i] = 0.01f * (float)((i * 37) % 101) - 0.5f;
  __syncthreads();
  float2 r2[KP][NR];
  for (int j = 0; j < KP; ++j)
    for (int f = 0; f < NR; ++f) {
      r2[j][f].x = rec[(f * 4 + 2 * j) * 128 + threadIdx.x];
      r2[j][f].y = rec[(f * 4 + 2 * j + 1) * 128 + threadIdx.x];
    }
  uint32_t mask[4] = {0, 0, 0, 0};
  float mabs[4] = {1e30f, 1e30f, 1e30f, 1e30f};
  float2 keep = make_float2(0.f, 0.f);
  float2 acc[KP] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll 1
    for (int i4 = 0; i4 < kPts; i4 += 4) {
      const float* gp = pts + i4;
      const float4 X = *reinterpret_cast<const float4*>(gp);
      const float4 Y = *reinterpret_cast<const float4*>(gp + kPts);
      const float4 Z = *reinterpret_cast<const float4*>(gp + 2 * kPts);
      const float4 U4 = *reinterpret_cast<const float4*>(gp + 3 * kPts);
      const float4 V = *reinterpret_cast<const float4*>(gp + 4 * kPts);
      const float4 W = *reinterpret_cast<const float4*>(gp + 5 * kPts);
      const float px[4] = {X.x, X.y, X.z, X.w}, py[4] = {Y.x, Y.y, Y.z, Y.w}, pz[4] = {Z.x, Z.y, Z.z, Z.w};
      const float nx[4] = {U4.x, U4.y, U4.z, U4.w}, ny[4] = {V.x, V.y, V.z, V.w}, nz[4] = {W.x, W.y, W.z, W.w};
      if constexpr (VAR >= 4) {
        if constexpr (T == RSC_PLANE) {
          float2 e[2][4], nt[2][4];
          plane_lockstep(r2, px, py, pz, nx, ny, nz, eps, cosa, e, nt);
#pragma unroll
          for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < KP; ++j) {
              if constexpr (VAR == 4) {
                acc[j] = __ffma2_rn(e[j][q], make_float2(1e-9f, 1e-9f), acc[j]);
                acc[j] = __ffma2_rn(nt[j][q], make_float2(1e-9f, 1e-9f), acc[j]);
              } else {
                const float2 m = max2_nan(e[j][q], nt[j][q]);
                mask[2 * j] = __funnelshift_l(__float_as_uint(m.x), mask[2 * j], 1);
                mask[2 * j + 1] = __funnelshift_l(__float_as_uint(m.y), mask[2 * j + 1], 1);
                mabs[2 * j] = fmin_nan(mabs[2 * j], fabsf(m.x));
                mabs[2 * j + 1] = fmin_nan(mabs[2 * j + 1], fabsf(m.y));
              }
            }
        }
        continue;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int j = 0; j < KP; ++j) {
          if constexpr (VAR == 0) {
            // everything up to (not including) the max: re-derive e and nt by calling eval2 and
            // keeping its result alive costs the two FMNMX; instead keep the packed margin terms
            const float2 m = eval2_terms<T>(r2[j], px[q], py[q], pz[q], nx[q], ny[q], nz[q], eps, cosa, &keep);
            acc[j] = __ffma2_rn(m, make_float2(1e-9f, 1e-9f), acc[j]);  // two extra packed ops keep e and nt alive
            acc[j] = __ffma2_rn(keep, make_float2(1e-9f, 1e-9f), acc[j]);
          } else {
            const float2 m = eval2<T>(r2[j], px[q], py[q], pz[q], nx[q], ny[q], nz[q], eps, cosa);
            if constexpr (VAR == 1) acc[j] = __ffma2_rn(m, make_float2(1e-9f, 1e-9f), acc[j]);  // one extra packed op
            if constexpr (VAR >= 2) {
              mask[2 * j] = __funnelshift_l(__float_as_uint(m.x), mask[2 * j], 1);
              mask[2 * j + 1] = __funnelshift_l(__float_as_uint(m.y), mask[2 * j + 1], 1);
            }
            if constexpr (VAR >= 3) {
              mabs[2 * j] = fmin_nan(mabs[2 * j], fabsf(m.x));
              mabs[2 * j + 1] = fmin_nan(mabs[2 * j + 1], fabsf(m.y));
            }
          }
        }
      }
    }
  }
  float s = 0;
  for (int k = 0; k < 4; ++k) s += (float)mask[k] + mabs[k];
  s += acc[0].x + acc[0].y + acc[1].x + acc[1].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int T, int VAR, int MINB>
void run(const char* tname, const float* rec, float* out, int sms, float mhz) {
  const int blocks = sms * MINB;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  loop_kernel<T, VAR, MINB><<<blocks, 128>>>(rec, out, 0.3f, 0.996f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 3; ++r) loop_kernel<T, VAR, MINB><<<blocks, 128>>>(rec, out, 0.3f, 0.996f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= 3;
  // per SMSP: MINB warps, each ITERS*kPts points x 2 pairs
  const double pair_points = (double)MINB * ITERS * kPts * 2;
  const double cyc = ms * 1e-3 * mhz * 1e6 / pair_points;
  static const char* vn[] = {"fp32 ops + 2 packed fma (no ALU)", "fp32 ops + max + 1 packed fma", "fp32 ops + max + shf", "fp32 ops + max + shf + min", "lockstep snake order: fp32 ops + 2 packed fma (no ALU)", "lockstep snake order: full"};
  printf("{\"type\": \"%s\", \"variant\": \"%s\", \"warps_per_smsp\": %d, \"smsp_cycles_per_pair_point\": %.2f, \"G_evals_s_chip\": %.0f}\n", tname,
         vn[VAR], MINB, cyc, (double)sms * 4 * mhz * 1e6 * 64 / cyc / 1e9);
  fflush(stdout);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  const float mhz = p.clockRate / 1000.f;
  float *rec, *out;
  cudaMalloc(&rec, 12 * 4 * 128 * 4);
  cudaMalloc(&out, (size_t)sms * 8 * 128 * 4);
  float h[12 * 4 * 128];
  for (int i = 0; i < 12 * 4 * 128; ++i) h[i] = 0.001f * (float)((i * 131) % 997) + 0.1f;
  cudaMemcpy(rec, h, sizeof(h), cudaMemcpyHostToDevice);
#define ALLV(T, NAME, MINB)          \
  run<T, 0, MINB>(NAME, rec, out, sms, mhz); \
  run<T, 1, MINB>(NAME, rec, out, sms, mhz); \
  run<T, 2, MINB>(NAME, rec, out, sms, mhz); \
  run<T, 3, MINB>(NAME, rec, out, sms, mhz);
  ALLV(RSC_PLANE, "plane", 4)
  run<RSC_PLANE, 4, 4>("plane", rec, out, sms, mhz);
  run<RSC_PLANE, 5, 4>("plane", rec, out, sms, mhz);
  ALLV(RSC_SPHERE, "sphere", 4)
  ALLV(RSC_CYLINDER, "cylinder", 4)
  ALLV(RSC_CONE, "cone", 4)
  ALLV(RSC_PLANE, "plane", 3)
  ALLV(RSC_CONE, "cone", 3)
  return 0;
}
