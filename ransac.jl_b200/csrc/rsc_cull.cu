// rsc_cull.cu -- whole-cloud scoring with Morton-tile culling (extension; same counts as rsc_score).
//
// The dense K2 kernel (rsc_score.cu) evaluates every (candidate, point) pair, like the reference's
// compatibles*.  Most of those pairs are far from the candidate's surface.  With the cloud in Morton
// order (rsc_octree.cu) a tile of 512 consecutive points has a small bounding sphere (c, r), and every
// distance function of the path (plane.jl:82-103, sphere.jl:163-166, cylinder.jl:209-214, the cone's
// h sin - rho cos) is 1-Lipschitz in the point, so
//       |dist(c)| > eps + r   =>   no point of the tile is within eps   =>   the tile adds nothing
// to the candidate's count.  tools/cull_estimate.py: 8 % of the pairs of config c3 survive that test.
//
// One kernel, no work lists: a CTA owns a tile (4 points per thread in registers) and walks the
// candidates in chunks of 128 -- BROAD PHASE: thread t tests candidate chunk+t against the tile sphere
// (the candidate's compiled FP32 record, one 48-byte load) and stages the survivors' records in shared
// memory; NARROW PHASE: all threads loop over the survivors (warp-uniform), evaluate their 4 points
// with the same eval<T>() forms and guard band as the dense kernel, decide in-band pairs in FP64 in the
// reference's operation order -- queued through a per-CTA staging buffer for cull_fix_kernel, one thread per
// pair (rsc_exact.cuh) -- and add the warp's count with one REDUX + atomic.
// Counts only (masks would come out in Morton order), whole cloud only.
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "rsc_eval.cuh"
#include "rsc_exact.cuh"

namespace rsc {

constexpr int kCullThreads = 128;
constexpr int kCullWarps = kCullThreads / 32;
constexpr int kCullTile = 128;                   // points of a warp's tile: one bounding sphere each
constexpr int kCullPts = kCullTile / 32;         // 4 points per lane = 2 packed pairs
constexpr int kCullChunk = 256;                  // candidates staged in shared memory per step
constexpr int kCullSQ = 1024;                    // per-CTA staging of queued pairs (one global atomic per flush, not per pair)
#ifndef RSC_CULL_MINB
#define RSC_CULL_MINB 4
#endif
constexpr int kCullMinB = RSC_CULL_MINB;  // CTAs per SM the kernel is compiled for (registers: 65536 / (128 kCullMinB))

struct CullArgs {
  PointSet ps;          // the points in Morton order (rows of n_pad floats; n_pad a multiple of 512)
  const float4* tiles;  // bounding sphere of every 128-point tile: centre, radius
  int ngroups;          // n_pad / 512: a CTA takes the four tiles of a group, one per warp
  int nranges;          // the candidates are split into nranges runs of chunks_per_range chunks
  int chunks_per_range;
  const float* rec;     // [C][kRecFields] compiled records, candidate-major
  const uint8_t* col;   // [C] column types
  const rsc_cand* cands;
  const int32_t* d_C;   // number of candidates on the device (nullptr: C)
  Thresh th;
  int C;
  int32_t* cv;  // [C] compatible real points
  int32_t* ce;  // [C] compatible enabled points
  unsigned long long* stats;  // [0] surviving (candidate, tile) pairs, [1] pairs decided in FP64
  uint32_t* work;             // dynamic work counter (items handed out)
  uint2* queue;               // in-band pairs (candidate, position) waiting for their float64 decision
  uint32_t qcap;
  uint32_t* qn;               // entries appended (may exceed qcap: the excess was decided inline)
  int inline_fp64;            // != 0: no queue, every in-band pair is decided on the spot (RSC_CULL_INLINE=1)
};

__device__ __forceinline__ int cull_count(const CullArgs& a) { return a.d_C ? min(*a.d_C, a.C) : a.C; }

__global__ void cull_compile_kernel(const rsc_cand* __restrict__ cands, int C, const int32_t* __restrict__ d_C, Thresh th, float pmax,
                                    float nmax, float* __restrict__ rec, uint8_t* __restrict__ col) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (d_C ? min(*d_C, C) : C)) return;
  const rsc_cand c = cands[i];
  const int ct = col_type(c);
  float r[kRecFields];
  compile_record(c, ct, th, pmax, nmax, r);
#pragma unroll
  for (int f = 0; f < kRecFields; ++f) rec[(size_t)i * kRecFields + f] = r[f];
  col[i] = (uint8_t)ct;
}

__global__ void morton_gather_kernel(const float* __restrict__ soa, int64_t n_pad, const uint32_t* __restrict__ perm, int64_t n,
                                     float* __restrict__ msoa) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  const bool real = i < n;
  const int64_t j = real ? (int64_t)perm[i] : 0;
#pragma unroll
  for (int f = 0; f < 6; ++f) msoa[f * n_pad + i] = real ? soa[f * n_pad + j] : 0.f;
}

// one warp per 128-point tile: centre of the bounding box, radius = largest distance of a point to it (rounded up)
__global__ void __launch_bounds__(256) tile_sphere_kernel(const float* __restrict__ X, const float* __restrict__ Y,
                                                          const float* __restrict__ Z, int64_t n, int ntiles,
                                                          float4* __restrict__ tiles) {
  const int tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (tile >= ntiles) return;
  const int64_t base = (int64_t)tile * kCullTile;
  float lo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, hi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
  for (int k = lane; k < kCullTile; k += 32) {
    const int64_t j = base + k;
    if (j < n) {
      const float p[3] = {X[j], Y[j], Z[j]};
#pragma unroll
      for (int d = 0; d < 3; ++d) lo[d] = fminf(lo[d], p[d]), hi[d] = fmaxf(hi[d], p[d]);
    }
  }
#pragma unroll
  for (int d = 0; d < 3; ++d)
#pragma unroll
    for (int s = 16; s; s >>= 1) {
      lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], s));
      hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], s));
    }
  const float cx = 0.5f * (lo[0] + hi[0]), cy = 0.5f * (lo[1] + hi[1]), cz = 0.5f * (lo[2] + hi[2]);
  float r2 = 0.f;
  for (int k = lane; k < kCullTile; k += 32) {
    const int64_t j = base + k;
    if (j < n) {
      const float dx = X[j] - cx, dy = Y[j] - cy, dz = Z[j] - cz;
      r2 = fmaxf(r2, dx * dx + dy * dy + dz * dz);
    }
  }
#pragma unroll
  for (int s = 16; s; s >>= 1) r2 = fmaxf(r2, __shfl_xor_sync(0xffffffffu, r2, s));
  if (lane == 0) {
    // NaN / Inf coordinates: an infinite radius, the tile is never culled; a tile of padding only: culled by everyone
    float r = sqrtf(r2) * 1.00001f + 1e-30f;
    if (!(r < 3.0e38f) || !(fabsf(cx) < 3.0e38f) || !(fabsf(cy) < 3.0e38f) || !(fabsf(cz) < 3.0e38f)) r = __int_as_float(0x7f800000);
    if (base >= n) r = -1.f;
    tiles[tile] = make_float4(base >= n ? 0.f : cx, base >= n ? 0.f : cy, base >= n ? 0.f : cz, r);
  }
}

// true: provably no point within eps of the candidate inside the sphere ts (NaN anywhere -> false)
__device__ __forceinline__ bool cull_far(int col, const float* r, float4 ts, float eps) {
  const float band = r[kBandField];
  const float rt = ts.w;
  float d, lim;
  if (col == RSC_PLANE) {
    d = fabsf(fmaf(r[0], ts.x, fmaf(r[1], ts.y, fmaf(r[2], ts.z, r[3]))));
    lim = eps + rt;
  } else {
    const float vx = fmaf(r[0], ts.x, r[1]), vy = fmaf(r[0], ts.y, r[2]), vz = fmaf(r[0], ts.z, r[3]);
    if (col == RSC_SPHERE) {
      d = fabsf(sqrtf(fmaf(vx, vx, fmaf(vy, vy, vz * vz))) + r[4]);
      lim = eps + rt;
    } else {
      const float h = fmaf(r[4], vx, fmaf(r[5], vy, r[6] * vz));
      const float wx = fmaf(-r[4], h, vx), wy = fmaf(-r[5], h, vy), wz = fmaf(-r[6], h, vz);
      const float rho = sqrtf(fmaf(wx, wx, fmaf(wy, wy, wz * wz)));
      if (col == RSC_CYLINDER) {
        // the reference uses the axis as given (Q8): p -> |v - a (a.v)| has Lipschitz constant
        // max(1, | |a|^2 - 1 |), which is 1 for the unit axes the fits produce
        const float a2 = fmaf(r[4], r[4], fmaf(r[5], r[5], r[6] * r[6]));
        d = fabsf(rho + r[7]);
        lim = eps + rt * fmaxf(1.f, fabsf(a2 - 1.f) * 1.0001f);
      } else {
        // the cone records are scaled by 1/cos (1/sin for wide cones) of the half angle: r8 = -eps/scale
        d = (col == kConeWide) ? fabsf(fmaf(-rho, r[7], h)) : fabsf(fmaf(h, r[7], -rho));
        lim = -r[8] * (1.f + rt / eps);
      }
    }
  }
  return d > lim * 1.0001f + 8.f * band;
}

// the points of a lane: 4 of its warp's 128-point tile as two packed pairs (points q = 2i, 2i + 1 sit at
// tile base + 32 q + lane), plus their valid / enabled bits in the order the margins' signs are collected
struct CullPoints {
  float2 x[kCullPts / 2], y[kCullPts / 2], z[kCullPts / 2], nx[kCullPts / 2], ny[kCullPts / 2], nz[kCullPts / 2];
  uint32_t valid, enabled;  // bit (kCullPts - 1 - q): point q of this lane
};

// the FP64 decision of one in-band pair; out of line so that its registers (and code) do not weigh on the FP32 loop
__device__ __noinline__ uint32_t cull_exact(const rsc_cand* __restrict__ cp, const Thresh* th, float px, float py, float pz, float nx,
                                            float ny, float nz) {
  const rsc_cand c = *cp;
  ex::ConeTrig tr{1.0, 0.0};
  if (c.type == RSC_CONE) {
    tr.ct = cos(-c.p[6] / 2);
    tr.st = sin(-c.p[6] / 2);
  }
  const ex::V3 p = {(double)px, (double)py, (double)pz};
  const ex::V3 n = {(double)nx, (double)ny, (double)nz};
  return ex::compat(c, tr, *th, p, n) ? 1u : 0u;
}

// float64 decision of one queued pair, added to the counts (the FP32 pass counted nothing for it)
__device__ __forceinline__ void cull_fix_one(const CullArgs& a, uint2 e) {
  const uint32_t cand = e.x;
  const int64_t j = (int64_t)e.y;
  const uint32_t ok = cull_exact(a.cands + cand, &a.th, a.ps.x[j], a.ps.y[j], a.ps.z[j], a.ps.nx[j], a.ps.ny[j], a.ps.nz[j]);
  if (ok) {
    atomicAdd(a.cv + cand, 1);
    if ((a.ps.enabled[j >> 5] >> (j & 31)) & 1u) atomicAdd(a.ce + cand, 1);
  }
}

// the rare continuation of cull_narrow: some margin of this lane is inside the guard band (or NaN).  Those points
// leave the sign word and go to the float64 queue (or are decided here when the staging area is full).
__device__ __noinline__ uint32_t cull_ambiguous(const CullArgs* a, float band, int cand, float m0, float m1, float m2, float m3,
                                                uint32_t acc, uint32_t valid, uint32_t enabled, uint32_t j0, uint2* sq, uint32_t* sqn,
                                                int* nexact) {
  const float m[kCullPts] = {m0, m1, m2, m3};
#pragma unroll
  for (int q = 0; q < kCullPts; ++q) {
    const uint32_t bit = 1u << (kCullPts - 1 - q);
    if (fabsf(m[q]) > band) continue;  // sure (false for NaN)
    acc &= ~bit;
    if (!(valid & bit)) continue;  // padding counts for nothing
    const uint32_t j = j0 + 32u * q;
    const uint32_t slot = a->inline_fp64 ? (uint32_t)kCullSQ : atomicAdd(sqn, 1u);
    if (slot < (uint32_t)kCullSQ) {
      sq[slot] = make_uint2((uint32_t)cand, j);
    } else {  // staging full (e.g. a needle cone: every pair is in-band): decide on the spot
      const uint32_t ok = cull_exact(a->cands + cand, &a->th, a->ps.x[j], a->ps.y[j], a->ps.z[j], a->ps.nx[j], a->ps.ny[j], a->ps.nz[j]);
      if (ok) acc |= bit;
      ++*nexact;
    }
  }
  return acc;
}

// one surviving (candidate, tile) pair: the warp's 128 points against the record at sr (shared memory, broadcast
// reads).  Returns this lane's counts: compatible real points | compatible enabled points << 16.
template <int T>
__device__ __forceinline__ uint32_t cull_narrow(const CullArgs& a, const float* __restrict__ sr, int cand, const CullPoints& P,
                                                uint32_t j0, uint2* sq, uint32_t* sqn, int* nexact) {
  float r[12];
  {
    const float4* s4 = reinterpret_cast<const float4*>(sr);
    const float4 a0 = s4[0], a1 = s4[1], a2 = s4[2];
    r[0] = a0.x, r[1] = a0.y, r[2] = a0.z, r[3] = a0.w, r[4] = a1.x, r[5] = a1.y, r[6] = a1.z, r[7] = a1.w;
    r[8] = a2.x, r[9] = a2.y, r[10] = a2.z, r[11] = a2.w;
  }
  const float band = r[kBandField];
  constexpr int PT = public_type(T);
  const float eps = a.th.eps[PT], cosa = a.th.cosa[PT];
  const float2 m0 = evalp<T>(r, P.x[0], P.y[0], P.z[0], P.nx[0], P.ny[0], P.nz[0], eps, cosa);
  const float2 m1 = evalp<T>(r, P.x[1], P.y[1], P.z[1], P.nx[1], P.ny[1], P.nz[1], eps, cosa);
  // the inlier bit is the sign of the margin: funnel-shifted into a word, point 0 ends up in bit 3
  uint32_t acc = __funnelshift_l(__float_as_uint(m0.x), 0u, 1);
  acc = __funnelshift_l(__float_as_uint(m0.y), acc, 1);
  acc = __funnelshift_l(__float_as_uint(m1.x), acc, 1);
  acc = __funnelshift_l(__float_as_uint(m1.y), acc, 1);
  const float amin = fmin_nan(fmin_nan(fabsf(m0.x), fabsf(m0.y)), fmin_nan(fabsf(m1.x), fabsf(m1.y)));
  if (!(amin > band)) acc = cull_ambiguous(&a, band, cand, m0.x, m0.y, m1.x, m1.y, acc, P.valid, P.enabled, j0, sq, sqn, nexact);
  return (uint32_t)__popc(acc & P.valid) | ((uint32_t)__popc(acc & P.enabled) << 16);
}

__global__ void __launch_bounds__(kCullThreads, kCullMinB) cull_score_kernel(const __grid_constant__ CullArgs a) {
  __shared__ __align__(16) float srec[kCullChunk][kRecFields];
  __shared__ uint8_t scol[kCullChunk];
  __shared__ uint2 sq[kCullSQ];
  __shared__ uint32_t sqn, sbase, sitem;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = cull_count(a);
  if (tid == 0) sqn = 0;
  __syncthreads();
  // append the staged pairs to the global queue (entries beyond its capacity are decided here)
  auto flush = [&]() {
    const uint32_t cnt = sqn < (uint32_t)kCullSQ ? sqn : (uint32_t)kCullSQ;
    if (tid == 0) sbase = atomicAdd(a.qn, cnt);
    __syncthreads();
    for (uint32_t i = tid; i < cnt; i += kCullThreads) {
      const uint32_t g = sbase + i;
      if (g < a.qcap)
        a.queue[g] = sq[i];
      else
        cull_fix_one(a, sq[i]);
    }
    __syncthreads();
    if (tid == 0) sqn = 0;
    __syncthreads();
  };
  unsigned long long n_surv = 0;
  int n_exact = 0;
  const uint32_t nitems = (uint32_t)a.ngroups * (uint32_t)a.nranges;
  for (;;) {
    __syncthreads();  // everybody is done with sitem / srec of the previous item
    if (tid == 0) sitem = atomicAdd(a.work, 1u);
    __syncthreads();
    const uint32_t item = sitem;
    if (item >= nitems) break;
    const int group = (int)(item % (uint32_t)a.ngroups), range = (int)(item / (uint32_t)a.ngroups);
    const int tile = group * kCullWarps + warp;
    const int64_t base = (int64_t)tile * kCullTile;
    CullPoints P;
    {
      float px[kCullPts], py[kCullPts], pz[kCullPts], qx[kCullPts], qy[kCullPts], qz[kCullPts];
      P.valid = 0, P.enabled = 0;
#pragma unroll
      for (int q = 0; q < kCullPts; ++q) {
        const int64_t j = base + q * 32 + lane;
        px[q] = a.ps.x[j], py[q] = a.ps.y[j], pz[q] = a.ps.z[j];
        qx[q] = a.ps.nx[j], qy[q] = a.ps.ny[j], qz[q] = a.ps.nz[j];
        const uint32_t bit = 1u << (kCullPts - 1 - q);
        if (j < a.ps.n) P.valid |= bit;
        if ((a.ps.enabled[j >> 5] >> lane) & 1u) P.enabled |= bit;
      }
      P.enabled &= P.valid;
#pragma unroll
      for (int i = 0; i < kCullPts / 2; ++i) {
        P.x[i] = make_float2(px[2 * i], px[2 * i + 1]), P.y[i] = make_float2(py[2 * i], py[2 * i + 1]);
        P.z[i] = make_float2(pz[2 * i], pz[2 * i + 1]), P.nx[i] = make_float2(qx[2 * i], qx[2 * i + 1]);
        P.ny[i] = make_float2(qy[2 * i], qy[2 * i + 1]), P.nz[i] = make_float2(qz[2 * i], qz[2 * i + 1]);
      }
    }
    const float4 ts = a.tiles[tile];
    const uint32_t j0 = (uint32_t)(base + lane);
    const int c_lo = range * a.chunks_per_range * kCullChunk;
    const int c_hi = min(C, c_lo + a.chunks_per_range * kCullChunk);
    for (int c0 = c_lo; c0 < c_hi; c0 += kCullChunk) {
      if (c0 != c_lo) __syncthreads();  // the previous chunk's records are no longer read
      if (sqn > (uint32_t)kCullSQ / 2) flush();  // uniform: every thread reads the same sqn after a barrier
      // ---- stage the chunk's records (coalesced 16-byte copies) ----
      const int nc = min(kCullChunk, c_hi - c0);
      {
        const float4* g = reinterpret_cast<const float4*>(a.rec + (size_t)c0 * kRecFields);
        float4* s = reinterpret_cast<float4*>(&srec[0][0]);
        for (int i = tid; i < nc * (kRecFields / 4); i += kCullThreads) s[i] = g[i];
        for (int i = tid; i < nc; i += kCullThreads) scol[i] = a.col[c0 + i];
      }
      __syncthreads();
      // ---- per warp: broad phase (a candidate per lane against the tile's sphere), then the survivors ----
#pragma unroll 1
      for (int k0 = 0; k0 < nc; k0 += 32) {
        const int s = k0 + lane;
        bool keep = false;
        if (s < nc) {
          float r[kRecFields];
          const float4* s4 = reinterpret_cast<const float4*>(srec[s]);
          const float4 r0 = s4[0], r1 = s4[1], r2 = s4[2];
          r[0] = r0.x, r[1] = r0.y, r[2] = r0.z, r[3] = r0.w, r[4] = r1.x, r[5] = r1.y, r[6] = r1.z, r[7] = r1.w;
          r[8] = r2.x, r[9] = r2.y, r[10] = r2.z, r[11] = r2.w;
          const int ct = scol[s];
          keep = !cull_far(ct, r, ts, a.th.eps[public_type(ct)]);
        }
        uint32_t mm = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) n_surv += __popc(mm);
        while (mm) {
          const int b = __ffs(mm) - 1;
          mm &= mm - 1;
          const int sv = k0 + b;
          const int cand = c0 + sv;
          uint32_t cnt;
          switch (scol[sv]) {
            case RSC_PLANE:
              cnt = cull_narrow<RSC_PLANE>(a, srec[sv], cand, P, j0, sq, &sqn, &n_exact);
              break;
            case RSC_SPHERE:
              cnt = cull_narrow<RSC_SPHERE>(a, srec[sv], cand, P, j0, sq, &sqn, &n_exact);
              break;
            case RSC_CYLINDER:
              cnt = cull_narrow<RSC_CYLINDER>(a, srec[sv], cand, P, j0, sq, &sqn, &n_exact);
              break;
            case kConeWide:
              cnt = cull_narrow<kConeWide>(a, srec[sv], cand, P, j0, sq, &sqn, &n_exact);
              break;
            default:
              cnt = cull_narrow<RSC_CONE>(a, srec[sv], cand, P, j0, sq, &sqn, &n_exact);
              break;
          }
          cnt = __reduce_add_sync(0xffffffffu, cnt);
          if (lane == 0 && cnt) {
            if (cnt & 0xffffu) atomicAdd(a.cv + cand, (int)(cnt & 0xffffu));
            if (cnt >> 16) atomicAdd(a.ce + cand, (int)(cnt >> 16));
          }
        }
      }
    }
  }
  __syncthreads();
  if (sqn) flush();
  n_exact = __reduce_add_sync(0xffffffffu, n_exact);
  if (lane == 0 && n_exact) atomicAdd(a.stats + 1, (unsigned long long)n_exact);
  if (lane == 0 && n_surv) atomicAdd(a.stats, n_surv);
}

// float64 decisions of the queued pairs: one thread per pair, all lanes busy
__global__ void __launch_bounds__(256) cull_fix_kernel(const __grid_constant__ CullArgs a) {
  const uint32_t n = *a.qn < a.qcap ? *a.qn : a.qcap;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) cull_fix_one(a, a.queue[i]);
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(a.stats + 1, (unsigned long long)*a.qn);
}

__global__ void cull_policy_kernel(const rsc_cand* __restrict__ cands, int C, const int32_t* __restrict__ cv,
                                   const int32_t* __restrict__ ce, uint32_t honour_enabled, int32_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C) return;
  out[i] = ((honour_enabled >> cands[i].type) & 1u) ? ce[i] : cv[i];
}

// Morton-ordered copy of the cloud + tile spheres, kept with the flattened octree (dropped with it)
static int32_t cull_prepare(rsc_cloud* cloud, cudaStream_t st) {
  rsc_ctx* ctx = cloud->ctx;
  rsc_cells& c = cloud->cells;
  if (c.msoa) return RSC_OK;
  const int ntiles = (int)(cloud->n_pad / kCullTile);
  RSC_CUDA(ctx, cudaMalloc(&c.msoa, (size_t)6 * cloud->n_pad * sizeof(float)));
  cudaError_t e = cudaMalloc(&c.tiles, (size_t)ntiles * sizeof(float4));
  if (e != cudaSuccess) {
    cudaFree(c.msoa);
    c.msoa = nullptr;
    return fail_cuda(ctx, e, "score_culled: cudaMalloc");
  }
  morton_gather_kernel<<<(unsigned)((cloud->n_pad + 255) / 256), 256, 0, st>>>(cloud->soa, cloud->n_pad, c.perm, cloud->n, c.msoa);
  RSC_CUDA(ctx, cudaGetLastError());
  tile_sphere_kernel<<<(ntiles + 7) / 8, 256, 0, st>>>(c.msoa, c.msoa + cloud->n_pad, c.msoa + 2 * cloud->n_pad, cloud->n, ntiles,
                                                       reinterpret_cast<float4*>(c.tiles));
  RSC_CUDA(ctx, cudaGetLastError());
  return RSC_OK;
}

// bounding spheres of the 128-point tiles of a point set that already is in a spatially coherent order
int32_t cull_tile_spheres(rsc_ctx* ctx, const PointSet& ps, float4* tiles, cudaStream_t st) {
  const int ntiles = (int)(ps.n_pad / kCullTile);
  if (ntiles == 0) return RSC_OK;
  tile_sphere_kernel<<<(ntiles + 7) / 8, 256, 0, st>>>(ps.x, ps.y, ps.z, ps.n, ntiles, tiles);
  RSC_CUDA(ctx, cudaGetLastError());
  return RSC_OK;
}

// Enqueue the culled scoring of up to C_cap candidates on the device (their number may itself live on the
// device: d_C) against the Morton-ordered point set ps with tile spheres `tiles`.  cv / ce [C_cap] receive the
// compatible real / enabled points of every candidate.  No synchronisation; d_stats (2 x u64, optional) gets
// the surviving (candidate, tile) pairs and the pairs decided in float64.
int32_t cull_enqueue(rsc_ctx* ctx, const rsc_cloud* cloud, const PointSet& ps, const float4* tiles, const Thresh& th,
                     const rsc_cand* d_cands, int C_cap, const int32_t* d_C, int32_t* cv, int32_t* ce, unsigned long long* d_stats,
                     cudaStream_t st) {
  if (C_cap <= 0 || ps.n_pad <= 0) return RSC_OK;
  // scratch: [rec][col][stats 2 x u64][qn, work][queue]
  const size_t o_col = (size_t)C_cap * kRecFields * sizeof(float);
  const size_t o_ctr = (o_col + (size_t)C_cap + 15) / 16 * 16;
  const size_t o_queue = o_ctr + 32;
  int64_t qcap = (int64_t)((double)C_cap * (double)ps.n / 4096.0);
  qcap = qcap < (1 << 16) ? (1 << 16) : qcap > (8 << 20) ? (8 << 20) : qcap;
  RSC_CUDA(ctx, ctx->cullbuf.ensure(o_queue + (size_t)qcap * sizeof(uint2)));
  char* b = ctx->cullbuf.as<char>();
  float* d_rec = reinterpret_cast<float*>(b);
  uint8_t* d_col = reinterpret_cast<uint8_t*>(b + o_col);
  unsigned long long* ctr = reinterpret_cast<unsigned long long*>(b + o_ctr);
  RSC_CUDA(ctx, cudaMemsetAsync(ctr, 0, 32, st));
  RSC_CUDA(ctx, cudaMemsetAsync(cv, 0, (size_t)C_cap * sizeof(int32_t), st));
  RSC_CUDA(ctx, cudaMemsetAsync(ce, 0, (size_t)C_cap * sizeof(int32_t), st));
  cull_compile_kernel<<<(C_cap + 127) / 128, 128, 0, st>>>(d_cands, C_cap, d_C, th, cloud->pmax, cloud->nmax, d_rec, d_col);
  RSC_CUDA(ctx, cudaGetLastError());
  CullArgs a;
  a.ps = ps;
  a.tiles = tiles;
  a.ngroups = (int)(ps.n_pad / (kCullTile * kCullWarps));
  const int chunks = (C_cap + kCullChunk - 1) / kCullChunk;
  // enough work items to balance the persistent grid: split the candidates when the point set is small
  const int cap = ctx->sm_count * kCullMinB;
  int nranges = 1;
  while (nranges < chunks && (int64_t)a.ngroups * nranges < 8ll * cap) nranges *= 2;
  a.chunks_per_range = (chunks + nranges - 1) / nranges;
  a.nranges = (chunks + a.chunks_per_range - 1) / a.chunks_per_range;
  a.rec = d_rec, a.col = d_col, a.cands = d_cands, a.d_C = d_C;
  a.th = th;
  a.C = C_cap;
  a.cv = cv, a.ce = ce;
  a.stats = d_stats ? d_stats : ctr;
  a.qn = reinterpret_cast<uint32_t*>(ctr + 2);
  a.work = a.qn + 1;
  a.queue = reinterpret_cast<uint2*>(b + o_queue);
  a.qcap = (uint32_t)qcap;
  // in-band pairs: queued for cull_fix_kernel (default) or decided on the spot (RSC_CULL_INLINE=1) -- both
  // validated against the dense path (tests/test_cull_gpu.py)
  a.inline_fp64 = getenv("RSC_CULL_INLINE") ? atoi(getenv("RSC_CULL_INLINE")) : 0;
  const int64_t items = (int64_t)a.ngroups * a.nranges;
  const int grid = (int)(items < cap ? items : cap);
  cull_score_kernel<<<grid, kCullThreads, 0, st>>>(a);
  RSC_CUDA(ctx, cudaGetLastError());
  cull_fix_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(a);
  RSC_CUDA(ctx, cudaGetLastError());
  return RSC_OK;
}

}  // namespace rsc

using namespace rsc;

extern "C" int32_t rsc_score_culled(rsc_cloud* cloud, const rsc_params* params, const rsc_cand* cands, int32_t C, int32_t* counts,
                                    int64_t* pairs_total, int64_t* pairs_survived, double* kernel_ms) {
  if (!cloud) return RSC_E_ARG;
  rsc_ctx* ctx = cloud->ctx;
  if (!params) return fail(ctx, RSC_E_ARG, "score_culled: params is null");
  if (C < 0 || (C > 0 && (!cands || !counts))) return fail(ctx, RSC_E_ARG, "score_culled: null candidates/counts");
  for (int t = 0; t < RSC_NTYPES; ++t)
    if (!(params->eps[t] == params->eps[t]) || !(params->alpha[t] == params->alpha[t]))
      return fail(ctx, RSC_E_ARG, "params: NaN threshold");
  for (int i = 0; i < C; ++i)
    if (cands[i].type < 0 || cands[i].type >= RSC_NTYPES) return fail(ctx, RSC_E_ARG, "score_culled: unknown shape type");
  if (pairs_total) *pairs_total = 0;
  if (pairs_survived) *pairs_survived = 0;
  if (kernel_ms) *kernel_ms = 0.0;
  if (C == 0) return RSC_OK;
  RSC_CUDA(ctx, cudaSetDevice(ctx->device));
  if (int32_t rc = cloud_ready(cloud)) return rc;
  if (cloud->cells.nlevels == 0) return fail(ctx, RSC_E_STATE, "score_culled: needs rsc_cloud_build_cells (Morton order) first");
  cudaStream_t st = ctx->stream;
  if (int32_t rc = cull_prepare(cloud, st)) return rc;
  if (int32_t rc = cells_refresh_enabled(cloud, st)) return rc;
  const Thresh th = make_thresh(params);
  // scratch: [cands][cv][ce][policy][stats]
  const size_t o_cv = (size_t)C * sizeof(rsc_cand);
  const size_t o_stats = (o_cv + (size_t)3 * C * sizeof(int32_t) + 15) / 16 * 16;
  RSC_CUDA(ctx, ctx->misc.ensure(o_stats + 16));
  char* b = ctx->misc.as<char>();
  rsc_cand* d_c = reinterpret_cast<rsc_cand*>(b);
  int32_t* d_cv = reinterpret_cast<int32_t*>(b + o_cv);
  unsigned long long* d_stats = reinterpret_cast<unsigned long long*>(b + o_stats);
  RSC_CUDA(ctx, cudaMemcpyAsync(d_c, cands, (size_t)C * sizeof(rsc_cand), cudaMemcpyHostToDevice, st));
  RSC_CUDA(ctx, cudaMemsetAsync(d_stats, 0, 16, st));
  PointSet ps;
  const float* m = cloud->cells.msoa;
  ps.x = m, ps.y = m + cloud->n_pad, ps.z = m + 2 * cloud->n_pad;
  ps.nx = m + 3 * cloud->n_pad, ps.ny = m + 4 * cloud->n_pad, ps.nz = m + 5 * cloud->n_pad;
  ps.enabled = cloud->cells.en_sorted, ps.valid = nullptr;
  ps.n = cloud->n, ps.n_pad = cloud->n_pad;
  RSC_CUDA(ctx, cudaEventRecord(ctx->evk0, st));
  if (int32_t rc = cull_enqueue(ctx, cloud, ps, reinterpret_cast<const float4*>(cloud->cells.tiles), th, d_c, C, nullptr, d_cv, d_cv + C,
                                d_stats, st))
    return rc;
  RSC_CUDA(ctx, cudaEventRecord(ctx->evk1, st));
  cull_policy_kernel<<<(C + 255) / 256, 256, 0, st>>>(d_c, C, d_cv, d_cv + C, th.honour_enabled, d_cv + 2 * (size_t)C);
  RSC_CUDA(ctx, cudaGetLastError());
  unsigned long long hs[2] = {0, 0};
  RSC_CUDA(ctx, cudaMemcpyAsync(counts, d_cv + 2 * (size_t)C, (size_t)C * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  RSC_CUDA(ctx, cudaMemcpyAsync(hs, d_stats, sizeof(hs), cudaMemcpyDeviceToHost, st));
  RSC_CUDA(ctx, cudaStreamSynchronize(st));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, ctx->evk0, ctx->evk1);
  if (kernel_ms) *kernel_ms = ms;
  if (pairs_total) *pairs_total = (int64_t)C * (cloud->n_pad / kCullTile);
  if (pairs_survived) *pairs_survived = (int64_t)hs[0];
  ctx->stats.evals += (int64_t)C * cloud->n;
  ctx->stats.cands_scored += C;
  ctx->stats.exact_pairs += (int64_t)hs[1];
  return RSC_OK;
}
