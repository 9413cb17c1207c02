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
constexpr int kCullPts = kTile / kCullThreads;  // 4 points per thread

struct CullArgs {
  const float* msoa;  // Morton-ordered SoA copy of the cloud, rows of n_pad floats
  int64_t n, n_pad;
  const uint32_t* en;   // pc.isenabled in Morton order
  const float4* tiles;  // bounding sphere per tile: centre, radius
  int ntiles;
  const float* rec;     // [C][kRecFields] compiled records, candidate-major
  const uint8_t* col;   // [C] column types
  const rsc_cand* cands;
  const double* trig;   // [2C] cos/sin(-opang/2) for the FP64 cone
  Thresh th;
  int C;
  int32_t* cv;  // [C] compatible real points
  int32_t* ce;  // [C] compatible enabled points
  unsigned long long* stats;  // [0] surviving (candidate, tile) pairs, [1] pairs decided in FP64
  uint2* queue;               // in-band pairs (candidate, Morton position) waiting for their float64 decision
  uint32_t qcap;
  uint32_t* qn;               // entries appended (may exceed qcap: the excess was decided inline)
  int inline_fp64;            // != 0: no queue, every in-band pair is decided on the spot (RSC_CULL_INLINE=1)
};

constexpr int kCullSQ = 1024;  // per-CTA staging of queued pairs (one global atomic per flush, not per pair)

__global__ void cull_compile_kernel(const rsc_cand* __restrict__ cands, int C, Thresh th, float pmax, float nmax,
                                    float* __restrict__ rec, uint8_t* __restrict__ col) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C) return;
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

// one warp per tile: centre of the bounding box, radius = largest distance of a point to it (rounded up)
__global__ void __launch_bounds__(256) tile_sphere_kernel(const float* __restrict__ msoa, int64_t n, int64_t n_pad, int ntiles,
                                                          float4* __restrict__ tiles) {
  const int tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (tile >= ntiles) return;
  const int64_t base = (int64_t)tile * kTile;
  const float* X = msoa;
  const float* Y = msoa + n_pad;
  const float* Z = msoa + 2 * n_pad;
  float lo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, hi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
  for (int k = lane; k < kTile; k += 32) {
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
  for (int k = lane; k < kTile; k += 32) {
    const int64_t j = base + k;
    if (j < n) {
      const float dx = X[j] - cx, dy = Y[j] - cy, dz = Z[j] - cz;
      r2 = fmaxf(r2, dx * dx + dy * dy + dz * dz);
    }
  }
#pragma unroll
  for (int s = 16; s; s >>= 1) r2 = fmaxf(r2, __shfl_xor_sync(0xffffffffu, r2, s));
  if (lane == 0) {
    // NaN / Inf coordinates: an infinite radius, the tile is never culled
    float r = sqrtf(r2) * 1.00001f + 1e-30f;
    if (!(r < 3.0e38f) || !(fabsf(cx) < 3.0e38f) || !(fabsf(cy) < 3.0e38f) || !(fabsf(cz) < 3.0e38f)) r = __int_as_float(0x7f800000);
    tiles[tile] = make_float4(cx, cy, cz, r);
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

struct CullPoints {
  float px[kCullPts], py[kCullPts], pz[kCullPts], nx[kCullPts], ny[kCullPts], nz[kCullPts];
  uint32_t valid, enabled;  // bit q: point q of this thread
};

// the FP64 decision of one in-band pair; kept out of line so that its registers (and code) do not
// weigh on the FP32 loop -- the kernel's occupancy is set by the narrow phase, not by this path
__device__ __noinline__ uint32_t cull_exact(const rsc_cand* __restrict__ cp, const double* __restrict__ trig, const Thresh* th,
                                            float px, float py, float pz, float nx, float ny, float nz) {
  const rsc_cand c = *cp;
  const ex::ConeTrig tr = {trig[0], trig[1]};
  const ex::V3 p = {(double)px, (double)py, (double)pz};
  const ex::V3 n = {(double)nx, (double)ny, (double)nz};
  return ex::compat(c, tr, *th, p, n) ? 1u : 0u;
}

// float64 decision of one queued pair, added to the counts (the FP32 pass counted nothing for it)
__device__ __forceinline__ void cull_fix_one(const CullArgs& a, uint2 e) {
  const uint32_t cand = e.x;
  const int64_t j = (int64_t)e.y;
  const float* X = a.msoa;
  const uint32_t ok = cull_exact(a.cands + cand, a.trig + 2 * cand, &a.th, X[j], X[a.n_pad + j], X[2 * a.n_pad + j], X[3 * a.n_pad + j],
                                 X[4 * a.n_pad + j], X[5 * a.n_pad + j]);
  if (ok) {
    atomicAdd(a.cv + cand, 1);
    if ((a.en[j >> 5] >> (j & 31)) & 1u) atomicAdd(a.ce + cand, 1);
  }
}

template <int T>
__device__ __forceinline__ void cull_narrow(const CullArgs& a, const float* __restrict__ sr, int cand, const CullPoints& P, int& cv,
                                            int& ce, int& nexact, uint32_t jbase, uint2* sq, uint32_t* sqn) {
  float r[RecN<T>::n];
#pragma unroll
  for (int i = 0; i < RecN<T>::n; ++i) r[i] = sr[i];
  const float band = sr[kBandField];
  constexpr int PT = public_type(T);
  const float eps = a.th.eps[PT], cosa = a.th.cosa[PT];
  uint32_t amb = 0;
#pragma unroll
  for (int q = 0; q < kCullPts; ++q) {
    const float m = eval<T>(r, P.px[q], P.py[q], P.pz[q], P.nx[q], P.ny[q], P.nz[q], eps, cosa);
    const bool sure = fabsf(m) > band;  // false for NaN
    const uint32_t ok = (sure && m < 0.f) ? 1u : 0u;
    cv += (int)(ok & (P.valid >> q));
    ce += (int)(ok & (P.enabled >> q));
    amb |= (sure ? 0u : 1u) << q;
  }
  amb &= P.valid;  // padding points count for nothing
  if (amb) {       // inside the FP32 guard band: queued for the reference's float64 decision
#pragma unroll
    for (int q = 0; q < kCullPts; ++q)
      if ((amb >> q) & 1u) {
        const uint32_t slot = a.inline_fp64 ? (uint32_t)kCullSQ : atomicAdd(sqn, 1u);
        if (slot < (uint32_t)kCullSQ) {
          sq[slot] = make_uint2((uint32_t)cand, jbase + (uint32_t)(q * kCullThreads));
        } else {  // staging full (e.g. a needle cone: every pair is in-band): decide on the spot
          const uint32_t ok = cull_exact(a.cands + cand, a.trig + 2 * cand, &a.th, P.px[q], P.py[q], P.pz[q], P.nx[q], P.ny[q], P.nz[q]);
          cv += (int)ok;
          ce += (int)(ok & (P.enabled >> q));
          ++nexact;
        }
      }
  }
}

__global__ void __launch_bounds__(kCullThreads, 4) cull_score_kernel(const __grid_constant__ CullArgs a) {
  __shared__ __align__(16) float srec[kCullThreads][kRecFields];
  __shared__ uint8_t scol[kCullThreads];
  __shared__ uint32_t surv[kCullThreads / 32];
  __shared__ uint2 sq[kCullSQ];
  __shared__ uint32_t sqn, sbase;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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
  const float* X = a.msoa;
  for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
    CullPoints P;
    P.valid = 0, P.enabled = 0;
    const int64_t base = (int64_t)tile * kTile;
#pragma unroll
    for (int q = 0; q < kCullPts; ++q) {
      const int64_t j = base + q * kCullThreads + tid;
      P.px[q] = X[j], P.py[q] = X[a.n_pad + j], P.pz[q] = X[2 * a.n_pad + j];
      P.nx[q] = X[3 * a.n_pad + j], P.ny[q] = X[4 * a.n_pad + j], P.nz[q] = X[5 * a.n_pad + j];
      P.valid |= (j < a.n ? 1u : 0u) << q;
      P.enabled |= ((a.en[j >> 5] >> (j & 31)) & 1u) << q;
    }
    P.enabled &= P.valid;
    const float4 ts = a.tiles[tile];
    const uint32_t jbase = (uint32_t)(base + tid);
    for (int c0 = 0; c0 < a.C; c0 += kCullThreads) {
      // ---- broad phase: one candidate per thread against the tile sphere ----
      const int ci = c0 + tid;
      bool keep = false;
      if (ci < a.C) {
        float r[kRecFields];
        const float4* g = reinterpret_cast<const float4*>(a.rec + (size_t)ci * kRecFields);
        const float4 r0 = g[0], r1 = g[1], r2 = g[2];
        r[0] = r0.x, r[1] = r0.y, r[2] = r0.z, r[3] = r0.w, r[4] = r1.x, r[5] = r1.y, r[6] = r1.z, r[7] = r1.w;
        r[8] = r2.x, r[9] = r2.y, r[10] = r2.z, r[11] = r2.w;
        const int ct = a.col[ci];
        keep = !cull_far(ct, r, ts, a.th.eps[public_type(ct)]);
        if (keep) {
          float4* s = reinterpret_cast<float4*>(srec[tid]);
          s[0] = r0, s[1] = r1, s[2] = r2;
          scol[tid] = (uint8_t)ct;
        }
      }
      const uint32_t m = __ballot_sync(0xffffffffu, keep);
      if (lane == 0) surv[warp] = m;
      __syncthreads();
      // ---- narrow phase: every thread evaluates its 4 points against each survivor ----
#pragma unroll 1
      for (int w = 0; w < kCullThreads / 32; ++w) {
        uint32_t mm = surv[w];
        if (tid == 0) n_surv += __popc(mm);
        while (mm) {
          const int b = __ffs(mm) - 1;
          mm &= mm - 1;
          const int s = w * 32 + b;
          const int cand = c0 + s;
          int cv = 0, ce = 0;
          switch (scol[s]) {
            case RSC_PLANE:
              cull_narrow<RSC_PLANE>(a, srec[s], cand, P, cv, ce, n_exact, jbase, sq, &sqn);
              break;
            case RSC_SPHERE:
              cull_narrow<RSC_SPHERE>(a, srec[s], cand, P, cv, ce, n_exact, jbase, sq, &sqn);
              break;
            case RSC_CYLINDER:
              cull_narrow<RSC_CYLINDER>(a, srec[s], cand, P, cv, ce, n_exact, jbase, sq, &sqn);
              break;
            case kConeWide:
              cull_narrow<kConeWide>(a, srec[s], cand, P, cv, ce, n_exact, jbase, sq, &sqn);
              break;
            default:
              cull_narrow<RSC_CONE>(a, srec[s], cand, P, cv, ce, n_exact, jbase, sq, &sqn);
              break;
          }
          cv = __reduce_add_sync(0xffffffffu, cv);
          ce = __reduce_add_sync(0xffffffffu, ce);
          if (lane == 0) {
            if (cv) atomicAdd(a.cv + cand, cv);
            if (ce) atomicAdd(a.ce + cand, ce);
          }
        }
      }
      __syncthreads();  // the next chunk overwrites srec / surv
      if (sqn > (uint32_t)kCullSQ / 2) flush();  // uniform: every thread reads the same sqn after the barrier
    }
  }
  if (sqn) flush();
  n_exact = __reduce_add_sync(0xffffffffu, n_exact);
  if (lane == 0 && n_exact) atomicAdd(a.stats + 1, (unsigned long long)n_exact);
  if (tid == 0 && n_surv) atomicAdd(a.stats, n_surv);
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
  const int ntiles = (int)(cloud->n_pad / kTile);
  RSC_CUDA(ctx, cudaMalloc(&c.msoa, (size_t)6 * cloud->n_pad * sizeof(float)));
  cudaError_t e = cudaMalloc(&c.tiles, (size_t)ntiles * sizeof(float4));
  if (e != cudaSuccess) {
    cudaFree(c.msoa);
    c.msoa = nullptr;
    return fail_cuda(ctx, e, "score_culled: cudaMalloc");
  }
  morton_gather_kernel<<<(unsigned)((cloud->n_pad + 255) / 256), 256, 0, st>>>(cloud->soa, cloud->n_pad, c.perm, cloud->n, c.msoa);
  RSC_CUDA(ctx, cudaGetLastError());
  tile_sphere_kernel<<<(ntiles + 7) / 8, 256, 0, st>>>(c.msoa, cloud->n, cloud->n_pad, ntiles, reinterpret_cast<float4*>(c.tiles));
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
  std::vector<double> trig((size_t)2 * C, 0.0);
  for (int i = 0; i < C; ++i)
    if (cands[i].type == RSC_CONE) trig[2 * i] = cos(-cands[i].p[6] / 2), trig[2 * i + 1] = sin(-cands[i].p[6] / 2);
  // scratch: [cands][trig][rec][cv][ce][policy][stats][col]
  const size_t o_trig = (size_t)C * sizeof(rsc_cand);
  const size_t o_rec = o_trig + (size_t)2 * C * sizeof(double);
  const size_t o_cv = o_rec + (size_t)C * kRecFields * sizeof(float);
  const size_t o_stats = o_cv + (size_t)3 * C * sizeof(int32_t);
  const size_t o_stats_al = (o_stats + 15) / 16 * 16;
  const size_t o_col = o_stats_al + 32;  // stats[2], qn
  int64_t qcap = (int64_t)C * cloud->n / 4096;
  qcap = qcap < (1 << 16) ? (1 << 16) : qcap > (8 << 20) ? (8 << 20) : qcap;
  const size_t o_queue = (o_col + (size_t)C + 15) / 16 * 16;
  RSC_CUDA(ctx, ctx->cullbuf.ensure(o_queue + (size_t)qcap * sizeof(uint2)));
  char* b = ctx->cullbuf.as<char>();
  rsc_cand* d_c = reinterpret_cast<rsc_cand*>(b);
  double* d_trig = reinterpret_cast<double*>(b + o_trig);
  float* d_rec = reinterpret_cast<float*>(b + o_rec);
  int32_t* d_cv = reinterpret_cast<int32_t*>(b + o_cv);
  unsigned long long* d_stats = reinterpret_cast<unsigned long long*>(b + o_stats_al);
  uint8_t* d_col = reinterpret_cast<uint8_t*>(b + o_col);
  RSC_CUDA(ctx, cudaMemcpyAsync(d_c, cands, (size_t)C * sizeof(rsc_cand), cudaMemcpyHostToDevice, st));
  RSC_CUDA(ctx, cudaMemcpyAsync(d_trig, trig.data(), (size_t)2 * C * sizeof(double), cudaMemcpyHostToDevice, st));
  RSC_CUDA(ctx, cudaMemsetAsync(d_cv, 0, (size_t)3 * C * sizeof(int32_t), st));
  RSC_CUDA(ctx, cudaMemsetAsync(d_stats, 0, 32, st));
  cull_compile_kernel<<<(C + 127) / 128, 128, 0, st>>>(d_c, C, th, cloud->pmax, cloud->nmax, d_rec, d_col);
  RSC_CUDA(ctx, cudaGetLastError());
  CullArgs a;
  a.msoa = cloud->cells.msoa;
  a.n = cloud->n, a.n_pad = cloud->n_pad;
  a.en = cloud->cells.en_sorted;
  a.tiles = reinterpret_cast<const float4*>(cloud->cells.tiles);
  a.ntiles = (int)(cloud->n_pad / kTile);
  a.rec = d_rec, a.col = d_col, a.cands = d_c, a.trig = d_trig;
  a.th = th;
  a.C = C;
  a.cv = d_cv, a.ce = d_cv + C;
  a.stats = d_stats;
  a.qn = reinterpret_cast<uint32_t*>(d_stats + 2);
  a.queue = reinterpret_cast<uint2*>(b + o_queue);
  a.qcap = (uint32_t)qcap;
  // in-band pairs: queued for cull_fix_kernel (default; 12.2 ms on c3) or decided on the spot
  // (RSC_CULL_INLINE=1; 13.5 ms) -- both validated against the dense path (tests/test_cull_gpu.py)
  a.inline_fp64 = getenv("RSC_CULL_INLINE") ? atoi(getenv("RSC_CULL_INLINE")) : 0;
  const int grid = a.ntiles < ctx->sm_count * 8 ? a.ntiles : ctx->sm_count * 8;
  RSC_CUDA(ctx, cudaEventRecord(ctx->evk0, st));
  cull_score_kernel<<<grid, kCullThreads, 0, st>>>(a);
  RSC_CUDA(ctx, cudaGetLastError());
  cull_fix_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(a);
  RSC_CUDA(ctx, cudaGetLastError());
  RSC_CUDA(ctx, cudaEventRecord(ctx->evk1, st));
  cull_policy_kernel<<<(C + 255) / 256, 256, 0, st>>>(d_c, C, a.cv, a.ce, th.honour_enabled, d_cv + 2 * (size_t)C);
  RSC_CUDA(ctx, cudaGetLastError());
  unsigned long long hs[2] = {0, 0};
  RSC_CUDA(ctx, cudaMemcpyAsync(counts, d_cv + 2 * (size_t)C, (size_t)C * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  RSC_CUDA(ctx, cudaMemcpyAsync(hs, d_stats, sizeof(hs), cudaMemcpyDeviceToHost, st));
  RSC_CUDA(ctx, cudaStreamSynchronize(st));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, ctx->evk0, ctx->evk1);
  if (kernel_ms) *kernel_ms = ms;
  if (pairs_total) *pairs_total = (int64_t)C * a.ntiles;
  if (pairs_survived) *pairs_survived = (int64_t)hs[0];
  ctx->stats.evals += (int64_t)C * cloud->n;
  ctx->stats.cands_scored += C;
  ctx->stats.exact_pairs += (int64_t)hs[1];
  return RSC_OK;
}
