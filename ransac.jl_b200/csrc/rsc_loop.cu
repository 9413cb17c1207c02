// rsc_loop.cu -- ransac(pc, params) for the built-in shapes with the whole bookkeeping on the device
// (iterations.jl:35-162): the default loop behind rsc_ransac_run.
//
// Per speculative batch of nb iterations (rsc_run.cu explains the batching):
//   K1  sample + fit                      rsc_fit.cu        -> candidates, per-iteration segment bounds
//   K2  score the new candidates on subset 1 (iterations.jl:95), counts all-reduced when sharded
//   K3  finish_new / argmax / seg_argmax  -> best of the old store, best of every iteration's candidates
//       decide_kernel                     the reference's bookkeeping, iteration by iteration:
//                                         countcandidates (iterations.jl:94-102), findhighestscore
//                                         (fitting.jl:140-150), estimatescore (confidenceintervals.jl:
//                                         53-74, Int64 wrap included), prob / chooseS (utilities.jl:262,
//                                         :297-300), the extraction test (iterations.jl:113) and the
//                                         termination test (iterations.jl:151-156) -- it cuts the batch
//                                         at the first extraction / termination and leaves a 128-byte
//                                         record (iterations used, extract?, which candidate, its 64-byte
//                                         shape, counters) that the host reads with ONE synchronisation
//   on extraction:
//   K4  refit over the rank's points, ascending inlier indices by stream compaction, enabled bits
//       cleared (rsc_extract.cu)
//   K5  removeinvalidshapes!: the stored candidates are re-scored with K2 against the subset-1 copy
//       under the mask "enabled before the extraction and not after" -- a candidate has a now-disabled
//       inlier iff that count is non-zero (see rsc_run.cu for the equivalence, Q4 included) -- then
//       invalidate + order-preserving compaction.  No gather, no intermediate host round trip.
// Host synchronisations: 2 per batch (the number of new candidates sizes the K2 launch; the decision
// record) + 1 per extraction (store size, inlier count).
//
// Multi-GPU, one process per GPU (include/rsc.h "point-range sharding"):
//   sharded storage  every rank holds its point range only; minimal sets are drawn from the replicated
//                    whole-cloud enabled mask (identical on all ranks), their coordinates gathered by one
//                    all-reduce per batch; counts / K5 hits all-reduced; the inlier list of a shape stays
//                    distributed (each rank holds the ascending indices of its range)
//   replicated       every rank holds the cloud and scores / refits its range (rsc_cloud_set_range)
// All ranks take identical decisions: decide_kernel runs replicated on identical (all-reduced) counts.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>

#include "rsc_loop.cuh"

namespace rsc {

constexpr int kMaxBatch = 64;

struct DecideParams {
  long long N, M, tau;
  double prob_det;
  int S, drawN, extract_s, terminate_s;
};

struct BatchRec {  // device -> host, once per batch
  long long counters[3];
  int32_t used, extract, terminated, iterations;
  int32_t best_idx, best_score, store_n, ovf;  // ovf: 1 = a guard-band queue overflowed, 2 = more new candidates than the launch was sized for
  int32_t n_new, pad;
  rsc_cand best;
};

struct LoopDev {  // device-resident scratch + state of one run
  int64_t oldbest[2];             // arg-max over the store as it was before the batch: index, score
  long long seg_keys[kMaxBatch];  // per batch iteration: arg-max key over its new candidates
  int32_t seg[kMaxBatch + 2];     // per batch iteration: first new candidate
  long long counters[3];          // countcandidates: lengthC, allcand, nofminset (iterations.jl:70)
  BatchRec rec;
  unsigned long long kept;        // store size after K5
  int32_t tot_global[2];          // sharded storage: inliers of the extraction summed over the ranks
};

struct K5Host {  // what the host reads after an extraction
  unsigned long long kept, total_local;
  int32_t ovf, tot_global;
};

// estimatescore(...).E with Julia's Int64 arithmetic (confidenceintervals.jl:53-74; the products wrap, Q9)
__device__ double estimate_E(long long M, long long N, long long count) {
  const long long Np = -2 - M, x = -2 - N, n = -1 - count;
  const long long xn = (long long)((unsigned long long)x * (unsigned long long)n);
  const long long prod =
      (long long)((unsigned long long)((long long)((unsigned long long)xn * (unsigned long long)(Np - x))) * (unsigned long long)(Np - n));
  const double sq_ = (double)prod / (double)(Np - 1);
  const double sq = sq_ < 0 ? 0.0 : sqrt(sq_);
  const double a = -1 - ((double)xn + sq) / (double)Np, b = -1 - ((double)xn - sq) / (double)Np;
  if (a != a || b != b) return NAN;
  const double lo = a < b ? a : b, hi = a < b ? b : a;
  return (lo + hi) / 2;
}

__device__ double prob_dev(double n, double s, double N, double k) { return 1 - pow(1 - pow(n / N, k), s); }  // utilities.jl:262

// K3: the bookkeeping of one batch.  The walk over the batch's iterations is sequential by nature (cut at the
// first extraction / termination), but what it compares is not: thread j evaluates iteration j's counters,
// running best (a prefix maximum of the per-iteration arg-max keys), estimatescore, and the two prob() tests
// (four FP64 pow calls); thread 0 then takes the first iteration that extracts or terminates.
__global__ void __launch_bounds__(kMaxBatch) decide_kernel(LoopDev* __restrict__ d, const int32_t* __restrict__ ovf, int store_n0,
                                                           int nb, int k0, DecideParams q, const rsc_cand* __restrict__ cands,
                                                           int cap_new) {
  __shared__ long long s_key[kMaxBatch];
  __shared__ int s_ext[kMaxBatch], s_term[kMaxBatch];
  const int j = threadIdx.x;
  const int over = ovf ? (*ovf ? 1 : 0) : 0;
  const int n_new = d->seg[nb];
  const bool redo = over || (cap_new >= 0 && n_new > cap_new);
  if (j < nb && !redo) {
    long long key = (store_n0 >= 1 && d->oldbest[0] >= 0) ? ((d->oldbest[1] << 32) | (long long)(0x7fffffff - d->oldbest[0])) : -1;
    for (int i = 0; i <= j; ++i)
      if (d->seg_keys[i] > key) key = d->seg_keys[i];  // strict >: the first maximum wins (the key carries the store index)
    const long long c0 = store_n0 + d->seg[j + 1];                 // iterations.jl:102
    const long long c1 = d->counters[1] + d->seg[j + 1];           // iterations.jl:94 (accumulated over the batch)
    const long long c2 = (long long)(k0 + j) * q.S;                // iterations.jl:99
    const long long cc[3] = {c0, c1, c2};
    int ext = 0;
    if (c0 >= 1 && key >= 0) {
      const double E = estimate_E(q.M, q.N, (long long)(key >> 32));
      ext = prob_dev(E, (double)cc[q.extract_s], (double)q.N, (double)q.drawN) > q.prob_det;  // iterations.jl:113
    }
    s_key[j] = key;
    s_ext[j] = ext;
    s_term[j] = prob_dev((double)q.tau, (double)cc[q.terminate_s], (double)q.N, (double)q.drawN) > q.prob_det;  // iterations.jl:151-156
  }
  __syncthreads();
  if (j != 0) return;
  BatchRec r;
  r.ovf = over ? 1 : ((cap_new >= 0 && n_new > cap_new) ? 2 : 0);  // 2: the sync-free launch was sized for fewer candidates
  r.n_new = n_new;
  r.pad = 0;
  r.used = 0, r.extract = 0, r.terminated = 0, r.iterations = k0 - 1, r.best_idx = -1, r.best_score = 0, r.store_n = store_n0;
  for (int i = 0; i < 3; ++i) r.counters[i] = d->counters[i];
  if (redo) {  // the counts are incomplete (queue overflow on some rank) or some candidates were not scored: the host repeats K2
    d->rec = r;
    return;
  }
  int last = nb - 1;
  for (int i = 0; i < nb; ++i)
    if (s_ext[i] || s_term[i]) {
      last = i;
      break;
    }
  if (nb > 0) {
    r.used = last + 1;
    r.iterations = k0 + last;
    r.store_n = store_n0 + d->seg[last + 1];
    r.counters[0] = r.store_n;
    r.counters[1] = d->counters[1] + d->seg[last + 1];
    r.counters[2] = (long long)(k0 + last) * q.S;
    r.extract = s_ext[last];
    r.terminated = s_term[last];
    if (r.extract) {
      r.best_idx = 0x7fffffff - (int)(s_key[last] & 0xffffffffll);
      r.best_score = (int)(s_key[last] >> 32);
      r.best = cands[r.best_idx];
    }
  }
  for (int i = 0; i < 3; ++i) d->counters[i] = r.counters[i];
  d->rec = r;
}

// the sync-free batch path: the number of new candidates stays on the device (the fit's compaction total)
__global__ void append_new_kernel(const rsc_cand* __restrict__ src, const unsigned long long* __restrict__ total, int cap,
                                  rsc_cand* __restrict__ dst) {
  const int n = (int)min(*total, (unsigned long long)cap);
  const uint4* s = reinterpret_cast<const uint4*>(src);
  uint4* d = reinterpret_cast<uint4*>(dst);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n * 4; i += gridDim.x * blockDim.x) d[i] = s[i];
}

__global__ void finish_new_dev_kernel(const rsc_cand* __restrict__ cands, const unsigned long long* __restrict__ total, int cap,
                                      const int32_t* __restrict__ cv, const int32_t* __restrict__ ce, uint32_t honour_enabled,
                                      int32_t* __restrict__ score, uint8_t* __restrict__ flags) {
  const int n = (int)min(*total, (unsigned long long)cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int t = cands[i].type;
    const bool honour = (honour_enabled >> t) & 1u;
    score[i] = honour ? ce[i] : cv[i];
    flags[i] = (uint8_t)(1u | ((!honour && cv[i] != ce[i]) ? 2u : 0u));
  }
}

// K5: mask of the subset points the extraction just disabled, in place: old &= ~new
__global__ void newly_mask_kernel(uint32_t* __restrict__ old_en, const uint32_t* __restrict__ new_en, int64_t words) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w < words) old_en[w] &= ~new_en[w];
}

// K5, gathered form: the newly disabled subset points (old & ~new) into a compact SoA scratch set of
// `cap` points (the host sizes it by the extracted candidate's subset score, an upper bound of their
// number); *over = 1 if they do not fit (then the host repeats K5 in the masked form)
__global__ void newly_gather_capped_kernel(const uint32_t* __restrict__ old_en, const uint32_t* __restrict__ new_en, int64_t words,
                                           const unsigned long long* __restrict__ offs, const unsigned long long* __restrict__ total,
                                           const float* __restrict__ soa, int64_t m_pad, float* __restrict__ out, int64_t cap,
                                           int32_t* __restrict__ over) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w == 0) *over = (*total > (unsigned long long)cap) ? 1 : 0;
  if (w >= words) return;
  uint32_t bits = old_en[w] & ~new_en[w];
  unsigned long long o = offs[w];
  while (bits) {
    const int b = __ffs(bits) - 1;
    bits &= bits - 1;
    const int64_t j = w * 32 + b;
    if (o < (unsigned long long)cap) {
#pragma unroll
      for (int f = 0; f < 6; ++f) out[f * cap + o] = soa[f * m_pad + j];
    }
    ++o;
  }
}

// valid (= enabled) words of the scratch set: the first *n points are real
__global__ void fill_valid_dev_kernel(uint32_t* __restrict__ valid, const unsigned long long* __restrict__ n, int64_t words) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= words) return;
  const long long nn = (long long)*n, lo = w * 32;
  valid[w] = lo + 32 <= nn ? 0xffffffffu : (lo < nn ? (1u << (nn - lo)) - 1u : 0u);
}

__global__ void k5_pack_kernel(LoopDev* __restrict__ d, const unsigned long long* __restrict__ ktot, const int32_t* __restrict__ ovf,
                               const int32_t* __restrict__ over, const unsigned long long* __restrict__ total_local,
                               K5Host* __restrict__ out) {
  out->kept = *ktot;
  out->total_local = *total_local;
  out->ovf = (*ovf ? 1 : 0) + (*over ? 2 : 0);  // bit 0: guard-band queue overflow, bit 1: scratch set too small (any rank)
  out->tot_global = d->tot_global[0];
}

int32_t ransac_loop_device(rsc_cloud* cloud, const rsc_params* p, uint64_t seed, rsc_run* run) {
  rsc_ctx* ctx = cloud->ctx;
  cudaStream_t st = ctx->stream;
  if (!ctx->loop_scratch) ctx->loop_scratch = new LoopScratch();
  LoopScratch& ls = *static_cast<LoopScratch*>(ctx->loop_scratch);
  Store& store = ls.store;
  store.n = 0, store.cur = 0;
  int32_t rc = RSC_OK;
  const Thresh th = make_thresh(p);
  rsc_subset& sub = cloud->subsets[0];
  const bool shard = cloud->is_shard();
  const bool ranged = !shard && ctx->allreduce != nullptr && cloud->range_set;
  const bool coll = shard || ranged;  // counts are summed over the ranks
  if (shard && !ctx->allreduce)
    return fail(ctx, RSC_E_STATE, "ransac_run: sharded storage needs a communicator (rsc_ctx_comm_init or rsc_ctx_set_allreduce)");
  if (shard && cloud->n_global >= 2147483647LL) return fail(ctx, RSC_E_ARG, "ransac_run: sharded storage supports clouds below 2^31 points");
  if (shard && (p->compat_flags & (RSC_REFIT_LSQ | RSC_EXTRACT_BITMAP)))
    return fail(ctx, RSC_E_STATE, "ransac_run: RSC_REFIT_LSQ / RSC_EXTRACT_BITMAP are not available on sharded storage");
  if ((p->compat_flags & RSC_EXTRACT_BITMAP) && !(ctx->bitmap_beta > 0.0))
    return fail(ctx, RSC_E_STATE, "ransac_run: RSC_EXTRACT_BITMAP needs rsc_ctx_set_bitmap(ctx, beta, eight) first");
  const int64_t N = cloud->n_global > 0 ? cloud->n_global : cloud->n;
  const int64_t M = sub.m_global > 0 ? sub.m_global : sub.m;
  const int S = p->minsubsetN;
  const int maxnew = S * p->n_shape_types;
  // this rank's slice of the subset copy: all of it (single GPU, or a shard's local entries) or, on
  // replicated storage, the same fraction of it as the rank's range of the cloud
  PointSet sps = view_subset(&sub);
  if (ranged && cloud->n_pad > 0) {
    const int64_t lo = (int64_t)((double)cloud->range_lo / cloud->n_pad * sub.m_pad) / kTile * kTile;
    const int64_t hi =
        cloud->range_hi >= cloud->n_pad ? sub.m_pad : (int64_t)((double)cloud->range_hi / cloud->n_pad * sub.m_pad) / kTile * kTile;
    sps.x += lo, sps.y += lo, sps.z += lo, sps.nx += lo, sps.ny += lo, sps.nz += lo;
    sps.enabled += lo / 32, sps.valid += lo / 32;
    sps.n_pad = std::max<int64_t>(0, hi - lo);
    sps.n = std::max<int64_t>(0, (sub.m < hi ? sub.m : hi) - lo);
  }
  const int64_t sps_word0 = sps.enabled - sub.enabled;
  // small-batch scorer (rsc_small.cu) limits: beyond them the tiled kernel is the faster one
  const bool use_small = getenv("RSC_NO_SMALL") == nullptr;
  // (measured on c4, profiles/r2h_launches_ransac_c4_summary.json: small kernel ~60 us at ~100 candidates x 312 k
  // points and linear in the candidates, tiled path ~190 us flat at this size -> cross-over near 1e8 evaluations)
  constexpr int kSmallMaxCands = 4096;
  constexpr double kSmallMaxEvals = 1.2e8;
  long long rate_seen = -1;  // recent number of new candidates per iteration (-1: no batch yet)
  const int Bmax = getenv("RSC_BATCH") ? std::max(1, std::min(kMaxBatch, atoi(getenv("RSC_BATCH")))) : 16;
  const int Bmin = getenv("RSC_BATCH_MIN") ? std::max(1, std::min(Bmax, atoi(getenv("RSC_BATCH_MIN")))) : std::min(8, Bmax);
  // Without the two switches the batch sizes follow the run: a batch is cut at the first extraction and the sets drawn
  // behind the cut are fitted for nothing, so batches should not be much longer than the usual gap between two
  // extractions (c4: an extraction every ~11 iterations -> 6 rising to 12 measured best (13.7 ms of K1 + K2 against 15.0
  // at 8..16); c5: every ~17 -> 8..16).  `gap` = smoothed iterations per extraction; results do not depend on any of this.
  const bool adapt = !getenv("RSC_BATCH") && !getenv("RSC_BATCH_MIN");
  double gap = 16.0;
  int since_extract = 0;
  int B = 1;
  const bool trace = getenv("RSC_TRACE") != nullptr;
  double t_fit = 0, t_score = 0, t_extract = 0;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<double>(b - a).count();
  };
  auto sync = [&]() {
    ++run->syncs;
    return cudaStreamSynchronize(st);
  };
#define RUN_CUDA(expr)                \
  do {                                \
    cudaError_t _e = (expr);          \
    if (_e != cudaSuccess) return fail_cuda(ctx, _e, #expr); \
  } while (0)

  // pinned mirror of what the host reads (seg bounds, decision record, K5 summary)
  struct HostIo {
    int32_t seg[kMaxBatch + 2];
    BatchRec rec;
    K5Host k5;
  };
  if (ctx->pinned_cap < sizeof(HostIo)) {
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    ctx->pinned = nullptr, ctx->pinned_cap = 0;
    RUN_CUDA(cudaMallocHost(&ctx->pinned, sizeof(HostIo) + 256));
    ctx->pinned_cap = sizeof(HostIo) + 256;
  }
  HostIo* hio = static_cast<HostIo*>(ctx->pinned);
  RUN_CUDA(ls.hostio.ensure(sizeof(LoopDev) + sizeof(K5Host) + 64));
  LoopDev* dev = ls.hostio.as<LoopDev>();
  K5Host* dk5 = reinterpret_cast<K5Host*>(dev + 1);
  RUN_CUDA(cudaMemsetAsync(dev, 0, sizeof(LoopDev) + sizeof(K5Host), st));

  // sharded storage: whatever the ranks did to their local masks, the replicated one follows them
  if (shard && (rc = shard_sync_enabled(cloud, st))) return rc;
  int64_t n_enabled = 0, n_enabled_local = rsc_cloud_count_enabled(cloud);
  if (n_enabled_local < 0) return fail(ctx, RSC_E_CUDA, "ransac_run: counting the enabled points failed");
  ++run->syncs;
  if (shard) {
    n_enabled = count_mask_bits(ctx, cloud->g_enabled, cloud->g_words);
    if (n_enabled < 0) return fail(ctx, RSC_E_CUDA, "ransac_run: counting the enabled points failed");
    ++run->syncs;
  } else {
    n_enabled = n_enabled_local;
  }
  run->device = ctx->device;
  RUN_CUDA(cudaMalloc(&run->d_idx, (size_t)(n_enabled_local > 0 ? n_enabled_local : 1) * sizeof(int64_t)));  // a point is extracted at most once

  RUN_CUDA(ls.newcnt.ensure((size_t)2 * maxnew * Bmax * 4 + 64));
  if ((rc = fit_reserve(ctx, p, S * Bmax))) return rc;
  RUN_CUDA(store.reserve((size_t)std::min<int64_t>((int64_t)maxnew * Bmax, 1 << 20), st));
  DecideParams q;
  q.N = N, q.M = M, q.tau = p->tau, q.prob_det = p->prob_det;
  q.S = S, q.drawN = p->drawN, q.extract_s = p->extract_s, q.terminate_s = p->terminate_s;
  bool terminated = false;

  for (int k = 1; k <= p->itermax && !terminated;) {
    if (n_enabled < p->tau) break;  // iterations.jl:75
    const int nb = std::min(B, p->itermax - k + 1);
    const auto tk0 = now();
    ++run->batches;
    // ---- K1: nb * minsubsetN minimal sets -> candidates (device, compacted in reference order) ----
    FitScratch fs;
    // (segment bounds of the batch's iterations: from the compaction itself when an iteration is whole groups of sets)
    const bool seg_fused = S % 128 == 0;
    if ((rc = fit_enqueue(ctx, cloud, 2, p, p->drawN, nullptr, nullptr, nullptr, S * nb, seed, (uint64_t)(k - 1) * S, st, &fs, nullptr,
                          seg_fused ? dev->seg : nullptr, S, nb)))
      return rc;
    if (!seg_fused) {
      seg_bounds_kernel<<<1, 96, 0, st>>>(fs.out_set, fs.total, S, nb, dev->seg);
      RUN_CUDA(cudaGetLastError());
    }
    const int store_n0 = store.n;
    // ---- sync-free path: K2 for a small batch, sized by a PREDICTION of the number of new candidates ----
    // (4 x the largest per-iteration yield seen so far; the true number stays on the device.  If more turn up,
    // decide_kernel says so and the batch is scored again the classic way.)
    int cap_new = -1;
    if (use_small && rate_seen >= 0) {
      long long want = std::max<long long>(128, 4ll * rate_seen * nb + 64);
      if (const char* e = getenv("RSC_SMALL_CAP")) want = std::max(1, atoi(e));  // test hook: force undersized launches
      const long long cap = std::min<long long>((want + 127) / 128 * 128, (long long)maxnew * nb);
      // the launch is sized for `cap` but its cost follows the ACTUAL number (expected: cap / 4)
      if (cap <= kSmallMaxCands && (double)(cap / 4) * (double)std::max<int64_t>(sps.n_pad, 1) <= kSmallMaxEvals) cap_new = (int)cap;
    }
    bool scored = false;
    if (cap_new > 0) {
      RUN_CUDA(store.reserve((size_t)store_n0 + cap_new, st));
      RUN_CUDA(ls.newcnt.ensure((size_t)2 * cap_new * 4 + 64));
      rsc_cand* dst = store.cands[store.cur].as<rsc_cand>() + store_n0;
      int32_t* cv = ls.newcnt.as<int32_t>();
      int32_t* ce = cv + cap_new;
      append_new_kernel<<<std::max(1, std::min(cap_new * 4 / 256, 64)), 256, 0, st>>>(fs.out, fs.total, cap_new, dst);
      RUN_CUDA(cudaGetLastError());
      if ((rc = score_small_enqueue(ctx, cloud, sps, th, dst, cap_new, fs.total, cv, ce, st))) return rc;
      if (coll && ctx->allreduce(ctx->allreduce_user, cv, (int64_t)2 * cap_new, (void*)st))
        return fail(ctx, RSC_E_NCCL, "ransac_run: all-reduce of the counts failed");
      finish_new_dev_kernel<<<std::max(1, std::min(cap_new / 256, 64)), 256, 0, st>>>(
          dst, fs.total, cap_new, cv, ce, th.honour_enabled, store.score[store.cur].as<int32_t>() + store_n0,
          store.flags[store.cur].as<uint8_t>() + store_n0);
      RUN_CUDA(cudaGetLastError());
      seg_argmax_kernel<<<nb, 256, 0, st>>>(store.score[store.cur].as<int32_t>(), store.flags[store.cur].as<uint8_t>(), dev->seg, store_n0,
                                             dev->seg_keys, cap_new);  // never reads past the candidates this launch was sized for
      RUN_CUDA(cudaGetLastError());
      if (store_n0 >= 1) {
        RUN_CUDA(argmax_enqueue(store.score[store.cur].as<int32_t>(), store.flags[store.cur].as<uint8_t>(), store_n0, dev->oldbest, ctx->sm_count, st));
      }
      decide_kernel<<<1, kMaxBatch, 0, st>>>(dev, nullptr, store_n0, nb, k, q, store.cands[store.cur].as<rsc_cand>(), cap_new);
      RUN_CUDA(cudaGetLastError());
      RUN_CUDA(cudaMemcpyAsync(&hio->rec, &dev->rec, sizeof(BatchRec), cudaMemcpyDeviceToHost, st));
      RUN_CUDA(sync());
      scored = hio->rec.ovf == 0;
      if (scored) ctx->stats.evals += (int64_t)hio->rec.n_new * sps.n, ctx->stats.cands_scored += hio->rec.n_new;
    }
    int n_new = 0;
    if (scored) {
      n_new = hio->rec.n_new;
    } else if (cap_new > 0) {
      n_new = hio->rec.n_new;  // known from the record of the undersized attempt
    } else {
      RUN_CUDA(cudaMemcpyAsync(hio->seg, dev->seg, (size_t)(nb + 1) * 4, cudaMemcpyDeviceToHost, st));
      RUN_CUDA(sync());
      n_new = hio->seg[nb];
    }
    rate_seen = std::max<long long>((n_new + nb - 1) / nb, rate_seen / 2);  // the yield falls as shapes are extracted: follow it
    const auto tk1 = now();
    t_fit += secs(tk0, tk1);
    // ---- classic path: K2 (tiled) on subset 1 + K3 (repeated with a larger guard-band queue if that overflowed) ----
    for (int attempt = 0; !scored; ++attempt) {
      const int32_t* d_ovf = nullptr;
      if (n_new > 0) {
        RUN_CUDA(store.reserve((size_t)store_n0 + n_new, st));
        rsc_cand* dst = store.cands[store.cur].as<rsc_cand>() + store_n0;
        RUN_CUDA(cudaMemcpyAsync(dst, fs.out, (size_t)n_new * sizeof(rsc_cand), cudaMemcpyDeviceToDevice, st));
        int32_t* cv = ls.newcnt.as<int32_t>();
        int32_t* ce = cv + n_new;
        // (the culled scorer on the subset's Morton view when the batch is large: same counts)
        // the overflow flag rides behind the counts, so that a sharded run decides to repeat collectively
        if ((rc = loop_score_new(ctx, cloud, sub, sps, !ranged, th, dst, n_new, cv, ce, cv + 2 * n_new, st))) return rc;
        if (coll && ctx->allreduce(ctx->allreduce_user, cv, (int64_t)2 * n_new + 1, (void*)st))
          return fail(ctx, RSC_E_NCCL, "ransac_run: all-reduce of the counts failed");
        finish_new_kernel<<<(n_new + 255) / 256, 256, 0, st>>>(dst, n_new, cv, ce, th.honour_enabled,
                                                               store.score[store.cur].as<int32_t>() + store_n0,
                                                               store.flags[store.cur].as<uint8_t>() + store_n0);
        RUN_CUDA(cudaGetLastError());
        seg_argmax_kernel<<<nb, 256, 0, st>>>(store.score[store.cur].as<int32_t>(), store.flags[store.cur].as<uint8_t>(), dev->seg,
                                               store_n0, dev->seg_keys);
        RUN_CUDA(cudaGetLastError());
        d_ovf = cv + 2 * n_new;
      } else {
        RUN_CUDA(cudaMemsetAsync(dev->seg_keys, 0xff, sizeof(dev->seg_keys), st));  // -1: no new candidate
      }
      if (store_n0 >= 1) {
        RUN_CUDA(argmax_enqueue(store.score[store.cur].as<int32_t>(), store.flags[store.cur].as<uint8_t>(), store_n0,
                                          dev->oldbest, ctx->sm_count, st));
      }
      decide_kernel<<<1, kMaxBatch, 0, st>>>(dev, d_ovf, store_n0, nb, k, q, store.cands[store.cur].as<rsc_cand>(), -1);
      RUN_CUDA(cudaGetLastError());
      RUN_CUDA(cudaMemcpyAsync(&hio->rec, &dev->rec, sizeof(BatchRec), cudaMemcpyDeviceToHost, st));
      RUN_CUDA(sync());
      if (!hio->rec.ovf) break;
      if (attempt >= 4) return fail(ctx, RSC_E_STATE, "ransac_run: guard-band queue kept overflowing");
      if ((rc = grow_guard_queue(ctx))) return rc;
    }
    const BatchRec rec = hio->rec;
    const auto tk2 = now();
    t_score += secs(tk1, tk2);
    run->iterations = rec.iterations;
    store.n = rec.store_n;
    terminated = rec.terminated != 0;
    if (rec.extract) {
      // ---- K4: refit over this rank's points, invalidate them ----
      rsc_cand shape = rec.best;
      if (p->compat_flags & RSC_REFIT_LSQ) {  // extension: the paper's least-squares refit within 3 eps
        if ((rc = lsq_refine(cloud, p, 3.0, &shape, nullptr, nullptr, st))) return rc;
        ++run->syncs;
      }
      Thresh thr = th;
      thr.honour_enabled = 0xFu;
      const int64_t swords = sub.m_pad / 32;
      RUN_CUDA(ls.olden.ensure((size_t)swords * 4));
      RUN_CUDA(cudaMemcpyAsync(ls.olden.p, sub.enabled, (size_t)swords * 4, cudaMemcpyDeviceToDevice, st));
      if ((rc = refit_mask_enqueue(cloud, thr, shape, st))) return rc;
      if (p->compat_flags & RSC_EXTRACT_BITMAP) {  // extension: only the largest connected component in parameter space
        unsigned long long tot = 0;
        RUN_CUDA(cudaMemcpyAsync(&tot, ctx->misc2.p, 8, cudaMemcpyDeviceToHost, st));
        RUN_CUDA(sync());
        if (tot > 0) {
          RUN_CUDA(ctx->misc.ensure((size_t)tot * 16));
          int64_t* lst = ctx->misc.as<int64_t>();
          if ((rc = refit_write_enqueue(cloud, lst, false, st))) return rc;  // the list, nothing disabled yet
          int64_t kept = 0;
          if ((rc = bitmap_filter_dev(cloud, shape, ctx->bitmap_beta, ctx->bitmap_eight != 0, lst, (int64_t)tot, lst + tot, &kept, nullptr, st)))
            return rc;
          ++run->syncs;
          if ((rc = refit_mask_from_list(cloud, lst + tot, kept, st))) return rc;
        }
      }
      // the inlier mask words (ctx->idxbuf) and this rank's inlier count (ctx->misc2) are on the device
      if ((rc = refit_write_enqueue(cloud, run->d_idx + run->off.back(), true, st))) return rc;
      if (shard) {  // the replicated whole-cloud mask follows; the ranks' inlier counts ride along
        RUN_CUDA(cudaMemcpyAsync(dev->tot_global, ctx->misc2.p, 8, cudaMemcpyDeviceToDevice, st));
        if ((rc = shard_clear_enabled(cloud, ctx->idxbuf.as<uint32_t>(), dev->tot_global, 1, dev->tot_global, st))) return rc;
      }
      // ---- K5: drop the best and every candidate compatible with a newly disabled subset point ----
      // Gathered form: the newly disabled points of this rank's subset slice go into a compact scratch
      // set.  Their number is only known on the device, but it cannot exceed the candidate's subset
      // score (every one of them was a compatible, enabled subset point when the candidate was scored;
      // the enabled set only shrinks), which the host has from the decision record: that sizes the
      // launch without a round trip.  Masked form (the least-squares refit moves the shape, so the bound
      // does not hold; also the fallback): score against the whole slice under the mask old & ~new.
      const int nst = store.n;
      // (the bitmap filter only removes points from the list, so the bound still holds with it)
      bool masked = (p->compat_flags & RSC_REFIT_LSQ) != 0 || getenv("RSC_K5_MASKED") != nullptr;
      PointSet dps = sps;
      // hit counts [nst] + two flags that ride with them through the all-reduce: guard-band queue overflow,
      // scratch set too small
      RUN_CUDA(ctx->counts.ensure(((size_t)3 * nst + 8) * 4));
      int32_t* hit = ctx->counts.as<int32_t>() + 2 * (size_t)nst;
      int32_t* d_over = hit + nst + 1;
      RUN_CUDA(cudaMemsetAsync(d_over, 0, 4, st));
      const int64_t sl_words = sps.n_pad / 32;
      if (!masked && sl_words > 0) {
        const int64_t cap = ((int64_t)std::max(rec.best_score, 1) + kTile - 1) / kTile * kTile;
        const size_t o_woff = ((size_t)sl_words * 4 + 255) / 256 * 256;
        RUN_CUDA(ls.prog.ensure(o_woff + (size_t)(sl_words + 2) * 8 + 64));
        uint32_t* wcnt = ls.prog.as<uint32_t>();
        unsigned long long* woff = (unsigned long long*)(ls.prog.as<char>() + o_woff);
        unsigned long long* wtot = woff + sl_words;
        const uint32_t* old_sl = ls.olden.as<uint32_t>() + sps_word0;
        newly_count_kernel<<<(unsigned)((sl_words + 255) / 256), 256, 0, st>>>(old_sl, sps.enabled, sl_words, wcnt);
        RUN_CUDA(cudaGetLastError());
        if ((rc = scan_u32(ctx, wcnt, (int)sl_words, woff, wtot, st))) return rc;
        RUN_CUDA(ls.nscratch.ensure((size_t)6 * cap * 4));
        RUN_CUDA(ls.nvalid.ensure((size_t)(cap / 32) * 4));
        RUN_CUDA(cudaMemsetAsync(ls.nscratch.p, 0, (size_t)6 * cap * 4, st));
        newly_gather_capped_kernel<<<(unsigned)((sl_words + 255) / 256), 256, 0, st>>>(old_sl, sps.enabled, sl_words, woff, wtot, sps.x,
                                                                                      sps.y - sps.x, ls.nscratch.as<float>(), cap, d_over);
        RUN_CUDA(cudaGetLastError());
        fill_valid_dev_kernel<<<(unsigned)((cap / 32 + 255) / 256), 256, 0, st>>>(ls.nvalid.as<uint32_t>(), wtot, cap / 32);
        RUN_CUDA(cudaGetLastError());
        float* b = ls.nscratch.as<float>();
        dps.x = b, dps.y = b + cap, dps.z = b + 2 * cap, dps.nx = b + 3 * cap, dps.ny = b + 4 * cap, dps.nz = b + 5 * cap;
        dps.enabled = dps.valid = ls.nvalid.as<uint32_t>();
        dps.n = dps.n_pad = cap;
      }
      auto mask_form = [&]() -> cudaError_t {
        newly_mask_kernel<<<(unsigned)((swords + 255) / 256), 256, 0, st>>>(ls.olden.as<uint32_t>(), sub.enabled, swords);
        dps = sps;
        dps.enabled = ls.olden.as<uint32_t>() + sps_word0;
        masked = true;
        cudaMemsetAsync(d_over, 0, 4, st);
        return cudaGetLastError();
      };
      if (masked) RUN_CUDA(mask_form());
      const size_t o_koff = ((size_t)nst * 4 + 255) / 256 * 256;
      RUN_CUDA(ls.nmeta.ensure(o_koff + (size_t)(nst + 1) * 8));
      uint32_t* keep = ls.nmeta.as<uint32_t>();
      unsigned long long* koff = (unsigned long long*)(ls.nmeta.as<char>() + o_koff);
      unsigned long long* ktot = koff + nst;
      int nxt = store.cur ^ 1;
      for (int attempt = 0;; ++attempt) {
        // enabled-gated counts over the newly disabled points = hits
        if (use_small && nst <= kSmallMaxCands && (double)nst * (double)std::max<int64_t>(dps.n_pad, 1) <= kSmallMaxEvals) {
          if ((rc = score_small_enqueue(ctx, cloud, dps, th, store.cands[store.cur].as<rsc_cand>(), nst, nullptr,
                                        ctx->counts.as<int32_t>(), hit, st)))
            return rc;
          RUN_CUDA(cudaMemsetAsync(hit + nst, 0, 4, st));  // no guard-band queue on this path
        } else {
          if ((rc = score_enqueue(ctx, cloud, dps, th, store.cands[store.cur].as<rsc_cand>(), nst, nullptr, false, st,
                                  ctx->counts.as<int32_t>(), hit)))
            return rc;
          queue_overflow_kernel<<<1, 1, 0, st>>>(ctx->wl_count.as<uint32_t>(), (uint32_t)ctx->wl_cap, hit + nst);
          RUN_CUDA(cudaGetLastError());
        }
        if (coll && ctx->allreduce(ctx->allreduce_user, hit, (int64_t)nst + 2, (void*)st))
          return fail(ctx, RSC_E_NCCL, "ransac_run: all-reduce of the K5 hits failed");
        invalidate_kernel<<<(nst + 255) / 256, 256, 0, st>>>(hit, store.flags[store.cur].as<uint8_t>(), nst, rec.best_idx, keep);
        RUN_CUDA(cudaGetLastError());
        if ((rc = scan_u32(ctx, keep, nst, koff, ktot, st))) return rc;
        compact_store_kernel<<<(nst + 255) / 256, 256, 0, st>>>(
            store.cands[store.cur].as<rsc_cand>(), store.score[store.cur].as<int32_t>(), store.flags[store.cur].as<uint8_t>(), keep,
            koff, nst, store.cands[nxt].as<rsc_cand>(), store.score[nxt].as<int32_t>(), store.flags[nxt].as<uint8_t>());
        RUN_CUDA(cudaGetLastError());
        k5_pack_kernel<<<1, 1, 0, st>>>(dev, ktot, hit + nst, d_over, ctx->misc2.as<unsigned long long>(), dk5);
        RUN_CUDA(cudaGetLastError());
        RUN_CUDA(cudaMemcpyAsync(&hio->k5, dk5, sizeof(K5Host), cudaMemcpyDeviceToHost, st));
        RUN_CUDA(sync());
        if (!hio->k5.ovf) break;
        if (hio->k5.ovf & 2) {  // (not expected) more newly disabled points than the bound: masked form
          RUN_CUDA(mask_form());
          continue;
        }
        // invalidate_kernel marked flags of this attempt: they only ever go from alive to dead on hits that
        // can only grow with the complete counts, so repeating on the same flags is safe
        if (attempt >= 4) return fail(ctx, RSC_E_STATE, "ransac_run: guard-band queue kept overflowing");
        if ((rc = grow_guard_queue(ctx))) return rc;
      }
      const int64_t total_local = (int64_t)hio->k5.total_local;
      const int64_t total = shard ? (int64_t)hio->k5.tot_global : total_local;
      run->shapes.push_back(shape);
      run->off.push_back(run->off.back() + total_local);
      run->total.push_back(total);
      n_enabled -= total;
      store.cur = nxt;
      store.n = (int)hio->k5.kept;
      t_extract += secs(tk2, now());
    }
    k += rec.used;
    since_extract += rec.used;
    if (rec.extract) {
      gap = 0.5 * gap + 0.5 * (double)since_extract;
      since_extract = 0;
    }
    if (adapt) {
      const int bmax = std::max(8, std::min(Bmax, (int)(gap + 1.5)));
      B = rec.extract ? std::max(4, std::min(bmax, (int)(0.5 * gap + 0.5))) : std::min(2 * B, bmax);
    } else {
      B = rec.extract ? Bmin : std::min(2 * B, Bmax);
    }
  }
  RUN_CUDA(sync());
  if (trace)
    fprintf(stderr,
            "[rsc_ransac_run/device] iterations %d shapes %zu batches %d syncs %d | sample+fit %.1f ms, score+decide %.1f ms, "
            "refit+invalidate %.1f ms | all-reduces %lld (%.1f MB)\n",
            run->iterations, run->shapes.size(), run->batches, run->syncs, 1e3 * t_fit, 1e3 * t_score, 1e3 * t_extract,
            (long long)ctx->allreduce_calls, ctx->allreduce_bytes / 1e6);
  return RSC_OK;
#undef RUN_CUDA
}

}  // namespace rsc
