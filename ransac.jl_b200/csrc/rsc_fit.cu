// rsc_fit.cu -- K1 (placeholder until the batched fit kernel lands in this file)
#include "rsc_common.cuh"
using namespace rsc;
extern "C" {
int32_t rsc_fit_batch(rsc_cloud* cloud, const rsc_params*, const int64_t*, int32_t, rsc_cand*, int32_t*, int32_t*) {
  return cloud ? fail(cloud->ctx, RSC_E_STATE, "rsc_fit_batch: not built yet") : RSC_E_ARG;
}
int32_t rsc_fit_points(rsc_ctx* ctx, const rsc_params*, const double*, const double*, int32_t, int32_t, rsc_cand*,
                       int32_t*, int32_t*) {
  return fail(ctx, RSC_E_STATE, "rsc_fit_points: not built yet");
}
int32_t rsc_sample_fit(rsc_cloud* cloud, const rsc_params*, uint64_t, uint64_t, int32_t, rsc_cand*, int32_t*, int64_t*,
                       int32_t*) {
  return cloud ? fail(cloud->ctx, RSC_E_STATE, "rsc_sample_fit: not built yet") : RSC_E_ARG;
}
}
