// rsc_run.cu -- the whole ransac loop (placeholder until the loop lands in this file)
#include "rsc_common.cuh"
using namespace rsc;
struct rsc_run {
  int dummy;
};
extern "C" {
int32_t rsc_ransac_run(rsc_cloud* cloud, const rsc_params*, uint64_t, rsc_run**) {
  return cloud ? fail(cloud->ctx, RSC_E_STATE, "rsc_ransac_run: not built yet") : RSC_E_ARG;
}
int32_t rsc_run_nshapes(const rsc_run*) { return 0; }
int32_t rsc_run_iterations(const rsc_run*) { return 0; }
double rsc_run_seconds(const rsc_run*) { return 0.0; }
int32_t rsc_run_shape(const rsc_run*, int32_t, rsc_cand*, int64_t*) { return RSC_E_STATE; }
int32_t rsc_run_inpoints(const rsc_run*, int32_t, int64_t*) { return RSC_E_STATE; }
void rsc_run_destroy(rsc_run* r) { delete r; }
}
