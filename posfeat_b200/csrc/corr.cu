// Subsystem (4): training-side correlation + softmax expectation (placeholder:
// entry points exist so the ABI is complete; kernels land next).
#include "common.cuh"
using namespace posfeat;

extern "C" int posfeat_corr_expect_fwd_f32(const float*, const float*, const float*, int, int, int, int, int, int,
                                           float, float*, float*, void*) {
  return set_error(POSFEAT_EUNSUPPORTED, "corr_expect_fwd not built yet");
}
extern "C" int posfeat_corr_expect_bwd_f32(const float*, const float*, const float*, int, int, int, int, int, int,
                                           float, const float*, const float*, const float*, float*, float*, void*) {
  return set_error(POSFEAT_EUNSUPPORTED, "corr_expect_bwd not built yet");
}
extern "C" int posfeat_window_expect_fwd_f32(const float*, int, int, int, int, int64_t, int64_t, int64_t, int64_t,
                                             const float*, const float*, int, const float*, int, float*, float*,
                                             float*, float*, void*) {
  return set_error(POSFEAT_EUNSUPPORTED, "window_expect_fwd not built yet");
}
extern "C" int posfeat_window_expect_bwd_f32(const float*, int, int, int, int, int64_t, int64_t, int64_t, int64_t,
                                             const float*, const float*, int, const float*, int, const float*,
                                             const float*, const float*, const float*, float*, float*, void*) {
  return set_error(POSFEAT_EUNSUPPORTED, "window_expect_bwd not built yet");
}
