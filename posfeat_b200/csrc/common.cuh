// Shared helpers for the posfeat_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/posfeat_b200.h"

namespace posfeat {

// thread-local last-error buffer, filled by set_error (api.cu)
int set_error(int code, const char* fmt, ...);

#define PF_CHECK_ARG(cond, ...)                                   \
  do {                                                            \
    if (!(cond)) return ::posfeat::set_error(POSFEAT_EINVAL, __VA_ARGS__); \
  } while (0)

#define PF_CUDA(call)                                                         \
  do {                                                                        \
    cudaError_t e__ = (call);                                                 \
    if (e__ != cudaSuccess)                                                   \
      return ::posfeat::set_error(POSFEAT_ECUDA, "%s failed: %s (%s:%d)", #call, \
                                  cudaGetErrorString(e__), __FILE__, __LINE__);  \
  } while (0)

#define PF_LAUNCH_CHECK(name)                                                     \
  do {                                                                            \
    ::posfeat::count_launch();                                                    \
    cudaError_t e__ = cudaGetLastError();                                         \
    if (e__ != cudaSuccess)                                                       \
      return ::posfeat::set_error(POSFEAT_ECUDA, "launch of %s failed: %s", name, \
                                  cudaGetErrorString(e__));                       \
  } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__device__ __forceinline__ int reflect_idx(int i, int n) {
  // F.pad(mode='reflect'): -k -> k, (n-1)+k -> (n-1)-k   (|k| <= n-1)
  i = i < 0 ? -i : i;
  return i > n - 1 ? 2 * (n - 1) - i : i;
}

// torch.linspace(-1, 1, n)[i] in float32, bit exact (see oracle linspace_f32)
__device__ __forceinline__ float linspace_pm1(int i, int n, float step) {
  return (i < n / 2) ? __fmaf_rn(step, (float)i, -1.0f) : __fmaf_rn(-step, (float)(n - 1 - i), 1.0f);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

int sm_count();

// launch accounting and optional per-kernel CUDA-event timing (api.cu)
void count_launch();
enum ProfSlot {
  PROF_NMS = 0, PROF_SELECT, PROF_SAMPLE, PROF_MNN_PREP, PROF_MNN_TC, PROF_MNN_RESCORE, PROF_MNN_SIMT,
  PROF_MNN_COMPACT, PROF_CORR_FWD, PROF_CORR_BWD, PROF_WIN_FWD, PROF_WIN_BWD, PROF_MNN_SCAN, PROF_MNN_VERIFY,
  PROF_FETCH, PROF_KPOUT, PROF_COUNT
};
void prof_begin(int slot, cudaStream_t s);
void prof_end(int slot, cudaStream_t s);
struct ProfScope {
  int slot; cudaStream_t s;
  ProfScope(int slot_, cudaStream_t s_) : slot(slot_), s(s_) { prof_begin(slot, s); }
  ~ProfScope() { prof_end(slot, s); }
};

}  // namespace posfeat
