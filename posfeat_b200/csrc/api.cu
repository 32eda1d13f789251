// C-ABI plumbing: version, thread-local error string, device query, and the
// matcher entry points that dispatch between the exact SIMT kernel and the
// tcgen05 tensor-core kernel.
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "mnn_common.cuh"

namespace posfeat {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mu;
struct ProfPair { cudaEvent_t a, b; };
static std::vector<ProfPair> g_prof[PROF_COUNT];
static cudaEvent_t g_prof_open[PROF_COUNT];
static const char* kProfNames[PROF_COUNT] = {"nms_candidates", "select_topk", "sample_l2norm", "mnn_prep", "mnn_tc",
                                             "mnn_rescore", "mnn_simt", "mnn_compact", "corr_fwd", "corr_bwd",
                                             "window_fwd", "window_bwd", "mnn_scan", "mnn_verify", "fetch_taps",
                                             "keypoint_outputs"};

void prof_begin(int slot, cudaStream_t s) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, s);
  g_prof_open[slot] = e;
}
void prof_end(int slot, cudaStream_t s) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_prof_open[slot]) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, s);
  g_prof[slot].push_back({g_prof_open[slot], e});
  g_prof_open[slot] = nullptr;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

struct MnnScratch {
  float *A, *B;
  int32_t *nn12, *nn21, *n_matches;
  int64_t* matches;
  void* ws;
  size_t ws_bytes, total;
};

static MnnScratch carve_host_scratch(void* base, int N, int M, int D, int algo) {
  MnnScratch s;
  size_t off = 0;
  char* p = (char*)base;
  auto take = [&](size_t bytes) {
    char* r = p ? p + off : nullptr;
    off += align_up(bytes, 256);
    return r;
  };
  s.A = (float*)take(sizeof(float) * (size_t)N * D);
  s.B = (float*)take(sizeof(float) * (size_t)M * D);
  s.nn12 = (int32_t*)take(sizeof(int32_t) * N);
  s.nn21 = (int32_t*)take(sizeof(int32_t) * M);
  s.n_matches = (int32_t*)take(sizeof(int32_t));
  s.matches = (int64_t*)take(sizeof(int64_t) * 2 * (size_t)N);
  s.ws_bytes = posfeat_mnn_workspace_bytes(N, M, D, algo);
  s.ws = take(s.ws_bytes);
  s.total = off;
  return s;
}

}  // namespace posfeat

using namespace posfeat;

extern "C" int posfeat_version(void) { return 100; }

extern "C" int posfeat_last_error(char* buf, int n) {
  if (!buf || n <= 0) return POSFEAT_EINVAL;
  strncpy(buf, g_err, (size_t)n - 1);
  buf[n - 1] = 0;
  return POSFEAT_OK;
}

extern "C" int64_t posfeat_launch_count(void) { return (int64_t)g_launches.load(); }

extern "C" int posfeat_host_device_pointer(const void* host, void** dev_out) {
  PF_CHECK_ARG(host && dev_out, "NULL pointer");
  void* d = nullptr;
  cudaError_t e = cudaHostGetDevicePointer(&d, const_cast<void*>(host), 0);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_error(POSFEAT_EINVAL, "host pointer %p is not page-locked, mapped memory: %s", host,
                     cudaGetErrorString(e));
  }
  *dev_out = d;
  return POSFEAT_OK;
}
extern "C" int posfeat_profile_enable(int on) {
  g_prof_on.store(on ? 1 : 0);
  return POSFEAT_OK;
}
extern "C" int posfeat_profile_slot_count(void) { return PROF_COUNT; }
extern "C" const char* posfeat_profile_slot_name(int slot) {
  return (slot >= 0 && slot < PROF_COUNT) ? kProfNames[slot] : "";
}
extern "C" int posfeat_profile_read(int slot, double* total_ms, int32_t* launches) {
  PF_CHECK_ARG(slot >= 0 && slot < PROF_COUNT && total_ms && launches, "bad profile slot %d", slot);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  double ms = 0.0;
  int n = 0;
  for (ProfPair& pr : g_prof[slot]) {
    float t = 0.f;
    if (cudaEventSynchronize(pr.b) == cudaSuccess && cudaEventElapsedTime(&t, pr.a, pr.b) == cudaSuccess) {
      ms += t;
      ++n;
    }
    cudaEventDestroy(pr.a);
    cudaEventDestroy(pr.b);
  }
  g_prof[slot].clear();
  *total_ms = ms;
  *launches = n;
  return POSFEAT_OK;
}

extern "C" int posfeat_device_sm_count(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  return n;
}

static int resolve_algo(int N, int M, int D, int algo) {
  if (algo == POSFEAT_MNN_AUTO) return (tc_supported(N, M, D) && (int64_t)N * M >= (int64_t)1024 * 1024) ? POSFEAT_MNN_TC : POSFEAT_MNN_SIMT;
  return algo;
}

extern "C" size_t posfeat_mnn_batched_workspace_bytes(int P, int N, int M, int D, int algo) {
  if (P < 1 || N < 1 || M < 1 || D < 1) return 0;
  size_t need = simt_workspace_bytes(N, M);
  const int a = resolve_algo(N, M, D, algo);
  if (a == POSFEAT_MNN_TC && tc_supported(N, M, D)) {
    const size_t t = tc_workspace_bytes(P, N, M);
    if (t > need) need = t;
  }
  return need;
}

extern "C" size_t posfeat_mnn_workspace_bytes(int N, int M, int D, int algo) {
  return posfeat_mnn_batched_workspace_bytes(1, N, M, D, algo);
}

extern "C" int posfeat_mnn_batched_f32(const float* A, int64_t stride_a, int N, int64_t lda, const float* Bm,
                                       int64_t stride_b, int M, int64_t ldb, int D, int P, int algo, int32_t* nn12,
                                       int32_t* nn21, int64_t* matches, int32_t* n_matches, void* workspace,
                                       size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PF_CHECK_ARG(A && Bm && nn12 && matches && n_matches && workspace, "NULL pointer");
  PF_CHECK_ARG(P >= 1 && P <= 65535, "pairs=%d outside [1, 65535]", P);
  PF_CHECK_ARG(N >= 1 && M >= 1 && D >= 1, "empty operand (N=%d M=%d D=%d): the reference's torch.max raises on an empty reduction too", N, M, D);
  PF_CHECK_ARG(lda >= D && ldb >= D, "row stride smaller than D");
  const bool prepared = (algo & POSFEAT_MNN_PREPARED) != 0;
  algo &= ~POSFEAT_MNN_PREPARED;
  PF_CHECK_ARG(algo >= POSFEAT_MNN_AUTO && algo <= POSFEAT_MNN_TC, "unknown algo %d", algo);
  PF_CHECK_ARG(!prepared || (algo == POSFEAT_MNN_TC && N == M), "POSFEAT_MNN_PREPARED goes with POSFEAT_MNN_TC and N == M");
  const int a = resolve_algo(N, M, D, algo);
  const size_t need = posfeat_mnn_batched_workspace_bytes(P, N, M, D, a);
  if (ws_bytes < need) return set_error(POSFEAT_EWORKSPACE, "mnn workspace: need %zu bytes, got %zu", need, ws_bytes);
  if (a == POSFEAT_MNN_TC) {
    if (!tc_supported(N, M, D))
      return set_error(POSFEAT_EUNSUPPORTED, "tensor-core matcher needs D == 128 (got D=%d)", D);
    return mnn_tc(A, stride_a, N, lda, Bm, stride_b, M, ldb, D, P, nn12, nn21, matches, n_matches, workspace, ws_bytes,
                  stream, nullptr, nullptr, prepared);
  } else {
    PF_CHECK_ARG(nn21 != nullptr, "the exact SIMT matcher needs an nn21 buffer");
    for (int p = 0; p < P; ++p)
      if (int e = mnn_simt(A + p * stride_a, N, lda, Bm + p * stride_b, M, ldb, D, nn12 + (size_t)p * N,
                           nn21 + (size_t)p * M, workspace, stream))
        return e;
  }
  return launch_mutual_compact_batched(nn12, nn21, P, N, M, matches, n_matches, stream);
}

extern "C" int posfeat_mnn_f32(const float* A, int N, int64_t lda, const float* Bm, int M, int64_t ldb, int D,
                               int algo, int32_t* nn12, int32_t* nn21, int64_t* matches, int32_t* n_matches,
                               void* workspace, size_t ws_bytes, void* stream) {
  return posfeat_mnn_batched_f32(A, 0, N, lda, Bm, 0, M, ldb, D, 1, algo, nn12, nn21, matches, n_matches, workspace,
                                 ws_bytes, stream);
}

// ---- ratio-test matchers (evaluations/aachen/matchers.py:17-75, ETH custom_matcher.py:16-73)
// Same algorithm choice as the mutual-NN matcher: large D == 128 problems contract on the tensor cores
// (both directions, top-2 rescoring), everything else on the exact SIMT kernel.
static size_t ratio_core_bytes(int N, int M, int D) {
  const bool tc = resolve_algo(N, M, D, POSFEAT_MNN_AUTO) == POSFEAT_MNN_TC;
  return align_up(tc ? tc_workspace_bytes(1, N, M) : simt_workspace_bytes(N, M), 256);
}

extern "C" size_t posfeat_ratio_match_workspace_bytes(int N, int M, int D) {
  if (N < 2 || M < 2 || D < 1) return 0;
  return ratio_core_bytes(N, M, D) + align_up(sizeof(int32_t) * (size_t)M, 256) +
         align_up(sizeof(float) * 2 * (size_t)N, 256) + align_up(sizeof(float) * 2 * (size_t)M, 256) +
         align_up((size_t)N, 256);
}

extern "C" int posfeat_ratio_match_f32(const float* A, int N, int64_t lda, const float* Bm, int M, int64_t ldb, int D,
                                       float ratio, int mutual, int32_t* nn12, int64_t* matches, int32_t* n_matches,
                                       void* workspace, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PF_CHECK_ARG(A && Bm && nn12 && matches && n_matches && workspace, "NULL pointer");
  PF_CHECK_ARG(N >= 2 && M >= 2 && D >= 1, "the ratio test needs at least two descriptors on each side (torch.topk(2) raises otherwise)");
  PF_CHECK_ARG(lda >= D && ldb >= D, "row stride smaller than D");
  PF_CHECK_ARG(((uintptr_t)workspace & 255) == 0, "ratio workspace must be 256-byte aligned");
  const size_t need = posfeat_ratio_match_workspace_bytes(N, M, D);
  if (ws_bytes < need) return set_error(POSFEAT_EWORKSPACE, "ratio workspace: need %zu bytes, got %zu", need, ws_bytes);
  char* w = (char*)workspace;
  const size_t core = ratio_core_bytes(N, M, D);
  void* core_ws = w; w += core;
  int32_t* nn21 = (int32_t*)w; w += align_up(sizeof(int32_t) * (size_t)M, 256);
  float* top12 = (float*)w; w += align_up(sizeof(float) * 2 * (size_t)N, 256);
  float* top21 = (float*)w; w += align_up(sizeof(float) * 2 * (size_t)M, 256);
  unsigned char* flags = (unsigned char*)w;
  if (resolve_algo(N, M, D, POSFEAT_MNN_AUTO) == POSFEAT_MNN_TC && !getenv("POSFEAT_RATIO_SIMT")) {
    if (int e = mnn_tc(A, 0, N, lda, Bm, 0, M, ldb, D, 1, nn12, nn21, matches, n_matches, core_ws, core, stream, top12, top21))
      return e;
  } else {
    if (int e = mnn_simt_top2(A, N, lda, Bm, M, ldb, D, nn12, nn21, top12, top21, core_ws, stream)) return e;
  }
  if (int e = launch_ratio_flags(nn12, nn21, top12, top21, N, M, ratio, mutual, flags, stream)) return e;
  return launch_compact_flags(nn12, flags, 1, N, matches, n_matches, stream);
}

extern "C" size_t posfeat_mnn_host_scratch_bytes(int N, int M, int D, int algo) {
  if (N < 1 || M < 1 || D < 1) return 0;
  return carve_host_scratch(nullptr, N, M, D, algo).total;
}

extern "C" int posfeat_mnn_host_f32(const float* A_host, int N, const float* B_host, int M, int D, int algo,
                                    int64_t* matches_host, int32_t* n_matches_host, void* dev_scratch,
                                    size_t scratch_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PF_CHECK_ARG(A_host && B_host && matches_host && n_matches_host && dev_scratch, "NULL pointer");
  PF_CHECK_ARG(N >= 1 && M >= 1 && D >= 1, "empty operand (N=%d M=%d D=%d)", N, M, D);
  MnnScratch s = carve_host_scratch(dev_scratch, N, M, D, algo);
  if (scratch_bytes < s.total) return set_error(POSFEAT_EWORKSPACE, "mnn host scratch: need %zu bytes, got %zu", s.total, scratch_bytes);
  PF_CUDA(cudaMemcpyAsync(s.A, A_host, sizeof(float) * (size_t)N * D, cudaMemcpyHostToDevice, stream));
  PF_CUDA(cudaMemcpyAsync(s.B, B_host, sizeof(float) * (size_t)M * D, cudaMemcpyHostToDevice, stream));
  int e = posfeat_mnn_f32(s.A, N, D, s.B, M, D, D, algo, s.nn12, s.nn21, s.matches, s.n_matches, s.ws, s.ws_bytes, stream);
  if (e) return e;
  PF_CUDA(cudaMemcpyAsync(n_matches_host, s.n_matches, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  PF_CUDA(cudaStreamSynchronize(stream));
  const int k = *n_matches_host;
  if (k > 0) {
    PF_CUDA(cudaMemcpyAsync(matches_host, s.matches, sizeof(int64_t) * 2 * (size_t)k, cudaMemcpyDeviceToHost, stream));
    PF_CUDA(cudaStreamSynchronize(stream));
  }
  return POSFEAT_OK;
}
