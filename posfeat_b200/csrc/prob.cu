// compute_prob (losses/preprocess_utils.py:89-115) as a stand-alone entry point.
//
// The three training-side expectations (corr.cu / corr_tc.cu) fuse this function away -- they never write
// the [B,m,n] probabilities.  Callers that want the tensor itself (the reference returns it from
// get_expected_correspondence_locs(with_std=True), and compute_prob is importable on its own) get it here:
//   'cos':  prob = softmax_j(scale * <f1_i, f2_j>)            (scale = sqrt(n) with with_scale, else 1)
//   'euc':  prob = softmax_j(-(|f1_i|^2 + |f2_j|^2 - 2 <f1_i, f2_j>))
// Materialising B*m*n floats is the job, so the op is bound by writing them: pass 1 (a 64x64-tile SIMT
// contraction, float32 FMA) stores the logits (and the raw similarities when asked), pass 2 turns every
// row into probabilities in place (row kept in registers when n <= 8192, else three sweeps).
#include "common.cuh"

namespace posfeat {

constexpr int kPT = 64;      // logits tile (rows and columns)
constexpr int kPK = 16;      // K slice staged in shared memory

__global__ void __launch_bounds__(256)
prob_logits_kernel(const float* __restrict__ f1, const float* __restrict__ f2, int m, int n, int D, int mode,
                   float scale, float* __restrict__ logits, float* __restrict__ sim) {
  __shared__ float sa[kPK][kPT + 1], sb[kPK][kPT + 1];
  __shared__ float na[kPT], nb[kPT];
  const int b = blockIdx.z;
  const int i0 = blockIdx.y * kPT, j0 = blockIdx.x * kPT;
  const float* A = f1 + (size_t)b * m * D;
  const float* Bm = f2 + (size_t)b * n * D;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;      // 16 x 16 threads, 4 x 4 outputs each
  float acc[4][4] = {};
  float sq = 0.f;                                             // 'euc': |row|^2 of the row this thread stages
  const int lr = threadIdx.x >> 2, lk = (threadIdx.x & 3) * 4;   // staging: 64 rows x 16 k, 4 k per thread
  for (int k0 = 0; k0 < D; k0 += kPK) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = k0 + lk + q;
      const float va = (i0 + lr < m && k < D) ? __ldg(A + (size_t)(i0 + lr) * D + k) : 0.f;
      const float vb = (j0 + lr < n && k < D) ? __ldg(Bm + (size_t)(j0 + lr) * D + k) : 0.f;
      sa[lk + q][lr] = va;
      sb[lk + q][lr] = vb;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kPK; ++k) {
      float a[4], c[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) { a[r] = sa[k][ty * 4 + r]; c[r] = sb[k][tx * 4 + r]; }
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int s = 0; s < 4; ++s) acc[r][s] = fmaf(a[r], c[s], acc[r][s]);
    }
    __syncthreads();
  }
  if (mode == 1) {   // squared norms of the tile's rows / columns (threads 0..63: rows of A, 64..127: rows of B)
    if (threadIdx.x < 2 * kPT) {
      const bool second = threadIdx.x >= kPT;
      const int r = threadIdx.x & (kPT - 1);
      const int row = (second ? j0 : i0) + r;
      const float* src = (second ? Bm : A) + (size_t)row * D;
      if (row < (second ? n : m))
        for (int k = 0; k < D; ++k) { const float v = __ldg(src + k); sq = fmaf(v, v, sq); }
      (second ? nb : na)[r] = sq;
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ty * 4 + r;
    if (i >= m) continue;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int j = j0 + tx * 4 + s;
      if (j >= n) continue;
      const size_t o = ((size_t)b * m + i) * n + j;
      const float d = acc[r][s];
      if (sim) sim[o] = d;
      logits[o] = mode == 1 ? -((na[ty * 4 + r] + nb[tx * 4 + s]) - 2.f * d) : scale * d;
    }
  }
}

// one CTA per row: softmax in place
__global__ void __launch_bounds__(256)
prob_softmax_kernel(float* __restrict__ logits, int n) {
  __shared__ float red[8];
  __shared__ float s_bcast;
  float* row = logits + (size_t)blockIdx.x * n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  auto block_reduce = [&](float v, bool is_max) {
    v = is_max ? warp_max(v) : warp_sum(v);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = red[0];
      for (int k = 1; k < 8; ++k) t = is_max ? fmaxf(t, red[k]) : t + red[k];
      s_bcast = t;
    }
    __syncthreads();
    const float out = s_bcast;
    __syncthreads();
    return out;
  };
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < n; j += 256) mx = fmaxf(mx, row[j]);
  mx = block_reduce(mx, true);
  float sum = 0.f;
  for (int j = threadIdx.x; j < n; j += 256) {
    const float e = expf(row[j] - mx);
    row[j] = e;
    sum += e;
  }
  sum = block_reduce(sum, false);
  const float inv = 1.f / sum;
  for (int j = threadIdx.x; j < n; j += 256) row[j] *= inv;
}

}  // namespace posfeat

using namespace posfeat;

extern "C" int posfeat_compute_prob_f32(const float* f1, const float* f2, int B, int m, int n, int D, int mode,
                                        float scale, float* prob, float* sim, void* stream) {
  PF_CHECK_ARG(f1 && f2 && prob, "compute_prob: NULL pointer");
  PF_CHECK_ARG(B >= 0 && m >= 0 && n >= 1 && D >= 1, "compute_prob: bad shape B=%d m=%d n=%d D=%d", B, m, n, D);
  PF_CHECK_ARG(mode == 0 || mode == 1, "compute_prob: mode must be 0 ('cos') or 1 ('euc')");
  PF_CHECK_ARG(!(sim && mode == 1), "compute_prob: return_sim needs loss_distance 'cos'");
  PF_CHECK_ARG(B <= 65535, "compute_prob: at most 65535 batches");
  if (B == 0 || m == 0) return POSFEAT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 grid((n + kPT - 1) / kPT, (m + kPT - 1) / kPT, B);
  PF_CHECK_ARG(grid.y <= 65535, "compute_prob: m too large");
  prob_logits_kernel<<<grid, 256, 0, st>>>(f1, f2, m, n, D, mode, scale, prob, sim);
  PF_LAUNCH_CHECK("prob_logits_kernel");
  prob_softmax_kernel<<<(unsigned)((size_t)B * m), 256, 0, st>>>(prob, n);
  PF_LAUNCH_CHECK("prob_softmax_kernel");
  return POSFEAT_OK;
}
