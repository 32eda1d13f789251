// Subsystem (3), exact path: mutual nearest neighbours without materialising
// the N x M similarity matrix, CUDA-core kernel with float64 accumulation.
//
// Replaces mnn_matcher / mutual_nn_matcher (reference
// evaluations/hpatches/evaluation.py:27-38, evaluations/aachen/matchers.py:5-13,
// evaluations/ETH_local_feature/custom_matcher.py:5-13,
// losses/preprocess_utils.py:795-803):
//   sim = A @ B.T ; nn12 = argmax_j ; nn21 = argmax_i ; keep i with nn21[nn12[i]] == i.
// Products of float32 numbers are exact in float64, so the accumulated value is
// the real dot product to ~1e-16: the argmax is independent of summation order
// and exact ties (duplicated descriptors) resolve to the first index like
// torch.max.  This kernel is also the rescoring reference of the tensor-core path.
#include "common.cuh"
#include "mnn_common.cuh"

namespace posfeat {

constexpr int kBM = 64, kBN = 64, kBK = 32, kPitch = 66;

// rows of X (64 per CTA) against a slice of the rows of Y; per-row best (value,
// index) of the slice goes to part_val/part_idx[split][row].
__global__ void __launch_bounds__(256)
rowbest_simt_kernel(const float* __restrict__ X, int NX, int64_t ldx, const float* __restrict__ Y, int NY,
                    int64_t ldy, int D, int cols_per_split, double* __restrict__ part_val,
                    int32_t* __restrict__ part_idx, double* __restrict__ part_sec) {
  __shared__ __align__(16) double Xs[kBK][kPitch];
  __shared__ __align__(16) double Ys[kBK][kPitch];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int row0 = blockIdx.x * kBM;
  const int split = blockIdx.y;
  const int c_begin = split * cols_per_split;
  const int c_end = min(NY, c_begin + cols_per_split);

  double bestv[4], secv[4];      // best and second best value per row (second: Lowe ratio test)
  int besti[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { bestv[i] = -INFINITY; secv[i] = -INFINITY; besti[i] = 0x7fffffff; }

  for (int col0 = c_begin; col0 < c_end; col0 += kBN) {
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    for (int k0 = 0; k0 < D; k0 += kBK) {
      // 64 rows x 32 k per operand, coalesced along k
      for (int e = tid; e < kBM * kBK; e += 256) {
        const int r = e >> 5, k = e & 31;
        float xv = 0.f, yv = 0.f;
        if (k0 + k < D) {
          if (row0 + r < NX) xv = __ldg(X + (int64_t)(row0 + r) * ldx + k0 + k);
          if (col0 + r < c_end) yv = __ldg(Y + (int64_t)(col0 + r) * ldy + k0 + k);
        }
        Xs[k][r] = (double)xv;
        Ys[k][r] = (double)yv;
      }
      __syncthreads();
#pragma unroll 8
      for (int k = 0; k < kBK; ++k) {
        const double2 x01 = *reinterpret_cast<const double2*>(&Xs[k][ty * 4]);
        const double2 x23 = *reinterpret_cast<const double2*>(&Xs[k][ty * 4 + 2]);
        const double2 y01 = *reinterpret_cast<const double2*>(&Ys[k][tx * 4]);
        const double2 y23 = *reinterpret_cast<const double2*>(&Ys[k][tx * 4 + 2]);
        const double xs[4] = {x01.x, x01.y, x23.x, x23.y};
        const double ys[4] = {y01.x, y01.y, y23.x, y23.y};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fma(xs[i], ys[j], acc[i][j]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = col0 + tx * 4 + j;
        if (c < c_end) {
          if (acc[i][j] > bestv[i]) { secv[i] = bestv[i]; bestv[i] = acc[i][j]; besti[i] = c; }
          else if (acc[i][j] > secv[i]) secv[i] = acc[i][j];
        }
      }
  }
  // reduce over the 16 threads (tx) that share a row group: half-warp butterflies
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bestv[i], o);
      const int oi = __shfl_xor_sync(0xffffffffu, besti[i], o);
      const double os = __shfl_xor_sync(0xffffffffu, secv[i], o);
      secv[i] = fmax(fmax(secv[i], os), fmin(bestv[i], ov));
      if (ov > bestv[i] || (ov == bestv[i] && oi < besti[i])) { bestv[i] = ov; besti[i] = oi; }
    }
    const int r = row0 + ty * 4 + i;
    if (tx == 0 && r < NX) {
      part_val[(int64_t)split * NX + r] = bestv[i];
      part_idx[(int64_t)split * NX + r] = besti[i];
      if (part_sec) part_sec[(int64_t)split * NX + r] = secv[i];
    }
  }
}

__global__ void rowbest_finalize_kernel(const double* __restrict__ part_val, const int32_t* __restrict__ part_idx,
                                        const double* __restrict__ part_sec, int NX, int splits,
                                        int32_t* __restrict__ nn, float* __restrict__ top2) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= NX) return;
  double bv = -INFINITY, sv = -INFINITY;
  int bi = 0x7fffffff;
  for (int s = 0; s < splits; ++s) {
    const double v = part_val[(int64_t)s * NX + r];
    const int i = part_idx[(int64_t)s * NX + r];
    if (part_sec) sv = fmax(fmax(sv, part_sec[(int64_t)s * NX + r]), fmin(bv, v));
    if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
  }
  nn[r] = bi == 0x7fffffff ? 0 : bi;   // all-NaN rows: torch.max also reports an index
  if (top2) { top2[2 * r] = (float)bv; top2[2 * r + 1] = (float)sv; }   // the reference's sim is float32
}

// Lowe ratio test of evaluations/aachen/matchers.py:17-75 on the top-2 similarities:
// dist = sqrt(2 - 2 sim), ratio = d0 / (d1 + 1e-8); keep i iff ratio12[i] <= ratio and
// ratio21[nn12[i]] <= ratio (and, for the mutual variant, nn21[nn12[i]] == i).
__global__ void ratio_flags_kernel(const int32_t* __restrict__ nn12, const int32_t* __restrict__ nn21,
                                   const float* __restrict__ top12, const float* __restrict__ top21, int N, int M,
                                   float ratio, int mutual, unsigned char* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  auto lowe = [](float s0, float s1) {
    const float d0 = sqrtf(2.f - 2.f * s0), d1 = sqrtf(2.f - 2.f * s1);
    return d0 / (d1 + 1e-8f);
  };
  const int j = nn12[i];
  bool keep = lowe(top12[2 * i], top12[2 * i + 1]) <= ratio;
  keep = keep && j >= 0 && j < M && lowe(top21[2 * j], top21[2 * j + 1]) <= ratio;   // NaN ratios fail like torch
  if (mutual) keep = keep && nn21[j] == i;
  flags[i] = keep ? 1 : 0;
}

// keep rows with nn21[nn12[i]] == i, ordered compaction (ascending i); one CTA per
// pair, kItems consecutive rows per thread so all loads of a pass are in flight together
constexpr int kItems = 8;
__global__ void __launch_bounds__(1024)
mutual_compact_kernel(const int32_t* __restrict__ nn12_, const int32_t* __restrict__ nn21_, int N, int M,
                      int64_t* __restrict__ matches_, int32_t* __restrict__ n_matches) {
  __shared__ int s_warp[32];
  __shared__ int s_base, s_total;
  const int pair = blockIdx.x;
  const int32_t* nn12 = nn12_ + (size_t)pair * N;
  const int32_t* nn21 = nn21_ + (size_t)pair * M;
  int64_t* matches = matches_ + (size_t)pair * N * 2;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int i0 = 0; i0 < N; i0 += 1024 * kItems) {
    const int first = i0 + tid * kItems;
    int j[kItems], back[kItems];
#pragma unroll
    for (int k = 0; k < kItems; ++k) j[k] = (first + k < N) ? nn12[first + k] : -1;
#pragma unroll
    for (int k = 0; k < kItems; ++k) back[k] = (j[k] >= 0 && j[k] < M) ? nn21[j[k]] : -1;
    unsigned keep = 0;
#pragma unroll
    for (int k = 0; k < kItems; ++k) keep |= (back[k] == first + k && first + k < N) ? (1u << k) : 0u;
    const int mine = __popc(keep);
    int inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      const int v = s_warp[lane];
      int winc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += t;
      }
      s_warp[lane] = winc - v;  // exclusive prefix over warps
      if (lane == 31) s_total = winc;
    }
    __syncthreads();
    int pos = s_base + s_warp[wid] + inc - mine;
#pragma unroll
    for (int k = 0; k < kItems; ++k)
      if (keep & (1u << k)) {
        matches[2 * (int64_t)pos] = first + k;
        matches[2 * (int64_t)pos + 1] = j[k];
        ++pos;
      }
    __syncthreads();
    if (tid == 0) s_base += s_total;
    __syncthreads();
  }
  if (tid == 0) n_matches[pair] = s_base;
}

int launch_mutual_compact_batched(const int32_t* nn12, const int32_t* nn21, int P, int N, int M, int64_t* matches,
                                  int32_t* n_matches, cudaStream_t stream) {
  ProfScope prof(PROF_MNN_COMPACT, stream);
  mutual_compact_kernel<<<P, 1024, 0, stream>>>(nn12, nn21, N, M, matches, n_matches);
  PF_LAUNCH_CHECK("mutual_compact_kernel");
  return POSFEAT_OK;
}

int launch_mutual_compact(const int32_t* nn12, const int32_t* nn21, int N, int M, int64_t* matches,
                          int32_t* n_matches, cudaStream_t stream) {
  return launch_mutual_compact_batched(nn12, nn21, 1, N, M, matches, n_matches, stream);
}

static int choose_splits(int NX, int NY) {
  const int row_blocks = (NX + kBM - 1) / kBM;
  const int col_tiles = (NY + kBN - 1) / kBN;
  int want = (2 * sm_count() + row_blocks - 1) / row_blocks;
  if (want < 1) want = 1;
  if (want > col_tiles) want = col_tiles;
  if (want > 64) want = 64;
  return want;
}

size_t simt_workspace_bytes(int N, int M) {
  // partial (value, second value, index) per split and row, both directions (splits <= 64)
  const size_t per = 2 * sizeof(double) + sizeof(int32_t);
  return align_up((size_t)64 * N * per, 256) + align_up((size_t)64 * M * per, 256) + 1024;
}

int run_rowbest_simt(const float* X, int NX, int64_t ldx, const float* Y, int NY, int64_t ldy, int D,
                     int32_t* nn, float* top2, void* ws, cudaStream_t stream) {
  const int splits = choose_splits(NX, NY);
  const int col_tiles = (NY + kBN - 1) / kBN;
  const int cols_per_split = ((col_tiles + splits - 1) / splits) * kBN;
  const int used = (NY + cols_per_split - 1) / cols_per_split;
  double* pv = (double*)ws;
  double* ps = top2 ? pv + (size_t)64 * NX : nullptr;
  int32_t* pi = (int32_t*)((char*)ws + align_up(2 * sizeof(double) * (size_t)64 * NX, 256));
  dim3 grid((NX + kBM - 1) / kBM, used);
  ProfScope prof(PROF_MNN_SIMT, stream);
  rowbest_simt_kernel<<<grid, 256, 0, stream>>>(X, NX, ldx, Y, NY, ldy, D, cols_per_split, pv, pi, ps);
  PF_LAUNCH_CHECK("rowbest_simt_kernel");
  rowbest_finalize_kernel<<<(NX + 255) / 256, 256, 0, stream>>>(pv, pi, ps, NX, used, nn, top2);
  PF_LAUNCH_CHECK("rowbest_finalize_kernel");
  return POSFEAT_OK;
}

int mnn_simt(const float* A, int N, int64_t lda, const float* Bm, int M, int64_t ldb, int D, int32_t* nn12,
             int32_t* nn21, void* ws, cudaStream_t stream) {
  return mnn_simt_top2(A, N, lda, Bm, M, ldb, D, nn12, nn21, nullptr, nullptr, ws, stream);
}

int mnn_simt_top2(const float* A, int N, int64_t lda, const float* Bm, int M, int64_t ldb, int D, int32_t* nn12,
                  int32_t* nn21, float* top12, float* top21, void* ws, cudaStream_t stream) {
  char* w = (char*)ws;
  const size_t per = 2 * sizeof(double) + sizeof(int32_t);
  if (int e = run_rowbest_simt(A, N, lda, Bm, M, ldb, D, nn12, top12, w, stream)) return e;
  w += align_up((size_t)64 * N * per, 256) + 512;
  return run_rowbest_simt(Bm, M, ldb, A, N, lda, D, nn21, top21, w, stream);
}

int launch_ratio_flags(const int32_t* nn12, const int32_t* nn21, const float* top12, const float* top21, int N, int M,
                       float ratio, int mutual, unsigned char* flags, cudaStream_t stream) {
  ratio_flags_kernel<<<(N + 255) / 256, 256, 0, stream>>>(nn12, nn21, top12, top21, N, M, ratio, mutual, flags);
  PF_LAUNCH_CHECK("ratio_flags_kernel");
  return POSFEAT_OK;
}

}  // namespace posfeat
