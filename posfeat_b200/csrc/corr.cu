// Subsystem (4): training-side correlation + softmax expectation for sm_100a.
//
// Dense variant (get_expected_correspondence_locs, reference
// losses/preprocess_utils.py:55-82, and the grid<->grid stage of
// Preprocess_Line2Window.forward, losses/preprocess.py:59-81):
//     out[b,i,:] = sum_j softmax_j(scale * <q_i, k_j>) * v[j,:]
// computed flash-style: 64x64 logit tiles in registers, online softmax, the
// [B,n,m] probability tensor is never written.  The backward pass recomputes
// the logits from (q, k, lse); one kernel owns rows of q, one owns rows of k (atomics
// only when a long sweep over the other side is split across CTAs to fill the chip).
//
// Window / line variant (get_expected_correspondence_within_window,
// losses/preprocess_utils.py:721-758; the sampling half of epipolar_line_search,
// :662-675): per query, m sampling positions (centre + offset table, or a line
// between two endpoints) are bilinearly gathered on the fly from the feature
// map, dotted with the query and soft-maxed; the [B,n,m,D] gathered tensor of
// the reference (943 MB at B=8, n=1200) is never materialised.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

namespace posfeat {

// tensor-core forward path (corr_tc.cu)
bool corr_tc_eligible(int B, int n, int m, int D, int C);
size_t corr_tc_workspace_bytes(int B, int n, int m, int D, int C);
int corr_tc_fwd(const float* q, const float* k, const float* v, int v_batched, int B, int n, int m, int D, int C,
                float scale, float* out, float* lse, void* ws, size_t ws_bytes, cudaStream_t stream);
size_t corr_tc_bwd_workspace_bytes(int B, int n, int m, int D, int C);
bool corr_tc_bwd_eligible(int B, int n, int m, int D, int C);
size_t corr_disk_workspace_bytes(int B, int n, int m, int D);
int corr_disk_rows(const float* q, const float* k, const float* rowtab, const float* coltab, int B, int n, int m, int D,
                   float temp, float thr_own, float thr_other, float good_reward, float bad_reward, int dynamic_reward,
                   float* rows_out, void* ws, size_t ws_bytes, cudaStream_t stream);
int corr_tc_bwd(const float* q, const float* k, const float* v, int v_batched, int B, int n, int m, int D, int C,
                float scale, const float* out, const float* lse, const float* g_out, float* g_q, float* g_k, void* ws,
                size_t ws_bytes, cudaStream_t stream);

constexpr int kCT = 64;          // logit tile edge
constexpr int kCDmax = 128;      // descriptor length supported by the dense kernels
constexpr int kCmax = 4;         // value-table width

// ---------------------------------------------------------------------------
// dense forward
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
corr_expect_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                       int v_batched, int n, int m, int D, int C, float scale, float* __restrict__ out,
                       float* __restrict__ lse) {
  __shared__ __align__(16) float Qs[32][kCT + 4];   // [kk][row]
  __shared__ __align__(16) float Ks[32][kCT + 4];
  __shared__ float Vs[kCT][kCmax];
  const int b = blockIdx.y, row0 = blockIdx.x * kCT;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const float* qb = q + (size_t)b * n * D;
  const float* kb = k + (size_t)b * m * D;
  const float* vb = v + (v_batched ? (size_t)b * m * C : 0);

  float mrun[4], lpart[4], acc[4][kCmax];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    mrun[i] = -INFINITY; lpart[i] = 0.f;
#pragma unroll
    for (int c = 0; c < kCmax; ++c) acc[i][c] = 0.f;
  }
  for (int col0 = 0; col0 < m; col0 += kCT) {
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
    for (int k0 = 0; k0 < D; k0 += 32) {
      for (int e = tid; e < kCT * 32; e += 256) {
        const int r = e >> 5, kk = e & 31;
        float qv = 0.f, kv = 0.f;
        if (k0 + kk < D) {
          if (row0 + r < n) qv = __ldg(qb + (size_t)(row0 + r) * D + k0 + kk);
          if (col0 + r < m) kv = __ldg(kb + (size_t)(col0 + r) * D + k0 + kk);
        }
        Qs[kk][r] = qv;
        Ks[kk][r] = kv;
      }
      if (k0 == 0) {
        for (int e = tid; e < kCT * kCmax; e += 256) {
          const int r = e / kCmax, c = e - r * kCmax;
          Vs[r][c] = (c < C && col0 + r < m) ? __ldg(vb + (size_t)(col0 + r) * C + c) : 0.f;
        }
      }
      __syncthreads();
#pragma unroll 8
      for (int kk = 0; kk < 32; ++kk) {
        const float4 a = *reinterpret_cast<const float4*>(&Qs[kk][ty * 4]);
        const float4 bq = *reinterpret_cast<const float4*>(&Ks[kk][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) s[i][j] = fmaf(av[i], bv[j], s[i][j]);
      }
      __syncthreads();
    }
    // online softmax update (rows are shared by the 16 threads with the same ty)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float tmax = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[i][j] = (col0 + tx * 4 + j < m) ? s[i][j] * scale : -INFINITY;
        tmax = fmaxf(tmax, s[i][j]);
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
      const float mnew = fmaxf(mrun[i], tmax);
      const float corr = (mrun[i] == -INFINITY) ? 0.f : expf(mrun[i] - mnew);
      lpart[i] *= corr;
#pragma unroll
      for (int c = 0; c < kCmax; ++c) acc[i][c] *= corr;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float p = (s[i][j] == -INFINITY) ? 0.f : expf(s[i][j] - mnew);
        lpart[i] += p;
#pragma unroll
        for (int c = 0; c < kCmax; ++c) acc[i][c] = fmaf(p, Vs[tx * 4 + j][c], acc[i][c]);
      }
      mrun[i] = mnew;
    }
    __syncthreads();   // Vs is rewritten by the next tile
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      lpart[i] += __shfl_xor_sync(0xffffffffu, lpart[i], o);
#pragma unroll
      for (int c = 0; c < kCmax; ++c) acc[i][c] += __shfl_xor_sync(0xffffffffu, acc[i][c], o);
    }
    const int r = row0 + ty * 4 + i;
    if (tx == 0 && r < n) {
      const float inv = 1.f / lpart[i];
      for (int c = 0; c < C; ++c) out[((size_t)b * n + r) * C + c] = acc[i][c] * inv;
      lse[(size_t)b * n + r] = mrun[i] + logf(lpart[i]);
    }
  }
}

// ---------------------------------------------------------------------------
// dense backward: rows of X are owned by the CTA, Y is swept in 64-row tiles.
//   kXisQ = true : X = q, Y = k, result g_q      kXisQ = false : X = k, Y = q, result g_k
//   W(x,y) = exp(scale*<x,y> - lse[qi]) * (<g[qi], v[ki]> - <g[qi], out[qi]>)
//   gX[x,:] = scale * sum_y W(x,y) * Y[y,:]
// ---------------------------------------------------------------------------
template <bool kXisQ>
__global__ void __launch_bounds__(256)
corr_expect_bwd_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                       int v_batched, int n, int m, int D, int C, float scale, const float* __restrict__ out,
                       const float* __restrict__ lse, const float* __restrict__ g_out, float* __restrict__ gX,
                       int ytiles_per_split) {
  extern __shared__ __align__(16) float smem[];
  float* Xs = smem;                        // [kCDmax][kCT+4]  (k-major)
  float* Ys = Xs + kCDmax * (kCT + 4);     // [kCT][kCDmax+4]  (row-major: second GEMM reads rows)
  float* Ws = Ys + kCT * (kCDmax + 4);     // [kCT x][kCT+1 y]
  float* Yt = Ws + kCT * (kCT + 1);        // [32][kCT+4] k-major chunk of Y for the first GEMM
  __shared__ float s_lse[kCT], s_go[kCT], s_g[kCT][kCmax], s_v[kCT][kCmax];

  const int b = blockIdx.y, x0 = blockIdx.x * kCT;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int nX = kXisQ ? n : m, nY = kXisQ ? m : n;
  const float* Xg = (kXisQ ? q + (size_t)b * n * D : k + (size_t)b * m * D);
  const float* Yg = (kXisQ ? k + (size_t)b * m * D : q + (size_t)b * n * D);
  const float* vb = v + (v_batched ? (size_t)b * m * C : 0);
  const float* ob = out + (size_t)b * n * C;
  const float* gb = g_out + (size_t)b * n * C;
  const float* lb = lse + (size_t)b * n;

  // X tile, k-major, zero padded to kCDmax
  for (int e = tid; e < kCT * kCDmax; e += 256) {
    const int r = e / kCDmax, kk = e - r * kCDmax;
    Xs[kk * (kCT + 4) + r] = (kk < D && x0 + r < nX) ? __ldg(Xg + (size_t)(x0 + r) * D + kk) : 0.f;
  }
  // per-row side data of the owner side
  auto load_side = [&](int base, int count_q_side, bool is_q_side) {
    if (tid < kCT) {
      const int r = base + tid;
      if (is_q_side) {
        float go = 0.f;
        for (int c = 0; c < kCmax; ++c) {
          const float g = (c < C && r < n) ? __ldg(gb + (size_t)r * C + c) : 0.f;
          const float o = (c < C && r < n) ? __ldg(ob + (size_t)r * C + c) : 0.f;
          s_g[tid][c] = g;
          go = fmaf(g, o, go);
        }
        s_go[tid] = go;
        s_lse[tid] = r < n ? __ldg(lb + r) : INFINITY;      // exp(. - inf) = 0 for padding rows
      } else {
        for (int c = 0; c < kCmax; ++c) s_v[tid][c] = (c < C && r < m) ? __ldg(vb + (size_t)r * C + c) : 0.f;
      }
    }
    (void)count_q_side;
  };
  load_side(x0, 0, kXisQ);

  float acc[4][8];   // rows ty*4.., cols tx*8..  of gX tile [64][128]
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  // the Y sweep may be split over blockIdx.z (few X tiles, long Y): partial sums are then added atomically
  const int ybeg = blockIdx.z * ytiles_per_split * kCT;
  const int yend = min(nY, ybeg + ytiles_per_split * kCT);
  for (int y0 = ybeg; y0 < yend; y0 += kCT) {
    __syncthreads();
    // Y tile row-major (for W*Y) ...
    for (int e = tid; e < kCT * kCDmax; e += 256) {
      const int r = e / kCDmax, kk = e - r * kCDmax;
      Ys[r * (kCDmax + 4) + kk] = (kk < D && y0 + r < nY) ? __ldg(Yg + (size_t)(y0 + r) * D + kk) : 0.f;
    }
    load_side(y0, 0, !kXisQ);
    __syncthreads();
    // first GEMM: S = X * Y^T (64x64), Y re-staged k-major in 32-wide chunks
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
    for (int k0 = 0; k0 < D; k0 += 32) {
      for (int e = tid; e < kCT * 32; e += 256) {
        const int r = e >> 5, kk = e & 31;
        Yt[kk * (kCT + 4) + r] = Ys[r * (kCDmax + 4) + k0 + kk];
      }
      __syncthreads();
#pragma unroll 8
      for (int kk = 0; kk < 32; ++kk) {
        const float4 a = *reinterpret_cast<const float4*>(&Xs[(k0 + kk) * (kCT + 4) + ty * 4]);
        const float4 bq = *reinterpret_cast<const float4*>(&Yt[kk * (kCT + 4) + tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) s[i][j] = fmaf(av[i], bv[j], s[i][j]);
      }
      __syncthreads();
    }
    // W tile
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int xr = ty * 4 + i, yr = tx * 4 + j;
        const int qi = kXisQ ? xr : yr, ki = kXisQ ? yr : xr;
        float gv = 0.f;
#pragma unroll
        for (int c = 0; c < kCmax; ++c) gv = fmaf(s_g[qi][c], s_v[ki][c], gv);
        const bool valid = (x0 + xr < nX) && (y0 + yr < nY);
        const float p = valid ? expf(s[i][j] * scale - s_lse[qi]) : 0.f;
        Ws[xr * (kCT + 1) + yr] = p * (gv - s_go[qi]);
      }
    __syncthreads();
    // second GEMM: acc[64 x 128] += W[64 x 64] * Y[64 x 128]
#pragma unroll 4
    for (int yy = 0; yy < kCT; ++yy) {
      const float4 y0v = *reinterpret_cast<const float4*>(&Ys[yy * (kCDmax + 4) + tx * 8]);
      const float4 y1v = *reinterpret_cast<const float4*>(&Ys[yy * (kCDmax + 4) + tx * 8 + 4]);
      const float yv[8] = {y0v.x, y0v.y, y0v.z, y0v.w, y1v.x, y1v.y, y1v.z, y1v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float wv = Ws[(ty * 4 + i) * (kCT + 1) + yy];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(wv, yv[j], acc[i][j]);
      }
    }
  }
  float* gXb = gX + (size_t)b * nX * D;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = x0 + ty * 4 + i;
    if (r < nX)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = tx * 8 + j;
        if (c < D) {
          if (gridDim.z == 1) gXb[(size_t)r * D + c] = acc[i][j] * scale;
          else atomicAdd(gXb + (size_t)r * D + c, acc[i][j] * scale);
        }
      }
  }
}

// ---------------------------------------------------------------------------
// window / line gather variant
// ---------------------------------------------------------------------------
struct TapSet {
  int x0, y0;
  float w00, w01, w10, w11;
  bool in00, in01, in10, in11;
};
// align_corners=False unnormalisation; border=1 clamps the sampling coordinate
__device__ __forceinline__ TapSet taps_of(float gx, float gy, int h, int w, int border) {
  float ix = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)w), 1.f), 0.5f);
  float iy = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)h), 1.f), 0.5f);
  if (border) {
    ix = fminf(fmaxf(ix, 0.f), (float)(w - 1));
    iy = fminf(fmaxf(iy, 0.f), (float)(h - 1));
  }
  const float fx = floorf(ix), fy = floorf(iy);
  TapSet t;
  t.x0 = (int)fminf(fmaxf(fx, -2.f), (float)w);
  t.y0 = (int)fminf(fmaxf(fy, -2.f), (float)h);
  const float wx1 = ix - fx, wx0 = (fx + 1.f) - ix, wy1 = iy - fy, wy0 = (fy + 1.f) - iy;
  const bool xin0 = t.x0 >= 0 && t.x0 < w, xin1 = t.x0 + 1 >= 0 && t.x0 + 1 < w;
  const bool yin0 = t.y0 >= 0 && t.y0 < h, yin1 = t.y0 + 1 >= 0 && t.y0 + 1 < h;
  t.in00 = xin0 && yin0; t.in01 = xin1 && yin0; t.in10 = xin0 && yin1; t.in11 = xin1 && yin1;
  t.w00 = wy0 * wx0; t.w01 = wy0 * wx1; t.w10 = wy1 * wx0; t.w11 = wy1 * wx1;
  return t;
}

constexpr int kWinThreads = 128;
constexpr int kWinMaxPts = 1024;
constexpr int kWinCPL = 4;     // channels per lane: D <= 128

// position p of query (b,i): mode 0 = centre + offsets[p];  mode 1 = e1 + (e2-e1)*linspace(0,1,m)[p]
__device__ __forceinline__ float2 win_pos(int mode, const float* __restrict__ centre, const float* __restrict__ offsets,
                                          size_t qi, int p, int m) {
  if (mode == 0) {
    const float2 c = *reinterpret_cast<const float2*>(centre + qi * 2);
    const float2 o = *reinterpret_cast<const float2*>(offsets + (size_t)p * 2);
    return make_float2(c.x + o.x, c.y + o.y);
  }
  const float4 e = *reinterpret_cast<const float4*>(centre + qi * 4);   // (x1, y1, x2, y2)
  const float step = 1.0f / (float)(m - 1);
  const float t = (p < m / 2) ? step * (float)p : 1.0f - step * (float)(m - 1 - p);   // torch.linspace(0, 1, m)
  return make_float2(__fadd_rn(__fmul_rn(e.z - e.x, t), e.x), __fadd_rn(__fmul_rn(e.w - e.y, t), e.y));
}

__device__ __forceinline__ float gather_dot(const float* __restrict__ fb, const TapSet& t, int D, int64_t sc,
                                            int64_t sy, int64_t sx, const float (&qv)[kWinCPL], int lane,
                                            float (&sv)[kWinCPL]) {
  const float* base = fb + (int64_t)t.y0 * sy + (int64_t)t.x0 * sx;
  float dot = 0.f;
  if (sc == 1 && ((sx | sy) & 3) == 0 && (reinterpret_cast<uintptr_t>(fb) & 15) == 0 && (D & 3) == 0) {
    // channels innermost: one 128-bit load per corner and lane
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 a00 = z4, a01 = z4, a10 = z4, a11 = z4;
    if (lane * 4 < D) {
      const float4* pc = reinterpret_cast<const float4*>(base) + lane;
      if (t.in00) a00 = __ldg(pc);
      if (t.in01) a01 = __ldg(reinterpret_cast<const float4*>(base + sx) + lane);
      if (t.in10) a10 = __ldg(reinterpret_cast<const float4*>(base + sy) + lane);
      if (t.in11) a11 = __ldg(reinterpret_cast<const float4*>(base + sy + sx) + lane);
    }
    sv[0] = a00.x * t.w00 + a01.x * t.w01 + a10.x * t.w10 + a11.x * t.w11;
    sv[1] = a00.y * t.w00 + a01.y * t.w01 + a10.y * t.w10 + a11.y * t.w11;
    sv[2] = a00.z * t.w00 + a01.z * t.w01 + a10.z * t.w10 + a11.z * t.w11;
    sv[3] = a00.w * t.w00 + a01.w * t.w01 + a10.w * t.w10 + a11.w * t.w11;
#pragma unroll
    for (int j = 0; j < kWinCPL; ++j) dot = fmaf(sv[j], qv[j], dot);
    return warp_sum(dot);
  }
#pragma unroll
  for (int j = 0; j < kWinCPL; ++j) {
    const int c = (sc == 1) ? lane * kWinCPL + j : lane + 32 * j;
    float a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f;
    if (c < D) {
      const float* pc = base + (int64_t)c * sc;
      if (t.in00) a00 = __ldg(pc);
      if (t.in01) a01 = __ldg(pc + sx);
      if (t.in10) a10 = __ldg(pc + sy);
      if (t.in11) a11 = __ldg(pc + sy + sx);
    }
    sv[j] = a00 * t.w00 + a01 * t.w01 + a10 * t.w10 + a11 * t.w11;
    dot = fmaf(sv[j], qv[j], dot);
  }
  return warp_sum(dot);
}


// ---- window variant, "box" formulation (mode 0, channels-last map) -------------------------------------
// The m window positions of a query are a dense little grid (12 x 16 taps about one pixel apart), so their
// 4 m bilinear corners fall on only ~(w_x + 2)(w_y + 2) distinct pixels.  Because the dot product is linear,
//   <q, sum_k w_k f_k> = sum_k w_k <q, f_k>,
// the CTA first evaluates <q, f> once per pixel of the taps' bounding box (one coalesced 512-byte read each,
// pixels outside the map count as zero = zeros padding) and then blends those scalars per tap: 3x fewer
// descriptor reads than gathering every corner of every tap; the backward pass likewise accumulates one
// scalar per pixel in shared memory before it touches g_fmap (3x fewer atomics).
constexpr int kBoxMax = 1024;      // pixels in the bounding box handled this way (else: per-tap gather)

struct BoxInfo { int xmin, ymin, bw, bh; };

// bounding box of the corner pixels of all m taps (block-wide); returns false when it exceeds kBoxMax
__device__ __forceinline__ bool window_box(const float* __restrict__ centre, const float* __restrict__ offsets, size_t qi,
                                           int m, int h, int w, int (*red)[kWinThreads / 32], BoxInfo& box) {
  int x0 = 1 << 30, x1 = -(1 << 30), y0 = 1 << 30, y1 = -(1 << 30);
  for (int p = threadIdx.x; p < m; p += kWinThreads) {
    const float2 pos = win_pos(0, centre, offsets, qi, p, m);
    const TapSet t = taps_of(pos.x, pos.y, h, w, 0);
    x0 = min(x0, t.x0); x1 = max(x1, t.x0 + 1); y0 = min(y0, t.y0); y1 = max(y1, t.y0 + 1);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    x0 = min(x0, __shfl_xor_sync(0xffffffffu, x0, o)); x1 = max(x1, __shfl_xor_sync(0xffffffffu, x1, o));
    y0 = min(y0, __shfl_xor_sync(0xffffffffu, y0, o)); y1 = max(y1, __shfl_xor_sync(0xffffffffu, y1, o));
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = x0; red[1][warp] = x1; red[2][warp] = y0; red[3][warp] = y1; }
  __syncthreads();
  x0 = min(min(red[0][0], red[0][1]), min(red[0][2], red[0][3]));
  x1 = max(max(red[1][0], red[1][1]), max(red[1][2], red[1][3]));
  y0 = min(min(red[2][0], red[2][1]), min(red[2][2], red[2][3]));
  y1 = max(max(red[3][0], red[3][1]), max(red[3][2], red[3][3]));
  box.xmin = x0; box.ymin = y0; box.bw = x1 - x0 + 1; box.bh = y1 - y0 + 1;
  return (int64_t)box.bw * box.bh <= kBoxMax;
}

// sd[e] = <q, f[pixel e]> for every pixel of the box (0 outside the map); D % 4 == 0, channels innermost
__device__ __forceinline__ void box_dots(const float* __restrict__ fb, int D, int h, int w, int64_t sy, int64_t sx,
                                         const BoxInfo& box, const float4 q4, float* __restrict__ sd) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int npix = box.bw * box.bh;
  const bool act = lane * 4 < D;
  for (int e0 = warp * 4; e0 < npix; e0 += (kWinThreads / 32) * 4) {
    float d[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int e = e0 + k;
      const int by = e / box.bw, bx = e - by * box.bw;
      const int x = box.xmin + bx, y = box.ymin + by;
      float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e < npix && act && x >= 0 && x < w && y >= 0 && y < h)
        f = __ldg(reinterpret_cast<const float4*>(fb + (int64_t)y * sy + (int64_t)x * sx) + lane);
      d[k] = fmaf(f.w, q4.w, fmaf(f.z, q4.z, fmaf(f.y, q4.y, f.x * q4.x)));
    }
    // four warp sums with six shuffles: halve the value set while halving the lane set
    const bool h16 = lane & 16, h8 = lane & 8;
    float k0 = (h16 ? d[2] : d[0]) + __shfl_xor_sync(0xffffffffu, h16 ? d[0] : d[2], 16);
    float k1 = (h16 ? d[3] : d[1]) + __shfl_xor_sync(0xffffffffu, h16 ? d[1] : d[3], 16);
    float r = (h8 ? k1 : k0) + __shfl_xor_sync(0xffffffffu, h8 ? k0 : k1, 8);
    r += __shfl_xor_sync(0xffffffffu, r, 4);
    r += __shfl_xor_sync(0xffffffffu, r, 2);
    r += __shfl_xor_sync(0xffffffffu, r, 1);
    const int idx = (h16 ? 2 : 0) + (h8 ? 1 : 0);
    if ((lane & 7) == 0 && e0 + idx < npix) sd[e0 + idx] = r;
  }
}

// one CTA (4 warps) per query
__global__ void __launch_bounds__(kWinThreads)
window_expect_fwd_kernel(const float* __restrict__ fmap, int D, int h, int w, int64_t sb, int64_t sc, int64_t sy,
                         int64_t sx, const float* __restrict__ q, const float* __restrict__ centre, int n,
                         const float* __restrict__ offsets, int m, int mode, float* __restrict__ exp_xy,
                         float* __restrict__ std_out, float* __restrict__ prob, float* __restrict__ lse_out) {
  __shared__ float z[kWinMaxPts], px[kWinMaxPts], py[kWinMaxPts];
  __shared__ float red[8][kWinThreads / 32];
  const int b = blockIdx.y, i = blockIdx.x;
  const size_t qi = (size_t)b * n + i;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* fb = fmap + b * sb;
  float qv[kWinCPL];
#pragma unroll
  for (int j = 0; j < kWinCPL; ++j) {
    const int c = (sc == 1) ? lane * kWinCPL + j : lane + 32 * j;
    qv[j] = c < D ? __ldg(q + qi * D + c) : 0.f;
  }
  __shared__ float sd[kBoxMax];
  __shared__ int ired[4][kWinThreads / 32];
  BoxInfo box;
  bool boxed = false;
  if (mode == 0 && sc == 1 && (D & 3) == 0 && D <= 128 && (sx & 3) == 0 && (sy & 3) == 0 && (sb & 3) == 0 &&
      (reinterpret_cast<uintptr_t>(fmap) & 15) == 0)
    boxed = window_box(centre, offsets, qi, m, h, w, ired, box);          // block uniform
  if (boxed) {
    const float4 q4 = make_float4(qv[0], qv[1], qv[2], qv[3]);            // sc == 1: lane holds channels 4*lane .. +3
    box_dots(fb, D, h, w, sy, sx, box, q4, sd);
    __syncthreads();
    for (int p = threadIdx.x; p < m; p += kWinThreads) {
      const float2 pos = win_pos(0, centre, offsets, qi, p, m);
      const TapSet t = taps_of(pos.x, pos.y, h, w, 0);
      const float* s0 = sd + (t.y0 - box.ymin) * box.bw + (t.x0 - box.xmin);
      z[p] = s0[0] * t.w00 + s0[1] * t.w01 + s0[box.bw] * t.w10 + s0[box.bw + 1] * t.w11;
      px[p] = pos.x; py[p] = pos.y;
    }
  } else {
    for (int p = warp; p < m; p += kWinThreads / 32) {
      const float2 pos = win_pos(mode, centre, offsets, qi, p, m);
      const TapSet t = taps_of(pos.x, pos.y, h, w, mode);
      float sv[kWinCPL];
      const float d = gather_dot(fb, t, D, sc, sy, sx, qv, lane, sv);
      if (lane == 0) { z[p] = d; px[p] = pos.x; py[p] = pos.y; }
    }
  }
  __syncthreads();
  // softmax statistics over the m positions (block reduction)
  float mx = -INFINITY;
  for (int p = threadIdx.x; p < m; p += kWinThreads) mx = fmaxf(mx, z[p]);
  mx = warp_max(mx);
  if (lane == 0) red[0][warp] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red[0][0], red[0][1]), fmaxf(red[0][2], red[0][3]));
  float se = 0.f, sxv = 0.f, syv = 0.f, sxx = 0.f, syy = 0.f;
  for (int p = threadIdx.x; p < m; p += kWinThreads) {
    const float e = expf(z[p] - mx);
    se += e; sxv = fmaf(e, px[p], sxv); syv = fmaf(e, py[p], syv);
    sxx = fmaf(e, px[p] * px[p], sxx); syy = fmaf(e, py[p] * py[p], syy);
  }
  se = warp_sum(se); sxv = warp_sum(sxv); syv = warp_sum(syv); sxx = warp_sum(sxx); syy = warp_sum(syy);
  if (lane == 0) { red[1][warp] = se; red[2][warp] = sxv; red[3][warp] = syv; red[4][warp] = sxx; red[5][warp] = syy; }
  __syncthreads();
  float tot[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) tot[k] = red[1 + k][0] + red[1 + k][1] + red[1 + k][2] + red[1 + k][3];
  const float inv = 1.f / tot[0];
  const float ex = tot[1] * inv, ey = tot[2] * inv;
  if (threadIdx.x == 0) {
    exp_xy[qi * 2 + 0] = ex;
    exp_xy[qi * 2 + 1] = ey;
    const float vx = tot[3] * inv - ex * ex, vy = tot[4] * inv - ey * ey;
    std_out[qi] = sqrtf(fmaxf(vx, 1e-10f)) + sqrtf(fmaxf(vy, 1e-10f));
    lse_out[qi] = mx + logf(tot[0]);
  }
  if (prob)
    for (int p = threadIdx.x; p < m; p += kWinThreads) prob[qi * m + p] = expf(z[p] - mx) * inv;
}

// ---- epipolar line search in one kernel (epipolar_line_search, losses/preprocess_utils.py:662-694, with
// get_endpoints :697-719 folded in).  Per query: the epipolar line F x~ is clipped to the image rectangle
// (two of the four border intersections must lie inside, else the left/right intersections are used and the
// query is flagged invalid), m positions between the two endpoints are sampled (bilinear, border padding),
// dotted with the query and soft-maxed.  Outputs per query: the endpoints, the validity flag, the soft
// expectation, the sum of the positions that attain the largest probability (the reference's
// `(prob == prob.max) * grids` sum) and sum_p prob_p * g_p^2 for the variance; the [B,n,m,D] gathered tensor
// and, unless asked for, the probabilities are never written.
__global__ void __launch_bounds__(kWinThreads)
line_search_kernel(const float* __restrict__ fmap, int D, int h, int w, int64_t sb, int64_t sc, int64_t sy, int64_t sx,
                   const float* __restrict__ q, const float* __restrict__ coord_px, const float* __restrict__ Fmat,
                   int n, int imh, int imw, int m, float* __restrict__ ends_out, unsigned char* __restrict__ valid_out,
                   float* __restrict__ exp_soft, float* __restrict__ nn_xy, float* __restrict__ m2_out,
                   float* __restrict__ prob) {
  __shared__ float z[kWinMaxPts], px[kWinMaxPts], py[kWinMaxPts];
  __shared__ float red[8][kWinThreads / 32];
  __shared__ float s_ends[4];
  const int b = blockIdx.y, i = blockIdx.x;
  const size_t qi = (size_t)b * n + i;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    // line = F [x, y, 1]^T accumulated like a K = 3 SGEMM: fma(F2, 1, fma(F1, y, F0 * x))
    const float* F = Fmat + (size_t)b * 9;
    const float x = coord_px[qi * 2], y = coord_px[qi * 2 + 1];
    const float la = __fadd_rn(__fmaf_rn(F[1], y, __fmul_rn(F[0], x)), F[2]);
    const float lb = __fadd_rn(__fmaf_rn(F[4], y, __fmul_rn(F[3], x)), F[5]);
    const float lc = __fadd_rn(__fmaf_rn(F[7], y, __fmul_rn(F[6], x)), F[8]);
    const float wm = (float)(imw - 1), hm = (float)(imh - 1);
    float ptx[4], pty[4];
    ptx[0] = 0.f;  pty[0] = __fdiv_rn(-lc, lb);                                             // left border
    ptx[1] = wm;   pty[1] = __fdiv_rn(-__fadd_rn(__fmul_rn(la, wm), lc), lb);               // right border
    ptx[2] = __fdiv_rn(-__fadd_rn(__fmul_rn(lb, hm), lc), la); pty[2] = hm;                 // y = h - 1
    ptx[3] = __fdiv_rn(-lc, la); pty[3] = 0.f;                                              // y = 0
    bool in[4];
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      in[k] = ptx[k] >= 0.f && ptx[k] <= wm && pty[k] >= 0.f && pty[k] <= hm;
      cnt += in[k];
    }
    const bool valid = cnt == 2;
    if (!valid) { in[0] = in[1] = true; in[2] = in[3] = false; }
    float e[4];
    int o = 0;
    const float cx = __fmul_rn(wm, 0.5f), cy = __fmul_rn(hm, 0.5f);     // (w-1)/2, (h-1)/2
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (in[k] && o < 4) {
        e[o++] = __fdiv_rn(__fsub_rn(ptx[k], cx), cx);
        e[o++] = __fdiv_rn(__fsub_rn(pty[k], cy), cy);
      }
#pragma unroll
    for (int k = 0; k < 4; ++k) { s_ends[k] = e[k]; ends_out[qi * 4 + k] = e[k]; }
    valid_out[qi] = valid ? 1 : 0;
  }
  __syncthreads();
  const float e0 = s_ends[0], e1 = s_ends[1], e2 = s_ends[2], e3 = s_ends[3];
  const float* fb = fmap + b * sb;
  float qv[kWinCPL];
#pragma unroll
  for (int j = 0; j < kWinCPL; ++j) {
    const int c = (sc == 1) ? lane * kWinCPL + j : lane + 32 * j;
    qv[j] = c < D ? __ldg(q + qi * D + c) : 0.f;
  }
  const float step = 1.0f / (float)(m - 1);
  for (int p = warp; p < m; p += kWinThreads / 32) {
    const float t = (p < m / 2) ? step * (float)p : 1.0f - step * (float)(m - 1 - p);     // torch.linspace(0, 1, m)
    const float gx = __fadd_rn(__fmul_rn(e2 - e0, t), e0), gy = __fadd_rn(__fmul_rn(e3 - e1, t), e1);
    const TapSet ts = taps_of(gx, gy, h, w, 1);
    float sv[kWinCPL];
    const float d = gather_dot(fb, ts, D, sc, sy, sx, qv, lane, sv);
    if (lane == 0) { z[p] = d; px[p] = gx; py[p] = gy; }
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int p = threadIdx.x; p < m; p += kWinThreads) mx = fmaxf(mx, z[p]);
  mx = warp_max(mx);
  if (lane == 0) red[0][warp] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red[0][0], red[0][1]), fmaxf(red[0][2], red[0][3]));
  float se = 0.f;
  for (int p = threadIdx.x; p < m; p += kWinThreads) se += expf(z[p] - mx);
  se = warp_sum(se);
  if (lane == 0) red[1][warp] = se;
  __syncthreads();
  const float inv = 1.f / (red[1][0] + red[1][1] + red[1][2] + red[1][3]);
  // probabilities as they would be stored; their maximum decides the nearest-neighbour position(s)
  float pm = 0.f, sxv = 0.f, syv = 0.f, sxx = 0.f, syy = 0.f;
  for (int p = threadIdx.x; p < m; p += kWinThreads) {
    const float pr = expf(z[p] - mx) * inv;
    z[p] = pr;
    pm = fmaxf(pm, pr);
    sxv = fmaf(pr, px[p], sxv); syv = fmaf(pr, py[p], syv);
    sxx = __fadd_rn(sxx, __fmul_rn(__fmul_rn(px[p], px[p]), pr));
    syy = __fadd_rn(syy, __fmul_rn(__fmul_rn(py[p], py[p]), pr));
  }
  pm = warp_max(pm); sxv = warp_sum(sxv); syv = warp_sum(syv); sxx = warp_sum(sxx); syy = warp_sum(syy);
  if (lane == 0) { red[2][warp] = pm; red[3][warp] = sxv; red[4][warp] = syv; red[5][warp] = sxx; red[6][warp] = syy; }
  __syncthreads();
  pm = fmaxf(fmaxf(red[2][0], red[2][1]), fmaxf(red[2][2], red[2][3]));
  float nx = 0.f, ny = 0.f;
  for (int p = threadIdx.x; p < m; p += kWinThreads)
    if (z[p] == pm) { nx += px[p]; ny += py[p]; }
  nx = warp_sum(nx); ny = warp_sum(ny);
  __syncthreads();
  if (lane == 0) { red[0][warp] = nx; red[1][warp] = ny; }
  __syncthreads();
  if (threadIdx.x == 0) {
    exp_soft[qi * 2 + 0] = red[3][0] + red[3][1] + red[3][2] + red[3][3];
    exp_soft[qi * 2 + 1] = red[4][0] + red[4][1] + red[4][2] + red[4][3];
    nn_xy[qi * 2 + 0] = red[0][0] + red[0][1] + red[0][2] + red[0][3];
    nn_xy[qi * 2 + 1] = red[1][0] + red[1][1] + red[1][2] + red[1][3];
    m2_out[qi * 2 + 0] = red[5][0] + red[5][1] + red[5][2] + red[5][3];
    m2_out[qi * 2 + 1] = red[6][0] + red[6][1] + red[6][2] + red[6][3];
  }
  if (prob)
    for (int p = threadIdx.x; p < m; p += kWinThreads) prob[qi * m + p] = z[p];
}

// backward of the window variant (mode 0): g_q and g_fmap from (g_exp, g_std)
__global__ void __launch_bounds__(kWinThreads)
window_expect_bwd_kernel(const float* __restrict__ fmap, int D, int h, int w, int64_t sb, int64_t sc, int64_t sy,
                         int64_t sx, const float* __restrict__ q, const float* __restrict__ centre, int n,
                         const float* __restrict__ offsets, int m, const float* __restrict__ exp_xy,
                         const float* __restrict__ prob, const float* __restrict__ g_exp,
                         const float* __restrict__ g_std, float* __restrict__ g_q, float* __restrict__ g_fmap) {
  __shared__ float dz[kWinMaxPts];
  __shared__ float red[kWinThreads / 32];
  __shared__ float gq_part[kWinThreads / 32][128];
  const int b = blockIdx.y, i = blockIdx.x;
  const size_t qi = (size_t)b * n + i;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float ex = exp_xy[qi * 2], ey = exp_xy[qi * 2 + 1];
  const float gex = g_exp[qi * 2], gey = g_exp[qi * 2 + 1], gs = g_std[qi];
  // variance of the window distribution (for the std gradient)
  float sxx = 0.f, syy = 0.f;
  for (int p = threadIdx.x; p < m; p += kWinThreads) {
    const float2 pos = win_pos(0, centre, offsets, qi, p, m);
    const float pr = prob[qi * m + p];
    sxx = fmaf(pr, pos.x * pos.x, sxx); syy = fmaf(pr, pos.y * pos.y, syy);
  }
  sxx = warp_sum(sxx); syy = warp_sum(syy);
  if (lane == 0) { red[warp] = sxx; }
  __syncthreads();
  const float vx = red[0] + red[1] + red[2] + red[3] - ex * ex;
  __syncthreads();
  if (lane == 0) { red[warp] = syy; }
  __syncthreads();
  const float vy = red[0] + red[1] + red[2] + red[3] - ey * ey;
  __syncthreads();
  const float hx = vx > 1e-10f ? gs * 0.5f / sqrtf(vx) : 0.f;   // d std / d var_x (clamp kills the gradient)
  const float hy = vy > 1e-10f ? gs * 0.5f / sqrtf(vy) : 0.f;
  // dL/dP_p and its probability-weighted mean
  float wsum = 0.f;
  for (int p = threadIdx.x; p < m; p += kWinThreads) {
    const float2 pos = win_pos(0, centre, offsets, qi, p, m);
    const float dP = gex * pos.x + gey * pos.y + hx * (pos.x * pos.x - 2.f * ex * pos.x) +
                     hy * (pos.y * pos.y - 2.f * ey * pos.y);
    dz[p] = dP;
    wsum = fmaf(prob[qi * m + p], dP, wsum);
  }
  wsum = warp_sum(wsum);
  if (lane == 0) red[warp] = wsum;
  __syncthreads();
  const float mean = red[0] + red[1] + red[2] + red[3];
  for (int p = threadIdx.x; p < m; p += kWinThreads) dz[p] = prob[qi * m + p] * (dz[p] - mean);
  __syncthreads();

  const float* fb = fmap + b * sb;
  float* gfb = g_fmap + b * sb;
  float qv[kWinCPL], gq[kWinCPL];
#pragma unroll
  for (int j = 0; j < kWinCPL; ++j) {
    const int c = (sc == 1) ? lane * kWinCPL + j : lane + 32 * j;
    qv[j] = c < D ? __ldg(q + qi * D + c) : 0.f;
    gq[j] = 0.f;
  }
  __shared__ float gpix[kBoxMax];
  __shared__ int ired[4][kWinThreads / 32];
  BoxInfo box;
  bool boxed = false;
  if (sc == 1 && (D & 3) == 0 && D <= 128 && (sx & 3) == 0 && (sy & 3) == 0 && (sb & 3) == 0 &&
      (reinterpret_cast<uintptr_t>(fmap) & 15) == 0 && (reinterpret_cast<uintptr_t>(g_fmap) & 15) == 0)
    boxed = window_box(centre, offsets, qi, m, h, w, ired, box);          // block uniform
  if (boxed) {
    // one scalar per pixel of the box: gpix = sum over taps of dz * bilinear weight
    const int npix = box.bw * box.bh;
    for (int e = threadIdx.x; e < npix; e += kWinThreads) gpix[e] = 0.f;
    __syncthreads();
    for (int p = threadIdx.x; p < m; p += kWinThreads) {
      const float d = dz[p];
      if (d == 0.f) continue;
      const float2 pos = win_pos(0, centre, offsets, qi, p, m);
      const TapSet t = taps_of(pos.x, pos.y, h, w, 0);
      float* g0 = gpix + (t.y0 - box.ymin) * box.bw + (t.x0 - box.xmin);
      if (t.in00) atomicAdd(g0, d * t.w00);
      if (t.in01) atomicAdd(g0 + 1, d * t.w01);
      if (t.in10) atomicAdd(g0 + box.bw, d * t.w10);
      if (t.in11) atomicAdd(g0 + box.bw + 1, d * t.w11);
    }
    __syncthreads();
    const float4 q4 = make_float4(qv[0], qv[1], qv[2], qv[3]);
    const bool act = lane * 4 < D;
    for (int e = warp; e < npix; e += kWinThreads / 32) {
      const float gcoef = gpix[e];
      if (gcoef == 0.f) continue;                                          // warp uniform (also every pixel outside the map)
      const int by = e / box.bw, bx = e - by * box.bw;
      const int64_t off = (int64_t)(box.ymin + by) * sy + (int64_t)(box.xmin + bx) * sx;
      if (act) {
        const float4 f = __ldg(reinterpret_cast<const float4*>(fb + off) + lane);
        gq[0] = fmaf(gcoef, f.x, gq[0]); gq[1] = fmaf(gcoef, f.y, gq[1]);
        gq[2] = fmaf(gcoef, f.z, gq[2]); gq[3] = fmaf(gcoef, f.w, gq[3]);
        atomicAdd(reinterpret_cast<float4*>(gfb + off) + lane,
                  make_float4(gcoef * q4.x, gcoef * q4.y, gcoef * q4.z, gcoef * q4.w));
      }
    }
  } else
  for (int p = warp; p < m; p += kWinThreads / 32) {
    const float d = dz[p];
    if (d == 0.f) continue;
    const float2 pos = win_pos(0, centre, offsets, qi, p, m);
    const TapSet t = taps_of(pos.x, pos.y, h, w, 0);
    float* base = gfb + (int64_t)t.y0 * sy + (int64_t)t.x0 * sx;
    const float* rbase = fb + (int64_t)t.y0 * sy + (int64_t)t.x0 * sx;
#pragma unroll
    for (int j = 0; j < kWinCPL; ++j) {
      const int c = (sc == 1) ? lane * kWinCPL + j : lane + 32 * j;
      if (c >= D) continue;
      const int64_t off = (int64_t)c * sc;
      float a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f;
      const float gd = d * qv[j];
      if (t.in00) { a00 = __ldg(rbase + off); atomicAdd(base + off, gd * t.w00); }
      if (t.in01) { a01 = __ldg(rbase + off + sx); atomicAdd(base + off + sx, gd * t.w01); }
      if (t.in10) { a10 = __ldg(rbase + off + sy); atomicAdd(base + off + sy, gd * t.w10); }
      if (t.in11) { a11 = __ldg(rbase + off + sy + sx); atomicAdd(base + off + sy + sx, gd * t.w11); }
      gq[j] = fmaf(d, a00 * t.w00 + a01 * t.w01 + a10 * t.w10 + a11 * t.w11, gq[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < kWinCPL; ++j) {
    const int c = (sc == 1) ? lane * kWinCPL + j : lane + 32 * j;
    gq_part[warp][c] = gq[j];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += kWinThreads)
    g_q[qi * D + c] = gq_part[0][c] + gq_part[1][c] + gq_part[2][c] + gq_part[3][c];
}

}  // namespace posfeat

using namespace posfeat;

static int check_dense(const void* q, const void* k, const void* v, int B, int n, int m, int D, int C) {
  PF_CHECK_ARG(q && k && v, "NULL pointer");
  PF_CHECK_ARG(B >= 1 && B <= 65535 && n >= 1 && m >= 1, "bad shape B=%d n=%d m=%d", B, n, m);
  PF_CHECK_ARG(D >= 1 && D <= kCDmax, "descriptor length %d outside [1, %d]", D, kCDmax);
  PF_CHECK_ARG(C >= 1 && C <= kCmax, "value width %d outside [1, %d]", C, kCmax);
  return 0;
}

extern "C" size_t posfeat_corr_expect_workspace_bytes(int B, int n, int m, int D, int C) {
  if (B < 1 || n < 1 || m < 1) return 0;
  return corr_tc_workspace_bytes(B, n, m, D, C);
}

extern "C" int posfeat_corr_expect_fwd_f32(const float* q, const float* k, const float* v, int v_batched, int B, int n,
                                           int m, int D, int C, float scale, float* out, float* lse, void* workspace,
                                           size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int e = check_dense(q, k, v, B, n, m, D, C)) return e;
  PF_CHECK_ARG(out && lse, "NULL output pointer");
  if (corr_tc_eligible(B, n, m, D, C) && !getenv("POSFEAT_CORR_SIMT"))
    return corr_tc_fwd(q, k, v, v_batched, B, n, m, D, C, scale, out, lse, workspace, ws_bytes, stream);
  dim3 grid((n + kCT - 1) / kCT, B);
  ProfScope prof(PROF_CORR_FWD, stream);
  corr_expect_fwd_kernel<<<grid, 256, 0, stream>>>(q, k, v, v_batched, n, m, D, C, scale, out, lse);
  PF_LAUNCH_CHECK("corr_expect_fwd_kernel");
  return POSFEAT_OK;
}

extern "C" size_t posfeat_corr_expect_bwd_workspace_bytes(int B, int n, int m, int D, int C) {
  if (B < 1 || n < 1 || m < 1) return 0;
  return corr_tc_bwd_workspace_bytes(B, n, m, D, C);
}

extern "C" int posfeat_corr_expect_bwd_f32(const float* q, const float* k, const float* v, int v_batched, int B, int n,
                                           int m, int D, int C, float scale, const float* out, const float* lse,
                                           const float* g_out, float* g_q, float* g_k, void* workspace, size_t ws_bytes,
                                           void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int e = check_dense(q, k, v, B, n, m, D, C)) return e;
  PF_CHECK_ARG(out && lse && g_out && (g_q || g_k), "NULL pointer");
  if (corr_tc_bwd_eligible(B, n, m, D, C) && !getenv("POSFEAT_CORR_SIMT"))
    return corr_tc_bwd(q, k, v, v_batched, B, n, m, D, C, scale, out, lse, g_out, g_q, g_k, workspace, ws_bytes, stream);
  const size_t smem = sizeof(float) * (kCDmax * (kCT + 4) + kCT * (kCDmax + 4) + kCT * (kCT + 1) + 32 * (kCT + 4));
  PF_CUDA(cudaFuncSetAttribute(corr_expect_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PF_CUDA(cudaFuncSetAttribute(corr_expect_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope prof(PROF_CORR_BWD, stream);
  // rows of X are owned by a CTA; when there are too few X tiles to fill the chip the Y sweep is split as well
  auto launch = [&](bool x_is_q, float* gX) -> int {
    const int nX = x_is_q ? n : m, nY = x_is_q ? m : n;
    const int xt = (nX + kCT - 1) / kCT, yt = (nY + kCT - 1) / kCT;
    int splits = std::max(1, std::min(yt, (4 * 148 + xt * B - 1) / (xt * B)));
    const int tps = (yt + splits - 1) / splits;
    splits = (yt + tps - 1) / tps;
    if (splits > 1) PF_CUDA(cudaMemsetAsync(gX, 0, sizeof(float) * (size_t)B * nX * D, stream));
    dim3 grid(xt, B, splits);
    if (x_is_q) corr_expect_bwd_kernel<true><<<grid, 256, smem, stream>>>(q, k, v, v_batched, n, m, D, C, scale, out, lse, g_out, gX, tps);
    else corr_expect_bwd_kernel<false><<<grid, 256, smem, stream>>>(q, k, v, v_batched, n, m, D, C, scale, out, lse, g_out, gX, tps);
    PF_LAUNCH_CHECK("corr_expect_bwd_kernel");
    return POSFEAT_OK;
  };
  if (g_q) if (int e = launch(true, g_q)) return e;
  if (g_k) if (int e = launch(false, g_k)) return e;
  return POSFEAT_OK;
}

static int check_window(const void* fmap, int B, int D, int h, int w, const void* q, const void* centre, int n, int m) {
  PF_CHECK_ARG(fmap && q && centre, "NULL pointer");
  PF_CHECK_ARG(B >= 1 && B <= 65535 && h >= 1 && w >= 1 && n >= 1, "bad shape B=%d h=%d w=%d n=%d", B, h, w, n);
  PF_CHECK_ARG(D >= 1 && D <= 32 * kWinCPL, "descriptor length %d outside [1, %d]", D, 32 * kWinCPL);
  PF_CHECK_ARG(m >= 1 && m <= kWinMaxPts, "number of sampling positions %d outside [1, %d]", m, kWinMaxPts);
  return 0;
}

extern "C" int posfeat_window_expect_fwd_f32(const float* fmap, int B, int D, int h, int w, int64_t sb, int64_t sc,
                                             int64_t sy, int64_t sx, const float* q, const float* centre, int n,
                                             const float* offsets, int m, int mode, float* exp_xy, float* std_out,
                                             float* prob, float* lse, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int e = check_window(fmap, B, D, h, w, q, centre, n, m)) return e;
  PF_CHECK_ARG(mode == 0 || mode == 1, "mode must be 0 (window) or 1 (line)");
  PF_CHECK_ARG(mode == 1 || offsets, "window mode needs an offset table");
  PF_CHECK_ARG(mode == 0 || m >= 2, "line mode needs at least 2 samples");
  PF_CHECK_ARG(exp_xy && std_out && lse, "NULL output pointer");
  PF_CHECK_ARG(sc == 1 ? (D % kWinCPL == 0 || D <= 32 * kWinCPL) : true, "bad D");
  dim3 grid(n, B);
  ProfScope prof(PROF_WIN_FWD, stream);
  window_expect_fwd_kernel<<<grid, kWinThreads, 0, stream>>>(fmap, D, h, w, sb, sc, sy, sx, q, centre, n, offsets, m, mode,
                                                              exp_xy, std_out, prob, lse);
  PF_LAUNCH_CHECK("window_expect_fwd_kernel");
  return POSFEAT_OK;
}

// ---- DiskLoss dense affinity (losses/kploss.py:158-182): row-side sums of the dual-softmax match distribution
extern "C" size_t posfeat_dual_softmax_reward_workspace_bytes(int B, int n, int m, int D) {
  if (B < 1 || n < 1 || m < 1 || D < 1 || D > 128) return 0;
  return corr_disk_workspace_bytes(B, n, m, D);
}

extern "C" int posfeat_dual_softmax_reward_f32(const float* q, const float* k, const float* rowtab, const float* coltab,
                                               int B, int n, int m, int D, float temperature, float thr_own,
                                               float thr_other, float good_reward, float bad_reward, int dynamic_reward,
                                               float* rows_out, void* workspace, size_t ws_bytes, void* stream) {
  PF_CHECK_ARG(q && k && rowtab && coltab && rows_out && workspace, "NULL pointer");
  PF_CHECK_ARG(B >= 1 && B <= 65535 && n >= 1 && m >= 1 && D >= 1 && D <= 128, "bad shape B=%d n=%d m=%d D=%d", B, n, m, D);
  PF_CHECK_ARG(thr_own > 0.f && thr_other > 0.f, "reward thresholds must be positive");
  return corr_disk_rows(q, k, rowtab, coltab, B, n, m, D, temperature, thr_own, thr_other, good_reward, bad_reward,
                        dynamic_reward, rows_out, workspace, ws_bytes, (cudaStream_t)stream);
}

extern "C" int posfeat_line_search_f32(const float* fmap, int B, int D, int h, int w, int64_t sb, int64_t sc, int64_t sy,
                                       int64_t sx, const float* q, const float* coord_px, const float* Fmat, int n,
                                       int img_h, int img_w, int line_step, float* ends, unsigned char* valid,
                                       float* exp_soft, float* nn_xy, float* m2, float* prob, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int e = check_window(fmap, B, D, h, w, q, coord_px, n, line_step)) return e;
  PF_CHECK_ARG(Fmat && ends && valid && exp_soft && nn_xy && m2, "NULL pointer");
  PF_CHECK_ARG(line_step >= 2 && img_h >= 2 && img_w >= 2, "line search needs at least 2 samples and a 2x2 image");
  PF_CHECK_ARG(sc == 1 ? (D % kWinCPL == 0 || D <= 32 * kWinCPL) : true, "bad D");
  dim3 grid(n, B);
  ProfScope prof(PROF_WIN_FWD, stream);
  line_search_kernel<<<grid, kWinThreads, 0, stream>>>(fmap, D, h, w, sb, sc, sy, sx, q, coord_px, Fmat, n, img_h, img_w,
                                                       line_step, ends, valid, exp_soft, nn_xy, m2, prob);
  PF_LAUNCH_CHECK("line_search_kernel");
  return POSFEAT_OK;
}

extern "C" int posfeat_window_expect_bwd_f32(const float* fmap, int B, int D, int h, int w, int64_t sb, int64_t sc,
                                             int64_t sy, int64_t sx, const float* q, const float* centre, int n,
                                             const float* offsets, int m, const float* exp_xy, const float* prob,
                                             const float* g_exp, const float* g_std, float* g_q, float* g_fmap,
                                             void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int e = check_window(fmap, B, D, h, w, q, centre, n, m)) return e;
  PF_CHECK_ARG(offsets && exp_xy && prob && g_exp && g_std && g_q && g_fmap, "NULL pointer");
  dim3 grid(n, B);
  ProfScope prof(PROF_WIN_BWD, stream);
  window_expect_bwd_kernel<<<grid, kWinThreads, 0, stream>>>(fmap, D, h, w, sb, sc, sy, sx, q, centre, n, offsets, m, exp_xy,
                                                              prob, g_exp, g_std, g_q, g_fmap);
  PF_LAUNCH_CHECK("window_expect_bwd_kernel");
  return POSFEAT_OK;
}
