// Micro-benchmark: what read bandwidth does the NMS access pattern reach with no compute?
//  mode 0: linear grid-stride float4 reads (the ceiling)
//  mode 1: strip pattern -- a warp reads 512-byte row segments of a [B][H][W] float map, ROWS rows
//          per task, BATCH rows in flight, tasks laid out like nms_quad_r1_kernel
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench_stream tools/ubench_stream.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void linear_read(const float4* __restrict__ p, size_t n, float* out) {
  float acc = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = __ldg(p + i);
    acc += v.x + v.y + v.z + v.w;
  }
  if (acc == 12345.678f) *out = acc;
}

template <int BATCH>
__global__ void strip_read(const float* __restrict__ p, int H, int W, int rows, int ncw, int ntasks, int img_fast,
                           float* out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int b = img_fast ? blockIdx.x : blockIdx.y;
  const int task = (img_fast ? blockIdx.y : blockIdx.x) * wpb + warp;
  if (task >= ntasks) return;
  const int cw = task % ncw, strip = task / ncw;
  const int x0 = cw * 124 + 4 * lane;
  const float* base = p + (size_t)b * H * W + (x0 + 3 < W ? x0 : 0);
  const int y0 = strip * rows;
  float acc = 0.f;
  for (int y = y0; y < y0 + rows + 2; y += BATCH) {
    float4 v[BATCH];
#pragma unroll
    for (int k = 0; k < BATCH; ++k) {
      int yy = min(y + k, H - 1);
      v[k] = __ldg(reinterpret_cast<const float4*>(base + (size_t)yy * W));
    }
#pragma unroll
    for (int k = 0; k < BATCH; ++k) acc += v[k].x + v[k].y + v[k].z + v[k].w;
  }
  if (acc == 12345.678f) *out = acc;
}

int main() {
  const int B = 128, H = 896, W = 1200;
  const size_t n = (size_t)B * H * W;
  float *d, *o;
  cudaMalloc(&d, n * 4);
  cudaMalloc(&o, 4);
  cudaMemset(d, 0, n * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  auto time = [&](auto fn, const char* name) {
    for (int i = 0; i < 3; ++i) fn();
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) fn();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%-48s %7.1f us  %6.0f GB/s  (%s)\n", name, ms * 100, n * 4 / (ms / 10 * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
  };
  time([&] { linear_read<<<148 * 8, 256>>>((const float4*)d, n / 4, o); }, "linear float4, 148x8 CTAs x 256");
  time([&] { linear_read<<<148 * 16, 512>>>((const float4*)d, n / 4, o); }, "linear float4, 148x16 CTAs x 512");
  const int ncw = (W - 2 + 123) / 124;
  char name[128];
  for (int rows : {32, 62, 126}) {
    const int nstrips = (H - 2 + rows - 1) / rows, ntasks = ncw * nstrips;
    for (int wpb : {4, 8}) {
      for (int img_fast : {0, 1}) {
        dim3 g = img_fast ? dim3(B, (ntasks + wpb - 1) / wpb) : dim3((ntasks + wpb - 1) / wpb, B);
        snprintf(name, sizeof name, "strip rows=%d batch=8 warps/cta=%d img_fast=%d", rows, wpb, img_fast);
        time([&] { strip_read<8><<<g, wpb * 32>>>(d, H, W, rows, ncw, ntasks, img_fast, o); }, name);
        snprintf(name, sizeof name, "strip rows=%d batch=16 warps/cta=%d img_fast=%d", rows, wpb, img_fast);
        time([&] { strip_read<16><<<g, wpb * 32>>>(d, H, W, rows, ncw, ntasks, img_fast, o); }, name);
      }
    }
  }
  return 0;
}
