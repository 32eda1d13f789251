// Micro-benchmark: tcgen05.ld (LDTM) throughput per SM on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench_tmem tools/ubench_tmem.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}

template <int MODE>  // 0: ld+wait each; 1: two loads per wait; 2: ld + max-reduce work
__global__ void bench(int iters, long long* cycles, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t v[32], w[32];
    const uint32_t col = (uint32_t)((i * 64 + (warp >> 2) * 32) & 511);
    ld32(base + col, v);
    if (MODE == 1) ld32(base + ((col + 32) & 511), w);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (MODE == 2) {
      float m = -1e30f;
#pragma unroll
      for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(v[j]));
      acc += m;
    } else {
      acc += __uint_as_float(v[0] ^ v[31]);
      if (MODE == 1) acc += __uint_as_float(w[0] ^ w[31]);
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 12345.f) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512));
}

int main() {
  long long* d_c; float* d_s;
  cudaMalloc(&d_c, 8 * 256); cudaMalloc(&d_s, 4);
  const int iters = 4096;
  for (int mode = 0; mode < 3; ++mode)
    for (int warps : {4, 8, 16, 32})
      for (int grid : {1, 148}) {
        if (mode == 0) bench<0><<<grid, warps * 32>>>(iters, d_c, d_s);
        if (mode == 1) bench<1><<<grid, warps * 32>>>(iters, d_c, d_s);
        if (mode == 2) bench<2><<<grid, warps * 32>>>(iters, d_c, d_s);
        cudaError_t e = cudaDeviceSynchronize();
        long long c[256];
        cudaMemcpy(c, d_c, 8 * grid, cudaMemcpyDeviceToHost);
        double loads = (double)iters * warps * (mode == 1 ? 2 : 1);
        double bytes = loads * 32 * 32 * 4;
        printf("mode %d warps %2d grid %3d: %lld cycles, %.1f B/cycle/SM, %.1f cycles per LDTM.x32 per SM (%s)\n", mode,
               warps, grid, c[0], bytes / c[0], c[0] / loads, cudaGetErrorString(e));
      }
  return 0;
}
