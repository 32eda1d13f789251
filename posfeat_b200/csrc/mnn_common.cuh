// Declarations shared by the matcher translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace posfeat {

size_t simt_workspace_bytes(int N, int M);
int mnn_simt(const float* A, int N, int64_t lda, const float* Bm, int M, int64_t ldb, int D, int32_t* nn12,
             int32_t* nn21, void* ws, cudaStream_t stream);
int run_rowbest_simt(const float* X, int NX, int64_t ldx, const float* Y, int NY, int64_t ldy, int D,
                     int32_t* nn, float* top2, void* ws, cudaStream_t stream);
// both directions, optionally with the top-2 similarities per row ([rows][2] float32) for the ratio test
int mnn_simt_top2(const float* A, int N, int64_t lda, const float* Bm, int M, int64_t ldb, int D, int32_t* nn12,
                  int32_t* nn21, float* top12, float* top21, void* ws, cudaStream_t stream);
int launch_ratio_flags(const int32_t* nn12, const int32_t* nn21, const float* top12, const float* top21, int N, int M,
                       float ratio, int mutual, unsigned char* flags, cudaStream_t stream);
int launch_compact_flags(const int32_t* nn12, const unsigned char* flags, int P, int N, int64_t* matches,
                         int32_t* n_matches, cudaStream_t stream);
int launch_mutual_compact(const int32_t* nn12, const int32_t* nn21, int N, int M, int64_t* matches,
                          int32_t* n_matches, cudaStream_t stream);

// tensor-core path (mnn_tc.cu)
bool tc_supported(int N, int M, int D);
size_t tc_workspace_bytes(int P, int N, int M);
// batched over P pairs: pair p uses A + p*strideA, Bm + p*strideB (elements), nn12 + p*N, nn21 + p*M
// nn21 may be NULL (matches only); the call also produces matches / n_matches
int mnn_tc(const float* A, int64_t strideA, int N, int64_t lda, const float* Bm, int64_t strideB, int M, int64_t ldb,
           int D, int P, int32_t* nn12, int32_t* nn21, int64_t* matches, int32_t* n_matches, void* ws, size_t ws_bytes,
           cudaStream_t stream, float* top12 = nullptr, float* top21 = nullptr);
int launch_mutual_compact_batched(const int32_t* nn12, const int32_t* nn21, int P, int N, int M, int64_t* matches,
                                  int32_t* n_matches, cudaStream_t stream);

}  // namespace posfeat
