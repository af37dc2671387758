// Exact-fp32 SIMT GEMM used for small / latency-bound products and as the on-device
// cross-check of the tcgen05 path.  C[M,N] = alpha * op(A)[M,K] * op(B)[K,N] + beta * C.
// All matrices row-major with explicit leading dimensions.
//   transA == 0: A is [M,K] (lda >= K);  transA == 1: A is stored [K,M] (lda >= M)
//   transB == 0: B is [K,N] (ldb >= N);  transB == 1: B is stored [N,K] (ldb >= K)
#pragma once
#include "common.cuh"

int mb200_sgemm_launch(int transA, int transB, int M, int N, int K, float alpha, const float* A, int lda,
                       const float* B, int ldb, float beta, float* C, int ldc, cudaStream_t stream);
