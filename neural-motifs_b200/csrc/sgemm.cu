// Exact-fp32 SIMT GEMM (see sgemm.cuh). 64x64x16 tiles, 256 threads, 4x4 register tile.
#include "sgemm.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;

template <int TA, int TB>
__global__ void __launch_bounds__(256)
sgemm_kernel(int M, int N, int K, float alpha, const float* __restrict__ A, int lda,
             const float* __restrict__ B, int ldb, float beta, float* __restrict__ C, int ldc) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tr = (tid / 16) * TM, tc = (tid % 16) * TN;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    // A tile -> As[k][m]
#pragma unroll
    for (int i = 0; i < (BM * BK) / 256; ++i) {
      const int e = tid + i * 256;
      int m, k;
      if (TA == 0) { k = e % BK; m = e / BK; } else { m = e % BM; k = e / BM; }
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < K) v = (TA == 0) ? A[(size_t)gm * lda + gk] : A[(size_t)gk * lda + gm];
      As[k][m] = v;
    }
    // B tile -> Bs[k][n]
#pragma unroll
    for (int i = 0; i < (BN * BK) / 256; ++i) {
      const int e = tid + i * 256;
      int n, k;
      if (TB == 0) { n = e % BN; k = e / BN; } else { k = e % BK; n = e / BK; }
      const int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < N && gk < K) v = (TB == 0) ? B[(size_t)gk * ldb + gn] : B[(size_t)gn * ldb + gk];
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[k][tr + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[k][tc + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int gm = m0 + tr + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int gn = n0 + tc + j;
      if (gn >= N) continue;
      float* c = C + (size_t)gm * ldc + gn;
      const float v = alpha * acc[i][j];
      *c = (beta == 0.f) ? v : fmaf(beta, *c, v);
    }
  }
}

}  // namespace

int mb200_sgemm_launch(int transA, int transB, int M, int N, int K, float alpha, const float* A, int lda,
                       const float* B, int ldb, float beta, float* C, int ldc, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return MB200_OK;
  dim3 grid(mb200_div_up(N, BN), mb200_div_up(M, BM));
  if (grid.y > 65535) return MB200_ERR_UNSUPPORTED;
  if (!transA && !transB) sgemm_kernel<0, 0><<<grid, 256, 0, stream>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  else if (!transA && transB) sgemm_kernel<0, 1><<<grid, 256, 0, stream>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  else if (transA && !transB) sgemm_kernel<1, 0><<<grid, 256, 0, stream>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  else sgemm_kernel<1, 1><<<grid, 256, 0, stream>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  MB200_CHECK_LAUNCH("mb200_sgemm");
  return MB200_OK;
}

extern "C" int mb200_sgemm(int transA, int transB, int M, int N, int K, float alpha, const float* A, int lda,
                           const float* B, int ldb, float beta, float* C, int ldc, cudaStream_t stream) {
  return mb200_sgemm_launch(transA, transB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, stream);
}
