// EXPERIMENTAL (not yet run on a B200; behind MOTIFS_GEMM_MN=1 on the Python side): the bf16x3 tcgen05 GEMM of
// gemm_tc.cu for operands whose REDUCTION dimension is the row index — C[M,N] = A^T B with A stored [K, M] and B stored
// [K, N] (row-major, bf16 (hi, lo) pairs). This is the shape of every weight-gradient product on the path,
//   dW[out, in] = sum_rows dY[row, out] * X[row, in]        (fc6/fc7, post_lstm, rel_compress, LSTM projections,
//                                                             lib/rel_model.py:360-390, 503-524 under autograd)
// where today both operands are first transposed into K-major copies (split_transpose_kernel: 1.04 ms per step in
// profiles/r01_ncu_launches_profile_step_v5.csv). Here the activations' natural [rows, channels] layout is consumed
// directly as "MN-major" UMMA operands:
//   * TMA box {64 channels (128 B), 64 rows} with SWIZZLE_128B lands in shared memory as 8-row x 128-byte atoms —
//     the canonical MN-major SW128 layout  Swizzle<3,4,3> o ((8,n),(8,k)) : ((1,LBO),(8,SBO))  in 16-byte units
//     (cute/atom/mma_traits_sm100.hpp), with SBO = 1024 B (next 8 K-rows) and LBO = 8192 B (next 64-channel block);
//   * a 128-wide A tile is two such boxes, a 128/256-wide B tile two/four; one UMMA (K = 16) advances the start
//     address by 16 rows x 128 B = 2048 B; instruction-descriptor bits 15/16 (a_major / b_major) = 1 (MN).
// Everything after the MMA (TMEM accumulators, epilogue, split-K) is the scheme of gemm_tc.cu; that file is untouched.
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int BM = 128, BK = 64;
constexpr int ACC_STAGES = 2;
constexpr int TILE_A = BM * BK * 2;            // 16 KB = two 64-channel boxes of 8 KB
constexpr int BOX_BYTES = 64 * BK * 2;         // one {64 ch, 64 rows} box
constexpr int kThreads = 192;

template <int BN_> struct Cfg {
  static constexpr int BN = BN_;
  static constexpr int STAGES = BN_ == 128 ? 3 : 2;
  static constexpr int TILE_B = BN_ * BK * 2;
  static constexpr int STAGE_BYTES = 2 * TILE_A + 2 * TILE_B;
  static constexpr int TMEM_COLS = ACC_STAGES * BN_;
  static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 + 256;
};

struct Params {
  int M, N;
  int kblocks, splits, m_tiles, n_tiles;
  float* C; long long ldc;
  float* partial;
};

// MN-major, 128-byte swizzle: LBO [16,30) = 8192 >> 4, SBO [32,46) = 1024 >> 4, version 1 @46, layout SW128 (2) @61.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(8192 >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__host__ __device__ constexpr uint32_t idesc_bf16_f32_mn(int M, int N) {
  return tc::umma_idesc_bf16_f32(M, N) | (1u << 15) | (1u << 16);
}

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16x3_mn_kernel(const __grid_constant__ CUtensorMap tmAhi, const __grid_constant__ CUtensorMap tmAlo,
                      const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo,
                      const Params p) {
  constexpr int STAGES = Cfg<BN>::STAGES, TILE_B = Cfg<BN>::TILE_B, STAGE_BYTES = Cfg<BN>::STAGE_BYTES;
  constexpr int TMEM_COLS = Cfg<BN>::TMEM_COLS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + (size_t)STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + ACC_STAGES;
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + ACC_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_split = p.m_tiles * p.n_tiles;
  const int total_tiles = p.splits * per_split;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmAhi); tc::prefetch_tmap(&tmAlo); tc::prefetch_tmap(&tmBhi); tc::prefetch_tmap(&tmBlo);
    for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < ACC_STAGES; ++s) { tc::mbar_init(&tfull_bar[s], 1); tc::mbar_init(&tempty_bar[s], 4); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, TMEM_COLS);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------------------------------------------------------- TMA producer: 64-channel x 64-row boxes
      int stage = 0; uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int split = t / per_split, r = t - split * per_split;
        const int mi = r / p.n_tiles, ni = r - mi * p.n_tiles;
        const int m0 = mi * BM, n0 = ni * BN, kb0 = split * p.kblocks;
        for (int kb = 0; kb < p.kblocks; ++kb) {
          tc::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + (size_t)stage * STAGE_BYTES;
          tc::mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          const int k0 = (kb0 + kb) * BK;
#pragma unroll
          for (int j = 0; j < BM / 64; ++j) {
            tc::tma_load_2d(st + j * BOX_BYTES, &tmAhi, &full_bar[stage], m0 + 64 * j, k0);
            tc::tma_load_2d(st + TILE_A + j * BOX_BYTES, &tmAlo, &full_bar[stage], m0 + 64 * j, k0);
          }
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) {
            tc::tma_load_2d(st + 2 * TILE_A + j * BOX_BYTES, &tmBhi, &full_bar[stage], n0 + 64 * j, k0);
            tc::tma_load_2d(st + 2 * TILE_A + TILE_B + j * BOX_BYTES, &tmBlo, &full_bar[stage], n0 + 64 * j, k0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------------------------------------------------------- MMA issuer
      constexpr uint32_t idesc = idesc_bf16_f32_mn(BM, BN);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        tc::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc::tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < p.kblocks; ++kb) {
          tc::mbar_wait(&full_bar[stage], phase);
          tc::tc_fence_after();
          const uint32_t sa = tc::smem_u32(smem + (size_t)stage * STAGE_BYTES);
          const uint64_t a_hi = umma_desc_mn_sw128(sa), a_lo = umma_desc_mn_sw128(sa + TILE_A);
          const uint64_t b_hi = umma_desc_mn_sw128(sa + 2 * TILE_A), b_lo = umma_desc_mn_sw128(sa + 2 * TILE_A + TILE_B);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adv = (uint64_t)((k * 16 * 128) >> 4);     // 16 K-rows of 128 bytes per UMMA
            tc::umma_bf16(tmem_d, a_hi + adv, b_hi + adv, idesc, (kb | k) != 0);
            tc::umma_bf16(tmem_d, a_hi + adv, b_lo + adv, idesc, 1);
            tc::umma_bf16(tmem_d, a_lo + adv, b_hi + adv, idesc, 1);
          }
          tc::umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc::umma_commit(&tfull_bar[acc]);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;
    const int r = q * 32 + lane;
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int split = t / per_split, rr = t - split * per_split;
      const int mi = rr / p.n_tiles, ni = rr - mi * p.n_tiles;
      const int m0 = mi * BM, n0 = ni * BN;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      tc::mbar_wait(&tfull_bar[acc], acc_phase);
      tc::tc_fence_after();
      const long long row = (m0 + r < p.M) ? (long long)(m0 + r) : -1;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        __syncwarp();
        uint32_t v[32];
        tc::tmem_ld_32x32(taddr + (uint32_t)(c * 32), v);
        tc::tmem_ld_wait();
        const int nb = n0 + c * 32;
        if (row < 0 || nb >= p.N) continue;
        const int ncols = min(32, p.N - nb);
        float* dst = p.splits > 1 ? p.partial + ((size_t)split * p.M + row) * p.N + nb : p.C + row * p.ldc + nb;
        if (ncols == 32 && ((((uintptr_t)dst) & 15) == 0)) {
#pragma unroll
          for (int j = 0; j < 8; ++j) ((uint4*)dst)[j] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) if (j < ncols) dst[j] = __uint_as_float(v[j]);
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tempty_bar[acc]);
    }
  }
  __syncthreads();
  if (warp == 1) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, TMEM_COLS); }
}

__global__ void splitk_sum_kernel(const float* __restrict__ partial, int splits, long long MN, int N, float* __restrict__ C,
                                  long long ldc) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < MN; i += (long long)blockDim.x * gridDim.x) {
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += partial[(size_t)z * MN + i];
    const long long m = i / N;
    C[m * ldc + (i - m * N)] = s;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_mn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &st) == cudaSuccess &&
        st == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// [rows = K, cols] bf16 matrix with `cols` contiguous (pitch ld elements) -> box {64 cols, 64 rows}, 128B swizzle, zero fill.
bool make_tmap_mn(CUtensorMap* m, const void* ptr, long long rows, long long cols, long long ld) {
  EncodeTiledFn enc = get_encode_mn();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)BK};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int splits_for(int M, int N, int kblocks) {
  const long long tiles = (long long)mb200_div_up(M, BM) * mb200_div_up(N, 128);
  if (tiles >= kNumSMs / 2 || kblocks < 8) return 1;
  int s = (int)min((long long)kblocks / 4, (long long)(kNumSMs / tiles));
  return s < 2 ? 1 : s;
}

}  // namespace

extern "C" {

// Floats of split-K workspace mb200_gemm_bf16x3_mn may need (0 when it will not split).
long long mb200_gemm_mn_workspace_floats(int M, int N, int K) {
  const int s = splits_for(M, N, mb200_div_up(K, BK));
  return s > 1 ? (long long)s * M * N : 0;
}

// C[M,N] (fp32, row pitch ldc) = A^T B with A stored [K, M] (pitch lda), B stored [K, N] (pitch ldb), both as (hi, lo)
// bf16 pairs, M / N contiguous. lda, ldb multiples of 8 (16-byte TMA pitch); K arbitrary (rows beyond K read as zero).
int mb200_gemm_bf16x3_mn(const void* Ahi, const void* Alo, long long lda, const void* Bhi, const void* Blo, long long ldb,
                         int M, int N, int K, float* C, long long ldc, float* workspace, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return MB200_OK;
  if (K <= 0 || lda % 8 || ldb % 8 || lda < M || ldb < N) return MB200_ERR_ARG;
  static bool attr_done = false;
  if (!attr_done) {
    MB200_CHECK(cudaFuncSetAttribute(gemm_bf16x3_mn_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<128>::SMEM_BYTES));
    MB200_CHECK(cudaFuncSetAttribute(gemm_bf16x3_mn_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<256>::SMEM_BYTES));
    attr_done = true;
  }
  const int kblocks = mb200_div_up(K, BK);
  Params p = {};
  p.M = M; p.N = N; p.C = C; p.ldc = ldc;
  p.splits = workspace ? splits_for(M, N, kblocks) : 1;
  // every split takes the same number of k-blocks; the last ones may run past K (zero-filled rows)
  p.kblocks = mb200_div_up(kblocks, p.splits);
  p.partial = workspace;
  const int bn = (p.splits > 1 || N < 256) ? 128 : 256;
  p.m_tiles = mb200_div_up(M, BM); p.n_tiles = mb200_div_up(N, bn);
  CUtensorMap ta, tal, tb, tbl;
  if (!make_tmap_mn(&ta, Ahi, K, M, lda) || !make_tmap_mn(&tal, Alo, K, M, lda) ||
      !make_tmap_mn(&tb, Bhi, K, N, ldb) || !make_tmap_mn(&tbl, Blo, K, N, ldb)) {
    mb200_set_error("cuTensorMapEncodeTiled (mn)", cudaErrorInvalidValue);
    return MB200_ERR_CUDA;
  }
  const long long tiles = (long long)p.splits * p.m_tiles * p.n_tiles;
  if (tiles > 0x7fffffffLL) return MB200_ERR_UNSUPPORTED;
  const int grid = (int)min(tiles, (long long)g_mb200_sm_budget);
  if (bn == 256) gemm_bf16x3_mn_kernel<256><<<grid, kThreads, Cfg<256>::SMEM_BYTES, stream>>>(ta, tal, tb, tbl, p);
  else gemm_bf16x3_mn_kernel<128><<<grid, kThreads, Cfg<128>::SMEM_BYTES, stream>>>(ta, tal, tb, tbl, p);
  MB200_CHECK_LAUNCH("gemm_bf16x3_mn_kernel");
  if (p.splits > 1) {
    const long long MN = (long long)M * N;
    splitk_sum_kernel<<<(int)min((long long)kNumSMs * 8, (MN + 255) / 256), 256, 0, stream>>>(workspace, p.splits, MN, N, C, ldc);
    MB200_CHECK_LAUNCH("splitk_sum_kernel");
  }
  return MB200_OK;
}

}  // extern "C"
