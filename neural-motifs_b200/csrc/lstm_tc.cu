// Highway-LSTM recurrence on the tensor cores (tcgen05, bf16x3) for LARGE batches — BASELINE.json configs[4]
// (H = 512, B = 256, T up to 256) and the north star's "runs the 4x gate GEMM on tensor cores per step".
//
// Per step the reference does  tmp_h[B,5H] = h_{t-1}[B,H] x W_h[H,5H]  with cublasSgemm plus an element-wise kernel and two
// device-wide host syncs (lib/lstm/highway_lstm_cuda/src/highway_lstm_kernel.cu:441-465); csrc/lstm.cu does it as a SIMT
// matvec, which is the right thing for the production shape (B = 6: latency-bound, 4x the reference) but 2560 x 512 MACs for
// each of 256 rows is tensor-core work. Here ONE persistent cooperative kernel runs all T steps of a layer:
//   * CTA (s, m) owns hidden units [16 s, 16 s + 16) for batch rows [128 m, 128 m + 128): its slice of W_h — 5 gates x 16
//     units = 80 columns, as K-major bf16 (hi, lo) pairs, 160 KB — is loaded into shared memory ONCE and stays there;
//   * per step it streams its rows of h_{t-1} (kept as a bf16 pair beside the fp32 state) through a 2-stage TMA ring,
//     issues 8 k-blocks x 4 x 3 tcgen05.mma (128 x 80 x 16, fp32 accumulation in TMEM: h_hi W_hi + h_hi W_lo + h_lo W_hi),
//     and four epilogue warps (thread = batch row) read the 80 gate pre-activations of their row back from TMEM, add the
//     hoisted input projection + bias, apply the gate math in the reference's expression order and write h_t (fp32 and
//     bf16 pair), c_t and — in training — the six gate activations the backward kernel needs;
//   * one grid barrier per step (the h_t pair must be complete before any CTA's TMA of step t+1 reads it).
// 32 slices x ceil(B/128) CTAs (64 at B = 256). Slot / direction / zero-state conventions are csrc/lstm.cu's, so the
// existing backward kernel consumes this kernel's h / c / gates unchanged.
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int kU = 16;                     // hidden units per CTA
constexpr int kN = 5 * kU;                 // 80 accumulator columns: column = gate * 16 + unit
constexpr int kBK = 64, kBM = 128;
constexpr int kThreadsTc = 192;            // warp 0: TMA, warp 1: MMA + TMEM, warps 2..5: epilogue
constexpr int kStages = 2;
constexpr int kTileA = kBM * kBK * 2;      // 16 KB per operand half
constexpr int kTileW = kN * kBK * 2;       // 10 240 B per k-block per operand half (10 swizzle atoms)
constexpr int kTmemCols = 128;

__device__ unsigned int g_lstm_tc_barrier;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ void grid_barrier_all(unsigned int* bar, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
    } while (v < target);
  }
  __syncthreads();
}

__device__ __forceinline__ void tmem_ld_32x32_x4(uint32_t taddr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}

struct TcArgs {
  int H, B, T, dir;
  const float* P;          // [T,B,6H] hoisted input projection (aliases gates in training)
  const float* bias;       // [5H]
  const float* dropout;    // [B,H]
  float* h; float* c;      // [T+1,B,H]
  __nv_bfloat16* hb_hi; __nv_bfloat16* hb_lo;   // [T+1,B,H] bf16 pair of h (slot 0 and untouched rows are zero)
  float* gates;            // [T,B,6H] or nullptr
  const int* lengths;      // DEV [B], descending
  unsigned int* barrier;
};

__global__ void __launch_bounds__(kThreadsTc, 1)
lstm_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmHhi, const __grid_constant__ CUtensorMap tmHlo,
                   const __grid_constant__ CUtensorMap tmWhi, const __grid_constant__ CUtensorMap tmWlo, const TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int H = a.H, B = a.B, T = a.T;
  const int kblocks = H / kBK;
  uint8_t* smem_w = smem;                                         // [kblocks][hi | lo][80 rows x 128 B]
  uint8_t* smem_a = smem + (size_t)kblocks * 2 * kTileW;          // ring: [stage][hi | lo][128 rows x 128 B]
  uint64_t* full_bar = (uint64_t*)(smem_a + (size_t)kStages * 2 * kTileA);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* w_bar = empty_bar + kStages;
  uint64_t* acc_bar = w_bar + 1;
  uint32_t* tmem_slot = (uint32_t*)(acc_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slices = H / kU;
  const int s = blockIdx.x % slices, m = blockIdx.x / slices;
  const int j0 = s * kU, m0 = m * kBM;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmHhi); tc::prefetch_tmap(&tmHlo); tc::prefetch_tmap(&tmWhi); tc::prefetch_tmap(&tmWlo);
    for (int i = 0; i < kStages; ++i) { tc::mbar_init(&full_bar[i], 1); tc::mbar_init(&empty_bar[i], 1); }
    tc::mbar_init(w_bar, 1); tc::mbar_init(acc_bar, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, kTmemCols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {                                   // the weight slice: resident for the whole sequence
    tc::mbar_expect_tx(w_bar, (uint32_t)(kblocks * 2 * kTileW));
    for (int kb = 0; kb < kblocks; ++kb) {
      tc::tma_load_2d(smem_w + (size_t)(2 * kb) * kTileW, &tmWhi, w_bar, kb * kBK, s * kN);
      tc::tma_load_2d(smem_w + (size_t)(2 * kb + 1) * kTileW, &tmWlo, w_bar, kb * kBK, s * kN);
    }
  }

  // epilogue threads: row of the batch tile == TMEM lane
  const int q = warp & 3;
  const int b = m0 + q * 32 + lane;
  const int my_len = (warp >= 2 && b < B) ? a.lengths[b] : 0;
  const size_t BH = (size_t)B * H;
  int stage = 0; uint32_t phase = 0;             // TMA / MMA ring position (each keeps its own copy)
  uint32_t acc_phase = 0;

  for (int step = 0; step < T; ++step) {
    const int t = (a.dir == 0) ? step : T - 1 - step;
    const int prev = (a.dir == 0) ? t : (t + 2) % (T + 1);
    if (warp == 0) {
      if (lane == 0) {
        asm volatile("fence.proxy.async.global;" ::: "memory");   // h_{t-1} was written through the generic proxy
        for (int kb = 0; kb < kblocks; ++kb) {
          tc::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem_a + (size_t)stage * 2 * kTileA;
          tc::mbar_expect_tx(&full_bar[stage], 2 * kTileA);
          tc::tma_load_2d(st, &tmHhi, &full_bar[stage], kb * kBK, prev * B + m0);
          tc::tma_load_2d(st + kTileA, &tmHlo, &full_bar[stage], kb * kBK, prev * B + m0);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
      __syncwarp();
    } else if (warp == 1) {
      if (lane == 0) {
        constexpr uint32_t idesc = tc::umma_idesc_bf16_f32(kBM, kN);
        if (step == 0) tc::mbar_wait(w_bar, 0);
        tc::tc_fence_after();
        for (int kb = 0; kb < kblocks; ++kb) {
          tc::mbar_wait(&full_bar[stage], phase);
          tc::tc_fence_after();
          const uint32_t sa = tc::smem_u32(smem_a + (size_t)stage * 2 * kTileA);
          const uint32_t sw = tc::smem_u32(smem_w + (size_t)(2 * kb) * kTileW);
          const uint64_t a_hi = tc::umma_desc_k_sw128(sa), a_lo = tc::umma_desc_k_sw128(sa + kTileA);
          const uint64_t w_hi = tc::umma_desc_k_sw128(sw), w_lo = tc::umma_desc_k_sw128(sw + kTileW);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);
            tc::umma_bf16(tmem_base, a_hi + adv, w_hi + adv, idesc, (kb | k) != 0);
            tc::umma_bf16(tmem_base, a_hi + adv, w_lo + adv, idesc, 1);
            tc::umma_bf16(tmem_base, a_lo + adv, w_hi + adv, idesc, 1);
          }
          tc::umma_commit(&empty_bar[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        tc::umma_commit(acc_bar);
      }
      __syncwarp();
    } else {
      const bool active = b < B && my_len > t;
      const float* Pp = a.P + ((size_t)t * B + (active ? b : 0)) * 6 * H + j0;
      const float* cprev = a.c + (size_t)prev * BH + (size_t)(active ? b : 0) * H + j0;
      // everything that does not depend on the recurrence is fetched BEFORE the wait for the accumulator: the row's 6 x 16
      // projection values, c_{t-1}, the dropout mask and the bias (first version: 4 serial load round trips per step
      // behind the MMAs, 22 us per step at B = 256)
      float4 p[6][kU / 4], cp[kU / 4], dp[kU / 4];
      if (active) {
#pragma unroll
        for (int v = 0; v < kU / 4; ++v) {
#pragma unroll
          for (int gg = 0; gg < 6; ++gg) p[gg][v] = __ldcg((const float4*)(Pp + (size_t)gg * H + 4 * v));
          cp[v] = (step == 0) ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldcg((const float4*)(cprev + 4 * v));
          dp[v] = __ldg((const float4*)(a.dropout + (size_t)b * H + j0 + 4 * v));
        }
      }
      tc::mbar_wait(acc_bar, acc_phase);
      tc::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll
      for (int v = 0; v < kU / 4; ++v) {
        const int u0 = 4 * v;
        __syncwarp();
        uint32_t g[5][4];
#pragma unroll
        for (int gg = 0; gg < 5; ++gg) tmem_ld_32x32_x4(taddr + (uint32_t)(gg * kU + u0), g[gg]);
        tc::tmem_ld_wait();
        if (!active) continue;
        float hv[4], cv[4], gv[6][4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = j0 + u0 + e;
          float gt[5];
#pragma unroll
          for (int gg = 0; gg < 5; ++gg)
            gt[gg] = (((const float*)&p[gg][v])[e] + __uint_as_float(g[gg][e])) + __ldg(a.bias + gg * H + j);
          // elementWise_fp, highway_lstm_kernel.cu:125-159 (same expression order as csrc/lstm.cu)
          const float in_gate = sigmoidf_(gt[0]);
          const float forget_gate = sigmoidf_(gt[1]);
          const float act_gate = tanhf(gt[2]);
          const float out_gate = sigmoidf_(gt[3]);
          const float r_gate = sigmoidf_(gt[4]);
          const float lin_gate = ((const float*)&p[5][v])[e];
          float val = (forget_gate * ((const float*)&cp[v])[e]) + (in_gate * act_gate);
          cv[e] = val;
          val = out_gate * tanhf(val);
          val = (float)((double)(val * r_gate) + (1.0 - (double)r_gate) * (double)lin_gate);
          hv[e] = val * ((const float*)&dp[v])[e];
          gv[0][e] = in_gate; gv[1][e] = forget_gate; gv[2][e] = act_gate; gv[3][e] = out_gate; gv[4][e] = r_gate;
          gv[5][e] = lin_gate;
        }
        const size_t o = (size_t)(t + 1) * BH + (size_t)b * H + j0 + u0;
        *(float4*)(a.h + o) = make_float4(hv[0], hv[1], hv[2], hv[3]);
        *(float4*)(a.c + o) = make_float4(cv[0], cv[1], cv[2], cv[3]);
        __nv_bfloat16 hh[4], hl[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) tc::split_bf16(hv[e], hh[e], hl[e]);
        *(uint2*)(a.hb_hi + o) = *(const uint2*)hh;
        *(uint2*)(a.hb_lo + o) = *(const uint2*)hl;
        if (a.gates) {
          float* G = a.gates + ((size_t)t * B + b) * 6 * H + j0 + u0;
#pragma unroll
          for (int gg = 0; gg < 6; ++gg)
            *(float4*)(G + (size_t)gg * H) = make_float4(gv[gg][0], gv[gg][1], gv[gg][2], gv[gg][3]);
        }
      }
      acc_phase ^= 1;
      tc::tc_fence_before();
      asm volatile("fence.proxy.async.global;" ::: "memory");     // make the h pair visible to the other CTAs' TMA reads
    }
    grid_barrier_all(a.barrier, (unsigned int)(step + 1) * gridDim.x);
    tc::tc_fence_after();
  }
  __syncthreads();
  if (warp == 1) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, kTmemCols); }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_tc() {
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

bool make_tmap_rows(CUtensorMap* m, const void* ptr, long long rows, long long K, int box_rows) {
  EncodeTiledFn enc = get_encode_tc();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

unsigned int* tc_barrier_ptr(cudaStream_t stream) {
  static unsigned int* base[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!base[dev]) {
    void* p = nullptr;
    if (cudaGetSymbolAddress(&p, g_lstm_tc_barrier) != cudaSuccess) return nullptr;
    base[dev] = (unsigned int*)p;
  }
  cudaMemsetAsync(base[dev], 0, sizeof(unsigned int), stream);
  return base[dev];
}

size_t tc_smem_bytes(int H) {
  return (size_t)(H / kBK) * 2 * kTileW + (size_t)kStages * 2 * kTileA + 1024 + 256;
}

}  // namespace

extern "C" {

// 1 when mb200_highway_lstm_layer_forward_tc supports (H, B): H a multiple of 64 whose weight slice fits shared memory.
int mb200_highway_lstm_tc_supported(int hiddenSize, int miniBatch) {
  if (hiddenSize <= 0 || hiddenSize % kBK != 0 || miniBatch <= 0) return 0;
  if (tc_smem_bytes(hiddenSize) > 227 * 1024) return 0;
  const int grid = (hiddenSize / kU) * mb200_div_up(miniBatch, kBM);
  return grid <= kNumSMs ? 1 : 0;
}

// One layer of the highway-LSTM recurrence on the tensor cores (same contract as mb200_highway_lstm_layer_forward,
// csrc/lstm.cu). Wt_hi / Wt_lo: the recurrent weights as bf16 pairs [H/16 * 80, H], K contiguous, row
// (s * 80 + g * 16 + u) = column (g * H + 16 s + u) of W_h [H, 5H]; hb_hi / hb_lo: [T+1, B, H] bf16, ZERO on entry.
int mb200_highway_lstm_layer_forward_tc(int hiddenSize, int miniBatch, int seqLength, int dir, const float* P,
                                        const void* Wt_hi, const void* Wt_lo, const float* bias, const float* dropout,
                                        float* h, float* c, void* hb_hi, void* hb_lo, float* gates,
                                        const int* lengths_dev, cudaStream_t stream) {
  const int H = hiddenSize, B = miniBatch, TT = seqLength;
  if (H <= 0 || B <= 0 || TT <= 0) return MB200_OK;
  if (!mb200_highway_lstm_tc_supported(H, B)) return MB200_ERR_UNSUPPORTED;
  const int grid = (H / kU) * mb200_div_up(B, kBM);
  const size_t smem = tc_smem_bytes(H);
  static bool attr = false;
  if (!attr) {
    MB200_CHECK(cudaFuncSetAttribute(lstm_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr = true;
  }
  int per_sm = 0;
  MB200_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lstm_fwd_tc_kernel, kThreadsTc, smem));
  if (per_sm < 1) return MB200_ERR_UNSUPPORTED;
  CUtensorMap th, tl, wh, wl;
  const long long hrows = (long long)(TT + 1) * B;
  if (!make_tmap_rows(&th, hb_hi, hrows, H, kBM) || !make_tmap_rows(&tl, hb_lo, hrows, H, kBM) ||
      !make_tmap_rows(&wh, Wt_hi, (long long)(H / kU) * kN, H, kN) ||
      !make_tmap_rows(&wl, Wt_lo, (long long)(H / kU) * kN, H, kN)) {
    mb200_set_error("cuTensorMapEncodeTiled", cudaErrorInvalidValue);
    return MB200_ERR_CUDA;
  }
  TcArgs a;
  a.H = H; a.B = B; a.T = TT; a.dir = dir; a.P = P; a.bias = bias; a.dropout = dropout; a.h = h; a.c = c;
  a.hb_hi = (__nv_bfloat16*)hb_hi; a.hb_lo = (__nv_bfloat16*)hb_lo; a.gates = gates; a.lengths = lengths_dev;
  a.barrier = tc_barrier_ptr(stream);
  if (!a.barrier) return MB200_ERR_CUDA;
  void* args[] = {(void*)&th, (void*)&tl, (void*)&wh, (void*)&wl, (void*)&a};
  MB200_CHECK(cudaLaunchCooperativeKernel((const void*)lstm_fwd_tc_kernel, dim3(grid), dim3(kThreadsTc), args, smem, stream));
  return MB200_OK;
}

}  // extern "C"
