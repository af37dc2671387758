// fp32-faithful tensor-core GEMM and 3x3 convolution for sm_100a: tcgen05.mma (kind::f16, bf16
// operands, fp32 accumulation in TMEM) fed by TMA through an mbarrier ring.
//
// Precision scheme ("bf16x3"): every fp32 operand x is carried as a pair of bf16 tensors
// (hi = bf16(x), lo = bf16(x - hi)); a product is accumulated as hi*hi + hi*lo + lo*hi in the
// same TMEM accumulator.  That keeps ~16 mantissa bits per operand (relative error ~2^-16 per
// product, fp32 accumulation), which is what the parity bar of the north star needs (fp32 logits
// within 1e-3 of the reference's fp32 cuDNN/cuBLAS path through 15 stacked layers) — a single
// bf16 or tf32 pass does not (DESIGN.md "precision").  Per k-block the kernel loads FOUR tiles
// (A_hi, A_lo, B_hi, B_lo) and issues THREE MMAs per 16-wide k step.
//
// Replaces (a) cublasSgemm under nn.Linear: fc6/fc7 (lib/object_detector.py:102-103,129-138,
// lib/rel_model.py:360-374), post_lstm / rel_compress (lib/rel_model.py:377,390), the hoisted LSTM
// input projections; (b) cuDNN under nn.Conv2d for the 3x3/pad-1 VGG convolutions
// (lib/object_detector.py:110-127) as an implicit GEMM whose A tiles are shifted TMA boxes of
// the NHWC activation (out-of-bounds zero fill = the padding), no im2col buffer.
//
// Layout contract: A [M, Kp] and B [N, Kp] bf16, K contiguous ("K-major"), Kp % 64 == 0 with
// zero padding; C row-major. Tile 128 x 128 x 64, 128-byte swizzle, 3-stage ring (64 KB / stage).
// Persistent: one CTA per SM walks a static tile list; warp 0 = TMA producer, warp 1 = MMA issuer +
// TMEM owner, warps 2..5 = epilogue (one TMEM lane quarter each). Two TMEM accumulators (256
// columns) let the epilogue of tile i run under the MMAs of tile i+1.
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int BM = 128, BK = 64;
constexpr int ACC_STAGES = 2;                  // TMEM accumulators: epilogue of tile i overlaps the MMAs of tile i+1
constexpr int TILE_A = BM * BK * 2;            // 16 KB
constexpr int kGemmThreads = 320;          // warp 0: TMA, warp 1: MMA + TMEM owner, warps 2..9: epilogue (two per TMEM lane quarter)
constexpr int kEpiThreads = 256;
constexpr int TH = 8, TW = 16;                 // conv: spatial tile = 128 output pixels

// Two tile shapes. ncu on the 128x128 tile: tensor pipe 34 % with 7.4 TB/s of L2->SM traffic — the
// kernel is bound by operand delivery (64 KB per 768 MMA cycles), so wide outputs (N >= 256) use
// 128x256 tiles: 96 KB per 1536 MMA cycles (-27 % bytes per flop), 2 stages, all 512 TMEM columns.
template <int BN_> struct Cfg {
  static constexpr int BN = BN_;
  static constexpr int STAGES = BN_ == 128 ? 3 : 2;
  static constexpr int TILE_B = BN_ * BK * 2;
  static constexpr int STAGE_BYTES = 2 * TILE_A + 2 * TILE_B;
  static constexpr int TMEM_COLS = ACC_STAGES * BN_;
  static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

struct Params {
  int M, N;                  // logical output size (conv: M = B*H*W pixels, N = Cout)
  int kblocks;               // k-blocks per split
  int splits;
  int m_tiles, n_tiles;      // tile grid (conv: m_tiles = images * tiles_h * tiles_w)
  int m_fastest;             // tile order: 1 = consecutive tiles share the B tile (A is small and L2-resident)
  // epilogue targets (any may be null)
  float* C; long long ldc;
  __nv_bfloat16* Chi; __nv_bfloat16* Clo; long long ldsplit;
  const float* bias; int relu;
  float* partial;            // [splits][M][N] when splits > 1
  // conv mode
  int conv; int H, W, Cin, tiles_w, tiles_h;
};

struct TileCoord { int split, m0, n0, img, h0, w0; };

template <int BN>
__device__ __forceinline__ TileCoord decode_tile(const Params& p, int t) {
  TileCoord tc_;
  const int per_split = p.m_tiles * p.n_tiles;
  tc_.split = t / per_split;
  const int r = t - tc_.split * per_split;
  int mi, ni;
  if (p.m_fastest) { ni = r / p.m_tiles; mi = r - ni * p.m_tiles; }   // B tile streamed once, A re-read from L2
  else { mi = r / p.n_tiles; ni = r - mi * p.n_tiles; }               // neighbouring CTAs share the A tile in L2
  tc_.n0 = ni * BN;
  tc_.m0 = mi * BM; tc_.img = 0; tc_.h0 = 0; tc_.w0 = 0;
  if (p.conv) {
    const int per_img = p.tiles_w * p.tiles_h;
    tc_.img = mi / per_img;
    const int q = mi - tc_.img * per_img;
    const int th = q / p.tiles_w;
    tc_.h0 = th * TH;
    tc_.w0 = (q - th * p.tiles_w) * TW;
  }
  return tc_;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}


// ------------------------------------------------------------------ epilogue of one 128-row x BN tile (all kernels)
// ncu (source page, conv1_2): more than half of all warp samples sat in the epilogue — FADDs waiting for the per-element
// bias LDGs, moves waiting for tcgen05.ld, and a GPU-scope MEMBAR + ERRBAR in front of the release.cluster arrive — with
// one warp per scheduler nothing hides those latencies, and for short-K tiles the epilogue, not the MMA, set the pace.
//  * the tile's bias slice is staged in shared memory BEFORE the wait for the accumulator (broadcast LDS afterwards);
//  * the TMEM read of chunk c+1 is in flight while chunk c is converted and stored;
//  * the hand-back arrive carries no memory fence of its own (see mbar_arrive_remote_cta_release).
//  * EIGHT epilogue warps: a warp may only touch the TMEM lane quarter (warp % 4), so two warps share a quarter and take
//    half of the tile's columns each — twice the issue slots for the convert / store stream of short-K tiles.
// The 256 epilogue threads (warps 2..9) of a CTA call these together; named barrier 1 is theirs.
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

template <int BN>
__device__ __forceinline__ void epi_stage_bias(const Params& p, int n0, float* s_bias, int tid128) {
  epi_bar();                                   // every warp is done reading the previous tile's slice
  if (p.bias) {
    for (int i = tid128; i < BN; i += kEpiThreads) s_bias[i] = (n0 + i < p.N) ? __ldg(p.bias + n0 + i) : 0.f;
  }
  epi_bar();
}

__device__ __forceinline__ void epi_chunk(const Params& p, long long row, int nb, int split, const uint32_t (&v)[32],
                                          const float* s_bias_c) {
  if (row < 0 || nb >= p.N) return;
  const int ncols = min(32, p.N - nb);
  if (p.splits > 1) {
    float* dst = p.partial + ((size_t)split * p.M + row) * p.N + nb;
    if (ncols == 32 && ((((uintptr_t)dst) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < 8; ++j) ((uint4*)dst)[j] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) if (j < ncols) dst[j] = __uint_as_float(v[j]);
    }
    return;
  }
  float x[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    float tt = __uint_as_float(v[j]);
    if (p.bias) tt += s_bias_c[j];
    if (p.relu) tt = fmaxf(tt, 0.f);
    x[j] = tt;
  }
  if (p.C) {
    float* dst = p.C + row * p.ldc + nb;
    if (ncols == 32 && ((((uintptr_t)dst) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < 8; ++j) ((float4*)dst)[j] = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) if (j < ncols) dst[j] = x[j];
    }
  }
  if (p.Chi) {
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float a0 = x[2 * j], a1 = x[2 * j + 1];
      const __nv_bfloat16 h0 = __float2bfloat16_rn(a0), h1 = __float2bfloat16_rn(a1);
      hi[j] = pack_bf16x2(a0, a1);
      lo[j] = pack_bf16x2(a0 - __bfloat162float(h0), a1 - __bfloat162float(h1));
    }
    __nv_bfloat16* dh = p.Chi + row * p.ldsplit + nb;
    __nv_bfloat16* dl = p.Clo + row * p.ldsplit + nb;
    if (ncols == 32 && ((((uintptr_t)dh) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        ((uint4*)dh)[j] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
        ((uint4*)dl)[j] = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) if (j < ncols) {
        __nv_bfloat16 h, l; tc::split_bf16(x[j], h, l);
        dh[j] = h; dl[j] = l;
      }
    }
  }
}

// drain one accumulator (this warp's 32 TMEM lanes x BN columns): software-pipelined TMEM reads, two register buffers
template <int BN>
__device__ __forceinline__ void epi_tile(const Params& p, long long row, int n0, int split, uint32_t taddr, const float* s_bias) {
  uint32_t va[32], vb[32];
  __syncwarp();
  tc::tmem_ld_32x32(taddr, va);
#pragma unroll
  for (int c = 0; c < BN / 32; c += 2) {
    tc::tmem_ld_wait();
    if (c + 1 < BN / 32) { __syncwarp(); tc::tmem_ld_32x32(taddr + (uint32_t)((c + 1) * 32), vb); }
    epi_chunk(p, row, n0 + c * 32, split, va, s_bias + c * 32);
    if (c + 1 < BN / 32) {
      tc::tmem_ld_wait();
      if (c + 2 < BN / 32) { __syncwarp(); tc::tmem_ld_32x32(taddr + (uint32_t)((c + 2) * 32), va); }
      epi_chunk(p, row, n0 + (c + 1) * 32, split, vb, s_bias + (c + 1) * 32);
    }
  }
}

// Persistent: grid = min(#tiles, #SMs); every role loops over the same static tile sequence.
template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16x3_kernel(const __grid_constant__ CUtensorMap tmAhi, const __grid_constant__ CUtensorMap tmAlo,
                   const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo,
                   const Params p) {
  constexpr int STAGES = Cfg<BN>::STAGES, TILE_B = Cfg<BN>::TILE_B, STAGE_BYTES = Cfg<BN>::STAGE_BYTES;
  constexpr int TMEM_COLS = Cfg<BN>::TMEM_COLS;
  extern __shared__ uint8_t smem_raw[];
  __shared__ float s_bias[BN];                   // bias slice of the tile being drained (epi_stage_bias)
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + (size_t)STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + ACC_STAGES;
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + ACC_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.splits * p.m_tiles * p.n_tiles;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmAhi); tc::prefetch_tmap(&tmAlo); tc::prefetch_tmap(&tmBhi); tc::prefetch_tmap(&tmBlo);
    for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < ACC_STAGES; ++s) { tc::mbar_init(&tfull_bar[s], 1); tc::mbar_init(&tempty_bar[s], 8); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, TMEM_COLS);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------------------------------------------------------- TMA producer
      int stage = 0; uint32_t phase = 0;
      const int cblocks = p.conv ? p.Cin / BK : 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const TileCoord tl = decode_tile<BN>(p, t);
        const int kb0 = tl.split * p.kblocks;
        for (int kb = 0; kb < p.kblocks; ++kb) {
          tc::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + (size_t)stage * STAGE_BYTES;
          tc::mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          const int kg = kb0 + kb;
          if (!p.conv) {
            tc::tma_load_2d(st, &tmAhi, &full_bar[stage], kg * BK, tl.m0);
            tc::tma_load_2d(st + TILE_A, &tmAlo, &full_bar[stage], kg * BK, tl.m0);
          } else {
            const int tap = kg / cblocks, cb = kg - tap * cblocks;
            const int dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
            tc::tma_load_4d(st, &tmAhi, &full_bar[stage], cb * BK, tl.w0 + dx, tl.h0 + dy, tl.img);
            tc::tma_load_4d(st + TILE_A, &tmAlo, &full_bar[stage], cb * BK, tl.w0 + dx, tl.h0 + dy, tl.img);
          }
          tc::tma_load_2d(st + 2 * TILE_A, &tmBhi, &full_bar[stage], kg * BK, tl.n0);
          tc::tma_load_2d(st + 2 * TILE_A + TILE_B, &tmBlo, &full_bar[stage], kg * BK, tl.n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------------------------------------------------------- MMA issuer
      constexpr uint32_t idesc = tc::umma_idesc_bf16_f32(BM, BN);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        tc::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);      // epilogue has drained this accumulator
        tc::tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < p.kblocks; ++kb) {
          tc::mbar_wait(&full_bar[stage], phase);
          tc::tc_fence_after();
          const uint32_t sa = tc::smem_u32(smem + (size_t)stage * STAGE_BYTES);
          const uint64_t a_hi = tc::umma_desc_k_sw128(sa), a_lo = tc::umma_desc_k_sw128(sa + TILE_A);
          const uint64_t b_hi = tc::umma_desc_k_sw128(sa + 2 * TILE_A), b_lo = tc::umma_desc_k_sw128(sa + 2 * TILE_A + TILE_B);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);   // 32 bytes per 16-wide k step
            tc::umma_bf16(tmem_d, a_hi + adv, b_hi + adv, idesc, (kb | k) != 0);
            tc::umma_bf16(tmem_d, a_hi + adv, b_lo + adv, idesc, 1);
            tc::umma_bf16(tmem_d, a_lo + adv, b_hi + adv, idesc, 1);
          }
          tc::umma_commit(&empty_bar[stage]);          // frees the smem slot when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc::umma_commit(&tfull_bar[acc]);              // accumulator complete -> epilogue
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;               // tile row == TMEM lane
    const int half = (warp - 2) >> 2;          // which half of the tile's columns (two warps per quarter)
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const TileCoord tl = decode_tile<BN>(p, t);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      epi_stage_bias<BN>(p, tl.n0, s_bias, (int)threadIdx.x - 64);      // while the MMAs of this tile still run
      tc::mbar_wait(&tfull_bar[acc], acc_phase);
      tc::tc_fence_after();
      long long row;                             // output row index (M axis), or -1 when masked
      if (!p.conv) {
        row = (tl.m0 + r < p.M) ? (long long)(tl.m0 + r) : -1;
      } else {
        const int h = tl.h0 + r / TW, w = tl.w0 + (r % TW);
        row = (h < p.H && w < p.W) ? ((long long)tl.img * p.H + h) * p.W + w : -1;
      }
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
      epi_tile<BN / 2>(p, row, tl.n0 + half * (BN / 2), tl.split, taddr + (uint32_t)(half * (BN / 2)), s_bias + half * (BN / 2));
      // accumulator drained: hand it back to the MMA warp (4 arrivals, one per epilogue warp)
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tempty_bar[acc]);
    }
  }
  __syncthreads();
  if (warp == 1) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, TMEM_COLS); }
}


// ====================================================================================== CTA-pair variant
// Same pipeline with cta_group::2 (tc_common.cuh "CTA pair"): a cluster of two CTAs computes a 256 x BN tile with ONE
// tcgen05.mma (M = 256) per k step and operand pair. CTA r of the pair owns rows [128 r, 128 r + 128) of the tile
// (conv: spatial tile 2*pm + r), loads its own A tiles and HALF of the B tile (rows [r*BN/2, (r+1)*BN/2) of the
// weights), and drains its own half of the accumulator. Per SM and k-block: 32 KB of A + 2 * BN/2 * 128 B of B land in
// shared memory instead of 32 KB + 2 * BN * 128 B, and each MMA reads BN/2 instead of BN rows of B per SM — the
// 1-CTA kernel sits at 54-73 % tensor pipe with its shared-memory pipe saturated (profiles/r01_SUMMARY.md).
// Barriers: full[s] of the LEADER counts one arrive.expect_tx (2 * stage bytes) and receives the TMA bytes of both
// CTAs; empty[s] / tfull[a] exist in both CTAs and are signalled by multicast commits; tempty[a] of the leader
// collects the 8 epilogue warps of the pair.
template <int BN_> struct Cfg2 {
  static constexpr int BN = BN_;
  static constexpr int HALF_B = (BN_ / 2) * BK * 2;              // bytes of one B operand half-tile
  static constexpr int STAGE_BYTES = 2 * TILE_A + 2 * HALF_B;    // per CTA
  static constexpr int STAGES = (BN_ == 256) ? 3 : (BN_ == 128 ? 4 : 5);
  static constexpr int TMEM_COLS = ACC_STAGES * BN_;
  static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

// pair-tile t -> coordinates of THIS CTA's half (rank r): p.m_tiles counts 128-row tiles, pairs walk ceil(m_tiles / 2)
template <int BN>
__device__ __forceinline__ TileCoord decode_tile2(const Params& p, int t, int rank, int pm_tiles) {
  TileCoord tc_;
  const int per_split = pm_tiles * p.n_tiles;
  tc_.split = t / per_split;
  const int r = t - tc_.split * per_split;
  int pmi, ni;
  if (p.m_fastest) { ni = r / pm_tiles; pmi = r - ni * pm_tiles; }
  else { pmi = r / p.n_tiles; ni = r - pmi * p.n_tiles; }
  const int mi = 2 * pmi + rank;                 // may equal p.m_tiles for the last pair: all rows masked, TMA zero fill
  tc_.n0 = ni * BN;
  tc_.m0 = mi * BM; tc_.img = 0; tc_.h0 = 0; tc_.w0 = 0;
  if (p.conv) {
    const int per_img = p.tiles_w * p.tiles_h;
    tc_.img = mi / per_img;                      // == B for the dummy half: out of bounds in the image dimension
    const int q = mi - tc_.img * per_img;
    const int th = q / p.tiles_w;
    tc_.h0 = th * TH;
    tc_.w0 = (q - th * p.tiles_w) * TW;
  }
  return tc_;
}

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_bf16x3_2cta_kernel(const __grid_constant__ CUtensorMap tmAhi, const __grid_constant__ CUtensorMap tmAlo,
                        const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo,
                        const Params p, const int num_images) {
  constexpr int STAGES = Cfg2<BN>::STAGES, HALF_B = Cfg2<BN>::HALF_B, STAGE_BYTES = Cfg2<BN>::STAGE_BYTES;
  constexpr int TMEM_COLS = Cfg2<BN>::TMEM_COLS;
  extern __shared__ uint8_t smem_raw[];
  __shared__ float s_bias[BN];                   // bias slice of the tile being drained (epi_stage_bias)
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + (size_t)STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + ACC_STAGES;
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + ACC_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = tc::cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int pm_tiles = (p.m_tiles + 1) >> 1;
  const int total_tiles = p.splits * pm_tiles * p.n_tiles;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmAhi); tc::prefetch_tmap(&tmAlo); tc::prefetch_tmap(&tmBhi); tc::prefetch_tmap(&tmBlo);
    for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < ACC_STAGES; ++s) { tc::mbar_init(&tfull_bar[s], 1); tc::mbar_init(&tempty_bar[s], 16); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc_2sm(tmem_slot, TMEM_COLS);
  tc::tc_fence_before();
  tc::cluster_sync_all();                       // barriers of BOTH CTAs initialised before any remote arrive / TMA
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------------------------------------------------------- TMA producer (both CTAs)
      int stage = 0; uint32_t phase = 0;
      const int cblocks = p.conv ? p.Cin / BK : 0;
      for (int t = pair; t < total_tiles; t += npairs) {
        const TileCoord tl = decode_tile2<BN>(p, t, (int)rank, pm_tiles);
        const int kb0 = tl.split * p.kblocks;
        const int nrow = tl.n0 + (int)rank * (BN / 2);          // this CTA's half of the B tile
        for (int kb = 0; kb < p.kblocks; ++kb) {
          tc::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + (size_t)stage * STAGE_BYTES;
          if (rank == 0) tc::mbar_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
          const uint32_t fb = tc::mapa_shared(tc::smem_u32(&full_bar[stage]), 0);
          const int kg = kb0 + kb;
          if (!p.conv) {
            tc::tma_load_2d_2sm(st, &tmAhi, fb, kg * BK, tl.m0);
            tc::tma_load_2d_2sm(st + TILE_A, &tmAlo, fb, kg * BK, tl.m0);
          } else {
            const int tap = kg / cblocks, cb = kg - tap * cblocks;
            const int dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
            tc::tma_load_4d_2sm(st, &tmAhi, fb, cb * BK, tl.w0 + dx, tl.h0 + dy, tl.img);
            tc::tma_load_4d_2sm(st + TILE_A, &tmAlo, fb, cb * BK, tl.w0 + dx, tl.h0 + dy, tl.img);
          }
          tc::tma_load_2d_2sm(st + 2 * TILE_A, &tmBhi, fb, kg * BK, nrow);
          tc::tma_load_2d_2sm(st + 2 * TILE_A + HALF_B, &tmBlo, fb, kg * BK, nrow);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ---------------------------------------------------------------- MMA issuer (leader CTA only)
      constexpr uint32_t idesc = tc::umma_idesc_bf16_f32(2 * BM, BN);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int t = pair; t < total_tiles; t += npairs, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        tc::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);      // both CTAs' epilogues have drained this accumulator
        tc::tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < p.kblocks; ++kb) {
          tc::mbar_wait(&full_bar[stage], phase);            // bytes of BOTH CTAs have landed
          tc::tc_fence_after();
          const uint32_t sa = tc::smem_u32(smem + (size_t)stage * STAGE_BYTES);
          const uint64_t a_hi = tc::umma_desc_k_sw128(sa), a_lo = tc::umma_desc_k_sw128(sa + TILE_A);
          const uint64_t b_hi = tc::umma_desc_k_sw128(sa + 2 * TILE_A), b_lo = tc::umma_desc_k_sw128(sa + 2 * TILE_A + HALF_B);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);
            tc::umma_bf16_2sm(tmem_d, a_hi + adv, b_hi + adv, idesc, (kb | k) != 0);
            tc::umma_bf16_2sm(tmem_d, a_hi + adv, b_lo + adv, idesc, 1);
            tc::umma_bf16_2sm(tmem_d, a_lo + adv, b_hi + adv, idesc, 1);
          }
          tc::umma_commit_2sm(&empty_bar[stage], 3);          // frees the slot in both CTAs
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc::umma_commit_2sm(&tfull_bar[acc], 3);              // accumulator complete -> both epilogues
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5 of both CTAs)
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int half = (warp - 2) >> 2;
    const uint32_t tempty_leader0 = tc::mapa_shared(tc::smem_u32(&tempty_bar[0]), 0);
    int it = 0;
    for (int t = pair; t < total_tiles; t += npairs, ++it) {
      const TileCoord tl = decode_tile2<BN>(p, t, (int)rank, pm_tiles);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      epi_stage_bias<BN>(p, tl.n0, s_bias, (int)threadIdx.x - 64);
      tc::mbar_wait(&tfull_bar[acc], acc_phase);
      tc::tc_fence_after();
      long long row;
      if (!p.conv) {
        row = (tl.m0 + r < p.M) ? (long long)(tl.m0 + r) : -1;
      } else {
        const int h = tl.h0 + r / TW, w = tl.w0 + (r % TW);
        row = (tl.img < num_images && h < p.H && w < p.W) ? ((long long)tl.img * p.H + h) * p.W + w : -1;
      }
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
      epi_tile<BN / 2>(p, row, tl.n0 + half * (BN / 2), tl.split, taddr + (uint32_t)(half * (BN / 2)), s_bias + half * (BN / 2));
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive_remote(tempty_leader0 + (uint32_t)(acc * sizeof(uint64_t)));
    }
  }
  // neither CTA may leave (or free TMEM) while the other can still signal its barriers / read its shared memory
  tc::tc_fence_before();
  tc::cluster_sync_all();
  if (warp == 1) { tc::tc_fence_after(); tc::tmem_dealloc_2sm(tmem_base, TMEM_COLS); }
}

// ====================================================================================== conv with a shared-memory halo
// ncu / timing of the tap-by-tap kernels above: conv1_2 (64 -> 64 channels) takes the same 790 us with 64-, 128-wide or
// 1-CTA tiles — nothing tensor-bound about it. Every 128-pixel A tile is fetched NINE times from L2 (once per tap,
// shifted), 4.8 GB for that layer = 6.1 TB/s of L2 -> SM traffic, which is the B200's L2 bandwidth; the 256/512-channel
// layers sit at the same wall (activation tiles 9x + weight tiles once per pixel tile).
// Here a CTA stages the input HALO of its tile once per 64-channel block — 18 lines x 16 pixels (10 used) x 64 ch, one
// 4-D TMA box per operand half — and all nine taps are read out of it: the A descriptor of tap (dy, dx) simply starts
// (dy+1) lines and (dx+1) pixels into the staged block. Tile = 16 lines x 8 pixels, so that one 8-row group of the MMA
// operand is one line of the tile and consecutive groups are one staged line (16 px x 128 B = 2048 B, a multiple of the
// 1024-byte swizzle atom) apart: SBO = 2048. L2 traffic for A drops from 9x to 2.25x of the input.
// Weights still stream per tap through their own ring. CTA pairs (cta_group::2) as above.
constexpr int HTH = 16, HTW = 8;                       // output tile of one CTA: 16 lines x 8 pixels = 128 rows
constexpr int HALO_LINES = HTH + 2, HALO_PX = 16;      // staged block: 18 lines x 16 pixels (pixels w0-1 .. w0+14)
constexpr int HALO_BYTES = HALO_LINES * HALO_PX * 128; // 36 864 B per operand half
constexpr int HALO_STAGES = 2;

template <int BN_> struct CfgH {
  static constexpr int HALF_B = (BN_ / 2) * BK * 2;
  static constexpr int B_STAGE = 2 * HALF_B;
  static constexpr int B_STAGES = (BN_ == 256) ? 2 : (BN_ == 128 ? 4 : 6);
  static constexpr int A_STAGE = 2 * HALO_BYTES;
  static constexpr int TMEM_COLS = ACC_STAGES * BN_;
  static constexpr size_t SMEM_BYTES = (size_t)HALO_STAGES * A_STAGE + (size_t)B_STAGES * B_STAGE + 1024 + 256;
};

// 0: tap-by-tap kernels; 1 (default): halo kernel where it measured faster (large maps, few input channels: conv1_2 792 ->
// 692 us, conv2_2 401 -> 390 us; it loses 5-10 % on the 74x74 / 37x37 maps); 2: halo kernel for every conv (tests).
int g_halo_mode = 1;

// The start address is a whole number of 128-byte rows into a 1024-byte swizzle atom. Measured on the B200: the 128-byte
// swizzle is a function of the ABSOLUTE shared-memory address (TMA writes and tcgen05 reads agree without further ado) —
// with the descriptor's base-offset field set to (addr >> 7) & 7 the results are wrong, with 0 they are exact.
__device__ __forceinline__ uint64_t umma_desc_halo(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((HALO_PX * 128) >> 4) << 32;         // SBO: the next 8-row group is the next staged line
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
conv3x3_halo_2cta_kernel(const __grid_constant__ CUtensorMap tmAhi, const __grid_constant__ CUtensorMap tmAlo,
                         const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo,
                         const Params p, const int num_images) {
  constexpr int HALF_B = CfgH<BN>::HALF_B, B_STAGE = CfgH<BN>::B_STAGE, B_STAGES = CfgH<BN>::B_STAGES;
  constexpr int A_STAGE = CfgH<BN>::A_STAGE, TMEM_COLS = CfgH<BN>::TMEM_COLS;
  extern __shared__ uint8_t smem_raw[];
  __shared__ float s_bias[BN];                   // bias slice of the tile being drained (epi_stage_bias)
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_b = smem + (size_t)HALO_STAGES * A_STAGE;
  uint64_t* afull = (uint64_t*)(smem_b + (size_t)B_STAGES * B_STAGE);
  uint64_t* aempty = afull + HALO_STAGES;
  uint64_t* bfull = aempty + HALO_STAGES;
  uint64_t* bempty = bfull + B_STAGES;
  uint64_t* tfull_bar = bempty + B_STAGES;
  uint64_t* tempty_bar = tfull_bar + ACC_STAGES;
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + ACC_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = tc::cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int pm_tiles = (p.m_tiles + 1) >> 1;
  const int total_tiles = pm_tiles * p.n_tiles;
  const int cblocks = p.Cin / BK;
  const int per_img = p.tiles_w * p.tiles_h;

  // this CTA's half of pair-tile t: spatial tile mi = 2*pmi + rank (== m_tiles for the dummy half of an odd last pair)
  auto decode = [&](int t, int& n0, int& img, int& h0, int& w0) {
    const int pmi = t / p.n_tiles, ni = t - pmi * p.n_tiles;      // n fastest: the pair's staged activations serve all n tiles from L2
    const int mi = 2 * pmi + (int)rank;
    n0 = ni * BN;
    img = mi / per_img;
    const int q = mi - img * per_img;
    const int th = q / p.tiles_w;
    h0 = th * HTH; w0 = (q - th * p.tiles_w) * HTW;
  };

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmAhi); tc::prefetch_tmap(&tmAlo); tc::prefetch_tmap(&tmBhi); tc::prefetch_tmap(&tmBlo);
    for (int s = 0; s < HALO_STAGES; ++s) { tc::mbar_init(&afull[s], 1); tc::mbar_init(&aempty[s], 1); }
    for (int s = 0; s < B_STAGES; ++s) { tc::mbar_init(&bfull[s], 1); tc::mbar_init(&bempty[s], 1); }
    for (int s = 0; s < ACC_STAGES; ++s) { tc::mbar_init(&tfull_bar[s], 1); tc::mbar_init(&tempty_bar[s], 16); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc_2sm(tmem_slot, TMEM_COLS);
  tc::tc_fence_before();
  tc::cluster_sync_all();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------------------------------------------------------- TMA producer (both CTAs)
      int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
      for (int t = pair; t < total_tiles; t += npairs) {
        int n0, img, h0, w0;
        decode(t, n0, img, h0, w0);
        const int nrow = n0 + (int)rank * (BN / 2);
        for (int cb = 0; cb < cblocks; ++cb) {
          tc::mbar_wait(&aempty[sa], pa ^ 1);
          uint8_t* st = smem + (size_t)sa * A_STAGE;
          if (rank == 0) tc::mbar_expect_tx(&afull[sa], 2 * A_STAGE);
          const uint32_t fa = tc::mapa_shared(tc::smem_u32(&afull[sa]), 0);
          tc::tma_load_4d_2sm(st, &tmAhi, fa, cb * BK, w0 - 1, h0 - 1, img);               // out of bounds = zero padding
          tc::tma_load_4d_2sm(st + HALO_BYTES, &tmAlo, fa, cb * BK, w0 - 1, h0 - 1, img);
          if (++sa == HALO_STAGES) { sa = 0; pa ^= 1; }
          for (int tap = 0; tap < 9; ++tap) {
            tc::mbar_wait(&bempty[sb], pb ^ 1);
            uint8_t* sbp = smem_b + (size_t)sb * B_STAGE;
            if (rank == 0) tc::mbar_expect_tx(&bfull[sb], 2 * B_STAGE);
            const uint32_t fb = tc::mapa_shared(tc::smem_u32(&bfull[sb]), 0);
            const int kg = tap * cblocks + cb;                                               // weight K order: (kh, kw, cin)
            tc::tma_load_2d_2sm(sbp, &tmBhi, fb, kg * BK, nrow);
            tc::tma_load_2d_2sm(sbp + HALF_B, &tmBlo, fb, kg * BK, nrow);
            if (++sb == B_STAGES) { sb = 0; pb ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ---------------------------------------------------------------- MMA issuer (leader CTA)
      constexpr uint32_t idesc = tc::umma_idesc_bf16_f32(2 * BM, BN);
      int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
      int it = 0;
      for (int t = pair; t < total_tiles; t += npairs, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        tc::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc::tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int cb = 0; cb < cblocks; ++cb) {
          tc::mbar_wait(&afull[sa], pa);
          tc::tc_fence_after();
          const uint32_t a_base = tc::smem_u32(smem + (size_t)sa * A_STAGE);
          for (int tap = 0; tap < 9; ++tap) {
            tc::mbar_wait(&bfull[sb], pb);
            tc::tc_fence_after();
            const int ky = tap / 3, kx = tap - ky * 3;                                       // = dy + 1, dx + 1
            const uint32_t a_off = (uint32_t)(ky * HALO_PX * 128 + kx * 128);
            const uint32_t sbb = tc::smem_u32(smem_b + (size_t)sb * B_STAGE);
            const uint64_t b_hi = tc::umma_desc_k_sw128(sbb), b_lo = tc::umma_desc_k_sw128(sbb + HALF_B);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t a_hi = umma_desc_halo(a_base + a_off + k * 32);
              const uint64_t a_lo = umma_desc_halo(a_base + HALO_BYTES + a_off + k * 32);
              const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);
              tc::umma_bf16_2sm(tmem_d, a_hi, b_hi + adv, idesc, (cb | tap | k) != 0);
              tc::umma_bf16_2sm(tmem_d, a_hi, b_lo + adv, idesc, 1);
              tc::umma_bf16_2sm(tmem_d, a_lo, b_hi + adv, idesc, 1);
            }
            tc::umma_commit_2sm(&bempty[sb], 3);
            if (++sb == B_STAGES) { sb = 0; pb ^= 1; }
          }
          tc::umma_commit_2sm(&aempty[sa], 3);
          if (++sa == HALO_STAGES) { sa = 0; pa ^= 1; }
        }
        tc::umma_commit_2sm(&tfull_bar[acc], 3);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5 of both CTAs)
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int half = (warp - 2) >> 2;
    const uint32_t tempty_leader0 = tc::mapa_shared(tc::smem_u32(&tempty_bar[0]), 0);
    int it = 0;
    for (int t = pair; t < total_tiles; t += npairs, ++it) {
      int n0, img, h0, w0;
      decode(t, n0, img, h0, w0);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      epi_stage_bias<BN>(p, n0, s_bias, (int)threadIdx.x - 64);
      tc::mbar_wait(&tfull_bar[acc], acc_phase);
      tc::tc_fence_after();
      const int h = h0 + r / HTW, w = w0 + (r % HTW);
      const long long row = (img < num_images && h < p.H && w < p.W) ? ((long long)img * p.H + h) * p.W + w : -1;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
      epi_tile<BN / 2>(p, row, n0 + half * (BN / 2), 0, taddr + (uint32_t)(half * (BN / 2)), s_bias + half * (BN / 2));
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive_remote(tempty_leader0 + (uint32_t)(acc * sizeof(uint64_t)));
    }
  }
  tc::tc_fence_before();
  tc::cluster_sync_all();
  if (warp == 1) { tc::tc_fence_after(); tc::tmem_dealloc_2sm(tmem_base, TMEM_COLS); }
}

// split-K second pass: out = sum_z partial[z] (+ bias) (relu) -> fp32 and/or split bf16
__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int splits, long long MN, int N,
                                     const float* __restrict__ bias, int relu, float* __restrict__ C, long long ldc,
                                     __nv_bfloat16* __restrict__ Chi, __nv_bfloat16* __restrict__ Clo, long long ldsplit) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < MN; i += (long long)blockDim.x * gridDim.x) {
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += partial[(size_t)z * MN + i];
    const long long m = i / N; const int n = (int)(i - m * N);
    if (bias) s += bias[n];
    if (relu) s = fmaxf(s, 0.f);
    if (C) C[m * ldc + n] = s;
    if (Chi) { __nv_bfloat16 h, l; tc::split_bf16(s, h, l); Chi[m * ldsplit + n] = h; Clo[m * ldsplit + n] = l; }
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
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

// 2D K-major bf16 matrix [rows, Kp] -> box {64, box_rows}, 128B swizzle, zero fill.
bool make_tmap_2d(CUtensorMap* m, const void* ptr, long long rows, long long Kp, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)Kp * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// 4D NHWC bf16 activation [B,H,W,C] -> box {64, TW, TH, 1}
bool make_tmap_nhwc(CUtensorMap* m, const void* ptr, int B, int H, int W, int C) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)TW, (cuuint32_t)TH, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// 4D NHWC bf16 activation [B,H,W,C] -> halo box {64, 16 px, 18 lines, 1}
bool make_tmap_halo(CUtensorMap* m, const void* ptr, int B, int H, int W, int C) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)HALO_PX, (cuuint32_t)HALO_LINES, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int ensure_attr() {
  static bool done = false;
  if (!done) {
    MB200_CHECK(cudaFuncSetAttribute(gemm_bf16x3_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<128>::SMEM_BYTES));
    MB200_CHECK(cudaFuncSetAttribute(gemm_bf16x3_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<256>::SMEM_BYTES));
    MB200_CHECK(cudaFuncSetAttribute(gemm_bf16x3_2cta_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg2<64>::SMEM_BYTES));
    MB200_CHECK(cudaFuncSetAttribute(gemm_bf16x3_2cta_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg2<128>::SMEM_BYTES));
    MB200_CHECK(cudaFuncSetAttribute(gemm_bf16x3_2cta_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg2<256>::SMEM_BYTES));
    MB200_CHECK(cudaFuncSetAttribute(conv3x3_halo_2cta_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CfgH<64>::SMEM_BYTES));
    MB200_CHECK(cudaFuncSetAttribute(conv3x3_halo_2cta_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CfgH<128>::SMEM_BYTES));
    MB200_CHECK(cudaFuncSetAttribute(conv3x3_halo_2cta_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CfgH<256>::SMEM_BYTES));
    done = true;
  }
  return MB200_OK;
}

// 0: 1-CTA kernels only; 1: choose per shape (default); 2: CTA-pair kernel whenever the shape allows (tests, A/B runs)
int g_pair_mode = 1;

inline double wave_eff(long long tiles, int slots) {
  return (double)tiles / (double)(((tiles + slots - 1) / slots) * slots);
}

// Tile width for an N-wide output: 128x256 tiles move 27 % fewer operand bytes per flop, but the
// persistent grid is quantised in waves of one tile per SM — weigh both.
inline int pick_bn(long long m_tiles, int N) {
  if (N < 256) return 128;
  const long long t128 = m_tiles * ((N + 127) / 128), t256 = m_tiles * ((N + 255) / 256);
  const double waste256 = (double)N / (double)(((N + 255) / 256) * 256);     // column padding
  const double waste128 = (double)N / (double)(((N + 127) / 128) * 128);
  return (wave_eff(t256, kNumSMs) * waste256 * 1.3 > wave_eff(t128, kNumSMs) * waste128) ? 256 : 128;
}

// CTA-pair plan for m_tiles 128-row tiles and an N-wide output: BN (0 = use the 1-CTA kernel). The pair kernel
// runs the tensor pipe faster per tile (operand delivery no longer binds) but needs >= 2 row tiles and
// quantises in waves of 74 pair-tiles.
inline int pick_pair_bn(long long m_tiles, int N, int splits, int bn1) {
  if (g_pair_mode == 0 || m_tiles < 2 || splits > 1) return 0;
  const long long pm = (m_tiles + 1) / 2;
  const int pairs = kNumSMs / 2;
  const double rows = (double)m_tiles / (double)(2 * pm);                     // the dummy half of an odd last pair
  double best = 0.0; int best_bn = 0;
  for (int bn = 64; bn <= 256; bn *= 2) {       // 64-wide tiles only where a 128-wide one would be half padding (conv1_2: Cout = 64)
    if ((bn == 256 && N < 256) || (bn == 64 && N > 64)) continue;
    const double colw = (double)N / (double)(((N + bn - 1) / bn) * bn);
    const double e = wave_eff(pm * ((N + bn - 1) / bn), pairs) * colw * rows * (bn == 256 ? 1.0 : (bn == 128 ? 0.93 : 0.8));
    if (e > best) { best = e; best_bn = bn; }
  }
  if (g_pair_mode == 2) return best_bn;
  const double colw1 = (double)N / (double)(((N + bn1 - 1) / bn1) * bn1);
  const double e1 = wave_eff(m_tiles * ((N + bn1 - 1) / bn1), kNumSMs) * colw1;
  return (best * 1.15 > e1) ? best_bn : 0;
}

int launch(const CUtensorMap& ahi, const CUtensorMap& alo, const CUtensorMap& bhi, const CUtensorMap& blo,
           const Params& p, int bn, int pair_bn, int num_images, cudaStream_t stream) {
  int rc = ensure_attr();
  if (rc != MB200_OK) return rc;
  if (pair_bn) {
    const long long tiles = (long long)p.splits * ((p.m_tiles + 1) / 2) * p.n_tiles;
    if (tiles > 0x7fffffffLL) return MB200_ERR_UNSUPPORTED;
    const int grid = 2 * (int)min(tiles, (long long)(g_mb200_sm_budget / 2));   // persistent: one CTA pair per TPC
    if (pair_bn == 256)
      gemm_bf16x3_2cta_kernel<256><<<grid, kGemmThreads, Cfg2<256>::SMEM_BYTES, stream>>>(ahi, alo, bhi, blo, p, num_images);
    else if (pair_bn == 128)
      gemm_bf16x3_2cta_kernel<128><<<grid, kGemmThreads, Cfg2<128>::SMEM_BYTES, stream>>>(ahi, alo, bhi, blo, p, num_images);
    else
      gemm_bf16x3_2cta_kernel<64><<<grid, kGemmThreads, Cfg2<64>::SMEM_BYTES, stream>>>(ahi, alo, bhi, blo, p, num_images);
    MB200_CHECK_LAUNCH("gemm_bf16x3_2cta_kernel");
    return MB200_OK;
  }
  const long long tiles = (long long)p.splits * p.m_tiles * p.n_tiles;
  if (tiles > 0x7fffffffLL) return MB200_ERR_UNSUPPORTED;
  const int grid = (int)min(tiles, (long long)g_mb200_sm_budget);      // persistent: one CTA per SM
  if (bn == 256) gemm_bf16x3_kernel<256><<<grid, kGemmThreads, Cfg<256>::SMEM_BYTES, stream>>>(ahi, alo, bhi, blo, p);
  else gemm_bf16x3_kernel<128><<<grid, kGemmThreads, Cfg<128>::SMEM_BYTES, stream>>>(ahi, alo, bhi, blo, p);
  MB200_CHECK_LAUNCH("gemm_bf16x3_kernel");
  return MB200_OK;
}

int finish_splitk(const Params& p, cudaStream_t stream) {
  if (p.splits > 1) {
    const long long MN = (long long)p.M * p.N;
    const int blocks = (int)min((long long)kNumSMs * 8, (MN + 255) / 256);
    splitk_reduce_kernel<<<blocks, 256, 0, stream>>>(p.partial, p.splits, MN, p.N, p.bias, p.relu, p.C, p.ldc,
                                                     p.Chi, p.Clo, p.ldsplit);
    MB200_CHECK_LAUNCH("splitk_reduce_kernel");
  }
  return MB200_OK;
}

}  // namespace

extern "C" {

// Floats of split-K workspace mb200_gemm_bf16x3 may need for an [M,N] output (0 when it will not split).
long long mb200_gemm_workspace_floats(int M, int N, int Kp) {
  const long long tiles = (long long)mb200_div_up(M, BM) * mb200_div_up(N, 128);
  const int kblocks = Kp / BK;
  if (tiles >= kNumSMs / 2 || kblocks < 8) return 0;
  int splits = (int)min((long long)kblocks / 4, (long long)(kNumSMs / tiles));
  if (splits < 2) return 0;
  while (kblocks % splits) --splits;
  return splits > 1 ? (long long)splits * M * N : 0;
}

// C[M,N] = A[M,K] * B[N,K]^T (+ bias[N]) (ReLU), A and B given as (hi, lo) bf16 pairs with row
// pitch Kp (multiple of 64, zero padded). Outputs: C fp32 (ldc) and/or (Chi, Clo) bf16 pair
// (ldsplit), any may be NULL. workspace: mb200_gemm_workspace_floats() floats or NULL.
int mb200_gemm_bf16x3(const void* Ahi, const void* Alo, const void* Bhi, const void* Blo, int M, int N, int Kp,
                      const float* bias, int relu, float* C, long long ldc, void* Chi, void* Clo,
                      long long ldsplit, float* workspace, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return MB200_OK;
  if (Kp <= 0 || Kp % BK != 0) return MB200_ERR_ARG;
  CUtensorMap ta, tal, tb, tbl;
  Params p = {};
  p.M = M; p.N = N;
  const int kblocks = Kp / BK;
  p.splits = 1;
  if (workspace) {
    const long long ws = mb200_gemm_workspace_floats(M, N, Kp);
    if (ws > 0) p.splits = (int)(ws / ((long long)M * N));
  }
  const int bn = p.splits > 1 ? 128 : pick_bn(mb200_div_up(M, BM), N);
  const int pair_bn = pick_pair_bn(mb200_div_up(M, BM), N, p.splits, bn);
  const int box_b = pair_bn ? pair_bn / 2 : bn;
  if (!make_tmap_2d(&ta, Ahi, M, Kp, BM) || !make_tmap_2d(&tal, Alo, M, Kp, BM) ||
      !make_tmap_2d(&tb, Bhi, N, Kp, box_b) || !make_tmap_2d(&tbl, Blo, N, Kp, box_b)) {
    mb200_set_error("cuTensorMapEncodeTiled", cudaErrorInvalidValue);
    return MB200_ERR_CUDA;
  }
  p.kblocks = kblocks / p.splits;
  p.C = C; p.ldc = ldc; p.Chi = (__nv_bfloat16*)Chi; p.Clo = (__nv_bfloat16*)Clo; p.ldsplit = ldsplit;
  p.bias = bias; p.relu = relu; p.partial = workspace; p.conv = 0;
  p.m_tiles = mb200_div_up(M, BM); p.n_tiles = mb200_div_up(N, pair_bn ? pair_bn : bn);
  // ncu: fc6 dX / dW re-streamed the 150-400 MB B operand once per m-tile (2.7 / 3.3 GB of DRAM reads).
  // When A (hi+lo) fits comfortably in L2 and B is the larger operand, walk m fastest instead.
  const double a_bytes = 4.0 * M * Kp, b_bytes = 4.0 * N * Kp;
  p.m_fastest = (a_bytes <= 48e6 && b_bytes > a_bytes) ? 1 : 0;
  int rc = launch(ta, tal, tb, tbl, p, bn, pair_bn, 0, stream);
  return rc != MB200_OK ? rc : finish_splitk(p, stream);
}

// 3x3 / stride 1 / pad 1 convolution as implicit GEMM. x: NHWC bf16 pair [B,H,W,Cin] (Cin % 64 == 0);
// w: [Cout, 9*Cin] bf16 pair, K order (kh, kw, cin); outputs NHWC: y fp32 and/or (yhi, ylo), any NULL.
int mb200_conv3x3_bf16x3(const void* xhi, const void* xlo, const void* whi, const void* wlo, int B, int H, int W,
                         int Cin, int Cout, const float* bias, int relu, float* y, void* yhi, void* ylo,
                         cudaStream_t stream) {
  if (B <= 0 || H <= 0 || W <= 0 || Cout <= 0) return MB200_OK;
  if (Cin % BK != 0) return MB200_ERR_ARG;
  CUtensorMap ta, tal, tb, tbl;
  const long long Kp = 9LL * Cin;
  if (g_pair_mode != 0 && (g_halo_mode == 2 || (g_halo_mode == 1 && Cin <= 128 && (long long)H * W >= 80000))) {
    // halo kernel: 16 x 8 pixel tiles, always CTA pairs (a single tile gets a dummy partner)
    const int tw = mb200_div_up(W, HTW), th = mb200_div_up(H, HTH);
    const long long sp = (long long)B * tw * th;
    int bn = 64;
    if (Cout > 64) {
      const long long pm = (sp + 1) / 2;
      const double e128 = wave_eff(pm * ((Cout + 127) / 128), kNumSMs / 2) * ((double)Cout / (((Cout + 127) / 128) * 128)) * 0.93;
      const double e256 = wave_eff(pm * ((Cout + 255) / 256), kNumSMs / 2) * ((double)Cout / (((Cout + 255) / 256) * 256));
      bn = (Cout >= 256 && e256 >= e128) ? 256 : 128;
    }
    if (!make_tmap_halo(&ta, xhi, B, H, W, Cin) || !make_tmap_halo(&tal, xlo, B, H, W, Cin) ||
        !make_tmap_2d(&tb, whi, Cout, Kp, bn / 2) || !make_tmap_2d(&tbl, wlo, Cout, Kp, bn / 2)) {
      mb200_set_error("cuTensorMapEncodeTiled", cudaErrorInvalidValue);
      return MB200_ERR_CUDA;
    }
    int rc = ensure_attr();
    if (rc != MB200_OK) return rc;
    Params p = {};
    p.M = B * H * W; p.N = Cout; p.splits = 1; p.kblocks = (int)(Kp / BK);
    p.C = y; p.ldc = Cout; p.Chi = (__nv_bfloat16*)yhi; p.Clo = (__nv_bfloat16*)ylo; p.ldsplit = Cout;
    p.bias = bias; p.relu = relu; p.partial = nullptr;
    p.conv = 1; p.H = H; p.W = W; p.Cin = Cin; p.tiles_w = tw; p.tiles_h = th;
    p.m_tiles = (int)sp; p.n_tiles = mb200_div_up(Cout, bn);
    const long long tiles = ((sp + 1) / 2) * p.n_tiles;
    if (tiles > 0x7fffffffLL) return MB200_ERR_UNSUPPORTED;
    const int grid = 2 * (int)min(tiles, (long long)(g_mb200_sm_budget / 2));
    if (bn == 256) conv3x3_halo_2cta_kernel<256><<<grid, kGemmThreads, CfgH<256>::SMEM_BYTES, stream>>>(ta, tal, tb, tbl, p, B);
    else if (bn == 128) conv3x3_halo_2cta_kernel<128><<<grid, kGemmThreads, CfgH<128>::SMEM_BYTES, stream>>>(ta, tal, tb, tbl, p, B);
    else conv3x3_halo_2cta_kernel<64><<<grid, kGemmThreads, CfgH<64>::SMEM_BYTES, stream>>>(ta, tal, tb, tbl, p, B);
    MB200_CHECK_LAUNCH("conv3x3_halo_2cta_kernel");
    return MB200_OK;
  }
  const long long sp_tiles = (long long)B * mb200_div_up(W, TW) * mb200_div_up(H, TH);
  const int bn = pick_bn(sp_tiles, Cout);
  const int pair_bn = pick_pair_bn(sp_tiles, Cout, 1, bn);
  const int box_b = pair_bn ? pair_bn / 2 : bn;
  if (!make_tmap_nhwc(&ta, xhi, B, H, W, Cin) || !make_tmap_nhwc(&tal, xlo, B, H, W, Cin) ||
      !make_tmap_2d(&tb, whi, Cout, Kp, box_b) || !make_tmap_2d(&tbl, wlo, Cout, Kp, box_b)) {
    mb200_set_error("cuTensorMapEncodeTiled", cudaErrorInvalidValue);
    return MB200_ERR_CUDA;
  }
  Params p = {};
  p.M = B * H * W; p.N = Cout; p.splits = 1; p.kblocks = (int)(Kp / BK);
  p.C = y; p.ldc = Cout; p.Chi = (__nv_bfloat16*)yhi; p.Clo = (__nv_bfloat16*)ylo; p.ldsplit = Cout;
  p.bias = bias; p.relu = relu; p.partial = nullptr;
  p.conv = 1; p.H = H; p.W = W; p.Cin = Cin; p.tiles_w = mb200_div_up(W, TW); p.tiles_h = mb200_div_up(H, TH);
  p.m_tiles = B * p.tiles_w * p.tiles_h; p.n_tiles = mb200_div_up(Cout, pair_bn ? pair_bn : bn);
  return launch(ta, tal, tb, tbl, p, bn, pair_bn, B, stream);
}

/* conv3x3 kernel selection: 0 = tap-by-tap implicit GEMM (shifted TMA boxes), 1 = per layer (default: the shared-memory
 * halo kernel for large maps with <= 128 input channels), 2 = halo kernel always. Returns the previous mode. */
int mb200_conv_set_halo_mode(int mode) {
  const int old = g_halo_mode;
  if (mode >= 0 && mode <= 2) g_halo_mode = mode;
  return old;
}

/* 0: 1-CTA kernels only; 1: per-shape choice (default); 2: the CTA-pair (cta_group::2) kernel whenever the shape allows.
 * Returns the previous mode. For tests and A/B measurements. */
int mb200_gemm_set_pair_mode(int mode) {
  const int old = g_pair_mode;
  if (mode >= 0 && mode <= 2) g_pair_mode = mode;
  return old;
}

}  // extern "C"
