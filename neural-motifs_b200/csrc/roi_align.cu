// RoIAlign (TF crop_and_resize semantics) for sm_100a.
//
// Replaces lib/fpn/roi_align/src/cuda/roi_align_kernel.cu:15-80 (forward) and :103-170
// (backward) of the reference, behind the same extern "C" launchers
// (roi_align_kernel.h:11-23).  Semantics kept: one bilinear sample per output bin on a
// PH x PW grid spanning the normalised box inclusive of both ends, extrapolation value
// outside the image, rois whose batch index is out of range produce zeros.
//
// Design (HBM-bound op; see DESIGN.md "RoIAlign"):
//   * one CTA per (roi, 64-channel chunk); the per-roi sample table (4 window offsets + 2 lerp
//     weights per bin) is computed once by one warp instead of once per element with integer
//     div/mod as the reference does;
//   * NCHW: the roi's window of each feature plane is staged in shared memory transposed to
//     [pixel][channel] (row stride 33 words), so that in the compute phase the 32 lanes of a
//     warp are 32 channels of ONE bin: every shared-memory read is conflict-free and the bin's
//     table entry is a broadcast; results go through a [channel][bin] tile and leave as
//     contiguous 16-byte vector stores (the output of a chunk is one contiguous run);
//   * NHWC: lanes are channel quads, loads and stores are 16-byte vectors straight from L2;
//   * windows larger than 256 px (bins >= 1 px apart, no re-use to exploit) gather from global.
#include "common.cuh"

namespace {

constexpr int kMaxCrop = 32;     // fast kernels: crop_h, crop_w <= 32 and crop_h*crop_w <= 256
constexpr int kMaxBins = 256;
constexpr int kThreads = 256;
constexpr int kChunk   = 32;     // channels per CTA (backward kernel)

struct BinTab {
  int   o00[kMaxBins], o01[kMaxBins], o10[kMaxBins], o11[kMaxBins];
  float wx[kMaxBins], wy[kMaxBins];
  int   ok[kMaxBins];
};

struct AxisTab {
  int lo[kMaxCrop], hi[kMaxCrop], ok[kMaxCrop];
  float lerp[kMaxCrop];
};

// Sample position along one axis — the arithmetic of roi_align_kernel.cu:37-55.
__device__ __forceinline__ void axis_sample(float a1, float a2, int size, int crop, int i,
                                            int* lo, int* hi, float* lerp, int* ok) {
  // Operation order and fused multiply-adds are those nvcc emits for the reference source
  // (checked in its sm_100a SASS): in = fma(a1, size-1, i*scale); see DESIGN.md "fp contract".
  const float scale = (crop > 1) ? __fdiv_rn(__fmul_rn(a2 - a1, (float)(size - 1)), (float)(crop - 1)) : 0.f;
  const float in = (crop > 1) ? __fmaf_rn(a1, (float)(size - 1), __fmul_rn((float)i, scale))
                              : (float)(0.5 * (double)(a1 + a2) * (double)(size - 1));
  if (in < 0 || in > size - 1) { *ok = 0; *lo = 0; *hi = 0; *lerp = 0.f; return; }
  const int l = (int)floorf(in);
  const int h = (int)ceilf(in);
  *lo = l; *hi = h; *lerp = in - l; *ok = 1;
}

__device__ __forceinline__ float bilerp(float tl, float tr, float bl, float br, float wx, float wy) {
  const float top = __fmaf_rn(wx, tr - tl, tl);
  const float bottom = __fmaf_rn(wx, br - bl, bl);
  return __fmaf_rn(wy, bottom - top, top);
}

// Builds the per-roi tables. Returns (via shared memory) the window and whether any sample
// is inside the image. All threads of the CTA must call this.
struct RoiInfo { int b_in; int y_lo, x_lo, wh, ww; int any_ok; };

__device__ __forceinline__ void build_tables(const float* __restrict__ boxes, int n, int batch,
                                             int H, int W, int PH, int PW,
                                             AxisTab& ty, AxisTab& tx, RoiInfo& info) {
  const int tid = threadIdx.x;
  const float* bx = boxes + (size_t)n * 5;
  const int b_in = (int)bx[0];
  const float x1 = bx[1], y1 = bx[2], x2 = bx[3], y2 = bx[4];
  if (tid < PH) axis_sample(y1, y2, H, PH, tid, &ty.lo[tid], &ty.hi[tid], &ty.lerp[tid], &ty.ok[tid]);
  if (tid >= 32 && tid < 32 + PW) {
    const int i = tid - 32;
    axis_sample(x1, x2, W, PW, i, &tx.lo[i], &tx.hi[i], &tx.lerp[i], &tx.ok[i]);
  }
  __syncthreads();
  if (tid == 0) {
    int ylo = H, yhi = -1, xlo = W, xhi = -1;
    for (int i = 0; i < PH; ++i) if (ty.ok[i]) { ylo = min(ylo, ty.lo[i]); yhi = max(yhi, ty.hi[i]); }
    for (int i = 0; i < PW; ++i) if (tx.ok[i]) { xlo = min(xlo, tx.lo[i]); xhi = max(xhi, tx.hi[i]); }
    info.b_in = b_in;
    info.any_ok = (yhi >= 0 && xhi >= 0 && b_in >= 0 && b_in < batch);
    info.y_lo = ylo; info.x_lo = xlo;
    info.wh = yhi - ylo + 1; info.ww = xhi - xlo + 1;
  }
  __syncthreads();
}

// ---------------------------------------------------------------- forward, NCHW -> [N,C,PH,PW]
// v4 (ncu on v3: issue slots 70 % busy, 108 thread-instructions per output -> instruction bound):
// one CTA per (roi, 64-channel chunk); the window of all 64 planes is staged ONCE, transposed to
// [pixel][channel] with an odd row stride (65 words) so that
//   * staging stores (lanes = consecutive window pixels of one plane) are conflict-free,
//   * compute loads (lanes = channels c and c+32 of ONE bin) are conflict-free and the bin's
//     descriptor is a broadcast read amortised over 2 outputs per lane,
//   * results go to a [channel][bin] tile (stride bins, odd for 7x7) and leave as one contiguous
//     run of 16-byte vector stores.
constexpr int kChunkN = 64;               // channels per CTA
constexpr int kWinStride = kChunkN + 1;   // 65 words
constexpr int kStageMaxArea = 196;        // windows up to 14x14 px are staged (boxes up to ~200 px)

struct __align__(16) BinDesc { int o00, o01, o10, o11; };   // window offsets (pixel * kWinStride) or plane offsets
struct __align__(8) BinW { float wx, wy; };
struct RoiHead { int b_in, y_lo, x_lo, wh, ww, any_ok, bad_batch; };

// Warp 0 builds the axis tables and the window.
__device__ __forceinline__ void roi_preamble(const float* __restrict__ boxes, int n, int batch, int H, int W,
                                             int PH, int PW, AxisTab& ty, AxisTab& tx, RoiHead& hd) {
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    const float* bx = boxes + (size_t)n * 5;
    const int b_in = (int)bx[0];
    const float x1 = bx[1], y1 = bx[2], x2 = bx[3], y2 = bx[4];
    int ylo = H, yhi = -1, xlo = W, xhi = -1;
    if (lane < PH) {
      int lo, hi, ok; float lerp;
      axis_sample(y1, y2, H, PH, lane, &lo, &hi, &lerp, &ok);
      ty.lo[lane] = lo; ty.hi[lane] = hi; ty.lerp[lane] = lerp; ty.ok[lane] = ok;
      if (ok) { ylo = lo; yhi = hi; }
    }
    if (lane < PW) {
      int lo, hi, ok; float lerp;
      axis_sample(x1, x2, W, PW, lane, &lo, &hi, &lerp, &ok);
      tx.lo[lane] = lo; tx.hi[lane] = hi; tx.lerp[lane] = lerp; tx.ok[lane] = ok;
      if (ok) { xlo = lo; xhi = hi; }
    }
    ylo = __reduce_min_sync(0xffffffffu, ylo); yhi = __reduce_max_sync(0xffffffffu, yhi);
    xlo = __reduce_min_sync(0xffffffffu, xlo); xhi = __reduce_max_sync(0xffffffffu, xhi);
    if (lane == 0) {
      hd.b_in = b_in;
      hd.bad_batch = (b_in < 0 || b_in >= batch);
      hd.any_ok = (yhi >= 0 && xhi >= 0 && !hd.bad_batch);
      hd.y_lo = ylo; hd.x_lo = xlo; hd.wh = yhi - ylo + 1; hd.ww = xhi - xlo + 1;
    }
  }
  __syncthreads();
}

// v5 (ncu on v3/v4: issue-bound, ~50-100 thread-instructions per output, most of them in the
// register-staged window copy and in shared-memory table reads): the window is copied with
// cp.async (one instruction per element, no register round trip) into [channel][pixel]; in the
// compute phase a lane IS a bin — its 4 window offsets and 2 weights live in registers for the whole
// CTA — and loops over channels: 4 LDS + 3 FMA + 1 coalesced STG per output, no output tile.
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}

__global__ void __launch_bounds__(kThreads)
roi_align_fwd_nchw_kernel(const float* __restrict__ feat, const float* __restrict__ boxes,
                          int num_boxes, int batch, int H, int W, int PH, int PW, int C,
                          float extrap, float* __restrict__ out) {
  extern __shared__ __align__(16) float s_win[];   // [64][area]
  __shared__ AxisTab ty, tx;
  __shared__ RoiHead hd;
  __shared__ int s_goff[kStageMaxArea];
  const int bins = PH * PW;
  const int n = blockIdx.x;
  const int c0 = blockIdx.y * kChunkN;
  const int nc = min(kChunkN, C - c0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kThreads / 32;

  roi_preamble(boxes, n, batch, H, W, PH, PW, ty, tx, hd);

  float* o = out + ((size_t)n * C + c0) * bins;
  const int total = nc * bins;
  if (!hd.any_ok) {
    const float v = hd.bad_batch ? 0.f : extrap;
    for (int i = tid; i < total; i += kThreads) o[i] = v;
    return;
  }
  const int area = hd.wh * hd.ww;
  const bool staged = area <= kStageMaxArea;
  const int ww = hd.ww;
  const size_t HW = (size_t)H * W;
  const float* plane0 = feat + ((size_t)hd.b_in * C + c0) * HW;

  // bin slot of this thread: slots = bins rounded up to a warp multiple, groups = 256 / slots
  const int slots = (bins + 31) & ~31;
  const int groups = kThreads / slots;           // >= 1 because bins <= 256
  const int grp = tid / slots, b = tid - grp * slots;
  const bool has_bin = (grp < groups) && (b < bins);
  int o00 = 0, o01 = 0, o10 = 0, o11 = 0, okb = 0;
  float wx = 0.f, wy = 0.f;
  if (has_bin) {
    const int y = b / PW, x = b - y * PW;
    okb = ty.ok[y] & tx.ok[x];
    wx = tx.lerp[x]; wy = ty.lerp[y];
    if (okb) {
      if (staged) {
        const int yt = ty.lo[y] - hd.y_lo, yb = ty.hi[y] - hd.y_lo;
        const int xl = tx.lo[x] - hd.x_lo, xr = tx.hi[x] - hd.x_lo;
        o00 = yt * ww + xl; o01 = yt * ww + xr; o10 = yb * ww + xl; o11 = yb * ww + xr;
      } else {
        o00 = ty.lo[y] * W + tx.lo[x]; o01 = ty.lo[y] * W + tx.hi[x];
        o10 = ty.hi[y] * W + tx.lo[x]; o11 = ty.hi[y] * W + tx.hi[x];
      }
    }
  }
  if (!staged) {
    // large window (bins >= 1 px apart, nothing to re-use): gather straight from global / L1
    if (has_bin)
      for (int c = grp; c < nc; c += groups) {
        const float* p = plane0 + (size_t)c * HW;
        float v = extrap;
        if (okb) v = bilerp(__ldg(p + o00), __ldg(p + o01), __ldg(p + o10), __ldg(p + o11), wx, wy);
        o[(size_t)c * bins + b] = v;
      }
    return;
  }
  for (int r = tid; r < area; r += kThreads) {
    const int wy_ = r / ww, wx_ = r - wy_ * ww;
    s_goff[r] = (hd.y_lo + wy_) * W + hd.x_lo + wx_;
  }
  __syncthreads();
  // ---- stage with cp.async: warp -> plane, lanes -> window pixels
  for (int r = lane; r < area; r += 32) {
    const int g = s_goff[r];
    float* d = s_win + r;
    const float* p = plane0 + g;
#pragma unroll 4
    for (int c = warp; c < nc; c += kWarps) cp_async4(d + c * area, p + (size_t)c * HW);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  // ---- compute: lane = bin (descriptor in registers), loop over channels; coalesced stores
  if (has_bin) {
    const float* w = s_win + (size_t)grp * area;
    float* op = o + (size_t)grp * bins + b;
    const int cstep = groups * area, ostep = groups * bins;
    if (okb) {
#pragma unroll 4
      for (int c = grp; c < nc; c += groups) {
        *op = bilerp(w[o00], w[o01], w[o10], w[o11], wx, wy);
        w += cstep; op += ostep;
      }
    } else {
      for (int c = grp; c < nc; c += groups) { *op = extrap; op += ostep; }
    }
  }
}

// ---------------------------------------------------------------- forward, NHWC -> [N, PH*PW, C]
// Pipeline variant: the backbone epilogue leaves conv5_3 as NHWC fp32; pooled features come out
// bin-major / channel-minor, the K order the fc6 tensor-core GEMM consumes. One CTA per
// (roi, 128-channel chunk): a warp owns one bin at a time, each lane 4 consecutive channels.
constexpr int kChunkNHWC = 256;    // channels per CTA: each lane owns two groups of 4 channels (8 x 16-byte loads per bin)
constexpr int kNhwcStageArea = 96;       // windows up to 96 px are staged: 96 * 512 B = 48 KB
constexpr bool kNhwcUseStaging = false;  // measured on B200: L1 already serves the re-reads (39 us vs 49 us staged)
// CHW == false: out [N, bins, C] (bin-major).  CHW == true: out [N, C, bins] — the reference's
// [N,C,PH,PW] layout — produced through a shared [128][bins] tile and contiguous vector stores.
// ncu on the first version: issue-bound (34 thread-instructions per output, 0.74 issue slots/cycle):
// per-bin index math is now done ONCE per CTA into a shared descriptor table (two 16-byte broadcast
// reads per bin), offsets are 32-bit element offsets, and the bin loop is unrolled by two so eight
// 16-byte loads are in flight per lane.
struct __align__(16) NhwcBin { int o00, o01, o10, o11; };           // element offsets of the 4 corners
struct __align__(16) NhwcBinW { float wx, wy; int ok; int pad; };

template <bool CHW>
__global__ void __launch_bounds__(kThreads)
roi_align_fwd_nhwc_kernel(const float* __restrict__ feat, const float* __restrict__ boxes,
                          int num_boxes, int batch, int H, int W, int PH, int PW, int C,
                          float extrap, float* __restrict__ out) {
  extern __shared__ __align__(16) float s_tile[];          // CHW only: [kChunkNHWC][bins]
  __shared__ AxisTab ty, tx;
  __shared__ RoiHead hd;
  __shared__ NhwcBin s_bin[kMaxBins];
  __shared__ NhwcBinW s_binw[kMaxBins];
  const int n = blockIdx.x;
  const int c0 = blockIdx.y * kChunkNHWC;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kThreads / 32;
  constexpr int kQ = kChunkNHWC / 128;                      // float4 groups per lane (128 channels each)
  const int bins = PH * PW;
  roi_preamble(boxes, n, batch, H, W, PH, PW, ty, tx, hd);
  for (int b = tid; b < bins; b += kThreads) {
    const int y = b / PW, x = b - y * PW;
    NhwcBin d; NhwcBinW w;
    w.ok = hd.bad_batch ? 2 : ((ty.ok[y] & tx.ok[x]) ? 1 : 0);
    w.wx = tx.lerp[x]; w.wy = ty.lerp[y]; w.pad = 0;
    d.o00 = (ty.lo[y] * W + tx.lo[x]) * C; d.o01 = (ty.lo[y] * W + tx.hi[x]) * C;
    d.o10 = (ty.hi[y] * W + tx.lo[x]) * C; d.o11 = (ty.hi[y] * W + tx.hi[x]) * C;
    if (w.ok != 1) { d.o00 = d.o01 = d.o10 = d.o11 = 0; }
    s_bin[b] = d; s_binw[b] = w;
  }
  __syncthreads();
  const float* img0 = feat + (size_t)(hd.bad_batch ? 0 : hd.b_in) * H * W * C;
  float* o0 = out + (size_t)n * bins * C;

  // a warp owns whole bins; two bins are processed per iteration so that 2 x kQ x 4 = 16 independent
  // 16-byte loads are in flight per lane (the CTA's duration, and with it the wave quantisation of
  // short launches, is set by the number of dependent L2 round trips).
  auto load_bin = [&](int b, float4 (&tl)[kQ], float4 (&tr)[kQ], float4 (&bl)[kQ], float4 (&br)[kQ]) {
    const NhwcBin d = s_bin[b];
#pragma unroll
    for (int q = 0; q < kQ; ++q) {
      const int c = c0 + q * 128 + 4 * lane;
      const float* img = img0 + (c < C ? c : 0);
      tl[q] = __ldg((const float4*)(img + d.o00)); tr[q] = __ldg((const float4*)(img + d.o01));
      bl[q] = __ldg((const float4*)(img + d.o10)); br[q] = __ldg((const float4*)(img + d.o11));
    }
  };
  auto finish_bin = [&](int b, const float4 (&tl)[kQ], const float4 (&tr)[kQ], const float4 (&bl)[kQ], const float4 (&br)[kQ]) {
    const NhwcBinW w = s_binw[b];
#pragma unroll
    for (int q = 0; q < kQ; ++q) {
      const int c = c0 + q * 128 + 4 * lane;
      float4 v;
      v.x = bilerp(tl[q].x, tr[q].x, bl[q].x, br[q].x, w.wx, w.wy); v.y = bilerp(tl[q].y, tr[q].y, bl[q].y, br[q].y, w.wx, w.wy);
      v.z = bilerp(tl[q].z, tr[q].z, bl[q].z, br[q].z, w.wx, w.wy); v.w = bilerp(tl[q].w, tr[q].w, bl[q].w, br[q].w, w.wx, w.wy);
      if (w.ok != 1) { const float e = (w.ok == 2) ? 0.f : extrap; v = make_float4(e, e, e, e); }
      if (CHW) {
        float* t = s_tile + (size_t)(q * 128 + 4 * lane) * bins + b;
        t[0] = v.x; t[bins] = v.y; t[2 * bins] = v.z; t[3 * bins] = v.w;
      } else if (c < C) {
        *(float4*)(o0 + (size_t)b * C + c) = v;
      }
    }
  };
  int b = warp;
  for (; b + kWarps < bins; b += 2 * kWarps) {
    float4 tl0[kQ], tr0[kQ], bl0[kQ], br0[kQ], tl1[kQ], tr1[kQ], bl1[kQ], br1[kQ];
    load_bin(b, tl0, tr0, bl0, br0);
    load_bin(b + kWarps, tl1, tr1, bl1, br1);
    finish_bin(b, tl0, tr0, bl0, br0);
    finish_bin(b + kWarps, tl1, tr1, bl1, br1);
  }
  if (b < bins) {
    float4 tl0[kQ], tr0[kQ], bl0[kQ], br0[kQ];
    load_bin(b, tl0, tr0, bl0, br0);
    finish_bin(b, tl0, tr0, bl0, br0);
  }
  if (CHW) {   // the chunk's [nc][bins] block is one contiguous run in global memory
    __syncthreads();
    const int nc = min(kChunkNHWC, C - c0);
    float* dst = out + ((size_t)n * C + c0) * bins;
    const int total = nc * bins;
    if ((total % 4 == 0) && ((((uintptr_t)dst) & 15) == 0)) {
      for (int i = tid; i < total / 4; i += kThreads) ((float4*)dst)[i] = ((const float4*)s_tile)[i];
    } else {
      for (int i = tid; i < total; i += kThreads) dst[i] = s_tile[i];
    }
  }
}

// Tried and removed (round 2, profiles/r02_microbench_roi_{perbin,cols}.json): a SEPARABLE variant — a warp owns one bin
// column, loads the two pixels of every distinct feature row once (2 x rows instead of 4 x PH loads) and forms the
// horizontal lerps once, bit-identical results. ncu had shown the per-bin kernel bound by L1 load wavefronts (411 MB of
// corner loads for 103 MB of output), so 35-60 % fewer loads looked like the lever; measured it was 2.3x SLOWER (78.8 vs
// 34.8 us at N=1024): 80-128 registers for the row cache leave 14 warps per SM, and the dependent emission loop after the
// load batch exposes latency the per-bin kernel hides with 24 warps x 16 independent loads.
// Also tried and removed: a WARP-AUTONOMOUS prologue (every warp samples both axes itself and reads the bin descriptors with
// shuffles: no shared tables, no block barrier in front of the loads) — 40.0 vs 35.9 us bin-major, 49.1 vs 49.2 us in the
// [N,C,7,7] layout (gpurun call of 2026-09-23, tools/run_roi.py): the prologue is not what bounds the kernel either.

// scalar NHWC fallback for channel counts that are not a multiple of 4
__global__ void __launch_bounds__(kThreads)
roi_align_fwd_nhwc_scalar_kernel(const float* __restrict__ feat, const float* __restrict__ boxes,
                                 int num_boxes, int batch, int H, int W, int PH, int PW, int C,
                                 float extrap, float* __restrict__ out) {
  __shared__ AxisTab ty, tx;
  __shared__ RoiHead hd;
  const int n = blockIdx.x;
  const int bins = PH * PW;
  roi_preamble(boxes, n, batch, H, W, PH, PW, ty, tx, hd);
  float* o = out + (size_t)n * bins * C;
  const float* img = feat + (size_t)(hd.bad_batch ? 0 : hd.b_in) * H * W * C;
  for (int i = threadIdx.x; i < bins * C; i += kThreads) {
    const int b = i / C, c = i - b * C;
    const int y = b / PW, x = b - y * PW;
    float v;
    if (hd.bad_batch) v = 0.f;
    else if (!(ty.ok[y] & tx.ok[x])) v = extrap;
    else {
      const float* r0 = img + (size_t)(ty.lo[y] * W) * C + c;
      const float* r1 = img + (size_t)(ty.hi[y] * W) * C + c;
      v = bilerp(__ldg(r0 + (size_t)tx.lo[x] * C), __ldg(r0 + (size_t)tx.hi[x] * C),
                 __ldg(r1 + (size_t)tx.lo[x] * C), __ldg(r1 + (size_t)tx.hi[x] * C), tx.lerp[x], ty.lerp[y]);
    }
    o[i] = v;
  }
}

// ---------------------------------------------------------------- backward, NCHW
// grads [N,C,PH,PW] -> grads_image [B,C,H,W] (accumulated with atomics, as
// roi_align_kernel.cu:157-168; the caller pre-zeroes grads_image).
__global__ void __launch_bounds__(kThreads)
roi_align_bwd_nchw_kernel(const float* __restrict__ grads, const float* __restrict__ boxes,
                          int num_boxes, int batch, int H, int W, int PH, int PW, int C,
                          float* __restrict__ gimg) {
  __shared__ AxisTab ty, tx;
  __shared__ BinTab tb;
  __shared__ RoiInfo info;
  const int n = blockIdx.x;
  const int c0 = blockIdx.y * kChunk;
  const int nc = min(kChunk, C - c0);
  const int tid = threadIdx.x;
  const int bins = PH * PW;
  build_tables(boxes, n, batch, H, W, PH, PW, ty, tx, info);
  if (!info.any_ok) return;
  // Tried and removed (round 2): accumulating the roi's window in shared memory first (separable transpose, no shared
  // atomics) and issuing one global atomic per window pixel instead of four per bin — fewer atomics (30-100 vs 196 per
  // channel) but twice as slow (587 vs 285 us at N=1024, 4.5 vs 2.9 ms at N=8192): the two extra block-wide phases and
  // their index arithmetic cost more than the L2 atomics they save.
  for (int b = tid; b < bins; b += kThreads) {
    const int y = b / PW, x = b - y * PW;
    tb.ok[b] = ty.ok[y] & tx.ok[x];
    tb.wx[b] = tx.lerp[x]; tb.wy[b] = ty.lerp[y];
    tb.o00[b] = ty.lo[y] * W + tx.lo[x]; tb.o01[b] = ty.lo[y] * W + tx.hi[x];
    tb.o10[b] = ty.hi[y] * W + tx.lo[x]; tb.o11[b] = ty.hi[y] * W + tx.hi[x];
  }
  __syncthreads();
  const float* g = grads + ((size_t)n * C + c0) * bins;
  float* plane0 = gimg + ((size_t)info.b_in * C + c0) * H * W;
  const int total = nc * bins;
  for (int i = tid; i < total; i += kThreads) {
    const int c = i / bins, b = i - c * bins;
    if (!tb.ok[b]) continue;
    float* p = plane0 + (size_t)c * H * W;
    const float go = g[i];
    const float wx = tb.wx[b], wy = tb.wy[b];
    const float dtop = (1 - wy) * go;
    atomicAdd(p + tb.o00[b], (1 - wx) * dtop);
    atomicAdd(p + tb.o01[b], wx * dtop);
    const float dbottom = wy * go;
    atomicAdd(p + tb.o10[b], (1 - wx) * dbottom);
    atomicAdd(p + tb.o11[b], wx * dbottom);
  }
}

// ---------------------------------------------------------------- generic fallback (large crops)
__global__ void roi_align_fwd_generic_kernel(const long long nthreads, const float* __restrict__ feat,
                                             const float* __restrict__ boxes, int batch, int H, int W,
                                             int PH, int PW, int C, float extrap, float* __restrict__ out) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < nthreads;
       idx += (long long)blockDim.x * gridDim.x) {
    long long r = idx;
    const int x = r % PW; r /= PW;
    const int y = r % PH; r /= PH;
    const int d = r % C;
    const int n = r / C;
    const float* bx = boxes + (size_t)n * 5;
    const int b_in = (int)bx[0];
    if (b_in < 0 || b_in >= batch) { out[idx] = 0.f; continue; }
    int yt, yb, xl, xr, oky, okx; float wy, wx;
    axis_sample(bx[2], bx[4], H, PH, y, &yt, &yb, &wy, &oky);
    axis_sample(bx[1], bx[3], W, PW, x, &xl, &xr, &wx, &okx);
    if (!(oky && okx)) { out[idx] = extrap; continue; }
    const float* p = feat + ((size_t)b_in * C + d) * H * W;
    out[idx] = bilerp(__ldg(p + yt * W + xl), __ldg(p + yt * W + xr), __ldg(p + yb * W + xl),
                      __ldg(p + yb * W + xr), wx, wy);
  }
}

__global__ void roi_align_bwd_generic_kernel(const long long nthreads, const float* __restrict__ grads,
                                             const float* __restrict__ boxes, int batch, int H, int W,
                                             int PH, int PW, int C, float* __restrict__ gimg) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < nthreads;
       idx += (long long)blockDim.x * gridDim.x) {
    long long r = idx;
    const int x = r % PW; r /= PW;
    const int y = r % PH; r /= PH;
    const int d = r % C;
    const int n = r / C;
    const float* bx = boxes + (size_t)n * 5;
    const int b_in = (int)bx[0];
    if (b_in < 0 || b_in >= batch) continue;
    int yt, yb, xl, xr, oky, okx; float wy, wx;
    axis_sample(bx[2], bx[4], H, PH, y, &yt, &yb, &wy, &oky);
    axis_sample(bx[1], bx[3], W, PW, x, &xl, &xr, &wx, &okx);
    if (!(oky && okx)) continue;
    float* p = gimg + ((size_t)b_in * C + d) * H * W;
    const float go = grads[idx];
    const float dtop = (1 - wy) * go;
    atomicAdd(p + yt * W + xl, (1 - wx) * dtop);
    atomicAdd(p + yt * W + xr, wx * dtop);
    const float dbottom = wy * go;
    atomicAdd(p + yb * W + xl, (1 - wx) * dbottom);
    atomicAdd(p + yb * W + xr, wx * dbottom);
  }
}

}  // namespace

extern "C" {

// Drop-in for roi_align_kernel.h:11-15 (same name, argument order and return value).
int ROIAlignForwardLaucher(const float* image_ptr, const float* boxes_ptr, int num_boxes, int batch,
                           int image_height, int image_width, int crop_height, int crop_width,
                           int depth, float extrapolation_value, float* crops_ptr, cudaStream_t stream) {
  if (num_boxes <= 0 || depth <= 0) return MB200_OK;
  if (crop_height <= 0 || crop_width <= 0 || image_height <= 0 || image_width <= 0) return MB200_ERR_ARG;
  const int bins = crop_height * crop_width;
  const int chunks = mb200_div_up(depth, kChunkN);
  if (crop_height <= kMaxCrop && crop_width <= kMaxCrop && bins <= kMaxBins && chunks <= 65535) {
    dim3 grid(num_boxes, chunks);
    const size_t smem = (size_t)kChunkN * kStageMaxArea * sizeof(float);   // 49 KB
    static bool attr_set = false;
    if (!attr_set) {
      MB200_CHECK(cudaFuncSetAttribute(roi_align_fwd_nchw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_set = true;
    }
    roi_align_fwd_nchw_kernel<<<grid, kThreads, smem, stream>>>(
        image_ptr, boxes_ptr, num_boxes, batch, image_height, image_width, crop_height, crop_width,
        depth, extrapolation_value, crops_ptr);
  } else {
    const long long total = (long long)num_boxes * depth * bins;
    const int blocks = (int)min((long long)kNumSMs * 16, (total + 255) / 256);
    roi_align_fwd_generic_kernel<<<blocks, 256, 0, stream>>>(total, image_ptr, boxes_ptr, batch,
        image_height, image_width, crop_height, crop_width, depth, extrapolation_value, crops_ptr);
  }
  MB200_CHECK_LAUNCH("ROIAlignForwardLaucher");
  return MB200_OK;
}

// Drop-in for roi_align_kernel.h:21-23. grads_image must be pre-zeroed by the caller
// (functions/roi_align.py:66-67 does so); gradients are accumulated into it.
int ROIAlignBackwardLaucher(const float* grads_ptr, const float* boxes_ptr, int num_boxes, int batch,
                            int image_height, int image_width, int crop_height, int crop_width,
                            int depth, float* grads_image_ptr, cudaStream_t stream) {
  if (num_boxes <= 0 || depth <= 0) return MB200_OK;
  if (crop_height <= 0 || crop_width <= 0 || image_height <= 0 || image_width <= 0) return MB200_ERR_ARG;
  const int bins = crop_height * crop_width;
  const int chunks = mb200_div_up(depth, kChunk);
  if (crop_height <= kMaxCrop && crop_width <= kMaxCrop && bins <= kMaxBins && chunks <= 65535) {
    dim3 grid(num_boxes, chunks);
    roi_align_bwd_nchw_kernel<<<grid, kThreads, 0, stream>>>(grads_ptr, boxes_ptr, num_boxes, batch,
        image_height, image_width, crop_height, crop_width, depth, grads_image_ptr);
  } else {
    const long long total = (long long)num_boxes * depth * bins;
    const int blocks = (int)min((long long)kNumSMs * 16, (total + 255) / 256);
    roi_align_bwd_generic_kernel<<<blocks, 256, 0, stream>>>(total, grads_ptr, boxes_ptr, batch,
        image_height, image_width, crop_height, crop_width, depth, grads_image_ptr);
  }
  MB200_CHECK_LAUNCH("ROIAlignBackwardLaucher");
  return MB200_OK;
}

static int roi_align_nhwc_launch(bool chw, const float* image_nhwc, const float* boxes_ptr, int num_boxes,
                                 int batch, int image_height, int image_width, int crop_height, int crop_width,
                                 int depth, float extrapolation_value, float* crops, cudaStream_t stream) {
  if (num_boxes <= 0 || depth <= 0) return MB200_OK;
  if (crop_height <= 0 || crop_width <= 0 || crop_height > kMaxCrop || crop_width > kMaxCrop ||
      crop_height * crop_width > kMaxBins)
    return MB200_ERR_ARG;
  const int bins = crop_height * crop_width;
  if (depth % 4 == 0 && ((((uintptr_t)image_nhwc) | ((uintptr_t)crops)) & 15) == 0) {
    dim3 grid(num_boxes, mb200_div_up(depth, kChunkNHWC));
    const size_t win = 0;
    const size_t max_tile = 200 * 1024;      // opt-in ceiling; larger crops take the bin-major variant
    static bool attr_set = false;
    if (!attr_set) {
      MB200_CHECK(cudaFuncSetAttribute(roi_align_fwd_nhwc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)win));
      MB200_CHECK(cudaFuncSetAttribute(roi_align_fwd_nhwc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(win + max_tile)));
      attr_set = true;
    }
    if (chw && (size_t)kChunkNHWC * bins * sizeof(float) > max_tile) return MB200_ERR_UNSUPPORTED;
    if (chw)
      roi_align_fwd_nhwc_kernel<true><<<grid, kThreads, win + (size_t)kChunkNHWC * bins * sizeof(float), stream>>>(
          image_nhwc, boxes_ptr, num_boxes, batch, image_height, image_width, crop_height, crop_width, depth,
          extrapolation_value, crops);
    else
      roi_align_fwd_nhwc_kernel<false><<<grid, kThreads, win, stream>>>(
          image_nhwc, boxes_ptr, num_boxes, batch, image_height, image_width, crop_height, crop_width, depth,
          extrapolation_value, crops);
  } else {
    if (chw) return MB200_ERR_UNSUPPORTED;
    roi_align_fwd_nhwc_scalar_kernel<<<num_boxes, kThreads, 0, stream>>>(image_nhwc, boxes_ptr, num_boxes, batch,
        image_height, image_width, crop_height, crop_width, depth, extrapolation_value, crops);
  }
  MB200_CHECK_LAUNCH("mb200_roi_align_forward_nhwc");
  return MB200_OK;
}

// Superset: NHWC feature map in, [N, crop_h*crop_w, depth] out (channels-last pooled features).
int mb200_roi_align_forward_nhwc(const float* image_nhwc, const float* boxes_ptr, int num_boxes,
                                 int batch, int image_height, int image_width, int crop_height,
                                 int crop_width, int depth, float extrapolation_value,
                                 float* crops_nhwc, cudaStream_t stream) {
  return roi_align_nhwc_launch(false, image_nhwc, boxes_ptr, num_boxes, batch, image_height, image_width,
                               crop_height, crop_width, depth, extrapolation_value, crops_nhwc, stream);
}

// Superset: NHWC feature map in, the reference's [N, depth, crop_h, crop_w] out (depth % 4 == 0).
int mb200_roi_align_forward_nhwc_to_nchw(const float* image_nhwc, const float* boxes_ptr, int num_boxes,
                                         int batch, int image_height, int image_width, int crop_height,
                                         int crop_width, int depth, float extrapolation_value,
                                         float* crops_nchw, cudaStream_t stream) {
  return roi_align_nhwc_launch(true, image_nhwc, boxes_ptr, num_boxes, batch, image_height, image_width,
                               crop_height, crop_width, depth, extrapolation_value, crops_nchw, stream);
}

}  // extern "C"
