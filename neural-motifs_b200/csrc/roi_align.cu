// RoIAlign (TF crop_and_resize semantics) for sm_100a.
//
// Replaces lib/fpn/roi_align/src/cuda/roi_align_kernel.cu:15-80 (forward) and :103-170
// (backward) of the reference, behind the same extern "C" launchers
// (roi_align_kernel.h:11-23).  Semantics kept: one bilinear sample per output bin on a
// PH x PW grid spanning the normalised box inclusive of both ends, extrapolation value
// outside the image, rois whose batch index is out of range produce zeros.
//
// Design (HBM-bound op; see DESIGN.md "RoIAlign"):
//   * one CTA per (roi, 32-channel chunk); the per-roi sample table (4 offsets + 2 lerp
//     weights per bin) is computed once into shared memory instead of once per element
//     with integer div/mod as the reference does;
//   * the roi's window of the feature plane is staged in shared memory (small boxes
//     re-read every pixel ~3x), then every output is produced from shared memory;
//   * outputs are written bin-fastest so a warp stores one contiguous 128 B line.
#include "common.cuh"

namespace {

constexpr int kMaxCrop = 32;     // fast kernels: crop_h, crop_w <= 32 and crop_h*crop_w <= 256
constexpr int kMaxBins = 256;
constexpr int kThreads = 256;
constexpr int kChunk   = 32;     // channels per CTA
constexpr int kWinFloats = 8192; // 32 KB dynamic shared memory for the staged window
constexpr int kStageMaxArea = kWinFloats / kChunk;  // 256 px: one pass covers the whole chunk

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
__global__ void __launch_bounds__(kThreads)
roi_align_fwd_nchw_kernel(const float* __restrict__ feat, const float* __restrict__ boxes,
                          int num_boxes, int batch, int H, int W, int PH, int PW, int C,
                          float extrap, float* __restrict__ out) {
  extern __shared__ float s_win[];
  __shared__ AxisTab ty, tx;
  __shared__ BinTab tb;
  __shared__ RoiInfo info;

  const int n = blockIdx.x;
  const int c0 = blockIdx.y * kChunk;
  const int nc = min(kChunk, C - c0);
  const int tid = threadIdx.x;
  const int bins = PH * PW;

  build_tables(boxes, n, batch, H, W, PH, PW, ty, tx, info);

  float* o = out + ((size_t)n * C + c0) * bins;
  const int total = nc * bins;
  if (!info.any_ok) {
    // roi outside the batch -> zeros (the reference leaves the caller's zero fill untouched);
    // every sample outside the image -> extrapolation value.
    const bool bad_batch = (info.b_in < 0 || info.b_in >= batch);
    const float v = bad_batch ? 0.f : extrap;
    for (int i = tid; i < total; i += kThreads) o[i] = v;
    return;
  }

  const int area = info.wh * info.ww;
  const bool staged = area <= kStageMaxArea;
  const int ww = info.ww;
  for (int b = tid; b < bins; b += kThreads) {
    const int y = b / PW, x = b - y * PW;
    const int ok = ty.ok[y] & tx.ok[x];
    tb.ok[b] = ok;
    tb.wx[b] = tx.lerp[x];
    tb.wy[b] = ty.lerp[y];
    if (staged) {
      const int yt = ty.lo[y] - info.y_lo, yb = ty.hi[y] - info.y_lo;
      const int xl = tx.lo[x] - info.x_lo, xr = tx.hi[x] - info.x_lo;
      tb.o00[b] = ok ? yt * ww + xl : 0; tb.o01[b] = ok ? yt * ww + xr : 0;
      tb.o10[b] = ok ? yb * ww + xl : 0; tb.o11[b] = ok ? yb * ww + xr : 0;
    } else {
      tb.o00[b] = ty.lo[y] * W + tx.lo[x]; tb.o01[b] = ty.lo[y] * W + tx.hi[x];
      tb.o10[b] = ty.hi[y] * W + tx.lo[x]; tb.o11[b] = ty.hi[y] * W + tx.hi[x];
    }
  }

  const float* plane0 = feat + ((size_t)info.b_in * C + c0) * H * W;
  if (staged) {
    // Stage the window of every channel of the chunk: s_win[c][r], r = wy*ww + wx.
    __syncthreads();
    const size_t base = (size_t)info.y_lo * W + info.x_lo;
    // Each warp walks channels; lanes walk window elements.
    const int warp = tid >> 5, lane = tid & 31, nwarps = kThreads >> 5;
    for (int c = warp; c < nc; c += nwarps) {
      const float* p = plane0 + (size_t)c * H * W + base;
      float* d = s_win + c * area;
      for (int r = lane; r < area; r += 32) {
        const int wy_ = r / ww, wx_ = r - wy_ * ww;
        d[r] = __ldg(p + wy_ * W + wx_);
      }
    }
    __syncthreads();
    for (int i = tid; i < total; i += kThreads) {
      const int c = i / bins, b = i - c * bins;
      const float* d = s_win + c * area;
      float v = extrap;
      if (tb.ok[b]) v = bilerp(d[tb.o00[b]], d[tb.o01[b]], d[tb.o10[b]], d[tb.o11[b]], tb.wx[b], tb.wy[b]);
      o[i] = v;
    }
  } else {
    __syncthreads();
    for (int i = tid; i < total; i += kThreads) {
      const int c = i / bins, b = i - c * bins;
      const float* p = plane0 + (size_t)c * H * W;
      float v = extrap;
      if (tb.ok[b])
        v = bilerp(__ldg(p + tb.o00[b]), __ldg(p + tb.o01[b]), __ldg(p + tb.o10[b]), __ldg(p + tb.o11[b]),
                   tb.wx[b], tb.wy[b]);
      o[i] = v;
    }
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

// ---------------------------------------------------------------- forward, NHWC -> [N, PH*PW, C]
// Pipeline variant: the backbone epilogue leaves conv5_3 as NHWC fp32; pooled features
// come out bin-major / channel-minor ("channels last"), which is the K order the fc6
// tensor-core GEMM consumes.  One CTA per (roi, 64-channel chunk).
constexpr int kChunkNHWC = 64;
__global__ void __launch_bounds__(kThreads)
roi_align_fwd_nhwc_kernel(const float* __restrict__ feat, const float* __restrict__ boxes,
                          int num_boxes, int batch, int H, int W, int PH, int PW, int C,
                          float extrap, float* __restrict__ out) {
  __shared__ AxisTab ty, tx;
  __shared__ RoiInfo info;
  const int n = blockIdx.x;
  const int c0 = blockIdx.y * kChunkNHWC;
  const int nc = min(kChunkNHWC, C - c0);
  const int tid = threadIdx.x;
  const int bins = PH * PW;
  build_tables(boxes, n, batch, H, W, PH, PW, ty, tx, info);
  float* o = out + (size_t)n * bins * C + c0;
  const bool bad_batch = (info.b_in < 0 || info.b_in >= batch);
  const float* img = feat + (size_t)(bad_batch ? 0 : info.b_in) * H * W * C + c0;
  // thread -> (bin group, channel): channel fastest so loads and stores are coalesced.
  const int cl = tid % kChunkNHWC;
  const int bg = tid / kChunkNHWC;               // 0..3
  const int bstep = kThreads / kChunkNHWC;
  if (cl >= nc) return;
  for (int b = bg; b < bins; b += bstep) {
    const int y = b / PW, x = b - y * PW;
    float v;
    if (bad_batch) v = 0.f;
    else if (!(ty.ok[y] & tx.ok[x])) v = extrap;
    else {
      const float* r0 = img + (size_t)(ty.lo[y] * W) * C;
      const float* r1 = img + (size_t)(ty.hi[y] * W) * C;
      const int xl = tx.lo[x] * C + cl, xr = tx.hi[x] * C + cl;
      v = bilerp(__ldg(r0 + xl), __ldg(r0 + xr), __ldg(r1 + xl), __ldg(r1 + xr), tx.lerp[x], ty.lerp[y]);
    }
    o[(size_t)b * C + cl] = v;
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
  const int chunks = mb200_div_up(depth, kChunk);
  if (crop_height <= kMaxCrop && crop_width <= kMaxCrop && bins <= kMaxBins && chunks <= 65535) {
    dim3 grid(num_boxes, chunks);
    roi_align_fwd_nchw_kernel<<<grid, kThreads, kWinFloats * sizeof(float), stream>>>(
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

// Superset: NHWC feature map in, [N, crop_h*crop_w, depth] out (channels-last pooled features).
int mb200_roi_align_forward_nhwc(const float* image_nhwc, const float* boxes_ptr, int num_boxes,
                                 int batch, int image_height, int image_width, int crop_height,
                                 int crop_width, int depth, float extrapolation_value,
                                 float* crops_nhwc, cudaStream_t stream) {
  if (num_boxes <= 0 || depth <= 0) return MB200_OK;
  if (crop_height <= 0 || crop_width <= 0 || crop_height > kMaxCrop || crop_width > kMaxCrop)
    return MB200_ERR_ARG;
  dim3 grid(num_boxes, mb200_div_up(depth, kChunkNHWC));
  roi_align_fwd_nhwc_kernel<<<grid, kThreads, 0, stream>>>(image_nhwc, boxes_ptr, num_boxes, batch,
      image_height, image_width, crop_height, crop_width, depth, extrapolation_value, crops_nhwc);
  MB200_CHECK_LAUNCH("mb200_roi_align_forward_nhwc");
  return MB200_OK;
}

}  // extern "C"
