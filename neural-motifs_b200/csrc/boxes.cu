// Box geometry kernels for sm_100a: pairwise IoU (fp32 torch formula and the float64
// Cython formula), union-box rois, and the union-box mask rasteriser.
//
// Replaces, on the device:
//   lib/fpn/box_utils.py:85-131            bbox_intersections / bbox_overlaps (torch, fp32)
//   lib/fpn/box_intersections_cpu/bbox.pyx:15-62 bbox_overlaps (Cython, float64)
//   lib/get_union_boxes.py:82-87           union roi = (min x1,y1 ; max x2,y2)
//   lib/draw_rectangles/draw_rectangles.pyx:27-67 draw_union_boxes (CPU rasteriser that the
//       reference reaches through a D2H copy + numpy + H2D copy in the middle of forward)
#include "common.cuh"

namespace {

// box_utils.py:109-131. Separate torch kernels in the reference => no fused multiply-add.
__global__ void bbox_overlaps_f32_kernel(const float4* __restrict__ a, int A, const float4* __restrict__ b,
                                         int B, float* __restrict__ out) {
  const long long total = (long long)A * B;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)blockDim.x * gridDim.x) {
    const int i = idx / B, j = idx - (long long)i * B;
    const float4 p = a[i], q = b[j];
    const float iw = fmaxf(__fadd_rn(__fadd_rn(fminf(p.z, q.z), -fmaxf(p.x, q.x)), 1.0f), 0.f);
    const float ih = fmaxf(__fadd_rn(__fadd_rn(fminf(p.w, q.w), -fmaxf(p.y, q.y)), 1.0f), 0.f);
    const float inter = __fmul_rn(iw, ih);
    const float area_a = __fmul_rn(__fadd_rn(__fadd_rn(p.z, -p.x), 1.0f), __fadd_rn(__fadd_rn(p.w, -p.y), 1.0f));
    const float area_b = __fmul_rn(__fadd_rn(__fadd_rn(q.z, -q.x), 1.0f), __fadd_rn(__fadd_rn(q.w, -q.y), 1.0f));
    const float uni = __fadd_rn(__fadd_rn(area_a, area_b), -inter);
    out[idx] = __fdiv_rn(inter, uni);
  }
}

// bbox.pyx:21-62: float64, overlaps[n,k] = iw*ih / ua when iw>0 and ih>0, else 0.
// mode 0: IoU (bbox_overlaps); mode 1: intersection / query-box area (bbox_intersections :71-107)
__global__ void bbox_overlaps_f64_kernel(const double* __restrict__ boxes, int N, const double* __restrict__ query,
                                         int K, int mode, double* __restrict__ out) {
  const long long total = (long long)N * K;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)blockDim.x * gridDim.x) {
    const int n = idx / K, k = idx - (long long)n * K;
    const double* bb = boxes + (size_t)n * 4;
    const double* qq = query + (size_t)k * 4;
    const double box_area = __dmul_rn(qq[2] - qq[0] + 1, qq[3] - qq[1] + 1);
    double r = 0.0;
    const double iw = fmin(bb[2], qq[2]) - fmax(bb[0], qq[0]) + 1;
    if (iw > 0) {
      const double ih = fmin(bb[3], qq[3]) - fmax(bb[1], qq[1]) + 1;
      if (ih > 0) {
        const double inter = __dmul_rn(iw, ih);
        if (mode == 0) {
          const double ua = __dadd_rn(__dadd_rn(__dmul_rn(bb[2] - bb[0] + 1, bb[3] - bb[1] + 1), box_area), -inter);
          r = __ddiv_rn(inter, ua);
        } else {
          r = __ddiv_rn(inter, box_area);
        }
      }
    }
    out[idx] = r;
  }
}

// get_union_boxes.py:82-87 + the [N,8] pair layout of :47.
__global__ void union_rois_kernel(const float* __restrict__ rois, const long long* __restrict__ pairs, int R,
                                  float* __restrict__ union_rois, float* __restrict__ pair_boxes) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float* a = rois + (size_t)pairs[2 * r] * 5;
  const float* b = rois + (size_t)pairs[2 * r + 1] * 5;
  float* u = union_rois + (size_t)r * 5;
  u[0] = a[0];
  u[1] = fminf(a[1], b[1]); u[2] = fminf(a[2], b[2]);
  u[3] = fmaxf(a[3], b[3]); u[4] = fmaxf(a[4], b[4]);
  if (pair_boxes) {
    float* p = pair_boxes + (size_t)r * 8;
    p[0] = a[1]; p[1] = a[2]; p[2] = a[3]; p[3] = a[4];
    p[4] = b[1]; p[5] = b[2]; p[6] = b[3]; p[7] = b[4];
  }
}

__device__ __forceinline__ float minmax01(float x) { return fminf(fmaxf(x, 0.f), 1.f); }

// draw_rectangles.pyx:45-66, float32 arithmetic in the same order:
//   x1_box = (box_x1 - x1_union) * P / w ;  contrib = minmax(k+1-x1_box) * minmax(x2_box-k)
// One CTA per pair; `offset` is subtracted from every pixel (get_union_boxes.py:49 uses 0.5).
__global__ void __launch_bounds__(256)
draw_union_boxes_kernel(const float* __restrict__ pairs, int N, int P, float offset, float* __restrict__ out) {
  const int n = blockIdx.x;
  __shared__ float s_box[2][4];
  if (threadIdx.x < 2) {
    const float* p = pairs + (size_t)n * 8;
    const float x1u = fminf(p[0], p[4]), y1u = fminf(p[1], p[5]);
    const float x2u = fmaxf(p[2], p[6]), y2u = fmaxf(p[3], p[7]);
    const float w = x2u - x1u, h = y2u - y1u;
    const int i = threadIdx.x;
    const float Pf = (float)P;
    s_box[i][0] = __fdiv_rn(__fmul_rn(p[0 + 4 * i] - x1u, Pf), w);
    s_box[i][1] = __fdiv_rn(__fmul_rn(p[1 + 4 * i] - y1u, Pf), h);
    s_box[i][2] = __fdiv_rn(__fmul_rn(p[2 + 4 * i] - x1u, Pf), w);
    s_box[i][3] = __fdiv_rn(__fmul_rn(p[3 + 4 * i] - y1u, Pf), h);
  }
  __syncthreads();
  const int total = 2 * P * P;
  float* o = out + (size_t)n * total;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int i = idx / (P * P);
    const int r = idx - i * P * P;
    const int j = r / P, k = r - j * P;
    const float yc = __fmul_rn(minmax01((float)(j + 1) - s_box[i][1]), minmax01(s_box[i][3] - (float)j));
    const float xc = __fmul_rn(minmax01((float)(k + 1) - s_box[i][0]), minmax01(s_box[i][2] - (float)k));
    o[idx] = __fmul_rn(xc, yc) - offset;
  }
}

// box_utils.py:28-48 bbox_preds (delta decode with the +1 pixel convention of center_size /
// point_form :51-82), optionally clamped to [0, w-1] x [0, h-1] of the roi's image
// (object_detector.py:383-387, :586-590). boxes [N,4]; deltas [N*K,4]; out [N*K,4];
// prior of row r is boxes[r / K]. im_hw: per-row-image (h,w) float pairs or NULL; im_idx [N] or NULL.
__global__ void bbox_preds_kernel(const float4* __restrict__ boxes, const float4* __restrict__ deltas,
                                  long long total, int K, const float* __restrict__ im_hw,
                                  const int* __restrict__ im_idx, float4* __restrict__ out) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < total;
       r += (long long)blockDim.x * gridDim.x) {
    const long long n = r / K;
    const float4 b = boxes[n];
    const float4 d = deltas[r];
    // center_size: wh = b[2:] - b[:2] + 1 ; c = b[:2] + 0.5*wh
    const float w = __fadd_rn(__fadd_rn(b.z, -b.x), 1.0f), h = __fadd_rn(__fadd_rn(b.w, -b.y), 1.0f);
    const float cx = __fadd_rn(b.x, __fmul_rn(0.5f, w)), cy = __fadd_rn(b.y, __fmul_rn(0.5f, h));
    const float nx = __fadd_rn(cx, __fmul_rn(w, d.x)), ny = __fadd_rn(cy, __fmul_rn(h, d.y));
    const float nw = __fmul_rn(expf(d.z), w), nh = __fmul_rn(expf(d.w), h);
    // point_form: (c - 0.5*wh, c + 0.5*(wh - 2))
    float4 o;
    o.x = __fadd_rn(nx, -__fmul_rn(0.5f, nw)); o.y = __fadd_rn(ny, -__fmul_rn(0.5f, nh));
    o.z = __fadd_rn(nx, __fmul_rn(0.5f, __fadd_rn(nw, -2.0f))); o.w = __fadd_rn(ny, __fmul_rn(0.5f, __fadd_rn(nh, -2.0f)));
    if (im_hw) {
      const int im = im_idx ? im_idx[n] : 0;
      const float hh = im_hw[2 * im] - 1.f, ww = im_hw[2 * im + 1] - 1.f;
      o.x = fminf(fmaxf(o.x, 0.f), ww); o.z = fminf(fmaxf(o.z, 0.f), ww);
      o.y = fminf(fmaxf(o.y, 0.f), hh); o.w = fminf(fmaxf(o.w, 0.f), hh);
    }
    out[r] = o;
  }
}

inline int grid_for(long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  const long long cap = (long long)kNumSMs * 32;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}


// ------------------------------------------------------------------ RPN anchor-target assignment
// lib/fpn/anchor_targets.py:50-67 of the reference (numpy in the DataLoader worker): IoU of every inside-image
// anchor with every GT box in float64 (bbox.pyx:21-62), per-anchor max / FIRST arg-max (np.argmax), per-GT-box
// max, then labels: 0 where max < neg_thr, 1 where the anchor attains some GT box's maximum (`overlaps ==
// gt_max_overlaps`, which also fires for a GT box that no anchor overlaps: every anchor then "attains" 0),
// 1 where max >= pos_thr; -1 otherwise. One WARP per anchor, lanes stride over the GT boxes; the [N,G] IoU
// matrix is never written (27 380 x G doubles in the reference). Pass 1: shuffle reduce (value, then lower
// index) + the column maxima through 64-bit atomicMax on the bit pattern (IoU >= 0, so the unsigned order of
// the bits is the order of the values). Pass 2 re-derives each IoU bit-identically and votes with __any_sync.
__device__ __forceinline__ double iou_f64(const double* __restrict__ bb, const double* __restrict__ qq) {
  const double iw = fmin(bb[2], qq[2]) - fmax(bb[0], qq[0]) + 1;
  if (!(iw > 0)) return 0.0;
  const double ih = fmin(bb[3], qq[3]) - fmax(bb[1], qq[1]) + 1;
  if (!(ih > 0)) return 0.0;
  const double box_area = __dmul_rn(qq[2] - qq[0] + 1, qq[3] - qq[1] + 1);
  const double inter = __dmul_rn(iw, ih);
  const double ua = __dadd_rn(__dadd_rn(__dmul_rn(bb[2] - bb[0] + 1, bb[3] - bb[1] + 1), box_area), -inter);
  return __ddiv_rn(inter, ua);
}

__global__ void anchor_rowmax_kernel(const double* __restrict__ anchors, int N, const double* __restrict__ gt, int G,
                                     double* __restrict__ max_ov, int* __restrict__ argmax,
                                     unsigned long long* __restrict__ gt_max_bits) {
  const int lane = threadIdx.x & 31;
  const int warps = (blockDim.x >> 5) * gridDim.x;
  for (int a = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); a < N; a += warps) {
    const double* bb = anchors + (size_t)a * 4;
    double best = -1.0; int arg = 0x7fffffff;
    for (int g = lane; g < G; g += 32) {
      const double v = iou_f64(bb, gt + (size_t)g * 4);
      if (v > best) { best = v; arg = g; }                      // strict: the first maximum of this lane's stride
      if (v > 0.0) atomicMax(gt_max_bits + g, (unsigned long long)__double_as_longlong(v));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
    }
    if (lane == 0) { max_ov[a] = best; argmax[a] = arg; }
  }
}

__global__ void anchor_label_kernel(const double* __restrict__ anchors, int N, const double* __restrict__ gt, int G,
                                    const double* __restrict__ max_ov, const unsigned long long* __restrict__ gt_max_bits,
                                    double neg_thr, double pos_thr, long long* __restrict__ labels) {
  const int lane = threadIdx.x & 31;
  const int warps = (blockDim.x >> 5) * gridDim.x;
  for (int a = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); a < N; a += warps) {
    const double* bb = anchors + (size_t)a * 4;
    bool hit = false;
    for (int g = lane; g < G; g += 32)
      hit |= (iou_f64(bb, gt + (size_t)g * 4) == __longlong_as_double((long long)gt_max_bits[g]));
    hit = __any_sync(0xffffffffu, hit);
    if (lane == 0) {
      const double m = max_ov[a];
      long long l = -1;
      if (m < neg_thr) l = 0;
      if (hit) l = 1;
      if (m >= pos_thr) l = 1;
      labels[a] = l;
    }
  }
}


// ------------------------------------------------------------------ decoder: overlap-aware greedy commitment
// lib/lstm/decoder_rnn.py:230-247 of the reference — a host numpy loop over the detections there (D2H of the [N,N,C]
// overlap tensor and of the class probabilities, then N iterations of argmax / suppress). One CTA here: the [N,C]
// probabilities live in shared memory; every iteration takes the global arg-max (np.argmax order: the FIRST maximum in
// row-major order), commits that (box, class), zeroes the class for every box whose class-specific box overlaps the
// winner's by IoU >= thresh (nms_overlaps, lib/fpn/box_utils.py:134-154: separate fp32 ops, no fused multiply-add),
// and retires the winner's row with -1. boxes [N,C,4], probs [N,C] (softmax; column 0 is cleared here), out [N] int64.
__global__ void __launch_bounds__(512) decoder_commit_kernel(const float* __restrict__ boxes, const float* __restrict__ probs,
                                                             int N, int C, float thresh, long long* __restrict__ commit) {
  extern __shared__ float s_p[];                 // [N*C]
  __shared__ float s_val[16];
  __shared__ int s_idx[16];
  __shared__ int s_win;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int total = N * C;
  for (int i = tid; i < total; i += blockDim.x) s_p[i] = (i % C == 0) ? 0.f : probs[i];
  for (int i = tid; i < N; i += blockDim.x) commit[i] = 0;
  __syncthreads();
  for (int it = 0; it < N; ++it) {
    float best = -INFINITY; int bi = 0x7fffffff;
    for (int i = tid; i < total; i += blockDim.x) {
      const float v = s_p[i];
      if (v > best) { best = v; bi = i; }        // ascending i per thread: keeps this thread's first maximum
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) { s_val[warp] = best; s_idx[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
      best = lane < (blockDim.x >> 5) ? s_val[lane] : -INFINITY;
      bi = lane < (blockDim.x >> 5) ? s_idx[lane] : 0x7fffffff;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
      }
      if (lane == 0) s_win = bi;
    }
    __syncthreads();
    const int b = s_win / C, cls = s_win - b * C;
    if (tid == 0) commit[b] = cls;
    const float* wb = boxes + ((size_t)b * C + cls) * 4;
    const float wx1 = wb[0], wy1 = wb[1], wx2 = wb[2], wy2 = wb[3];
    const float warea = __fmul_rn(__fadd_rn(__fadd_rn(wx2, -wx1), 1.0f), __fadd_rn(__fadd_rn(wy2, -wy1), 1.0f));
    for (int j = tid; j < N; j += blockDim.x) {
      const float* q = boxes + ((size_t)j * C + cls) * 4;
      // is_overlap[b, j, cls]: inter from min/max of the corners, union = (-inter + area[j]) + area[b]
      const float iw = fmaxf(__fadd_rn(__fadd_rn(fminf(wx2, q[2]), -fmaxf(wx1, q[0])), 1.0f), 0.f);
      const float ih = fmaxf(__fadd_rn(__fadd_rn(fminf(wy2, q[3]), -fmaxf(wy1, q[1])), 1.0f), 0.f);
      const float inter = __fmul_rn(iw, ih);
      const float qarea = __fmul_rn(__fadd_rn(__fadd_rn(q[2], -q[0]), 1.0f), __fadd_rn(__fadd_rn(q[3], -q[1]), 1.0f));
      const float uni = __fadd_rn(__fadd_rn(-inter, qarea), warea);
      if (__fdiv_rn(inter, uni) >= thresh) s_p[j * C + cls] = 0.f;
    }
    __syncthreads();
    for (int c = tid; c < C; c += blockDim.x) s_p[b * C + c] = -1.f;
    __syncthreads();
  }
}

}  // namespace

extern "C" {

int mb200_bbox_overlaps_f32(const float* boxes_a, int A, const float* boxes_b, int B, float* out,
                            cudaStream_t stream) {
  if (A <= 0 || B <= 0) return MB200_OK;
  bbox_overlaps_f32_kernel<<<grid_for((long long)A * B, 256), 256, 0, stream>>>(
      (const float4*)boxes_a, A, (const float4*)boxes_b, B, out);
  MB200_CHECK_LAUNCH("mb200_bbox_overlaps_f32");
  return MB200_OK;
}

int mb200_bbox_overlaps_f64(const double* boxes, int N, const double* query, int K, int mode, double* out,
                            cudaStream_t stream) {
  if (N <= 0 || K <= 0) return MB200_OK;
  if (mode != 0 && mode != 1) return MB200_ERR_ARG;
  bbox_overlaps_f64_kernel<<<grid_for((long long)N * K, 256), 256, 0, stream>>>(boxes, N, query, K, mode, out);
  MB200_CHECK_LAUNCH("mb200_bbox_overlaps_f64");
  return MB200_OK;
}

int mb200_anchor_targets(const double* anchors, int N, const double* gt_boxes, int G, double neg_thr, double pos_thr,
                         unsigned long long* gt_max_ws, double* max_overlaps, int* argmax, long long* labels,
                         cudaStream_t stream) {
  if (N <= 0) return MB200_OK;
  if (G <= 0) return MB200_ERR_ARG;
  MB200_CHECK(cudaMemsetAsync(gt_max_ws, 0, sizeof(unsigned long long) * (size_t)G, stream));
  const int blocks = min(mb200_div_up(N, 8), kNumSMs * 8);
  anchor_rowmax_kernel<<<blocks, 256, 0, stream>>>(anchors, N, gt_boxes, G, max_overlaps, argmax, gt_max_ws);
  MB200_CHECK_LAUNCH("anchor_rowmax_kernel");
  anchor_label_kernel<<<blocks, 256, 0, stream>>>(anchors, N, gt_boxes, G, max_overlaps, gt_max_ws, neg_thr, pos_thr, labels);
  MB200_CHECK_LAUNCH("anchor_label_kernel");
  return MB200_OK;
}

int mb200_decoder_commit(const float* boxes, const float* probs, int N, int C, float thresh, long long* commit,
                         cudaStream_t stream) {
  if (N <= 0) return MB200_OK;
  if (C <= 1) return MB200_ERR_ARG;
  const size_t smem = (size_t)N * C * sizeof(float);
  if (smem > 200 * 1024) return MB200_ERR_UNSUPPORTED;        // > ~330 detections x 151 classes: caller keeps the host loop
  static bool attr = false;
  if (!attr) {
    MB200_CHECK(cudaFuncSetAttribute(decoder_commit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  decoder_commit_kernel<<<1, 512, smem, stream>>>(boxes, probs, N, C, thresh, commit);
  MB200_CHECK_LAUNCH("decoder_commit_kernel");
  return MB200_OK;
}

int mb200_union_rois(const float* rois, const long long* pairs, int num_pairs, float* union_rois,
                     float* pair_boxes /* may be NULL */, cudaStream_t stream) {
  if (num_pairs <= 0) return MB200_OK;
  union_rois_kernel<<<mb200_div_up(num_pairs, 128), 128, 0, stream>>>(rois, pairs, num_pairs, union_rois, pair_boxes);
  MB200_CHECK_LAUNCH("mb200_union_rois");
  return MB200_OK;
}

int mb200_draw_union_boxes(const float* pair_boxes, int num_pairs, int pooling_size, float offset, float* out,
                           cudaStream_t stream) {
  if (num_pairs <= 0) return MB200_OK;
  if (pooling_size <= 0) return MB200_ERR_ARG;
  draw_union_boxes_kernel<<<num_pairs, 256, 0, stream>>>(pair_boxes, num_pairs, pooling_size, offset, out);
  MB200_CHECK_LAUNCH("mb200_draw_union_boxes");
  return MB200_OK;
}

int mb200_bbox_preds(const float* boxes, const float* deltas, long long num_rows, int rows_per_box,
                     const float* im_hw, const int* im_idx, float* out, cudaStream_t stream) {
  if (num_rows <= 0) return MB200_OK;
  if (rows_per_box <= 0) return MB200_ERR_ARG;
  bbox_preds_kernel<<<grid_for(num_rows, 256), 256, 0, stream>>>((const float4*)boxes, (const float4*)deltas,
      num_rows, rows_per_box, im_hw, im_idx, (float4*)out);
  MB200_CHECK_LAUNCH("mb200_bbox_preds");
  return MB200_OK;
}

}  // extern "C"
