// Layout / precision conversion kernels feeding the bf16x3 tensor-core path (gemm_tc.cu):
// fp32 -> (hi, lo) bf16 pairs with K padded to a multiple of 64, plain and transposed; conv
// weight re-layout; the 3-channel stem im2col; 2x2 max-pool on NHWC pairs. All HBM-bound,
// coalesced, vectorised where the layout allows.
#include "common.cuh"
#include "tc_common.cuh"

namespace {

// src [rows, cols] fp32 (row pitch ld) -> hi, lo [rows, Kp] bf16, zero padded for cols..Kp.
__global__ void split_kernel(const float* __restrict__ src, long long rows, int cols, long long ld, int Kp,
                             __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const long long total = rows * (long long)(Kp / 2);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)blockDim.x * gridDim.x) {
    const long long r = i / (Kp / 2);
    const int c = (int)(i - r * (Kp / 2)) * 2;
    const float a = c < cols ? src[r * ld + c] : 0.f;
    const float b = c + 1 < cols ? src[r * ld + c + 1] : 0.f;
    __nv_bfloat16 ah, al, bh, bl;
    tc::split_bf16(a, ah, al); tc::split_bf16(b, bh, bl);
    __nv_bfloat162 h; h.x = ah; h.y = bh;
    __nv_bfloat162 l; l.x = al; l.y = bl;
    *(__nv_bfloat162*)(hi + r * Kp + c) = h;
    *(__nv_bfloat162*)(lo + r * Kp + c) = l;
  }
}

// src [rows, cols] fp32 -> hi, lo [cols, Rp] (transposed), zero padded for rows..Rp.
__global__ void split_transpose_kernel(const float* __restrict__ src, int rows, int cols, long long ld, int Rp,
                                       __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? src[(long long)r * ld + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;       // output row = c, output col = r
    if (c < cols && r < Rp) {
      __nv_bfloat16 h, l;
      tc::split_bf16(tile[threadIdx.x][i], h, l);
      hi[(long long)c * Rp + r] = h;
      lo[(long long)c * Rp + r] = l;
    }
  }
}

// conv weight OIHW fp32 [O, I, 3, 3] -> hi, lo [O, 9*Ip] with K order (kh, kw, i); Ip >= I zero padded.
__global__ void conv_weight_split_kernel(const float* __restrict__ w, int O, int I, int Ip,
                                         __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const long long total = (long long)O * 9 * Ip;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)blockDim.x * gridDim.x) {
    const int i = idx % Ip;
    const int tap = (idx / Ip) % 9;
    const long long o = idx / (9LL * Ip);
    const float v = i < I ? w[(o * I + i) * 9 + tap] : 0.f;
    __nv_bfloat16 h, l; tc::split_bf16(v, h, l);
    hi[idx] = h; lo[idx] = l;
  }
}

// Stem (Cin = 3): x fp32 NCHW [B,3,H,W] -> im2col pair [B*H*W, 64], k = (kh*3+kw)*3 + c (27 used).
__global__ void im2col3_split_kernel(const float* __restrict__ x, int B, int H, int W,
                                     __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const long long total = (long long)B * H * W * 32;   // one thread per (pixel, k pair)
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)blockDim.x * gridDim.x) {
    const int kp = (int)(idx & 31);
    const long long pix = idx >> 5;
    const int w = pix % W; const int h = (pix / W) % H; const long long b = pix / ((long long)W * H);
    float v[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int k = kp * 2 + e;
      float t = 0.f;
      if (k < 27) {
        const int tap = k / 3, c = k - tap * 3;
        const int yy = h + tap / 3 - 1, xx = w + (tap % 3) - 1;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) t = __ldg(x + ((b * 3 + c) * H + yy) * W + xx);
      }
      v[e] = t;
    }
    __nv_bfloat16 ah, al, bh, bl;
    tc::split_bf16(v[0], ah, al); tc::split_bf16(v[1], bh, bl);
    __nv_bfloat162 hh; hh.x = ah; hh.y = bh;
    __nv_bfloat162 ll; ll.x = al; ll.y = bl;
    *(__nv_bfloat162*)(hi + pix * 64 + kp * 2) = hh;
    *(__nv_bfloat162*)(lo + pix * 64 + kp * 2) = ll;
  }
}

// stem weight [O, 3, 3, 3] OIHW -> [O, 64] with k = (kh*3+kw)*3 + c
__global__ void stem_weight_split_kernel(const float* __restrict__ w, int O, __nv_bfloat16* __restrict__ hi,
                                         __nv_bfloat16* __restrict__ lo) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= O * 64) return;
  const int k = idx & 63, o = idx >> 6;
  float v = 0.f;
  if (k < 27) { const int tap = k / 3, c = k - tap * 3; v = w[(o * 3 + c) * 9 + tap]; }
  __nv_bfloat16 h, l; tc::split_bf16(v, h, l);
  hi[idx] = h; lo[idx] = l;
}

// 2x2 / stride 2 max pool on an NHWC (hi, lo) pair (floor mode, as nn.MaxPool2d(2, 2)).
// hi + lo is exact in fp32, so the max is taken on reconstructed values and re-split.
__global__ void maxpool2_nhwc_split_kernel(const __nv_bfloat16* __restrict__ xhi, const __nv_bfloat16* __restrict__ xlo,
                                           int B, int H, int W, int C, __nv_bfloat16* __restrict__ yhi,
                                           __nv_bfloat16* __restrict__ ylo) {
  const int Ho = H / 2, Wo = W / 2, C2 = C / 2;
  const long long total = (long long)B * Ho * Wo * C2;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)blockDim.x * gridDim.x) {
    const int c = (int)(idx % C2) * 2;
    const long long pix = idx / C2;
    const int wo = pix % Wo; const int ho = (pix / Wo) % Ho; const long long b = pix / ((long long)Wo * Ho);
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const long long off = (((b * H + 2 * ho + dy) * W) + 2 * wo + dx) * C + c;
        const __nv_bfloat162 h = *(const __nv_bfloat162*)(xhi + off);
        const __nv_bfloat162 l = *(const __nv_bfloat162*)(xlo + off);
        m0 = fmaxf(m0, __bfloat162float(h.x) + __bfloat162float(l.x));
        m1 = fmaxf(m1, __bfloat162float(h.y) + __bfloat162float(l.y));
      }
    __nv_bfloat16 ah, al, bh, bl;
    tc::split_bf16(m0, ah, al); tc::split_bf16(m1, bh, bl);
    __nv_bfloat162 hh; hh.x = ah; hh.y = bh;
    __nv_bfloat162 ll; ll.x = al; ll.y = bl;
    const long long o = ((b * Ho + ho) * Wo + wo) * C + c;
    *(__nv_bfloat162*)(yhi + o) = hh;
    *(__nv_bfloat162*)(ylo + o) = ll;
  }
}

inline int blocks_for(long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  const long long cap = (long long)kNumSMs * 32;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

extern "C" {

int mb200_split_bf16(const float* src, long long rows, int cols, long long ld, int Kp, void* hi, void* lo,
                     cudaStream_t stream) {
  if (rows <= 0) return MB200_OK;
  if (Kp < cols || Kp % 2) return MB200_ERR_ARG;
  split_kernel<<<blocks_for(rows * (Kp / 2), 256), 256, 0, stream>>>(src, rows, cols, ld, Kp, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo);
  MB200_CHECK_LAUNCH("mb200_split_bf16");
  return MB200_OK;
}

int mb200_split_transpose_bf16(const float* src, int rows, int cols, long long ld, int Rp, void* hi, void* lo,
                               cudaStream_t stream) {
  if (rows <= 0 || cols <= 0) return MB200_OK;
  if (Rp < rows) return MB200_ERR_ARG;
  dim3 grid(mb200_div_up(cols, 32), mb200_div_up(Rp, 32));
  if (grid.y > 65535) return MB200_ERR_UNSUPPORTED;
  split_transpose_kernel<<<grid, dim3(32, 8), 0, stream>>>(src, rows, cols, ld, Rp, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo);
  MB200_CHECK_LAUNCH("mb200_split_transpose_bf16");
  return MB200_OK;
}

int mb200_conv_weight_split(const float* w_oihw, int O, int I, int Ip, void* hi, void* lo, cudaStream_t stream) {
  if (O <= 0) return MB200_OK;
  conv_weight_split_kernel<<<blocks_for((long long)O * 9 * Ip, 256), 256, 0, stream>>>(w_oihw, O, I, Ip, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo);
  MB200_CHECK_LAUNCH("mb200_conv_weight_split");
  return MB200_OK;
}

int mb200_im2col3_split(const float* x_nchw, int B, int H, int W, void* hi, void* lo, cudaStream_t stream) {
  if (B <= 0) return MB200_OK;
  im2col3_split_kernel<<<blocks_for((long long)B * H * W * 32, 256), 256, 0, stream>>>(x_nchw, B, H, W, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo);
  MB200_CHECK_LAUNCH("mb200_im2col3_split");
  return MB200_OK;
}

int mb200_stem_weight_split(const float* w_oihw, int O, void* hi, void* lo, cudaStream_t stream) {
  if (O <= 0) return MB200_OK;
  stem_weight_split_kernel<<<mb200_div_up(O * 64, 256), 256, 0, stream>>>(w_oihw, O, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo);
  MB200_CHECK_LAUNCH("mb200_stem_weight_split");
  return MB200_OK;
}

int mb200_maxpool2_nhwc_split(const void* xhi, const void* xlo, int B, int H, int W, int C, void* yhi, void* ylo,
                              cudaStream_t stream) {
  if (B <= 0) return MB200_OK;
  if (C % 2) return MB200_ERR_ARG;
  maxpool2_nhwc_split_kernel<<<blocks_for((long long)B * (H / 2) * (W / 2) * (C / 2), 256), 256, 0, stream>>>(
      (const __nv_bfloat16*)xhi, (const __nv_bfloat16*)xlo, B, H, W, C, (__nv_bfloat16*)yhi, (__nv_bfloat16*)ylo);
  MB200_CHECK_LAUNCH("mb200_maxpool2_nhwc_split");
  return MB200_OK;
}

}  // extern "C"
