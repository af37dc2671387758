// The union-box mask branch of UnionBoxesAndFeats (lib/get_union_boxes.py:28-37 of the reference):
//   conv7x7/s2 (2->C1) + ReLU + BatchNorm + maxpool3x3/s2 + conv3x3 (C1->C2) + ReLU + BatchNorm
// on [R,2,27,27] box masks, forward and backward, as NHWC streaming kernels around the tcgen05 GEMM
// (gemm_tc.cu): the two convolutions become (explicit im2col) x (weight matrix) products, everything
// else here is HBM-bound layout / normalisation work:
//   * im2col of the 7x7/s2 stem and of a 3x3/p1 NHWC map, plain ([P, K]) or transposed ([K, P], the
//     K-major operand of the weight-gradient GEMM), emitted directly as (hi, lo) bf16 pairs;
//   * training-mode BatchNorm statistics (two-pass, double accumulation), BN + max-pool fused,
//     BN + NHWC->NCHW + residual add fused; BN/ReLU backward (reduce + apply), col2im, un-pool.
// ncu launch list of the SGCls step before this file: cuDNN's fp32 SIMT forward convs took 5.9 ms of a
// 28 ms step (TF32 had to be off to stay inside the fp32 parity bar).
#include "common.cuh"
#include "tc_common.cuh"

namespace {

inline int blocks_for(long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  const long long cap = (long long)kNumSMs * 32;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

__device__ __forceinline__ void store_pair(__nv_bfloat16* hi, __nv_bfloat16* lo, long long off, float a, float b) {
  __nv_bfloat16 ah, al, bh, bl;
  tc::split_bf16(a, ah, al); tc::split_bf16(b, bh, bl);
  __nv_bfloat162 h; h.x = ah; h.y = bh;
  __nv_bfloat162 l; l.x = al; l.y = bl;
  *(__nv_bfloat162*)(hi + off) = h;
  *(__nv_bfloat162*)(lo + off) = l;
}

// ---------------------------------------------------------------- 7x7 / stride 2 / pad 3 im2col of [R,2,S,S]
// k = (ky*7 + kx)*2 + c (98 used of 128). Plain: out [R*Ho*Wo, 128]; one thread per (p, k pair).
__global__ void im2col7s2_kernel(const float* __restrict__ m, int R, int S, int Ho, __nv_bfloat16* __restrict__ hi,
                                 __nv_bfloat16* __restrict__ lo) {
  const unsigned total = (unsigned)R * Ho * Ho * 64u;          // host guarantees < 2^31
  for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += blockDim.x * gridDim.x) {
    const unsigned kp = idx & 63u;
    const unsigned p = idx >> 6;
    const unsigned wo = p % Ho, t = p / Ho, ho = t % Ho, r = t / Ho;
    float v[2] = {0.f, 0.f};
    if (kp < 49) {                                   // k pair = the two channels of one tap
      const int ky = kp / 7, kx = kp - ky * 7;
      const int yy = 2 * (int)ho + ky - 3, xx = 2 * (int)wo + kx - 3;
      if (yy >= 0 && yy < S && xx >= 0 && xx < S) {
        const float* src = m + ((size_t)r * 2 * S + yy) * S + xx;
        v[0] = __ldg(src);
        v[1] = __ldg(src + S * S);
      }
    }
    store_pair(hi, lo, (long long)p * 128 + kp * 2, v[0], v[1]);
  }
}

// Transposed: out [128, Pp]; one thread per (k, p pair), p fastest; columns >= P and rows >= 98 are zero.
__global__ void im2col7s2_t_kernel(const float* __restrict__ m, int R, int S, int Ho, long long P, long long Pp,
                                   __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const unsigned half = (unsigned)(Pp / 2);                    // host guarantees 128 * half < 2^32
  const unsigned k = blockIdx.y;                               // one k row per blockIdx.y
  const int tap = k >> 1, c = k & 1;
  const int ky = tap / 7, kx = tap - ky * 7;
  for (unsigned j = blockIdx.x * blockDim.x + threadIdx.x; j < half; j += blockDim.x * gridDim.x) {
    const unsigned p0 = j * 2;
    float v[2] = {0.f, 0.f};
    if (k < 98) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const unsigned p = p0 + e;
        if (p < (unsigned)P) {
          const unsigned wo = p % Ho, t = p / Ho, ho = t % Ho, r = t / Ho;
          const int yy = 2 * (int)ho + ky - 3, xx = 2 * (int)wo + kx - 3;
          if (yy >= 0 && yy < S && xx >= 0 && xx < S) v[e] = __ldg(m + (((size_t)r * 2 + c) * S + yy) * S + xx);
        }
      }
    }
    store_pair(hi, lo, (long long)k * Pp + p0, v[0], v[1]);
  }
}

// ---------------------------------------------------------------- 3x3 / stride 1 / pad 1 im2col of NHWC [R,H,W,C]
// k = tap*C + c (tap = ky*3 + kx), matching conv_weight_split_kernel. Plain: out [P, 9C].
__global__ void im2col3_nhwc_kernel(const float* __restrict__ x, int R, int H, int W, int C,
                                    __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const unsigned C4 = C / 4;
  const unsigned total = (unsigned)R * H * W * C4;             // host guarantees < 2^31
  for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += blockDim.x * gridDim.x) {
    const unsigned c = (idx % C4) * 4;
    const unsigned p = idx / C4;
    const int w = p % W; const unsigned t = p / W; const int h = t % H; const unsigned r = t / H;
    const long long orow = (long long)p * 9 * C + c;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int yy = h + tap / 3 - 1, xx = w + tap % 3 - 1;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = __ldg((const float4*)(x + (((size_t)r * H + yy) * W + xx) * C + c));
      const long long off = orow + (long long)tap * C;
      store_pair(hi, lo, off, v.x, v.y);
      store_pair(hi, lo, off + 2, v.z, v.w);
    }
  }
}

// Transposed: out [9C, Pp]. Block = one tap x 32 channels x 64 pixels, transposed through shared memory.
__global__ void im2col3_nhwc_t_kernel(const float* __restrict__ x, int R, int H, int W, int C, long long P, long long Pp,
                                      __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  __shared__ float tile[64][33];
  __shared__ int pix[64];                                      // (r*H + h)*W + w of each pixel, -1 beyond P
  __shared__ short ph[64], pw[64];
  const int c0 = blockIdx.y * 32;
  const long long p0 = (long long)blockIdx.x * 64;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  if (tid < 64) {
    const long long p = p0 + tid;
    int q = -1, h = 0, w = 0;
    if (p < P) {
      const unsigned pu = (unsigned)p;                         // host guarantees P < 2^31
      w = pu % W; h = (pu / W) % H; q = (int)pu;
    }
    pix[tid] = q; ph[tid] = (short)h; pw[tid] = (short)w;
  }
  __syncthreads();
  for (int tap = 0; tap < 9; ++tap) {
    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
#pragma unroll
    for (int i = threadIdx.y; i < 64; i += 8) {
      const int q = pix[i], yy = ph[i] + dy, xx = pw[i] + dx;
      float v = 0.f;
      if (q >= 0 && yy >= 0 && yy < H && xx >= 0 && xx < W)
        v = __ldg(x + ((long long)q + dy * W + dx) * C + c0 + threadIdx.x);
      tile[i][threadIdx.x] = v;
    }
    __syncthreads();
#pragma unroll
    for (int i = threadIdx.y; i < 32; i += 8) {                // i = channel within the block, threadIdx.x = pixel pair
      const long long off = ((long long)tap * C + c0 + i) * Pp + p0 + 2 * threadIdx.x;
      store_pair(hi, lo, off, tile[2 * threadIdx.x][i], tile[2 * threadIdx.x + 1][i]);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------- per-channel sums over the rows of [P, C]
// sums[c] += sum_p f(p,c), sums[C+c] += sum_p g(p,c); block = 32 channels x 8 row lanes, fp32 per thread
// (a few hundred terms), double across threads / blocks.
template <int MODE>   // 0: (x - shift, (x - shift)^2)   1: (g, g * (x - mean) * invstd)
__global__ void colsum2_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ s0,
                               const float* __restrict__ s1, long long P, int C, double* __restrict__ sums) {
  __shared__ double red[2][8][32];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float acc0 = 0.f, acc1 = 0.f;
  if (c < C) {
    const float sh = s0 ? s0[c] : 0.f;
    const float is = (MODE == 1) ? s1[c] : 0.f;
    for (long long p = blockIdx.y * 8 + threadIdx.y; p < P; p += 8LL * gridDim.y) {
      const float v = __ldg(a + p * C + c);
      if (MODE == 0) {
        const float d = v - sh;
        acc0 += d; acc1 = fmaf(d, d, acc1);
      } else {
        const float xh = (__ldg(b + p * C + c) - sh) * is;
        acc0 += v; acc1 = fmaf(v, xh, acc1);
      }
    }
  }
  red[0][threadIdx.y][threadIdx.x] = (double)acc0;
  red[1][threadIdx.y][threadIdx.x] = (double)acc1;
  __syncthreads();
  if (threadIdx.y < 2 && c < C) {
    double t = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += red[threadIdx.y][j][threadIdx.x];
    atomicAdd(sums + (long long)threadIdx.y * C + c, t);
  }
}

// Turns the pass sums into statistics. pass 0: mean[c] = S0/P (first estimate used as the shift of pass 1).
// pass 1: mean += S0/P; var = S1/P - (S0/P)^2 (biased, as the normalisation uses); invstd = rsqrt(var + eps);
// running stats (nullable) move by `momentum` with the unbiased variance, as nn.BatchNorm2d in training.
__global__ void bn_finalize_kernel(const double* __restrict__ sums, long long P, int C, int pass, float eps, float momentum,
                                   float* __restrict__ mean, float* __restrict__ invstd, float* __restrict__ run_mean,
                                   float* __restrict__ run_var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double s0 = sums[c] / (double)P;
  if (pass == 0) { mean[c] = (float)s0; return; }
  const double m = (double)mean[c] + s0;
  double var = sums[C + c] / (double)P - s0 * s0;
  if (var < 0.0) var = 0.0;
  mean[c] = (float)m;
  invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (run_mean) {
    const double unb = P > 1 ? var * (double)P / (double)(P - 1) : var;
    run_mean[c] = (float)((1.0 - momentum) * (double)run_mean[c] + (double)momentum * m);
    run_var[c] = (float)((1.0 - momentum) * (double)run_var[c] + (double)momentum * unb);
  }
}

// ---------------------------------------------------------------- BN + 3x3/s2/p1 max-pool, NHWC
__global__ void bn_pool3s2_nhwc_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                       const float* __restrict__ invstd, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, int R, int H, int W, int C, int Ho, int Wo,
                                       float* __restrict__ y, unsigned char* __restrict__ arg) {
  const int C4 = C / 4;
  const long long total = (long long)R * Ho * Wo * C4;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)blockDim.x * gridDim.x) {
    const int c = (int)(idx % C4) * 4;
    const long long q = idx / C4;
    const int ox = (int)(q % Wo), oy = (int)((q / Wo) % Ho);
    const long long r = q / ((long long)Wo * Ho);
    const float4 mu = *(const float4*)(mean + c), is = *(const float4*)(invstd + c);
    const float4 ga = *(const float4*)(gamma + c), be = *(const float4*)(beta + c);
    float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    int bi[4] = {0, 0, 0, 0};
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int yy = 2 * oy - 1 + dy, xx = 2 * ox - 1 + dx;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
          const float4 v = __ldg((const float4*)(x + ((r * H + yy) * W + xx) * C + c));
          const float t[4] = {(v.x - mu.x) * is.x * ga.x + be.x, (v.y - mu.y) * is.y * ga.y + be.y,
                              (v.z - mu.z) * is.z * ga.z + be.z, (v.w - mu.w) * is.w * ga.w + be.w};
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (t[e] > best[e] || (t[e] != t[e])) { best[e] = t[e]; bi[e] = dy * 3 + dx; }   // first maximum (ATen)
        }
      }
    *(float4*)(y + q * C + c) = make_float4(best[0], best[1], best[2], best[3]);
    *(uchar4*)(arg + q * C + c) = make_uchar4((unsigned char)bi[0], (unsigned char)bi[1], (unsigned char)bi[2],
                                              (unsigned char)bi[3]);
  }
}

// dx[r,Y,X,c] = sum over the <= 4 windows covering (Y,X) whose arg-max is (Y,X) of dy[r,oy,ox,c]
__global__ void unpool3s2_nhwc_kernel(const float* __restrict__ gy, const unsigned char* __restrict__ arg, int R, int H,
                                      int W, int C, int Ho, int Wo, float* __restrict__ gx) {
  const int C4 = C / 4;
  const long long total = (long long)R * H * W * C4;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)blockDim.x * gridDim.x) {
    const int c = (int)(idx % C4) * 4;
    const long long q = idx / C4;
    const int xx = (int)(q % W), yy = (int)((q / W) % H);
    const long long r = q / ((long long)W * H);
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    const int oy0 = yy / 2, oy1 = (yy + 1) / 2, ox0 = xx / 2, ox1 = (xx + 1) / 2;
    for (int oy = oy0; oy <= oy1; ++oy) {
      if (oy >= Ho) continue;
      const int dy = yy - (2 * oy - 1);
      for (int ox = ox0; ox <= ox1; ++ox) {
        if (ox >= Wo) continue;
        const int code = dy * 3 + (xx - (2 * ox - 1));
        const long long o = ((r * Ho + oy) * Wo + ox) * C + c;
        const uchar4 a = *(const uchar4*)(arg + o);
        const float4 g = __ldg((const float4*)(gy + o));
        if (a.x == code) s[0] += g.x;
        if (a.y == code) s[1] += g.y;
        if (a.z == code) s[2] += g.z;
        if (a.w == code) s[3] += g.w;
      }
    }
    *(float4*)(gx + q * C + c) = make_float4(s[0], s[1], s[2], s[3]);
  }
}

// ---------------------------------------------------------------- BN apply + NHWC -> NCHW (+ addend), per (roi, 64 ch)
// x [R, HW, C] -> out [R, C, HW]; HW <= 64.
__global__ void bn_nhwc_to_nchw_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                       const float* __restrict__ invstd, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, const float* __restrict__ addend, int HW, int C,
                                       float* __restrict__ out) {
  __shared__ float tile[64][65];
  const long long r = blockIdx.x;
  const int c0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;     // 256 threads: 64 x 4
  const int c = c0 + tx;
  if (c < C) {
    const float mu = mean[c], sc = invstd[c] * gamma[c], be = beta[c];
    for (int p = ty; p < HW; p += 4) tile[p][tx] = (__ldg(x + (r * HW + p) * C + c) - mu) * sc + be;
  }
  __syncthreads();
  const int nch = min(64, C - c0);
  const long long base = (r * C + c0) * HW;                    // contiguous [nch, HW] block of the output
  for (int i = threadIdx.x; i < nch * HW; i += blockDim.x) {
    const int ci = i / HW, p = i - ci * HW;
    float v = tile[p][ci];
    if (addend) v += __ldg(addend + base + i);
    out[base + i] = v;
  }
}

// g [R, C, HW] -> out [R, HW, C]
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ g, int HW, int C, float* __restrict__ out) {
  __shared__ float tile[64][65];
  const long long r = blockIdx.x;
  const int c0 = blockIdx.y * 64;
  const int nch = min(64, C - c0);
  const long long base = (r * C + c0) * HW;
  for (int i = threadIdx.x; i < nch * HW; i += blockDim.x) {
    const int ci = i / HW, p = i - ci * HW;
    tile[p][ci] = __ldg(g + base + i);
  }
  __syncthreads();
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  if (c0 + tx < C)
    for (int p = ty; p < HW; p += 4) out[(r * HW + p) * C + c0 + tx] = tile[p][tx];
}

// ---------------------------------------------------------------- (conv -> ReLU -> BN) backward, element pass
// dz = (x > 0) * gamma * invstd * (g - Sg/P - xhat * Sgx/P); also column sums of dz (the conv bias gradient).
__global__ void bn_relu_bwd_kernel(const float* __restrict__ g, const float* __restrict__ x, const float* __restrict__ mean,
                                   const float* __restrict__ invstd, const float* __restrict__ gamma,
                                   const double* __restrict__ sums, long long P, int C, float* __restrict__ dz,
                                   double* __restrict__ dbias) {
  __shared__ double red[8][32];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (c < C) {
    const float mu = mean[c], is = invstd[c];
    const float k = gamma[c] * is;
    const float mg = (float)(sums[c] / (double)P), mgx = (float)(sums[C + c] / (double)P);
    for (long long p = blockIdx.y * 8 + threadIdx.y; p < P; p += 8LL * gridDim.y) {
      const float xv = __ldg(x + p * C + c);
      const float xh = (xv - mu) * is;
      float d = k * (__ldg(g + p * C + c) - mg - xh * mgx);
      if (!(xv > 0.f)) d = 0.f;
      dz[p * C + c] = d;
      acc += d;
    }
  }
  red[threadIdx.y][threadIdx.x] = (double)acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    double t = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += red[j][threadIdx.x];
    atomicAdd(dbias + c, t);
  }
}

// ---------------------------------------------------------------- fused variants of the backward pair above
// g is either a tensor [P,C] (UNPOOL == false) or gathered on the fly from the pooled gradient
// gy [R,Ho,Wo,C] and the arg-max codes (UNPOOL == true: P = R*H*W rows, x is the pool's input).
template <bool UNPOOL>
__device__ __forceinline__ float grad_at(const float* __restrict__ g, const unsigned char* __restrict__ arg, long long p,
                                         int c, int C, int H, int W, int Ho, int Wo) {
  if (!UNPOOL) return __ldg(g + p * C + c);
  const unsigned pu = (unsigned)p;
  const int xx = pu % W, yy = (pu / W) % H;
  const long long r = pu / ((unsigned)W * H);
  float s = 0.f;
  const int oy0 = yy / 2, oy1 = (yy + 1) / 2, ox0 = xx / 2, ox1 = (xx + 1) / 2;
  for (int oy = oy0; oy <= oy1; ++oy) {
    if (oy >= Ho) continue;
    const int dy = yy - (2 * oy - 1);
    for (int ox = ox0; ox <= ox1; ++ox) {
      if (ox >= Wo) continue;
      const long long o = ((r * Ho + oy) * Wo + ox) * C + c;
      if (arg[o] == dy * 3 + (xx - (2 * ox - 1))) s += __ldg(g + o);
    }
  }
  return s;
}

template <bool UNPOOL>
__global__ void bn_bwd_reduce_kernel(const float* __restrict__ g, const unsigned char* __restrict__ arg,
                                     const float* __restrict__ x, const float* __restrict__ mean,
                                     const float* __restrict__ invstd, long long P, int C, int H, int W, int Ho, int Wo,
                                     double* __restrict__ sums) {
  __shared__ double red[2][8][32];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float acc0 = 0.f, acc1 = 0.f;
  if (c < C) {
    const float mu = mean[c], is = invstd[c];
    for (long long p = blockIdx.y * 8 + threadIdx.y; p < P; p += 8LL * gridDim.y) {
      const float gv = grad_at<UNPOOL>(g, arg, p, c, C, H, W, Ho, Wo);
      const float xh = (__ldg(x + p * C + c) - mu) * is;
      acc0 += gv; acc1 = fmaf(gv, xh, acc1);
    }
  }
  red[0][threadIdx.y][threadIdx.x] = (double)acc0;
  red[1][threadIdx.y][threadIdx.x] = (double)acc1;
  __syncthreads();
  if (threadIdx.y < 2 && c < C) {
    double t = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += red[threadIdx.y][j][threadIdx.x];
    atomicAdd(sums + (long long)threadIdx.y * C + c, t);
  }
}

// dz as bf16 pairs: transposed [C, Pp] (always; columns P..Pp zeroed) and, when plain_hi != NULL, plain [P, C].
// Grid (C/32, NB): a block owns 32 channels and walks 64-pixel tiles, so the bias gradient needs NB*C atomics.
template <bool UNPOOL>
__global__ void bn_relu_bwd_split_kernel(const float* __restrict__ g, const unsigned char* __restrict__ arg,
                                         const float* __restrict__ x, const float* __restrict__ mean,
                                         const float* __restrict__ invstd, const float* __restrict__ gamma,
                                         const double* __restrict__ sums, long long P, long long Pp, int C, int H, int W,
                                         int Ho, int Wo, __nv_bfloat16* __restrict__ t_hi, __nv_bfloat16* __restrict__ t_lo,
                                         __nv_bfloat16* __restrict__ p_hi, __nv_bfloat16* __restrict__ p_lo,
                                         double* __restrict__ dbias) {
  __shared__ float tile[64][33];
  __shared__ double red[8][32];
  const int c0 = blockIdx.x * 32, c = c0 + threadIdx.x;        // host guarantees C % 32 == 0
  const float mu = mean[c], is = invstd[c];
  const float k = gamma[c] * is;
  const float mg = (float)(sums[c] / (double)P), mgx = (float)(sums[C + c] / (double)P);
  float acc = 0.f;
  for (long long p0 = (long long)blockIdx.y * 64; p0 < Pp; p0 += 64LL * gridDim.y) {
#pragma unroll 2
    for (int i = threadIdx.y; i < 64; i += 8) {
      const long long p = p0 + i;
      float d = 0.f;
      if (p < P) {
        const float xv = __ldg(x + p * C + c);
        const float gv = grad_at<UNPOOL>(g, arg, p, c, C, H, W, Ho, Wo);
        d = k * (gv - mg - (xv - mu) * is * mgx);
        if (!(xv > 0.f)) d = 0.f;
        if (p_hi) {
          __nv_bfloat16 dh, dl; tc::split_bf16(d, dh, dl);
          p_hi[p * C + c] = dh; p_lo[p * C + c] = dl;
        }
      }
      tile[i][threadIdx.x] = d;
      acc += d;
    }
    __syncthreads();
#pragma unroll
    for (int i = threadIdx.y; i < 32; i += 8) {                // i = channel within the block, threadIdx.x = pixel pair
      const long long off = ((long long)(c0 + i)) * Pp + p0 + 2 * threadIdx.x;
      store_pair(t_hi, t_lo, off, tile[2 * threadIdx.x][i], tile[2 * threadIdx.x + 1][i]);
    }
    __syncthreads();
  }
  red[threadIdx.y][threadIdx.x] = (double)acc;
  __syncthreads();
  if (threadIdx.y == 0) {
    double t = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += red[j][threadIdx.x];
    atomicAdd(dbias + c, t);
  }
}

// ---------------------------------------------------------------- col2im of the 3x3/p1 im2col layout
// dx[r,y,x,c] = sum_tap dcol[(r, y - dy, x - dx), tap*C + c]
__global__ void col2im3_nhwc_kernel(const float* __restrict__ dcol, int R, int H, int W, int C, float* __restrict__ dx) {
  const int C4 = C / 4;
  const long long total = (long long)R * H * W * C4;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)blockDim.x * gridDim.x) {
    const int c = (int)(idx % C4) * 4;
    const long long q = idx / C4;
    const int xx = (int)(q % W), yy = (int)((q / W) % H);
    const long long r = q / ((long long)W * H);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int sy = yy - (tap / 3 - 1), sx = xx - (tap % 3 - 1);
      if (sy >= 0 && sy < H && sx >= 0 && sx < W) {
        const float4 v = __ldg((const float4*)(dcol + ((r * H + sy) * W + sx) * 9LL * C + (long long)tap * C + c));
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
    }
    *(float4*)(dx + q * C + c) = s;
  }
}

inline dim3 colsum_grid(long long P, int C) {
  const int gx = mb200_div_up(C, 32);
  long long gy = (kNumSMs * 8 + gx - 1) / gx;
  const long long maxy = (P + 7) / 8;
  if (gy > maxy) gy = maxy;
  if (gy < 1) gy = 1;
  return dim3(gx, (unsigned)gy);
}

}  // namespace

extern "C" {

int mb200_im2col7s2_split(const float* masks, int R, int S, int transposed, long long Pp, void* hi, void* lo,
                          cudaStream_t stream) {
  if (R <= 0) return MB200_OK;
  const int Ho = (S + 6 - 7) / 2 + 1;
  const long long P = (long long)R * Ho * Ho;
  if (transposed) {
    if (Pp < P || Pp % 2) return MB200_ERR_ARG;
    if (Pp >= (1LL << 31)) return MB200_ERR_UNSUPPORTED;
    im2col7s2_t_kernel<<<dim3(blocks_for(Pp / 2, 256) < 64 ? blocks_for(Pp / 2, 256) : 64, 128), 256, 0, stream>>>(
        masks, R, S, Ho, P, Pp, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo);
  } else {
    if (P * 64 >= (1LL << 31)) return MB200_ERR_UNSUPPORTED;
    im2col7s2_kernel<<<blocks_for(P * 64, 256), 256, 0, stream>>>(masks, R, S, Ho, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo);
  }
  MB200_CHECK_LAUNCH("mb200_im2col7s2_split");
  return MB200_OK;
}

int mb200_im2col3_nhwc_split(const float* x, int R, int H, int W, int C, int transposed, long long Pp, void* hi,
                             void* lo, cudaStream_t stream) {
  if (R <= 0) return MB200_OK;
  const long long P = (long long)R * H * W;
  if (transposed) {
    if (C % 32 || Pp < P || Pp % 64) return MB200_ERR_ARG;
    if (Pp >= (1LL << 31)) return MB200_ERR_UNSUPPORTED;
    dim3 grid((unsigned)(Pp / 64), C / 32);
    im2col3_nhwc_t_kernel<<<grid, dim3(32, 8), 0, stream>>>(x, R, H, W, C, P, Pp, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo);
  } else {
    if (C % 4) return MB200_ERR_ARG;
    if (P * (C / 4) >= (1LL << 31)) return MB200_ERR_UNSUPPORTED;
    im2col3_nhwc_kernel<<<blocks_for(P * (C / 4), 256), 256, 0, stream>>>(x, R, H, W, C, (__nv_bfloat16*)hi,
                                                                               (__nv_bfloat16*)lo);
  }
  MB200_CHECK_LAUNCH("mb200_im2col3_nhwc_split");
  return MB200_OK;
}

// Training-mode BatchNorm statistics of x [P, C]: mean, invstd (biased variance + eps) and the running
// statistics update (running_* may be NULL). `sums` is scratch for 4*C doubles.
int mb200_bn_stats(const float* x, long long P, int C, float eps, float momentum, double* sums, float* mean,
                   float* invstd, float* running_mean, float* running_var, cudaStream_t stream) {
  if (P <= 0 || C <= 0) return MB200_ERR_ARG;
  MB200_CHECK(cudaMemsetAsync(sums, 0, sizeof(double) * 4 * C, stream));
  const dim3 grid = colsum_grid(P, C);
  colsum2_kernel<0><<<grid, dim3(32, 8), 0, stream>>>(x, nullptr, nullptr, nullptr, P, C, sums);
  bn_finalize_kernel<<<mb200_div_up(C, 128), 128, 0, stream>>>(sums, P, C, 0, eps, momentum, mean, invstd, nullptr, nullptr);
  colsum2_kernel<0><<<grid, dim3(32, 8), 0, stream>>>(x, nullptr, mean, nullptr, P, C, sums + 2 * C);
  bn_finalize_kernel<<<mb200_div_up(C, 128), 128, 0, stream>>>(sums + 2 * C, P, C, 1, eps, momentum, mean, invstd,
                                                                 running_mean, running_var);
  MB200_CHECK_LAUNCH("mb200_bn_stats");
  return MB200_OK;
}

int mb200_bn_pool3s2_nhwc(const float* x, const float* mean, const float* invstd, const float* gamma, const float* beta,
                          int R, int H, int W, int C, float* y, unsigned char* argmax, cudaStream_t stream) {
  if (R <= 0) return MB200_OK;
  if (C % 4) return MB200_ERR_ARG;
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  bn_pool3s2_nhwc_kernel<<<blocks_for((long long)R * Ho * Wo * (C / 4), 256), 256, 0, stream>>>(
      x, mean, invstd, gamma, beta, R, H, W, C, Ho, Wo, y, argmax);
  MB200_CHECK_LAUNCH("mb200_bn_pool3s2_nhwc");
  return MB200_OK;
}

int mb200_unpool3s2_nhwc(const float* grad_y, const unsigned char* argmax, int R, int H, int W, int C, float* grad_x,
                         cudaStream_t stream) {
  if (R <= 0) return MB200_OK;
  if (C % 4) return MB200_ERR_ARG;
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  unpool3s2_nhwc_kernel<<<blocks_for((long long)R * H * W * (C / 4), 256), 256, 0, stream>>>(grad_y, argmax, R, H, W, C,
                                                                                               Ho, Wo, grad_x);
  MB200_CHECK_LAUNCH("mb200_unpool3s2_nhwc");
  return MB200_OK;
}

int mb200_bn_nhwc_to_nchw(const float* x, const float* mean, const float* invstd, const float* gamma, const float* beta,
                          const float* addend, int R, int HW, int C, float* out, cudaStream_t stream) {
  if (R <= 0) return MB200_OK;
  if (HW > 64 || HW <= 0) return MB200_ERR_UNSUPPORTED;
  bn_nhwc_to_nchw_kernel<<<dim3(R, mb200_div_up(C, 64)), 256, 0, stream>>>(x, mean, invstd, gamma, beta, addend, HW, C, out);
  MB200_CHECK_LAUNCH("mb200_bn_nhwc_to_nchw");
  return MB200_OK;
}

int mb200_nchw_to_nhwc(const float* x, int R, int C, int HW, float* out, cudaStream_t stream) {
  if (R <= 0) return MB200_OK;
  if (HW > 64 || HW <= 0) return MB200_ERR_UNSUPPORTED;
  nchw_to_nhwc_kernel<<<dim3(R, mb200_div_up(C, 64)), 256, 0, stream>>>(x, HW, C, out);
  MB200_CHECK_LAUNCH("mb200_nchw_to_nhwc");
  return MB200_OK;
}

// Backward of y = BN_train(x), x = ReLU(conv): given g = dL/dy [P,C] and the saved x, mean, invstd:
// sums[0:C] = sum g (= dbeta), sums[C:2C] = sum g*xhat (= dgamma), dz [P,C] = dL/d(conv output),
// dbias[C] = column sums of dz. `sums` (2C doubles) and `dbias` (C doubles) are overwritten.
int mb200_bn_relu_backward(const float* g, const float* x, const float* mean, const float* invstd, const float* gamma,
                           long long P, int C, double* sums, float* dz, double* dbias, cudaStream_t stream) {
  if (P <= 0 || C <= 0) return MB200_ERR_ARG;
  MB200_CHECK(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, stream));
  MB200_CHECK(cudaMemsetAsync(dbias, 0, sizeof(double) * C, stream));
  const dim3 grid = colsum_grid(P, C);
  colsum2_kernel<1><<<grid, dim3(32, 8), 0, stream>>>(g, x, mean, invstd, P, C, sums);
  bn_relu_bwd_kernel<<<grid, dim3(32, 8), 0, stream>>>(g, x, mean, invstd, gamma, sums, P, C, dz, dbias);
  MB200_CHECK_LAUNCH("mb200_bn_relu_backward");
  return MB200_OK;
}

// Fused form of mb200_bn_relu_backward for the weight-gradient GEMMs: dz leaves as bf16 (hi, lo) pairs, transposed
// [C, Pp] (Pp % 64 == 0, zero padded) and optionally plain [P, C]; with argmax != NULL, g is the POOLED gradient
// [R,Ho,Wo,C] and is routed through the 3x3/s2/p1 max-pool on the fly (P must be R*H*W).
int mb200_bn_relu_backward_split(const float* g, const unsigned char* argmax, const float* x, const float* mean,
                                 const float* invstd, const float* gamma, long long P, long long Pp, int C, int H, int W,
                                 double* sums, void* t_hi, void* t_lo, void* p_hi, void* p_lo, double* dbias,
                                 cudaStream_t stream) {
  if (P <= 0 || C <= 0 || C % 32 || Pp < P || Pp % 64 || P >= (1LL << 31)) return MB200_ERR_ARG;
  MB200_CHECK(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, stream));
  MB200_CHECK(cudaMemsetAsync(dbias, 0, sizeof(double) * C, stream));
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const dim3 rgrid = colsum_grid(P, C);
  long long nb = (kNumSMs * 6) / (C / 32);
  if (nb > Pp / 64) nb = Pp / 64;
  if (nb < 1) nb = 1;
  const dim3 agrid(C / 32, (unsigned)nb);
  __nv_bfloat16 *th = (__nv_bfloat16*)t_hi, *tl = (__nv_bfloat16*)t_lo, *ph = (__nv_bfloat16*)p_hi, *pl = (__nv_bfloat16*)p_lo;
  if (argmax) {
    bn_bwd_reduce_kernel<true><<<rgrid, dim3(32, 8), 0, stream>>>(g, argmax, x, mean, invstd, P, C, H, W, Ho, Wo, sums);
    bn_relu_bwd_split_kernel<true><<<agrid, dim3(32, 8), 0, stream>>>(g, argmax, x, mean, invstd, gamma, sums, P, Pp, C, H, W,
                                                                       Ho, Wo, th, tl, ph, pl, dbias);
  } else {
    bn_bwd_reduce_kernel<false><<<rgrid, dim3(32, 8), 0, stream>>>(g, nullptr, x, mean, invstd, P, C, H, W, Ho, Wo, sums);
    bn_relu_bwd_split_kernel<false><<<agrid, dim3(32, 8), 0, stream>>>(g, nullptr, x, mean, invstd, gamma, sums, P, Pp, C, H,
                                                                        W, Ho, Wo, th, tl, ph, pl, dbias);
  }
  MB200_CHECK_LAUNCH("mb200_bn_relu_backward_split");
  return MB200_OK;
}

int mb200_col2im3_nhwc(const float* dcol, int R, int H, int W, int C, float* dx, cudaStream_t stream) {
  if (R <= 0) return MB200_OK;
  if (C % 4) return MB200_ERR_ARG;
  col2im3_nhwc_kernel<<<blocks_for((long long)R * H * W * (C / 4), 256), 256, 0, stream>>>(dcol, R, H, W, C, dx);
  MB200_CHECK_LAUNCH("mb200_col2im3_nhwc");
  return MB200_OK;
}

}  // extern "C"
