// Fused global-norm clip + weight decay + momentum SGD update over a flat parameter buffer.
// Replaces, for the trainable parameters of the relation model, the caller-side sequence of
// models/train_rels.py:145-150 — clip_grad_norm (lib/pytorch_misc.py:416-459, one host sync per
// parameter there) followed by torch.optim.SGD.step (momentum 0.9, weight decay, no dampening,
// no nesterov) — with ONE pass over (param, grad, momentum): 20 bytes per parameter instead of
// the ~44 the foreach kernels move, and the gradient is zeroed in the same pass.
#include "common.cuh"
#include "tc_common.cuh"

namespace {

// SPLIT: also emit the bf16 (hi, lo) pair of the UPDATED parameter at the same flat index — for every weight matrix whose
// row length is a multiple of 64 that IS the K-major operand layout of the tcgen05 GEMM, so the per-step re-split of the
// trainable weights (fc6 / fc7 copies: 2.2 GB of traffic, 0.3 ms on the compute stream) rides along with the update.
template <bool SPLIT>
__global__ void sgd_momentum_clip_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ buf,
                                         __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                         long long n, float lr, float momentum, float weight_decay,
                                         const float* __restrict__ total_norm, float max_norm, float grad_scale,
                                         int first_step, int zero_grad) {
  // g holds grad_scale^-1 times the gradient (data parallel: the all-reduced SUM, grad_scale = 1/world);
  // *total_norm is the norm of the SCALED gradient.
  float coef = 1.f;
  if (total_norm) {                       // clip_coef = max_norm / (norm + 1e-6), applied when < 1
    const float c = max_norm / (*total_norm + 1e-6f);
    coef = c < 1.f ? c : 1.f;
  }
  coef *= grad_scale;
  const long long n4 = n >> 2;
  const long long stride = (long long)blockDim.x * gridDim.x;
  float4* p4 = (float4*)p; float4* g4 = (float4*)g; float4* b4 = (float4*)buf;
  constexpr int U = 4;                     // 12 independent 16-byte loads in flight per thread before any store
  for (long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; i0 < n4; i0 += stride * U) {
    float4 pv[U], gv[U], bv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i < n4) {
        pv[u] = __ldcs(p4 + i); gv[u] = __ldcs(g4 + i);
        bv[u] = first_step ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldcs(b4 + i);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i < n4) {
        float* pp = (float*)&pv[u]; float* gg = (float*)&gv[u]; float* bb = (float*)&bv[u];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float d = gg[k] * coef;
          d = fmaf(weight_decay, pp[k], d);                      // d_p = g + wd * p
          bb[k] = first_step ? d : fmaf(momentum, bb[k], d);     // buf = momentum * buf + d_p
          pp[k] = fmaf(-lr, bb[k], pp[k]);                       // p -= lr * buf
        }
        __stcs(p4 + i, pv[u]); __stcs(b4 + i, bv[u]);
        if (zero_grad) __stcs(g4 + i, make_float4(0.f, 0.f, 0.f, 0.f));
        if (SPLIT) {
          __nv_bfloat16 h[4], l[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) tc::split_bf16(pp[k], h[k], l[k]);
          *(uint2*)(hi + 4 * i) = *(const uint2*)h;
          *(uint2*)(lo + 4 * i) = *(const uint2*)l;
        }
      }
    }
  }
  for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    float d = g[i] * coef;
    d = fmaf(weight_decay, p[i], d);
    const float b = first_step ? d : fmaf(momentum, buf[i], d);
    buf[i] = b;
    p[i] = fmaf(-lr, b, p[i]);
    if (zero_grad) g[i] = 0.f;
    if (SPLIT) tc::split_bf16(p[i], hi[i], lo[i]);
  }
}

// sum of squares of a flat fp32 buffer, accumulated in double: fp32 per thread (a few hundred terms),
// double across the block and across blocks (one atomic per block).
__global__ void sumsq_kernel(const float* __restrict__ x, long long n, double* __restrict__ acc) {
  const long long n4 = n >> 2;
  const long long stride = (long long)blockDim.x * gridDim.x;
  const float4* x4 = (const float4*)x;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  constexpr int U = 4;
  for (long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; i0 < n4; i0 += stride * U) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      v[u] = i < n4 ? __ldg(x4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      s0 = fmaf(v[u].x, v[u].x, s0); s1 = fmaf(v[u].y, v[u].y, s1);
      s2 = fmaf(v[u].z, v[u].z, s2); s3 = fmaf(v[u].w, v[u].w, s3);
    }
  }
  for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) s0 = fmaf(x[i], x[i], s0);
  double t = (double)s0 + (double)s1 + (double)s2 + (double)s3;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  __shared__ double red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    double b = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) b += red[j];
    atomicAdd(acc, b);
  }
}

}  // namespace

extern "C" int mb200_sumsq_accum(const float* x, long long n, double* acc, cudaStream_t stream) {
  if (n <= 0) return MB200_OK;
  if (((uintptr_t)x) & 15) return MB200_ERR_ARG;
  const int blocks = (int)min((long long)kNumSMs * 8, (n / 16 + 255) / 256 + 1);
  sumsq_kernel<<<blocks, 256, 0, stream>>>(x, n, acc);
  MB200_CHECK_LAUNCH("mb200_sumsq_accum");
  return MB200_OK;
}

extern "C" int mb200_sgd_momentum_clip_scaled(float* params, float* grads, float* momentum_buf, long long n, float lr,
                                              float momentum, float weight_decay, const float* total_norm_dev,
                                              float max_norm, float grad_scale, int first_step, int zero_grad,
                                              cudaStream_t stream) {
  if (n <= 0) return MB200_OK;
  if ((((uintptr_t)params) | ((uintptr_t)grads) | ((uintptr_t)momentum_buf)) & 15) return MB200_ERR_ARG;
  const int blocks = (int)min((long long)kNumSMs * 8, (n / 4 + 255) / 256 + 1);
  sgd_momentum_clip_kernel<false><<<blocks, 256, 0, stream>>>(params, grads, momentum_buf, nullptr, nullptr, n, lr, momentum,
                                                              weight_decay, total_norm_dev, max_norm, grad_scale, first_step,
                                                              zero_grad);
  MB200_CHECK_LAUNCH("mb200_sgd_momentum_clip");
  return MB200_OK;
}

extern "C" int mb200_sgd_momentum_clip_split(float* params, float* grads, float* momentum_buf, void* hi, void* lo, long long n,
                                             float lr, float momentum, float weight_decay, const float* total_norm_dev,
                                             float max_norm, float grad_scale, int first_step, int zero_grad,
                                             cudaStream_t stream) {
  if (n <= 0) return MB200_OK;
  if ((((uintptr_t)params) | ((uintptr_t)grads) | ((uintptr_t)momentum_buf)) & 15) return MB200_ERR_ARG;
  if ((((uintptr_t)hi) | ((uintptr_t)lo)) & 7) return MB200_ERR_ARG;
  const int blocks = (int)min((long long)kNumSMs * 8, (n / 4 + 255) / 256 + 1);
  sgd_momentum_clip_kernel<true><<<blocks, 256, 0, stream>>>(params, grads, momentum_buf, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo,
                                                             n, lr, momentum, weight_decay, total_norm_dev, max_norm,
                                                             grad_scale, first_step, zero_grad);
  MB200_CHECK_LAUNCH("mb200_sgd_momentum_clip_split");
  return MB200_OK;
}

extern "C" int mb200_sgd_momentum_clip(float* params, float* grads, float* momentum_buf, long long n, float lr,
                                       float momentum, float weight_decay, const float* total_norm_dev,
                                       float max_norm, int first_step, int zero_grad, cudaStream_t stream) {
  return mb200_sgd_momentum_clip_scaled(params, grads, momentum_buf, n, lr, momentum, weight_decay, total_norm_dev,
                                        max_norm, 1.f, first_step, zero_grad, stream);
}
