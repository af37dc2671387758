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
//
// MC ("multicast"): data-parallel sharded update over NVSwitch. p_st / hi_st / lo_st are then the MULTICAST addresses of
// the symmetric parameter buffers: one `multimem.st` per 16 bytes lands the updated values in every rank's copy, i.e. the
// all-gather of the sharded update happens inside the store (lib/fused_optim.py "nvls" mode). Loads stay local.
//
// Launch shapes: the foreground one (8 CTAs of 256 per SM) owns the GPU; the BACKGROUND one (one CTA of 128 threads per SM,
// <= 80 registers) fits beside a resident tcgen05 GEMM CTA (320 threads x 168 registers + ~200 KB of shared memory leave
// 11.7 K registers per SM), so that a deferred update really runs underneath the next step's backbone: with the
// foreground shape the GEMM CTAs could not be placed until the whole update had drained (r02_trace_gaps_n2.log: a 0.98 ms
// hole in the compute stream).
__device__ __forceinline__ void multimem_st_v4(float* a, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void multimem_st_v2(void* a, const uint2& v) {
  asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1, %2};" ::"l"(a), "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)) : "memory");
}
__device__ __forceinline__ float4 multimem_ld_reduce_add_v4(const float* a) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(a) : "memory");
  return v;
}

template <bool SPLIT, bool MC, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
sgd_momentum_clip_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ buf,
                         __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                         float* p_st, __nv_bfloat16* hi_st, __nv_bfloat16* lo_st,
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
        if (MC) multimem_st_v4(p_st + 4 * i, pv[u]); else __stcs(p4 + i, pv[u]);
        __stcs(b4 + i, bv[u]);
        if (zero_grad) __stcs(g4 + i, make_float4(0.f, 0.f, 0.f, 0.f));
        if (SPLIT) {
          __nv_bfloat16 h[4], l[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) tc::split_bf16(pp[k], h[k], l[k]);
          if (MC) {
            multimem_st_v2(hi_st + 4 * i, *(const uint2*)h);
            multimem_st_v2(lo_st + 4 * i, *(const uint2*)l);
          } else {
            *(uint2*)(hi + 4 * i) = *(const uint2*)h;
            *(uint2*)(lo + 4 * i) = *(const uint2*)l;
          }
        }
      }
    }
  }
  if (MC) return;                            // the launcher takes multicast ranges in multiples of 4 elements only
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
// REDUCE (data-parallel "nvls" mode): x is the MULTICAST address of this rank's shard of the symmetric gradient buffer —
// `multimem.ld_reduce.add` returns the SUM over all ranks, formed inside the NVSwitch — which is written to `out` (the
// local copy of the shard) and squared: reduce-scatter + norm in one pass, no staging buffer.
template <bool REDUCE>
__global__ void sumsq_kernel(const float* __restrict__ x, float* __restrict__ out, long long n, double* __restrict__ acc) {
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
      if (REDUCE) v[u] = i < n4 ? multimem_ld_reduce_add_v4(x + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
      else v[u] = i < n4 ? __ldg(x4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (REDUCE && i < n4) ((float4*)out)[i] = v[u];
      s0 = fmaf(v[u].x, v[u].x, s0); s1 = fmaf(v[u].y, v[u].y, s1);
      s2 = fmaf(v[u].z, v[u].z, s2); s3 = fmaf(v[u].w, v[u].w, s3);
    }
  }
  if (!REDUCE)
    for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) s0 = fmaf(x[i], x[i], s0);
  double t = (double)s0 + (double)s1 + (double)s2 + (double)s3;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  __shared__ double red[32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    double b = 0.0;
    for (int j = 0; j < (int)(blockDim.x >> 5); ++j) b += red[j];
    atomicAdd(acc, b);
  }
}

// Data-parallel "ce" mode: the other ranks' copies of THIS rank's gradient shard have been landed by the copy engines in
// `stage` (nslots slots, `stride` floats apart); g[i] += sum_k stage[k][i], and the squared norm of the result.
__global__ void reduce_staged_sumsq_kernel(float* __restrict__ g, const float* __restrict__ stage, long long stride, int nslots,
                                           long long n, double* __restrict__ acc) {
  const long long n4 = n >> 2;
  const long long step = (long long)blockDim.x * gridDim.x;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  constexpr int U = 2;
  for (long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; i0 < n4; i0 += step * U) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * step;
      v[u] = i < n4 ? __ldcs((const float4*)g + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int k = 0; k < nslots; ++k) {
      const float4* sk = (const float4*)(stage + (long long)k * stride);
      float4 w[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long i = i0 + u * step;
        w[u] = i < n4 ? __ldcs(sk + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) { v[u].x += w[u].x; v[u].y += w[u].y; v[u].z += w[u].z; v[u].w += w[u].w; }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * step;
      if (i < n4) ((float4*)g)[i] = v[u];
      s0 = fmaf(v[u].x, v[u].x, s0); s1 = fmaf(v[u].y, v[u].y, s1);
      s2 = fmaf(v[u].z, v[u].z, s2); s3 = fmaf(v[u].w, v[u].w, s3);
    }
  }
  double t = (double)s0 + (double)s1 + (double)s2 + (double)s3;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  __shared__ double red[32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    double b = 0.0;
    for (int j = 0; j < (int)(blockDim.x >> 5); ++j) b += red[j];
    atomicAdd(acc, b);
  }
}

// one double into slot `idx` of a symmetric array on EVERY rank (multicast store): the ranks' partial squared norms
__global__ void bcast_slot_kernel(const double* __restrict__ v, double* slots_mc, int idx) {
  if (threadIdx.x == 0 && blockIdx.x == 0)
    asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(slots_mc + idx), "d"(*v) : "memory");
}

int g_background = 0;     // mb200_optim_set_background: launch shapes that co-reside with a tcgen05 GEMM CTA
constexpr int kBgThreads = 128;

}  // namespace

extern "C" int mb200_optim_set_background(int on) {
  g_background = on ? 1 : 0;
  return MB200_OK;
}

extern "C" int mb200_sumsq_accum(const float* x, long long n, double* acc, cudaStream_t stream) {
  if (n <= 0) return MB200_OK;
  if (((uintptr_t)x) & 15) return MB200_ERR_ARG;
  if (g_background) {
    sumsq_kernel<false><<<kNumSMs, kBgThreads, 0, stream>>>(x, nullptr, n, acc);
  } else {
    const int blocks = (int)min((long long)kNumSMs * 8, (n / 16 + 255) / 256 + 1);
    sumsq_kernel<false><<<blocks, 256, 0, stream>>>(x, nullptr, n, acc);
  }
  MB200_CHECK_LAUNCH("mb200_sumsq_accum");
  return MB200_OK;
}

// Data-parallel "nvls" mode, pass 1: out[i] = sum over ranks of the symmetric gradient buffer at the multicast address
// x_mc (this rank's shard), *acc += sum of squares of the result. n must be a multiple of 4, pointers 16-byte aligned.
extern "C" int mb200_dp_reduce_shard_sumsq(const float* x_mc, float* out, long long n, double* acc, cudaStream_t stream) {
  if (n <= 0) return MB200_OK;
  if (((((uintptr_t)x_mc) | ((uintptr_t)out)) & 15) || (n & 3)) return MB200_ERR_ARG;
  sumsq_kernel<true><<<kNumSMs, g_background ? kBgThreads : 256, 0, stream>>>(x_mc, out, n, acc);
  MB200_CHECK_LAUNCH("mb200_dp_reduce_shard_sumsq");
  return MB200_OK;
}

// Data-parallel "ce" mode, the reduce step: g[i] += sum over the nslots staged copies (stage + k * stride), *acc += |g|^2.
extern "C" int mb200_dp_reduce_staged_sumsq(float* g, const float* stage, long long stride, int nslots, long long n, double* acc,
                                            cudaStream_t stream) {
  if (n <= 0) return MB200_OK;
  if (((((uintptr_t)g) | ((uintptr_t)stage)) & 15) || (n & 3) || (stride & 3) || nslots < 0) return MB200_ERR_ARG;
  if (g_background) {
    reduce_staged_sumsq_kernel<<<kNumSMs, kBgThreads, 0, stream>>>(g, stage, stride, nslots, n, acc);
  } else {
    const int blocks = (int)min((long long)kNumSMs * 8, (n / 8 + 255) / 256 + 1);
    reduce_staged_sumsq_kernel<<<blocks, 256, 0, stream>>>(g, stage, stride, nslots, n, acc);
  }
  MB200_CHECK_LAUNCH("mb200_dp_reduce_staged_sumsq");
  return MB200_OK;
}

// *v -> slots[idx] on every rank (slots_mc: multicast address of a symmetric array of doubles)
extern "C" int mb200_dp_bcast_slot(const double* v, double* slots_mc, int idx, cudaStream_t stream) {
  if (idx < 0 || (((uintptr_t)slots_mc) & 7)) return MB200_ERR_ARG;
  bcast_slot_kernel<<<1, 32, 0, stream>>>(v, slots_mc, idx);
  MB200_CHECK_LAUNCH("mb200_dp_bcast_slot");
  return MB200_OK;
}

template <bool SPLIT, bool MC>
static int launch_sgd(float* params, float* grads, float* momentum_buf, void* hi, void* lo, float* p_st, void* hi_st, void* lo_st,
                      long long n, float lr, float momentum, float weight_decay, const float* total_norm_dev, float max_norm,
                      float grad_scale, int first_step, int zero_grad, cudaStream_t stream) {
  if (g_background)
    sgd_momentum_clip_kernel<SPLIT, MC, kBgThreads, 6><<<kNumSMs, kBgThreads, 0, stream>>>(
        params, grads, momentum_buf, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, p_st, (__nv_bfloat16*)hi_st, (__nv_bfloat16*)lo_st, n,
        lr, momentum, weight_decay, total_norm_dev, max_norm, grad_scale, first_step, zero_grad);
  else {
    const int blocks = (int)min((long long)kNumSMs * 8, (n / 4 + 255) / 256 + 1);
    sgd_momentum_clip_kernel<SPLIT, MC, 256, 2><<<blocks, 256, 0, stream>>>(
        params, grads, momentum_buf, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, p_st, (__nv_bfloat16*)hi_st, (__nv_bfloat16*)lo_st, n,
        lr, momentum, weight_decay, total_norm_dev, max_norm, grad_scale, first_step, zero_grad);
  }
  return MB200_OK;
}

extern "C" int mb200_sgd_momentum_clip_scaled(float* params, float* grads, float* momentum_buf, long long n, float lr,
                                              float momentum, float weight_decay, const float* total_norm_dev,
                                              float max_norm, float grad_scale, int first_step, int zero_grad,
                                              cudaStream_t stream) {
  if (n <= 0) return MB200_OK;
  if ((((uintptr_t)params) | ((uintptr_t)grads) | ((uintptr_t)momentum_buf)) & 15) return MB200_ERR_ARG;
  launch_sgd<false, false>(params, grads, momentum_buf, nullptr, nullptr, nullptr, nullptr, nullptr, n, lr, momentum, weight_decay,
                           total_norm_dev, max_norm, grad_scale, first_step, zero_grad, stream);
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
  launch_sgd<true, false>(params, grads, momentum_buf, hi, lo, nullptr, nullptr, nullptr, n, lr, momentum, weight_decay,
                          total_norm_dev, max_norm, grad_scale, first_step, zero_grad, stream);
  MB200_CHECK_LAUNCH("mb200_sgd_momentum_clip_split");
  return MB200_OK;
}

// Data-parallel "nvls" mode, pass 2: the fused update of THIS rank's shard (params / grads / momentum_buf: local
// addresses of the shard) whose results — the parameters and, when hi_mc / lo_mc are given, their bf16 operand pairs —
// are stored through the multicast addresses p_mc / hi_mc / lo_mc into every rank's copy. n % 4 == 0.
extern "C" int mb200_sgd_momentum_clip_mc(float* params, float* grads, float* momentum_buf, float* p_mc, void* hi_mc, void* lo_mc,
                                          long long n, float lr, float momentum, float weight_decay,
                                          const float* total_norm_dev, float max_norm, float grad_scale, int first_step,
                                          int zero_grad, cudaStream_t stream) {
  if (n <= 0) return MB200_OK;
  if (((((uintptr_t)params) | ((uintptr_t)grads) | ((uintptr_t)momentum_buf) | ((uintptr_t)p_mc)) & 15) || (n & 3)) return MB200_ERR_ARG;
  if ((hi_mc == nullptr) != (lo_mc == nullptr)) return MB200_ERR_ARG;
  if ((((uintptr_t)hi_mc) | ((uintptr_t)lo_mc)) & 7) return MB200_ERR_ARG;
  if (hi_mc)
    launch_sgd<true, true>(params, grads, momentum_buf, nullptr, nullptr, p_mc, hi_mc, lo_mc, n, lr, momentum, weight_decay,
                           total_norm_dev, max_norm, grad_scale, first_step, zero_grad, stream);
  else
    launch_sgd<false, true>(params, grads, momentum_buf, nullptr, nullptr, p_mc, nullptr, nullptr, n, lr, momentum, weight_decay,
                            total_norm_dev, max_norm, grad_scale, first_step, zero_grad, stream);
  MB200_CHECK_LAUNCH("mb200_sgd_momentum_clip_mc");
  return MB200_OK;
}

extern "C" int mb200_sgd_momentum_clip(float* params, float* grads, float* momentum_buf, long long n, float lr,
                                       float momentum, float weight_decay, const float* total_norm_dev,
                                       float max_norm, int first_step, int zero_grad, cudaStream_t stream) {
  return mb200_sgd_momentum_clip_scaled(params, grads, momentum_buf, n, lr, momentum, weight_decay, total_norm_dev,
                                        max_norm, 1.f, first_step, zero_grad, stream);
}
