// 3x3 / stride 2 / pad 1 max-pool (forward with arg-max, backward by gather) on NCHW fp32 — the pool
// inside the union-box mask branch (`nn.MaxPool2d(kernel_size=3, stride=2, padding=1)`,
// lib/get_union_boxes.py:34). torch's NCHW max_pool backward takes 0.94 ms on the [1536,256,14,14]
// tensor of the SGCls step (ncu launch list); this pair is bandwidth-bound: backward is a deterministic
// gather (each input reads the <= 4 windows covering it), no atomics.
#include "common.cuh"

namespace {

__global__ void maxpool3s2_fwd_kernel(const float* __restrict__ x, long long planes, int H, int W, int Ho, int Wo,
                                      float* __restrict__ y, unsigned char* __restrict__ arg) {
  const long long total = planes * Ho * Wo;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)blockDim.x * gridDim.x) {
    const int ox = (int)(i % Wo);
    const int oy = (int)((i / Wo) % Ho);
    const long long p = i / ((long long)Wo * Ho);
    const float* xp = x + p * H * W;
    float best = -INFINITY; int bi = 0;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int yy = 2 * oy - 1 + dy, xx = 2 * ox - 1 + dx;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
          const float v = __ldg(xp + yy * W + xx);
          if (v > best || (v != v)) { best = v; bi = dy * 3 + dx; }   // first maximum, NaN propagates (as ATen)
        }
      }
    y[i] = best; arg[i] = (unsigned char)bi;
  }
}

__global__ void maxpool3s2_bwd_kernel(const float* __restrict__ gy, const unsigned char* __restrict__ arg,
                                      long long planes, int H, int W, int Ho, int Wo, float* __restrict__ gx) {
  const long long total = planes * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)blockDim.x * gridDim.x) {
    const int xx = (int)(i % W);
    const int yy = (int)((i / W) % H);
    const long long p = i / ((long long)W * H);
    const float* gp = gy + p * Ho * Wo;
    const unsigned char* ap = arg + p * Ho * Wo;
    float s = 0.f;
    // windows (oy, ox) with 2*o - 1 <= coord <= 2*o + 1
    const int oy0 = (yy) / 2, oy1 = (yy + 1) / 2;      // equal for even yy
    const int ox0 = (xx) / 2, ox1 = (xx + 1) / 2;
    for (int oy = oy0; oy <= oy1; ++oy) {
      if (oy >= Ho) continue;
      const int dy = yy - (2 * oy - 1);
      for (int ox = ox0; ox <= ox1; ++ox) {
        if (ox >= Wo) continue;
        const int dx = xx - (2 * ox - 1);
        if (ap[oy * Wo + ox] == dy * 3 + dx) s += gp[oy * Wo + ox];
      }
    }
    gx[i] = s;
  }
}

inline int pool_blocks(long long total) {
  long long b = (total + 255) / 256;
  const long long cap = (long long)kNumSMs * 32;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

extern "C" {

int mb200_maxpool3s2_forward(const float* x, long long planes, int H, int W, float* y, unsigned char* argmax,
                             cudaStream_t stream) {
  if (planes <= 0) return MB200_OK;
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  maxpool3s2_fwd_kernel<<<pool_blocks(planes * Ho * Wo), 256, 0, stream>>>(x, planes, H, W, Ho, Wo, y, argmax);
  MB200_CHECK_LAUNCH("mb200_maxpool3s2_forward");
  return MB200_OK;
}

int mb200_maxpool3s2_backward(const float* grad_y, const unsigned char* argmax, long long planes, int H, int W,
                              float* grad_x, cudaStream_t stream) {
  if (planes <= 0) return MB200_OK;
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  maxpool3s2_bwd_kernel<<<pool_blocks(planes * H * W), 256, 0, stream>>>(grad_y, argmax, planes, H, W, Ho, Wo, grad_x);
  MB200_CHECK_LAUNCH("mb200_maxpool3s2_backward");
  return MB200_OK;
}

}  // extern "C"
