// VGG stem: conv1_1 (3 -> 64 channels, 3x3, pad 1) + bias + ReLU, fp32 NCHW image in, NHWC (hi, lo) bf16
// pair out — one direct-convolution kernel instead of im2col (538 MB written + read) followed by a GEMM.
// K = 27 is too thin for the tensor pipe (a k-block is 64 wide) and the layer is 0.6 % of the backbone's
// FLOPs, so it runs as exact fp32 FMAs on the CUDA cores: a lane owns one pixel and its 27 taps in
// registers, weights are 16-byte broadcast reads from shared memory, and the 64 outputs of 32 pixels are
// staged through a swizzled shared tile so that every global store instruction writes four complete
// 128-byte NHWC pixels. Replaces the first nn.Conv2d of lib/object_detector.py:110-127 (cuDNN in the reference).
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int kStemWarps = 4;
constexpr int kCout = 64;

__global__ void __launch_bounds__(kStemWarps * 32)
stem_conv3x3_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                    int B, int H, int W, int relu, __nv_bfloat16* __restrict__ yhi, __nv_bfloat16* __restrict__ ylo) {
  __shared__ __align__(16) float s_w[27][kCout];       // [tap*3 + c][cout]
  __shared__ __align__(16) float s_b[kCout];
  __shared__ __align__(16) uint4 s_hi[kStemWarps][32 * 8];   // per warp: 32 pixels x 8 chunks of 8 bf16
  __shared__ __align__(16) uint4 s_lo[kStemWarps][32 * 8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 27 * kCout; i += kStemWarps * 32) {
    const int k = i / kCout, o = i - k * kCout;          // k = tap*3 + c  (tap = kh*3 + kw)
    const int tap = k / 3, c = k - tap * 3;
    s_w[k][o] = w[(o * 3 + c) * 9 + tap];                // OIHW
  }
  if (tid < kCout) s_b[tid] = bias ? bias[tid] : 0.f;
  __syncthreads();

  const int segs = (W + 31) / 32;
  const long long items = (long long)B * H * segs;
  for (long long item = (long long)blockIdx.x * kStemWarps + warp; item < items; item += (long long)gridDim.x * kStemWarps) {
    const int seg = (int)(item % segs);
    const int h = (int)((item / segs) % H);
    const int b = (int)(item / ((long long)segs * H));
    const int w0 = seg * 32, px = w0 + lane;
    // 27 taps of this lane's pixel (zero padding)
    float in[27];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int yy = h + tap / 3 - 1, xx = px + tap % 3 - 1;
      const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;
#pragma unroll
      for (int c = 0; c < 3; ++c)
        in[tap * 3 + c] = ok ? __ldg(x + (((size_t)b * 3 + c) * H + yy) * W + xx) : 0.f;
    }
#pragma unroll 1
    for (int g = 0; g < 8; ++g) {
      float acc[8];
      const float4 b0 = *(const float4*)&s_b[g * 8], b1 = *(const float4*)&s_b[g * 8 + 4];
      acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w; acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
#pragma unroll
      for (int k = 0; k < 27; ++k) {
        const float4 w0v = *(const float4*)&s_w[k][g * 8], w1v = *(const float4*)&s_w[k][g * 8 + 4];
        const float v = in[k];
        acc[0] = fmaf(v, w0v.x, acc[0]); acc[1] = fmaf(v, w0v.y, acc[1]); acc[2] = fmaf(v, w0v.z, acc[2]); acc[3] = fmaf(v, w0v.w, acc[3]);
        acc[4] = fmaf(v, w1v.x, acc[4]); acc[5] = fmaf(v, w1v.y, acc[5]); acc[6] = fmaf(v, w1v.z, acc[6]); acc[7] = fmaf(v, w1v.w, acc[7]);
      }
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float a0 = acc[2 * j], a1 = acc[2 * j + 1];
        if (relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); }
        const __nv_bfloat16 h0 = __float2bfloat16_rn(a0), h1 = __float2bfloat16_rn(a1);
        __nv_bfloat162 hv; hv.x = h0; hv.y = h1;
        __nv_bfloat162 lv = __floats2bfloat162_rn(a0 - __bfloat162float(h0), a1 - __bfloat162float(h1));
        hi[j] = *reinterpret_cast<uint32_t*>(&hv); lo[j] = *reinterpret_cast<uint32_t*>(&lv);
      }
      const int slot = lane * 8 + (g ^ (lane & 7));        // XOR swizzle: conflict-free 16-byte stores
      s_hi[warp][slot] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      s_lo[warp][slot] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
    __syncwarp();
    // coalesced store: 32 pixels x 128 B are contiguous in NHWC; each instruction writes 4 whole pixels
    const int valid = min(32, W - w0);
    uint4* dh = (uint4*)(yhi + (((size_t)b * H + h) * W + w0) * kCout);
    uint4* dl = (uint4*)(ylo + (((size_t)b * H + h) * W + w0) * kCout);
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int i = it * 32 + lane;                        // 16-byte chunk index in the warp's tile
      const int p = i >> 3, cpos = i & 7;
      if (p < valid) {
        const int gch = cpos ^ (p & 7);                    // channel group stored at this slot
        dh[p * 8 + gch] = s_hi[warp][i];
        dl[p * 8 + gch] = s_lo[warp][i];
      }
    }
    __syncwarp();
  }
}

}  // namespace

extern "C" int mb200_conv3x3_stem_split(const float* x_nchw, const float* w_oihw, const float* bias, int B, int H,
                                        int W, int Cout, int relu, void* yhi, void* ylo, cudaStream_t stream) {
  if (B <= 0 || H <= 0 || W <= 0) return MB200_OK;
  if (Cout != kCout) return MB200_ERR_UNSUPPORTED;
  const long long items = (long long)B * H * ((W + 31) / 32);
  const int blocks = (int)min((items + kStemWarps - 1) / kStemWarps, (long long)kNumSMs * 16);
  stem_conv3x3_kernel<<<blocks, kStemWarps * 32, 0, stream>>>(x_nchw, w_oihw, bias, B, H, W, relu,
                                                              (__nv_bfloat16*)yhi, (__nv_bfloat16*)ylo);
  MB200_CHECK_LAUNCH("mb200_conv3x3_stem_split");
  return MB200_OK;
}
