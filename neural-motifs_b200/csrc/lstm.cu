// Alternating highway LSTM for sm_100a: persistent recurrent kernels + hoisted projections.
//
// Replaces lib/lstm/highway_lstm_cuda/src/highway_lstm_kernel.cu of the reference:
//   highway_lstm_forward_ongpu :377-496 (per step: 2 cublasSgemm on fresh streams, 2
//   cudaDeviceSynchronize, elementWise_fp :108-160) and highway_lstm_backward_ongpu :162-375
//   (per step: elementWise_bp :46-104, 4 cublasSgemm + 1 Sgemv, 2 cudaDeviceSynchronize).
//
// B200 design (DESIGN.md "highway LSTM"):
//   * the input projection x_t * W_i does not depend on the recurrence, so it is hoisted out
//     of the time loop: ONE [T*B, In] x [In, 6H] GEMM per layer, written straight into the
//     `gates` buffer (which the recurrent kernel then overwrites in place with activations);
//   * the recurrence runs as ONE persistent cooperative kernel per layer: CTA i owns hidden
//     units [4i, 4i+4) and keeps the 20 matching columns of W_h in shared memory for the whole
//     sequence; per step it pulls h_{t-1} (L2, bypassing L1), does its slice of the matvec,
//     applies the gate math for its units and publishes h_t; one grid barrier per step,
//     no host round trip (the reference takes two device-wide host syncs per step);
//   * backward mirrors it: gate gradients for all steps are kept in a [T,B,6H] buffer so that
//     dX, dW_i, dW_h and db become four large GEMM/reductions after the time loop; only
//     dH_{t-1} = dG_t * W_h^T stays inside the persistent kernel.
// Slot/zero-state/direction conventions are the reference's: slot t+1 holds
// the output of time t, even layers run forward reading slot t, odd layers run backward reading
// slot (t+2)%(T+1); rows b >= covered(t) are never touched and stay zero.
#include "common.cuh"
#include "sgemm.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace {

constexpr int kThreads = 256;
constexpr int kUJ = 4;        // hidden units per CTA
constexpr int kTPU = 64;      // threads per unit
constexpr int kBTf = 32;      // batch rows per tile, forward (h rows of H floats)
constexpr int kBTb = 8;       // batch rows per tile, backward (dG rows of 5H floats)

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// Grid-wide barrier for the cooperative (co-resident) launch: monotonically increasing arrival
// counter in global memory, release by fence + atomic, acquire by ld.acquire spinning. Cheaper than
// cooperative_groups' grid.sync() for 128 CTAs, and it is the only thing on the per-step critical path.
__device__ unsigned int g_lstm_barrier[2];

__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
    } while (v < target);
  }
  __syncthreads();
}

struct FwdArgs {
  int H, B, T, dir, training;
  const float* P;        // [T,B,6H] input projection (aliases gates when training)
  const float* Wh;       // [H,5H]
  const float* bias;     // [5H]
  const float* dropout;  // [B,H]
  float* h;              // [T+1,B,H]
  float* c;              // [T+1,B,H]
  float* gates;          // [T,B,6H] or nullptr
  const int* lengths;    // DEV [B], descending
  unsigned int* barrier; // zeroed before launch
};

__device__ __forceinline__ int pick_ks(int rows) {
  int p = 1;
  while (p < rows) p <<= 1;
  int ks = kTPU / p;
  return ks < 2 ? 2 : (ks > 32 ? 32 : ks);
}

__global__ void __launch_bounds__(kThreads, 1) lstm_fwd_kernel(FwdArgs a) {
  extern __shared__ float smem[];
  const int H = a.H, B = a.B, T = a.T;
  const int ldh = H + 4;
  float* Ws = smem;                         // [H][5*kUJ]
  float* hs = Ws + (size_t)H * 5 * kUJ;     // [kBTf][ldh]
  int* lens = (int*)(hs + (size_t)kBTf * ldh);  // [B]
  const int tid = threadIdx.x;
  const int j0 = blockIdx.x * kUJ;
  const int u = tid / kTPU;                 // unit within CTA (warp-uniform)
  const int t64 = tid % kTPU;
  const int j = j0 + u;

  for (int idx = tid; idx < H * 5 * kUJ; idx += kThreads) {
    const int k = idx / (5 * kUJ), r = idx - k * 5 * kUJ;
    const int g = r / kUJ, uu = r - g * kUJ;
    Ws[idx] = a.Wh[(size_t)k * 5 * H + g * H + j0 + uu];
  }
  for (int idx = tid; idx < B; idx += kThreads) lens[idx] = a.lengths[idx];
  float bias_r[5];
#pragma unroll
  for (int g = 0; g < 5; ++g) bias_r[g] = a.bias[g * H + j];
  __syncthreads();

  int cov = (a.dir == 0) ? B : 0;
  const size_t BH = (size_t)B * H;
  for (int step = 0; step < T; ++step) {
    const int t = (a.dir == 0) ? step : T - 1 - step;
    int prev;
    if (a.dir == 0) { while (cov > 0 && lens[cov - 1] <= t) --cov; prev = t; }
    else { while (cov < B && lens[cov] > t) ++cov; prev = (t + 2) % (T + 1); }
    const float* hprev = a.h + (size_t)prev * BH;
    const float* cprev = a.c + (size_t)prev * BH;
    float* hout = a.h + (size_t)(t + 1) * BH;
    float* cout = a.c + (size_t)(t + 1) * BH;

    for (int b0 = 0; b0 < cov; b0 += kBTf) {
      const int rows = min(kBTf, cov - b0);
      // h_{t-1} tile -> shared (float4, L2 path: other SMs wrote it)
      const int vec_per_row = H / 4;
      for (int idx = tid; idx < rows * vec_per_row; idx += kThreads) {
        const int r = idx / vec_per_row, v = idx - r * vec_per_row;
        const float4 x = __ldcg((const float4*)(hprev + (size_t)(b0 + r) * H) + v);
        *(float4*)(hs + (size_t)r * ldh + 4 * v) = x;
      }
      __syncthreads();
      if (rows > 8) {
        // Large batch tile: each 8-lane group owns FOUR rows, so a k-step costs 4 + 5 shared loads for 20
        // FMAs (0.45 LDS/FMA instead of 1.2) — the B=256 microbenchmark shape is LDS-bound otherwise.
        constexpr int RB = 4, KS = 8;
        const int ks = t64 % KS, grp = t64 / KS;           // 8 groups x 4 rows = 32 rows
        const int r0 = grp * RB;
        float acc[RB][5];
#pragma unroll
        for (int q = 0; q < RB; ++q)
#pragma unroll
          for (int g = 0; g < 5; ++g) acc[q][g] = 0.f;
        const float* hbase = hs + (size_t)r0 * ldh;         // rows beyond `rows` hold stale data; never written back
        for (int k = ks; k < H; k += KS) {
          float hv[RB], wv[5];
          const float* w = Ws + (size_t)k * 5 * kUJ + u;
#pragma unroll
          for (int g = 0; g < 5; ++g) wv[g] = w[g * kUJ];
#pragma unroll
          for (int q = 0; q < RB; ++q) hv[q] = hbase[(size_t)q * ldh + k];
#pragma unroll
          for (int q = 0; q < RB; ++q)
#pragma unroll
            for (int g = 0; g < 5; ++g) acc[q][g] = fmaf(hv[q], wv[g], acc[q][g]);
        }
#pragma unroll
        for (int o = KS >> 1; o > 0; o >>= 1)
#pragma unroll
          for (int q = 0; q < RB; ++q)
#pragma unroll
            for (int g = 0; g < 5; ++g) acc[q][g] += __shfl_xor_sync(0xffffffffu, acc[q][g], o);
        // lanes ks = 0..3 of the group finish one row each
        float mine[5];
#pragma unroll
        for (int g = 0; g < 5; ++g) {
          float v = acc[0][g];
          v = (ks == 1) ? acc[1][g] : v; v = (ks == 2) ? acc[2][g] : v; v = (ks == 3) ? acc[3][g] : v;
          mine[g] = v;
        }
        const int r = r0 + ks;
        if (ks < RB && r < rows) {
          const int b = b0 + r;
          const float* P = a.P + ((size_t)t * B + b) * 6 * H + j;
          float p[6];
#pragma unroll
          for (int g = 0; g < 6; ++g) p[g] = __ldcg(P + (size_t)g * H);
          const float cp = __ldcg(cprev + (size_t)b * H + j);
          const float dp = a.dropout[(size_t)b * H + j];
          float gt[5];
#pragma unroll
          for (int g = 0; g < 5; ++g) gt[g] = (p[g] + mine[g]) + bias_r[g];
          const float in_gate = sigmoidf_(gt[0]);
          const float forget_gate = sigmoidf_(gt[1]);
          const float act_gate = tanhf(gt[2]);
          const float out_gate = sigmoidf_(gt[3]);
          const float r_gate = sigmoidf_(gt[4]);
          const float lin_gate = p[5];
          if (a.gates) {
            float* G = a.gates + ((size_t)t * B + b) * 6 * H + j;
            G[0] = in_gate; G[(size_t)H] = forget_gate; G[(size_t)2 * H] = act_gate;
            G[(size_t)3 * H] = out_gate; G[(size_t)4 * H] = r_gate; G[(size_t)5 * H] = lin_gate;
          }
          float val = (forget_gate * cp) + (in_gate * act_gate);
          cout[(size_t)b * H + j] = val;
          val = out_gate * tanhf(val);
          val = (float)((double)(val * r_gate) + (1.0 - (double)r_gate) * (double)lin_gate);
          val = val * dp;
          hout[(size_t)b * H + j] = val;
        }
      } else {
      const int KS = pick_ks(rows);
      const int ks = t64 % KS, slot = t64 / KS, nslots = kTPU / KS;
      const int iters = (rows + nslots - 1) / nslots;
      for (int m = 0; m < iters; ++m) {
        const int r = slot + m * nslots;
        const bool active = r < rows;
        const int rc = active ? r : 0;
        const int b = b0 + rc;
        const bool lead = active && ks == 0;
        float p[6], cp = 0.f, dp = 0.f;
        if (lead) {
          const float* P = a.P + ((size_t)t * B + b) * 6 * H + j;
#pragma unroll
          for (int g = 0; g < 6; ++g) p[g] = __ldcg(P + (size_t)g * H);
          cp = __ldcg(cprev + (size_t)b * H + j);
          dp = a.dropout[(size_t)b * H + j];
        }
        float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        const float* hrow = hs + (size_t)rc * ldh;
        // 4 k-steps per iteration, all 24 shared-memory loads issued before the FMAs (ncu: the
        // non-unrolled loop was LDS-latency bound); H / KS is a multiple of 4 (KS <= 32, H % 128 == 0) or
        // the tail loop finishes it.
        int k = ks;
        for (; k + 3 * KS < H; k += 4 * KS) {
          float hv[4], wv[4][5];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            hv[q] = hrow[k + q * KS];
            const float* w = Ws + (size_t)(k + q * KS) * 5 * kUJ + u;
#pragma unroll
            for (int g = 0; g < 5; ++g) wv[q][g] = w[g * kUJ];
          }
#pragma unroll
          for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int g = 0; g < 5; ++g) acc[g] = fmaf(hv[q], wv[q][g], acc[g]);
        }
        for (; k < H; k += KS) {
          const float hv = hrow[k];
          const float* w = Ws + (size_t)k * 5 * kUJ + u;
#pragma unroll
          for (int g = 0; g < 5; ++g) acc[g] = fmaf(hv, w[g * kUJ], acc[g]);
        }
        for (int o = KS >> 1; o > 0; o >>= 1) {
#pragma unroll
          for (int g = 0; g < 5; ++g) acc[g] += __shfl_xor_sync(0xffffffffu, acc[g], o);
        }
        if (lead) {
          // elementWise_fp, highway_lstm_kernel.cu:125-159
          float gt[5];
#pragma unroll
          for (int g = 0; g < 5; ++g) gt[g] = (p[g] + acc[g]) + bias_r[g];
          const float in_gate = sigmoidf_(gt[0]);
          const float forget_gate = sigmoidf_(gt[1]);
          const float act_gate = tanhf(gt[2]);
          const float out_gate = sigmoidf_(gt[3]);
          const float r_gate = sigmoidf_(gt[4]);
          const float lin_gate = p[5];
          if (a.gates) {
            float* G = a.gates + ((size_t)t * B + b) * 6 * H + j;
            G[0] = in_gate; G[(size_t)H] = forget_gate; G[(size_t)2 * H] = act_gate;
            G[(size_t)3 * H] = out_gate; G[(size_t)4 * H] = r_gate; G[(size_t)5 * H] = lin_gate;
          }
          float val = (forget_gate * cp) + (in_gate * act_gate);
          cout[(size_t)b * H + j] = val;
          val = out_gate * tanhf(val);
          // the reference mixes a double literal in here (:155): val*r + (1. - r)*lin
          val = (float)((double)(val * r_gate) + (1.0 - (double)r_gate) * (double)lin_gate);
          val = val * dp;
          hout[(size_t)b * H + j] = val;
        }
      }
      }
      __syncthreads();
    }
    grid_barrier(a.barrier, (unsigned int)(step + 1) * gridDim.x);
  }
}

struct BwdArgs {
  int H, B, T, dir;
  const float* out_grad;  // [T,B,H] upstream gradient of this layer's output
  const float* Wh;        // [H,5H]
  const float* h;         // [T+1,B,H]
  const float* c;         // [T+1,B,H]
  const float* gates;     // [T,B,6H] saved activations
  const float* dropout;   // [B,H]
  float* h_grad;          // [T+1,B,H] (zero-initialised)
  float* c_grad;          // [T+1,B,H] (zero-initialised)
  float* dG;              // [T,B,6H]  (zero-initialised) gate gradients of every step
  const int* lengths;
  unsigned int* barrier;  // zeroed before launch
};

__global__ void __launch_bounds__(kThreads, 1) lstm_bwd_kernel(BwdArgs a) {
  extern __shared__ float smem[];
  const int H = a.H, B = a.B, T = a.T;
  const int H5 = 5 * H;
  const int ldg = H5 + 4;
  float* Ws = smem;                          // [kUJ][5H]: rows k0..k0+3 of W_h
  float* gs = Ws + (size_t)kUJ * H5;         // [kBTb][ldg]
  int* lens = (int*)(gs + (size_t)kBTb * ldg);
  __shared__ float red[(kThreads / 32) * kBTb * kUJ];
  const int tid = threadIdx.x;
  const int j0 = blockIdx.x * kUJ;
  const int u = tid / kTPU;
  const int t64 = tid % kTPU;

  for (int idx = tid; idx < kUJ * H5; idx += kThreads) {
    const int uu = idx / H5, r = idx - uu * H5;
    Ws[idx] = a.Wh[(size_t)(j0 + uu) * H5 + r];
  }
  for (int idx = tid; idx < B; idx += kThreads) lens[idx] = a.lengths[idx];
  __syncthreads();

  // Backward through time runs opposite to the layer's forward direction (:198-211).
  int cov = (a.dir == 0) ? 0 : B;
  const size_t BH = (size_t)B * H;
  for (int step = 0; step < T; ++step) {
    const int t = (a.dir == 0) ? T - 1 - step : step;
    int prev, prevg;
    if (a.dir == 0) { while (cov < B && lens[cov] > t) ++cov; prevg = (t + 2) % (T + 1); prev = t; }
    else { while (cov > 0 && lens[cov - 1] <= t) --cov; prevg = t; prev = (t + 2) % (T + 1); }

    // (A) elementWise_bp (:46-104) for the units this CTA owns
    for (int idx = tid; idx < cov * kUJ; idx += kThreads) {
      const int b = idx / kUJ, j = j0 + (idx - b * kUJ);
      const size_t e = (size_t)b * H + j;
      float d_h = a.out_grad[(size_t)t * BH + e] + __ldcg(a.h_grad + (size_t)prevg * BH + e);
      d_h = d_h * a.dropout[e];
      const float* G = a.gates + ((size_t)t * B + b) * 6 * H + j;
      const float in_gate = G[0], forget_gate = G[(size_t)H], act_gate = G[(size_t)2 * H];
      const float out_gate = G[(size_t)3 * H], r_gate = G[(size_t)4 * H], lin_gate = G[(size_t)5 * H];
      const float c_out = a.c[(size_t)(t + 1) * BH + e];
      const float c_in = a.c[(size_t)prev * BH + e];
      const float th = tanhf(c_out);
      const float d_out = d_h * r_gate;
      const float d_c = d_out * out_gate * (1.f - th * th) + __ldcg(a.c_grad + (size_t)prevg * BH + e);
      const float h_prime = out_gate * th;
      float* D = a.dG + ((size_t)t * B + b) * 6 * H + j;
      D[0] = d_c * act_gate * in_gate * (1.f - in_gate);
      D[(size_t)H] = d_c * c_in * forget_gate * (1.f - forget_gate);
      D[(size_t)2 * H] = d_c * in_gate * (1.f - act_gate * act_gate);
      D[(size_t)3 * H] = d_out * th * out_gate * (1.f - out_gate);
      D[(size_t)4 * H] = d_h * (h_prime - lin_gate) * r_gate * (1.f - r_gate);
      D[(size_t)5 * H] = d_h * (1 - r_gate);
      a.c_grad[(size_t)(t + 1) * BH + e] = forget_gate * d_c;
    }
    grid_barrier(a.barrier, (unsigned int)(step + 1) * gridDim.x);
    // (B) h_grad[t+1][b][k] = sum_{5H} dG[t][b][:5H] * W_h[k][:]  for k in this CTA's units
    float* hg = a.h_grad + (size_t)(t + 1) * BH;
    for (int b0 = 0; b0 < cov; b0 += kBTb) {
      const int rows = min(kBTb, cov - b0);
      const int vec_per_row = H5 / 4;
      // all loads of a row batch are issued before the first shared-memory store (L2 latency ~1 us
      // per dependent round trip is the cost that matters here)
      for (int v0 = tid; v0 < vec_per_row; v0 += kThreads) {
        float4 x[kBTb];
#pragma unroll
        for (int r = 0; r < kBTb; ++r)
          if (r < rows) x[r] = __ldcg((const float4*)(a.dG + ((size_t)t * B + b0 + r) * 6 * H) + v0);
#pragma unroll
        for (int r = 0; r < kBTb; ++r)
          if (r < rows) *(float4*)(gs + (size_t)r * ldg + 4 * v0) = x[r];
      }
      __syncthreads();
      // All 256 threads split K = 5H; each keeps a [rows][4 units] accumulator tile, so a k-step costs
      // rows + 4 shared loads for 4*rows FMAs (ncu on the first version: one LDS pair per FMA in a
      // non-unrolled loop, 50 % of all stall samples). Then warp shuffles + one shared round reduce.
      float acc[kBTb][kUJ];
#pragma unroll
      for (int r = 0; r < kBTb; ++r)
#pragma unroll
        for (int q = 0; q < kUJ; ++q) acc[r][q] = 0.f;
      for (int k = tid; k < H5; k += kThreads) {
        float wv[kUJ], gv[kBTb];
#pragma unroll
        for (int q = 0; q < kUJ; ++q) wv[q] = Ws[(size_t)q * H5 + k];
#pragma unroll
        for (int r = 0; r < kBTb; ++r) gv[r] = (r < rows) ? gs[(size_t)r * ldg + k] : 0.f;
#pragma unroll
        for (int r = 0; r < kBTb; ++r)
#pragma unroll
          for (int q = 0; q < kUJ; ++q) acc[r][q] = fmaf(gv[r], wv[q], acc[r][q]);
      }
#pragma unroll
      for (int r = 0; r < kBTb; ++r)
        if (r < rows) {                       // uniform across the CTA
#pragma unroll
          for (int q = 0; q < kUJ; ++q) {
            float v = acc[r][q];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            acc[r][q] = v;
          }
        }
      const int lane_ = tid & 31, warp_ = tid >> 5;
      if (lane_ == 0) {
#pragma unroll
        for (int r = 0; r < kBTb; ++r)
#pragma unroll
          for (int q = 0; q < kUJ; ++q) red[(warp_ * kBTb + r) * kUJ + q] = acc[r][q];
      }
      __syncthreads();
      if (tid < rows * kUJ) {
        const int r = tid / kUJ, q = tid - r * kUJ;
        float v = 0.f;
#pragma unroll
        for (int wq = 0; wq < kThreads / 32; ++wq) v += red[(wq * kBTb + r) * kUJ + q];
        hg[(size_t)(b0 + r) * H + j0 + q] = v;
      }
      __syncthreads();
    }
    // next step's (A) reads only h_grad/c_grad entries this CTA wrote: a CTA barrier suffices
    __syncthreads();
  }
}

// column sums: out[c] += sum_r in[r*ld + c], r < rows, c < cols
__global__ void colsum_accum_kernel(const float* __restrict__ in, int rows, int cols, int ld,
                                    float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float s = 0.f;
  for (int r = blockIdx.y; r < rows; r += gridDim.y) s += in[(size_t)r * ld + c];
  atomicAdd(out + c, s);
}

size_t fwd_smem_bytes(int H, int B) { return ((size_t)H * 5 * kUJ + (size_t)kBTf * (H + 4) + B) * 4; }
size_t bwd_smem_bytes(int H, int B) { return ((size_t)kUJ * 5 * H + (size_t)kBTb * (5 * H + 4) + B) * 4; }

// The barrier words are __device__ globals: one instance PER DEVICE, so the address is cached per device (a process that
// drives a second GPU must not spin on the first one's counter). One counter per direction and device: launches of the
// same direction are serialised by the single compute stream this library runs on (they must not overlap on two streams).
unsigned int* barrier_ptr(int which, cudaStream_t stream) {
  static unsigned int* base[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!base[dev]) {
    void* p = nullptr;
    if (cudaGetSymbolAddress(&p, g_lstm_barrier) != cudaSuccess) return nullptr;
    base[dev] = (unsigned int*)p;
  }
  cudaMemsetAsync(base[dev] + which, 0, sizeof(unsigned int), stream);
  return base[dev] + which;
}

int check_coop(const void* fn, int grid, size_t smem) {
  int dev = 0, coop = 0, sms = 0, per_sm = 0;
  MB200_CHECK(cudaGetDevice(&dev));
  MB200_CHECK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  MB200_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (!coop) return MB200_ERR_UNSUPPORTED;
  MB200_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MB200_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kThreads, smem));
  if (per_sm * sms < grid) return MB200_ERR_UNSUPPORTED;
  return MB200_OK;
}

inline size_t weight_start(int layer, int In, int H) {
  // alternating_highway_lstm.py:212-221 / highway_lstm_kernel.cu:435
  if (layer == 0) return 0;
  return (size_t)6 * H * In + (size_t)5 * H * H + (size_t)(layer - 1) * 11 * H * H;
}

}  // namespace

extern "C" {

size_t mb200_highway_lstm_scratch_floats(int hiddenSize, int miniBatch, int seqLength) {
  return (size_t)seqLength * miniBatch * 6 * hiddenSize;
}

// Forward with device-side lengths and caller-provided scratch.
//   proj_scratch DEV [T,B,6H] floats, only used when gates == NULL (inference).
int mb200_highway_lstm_forward(int inputSize, int hiddenSize, int miniBatch, int numLayers, int seqLength,
                               const float* x, const int* lengths_dev, float* h_data, float* c_data,
                               const float* T, const float* bias, const float* dropout, float* gates,
                               float* proj_scratch, cudaStream_t stream) {
  const int H = hiddenSize, B = miniBatch, TT = seqLength;
  if (H <= 0 || B <= 0 || TT <= 0 || numLayers <= 0) return MB200_OK;
  if (H % kUJ != 0 || (H % 4) != 0) return MB200_ERR_UNSUPPORTED;
  if (!gates && !proj_scratch) return MB200_ERR_ARG;
  const int grid = H / kUJ;
  const size_t smem = fwd_smem_bytes(H, B);
  int rc = check_coop((const void*)lstm_fwd_kernel, grid, smem);
  if (rc != MB200_OK) return rc;
  const size_t acc = (size_t)(TT + 1) * B * H;
  for (int layer = 0; layer < numLayers; ++layer) {
    const int In = layer == 0 ? inputSize : H;
    const float* Wi = T + weight_start(layer, inputSize, H);
    const float* Wh = Wi + (size_t)6 * H * In;
    const float* xin = layer == 0 ? x : h_data + (size_t)(layer - 1) * acc + (size_t)B * H;  // slots 1..T
    float* P = gates ? gates + (size_t)layer * TT * B * 6 * H : proj_scratch;
    rc = mb200_sgemm_launch(0, 0, TT * B, 6 * H, In, 1.f, xin, In, Wi, 6 * H, 0.f, P, 6 * H, stream);
    if (rc != MB200_OK) return rc;
    FwdArgs a;
    a.H = H; a.B = B; a.T = TT; a.dir = layer % 2; a.training = gates != nullptr;
    a.P = P; a.Wh = Wh; a.bias = bias + (size_t)layer * 5 * H; a.dropout = dropout + (size_t)layer * B * H;
    a.h = h_data + (size_t)layer * acc; a.c = c_data + (size_t)layer * acc;
    a.gates = gates ? P : nullptr; a.lengths = lengths_dev;
    a.barrier = barrier_ptr(0, stream);
    if (!a.barrier) return MB200_ERR_CUDA;
    void* args[] = {&a};
    MB200_CHECK(cudaLaunchCooperativeKernel((const void*)lstm_fwd_kernel, dim3(grid), dim3(kThreads), args, smem, stream));
  }
  return MB200_OK;
}

// Backward with device-side lengths and caller-provided scratch dG [T,B,6H] (any contents).
// h_data_grad / c_data_grad [L,T+1,B,H] and x_grad must be zero on entry; T_grad / bias_grad are
// accumulated into. h_out_grad [L,T,B,H] is scratch for the inter-layer gradients.
int mb200_highway_lstm_backward(int inputSize, int hiddenSize, int miniBatch, int numLayers, int seqLength,
                                const float* out_grad, const int* lengths_dev, float* h_data_grad,
                                float* c_data_grad, const float* x, const float* h_data, const float* c_data,
                                const float* T, const float* gates_out, const float* dropout_in,
                                float* h_out_grad, float* x_grad, float* T_grad, float* bias_grad,
                                int do_weight_grad, float* dG_scratch, cudaStream_t stream) {
  const int H = hiddenSize, B = miniBatch, TT = seqLength;
  if (H <= 0 || B <= 0 || TT <= 0 || numLayers <= 0) return MB200_OK;
  if (H % kUJ != 0 || (H % 4) != 0) return MB200_ERR_UNSUPPORTED;
  if (!gates_out || !dG_scratch) return MB200_ERR_ARG;
  const int grid = H / kUJ;
  const size_t smem = bwd_smem_bytes(H, B);
  int rc = check_coop((const void*)lstm_bwd_kernel, grid, smem);
  if (rc != MB200_OK) return rc;
  const size_t acc = (size_t)(TT + 1) * B * H;
  const size_t BH = (size_t)B * H;
  const size_t TB = (size_t)TT * B;
  for (int layer = numLayers - 1; layer >= 0; --layer) {
    const int In = layer == 0 ? inputSize : H;
    const size_t ws = weight_start(layer, inputSize, H);
    const float* Wi = T + ws;
    const float* Wh = Wi + (size_t)6 * H * In;
    MB200_CHECK(cudaMemsetAsync(dG_scratch, 0, TB * 6 * H * sizeof(float), stream));
    BwdArgs a;
    a.H = H; a.B = B; a.T = TT; a.dir = layer % 2;
    a.out_grad = (layer == numLayers - 1) ? out_grad : h_out_grad + (size_t)layer * TB * H;
    a.Wh = Wh; a.h = h_data + (size_t)layer * acc; a.c = c_data + (size_t)layer * acc;
    a.gates = gates_out + (size_t)layer * TB * 6 * H; a.dropout = dropout_in + (size_t)layer * BH;
    a.h_grad = h_data_grad + (size_t)layer * acc; a.c_grad = c_data_grad + (size_t)layer * acc;
    a.dG = dG_scratch; a.lengths = lengths_dev;
    a.barrier = barrier_ptr(1, stream);
    if (!a.barrier) return MB200_ERR_CUDA;
    void* args[] = {&a};
    MB200_CHECK(cudaLaunchCooperativeKernel((const void*)lstm_bwd_kernel, dim3(grid), dim3(kThreads), args, smem, stream));
    // dX = dG * W_i^T   (:278-289), all steps at once
    float* dx = layer == 0 ? x_grad : h_out_grad + (size_t)(layer - 1) * TB * H;
    rc = mb200_sgemm_launch(0, 1, (int)TB, In, 6 * H, 1.f, dG_scratch, 6 * H, Wi, 6 * H, 0.f, dx, In, stream);
    if (rc != MB200_OK) return rc;
    if (do_weight_grad) {
      const float* xin = layer == 0 ? x : h_data + (size_t)(layer - 1) * acc + BH;
      // dW_i += X^T dG (:314-325)
      rc = mb200_sgemm_launch(1, 0, In, 6 * H, (int)TB, 1.f, xin, In, dG_scratch, 6 * H, 1.f, T_grad + ws, 6 * H, stream);
      if (rc != MB200_OK) return rc;
      // dW_h += Hprev^T dG[:, :5H] (:329-340). Even layers: Hprev(t) = slot t. Odd layers:
      // Hprev(t) = slot t+2 for t <= T-2 and the zero slot for t = T-1.
      float* dWh = T_grad + ws + (size_t)6 * H * In;
      if (layer % 2 == 0)
        rc = mb200_sgemm_launch(1, 0, H, 5 * H, (int)TB, 1.f, a.h, H, dG_scratch, 6 * H, 1.f, dWh, 5 * H, stream);
      else if (TT > 1)
        rc = mb200_sgemm_launch(1, 0, H, 5 * H, (int)((TT - 1) * (size_t)B), 1.f, a.h + 2 * BH, H, dG_scratch, 6 * H, 1.f, dWh, 5 * H, stream);
      if (rc != MB200_OK) return rc;
      // db += sum over (t,b) of dG[:, :5H] (:344-354)
      dim3 g(mb200_div_up(5 * H, 128), (unsigned)min((size_t)64, TB));
      colsum_accum_kernel<<<g, 128, 0, stream>>>(dG_scratch, (int)TB, 5 * H, 6 * H, bias_grad + (size_t)layer * 5 * H);
      MB200_CHECK_LAUNCH("colsum_accum_kernel");
    }
  }
  return MB200_OK;
}


// Single layer, projection already done (P = x_t W_i for every step, [T,B,6H]): just the persistent
// recurrence. h / c [T+1,B,H] zero on entry; gates may alias P (in place) or be NULL.
int mb200_highway_lstm_layer_forward(int hiddenSize, int miniBatch, int seqLength, int dir, const float* P,
                                     const float* Wh, const float* bias, const float* dropout, float* h, float* c,
                                     float* gates, const int* lengths_dev, cudaStream_t stream) {
  const int H = hiddenSize, B = miniBatch, TT = seqLength;
  if (H <= 0 || B <= 0 || TT <= 0) return MB200_OK;
  if (H % kUJ != 0 || (H % 4) != 0) return MB200_ERR_UNSUPPORTED;
  const int grid = H / kUJ;
  const size_t smem = fwd_smem_bytes(H, B);
  int rc = check_coop((const void*)lstm_fwd_kernel, grid, smem);
  if (rc != MB200_OK) return rc;
  FwdArgs a;
  a.H = H; a.B = B; a.T = TT; a.dir = dir; a.training = gates != nullptr;
  a.P = P; a.Wh = Wh; a.bias = bias; a.dropout = dropout; a.h = h; a.c = c; a.gates = gates; a.lengths = lengths_dev;
  a.barrier = barrier_ptr(0, stream);
  if (!a.barrier) return MB200_ERR_CUDA;
  void* args[] = {&a};
  MB200_CHECK(cudaLaunchCooperativeKernel((const void*)lstm_fwd_kernel, dim3(grid), dim3(kThreads), args, smem, stream));
  return MB200_OK;
}

// Single layer backward through time: fills dG [T,B,6H] (zeroed here), h_grad / c_grad [T+1,B,H]
// (zero on entry). The caller turns dG into dX, dW_i, dW_h, db with large GEMMs.
int mb200_highway_lstm_layer_backward(int hiddenSize, int miniBatch, int seqLength, int dir, const float* out_grad,
                                      const float* Wh, const float* h, const float* c, const float* gates,
                                      const float* dropout, float* h_grad, float* c_grad, float* dG,
                                      const int* lengths_dev, cudaStream_t stream) {
  const int H = hiddenSize, B = miniBatch, TT = seqLength;
  if (H <= 0 || B <= 0 || TT <= 0) return MB200_OK;
  if (H % kUJ != 0 || (H % 4) != 0) return MB200_ERR_UNSUPPORTED;
  const int grid = H / kUJ;
  const size_t smem = bwd_smem_bytes(H, B);
  int rc = check_coop((const void*)lstm_bwd_kernel, grid, smem);
  if (rc != MB200_OK) return rc;
  MB200_CHECK(cudaMemsetAsync(dG, 0, (size_t)TT * B * 6 * H * sizeof(float), stream));
  BwdArgs a;
  a.H = H; a.B = B; a.T = TT; a.dir = dir; a.out_grad = out_grad; a.Wh = Wh; a.h = h; a.c = c; a.gates = gates;
  a.dropout = dropout; a.h_grad = h_grad; a.c_grad = c_grad; a.dG = dG; a.lengths = lengths_dev;
  a.barrier = barrier_ptr(1, stream);
  if (!a.barrier) return MB200_ERR_CUDA;
  void* args[] = {&a};
  MB200_CHECK(cudaLaunchCooperativeKernel((const void*)lstm_bwd_kernel, dim3(grid), dim3(kThreads), args, smem, stream));
  return MB200_OK;
}

// ---- drop-in launchers (highway_lstm_kernel.h:7-9): host lengths, internal scratch.
void highway_lstm_forward_ongpu(int inputSize, int hiddenSize, int miniBatch, int numLayers, int seqLength,
                                float* x, int* lengths, float* h_data, float* c_data, float* tmp_i,
                                float* tmp_h, float* T, float* bias, float* dropout, float* gates,
                                int is_training, cudaStream_t stream, void* handle) {
  (void)tmp_i; (void)tmp_h; (void)handle;
  int* len_dev = nullptr; float* proj = nullptr;
  if (cudaMallocAsync((void**)&len_dev, sizeof(int) * miniBatch, stream) != cudaSuccess) { mb200_set_error("lstm fwd alloc", cudaGetLastError()); return; }
  cudaMemcpyAsync(len_dev, lengths, sizeof(int) * miniBatch, cudaMemcpyHostToDevice, stream);
  float* g = is_training ? gates : nullptr;
  if (!g) cudaMallocAsync((void**)&proj, sizeof(float) * mb200_highway_lstm_scratch_floats(hiddenSize, miniBatch, seqLength), stream);
  mb200_highway_lstm_forward(inputSize, hiddenSize, miniBatch, numLayers, seqLength, x, len_dev, h_data, c_data,
                             T, bias, dropout, g, proj, stream);
  if (proj) cudaFreeAsync(proj, stream);
  cudaFreeAsync(len_dev, stream);
}

void highway_lstm_backward_ongpu(int inputSize, int hiddenSize, int miniBatch, int numLayers, int seqLength,
                                 float* out_grad, int* lengths, float* h_data_grad, float* c_data_grad,
                                 float* x, float* h_data, float* c_data, float* T, float* gates_out,
                                 float* dropout_in, float* h_gates_grad, float* i_gates_grad,
                                 float* h_out_grad, float* x_grad, float* T_grad, float* bias_grad,
                                 int isTraining, int do_weight_grad, cudaStream_t stream, void* handle) {
  (void)h_gates_grad; (void)i_gates_grad; (void)isTraining; (void)handle;
  int* len_dev = nullptr; float* dG = nullptr;
  if (cudaMallocAsync((void**)&len_dev, sizeof(int) * miniBatch, stream) != cudaSuccess) { mb200_set_error("lstm bwd alloc", cudaGetLastError()); return; }
  cudaMemcpyAsync(len_dev, lengths, sizeof(int) * miniBatch, cudaMemcpyHostToDevice, stream);
  cudaMallocAsync((void**)&dG, sizeof(float) * mb200_highway_lstm_scratch_floats(hiddenSize, miniBatch, seqLength), stream);
  mb200_highway_lstm_backward(inputSize, hiddenSize, miniBatch, numLayers, seqLength, out_grad, len_dev,
                              h_data_grad, c_data_grad, x, h_data, c_data, T, gates_out, dropout_in,
                              h_out_grad, x_grad, T_grad, bias_grad, do_weight_grad, dG, stream);
  cudaFreeAsync(dG, stream);
  cudaFreeAsync(len_dev, stream);
}

}  // extern "C"
