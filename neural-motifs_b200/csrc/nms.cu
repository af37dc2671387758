// Greedy NMS for sm_100a, fully on device.
//
// Replaces lib/fpn/nms/src/cuda/nms_kernel.cu of the reference: `nms_kernel` :33-75 (the
// 64x64-tile IoU bitmask) and the host half of `ApplyNMSGPU` :88-131 (cudaMalloc, blocking
// D2H copy of the N x ceil(N/64) mask, serial CPU suppression loop, cudaFree).
//
// Bit-exactness contract: the IoU expression is the one of `devIoU` :23-31, evaluated in
// fp32 with IEEE division and compared with strict `>`; boxes arrive already sorted by
// score (functions/nms.py:37-40) and the greedy scan runs in index order, so keep lists are
// identical to the reference's.
//
// Differences by design: (1) only the upper triangle of the mask is computed (the reference
// computes all tiles but reads only j >= i/64 words, :124); (2) the greedy reduce runs on
// the device in 64-box chunks, so nothing but the final keep list ever crosses PCIe;
// (3) a segmented entry point runs many independent problems (images x classes) in one launch.
#include "common.cuh"

namespace {

constexpr int kTile = 64;  // == sizeof(unsigned long long) * 8, nms_kernel.cu:21

// devIoU, nms_kernel.cu:23-31, with the exact operation order nvcc emits for the reference
// source on sm_100a (checked in its SASS): Sa is a plain product, Sa+Sb is contracted to
// fma(wb, hb, Sa), the intersection is a plain product and the division is IEEE.
__device__ __forceinline__ float dev_iou(const float4 a, const float4 b) {
  const float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
  const float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
  const float width = fmaxf(right - left + 1, 0.f), height = fmaxf(bottom - top + 1, 0.f);
  const float interS = __fmul_rn(width, height);
  const float Sa = __fmul_rn(a.z - a.x + 1, a.w - a.y + 1);
  const float SaSb = __fmaf_rn(b.z - b.x + 1, b.w - b.y + 1, Sa);
  return __fdiv_rn(interS, SaSb - interS);
}

// Segment s covers boxes [seg_off[s], seg_off[s+1]) ; its mask starts at mask_off[s]
// (in 64-bit words) and has ceil(n/64) words per row.
// grid: (max col tiles, max row tiles, segments)
__global__ void __launch_bounds__(kTile)
nms_mask_kernel(const float4* __restrict__ boxes, const int* __restrict__ seg_off,
                const long long* __restrict__ mask_off, float thresh,
                unsigned long long* __restrict__ mask) {
  const int s = blockIdx.z;
  const int beg = seg_off[s];
  const int n = seg_off[s + 1] - beg;
  const int row_start = blockIdx.y, col_start = blockIdx.x;
  const int col_blocks = (n + kTile - 1) / kTile;
  if (row_start >= col_blocks || col_start >= col_blocks || col_start < row_start) return;
  const int row_size = min(n - row_start * kTile, kTile);
  const int col_size = min(n - col_start * kTile, kTile);
  __shared__ float4 block_boxes[kTile];
  const float4* b = boxes + beg;
  if ((int)threadIdx.x < col_size) block_boxes[threadIdx.x] = b[kTile * col_start + threadIdx.x];
  __syncthreads();
  if ((int)threadIdx.x < row_size) {
    const int cur = kTile * row_start + threadIdx.x;
    const float4 cur_box = b[cur];
    unsigned long long t = 0;
    const int start = (row_start == col_start) ? threadIdx.x + 1 : 0;
    for (int i = start; i < col_size; ++i)
      if (dev_iou(cur_box, block_boxes[i]) > thresh) t |= 1ULL << i;
    mask[mask_off[s] + (long long)cur * col_blocks + col_start] = t;
  }
}

// Greedy suppression (nms_kernel.cu:113-128) on the device. One CTA per segment.
// Thread j owns remv word j (j = threadIdx.x + k*blockDim). Boxes are scanned in chunks of
// 64: warp 0 resolves the chunk's diagonal word serially (64 dependent steps, registers
// only), then all threads OR the rows of the kept boxes into their remv words.
constexpr int kReduceThreads = 256;
constexpr int kMaxWordsPerThread = 8;  // supports n <= 256*8*64 = 131072 boxes per segment

__global__ void __launch_bounds__(kReduceThreads)
nms_reduce_kernel(const unsigned long long* __restrict__ mask, const int* __restrict__ seg_off,
                  const long long* __restrict__ mask_off, int max_keep,
                  int* __restrict__ keep, int* __restrict__ num_keep) {
  const int s = blockIdx.x;
  const int beg = seg_off[s];
  const int n = seg_off[s + 1] - beg;
  const int col_blocks = (n + kTile - 1) / kTile;
  const unsigned long long* m = mask + mask_off[s];
  int* keep_s = keep + beg;   // keep list of segment s lives at its box offset (local indices)

  __shared__ unsigned long long s_keepmask;   // kept boxes of the current chunk
  __shared__ unsigned long long s_remv_cur;   // remv word of the current chunk
  __shared__ int s_count;
  __shared__ int s_rows[kTile];
  unsigned long long remv[kMaxWordsPerThread];
#pragma unroll
  for (int k = 0; k < kMaxWordsPerThread; ++k) remv[k] = 0ULL;
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();

  for (int chunk = 0; chunk < col_blocks; ++chunk) {
    // owner of remv[chunk] publishes it
    if ((chunk % kReduceThreads) == (int)threadIdx.x) s_remv_cur = remv[chunk / kReduceThreads];
    __syncthreads();
    if (threadIdx.x < 32) {
      const int lane = threadIdx.x;
      const int base = chunk * kTile;
      const int csize = min(n - base, kTile);
      // lane holds the diagonal words of boxes base+lane and base+lane+32
      unsigned long long d0 = (lane < csize) ? m[(long long)(base + lane) * col_blocks + chunk] : 0ULL;
      unsigned long long d1 = (lane + 32 < csize) ? m[(long long)(base + lane + 32) * col_blocks + chunk] : 0ULL;
      unsigned long long r = s_remv_cur;
      unsigned long long kept = 0ULL;
      for (int i = 0; i < csize; ++i) {
        const unsigned long long di = __shfl_sync(0xffffffffu, (i < 32) ? d0 : d1, i & 31);
        if (!(r & (1ULL << i))) { kept |= 1ULL << i; r |= di; }
      }
      if (lane == 0) {
        s_keepmask = kept;
        // append kept indices in order (local to the segment)
        int c = s_count;
        unsigned long long k = kept;
        while (k) {
          const int i = __ffsll((long long)k) - 1;
          k &= k - 1;
          if (c < max_keep) keep_s[c] = base + i;
          ++c;
        }
        s_count = c;
      }
    }
    __syncthreads();
    const unsigned long long kept = s_keepmask;
    if (kept) {
      // OR the rows of the kept boxes into the words this thread owns (only words > chunk
      // matter). Loads are issued 8 at a time so their L2 latencies overlap.
      const int nk = __popcll(kept);
      if (threadIdx.x < 64) {   // compact list of kept rows of this chunk
        const unsigned long long below = kept & ((1ULL << threadIdx.x) - 1ULL);
        if (kept & (1ULL << threadIdx.x)) s_rows[__popcll(below)] = chunk * kTile + threadIdx.x;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < kMaxWordsPerThread; ++k) {
        const int w = threadIdx.x + k * kReduceThreads;
        if (w > chunk && w < col_blocks) {
          unsigned long long acc = remv[k];
          int i = 0;
          for (; i + 8 <= nk; i += 8) {
            unsigned long long v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = m[(long long)s_rows[i + q] * col_blocks + w];
#pragma unroll
            for (int q = 0; q < 8; ++q) acc |= v[q];
          }
          for (; i < nk; ++i) acc |= m[(long long)s_rows[i] * col_blocks + w];
          remv[k] = acc;
        }
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) num_keep[s] = min(s_count, max_keep);
}

}  // namespace

extern "C" {

// Words of scratch the caller must provide for a set of segments with the given sizes.
long long mb200_nms_mask_words(const int* seg_sizes_host, int num_segments) {
  long long w = 0;
  for (int s = 0; s < num_segments; ++s) {
    const long long n = seg_sizes_host[s];
    w += n * ((n + kTile - 1) / kTile);
  }
  return w;
}

// Segmented on-device NMS.
//   boxes_dev      [total,4] fp32, each segment already sorted by descending score
//   seg_off_dev    [S+1] int32 box offsets, mask_off_dev [S] int64 word offsets into mask_dev
//   max_seg        largest segment size (host value, sizes the grid)
//   keep_dev       [total] int32: segment s writes its kept LOCAL indices at keep_dev[seg_off[s]..]
//   num_keep_dev   [S] int32
// Returns MB200_OK / error code. Nothing is copied to the host.
int mb200_nms_segmented(const float* boxes_dev, const int* seg_off_dev, const long long* mask_off_dev,
                        int num_segments, int max_seg, float thresh, int max_keep,
                        unsigned long long* mask_dev, int* keep_dev, int* num_keep_dev,
                        cudaStream_t stream) {
  if (num_segments <= 0) return MB200_OK;
  if (max_seg > kReduceThreads * kMaxWordsPerThread * kTile) return MB200_ERR_UNSUPPORTED;
  if (max_seg > 0) {
    const int tiles = mb200_div_up(max_seg, kTile);
    dim3 grid(tiles, tiles, num_segments);
    if (tiles > 65535 || num_segments > 65535) return MB200_ERR_UNSUPPORTED;
    nms_mask_kernel<<<grid, kTile, 0, stream>>>((const float4*)boxes_dev, seg_off_dev, mask_off_dev,
                                                thresh, mask_dev);
    MB200_CHECK_LAUNCH("nms_mask_kernel");
  }
  nms_reduce_kernel<<<num_segments, kReduceThreads, 0, stream>>>(mask_dev, seg_off_dev, mask_off_dev,
                                                                 max_keep, keep_dev, num_keep_dev);
  MB200_CHECK_LAUNCH("nms_reduce_kernel");
  return MB200_OK;
}

// Drop-in for nms_kernel.h:1-2 / nms_kernel.cu:88: boxes on the device (sorted), keep list on
// the HOST, returns the number kept. Synchronous like the reference. Unlike the reference
// it runs on the caller's device without cudaSetDevice side effects unless device_id differs.
int ApplyNMSGPU(int* keep_out, const float* boxes_dev, const int boxes_num, float nms_overlap_thresh,
                int device_id) {
  if (boxes_num <= 0) return 0;
  int cur = -1;
  if (cudaGetDevice(&cur) != cudaSuccess) return MB200_ERR_CUDA;
  if (device_id >= 0 && cur != device_id) MB200_CHECK(cudaSetDevice(device_id));
  const long long col_blocks = (boxes_num + kTile - 1) / kTile;
  const size_t mask_bytes = sizeof(unsigned long long) * boxes_num * col_blocks;
  // one allocation: [seg_off(2 int) pad][mask_off (1 ll)][num_keep][keep n][mask]
  char* scratch = nullptr;
  const size_t head = 64;
  const size_t keep_bytes = ((size_t)boxes_num * sizeof(int) + 63) / 64 * 64;
  MB200_CHECK(cudaMalloc(&scratch, head + keep_bytes + mask_bytes));
  struct Head { int seg_off[2]; int num_keep; int pad; long long mask_off; } h;
  h.seg_off[0] = 0; h.seg_off[1] = boxes_num; h.num_keep = 0; h.pad = 0; h.mask_off = 0;
  cudaStream_t stream = 0;  // legacy default stream, as the reference (nms_kernel.cu:102)
  cudaError_t e = cudaMemcpyAsync(scratch, &h, sizeof(h), cudaMemcpyHostToDevice, stream);
  int rc = MB200_OK;
  if (e == cudaSuccess) {
    Head* dh = (Head*)scratch;
    rc = mb200_nms_segmented(boxes_dev, dh->seg_off, &dh->mask_off, 1, boxes_num, nms_overlap_thresh,
                             boxes_num, (unsigned long long*)(scratch + head + keep_bytes),
                             (int*)(scratch + head), &dh->num_keep, stream);
  }
  int num = 0;
  if (e == cudaSuccess && rc == MB200_OK) {
    e = cudaMemcpy(&num, &((Head*)scratch)->num_keep, sizeof(int), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && num > 0)
      e = cudaMemcpy(keep_out, scratch + head, sizeof(int) * num, cudaMemcpyDeviceToHost);
  }
  cudaFree(scratch);
  if (device_id >= 0 && cur != device_id) cudaSetDevice(cur);
  if (e != cudaSuccess) { mb200_set_error("ApplyNMSGPU", e); return MB200_ERR_CUDA; }
  if (rc != MB200_OK) return rc;
  return num;
}

}  // extern "C"
