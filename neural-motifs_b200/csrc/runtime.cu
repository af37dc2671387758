// Error bookkeeping and version probe for libmotifs_b200.so.
#include "common.cuh"
#include <string.h>

static thread_local char g_err[512] = "";
int g_mb200_sm_budget = kNumSMs;

void mb200_set_error(const char* what, cudaError_t err) {
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(err));
}

extern "C" {

const char* mb200_last_error() { return g_err; }

// ABI version of include/motifs_b200.h this library implements.
int mb200_abi_version() { return 1; }

// SM budget of the persistent tcgen05 kernels (see common.cuh); n is clamped to [16, 148] and rounded down to even.
// Returns the previous budget.
int mb200_set_sm_budget(int n) {
  const int old = g_mb200_sm_budget;
  if (n > kNumSMs) n = kNumSMs;
  if (n < 16) n = 16;
  g_mb200_sm_budget = n & ~1;
  return old;
}

// cudaMemsetAsync(ptr, 0, bytes): zero-fill by the memset engine, no SM involved (a fill KERNEL queued beside the
// persistent tcgen05 GEMMs would hold up their CTAs; lib/fused_optim.py zeroes gradient shards with this).
int mb200_zero_async(void* ptr, long long bytes, cudaStream_t stream) {
  if (bytes <= 0) return MB200_OK;
  MB200_CHECK(cudaMemsetAsync(ptr, 0, (size_t)bytes, stream));
  return MB200_OK;
}

// Compiled architecture (100 => sm_100a). Lets the host fail loudly on a mismatched device.
int mb200_compiled_arch() { return 100; }

// Returns 1 when the current device is a compute-capability 10.x part, 0 otherwise,
// MB200_ERR_CUDA when no device is usable.
int mb200_device_ok() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return MB200_ERR_CUDA;
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return MB200_ERR_CUDA;
  return p.major == 10 ? 1 : 0;
}

}  // extern "C"
