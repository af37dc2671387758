// Shared helpers for the motifs_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define MB200_OK 1               // the reference launchers return 1 on success
#define MB200_ERR_ARG 0          // roi_align_cuda.c:19-22 returns 0 on a bad shape
#define MB200_ERR_CUDA (-1)      // superset: CUDA failure (the reference prints / exit(-1))
#define MB200_ERR_UNSUPPORTED (-2)

// Last CUDA error string, readable through mb200_last_error().
extern "C" const char* mb200_last_error();
void mb200_set_error(const char* what, cudaError_t err);

#define MB200_CHECK_LAUNCH(what)                                   \
  do {                                                             \
    cudaError_t e__ = cudaGetLastError();                          \
    if (e__ != cudaSuccess) {                                      \
      mb200_set_error(what, e__);                                  \
      return MB200_ERR_CUDA;                                       \
    }                                                              \
  } while (0)

#define MB200_CHECK(call)                                          \
  do {                                                             \
    cudaError_t e__ = (call);                                      \
    if (e__ != cudaSuccess) {                                      \
      mb200_set_error(#call, e__);                                 \
      return MB200_ERR_CUDA;                                       \
    }                                                              \
  } while (0)

static inline int mb200_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

constexpr int kNumSMs = 148;  // B200

// SMs the persistent tensor-core kernels may occupy (<= kNumSMs, even). While a gradient all-reduce is in flight NCCL's
// channel CTAs hold SMs for milliseconds; a persistent grid of 148 CTAs with 200 KB of shared memory each then runs its
// last CTAs in a SECOND wave (they cannot co-reside with NCCL's), i.e. the kernel takes twice as long. The host lowers the
// budget for launches that may overlap a collective (mb200_set_sm_budget) and restores it afterwards.
extern int g_mb200_sm_budget;
