// b2_cuda.h -- launch / shared-memory macros.  Under nvcc these are the CUDA constructs; with
// -DB2_EMULATE (tests/hostsim only) the same kernel sources run on the CPU emulator so that
// `pytest -m "not gpu"` can exercise them.  The product library is always the nvcc build.
#pragma once
#include <cstdint>
#include <cstdio>

#ifdef B2_EMULATE
#include "cuda_emu.h"
#define B2_LAUNCH(kern, grid, block, smem, stream, ...) \
  emu::launch(kern, dim3(grid), dim3(block), (size_t)(smem), __VA_ARGS__)
#define B2_DYN_SMEM(name) unsigned char* name = emu::S().dyn_smem
#define B2_UNROLL
#else
#include <cuda_runtime.h>
#define B2_LAUNCH(kern, grid, block, smem, stream, ...) \
  kern<<<grid, block, smem, stream>>>(__VA_ARGS__)
#define B2_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define B2_UNROLL _Pragma("unroll")
#endif

#define B2_CUDA_OK(expr)                                                                  \
  do {                                                                                    \
    cudaError_t e_ = (expr);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      snprintf(b2::g_last_error, sizeof(b2::g_last_error), "%s:%d: %s -> %s", __FILE__,   \
               __LINE__, #expr, cudaGetErrorString(e_));                                  \
      return -1;                                                                          \
    }                                                                                     \
  } while (0)

namespace b2 {
// one message buffer per host thread: handles driven from different threads (one rank per thread)
// never read each other's errors
extern thread_local char g_last_error[512];
}
