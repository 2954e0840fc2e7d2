// mma.cuh -- FP64 tensor-core fragment op shared by the single-system and the batched kernels.
#pragma once
#include "b2_cuda.h"

namespace b2 {

#ifdef B2_TIMING
__device__ long long b2_dbg[64];
#define B2_TICK(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) b2_dbg[i] = clock64(); } while (0)
#define B2_ACC(i, t0) do { if (threadIdx.x == 0 && blockIdx.x == 0) b2_dbg[i] += clock64() - (t0); } while (0)
#else
#define B2_TICK(i)
#define B2_ACC(i, t0)
#endif

// D(8x8) += A(8x4) * B(4x8) on the FP64 tensor cores (PTX mma.m8n8k4.f64 = SASS DMMA).
// Fragments, g = lane/4, t = lane%4:  a = A[g][t], b = B[t][g], c0/c1 = C[g][2t + {0,1}].
__device__ __forceinline__ void dmma_8x8x4(double& c0, double& c1, double a, double b) {
#ifdef B2_EMULATE
  const int lane = threadIdx.x & 31;
  const int row = lane >> 2, cp = (lane & 3) * 2;
  double s0 = 0.0, s1 = 0.0;
  for (int k = 0; k < 4; k++) {
    const double av = __shfl_sync(0xffffffffu, a, row * 4 + k);
    const double b0 = __shfl_sync(0xffffffffu, b, cp * 4 + k);
    const double b1 = __shfl_sync(0xffffffffu, b, (cp + 1) * 4 + k);
    s0 += av * b0;
    s1 += av * b1;
  }
  c0 += s0;
  c1 += s1;
#else
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
#endif
}

}  // namespace b2
