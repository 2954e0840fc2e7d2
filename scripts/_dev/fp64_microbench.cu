// Developer aid: FP64 pipe microbenchmarks on one SM (clock64 around unrolled instruction streams).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_microbench fp64_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int ILP>
__global__ void k_dfma(double* out, long long* cyc, int iters, double x0) {
  double a[ILP];
  for (int i = 0; i < ILP; i++) a[i] = x0 + i + threadIdx.x;
  const double m = 1.0000001, c = 1e-9;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 16; u++)
#pragma unroll
      for (int i = 0; i < ILP; i++) a[i] = __fma_rn(a[i], m, c);
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < ILP; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int ILP>
__global__ void k_dmma(double* out, long long* cyc, int iters, double x0) {
  double c0[ILP], c1[ILP];
  for (int i = 0; i < ILP; i++) { c0[i] = x0 + i; c1[i] = x0 - i; }
  double a = 1e-3 * threadIdx.x, b = 1e-3;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 16; u++)
#pragma unroll
      for (int i = 0; i < ILP; i++) dmma(c0[i], c1[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < ILP; i++) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_rcp(double* out, long long* cyc, int iters, double x0) {
  double a = x0 + threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 16; u++) { double x; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a)); a = x; }
  }
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_shfl(double* out, long long* cyc, int iters, double x0) {
  double a = x0 + threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 16; u++) a = __shfl_sync(0xffffffffu, a, (threadIdx.x + 1) & 31);
  }
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_ffma(float* out, long long* cyc, int iters, float x0) {
  float a = x0 + threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 16; u++) a = __fmaf_rn(a, 1.0000001f, 1e-9f);
  }
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double* out; long long* cyc; long long h[8];
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 64);
  const int iters = 256;
#define RUN(name, kern, threads, ninstr)                                                       \
  kern<<<1, threads>>>((decltype(out))out, cyc, iters, 1.0); cudaDeviceSynchronize();          \
  kern<<<1, threads>>>((decltype(out))out, cyc, iters, 1.0); cudaDeviceSynchronize();          \
  cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);                                               \
  printf("%-34s threads %4d: %8.2f cycles per warp-instruction (%s)\n", name, threads, (double)h[0] / (iters * 16.0 * (ninstr)), cudaGetErrorString(cudaGetLastError()));
  for (int th : {32, 128, 256, 512}) {
    RUN("DFMA dependent (ILP 1)", k_dfma<1>, th, 1);
    RUN("DFMA ILP 4", k_dfma<4>, th, 4);
    RUN("DFMA ILP 8", k_dfma<8>, th, 8);
    RUN("DFMA ILP 16", k_dfma<16>, th, 16);
    RUN("DMMA dependent (ILP 1)", k_dmma<1>, th, 1);
    RUN("DMMA ILP 4", k_dmma<4>, th, 4);
    RUN("DMMA ILP 8", k_dmma<8>, th, 8);
  }
  { float* fo = (float*)out; k_ffma<<<1, 32>>>(fo, cyc, iters, 1.0f); cudaDeviceSynchronize(); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("FFMA dependent: %.2f cycles\n", (double)h[0] / (iters * 16.0)); }
  k_rcp<<<1, 32>>>(out, cyc, iters, 1.5); cudaDeviceSynchronize(); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("MUFU.RCP64H dependent: %.2f cycles\n", (double)h[0] / (iters * 16.0));
  k_shfl<<<1, 32>>>(out, cyc, iters, 1.5); cudaDeviceSynchronize(); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("SHFL.f64 dependent: %.2f cycles\n", (double)h[0] / (iters * 16.0));
  return 0;
}
