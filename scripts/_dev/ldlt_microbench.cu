// Developer aid: cycles of cta_ldlt64 on one SM, alone and next to a CTA that saturates the FP64
// tensor pipe (what k_front_dag's diagonal task sees).  build: see build line in the file header
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../cannoles_b200/csrc [-DKERNELS_H='"/tmp/old_kernels.cuh"'] -o ldlt_microbench ldlt_microbench.cu
#include <cstdio>
#include <vector>
#include <cmath>
#ifndef KERNELS_H
#define KERNELS_H "kernels.cuh"
#endif
#include KERNELS_H
namespace b2 { char g_last_error[512]; }
using namespace b2;
__global__ void __launch_bounds__(256) k_one(double* A, int nb, long long* cyc, int* flags) {
  extern __shared__ double sm[];
  double* S = sm; double* Wd = S + NB * DIAG_LD;
  for (int idx = threadIdx.x; idx < NB * NB; idx += 256) { int i = idx % NB, j = idx / NB; S[i + j * DIAG_LD] = A[idx]; }
  __syncthreads();
  long long t0 = clock64();
  cta_ldlt64<256>(S, nb, Wd, flags);
  long long t1 = clock64();
#ifdef REPS
  for (int rep = 0; rep < REPS; rep++) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < NB * NB; idx += 256) { int i = idx % NB, j = idx / NB; S[i + j * DIAG_LD] = A[idx]; }
    __syncthreads();
    cta_ldlt64<256>(S, nb, Wd, flags);
  }
#endif
#ifdef TWICE
  __syncthreads();
  for (int idx = threadIdx.x; idx < NB * NB; idx += 256) { int i = idx % NB, j = idx / NB; S[i + j * DIAG_LD] = A[idx]; }
  __syncthreads();
  if (threadIdx.x == 0) { b2_dbg[10] = 0; b2_dbg[11] = 0; }
  t0 = clock64();
  cta_ldlt64<256>(S, nb, Wd, flags);
  t1 = clock64();
  if (threadIdx.x == 0) { cyc[1] = b2_dbg[10]; cyc[2] = b2_dbg[11]; }
#endif
  for (int idx = threadIdx.x; idx < NB * NB; idx += 256) { int i = idx % NB, j = idx / NB; A[idx] = S[i + j * DIAG_LD]; }
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void __launch_bounds__(256) k_hog(double* out, volatile int* stop) {
  double c0[4] = {1, 2, 3, 4}, c1[4] = {1, 2, 3, 4}; double a = 1e-3 * threadIdx.x, b = 1e-3;
  while (!*stop) {
#pragma unroll
    for (int u = 0; u < 64; u++) dmma_8x8x4(c0[u & 3], c1[u & 3], a, b);
  }
  out[blockIdx.x * 256 + threadIdx.x] = c0[0] + c0[1] + c0[2] + c0[3] + c1[0] + c1[1] + c1[2] + c1[3];
}
int main() {
  std::vector<double> A(NB * NB), L(NB * NB);
  // SPD-ish quasi-definite test block: A = M M^T + I with sign flips on half the diagonal
  for (int j = 0; j < NB; j++) for (int i = 0; i < NB; i++) A[i + j * NB] = (i == j) ? (i % 2 ? -70.0 : 70.0) : sin(0.37 * i + 1.3 * j) + sin(0.37 * j + 1.3 * i);
  double* dA; long long* dc; int* df; double* dout; int* dstop;
  cudaMalloc(&dA, NB * NB * 8); cudaMalloc(&dc, 64); cudaMalloc(&df, 64); cudaMalloc(&dout, 148 * 2 * 256 * 8);
  cudaMallocHost(&dstop, 4); *dstop = 0;
  cudaFuncSetAttribute(k_one, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  long long h;
  for (int rep = 0; rep < 3; rep++) {
    cudaMemcpy(dA, A.data(), NB * NB * 8, cudaMemcpyHostToDevice);
    k_one<<<1, 256, (NB * DIAG_LD + 2 * NB * 8) * 8>>>(dA, NB, dc, df); cudaDeviceSynchronize();
    long long h3[3]; cudaMemcpy(h3, dc, 24, cudaMemcpyDeviceToHost); h = h3[0];
    printf("alone: %lld cycles, panel part %lld, next-panel update %lld (%s)\n", h, h3[1], h3[2], cudaGetErrorString(cudaGetLastError()));
  }
  cudaMemcpy(L.data(), dA, NB * NB * 8, cudaMemcpyDeviceToHost);
  // check: reconstruct A from L D L^T
  double err = 0;
  for (int i = 0; i < NB; i++) for (int j = 0; j <= i; j++) {
    double s = 0; for (int k = 0; k <= j; k++) { double lik = (k == i) ? 1.0 : L[i + k * NB], ljk = (k == j) ? 1.0 : L[j + k * NB]; s += lik * L[k + k * NB] * ljk; }
    err = fmax(err, fabs(s - A[i + j * NB]));
  }
  printf("max |L D L^T - A| = %.3e\n", err);
  cudaStream_t s1, s2; cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking); fflush(stdout);
  return 0;
  k_hog<<<148, 256, 0, s1>>>(dout, dstop);
  for (int rep = 0; rep < 3; rep++) {
    cudaMemcpyAsync(dA, A.data(), NB * NB * 8, cudaMemcpyHostToDevice, s2);
    k_one<<<1, 256, (NB * DIAG_LD + 2 * NB * 8) * 8, s2>>>(dA, NB, dc, df);
    cudaMemcpyAsync(&h, dc, 8, cudaMemcpyDeviceToHost, s2); cudaStreamSynchronize(s2);
    printf("next to a DMMA-saturating CTA: %lld cycles (%s)\n", h, cudaGetErrorString(cudaGetLastError())); fflush(stdout);
  }
  *dstop = 1; cudaDeviceSynchronize();
  return 0;
}
