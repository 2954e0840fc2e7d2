// measure.cu -- roofline denominators measured on the box (MEASURED_PEAKS.json carries no FP64
// entry): cuBLAS DGEMM throughput and a plain device copy.  Library calls are fine here: this
// is measurement plumbing, not the hot path.
#include <cublas_v2.h>
#include <cuda_runtime.h>

#include <cstdio>

#include "../../include/cannoles_b200.h"
#include "b2_cuda.h"

extern "C" {

int b2_measure_dgemm(int n, int reps, double* tflops_own, double* tflops_cublas) {
  if (tflops_own) *tflops_own = 0.0;
  if (tflops_cublas) *tflops_cublas = 0.0;
  double *A = nullptr, *B = nullptr, *Cm = nullptr;
  size_t bytes = (size_t)n * n * sizeof(double);
  B2_CUDA_OK(cudaMalloc((void**)&A, bytes));
  B2_CUDA_OK(cudaMalloc((void**)&B, bytes));
  B2_CUDA_OK(cudaMalloc((void**)&Cm, bytes));
  B2_CUDA_OK(cudaMemset(A, 0, bytes));
  B2_CUDA_OK(cudaMemset(B, 0, bytes));
  B2_CUDA_OK(cudaMemset(Cm, 0, bytes));
  cublasHandle_t hb;
  if (cublasCreate(&hb) != CUBLAS_STATUS_SUCCESS) {
    snprintf(b2::g_last_error, sizeof(b2::g_last_error), "cublasCreate failed");
    return -1;
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const double one = 1.0, zero = 0.0;
  double best = 1e30;
  for (int r = 0; r < reps + 2; r++) {
    cudaEventRecord(e0);
    cublasDgemm(hb, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &one, A, n, B, n, &zero, Cm, n);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r >= 2 && ms < best) best = ms;
  }
  if (tflops_cublas) *tflops_cublas = 2.0 * n * (double)n * n / (best * 1e-3) / 1e12;
  cublasDestroy(hb);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(A); cudaFree(B); cudaFree(Cm);
  return 0;
}

int b2_measure_hbm(size_t bytes, int reps, double* gbs_copy) {
  if (gbs_copy) *gbs_copy = 0.0;
  char *a = nullptr, *b = nullptr;
  B2_CUDA_OK(cudaMalloc((void**)&a, bytes));
  B2_CUDA_OK(cudaMalloc((void**)&b, bytes));
  B2_CUDA_OK(cudaMemset(a, 1, bytes));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 1e30;
  for (int r = 0; r < reps + 2; r++) {
    cudaEventRecord(e0);
    cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice, 0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r >= 2 && ms < best) best = ms;
  }
  if (gbs_copy) *gbs_copy = 2.0 * (double)bytes / (best * 1e-3) / 1e9;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(a); cudaFree(b);
  return 0;
}

}  // extern "C"
