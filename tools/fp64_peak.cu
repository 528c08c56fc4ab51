// Micro-benchmark: FP64 DFMA and DMMA (mma.sync.m8n8k4.f64) peak on this GPU, and LDS.128 broadcast cost.
// Used for the FP64 roofline denominator of the vertical-implicit kernel (MEASURED_PEAKS.json has no FP64 figure).
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma_kernel(double* out, int iters, double a, double b) {
  double c[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) { c[i][0] = threadIdx.x * 1e-3; c[i][1] = i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  double* d;
  cudaMalloc(&d, sizeof(double) * 148 * 8 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int threads : {256, 512, 1024}) {
    const int blocks = 148 * (2048 / threads), iters = 20000;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      dfma_kernel<<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double fl = 2.0 * 8 * iters * double(blocks) * threads;
      if (rep) printf("DFMA threads/block %4d: %.2f TFLOP/s (%.3f ms)\n", threads, fl / ms * 1e-9, ms);
      cudaEventRecord(e0);
      dmma_kernel<<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
      fl = 2.0 * 256 * 4 * iters * double(blocks) * threads / 32;
      if (rep) printf("DMMA threads/block %4d: %.2f TFLOP/s (%.3f ms)\n", threads, fl / ms * 1e-9, ms);
    }
  }
  printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
