// Microbenchmark: FP64 throughput of mma.sync.m8n8k4.f64 versus plain DFMA on this GPU (decides how the d >= 30
// GEMMs of the generic path are issued).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_rate dmma_rate.cu
// Measured on this pool's B200 (1965 MHz): DMMA 37.1 TFLOP/s, DFMA 33.9 TFLOP/s at 8-32 warps per CTA, two CTAs per SM.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dmma_kernel(double* out, int iters) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void dfma_kernel(double* out, int iters) {
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = i;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double* out; cudaMalloc(&out, 148 * 8 * 1024 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int warps = 4; warps <= 32; warps *= 2) {
    const int iters = 20000, grid = 148 * 2;
    float ms;
    dmma_kernel<<<grid, warps * 32>>>(out, iters); cudaDeviceSynchronize();
    cudaEventRecord(e0); dmma_kernel<<<grid, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    const double fl_mma = 2.0 * 8 * 8 * 4 * 8.0 * iters * warps * grid;
    printf("dmma  warps/CTA %2d (2 CTA/SM): %.3f ms  %.2f TFLOP/s\n", warps, ms, fl_mma / ms * 1e-9);
    dfma_kernel<<<grid, warps * 32>>>(out, iters); cudaDeviceSynchronize();
    cudaEventRecord(e0); dfma_kernel<<<grid, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    const double fl_fma = 2.0 * 16 * 32.0 * iters * warps * grid;
    printf("dfma  warps/CTA %2d (2 CTA/SM): %.3f ms  %.2f TFLOP/s\n", warps, ms, fl_fma / ms * 1e-9);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
