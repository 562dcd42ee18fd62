// DMMA throughput as a function of warps per SM sub-partition (is one warp enough to saturate the pipe?)
#include <cuda_runtime.h>
#include <cstdio>
constexpr int ITERS = 8192;
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NACC>
__global__ void k(double* out, double a, double b) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  if (s == 1.2345) out[0] = s;
}
template <int NACC>
void run(double* d, int wps) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = 148 * wps;  // 128-thread blocks: one warp per sub-partition per block
  k<NACC><<<grid, 128>>>(d, 1.0000001, 1e-9); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<NACC><<<grid, 128>>>(d, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("{\"nacc\": %d, \"warps_per_smsp\": %d, \"tflops\": %.2f}\n", NACC, wps, (double)grid * 4 * ITERS * NACC * 512.0 / ms * 1e-9);
}
int main() {
  double* d; cudaMalloc(&d, 64);
  for (int w : {1, 2, 3, 4, 6, 8}) run<8>(d, w);
  for (int w : {1, 2, 4}) run<16>(d, w);
  for (int w : {1, 2, 4}) run<4>(d, w);
  return 0;
}
