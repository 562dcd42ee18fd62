// Legacy mma.sync int8 (m16n8k32, s32 accumulate) issue rate on sm_100a.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
constexpr int ITERS = 8192;
__device__ __forceinline__ void imma(int (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__global__ void k(int* out, uint32_t a0, uint32_t b0) {
  int c[8][4];
  uint32_t a[4] = {a0, a0 + 1, a0 + 2, a0 + 3}, b[2] = {b0, b0 + 1};
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) imma(c[i], a, b);
  }
  int s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 12345) out[0] = s;
}
int main() {
  int* d; cudaMalloc(&d, 64);
  for (int wps : {1, 2, 4, 8}) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * wps;
    k<<<grid, 128>>>(d, 0x01010101u, 0x01020304u); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<<<grid, 128>>>(d, 0x01010101u, 0x01020304u); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double n = (double)grid * 4 * ITERS * 8;
    printf("{\"warps_per_smsp\": %d, \"tops\": %.1f, \"cycles_per_imma_per_smsp\": %.2f}\n", wps, n * 16 * 8 * 32 * 2 / ms * 1e-9,
           ms * 1e-3 * 1.965e9 / (ITERS * 8.0 * wps));
  }
  return 0;
}
