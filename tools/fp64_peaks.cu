// Microbenchmarks that measure the FP64 roofline denominators the RBM kernels are judged
// against on one B200: DFMA / DADD / FFMA pipe peaks, FP64 tensor (DMMA) peaks, DFMA+DMMA
// overlap, the predicated-DADD + broadcast-LDS loop shape used by the Gibbs kernel, and the
// Philox4x32-10 rate. Prints one JSON object. Build:
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/fp64_peaks tools/fp64_peaks.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int ITERS = 4096;

__global__ void k_dfma(double* out, double a, double b) {
  double acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  if (s == 1.2345) out[0] = s;
}

__global__ void k_dadd(double* out, double a) {
  double acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = acc[i] + a;
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  if (s == 1.2345) out[0] = s;
}

__global__ void k_ffma(float* out, float a, float b) {
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  if (s == 1.2345f) out[0] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, "
               "{%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, "
               "{%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

__global__ void k_dmma884(double* out, double a, double b) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  if (s == 1.2345) out[0] = s;
}

__global__ void k_dmma16816(double* out, double a0, double b0) {
  double c[4][4]; double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = a0 + i;
#pragma unroll
  for (int i = 0; i < 4; ++i) b[i] = b0 + i;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[i][j] = threadIdx.x + i + j;
  for (int it = 0; it < ITERS / 4; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) dmma16816(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s += c[i][j];
  if (s == 1.2345) out[0] = s;
}

__global__ void k_dmma1688(double* out, double a0, double b0) {
  double c[4][4]; double a[4], b[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = a0 + i;
#pragma unroll
  for (int i = 0; i < 2; ++i) b[i] = b0 + i;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[i][j] = threadIdx.x + i + j;
  for (int it = 0; it < ITERS / 2; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) dmma1688(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s += c[i][j];
  if (s == 1.2345) out[0] = s;
}

// DFMA and DMMA issued from the same warp, interleaved (are the pipes separate?)
__global__ void k_mix_same_warp(double* out, double a, double b) {
  double c[4][2]; double acc[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) dmma884(c[i][0], c[i][1], a, b);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  if (s == 1.2345) out[0] = s;
}

// even warps DMMA, odd warps DFMA
__global__ void k_mix_split_warp(double* out, double a, double b) {
  if ((threadIdx.x >> 5) & 1) {
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    if (s == 1.2345) out[0] = s;
  } else {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 1.2345) out[0] = s;
  }
}

// Gibbs inner-loop shape: per input bit one predicate, OUT accumulators, W broadcast from smem.
template <int OUT>
__global__ void k_pred_dadd_lds(double* out, const double* __restrict__ w, unsigned long long seedbits) {
  __shared__ double sw[64 * 64];
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  unsigned long long bits = seedbits * (threadIdx.x + 1 + blockIdx.x * 977ull) | 1ull;
  double tot = 0;
  for (int it = 0; it < ITERS / 64; ++it) {
    for (int ob = 0; ob < 64; ob += OUT) {
      double acc[OUT];
#pragma unroll
      for (int o = 0; o < OUT; ++o) acc[o] = 0.0;
#pragma unroll 16
      for (int j = 0; j < 64; ++j) {
        const bool p = (bits >> j) & 1ull;
        const double* row = sw + j * 64 + ob;
#pragma unroll
        for (int o = 0; o < OUT; ++o) acc[o] += p ? row[o] : 0.0;
      }
#pragma unroll
      for (int o = 0; o < OUT; ++o) tot += acc[o];
    }
    bits = bits * 6364136223846793005ull + 1442695040888963407ull;
  }
  if (tot == 1.2345) out[0] = tot;
}

__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    c[0] = hi1 ^ c[1] ^ k0; c[1] = lo1; c[2] = hi0 ^ c[3] ^ k1; c[3] = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
__global__ void k_philox(uint32_t* out) {
  uint32_t acc = 0;
  for (int it = 0; it < ITERS; ++it) {
    uint32_t c[4] = {(uint32_t)it, threadIdx.x, blockIdx.x, 0u};
    philox4x32_10(c, 0x1234u, 0x5678u);
    acc ^= c[0] ^ c[1] ^ c[2] ^ c[3];
  }
  if (acc == 0x12345u) out[0] = acc;
}

template <typename F>
double time_ms(F launch, int reps = 5) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch(); CK(cudaDeviceSynchronize());
  double best = 1e30;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  const int blocks = sms * 8, threads = 256;
  double* dout; CK(cudaMalloc(&dout, 1024));
  double* dw; CK(cudaMalloc(&dw, 64 * 64 * 8)); CK(cudaMemset(dw, 0, 64 * 64 * 8));
  const double nthr = (double)blocks * threads;
  const double nwarp = nthr / 32;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d", prop.name, sms, prop.clockRate);

  double ms;
  ms = time_ms([&] { k_dfma<<<blocks, threads>>>(dout, 1.0000001, 1e-9); });
  printf(", \"dfma_tflops\": %.3f", nthr * ITERS * 8 * 2 / ms * 1e-9);
  ms = time_ms([&] { k_dadd<<<blocks, threads>>>(dout, 1e-9); });
  printf(", \"dadd_tops\": %.3f", nthr * ITERS * 8 / ms * 1e-9);
  ms = time_ms([&] { k_ffma<<<blocks, threads>>>((float*)dout, 1.0000001f, 1e-9f); });
  printf(", \"ffma_tflops\": %.3f", nthr * ITERS * 16 * 2 / ms * 1e-9);
  ms = time_ms([&] { k_dmma884<<<blocks, threads>>>(dout, 1.0000001, 1e-9); });
  printf(", \"dmma_m8n8k4_tflops\": %.3f", nwarp * ITERS * 8 * (8 * 8 * 4 * 2.0) / ms * 1e-9);
  ms = time_ms([&] { k_dmma1688<<<blocks, threads>>>(dout, 1.0000001, 1e-9); });
  printf(", \"dmma_m16n8k8_tflops\": %.3f", nwarp * (ITERS / 2) * 4 * (16 * 8 * 8 * 2.0) / ms * 1e-9);
  ms = time_ms([&] { k_dmma16816<<<blocks, threads>>>(dout, 1.0000001, 1e-9); });
  printf(", \"dmma_m16n8k16_tflops\": %.3f", nwarp * (ITERS / 4) * 4 * (16 * 8 * 16 * 2.0) / ms * 1e-9);
  ms = time_ms([&] { k_mix_same_warp<<<blocks, threads>>>(dout, 1.0000001, 1e-9); });
  printf(", \"mix_same_warp_tflops\": %.3f", (nwarp * ITERS * 4 * 512.0 + nthr * ITERS * 8 * 2) / ms * 1e-9);
  ms = time_ms([&] { k_mix_split_warp<<<blocks, threads>>>(dout, 1.0000001, 1e-9); });
  printf(", \"mix_split_warp_tflops\": %.3f", (nwarp / 2 * ITERS * 8 * 512.0 + nthr / 2 * ITERS * 8 * 2) / ms * 1e-9);
  ms = time_ms([&] { k_pred_dadd_lds<4><<<blocks, threads>>>(dout, dw, 0x9E3779B97F4A7C15ull); });
  printf(", \"pred_dadd_lds_out4_tops\": %.3f", nthr * (ITERS / 64) * 64.0 * 64 / ms * 1e-9);
  ms = time_ms([&] { k_pred_dadd_lds<8><<<blocks, threads>>>(dout, dw, 0x9E3779B97F4A7C15ull); });
  printf(", \"pred_dadd_lds_out8_tops\": %.3f", nthr * (ITERS / 64) * 64.0 * 64 / ms * 1e-9);
  ms = time_ms([&] { k_pred_dadd_lds<16><<<blocks, threads>>>(dout, dw, 0x9E3779B97F4A7C15ull); });
  printf(", \"pred_dadd_lds_out16_tops\": %.3f", nthr * (ITERS / 64) * 64.0 * 64 / ms * 1e-9);
  ms = time_ms([&] { k_philox<<<blocks, threads>>>((uint32_t*)dout); });
  printf(", \"philox4x32_10_gcalls\": %.3f", nthr * ITERS / ms * 1e-6);
  printf("}\n");
  return 0;
}
