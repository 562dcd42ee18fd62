// Probe for the Blackwell tensor path of the integer Gibbs filter: tcgen05.mma.cta_group::1.kind::i8 (u8 x s8 -> s32) with
// the accumulator in TMEM, operands in shared memory in the canonical K-major no-swizzle layout, results read back with
// tcgen05.ld.  Checks the descriptors / layouts against a CPU product and times the round trip a Gibbs half-sweep needs:
// write the A tile (state bytes) -> fence -> 4 MMAs (M = 128, N = 128, K = 4 x 32) -> commit -> wait -> tcgen05.ld.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/umma_i8_probe tools/umma_i8_probe.cu
//
// Layout used (bytes): element (row r, k) of an operand with R rows lives at
//   (k / 16) * (R / 8) * 128 + (r / 8) * 128 + (r % 8) * 16 + k % 16
// i.e. 8 x 16-byte core matrices, 8-row groups 128 B apart (SBO), 16-byte K chunks (R / 8) * 128 B apart (LBO).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int M = 128, N = 128, K = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__host__ __device__ inline size_t operand_offset(int rows, int r, int k) {
  return (size_t)(k / 16) * (rows / 8) * 128 + (size_t)(r / 8) * 128 + (r % 8) * 16 + k % 16;
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // base offset 0, layout type 0 = no swizzle
}

__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(
          smem_u32(bar)),
      "r"(parity)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) probe_kernel(const uint8_t* __restrict__ A, const int8_t* __restrict__ B, int* __restrict__ D,
                                                       int iters, long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                    // M x K bytes
  uint8_t* sB = smem + M * K;            // N x K bytes
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "r"(128u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // B: copied once (already in the canonical layout in global memory)
  for (int i = tid; i < N * K / 16; i += blockDim.x) reinterpret_cast<uint4*>(sB)[i] = reinterpret_cast<const uint4*>(B)[i];
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t tmem = tmem_base_s;
  // instruction descriptor: D = s32, A = u8, B = s8, both K-major, N, M
  const uint32_t idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  const uint32_t lboA = (M / 8) * 128, lboB = (N / 8) * 128;

  uint32_t phase = 0;
  long long t0 = 0, total = 0;
  int acc[32];
  for (int it = 0; it < iters; ++it) {
    if (it == 1) t0 = clock64();
    // each thread writes its row of A (row = tid): K bytes as K / 16 16-byte chunks — what a Gibbs thread does with its chain's bits
    const uint4* arow = reinterpret_cast<const uint4*>(A + (size_t)tid * K);
#pragma unroll
    for (int kc = 0; kc < K / 16; ++kc) {
      uint4 v = arow[kc];
      v.x += (uint32_t)it * 0u;  // keep the load inside the loop
      *reinterpret_cast<uint4*>(sA + operand_offset(M, tid, 16 * kc)) = v;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    if (warp == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;\n");
      if (lane == 0) {
#pragma unroll
        for (int ks = 0; ks < K / 32; ++ks) {
          const uint64_t ad = make_desc(smem_u32(sA) + ks * 2 * lboA, lboA, 128);
          const uint64_t bd = make_desc(smem_u32(sB) + ks * 2 * lboB, lboB, 128);
          mma_i8(tmem, ad, bd, idesc, ks > 0 ? 1u : 0u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
      }
      __syncwarp();
    }
    mbar_wait(&mbar, phase);
    phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    // every thread reads its row: lane quarter of this warp, 32 columns at a time
    const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < N; c0 += 32) {
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,"
          "%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
          : "=r"(acc[0]), "=r"(acc[1]), "=r"(acc[2]), "=r"(acc[3]), "=r"(acc[4]), "=r"(acc[5]), "=r"(acc[6]), "=r"(acc[7]), "=r"(acc[8]),
            "=r"(acc[9]), "=r"(acc[10]), "=r"(acc[11]), "=r"(acc[12]), "=r"(acc[13]), "=r"(acc[14]), "=r"(acc[15]), "=r"(acc[16]),
            "=r"(acc[17]), "=r"(acc[18]), "=r"(acc[19]), "=r"(acc[20]), "=r"(acc[21]), "=r"(acc[22]), "=r"(acc[23]), "=r"(acc[24]),
            "=r"(acc[25]), "=r"(acc[26]), "=r"(acc[27]), "=r"(acc[28]), "=r"(acc[29]), "=r"(acc[30]), "=r"(acc[31])
          : "r"(taddr + (uint32_t)c0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (it == iters - 1) {
#pragma unroll
        for (int j = 0; j < 32; ++j) D[(size_t)tid * N + c0 + j] = acc[j];
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
  }
  total = clock64() - t0;
  if (tid == 0 && iters > 1) cycles[0] = total / (iters - 1);
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(128u));
}

// Throughput of the same MMA shape: `reps` batches of 4 x tcgen05.mma (M = N = 128, K = 32 each) issued back to back by one thread
// per CTA, one commit + wait at the end; one CTA per SM (the tensor pipe is per SM).  Operand contents do not matter.
__global__ void __launch_bounds__(128, 1) rate_kernel(int reps, long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (M * K + N * K) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "r"(128u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  const uint32_t lbo = (M / 8) * 128;
  if (tid == 0) {
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int ks = 0; ks < K / 32; ++ks)
        mma_i8(tmem, make_desc(smem_u32(smem) + ks * 2 * lbo, lbo, 128), make_desc(smem_u32(smem + M * K) + ks * 2 * lbo, lbo, 128), idesc, 1u);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    mbar_wait(&mbar, 0);
    if (blockIdx.x == 0) cycles[0] = clock64() - t0;
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(128u));
}

int main() {
  std::vector<uint8_t> hA((size_t)M * K);
  std::vector<int8_t> hB((size_t)N * K), hBl((size_t)N * K);
  srand(1);
  for (auto& x : hA) x = (uint8_t)(rand() % 256);
  for (auto& x : hB) x = (int8_t)(rand() % 256 - 128);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) hBl[operand_offset(N, n, k)] = hB[(size_t)n * K + k];
  uint8_t* dA; int8_t* dB; int* dD; long long* dC;
  cudaMalloc(&dA, hA.size()); cudaMalloc(&dB, hBl.size()); cudaMalloc(&dD, (size_t)M * N * 4); cudaMalloc(&dC, 8);
  cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hBl.data(), hBl.size(), cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xff, (size_t)M * N * 4);
  const size_t smem = (size_t)M * K + (size_t)N * K;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_kernel<<<1, 128, smem>>>(dA, dB, dD, 1001, dC);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; }
  std::vector<int> hD((size_t)M * N);
  long long cyc = 0;
  cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
  long long bad = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      int ref = 0;
      for (int k = 0; k < K; ++k) ref += (int)hA[(size_t)m * K + k] * (int)hB[(size_t)n * K + k];
      if (ref != hD[(size_t)m * N + n]) {
        if (bad < 5) printf("mismatch at (%d,%d): got %d want %d\n", m, n, hD[(size_t)m * N + n], ref);
        ++bad;
      }
    }
  // rate: all SMs, timed with events
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int reps = 20000;
  rate_kernel<<<sms, 128, smem>>>(100, dC);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  rate_kernel<<<sms, 128, smem>>>(reps, dC);
  cudaEventRecord(e1);
  e = cudaDeviceSynchronize();
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  long long rcyc = 0;
  cudaMemcpy(&rcyc, dC, 8, cudaMemcpyDeviceToHost);
  const double ops = 2.0 * M * N * K * (double)reps * sms;
  printf("{\"probe\": \"tcgen05.mma.kind::i8 M=%d N=%d K=%d\", \"mismatches\": %lld, \"round_trip_cycles\": %lld, "
         "\"rate\": {\"sms\": %d, \"ms\": %.4f, \"tops\": %.1f, \"cycles_per_4_mma\": %.1f, \"error\": \"%s\"}}\n",
         M, N, K, bad, cyc, sms, ms, ops / (ms * 1e-3) / 1e12, (double)rcyc / reps, e == cudaSuccess ? "" : cudaGetErrorString(e));
  return bad == 0 ? 0 : 2;
}
