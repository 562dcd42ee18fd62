// Shared device/host helpers for libqucumber_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include "../../include/qcb.h"

namespace qcb {

// ---- error plumbing (no exception crosses the C ABI) -----------------------------------------
void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

#define QCB_CUDA(expr)                                                                            \
  do {                                                                                            \
    cudaError_t e__ = (expr);                                                                     \
    if (e__ != cudaSuccess)                                                                       \
      return ::qcb::fail(QCB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),   \
                         __FILE__, __LINE__);                                                     \
  } while (0)

#define QCB_LAUNCH_CHECK()                                                                        \
  do {                                                                                            \
    ::qcb::g_launches.fetch_add(1, std::memory_order_relaxed);                                    \
    cudaError_t e__ = cudaGetLastError();                                                         \
    if (e__ != cudaSuccess)                                                                       \
      return ::qcb::fail(QCB_ERR_CUDA, "kernel launch failed: %s (%s:%d)",                        \
                         cudaGetErrorString(e__), __FILE__, __LINE__);                            \
  } while (0)

inline int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// Function attributes (dynamic shared memory limit, cluster size) and occupancy figures are cached per device: a
// process may drive several GPUs, and cudaFuncSetAttribute applies to the current device only.
constexpr int QCB_MAX_DEVICES = 64;
inline int device_slot() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < QCB_MAX_DEVICES) ? dev : 0;
}

__host__ __device__ constexpr size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
__host__ __device__ constexpr int round_up(int x, int a) { return (x + a - 1) / a * a; }

// Stacked "hidden-like" parameter view: rows [0, nh) come from W/c, rows [nh, nh+na) from U/d.
struct StackedParams {
  const double* W;
  const double* U;
  const double* b;
  const double* c;
  const double* d;
  int N, nh, na;
  __host__ __device__ int nhid() const { return nh + na; }
  __device__ __forceinline__ double w(int hid, int vis) const {
    return hid < nh ? W[(size_t)hid * N + vis] : U[(size_t)(hid - nh) * N + vis];
  }
  __device__ __forceinline__ double hbias(int hid) const { return hid < nh ? c[hid] : d[hid - nh]; }
};

// Where the chain states of a Gibbs call come from and where they go.  Exactly one source: v_in (float64 rows),
// bits_in (packed rows, see qcb.h "Compact states") or philox_init (v0 ~ Bernoulli(1/2) drawn in-kernel, Philox
// stream 5); v_out and / or bits_out receive v_k.
struct ChainIO {
  const double* v_in = nullptr;
  double* v_out = nullptr;
  const uint8_t* bits_in = nullptr;
  uint8_t* bits_out = nullptr;
  int philox_init = 0;
  __host__ __device__ bool plain() const { return bits_in == nullptr && bits_out == nullptr && !philox_init; }
};
__host__ __device__ inline int packed_row_bytes(int N) { return (N + 7) / 8; }
constexpr uint32_t PHILOX_STREAM_V0 = 5u;  // counter word 3 = (5 << 29) | call: the initial-state draws

#ifdef __CUDACC__
// ---- Philox4x32-10 ---------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ double uniform53(uint32_t hi, uint32_t lo) {
  const uint64_t bits = ((uint64_t)hi << 21) | (uint64_t)(lo >> 11);
  return (double)bits * 0x1.0p-53;
}

// torch.sigmoid in float64
__device__ __forceinline__ double sigmoid_exact(double x) { return 1.0 / (1.0 + exp(-x)); }

// Logistic function for the reduced-gradient kernels: same value as sigmoid_exact to ~2 ulp at a third of the
// FP64 instructions (the gradient sums only need 1e-10; the Gibbs decisions keep sigmoid_exact).
//   e^t = 2^(n/64) e^r,  n = rint(64 t / ln 2),  |r| <= ln2 / 128: 64-entry table (shared memory) x degree-5 Taylor;
//   1 / (1 + e) by the hardware reciprocal seed and two Newton steps.
constexpr int EXP2_TABLE = 64;
__device__ __forceinline__ void fill_exp2_table(double* tab) {  // call with all threads of the CTA, then sync
  for (int j = threadIdx.x; j < EXP2_TABLE; j += blockDim.x) tab[j] = exp2((double)j / EXP2_TABLE);
}
// e^t for |t| <= 700 (callers clamp), ~1.5 ulp
__device__ __forceinline__ double exp_tab(double t, const double* __restrict__ tab) {
  const double shifted = fma(t, 92.332482616893656, 0x1.8p52);           // 64 / ln 2
  const int n = __double2loint(shifted);
  const double nf = shifted - 0x1.8p52;
  double r = fma(nf, -0.010830424696249145, t);                          // ln2/64 split in two parts
  r = fma(nf, -3.623510646634843e-19, r);
  double p = fma(r, 1.0 / 120.0, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const double m = tab[n & (EXP2_TABLE - 1)] * p;                        // in [1, 2 e^{ln2/128})
  return __hiloint2double(__double2hiint(m) + ((n >> 6) << 20), __double2loint(m));
}
__device__ __forceinline__ double sigmoid_tab(double x, const double* __restrict__ tab) {
  const double e = exp_tab(fmin(fmax(-x, -700.0), 700.0), tab);
  const double d = 1.0 + e;
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  y = fma(y, fma(-d, y, 1.0), y);
  y = fma(y, fma(-d, y, 1.0), y);
  return y;
}

// Natural logarithm of a normal positive double to ~1e-14 absolute (the exact KL / NLL sums weight it with
// probabilities that add up to 1): x = 2^e m, m in [1, 2), c_j = 1 + (j + 1/2) / 64 the centre of m's cell,
// log x = e ln 2 + log c_j + log1p(m / c_j - 1) with |m / c_j - 1| <= 2^-7 and a degree-6 series; the 64-entry table of
// (1 / c_j, log c_j) lives in shared memory.  A third of the instructions of log().
constexpr int LOG_TABLE = 64;
__device__ __forceinline__ void fill_log_table(double2* tab) {  // call with all threads of the CTA, then sync
  for (int j = threadIdx.x; j < LOG_TABLE; j += blockDim.x) {
    const double c = 1.0 + ((double)j + 0.5) / LOG_TABLE;
    tab[j] = make_double2(1.0 / c, log(c));
  }
}
__device__ __forceinline__ double log_tab(double x, const double2* __restrict__ tab) {
  const int hi = __double2hiint(x);
  const int e = (hi >> 20) - 1023;
  const double m = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, __double2loint(x));
  const double2 t = tab[(hi >> 14) & (LOG_TABLE - 1)];
  const double r = fma(m, t.x, -1.0);
  double p = fma(r, -1.0 / 6.0, 0.2);
  p = fma(p, r, -0.25);
  p = fma(p, r, 1.0 / 3.0);
  p = fma(p, r, -0.5);
  p = fma(p * r, r, r);
  return fma((double)e, 0.6931471805599453, t.y + p);
}

// F.softplus(beta=1, threshold=20)
__device__ __forceinline__ double softplus20(double x) { return x > 20.0 ? x : log1p(exp(x)); }

// ---- warp helpers ----------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- mbarrier + 1-D TMA bulk copy (global -> shared::cta) --------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// FP64 tensor-core tile: D(8x8) += A(8x4) * B(4x8).  Lane l holds A[l>>2][l&3], B[l&3][l>>2],
// C[l>>2][2*(l&3) + {0,1}]  (PTX ISA, mma.m8n8k4 .f64 fragment layout).
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
#endif  // __CUDACC__

}  // namespace qcb
