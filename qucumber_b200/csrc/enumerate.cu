// Exact enumeration over the visible Hilbert space (partition function, fidelity / KL partial sums)
// for sm_100a.
//
// Replaces generate_hilbert_space + partition (nn_states/neural_state.py:156-197,
// rbm/binary_rbm.py:230-241, rbm/purification_rbm.py:453-463) and the full-space passes of
// utils/training_statistics.py:27-76,131-183.  The reference materialises a (2^N, N) float64
// tensor (capped at N = 20) and pushes it through a GEMM; here a basis state is only an index:
//
//  * a thread owns an aligned block of 2^m consecutive indices and walks it in Gray-code order,
//    so consecutive states differ in ONE visible unit and the hidden pre-activations
//    theta_i = c_i + W_i . v are updated with nh adds per state instead of an N x nh product;
//  * all threads of a CTA flip the same bit in the same direction at the same step, so the
//    needed column of W is a shared-memory broadcast;
//  * -E(v) = b.v + sum_i softplus(theta_i) (torch threshold 20) feeds an online log-sum-exp kept
//    as (max, sum exp(x - max)); CTAs write partial pairs, a final kernel combines them in a
//    fixed order.  Index r maps to visible unit j = (r >> (N-1-j)) & 1, the reference's row order.
#include "common.cuh"

namespace qcb {

constexpr int ENUM_THREADS = 128;

struct EnumArgs {
  StackedParams p;
  uint64_t first, count;
  int m;                 // log2 of states per thread
  uint64_t nblocks;      // aligned 2^m blocks covering [first, first + count)
  uint64_t block0;       // index of the first aligned block
  double logZ;           // stats kernel only
  const double* t_re;
  const double* t_im;
};

struct LSE {
  double m, s;
};
__device__ __forceinline__ void lse_add(LSE& a, double x) {
  if (x > a.m) {
    a.s = a.s * exp(a.m - x) + 1.0;  // exp(-inf) = 0 handles the first element
    a.m = x;
  } else {
    a.s += exp(x - a.m);
  }
}
__device__ __forceinline__ LSE lse_merge(LSE a, LSE b) {
  if (b.m == -INFINITY) return a;
  if (a.m == -INFINITY) return b;
  LSE r;
  r.m = fmax(a.m, b.m);
  r.s = a.s * exp(a.m - r.m) + b.s * exp(b.m - r.m);
  return r;
}

// Shared-memory staging: Wt[j][i] (visible-major, NHT stride), hb[NHT], vb[N]
template <int NHT>
__device__ __forceinline__ void stage_enum(const StackedParams& p, double* Wt, double* hb, double* vb) {
  const int H = p.nhid();
  for (int idx = threadIdx.x; idx < p.N * NHT; idx += blockDim.x) {
    const int j = idx / NHT, i = idx - j * NHT;
    Wt[idx] = i < H ? p.w(i, j) : 0.0;
  }
  for (int i = threadIdx.x; i < NHT; i += blockDim.x) hb[i] = i < H ? p.hbias(i) : 0.0;
  for (int j = threadIdx.x; j < p.N; j += blockDim.x) vb[j] = p.b[j];
}

// MODE 0: log-sum-exp partials.  MODE 1: statistics partials (5 sums) given logZ.
template <int NHT, int MODE>
__global__ void __launch_bounds__(ENUM_THREADS) enum_kernel(const EnumArgs a, double* __restrict__ partials,
                                                            const int* __restrict__ flags) {
  if (flags != nullptr && flags[0] != 0) return;  // the product-form kernel handles this launch
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* Wt = reinterpret_cast<double*>(smem_raw);
  double* hb = Wt + (size_t)a.p.N * NHT;
  double* vb = hb + NHT;
  double* red = vb + round_up(a.p.N, 2);  // reduction scratch: 8 doubles per warp
  stage_enum<NHT>(a.p, Wt, hb, vb);
  __syncthreads();

  const int N = a.p.N, H = a.p.nhid();
  const uint64_t last = a.first + a.count;  // exclusive
  const uint32_t steps = 1u << a.m;

  LSE acc{-INFINITY, 0.0};
  double sums[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  const double eps = 0x1.0p-52;

  for (uint64_t blk = (uint64_t)blockIdx.x * ENUM_THREADS + threadIdx.x; blk < a.nblocks;
       blk += (uint64_t)gridDim.x * ENUM_THREADS) {
    const uint64_t base = (a.block0 + blk) << a.m;
    // theta for the block's first state (gray(0) = 0 -> index = base)
    double th[NHT];
#pragma unroll
    for (int i = 0; i < NHT; ++i) th[i] = hb[i];
    double bv = 0.0;
    for (int j = 0; j < N; ++j) {
      if ((base >> (N - 1 - j)) & 1ull) {
        const double* col = Wt + (size_t)j * NHT;
#pragma unroll
        for (int i = 0; i < NHT; ++i) th[i] += col[i];
        bv += vb[j];
      }
    }
    for (uint32_t k = 0; k < steps; ++k) {
      const uint32_t gk = k ^ (k >> 1);
      const uint64_t idx = base | gk;
      double x = bv;
#pragma unroll
      for (int i = 0; i < NHT; ++i)
        if (i < H) x += softplus20(th[i]);
      if (idx >= a.first && idx < last) {
        if (MODE == 0) {
          lse_add(acc, x);
        } else {
          const double lp = x - a.logZ;
          const double p = exp(lp);
          sums[4] += p;
          if (a.t_re != nullptr) {
            const double tr = a.t_re[idx - a.first];
            const double ti = a.t_im != nullptr ? a.t_im[idx - a.first] : 0.0;
            const double psi = exp(0.5 * lp);
            sums[0] = fma(tr, psi, sums[0]);
            sums[1] = fma(ti, psi, sums[1]);
            const double pt = tr * tr + ti * ti;
            sums[2] = fma(pt, log(fmin(fmax(pt, eps), 1.0 - eps)), sums[2]);
            sums[3] = fma(pt, log(fmin(fmax(p, eps), 1.0 - eps)), sums[3]);
          }
        }
      }
      if (k + 1 < steps) {
        const uint32_t k1 = k + 1;
        const int pbit = __ffs(k1) - 1;
        const uint32_t g1 = k1 ^ (k1 >> 1);
        const double sgn = ((g1 >> pbit) & 1u) ? 1.0 : -1.0;
        const int j = N - 1 - pbit;
        const double* col = Wt + (size_t)j * NHT;
#pragma unroll
        for (int i = 0; i < NHT; ++i) th[i] = fma(sgn, col[i], th[i]);
        bv = fma(sgn, vb[j], bv);
      }
    }
  }

  // CTA reduction -> partials[blockIdx]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (MODE == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      LSE other;
      other.m = __shfl_xor_sync(0xffffffffu, acc.m, o);
      other.s = __shfl_xor_sync(0xffffffffu, acc.s, o);
      acc = lse_merge(acc, other);
    }
    if (lane == 0) { red[2 * warp] = acc.m; red[2 * warp + 1] = acc.s; }
    __syncthreads();
    if (threadIdx.x == 0) {
      LSE t{red[0], red[1]};
      for (int w = 1; w < ENUM_THREADS / 32; ++w) t = lse_merge(t, LSE{red[2 * w], red[2 * w + 1]});
      partials[2 * blockIdx.x] = t.m;
      partials[2 * blockIdx.x + 1] = t.s;
    }
  } else {
#pragma unroll
    for (int q = 0; q < 5; ++q) sums[q] = warp_sum(sums[q]);
    if (lane == 0)
      for (int q = 0; q < 5; ++q) red[8 * warp + q] = sums[q];
    __syncthreads();
    if (threadIdx.x < 5) {
      double t = 0.0;
      for (int w = 0; w < ENUM_THREADS / 32; ++w) t += red[8 * w + threadIdx.x];
      partials[8 * blockIdx.x + threadIdx.x] = t;
    }
  }
}


// ---- product-form fast path for log Z ------------------------------------------------------------
// sum_v exp(-E(v)) = sum_v e^{b.v} prod_i f_i(theta_i),  f_i = theta_i > 20 ? e^{theta_i} : 1 + e^{theta_i}
// (the branch reproduces torch's softplus threshold exactly).  t_i = e^{theta_i} is carried along
// the Gray-code walk multiplicatively: a flip of visible unit j multiplies t_i by e^{+-W_ij} (tables
// in shared memory, broadcast), so a state costs ~3 nh FP64 operations and no exp/log at all.
// Each thread re-synchronises t_i from an exact evaluation at the start of its 2^m-state block
// (m <= 12), which bounds the multiplicative drift to ~1e-13 relative.
// Only used when a pre-pass proves that nothing can overflow or underflow (enum_bounds_kernel);
// otherwise the log-domain kernel above runs.  flags[0] = eligible, flags[1] = no pre-activation can
// exceed the softplus threshold (the select is skipped).
__global__ void enum_bounds_kernel(StackedParams p, int* __restrict__ flags) {
  __shared__ double red[4][32];
  const int H = p.nhid(), N = p.N;
  double sum_sp = 0.0, max_th = -INFINITY, max_abs = 0.0;
  int bad = 0;
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    double hi = p.hbias(i), lo = hi;
    for (int j = 0; j < N; ++j) {
      const double w = p.w(i, j);
      if (!(fabs(w) < 1e300)) bad = 1;
      hi += fmax(w, 0.0);
      lo += fmin(w, 0.0);
    }
    if (!(fabs(hi) < 1e300) || !(fabs(lo) < 1e300)) bad = 1;
    sum_sp += fmax(hi, 0.0) + 0.6931471805599453;
    max_th = fmax(max_th, hi);
    max_abs = fmax(max_abs, fmax(fabs(hi), fabs(lo)));
  }
  double bpos = 0.0, bneg = 0.0;
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    const double b = p.b[j];
    if (!(fabs(b) < 1e300)) bad = 1;
    bpos += fmax(b, 0.0);
    bneg += fmin(b, 0.0);
  }
  sum_sp = warp_sum(sum_sp);
  bpos = warp_sum(bpos);
  bneg = warp_sum(bneg);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    max_th = fmax(max_th, __shfl_xor_sync(0xffffffffu, max_th, o));
    max_abs = fmax(max_abs, __shfl_xor_sync(0xffffffffu, max_abs, o));
    bad |= __shfl_xor_sync(0xffffffffu, bad, o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[0][warp] = sum_sp; red[1][warp] = bpos; red[2][warp] = bneg; red[3][warp] = max_th; }
  __shared__ double red2[32];
  __shared__ int redb[32];
  if (lane == 0) { red2[warp] = max_abs; redb[warp] = bad; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0, bp = 0, bn = 0, mt = -INFINITY, ma = 0; int b = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      s += red[0][w]; bp += red[1][w]; bn += red[2][w]; mt = fmax(mt, red[3][w]); ma = fmax(ma, red2[w]); b |= redb[w];
    }
    const bool ok = !b && (s + bp < 650.0) && (bn > -650.0) && (ma < 650.0);
    flags[0] = ok ? 1 : 0;
    flags[1] = (mt <= 19.99) ? 1 : 0;
  }
}

template <int NHT>
__global__ void __launch_bounds__(ENUM_THREADS) enum_fast_kernel(const EnumArgs a, double* __restrict__ partials,
                                                                 const int* __restrict__ flags) {
  if (flags[0] == 0) return;
  const bool no_thresh = flags[1] != 0;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = a.p.N, H = a.p.nhid();
  double* Wt = reinterpret_cast<double*>(smem_raw);   // [N][NHT]      W_ij      (block re-synchronisation)
  double* Ep = Wt + (size_t)N * NHT;                  // [N][NHT]  exp(+W_ij)
  double* Em = Ep + (size_t)N * NHT;                  // [N][NHT]  exp(-W_ij)
  double* hb = Em + (size_t)N * NHT;                  // [NHT]
  double* vb = hb + NHT;                              // [N]
  double* ebp = vb + round_up(N, 2);                  // [N] exp(+b_j)
  double* ebm = ebp + round_up(N, 2);                 // [N] exp(-b_j)
  double* red = ebm + round_up(N, 2);
  stage_enum<NHT>(a.p, Wt, hb, vb);
  __syncthreads();
  for (int idx = threadIdx.x; idx < N * NHT; idx += blockDim.x) {
    const int i = idx % NHT;
    Ep[idx] = i < H ? exp(Wt[idx]) : 1.0;
    Em[idx] = i < H ? exp(-Wt[idx]) : 1.0;
  }
  for (int j = threadIdx.x; j < N; j += blockDim.x) { ebp[j] = exp(vb[j]); ebm[j] = exp(-vb[j]); }
  __syncthreads();

  const uint64_t last = a.first + a.count;
  const uint32_t steps = 1u << a.m;
  const double E20 = 485165195.40979028;  // e^20
  double S = 0.0;
  for (uint64_t blk = (uint64_t)blockIdx.x * ENUM_THREADS + threadIdx.x; blk < a.nblocks;
       blk += (uint64_t)gridDim.x * ENUM_THREADS) {
    const uint64_t base = (a.block0 + blk) << a.m;
    double t[NHT];
#pragma unroll
    for (int i = 0; i < NHT; ++i) t[i] = hb[i];
    double bv = 0.0;
    for (int j = 0; j < N; ++j) {
      if ((base >> (N - 1 - j)) & 1ull) {
        const double* col = Wt + (size_t)j * NHT;
#pragma unroll
        for (int i = 0; i < NHT; ++i) t[i] += col[i];
        bv += vb[j];
      }
    }
#pragma unroll
    for (int i = 0; i < NHT; ++i) t[i] = i < H ? exp(t[i]) : 0.0;   // padded units contribute a factor 1
    double eb = exp(bv);
    for (uint32_t k = 0; k < steps; ++k) {
      const uint64_t idx = base | (uint64_t)(k ^ (k >> 1));
      double p0 = eb, p1 = 1.0, p2 = 1.0, p3 = 1.0;
      if (no_thresh) {
#pragma unroll
        for (int i = 0; i < NHT; i += 4) {
          p0 *= 1.0 + t[i]; p1 *= 1.0 + t[i + 1]; p2 *= 1.0 + t[i + 2]; p3 *= 1.0 + t[i + 3];
        }
      } else {
#pragma unroll
        for (int i = 0; i < NHT; i += 4) {
          p0 *= t[i] > E20 ? t[i] : 1.0 + t[i];
          p1 *= t[i + 1] > E20 ? t[i + 1] : 1.0 + t[i + 1];
          p2 *= t[i + 2] > E20 ? t[i + 2] : 1.0 + t[i + 2];
          p3 *= t[i + 3] > E20 ? t[i + 3] : 1.0 + t[i + 3];
        }
      }
      const double P = (p0 * p1) * (p2 * p3);
      if (idx >= a.first && idx < last) S += P;
      if (k + 1 < steps) {
        const uint32_t k1 = k + 1;
        const int pbit = __ffs(k1) - 1;
        const uint32_t g1 = k1 ^ (k1 >> 1);
        const bool up = (g1 >> pbit) & 1u;
        const int j = N - 1 - pbit;
        const double* col = (up ? Ep : Em) + (size_t)j * NHT;
#pragma unroll
        for (int i = 0; i < NHT; ++i) t[i] *= col[i];
        eb *= up ? ebp[j] : ebm[j];
      }
    }
  }
  S = warp_sum(S);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[warp] = S;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < ENUM_THREADS / 32; ++w) tot += red[w];
    // (m, s) pair in the common format: log-sum = m + log(s)
    partials[2 * blockIdx.x] = tot > 0.0 ? 0.0 : -INFINITY;
    partials[2 * blockIdx.x + 1] = tot;
  }
}

// ---- product-form fast path for the fidelity / KL sums ---------------------------------------------
// Same walk as enum_fast_kernel, but a WARP owns an aligned block of 2^(m+5) indices: lane l holds the five
// lowest index bits (its own visible bits for the last five units) and all lanes flip the same higher bit at
// every step, so the per-state reads of the target amplitudes are 256 contiguous bytes per warp.
// Per state: p = e^{-E} / Z from the running product, then the five sums of MODE 1 above.
template <int NHT>
__global__ void __launch_bounds__(ENUM_THREADS) enum_stats_fast_kernel(const EnumArgs a, double* __restrict__ partials,
                                                                       const int* __restrict__ flags) {
  if (flags[0] == 0) return;
  const bool no_thresh = flags[1] != 0;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = a.p.N, H = a.p.nhid();
  double* Wt = reinterpret_cast<double*>(smem_raw);
  double* Ep = Wt + (size_t)N * NHT;
  double* Em = Ep + (size_t)N * NHT;
  double* hb = Em + (size_t)N * NHT;
  double* vb = hb + NHT;
  double* ebp = vb + round_up(N, 2);
  double* ebm = ebp + round_up(N, 2);
  double* red = ebm + round_up(N, 2);  // 8 doubles per warp
  stage_enum<NHT>(a.p, Wt, hb, vb);
  __shared__ double2 ltab[LOG_TABLE];
  fill_log_table(ltab);
  __syncthreads();
  for (int idx = threadIdx.x; idx < N * NHT; idx += blockDim.x) {
    const int i = idx % NHT;
    Ep[idx] = i < H ? exp(Wt[idx]) : 1.0;
    Em[idx] = i < H ? exp(-Wt[idx]) : 1.0;
  }
  for (int j = threadIdx.x; j < N; j += blockDim.x) { ebp[j] = exp(vb[j]); ebm[j] = exp(-vb[j]); }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int m = a.m < N - 5 ? a.m : N - 5;  // per-lane walk length (log2); the caller guarantees N >= 7
  const int wshift = m + 5;
  const uint64_t last = a.first + a.count;
  const uint64_t wb0 = a.first >> wshift, nwb = ((last - 1) >> wshift) - wb0 + 1;
  const uint32_t steps = 1u << m;
  const double E20 = 485165195.40979028;  // e^20
  const double Zinv = exp(-a.logZ);
  const double eps = 0x1.0p-52;
  double sums[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  for (uint64_t wb = (uint64_t)blockIdx.x * (ENUM_THREADS / 32) + warp; wb < nwb; wb += (uint64_t)gridDim.x * (ENUM_THREADS / 32)) {
    const uint64_t base = ((wb0 + wb) << wshift) | (uint64_t)lane;
    double t[NHT];
#pragma unroll
    for (int i = 0; i < NHT; ++i) t[i] = hb[i];
    double bv = 0.0;
    for (int j = 0; j < N; ++j) {
      if ((base >> (N - 1 - j)) & 1ull) {
        const double* col = Wt + (size_t)j * NHT;
#pragma unroll
        for (int i = 0; i < NHT; ++i) t[i] += col[i];
        bv += vb[j];
      }
    }
#pragma unroll
    for (int i = 0; i < NHT; ++i) t[i] = i < H ? exp(t[i]) : 0.0;
    double eb = exp(bv);
    // the target amplitudes of a state are fetched two steps ahead (few warps are resident: 64 registers hold t_i)
    auto fetch = [&](uint32_t k, double& re, double& im) {
      const uint64_t idx = base | ((uint64_t)(k ^ (k >> 1)) << 5);
      const bool live = k < steps && a.t_re != nullptr && idx >= a.first && idx < last;
      re = live ? a.t_re[idx - a.first] : 0.0;
      im = (live && a.t_im != nullptr) ? a.t_im[idx - a.first] : 0.0;
    };
    double tr0, ti0, tr1, ti1;
    fetch(0u, tr0, ti0);
    fetch(1u, tr1, ti1);
    for (uint32_t k = 0; k < steps; ++k) {
      const uint64_t idx = base | ((uint64_t)(k ^ (k >> 1)) << 5);
      const double tr = tr0, ti = ti0;
      tr0 = tr1;
      ti0 = ti1;
      fetch(k + 2u, tr1, ti1);
      double p0 = eb, p1 = 1.0, p2 = 1.0, p3 = 1.0;
      if (no_thresh) {
#pragma unroll
        for (int i = 0; i < NHT; i += 4) {
          p0 *= 1.0 + t[i]; p1 *= 1.0 + t[i + 1]; p2 *= 1.0 + t[i + 2]; p3 *= 1.0 + t[i + 3];
        }
      } else {
#pragma unroll
        for (int i = 0; i < NHT; i += 4) {
          p0 *= t[i] > E20 ? t[i] : 1.0 + t[i];
          p1 *= t[i + 1] > E20 ? t[i + 1] : 1.0 + t[i + 1];
          p2 *= t[i + 2] > E20 ? t[i + 2] : 1.0 + t[i + 2];
          p3 *= t[i + 3] > E20 ? t[i + 3] : 1.0 + t[i + 3];
        }
      }
      const double p = ((p0 * p1) * (p2 * p3)) * Zinv;
      if (idx >= a.first && idx < last) {
        sums[4] += p;
        if (a.t_re != nullptr) {
          const double psi = sqrt(p);
          sums[0] = fma(tr, psi, sums[0]);
          sums[1] = fma(ti, psi, sums[1]);
          const double pt = tr * tr + ti * ti;
          sums[2] = fma(pt, log_tab(fmin(fmax(pt, eps), 1.0 - eps), ltab), sums[2]);  // arguments in [2^-52, 1): normal
          sums[3] = fma(pt, log_tab(fmin(fmax(p, eps), 1.0 - eps), ltab), sums[3]);
        }
      }
      if (k + 1 < steps) {
        const uint32_t k1 = k + 1;
        const int pbit = __ffs(k1) - 1;
        const uint32_t g1 = k1 ^ (k1 >> 1);
        const bool up = (g1 >> pbit) & 1u;
        const int j = N - 1 - (pbit + 5);
        const double* col = (up ? Ep : Em) + (size_t)j * NHT;
#pragma unroll
        for (int i = 0; i < NHT; ++i) t[i] *= col[i];
        eb *= up ? ebp[j] : ebm[j];
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 5; ++q) sums[q] = warp_sum(sums[q]);
  if (lane == 0)
    for (int q = 0; q < 5; ++q) red[8 * warp + q] = sums[q];
  __syncthreads();
  if (threadIdx.x < 5) {
    double tot = 0.0;
    for (int w = 0; w < ENUM_THREADS / 32; ++w) tot += red[8 * w + threadIdx.x];
    partials[8 * blockIdx.x + threadIdx.x] = tot;
  }
}

// One block of 256 threads; strided partials per thread, then warps, then the block — a fixed order.
__global__ void __launch_bounds__(256) lse_combine_kernel(const double* __restrict__ partials, int n, double* __restrict__ out2) {
  __shared__ double red[16];
  LSE acc{-INFINITY, 0.0};
  for (int i = threadIdx.x; i < n; i += 256) acc = lse_merge(acc, LSE{partials[2 * i], partials[2 * i + 1]});
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    LSE other;
    other.m = __shfl_xor_sync(0xffffffffu, acc.m, o);
    other.s = __shfl_xor_sync(0xffffffffu, acc.s, o);
    acc = lse_merge(acc, other);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[2 * warp] = acc.m; red[2 * warp + 1] = acc.s; }
  __syncthreads();
  if (threadIdx.x == 0) {
    LSE t{red[0], red[1]};
    for (int w = 1; w < 8; ++w) t = lse_merge(t, LSE{red[2 * w], red[2 * w + 1]});
    out2[0] = t.m;
    out2[1] = t.s;
  }
}

// One block of 256 threads: strided partials per thread, then warps, then the block — a fixed order.
__global__ void __launch_bounds__(256) sums_combine_kernel(const double* __restrict__ partials, int n, double* __restrict__ out8) {
  __shared__ double red[8][8];
  double s[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  for (int i = threadIdx.x; i < n; i += 256)
#pragma unroll
    for (int q = 0; q < 5; ++q) s[q] += partials[8 * i + q];
#pragma unroll
  for (int q = 0; q < 5; ++q) s[q] = warp_sum(s[q]);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0)
    for (int q = 0; q < 5; ++q) red[warp][q] = s[q];
  __syncthreads();
  if (threadIdx.x < 8) {
    double t = 0.0;
    if (threadIdx.x < 5)
      for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    out8[threadIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256) neg_lse_kernel(const double* __restrict__ E, int64_t n, double* __restrict__ partials) {
  __shared__ double red[16];
  LSE acc{-INFINITY, 0.0};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) lse_add(acc, -E[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    LSE other;
    other.m = __shfl_xor_sync(0xffffffffu, acc.m, o);
    other.s = __shfl_xor_sync(0xffffffffu, acc.s, o);
    acc = lse_merge(acc, other);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[2 * warp] = acc.m; red[2 * warp + 1] = acc.s; }
  __syncthreads();
  if (threadIdx.x == 0) {
    LSE t{red[0], red[1]};
    for (int w = 1; w < 8; ++w) t = lse_merge(t, LSE{red[2 * w], red[2 * w + 1]});
    partials[2 * blockIdx.x] = t.m;
    partials[2 * blockIdx.x + 1] = t.s;
  }
}

static int pick_nht(int H) {
  if (H <= 8) return 8;
  if (H <= 16) return 16;
  if (H <= 32) return 32;
  if (H <= 64) return 64;
  return 0;
}

static int enum_max_grid() { return sm_count() * 8; }

static size_t enum_partials_bytes() { return align_up((size_t)enum_max_grid() * 8 * sizeof(double), 256); }

size_t enum_workspace_bytes() { return enum_partials_bytes() + 256; }  // + flags

// ---- more than 64 hidden-like units: the reference's own structure (binary_rbm.py:230-241: energies of a block of states, then
// a log-sum-exp) in chunks of 2^20 states — energies from the row-tiled kernel of large.cu, states still generated from their
// index, (max, sum) pairs / the five sums merged in a fixed order ----
int run_energy_enum_tiled(const StackedParams& p, uint64_t first, int64_t rows, double* E, cudaStream_t st);
constexpr int64_t ENUM_WIDE_CHUNK = (int64_t)1 << 20;
constexpr int ENUM_WIDE_MAX_CHUNKS = 1 << 14;
static size_t enum_wide_bytes() { return align_up((size_t)ENUM_WIDE_CHUNK * sizeof(double), 256) + align_up((size_t)ENUM_WIDE_MAX_CHUNKS * 2 * sizeof(double), 256); }
size_t enum_workspace_bytes_for(int nhid) { return enum_workspace_bytes() + (pick_nht(nhid) == 0 ? enum_wide_bytes() : 0); }

// the five sums of enum_kernel<.., 1> from explicit energies; block partials are stored by the first chunk and added to by the rest
__global__ void __launch_bounds__(256) stats_from_energy_kernel(const double* __restrict__ E, int64_t n, double logZ, const double* __restrict__ t_re,
                                                                const double* __restrict__ t_im, double* __restrict__ partials, int first_chunk) {
  __shared__ double red[8][8];
  double sums[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  const double eps = 0x1.0p-52;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double lp = -E[i] - logZ;
    const double p = exp(lp);
    sums[4] += p;
    if (t_re != nullptr) {
      const double tr = t_re[i], ti = t_im != nullptr ? t_im[i] : 0.0;
      const double psi = exp(0.5 * lp);
      sums[0] = fma(tr, psi, sums[0]);
      sums[1] = fma(ti, psi, sums[1]);
      const double pt = tr * tr + ti * ti;
      sums[2] = fma(pt, log(fmin(fmax(pt, eps), 1.0 - eps)), sums[2]);
      sums[3] = fma(pt, log(fmin(fmax(p, eps), 1.0 - eps)), sums[3]);
    }
  }
#pragma unroll
  for (int q = 0; q < 5; ++q) sums[q] = warp_sum(sums[q]);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0)
    for (int q = 0; q < 5; ++q) red[warp][q] = sums[q];
  __syncthreads();
  if (threadIdx.x < 5) {
    double tot = 0.0;
    for (int w = 0; w < 8; ++w) tot += red[w][threadIdx.x];
    double* dst = partials + 8 * blockIdx.x + threadIdx.x;
    *dst = first_chunk ? tot : *dst + tot;
  }
}

static int wide_check(const StackedParams& p, uint64_t first, uint64_t count, size_t ws_bytes, const void* ws) {
  if (p.N < 1 || p.N > 62) return fail(QCB_ERR_UNSUPPORTED, "enumeration: N=%d outside [1, 62]", p.N);
  const uint64_t total = 1ull << p.N;
  if (first > total || count > total - first) return fail(QCB_ERR_BAD_ARG, "enumeration: range exceeds 2^N states");
  if (count > (uint64_t)ENUM_WIDE_CHUNK * ENUM_WIDE_MAX_CHUNKS)
    return fail(QCB_ERR_UNSUPPORTED, "enumeration: more than 2^34 states per call with nh + na = %d > 64 (split the range)", p.nhid());
  if (ws == nullptr || ws_bytes < enum_workspace_bytes_for(p.nhid())) return fail(QCB_ERR_BAD_ARG, "enumeration: workspace too small");
  return QCB_OK;
}

static int run_logz_range_wide(const StackedParams& p, uint64_t first, uint64_t count, double* out2, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (int rc = wide_check(p, first, count, ws_bytes, ws)) return rc;
  double* partials = reinterpret_cast<double*>(ws);
  double* E = reinterpret_cast<double*>(static_cast<char*>(ws) + enum_workspace_bytes());
  double* pairs = reinterpret_cast<double*>(reinterpret_cast<char*>(E) + align_up((size_t)ENUM_WIDE_CHUNK * sizeof(double), 256));
  int nchunks = 0;
  for (uint64_t done = 0; done < count; done += (uint64_t)ENUM_WIDE_CHUNK, ++nchunks) {
    const int64_t n = (int64_t)(count - done < (uint64_t)ENUM_WIDE_CHUNK ? count - done : (uint64_t)ENUM_WIDE_CHUNK);
    if (int rc = run_energy_enum_tiled(p, first + done, n, E, st)) return rc;
    const int64_t need = (n + 255) / 256;
    const int grid = (int)(need < enum_max_grid() ? need : enum_max_grid());
    neg_lse_kernel<<<grid, 256, 0, st>>>(E, n, partials);
    QCB_LAUNCH_CHECK();
    lse_combine_kernel<<<1, 256, 0, st>>>(partials, grid, pairs + 2 * nchunks);
    QCB_LAUNCH_CHECK();
  }
  lse_combine_kernel<<<1, 256, 0, st>>>(pairs, nchunks, out2);
  QCB_LAUNCH_CHECK();
  return QCB_OK;
}

static int run_exact_stats_range_wide(const StackedParams& p, uint64_t first, uint64_t count, double logZ, const double* t_re, const double* t_im,
                                      double* out8, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (int rc = wide_check(p, first, count, ws_bytes, ws)) return rc;
  double* partials = reinterpret_cast<double*>(ws);
  double* E = reinterpret_cast<double*>(static_cast<char*>(ws) + enum_workspace_bytes());
  const int grid = sm_count() * 4;
  for (uint64_t done = 0; done < count; done += (uint64_t)ENUM_WIDE_CHUNK) {
    const int64_t n = (int64_t)(count - done < (uint64_t)ENUM_WIDE_CHUNK ? count - done : (uint64_t)ENUM_WIDE_CHUNK);
    if (int rc = run_energy_enum_tiled(p, first + done, n, E, st)) return rc;
    stats_from_energy_kernel<<<grid, 256, 0, st>>>(E, n, logZ, t_re != nullptr ? t_re + done : nullptr, t_im != nullptr ? t_im + done : nullptr,
                                                   partials, done == 0 ? 1 : 0);
    QCB_LAUNCH_CHECK();
  }
  sums_combine_kernel<<<1, 256, 0, st>>>(partials, count > 0 ? grid : 0, out8);
  QCB_LAUNCH_CHECK();
  return QCB_OK;
}

template <int NHT, int MODE>
static int launch_enum(const EnumArgs& a, double* partials, int* grid_out, cudaStream_t st, const int* flags = nullptr) {
  const size_t smem = ((size_t)a.p.N * NHT + NHT + round_up(a.p.N, 2) + 8 * (ENUM_THREADS / 32)) * sizeof(double);
  if (smem > 200 * 1024) return fail(QCB_ERR_UNSUPPORTED, "enumeration: parameters do not fit in shared memory");
  QCB_CUDA(cudaFuncSetAttribute(enum_kernel<NHT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const uint64_t need = (a.nblocks + ENUM_THREADS - 1) / ENUM_THREADS;
  const uint64_t cap = (uint64_t)enum_max_grid();
  const int grid = (int)(need < cap ? need : cap);
  enum_kernel<NHT, MODE><<<grid, ENUM_THREADS, smem, st>>>(a, partials, flags);
  QCB_LAUNCH_CHECK();
  *grid_out = grid;
  return QCB_OK;
}

template <int NHT>
static int launch_enum_fast(const EnumArgs& a, double* partials, const int* flags, cudaStream_t st) {
  const int N = a.p.N;
  const size_t smem = ((size_t)3 * N * NHT + NHT + 3 * round_up(N, 2) + 8) * sizeof(double);
  if (smem > 200 * 1024) return QCB_OK;  // the general kernel covers it (flags stay as computed; see caller)
  QCB_CUDA(cudaFuncSetAttribute(enum_fast_kernel<NHT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const uint64_t need = (a.nblocks + ENUM_THREADS - 1) / ENUM_THREADS;
  const uint64_t cap = (uint64_t)enum_max_grid();
  const int grid = (int)(need < cap ? need : cap);
  enum_fast_kernel<NHT><<<grid, ENUM_THREADS, smem, st>>>(a, partials, flags);
  QCB_LAUNCH_CHECK();
  return QCB_OK;
}

static int setup_enum(const StackedParams& p, uint64_t first, uint64_t count, EnumArgs& a) {
  if (p.N < 1 || p.N > 62) return fail(QCB_ERR_UNSUPPORTED, "enumeration: N=%d outside [1, 62]", p.N);
  const uint64_t total = 1ull << p.N;
  if (first > total || count > total - first) return fail(QCB_ERR_BAD_ARG, "enumeration: range exceeds 2^N states");
  if (pick_nht(p.nhid()) == 0) return fail(QCB_ERR_UNSUPPORTED, "enumeration: nh + na = %d > 64", p.nhid());
  a.p = p;
  a.first = first;
  a.count = count;
  // states per thread: enough threads to fill the GPU (>= ~150k threads) but long Gray walks when possible
  int m = 0;
  const uint64_t want_threads = (uint64_t)sm_count() * 1024;
  while (m < 12 && m < p.N && (count >> (m + 1)) >= want_threads) ++m;
  if (m < 2 && p.N >= 2) m = 2;
  if (m > p.N) m = p.N;
  a.m = m;
  a.block0 = first >> m;
  a.nblocks = count == 0 ? 0 : (((first + count - 1) >> m) - a.block0 + 1);
  a.logZ = 0.0;
  a.t_re = a.t_im = nullptr;
  return QCB_OK;
}

int run_logz_range(const StackedParams& p, uint64_t first, uint64_t count, double* out2, void* ws, size_t ws_bytes,
                   cudaStream_t st) {
  if (pick_nht(p.nhid()) == 0) return run_logz_range_wide(p, first, count, out2, ws, ws_bytes, st);
  EnumArgs a;
  int rc = setup_enum(p, first, count, a);
  if (rc != QCB_OK) return rc;
  if (ws == nullptr || ws_bytes < enum_workspace_bytes()) return fail(QCB_ERR_BAD_ARG, "logz_range: workspace too small");
  double* partials = reinterpret_cast<double*>(ws);
  int* flags = reinterpret_cast<int*>(static_cast<char*>(ws) + enum_partials_bytes());
  int grid = 0;
  if (count > 0) {
    const int nht = pick_nht(p.nhid());
    const bool fast_fits = ((size_t)3 * p.N * nht + nht + 3 * round_up(p.N, 2) + 8) * sizeof(double) <= 200 * 1024;
    const int* gflags = nullptr;
    if (fast_fits) {
      // pre-pass decides on the device (no host sync) whether the product form is safe; exactly one of
      // the two enumeration kernels then does the work, the other returns at once
      enum_bounds_kernel<<<1, 128, 0, st>>>(p, flags);
      QCB_LAUNCH_CHECK();
      switch (nht) {
        case 8: rc = launch_enum_fast<8>(a, partials, flags, st); break;
        case 16: rc = launch_enum_fast<16>(a, partials, flags, st); break;
        case 32: rc = launch_enum_fast<32>(a, partials, flags, st); break;
        default: rc = launch_enum_fast<64>(a, partials, flags, st); break;
      }
      if (rc != QCB_OK) return rc;
      gflags = flags;
    }
    switch (nht) {
      case 8: rc = launch_enum<8, 0>(a, partials, &grid, st, gflags); break;
      case 16: rc = launch_enum<16, 0>(a, partials, &grid, st, gflags); break;
      case 32: rc = launch_enum<32, 0>(a, partials, &grid, st, gflags); break;
      default: rc = launch_enum<64, 0>(a, partials, &grid, st, gflags); break;
    }
    if (rc != QCB_OK) return rc;
  }
  lse_combine_kernel<<<1, 256, 0, st>>>(partials, grid, out2);
  QCB_LAUNCH_CHECK();
  return QCB_OK;
}

int run_neg_logsumexp(const double* E, int64_t n, double* out2, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (ws == nullptr || ws_bytes < enum_workspace_bytes()) return fail(QCB_ERR_BAD_ARG, "neg_logsumexp: workspace too small");
  double* partials = reinterpret_cast<double*>(ws);
  int grid = 0;
  if (n > 0) {
    const int64_t need = (n + 255) / 256;
    grid = (int)(need < enum_max_grid() ? need : enum_max_grid());
    neg_lse_kernel<<<grid, 256, 0, st>>>(E, n, partials);
    QCB_LAUNCH_CHECK();
  }
  lse_combine_kernel<<<1, 256, 0, st>>>(partials, grid, out2);
  QCB_LAUNCH_CHECK();
  return QCB_OK;
}

int run_exact_stats_range(const StackedParams& p, uint64_t first, uint64_t count, double logZ, const double* t_re,
                          const double* t_im, double* out8, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (pick_nht(p.nhid()) == 0) return run_exact_stats_range_wide(p, first, count, logZ, t_re, t_im, out8, ws, ws_bytes, st);
  EnumArgs a;
  int rc = setup_enum(p, first, count, a);
  if (rc != QCB_OK) return rc;
  if (ws == nullptr || ws_bytes < enum_workspace_bytes()) return fail(QCB_ERR_BAD_ARG, "exact_stats: workspace too small");
  a.logZ = logZ;
  a.t_re = t_re;
  a.t_im = t_im;
  double* partials = reinterpret_cast<double*>(ws);
  int* flags = reinterpret_cast<int*>(static_cast<char*>(ws) + enum_partials_bytes());
  int grid = 0;
  if (count > 0) {
    const int nht = pick_nht(p.nhid());
    const size_t fast_smem = ((size_t)3 * p.N * nht + nht + 3 * round_up(p.N, 2) + 8 * (ENUM_THREADS / 32)) * sizeof(double);
    const bool fast = p.N >= 7 && fast_smem <= 200 * 1024 && logZ > -600.0 && logZ < 700.0;
    // both kernels are launched on the same grid; the device-side pre-pass picks the one that works
    const uint64_t need = (a.nblocks + ENUM_THREADS - 1) / ENUM_THREADS;
    const uint64_t cap = (uint64_t)enum_max_grid();
    const int fgrid = (int)(need < cap ? need : cap);
    if (fast) {
      enum_bounds_kernel<<<1, 128, 0, st>>>(p, flags);
      QCB_LAUNCH_CHECK();
#define QCB_STATS_FAST(NHT)                                                                                              \
  QCB_CUDA(cudaFuncSetAttribute(enum_stats_fast_kernel<NHT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
  enum_stats_fast_kernel<NHT><<<fgrid, ENUM_THREADS, fast_smem, st>>>(a, partials, flags);
      switch (nht) {
        case 8: { QCB_STATS_FAST(8) } break;
        case 16: { QCB_STATS_FAST(16) } break;
        case 32: { QCB_STATS_FAST(32) } break;
        default: { QCB_STATS_FAST(64) } break;
      }
#undef QCB_STATS_FAST
      QCB_LAUNCH_CHECK();
    }
    const int* gflags = fast ? flags : nullptr;
    switch (nht) {
      case 8: rc = launch_enum<8, 1>(a, partials, &grid, st, gflags); break;
      case 16: rc = launch_enum<16, 1>(a, partials, &grid, st, gflags); break;
      case 32: rc = launch_enum<32, 1>(a, partials, &grid, st, gflags); break;
      default: rc = launch_enum<64, 1>(a, partials, &grid, st, gflags); break;
    }
    if (rc != QCB_OK) return rc;
  }
  sums_combine_kernel<<<1, 256, 0, st>>>(partials, grid, out8);
  QCB_LAUNCH_CHECK();
  return QCB_OK;
}

}  // namespace qcb
