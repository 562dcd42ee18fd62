// Effective energy, conditionals and per-row energy gradients for BinaryRBM / PurificationRBM (sm_100a).
//
// Replaces rbm/binary_rbm.py:75-168 and rbm/purification_rbm.py:131-259 (F.linear + softplus +
// sum; matmul + add_ + sigmoid_ + clamp_) with fused kernels: the stacked weight matrix [W; U]
// stays in shared memory (odd row stride, conflict-free), one warp walks one row of v at a time
// and the elementwise tail (softplus / sigmoid) never leaves the SM.  Inputs may be arbitrary
// real vectors, as in the reference.  The reduced gradient lives in grad_reduce.cu.
#include "rows.cuh"

namespace qcb {

// MODE 0: energy  E[r] = -(vb . x_r + sum_o softplus(theta_o))      (out: rows)
// MODE 1: sigmoid out[r][o] = clamp(sigmoid(theta_o), 0, 1)          (out: rows x n_out)
// MODE 2: affine  out[r][o] = theta_o                                (out: rows x n_out)
template <int MODE>
__global__ void __launch_bounds__(ROWS_THREADS) rows_affine_kernel(AffineSpec s, const double* __restrict__ x,
                                                                   const double* __restrict__ x2,
                                                                   const double* __restrict__ vb, int64_t rows,
                                                                   double* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* Ws = reinterpret_cast<double*>(smem_raw);
  const int nin = s.n_in();
  const int LD = nin | 1;
  double* bs = Ws + (size_t)s.n_out * LD;
  double* vbs = bs + s.n_out;                           // visible bias (MODE 0)
  double* xbuf = vbs + nin + (threadIdx.x >> 5) * nin;  // per-warp input row
  stage_affine(s, Ws, LD, bs);
  if (MODE == 0)
    for (int i = threadIdx.x; i < nin; i += blockDim.x) vbs[i] = vb[i];
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t r = (int64_t)blockIdx.x * ROWS_WARPS + warp; r < rows; r += (int64_t)gridDim.x * ROWS_WARPS) {
    for (int i = lane; i < s.n_in1; i += 32) xbuf[i] = x[r * s.n_in1 + i];
    for (int i = lane; i < s.n_in2; i += 32) xbuf[s.n_in1 + i] = x2[r * s.n_in2 + i];
    __syncwarp();
    double part = 0.0;
    for (int o = lane; o < s.n_out; o += 32) {
      const double* wrow = Ws + (size_t)o * LD;
      double th = bs[o];
      for (int i = 0; i < nin; ++i) th = fma(wrow[i], xbuf[i], th);
      if (MODE == 0) {
        part += softplus20(th);
      } else if (MODE == 1) {
        out[r * s.n_out + o] = fmin(fmax(sigmoid_exact(th), 0.0), 1.0);
      } else {
        out[r * s.n_out + o] = th;
      }
    }
    if (MODE == 0) {
      for (int i = lane; i < nin; i += 32) part = fma(vbs[i], xbuf[i], part);
      part = warp_sum(part);
      if (lane == 0) out[r] = -part;
    }
    __syncwarp();
  }
}

static size_t rows_affine_smem(const AffineSpec& s) {
  const int nin = s.n_in();
  return ((size_t)s.n_out * (nin | 1) + s.n_out + nin + (size_t)ROWS_WARPS * nin) * sizeof(double);
}

// layers beyond the shared-memory-resident kernels: row-tiled kernels with the weights in global memory (large.cu)
int run_rows_affine_tiled(int mode, const AffineSpec& s, const double* x, const double* x2, const double* vb, int64_t rows, double* out,
                          cudaStream_t st);
int run_grad_batch_tiled(const StackedParams& p, const double* v, int64_t rows, double* g_rows, cudaStream_t st);
int run_flip_logratio_tiled(const StackedParams& p, const double* v, int64_t rows, double* out, cudaStream_t st);

template <int MODE>
static int launch_rows_affine(const AffineSpec& s, const double* x, const double* x2, const double* vb, int64_t rows,
                              double* out, cudaStream_t st) {
  const size_t smem = rows_affine_smem(s);
  if (smem > SMEM_LIMIT) return run_rows_affine_tiled(MODE, s, x, x2, vb, rows, out, st);
  QCB_CUDA(cudaFuncSetAttribute(rows_affine_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
  const int64_t need = (rows + ROWS_WARPS - 1) / ROWS_WARPS;
  const int64_t cap = (int64_t)sm_count() * 4;
  const int grid = (int)(need < cap ? need : cap);
  rows_affine_kernel<MODE><<<grid, ROWS_THREADS, smem, st>>>(s, x, x2, vb, rows, out);
  QCB_LAUNCH_CHECK();
  return QCB_OK;
}

int try_energy_dmma(const AffineSpec& s, const double* v, const double* vb, int64_t rows, double* E, int* handled,
                    cudaStream_t st);

int run_energy(const StackedParams& p, const double* v, int64_t rows, double* E, cudaStream_t st) {
  if (rows == 0) return QCB_OK;
  int handled = 0;  // large batches: theta on the FP64 tensor cores (grad_reduce_dmma.cu)
  const int rc = try_energy_dmma(stacked_hidden_spec(p), v, p.b, rows, E, &handled, st);
  if (rc != QCB_OK || handled) return rc;
  return launch_rows_affine<0>(stacked_hidden_spec(p), v, nullptr, p.b, rows, E, st);
}

// mode: 1 = sigmoid (conditionals), 2 = raw affine (PurificationRBM.mixing_term and friends)
int run_linear(int mode, const double* M, int m_rows, int m_cols, int transpose_m, const double* bias, const double* x,
               const double* M2, int m2_rows, const double* x2, int64_t rows, double* out, cudaStream_t st) {
  if (rows == 0) return QCB_OK;
  AffineSpec s{};
  s.M1 = M; s.m1_rows = m_rows; s.m1_cols = m_cols; s.t1 = transpose_m;
  s.M2 = M2; s.m2_rows = m2_rows; s.m2_cols = m_cols; s.t2 = transpose_m;
  s.n_out = transpose_m ? m_cols : m_rows;
  s.n_in1 = transpose_m ? m_rows : m_cols;
  s.n_in2 = (M2 != nullptr) ? (transpose_m ? m2_rows : m_cols) : 0;
  s.bias1 = bias; s.nb1 = s.n_out; s.bias2 = nullptr;
  s.stacked_units = 0;
  if (M2 != nullptr && !transpose_m) return fail(QCB_ERR_BAD_ARG, "linear: second operand needs transpose_m=1");
  if (mode == 1) return launch_rows_affine<1>(s, x, x2, nullptr, rows, out, st);
  return launch_rows_affine<2>(s, x, x2, nullptr, rows, out, st);
}

// ---- Bernoulli thresholding (sample_* single steps) --------------------------------------------
__global__ void bernoulli_kernel(const double* __restrict__ p, int64_t n, const double* __restrict__ u, uint32_t k0,
                                 uint32_t k1, uint64_t chain, uint32_t sweep, double* __restrict__ out) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one Philox call = 4 consecutive slots
  if (q * 4 >= n) return;
  double uu[4];
  if (u != nullptr) {
#pragma unroll
    for (int w = 0; w < 4; ++w) uu[w] = (q * 4 + w < n) ? u[q * 4 + w] : 2.0;
  } else {
    uint32_t hi[4], lo[4];
    philox4x32_10((uint32_t)chain, (uint32_t)(chain >> 32), sweep, (uint32_t)q, k0, k1, hi);
    philox4x32_10((uint32_t)chain, (uint32_t)(chain >> 32), sweep, (uint32_t)q | 0x80000000u, k0, k1, lo);
#pragma unroll
    for (int w = 0; w < 4; ++w) uu[w] = uniform53(hi[w], lo[w]);
  }
#pragma unroll
  for (int w = 0; w < 4; ++w)
    if (q * 4 + w < n) out[q * 4 + w] = (uu[w] < p[q * 4 + w]) ? 1.0 : 0.0;
}

int run_bernoulli(const double* p, int64_t n, const double* uniforms, uint64_t seed, uint64_t chain_id0, uint64_t sweep0,
                  double* out, cudaStream_t st) {
  if (n == 0) return QCB_OK;
  if (uniforms == nullptr && n > (int64_t)1 << 33) return fail(QCB_ERR_BAD_ARG, "bernoulli: n too large for one Philox stream");
  const int64_t calls = (n + 3) / 4;
  bernoulli_kernel<<<(unsigned)((calls + 255) / 256), 256, 0, st>>>(p, n, uniforms, (uint32_t)seed, (uint32_t)(seed >> 32),
                                                                     chain_id0, (uint32_t)sweep0, out);
  QCB_LAUNCH_CHECK();
  return QCB_OK;
}

// ---- per-row gradients (reduce=False) -----------------------------------------------------------
__global__ void __launch_bounds__(ROWS_THREADS) grad_batch_kernel(AffineSpec s, const double* __restrict__ v, int64_t rows,
                                                                  double* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* Ws = reinterpret_cast<double*>(smem_raw);
  const int N = s.n_in1, H = s.n_out;
  const int LD = N | 1;
  double* bs = Ws + (size_t)H * LD;
  double* vbuf = bs + H + (threadIdx.x >> 5) * (N + H);
  double* pbuf = vbuf + N;
  stage_affine(s, Ws, LD, bs);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t G = (size_t)H * N + N + H;
  for (int64_t r = (int64_t)blockIdx.x * ROWS_WARPS + warp; r < rows; r += (int64_t)gridDim.x * ROWS_WARPS) {
    for (int i = lane; i < N; i += 32) vbuf[i] = v[r * N + i];
    __syncwarp();
    for (int o = lane; o < H; o += 32) {
      const double* wrow = Ws + (size_t)o * LD;
      double th = bs[o];
      for (int i = 0; i < N; ++i) th = fma(wrow[i], vbuf[i], th);
      pbuf[o] = fmin(fmax(sigmoid_exact(th), 0.0), 1.0);
    }
    __syncwarp();
    double* orow = out + (size_t)r * G;
    for (int e = lane; e < H * N; e += 32) {
      const int i = e / N, j = e - i * N;
      orow[e] = -(pbuf[i] * vbuf[j]);
    }
    for (int j = lane; j < N; j += 32) orow[(size_t)H * N + j] = -vbuf[j];
    for (int i = lane; i < H; i += 32) orow[(size_t)H * N + N + i] = -pbuf[i];
    __syncwarp();
  }
}

int run_grad_batch(const StackedParams& p, const double* v, int64_t rows, double* g_rows, cudaStream_t st) {
  if (rows == 0) return QCB_OK;
  const int N = p.N, H = p.nhid();
  const size_t smem = ((size_t)H * (N | 1) + H + (size_t)ROWS_WARPS * (N + H)) * sizeof(double);
  if (smem > SMEM_LIMIT) return run_grad_batch_tiled(p, v, rows, g_rows, st);
  QCB_CUDA(cudaFuncSetAttribute(grad_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
  const int64_t need = (rows + ROWS_WARPS - 1) / ROWS_WARPS;
  const int64_t cap = (int64_t)sm_count() * 4;
  grad_batch_kernel<<<(int)(need < cap ? need : cap), ROWS_THREADS, smem, st>>>(stacked_hidden_spec(p), v, rows, g_rows);
  QCB_LAUNCH_CHECK();
  return QCB_OK;
}

// ---- single-spin-flip log-ratios -----------------------------------------------------------------
// out[r][i] = E(v_r) - E(v_r with visible unit i flipped) = log p(v^(i)) - log p(v): the quantity
// every local estimator built from psi(v^(i)) / psi(v) needs (SigmaX / SigmaY, TFIM energy:
// observables/pauli.py:50-83, examples/Tutorial4/quantum_ising_chain.py:45-78).  The reference
// evaluates the full energy of N flipped copies of the batch; here theta = Wst v + hb is computed
// once per row and a flip of unit i changes hidden unit o's softplus by
//   log(1 + e^(theta_o +- W_oi)) - log(1 + e^theta_o) = log(tau_o + sigma_o e^(+-W_oi)),  sigma = sigmoid(theta), tau = 1 - sigma,
// so one flip is a product of nh FMAs against tables of e^(+-W) and ONE log (TAB = true).  Both terms of each FMA
// are positive, i.e. no cancellation.  The identity needs the reference's softplus to be in its log1p(exp) branch
// (F.softplus threshold 20) on both sides and binary inputs: a row with theta_o + max_i |W_oi| > 20 for some o, or a
// non-binary entry, or a layer whose products could leave the double range (sum_o max_i |W_oi| > 700, non-finite
// weights), takes the term-by-term path with the thresholded softplus instead (always when TAB = false: tables do
// not fit in shared memory).  A warp handles RW = 32 / N rows at a time so that small layers use all lanes; for
// binary rows of 33..64 units (one row per warp pass) theta is accumulated over the SET bits only (adding W_oi in
// ascending i: bit-identical to the dense FMA loop for finite weights) and each lane takes two columns that share the
// (sigma, tau) loads.  The kernel is bound by instruction issue (N = 10: 72 % of the issue slots) and shared-memory
// wavefronts, not by the FP64 pipe; sigma / tau use the table exponential and a Newton reciprocal.
struct FlipGeom { int RW, PW; };  // rows per warp pass, doubles of per-warp scratch (even)
inline FlipGeom flip_geom(int N, int H) {
  int rw = 32 / N;
  if (rw < 1) rw = 1;
  const int cap = 256 / H > 1 ? 256 / H : 1;
  if (rw > cap) rw = cap;
  int pw = 3 * rw * H + rw * N;
  pw += pw & 1;
  return {rw, pw};
}

template <bool TAB>
__global__ void __launch_bounds__(1024) rows_flip_kernel(AffineSpec s, const double* __restrict__ v,
                                                                 const double* __restrict__ vb, int64_t rows,
                                                                 double* __restrict__ out, int RW, int PW) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* Ws = reinterpret_cast<double*>(smem_raw);
  const int N = s.n_in1, H = s.n_out;
  const int LD = N | 1;
  double* bs = Ws + (size_t)H * LD;
  double* vbs = bs + H;
  double* Ep = vbs + N;                                   // e^{+W}, e^{-W}, max_i |W_oi| (TAB only)
  double* Em = Ep + (TAB ? (size_t)H * LD : 0);
  double* wmax = Em + (TAB ? (size_t)H * LD : 0);
  const int WARPS = blockDim.x >> 5;
  size_t off = (size_t)H * LD * (TAB ? 3 : 1) + 2 * H * (TAB ? 1 : 0) + (TAB ? 0 : H) + N;  // doubles before the scratch
  off += off & 1;                                         // (sigma, tau) pairs are read as 16-byte words
  double* scratch = Ws + off;
  double* wbuf = scratch + (size_t)(threadIdx.x >> 5) * PW;
  double2* st = reinterpret_cast<double2*>(wbuf);         // (RW, H): (sigma, tau) on product rows, (softplus(theta), -) otherwise
  double* th = wbuf + 2 * RW * H;                         // (RW, H) theta
  double* vrow = th + RW * H;                             // (RW, N)
  unsigned* words = reinterpret_cast<unsigned*>(scratch + (size_t)WARPS * PW);
  unsigned* cta_slow = words;                             // all ones: every row term by term
  unsigned* flag = words + 1 + (threadIdx.x >> 5);        // per warp: mask of term-by-term rows
  unsigned* bits = words + 1 + WARPS + (threadIdx.x >> 5) * 2;  // per warp: set bits of the pass, flattened (row, unit)
  stage_affine(s, Ws, LD, bs);
  for (int i = threadIdx.x; i < N; i += blockDim.x) vbs[i] = vb[i];
  if (threadIdx.x == 0) *cta_slow = TAB ? 0u : 0xffffffffu;
  __shared__ double e2tab[EXP2_TABLE];
  __shared__ double2 ltab[LOG_TABLE];
  fill_exp2_table(e2tab);
  fill_log_table(ltab);
  __syncthreads();
  if (TAB) {
    for (int idx = threadIdx.x; idx < H * N; idx += blockDim.x) {
      const int o = idx / N, i = idx - o * N;
      const double w = Ws[o * LD + i];
      Ep[o * LD + i] = exp(w);
      Em[o * LD + i] = exp(-w);
    }
    for (int o = threadIdx.x; o < H; o += blockDim.x) {
      double m = 0.0;
      for (int i = 0; i < N; ++i) {
        const double a = fabs(Ws[o * LD + i]);
        if (!(a <= m) && m == m) m = a;  // a NaN sticks
      }
      wmax[o] = m;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int o = 0; o < H; ++o) tot += wmax[o];
      if (!(tot <= 700.0)) *cta_slow = 0xffffffffu;  // also inf / NaN weights
    }
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  // set-bit accumulation of theta pays when a warp walks ONE row (33..64 units: uniform trip count, half the loads);
  // with several short rows per warp the lanes' trip counts diverge and the dense loop is cheaper
  const bool sparse_ok = TAB && N > 32 && N <= 64 && *cta_slow == 0u;
  const int rrH0 = lane / H, oH0 = lane - rrH0 * H;  // (row, unit) of a lane in the first pass of each loop, hoisted
  const int rrN0 = lane / N, iN0 = lane - rrN0 * N;
  for (int64_t g = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5); g * RW < rows; g += (int64_t)gridDim.x * WARPS) {
    const int64_t r0 = g * RW;
    const int nr = (int)(rows - r0 < RW ? rows - r0 : RW);
    if (lane == 0) *flag = *cta_slow;
    __syncwarp();
    for (int e0 = 0; e0 < nr * N; e0 += 32) {
      const int e = e0 + lane;
      const bool valid = e < nr * N;
      const double x = valid ? v[r0 * N + e] : 0.0;
      if (valid) vrow[e] = x;
      if (TAB) {
        if (x != 0.0 && x != 1.0) atomicOr(flag, 1u << (e / N));
        const unsigned ones = __ballot_sync(0xffffffffu, x == 1.0);  // bit = flattened (row, unit) index of the pass
        if (sparse_ok && lane == 0 && e0 < 64) bits[e0 >> 5] = ones;
      }
    }
    __syncwarp();
    const unsigned f0 = *flag;  // non-binary rows so far
    for (int e = lane; e < nr * H; e += 32) {
      const int rr = e < 32 ? rrH0 : e / H, o = e < 32 ? oH0 : e - rr * H;
      const double* wrow = Ws + (size_t)o * LD;
      double t = bs[o];
      if (sparse_ok && !((f0 >> rr) & 1u)) {
        unsigned m = bits[0];  // one row per pass here (N > 32)
        while (m) {
          const int i = __ffs(m) - 1;
          m &= m - 1;
          t += wrow[i];
        }
        m = bits[1];
        while (m) {
          const int i = __ffs(m) - 1;
          m &= m - 1;
          t += wrow[32 + i];
        }
      } else {
        const double* x = vrow + rr * N;
        for (int i = 0; i < N; ++i) t = fma(wrow[i], x[i], t);
      }
      th[e] = t;
      if (TAB) {
        if (!(t + wmax[o] <= 20.0)) atomicOr(flag, 1u << rr);
        // sigma, tau with the table exponential and a Newton reciprocal (only used when t <= 20, see above)
        const double ex = exp_tab(fmin(fmax(t, -700.0), 700.0), e2tab);
        const double dn = 1.0 + ex;
        double tau;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(tau) : "d"(dn));
        tau = fma(tau, fma(-dn, tau, 1.0), tau);
        tau = fma(tau, fma(-dn, tau, 1.0), tau);
        st[e] = make_double2(ex * tau, tau);
      }
    }
    __syncwarp();
    const unsigned f = *flag;
    if (f != 0u) {
      for (int e = lane; e < nr * H; e += 32)
        if ((f >> (e / H)) & 1u) st[e].x = softplus20(th[e]);
      __syncwarp();
    }
    if (TAB && RW == 1 && N > 32 && f == 0u) {
      // one product row per pass: columns i and i + 32 of a lane share the (sigma, tau) loads
      for (int i = lane; i < N; i += 64) {
        const bool two = i + 32 < N;
        const int i2 = two ? i + 32 : i;
        const double x1 = vrow[i], x2 = vrow[i2];
        const double* T1 = (x1 == 0.0 ? Ep : Em) + i;
        const double* T2 = (x2 == 0.0 ? Ep : Em) + i2;
        double p = 1.0, q = 1.0;
#pragma unroll 4
        for (int o = 0; o < H; ++o) {
          const double2 a = st[o];
          p *= fma(a.x, T1[(size_t)o * LD], a.y);
          q *= fma(a.x, T2[(size_t)o * LD], a.y);
        }
        out[r0 * N + i] = fma(1.0 - 2.0 * x1, vbs[i], log_tab(p, ltab));
        if (two) out[r0 * N + i2] = fma(1.0 - 2.0 * x2, vbs[i2], log_tab(q, ltab));
      }
    } else {
      for (int e = lane; e < nr * N; e += 32) {
        const int rr = e < 32 ? rrN0 : e / N, i = e < 32 ? iN0 : e - rr * N;
        const double x = vrow[e];
        const double sgn = 1.0 - 2.0 * x;  // flipping a binary unit changes it by +-1
        double d = sgn * vbs[i];
        if (!TAB || ((f >> rr) & 1u)) {
          const double* t = th + rr * H;
          const double2* sp = st + rr * H;
          for (int o = 0; o < H; ++o) d += softplus20(fma(sgn, Ws[(size_t)o * LD + i], t[o])) - sp[o].x;
        } else {
          const double* T = (x == 0.0 ? Ep : Em) + i;
          const double2* a = st + rr * H;
          double p0 = 1.0, p1 = 1.0;
          int o = 0;
          for (; o + 1 < H; o += 2) {
            p0 *= fma(a[o].x, T[(size_t)o * LD], a[o].y);
            p1 *= fma(a[o + 1].x, T[(size_t)(o + 1) * LD], a[o + 1].y);
          }
          if (o < H) p0 *= fma(a[o].x, T[(size_t)o * LD], a[o].y);
          d += log_tab(p0 * p1, ltab);  // the product is a normal positive double (sum of max |W| <= 700)
        }
        out[r0 * N + e] = d;
      }
    }
    __syncwarp();
  }
}

static size_t flip_smem(int N, int H, bool tab, const FlipGeom& g, int warps) {
  const size_t LD = (size_t)(N | 1);
  return ((size_t)H * LD * (tab ? 3 : 1) + H + N + (tab ? H : 0) + 1 + (size_t)warps * g.PW) * sizeof(double) +
         (1 + 3 * (size_t)warps) * sizeof(unsigned);
}

int run_flip_logratio(const StackedParams& p, const double* v, int64_t rows, double* out, cudaStream_t st) {
  if (rows == 0) return QCB_OK;
  const int N = p.N, H = p.nhid();
  const FlipGeom g = flip_geom(N, H);
  // CTA size: 8 warps while eight such CTAs fit on an SM; layers whose tables take most of the shared memory get
  // up to 32 warps around ONE copy of the tables instead (the row loop is latency-bound: it wants the warps)
  const size_t sm_bytes = 227 * 1024;
  bool tab = true;
  int warps = 8;
  if (flip_smem(N, H, true, g, 8) * 8 > sm_bytes) {
    warps = 32;
    while (warps > 8 && flip_smem(N, H, true, g, warps) > SMEM_LIMIT) warps >>= 1;
    if (flip_smem(N, H, true, g, warps) > SMEM_LIMIT) {
      tab = false;
      warps = 32;
      while (warps > 8 && flip_smem(N, H, false, g, warps) > SMEM_LIMIT) warps >>= 1;
    }
  }
  const size_t smem = flip_smem(N, H, tab, g, warps);
  if (smem > SMEM_LIMIT) return run_flip_logratio_tiled(p, v, rows, out, st);
  const int64_t groups = (rows + g.RW - 1) / g.RW;
  const int64_t need = (groups + warps - 1) / warps;
  const int64_t cap = (int64_t)sm_count() * (64 / warps);
  const int grid = (int)(need < cap ? need : cap);
  if (tab) {
    QCB_CUDA(cudaFuncSetAttribute(rows_flip_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
    rows_flip_kernel<true><<<grid, warps * 32, smem, st>>>(stacked_hidden_spec(p), v, p.b, rows, out, g.RW, g.PW);
  } else {
    QCB_CUDA(cudaFuncSetAttribute(rows_flip_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
    rows_flip_kernel<false><<<grid, warps * 32, smem, st>>>(stacked_hidden_spec(p), v, p.b, rows, out, g.RW, g.PW);
  }
  QCB_LAUNCH_CHECK();
  return QCB_OK;
}

}  // namespace qcb
