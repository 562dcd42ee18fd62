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

template <int MODE>
static int launch_rows_affine(const AffineSpec& s, const double* x, const double* x2, const double* vb, int64_t rows,
                              double* out, cudaStream_t st) {
  const size_t smem = rows_affine_smem(s);
  if (smem > SMEM_LIMIT) return fail(QCB_ERR_UNSUPPORTED, "layer %d x %d does not fit in shared memory", s.n_out, s.n_in());
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
  if (smem > SMEM_LIMIT) return fail(QCB_ERR_UNSUPPORTED, "grad_batch: layer %d x %d does not fit in shared memory", H, N);
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
// once per row and each flip costs nh softplus evaluations of theta_h +- W_hi.
__global__ void __launch_bounds__(ROWS_THREADS) rows_flip_kernel(AffineSpec s, const double* __restrict__ v,
                                                                 const double* __restrict__ vb, int64_t rows,
                                                                 double* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* Ws = reinterpret_cast<double*>(smem_raw);
  const int N = s.n_in1, H = s.n_out;
  const int LD = N | 1;
  double* bs = Ws + (size_t)H * LD;
  double* vbs = bs + H;
  double* wbuf = vbs + N + (threadIdx.x >> 5) * (N + 2 * H);  // per warp: v (N), theta (H), softplus(theta) (H)
  double* vrow = wbuf;
  double* th = wbuf + N;
  double* sp = th + H;
  stage_affine(s, Ws, LD, bs);
  for (int i = threadIdx.x; i < N; i += blockDim.x) vbs[i] = vb[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t r = (int64_t)blockIdx.x * ROWS_WARPS + warp; r < rows; r += (int64_t)gridDim.x * ROWS_WARPS) {
    for (int i = lane; i < N; i += 32) vrow[i] = v[r * N + i];
    __syncwarp();
    for (int o = lane; o < H; o += 32) {
      const double* wrow = Ws + (size_t)o * LD;
      double t = bs[o];
      for (int i = 0; i < N; ++i) t = fma(wrow[i], vrow[i], t);
      th[o] = t;
      sp[o] = softplus20(t);
    }
    __syncwarp();
    for (int i = lane; i < N; i += 32) {
      const double sgn = 1.0 - 2.0 * vrow[i];  // flipping a binary unit changes it by +-1
      double d = sgn * vbs[i];
      for (int o = 0; o < H; ++o) d += softplus20(fma(sgn, Ws[(size_t)o * LD + i], th[o])) - sp[o];
      out[r * N + i] = d;
    }
    __syncwarp();
  }
}

int run_flip_logratio(const StackedParams& p, const double* v, int64_t rows, double* out, cudaStream_t st) {
  if (rows == 0) return QCB_OK;
  const int N = p.N, H = p.nhid();
  const size_t smem = ((size_t)H * (N | 1) + H + N + (size_t)ROWS_WARPS * (N + 2 * H)) * sizeof(double);
  if (smem > SMEM_LIMIT) return fail(QCB_ERR_UNSUPPORTED, "flip_logratio: layer %d x %d does not fit in shared memory", H, N);
  QCB_CUDA(cudaFuncSetAttribute(rows_flip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
  const int64_t need = (rows + ROWS_WARPS - 1) / ROWS_WARPS;
  const int64_t cap = (int64_t)sm_count() * 4;
  rows_flip_kernel<<<(int)(need < cap ? need : cap), ROWS_THREADS, smem, st>>>(stacked_hidden_spec(p), v, p.b, rows, out);
  QCB_LAUNCH_CHECK();
  return QCB_OK;
}

}  // namespace qcb
