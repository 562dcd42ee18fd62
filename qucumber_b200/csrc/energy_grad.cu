// Effective energy, conditionals and energy gradients for BinaryRBM / PurificationRBM (sm_100a).
//
// Replaces rbm/binary_rbm.py:75-168 and rbm/purification_rbm.py:131-259 (F.linear + softplus +
// sum, matmul + add_ + sigmoid_ + clamp_, and the two GEMMs + reductions of
// effective_energy_gradient) with fused kernels: the stacked weight matrix [W; U] stays in shared
// memory (odd row stride, conflict-free), one warp walks one row of v at a time and the
// elementwise tail (softplus / sigmoid / outer-product accumulation) never leaves the SM.
// Inputs may be arbitrary real vectors, as in the reference.
#include "common.cuh"

namespace qcb {

constexpr int ROWS_THREADS = 256;
constexpr int ROWS_WARPS = ROWS_THREADS / 32;
constexpr size_t SMEM_LIMIT = 200 * 1024;

// Generic "affine rows" operand: out unit o = bias[o] + sum_i Mop[o][i] * xin[i], where the input is
// the concatenation [x (n_in1) | x2 (n_in2)] and Mop[o][i] is gathered from up to two stored matrices.
struct AffineSpec {
  const double* M1; int m1_rows, m1_cols, t1;   // t1 = 1: Mop[o][i] = M1[i][o]
  const double* M2; int m2_rows, m2_cols, t2;   // second operand (may be NULL)
  const double* bias1; int nb1;                 // bias for units [0, nb1)
  const double* bias2;                          // bias for units [nb1, n_out) (may be NULL -> 0)
  int n_out, n_in1, n_in2;
  // mode A ("stacked units"): units [0,m1_rows) from M1 rows, [m1_rows, ..) from M2 rows; one input x
  // mode B ("stacked inputs"): all units see M1^T over x and M2^T over x2
  int stacked_units;
  __device__ __forceinline__ double mop(int o, int i) const {
    if (stacked_units) {
      if (o < m1_rows) return M1[(size_t)o * m1_cols + i];
      return M2[(size_t)(o - m1_rows) * m2_cols + i];
    }
    if (i < n_in1) return t1 ? M1[(size_t)i * m1_cols + o] : M1[(size_t)o * m1_cols + i];
    const int i2 = i - n_in1;
    return t2 ? M2[(size_t)i2 * m2_cols + o] : M2[(size_t)o * m2_cols + i2];
  }
  __device__ __forceinline__ double bias(int o) const {
    if (o < nb1) return bias1 ? bias1[o] : 0.0;
    return bias2 ? bias2[o - nb1] : 0.0;
  }
  __host__ __device__ int n_in() const { return n_in1 + n_in2; }
};

__device__ __forceinline__ void stage_affine(const AffineSpec& s, double* Ws, int LD, double* bs) {
  const int nin = s.n_in();
  for (int idx = threadIdx.x; idx < s.n_out * nin; idx += blockDim.x) {
    const int o = idx / nin, i = idx - o * nin;
    Ws[o * LD + i] = s.mop(o, i);
  }
  for (int o = threadIdx.x; o < s.n_out; o += blockDim.x) bs[o] = s.bias(o);
}

// MODE 0: energy  E[r] = -(vb . x_r + sum_o softplus(theta_o))      (out: rows)
// MODE 1: sigmoid out[r][o] = clamp(sigmoid(theta_o), 0, 1)          (out: rows x n_out)
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
  double* vbs = bs + s.n_out;                       // visible bias (MODE 0)
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
      } else {
        double p = sigmoid_exact(th);
        p = fmin(fmax(p, 0.0), 1.0);
        out[r * s.n_out + o] = p;
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

static AffineSpec stacked_hidden_spec(const StackedParams& p) {
  AffineSpec s{};
  s.M1 = p.W; s.m1_rows = p.nh; s.m1_cols = p.N; s.t1 = 0;
  s.M2 = p.U; s.m2_rows = p.na; s.m2_cols = p.N; s.t2 = 0;
  s.bias1 = p.c; s.nb1 = p.nh; s.bias2 = p.d;
  s.n_out = p.nhid(); s.n_in1 = p.N; s.n_in2 = 0;
  s.stacked_units = 1;
  return s;
}

int run_energy(const StackedParams& p, const double* v, int64_t rows, double* E, cudaStream_t st) {
  if (rows == 0) return QCB_OK;
  return launch_rows_affine<0>(stacked_hidden_spec(p), v, nullptr, p.b, rows, E, st);
}

int run_linear_sigmoid(const double* M, int m_rows, int m_cols, int transpose_m, const double* bias, const double* x,
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
  if (M2 != nullptr && !transpose_m) return fail(QCB_ERR_BAD_ARG, "linear_sigmoid: second operand needs transpose_m=1");
  return launch_rows_affine<1>(s, x, x2, nullptr, rows, out, st);
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

// ---- reduced gradient ---------------------------------------------------------------------------
// g = [acc] + scale * sum_r -[p_r (x) v_r ; v_r ; p_r].  Persistent CTAs walk tiles of GR rows:
// phase 1 (warp per row) computes p = sigmoid(Wst v + hb) into shared memory, phase 2 lets every
// thread accumulate its 4x4 sub-tiles of the (nhid x N) outer-product sum in registers.  Per-CTA
// partials go to the workspace and a second kernel sums them in a fixed order (deterministic).
constexpr int GR = 32;  // rows per tile

//
// ENUM = true turns the same kernel into the exact negative phase of
// nn_states/neural_state.py:406-413: row r is basis state (first + r) generated from its index and
// every row is weighted by p(r) = exp(-E(r) - logZ).
template <int TPT, bool ENUM>  // TPT = 4x4 tiles per thread
__global__ void __launch_bounds__(ROWS_THREADS) grad_reduce_kernel(AffineSpec s, const double* __restrict__ v, int64_t rows,
                                                                   double* __restrict__ partials, uint64_t first,
                                                                   double logZ, const double* __restrict__ vb) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* Ws = reinterpret_cast<double*>(smem_raw);
  const int N = s.n_in1, H = s.n_out;
  const int LD = N | 1;
  const int Np = round_up(N, 4), Hp = round_up(H, 4);
  double* bs = Ws + round_up(H * LD, 2);
  double* Vt = bs + round_up(H, 2);     // [GR][Np]
  double* Pt = Vt + (size_t)GR * Np;    // [GR][Hp]
  double* wts = Pt + (size_t)GR * Hp;   // [GR] row weights (ENUM)
  double* vbs = wts + GR;               // [N] visible bias (ENUM)
  stage_affine(s, Ws, LD, bs);
  if (ENUM)
    for (int i = threadIdx.x; i < N; i += blockDim.x) vbs[i] = vb[i];
  for (int i = threadIdx.x; i < GR * Np; i += blockDim.x) Vt[i] = 0.0;
  for (int i = threadIdx.x; i < GR * Hp; i += blockDim.x) Pt[i] = 0.0;
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tiles_j = Np / 4, ntiles = (Hp / 4) * tiles_j;
  double acc[TPT][16];
#pragma unroll
  for (int m = 0; m < TPT; ++m)
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[m][e] = 0.0;
  double bsum = 0.0;  // thread tid < N: sum_r v[r][tid]; N <= tid < N+H: sum_r p[r][tid-N]

  const int64_t ntile_rows = (rows + GR - 1) / GR;
  for (int64_t tr = blockIdx.x; tr < ntile_rows; tr += gridDim.x) {
    const int64_t r0 = tr * GR;
    // phase 1
    for (int rr = warp; rr < GR; rr += ROWS_WARPS) {
      const int64_t r = r0 + rr;
      double* vrow = Vt + rr * Np;
      double* prow = Pt + rr * Hp;
      if (r < rows) {
        if (ENUM) {
          const uint64_t idx = first + (uint64_t)r;
          for (int i = lane; i < N; i += 32) vrow[i] = ((idx >> (N - 1 - i)) & 1ull) ? 1.0 : 0.0;
        } else {
          for (int i = lane; i < N; i += 32) vrow[i] = v[r * N + i];
        }
        __syncwarp();
        double xpart = 0.0;
        for (int o = lane; o < H; o += 32) {
          const double* wrow = Ws + (size_t)o * LD;
          double th = bs[o];
          for (int i = 0; i < N; ++i) th = fma(wrow[i], vrow[i], th);
          prow[o] = fmin(fmax(sigmoid_exact(th), 0.0), 1.0);
          if (ENUM) xpart += softplus20(th);
        }
        if (ENUM) {
          for (int i = lane; i < N; i += 32) xpart = fma(vbs[i], vrow[i], xpart);
          const double wgt = exp(warp_sum(xpart) - logZ);
          for (int o = lane; o < H; o += 32) prow[o] *= wgt;
          if (lane == 0) wts[rr] = wgt;
        }
      } else {
        if (ENUM && lane == 0) wts[rr] = 0.0;
        for (int i = lane; i < N; i += 32) vrow[i] = 0.0;
        for (int o = lane; o < H; o += 32) prow[o] = 0.0;
      }
    }
    __syncthreads();
    // phase 2
#pragma unroll
    for (int m = 0; m < TPT; ++m) {
      const int tile = threadIdx.x + m * ROWS_THREADS;
      if (tile < ntiles) {
        const int ib = tile / tiles_j, jb = tile - ib * tiles_j;
        const double* pp = Pt + 4 * ib;
        const double* vp = Vt + 4 * jb;
#pragma unroll 4
        for (int rr = 0; rr < GR; ++rr) {
          const double2 p01 = *reinterpret_cast<const double2*>(pp + rr * Hp);
          const double2 p23 = *reinterpret_cast<const double2*>(pp + rr * Hp + 2);
          const double2 v01 = *reinterpret_cast<const double2*>(vp + rr * Np);
          const double2 v23 = *reinterpret_cast<const double2*>(vp + rr * Np + 2);
          const double pv[4] = {p01.x, p01.y, p23.x, p23.y};
          const double vv[4] = {v01.x, v01.y, v23.x, v23.y};
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[m][a * 4 + b] = fma(pv[a], vv[b], acc[m][a * 4 + b]);
        }
      }
    }
    if (threadIdx.x < N) {
      if (ENUM) {
        for (int rr = 0; rr < GR; ++rr) bsum = fma(wts[rr], Vt[rr * Np + threadIdx.x], bsum);
      } else {
        for (int rr = 0; rr < GR; ++rr) bsum += Vt[rr * Np + threadIdx.x];
      }
    } else if (threadIdx.x < N + H) {
      for (int rr = 0; rr < GR; ++rr) bsum += Pt[rr * Hp + (threadIdx.x - N)];
    }
    __syncthreads();
  }

  // write this CTA's partial in flat order [Wst (H x N) | vb (N) | hb (H)]
  double* out = partials + (size_t)blockIdx.x * ((size_t)H * N + N + H);
#pragma unroll
  for (int m = 0; m < TPT; ++m) {
    const int tile = threadIdx.x + m * ROWS_THREADS;
    if (tile < ntiles) {
      const int ib = tile / tiles_j, jb = tile - ib * tiles_j;
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int i = 4 * ib + a, j = 4 * jb + b;
          if (i < H && j < N) out[(size_t)i * N + j] = acc[m][a * 4 + b];
        }
    }
  }
  if (threadIdx.x < N + H) out[(size_t)H * N + threadIdx.x] = bsum;
}

// g[e] = (accumulate ? g[e] : 0) + scale * (-(sum over CTAs of partial[e])).
__global__ void grad_finalize_kernel(const double* __restrict__ partials, int nparts, int N, int nh, int na, double scale,
                                     int accumulate, double* __restrict__ g) {
  const int H = nh + na;
  const int G = H * N + N + H;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= G) return;
  double sum = 0.0;
  for (int p = 0; p < nparts; ++p) sum += partials[(size_t)p * G + e];
  // the stacked order [W rows, U rows | b | c, d] is the reference's flat order [W, U, b, c, d]
  const double val = -scale * sum;
  g[e] = accumulate ? g[e] + val : val;
}

static size_t grad_reduce_smem(int N, int H) {
  return ((size_t)round_up(H * (N | 1), 2) + round_up(H, 2) + (size_t)GR * round_up(N, 4) + (size_t)GR * round_up(H, 4) + GR + N) *
         sizeof(double);
}

static int grad_reduce_grid(int64_t rows) {
  const int64_t need = (rows + GR - 1) / GR;
  const int64_t cap = (int64_t)sm_count() * 2;
  return (int)(need < cap ? (need < 1 ? 1 : need) : cap);
}

size_t grad_reduce_workspace_bytes(int N, int nhid, int64_t rows) {
  const size_t G = (size_t)nhid * N + N + nhid;
  return align_up((size_t)sm_count() * 2 * G * sizeof(double), 256);
}

template <bool ENUM>
static int run_grad_reduce_impl(const StackedParams& p, const double* v, int64_t rows, double scale, double* g,
                                int accumulate, uint64_t first, double logZ, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int N = p.N, H = p.nhid();
  const int G = H * N + N + H;
  if (N + H > ROWS_THREADS) return fail(QCB_ERR_UNSUPPORTED, "grad_reduce: N + nh + na = %d > %d", N + H, ROWS_THREADS);
  const size_t smem = grad_reduce_smem(N, H);
  if (smem > SMEM_LIMIT) return fail(QCB_ERR_UNSUPPORTED, "grad_reduce: layer %d x %d does not fit in shared memory", H, N);
  const size_t need = grad_reduce_workspace_bytes(N, H, rows);
  if (ws == nullptr || ws_bytes < need) return fail(QCB_ERR_BAD_ARG, "grad_reduce: workspace too small (%zu < %zu)", ws_bytes, need);
  const int grid = grad_reduce_grid(rows);
  double* partials = reinterpret_cast<double*>(ws);
  if (rows > 0) {
    AffineSpec s = stacked_hidden_spec(p);
    const int ntiles = (round_up(H, 4) / 4) * (round_up(N, 4) / 4);
    const int tpt = (ntiles + ROWS_THREADS - 1) / ROWS_THREADS;
#define QCB_GR_CASE(T)                                                                                                        \
  if (tpt <= T) {                                                                                                             \
    QCB_CUDA(cudaFuncSetAttribute(grad_reduce_kernel<T, ENUM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT)); \
    grad_reduce_kernel<T, ENUM><<<grid, ROWS_THREADS, smem, st>>>(s, v, rows, partials, first, logZ, p.b);                    \
    QCB_LAUNCH_CHECK();                                                                                                       \
  } else
    QCB_GR_CASE(1) QCB_GR_CASE(2) QCB_GR_CASE(4)
    return fail(QCB_ERR_UNSUPPORTED, "grad_reduce: %d x %d weight matrix too large", H, N);
#undef QCB_GR_CASE
  }
  grad_finalize_kernel<<<(G + 255) / 256, 256, 0, st>>>(partials, rows > 0 ? grid : 0, N, p.nh, p.na, scale, accumulate, g);
  QCB_LAUNCH_CHECK();
  return QCB_OK;
}

int run_grad_reduce(const StackedParams& p, const double* v, int64_t rows, double scale, double* g, int accumulate,
                    void* ws, size_t ws_bytes, cudaStream_t st) {
  return run_grad_reduce_impl<false>(p, v, rows, scale, g, accumulate, 0, 0.0, ws, ws_bytes, st);
}

// g += sum_{r in range} p(r) grad E(r)  with p(r) = exp(-E(r) - logZ)
int run_exact_grad_range(const StackedParams& p, uint64_t first, uint64_t count, double logZ, double* g, void* ws,
                         size_t ws_bytes, cudaStream_t st) {
  if (count > (uint64_t)INT64_MAX) return fail(QCB_ERR_BAD_ARG, "exact_grad: range too large");
  return run_grad_reduce_impl<true>(p, nullptr, (int64_t)count, 1.0, g, 1, first, logZ, ws, ws_bytes, st);
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

}  // namespace qcb
