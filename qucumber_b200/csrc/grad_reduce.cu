// Reduced energy gradients: g = [acc] + scale * sum_r w_r * grad E(v_r), grad E(v) = -[p (x) v ; v ; p]
// with p = sigmoid(Wst v + hb) (rbm/binary_rbm.py:98-120, rbm/purification_rbm.py:160-185).
//
// One kernel serves every caller that needs "sum over rows of weighted outer products":
//   ROWS_EXPLICIT  rows given as a (rows, N) tensor, weight 1            (positive / negative phase)
//   ROWS_ENUM      rows are basis states generated from their index, weight exp(-E - logZ)
//                  (exact negative phase, nn_states/neural_state.py:406-413)
//   ROWS_EXPANDED  rows are samples expanded over their rotated sites with externally computed
//                  weights and optional extra hidden-like features (rotated-basis gradients,
//                  nn_states/complex_wavefunction.py:162-211, nn_states/density_matrix.py:281-330)
//
// Persistent CTAs walk tiles of GR rows: phase 1 (warp per row) builds the row v and its feature
// vector p in shared memory, phase 2 lets every thread accumulate its 4x4 sub-tiles of the
// (H x N) outer-product sum in registers.  Per-CTA partials go to the workspace and a second kernel
// sums them in a fixed order, so results are run-to-run deterministic.
#include "rows.cuh"

namespace qcb {

constexpr int GR = 32;  // rows per tile
constexpr int64_t GR_DIRECT_MAX_ROWS = 8 * GR;  // up to this many rows one CTA does the whole reduction and finalises in place
constexpr int GR_MAX_CTAS_PER_SM = 4;  // partial-sum slots per SM in the workspace (tensor-core kernel: up to 4 CTAs/SM)

int try_grad_reduce_small(const AffineSpec& s, const RowSrc& src, int mode, int H, int64_t rows, double* partials, int max_grid,
                          int* grid_out, void* hist_scratch, cudaStream_t st);
size_t grad_hist_scratch_bytes();
int try_grad_reduce_dmma(const AffineSpec& s, const RowSrc& src, int mode, int H, int64_t rows, double* partials, int max_grid,
                         int* grid_out, int* flag, unsigned char* umma_img, cudaStream_t st);
size_t grad_umma_workspace_bytes();

template <int TPT, int SRC>  // TPT = 4x4 tiles per thread
__global__ void __launch_bounds__(ROWS_THREADS) grad_reduce_kernel(AffineSpec s, RowSrc src, int64_t rows,
                                                                   double* __restrict__ partials, GradDirect direct) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* Ws = reinterpret_cast<double*>(smem_raw);
  const int N = s.n_in1, Hc = s.n_out;             // Hc: units whose p is computed here
  const int H = Hc + (SRC == ROWS_EXPANDED ? src.n_aux : 0);
  const int LD = N | 1;
  const int Np = round_up(N, 4), Hp = round_up(H, 4);
  double* bs = Ws + round_up(Hc * LD, 2);
  double* Vt = bs + round_up(Hc, 2);    // [GR][Np]
  double* Pt = Vt + (size_t)GR * Np;    // [GR][Hp]
  double* wts = Pt + (size_t)GR * Hp;   // [GR] row weights
  double* vbs = wts + GR;               // [N] visible bias (ROWS_ENUM)
  stage_affine(s, Ws, LD, bs);
  if (SRC == ROWS_ENUM)
    for (int i = threadIdx.x; i < N; i += blockDim.x) vbs[i] = src.vb[i];
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
  double bsum = 0.0;  // thread tid < N: sum_r w_r v[r][tid]; N <= tid < N+H: sum_r p[r][tid-N]

  const int64_t ntile_rows = (rows + GR - 1) / GR;
  for (int64_t tr = blockIdx.x; tr < ntile_rows; tr += gridDim.x) {
    const int64_t r0 = tr * GR;
    // phase 1
    for (int rr = warp; rr < GR; rr += ROWS_WARPS) {
      const int64_t r = r0 + rr;
      double* vrow = Vt + rr * Np;
      double* prow = Pt + rr * Hp;
      if (r < rows) {
        if (SRC == ROWS_ENUM) {
          const uint64_t idx = src.first + (uint64_t)r;
          for (int i = lane; i < N; i += 32) vrow[i] = ((idx >> (N - 1 - i)) & 1ull) ? 1.0 : 0.0;
        } else if (SRC == ROWS_EXPANDED) {
          for (int i = lane; i < N; i += 32) vrow[i] = src.ex.value(r, i);
        } else {
          for (int i = lane; i < N; i += 32) vrow[i] = src.v[r * N + i];
        }
        __syncwarp();
        double xpart = 0.0;
        for (int o = lane; o < Hc; o += 32) {
          const double* wrow = Ws + (size_t)o * LD;
          double th = bs[o];
          for (int i = 0; i < N; ++i) th = fma(wrow[i], vrow[i], th);
          prow[o] = fmin(fmax(sigmoid_exact(th), 0.0), 1.0);
          if (SRC == ROWS_ENUM) xpart += softplus20(th);
        }
        if (SRC == ROWS_ENUM) {
          for (int i = lane; i < N; i += 32) xpart = fma(vbs[i], vrow[i], xpart);
          const double wgt = exp(warp_sum(xpart) - src.logZ);
          for (int o = lane; o < Hc; o += 32) prow[o] *= wgt;
          if (lane == 0) wts[rr] = wgt;
        } else if (SRC == ROWS_EXPANDED) {
          const double wgt = src.wrow[r];
          for (int o = lane; o < Hc; o += 32) prow[o] *= wgt;
          for (int o = lane; o < src.n_aux; o += 32) prow[Hc + o] = src.aux[r * src.n_aux + o];
          if (lane == 0) wts[rr] = wgt;
        }
      } else {
        if (SRC != ROWS_EXPLICIT && lane == 0) wts[rr] = 0.0;
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
      if (SRC != ROWS_EXPLICIT) {
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
          if (i < H && j < N) direct.put(out, (size_t)i * N + j, acc[m][a * 4 + b]);
        }
    }
  }
  if (threadIdx.x < N + H) direct.put(out, (size_t)H * N + threadIdx.x, bsum);
}

template <int SRC>
__global__ void __launch_bounds__(ROWS_THREADS) grad_reduce_tiny_kernel(AffineSpec s, RowSrc src, int rows, GradDirect direct) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  grad_reduce_tiny_body<SRC>(s, src, rows, direct, smem_raw);
}

// The amplitude and the phase RBM of a rotated-basis gradient take the same expanded rows with their own weights: two CTAs
// of ONE launch (each what grad_reduce_tiny_kernel<ROWS_EXPANDED> does for its RBM).
__global__ void __launch_bounds__(ROWS_THREADS) grad_reduce_tiny_pair_kernel(TinyJob j0, TinyJob j1, int rows) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const TinyJob& job = blockIdx.x == 0 ? j0 : j1;
  grad_reduce_tiny_body<ROWS_EXPANDED>(job.s, job.src, rows, job.direct, smem_raw);
}


// g[e] = (accumulate ? g[e] : 0) + scale * (-(sum over CTAs of partial[e])); entries >= zero_from are
// forced to contribute nothing (the phase RBM's aux bias never receives a gradient,
// nn_states/density_matrix.py:200-211).  The stacked order [W rows, U rows | b | c, d] is the
// reference's flat order [W, U, b, c, d].
__global__ void grad_finalize_kernel(const double* __restrict__ partials, int nparts, int G, double scale, int accumulate,
                                     int zero_from, double* __restrict__ g) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= G) return;
  double sum = 0.0;
  for (int p = 0; p < nparts; ++p) sum += partials[(size_t)p * G + e];
  const double val = e >= zero_from ? 0.0 : -scale * sum;
  g[e] = accumulate ? g[e] + val : val;
}

int run_grad_finalize(const double* partials, int nparts, int G, double scale, int accumulate, int zero_from, double* g, cudaStream_t st) {
  grad_finalize_kernel<<<(G + 255) / 256, 256, 0, st>>>(partials, nparts, G, scale, accumulate, zero_from, g);
  QCB_LAUNCH_CHECK();
  return QCB_OK;
}

// layers beyond the resident kernels: large.cu
bool grad_reduce_needs_tiled(int N, int Hc, int H, size_t resident_smem);
size_t grad_reduce_tiled_workspace_bytes(int N, int H, int64_t rows);
template <int SRC>
int run_grad_reduce_tiled(const AffineSpec& s, const RowSrc& src, int H, int64_t rows, double scale, double* g, int accumulate, int zero_from,
                          void* ws, size_t ws_bytes, cudaStream_t st);

static size_t grad_reduce_smem(int N, int Hc, int H) {
  return ((size_t)round_up(Hc * (N | 1), 2) + round_up(Hc, 2) + (size_t)GR * round_up(N, 4) + (size_t)GR * round_up(H, 4) + GR + N) *
         sizeof(double);
}

static int grad_reduce_grid(int64_t rows) {
  const int64_t need = (rows + GR - 1) / GR;
  const int64_t cap = (int64_t)sm_count() * 2;
  return (int)(need < cap ? (need < 1 ? 1 : need) : cap);
}

size_t grad_reduce_workspace_bytes(int N, int nhid, int64_t rows) {
  (void)rows;
  const size_t G = (size_t)nhid * N + N + nhid;
  // partials + the weight image of the tcgen05 kernel + one flag word
  const size_t resident = align_up((size_t)sm_count() * GR_MAX_CTAS_PER_SM * G * sizeof(double), 256) + grad_umma_workspace_bytes() + 256;
  if (!grad_reduce_needs_tiled(N, nhid, nhid, grad_reduce_smem(N, nhid, nhid))) return resident;
  const size_t tiled = grad_reduce_tiled_workspace_bytes(N, nhid, rows);  // row-tiled kernels of large.cu
  return resident > tiled ? resident : tiled;
}

template <int SRC>
static int launch_grad_reduce(const AffineSpec& s, const RowSrc& src, int H, int64_t rows, double scale, double* g,
                              int accumulate, int zero_from, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int N = s.n_in1, Hc = s.n_out;
  const int G = H * N + N + H;
  const size_t smem = grad_reduce_smem(N, Hc, H);
  if (grad_reduce_needs_tiled(N, Hc, H, smem)) {  // layers beyond the resident kernels: weights stay in global memory (large.cu)
    if (src.bits != nullptr) return fail(QCB_ERR_UNSUPPORTED, "grad_reduce: packed rows need the tcgen05 kernel");
    return run_grad_reduce_tiled<SRC>(s, src, H, rows, scale, g, accumulate, zero_from, ws, ws_bytes, st);
  }
  const size_t need = grad_reduce_workspace_bytes(N, H, rows);
  if (ws == nullptr || ws_bytes < need) return fail(QCB_ERR_BAD_ARG, "grad_reduce: workspace too small (%zu < %zu)", ws_bytes, need);
  int grid = grad_reduce_grid(rows);
  double* partials = reinterpret_cast<double*>(ws);
  int grid_dmma = 0;
  if (rows > 0) {  // small layers, large batches of explicit rows: the streaming kernel of grad_reduce_small.cu
    // its histogram path borrows the scratch of the tcgen05 weight image (that kernel only takes layers of 33..64 units)
    void* hist_scratch = grad_umma_workspace_bytes() >= grad_hist_scratch_bytes() ? reinterpret_cast<char*>(ws) + need - 256 - grad_umma_workspace_bytes() : nullptr;
    const int rc = try_grad_reduce_small(s, src, SRC, H, rows, partials, sm_count() * GR_MAX_CTAS_PER_SM, &grid_dmma, hist_scratch, st);
    if (rc != QCB_OK) return rc;
  }
  if (rows > 0 && grid_dmma == 0) {
    // large batches of explicit / expanded rows run on the FP64 tensor cores (grad_reduce_dmma.cu)
    int* flag = reinterpret_cast<int*>(reinterpret_cast<char*>(ws) + need - 256);
    unsigned char* umma_img = reinterpret_cast<unsigned char*>(ws) + need - 256 - grad_umma_workspace_bytes();
    const int rc = try_grad_reduce_dmma(s, src, SRC, H, rows, partials, sm_count() * GR_MAX_CTAS_PER_SM, &grid_dmma, flag, umma_img, st);
    if (rc != QCB_OK) return rc;
  }
  GradDirect direct{nullptr, scale, accumulate, zero_from};
  if (grid_dmma > 0) {
    grid = grid_dmma;
  } else if (rows > 0) {
    if (rows <= GR_DIRECT_MAX_ROWS) {  // one CTA walks the few row tiles and finalises in place
      grid = 1;
      direct.g = g;
      if (SRC != ROWS_ENUM && src.bits == nullptr && N <= 16 && H <= 16 && grad_reduce_tiny_smem(N, Hc, H, rows) <= 48 * 1024) {
        grad_reduce_tiny_kernel<SRC><<<1, ROWS_THREADS, grad_reduce_tiny_smem(N, Hc, H, rows), st>>>(s, src, (int)rows, direct);
        QCB_LAUNCH_CHECK();
        return QCB_OK;
      }
    }
    const int ntiles = (round_up(H, 4) / 4) * (round_up(N, 4) / 4);
    const int tpt = (ntiles + ROWS_THREADS - 1) / ROWS_THREADS;
#define QCB_GR_CASE(T)                                                                                                       \
  if (tpt <= T) {                                                                                                            \
    QCB_CUDA(cudaFuncSetAttribute(grad_reduce_kernel<T, SRC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT)); \
    grad_reduce_kernel<T, SRC><<<grid, ROWS_THREADS, smem, st>>>(s, src, rows, partials, direct);                            \
    QCB_LAUNCH_CHECK();                                                                                                      \
  } else
    QCB_GR_CASE(1) QCB_GR_CASE(2) QCB_GR_CASE(4)
    return fail(QCB_ERR_UNSUPPORTED, "grad_reduce: %d x %d weight matrix too large", H, N);
#undef QCB_GR_CASE
  }
  if (direct.g != nullptr) return QCB_OK;  // the single CTA has written g itself
  grad_finalize_kernel<<<(G + 255) / 256, 256, 0, st>>>(partials, rows > 0 ? grid : 0, G, scale, accumulate, zero_from, g);
  QCB_LAUNCH_CHECK();
  return QCB_OK;
}

int run_grad_reduce(const StackedParams& p, const double* v, int64_t rows, double scale, double* g, int accumulate,
                    void* ws, size_t ws_bytes, cudaStream_t st) {
  RowSrc src{};
  src.v = v;
  const int H = p.nhid();
  return launch_grad_reduce<ROWS_EXPLICIT>(stacked_hidden_spec(p), src, H, rows, scale, g, accumulate, H * p.N + p.N + H, ws,
                                           ws_bytes, st);
}

// The same reduction over PACKED binary rows (v_k of a CD step straight from the Gibbs kernel); scratch_rows: rows x N doubles
// that the FP64 stand-in reads if the tcgen05 kernel gives the batch up.  Only for shapes grad_reduce_bits_supported() accepts.
bool grad_reduce_bits_supported(int N, int H, int64_t rows);
int run_grad_reduce_bits(const StackedParams& p, const uint8_t* bits, double* scratch_rows, int64_t rows, double scale, double* g,
                         int accumulate, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int H = p.nhid();
  if (!grad_reduce_bits_supported(p.N, H, rows)) return fail(QCB_ERR_UNSUPPORTED, "grad_reduce: packed rows not supported for this shape");
  RowSrc src{};
  src.v = scratch_rows;
  src.bits = bits;
  return launch_grad_reduce<ROWS_EXPLICIT>(stacked_hidden_spec(p), src, H, rows, scale, g, accumulate, H * p.N + p.N + H, ws,
                                           ws_bytes, st);
}

// g += sum_{r in range} p(r) grad E(r)  with p(r) = exp(-E(r) - logZ)
int run_exact_grad_range(const StackedParams& p, uint64_t first, uint64_t count, double logZ, double* g, void* ws,
                         size_t ws_bytes, cudaStream_t st) {
  if (count > (uint64_t)INT64_MAX) return fail(QCB_ERR_BAD_ARG, "exact_grad: range too large");
  RowSrc src{};
  src.first = first;
  src.logZ = logZ;
  src.vb = p.b;
  const int H = p.nhid();
  return launch_grad_reduce<ROWS_ENUM>(stacked_hidden_spec(p), src, H, (int64_t)count, 1.0, g, 1, H * p.N + p.N + H, ws, ws_bytes,
                                       st);
}

// g += scale * sum_rows [ w_r grad E_hidden(v_r) with the aux block of p replaced by aux features ]
// (rows = samples expanded over their rotated sites).  zero_tail > 0 zeroes the last zero_tail
// entries of the flat gradient.
int run_grad_reduce_expanded(const StackedParams& p, const Expansion& ex, int64_t rows, const double* wrow,
                             const double* aux, int n_aux, int zero_tail, double scale, double* g, void* ws, size_t ws_bytes,
                             cudaStream_t st) {
  RowSrc src{};
  src.ex = ex;
  src.wrow = wrow;
  src.aux = aux;
  src.n_aux = n_aux;
  const int H = p.nh + n_aux;
  const int G = H * p.N + p.N + H;
  return launch_grad_reduce<ROWS_EXPANDED>(hidden_only_spec(p), src, H, rows, scale, g, 1, G - zero_tail, ws, ws_bytes, st);
}

// Both RBMs of a rotated-basis gradient over the same expanded rows: g_am += scale * (...) with (w_am, aux_am), g_ph likewise
// with (w_ph, aux_ph) and its last zero_tail_ph entries untouched.  A handful of rows of a small layer: one launch for both
// (values as from two run_grad_reduce_expanded calls); otherwise those two calls.
int run_grad_reduce_expanded_pair(const StackedParams& am, const StackedParams& ph, const Expansion& ex, int64_t rows, const double* w_am,
                                  const double* aux_am, const double* w_ph, const double* aux_ph, int n_aux, int zero_tail_ph,
                                  double scale, double* g_am, double* g_ph, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int N = am.N, Hc = am.nh, H = am.nh + n_aux;
  const int G = H * N + N + H;
  if (rows > 0 && rows <= GR_DIRECT_MAX_ROWS && N <= 16 && H <= 16 && ph.N == N && ph.nh == Hc &&
      grad_reduce_tiny_smem(N, Hc, H, rows) <= 48 * 1024) {
    TinyJob j0{}, j1{};
    j0.s = hidden_only_spec(am);
    j1.s = hidden_only_spec(ph);
    j0.src.ex = ex; j0.src.wrow = w_am; j0.src.aux = aux_am; j0.src.n_aux = n_aux;
    j1.src.ex = ex; j1.src.wrow = w_ph; j1.src.aux = aux_ph; j1.src.n_aux = n_aux;
    j0.direct = GradDirect{g_am, scale, 1, G};
    j1.direct = GradDirect{g_ph, scale, 1, G - zero_tail_ph};
    grad_reduce_tiny_pair_kernel<<<2, ROWS_THREADS, grad_reduce_tiny_smem(N, Hc, H, rows), st>>>(j0, j1, (int)rows);
    QCB_LAUNCH_CHECK();
    return QCB_OK;
  }
  int rc = run_grad_reduce_expanded(am, ex, rows, w_am, aux_am, n_aux, 0, scale, g_am, ws, ws_bytes, st);
  if (rc != QCB_OK) return rc;
  return run_grad_reduce_expanded(ph, ex, rows, w_ph, aux_ph, n_aux, zero_tail_ph, scale, g_ph, ws, ws_bytes, st);
}

}  // namespace qcb
