// Layers beyond the shared-memory-resident kernels (the reference has no size limits: rbm/binary_rbm.py:48-72 allocates any
// nh x N).  The fused kernels keep a layer's weights in shared memory, which ends at about 25 000 weights; the kernels here
// leave the weights in global memory / L2 and tile the ROWS instead:
//   rows_affine_tiled_kernel   energies, conditionals and raw affine maps (rbm/binary_rbm.py:75-96, 129-168): a CTA stages
//                              32 rows in shared memory, lane = row, a warp walks four output units at a time and reads
//                              their weights as warp-wide broadcasts (one L1/L2 transaction feeds 32 rows)
//   outer_reduce_tiled_kernel  the reduced gradient's outer products sum_r a_r (x) v_r (rbm/binary_rbm.py:113-120) as 64 x 64
//                              output tiles with 4 x 4 register blocks, rows split over CTAs, partials summed in a fixed order
//   grad_batch_tiled / flip_tiled   per-row gradients and single-flip log-ratios, one CTA per row
// Same arithmetic as the resident kernels (sequential FMA over the inputs, exact logistic / thresholded softplus), so results
// agree to rounding; they are there for completeness, the resident kernels are the fast path.  FP64 FMA bound by design:
// 2 n_in n_out flop per row against n_in * 8 bytes.
#include "rows.cuh"

namespace qcb {

constexpr int LG_ROWS = 32;
constexpr int LG_THREADS = 256;
constexpr int LG_WARPS = LG_THREADS / 32;

// value j of absolute row r of a row source (rows.cuh)
template <int SRC>
__device__ __forceinline__ double src_value(const RowSrc& src, int N, int64_t r, int j) {
  if (SRC == ROWS_ENUM) return (((src.first + (uint64_t)r) >> (N - 1 - j)) & 1ull) ? 1.0 : 0.0;
  if (SRC == ROWS_EXPANDED) return src.ex.value(r, j);
  return src.v[r * N + j];
}

// where hidden-like unit o keeps its weights: Mop[o][i] = p1[i * st1] for i < n_in1, p2[(i - n_in1) * st2] beyond
__device__ __forceinline__ void unit_rows(const AffineSpec& s, int o, const double*& p1, int& st1, const double*& p2, int& st2) {
  if (s.stacked_units) {
    p1 = o < s.m1_rows ? s.M1 + (size_t)o * s.m1_cols : s.M2 + (size_t)(o - s.m1_rows) * s.m2_cols;
    st1 = 1;
    p2 = nullptr;
    st2 = 0;
  } else {
    p1 = s.t1 ? s.M1 + o : s.M1 + (size_t)o * s.m1_cols;
    st1 = s.t1 ? s.m1_cols : 1;
    p2 = s.M2 == nullptr ? nullptr : (s.t2 ? s.M2 + o : s.M2 + (size_t)o * s.m2_cols);
    st2 = s.t2 ? s.m2_cols : 1;
  }
}

// MODE 0: E[r] = -(vb . x_r + sum_o softplus(theta_o));  MODE 1: out[r][o] = clamp(sigmoid(theta_o));  MODE 2: out[r][o] = theta_o.
// Rows are row0 + [0, rows) of the source (x2: second input block of explicit rows, or NULL); out is indexed by the LOCAL row.
template <int MODE, int SRC>
__global__ void __launch_bounds__(LG_THREADS) rows_affine_tiled_kernel(AffineSpec s, RowSrc src, const double* __restrict__ x2,
                                                                       const double* __restrict__ vb, int64_t row0, int64_t rows,
                                                                       double* __restrict__ out, int ld_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* Xs = reinterpret_cast<double*>(smem_raw);  // [LG_ROWS][LDX]
  const int n1 = s.n_in1, n2 = s.n_in2, nin = n1 + n2, LDX = nin | 1;
  double* red = Xs + (size_t)LG_ROWS * LDX;          // [LG_WARPS][32] (MODE 0)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t ntiles = (rows + LG_ROWS - 1) / LG_ROWS;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t r0 = tile * LG_ROWS;
    const int nr = (int)(rows - r0 < LG_ROWS ? rows - r0 : (int64_t)LG_ROWS);
    for (int idx = threadIdx.x; idx < LG_ROWS * nin; idx += LG_THREADS) {
      const int r = idx / nin, i = idx - r * nin;
      double val = 0.0;
      if (r < nr) val = i < n1 ? src_value<SRC>(src, n1, row0 + r0 + r, i) : x2[(row0 + r0 + r) * n2 + (i - n1)];
      Xs[(size_t)r * LDX + i] = val;
    }
    __syncthreads();
    const double* xrow = Xs + (size_t)lane * LDX;
    double part = 0.0;
    for (int o4 = 4 * warp; o4 < s.n_out; o4 += 4 * LG_WARPS) {
      const double *p1[4], *p2[4];
      int st1[4], st2[4];
      double t[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int o = o4 + q < s.n_out ? o4 + q : s.n_out - 1;
        unit_rows(s, o, p1[q], st1[q], p2[q], st2[q]);
        t[q] = s.bias(o);
      }
      for (int i = 0; i < n1; ++i) {
        const double xv = xrow[i];
#pragma unroll
        for (int q = 0; q < 4; ++q) t[q] = fma(__ldg(p1[q] + (size_t)i * st1[q]), xv, t[q]);
      }
      for (int i = 0; i < n2; ++i) {
        const double xv = xrow[n1 + i];
#pragma unroll
        for (int q = 0; q < 4; ++q) t[q] = fma(__ldg(p2[q] + (size_t)i * st2[q]), xv, t[q]);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (o4 + q < s.n_out) {
          if (MODE == 0) part += softplus20(t[q]);
          else if (lane < nr) out[(size_t)(r0 + lane) * ld_out + o4 + q] = MODE == 1 ? fmin(fmax(sigmoid_exact(t[q]), 0.0), 1.0) : t[q];
        }
      }
    }
    if (MODE == 0) {
      red[warp * 32 + lane] = part;
      __syncthreads();
      if (warp == 0 && lane < nr) {
        double sum = 0.0;
#pragma unroll
        for (int w = 0; w < LG_WARPS; ++w) sum += red[w * 32 + lane];
        for (int i = 0; i < nin; ++i) sum = fma(vb[i], xrow[i], sum);
        out[r0 + lane] = -sum;
      }
    }
    __syncthreads();
  }
}

static size_t tiled_smem(int nin) { return ((size_t)LG_ROWS * (nin | 1) + LG_WARPS * 32) * sizeof(double); }
bool rows_tiled_fits(int nin) { return tiled_smem(nin) <= SMEM_LIMIT; }

template <int MODE, int SRC>
static int launch_rows_tiled(const AffineSpec& s, const RowSrc& src, const double* x2, const double* vb, int64_t row0, int64_t rows,
                             double* out, int ld_out, cudaStream_t st) {
  if (rows == 0) return QCB_OK;
  const size_t smem = tiled_smem(s.n_in());
  if (smem > SMEM_LIMIT)
    return fail(QCB_ERR_UNSUPPORTED, "layer %d x %d: more than %d input units per row tile", s.n_out, s.n_in(), 767);
  QCB_CUDA(cudaFuncSetAttribute(rows_affine_tiled_kernel<MODE, SRC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
  const int64_t ntiles = (rows + LG_ROWS - 1) / LG_ROWS;
  const int per_sm = (int)(SMEM_LIMIT / smem) < 4 ? (int)(SMEM_LIMIT / smem) : 4;
  const int64_t cap = (int64_t)sm_count() * (per_sm < 1 ? 1 : per_sm);
  rows_affine_tiled_kernel<MODE, SRC><<<(int)(ntiles < cap ? ntiles : cap), LG_THREADS, smem, st>>>(s, src, x2, vb, row0, rows, out, ld_out);
  QCB_LAUNCH_CHECK();
  return QCB_OK;
}

// explicit rows (x [| x2]) -> energy / sigmoid / affine, any layer with at most 767 inputs
int run_rows_affine_tiled(int mode, const AffineSpec& s, const double* x, const double* x2, const double* vb, int64_t rows, double* out,
                          cudaStream_t st) {
  RowSrc src{};
  src.v = x;
  if (mode == 0) return launch_rows_tiled<0, ROWS_EXPLICIT>(s, src, x2, vb, 0, rows, out, 1, st);
  if (mode == 1) return launch_rows_tiled<1, ROWS_EXPLICIT>(s, src, x2, vb, 0, rows, out, s.n_out, st);
  return launch_rows_tiled<2, ROWS_EXPLICIT>(s, src, x2, vb, 0, rows, out, s.n_out, st);
}

// energies of expanded rows row0 + [0, rows) (the rotated-basis amplitudes of layers beyond the fused kernels)
int run_energy_expanded_tiled(const StackedParams& p, const Expansion& ex, int64_t row0, int64_t rows, double* E, cudaStream_t st) {
  RowSrc src{};
  src.ex = ex;
  return launch_rows_tiled<0, ROWS_EXPANDED>(stacked_hidden_spec(p), src, nullptr, p.b, row0, rows, E, 1, st);
}

// energies of the basis states first + [0, rows) (exact enumeration of layers with more than 64 hidden-like units)
int run_energy_enum_tiled(const StackedParams& p, uint64_t first, int64_t rows, double* E, cudaStream_t st) {
  RowSrc src{};
  src.first = first;
  return launch_rows_tiled<0, ROWS_ENUM>(stacked_hidden_spec(p), src, nullptr, p.b, 0, rows, E, 1, st);
}

// ---- reduced gradient -------------------------------------------------------------------------------------------------
// G_ext = A^T B over the chunk's rows with A = [w p | aux | w] (rows x (H + 1)), B = [v | 1] (rows x (N + 1)): the weight block,
// the hidden-bias sums (column N) and the visible-bias sums (row H) of the flat gradient [Wst (H x N) | vb (N) | hb (H)] in one
// pass.  grid = (output tiles, row splits); partial slot = blockIdx.y.
constexpr int OT = 64;   // output tile
constexpr int OK_ROWS = 16;  // rows per shared-memory step

template <int SRC>
__global__ void __launch_bounds__(LG_THREADS) outer_reduce_tiled_kernel(const double* __restrict__ P, int Hc, int H, int N, RowSrc src,
                                                                        const double* __restrict__ wloc, int64_t row0, int64_t rows,
                                                                        double* __restrict__ partials, int first_chunk) {
  __shared__ double As[OK_ROWS][OT + 4], Bs[OK_ROWS][OT + 4];
  const int tiles_j = (N + 1 + OT - 1) / OT;
  const int ti = blockIdx.x / tiles_j, tj = blockIdx.x - ti * tiles_j;
  const int i0 = ti * OT, j0 = tj * OT;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t per = (rows + gridDim.y - 1) / gridDim.y;
  const int64_t rbeg = (int64_t)blockIdx.y * per, rend = rbeg + per < rows ? rbeg + per : rows;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
  for (int64_t rs = rbeg; rs < rend; rs += OK_ROWS) {
    for (int idx = threadIdx.x; idx < OK_ROWS * OT; idx += LG_THREADS) {
      const int rr = idx / OT, c = idx - rr * OT;
      const int64_t r = rs + rr;  // local row of the chunk
      double a = 0.0, b = 0.0;
      if (r < rend) {
        // weight of the row: 1 (explicit), the caller's (expanded) or exp(-E - logZ) from wloc (enumeration)
        const double w = SRC == ROWS_EXPLICIT ? 1.0 : (SRC == ROWS_EXPANDED ? src.wrow[row0 + r] : wloc[r]);
        const int i = i0 + c, j = j0 + c;
        if (i < Hc) a = P[(size_t)r * Hc + i] * w;
        else if (i < H) a = src.aux[(row0 + r) * src.n_aux + (i - Hc)];
        else if (i == H) a = w;
        if (j < N) b = src_value<SRC>(src, N, row0 + r, j);
        else if (j == N) b = 1.0;
      }
      As[rr][c] = a;
      Bs[rr][c] = b;
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < OK_ROWS; ++rr) {
      double av[4], bv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) av[a] = As[rr][4 * ty + a];
#pragma unroll
      for (int b = 0; b < 4; ++b) bv[b] = Bs[rr][4 * tx + b];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
    }
    __syncthreads();
  }
  double* out = partials + (size_t)blockIdx.y * ((size_t)H * N + N + H);
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int i = i0 + 4 * ty + a, j = j0 + 4 * tx + b;
      size_t e;
      if (i < H && j < N) e = (size_t)i * N + j;
      else if (i == H && j < N) e = (size_t)H * N + j;
      else if (i < H && j == N) e = (size_t)H * N + N + i;
      else continue;
      out[e] = first_chunk ? acc[a][b] : out[e] + acc[a][b];
    }
}

__global__ void enum_weights_kernel(const double* __restrict__ E, int64_t rows, double logZ, double* __restrict__ w) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < rows) w[r] = exp(-E[r] - logZ);
}

int run_grad_finalize(const double* partials, int nparts, int G, double scale, int accumulate, int zero_from, double* g, cudaStream_t st);

static int64_t large_chunk_rows(int N, int Hc) {
  int64_t c = ((int64_t)64 << 20) / ((int64_t)(Hc + 2) * 8);  // about 64 MiB of probabilities per chunk
  c = c / LG_ROWS * LG_ROWS;
  return c < 1024 ? 1024 : (c > (1 << 20) ? (1 << 20) : c);
}
static int large_splits(int N, int H) {
  const int tiles = ((H + 1 + OT - 1) / OT) * ((N + 1 + OT - 1) / OT);
  int s = (2 * sm_count() + tiles - 1) / tiles;
  return s < 1 ? 1 : (s > 64 ? 64 : s);
}
bool grad_reduce_needs_tiled(int N, int Hc, int H, size_t resident_smem) {
  return N > 128 || N + H > ROWS_THREADS || resident_smem > SMEM_LIMIT || ((round_up(H, 4) / 4) * (round_up(N, 4) / 4) + ROWS_THREADS - 1) / ROWS_THREADS > 4;
}
size_t grad_reduce_tiled_workspace_bytes(int N, int H, int64_t rows) {
  const size_t G = (size_t)H * N + N + H;
  int64_t chunk = large_chunk_rows(N, H);
  if (rows > 0 && rows < chunk) chunk = (rows + LG_ROWS - 1) / LG_ROWS * LG_ROWS;
  return align_up((size_t)large_splits(N, H) * G * sizeof(double), 256) + align_up((size_t)chunk * H * sizeof(double), 256) +
         2 * align_up((size_t)chunk * sizeof(double), 256);
}

template <int SRC>
int run_grad_reduce_tiled(const AffineSpec& s, const RowSrc& src, int H, int64_t rows, double scale, double* g, int accumulate, int zero_from,
                          void* ws, size_t ws_bytes, cudaStream_t st) {
  const int N = s.n_in1, Hc = s.n_out;
  const int G = H * N + N + H;
  if (!rows_tiled_fits(N)) return fail(QCB_ERR_UNSUPPORTED, "grad_reduce: N = %d > 767", N);
  const size_t need = grad_reduce_tiled_workspace_bytes(N, H, rows);
  if (ws == nullptr || ws_bytes < need) return fail(QCB_ERR_BAD_ARG, "grad_reduce: workspace too small (%zu < %zu)", ws_bytes, need);
  const int splits = large_splits(N, H);
  int64_t chunk = large_chunk_rows(N, H);
  if (rows > 0 && rows < chunk) chunk = (rows + LG_ROWS - 1) / LG_ROWS * LG_ROWS;
  char* base = static_cast<char*>(ws);
  double* partials = reinterpret_cast<double*>(base);
  double* P = reinterpret_cast<double*>(base + align_up((size_t)splits * G * sizeof(double), 256));
  double* Eloc = reinterpret_cast<double*>(reinterpret_cast<char*>(P) + align_up((size_t)chunk * H * sizeof(double), 256));
  double* wloc = reinterpret_cast<double*>(reinterpret_cast<char*>(Eloc) + align_up((size_t)chunk * sizeof(double), 256));
  const int tiles = ((H + 1 + OT - 1) / OT) * ((N + 1 + OT - 1) / OT);
  for (int64_t r0 = 0; r0 < rows; r0 += chunk) {
    const int64_t rc = rows - r0 < chunk ? rows - r0 : chunk;
    if (int rcode = launch_rows_tiled<1, SRC>(s, src, nullptr, nullptr, r0, rc, P, Hc, st)) return rcode;
    if (SRC == ROWS_ENUM) {
      if (int rcode = launch_rows_tiled<0, SRC>(s, src, nullptr, src.vb, r0, rc, Eloc, 1, st)) return rcode;
      enum_weights_kernel<<<(unsigned)((rc + 255) / 256), 256, 0, st>>>(Eloc, rc, src.logZ, wloc);
      QCB_LAUNCH_CHECK();
    }
    outer_reduce_tiled_kernel<SRC><<<dim3(tiles, splits), LG_THREADS, 0, st>>>(P, Hc, H, N, src, wloc, r0, rc, partials, r0 == 0 ? 1 : 0);
    QCB_LAUNCH_CHECK();
  }
  return run_grad_finalize(partials, rows > 0 ? splits : 0, G, scale, accumulate, zero_from, g, st);
}
template int run_grad_reduce_tiled<ROWS_EXPLICIT>(const AffineSpec&, const RowSrc&, int, int64_t, double, double*, int, int, void*, size_t, cudaStream_t);
template int run_grad_reduce_tiled<ROWS_ENUM>(const AffineSpec&, const RowSrc&, int, int64_t, double, double*, int, int, void*, size_t, cudaStream_t);
template int run_grad_reduce_tiled<ROWS_EXPANDED>(const AffineSpec&, const RowSrc&, int, int64_t, double, double*, int, int, void*, size_t, cudaStream_t);

// ---- per-row gradients (reduce = False) ----------------------------------------------------------------------------------
// out[r] = -[p (x) v ; v ; p].  p is first written (positive) into its own slot of the row by the tiled sigmoid pass; one CTA
// per row then expands it.
__global__ void __launch_bounds__(LG_THREADS) grad_batch_expand_kernel(const double* __restrict__ v, int N, int H, int64_t rows,
                                                                       double* __restrict__ out) {
  const size_t G = (size_t)H * N + N + H;
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    double* orow = out + (size_t)r * G;
    const double* p = orow + (size_t)H * N + N;
    const double* vr = v + r * N;
    for (size_t e = threadIdx.x; e < (size_t)H * N; e += LG_THREADS) {
      const int i = (int)(e / N), j = (int)(e - (size_t)i * N);
      orow[e] = -(p[i] * vr[j]);
    }
    for (int j = threadIdx.x; j < N; j += LG_THREADS) orow[(size_t)H * N + j] = -vr[j];
    __syncthreads();  // every product has read p
    for (int i = threadIdx.x; i < H; i += LG_THREADS) orow[(size_t)H * N + N + i] = -p[i];
    __syncthreads();
  }
}

int run_grad_batch_tiled(const StackedParams& p, const double* v, int64_t rows, double* g_rows, cudaStream_t st) {
  const int N = p.N, H = p.nhid();
  const size_t G = (size_t)H * N + N + H;
  if (G > (size_t)INT32_MAX) return fail(QCB_ERR_UNSUPPORTED, "grad_batch: %d x %d parameters per row", H, N);
  RowSrc src{};
  src.v = v;
  if (int rc = launch_rows_tiled<1, ROWS_EXPLICIT>(stacked_hidden_spec(p), src, nullptr, nullptr, 0, rows, g_rows + (size_t)H * N + N, (int)G, st))
    return rc;
  const int64_t cap = (int64_t)sm_count() * 8;
  grad_batch_expand_kernel<<<(int)(rows < cap ? rows : cap), LG_THREADS, 0, st>>>(v, N, H, rows, g_rows);
  QCB_LAUNCH_CHECK();
  return QCB_OK;
}

// ---- single-spin-flip log-ratios ---------------------------------------------------------------------------------------------
// out[r][i] = E(v_r) - E(v_r with unit i flipped) = sgn b_i + sum_o [softplus(theta_o + sgn W_oi) - softplus(theta_o)], sgn = 1 - 2 v_i
// (the term-by-term form of rows_flip_kernel, energy_grad.cu).  One CTA per row: theta by warps (coalesced weight rows), then
// thread i walks column i of the weights (consecutive threads read consecutive addresses).
__global__ void __launch_bounds__(LG_THREADS) flip_tiled_kernel(AffineSpec s, const double* __restrict__ v, const double* __restrict__ vb,
                                                                int64_t rows, double* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = s.n_in1, H = s.n_out;
  double* vs = reinterpret_cast<double*>(smem_raw);  // [N]
  double* th = vs + N;                               // [H]
  double* sp = th + H;                               // [H] softplus(theta)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    for (int j = threadIdx.x; j < N; j += LG_THREADS) vs[j] = v[r * N + j];
    __syncthreads();
    for (int o = warp; o < H; o += LG_WARPS) {
      const double *p1, *p2;
      int st1, st2;
      unit_rows(s, o, p1, st1, p2, st2);
      double t = 0.0;
      for (int i = lane; i < N; i += 32) t = fma(p1[(size_t)i * st1], vs[i], t);
      t = warp_sum(t) + s.bias(o);
      if (lane == 0) {
        th[o] = t;
        sp[o] = softplus20(t);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += LG_THREADS) {
      const double sgn = 1.0 - 2.0 * vs[i];
      double d = sgn * vb[i];
      for (int o = 0; o < H; ++o) {
        const double *p1, *p2;
        int st1, st2;
        unit_rows(s, o, p1, st1, p2, st2);
        d += softplus20(fma(sgn, p1[(size_t)i * st1], th[o])) - sp[o];
      }
      out[r * N + i] = d;
    }
    __syncthreads();
  }
}

int run_flip_logratio_tiled(const StackedParams& p, const double* v, int64_t rows, double* out, cudaStream_t st) {
  const int N = p.N, H = p.nhid();
  const size_t smem = ((size_t)N + 2 * H) * sizeof(double);
  if (smem > SMEM_LIMIT) return fail(QCB_ERR_UNSUPPORTED, "flip_logratio: N + 2 (nh + na) = %d exceeds one CTA's shared memory", N + 2 * H);
  QCB_CUDA(cudaFuncSetAttribute(flip_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
  const int64_t cap = (int64_t)sm_count() * 4;
  flip_tiled_kernel<<<(int)(rows < cap ? rows : cap), LG_THREADS, smem, st>>>(stacked_hidden_spec(p), v, p.b, rows, out);
  QCB_LAUNCH_CHECK();
  return QCB_OK;
}

}  // namespace qcb
