// Reduced energy gradient for SMALL layers (N, nh + na <= 16) and large batches of explicit rows:
//   g_partial = sum_rows [ p_r (x) v_r ; v_r ; p_r ],  p_r = sigmoid(Wst v_r + hb)        (rbm/binary_rbm.py:113-120)
// At N = nh = 10 a row is 80 bytes and ~470 FP64 operations: the tensor-core kernel (grad_reduce_dmma.cu) pads such a
// layer to 16 x 16 tiles and spends its time on tile bookkeeping (0.163 ms for 2^20 rows against an HBM floor of 0.013 ms).
// Here a CTA streams tiles of 512 rows through shared memory (one contiguous, coalesced copy per tile):
//   phase A  one thread = two rows: theta from the weights in shared memory (every load a warp-wide broadcast, each feeding
//            both rows), the table logistic of common.cuh, p into shared memory;
//   phase B  one thread = one entry of the flat gradient x one half of the tile's rows: the outer products and the two
//            bias sums are plain dot products over the rows, accumulated in a register for the whole kernel.
// Rows may hold arbitrary reals (the reference's contract).  Per-CTA partials go to the workspace in the flat order
// [Wst (H x N) | vb (N) | hb (H)] and are summed by grad_finalize_kernel in a fixed order, like the other two kernels.
#include "rows.cuh"

namespace qcb {

constexpr int GS_THREADS = 256;
constexpr int GS_TILE = 2 * GS_THREADS;  // rows per tile

template <int NT, int HT>
__global__ void __launch_bounds__(GS_THREADS, 2) grad_reduce_small_kernel(AffineSpec s, const double* __restrict__ v, int64_t rows,
                                                                          double* __restrict__ partials) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = s.n_in1, H = s.n_out;
  double* Vt = reinterpret_cast<double*>(smem_raw);   // [GS_TILE][N]  the tile as it lies in global memory
  double* Pt = Vt + (size_t)GS_TILE * N;              // [GS_TILE][H]
  __shared__ double Ws[HT * NT];                      // zero-padded weights
  __shared__ double hbs[HT];
  __shared__ double e2tab[EXP2_TABLE];
  for (int idx = threadIdx.x; idx < HT * NT; idx += GS_THREADS) {
    const int i = idx / NT, j = idx - i * NT;
    Ws[idx] = (i < H && j < N) ? s.mop(i, j) : 0.0;
  }
  for (int i = threadIdx.x; i < HT; i += GS_THREADS) hbs[i] = i < H ? s.bias(i) : 0.0;
  fill_exp2_table(e2tab);

  // phase B ownership: entry e of [Wst | vb | hb] and a half of the rows
  const int E = H * N + N + H;
  const int slices = GS_THREADS / E >= 2 ? 2 : 1;   // E <= 288; two row halves when the entries fit twice
  const int e = threadIdx.x % E, slice = threadIdx.x / E;
  const bool owner = slice < slices && (slices == 2 || threadIdx.x < E);
  // the two operands of this entry: a column of Pt (or the constant 1) and a column of Vt (or the constant 1)
  int pi = -1, vj = -1;
  if (e < H * N) { pi = e / N; vj = e - pi * N; }
  else if (e < H * N + N) vj = e - H * N;
  else pi = e - H * N - N;
  double acc = 0.0;
  // entries that do not fit one pass over the threads (E > 256) take a second one
  const int e2 = threadIdx.x + GS_THREADS;
  int pi2 = -1, vj2 = -1;
  const bool owner2 = slices == 1 && e2 < E;
  if (owner2) {
    if (e2 < H * N) { pi2 = e2 / N; vj2 = e2 - pi2 * N; }
    else if (e2 < H * N + N) vj2 = e2 - H * N;
    else pi2 = e2 - H * N - N;
  }
  double acc2 = 0.0;
  __syncthreads();

  const bool vec_ok = (reinterpret_cast<uintptr_t>(v) & 15) == 0 && ((GS_TILE * N) & 1) == 0;
  const int64_t ntiles = (rows + GS_TILE - 1) / GS_TILE;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t r0 = tile * GS_TILE;
    const int nrows = (int)(rows - r0 < GS_TILE ? rows - r0 : (int64_t)GS_TILE);
    const int ndoubles = nrows * N;
    const double* src = v + r0 * N;
    // ---- the tile is one contiguous span: coalesced 16-byte copies (r0 * N is even, so alignment follows the base)
    if (vec_ok) {
      const int nvec = ndoubles >> 1;
      for (int idx = threadIdx.x; idx < nvec; idx += GS_THREADS) reinterpret_cast<double2*>(Vt)[idx] = reinterpret_cast<const double2*>(src)[idx];
      if ((ndoubles & 1) && threadIdx.x == 0) Vt[ndoubles - 1] = src[ndoubles - 1];
    } else {
      for (int idx = threadIdx.x; idx < ndoubles; idx += GS_THREADS) Vt[idx] = src[idx];
    }
    for (int idx = ndoubles + threadIdx.x; idx < GS_TILE * N; idx += GS_THREADS) Vt[idx] = 0.0;  // ragged last tile
    __syncthreads();

    // ---- phase A: rows threadIdx.x and threadIdx.x + 256
    {
      double va[NT], vb2[NT];
      const double* ra = Vt + (size_t)threadIdx.x * N;
      const double* rb = Vt + (size_t)(threadIdx.x + GS_THREADS) * N;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        va[j] = j < N ? ra[j] : 0.0;
        vb2[j] = j < N ? rb[j] : 0.0;
      }
      double* pa = Pt + (size_t)threadIdx.x * H;
      double* pb = Pt + (size_t)(threadIdx.x + GS_THREADS) * H;
      const bool live_a = threadIdx.x < nrows, live_b = threadIdx.x + GS_THREADS < nrows;
#pragma unroll
      for (int i = 0; i < HT; ++i) {
        if (i < H) {
          double ta = hbs[i], tb = ta;
#pragma unroll
          for (int j = 0; j < NT; ++j) {
            const double w = Ws[i * NT + j];
            ta = fma(w, va[j], ta);
            tb = fma(w, vb2[j], tb);
          }
          pa[i] = live_a ? sigmoid_tab(ta, e2tab) : 0.0;   // rows beyond the batch contribute nothing
          pb[i] = live_b ? sigmoid_tab(tb, e2tab) : 0.0;
        }
      }
    }
    __syncthreads();

    // ---- phase B: this thread's entry over its rows
    if (owner) {
      const int rbeg = slices == 2 ? slice * (GS_TILE / 2) : 0, rend = slices == 2 ? rbeg + GS_TILE / 2 : GS_TILE;
      const int rlim = rend < nrows ? rend : nrows;
      const double* pp = pi >= 0 ? Pt + pi : nullptr;
      const double* vp = vj >= 0 ? Vt + vj : nullptr;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      int r = rbeg;
      if (pp != nullptr && vp != nullptr) {
        for (; r + 4 <= rlim; r += 4) {
          a0 = fma(pp[(size_t)r * H], vp[(size_t)r * N], a0);
          a1 = fma(pp[(size_t)(r + 1) * H], vp[(size_t)(r + 1) * N], a1);
          a2 = fma(pp[(size_t)(r + 2) * H], vp[(size_t)(r + 2) * N], a2);
          a3 = fma(pp[(size_t)(r + 3) * H], vp[(size_t)(r + 3) * N], a3);
        }
        for (; r < rlim; ++r) a0 = fma(pp[(size_t)r * H], vp[(size_t)r * N], a0);
      } else if (vp != nullptr) {
        for (; r + 4 <= rlim; r += 4) {
          a0 += vp[(size_t)r * N]; a1 += vp[(size_t)(r + 1) * N]; a2 += vp[(size_t)(r + 2) * N]; a3 += vp[(size_t)(r + 3) * N];
        }
        for (; r < rlim; ++r) a0 += vp[(size_t)r * N];
      } else {
        for (; r + 4 <= rlim; r += 4) {
          a0 += pp[(size_t)r * H]; a1 += pp[(size_t)(r + 1) * H]; a2 += pp[(size_t)(r + 2) * H]; a3 += pp[(size_t)(r + 3) * H];
        }
        for (; r < rlim; ++r) a0 += pp[(size_t)r * H];
      }
      acc += (a0 + a1) + (a2 + a3);
    }
    if (owner2) {
      const double* pp = pi2 >= 0 ? Pt + pi2 : nullptr;
      const double* vp = vj2 >= 0 ? Vt + vj2 : nullptr;
      double a0 = 0.0;
      for (int r = 0; r < nrows; ++r) a0 = fma(pp != nullptr ? pp[(size_t)r * H] : 1.0, vp != nullptr ? vp[(size_t)r * N] : 1.0, a0);
      acc2 += a0;
    }
    __syncthreads();
  }

  // ---- this CTA's partial: the two row halves are added in a fixed order through shared memory
  double* out = partials + (size_t)blockIdx.x * E;
  double* red = Vt;
  if (owner && slices == 2 && slice == 1) red[e] = acc;
  __syncthreads();
  if (owner && slice == 0) out[e] = slices == 2 ? acc + red[e] : acc;
  if (owner2) out[e2] = acc2;
}

static size_t gs_smem(int N, int H) { return (size_t)GS_TILE * (N + H) * sizeof(double); }

// *grid_out > 0 when this kernel took the batch (explicit rows, no aux features, N and H <= 16, >= 4096 rows)
int try_grad_reduce_small(const AffineSpec& s, const RowSrc& src, int mode, int H, int64_t rows, double* partials, int max_grid,
                          int* grid_out, cudaStream_t st) {
  *grid_out = 0;
  const int N = s.n_in1;
  if (mode != ROWS_EXPLICIT || H != s.n_out || rows < 4096 || N > 16 || H > 16 || N < 1 || H < 1) return QCB_OK;
  const int NT = (N + 3) / 4 * 4, HT = (H + 3) / 4 * 4;
  const size_t smem = gs_smem(N, H);
  const int64_t ntiles = (rows + GS_TILE - 1) / GS_TILE;
  const int64_t cap = (int64_t)sm_count() * 2;
  int grid = (int)(ntiles < cap ? ntiles : cap);
  if (grid > max_grid) grid = max_grid;
#define QCB_GS_CASE(NTx, HTx)                                                                                                  \
  if (NT == NTx && HT == HTx) {                                                                                                \
    QCB_CUDA(cudaFuncSetAttribute(grad_reduce_small_kernel<NTx, HTx>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    grad_reduce_small_kernel<NTx, HTx><<<grid, GS_THREADS, smem, st>>>(s, src.v, rows, partials);                              \
    QCB_LAUNCH_CHECK();                                                                                                        \
    *grid_out = grid;                                                                                                          \
    return QCB_OK;                                                                                                             \
  }
  QCB_GS_CASE(4, 4) QCB_GS_CASE(4, 8) QCB_GS_CASE(4, 12) QCB_GS_CASE(4, 16)
  QCB_GS_CASE(8, 4) QCB_GS_CASE(8, 8) QCB_GS_CASE(8, 12) QCB_GS_CASE(8, 16)
  QCB_GS_CASE(12, 4) QCB_GS_CASE(12, 8) QCB_GS_CASE(12, 12) QCB_GS_CASE(12, 16)
  QCB_GS_CASE(16, 4) QCB_GS_CASE(16, 8) QCB_GS_CASE(16, 12) QCB_GS_CASE(16, 16)
#undef QCB_GS_CASE
  return QCB_OK;
}

}  // namespace qcb
