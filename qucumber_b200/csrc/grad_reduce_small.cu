// Reduced energy gradient for SMALL layers (N, nh + na <= 16) and large batches of explicit rows:
//   g_partial = sum_rows [ p_r (x) v_r ; v_r ; p_r ],  p_r = sigmoid(Wst v_r + hb)        (rbm/binary_rbm.py:113-120)
// At N = nh = 10 a row is 80 bytes and ~470 FP64 operations: the tensor-core kernel (grad_reduce_dmma.cu) pads such a
// layer to 16 x 16 tiles and spends its time on tile bookkeeping (0.163 ms for 2^20 rows against an HBM floor of 0.013 ms).
// Here a CTA streams tiles of 512 rows through shared memory (one contiguous, coalesced copy per tile):
//   phase A  one thread = two rows: theta from the weights in shared memory (every load a warp-wide broadcast, each feeding
//            both rows), the table logistic of common.cuh, p into shared memory;
//   phase B  the outer products P^T V on the FP64 tensor cores (mma.sync m8n8k4): warp w owns rows 64 w .. 64 w + 63 of the
//            tile and keeps its share of the (at most 2 x 2) output blocks in registers for the whole kernel; the two bias
//            gradients are column sums taken by 8 (N + H) threads (one per column and 64-row slice).
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

  // phase B: G (H x N) += P^T (H x rows) . V (rows x N) as FP64 tensor-core tiles (mma.sync m8n8k4); warp w takes rows
  // 64 w .. 64 w + 63 of every tile and keeps its share of all (HT / 8) x (NT / 8) output blocks in registers.  The bias
  // gradients are column sums: thread c < N + H adds column c of [V | P] over the tile's rows.
  constexpr int MB = (HT + 7) / 8, NB = (NT + 7) / 8;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  double G[MB][NB][2];
#pragma unroll
  for (int m = 0; m < MB; ++m)
#pragma unroll
    for (int n = 0; n < NB; ++n) G[m][n][0] = G[m][n][1] = 0.0;
  double colsum = 0.0;
  const int cs_col = threadIdx.x % (N + H), cs_slice = threadIdx.x / (N + H);  // column of [V | P] and 64-row slice of the tile
  const int cs_ld = cs_col < N ? N : H;
  const int E = H * N + N + H;
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
      // two hidden units x two rows at a time: four independent FMA chains and four independent logistic evaluations keep
      // the FP64 pipe fed with 16 warps per SM (one chain per thread left 2.5 warps per issue waiting on a dependency)
#pragma unroll
      for (int i = 0; i < HT; i += 2) {
        if (i < H) {
          double ta0 = hbs[i], tb0 = ta0, ta1 = hbs[i + 1], tb1 = ta1;   // hbs and Ws are zero-padded to HT
#pragma unroll
          for (int j = 0; j < NT; ++j) {
            const double w0 = Ws[i * NT + j], w1 = Ws[(i + 1) * NT + j];
            ta0 = fma(w0, va[j], ta0);
            tb0 = fma(w0, vb2[j], tb0);
            ta1 = fma(w1, va[j], ta1);
            tb1 = fma(w1, vb2[j], tb1);
          }
          const double sa0 = sigmoid_tab(ta0, e2tab), sb0 = sigmoid_tab(tb0, e2tab);
          const double sa1 = sigmoid_tab(ta1, e2tab), sb1 = sigmoid_tab(tb1, e2tab);
          pa[i] = live_a ? sa0 : 0.0;   // rows beyond the batch contribute nothing
          pb[i] = live_b ? sb0 : 0.0;
          if (i + 1 < H) {
            pa[i + 1] = live_a ? sa1 : 0.0;
            pb[i + 1] = live_b ? sb1 : 0.0;
          }
        }
      }
    }
    __syncthreads();

    // ---- phase B: tensor-core outer products over this warp's 64 rows, column sums for the biases
    {
      const int rw = 64 * warp;
#pragma unroll 4
      for (int kk = 0; kk < 16; ++kk) {
        const int r = rw + 4 * kk + t;
        double afr[MB], bfr[NB];
#pragma unroll
        for (int m = 0; m < MB; ++m) afr[m] = 8 * m + g < H ? Pt[(size_t)r * H + 8 * m + g] : 0.0;
#pragma unroll
        for (int n = 0; n < NB; ++n) bfr[n] = 8 * n + g < N ? Vt[(size_t)r * N + 8 * n + g] : 0.0;
#pragma unroll
        for (int m = 0; m < MB; ++m)
#pragma unroll
          for (int n = 0; n < NB; ++n) dmma884(G[m][n][0], G[m][n][1], afr[m], bfr[n]);
      }
      if (cs_slice < 8) {  // rows beyond the batch are zero in both tiles
        const double* col = (cs_col < N ? Vt + cs_col : Pt + (cs_col - N)) + (size_t)(64 * cs_slice) * cs_ld;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll 4
        for (int r = 0; r < 64; r += 4) {
          a0 += col[(size_t)r * cs_ld]; a1 += col[(size_t)(r + 1) * cs_ld]; a2 += col[(size_t)(r + 2) * cs_ld]; a3 += col[(size_t)(r + 3) * cs_ld];
        }
        colsum += (a0 + a1) + (a2 + a3);
      }
    }
    __syncthreads();
  }

  // ---- this CTA's partial: the eight warps' shares of G are added in a fixed order through shared memory
  double* out = partials + (size_t)blockIdx.x * E;
  double* red = Vt;  // [8 warps][8 MB][8 NB]: the register blocks are whole 8 x 8 tiles (HT, NT are multiples of 4 only)
#pragma unroll
  for (int m = 0; m < MB; ++m)
#pragma unroll
    for (int n = 0; n < NB; ++n)
#pragma unroll
      for (int e2 = 0; e2 < 2; ++e2) red[((size_t)warp * (8 * MB) + 8 * m + g) * (8 * NB) + 8 * n + 2 * t + e2] = G[m][n][e2];
  __syncthreads();
  for (int idx = threadIdx.x; idx < H * N; idx += GS_THREADS) {
    const int i = idx / N, j = idx - i * N;
    double sum = 0.0;
#pragma unroll
    for (int w = 0; w < GS_THREADS / 32; ++w) sum += red[((size_t)w * (8 * MB) + i) * (8 * NB) + j];
    out[idx] = sum;
  }
  __syncthreads();
  if (cs_slice < 8) red[cs_slice * (N + H) + cs_col] = colsum;
  __syncthreads();
  if (threadIdx.x < N + H) {  // [vb (N) | hb (H)] follow the weights
    double sum = 0.0;
#pragma unroll
    for (int sl = 0; sl < 8; ++sl) sum += red[sl * (N + H) + threadIdx.x];
    out[H * N + threadIdx.x] = sum;
  }
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
