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

int g_grad_small_hist = 1;  // qcb_set_grad_reduce_mode(1 or 3) turns the histogram path off (the streaming kernel alone)
constexpr int GS_THREADS = 256;
constexpr int GS_TILE = 2 * GS_THREADS;  // rows per tile

template <int NT, int HT>
__global__ void __launch_bounds__(GS_THREADS, 2) grad_reduce_small_kernel(AffineSpec s, const double* __restrict__ v, int64_t rows,
                                                                          double* __restrict__ partials, const int* __restrict__ gate) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = s.n_in1, H = s.n_out;
  if (gate != nullptr && *gate == 0) {  // queued behind the histogram path (below), which took the batch: this slot counts as zero
    for (int idx = threadIdx.x; idx < H * N + N + H; idx += GS_THREADS) partials[(size_t)blockIdx.x * (H * N + N + H) + idx] = 0.0;
    return;
  }
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

// ---- binary rows of a very small visible layer: histogram, then one weighted pass over the 2^N states ----------------------
// At N = 10 a batch of 2^20 rows holds at most 1 024 DISTINCT rows.  sum_rows grad E(v_r) = sum_states count(s) grad E(s), so
// the batch is only COUNTED (one coalesced read of the rows: the HBM floor of this reduction, 80 B per row) and the logistic
// function is evaluated once per distinct state.  The counts are integers, so the result is the row sum in another order.
// A row that is not 0/1 raises *flag; the streaming kernel queued behind (gate = flag) then redoes the batch.
constexpr int GH_THREADS = 256;
constexpr int GH_MAX_N = 12;

__global__ void __launch_bounds__(GH_THREADS) state_hist_kernel(const double* __restrict__ v, int64_t rows, int N, unsigned int* __restrict__ hist,
                                                                int* __restrict__ flag) {
  extern __shared__ unsigned int sh[];                 // [2^N] counters, then [8 warps][GH_MAX_N] ballot words
  const int nstates = 1 << N;
  unsigned int* words = sh + nstates + (threadIdx.x >> 5) * (GH_MAX_N + 1);
  for (int i = threadIdx.x; i < nstates; i += GH_THREADS) sh[i] = 0u;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t nchunks = (rows + 31) >> 5;             // a warp takes 32 rows = 32 N consecutive doubles at a time
  const int64_t total = rows * N;
  bool nonbinary = false;
  for (int64_t chunk = (int64_t)blockIdx.x * (GH_THREADS / 32) + (threadIdx.x >> 5); chunk < nchunks;
       chunk += (int64_t)gridDim.x * (GH_THREADS / 32)) {
    const int64_t base = chunk * 32 * N;
    unsigned long long d[GH_MAX_N];
#pragma unroll
    for (int i = 0; i < GH_MAX_N; ++i) {                // all loads first: N x 256 contiguous bytes in flight per warp
      const int64_t at = base + 32 * i + lane;
      d[i] = (i < N && at < total) ? __ldg(reinterpret_cast<const unsigned long long*>(v) + at) : 0ull;
    }
#pragma unroll
    for (int i = 0; i < GH_MAX_N; ++i) {
      if (i < N) {
        nonbinary |= (d[i] << 1) != 0ull && d[i] != 0x3FF0000000000000ull;
        const unsigned int w = __ballot_sync(0xffffffffu, (d[i] << 1) != 0ull);  // bit l = double 32 i + l of the chunk
        if (lane == 0) words[i] = w;
      }
    }
    if (lane == 0) words[N] = 0u;
    __syncwarp();
    // lane l owns row 32 chunk + l: bits [l N, l N + N) of the chunk's bit string, unit j at bit j; the state index of the
    // enumeration has unit 0 as its MOST significant bit (nn_states/neural_state.py:173-176)
    const int at = lane * N;
    const unsigned int raw = __funnelshift_r(words[at >> 5], words[(at >> 5) + 1], at & 31) & (unsigned int)(nstates - 1);
    const unsigned int key = chunk * 32 + lane < rows ? (__brev(raw) >> (32 - N)) : 0xffffffffu;
    const unsigned int peers = __match_any_sync(0xffffffffu, key);  // equal rows of the warp are counted by one atomic
    if (key != 0xffffffffu && lane == __ffs(peers) - 1) atomicAdd(sh + key, (unsigned int)__popc(peers));
    __syncwarp();
  }
  if (nonbinary) *flag = 3;
  __syncthreads();
  for (int i = threadIdx.x; i < nstates; i += GH_THREADS)
    if (sh[i] != 0u) atomicAdd(hist + i, sh[i]);
}

// partial sums of count(s) [p(s) (x) s ; s ; p(s)] over GH_THREADS states per CTA, in the flat order of the other kernels
__global__ void __launch_bounds__(GH_THREADS) grad_hist_kernel(AffineSpec s, const unsigned int* __restrict__ hist, const int* __restrict__ flag,
                                                               double* __restrict__ partials) {
  const int N = s.n_in1, H = s.n_out;
  const int E = H * N + N + H;
  double* out = partials + (size_t)blockIdx.x * E;
  __shared__ double Ws[16 * GH_MAX_N], hbs[16];
  __shared__ double Pt[GH_THREADS][17];               // count * p, column 16 = the count itself
  if (*flag != 0) {                                   // the batch was not binary: the gated kernel behind this one redoes it
    for (int idx = threadIdx.x; idx < E; idx += GH_THREADS) out[idx] = 0.0;
    return;
  }
  for (int idx = threadIdx.x; idx < H * N; idx += GH_THREADS) Ws[idx] = s.mop(idx / N, idx % N);
  if (threadIdx.x < H) hbs[threadIdx.x] = s.bias(threadIdx.x);
  __syncthreads();
  const unsigned int state = blockIdx.x * GH_THREADS + threadIdx.x;   // unit j = bit N - 1 - j
  const unsigned int count = state < (1u << N) ? hist[state] : 0u;
  const double w = (double)count;
  for (int i = 0; i < H; ++i) {
    double pw = 0.0;
    if (count != 0u) {
      double th = hbs[i];
      for (int j = 0; j < N; ++j) th += ((state >> (N - 1 - j)) & 1u) ? Ws[i * N + j] : 0.0;
      pw = w * fmin(fmax(sigmoid_exact(th), 0.0), 1.0);
    }
    Pt[threadIdx.x][i] = pw;
  }
  Pt[threadIdx.x][16] = w;
  __syncthreads();
  // entry e of the flat gradient: a sum over the CTA's states in a fixed order
  for (int e = threadIdx.x; e < E; e += GH_THREADS) {
    int col, bit;  // column of Pt, and the unit whose bit selects the states (-1: all)
    if (e < H * N) { col = e / N; bit = e - col * N; }
    else if (e < H * N + N) { col = 16; bit = e - H * N; }
    else { col = e - H * N - N; bit = -1; }
    const unsigned int sel = bit < 0 ? 0u : 1u << (N - 1 - bit);
    const unsigned int s0 = blockIdx.x * GH_THREADS;
    double a0 = 0.0, a1 = 0.0;
    for (int r = 0; r < GH_THREADS; r += 2) {
      a0 += (bit < 0 || ((s0 + r) & sel)) ? Pt[r][col] : 0.0;
      a1 += (bit < 0 || ((s0 + r + 1) & sel)) ? Pt[r + 1][col] : 0.0;
    }
    out[e] = a0 + a1;
  }
}

size_t grad_hist_scratch_bytes() { return ((size_t)1 << GH_MAX_N) * sizeof(unsigned int) + 256; }

// *grid_out > 0 when this kernel took the batch (explicit rows, no aux features, N and H <= 16, >= 4096 rows)
// hist_scratch: grad_hist_scratch_bytes() of device memory (or NULL): enables the histogram path for N <= 12 and rows >= 16 * 2^N
int try_grad_reduce_small(const AffineSpec& s, const RowSrc& src, int mode, int H, int64_t rows, double* partials, int max_grid,
                          int* grid_out, void* hist_scratch, cudaStream_t st) {
  *grid_out = 0;
  const int N = s.n_in1;
  if (mode != ROWS_EXPLICIT || H != s.n_out || rows < 4096 || N > 16 || H > 16 || N < 1 || H < 1) return QCB_OK;
  const int* gate = nullptr;
  int first_slot = 0;
  if (hist_scratch != nullptr && g_grad_small_hist && N <= GH_MAX_N && rows >= ((int64_t)16 << N) && rows < ((int64_t)1 << 31)) {
    const int nstates = 1 << N;
    unsigned int* hist = reinterpret_cast<unsigned int*>(hist_scratch);
    int* flag = reinterpret_cast<int*>(hist + nstates);
    QCB_CUDA(cudaMemsetAsync(hist, 0, (size_t)(nstates + 1) * sizeof(unsigned int), st));
    const size_t hsmem = ((size_t)nstates + (GH_THREADS / 32) * (GH_MAX_N + 1)) * sizeof(unsigned int);
    const int64_t nchunks = (rows + 31) / 32, want = (nchunks + GH_THREADS / 32 - 1) / (GH_THREADS / 32), hcap = (int64_t)sm_count() * 4;
    state_hist_kernel<<<(int)(want < hcap ? want : hcap), GH_THREADS, hsmem, st>>>(src.v, rows, N, hist, flag);
    QCB_LAUNCH_CHECK();
    first_slot = (nstates + GH_THREADS - 1) / GH_THREADS;
    grad_hist_kernel<<<first_slot, GH_THREADS, 0, st>>>(s, hist, flag, partials);
    QCB_LAUNCH_CHECK();
    gate = flag;
    partials += (size_t)first_slot * ((size_t)H * N + N + H);
    max_grid -= first_slot;
  }
  const int NT = (N + 3) / 4 * 4, HT = (H + 3) / 4 * 4;
  const size_t smem = gs_smem(N, H);
  const int64_t ntiles = (rows + GS_TILE - 1) / GS_TILE;
  const int64_t cap = (int64_t)sm_count() * 2;
  int grid = (int)(ntiles < cap ? ntiles : cap);
  if (grid > max_grid) grid = max_grid;
#define QCB_GS_CASE(NTx, HTx)                                                                                                  \
  if (NT == NTx && HT == HTx) {                                                                                                \
    QCB_CUDA(cudaFuncSetAttribute(grad_reduce_small_kernel<NTx, HTx>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    grad_reduce_small_kernel<NTx, HTx><<<grid, GS_THREADS, smem, st>>>(s, src.v, rows, partials, gate);                        \
    QCB_LAUNCH_CHECK();                                                                                                        \
    *grid_out = first_slot + grid;                                                                                             \
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
