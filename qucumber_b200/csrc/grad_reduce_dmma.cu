// Reduced energy gradient on the FP64 tensor cores (large batches).
//
// g_partial = sum_rows [ p_r (x) v_r ; w_r v_r ; p_r ],  p_r = w_r * sigmoid(Wst v_r + hb)   (rbm/binary_rbm.py:113-120)
// Both contractions are DMMA (mma.sync m8n8k4 f64) tiles:
//   phase 1  Theta(64 rows x H) = V(64 x N) . Wst^T          A = V tile (shared memory), B = W fragments
//   phase 2  G(H x N)          += P^T(H x 64) . V(64 x N)     A = P tile, B = V tile, accumulators in registers
// A CTA of 8 warps walks tiles of 64 rows (persistent grid).  In phase 1 warp w owns rows 8w..8w+7 and
// applies the exact FP64 sigmoid to its accumulator fragments; in phase 2 warp w owns the 8-row
// blocks {w, w+8} of G, which stay in registers for the whole kernel.  Per-CTA partials go to the
// workspace in the same [Wst | vb | hb] layout as the FMA kernel and are summed by
// grad_finalize_kernel in a fixed order (deterministic).  Shapes that do not fit, tiny batches and
// index-generated rows keep using grad_reduce_kernel (grad_reduce.cu).
#include "rows.cuh"

namespace qcb {

#ifndef QCB_GD_WARPS  // overridable for occupancy experiments
#define QCB_GD_WARPS 8
#define QCB_GD_MIN_CTAS 2
#endif
constexpr int GD_WARPS = QCB_GD_WARPS;
constexpr int GD_THREADS = 32 * GD_WARPS;
constexpr int GD_TR = 8 * GD_WARPS;  // rows per tile: warp w owns rows 8w .. 8w+7 in phase 1

// ---- phase 1 on the int8 tensor cores (IMMA = true) ----------------------------------------------
// For binary rows theta = hb + sum of SELECTED weights.  The weights are written once per CTA as three 23-bit
// fixed-point groups  W 2^s = q2 + q1 2^-22 + q0 2^-44  (q = rint of the running remainder, exact in FP64: every
// weight within 2^-13 of the largest is represented exactly, the rest to 2^-66 of it), each group as signed
// digits of 8 / 7 / 8 bits like gibbs_imma.cu.  T_g = sum_k v_k q_g,ik is then an exact int32 from three chained
// mma.sync.m16n8k32 (u8 x s8) passes, and theta = hb + 2^-s (T2 + 2^-22 (T1 + 2^-22 T0)) costs three FP64 FMAs
// per unit (the innermost exact) instead of N: closer to the exact sum than the FP64 FMA chain, and the FP64 pipe
// keeps only the sigmoid and the outer products.  A tile with a non-binary entry (or weights that do not fit the
// image) raises *flag and the CTA stops; the FP64 kernel queued behind then redoes the whole batch.
constexpr long long GD_WQ_MAX = 127LL * 32768 + 63 * 256 + 127;

__device__ __forceinline__ void imma16832_u8s8(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// exact int32 -> double: 2^52 + 2^31 + t as a bit pattern, minus the offset (I2F.F64 measured the same)
__device__ __forceinline__ double int_to_double(int t) {
  return __hiloint2double(0x43300000, (int)((uint32_t)t ^ 0x80000000u)) - 4503601774854144.0;
}

// Digit images Bq[group][digit][k-step][n-block][lane] (uint2 = the two B registers of one m16n8k32) and the scale
// exponent; returns false when the weights do not fit (non-finite, or outside 1e-290 .. 1e290).
template <int KK, int HB8>
__device__ bool stage_theta_digits(const AffineSpec& s, int N, int Hc, uint2* Bq, double* red, int* s_out) {
  double mw = 0.0;
  bool bad = false;
  for (int idx = threadIdx.x; idx < Hc * N; idx += blockDim.x) {
    const double a = fabs(s.mop(idx / N, idx % N));
    bad = bad || !(a < 1e290);
    mw = fmax(mw, a);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mw = fmax(mw, __shfl_xor_sync(0xffffffffu, mw, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mw;
  bad = __syncthreads_or(bad) != 0;
  mw = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) mw = fmax(mw, red[w]);
  if (bad || (mw > 0.0 && mw < 1e-290)) return false;
  int sc = 0;
  if (mw > 0.0) {
    sc = 21 - ilogb(mw);
    if (llrint(scalbn(mw, sc)) > GD_WQ_MAX) sc -= 1;  // keep the top digit inside int8
  }
  *s_out = sc;
  for (int idx = threadIdx.x; idx < KK * HB8 * 32 * 2; idx += blockDim.x) {
    const int reg = idx & 1, ln = (idx >> 1) & 31, nb = (idx >> 6) % HB8, kk = (idx >> 6) / HB8;
    const int hid = 8 * nb + (ln >> 2);
    uint32_t word[3][3] = {{0u, 0u, 0u}, {0u, 0u, 0u}, {0u, 0u, 0u}};
    for (int i = 0; i < 4; ++i) {
      const int vis = 32 * kk + 16 * reg + 4 * (ln & 3) + i;
      if (hid < Hc && vis < N) {
        double x = scalbn(s.mop(hid, vis), sc);
#pragma unroll
        for (int gq = 2; gq >= 0; --gq) {
          const double q = rint(x);
          x = (x - q) * 4194304.0;  // remainder, exact; next group is 2^22 finer
          long long wq = (long long)q;
          const int d0 = (int)(((wq + 128) & 255) - 128);
          wq = (wq - d0) >> 8;
          const int d1 = (int)(((wq + 64) & 127) - 64);  // 7 bits: stored doubled, multiplied by the 0 / 128 operand
          const int d2 = (int)((wq - d1) >> 7);
          word[gq][0] |= (uint32_t)(d0 & 255) << (8 * i);
          word[gq][1] |= (uint32_t)((2 * d1) & 255) << (8 * i);
          word[gq][2] |= (uint32_t)(d2 & 255) << (8 * i);
        }
      }
    }
#pragma unroll
    for (int gq = 0; gq < 3; ++gq)
#pragma unroll
      for (int dg = 0; dg < 3; ++dg) {
        uint32_t* dst = reinterpret_cast<uint32_t*>(Bq + ((size_t)((gq * 3 + dg) * KK + kk) * HB8 + nb) * 32 + ln);
        dst[reg] = word[gq][dg];
      }
  }
  return true;
}

// theta of 16 rows x NBH hidden blocks for one warp: th[i][c] follows the m16n8 accumulator layout
// (c = 0, 1: row g, units 2t, 2t + 1 of block nb0 + i; c = 2, 3: row g + 8).  Vrows = the 16 rows of the V tile.
template <int Npad, int HB8, int NBH>
__device__ __forceinline__ bool theta_imma_tile(const double* __restrict__ Vrows, int LDV, const uint2* __restrict__ Bq, int nb0,
                                                const double* __restrict__ hb, double inv_scale, double (&th)[NBH][4]) {
  constexpr int KK = (Npad + 31) / 32;
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  uint32_t a[KK][4];
  bool nonbinary = false;
#pragma unroll
  for (int kk = 0; kk < KK; ++kk)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const double* src = Vrows + (g + 8 * (r & 1)) * LDV + 32 * kk + 16 * (r >> 1) + 4 * t;
      uint32_t w = 0u;
      if (32 * kk + 16 * (r >> 1) + 4 * t < Npad) {  // Npad is a multiple of 8 and 4 t + 3 < 8 k': all four or none
        const double2 x01 = *reinterpret_cast<const double2*>(src), x23 = *reinterpret_cast<const double2*>(src + 2);
        w = (x01.x != 0.0 ? 1u : 0u) | (x01.y != 0.0 ? 0x100u : 0u) | (x23.x != 0.0 ? 0x10000u : 0u) | (x23.y != 0.0 ? 0x1000000u : 0u);
        nonbinary |= ((x01.x != 0.0) & (x01.x != 1.0)) | ((x01.y != 0.0) & (x01.y != 1.0)) | ((x23.x != 0.0) & (x23.x != 1.0)) |
                     ((x23.y != 0.0) & (x23.y != 1.0));
      }
      a[kk][r] = w;
    }
#pragma unroll
  for (int gq = 0; gq < 3; ++gq) {
    int acc[NBH][4];
#pragma unroll
    for (int i = 0; i < NBH; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0;
#pragma unroll
    for (int dg = 2; dg >= 0; --dg) {
#pragma unroll
      for (int kk = 0; kk < KK; ++kk) {
        uint32_t av[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) av[r] = dg == 1 ? a[kk][r] << 7 : a[kk][r];
#pragma unroll
        for (int i = 0; i < NBH; ++i) {
          if (nb0 + i < HB8) {
            const uint2 b = Bq[((size_t)((gq * 3 + dg) * KK + kk) * HB8 + nb0 + i) * 32 + lane];
            imma16832_u8s8(acc[i], av, b.x, b.y);
          }
        }
      }
      if (dg == 2) {
#pragma unroll
        for (int i = 0; i < NBH; ++i)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[i][c] = (int)((uint32_t)acc[i][c] << 15);
      }
    }
#pragma unroll
    for (int i = 0; i < NBH; ++i)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const double td = int_to_double(acc[i][c]);
        th[i][c] = gq == 0 ? td : fma(th[i][c], 2.384185791015625e-07, td);  // 2^-22
      }
  }
#pragma unroll
  for (int i = 0; i < NBH; ++i)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int unit = 8 * (nb0 + i) + 2 * t + (c & 1);
      th[i][c] = fma(th[i][c], inv_scale, nb0 + i < HB8 ? hb[unit] : 0.0);
    }
  return nonbinary;
}

// NB8: 8-column blocks of the visible layer, HB8: 8-unit blocks of all hidden-like features (computed + aux)
// IMMA: theta from the integer images above; flag: see there (the FP64 instantiation returns at once when *flag == 0)
template <int NB8, int HB8, bool IMMA>
__global__ void __launch_bounds__(GD_THREADS, QCB_GD_MIN_CTAS) grad_reduce_dmma_kernel(AffineSpec s, RowSrc src, int mode, int64_t rows,
                                                                      double* __restrict__ partials, int* __restrict__ flag) {
  constexpr int Npad = 8 * NB8, Hpad = 8 * HB8;
  constexpr int KN4 = 2 * NB8;            // k-steps of phase 1
  constexpr int KK32 = (Npad + 31) / 32;  // k-steps of the integer phase 1
  constexpr int LDV = Npad + 4, LDP = Hpad + 4;
  static_assert(!IMMA || GD_WARPS == 8, "integer phase 1 splits a 64-row tile as 4 row groups x 2 column halves");
  if (!IMMA && flag != nullptr && *flag == 0) return;  // the integer kernel ahead of this one did the batch
  constexpr int MB = (HB8 + GD_WARPS - 1) / GD_WARPS;  // G row blocks per warp
  // Narrow hidden layers (HB8 < 8) would leave warps idle in phase 2: KG groups of HB8 warps then split the 64 rows
  // of a tile between them and their partial sums are combined in shared memory at the end.
  constexpr int KG = HB8 < GD_WARPS ? ((GD_WARPS / HB8) >= 4 ? 4 : (GD_WARPS / HB8 >= 2 ? 2 : 1)) : 1;
  constexpr int VPW = KG > 1 ? (NB8 + HB8 - 1) / HB8 : (NB8 + GD_WARPS - 1) / GD_WARPS;  // bias-sum columns per warp
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* Wf = reinterpret_cast<double*>(smem_raw);     // [KN4][HB8][32] B fragments of phase 1 (FP64) ...
  uint2* Bq = reinterpret_cast<uint2*>(smem_raw);       // ... or [3][3][KK32][HB8][32] digit images (integer)
  double* hb = Wf + (IMMA ? (size_t)9 * KK32 * HB8 * 32 : (size_t)KN4 * HB8 * 32);  // [Hpad]
  double* Vt = hb + Hpad;                               // [GD_TR][LDV]
  double* Pt = Vt + (size_t)GD_TR * LDV;                // [GD_TR][LDP]
  double* wts = Pt + (size_t)GD_TR * LDP;               // [GD_TR]

  const int N = s.n_in1, Hc = s.n_out;                  // Hc units are computed here, the rest are aux features
  const int n_aux = mode == ROWS_EXPANDED ? src.n_aux : 0;
  const int H = Hc + n_aux;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;

  __shared__ int8_t spos[128];
  __shared__ double e2tab[EXP2_TABLE];
  double inv_scale = 1.0;
  if (IMMA) {
    int sc = 0;
    if (!stage_theta_digits<KK32, HB8>(s, N, Hc, Bq, e2tab, &sc)) {  // e2tab doubles as the reduction scratch
      if (threadIdx.x == 0) *flag = 1;
      return;
    }
    inv_scale = scalbn(1.0, -sc);
    __syncthreads();
  } else {
    // stage W fragments: element (kk, nb, lane) = Wst[hid = 8 nb + g][vis = 4 kk + t]
    for (int idx = threadIdx.x; idx < KN4 * HB8 * 32; idx += GD_THREADS) {
      const int l = idx & 31, nb = (idx >> 5) % HB8, kk = (idx >> 5) / HB8;
      const int hid = 8 * nb + (l >> 2), vis = 4 * kk + (l & 3);
      Wf[idx] = (hid < Hc && vis < N) ? s.mop(hid, vis) : 0.0;
    }
  }
  for (int i = threadIdx.x; i < Hpad; i += GD_THREADS) hb[i] = i < Hc ? s.bias(i) : 0.0;
  fill_exp2_table(e2tab);
  if (threadIdx.x < 128) spos[threadIdx.x] = mode == ROWS_EXPANDED ? src.ex.sitepos[threadIdx.x] : (int8_t)-1;
  for (int i = threadIdx.x; i < GD_TR * LDV; i += GD_THREADS) Vt[i] = 0.0;
  for (int i = threadIdx.x; i < GD_TR * LDP; i += GD_THREADS) Pt[i] = 0.0;
  // explicit rows with 16-byte aligned row starts are pulled in by the TMA engine, one bulk copy per row into the
  // padded tile (no per-element instructions); everything else is filled element by element
  __shared__ __align__(8) uint64_t fill_bar;
  // (rows of at least 384 bytes: below that the per-copy overhead exceeds the instructions it saves)
  const bool bulk_fill = mode == ROWS_EXPLICIT && N >= 48 && (N & 1) == 0 && (reinterpret_cast<uintptr_t>(src.v) & 15) == 0;
  uint32_t fill_parity = 0;
  if (threadIdx.x == 0) mbar_init(&fill_bar, 1);
  __syncthreads();

  double G[MB][NB8][2];
  double hsum[MB], vsum[VPW];
  const int kgroup = KG > 1 ? warp / HB8 : 0, wing = KG > 1 ? warp % HB8 : warp;  // k-group and position inside it
  const bool p2_active = KG > 1 ? kgroup < KG : true;
  const int vstride = KG > 1 ? HB8 : GD_WARPS;
#pragma unroll
  for (int m = 0; m < MB; ++m) {
    hsum[m] = 0.0;
#pragma unroll
    for (int nb = 0; nb < NB8; ++nb) { G[m][nb][0] = 0.0; G[m][nb][1] = 0.0; }
  }
#pragma unroll
  for (int i = 0; i < VPW; ++i) vsum[i] = 0.0;

  const int fill_r0 = threadIdx.x / N, fill_j0 = threadIdx.x - fill_r0 * N;
  const int fill_dr = GD_THREADS / N, fill_dj = GD_THREADS - fill_dr * N;
  const int64_t ntiles = (rows + GD_TR - 1) / GD_TR;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t r0 = tile * GD_TR;
    // ---- fill the V tile, the row weights and the aux columns of the P tile
    if (mode == ROWS_EXPLICIT) {  // pull the CTA's next tile (one contiguous span) into L2 while this one is processed
      const int64_t rn = r0 + (int64_t)gridDim.x * GD_TR;
      const int64_t nbytes = ((rows - rn < GD_TR ? rows - rn : (int64_t)GD_TR)) * N * (int64_t)sizeof(double);
      const char* nbase = reinterpret_cast<const char*>(src.v + rn * N);
      for (int64_t off = (int64_t)threadIdx.x * 128; off < nbytes; off += (int64_t)GD_THREADS * 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(nbase + off));
    }
    if (bulk_fill) {
      const int nrows = (int)(rows - r0 < GD_TR ? rows - r0 : (int64_t)GD_TR);
      if (threadIdx.x < GD_TR) {
        if (threadIdx.x < nrows) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic accesses of the row vs. the async write
          tma_bulk_g2s(Vt + threadIdx.x * LDV, src.v + (r0 + threadIdx.x) * N, (uint32_t)(N * sizeof(double)), &fill_bar);
        } else {
          for (int jj = 0; jj < N; ++jj) Vt[threadIdx.x * LDV + jj] = 0.0;  // ragged last tile
        }
      }
      if (threadIdx.x == 0) mbar_expect_tx(&fill_bar, (uint32_t)(nrows * N * sizeof(double)));
    } else {
      // element idx = rr * N + j of the tile, stepped without divisions; the rotated-site table comes from shared memory
      // batches of 8 elements per thread: the 8 loads of a batch are issued before its first store
      constexpr int FILL = (GD_TR * Npad + GD_THREADS - 1) / GD_THREADS, FB = FILL < 8 ? FILL : 8;
      int rr = fill_r0, j = fill_j0;
      for (int i0 = 0; i0 < FILL; i0 += FB) {
        double vals[FB];
        int rr2 = rr, j2 = j;
#pragma unroll
        for (int i = 0; i < FB; ++i) {
          const int64_t r = r0 + rr2;
          double val = 0.0;
          if (threadIdx.x + (i0 + i) * GD_THREADS < GD_TR * N && r < rows) {
            if (mode == ROWS_EXPANDED) {
              const int tpos = (int)spos[j2];
              val = tpos < 0 ? src.ex.samples[(r >> src.ex.S) * N + j2]
                             : ((((uint32_t)r & ((1u << src.ex.S) - 1u)) >> (src.ex.S - 1 - tpos)) & 1u ? 1.0 : 0.0);
            } else {
              val = src.v[r * N + j2];
            }
          }
          vals[i] = val;
          rr2 += fill_dr;
          j2 += fill_dj;
          if (j2 >= N) { j2 -= N; ++rr2; }
        }
#pragma unroll
        for (int i = 0; i < FB; ++i) {
          if (threadIdx.x + (i0 + i) * GD_THREADS < GD_TR * N) Vt[rr * LDV + j] = vals[i];
          rr += fill_dr;
          j += fill_dj;
          if (j >= N) { j -= N; ++rr; }
        }
      }
    }
    for (int rr = threadIdx.x; rr < GD_TR; rr += GD_THREADS) {
      const int64_t r = r0 + rr;
      wts[rr] = r < rows ? (mode == ROWS_EXPANDED ? src.wrow[r] : 1.0) : 0.0;
    }
    for (int idx = threadIdx.x; idx < GD_TR * n_aux; idx += GD_THREADS) {
      const int rr = idx / n_aux, o = idx - rr * n_aux;
      const int64_t r = r0 + rr;
      Pt[rr * LDP + Hc + o] = r < rows ? src.aux[r * n_aux + o] : 0.0;
    }
    if (bulk_fill) {
      mbar_wait(&fill_bar, fill_parity);
      fill_parity ^= 1u;
    }
    __syncthreads();

    // ---- phase 1: theta for rows 8 warp .. 8 warp + 7, p = w * sigmoid(theta) -> P tile
    bool nonbinary = false;  // integer phase 1: an entry other than 0 and 1 was seen
    if (IMMA) {  // integer images: warp = (row group of 16, half of the hidden blocks)
      constexpr int NBH = (HB8 + 1) / 2;
      const int rg = warp >> 1, nb0 = (warp & 1) * NBH;
      double th[NBH][4];
      nonbinary = theta_imma_tile<Npad, HB8, NBH>(Vt + (size_t)16 * rg * LDV, LDV, Bq, nb0, hb, inv_scale, th);
      const double w0 = wts[16 * rg + g], w1 = wts[16 * rg + g + 8];
      double* prow0 = Pt + (16 * rg + g) * LDP;
      double* prow1 = prow0 + 8 * LDP;
#pragma unroll
      for (int i = 0; i < NBH; ++i) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int unit = 8 * (nb0 + i) + 2 * t + (c & 1);
          if (nb0 + i < HB8 && unit < Hc) {
            const double p = sigmoid_tab(th[i][c], e2tab);  // in [0, 1] by construction: the reference's clamp is a no-op
            if (c < 2) prow0[unit] = w0 * p; else prow1[unit] = w1 * p;
          }
        }
      }
    } else {
      double acc[HB8][2];
#pragma unroll
      for (int nb = 0; nb < HB8; ++nb) {
        acc[nb][0] = hb[8 * nb + 2 * t];
        acc[nb][1] = hb[8 * nb + 2 * t + 1];
      }
      const double* arow = Vt + (8 * warp + g) * LDV + t;
#pragma unroll 4
      for (int kk = 0; kk < KN4; ++kk) {
        const double a = arow[4 * kk];
        const double* bf = Wf + (size_t)kk * HB8 * 32 + lane;
#pragma unroll
        for (int nb = 0; nb < HB8; ++nb) dmma884(acc[nb][0], acc[nb][1], a, bf[nb * 32]);
      }
      const double wr = wts[8 * warp + g];
      double* prow = Pt + (8 * warp + g) * LDP;
#pragma unroll
      for (int nb = 0; nb < HB8; ++nb) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int unit = 8 * nb + 2 * t + e;
          if (unit < Hc) prow[unit] = wr * sigmoid_tab(acc[nb][e], e2tab);  // in [0, 1] by construction
        }
      }
    }
    if (IMMA) {
      if (threadIdx.x == 0) nonbinary |= *reinterpret_cast<volatile int*>(flag) != 0;  // another CTA gave up already
      if (__syncthreads_or(nonbinary)) {
        if (threadIdx.x == 0) *flag = 1;
        return;
      }
    } else {
      __syncthreads();
    }

    // ---- phase 2: G += P^T V for this warp's row blocks of G; column sums for the bias gradients
    constexpr int KPG = (GD_TR / 4) / KG;  // k-steps per group
#pragma unroll 2
    for (int kq = 0; kq < KPG; ++kq) {
      if (!p2_active) break;
      const int kk = kgroup * KPG + kq;
      const double* vrow = Vt + (4 * kk + t) * LDV + g;
      const double* prow = Pt + (4 * kk + t) * LDP + g;
      double b[NB8];
#pragma unroll
      for (int nb = 0; nb < NB8; ++nb) b[nb] = vrow[8 * nb];
      const double wk = wts[4 * kk + t];
#pragma unroll
      for (int i = 0; i < VPW; ++i) {
        const int nb = wing + vstride * i;
        if (nb < NB8) vsum[i] = fma(wk, vrow[8 * nb], vsum[i]);
      }
#pragma unroll
      for (int m = 0; m < MB; ++m) {
        const int mb = wing + GD_WARPS * m;
        if (mb < HB8) {
          const double a = prow[8 * mb];
          hsum[m] += a;
#pragma unroll
          for (int nb = 0; nb < NB8; ++nb) dmma884(G[m][nb][0], G[m][nb][1], a, b[nb]);
        }
      }
    }
    __syncthreads();
  }

  // ---- write this CTA's partial in flat order [Wst (H x N) | vb (N) | hb (H)]
  const size_t Gflat = (size_t)H * N + N + H;
  double* cta_out = partials + (size_t)blockIdx.x * Gflat;
  double* out = cta_out;
  if (KG > 1) {  // each k-group first writes its sums to shared memory (the tiles are no longer needed)
    __syncthreads();
    out = Vt + (size_t)kgroup * Gflat;
  }
#pragma unroll
  for (int m = 0; m < MB; ++m) {
    const int mb = wing + GD_WARPS * m;
    if (p2_active && mb < HB8) {
      const int hid = 8 * mb + g;
#pragma unroll
      for (int nb = 0; nb < NB8; ++nb)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int vis = 8 * nb + 2 * t + e;
          if (hid < H && vis < N) out[(size_t)hid * N + vis] = G[m][nb][e];
        }
      double hs = hsum[m];
      hs += __shfl_xor_sync(0xffffffffu, hs, 1);
      hs += __shfl_xor_sync(0xffffffffu, hs, 2);
      if (t == 0 && hid < H) out[(size_t)H * N + N + hid] = hs;
    }
  }
#pragma unroll
  for (int i = 0; i < VPW; ++i) {
    const int nb = wing + vstride * i;
    double vs = vsum[i];
    vs += __shfl_xor_sync(0xffffffffu, vs, 1);
    vs += __shfl_xor_sync(0xffffffffu, vs, 2);
    const int vis = 8 * nb + g;
    if (p2_active && nb < NB8 && t == 0 && vis < N) out[(size_t)H * N + vis] = vs;
  }
  if (KG > 1) {  // groups summed in a fixed order
    __syncthreads();
    for (size_t idx = threadIdx.x; idx < Gflat; idx += GD_THREADS) {
      double tsum = Vt[idx];
      for (int kg = 1; kg < KG; ++kg) tsum += Vt[(size_t)kg * Gflat + idx];
      cta_out[idx] = tsum;
    }
  }
}

static int pick8(int units) {
  if (units <= 16) return 2;
  if (units <= 24) return 3;
  if (units <= 32) return 4;
  if (units <= 64) return 8;
  if (units <= 128) return 16;
  return 0;
}

static size_t gd_smem(int NB8, int HB8, bool imma = false) {
  const size_t Npad = 8 * NB8, Hpad = 8 * HB8;
  const size_t wfrag = imma ? (size_t)9 * ((Npad + 31) / 32) * HB8 * 32 : (size_t)2 * NB8 * HB8 * 32;
  return (wfrag + Hpad + (size_t)GD_TR * (Npad + 4) + (size_t)GD_TR * (Hpad + 4) + GD_TR) * sizeof(double);
}

// resident CTAs per SM of one instantiation (cached per device)
template <int NB8, int HB8, bool IMMA>
static int gd_residency(int* out) {
  static int ctas_on[QCB_MAX_DEVICES] = {0};
  int& ctas_per_sm = ctas_on[device_slot()];
  if (ctas_per_sm == 0) {
    const size_t smem = gd_smem(NB8, HB8, IMMA);
    QCB_CUDA(cudaFuncSetAttribute(grad_reduce_dmma_kernel<NB8, HB8, IMMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    QCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, grad_reduce_dmma_kernel<NB8, HB8, IMMA>, GD_THREADS, smem));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
  }
  *out = ctas_per_sm;
  return QCB_OK;
}

// persistent grid = exactly the resident CTAs.  With a flag the integer-phase-1 kernel goes first and the FP64 kernel
// behind it only works when the first one gave up (non-binary rows, weights outside the image); both use one grid.
int try_grad_reduce_umma(const AffineSpec& s, const double* v, const unsigned char* bits, int H, int64_t rows, double* partials, int nslots,
                         int* flag, unsigned char* img, int* launched, cudaStream_t st);
bool grad_umma_shape(int N, int H, int64_t rows);
int run_unpack_states_if(const int* flag, const uint8_t* bits, int64_t rows, int N, double* v, cudaStream_t st);
int g_grad_mode = 0;

// Packed rows are only taken by the tcgen05 kernel (the FP64 kernel behind it reads the float64 scratch rows, which are
// unpacked on the device only when that kernel is needed): true when launch_gd below will route such a batch there.
bool grad_reduce_bits_supported(int N, int H, int64_t rows) {
  return g_grad_mode == 0 && grad_umma_shape(N, H, rows) && rows <= ((int64_t)1 << 29) && gd_smem(pick8(N), pick8(H)) <= SMEM_LIMIT;
}  // qcb_set_grad_reduce_mode: 0 automatic, 1 FP64 tensor cores only, 3 never the tcgen05 kernel

template <int NB8, int HB8, bool WITH_IMMA>
static int launch_gd(const AffineSpec& s, const RowSrc& src, int mode, int64_t rows, double* partials, int* grid_io, int* flag,
                     unsigned char* umma_img, cudaStream_t st) {
  int per_sm = 0;
  if (int rc = gd_residency<NB8, HB8, false>(&per_sm)) return rc;
  if (WITH_IMMA && flag != nullptr) {
    int per_sm_i = 0;
    if (int rc = gd_residency<NB8, HB8, WITH_IMMA>(&per_sm_i)) return rc;
    if (per_sm_i < per_sm) per_sm = per_sm_i;
  }
  const int64_t resident = (int64_t)sm_count() * per_sm;
  if (*grid_io > resident) *grid_io = (int)resident;
  const int grid = *grid_io;
  if (WITH_IMMA && flag != nullptr) {
    QCB_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
    // binary rows: both contractions on tcgen05 / TMEM (grad_reduce_umma.cu) where the shape fits, else integer theta + DMMA
    int on_umma = 0;
    if (mode == ROWS_EXPLICIT && g_grad_mode == 0) {
      if (int rc = try_grad_reduce_umma(s, src.v, src.bits, s.n_out, rows, partials, grid, flag, umma_img, &on_umma, st)) return rc;
      // packed source: the FP64 kernel queued below reads float64 rows, written from the bits only if *flag was raised
      if (on_umma && src.bits != nullptr)
        if (int rc = run_unpack_states_if(flag, src.bits, rows, s.n_in1, const_cast<double*>(src.v), st)) return rc;
    }
    if (src.bits != nullptr && !on_umma) return fail(QCB_ERR_UNSUPPORTED, "grad_reduce: packed rows need the tcgen05 kernel");
    if (!on_umma && g_grad_mode != 1) {
      grad_reduce_dmma_kernel<NB8, HB8, WITH_IMMA><<<grid, GD_THREADS, gd_smem(NB8, HB8, true), st>>>(s, src, mode, rows, partials, flag);
      QCB_LAUNCH_CHECK();
    } else if (!on_umma) {
      flag = nullptr;  // FP64 only
    }
  } else {
    flag = nullptr;
  }
  grad_reduce_dmma_kernel<NB8, HB8, false><<<grid, GD_THREADS, gd_smem(NB8, HB8), st>>>(s, src, mode, rows, partials, flag);
  QCB_LAUNCH_CHECK();
  return QCB_OK;
}

// Returns QCB_OK and sets *grid_out (> 0) when the tensor-core path took the batch; *grid_out = 0 means
// "not applicable, use the FMA kernel".  H = computed + aux hidden-like units.
// flag: one int of device scratch (or NULL): enables the integer phase 1 for 25..64 x 25..64 layers and >= 16384 rows.
int try_grad_reduce_dmma(const AffineSpec& s, const RowSrc& src, int mode, int H, int64_t rows, double* partials, int max_grid,
                         int* grid_out, int* flag, unsigned char* umma_img, cudaStream_t st) {
  *grid_out = 0;
  if (mode == ROWS_ENUM || rows < 2048) return QCB_OK;
  const int NB8 = pick8(s.n_in1), HB8 = pick8(H);
  if (NB8 == 0 || HB8 == 0 || gd_smem(NB8, HB8) > SMEM_LIMIT) return QCB_OK;
  const int64_t ntiles = (rows + GD_TR - 1) / GD_TR;
  int grid = (int)(ntiles < max_grid ? ntiles : max_grid);
  if (rows < 16384 || H != s.n_out) flag = nullptr;  // small batches: not worth a second launch; aux features: FP64 only
  int rc = QCB_OK;
#define QCB_GD_CASE(NB, HBx) \
  if (NB8 == NB && HB8 == HBx) rc = launch_gd<NB, HBx, (NB >= 4 && NB <= 8 && HBx >= 4 && HBx <= 8 && GD_WARPS == 8)>(s, src, mode, rows, partials, &grid, flag, umma_img, st); else
  QCB_GD_CASE(2, 2) QCB_GD_CASE(2, 3) QCB_GD_CASE(2, 4) QCB_GD_CASE(2, 8) QCB_GD_CASE(2, 16)
  QCB_GD_CASE(3, 2) QCB_GD_CASE(3, 3) QCB_GD_CASE(3, 4) QCB_GD_CASE(3, 8) QCB_GD_CASE(3, 16)
  QCB_GD_CASE(4, 2) QCB_GD_CASE(4, 3) QCB_GD_CASE(4, 4) QCB_GD_CASE(4, 8) QCB_GD_CASE(4, 16)
  QCB_GD_CASE(8, 2) QCB_GD_CASE(8, 3) QCB_GD_CASE(8, 4) QCB_GD_CASE(8, 8) QCB_GD_CASE(8, 16)
  QCB_GD_CASE(16, 2) QCB_GD_CASE(16, 3) QCB_GD_CASE(16, 4) QCB_GD_CASE(16, 8)
  return QCB_OK;  // (16, 16) does not fit: FMA kernel
#undef QCB_GD_CASE
  if (rc == QCB_OK) *grid_out = grid;
  return rc;
}

// ---- effective energy on the tensor cores (large batches) ---------------------------------------
// E[r] = -(b . v_r + sum_i softplus(theta_i)), theta = Wst v_r + hb as DMMA tiles (phase 1 above).
// The softplus sum is taken in product form per lane, exactly reproducing torch's threshold:
//   sum_i softplus(theta_i) = sum_{theta_i > 20} theta_i + log prod_{theta_i <= 20} (1 + e^{theta_i});
// a lane multiplies at most 32 factors <= 1 + e^20 before taking the log, so nothing can overflow.
template <int NB8, int HB8>
__global__ void __launch_bounds__(GD_THREADS) rows_energy_dmma_kernel(AffineSpec s, const double* __restrict__ v,
                                                                      const double* __restrict__ vb, int64_t rows,
                                                                      double* __restrict__ E) {
  constexpr int Npad = 8 * NB8, Hpad = 8 * HB8;
  constexpr int KN4 = 2 * NB8;
  constexpr int LDV = Npad + 4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* Wf = reinterpret_cast<double*>(smem_raw);     // [KN4][HB8][32]
  double* hb = Wf + (size_t)KN4 * HB8 * 32;             // [Hpad]
  double* vbs = hb + Hpad;                              // [Npad]
  double* Vt = vbs + Npad;                              // [GD_TR][LDV]
  const int N = s.n_in1, H = s.n_out;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  for (int idx = threadIdx.x; idx < KN4 * HB8 * 32; idx += GD_THREADS) {
    const int l = idx & 31, nb = (idx >> 5) % HB8, kk = (idx >> 5) / HB8;
    const int hid = 8 * nb + (l >> 2), vis = 4 * kk + (l & 3);
    Wf[idx] = (hid < H && vis < N) ? s.mop(hid, vis) : 0.0;
  }
  // padded hidden units get a bias of -inf-like size so that they contribute log(1 + 0) = 0
  for (int i = threadIdx.x; i < Hpad; i += GD_THREADS) hb[i] = i < H ? s.bias(i) : -1.0e300;
  for (int i = threadIdx.x; i < Npad; i += GD_THREADS) vbs[i] = i < N ? vb[i] : 0.0;
  for (int i = threadIdx.x; i < GD_TR * LDV; i += GD_THREADS) Vt[i] = 0.0;
  __shared__ double e2tab[EXP2_TABLE];
  fill_exp2_table(e2tab);
  __shared__ __align__(8) uint64_t fill_bar;  // wide, aligned rows come in through TMA row copies as in the gradient kernel
  const bool bulk_fill = N >= 48 && (N & 1) == 0 && (reinterpret_cast<uintptr_t>(v) & 15) == 0;
  uint32_t fill_parity = 0;
  if (threadIdx.x == 0) mbar_init(&fill_bar, 1);
  __syncthreads();
  const int fill_r0 = threadIdx.x / N, fill_j0 = threadIdx.x - fill_r0 * N;
  const int fill_dr = GD_THREADS / N, fill_dj = GD_THREADS - fill_dr * N;
  const int64_t ntiles = (rows + GD_TR - 1) / GD_TR;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t r0 = tile * GD_TR;
    if (bulk_fill) {
      const int nrows = (int)(rows - r0 < GD_TR ? rows - r0 : (int64_t)GD_TR);
      if (threadIdx.x < GD_TR) {
        if (threadIdx.x < nrows) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          tma_bulk_g2s(Vt + threadIdx.x * LDV, v + (r0 + threadIdx.x) * N, (uint32_t)(N * sizeof(double)), &fill_bar);
        } else {
          for (int jj = 0; jj < N; ++jj) Vt[threadIdx.x * LDV + jj] = 0.0;  // ragged last tile
        }
      }
      if (threadIdx.x == 0) mbar_expect_tx(&fill_bar, (uint32_t)(nrows * N * sizeof(double)));
      mbar_wait(&fill_bar, fill_parity);
      fill_parity ^= 1u;
    } else {
      // batches of 8 elements per thread, stepped without divisions: the loads of a batch are issued before its stores
      constexpr int FILL = (GD_TR * Npad + GD_THREADS - 1) / GD_THREADS, FB = FILL < 8 ? FILL : 8;
      int rr = fill_r0, j = fill_j0;
      for (int i0 = 0; i0 < FILL; i0 += FB) {
        double vals[FB];
        int rr2 = rr, j2 = j;
#pragma unroll
        for (int i = 0; i < FB; ++i) {
          const bool live = threadIdx.x + (i0 + i) * GD_THREADS < GD_TR * N && r0 + rr2 < rows;
          vals[i] = live ? v[(r0 + rr2) * N + j2] : 0.0;
          rr2 += fill_dr;
          j2 += fill_dj;
          if (j2 >= N) { j2 -= N; ++rr2; }
        }
#pragma unroll
        for (int i = 0; i < FB; ++i) {
          if (threadIdx.x + (i0 + i) * GD_THREADS < GD_TR * N) Vt[rr * LDV + j] = vals[i];
          rr += fill_dr;
          j += fill_dj;
          if (j >= N) { j -= N; ++rr; }
        }
      }
    }
    __syncthreads();
    double acc[HB8][2];
#pragma unroll
    for (int nb = 0; nb < HB8; ++nb) {
      acc[nb][0] = hb[8 * nb + 2 * t];
      acc[nb][1] = hb[8 * nb + 2 * t + 1];
    }
    const double* arow = Vt + (8 * warp + g) * LDV + t;
    double bv = 0.0;
#pragma unroll 4
    for (int kk = 0; kk < KN4; ++kk) {
      const double a = arow[4 * kk];
      bv = fma(vbs[4 * kk + t], a, bv);
      const double* bf = Wf + (size_t)kk * HB8 * 32 + lane;
#pragma unroll
      for (int nb = 0; nb < HB8; ++nb) dmma884(acc[nb][0], acc[nb][1], a, bf[nb * 32]);
    }
    double lin = bv, prod = 1.0;
#pragma unroll
    for (int nb = 0; nb < HB8; ++nb) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const double th = acc[nb][e];
        if (!(th <= 20.0)) lin += th; else prod *= 1.0 + exp_tab(fmax(th, -700.0), e2tab);  // a NaN stays a NaN
      }
    }
    double val = lin + log(prod);
    val += __shfl_xor_sync(0xffffffffu, val, 1);
    val += __shfl_xor_sync(0xffffffffu, val, 2);
    const int64_t r = r0 + 8 * warp + g;
    if (t == 0 && r < rows) E[r] = -val;
    __syncthreads();
  }
}

template <int NB8, int HB8>
static int launch_energy_dmma(const AffineSpec& s, const double* v, const double* vb, int64_t rows, double* E, int grid,
                              cudaStream_t st) {
  const size_t smem = ((size_t)2 * NB8 * HB8 * 32 + 8 * HB8 + 8 * NB8 + (size_t)GD_TR * (8 * NB8 + 4)) * sizeof(double);
  static int ctas_on[QCB_MAX_DEVICES] = {0};  // persistent grid = exactly the resident CTAs (a partial second wave costs more than it gives)
  int& ctas_per_sm = ctas_on[device_slot()];
  if (ctas_per_sm == 0) {
    QCB_CUDA(cudaFuncSetAttribute(rows_energy_dmma_kernel<NB8, HB8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    QCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, rows_energy_dmma_kernel<NB8, HB8>, GD_THREADS, smem));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
  }
  const int64_t resident = (int64_t)sm_count() * ctas_per_sm;
  if (grid > resident) grid = (int)resident;
  rows_energy_dmma_kernel<NB8, HB8><<<grid, GD_THREADS, smem, st>>>(s, v, vb, rows, E);
  QCB_LAUNCH_CHECK();
  return QCB_OK;
}

// *handled = 1 when the tensor-core path took the batch (rows >= 2048 and the layer fits)
int try_energy_dmma(const AffineSpec& s, const double* v, const double* vb, int64_t rows, double* E, int* handled,
                    cudaStream_t st) {
  *handled = 0;
  if (rows < 2048) return QCB_OK;
  const int NB8 = pick8(s.n_in1), HB8 = pick8(s.n_out);
  if (NB8 == 0 || HB8 == 0 || HB8 > 8) return QCB_OK;  // a lane multiplies 2 HB8 <= 16 factors before its log
  const size_t smem = ((size_t)2 * NB8 * HB8 * 32 + 8 * HB8 + 8 * NB8 + (size_t)GD_TR * (8 * NB8 + 4)) * sizeof(double);
  if (smem > SMEM_LIMIT) return QCB_OK;
  const int64_t ntiles = (rows + GD_TR - 1) / GD_TR;
  const int64_t cap = (int64_t)sm_count() * 8;
  const int grid = (int)(ntiles < cap ? ntiles : cap);
  int rc = QCB_OK;
#define QCB_EN_CASE(NB, HBx) \
  if (NB8 == NB && HB8 == HBx) rc = launch_energy_dmma<NB, HBx>(s, v, vb, rows, E, grid, st); else
  QCB_EN_CASE(2, 2) QCB_EN_CASE(2, 3) QCB_EN_CASE(2, 4) QCB_EN_CASE(2, 8)
  QCB_EN_CASE(3, 2) QCB_EN_CASE(3, 3) QCB_EN_CASE(3, 4) QCB_EN_CASE(3, 8)
  QCB_EN_CASE(4, 2) QCB_EN_CASE(4, 3) QCB_EN_CASE(4, 4) QCB_EN_CASE(4, 8)
  QCB_EN_CASE(8, 2) QCB_EN_CASE(8, 3) QCB_EN_CASE(8, 4) QCB_EN_CASE(8, 8)
  QCB_EN_CASE(16, 2) QCB_EN_CASE(16, 3) QCB_EN_CASE(16, 4) QCB_EN_CASE(16, 8)
  return QCB_OK;
#undef QCB_EN_CASE
  if (rc == QCB_OK) *handled = 1;
  return rc;
}

}  // namespace qcb
