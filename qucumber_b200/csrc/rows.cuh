// Row-wise affine layer staging shared by the energy / conditional / gradient kernels.
#pragma once
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
  // stacked_units = 1: units [0, m1_rows) are M1 rows, the rest M2 rows; one input x
  // stacked_units = 0: every unit sees M1^T over x and M2^T over x2 ("stacked inputs")
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

#ifdef __CUDACC__
__device__ __forceinline__ void stage_affine(const AffineSpec& s, double* Ws, int LD, double* bs) {
  const int nin = s.n_in();
  for (int idx = threadIdx.x; idx < s.n_out * nin; idx += blockDim.x) {
    const int o = idx / nin, i = idx - o * nin;
    Ws[o * LD + i] = s.mop(o, i);
  }
  for (int o = threadIdx.x; o < s.n_out; o += blockDim.x) bs[o] = s.bias(o);
}
#endif

// all hidden-like units of a (Purification)RBM: [W; U] rows with biases [c; d]
inline AffineSpec stacked_hidden_spec(const StackedParams& p) {
  AffineSpec s{};
  s.M1 = p.W; s.m1_rows = p.nh; s.m1_cols = p.N; s.t1 = 0;
  s.M2 = p.U; s.m2_rows = p.na; s.m2_cols = p.N; s.t2 = 0;
  s.bias1 = p.c; s.nb1 = p.nh; s.bias2 = p.d;
  s.n_out = p.nhid(); s.n_in1 = p.N; s.n_in2 = 0;
  s.stacked_units = 1;
  return s;
}

// only the hidden units (W, c) of a (Purification)RBM
inline AffineSpec hidden_only_spec(const StackedParams& p) {
  AffineSpec s{};
  s.M1 = p.W; s.m1_rows = p.nh; s.m1_cols = p.N; s.t1 = 0;
  s.M2 = nullptr; s.m2_rows = 0; s.m2_cols = p.N; s.t2 = 0;
  s.bias1 = p.c; s.nb1 = p.nh; s.bias2 = nullptr;
  s.n_out = p.nh; s.n_in1 = p.N; s.n_in2 = 0;
  s.stacked_units = 1;
  return s;
}

constexpr int ROT_MAX_N = 512;  // visible units of a state with rotated-basis data (the site tables travel as kernel arguments)

// Expansion of measured samples over their rotated sites (utils/unitaries.py:154-181): row
// r = b * 2^S + i is sample b with rotated site number t set to bit (i >> (S-1-t)) & 1.
struct Expansion {
  const double* samples;   // (B, N)
  int N, S;
  int8_t sitepos[ROT_MAX_N];  // visible unit -> index among the rotated sites, or -1
  __device__ __forceinline__ double value(int64_t row, int j) const {
    const int64_t b = row >> S;
    const int t = sitepos[j];
    if (t < 0) return samples[b * N + j];
    const uint32_t i = (uint32_t)(row & (((int64_t)1 << S) - 1));
    return ((i >> (S - 1 - t)) & 1u) ? 1.0 : 0.0;
  }
};

// The rotated sites of a measurement basis (utils/unitaries.py:154-181)
struct SiteInfo {
  int S;
  int16_t site[32];   // rotated site t -> visible unit
  uint8_t ucode[32];  // rotated site t -> unitary code
};

// Where the rows of a reduced-gradient pass come from, and how they are weighted.
enum { ROWS_EXPLICIT = 0, ROWS_ENUM = 1, ROWS_EXPANDED = 2 };
struct RowSrc {
  const double* v;       // ROWS_EXPLICIT: (rows, N)
  const uint8_t* bits;   // ROWS_EXPLICIT, optional: the same rows packed (qcb.h "Compact states"); v is then SCRATCH of rows x N
                         // doubles that is only filled (from the bits) when the tcgen05 kernel hands the batch to the FP64 kernel
  uint64_t first;        // ROWS_ENUM: index of row 0; weight = exp(-E - logZ)
  double logZ;
  const double* vb;      // ROWS_ENUM: visible bias
  Expansion ex;          // ROWS_EXPANDED
  const double* wrow;    // ROWS_EXPANDED: per-row weight (multiplies p_h and v)
  const double* aux;     // ROWS_EXPANDED: optional (rows, n_aux) features appended after p_h (already weighted)
  int n_aux;
};


inline size_t grad_reduce_tiny_smem(int N, int Hc, int H, int64_t rows) {
  return ((size_t)round_up(Hc * (N | 1), 2) + round_up(Hc, 2) + (size_t)rows * (N + H + 1)) * sizeof(double);
}

#ifdef __CUDACC__
// Tiny batches (one CTA): the kernel applies what grad_finalize_kernel would — g[e] = (accumulate ? g[e] : 0) - scale * sum[e],
// nothing for e >= zero_from — instead of writing a partial for a second launch to pick up (a rotated-basis batch of the
// two-qubit tutorials is up to nine bases x two reductions: the launches were the cost).  Same value bit for bit.
struct GradDirect {
  double* g;        // NULL: write the CTA's partial
  double scale;
  int accumulate;
  int zero_from;
  __device__ __forceinline__ void put(double* partial_out, size_t e, double sum) const {
    if (g == nullptr) {
      partial_out[e] = sum;
    } else {
      const double val = (long long)e >= (long long)zero_from ? 0.0 : -scale * sum;
      g[e] = accumulate ? g[e] + val : val;
    }
  }
};


// The same reduction for a handful of rows of a small layer (rows <= GR_DIRECT_MAX_ROWS, a few KB of tiles), one CTA, one
// THREAD per row in phase 1: in grad_reduce_kernel a warp walks its rows one after the other and every row waits for its own
// global loads — 16 us per launch for the 20..80 expanded rows of a two-qubit basis group, of which a training epoch has
// hundreds (profiles/two_qubit_launches_r02.txt).  Here the loads of all rows are in flight at once; phase 2 adds the rows in
// ascending order with the same fused multiply-adds, so g is bit for bit what the one-CTA walk of grad_reduce_kernel writes.
template <int SRC>
__device__ __forceinline__ void grad_reduce_tiny_body(const AffineSpec& s, const RowSrc& src, int rows, const GradDirect& direct,
                                                      unsigned char* smem_raw) {
  double* Ws = reinterpret_cast<double*>(smem_raw);
  const int N = s.n_in1, Hc = s.n_out;
  const int H = Hc + (SRC == ROWS_EXPANDED ? src.n_aux : 0);
  const int LD = N | 1;
  double* bs = Ws + round_up(Hc * LD, 2);
  double* Vt = bs + round_up(Hc, 2);        // [rows][N]
  double* Pt = Vt + (size_t)rows * N;       // [rows][H]
  double* wts = Pt + (size_t)rows * H;      // [rows]
  stage_affine(s, Ws, LD, bs);
  __syncthreads();
  for (int r = threadIdx.x; r < rows; r += blockDim.x) {
    double* vrow = Vt + r * N;
    double* prow = Pt + r * H;
    for (int i = 0; i < N; ++i) vrow[i] = SRC == ROWS_EXPANDED ? src.ex.value(r, i) : src.v[(size_t)r * N + i];
    const double wgt = SRC == ROWS_EXPANDED ? src.wrow[r] : 1.0;
    for (int o = 0; o < Hc; ++o) {
      const double* wrow = Ws + (size_t)o * LD;
      double th = bs[o];
      for (int i = 0; i < N; ++i) th = fma(wrow[i], vrow[i], th);
      const double pr = fmin(fmax(sigmoid_exact(th), 0.0), 1.0);
      prow[o] = SRC == ROWS_EXPANDED ? pr * wgt : pr;
    }
    if (SRC == ROWS_EXPANDED) {
      for (int o = 0; o < src.n_aux; ++o) prow[Hc + o] = src.aux[(size_t)r * src.n_aux + o];
      wts[r] = wgt;
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < H * N + N + H; e += blockDim.x) {
    double acc = 0.0;
    if (e < H * N) {
      const int i = e / N, j = e - i * N;
      for (int r = 0; r < rows; ++r) acc = fma(Pt[r * H + i], Vt[r * N + j], acc);
    } else if (e < H * N + N) {
      const int j = e - H * N;
      if (SRC == ROWS_EXPANDED) {
        for (int r = 0; r < rows; ++r) acc = fma(wts[r], Vt[r * N + j], acc);
      } else {
        for (int r = 0; r < rows; ++r) acc += Vt[r * N + j];
      }
    } else {
      const int i = e - H * N - N;
      for (int r = 0; r < rows; ++r) acc += Pt[r * H + i];
    }
    direct.put(nullptr, (size_t)e, acc);
  }
}


// one RBM's share of a two-CTA tiny reduction
struct TinyJob {
  AffineSpec s;
  RowSrc src;
  GradDirect direct;
};
#endif

}  // namespace qcb
