// Rotated-basis amplitudes / probabilities and their gradients (sm_100a).
//
// ComplexWaveFunction (nn_states/complex_wavefunction.py:162-211, utils/unitaries.py:154-233):
//   a sample measured in a basis with S non-Z sites is expanded over the 2^S basis states v_i it
//   overlaps with, T_i = Ut_i psi(v_i), D = sum_i T_i, and the gradients are
//     g_am = sum_b sum_i Re[T_i / D] grad E_am(v_i),   g_ph = sum_b sum_i (-Im[T_i / D]) grad E_ph(v_i)
//   (grad E is real, so the complex einsums of the reference collapse to two real weights per
//   expanded state).  Kernel 1 walks each sample's expansion in Gray-code order (one visible unit
//   flips per step, theta updated with nh adds) and writes T and D; kernel 2 turns them into the
//   two weights; the weighted outer products are reduced by grad_reduce.cu (ROWS_EXPANDED).
//   The reference builds v (2^S, B, N) and Ut with NumPy on the host; nothing leaves the GPU here.
//
// DensityMatrix (nn_states/density_matrix.py:113-330, rbm/purification_rbm.py:352-451,
// utils/unitaries.py:236-283): w_ij = Ut_i conj(Ut_j) rho~(v_i, v_j), P = sum_ij Re w_ij.  Using
// w_ji = conj(w_ij) and S_ji = conj(S_ij) the (2, 2^S, 2^S, B, G) gradient tensors of the
// reference reduce to per-expanded-state quantities
//     r_i = sum_j Re w_ij, s_i = sum_j Im w_ij, R_ik = sum_j Re(w_ij S_ijk), I_ik = sum_j Im(w_ij S_ijk)
// and  g_am = f sum_i [ r_i s_h(v_i) (x) v_i ; R_i (x) v_i ; r_i v_i ; r_i s_h(v_i) ; R_i ],
//      g_ph = -f sum_i [ s_i s_h'(v_i) (x) v_i ; I_i (x) v_i ; s_i v_i ; s_i s_h'(v_i) ; 0 ],  f = -1/(P + 1e-8),
// again a weighted outer-product reduction (S_ijk is the complex sigmoid of utils/cplx.py:319-334).
#include "rows.cuh"

namespace qcb {

int run_grad_reduce_expanded(const StackedParams& p, const Expansion& ex, int64_t rows, const double* wrow,
                             const double* aux, int n_aux, int zero_tail, double scale, double* g, void* ws, size_t ws_bytes,
                             cudaStream_t st);
int run_grad_reduce_expanded_pair(const StackedParams& am, const StackedParams& ph, const Expansion& ex, int64_t rows, const double* w_am,
                                  const double* aux_am, const double* w_ph, const double* aux_ph, int n_aux, int zero_tail_ph,
                                  double scale, double* g_am, double* g_ph, void* ws, size_t ws_bytes, cudaStream_t st);
size_t grad_reduce_workspace_bytes(int N, int nhid, int64_t rows);
size_t cwf_gray_workspace_bytes(int N);
int run_cwf_grad_gray(const StackedParams& am, const StackedParams& ph, const double* samples, int64_t Bc, const SiteInfo& si,
                      const double2* T, const double2* D, double* g_am, double* g_ph, void* ws, size_t ws_bytes, int* handled,
                      cudaStream_t st);
int g_cwf_mode = 0;  // 0: automatic (Gray-walk gradient where it applies), 1: always the expanded-row reduction (test / measurement hook)

constexpr int64_t ROT_MAX_ROWS = (int64_t)1 << 21;  // expanded rows processed per chunk
constexpr int ROT_MAX_S = 21;
constexpr int CWF_THREADS = 128;

__device__ __forceinline__ void cmul(double& ar, double& ai, double br, double bi) {
  const double r = ar * br - ai * bi;
  ai = ar * bi + ai * br;
  ar = r;
}

// Ut for expansion index i of a sample whose rotated sites read sigma_t:
// prod_t U[ucode_t][sigma_t][bit_t(i)]   (utils/unitaries.py:169-176)
__device__ __forceinline__ void unitary_product(const double* __restrict__ Usm, const SiteInfo& si, uint32_t sigbits,
                                                uint32_t i, double& re, double& im) {
  re = 1.0; im = 0.0;
  for (int t = 0; t < si.S; ++t) {
    const int row = (sigbits >> t) & 1u;
    const int col = (i >> (si.S - 1 - t)) & 1u;
    const double* u = Usm + si.ucode[t] * 8;
    cmul(re, im, u[row * 2 + col], u[4 + row * 2 + col]);
  }
}

// ---- ComplexWaveFunction kernel 1: T_i and D ---------------------------------------------------
template <int NHT>
__global__ void __launch_bounds__(CWF_THREADS) cwf_amp_kernel(StackedParams am, StackedParams ph, const double* __restrict__ samples,
                                                              int64_t b0, int64_t B, SiteInfo si, const double* __restrict__ U, int K,
                                                              int tps_log2, double2* __restrict__ T, double2* __restrict__ D,
                                                              double* __restrict__ Upsi, int64_t Btotal) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = am.N, nh = am.nh, S = si.S;
  double* Wa = reinterpret_cast<double*>(smem_raw);  // [N][NHT]
  double* Wp = Wa + (size_t)N * NHT;
  double* ca = Wp + (size_t)N * NHT;                 // [NHT]
  double* cp = ca + NHT;
  double* ba = cp + NHT;                             // [N]
  double* bp = ba + round_up(N, 2);
  double* Usm = bp + round_up(N, 2);                 // [K][8]
  double* red = Usm + (size_t)K * 8;                 // [CWF_THREADS][2]
  double* EaP = red + 2 * CWF_THREADS;               // [N][NHT] exp(+W_am), exp(-W_am), exp(+W_ph), exp(-W_ph)
  double* EaM = EaP + (size_t)N * NHT;
  double* EpP = EaM + (size_t)N * NHT;
  double* EpM = EpP + (size_t)N * NHT;
  double* eb = EpM + (size_t)N * NHT;                // [4][N] exp(+-b_am), exp(+-b_ph)
  __shared__ int fast_ok;
  for (int idx = threadIdx.x; idx < N * NHT; idx += blockDim.x) {
    const int j = idx / NHT, i = idx - j * NHT;
    Wa[idx] = i < nh ? am.W[(size_t)i * N + j] : 0.0;
    Wp[idx] = i < nh ? ph.W[(size_t)i * N + j] : 0.0;
  }
  for (int i = threadIdx.x; i < NHT; i += blockDim.x) {
    ca[i] = i < nh ? am.c[i] : 0.0;
    cp[i] = i < nh ? ph.c[i] : 0.0;
  }
  for (int j = threadIdx.x; j < N; j += blockDim.x) { ba[j] = am.b[j]; bp[j] = ph.b[j]; }
  for (int i = threadIdx.x; i < K * 8; i += blockDim.x) Usm[i] = U[i];
  if (threadIdx.x == 0) fast_ok = 1;
  __syncthreads();
  // Product form (see enumerate.cu): exp(theta_i) is carried multiplicatively along the Gray walk and
  // -E = log( e^{b.v} prod_i f_i ), so a state costs one sqrt, one log and one sincos instead of 2 nh
  // softplus evaluations.  Used only when no product can leave the double range for any state.
  for (int idx = threadIdx.x; idx < N * NHT; idx += blockDim.x) {
    const int i = idx % NHT;
    EaP[idx] = i < nh ? exp(Wa[idx]) : 1.0; EaM[idx] = i < nh ? exp(-Wa[idx]) : 1.0;
    EpP[idx] = i < nh ? exp(Wp[idx]) : 1.0; EpM[idx] = i < nh ? exp(-Wp[idx]) : 1.0;
  }
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    eb[j] = exp(ba[j]); eb[N + j] = exp(-ba[j]); eb[2 * N + j] = exp(bp[j]); eb[3 * N + j] = exp(-bp[j]);
  }
  if (threadIdx.x < 2) {  // thread 0: amplitude RBM, thread 1: phase RBM
    const double* Wt = threadIdx.x == 0 ? Wa : Wp;
    const double* cc = threadIdx.x == 0 ? ca : cp;
    const double* bb = threadIdx.x == 0 ? ba : bp;
    double sum_sp = 0.0, max_abs = 0.0, bpos = 0.0, bneg = 0.0;
    bool bad = false;
    for (int i = 0; i < nh; ++i) {
      double hi = cc[i], lo = cc[i];
      for (int j = 0; j < N; ++j) { hi += fmax(Wt[j * NHT + i], 0.0); lo += fmin(Wt[j * NHT + i], 0.0); }
      if (!(fabs(hi) < 1e300) || !(fabs(lo) < 1e300)) bad = true;
      sum_sp += fmax(hi, 0.0) + 0.6931471805599453;
      max_abs = fmax(max_abs, fmax(fabs(hi), fabs(lo)));
    }
    for (int j = 0; j < N; ++j) { if (!(fabs(bb[j]) < 1e300)) bad = true; bpos += fmax(bb[j], 0.0); bneg += fmin(bb[j], 0.0); }
    if (bad || !(sum_sp + bpos < 650.0) || !(bneg > -650.0) || !(max_abs < 650.0)) fast_ok = 0;
  }
  __syncthreads();
  const bool fast = fast_ok != 0;
  const double E20 = 485165195.40979028;  // e^20: torch's softplus threshold in the product form

  const int tps = 1 << tps_log2;            // threads per sample
  const int m = S - tps_log2;               // log2(states per thread)
  const int spc = CWF_THREADS >> tps_log2;  // samples per CTA
  const int sl = threadIdx.x >> tps_log2, tl = threadIdx.x & (tps - 1);
  const int64_t b = (int64_t)blockIdx.x * spc + sl;  // sample index within this chunk
  double dr = 0.0, di = 0.0;
  if (b < B) {
    const double* srow = samples + (b0 + b) * N;
    uint32_t sigbits = 0;  // measured values on the rotated sites
    for (int t = 0; t < S; ++t) sigbits |= (srow[si.site[t]] != 0.0 ? 1u : 0u) << t;
    const uint32_t ibase = (uint32_t)tl << m;
    // visible state of the first expanded index of this thread
    double tha[NHT], thp[NHT];
#pragma unroll
    for (int i = 0; i < NHT; ++i) { tha[i] = ca[i]; thp[i] = cp[i]; }
    double bva = 0.0, bvp = 0.0;
    for (int j = 0; j < N; ++j) {
      bool on = srow[j] != 0.0;
      for (int t = 0; t < S; ++t)
        if (si.site[t] == j) on = (ibase >> (S - 1 - t)) & 1u;
      if (on) {
#pragma unroll
        for (int i = 0; i < NHT; ++i) { tha[i] += Wa[j * NHT + i]; thp[i] += Wp[j * NHT + i]; }
        bva += ba[j]; bvp += bp[j];
      }
    }
    const uint32_t steps = 1u << m;
    double eba = 1.0, ebp = 1.0;
    if (fast) {  // tha / thp now hold exp(theta) (padded units: factor 1)
#pragma unroll
      for (int i = 0; i < NHT; ++i) { tha[i] = i < nh ? exp(tha[i]) : 0.0; thp[i] = i < nh ? exp(thp[i]) : 0.0; }
      eba = exp(bva);
      ebp = exp(bvp);
    }
    for (uint32_t k = 0; k < steps; ++k) {
      const uint32_t i_exp = ibase | (k ^ (k >> 1));
      double amp, phi;  // |psi| = sqrt(exp(-E_am)), phase = -E_ph / 2
      if (fast) {
        double pa0 = eba, pa1 = 1.0, pp0 = ebp, pp1 = 1.0;
#pragma unroll
        for (int i = 0; i < NHT; i += 2) {
          pa0 *= tha[i] > E20 ? tha[i] : 1.0 + tha[i];
          pa1 *= tha[i + 1] > E20 ? tha[i + 1] : 1.0 + tha[i + 1];
          pp0 *= thp[i] > E20 ? thp[i] : 1.0 + thp[i];
          pp1 *= thp[i + 1] > E20 ? thp[i + 1] : 1.0 + thp[i + 1];
        }
        amp = sqrt(pa0 * pa1);
        phi = 0.5 * log(pp0 * pp1);
      } else {
        double xa = bva, xp = bvp;
#pragma unroll
        for (int i = 0; i < NHT; ++i)
          if (i < nh) { xa += softplus20(tha[i]); xp += softplus20(thp[i]); }
        amp = exp(0.5 * xa);
        phi = 0.5 * xp;
      }
      double sn, cs;
      sincos(phi, &sn, &cs);
      double tr, ti;
      unitary_product(Usm, si, sigbits, i_exp, tr, ti);
      cmul(tr, ti, amp * cs, amp * sn);
      T[(b << S) + i_exp] = make_double2(tr, ti);
      dr += tr; di += ti;
      if (k + 1 < steps) {
        const uint32_t k1 = k + 1;
        const int pbit = __ffs(k1) - 1;
        const uint32_t g1 = k1 ^ (k1 >> 1);
        const bool up = (g1 >> pbit) & 1u;
        const double sgn = up ? 1.0 : -1.0;
        const int j = si.site[S - 1 - pbit];
        if (fast) {
          const double* ea = (up ? EaP : EaM) + j * NHT;
          const double* ep = (up ? EpP : EpM) + j * NHT;
#pragma unroll
          for (int i = 0; i < NHT; ++i) { tha[i] *= ea[i]; thp[i] *= ep[i]; }
          eba *= eb[(up ? 0 : N) + j];
          ebp *= eb[(up ? 2 * N : 3 * N) + j];
        } else {
#pragma unroll
          for (int i = 0; i < NHT; ++i) { tha[i] = fma(sgn, Wa[j * NHT + i], tha[i]); thp[i] = fma(sgn, Wp[j * NHT + i], thp[i]); }
          bva = fma(sgn, ba[j], bva);
          bvp = fma(sgn, bp[j], bvp);
        }
      }
    }
  }
  red[2 * threadIdx.x] = dr;
  red[2 * threadIdx.x + 1] = di;
  __syncthreads();
  if (tl == 0 && b < B) {
    double sr = 0.0, sim = 0.0;
    for (int q = 0; q < tps; ++q) { sr += red[2 * (threadIdx.x + q)]; sim += red[2 * (threadIdx.x + q) + 1]; }
    D[b] = make_double2(sr, sim);
    Upsi[b0 + b] = sr;
    Upsi[Btotal + b0 + b] = sim;
  }
}

// ---- ComplexWaveFunction kernel 2: per-row weights Re(T/D), -Im(T/D) -----------------------------
__global__ void cwf_weights_kernel(const double2* __restrict__ T, const double2* __restrict__ D, int S, int64_t rows,
                                   double* __restrict__ wa, double* __restrict__ wp) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const double2 t = T[r], d = D[r >> S];
  const double inv = 1.0 / (d.x * d.x + d.y * d.y);
  wa[r] = (t.x * d.x + t.y * d.y) * inv;
  wp[r] = -((t.y * d.x - t.x * d.y) * inv);
}

// ---- layers beyond the fused amplitude kernel (nh > 32, or tables that do not fit in shared memory): the reference's own
// structure — the energies of the expanded rows from the row-tiled kernel of large.cu, then T_i = Ut_i |psi(v_i)| e^{i phi(v_i)}
// with |psi| = sqrt(exp(-E_am)), phi = -E_ph / 2 (nn_states/wavefunction.py:33-46, complex_wavefunction.py:129-142) ----
int run_energy_expanded_tiled(const StackedParams& p, const Expansion& ex, int64_t row0, int64_t rows, double* E, cudaStream_t st);

__global__ void cwf_combine_kernel(const double* __restrict__ Ea, const double* __restrict__ Ep, const double* __restrict__ samples, int N,
                                   SiteInfo si, const double* __restrict__ U, int64_t rows, double2* __restrict__ T) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int S = si.S;
  const double* srow = samples + (r >> S) * N;
  uint32_t sigbits = 0;
  for (int t = 0; t < S; ++t) sigbits |= (srow[si.site[t]] != 0.0 ? 1u : 0u) << t;
  double tr, ti, sn, cs;
  unitary_product(U, si, sigbits, (uint32_t)(r & (((int64_t)1 << S) - 1)), tr, ti);
  const double amp = exp(-0.5 * Ea[r]);
  sincos(-0.5 * Ep[r], &sn, &cs);
  cmul(tr, ti, amp * cs, amp * sn);
  T[r] = make_double2(tr, ti);
}

// D_b = sum_i T_bi: one warp per sample, lanes stride the expansion, fixed order
__global__ void cwf_sum_kernel(const double2* __restrict__ T, int S, int64_t b0, int64_t Bc, double2* __restrict__ D, double* __restrict__ Upsi,
                               int64_t Btotal) {
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= Bc) return;
  double sr = 0.0, si_ = 0.0;
  for (int64_t i = lane; i < ((int64_t)1 << S); i += 32) {
    const double2 t = T[(b << S) + i];
    sr += t.x;
    si_ += t.y;
  }
  sr = warp_sum(sr);
  si_ = warp_sum(si_);
  if (lane == 0) {
    D[b] = make_double2(sr, si_);
    Upsi[b0 + b] = sr;
    Upsi[Btotal + b0 + b] = si_;
  }
}

// ---- a whole basis group of a SMALL state in one launch -----------------------------------------------------------------
// At most TINY_MAX_ROWS expanded rows, layers of at most 16 units (the two-qubit trainings of the reference's Tutorials 2 / 3:
// a batch is up to nine groups of ~10 samples x 4 rows): the three to five kernels of the general path are microseconds of
// work each and a training epoch launched hundreds of them.  Here CTA 0 serves the amplitude RBM and CTA 1 the phase RBM;
// both evaluate the (cheap) per-row quantities, thread per expanded row, and each reduces its own RBM's gradient with
// grad_reduce_tiny_body.  Same formulas as the general kernels; sums over a sample's rows run in ascending order.
constexpr int TINY_MAX_ROWS = ROWS_THREADS;
constexpr size_t DM_TINY_SMEM_MAX = 96 * 1024;

__device__ __forceinline__ void cwf_group_tiny_body(const StackedParams& am, const StackedParams& ph, const SiteInfo& si, const double* __restrict__ U,
                                                    int rows, int64_t Btotal, double* __restrict__ Upsi, double* w_am, double* w_ph,
                                                    const TinyJob& job, int cta, unsigned char* smem_raw) {
  __shared__ double2 Tsm[TINY_MAX_ROWS];
  const Expansion& ex = job.src.ex;
  const int N = am.N, nh = am.nh, S = si.S;
  const int r = threadIdx.x;
  const int mask = (1 << S) - 1;
  if (r < rows) {
    // T_i = Ut_i |psi(v_i)| e^{i phi(v_i)}, |psi| = sqrt(exp(-E_am)), phi = -E_ph / 2 (complex_wavefunction.py:129-142)
    double xa = 0.0, xp = 0.0;
    for (int j = 0; j < N; ++j) {
      const double vj = ex.value(r, j);
      xa = fma(am.b[j], vj, xa);
      xp = fma(ph.b[j], vj, xp);
    }
    for (int i = 0; i < nh; ++i) {
      double ta = am.c[i], tp = ph.c[i];
      for (int j = 0; j < N; ++j) {
        const double vj = ex.value(r, j);
        ta = fma(am.W[(size_t)i * N + j], vj, ta);
        tp = fma(ph.W[(size_t)i * N + j], vj, tp);
      }
      xa += softplus20(ta);
      xp += softplus20(tp);
    }
    const double amp = exp(0.5 * xa);
    double sn, cs;
    sincos(0.5 * xp, &sn, &cs);
    const int b = r >> S;
    uint32_t sigbits = 0;
    for (int t = 0; t < S; ++t) sigbits |= (ex.samples[(size_t)b * N + si.site[t]] != 0.0 ? 1u : 0u) << t;
    double tr, ti;
    unitary_product(U, si, sigbits, (uint32_t)(r & mask), tr, ti);
    cmul(tr, ti, amp * cs, amp * sn);
    Tsm[r] = make_double2(tr, ti);
  }
  __syncthreads();
  if (r < rows) {
    const int b = r >> S;
    double dr = 0.0, di = 0.0;
    for (int i = 0; i <= mask; ++i) { dr += Tsm[(b << S) + i].x; di += Tsm[(b << S) + i].y; }
    if (cta == 0 && (r & mask) == 0) {
      Upsi[b] = dr;
      Upsi[Btotal + b] = di;
    }
    const double2 t = Tsm[r];
    const double inv = 1.0 / (dr * dr + di * di);
    if (cta == 0) w_am[r] = (t.x * dr + t.y * di) * inv;   // Re(T / D)
    else w_ph[r] = -((t.y * dr - t.x * di) * inv);                // -Im(T / D)
  }
  __syncthreads();
  grad_reduce_tiny_body<ROWS_EXPANDED>(job.s, job.src, rows, job.direct, smem_raw);
}

__global__ void __launch_bounds__(ROWS_THREADS) cwf_group_tiny_kernel(StackedParams am, StackedParams ph, SiteInfo si, const double* __restrict__ U,
                                                                     int rows, int64_t Btotal, double* __restrict__ Upsi, double* w_am,
                                                                     double* w_ph, TinyJob j0, TinyJob j1) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const TinyJob& job = blockIdx.x == 0 ? j0 : j1;  // one inlined copy of the body: these kernels' time is instruction fetch
  cwf_group_tiny_body(am, ph, si, U, rows, Btotal, Upsi, w_am, w_ph, job, (int)blockIdx.x, smem_raw);
}

// ---- all basis groups of a BATCH in one launch: CTA pair g serves the g-th non-empty group (descriptors in global memory,
// written by the host for the whole epoch and copied over once); an all-Z group is the plain positive-phase reduction of the
// amplitude RBM (its second CTA has nothing to do).  Every pair writes its group's gradient to its own partial vector; the
// caller adds the partials in group order (fit_epoch_sum_kernel in api.cu): the sums of the group-by-group path, bit for bit.
struct alignas(16) TinyGroup {
  SiteInfo si;
  int rows, kind, pair_parallel, want_grad;   // kind 0: rotated group, 1: all-Z group
  unsigned long long body_bytes;
  long long Btotal;
  double* out;        // Upsi (2 x samples of the group) / UrhoU (samples of the group)
  double* w[2];
  double* aux[2];
  TinyJob job[2];
};

__device__ __forceinline__ const TinyGroup* load_tiny_group(const TinyGroup* __restrict__ groups, unsigned char* slot) {
  const uint32_t* src = reinterpret_cast<const uint32_t*>(groups + (blockIdx.x >> 1));
  uint32_t* dst = reinterpret_cast<uint32_t*>(slot);
  for (int i = threadIdx.x; i < (int)(sizeof(TinyGroup) / 4); i += blockDim.x) dst[i] = src[i];
  __syncthreads();
  return reinterpret_cast<const TinyGroup*>(slot);
}

__global__ void __launch_bounds__(ROWS_THREADS) cwf_batch_tiny_kernel(StackedParams am, StackedParams ph, const double* __restrict__ U,
                                                                     const TinyGroup* __restrict__ groups) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(16) unsigned char slot[sizeof(TinyGroup)];
  const TinyGroup& grp = *load_tiny_group(groups, slot);
  const int cta = blockIdx.x & 1;
  if (grp.kind == 1) {
    if (cta == 0) grad_reduce_tiny_body<ROWS_EXPLICIT>(grp.job[0].s, grp.job[0].src, grp.rows, grp.job[0].direct, smem_raw);
    return;
  }
  cwf_group_tiny_body(am, ph, grp.si, U, grp.rows, grp.Btotal, grp.out, grp.w[0], grp.w[1], grp.job[cta], cta, smem_raw);
}


static void make_tiny_jobs(const StackedParams& am, const StackedParams& ph, const Expansion& ex, const double* w_am, const double* aux_am,
                           const double* w_ph, const double* aux_ph, int n_aux, int zero_tail_ph, double scale, double* g_am, double* g_ph,
                           TinyJob& j0, TinyJob& j1) {
  const int G = (am.nh + n_aux) * am.N + am.N + am.nh + n_aux;
  j0.s = hidden_only_spec(am);
  j1.s = hidden_only_spec(ph);
  j0.src.ex = ex; j0.src.wrow = w_am; j0.src.aux = aux_am; j0.src.n_aux = n_aux;
  j1.src.ex = ex; j1.src.wrow = w_ph; j1.src.aux = aux_ph; j1.src.n_aux = n_aux;
  j0.direct = GradDirect{g_am, scale, 1, G};
  j1.direct = GradDirect{g_ph, scale, 1, G - zero_tail_ph};
}

int g_rotated_tiny = 1;  // 0: small groups take the general kernels too (test / measurement hook, qcb_set_rotated_mode bit 1)

static int make_sites(const uint8_t* basis_host, int N, int K, SiteInfo& si, Expansion& ex) {
  si.S = 0;
  for (int j = 0; j < N; ++j) {
    ex.sitepos[j] = -1;
    if (basis_host[j] >= K) return fail(QCB_ERR_BAD_ARG, "rotated: basis code %d at site %d >= K = %d", basis_host[j], j, K);
    if (basis_host[j] != 0) {
      if (si.S >= ROT_MAX_S) return fail(QCB_ERR_UNSUPPORTED, "rotated: more than %d rotated sites", ROT_MAX_S);
      si.site[si.S] = (int16_t)j;
      si.ucode[si.S] = basis_host[j];
      ex.sitepos[j] = (int8_t)si.S;
      ++si.S;
    }
  }
  ex.N = N;
  ex.S = si.S;
  return QCB_OK;
}

size_t cwf_workspace_bytes(int N, int nh, int64_t B) {
  const size_t red = grad_reduce_workspace_bytes(N, nh, 0), gray = cwf_gray_workspace_bytes(N);
  return align_up((size_t)ROT_MAX_ROWS * 32, 256) + align_up((size_t)(B > 0 ? B : 1) * 16, 256) + (red > gray ? red : gray);
}

template <int NHT>
static int launch_cwf_amp(const StackedParams& am, const StackedParams& ph, const double* samples, int64_t b0, int64_t Bc,
                          const SiteInfo& si, const double* U, int K, double2* T, double2* D, double* Upsi, int64_t Btotal,
                          cudaStream_t st) {
  const int N = am.N;
  const size_t smem = ((size_t)6 * N * NHT + 2 * NHT + 2 * round_up(N, 2) + (size_t)K * 8 + 2 * CWF_THREADS + 4 * N) * sizeof(double);
  if (smem > SMEM_LIMIT) return fail(QCB_ERR_UNSUPPORTED, "cwf_rotated: parameters do not fit in shared memory");
  QCB_CUDA(cudaFuncSetAttribute(cwf_amp_kernel<NHT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
  // threads per sample: long Gray walks amortise the per-thread start-up when there are many samples
  const int tps_cap = Bc >= 2048 ? 5 : 7;
  int tps_log2 = si.S < tps_cap ? si.S : tps_cap;
  const int spc = CWF_THREADS >> tps_log2;
  const int64_t grid = (Bc + spc - 1) / spc;
  cwf_amp_kernel<NHT><<<(unsigned)grid, CWF_THREADS, smem, st>>>(am, ph, samples, b0, Bc, si, U, K, tps_log2, T, D, Upsi, Btotal);
  QCB_LAUNCH_CHECK();
  return QCB_OK;
}

int run_cwf_rotated(const StackedParams& am, const StackedParams& ph, const double* samples, int64_t B,
                    const uint8_t* basis_host, const double* unitaries, int K, double* Upsi, double* g_am, double* g_ph,
                    void* ws, size_t ws_bytes, cudaStream_t st) {
  const int N = am.N, nh = am.nh;
  if (N > ROT_MAX_N) return fail(QCB_ERR_UNSUPPORTED, "cwf_rotated: N = %d > %d", N, ROT_MAX_N);
  // the fused Gray-walk amplitude kernel keeps six N x NHT tables in shared memory and 2 NHT pre-activations in registers
  const int NHT = nh <= 8 ? 8 : (nh <= 16 ? 16 : 32);
  const bool fused_amp = nh <= 32 && ((size_t)6 * N * NHT + 2 * NHT + 2 * round_up(N, 2) + (size_t)K * 8 + 2 * CWF_THREADS + 4 * N) * sizeof(double) <= SMEM_LIMIT;
  SiteInfo si{};
  Expansion ex{};
  int rc = make_sites(basis_host, N, K, si, ex);
  if (rc != QCB_OK) return rc;
  if (B == 0) return QCB_OK;
  if (ws == nullptr || ws_bytes < cwf_workspace_bytes(N, nh, B)) return fail(QCB_ERR_BAD_ARG, "cwf_rotated: workspace too small");
  char* base = static_cast<char*>(ws);
  double2* T = reinterpret_cast<double2*>(base);
  double* wa = reinterpret_cast<double*>(base + (size_t)ROT_MAX_ROWS * 16);
  double* wp = wa + ROT_MAX_ROWS;
  double2* D = reinterpret_cast<double2*>(base + align_up((size_t)ROT_MAX_ROWS * 32, 256));
  void* gws = base + align_up((size_t)ROT_MAX_ROWS * 32, 256) + align_up((size_t)B * 16, 256);
  const size_t gws_bytes = ws_bytes - (size_t)(static_cast<char*>(gws) - base);
  const int S = si.S;
  if (g_rotated_tiny && g_am != nullptr && S <= 8 && (B << S) <= TINY_MAX_ROWS && N <= 16 && nh <= 16 &&
      grad_reduce_tiny_smem(N, nh, nh, B << S) <= 40 * 1024) {  // a small group: amplitudes, weights and both reductions in one launch
    const int rows = (int)(B << S);
    ex.samples = samples;
    TinyJob j0{}, j1{};
    make_tiny_jobs(am, ph, ex, wa, nullptr, wp, nullptr, 0, 0, 1.0, g_am, g_ph, j0, j1);
    cwf_group_tiny_kernel<<<2, ROWS_THREADS, grad_reduce_tiny_smem(N, nh, nh, rows), st>>>(am, ph, si, unitaries, rows, B, Upsi, wa, wp, j0, j1);
    QCB_LAUNCH_CHECK();
    return QCB_OK;
  }
  const int64_t chunkB = ROT_MAX_ROWS >> S;
  for (int64_t b0 = 0; b0 < B; b0 += chunkB) {
    const int64_t Bc = B - b0 < chunkB ? B - b0 : chunkB;
    const int64_t rows = Bc << S;
    if (!fused_amp) {
      ex.samples = samples + b0 * N;
      rc = run_energy_expanded_tiled(am, ex, 0, rows, wa, st);  // the weight buffers hold the energies until the weights are formed
      if (rc == QCB_OK) rc = run_energy_expanded_tiled(ph, ex, 0, rows, wp, st);
      if (rc != QCB_OK) return rc;
      cwf_combine_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(wa, wp, samples + b0 * N, N, si, unitaries, rows, T);
      QCB_LAUNCH_CHECK();
      cwf_sum_kernel<<<(unsigned)((Bc + 7) / 8), 256, 0, st>>>(T, S, b0, Bc, D, Upsi, B);
      QCB_LAUNCH_CHECK();
    } else if (nh <= 8) rc = launch_cwf_amp<8>(am, ph, samples, b0, Bc, si, unitaries, K, T, D, Upsi, B, st);
    else if (nh <= 16) rc = launch_cwf_amp<16>(am, ph, samples, b0, Bc, si, unitaries, K, T, D, Upsi, B, st);
    else rc = launch_cwf_amp<32>(am, ph, samples, b0, Bc, si, unitaries, K, T, D, Upsi, B, st);
    if (rc != QCB_OK) return rc;
    if (g_am == nullptr) continue;
    if (g_cwf_mode == 0) {  // the rows of a sample differ only in the rotated sites: subset sums instead of outer products
      int handled = 0;
      rc = run_cwf_grad_gray(am, ph, samples + b0 * N, Bc, si, T, D, g_am, g_ph, gws, gws_bytes, &handled, st);
      if (rc != QCB_OK) return rc;
      if (handled) continue;
    }
    cwf_weights_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(T, D, S, rows, wa, wp);
    QCB_LAUNCH_CHECK();
    ex.samples = samples + b0 * N;
    rc = run_grad_reduce_expanded_pair(am, ph, ex, rows, wa, nullptr, wp, nullptr, 0, 0, 1.0, g_am, g_ph, gws, gws_bytes, st);
    if (rc != QCB_OK) return rc;
  }
  return QCB_OK;
}

// ================================= DensityMatrix =================================================
constexpr int DM_MAX_NA = 16;     // thread-per-row pair kernel (register arrays); wider auxiliary layers: warp per row, DM_WIDE_MAX_NA
constexpr int DM_WIDE_MAX_NA = 256;
constexpr int DM_MAX_S = 10;      // 4^S pairs per sample (the reference materialises (2, 2^S, 2^S, B, G) tensors)

// per expanded row: [f_am, f_ph, Ut_re, Ut_im, m_am[na], m_ph[na]]
// f = b.v + sum softplus(W v + c) (rbm/purification_rbm.py:375-381); m_am = U_am v + d_am,
// m_ph = U_ph v (no bias, nn_states/density_matrix.py:130-134)
__global__ void __launch_bounds__(256) dm_feat_kernel(StackedParams am, StackedParams ph, Expansion ex, SiteInfo si,
                                                      const double* __restrict__ U, int64_t rows, double* __restrict__ feat) {
  const int lane = threadIdx.x & 31;
  const int N = am.N, nh = am.nh, na = am.na, S = si.S;
  const int FW = 4 + 2 * na;
  for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * 8) {
    double fa = 0.0, fp = 0.0;
    for (int j = lane; j < N; j += 32) {
      const double vj = ex.value(r, j);
      fa = fma(am.b[j], vj, fa);
      fp = fma(ph.b[j], vj, fp);
    }
    for (int i = lane; i < nh; i += 32) {
      double ta = am.c[i], tp = ph.c[i];
      for (int j = 0; j < N; ++j) {
        const double vj = ex.value(r, j);
        ta = fma(am.W[(size_t)i * N + j], vj, ta);
        tp = fma(ph.W[(size_t)i * N + j], vj, tp);
      }
      fa += softplus20(ta);
      fp += softplus20(tp);
    }
    fa = warp_sum(fa);
    fp = warp_sum(fp);
    double* out = feat + r * FW;
    for (int k = lane; k < na; k += 32) {
      double ma = am.d[k], mp = 0.0;
      for (int j = 0; j < N; ++j) {
        const double vj = ex.value(r, j);
        ma = fma(am.U[(size_t)k * N + j], vj, ma);
        mp = fma(ph.U[(size_t)k * N + j], vj, mp);
      }
      out[4 + k] = ma;
      out[4 + na + k] = mp;
    }
    if (lane == 0) {
      const int64_t b = r >> S;
      uint32_t sigbits = 0;
      for (int t = 0; t < S; ++t) sigbits |= (ex.samples[b * N + si.site[t]] != 0.0 ? 1u : 0u) << t;
      double ur, ui;
      unitary_product(U, si, sigbits, (uint32_t)(r & (((int64_t)1 << S) - 1)), ur, ui);
      out[0] = fa; out[1] = fp; out[2] = ur; out[3] = ui;
    }
  }
}

// thread per expanded row i of sample b: loops over the partner index j
// stats row: [r_i, s_i, R_i[na], I_i[na]]
__device__ __forceinline__ void dm_pair_row(const double* fi, const double* fsample, int na, int S, double* out) {
  const int FW = 4 + 2 * na;
  double ri = 0.0, si = 0.0;
  double R[DM_MAX_NA], I[DM_MAX_NA], Sr[DM_MAX_NA], Si[DM_MAX_NA];
  for (int k = 0; k < na; ++k) { R[k] = 0.0; I[k] = 0.0; }
  const int nexp = 1 << S;
  for (int j = 0; j < nexp; ++j) {
    const double* fj = fsample + (size_t)j * FW;
    const double gp = 0.5 * (fi[0] + fj[0]);   // Gamma+ (amplitude RBM)
    const double gm = 0.5 * (fi[1] - fj[1]);   // Gamma- (phase RBM)
    double pre = 0.0, pim = 0.0;
    for (int k = 0; k < na; ++k) {
      const double x = 0.5 * (fi[4 + k] + fj[4 + k]);
      const double y = 0.5 * (fi[4 + na + k] - fj[4 + na + k]);
      double sn, cs;
      sincos(y, &sn, &cs);
      const double ex = exp(x);
      // Pi (nn_states/density_matrix.py:143-155): log sqrt(1 + 2 e^x cos y + e^2x) = log|1 + e^(x + iy)|, one exponential for
      // e^x, e^2x = (e^x)^2 and e^-x = 1 / e^x (x is an affine function of bounded parameters: no overflow concerns beyond exp's own)
      const double zr = 1.0 + ex * cs, zi = ex * sn;       // 1 + e^(x + iy)
      const double mod2 = zr * zr + zi * zi;               // = 1 + 2 e^x cos y + e^2x
      pre += 0.5 * log(mod2);
      pim += atan2(zi, zr);
      // complex sigmoid 1 / (1 + exp(-(x + iy))) = e^(x + iy) / (1 + e^(x + iy)) = conj(z) e^(x + iy) / |z|^2  (utils/cplx.py:319-334)
      const double inv = 1.0 / mod2;
      const double er = ex * cs, ei = zi;                  // e^(x + iy)
      Sr[k] = (er * zr + ei * zi) * inv;
      Si[k] = (ei * zr - er * zi) * inv;
    }
    const double ampl = exp(gp + pre), phs = gm + pim;
    double sn, cs;
    sincos(phs, &sn, &cs);
    // w = Ut_i conj(Ut_j) rho~
    double wr = fi[2], wi = fi[3];
    cmul(wr, wi, fj[2], -fj[3]);
    cmul(wr, wi, ampl * cs, ampl * sn);
    ri += wr;
    si += wi;
    for (int k = 0; k < na; ++k) {
      R[k] += wr * Sr[k] - wi * Si[k];
      I[k] += wr * Si[k] + wi * Sr[k];
    }
  }
  out[0] = ri;
  out[1] = si;
  for (int k = 0; k < na; ++k) { out[2 + k] = R[k]; out[2 + na + k] = I[k]; }
}

__global__ void __launch_bounds__(128) dm_pair_kernel(const double* __restrict__ feat, int na, int S, int64_t rows,
                                                      double* __restrict__ stats) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int FW = 4 + 2 * na, SW = 2 + 2 * na;
  dm_pair_row(feat + r * FW, feat + ((r >> S) << S) * FW, na, S, stats + r * SW);
}

// The same pair sums for auxiliary layers of more than DM_MAX_NA units: a WARP per expanded row i, lane l owns the auxiliary
// units k = l + 32 m (m < KPL); Pi is a warp sum, w_ij is formed by every lane, R_ik / I_ik stay with the lane that owns k.
template <int KPL>
__global__ void __launch_bounds__(256) dm_pair_wide_kernel(const double* __restrict__ feat, int na, int S, int64_t rows,
                                                           double* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const int FW = 4 + 2 * na, SW = 2 + 2 * na;
  const int nexp = 1 << S;
  for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * 8) {
    const int64_t b = r >> S;
    const double* fi = feat + r * FW;
    double xi[KPL], yi[KPL], R[KPL], I[KPL], Sr[KPL], Si[KPL];
#pragma unroll
    for (int m = 0; m < KPL; ++m) {
      const int k = lane + 32 * m;
      xi[m] = k < na ? fi[4 + k] : 0.0;
      yi[m] = k < na ? fi[4 + na + k] : 0.0;
      R[m] = 0.0;
      I[m] = 0.0;
    }
    const double fi0 = fi[0], fi1 = fi[1], fi2 = fi[2], fi3 = fi[3];
    double ri = 0.0, si = 0.0;
    for (int j = 0; j < nexp; ++j) {
      const double* fj = feat + ((b << S) + j) * FW;
      const double gp = 0.5 * (fi0 + fj[0]);   // Gamma+ (amplitude RBM)
      const double gm = 0.5 * (fi1 - fj[1]);   // Gamma- (phase RBM)
      double pre = 0.0, pim = 0.0;
#pragma unroll
      for (int m = 0; m < KPL; ++m) {
        const int k = lane + 32 * m;
        Sr[m] = 0.0;
        Si[m] = 0.0;
        if (k < na) {
          const double x = 0.5 * (xi[m] + fj[4 + k]);
          const double y = 0.5 * (yi[m] - fj[4 + na + k]);
          double sn, cs;
          sincos(y, &sn, &cs);
          const double ex = exp(x);
          const double zr = 1.0 + ex * cs, zi = ex * sn;            // 1 + e^(x + iy); Pi (nn_states/density_matrix.py:143-155)
          const double mod2 = zr * zr + zi * zi;
          pre += 0.5 * log(mod2);
          pim += atan2(zi, zr);
          const double inv = 1.0 / mod2;                            // complex sigmoid e^(x + iy) / (1 + e^(x + iy)) (utils/cplx.py:319-334)
          const double er = ex * cs, ei = zi;
          Sr[m] = (er * zr + ei * zi) * inv;
          Si[m] = (ei * zr - er * zi) * inv;
        }
      }
      pre = warp_sum(pre);
      pim = warp_sum(pim);
      const double ampl = exp(gp + pre), phs = gm + pim;
      double sn, cs;
      sincos(phs, &sn, &cs);
      double wr = fi2, wi = fi3;   // w = Ut_i conj(Ut_j) rho~
      cmul(wr, wi, fj[2], -fj[3]);
      cmul(wr, wi, ampl * cs, ampl * sn);
      ri += wr;
      si += wi;
#pragma unroll
      for (int m = 0; m < KPL; ++m) {
        R[m] += wr * Sr[m] - wi * Si[m];
        I[m] += wr * Si[m] + wi * Sr[m];
      }
    }
    double* out = stats + r * SW;
    if (lane == 0) {
      out[0] = ri;
      out[1] = si;
    }
#pragma unroll
    for (int m = 0; m < KPL; ++m) {
      const int k = lane + 32 * m;
      if (k < na) {
        out[2 + k] = R[m];
        out[2 + na + k] = I[m];
      }
    }
  }
}

static int launch_dm_pairs(const double* feat, int na, int S, int64_t rows, double* stats, cudaStream_t st) {
  if (na <= DM_MAX_NA) {
    dm_pair_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, st>>>(feat, na, S, rows, stats);
  } else {
    const int64_t want = (rows + 7) / 8, cap = (int64_t)sm_count() * 16;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    if (na <= 32) dm_pair_wide_kernel<1><<<grid, 256, 0, st>>>(feat, na, S, rows, stats);
    else if (na <= 64) dm_pair_wide_kernel<2><<<grid, 256, 0, st>>>(feat, na, S, rows, stats);
    else if (na <= 128) dm_pair_wide_kernel<4><<<grid, 256, 0, st>>>(feat, na, S, rows, stats);
    else dm_pair_wide_kernel<8><<<grid, 256, 0, st>>>(feat, na, S, rows, stats);
  }
  QCB_LAUNCH_CHECK();
  return QCB_OK;
}

// thread per sample: P_b = sum_i r_i (imaginary parts cancel, utils/unitaries.py:274-276)
__global__ void dm_prob_kernel(const double* __restrict__ stats, int na, int S, int64_t B, int64_t b0, double* __restrict__ UrhoU) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int SW = 2 + 2 * na;
  double p = 0.0;
  for (int i = 0; i < (1 << S); ++i) p += stats[((b << S) + i) * SW];
  UrhoU[b0 + b] = p;
}

// per row weights for the two weighted reductions
__global__ void dm_weights_kernel(const double* __restrict__ stats, const double* __restrict__ UrhoU, int na, int S,
                                  int64_t rows, int64_t b0, double* __restrict__ wam, double* __restrict__ auxam,
                                  double* __restrict__ wph, double* __restrict__ auxph) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int SW = 2 + 2 * na;
  const double f = -1.0 / (UrhoU[b0 + (r >> S)] + 1e-8);  // nn_states/density_matrix.py:296,300
  const double* s = stats + r * SW;
  wam[r] = f * s[0];
  wph[r] = -f * s[1];
  for (int k = 0; k < na; ++k) {
    auxam[r * na + k] = f * s[2 + k];
    auxph[r * na + k] = -f * s[2 + na + k];
  }
}

// DensityMatrix: features, pair sums, probabilities, row weights and (want_grad) the two reductions
__device__ __forceinline__ void dm_group_tiny_body(const StackedParams& am, const StackedParams& ph, const SiteInfo& si, const double* __restrict__ U,
                                                   int rows, double* __restrict__ UrhoU, double* w_am, double* aux_am, double* w_ph,
                                                   double* aux_ph, int want_grad, size_t body_bytes, int pair_parallel, const TinyJob& job,
                                                   int cta, unsigned char* smem_raw) {
  const Expansion& ex = job.src.ex;
  const int N = am.N, nh = am.nh, na = am.na, S = si.S;
  const int FW = 4 + 2 * na, SW = 2 + 2 * na;
  double* feat = reinterpret_cast<double*>(smem_raw + body_bytes);  // [rows][FW]
  double* stats = feat + (size_t)rows * FW;                          // [rows][SW]
  const int r = threadIdx.x;
  const int mask = (1 << S) - 1;
  if (r < rows) {
    double fa = 0.0, fp = 0.0;
    for (int j = 0; j < N; ++j) {
      const double vj = ex.value(r, j);
      fa = fma(am.b[j], vj, fa);
      fp = fma(ph.b[j], vj, fp);
    }
    for (int i = 0; i < nh; ++i) {
      double ta = am.c[i], tp = ph.c[i];
      for (int j = 0; j < N; ++j) {
        const double vj = ex.value(r, j);
        ta = fma(am.W[(size_t)i * N + j], vj, ta);
        tp = fma(ph.W[(size_t)i * N + j], vj, tp);
      }
      fa += softplus20(ta);
      fp += softplus20(tp);
    }
    double* out = feat + (size_t)r * FW;
    for (int k = 0; k < na; ++k) {
      double ma = am.d[k], mp = 0.0;
      for (int j = 0; j < N; ++j) {
        const double vj = ex.value(r, j);
        ma = fma(am.U[(size_t)k * N + j], vj, ma);
        mp = fma(ph.U[(size_t)k * N + j], vj, mp);
      }
      out[4 + k] = ma;
      out[4 + na + k] = mp;
    }
    const int b = r >> S;
    uint32_t sigbits = 0;
    for (int t = 0; t < S; ++t) sigbits |= (ex.samples[(size_t)b * N + si.site[t]] != 0.0 ? 1u : 0u) << t;
    double ur, ui;
    unitary_product(U, si, sigbits, (uint32_t)(r & mask), ur, ui);
    out[0] = fa; out[1] = fp; out[2] = ur; out[3] = ui;
  }
  __syncthreads();
  if (pair_parallel) {
    // the pair sums of dm_pair_row with the terms spread over the CTA: a row's 2^S pair terms x na auxiliary units are each a
    // chain of FP64 exp / sincos / log / atan2, and a two-qubit group has three warps' worth of rows.  Stage A: one
    // (row, partner, auxiliary unit) per thread; stage B: one (row, partner); stage C: a thread per row adds in dm_pair_row's order.
    const int nexp = 1 << S;
    double* pk = stats + (size_t)rows * SW;                // [rows * nexp * na][4]: Pi terms (re, im), complex sigmoid (re, im)
    double* wbuf = pk + (size_t)rows * nexp * na * 4;      // [rows * nexp][2]
    for (int item = threadIdx.x; item < rows * nexp * na; item += blockDim.x) {
      const int k = item % na, rj = item / na, ri_ = rj >> S, j = rj & mask;
      const double* fi = feat + (size_t)ri_ * FW;
      const double* fj = feat + (size_t)(((ri_ >> S) << S) + j) * FW;
      const double x = 0.5 * (fi[4 + k] + fj[4 + k]);
      const double y = 0.5 * (fi[4 + na + k] - fj[4 + na + k]);
      double sn, cs;
      sincos(y, &sn, &cs);
      const double ex_ = exp(x);
      const double zr = 1.0 + ex_ * cs, zi = ex_ * sn;
      const double mod2 = zr * zr + zi * zi;
      const double inv = 1.0 / mod2;
      const double er = ex_ * cs, ei = zi;
      double* o = pk + (size_t)item * 4;
      o[0] = 0.5 * log(mod2);
      o[1] = atan2(zi, zr);
      o[2] = (er * zr + ei * zi) * inv;
      o[3] = (ei * zr - er * zi) * inv;
    }
    __syncthreads();
    for (int rj = threadIdx.x; rj < rows * nexp; rj += blockDim.x) {
      const int ri_ = rj >> S, j = rj & mask;
      const double* fi = feat + (size_t)ri_ * FW;
      const double* fj = feat + (size_t)(((ri_ >> S) << S) + j) * FW;
      const double gp = 0.5 * (fi[0] + fj[0]);
      const double gm = 0.5 * (fi[1] - fj[1]);
      double pre = 0.0, pim = 0.0;
      for (int k = 0; k < na; ++k) { pre += pk[((size_t)rj * na + k) * 4]; pim += pk[((size_t)rj * na + k) * 4 + 1]; }
      const double ampl = exp(gp + pre), phs = gm + pim;
      double sn, cs;
      sincos(phs, &sn, &cs);
      double wr = fi[2], wi = fi[3];
      cmul(wr, wi, fj[2], -fj[3]);
      cmul(wr, wi, ampl * cs, ampl * sn);
      wbuf[2 * rj] = wr;
      wbuf[2 * rj + 1] = wi;
    }
    __syncthreads();
    if (r < rows) {
      double ri = 0.0, si_ = 0.0, R[DM_MAX_NA], I[DM_MAX_NA];
      for (int k = 0; k < na; ++k) { R[k] = 0.0; I[k] = 0.0; }
      for (int j = 0; j < nexp; ++j) {
        const int rj = (r << S) + j;
        const double wr = wbuf[2 * rj], wi = wbuf[2 * rj + 1];
        ri += wr;
        si_ += wi;
        for (int k = 0; k < na; ++k) {
          const double Sr = pk[((size_t)rj * na + k) * 4 + 2], Si = pk[((size_t)rj * na + k) * 4 + 3];
          R[k] += wr * Sr - wi * Si;
          I[k] += wr * Si + wi * Sr;
        }
      }
      double* out = stats + (size_t)r * SW;
      out[0] = ri;
      out[1] = si_;
      for (int k = 0; k < na; ++k) { out[2 + k] = R[k]; out[2 + na + k] = I[k]; }
    }
  } else if (r < rows) {
    dm_pair_row(feat + (size_t)r * FW, feat + (size_t)((r >> S) << S) * FW, na, S, stats + (size_t)r * SW);
  }
  __syncthreads();
  if (r < rows) {
    const int b = r >> S;
    double p = 0.0;
    for (int i = 0; i <= mask; ++i) p += stats[(size_t)((b << S) + i) * SW];
    if (cta == 0 && (r & mask) == 0) UrhoU[b] = p;
    if (want_grad) {
      const double f = -1.0 / (p + 1e-8);  // nn_states/density_matrix.py:296,300
      const double* st = stats + (size_t)r * SW;
      if (cta == 0) {
        w_am[r] = f * st[0];
        for (int k = 0; k < na; ++k) aux_am[(size_t)r * na + k] = f * st[2 + k];
      } else {
        w_ph[r] = -f * st[1];
        for (int k = 0; k < na; ++k) aux_ph[(size_t)r * na + k] = -f * st[2 + na + k];
      }
    }
  }
  if (!want_grad) return;
  __syncthreads();
  grad_reduce_tiny_body<ROWS_EXPANDED>(job.s, job.src, rows, job.direct, smem_raw);
}

__global__ void __launch_bounds__(ROWS_THREADS) dm_group_tiny_kernel(StackedParams am, StackedParams ph, SiteInfo si, const double* __restrict__ U,
                                                                    int rows, double* __restrict__ UrhoU, double* w_am, double* aux_am,
                                                                    double* w_ph, double* aux_ph, int want_grad, size_t body_bytes,
                                                                    int pair_parallel, TinyJob j0, TinyJob j1) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const TinyJob& job = blockIdx.x == 0 ? j0 : j1;  // one inlined copy of the body: these kernels' time is instruction fetch
  dm_group_tiny_body(am, ph, si, U, rows, UrhoU, w_am, aux_am, w_ph, aux_ph, want_grad, body_bytes, pair_parallel, job, (int)blockIdx.x, smem_raw);
}

__global__ void __launch_bounds__(ROWS_THREADS) dm_batch_tiny_kernel(StackedParams am, StackedParams ph, const double* __restrict__ U,
                                                                    const TinyGroup* __restrict__ groups) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(16) unsigned char slot[sizeof(TinyGroup)];
  const TinyGroup& grp = *load_tiny_group(groups, slot);
  const int cta = blockIdx.x & 1;
  if (grp.kind == 1) {
    if (cta == 0) grad_reduce_tiny_body<ROWS_EXPLICIT>(grp.job[0].s, grp.job[0].src, grp.rows, grp.job[0].direct, smem_raw);
    return;
  }
  dm_group_tiny_body(am, ph, grp.si, U, grp.rows, grp.out, grp.w[0], grp.aux[0], grp.w[1], grp.aux[1], 1, (size_t)grp.body_bytes,
                     grp.pair_parallel, grp.job[cta], cta, smem_raw);
}


// both of the above in one launch for few rotated sites (every row thread re-adds its sample's 2^S terms in the same order:
// same P_b, same weights; one launch less per basis group of a two-qubit training batch)
constexpr int DM_PW_MAX_S = 4;
__global__ void dm_prob_weights_kernel(const double* __restrict__ stats, int na, int S, int64_t rows, int64_t b0, double* __restrict__ UrhoU,
                                       double* __restrict__ wam, double* __restrict__ auxam, double* __restrict__ wph,
                                       double* __restrict__ auxph, int want_weights) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int SW = 2 + 2 * na;
  const int64_t b = r >> S;
  double p = 0.0;
  for (int i = 0; i < (1 << S); ++i) p += stats[((b << S) + i) * SW];
  if ((r & (((int64_t)1 << S) - 1)) == 0) UrhoU[b0 + b] = p;
  if (!want_weights) return;
  const double f = -1.0 / (p + 1e-8);  // nn_states/density_matrix.py:296,300
  const double* s = stats + r * SW;
  wam[r] = f * s[0];
  wph[r] = -f * s[1];
  for (int k = 0; k < na; ++k) {
    auxam[r * na + k] = f * s[2 + k];
    auxph[r * na + k] = -f * s[2 + na + k];
  }
}

constexpr int64_t DM_MAX_ROWS = (int64_t)1 << 18;
// expanded rows per chunk: 2^18, halved while the per-row scratch of a wide auxiliary layer would exceed 256 MiB
// (never below one sample's 2^DM_MAX_S rows)
static int64_t dm_chunk_rows(int na) {
  const size_t per_row = (size_t)(4 + 2 * na) + (2 + 2 * na) + (2 + 2 * na);
  int64_t rows = DM_MAX_ROWS;
  while (rows > ((int64_t)1 << DM_MAX_S) && (size_t)rows * per_row * sizeof(double) > ((size_t)256 << 20)) rows >>= 1;
  return rows;
}

size_t dm_workspace_bytes(int N, int nh, int na, int64_t B) {
  (void)B;
  const size_t per_row = (size_t)(4 + 2 * na) + (2 + 2 * na) + (2 + 2 * na);
  return align_up((size_t)dm_chunk_rows(na) * per_row * sizeof(double), 256) + grad_reduce_workspace_bytes(N, nh + na, dm_chunk_rows(na));
}

int run_dm_rotated(const StackedParams& am, const StackedParams& ph, const double* samples, int64_t B,
                   const uint8_t* basis_host, const double* unitaries, int K, double* UrhoU, double* g_am, double* g_ph,
                   void* ws, size_t ws_bytes, cudaStream_t st) {
  const int N = am.N, nh = am.nh, na = am.na;
  if (N > ROT_MAX_N) return fail(QCB_ERR_UNSUPPORTED, "dm_rotated: N = %d > %d", N, ROT_MAX_N);
  if (na > DM_WIDE_MAX_NA) return fail(QCB_ERR_UNSUPPORTED, "dm_rotated: na = %d > %d", na, DM_WIDE_MAX_NA);
  SiteInfo si{};
  Expansion ex{};
  int rc = make_sites(basis_host, N, K, si, ex);
  if (rc != QCB_OK) return rc;
  if (si.S > DM_MAX_S) return fail(QCB_ERR_UNSUPPORTED, "dm_rotated: %d rotated sites > %d (4^S pairs per sample)", si.S, DM_MAX_S);
  if (B == 0) return QCB_OK;
  if (ws == nullptr || ws_bytes < dm_workspace_bytes(N, nh, na, B)) return fail(QCB_ERR_BAD_ARG, "dm_rotated: workspace too small");
  const int S = si.S;
  const int FW = 4 + 2 * na, SW = 2 + 2 * na;
  const int64_t max_rows = dm_chunk_rows(na);
  double* feat = static_cast<double*>(ws);
  double* stats = feat + max_rows * FW;
  double* wam = stats + max_rows * SW;
  double* wph = wam + max_rows;
  double* auxam = wph + max_rows;
  double* auxph = auxam + max_rows * na;
  const size_t per_row = (size_t)FW + SW + SW;
  void* gws = static_cast<char*>(ws) + align_up((size_t)max_rows * per_row * sizeof(double), 256);
  const size_t gws_bytes = grad_reduce_workspace_bytes(N, nh + na, max_rows);
  if (g_rotated_tiny && S <= 8 && (B << S) <= TINY_MAX_ROWS && N <= 16 && nh + na <= 16 && na <= DM_MAX_NA) {
    const int rows = (int)(B << S);
    const size_t body = align_up(grad_reduce_tiny_smem(N, nh, nh + na, rows), 16);
    const size_t smem = body + (size_t)rows * (FW + SW) * sizeof(double);
    if (smem <= 44 * 1024) {  // a small group: features, pair sums, probabilities, weights and both reductions in one launch
      ex.samples = samples;
      TinyJob j0{}, j1{};
      make_tiny_jobs(am, ph, ex, wam, auxam, wph, auxph, na, na, -1.0, g_am, g_ph, j0, j1);
      // room for the pair terms of all rows (stage A / B of the kernel): spread over the CTA when they fit
      const size_t pair_bytes = (size_t)rows * ((size_t)1 << S) * ((size_t)na * 4 + 2) * sizeof(double);
      const int pair_parallel = smem + pair_bytes <= DM_TINY_SMEM_MAX ? 1 : 0;
      static std::atomic<uint64_t> attr_done{0};  // per device: cudaFuncSetAttribute applies to the current device only
      int dev = 0;
      QCB_CUDA(cudaGetDevice(&dev));
      if (!((attr_done.load(std::memory_order_relaxed) >> (dev & 63)) & 1ull)) {
        QCB_CUDA(cudaFuncSetAttribute(dm_group_tiny_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DM_TINY_SMEM_MAX));
        attr_done.fetch_or(1ull << (dev & 63), std::memory_order_relaxed);
      }
      dm_group_tiny_kernel<<<g_am != nullptr ? 2 : 1, ROWS_THREADS, smem + (pair_parallel ? pair_bytes : 0), st>>>(
          am, ph, si, unitaries, rows, UrhoU, wam, auxam, wph, auxph, g_am != nullptr ? 1 : 0, body, pair_parallel, j0, j1);
      QCB_LAUNCH_CHECK();
      return QCB_OK;
    }
  }
  const int64_t chunkB = max_rows >> S;
  for (int64_t b0 = 0; b0 < B; b0 += chunkB) {
    const int64_t Bc = B - b0 < chunkB ? B - b0 : chunkB;
    const int64_t rows = Bc << S;
    ex.samples = samples + b0 * N;
    const int64_t fgrid = (rows + 7) / 8;
    dm_feat_kernel<<<(unsigned)(fgrid < 65535 ? fgrid : 65535), 256, 0, st>>>(am, ph, ex, si, unitaries, rows, feat);
    QCB_LAUNCH_CHECK();
    rc = launch_dm_pairs(feat, na, S, rows, stats, st);
    if (rc != QCB_OK) return rc;
    if (S <= DM_PW_MAX_S && rows <= 4096) {
      dm_prob_weights_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(stats, na, S, rows, b0, UrhoU, wam, auxam, wph, auxph,
                                                                           g_am != nullptr ? 1 : 0);
      QCB_LAUNCH_CHECK();
      if (g_am == nullptr) continue;
    } else {
      dm_prob_kernel<<<(unsigned)((Bc + 127) / 128), 128, 0, st>>>(stats, na, S, Bc, b0, UrhoU);
      QCB_LAUNCH_CHECK();
      if (g_am == nullptr) continue;
      dm_weights_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(stats, UrhoU, na, S, rows, b0, wam, auxam, wph, auxph);
      QCB_LAUNCH_CHECK();
    }
    rc = run_grad_reduce_expanded_pair(am, ph, ex, rows, wam, auxam, wph, auxph, na, na, -1.0, g_am, g_ph, gws, gws_bytes, st);
    if (rc != QCB_OK) return rc;
  }
  return QCB_OK;
}

// ---- host side of the one-launch-per-batch path (used by qcb_fit_epoch_bases) ----------------------------------------------
size_t rotated_tiny_group_bytes() { return sizeof(TinyGroup); }

// Device memory one batch needs beside the descriptors: the groups' partial gradients and their row weights / aux features.
size_t rotated_batch_tiny_scratch_bytes(int ngroups, long long n_all, int na) {
  return align_up((size_t)ngroups * n_all * sizeof(double), 256) +
         align_up((size_t)ngroups * TINY_MAX_ROWS * (2 + 2 * (size_t)na) * sizeof(double), 256);
}

// Fills `host` (ngroups slots) with the descriptors of one batch if EVERY non-empty group qualifies for the small-group kernels;
// *n_used = number of descriptors written (0: the batch takes the group-by-group path), *smem = dynamic shared memory of the launch.
// Group u writes its gradient to gpart + u * n_all ([g_am ; g_ph] layout; the caller clears gpart first: an all-Z group leaves the
// phase part alone).  out: 2 doubles per sample of the batch (cwf) / 1 (dm); scratch: rotated_batch_tiny_scratch_bytes() minus gpart.
int prepare_rotated_batch_tiny(bool density, const StackedParams& am, const StackedParams& ph, const double* pos, const int64_t* off,
                               const uint8_t* bases, int ngroups, int K, double* out, double* gpart, long long n_all, void* scratch,
                               void* host_slots, int* n_used, size_t* smem_out) {
  *n_used = 0;
  *smem_out = 0;
  if (!g_rotated_tiny) return QCB_OK;
  const int N = am.N, nh = am.nh, na = density ? am.na : 0;
  const int H = nh + na;
  if (N > 16 || H > 16 || na > DM_MAX_NA) return QCB_OK;
  const long long n_am = (long long)H * N + N + H;
  const int G = (int)n_am;
  TinyGroup* slots = static_cast<TinyGroup*>(host_slots);
  const size_t cap = (size_t)ngroups * TINY_MAX_ROWS;
  double* w_am = static_cast<double*>(scratch);
  double* w_ph = w_am + cap;
  double* aux_am = w_ph + cap;
  double* aux_ph = aux_am + cap * na;
  const int FW = 4 + 2 * na, SW = 2 + 2 * na;
  size_t smem = 0, row_at = 0;
  int u = 0;
  for (int gi = 0; gi < ngroups; ++gi) {
    const int64_t first = off[gi], B = off[gi + 1] - first;
    if (B < 0) return fail(QCB_ERR_BAD_ARG, "fit_epoch_bases: offsets must not decrease");
    if (B == 0) continue;
    const uint8_t* basis = bases + (size_t)gi * N;
    TinyGroup grp{};
    bool plain = true;
    for (int j = 0; j < N; ++j) plain = plain && basis[j] == 0;
    size_t need;
    if (plain) {
      if (B > TINY_MAX_ROWS) return QCB_OK;
      need = grad_reduce_tiny_smem(N, H, H, B);
      if (need > 40 * 1024) return QCB_OK;
      grp.kind = 1;
      grp.rows = (int)B;
      grp.job[0].s = stacked_hidden_spec(am);
      grp.job[0].src.v = pos + first * N;
      grp.job[0].direct = GradDirect{gpart + (size_t)u * n_all, 1.0, 0, G};
    } else {
      SiteInfo si{};
      Expansion ex{};
      int rc = make_sites(basis, N, K, si, ex);
      if (rc != QCB_OK) return rc;
      const int S = si.S;
      if (S > 8 || (B << S) > TINY_MAX_ROWS) return QCB_OK;
      const int rows = (int)(B << S);
      ex.samples = pos + first * N;
      grp.kind = 0;
      grp.si = si;
      grp.rows = rows;
      grp.want_grad = 1;
      grp.Btotal = B;
      grp.w[0] = w_am + row_at;
      grp.w[1] = w_ph + row_at;
      double* gu = gpart + (size_t)u * n_all;
      if (!density) {
        need = grad_reduce_tiny_smem(N, nh, nh, rows);
        if (need > 40 * 1024) return QCB_OK;
        grp.out = out + 2 * (first - off[0]);
        make_tiny_jobs(am, ph, ex, grp.w[0], nullptr, grp.w[1], nullptr, 0, 0, 1.0, gu, gu + n_am, grp.job[0], grp.job[1]);
      } else {
        const size_t body = align_up(grad_reduce_tiny_smem(N, nh, H, rows), 16);
        need = body + (size_t)rows * (FW + SW) * sizeof(double);
        if (need > 44 * 1024) return QCB_OK;
        const size_t pair_bytes = (size_t)rows * ((size_t)1 << S) * ((size_t)na * 4 + 2) * sizeof(double);
        grp.pair_parallel = need + pair_bytes <= DM_TINY_SMEM_MAX ? 1 : 0;
        if (grp.pair_parallel) need += pair_bytes;
        grp.body_bytes = body;
        grp.out = out + (first - off[0]);
        grp.aux[0] = aux_am + row_at * na;
        grp.aux[1] = aux_ph + row_at * na;
        make_tiny_jobs(am, ph, ex, grp.w[0], grp.aux[0], grp.w[1], grp.aux[1], na, na, -1.0, gu, gu + n_am, grp.job[0], grp.job[1]);
      }
      grp.job[0].direct.accumulate = 0;
      grp.job[1].direct.accumulate = 0;
      row_at += (size_t)rows;
    }
    smem = need > smem ? need : smem;
    slots[u++] = grp;
  }
  *n_used = u;
  *smem_out = smem;
  return QCB_OK;
}

int launch_rotated_batch_tiny(bool density, const StackedParams& am, const StackedParams& ph, const double* unitaries, const void* dev_slots,
                              int n_used, size_t smem, cudaStream_t st) {
  const TinyGroup* groups = static_cast<const TinyGroup*>(dev_slots);
  if (density) {
    static std::atomic<uint64_t> attr_done{0};  // per device: cudaFuncSetAttribute applies to the current device only
    int dev = 0;
    QCB_CUDA(cudaGetDevice(&dev));
    if (!((attr_done.load(std::memory_order_relaxed) >> (dev & 63)) & 1ull)) {
      QCB_CUDA(cudaFuncSetAttribute(dm_batch_tiny_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DM_TINY_SMEM_MAX));
      attr_done.fetch_or(1ull << (dev & 63), std::memory_order_relaxed);
    }
    dm_batch_tiny_kernel<<<2 * n_used, ROWS_THREADS, smem, st>>>(am, ph, unitaries, groups);
  } else {
    cwf_batch_tiny_kernel<<<2 * n_used, ROWS_THREADS, smem, st>>>(am, ph, unitaries, groups);
  }
  QCB_LAUNCH_CHECK();
  return QCB_OK;
}

}  // namespace qcb
