// Rotated-basis gradient of ComplexWaveFunction without the expanded outer products
// (nn_states/complex_wavefunction.py:162-211: g_am = sum_b sum_i Re[T_i / D] grad E_am(v_bi), g_ph likewise with -Im).
//
// rotated.cu reduces the 2^S expanded rows of every sample through the generic weighted gradient reduction — 2^S N nh
// multiply-adds per sample and RBM, 80 % of the call at N = nh = 20, S = 10.  But the rows of ONE sample differ only in the S
// rotated sites.  With p_i = sigmoid(W v_i + c) and the per-row weight w_i:
//   columns of the weight gradient at FIXED sites j:    sum_i w_i p_i v_ij = sigma_j * Q,           Q   = sum_i w_i p_i
//   columns at the rotated site t:                      sum_i w_i p_i [bit_t(i) = 1]       =: Q^t
//   visible-bias entries: the same with p = 1 (a pseudo hidden unit), hidden-bias gradient = Q.
// So a sample needs Q and the S subset sums Q^t (nh numbers each), not an (nh x N) outer product per expanded row.
//  * cwf_gray_kernel: one warp walks a sample's expansion (or a 2^m-state sub-range of it) in Gray-code order, lane h = hidden
//    unit h of BOTH RBMs (lane nh = the pseudo unit).  Per state: e^theta_h carried multiplicatively (one multiply by a
//    shared-memory table entry), p = t / (1 + t) by a reciprocal with two Newton steps, X = w p added to the running sum R;
//    the subset sums need ONE update per state, for the site that just flipped: Q^t -+= R (a set bit contributes
//    R_end - R_start of every interval during which it is set).  The Q^t accumulate over all samples of the warp in shared
//    memory; Q per (sample, sub-range) goes to global memory.  When the products could leave the double range
//    (the same device-side test as cwf_amp_kernel) theta is carried additively and p is the table logistic.
//  * cwf_fixed_kernel: the fixed-site columns and the hidden-bias gradient, sum_b sigma_bj Q_b: one CTA per visible unit.
//  * cwf_partq_reduce_kernel / cwf_assemble_kernel: the per-CTA Q^t partials and the sample chunks of the fixed columns are
//    added in a fixed order, g += -(rotated-site columns | fixed columns).
#include "rows.cuh"

namespace qcb {

constexpr int GRAY_WARPS = 8;
constexpr int GRAY_THREADS = 32 * GRAY_WARPS;
constexpr int GRAY_MAX_SLOTS = 131072;  // (sample, sub-range) pairs per launch

struct GrayArgs {
  StackedParams am, ph;
  const double* samples;   // (Bc, N) rows of this chunk
  int64_t Bc;
  SiteInfo si;
  const double2* T;        // per expanded row of the chunk: T_i = Ut_i psi(v_i)
  const double2* D;        // per sample: sum_i T_i; the row weights are Re(T / D) and -Im(T / D)
  int sub_log2;            // sub-ranges per sample
  double* Qtot;            // [Bc << sub_log2][2][32]
  double* partQ;           // [grid][2][S][32]
};

__device__ __forceinline__ double rcp_newton(double d) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  y = fma(y, fma(-d, y, 1.0), y);
  y = fma(y, fma(-d, y, 1.0), y);
  return y;
}

__global__ void __launch_bounds__(GRAY_THREADS) cwf_gray_kernel(const GrayArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = a.am.N, nh = a.am.nh, S = a.si.S;
  double* Ea = reinterpret_cast<double*>(smem_raw);  // [S][2][32]: e^{+W_am[h][site t]}, e^{-W_am}; additive mode: +W, -W
  double* Ep = Ea + (size_t)S * 64;
  double* Qa = Ep + (size_t)S * 64;                  // [warps][S][32] subset sums, amplitude RBM
  double* Qp = Qa + (size_t)GRAY_WARPS * S * 32;
  __shared__ double e2tab[EXP2_TABLE];
  __shared__ int fast_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  fill_exp2_table(e2tab);
  if (threadIdx.x == 0) fast_s = 1;
  __syncthreads();
  if (threadIdx.x < 2 * nh) {  // can every e^theta stay inside the double range?  (one thread per hidden unit and RBM)
    const StackedParams& p = threadIdx.x < nh ? a.am : a.ph;
    const int h = threadIdx.x < nh ? threadIdx.x : threadIdx.x - nh;
    double hi = p.c[h], lo = p.c[h];
    for (int j = 0; j < N; ++j) { const double w = p.W[(size_t)h * N + j]; hi += fmax(w, 0.0); lo += fmin(w, 0.0); }
    if (!(hi < 600.0) || !(lo > -600.0)) fast_s = 0;
  }
  __syncthreads();
  const bool fast = fast_s != 0;
  for (int idx = threadIdx.x; idx < S * 64; idx += GRAY_THREADS) {
    const int t = idx >> 6, sg = (idx >> 5) & 1, h = idx & 31;
    const int j = a.si.site[t];
    const double wam = h < nh ? a.am.W[(size_t)h * N + j] : 0.0, wph = h < nh ? a.ph.W[(size_t)h * N + j] : 0.0;
    const double sa = sg ? -wam : wam, sp = sg ? -wph : wph;
    Ea[idx] = fast ? exp(sa) : sa;
    Ep[idx] = fast ? exp(sp) : sp;
  }
  for (int idx = threadIdx.x; idx < GRAY_WARPS * S * 32; idx += GRAY_THREADS) { Qa[idx] = 0.0; Qp[idx] = 0.0; }
  __syncthreads();

  const bool unit = lane < nh, ones = lane == nh;  // lane nh: the pseudo unit with p = 1 (visible-bias gradient)
  const int m = S - a.sub_log2;                    // log2(states per warp walk)
  const int64_t slots = a.Bc << a.sub_log2;
  double* myQa = Qa + (size_t)warp * S * 32 + lane;
  double* myQp = Qp + (size_t)warp * S * 32 + lane;
  for (int64_t slot = (int64_t)blockIdx.x * GRAY_WARPS + warp; slot < slots; slot += (int64_t)gridDim.x * GRAY_WARPS) {
    const int64_t b = slot >> a.sub_log2;
    const uint32_t ibase = (uint32_t)(slot & ((1 << a.sub_log2) - 1)) << m;
    const double* srow = a.samples + b * N;
    // theta of the first state of the walk: the sample with its rotated sites set to the bits of ibase
    double tha = 0.0, thp = 0.0;
    if (unit) {
      tha = a.am.c[lane];
      thp = a.ph.c[lane];
      for (int j = 0; j < N; ++j) {
        bool on = srow[j] != 0.0;
        for (int t = 0; t < S; ++t)
          if (a.si.site[t] == j) on = (ibase >> (S - 1 - t)) & 1u;
        if (on) { tha += a.am.W[(size_t)lane * N + j]; thp += a.ph.W[(size_t)lane * N + j]; }
      }
      if (fast) { tha = exp(tha); thp = exp(thp); }
    }
    const double2* Trow = a.T + (b << S);
    const double2 d = a.D[b];
    const double dinv = 1.0 / (d.x * d.x + d.y * d.y);
    double Ra = 0.0, Rp = 0.0;
    const uint32_t steps = 1u << m;
    double wa_l = 0.0, wp_l = 0.0;  // this lane's share of the next 32 states' weights
    for (uint32_t k = 0; k < steps; ++k) {
      if ((k & 31u) == 0u) {  // 32 states' weights at once: lane l fetches state k + l (scattered, all in flight together)
        const uint32_t kl = k + (uint32_t)lane;
        if (kl < steps) {
          const double2 t = Trow[ibase | (kl ^ (kl >> 1))];
          wa_l = (t.x * d.x + t.y * d.y) * dinv;
          wp_l = -((t.y * d.x - t.x * d.y) * dinv);
        }
      }
      const double wak = __shfl_sync(0xffffffffu, wa_l, (int)(k & 31u)), wpk = __shfl_sync(0xffffffffu, wp_l, (int)(k & 31u));
      const uint32_t k1 = k + 1;
      const uint32_t g1 = k1 ^ (k1 >> 1);
      double pa, pp;
      if (fast) {
        pa = unit ? tha * rcp_newton(1.0 + tha) : (ones ? 1.0 : 0.0);
        pp = unit ? thp * rcp_newton(1.0 + thp) : (ones ? 1.0 : 0.0);
      } else {
        pa = unit ? sigmoid_tab(tha, e2tab) : (ones ? 1.0 : 0.0);
        pp = unit ? sigmoid_tab(thp, e2tab) : (ones ? 1.0 : 0.0);
      }
      Ra = fma(wak, pa, Ra);
      Rp = fma(wpk, pp, Rp);
      if (k1 < steps) {
        const int pbit = __ffs(k1) - 1;
        const bool up = (g1 >> pbit) & 1u;
        const int t = S - 1 - pbit;              // rotated site that flips (expansion bit pbit <-> site S - 1 - pbit)
        // the flipped bit opens (up) or closes an interval of states in which it is set: Q^t -= R / += R
        myQa[t * 32] += up ? -Ra : Ra;
        myQp[t * 32] += up ? -Rp : Rp;
        const double ea = Ea[t * 64 + (up ? 0 : 32) + lane], ep = Ep[t * 64 + (up ? 0 : 32) + lane];
        if (fast) { tha *= ea; thp *= ep; } else { tha += ea; thp += ep; }
      }
    }
    // bits still set in the last state close their interval at the end of the walk; the sub-range's fixed high bits count fully
    const uint32_t glast = (steps - 1) ^ ((steps - 1) >> 1);
    const uint32_t setbits = ibase | glast;
    for (int t = 0; t < S; ++t)
      if ((setbits >> (S - 1 - t)) & 1u) { myQa[t * 32] += Ra; myQp[t * 32] += Rp; }
    a.Qtot[(slot * 2 + 0) * 32 + lane] = Ra;
    a.Qtot[(slot * 2 + 1) * 32 + lane] = Rp;
  }
  __syncthreads();
  // this CTA's share of the subset sums, warps added in a fixed order
  double* out = a.partQ + (size_t)blockIdx.x * 2 * S * 32;
  for (int idx = threadIdx.x; idx < S * 32; idx += GRAY_THREADS) {
    double sa = 0.0, sp = 0.0;
#pragma unroll
    for (int w = 0; w < GRAY_WARPS; ++w) { sa += Qa[(size_t)w * S * 32 + idx]; sp += Qp[(size_t)w * S * 32 + idx]; }
    out[idx] = sa;
    out[S * 32 + idx] = sp;
  }
}

constexpr int GRAY_FIX_CHUNKS = 16;

// Gfix[chunk][rbm][j][h] = sum over the chunk's samples b of sigma_bj sum_sub Q[b][sub][rbm][h] for j < N (used for the fixed sites);
// j = N: sigma = 1 (hidden-bias gradient).  grid (N + 1, GRAY_FIX_CHUNKS): the chunks are added in a fixed order by cwf_assemble_kernel.
__global__ void __launch_bounds__(256) cwf_fixed_kernel(const double* __restrict__ samples, int64_t Bc, int N, int sub_log2,
                                                        const double* __restrict__ Qtot, double* __restrict__ Gfix) {
  __shared__ double red[8][64];
  const int j = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nsub = 1 << sub_log2;
  const int64_t per = (Bc + GRAY_FIX_CHUNKS - 1) / GRAY_FIX_CHUNKS;
  const int64_t b_lo = blockIdx.y * per, b_hi = b_lo + per < Bc ? b_lo + per : Bc;
  double sa = 0.0, sp = 0.0;
  for (int64_t b = b_lo + warp; b < b_hi; b += 8) {
    if (j < N && samples[b * N + j] == 0.0) continue;
    for (int s = 0; s < nsub; ++s) {
      const double* q = Qtot + ((b * nsub + s) * 2) * 32;
      sa += q[lane];
      sp += q[32 + lane];
    }
  }
  red[warp][lane] = sa;
  red[warp][32 + lane] = sp;
  __syncthreads();
  if (threadIdx.x < 64) {
    double sum = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += red[w][threadIdx.x];
    Gfix[(((size_t)blockIdx.y * 2 + (threadIdx.x >> 5)) * (N + 1) + j) * 32 + (threadIdx.x & 31)] = sum;
  }
}

// Qsum[rbm][t][h] = sum over the CTAs of cwf_gray_kernel of their subset-sum partials, fixed order.  grid 2 S, 256 threads.
__global__ void __launch_bounds__(256) cwf_partq_reduce_kernel(const double* __restrict__ partQ, int nparts, int S, double* __restrict__ Qsum) {
  __shared__ double red[8][32];
  const int rt = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;  // rt = rbm * S + t
  double sum = 0.0;
  for (int p = warp; p < nparts; p += 8) sum += partQ[((size_t)p * 2 * S + rt) * 32 + lane];
  red[warp][lane] = sum;
  __syncthreads();
  if (threadIdx.x < 32) {
    double tot = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += red[w][threadIdx.x];
    Qsum[(size_t)rt * 32 + threadIdx.x] = tot;
  }
}

// g_am / g_ph += -(flat gradient [W (nh x N) | b (N) | c (nh)])
__global__ void cwf_assemble_kernel(const double* __restrict__ Qsum, const double* __restrict__ Gfix, SiteInfo si, int N, int nh,
                                    double* __restrict__ g_am, double* __restrict__ g_ph) {
  const int G = nh * N + N + nh;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 2 * G) return;
  const int rbm = e / G, f = e - rbm * G;
  int h, j;  // j = N: hidden-bias entry
  if (f < nh * N) { h = f / N; j = f - h * N; }
  else if (f < nh * N + N) { h = nh; j = f - nh * N; }  // visible bias: the pseudo unit
  else { h = f - nh * N - N; j = N; }
  int t = -1;
  if (j < N)
    for (int q = 0; q < si.S; ++q)
      if (si.site[q] == j) t = q;
  double val = 0.0;
  if (t >= 0) {
    val = Qsum[((size_t)rbm * si.S + t) * 32 + h];
  } else {
#pragma unroll
    for (int c = 0; c < GRAY_FIX_CHUNKS; ++c) val += Gfix[(((size_t)c * 2 + rbm) * (N + 1) + j) * 32 + h];
  }
  double* g = rbm == 0 ? g_am : g_ph;
  g[f] -= val;
}

static size_t gray_smem(int S) { return ((size_t)2 * S * 64 + (size_t)2 * GRAY_WARPS * S * 32) * sizeof(double); }

size_t cwf_gray_workspace_bytes(int N) {
  return align_up((size_t)GRAY_MAX_SLOTS * 64 * sizeof(double), 256) + align_up((size_t)sm_count() * 4 * 2 * 21 * 32 * sizeof(double), 256) +
         align_up((size_t)GRAY_FIX_CHUNKS * 2 * (N + 1) * 32 * sizeof(double), 256) + align_up((size_t)2 * 21 * 32 * sizeof(double), 256);
}

// *handled = 1 when this path took the chunk (5 <= S, nh <= 31, chunk small enough for the scratch)
int run_cwf_grad_gray(const StackedParams& am, const StackedParams& ph, const double* samples, int64_t Bc, const SiteInfo& si,
                      const double2* T, const double2* D, double* g_am, double* g_ph, void* ws, size_t ws_bytes, int* handled,
                      cudaStream_t st) {
  *handled = 0;
  const int N = am.N, nh = am.nh, S = si.S;
  if (S < 5 || nh > 31 || Bc < 1 || ws == nullptr || ws_bytes < cwf_gray_workspace_bytes(N)) return QCB_OK;
  // sub-ranges: enough warps to fill the GPU when there are few samples; walks of at most 2^12 states (drift of the products)
  int sub = S > 12 ? S - 12 : 0;
  while (sub < S - 3 && (Bc << sub) < 8192) ++sub;
  if ((Bc << sub) > GRAY_MAX_SLOTS) return QCB_OK;
  const size_t smem = gray_smem(S);
  static bool configured[QCB_MAX_DEVICES] = {false};
  if (!configured[device_slot()]) {
    QCB_CUDA(cudaFuncSetAttribute(cwf_gray_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gray_smem(21)));
    configured[device_slot()] = true;
  }
  char* base = static_cast<char*>(ws);
  GrayArgs a;
  a.am = am; a.ph = ph; a.samples = samples; a.Bc = Bc; a.si = si; a.T = T; a.D = D; a.sub_log2 = sub;
  a.Qtot = reinterpret_cast<double*>(base);
  a.partQ = reinterpret_cast<double*>(base + align_up((size_t)GRAY_MAX_SLOTS * 64 * sizeof(double), 256));
  double* Gfix = reinterpret_cast<double*>(reinterpret_cast<char*>(a.partQ) + align_up((size_t)sm_count() * 4 * 2 * 21 * 32 * sizeof(double), 256));
  const int64_t need = ((Bc << sub) + GRAY_WARPS - 1) / GRAY_WARPS;
  const int64_t cap = (int64_t)sm_count() * 4;
  const int grid = (int)(need < cap ? need : cap);
  cwf_gray_kernel<<<grid, GRAY_THREADS, smem, st>>>(a);
  QCB_LAUNCH_CHECK();
  double* Qsum = reinterpret_cast<double*>(reinterpret_cast<char*>(Gfix) + align_up((size_t)GRAY_FIX_CHUNKS * 2 * (N + 1) * 32 * sizeof(double), 256));
  cwf_fixed_kernel<<<dim3(N + 1, GRAY_FIX_CHUNKS), 256, 0, st>>>(samples, Bc, N, sub, a.Qtot, Gfix);
  QCB_LAUNCH_CHECK();
  cwf_partq_reduce_kernel<<<2 * S, 256, 0, st>>>(a.partQ, grid, S, Qsum);
  QCB_LAUNCH_CHECK();
  const int G2 = 2 * (nh * N + N + nh);
  cwf_assemble_kernel<<<(G2 + 127) / 128, 128, 0, st>>>(Qsum, Gfix, si, N, nh, g_am, g_ph);
  QCB_LAUNCH_CHECK();
  *handled = 1;
  return QCB_OK;
}

}  // namespace qcb
