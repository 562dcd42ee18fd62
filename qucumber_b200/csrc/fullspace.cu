// Full-space basis rotation and KL terms: the metrics QuCumber evaluates every `period`
// epochs for states trained on rotated-basis data.
//
//  * qcb_kron_rotate applies U_0 (x) U_1 (x) ... (x) U_{n-1} to the leading index of a planar complex array
//    x (2^n, batch) in place — what _kron_mult / rotate_psi / rotate_rho do with a Python loop over sites and
//    NumPy views (utils/unitaries.py:64-150).  One butterfly pass per non-identity site:
//    x'[.., a, ..] = sum_b U[a, b] x[.., b, ..]; site s is bit n-1-s of the index (row order of
//    generate_hilbert_space).  Memory-bound: 32 B read + 32 B written per pair and pass.
//  * qcb_kl_terms reduces  sum p_t log clamp(p_t)  and  sum p_t log clamp(p)  (probs_to_logits clamp,
//    utils/training_statistics.py:17,131-135) for p_t, p given as amplitudes or as probabilities.
#include "common.cuh"

namespace qcb {

constexpr int FS_THREADS = 256;

struct U2 {
  double re[2][2], im[2][2];
};

__global__ void __launch_bounds__(FS_THREADS) kron_site_kernel(double* __restrict__ re, double* __restrict__ im, int p, int64_t npairs,
                                                               int64_t batch, const double* __restrict__ unitary) {
  // unitary: (2, 2, 2) = [re/im][row][col]
  U2 u;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      u.re[a][b] = unitary[a * 2 + b];
      u.im[a][b] = unitary[4 + a * 2 + b];
    }
  const int64_t total = npairs * batch;
  const int64_t low_mask = ((int64_t)1 << p) - 1;
  for (int64_t w = (int64_t)blockIdx.x * FS_THREADS + threadIdx.x; w < total; w += (int64_t)gridDim.x * FS_THREADS) {
    const int64_t q = w / batch, t = w - q * batch;
    const int64_t i0 = ((q & ~low_mask) << 1) | (q & low_mask), i1 = i0 | ((int64_t)1 << p);
    const int64_t e0 = i0 * batch + t, e1 = i1 * batch + t;
    const double x0r = re[e0], x0i = im[e0], x1r = re[e1], x1i = im[e1];
    re[e0] = u.re[0][0] * x0r - u.im[0][0] * x0i + u.re[0][1] * x1r - u.im[0][1] * x1i;
    im[e0] = u.re[0][0] * x0i + u.im[0][0] * x0r + u.re[0][1] * x1i + u.im[0][1] * x1r;
    re[e1] = u.re[1][0] * x0r - u.im[1][0] * x0i + u.re[1][1] * x1r - u.im[1][1] * x1i;
    im[e1] = u.re[1][0] * x0i + u.im[1][0] * x0r + u.re[1][1] * x1i + u.im[1][1] * x1r;
  }
}

int run_kron_rotate(double* re, double* im, int n, int64_t batch, const unsigned char* basis_codes, const double* unitaries, int K,
                    cudaStream_t st) {
  if (n < 0 || n > 40 || batch < 0 || (n > 0 && basis_codes == nullptr) || unitaries == nullptr || K < 1)
    return fail(QCB_ERR_BAD_ARG, "kron_rotate: bad arguments");
  if (batch == 0) return QCB_OK;
  if (re == nullptr || im == nullptr) return fail(QCB_ERR_BAD_ARG, "kron_rotate: re / im must not be NULL");
  const int64_t npairs = n > 0 ? ((int64_t)1 << (n - 1)) : 0;
  for (int s = 0; s < n; ++s) {
    const int code = basis_codes[s];
    if (code >= K) return fail(QCB_ERR_BAD_ARG, "kron_rotate: basis code %d outside the unitary table", code);
    if (code == 0) continue;  // identity
    const int64_t total = npairs * batch;
    const int64_t need = (total + FS_THREADS - 1) / FS_THREADS;
    const int64_t cap = (int64_t)sm_count() * 16;
    kron_site_kernel<<<(int)(need < cap ? need : cap), FS_THREADS, 0, st>>>(re, im, n - 1 - s, npairs, batch, unitaries + (size_t)code * 8);
    QCB_LAUNCH_CHECK();
  }
  return QCB_OK;
}

// partials[2 * block + {0, 1}]
__global__ void __launch_bounds__(FS_THREADS) kl_terms_kernel(const double* __restrict__ t_re, const double* __restrict__ t_im,
                                                              const double* __restrict__ p_re, const double* __restrict__ p_im, int64_t n,
                                                              double zinv, double* __restrict__ partials) {
  __shared__ double red[2][FS_THREADS / 32];
  const double eps = 0x1.0p-52;
  double s0 = 0.0, s1 = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * FS_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * FS_THREADS) {
    const double tr = t_re[i];
    const double pt = t_im != nullptr ? tr * tr + t_im[i] * t_im[i] : tr;
    const double pr = p_re[i];
    const double p = (p_im != nullptr ? pr * pr + p_im[i] * p_im[i] : pr) * zinv;
    s0 = fma(pt, log(fmin(fmax(pt, eps), 1.0 - eps)), s0);
    s1 = fma(pt, log(fmin(fmax(p, eps), 1.0 - eps)), s1);
  }
  s0 = warp_sum(s0);
  s1 = warp_sum(s1);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[0][warp] = s0; red[1][warp] = s1; }
  __syncthreads();
  if (threadIdx.x < 2) {
    double t = 0.0;
    for (int w = 0; w < FS_THREADS / 32; ++w) t += red[threadIdx.x][w];
    partials[2 * blockIdx.x + threadIdx.x] = t;
  }
}

__global__ void kl_combine_kernel(const double* __restrict__ partials, int n, double* __restrict__ out2) {
  if (threadIdx.x >= 2) return;
  double t = 0.0;
  for (int i = 0; i < n; ++i) t += partials[2 * i + threadIdx.x];
  out2[threadIdx.x] = t;
}

static int kl_max_grid() { return sm_count() * 8; }

size_t kl_workspace_bytes() { return align_up((size_t)kl_max_grid() * 2 * sizeof(double), 256); }

int run_kl_terms(const double* t_re, const double* t_im, const double* p_re, const double* p_im, int64_t n, double logZ, double* out2,
                 void* ws, size_t ws_bytes, cudaStream_t st) {
  if (n < 0 || out2 == nullptr || (n > 0 && (t_re == nullptr || p_re == nullptr))) return fail(QCB_ERR_BAD_ARG, "kl_terms: bad arguments");
  if (ws == nullptr || ws_bytes < kl_workspace_bytes()) return fail(QCB_ERR_BAD_ARG, "kl_terms: workspace too small");
  double* partials = reinterpret_cast<double*>(ws);
  int grid = 0;
  if (n > 0) {
    const int64_t need = (n + FS_THREADS - 1) / FS_THREADS;
    grid = (int)(need < kl_max_grid() ? need : kl_max_grid());
    kl_terms_kernel<<<grid, FS_THREADS, 0, st>>>(t_re, t_im, p_re, p_im, n, exp(-logZ), partials);
    QCB_LAUNCH_CHECK();
  }
  kl_combine_kernel<<<1, 32, 0, st>>>(partials, grid, out2);
  QCB_LAUNCH_CHECK();
  return QCB_OK;
}

}  // namespace qcb
