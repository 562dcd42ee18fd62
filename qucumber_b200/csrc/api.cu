// extern "C" surface of libqucumber_b200.so (declared in include/qcb.h).
#include "common.cuh"
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

namespace qcb {

std::atomic<uint64_t> g_launches{0};
static thread_local std::string t_last_error;

void set_error(const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  t_last_error = buf;
}
int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  t_last_error = buf;
  return code;
}

// implemented in the other translation units
size_t gibbs_workspace_bytes(int N, int nhid);
size_t gibbs_workspace_bytes_rows(int N, int nhid, int64_t rows);
extern int g_gibbs_mode;
extern int g_grad_mode;
extern int g_grad_small_hist;
extern int g_cwf_mode;
extern int g_rotated_tiny;
extern int g_cd_epoch_mode;
int try_gibbs_imma(const StackedParams& p, const ChainIO& io, int64_t C, int k, uint64_t seed, uint64_t chain_id0,
                   uint64_t sweep0, void* ws, size_t ws_bytes, unsigned long long* dbg, int* handled, cudaStream_t st);
int run_gibbs_io(const StackedParams& p, const ChainIO& io, int64_t C, int k, uint64_t seed, uint64_t chain_id0, uint64_t sweep0,
                 void* ws, size_t ws_bytes, cudaStream_t st);
size_t gibbs_io_workspace_bytes(int N, int nhid, int64_t C);
int64_t gibbs_wave_chains(int N, int nhid);
int run_pack_states(const double* v, int64_t rows, int N, uint8_t* bits, cudaStream_t st);
int run_unpack_states(const uint8_t* bits, int64_t rows, int N, double* v, cudaStream_t st);
int run_gibbs(const StackedParams& p, const double* v_in, double* v, int64_t C, int k, uint64_t seed, uint64_t chain_id0, uint64_t sweep0,
              const double* uniforms, void* ws, size_t ws_bytes, cudaStream_t st);
int run_energy(const StackedParams& p, const double* v, int64_t rows, double* E, cudaStream_t st);
int run_linear(int mode, const double* M, int m_rows, int m_cols, int transpose_m, const double* bias, const double* x,
               const double* M2, int m2_rows, const double* x2, int64_t rows, double* out, cudaStream_t st);
int run_bernoulli(const double* p, int64_t n, const double* uniforms, uint64_t seed, uint64_t chain_id0, uint64_t sweep0,
                  double* out, cudaStream_t st);
int run_cd_epoch(double* W, double* b, double* c, int N, int nh, const double* pos, int64_t B_total, int pos_batch, const double* neg,
                 int64_t C_total, int neg_batch, int num_batches, int k, double lr, uint64_t seed, uint64_t sweep0, double* g_last,
                 cudaStream_t st);
size_t grad_reduce_workspace_bytes(int N, int nhid, int64_t rows);
size_t kl_workspace_bytes();
int run_kron_rotate(double* re, double* im, int n, int64_t batch, const unsigned char* basis_codes, const double* unitaries, int K,
                    cudaStream_t st);
int run_kl_terms(const double* t_re, const double* t_im, const double* p_re, const double* p_im, int64_t n, double logZ, double* out2,
                 void* ws, size_t ws_bytes, cudaStream_t st);
bool grad_reduce_bits_supported(int N, int H, int64_t rows);
size_t gibbs_imma_workspace_bytes(int N, int nhid);
int run_grad_reduce_bits(const StackedParams& p, const uint8_t* bits, double* scratch_rows, int64_t rows, double scale, double* g,
                         int accumulate, void* ws, size_t ws_bytes, cudaStream_t st);
int run_grad_reduce(const StackedParams& p, const double* v, int64_t rows, double scale, double* g, int accumulate,
                    void* ws, size_t ws_bytes, cudaStream_t st);
int run_exact_grad_range(const StackedParams& p, uint64_t first, uint64_t count, double logZ, double* g, void* ws,
                         size_t ws_bytes, cudaStream_t st);
int run_grad_batch(const StackedParams& p, const double* v, int64_t rows, double* g_rows, cudaStream_t st);
int run_flip_logratio(const StackedParams& p, const double* v, int64_t rows, double* out, cudaStream_t st);
size_t enum_workspace_bytes();
size_t enum_workspace_bytes_for(int nhid);
int run_logz_range(const StackedParams& p, uint64_t first, uint64_t count, double* out2, void* ws, size_t ws_bytes,
                   cudaStream_t st);
int run_neg_logsumexp(const double* E, int64_t n, double* out2, void* ws, size_t ws_bytes, cudaStream_t st);
int run_exact_stats_range(const StackedParams& p, uint64_t first, uint64_t count, double logZ, const double* t_re,
                          const double* t_im, double* out8, void* ws, size_t ws_bytes, cudaStream_t st);
size_t cwf_workspace_bytes(int N, int nh, int64_t B);
int run_cwf_rotated(const StackedParams& am, const StackedParams& ph, const double* samples, int64_t B,
                    const uint8_t* basis, const double* unitaries, int K, double* Upsi, double* g_am, double* g_ph,
                    void* ws, size_t ws_bytes, cudaStream_t st);
size_t dm_workspace_bytes(int N, int nh, int na, int64_t B);
int run_dm_rotated(const StackedParams& am, const StackedParams& ph, const double* samples, int64_t B,
                   const uint8_t* basis, const double* unitaries, int K, double* UrhoU, double* g_am, double* g_ph,
                   void* ws, size_t ws_bytes, cudaStream_t st);

static int make_params(StackedParams& p, const double* W, const double* U, const double* b, const double* c,
                       const double* d, int N, int nh, int na) {
  if (N < 1 || nh < 1 || na < 0) return fail(QCB_ERR_BAD_ARG, "bad sizes N=%d nh=%d na=%d", N, nh, na);
  if (W == nullptr || b == nullptr || c == nullptr) return fail(QCB_ERR_BAD_ARG, "W, b, c must not be NULL");
  if (na > 0 && (U == nullptr || d == nullptr)) return fail(QCB_ERR_BAD_ARG, "na > 0 needs U and d");
  p.W = W; p.U = U; p.b = b; p.c = c; p.d = d;
  p.N = N; p.nh = nh; p.na = na;
  return QCB_OK;
}

// A process may drive several GPUs: run each call on the device that owns the caller's stream and restore the
// caller's current device afterwards (the legacy NULL stream carries no device: the current one is used).
struct StreamDeviceGuard {
  int prev = -1;
  bool changed = false;
  // stream: the caller's stream; hint: any device pointer of the call (consulted for the legacy NULL stream,
  // which carries no device — PyTorch's default stream is that stream)
  StreamDeviceGuard(void* stream, const void* hint) {
    int dev = -1;
    if (stream != nullptr) {
      // a capturing stream must not be queried (it would invalidate the capture); graph capture runs on the current device
      cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
      if (cudaStreamIsCapturing((cudaStream_t)stream, &cap) != cudaSuccess) { (void)cudaGetLastError(); return; }
      if (cap != cudaStreamCaptureStatusNone) return;
      if (cudaStreamGetDevice((cudaStream_t)stream, &dev) != cudaSuccess) { (void)cudaGetLastError(); dev = -1; }
    } else if (hint != nullptr) {
      cudaPointerAttributes attr;
      if (cudaPointerGetAttributes(&attr, hint) == cudaSuccess && attr.type == cudaMemoryTypeDevice) dev = attr.device;
      else (void)cudaGetLastError();
    }
    if (dev >= 0 && cudaGetDevice(&prev) == cudaSuccess && prev != dev && cudaSetDevice(dev) == cudaSuccess) changed = true;
  }
  ~StreamDeviceGuard() {
    if (changed) (void)cudaSetDevice(prev);
  }
};

// ---- one training epoch over data with measurement bases (qcb_fit_epoch_bases) ----------------------
constexpr int FE_MAX_SEGS = 10;
struct FitEpochSegs {  // the parameter tensors of both RBMs and where their gradients sit in the flat vector [g_am ; g_ph]
  double* p[FE_MAX_SEGS];
  long long first[FE_MAX_SEGS + 1];
  int count;
};

size_t rotated_tiny_group_bytes();
size_t rotated_batch_tiny_scratch_bytes(int ngroups, long long n_all, int na);
int prepare_rotated_batch_tiny(bool density, const StackedParams& am, const StackedParams& ph, const double* pos, const int64_t* off,
                               const uint8_t* bases, int ngroups, int K, double* out, double* gpart, long long n_all, void* scratch,
                               void* host_slots, int* n_used, size_t* smem_out);
int launch_rotated_batch_tiny(bool density, const StackedParams& am, const StackedParams& ph, const double* unitaries, const void* dev_slots,
                              int n_used, size_t smem, cudaStream_t st);

// g <- scale * (p_0 + p_1 + ... ) over the per-group partial gradients of a batch, added in group order (what accumulating group
// after group into a cleared g and scaling it gives)
__global__ void fit_epoch_sum_kernel(double* __restrict__ g, const double* __restrict__ gpart, int nparts, long long n, double scale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double acc = 0.0;
    for (int u = 0; u < nparts; ++u) acc += gpart[(size_t)u * n + i];
    g[i] = acc * scale;
  }
}

__global__ void fit_epoch_scale_kernel(double* __restrict__ g, long long n, double scale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) g[i] *= scale;
}

// p <- p + (-lr) g, the arithmetic of the vanilla SGD step (torch.optim.SGD without momentum / weight decay)
__global__ void fit_epoch_sgd_kernel(FitEpochSegs segs, const double* __restrict__ g, double lr) {
  const long long n = segs.first[segs.count];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int s = 0;
    while (i >= segs.first[s + 1]) ++s;
    double* q = segs.p[s] + (i - segs.first[s]);
    *q = fma(-lr, g[i], *q);
  }
}

int run_fit_epoch_bases(const StackedParams& am, const StackedParams& ph, const FitEpochSegs& segs, bool density, const double* pos,
                        const int64_t* offsets, const uint8_t* bases, int ngroups, int nbatches, const double* unitaries, int K,
                        const double* neg, int64_t C, int k, double lr, uint64_t seed, uint64_t sweep0, double* scratch, double* g,
                        double* vk, void* ws, size_t ws_bytes, void* desc_host, void* desc_dev, size_t desc_bytes, cudaStream_t st) {
  const int N = am.N, nhid = am.nhid();
  const long long n_am = segs.first[segs.count / 2], n_all = segs.first[segs.count];
  double* g_am = g;
  double* g_ph = g + n_am;
  const size_t gb = gibbs_workspace_bytes_rows(N, nhid, C);
  if (C > 0 && ws_bytes < gb + grad_reduce_workspace_bytes(N, nhid, C)) return fail(QCB_ERR_BAD_ARG, "fit_epoch_bases: workspace too small");
  const int blocks = (int)((n_all + 255) / 256 < 148 ? (n_all + 255) / 256 : 148);
  // One launch per batch for the positive phase where every basis group of the batch is small (descriptors for the whole epoch
  // are written into the caller's pinned buffer now and cross in one copy); its device scratch sits at the END of the workspace,
  // past what the sweeps and the negative-phase reduction use.
  std::vector<int> used(nbatches > 0 ? nbatches : 0, 0);
  std::vector<size_t> smem_of(nbatches > 0 ? nbatches : 0, 0);
  const size_t slot = rotated_tiny_group_bytes();
  const size_t carve = rotated_batch_tiny_scratch_bytes(ngroups, n_all, density ? am.na : 0);
  char* carve_base = nullptr;
  if (desc_host != nullptr && desc_dev != nullptr && nbatches > 0 && desc_bytes >= (size_t)nbatches * ngroups * slot &&
      ws_bytes >= carve + 256 + (C > 0 ? gb + grad_reduce_workspace_bytes(N, nhid, C) : 0)) {
    carve_base = static_cast<char*>(ws) + ((ws_bytes - carve) / 256) * 256;
    double* gpart = reinterpret_cast<double*>(carve_base);
    void* wscratch = carve_base + align_up((size_t)ngroups * n_all * sizeof(double), 256);
    bool any = false;
    for (int b = 0; b < nbatches; ++b) {
      const int64_t* off = offsets + (size_t)b * ngroups;
      int rc = prepare_rotated_batch_tiny(density, am, ph, pos, off, bases, ngroups, K, scratch, gpart, n_all, wscratch,
                                          static_cast<char*>(desc_host) + (size_t)b * ngroups * slot, &used[b], &smem_of[b]);
      if (rc != QCB_OK) return rc;
      any = any || used[b] > 0;
    }
    if (any) QCB_CUDA(cudaMemcpyAsync(desc_dev, desc_host, (size_t)nbatches * ngroups * slot, cudaMemcpyHostToDevice, st));
  }
  for (int b = 0; b < nbatches; ++b) {
    const int64_t* off = offsets + (size_t)b * ngroups;
    const int64_t B = off[ngroups] - off[0];
    if (used[b] > 0) {
      double* gpart = reinterpret_cast<double*>(carve_base);
      QCB_CUDA(cudaMemsetAsync(gpart, 0, (size_t)used[b] * n_all * sizeof(double), st));
      int rc = launch_rotated_batch_tiny(density, am, ph, unitaries, static_cast<char*>(desc_dev) + (size_t)b * ngroups * slot, used[b],
                                         smem_of[b], st);
      if (rc != QCB_OK) return rc;
      fit_epoch_sum_kernel<<<blocks, 256, 0, st>>>(g, gpart, used[b], n_all, 1.0 / (double)B);
      QCB_LAUNCH_CHECK();
    } else {
    QCB_CUDA(cudaMemsetAsync(g, 0, (size_t)n_all * sizeof(double), st));
    for (int gi = 0; gi < ngroups; ++gi) {
      const int64_t first = off[gi], rows = off[gi + 1] - first;
      if (rows < 0) return fail(QCB_ERR_BAD_ARG, "fit_epoch_bases: offsets must not decrease");
      if (rows == 0) continue;
      const uint8_t* basis = bases + (size_t)gi * N;
      bool plain = true;
      for (int j = 0; j < N; ++j) plain = plain && basis[j] == 0;
      int rc;
      if (plain) rc = run_grad_reduce(am, pos + first * N, rows, 1.0, g_am, 1, ws, ws_bytes, st);
      else if (density) rc = run_dm_rotated(am, ph, pos + first * N, rows, basis, unitaries, K, scratch, g_am, g_ph, ws, ws_bytes, st);
      else rc = run_cwf_rotated(am, ph, pos + first * N, rows, basis, unitaries, K, scratch, g_am, g_ph, ws, ws_bytes, st);
      if (rc != QCB_OK) return rc;
    }
    if (B > 0) {
      fit_epoch_scale_kernel<<<blocks, 256, 0, st>>>(g, n_all, 1.0 / (double)B);
      QCB_LAUNCH_CHECK();
    }
    }
    if (C > 0) {
      int rc = run_gibbs(am, neg + (size_t)b * C * N, vk, C, k, seed, 0, sweep0 + (uint64_t)b * (uint64_t)k, nullptr, ws, gb, st);
      if (rc != QCB_OK) return rc;
      rc = run_grad_reduce(am, vk, C, -1.0 / (double)C, g_am, 1, static_cast<char*>(ws) + gb, ws_bytes - gb, st);
      if (rc != QCB_OK) return rc;
    }
    if (lr != 0.0) {
      fit_epoch_sgd_kernel<<<blocks, 256, 0, st>>>(segs, g, lr);
      QCB_LAUNCH_CHECK();
    }
  }
  return QCB_OK;
}

}  // namespace qcb

using namespace qcb;

extern "C" {

int qcb_version(void) { return QCB_VERSION; }

const char* qcb_last_error(void) { return t_last_error.c_str(); }

uint64_t qcb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

size_t qcb_workspace_bytes(int op, int N, int nh, int na, int64_t rows) {
  switch (op) {
    case QCB_OP_GIBBS: return gibbs_workspace_bytes_rows(N, nh + na, rows);
    case QCB_OP_ENERGY: return 0;
    case QCB_OP_GRAD_REDUCE: return grad_reduce_workspace_bytes(N, nh + na, rows);
    case QCB_OP_LOGZ: return enum_workspace_bytes_for(nh + na);
    case QCB_OP_EXACT_STATS: {
      const size_t a = enum_workspace_bytes_for(nh + na), b = grad_reduce_workspace_bytes(N, nh + na, rows);
      return a > b ? a : b;
    }
    case QCB_OP_CWF_ROTATED: return cwf_workspace_bytes(N, nh, rows);
    case QCB_OP_DM_ROTATED: return dm_workspace_bytes(N, nh, na, rows);
    case QCB_OP_CD_STEP: {
      const size_t a = gibbs_workspace_bytes_rows(N, nh + na, rows), b = grad_reduce_workspace_bytes(N, nh + na, rows);
      const size_t r = (size_t)(rows > 0 ? rows : 0);
      return a + b + align_up(r * N * sizeof(double), 256) + align_up(r * packed_row_bytes(N), 256);  // v_k as rows, and packed
    }
    case QCB_OP_KL_TERMS: return kl_workspace_bytes();
    case QCB_OP_SAMPLE: return gibbs_io_workspace_bytes(N, nh + na, rows);
    case QCB_OP_CDK_NEG: {
      const size_t a = gibbs_workspace_bytes_rows(N, nh + na, rows), b = grad_reduce_workspace_bytes(N, nh + na, rows);
      return a + b;
    }
    default: return 0;
  }
}

int qcb_rbm_gibbs(const double* W, const double* U, const double* b, const double* c, const double* d, int N, int nh,
                  int na, double* v, int64_t C, int k, uint64_t seed, uint64_t chain_id0, uint64_t sweep0,
                  const double* uniforms, void* ws, size_t ws_bytes, void* stream) {
  StreamDeviceGuard device_guard(stream, W);
  StackedParams p;
  int rc = make_params(p, W, U, b, c, d, N, nh, na);
  if (rc != QCB_OK) return rc;
  if (C < 0 || k < 0 || (C > 0 && v == nullptr)) return fail(QCB_ERR_BAD_ARG, "gibbs: bad C/k/v");
  return run_gibbs(p, v, v, C, k, seed, chain_id0, sweep0, uniforms, ws, ws_bytes, (cudaStream_t)stream);
}

int qcb_set_gibbs_mode(int mode) {
  if (mode < 0 || mode > 5)
    return fail(QCB_ERR_BAD_ARG, "gibbs mode must be 0 (automatic), 1 (FP64 tensor cores only), 2 (tcgen05 integer filter wherever it fits), "
                                 "3 (mma.sync integer filter, never tcgen05), 4 (the general tcgen05 kernel for every shape up to 128 units) or "
                                 "5 (the table kernel for every shape up to 16 units)");
  g_gibbs_mode = mode;
  return QCB_OK;
}

int qcb_set_cd_epoch_mode(int mode) {
  if (mode != 0 && mode != 1) return fail(QCB_ERR_BAD_ARG, "cd_epoch mode must be 0 (automatic) or 1 (never the warp-specialised kernel)");
  g_cd_epoch_mode = mode;
  return QCB_OK;
}

int qcb_set_grad_reduce_mode(int mode) {
  if (mode != 0 && mode != 1 && mode != 3)
    return fail(QCB_ERR_BAD_ARG, "grad_reduce mode must be 0 (automatic), 1 (FP64 tensor cores only) or 3 (integer theta + FP64 outer products, never tcgen05)");
  g_grad_mode = mode;
  g_grad_small_hist = mode == 0 ? 1 : 0;  // the histogram path of small binary layers is part of the automatic selection
  g_cwf_mode = mode != 0 ? 1 : 0;  // modes 1 and 3 also send the rotated-basis gradients through the expanded-row reduction
  g_rotated_tiny = mode == 0 ? 1 : 0;  // ... and small basis groups through the general kernels instead of the one-launch kernels
  return QCB_OK;
}

int qcb_debug_gibbs_filter(const double* W, const double* U, const double* b, const double* c, const double* d, int N, int nh,
                           int na, double* v, int64_t C, int k, uint64_t seed, unsigned long long* stats2, void* ws,
                           size_t ws_bytes, void* stream) {
  StreamDeviceGuard device_guard(stream, W);
  StackedParams p;
  int rc = make_params(p, W, U, b, c, d, N, nh, na);
  if (rc != QCB_OK) return rc;
  if (stats2 == nullptr || v == nullptr) return fail(QCB_ERR_BAD_ARG, "debug_gibbs_filter: NULL argument");
  int handled = 0;
  ChainIO io;
  io.v_in = v;
  io.v_out = v;
  rc = try_gibbs_imma(p, io, C, k, seed, 0, 0, ws, ws_bytes, stats2, &handled, (cudaStream_t)stream);
  if (rc != QCB_OK) return rc;
  if (!handled) return fail(QCB_ERR_UNSUPPORTED, "debug_gibbs_filter: shape not handled by the integer filter kernel");
  return QCB_OK;
}

int qcb_linear_sigmoid(const double* M, int m_rows, int m_cols, int transpose_m, const double* bias, const double* x,
                       const double* M2, int m2_rows, const double* x2, int64_t rows, double* out, void* stream) {
  StreamDeviceGuard device_guard(stream, M);
  if (M == nullptr || x == nullptr || out == nullptr || m_rows < 1 || m_cols < 1 || rows < 0)
    return fail(QCB_ERR_BAD_ARG, "linear_sigmoid: bad arguments");
  if ((M2 == nullptr) != (x2 == nullptr)) return fail(QCB_ERR_BAD_ARG, "linear_sigmoid: M2 and x2 go together");
  return run_linear(1, M, m_rows, m_cols, transpose_m, bias, x, M2, m2_rows, x2, rows, out, (cudaStream_t)stream);
}

int qcb_linear(const double* M, int m_rows, int m_cols, int transpose_m, const double* bias, const double* x, int64_t rows,
               double* out, void* stream) {
  StreamDeviceGuard device_guard(stream, M);
  if (M == nullptr || x == nullptr || out == nullptr || m_rows < 1 || m_cols < 1 || rows < 0)
    return fail(QCB_ERR_BAD_ARG, "linear: bad arguments");
  return run_linear(2, M, m_rows, m_cols, transpose_m, bias, x, nullptr, 0, nullptr, rows, out, (cudaStream_t)stream);
}

int qcb_bernoulli(const double* p, int64_t n, const double* uniforms, uint64_t seed, uint64_t chain_id0, uint64_t sweep0,
                  double* out, void* stream) {
  StreamDeviceGuard device_guard(stream, p);
  if (n < 0 || (n > 0 && (p == nullptr || out == nullptr))) return fail(QCB_ERR_BAD_ARG, "bernoulli: bad arguments");
  return run_bernoulli(p, n, uniforms, seed, chain_id0, sweep0, out, (cudaStream_t)stream);
}

int qcb_rbm_energy(const double* W, const double* U, const double* b, const double* c, const double* d, int N, int nh,
                   int na, const double* v, int64_t rows, double* E_out, void* ws, size_t ws_bytes, void* stream) {
  StreamDeviceGuard device_guard(stream, W);
  (void)ws; (void)ws_bytes;
  StackedParams p;
  int rc = make_params(p, W, U, b, c, d, N, nh, na);
  if (rc != QCB_OK) return rc;
  if (rows < 0 || (rows > 0 && (v == nullptr || E_out == nullptr))) return fail(QCB_ERR_BAD_ARG, "energy: bad rows/v/E_out");
  return run_energy(p, v, rows, E_out, (cudaStream_t)stream);
}

int qcb_rbm_grad_reduce(const double* W, const double* U, const double* b, const double* c, const double* d, int N,
                        int nh, int na, const double* v, int64_t rows, double scale, double* g, int accumulate, void* ws,
                        size_t ws_bytes, void* stream) {
  StreamDeviceGuard device_guard(stream, W);
  StackedParams p;
  int rc = make_params(p, W, U, b, c, d, N, nh, na);
  if (rc != QCB_OK) return rc;
  if (rows < 0 || g == nullptr || (rows > 0 && v == nullptr)) return fail(QCB_ERR_BAD_ARG, "grad_reduce: bad rows/v/g");
  return run_grad_reduce(p, v, rows, scale, g, accumulate, ws, ws_bytes, (cudaStream_t)stream);
}

int qcb_rbm_grad_batch(const double* W, const double* U, const double* b, const double* c, const double* d, int N,
                       int nh, int na, const double* v, int64_t rows, double* g_rows, void* stream) {
  StreamDeviceGuard device_guard(stream, W);
  StackedParams p;
  int rc = make_params(p, W, U, b, c, d, N, nh, na);
  if (rc != QCB_OK) return rc;
  if (rows < 0 || (rows > 0 && (v == nullptr || g_rows == nullptr))) return fail(QCB_ERR_BAD_ARG, "grad_batch: bad rows/v/out");
  return run_grad_batch(p, v, rows, g_rows, (cudaStream_t)stream);
}

int qcb_rbm_flip_logratio(const double* W, const double* U, const double* b, const double* c, const double* d, int N,
                          int nh, int na, const double* v, int64_t rows, double* out, void* stream) {
  StreamDeviceGuard device_guard(stream, W);
  StackedParams p;
  int rc = make_params(p, W, U, b, c, d, N, nh, na);
  if (rc != QCB_OK) return rc;
  if (rows < 0 || (rows > 0 && (v == nullptr || out == nullptr))) return fail(QCB_ERR_BAD_ARG, "flip_logratio: bad rows/v/out");
  return run_flip_logratio(p, v, rows, out, (cudaStream_t)stream);
}

int qcb_rbm_cdk_neg(const double* W, const double* U, const double* b, const double* c, const double* d, int N, int nh,
                    int na, double* v, int64_t C, int k, uint64_t seed, uint64_t chain_id0, uint64_t sweep0,
                    const double* uniforms, double scale, double* g, void* ws, size_t ws_bytes, void* stream) {
  StreamDeviceGuard device_guard(stream, W);
  StackedParams p;
  int rc = make_params(p, W, U, b, c, d, N, nh, na);
  if (rc != QCB_OK) return rc;
  if (C < 0 || k < 0 || g == nullptr || (C > 0 && v == nullptr)) return fail(QCB_ERR_BAD_ARG, "cdk_neg: bad arguments");
  const size_t gb = gibbs_workspace_bytes_rows(N, nh + na, C);
  if (ws == nullptr || ws_bytes < gb + grad_reduce_workspace_bytes(N, nh + na, C))
    return fail(QCB_ERR_BAD_ARG, "cdk_neg: workspace too small");
  rc = run_gibbs(p, v, v, C, k, seed, chain_id0, sweep0, uniforms, ws, gb, (cudaStream_t)stream);
  if (rc != QCB_OK) return rc;
  return run_grad_reduce(p, v, C, scale, g, 1, static_cast<char*>(ws) + gb, ws_bytes - gb, (cudaStream_t)stream);
}

int qcb_rbm_cd_step(const double* W, const double* U, const double* b, const double* c, const double* d, int N, int nh,
                    int na, const double* pos, int64_t B, const double* neg, int64_t C, int k, uint64_t seed,
                    uint64_t chain_id0, uint64_t sweep0, double pos_scale, double neg_scale, double* g, void* ws,
                    size_t ws_bytes, void* stream) {
  StreamDeviceGuard device_guard(stream, W);
  StackedParams p;
  int rc = make_params(p, W, U, b, c, d, N, nh, na);
  if (rc != QCB_OK) return rc;
  if (B < 0 || C < 0 || k < 0 || g == nullptr || (B > 0 && pos == nullptr) || (C > 0 && neg == nullptr))
    return fail(QCB_ERR_BAD_ARG, "cd_step: bad arguments");
  const size_t gb = gibbs_workspace_bytes_rows(N, nh + na, B > C ? B : C);  // sized like qcb_workspace_bytes(QCB_OP_CD_STEP, ..., max(B, C))
  const size_t rb = grad_reduce_workspace_bytes(N, nh + na, B > C ? B : C);
  const size_t vb = align_up((size_t)C * N * sizeof(double), 256);
  const size_t pb = align_up((size_t)C * packed_row_bytes(N), 256);
  if (ws == nullptr || ws_bytes < gb + rb + vb + pb) return fail(QCB_ERR_BAD_ARG, "cd_step: workspace too small");
  char* base = static_cast<char*>(ws);
  double* vk = reinterpret_cast<double*>(base + gb + rb);
  cudaStream_t st = (cudaStream_t)stream;
  rc = run_grad_reduce(p, pos, B, pos_scale, g, 0, base + gb, rb, st);
  if (rc != QCB_OK) return rc;
  // v_k is internal to the step.  Where the integer-filter Gibbs kernels and the tcgen05 gradient kernel both take the shape it
  // travels between them PACKED (8 bytes per chain at N = 64 instead of 512: the Gibbs kernel writes bits, the gradient kernel
  // reads them as the operand bytes it needs anyway); the float64 rows are only materialised if that kernel gives the batch up.
  if (k >= 1 && g_gibbs_mode != 1 && gibbs_imma_workspace_bytes(N, nh + na) != 0 && grad_reduce_bits_supported(N, nh + na, C)) {
    uint8_t* vk_bits = reinterpret_cast<uint8_t*>(base + gb + rb + vb);
    ChainIO io;
    io.v_in = neg;
    io.bits_out = vk_bits;
    rc = run_gibbs_io(p, io, C, k, seed, chain_id0, sweep0, base, gb, st);
    if (rc != QCB_OK) return rc;
    return run_grad_reduce_bits(p, vk_bits, vk, C, neg_scale, g, 1, base + gb, rb, st);
  }
  rc = run_gibbs(p, neg, vk, C, k, seed, chain_id0, sweep0, nullptr, base, gb, st);
  if (rc != QCB_OK) return rc;
  return run_grad_reduce(p, vk, C, neg_scale, g, 1, base + gb, rb, st);
}

int qcb_rbm_cd_step_packed(const double* W, const double* U, const double* b, const double* c, const double* d, int N, int nh,
                           int na, const uint8_t* pos_bits, int64_t B, const uint8_t* neg_bits, int64_t C, int k, uint64_t seed,
                           uint64_t chain_id0, uint64_t sweep0, double pos_scale, double neg_scale, double* g, void* ws,
                           size_t ws_bytes, void* stream) {
  StreamDeviceGuard device_guard(stream, W);
  StackedParams p;
  int rc = make_params(p, W, U, b, c, d, N, nh, na);
  if (rc != QCB_OK) return rc;
  if (B < 1 || C < 1 || k < 1 || g == nullptr || pos_bits == nullptr || neg_bits == nullptr)
    return fail(QCB_ERR_BAD_ARG, "cd_step_packed: bad arguments (needs B, C, k >= 1)");
  if (g_gibbs_mode == 1 || gibbs_imma_workspace_bytes(N, nh + na) == 0 || !grad_reduce_bits_supported(N, nh + na, B) ||
      !grad_reduce_bits_supported(N, nh + na, C))
    return fail(QCB_ERR_UNSUPPORTED, "cd_step_packed: packed rows need both layers of 33..64 units and at least 16384 rows per phase "
                                     "(use qcb_rbm_cd_step with float64 rows)");
  const int64_t R = B > C ? B : C;
  const size_t gb = gibbs_workspace_bytes_rows(N, nh + na, R);
  const size_t rb = grad_reduce_workspace_bytes(N, nh + na, R);
  const size_t vb = align_up((size_t)R * N * sizeof(double), 256);
  const size_t pb = align_up((size_t)R * packed_row_bytes(N), 256);
  if (ws == nullptr || ws_bytes < gb + rb + vb + pb) return fail(QCB_ERR_BAD_ARG, "cd_step_packed: workspace too small (size it as QCB_OP_CD_STEP for max(B, C) rows)");
  char* base = static_cast<char*>(ws);
  double* rows = reinterpret_cast<double*>(base + gb + rb);   // float64 stand-in rows, only written if the tcgen05 kernel gives a batch up
  uint8_t* vk_bits = reinterpret_cast<uint8_t*>(base + gb + rb + vb);
  cudaStream_t st = (cudaStream_t)stream;
  rc = run_grad_reduce_bits(p, pos_bits, rows, B, pos_scale, g, 0, base + gb, rb, st);
  if (rc != QCB_OK) return rc;
  ChainIO io;
  io.bits_in = neg_bits;
  io.bits_out = vk_bits;
  rc = run_gibbs_io(p, io, C, k, seed, chain_id0, sweep0, base, gb, st);
  if (rc != QCB_OK) return rc;
  return run_grad_reduce_bits(p, vk_bits, rows, C, neg_scale, g, 1, base + gb, rb, st);
}

int qcb_rbm_cd_epoch(double* W, double* b, double* c, int N, int nh, const double* pos, int64_t B_total, int pos_batch,
                     const double* neg, int64_t C_total, int neg_batch, int num_batches, int k, double lr, uint64_t seed,
                     uint64_t sweep0, double* g_last, void* stream) {
  StreamDeviceGuard device_guard(stream, W);
  return run_cd_epoch(W, b, c, N, nh, pos, B_total, pos_batch, neg, C_total, neg_batch, num_batches, k, lr, seed, sweep0, g_last,
                      (cudaStream_t)stream);
}

int qcb_rbm_logz_range(const double* W, const double* U, const double* b, const double* c, const double* d, int N,
                       int nh, int na, uint64_t first_state, uint64_t count, double* out2, void* ws, size_t ws_bytes,
                       void* stream) {
  StreamDeviceGuard device_guard(stream, W);
  StackedParams p;
  int rc = make_params(p, W, U, b, c, d, N, nh, na);
  if (rc != QCB_OK) return rc;
  if (out2 == nullptr) return fail(QCB_ERR_BAD_ARG, "logz_range: out2 is NULL");
  return run_logz_range(p, first_state, count, out2, ws, ws_bytes, (cudaStream_t)stream);
}

int qcb_kron_rotate(double* re, double* im, int n, int64_t batch, const uint8_t* basis, const double* unitaries, int K, void* stream) {
  StreamDeviceGuard device_guard(stream, re);
  return run_kron_rotate(re, im, n, batch, basis, unitaries, K, (cudaStream_t)stream);
}

int qcb_kl_terms(const double* t_re, const double* t_im, const double* p_re, const double* p_im, int64_t n, double logZ,
                 double* out2, void* ws, size_t ws_bytes, void* stream) {
  StreamDeviceGuard device_guard(stream, t_re);
  return run_kl_terms(t_re, t_im, p_re, p_im, n, logZ, out2, ws, ws_bytes, (cudaStream_t)stream);
}

int qcb_neg_logsumexp(const double* E, int64_t n, double* out2, void* ws, size_t ws_bytes, void* stream) {
  StreamDeviceGuard device_guard(stream, E);
  if (n < 0 || out2 == nullptr || (n > 0 && E == nullptr)) return fail(QCB_ERR_BAD_ARG, "neg_logsumexp: bad arguments");
  return run_neg_logsumexp(E, n, out2, ws, ws_bytes, (cudaStream_t)stream);
}

int qcb_rbm_exact_stats_range(const double* W, const double* U, const double* b, const double* c, const double* d, int N,
                              int nh, int na, uint64_t first_state, uint64_t count, double logZ, const double* target_re,
                              const double* target_im, double* out8, double* g_exact, void* ws, size_t ws_bytes,
                              void* stream) {
  StreamDeviceGuard device_guard(stream, W);
  StackedParams p;
  int rc = make_params(p, W, U, b, c, d, N, nh, na);
  if (rc != QCB_OK) return rc;
  if (out8 == nullptr) return fail(QCB_ERR_BAD_ARG, "exact_stats: out8 is NULL");
  rc = run_exact_stats_range(p, first_state, count, logZ, target_re, target_im, out8, ws, ws_bytes, (cudaStream_t)stream);
  if (rc != QCB_OK || g_exact == nullptr) return rc;
  return run_exact_grad_range(p, first_state, count, logZ, g_exact, ws, ws_bytes, (cudaStream_t)stream);
}

int qcb_cwf_rotated(const double* W_am, const double* b_am, const double* c_am, const double* W_ph, const double* b_ph,
                    const double* c_ph, int N, int nh, const double* samples, int64_t B, const uint8_t* basis,
                    const double* unitaries, int K, double* Upsi, double* g_am, double* g_ph, void* ws, size_t ws_bytes,
                    void* stream) {
  StreamDeviceGuard device_guard(stream, W_am);
  StackedParams am, ph;
  int rc = make_params(am, W_am, nullptr, b_am, c_am, nullptr, N, nh, 0);
  if (rc != QCB_OK) return rc;
  rc = make_params(ph, W_ph, nullptr, b_ph, c_ph, nullptr, N, nh, 0);
  if (rc != QCB_OK) return rc;
  if (B < 0 || basis == nullptr || unitaries == nullptr || K < 1 || (B > 0 && (samples == nullptr || Upsi == nullptr)))
    return fail(QCB_ERR_BAD_ARG, "cwf_rotated: bad arguments");
  if ((g_am == nullptr) != (g_ph == nullptr)) return fail(QCB_ERR_BAD_ARG, "cwf_rotated: g_am and g_ph go together");
  return run_cwf_rotated(am, ph, samples, B, basis, unitaries, K, Upsi, g_am, g_ph, ws, ws_bytes, (cudaStream_t)stream);
}

int qcb_dm_rotated(const double* W_am, const double* U_am, const double* b_am, const double* c_am, const double* d_am,
                   const double* W_ph, const double* U_ph, const double* b_ph, const double* c_ph, const double* d_ph,
                   int N, int nh, int na, const double* samples, int64_t B, const uint8_t* basis, const double* unitaries,
                   int K, double* UrhoU, double* g_am, double* g_ph, void* ws, size_t ws_bytes, void* stream) {
  StreamDeviceGuard device_guard(stream, W_am);
  StackedParams am, ph;
  if (na < 1) return fail(QCB_ERR_BAD_ARG, "dm_rotated: na must be >= 1");
  int rc = make_params(am, W_am, U_am, b_am, c_am, d_am, N, nh, na);
  if (rc != QCB_OK) return rc;
  rc = make_params(ph, W_ph, U_ph, b_ph, c_ph, d_ph, N, nh, na);
  if (rc != QCB_OK) return rc;
  if (B < 0 || basis == nullptr || unitaries == nullptr || K < 1 || (B > 0 && (samples == nullptr || UrhoU == nullptr)))
    return fail(QCB_ERR_BAD_ARG, "dm_rotated: bad arguments");
  if ((g_am == nullptr) != (g_ph == nullptr)) return fail(QCB_ERR_BAD_ARG, "dm_rotated: g_am and g_ph go together");
  return run_dm_rotated(am, ph, samples, B, basis, unitaries, K, UrhoU, g_am, g_ph, ws, ws_bytes, (cudaStream_t)stream);
}

// A batch whose rows are grouped by measurement basis, every group in ONE call: group gi holds rows offsets[gi] .. offsets[gi + 1]
// of `samples` and was measured in basis bases[gi * N ..]; an all-Z group adds the plain positive-phase gradient of the
// amplitude RBM, the others their rotated-basis gradients — what NeuralStateBase.gradient (nn_states/neural_state.py:326-369)
// does with one Python iteration and one C call per basis.  offsets and bases are HOST arrays.
int qcb_cwf_rotated_groups(const double* W_am, const double* b_am, const double* c_am, const double* W_ph, const double* b_ph,
                           const double* c_ph, int N, int nh, const double* samples, const int64_t* offsets, const uint8_t* bases,
                           int ngroups, const double* unitaries, int K, double* scratch, double* g_am, double* g_ph, void* ws,
                           size_t ws_bytes, void* stream) {
  StreamDeviceGuard device_guard(stream, W_am);
  StackedParams am, ph;
  int rc = make_params(am, W_am, nullptr, b_am, c_am, nullptr, N, nh, 0);
  if (rc != QCB_OK) return rc;
  rc = make_params(ph, W_ph, nullptr, b_ph, c_ph, nullptr, N, nh, 0);
  if (rc != QCB_OK) return rc;
  if (ngroups < 0 || offsets == nullptr || bases == nullptr || unitaries == nullptr || K < 1 || g_am == nullptr || g_ph == nullptr ||
      samples == nullptr || scratch == nullptr)
    return fail(QCB_ERR_BAD_ARG, "cwf_rotated_groups: bad arguments");
  for (int gi = 0; gi < ngroups; ++gi) {
    const int64_t first = offsets[gi], B = offsets[gi + 1] - first;
    if (B < 0) return fail(QCB_ERR_BAD_ARG, "cwf_rotated_groups: offsets must not decrease");
    if (B == 0) continue;
    const uint8_t* basis = bases + (size_t)gi * N;
    bool plain = true;
    for (int j = 0; j < N; ++j) plain = plain && basis[j] == 0;
    if (plain) rc = run_grad_reduce(am, samples + first * N, B, 1.0, g_am, 1, ws, ws_bytes, (cudaStream_t)stream);
    else rc = run_cwf_rotated(am, ph, samples + first * N, B, basis, unitaries, K, scratch, g_am, g_ph, ws, ws_bytes, (cudaStream_t)stream);
    if (rc != QCB_OK) return rc;
  }
  return QCB_OK;
}

int qcb_dm_rotated_groups(const double* W_am, const double* U_am, const double* b_am, const double* c_am, const double* d_am,
                          const double* W_ph, const double* U_ph, const double* b_ph, const double* c_ph, const double* d_ph,
                          int N, int nh, int na, const double* samples, const int64_t* offsets, const uint8_t* bases, int ngroups,
                          const double* unitaries, int K, double* scratch, double* g_am, double* g_ph, void* ws, size_t ws_bytes,
                          void* stream) {
  StreamDeviceGuard device_guard(stream, W_am);
  StackedParams am, ph;
  if (na < 1) return fail(QCB_ERR_BAD_ARG, "dm_rotated_groups: na must be >= 1");
  int rc = make_params(am, W_am, U_am, b_am, c_am, d_am, N, nh, na);
  if (rc != QCB_OK) return rc;
  rc = make_params(ph, W_ph, U_ph, b_ph, c_ph, d_ph, N, nh, na);
  if (rc != QCB_OK) return rc;
  if (ngroups < 0 || offsets == nullptr || bases == nullptr || unitaries == nullptr || K < 1 || g_am == nullptr || g_ph == nullptr ||
      samples == nullptr || scratch == nullptr)
    return fail(QCB_ERR_BAD_ARG, "dm_rotated_groups: bad arguments");
  for (int gi = 0; gi < ngroups; ++gi) {
    const int64_t first = offsets[gi], B = offsets[gi + 1] - first;
    if (B < 0) return fail(QCB_ERR_BAD_ARG, "dm_rotated_groups: offsets must not decrease");
    if (B == 0) continue;
    const uint8_t* basis = bases + (size_t)gi * N;
    bool plain = true;
    for (int j = 0; j < N; ++j) plain = plain && basis[j] == 0;
    if (plain) rc = run_grad_reduce(am, samples + first * N, B, 1.0, g_am, 1, ws, ws_bytes, (cudaStream_t)stream);
    else rc = run_dm_rotated(am, ph, samples + first * N, B, basis, unitaries, K, scratch, g_am, g_ph, ws, ws_bytes, (cudaStream_t)stream);
    if (rc != QCB_OK) return rc;
  }
  return QCB_OK;
}

// One epoch of CD-k training on data with measurement bases in ONE call: per batch the grouped positive phase (as
// qcb_cwf_rotated_groups / qcb_dm_rotated_groups), the k sweeps and the negative phase of the amplitude RBM (as qcb_rbm_cdk_neg)
// and the vanilla SGD step p <- p - lr g on both RBMs — the batch loop of NeuralStateBase.fit (nn_states/neural_state.py:593-623)
// with the reference's default optimizer.  na == 0: ComplexWaveFunction (U_*, d_* NULL); na >= 1: DensityMatrix.
int qcb_fit_epoch_bases(double* W_am, double* U_am, double* b_am, double* c_am, double* d_am, double* W_ph, double* U_ph, double* b_ph,
                        double* c_ph, double* d_ph, int N, int nh, int na, const double* pos, const int64_t* offsets,
                        const uint8_t* bases, int ngroups, int nbatches, const double* unitaries, int K, const double* neg,
                        int64_t neg_batch_size, int k, double lr, uint64_t seed, uint64_t sweep0, double* scratch, double* g,
                        double* vk, void* ws, size_t ws_bytes, void* desc_host, void* desc_dev, size_t desc_bytes, void* stream) {
  StreamDeviceGuard device_guard(stream, W_am);
  StackedParams am, ph;
  if (na < 0) return fail(QCB_ERR_BAD_ARG, "fit_epoch_bases: na must be >= 0");
  int rc = make_params(am, W_am, na ? U_am : nullptr, b_am, c_am, na ? d_am : nullptr, N, nh, na);
  if (rc != QCB_OK) return rc;
  rc = make_params(ph, W_ph, na ? U_ph : nullptr, b_ph, c_ph, na ? d_ph : nullptr, N, nh, na);
  if (rc != QCB_OK) return rc;
  if (ngroups < 1 || nbatches < 0 || offsets == nullptr || bases == nullptr || unitaries == nullptr || K < 1 || g == nullptr ||
      pos == nullptr || scratch == nullptr || neg_batch_size < 0 || k < 0 || (neg_batch_size > 0 && (neg == nullptr || vk == nullptr)) ||
      ws == nullptr)
    return fail(QCB_ERR_BAD_ARG, "fit_epoch_bases: bad arguments");
  FitEpochSegs segs;
  long long at = 0;
  int n = 0;
  for (int net = 0; net < 2; ++net) {  // the flat gradient of an RBM is [W ; U ; b ; c ; d], as its parameters() are ordered
    double* const t[5] = {net ? W_ph : W_am, net ? U_ph : U_am, net ? b_ph : b_am, net ? c_ph : c_am, net ? d_ph : d_am};
    const long long len[5] = {(long long)nh * N, (long long)na * N, N, nh, na};
    for (int j = 0; j < 5; ++j) {
      if (len[j] == 0) continue;
      segs.p[n] = t[j];
      segs.first[n] = at;
      at += len[j];
      ++n;
    }
  }
  segs.first[n] = at;
  segs.count = n;
  return run_fit_epoch_bases(am, ph, segs, na > 0, pos, offsets, bases, ngroups, nbatches, unitaries, K, neg, neg_batch_size, k, lr, seed,
                             sweep0, scratch, g, vk, ws, ws_bytes, desc_host, desc_dev, desc_bytes, (cudaStream_t)stream);
}

size_t qcb_fit_epoch_desc_bytes(int ngroups, int nbatches) {
  if (ngroups < 1 || nbatches < 1) return 0;
  return (size_t)ngroups * (size_t)nbatches * rotated_tiny_group_bytes();
}

// ---- compact states -------------------------------------------------------------------------------
int qcb_pack_states(const double* v, int64_t rows, int N, uint8_t* bits, void* stream) {
  StreamDeviceGuard device_guard(stream, v);
  if (N < 1 || rows < 0 || (rows > 0 && (v == nullptr || bits == nullptr))) return fail(QCB_ERR_BAD_ARG, "pack_states: bad arguments");
  return run_pack_states(v, rows, N, bits, (cudaStream_t)stream);
}

int qcb_unpack_states(const uint8_t* bits, int64_t rows, int N, double* v, void* stream) {
  StreamDeviceGuard device_guard(stream, v);
  if (N < 1 || rows < 0 || (rows > 0 && (v == nullptr || bits == nullptr))) return fail(QCB_ERR_BAD_ARG, "unpack_states: bad arguments");
  return run_unpack_states(bits, rows, N, v, (cudaStream_t)stream);
}

int qcb_rbm_sample(const double* W, const double* U, const double* b, const double* c, const double* d, int N, int nh, int na,
                   int64_t C, int k, uint64_t seed, uint64_t chain_id0, uint64_t sweep0, double* v_out, uint8_t* bits_out,
                   void* ws, size_t ws_bytes, void* stream) {
  StreamDeviceGuard device_guard(stream, W);
  StackedParams p;
  int rc = make_params(p, W, U, b, c, d, N, nh, na);
  if (rc != QCB_OK) return rc;
  if (C < 0 || k < 0 || (C > 0 && v_out == nullptr && bits_out == nullptr)) return fail(QCB_ERR_BAD_ARG, "sample: bad C/k or no output buffer");
  ChainIO io;
  io.philox_init = 1;
  io.v_out = v_out;
  io.bits_out = bits_out;
  return run_gibbs_io(p, io, C, k, seed, chain_id0, sweep0, ws, ws_bytes, (cudaStream_t)stream);
}

int qcb_rbm_gibbs_packed(const double* W, const double* U, const double* b, const double* c, const double* d, int N, int nh,
                         int na, uint8_t* bits, int64_t C, int k, uint64_t seed, uint64_t chain_id0, uint64_t sweep0, void* ws,
                         size_t ws_bytes, void* stream) {
  StreamDeviceGuard device_guard(stream, W);
  StackedParams p;
  int rc = make_params(p, W, U, b, c, d, N, nh, na);
  if (rc != QCB_OK) return rc;
  if (C < 0 || k < 0 || (C > 0 && bits == nullptr)) return fail(QCB_ERR_BAD_ARG, "gibbs_packed: bad C/k/bits");
  ChainIO io;
  io.bits_in = bits;
  io.bits_out = bits;
  return run_gibbs_io(p, io, C, k, seed, chain_id0, sweep0, ws, ws_bytes, (cudaStream_t)stream);
}

}  // extern "C"

// ---- host-buffer calls ----------------------------------------------------------------------------
namespace {

constexpr int MAX_SLICES = 64;
struct HostPipe {  // copy streams and events of one device (created on first use there); one call at a time per device
  std::mutex lock;
  cudaStream_t s_in = nullptr, s_out = nullptr;
  cudaEvent_t ev_in[MAX_SLICES], ev_k[MAX_SLICES], ev_params, ev_done;
};
HostPipe g_pipes[QCB_MAX_DEVICES];

int pipe_ready(HostPipe& hp) {
  if (hp.s_in != nullptr) return QCB_OK;
  QCB_CUDA(cudaStreamCreateWithFlags(&hp.s_in, cudaStreamNonBlocking));
  QCB_CUDA(cudaStreamCreateWithFlags(&hp.s_out, cudaStreamNonBlocking));
  for (int i = 0; i < MAX_SLICES; ++i) {
    QCB_CUDA(cudaEventCreateWithFlags(&hp.ev_in[i], cudaEventDisableTiming));
    QCB_CUDA(cudaEventCreateWithFlags(&hp.ev_k[i], cudaEventDisableTiming));
  }
  QCB_CUDA(cudaEventCreateWithFlags(&hp.ev_params, cudaEventDisableTiming));
  QCB_CUDA(cudaEventCreateWithFlags(&hp.ev_done, cudaEventDisableTiming));
  return QCB_OK;
}

size_t param_bytes(int N, int nh, int na) { return align_up(((size_t)(nh + na) * N + N + nh + na) * sizeof(double), 256); }

// copies the host parameters into dev_scratch on st and fills p with the device views
int stage_params(StackedParams& p, const double* W_host, const double* U_host, const double* b_host, const double* c_host,
                 const double* d_host, int N, int nh, int na, void* dev_scratch, cudaStream_t st) {
  double* dW = static_cast<double*>(dev_scratch);
  double* dU = dW + (size_t)nh * N;
  double* db = dU + (size_t)na * N;
  double* dc = db + N;
  double* dd = dc + nh;
  QCB_CUDA(cudaMemcpyAsync(dW, W_host, (size_t)nh * N * sizeof(double), cudaMemcpyHostToDevice, st));
  if (na > 0) {
    QCB_CUDA(cudaMemcpyAsync(dU, U_host, (size_t)na * N * sizeof(double), cudaMemcpyHostToDevice, st));
    QCB_CUDA(cudaMemcpyAsync(dd, d_host, (size_t)na * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  QCB_CUDA(cudaMemcpyAsync(db, b_host, (size_t)N * sizeof(double), cudaMemcpyHostToDevice, st));
  QCB_CUDA(cudaMemcpyAsync(dc, c_host, (size_t)nh * sizeof(double), cudaMemcpyHostToDevice, st));
  return make_params(p, dW, na > 0 ? dU : nullptr, db, dc, na > 0 ? dd : nullptr, N, nh, na);
}

// The chains are independent, so a host call is pipelined in slices: slice i+1 travels host->device and slice i-1
// device->host (PCIe is full duplex, two copy engines) while slice i is sampled.  Philox draws are keyed by global chain
// id, so slicing does not change any sample.  row_bytes: bytes of one chain in the host / device buffers; copy_in: whether
// the host buffer holds the initial states; run(c0, cn): enqueue the sampling of chains [c0, c0 + cn) on st.
template <class Run>
int host_pipeline(HostPipe& hp, char* host, char* dev, size_t row_bytes, int64_t C, int64_t per, bool copy_in, cudaStream_t st, Run run) {
  if (C <= 0) return QCB_OK;
  QCB_CUDA(cudaEventRecord(hp.ev_params, st));
  QCB_CUDA(cudaStreamWaitEvent(hp.s_in, hp.ev_params, 0));
  int i = 0;
  for (int64_t c0 = 0; c0 < C; c0 += per, ++i) {
    const int64_t cn = C - c0 < per ? C - c0 : per;
    const size_t off = (size_t)c0 * row_bytes, nb = (size_t)cn * row_bytes;
    if (copy_in) {
      QCB_CUDA(cudaMemcpyAsync(dev + off, host + off, nb, cudaMemcpyHostToDevice, hp.s_in));
      QCB_CUDA(cudaEventRecord(hp.ev_in[i], hp.s_in));
      QCB_CUDA(cudaStreamWaitEvent(st, hp.ev_in[i], 0));
    }
    const int rc = run(c0, cn);
    if (rc != QCB_OK) return rc;
    QCB_CUDA(cudaEventRecord(hp.ev_k[i], st));
    QCB_CUDA(cudaStreamWaitEvent(hp.s_out, hp.ev_k[i], 0));
    QCB_CUDA(cudaMemcpyAsync(host + off, dev + off, nb, cudaMemcpyDeviceToHost, hp.s_out));
  }
  QCB_CUDA(cudaEventRecord(hp.ev_done, hp.s_out));
  QCB_CUDA(cudaStreamWaitEvent(st, hp.ev_done, 0));
  return QCB_OK;
}

}  // namespace

extern "C" {

size_t qcb_gibbs_host_scratch_bytes(int N, int nh, int na, int64_t C) {
  return param_bytes(N, nh, na) + align_up((size_t)C * N * sizeof(double), 256) + gibbs_workspace_bytes_rows(N, nh + na, C);
}

int qcb_rbm_gibbs_host(const double* W_host, const double* U_host, const double* b_host, const double* c_host,
                       const double* d_host, int N, int nh, int na, double* v_host, int64_t C, int k, uint64_t seed,
                       uint64_t chain_id0, uint64_t sweep0, void* dev_scratch, size_t dev_scratch_bytes, void* stream) {
  StreamDeviceGuard device_guard(stream, dev_scratch);
  if (N < 1 || nh < 1 || na < 0 || C < 0 || k < 0) return fail(QCB_ERR_BAD_ARG, "gibbs_host: bad sizes");
  if (W_host == nullptr || b_host == nullptr || c_host == nullptr || (na > 0 && (U_host == nullptr || d_host == nullptr)) ||
      (C > 0 && v_host == nullptr))
    return fail(QCB_ERR_BAD_ARG, "gibbs_host: NULL argument");
  if (dev_scratch == nullptr || dev_scratch_bytes < qcb_gibbs_host_scratch_bytes(N, nh, na, C))
    return fail(QCB_ERR_BAD_ARG, "gibbs_host: device scratch too small");
  cudaStream_t st = (cudaStream_t)stream;
  HostPipe& hp = g_pipes[device_slot()];
  std::lock_guard<std::mutex> hold(hp.lock);
  int rc = pipe_ready(hp);
  if (rc != QCB_OK) return rc;
  StackedParams p;
  rc = stage_params(p, W_host, U_host, b_host, c_host, d_host, N, nh, na, dev_scratch, st);
  if (rc != QCB_OK) return rc;
  const size_t params = param_bytes(N, nh, na);
  double* dv = reinterpret_cast<double*>(static_cast<char*>(dev_scratch) + params);
  const size_t vbytes = (size_t)C * N * sizeof(double);
  void* ws = static_cast<char*>(dev_scratch) + params + align_up(vbytes, 256);
  // Slices of ~32 MiB: 8 MiB slices were measured slower (14.0 vs 12.2 ms for 2 x 512 MiB: per-slice copy set-up,
  // event hand-offs and two small launches outweigh the shorter fill and drain).
  int nslices = (int)(vbytes / ((size_t)32 << 20));
  nslices = nslices < 1 ? 1 : (nslices > MAX_SLICES ? MAX_SLICES : nslices);
  const int64_t per = ((C + nslices - 1) / nslices + 15) / 16 * 16;
  const size_t gws = gibbs_workspace_bytes_rows(N, nh + na, C);
  rc = host_pipeline(hp, reinterpret_cast<char*>(v_host), reinterpret_cast<char*>(dv), (size_t)N * sizeof(double), C, per, true, st,
                     [&](int64_t c0, int64_t cn) {
                       return run_gibbs(p, dv + c0 * N, dv + c0 * N, cn, k, seed, chain_id0 + (uint64_t)c0, sweep0, nullptr, ws, gws, st);
                     });
  if (rc != QCB_OK) return rc;
  QCB_CUDA(cudaStreamSynchronize(st));
  return QCB_OK;
}

size_t qcb_gibbs_host_packed_scratch_bytes(int N, int nh, int na, int64_t C) {
  const int64_t wave = gibbs_wave_chains(N, nh + na);
  const int64_t slice = C < 8 * wave ? C : 8 * wave;  // the float64 staging rows (shapes outside the integer kernel) cover one slice
  return param_bytes(N, nh, na) + align_up((size_t)(C > 0 ? C : 0) * packed_row_bytes(N), 256) + gibbs_io_workspace_bytes(N, nh + na, slice);
}

int qcb_rbm_gibbs_host_packed(const double* W_host, const double* U_host, const double* b_host, const double* c_host,
                              const double* d_host, int N, int nh, int na, uint8_t* bits_host, int64_t C, int k, uint64_t seed,
                              uint64_t chain_id0, uint64_t sweep0, int draw_initial, void* dev_scratch, size_t dev_scratch_bytes,
                              void* stream) {
  StreamDeviceGuard device_guard(stream, dev_scratch);
  if (N < 1 || nh < 1 || na < 0 || C < 0 || k < 0) return fail(QCB_ERR_BAD_ARG, "gibbs_host_packed: bad sizes");
  if (W_host == nullptr || b_host == nullptr || c_host == nullptr || (na > 0 && (U_host == nullptr || d_host == nullptr)) ||
      (C > 0 && bits_host == nullptr))
    return fail(QCB_ERR_BAD_ARG, "gibbs_host_packed: NULL argument");
  if (dev_scratch == nullptr || dev_scratch_bytes < qcb_gibbs_host_packed_scratch_bytes(N, nh, na, C))
    return fail(QCB_ERR_BAD_ARG, "gibbs_host_packed: device scratch too small");
  cudaStream_t st = (cudaStream_t)stream;
  HostPipe& hp = g_pipes[device_slot()];
  std::lock_guard<std::mutex> hold(hp.lock);
  int rc = pipe_ready(hp);
  if (rc != QCB_OK) return rc;
  StackedParams p;
  rc = stage_params(p, W_host, U_host, b_host, c_host, d_host, N, nh, na, dev_scratch, st);
  if (rc != QCB_OK) return rc;
  const size_t params = param_bytes(N, nh, na), rowb = (size_t)packed_row_bytes(N);
  uint8_t* dbits = reinterpret_cast<uint8_t*>(static_cast<char*>(dev_scratch) + params);
  void* ws = static_cast<char*>(dev_scratch) + params + align_up((size_t)C * rowb, 256);
  const size_t ws_bytes = dev_scratch_bytes - params - align_up((size_t)C * rowb, 256);
  // slices are whole waves of the persistent grid (no ragged tail per slice): 8 waves each, so that the first copy in
  // and the last copy out (a few tens of microseconds for ~2 MiB) are all that is not hidden behind sampling
  const int64_t wave = gibbs_wave_chains(N, nh + na);
  int64_t per = 8 * wave;
  if ((C + per - 1) / per > MAX_SLICES) per = ((C + MAX_SLICES - 1) / MAX_SLICES + wave - 1) / wave * wave;
  rc = host_pipeline(hp, reinterpret_cast<char*>(bits_host), reinterpret_cast<char*>(dbits), rowb, C, per, draw_initial == 0, st,
                     [&](int64_t c0, int64_t cn) {
                       ChainIO io;
                       io.bits_out = dbits + c0 * rowb;
                       if (draw_initial) io.philox_init = 1; else io.bits_in = io.bits_out;
                       return run_gibbs_io(p, io, cn, k, seed, chain_id0 + (uint64_t)c0, sweep0, ws, ws_bytes, st);
                     });
  if (rc != QCB_OK) return rc;
  QCB_CUDA(cudaStreamSynchronize(st));
  return QCB_OK;
}

}  // extern "C"
