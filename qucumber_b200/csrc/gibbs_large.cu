// Block Gibbs for layers that do not fit the fused kernels (more than 128 units, or both layers above 64): every
// conditional update is one affine + logistic pass over the chains (rows_affine_kernel: the layer's weights in shared memory)
// followed by a Bernoulli pass that follows the library's Philox convention for layers (include/qcb.h: unit u of a layer uses
// call q0 + 4 (u / 32) + (u % 16) / 4, word u % 4, 16-bit half (u / 16) % 2 of stream 0 and the low bits from two more streams)
// or the caller's fed uniforms.  Same draws and same exact FP64 decisions as the fused kernels would make — only slower
// (two launches and two HBM round trips of the state per conditional update, like the reference's own op chain,
// rbm/binary_rbm.py:129-228, rbm/purification_rbm.py:196-337).  Conditionals whose weights do not fit in shared memory run
// through the row-tiled kernel of large.cu (weights in global memory); what remains is that kernel's limit of 767 input units.
#include "rows.cuh"

namespace qcb {

int run_linear(int mode, const double* M, int m_rows, int m_cols, int transpose_m, const double* bias, const double* x,
               const double* M2, int m2_rows, const double* x2, int64_t rows, double* out, cudaStream_t st);

// out[c][u] = (uniform(c, unit0 + u) < p[c][u]); uniforms from Philox (layer convention, first call q0) or from `fed`
// (row stride fed_ld, column offset already applied by the caller)
__global__ void __launch_bounds__(256) bernoulli_layer_kernel(const double* __restrict__ p, int64_t C, int n_units, int unit0, uint32_t q0,
                                                              uint32_t key0, uint32_t key1, uint64_t chain_id0, uint32_t sweep,
                                                              const double* __restrict__ fed, int64_t fed_ld, double* __restrict__ out) {
  const int64_t total = C * n_units;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = idx / n_units;
    const int ul = (int)(idx - c * n_units);
    double uni;
    if (fed != nullptr) {
      uni = fed[c * fed_ld + ul];
    } else {
      const uint32_t u = (uint32_t)(unit0 + ul);
      const uint32_t q = q0 + 4u * (u >> 5) + ((u & 15u) >> 2), w = u & 3u, half = (u >> 4) & 1u;
      const uint64_t id = chain_id0 + (uint64_t)c;
      uint32_t hi[4], la[4], lb[4];
      philox4x32_10((uint32_t)id, (uint32_t)(id >> 32), sweep, (0u << 29) | q, key0, key1, hi);
      philox4x32_10((uint32_t)id, (uint32_t)(id >> 32), sweep, ((1u + 2u * half) << 29) | q, key0, key1, la);
      philox4x32_10((uint32_t)id, (uint32_t)(id >> 32), sweep, ((2u + 2u * half) << 29) | q, key0, key1, lb);
      const uint32_t hw = w == 0 ? hi[0] : (w == 1 ? hi[1] : (w == 2 ? hi[2] : hi[3]));
      const uint32_t aw = w == 0 ? la[0] : (w == 1 ? la[1] : (w == 2 ? la[2] : la[3]));
      const uint32_t bw = w == 0 ? lb[0] : (w == 1 ? lb[1] : (w == 2 ? lb[2] : lb[3]));
      const uint64_t h16 = (hw >> (16u * (1u - half))) & 0xFFFFu;
      uni = (double)((h16 << 37) | ((uint64_t)aw << 5) | (uint64_t)(bw >> 27)) * 0x1.0p-53;
    }
    out[idx] = uni < p[idx] ? 1.0 : 0.0;
  }
}

size_t gibbs_large_workspace_bytes(int N, int nhid, int64_t C) {
  const int widest = N > nhid ? N : nhid;
  // hidden-like states (C x nhid), probabilities (C x widest)
  return align_up((size_t)(C > 0 ? C : 0) * nhid * sizeof(double), 256) + align_up((size_t)(C > 0 ? C : 0) * widest * sizeof(double), 256);
}

static int bern(const double* p, int64_t C, int n_units, int unit0, uint32_t q0, uint64_t seed, uint64_t chain_id0, uint32_t sweep,
                const double* fed, int64_t fed_ld, double* out, cudaStream_t st) {
  const int64_t total = C * n_units;
  const int64_t need = (total + 255) / 256, cap = (int64_t)sm_count() * 8;
  bernoulli_layer_kernel<<<(unsigned)(need < cap ? (need < 1 ? 1 : need) : cap), 256, 0, st>>>(p, C, n_units, unit0, q0, (uint32_t)seed,
                                                                                               (uint32_t)(seed >> 32), chain_id0, sweep, fed,
                                                                                               fed_ld, out);
  QCB_LAUNCH_CHECK();
  return QCB_OK;
}

// v (C, N) is updated in place after a copy from v_in; uniforms (k, C, nh + na + N) or NULL
int run_gibbs_large(const StackedParams& p, const double* v_in, double* v, int64_t C, int k, uint64_t seed, uint64_t chain_id0,
                    uint64_t sweep0, const double* uniforms, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int N = p.N, nh = p.nh, na = p.na, nhid = p.nhid();
  if (ws == nullptr || ws_bytes < gibbs_large_workspace_bytes(N, nhid, C))
    return fail(QCB_ERR_BAD_ARG, "gibbs (large layers): workspace too small (%zu < %zu)", ws_bytes, gibbs_large_workspace_bytes(N, nhid, C));
  if (C == 0) return QCB_OK;
  if (v != v_in) QCB_CUDA(cudaMemcpyAsync(v, v_in, (size_t)C * N * sizeof(double), cudaMemcpyDeviceToDevice, st));
  double* hid = static_cast<double*>(ws);                                       // [h (C x nh) | a (C x na)] as two dense blocks
  double* aux = hid + (size_t)C * nh;
  double* prob = reinterpret_cast<double*>(static_cast<char*>(ws) + align_up((size_t)C * nhid * sizeof(double), 256));
  const uint32_t qv0 = (uint32_t)(4 * ((nhid + 31) / 32));
  const int64_t uld = nhid + N;
  for (int s = 0; s < k; ++s) {
    const uint32_t sweep = (uint32_t)(sweep0 + (uint64_t)s);
    const double* us = uniforms != nullptr ? uniforms + (size_t)s * C * uld : nullptr;
    // h | v
    int rc = run_linear(1, p.W, nh, N, 0, p.c, v, nullptr, 0, nullptr, C, prob, st);
    if (rc != QCB_OK) return rc;
    rc = bern(prob, C, nh, 0, 0u, seed, chain_id0, sweep, us, uld, hid, st);
    if (rc != QCB_OK) return rc;
    if (na > 0) {  // a | v: the auxiliary units continue the hidden-like layer's unit numbering
      rc = run_linear(1, p.U, na, N, 0, p.d, v, nullptr, 0, nullptr, C, prob, st);
      if (rc != QCB_OK) return rc;
      rc = bern(prob, C, na, nh, 0u, seed, chain_id0, sweep, us != nullptr ? us + nh : nullptr, uld, aux, st);
      if (rc != QCB_OK) return rc;
    }
    // v | h, a
    rc = run_linear(1, p.W, nh, N, 1, p.b, hid, na > 0 ? p.U : nullptr, na, na > 0 ? aux : nullptr, C, prob, st);
    if (rc != QCB_OK) return rc;
    rc = bern(prob, C, N, 0, qv0, seed, chain_id0, sweep, us != nullptr ? us + nhid : nullptr, uld, v, st);
    if (rc != QCB_OK) return rc;
  }
  return QCB_OK;
}

}  // namespace qcb
