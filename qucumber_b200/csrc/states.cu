// Compact chain states: conversion between the reference's float64 rows (one double per unit holding 0.0 / 1.0,
// rbm/binary_rbm.py:200-228 works on such tensors) and packed rows of ceil(N / 8) bytes, unit j = bit 7 - j % 8 of byte
// j / 8 (numpy.packbits order).  A chain state carries one bit of information per unit; shipping it as a double makes the
// host <-> device path 64 x larger than it has to be (qcb.h, "Compact states").  Also the in-kernel Bernoulli(1/2) initial
// state of NeuralStateBase.sample (nn_states/neural_state.py:124-131) for the kernels that take float64 rows.
// HBM-bound byte work: coalesced loads, one warp per group of rows, no tensor cores.
#include "common.cuh"

namespace qcb {

// one thread per output byte; the 8 doubles of a byte are one 64-byte segment, so a warp reads 2 KB contiguous per pass
__global__ void __launch_bounds__(256) pack_states_kernel(const double* __restrict__ v, int64_t rows, int N, uint8_t* __restrict__ bits) {
  const int rowbytes = packed_row_bytes(N);
  const int64_t total = rows * rowbytes;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / rowbytes;
    const int bb = (int)(idx - r * rowbytes);
    const double* src = v + r * N + 8 * bb;
    const int n = N - 8 * bb < 8 ? N - 8 * bb : 8;
    uint32_t byte = 0u;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < n && src[i] != 0.0) byte |= 0x80u >> i;
    bits[idx] = (uint8_t)byte;
  }
}

// one thread per double: consecutive threads write consecutive doubles (coalesced), the byte loads hit L1
__global__ void __launch_bounds__(256) unpack_states_kernel(const uint8_t* __restrict__ bits, int64_t rows, int N, double* __restrict__ v) {
  const int rowbytes = packed_row_bytes(N);
  const int64_t total = rows * N;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / N;
    const int j = (int)(idx - r * N);
    v[idx] = ((bits[r * rowbytes + (j >> 3)] >> (7 - (j & 7))) & 1u) ? 1.0 : 0.0;
  }
}

// the same, but only when *flag != 0: queued between the tcgen05 gradient kernel (packed rows) and its FP64 stand-in, which
// needs float64 rows in the rare case that it has to redo the batch (no host synchronisation, capturable)
__global__ void __launch_bounds__(256) unpack_states_if_kernel(const int* __restrict__ flag, const uint8_t* __restrict__ bits, int64_t rows,
                                                               int N, double* __restrict__ v) {
  if (*flag == 0) return;
  const int rowbytes = packed_row_bytes(N);
  const int64_t total = rows * N;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / N;
    const int j = (int)(idx - r * N);
    v[idx] = ((bits[r * rowbytes + (j >> 3)] >> (7 - (j & 7))) & 1u) ? 1.0 : 0.0;
  }
}

// v0 ~ Bernoulli(1/2): unit j of chain c = bit j % 32 of word j / 32 of Philox call (chain id, sweep0, stream 5, call 0)
__global__ void __launch_bounds__(256) philox_init_states_kernel(double* __restrict__ v, int64_t C, int N, uint32_t key0, uint32_t key1,
                                                                 uint64_t chain_id0, uint32_t sweep0) {
  const int64_t total = C * 4;  // one thread per (chain, word)
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = idx >> 2;
    const int w = (int)(idx & 3);
    if (32 * w >= N) continue;
    const uint64_t id = chain_id0 + (uint64_t)c;
    uint32_t rnd[4];
    philox4x32_10((uint32_t)id, (uint32_t)(id >> 32), sweep0, (PHILOX_STREAM_V0 << 29), key0, key1, rnd);
    const uint32_t word = w == 0 ? rnd[0] : (w == 1 ? rnd[1] : (w == 2 ? rnd[2] : rnd[3]));
    const int n = N - 32 * w < 32 ? N - 32 * w : 32;
    for (int i = 0; i < n; ++i) v[c * N + 32 * w + i] = ((word >> i) & 1u) ? 1.0 : 0.0;
  }
}

static int grid_for(int64_t work_items) {
  const int64_t need = (work_items + 255) / 256, cap = (int64_t)sm_count() * 8;
  return (int)(need < 1 ? 1 : (need < cap ? need : cap));
}

int run_pack_states(const double* v, int64_t rows, int N, uint8_t* bits, cudaStream_t st) {
  if (rows == 0) return QCB_OK;
  pack_states_kernel<<<grid_for(rows * packed_row_bytes(N)), 256, 0, st>>>(v, rows, N, bits);
  QCB_LAUNCH_CHECK();
  return QCB_OK;
}

int run_unpack_states(const uint8_t* bits, int64_t rows, int N, double* v, cudaStream_t st) {
  if (rows == 0) return QCB_OK;
  unpack_states_kernel<<<grid_for(rows * N), 256, 0, st>>>(bits, rows, N, v);
  QCB_LAUNCH_CHECK();
  return QCB_OK;
}

int run_unpack_states_if(const int* flag, const uint8_t* bits, int64_t rows, int N, double* v, cudaStream_t st) {
  if (rows == 0) return QCB_OK;
  unpack_states_if_kernel<<<grid_for(rows * N), 256, 0, st>>>(flag, bits, rows, N, v);
  QCB_LAUNCH_CHECK();
  return QCB_OK;
}

int run_philox_init_states(double* v, int64_t C, int N, uint64_t seed, uint64_t chain_id0, uint64_t sweep0, cudaStream_t st) {
  if (C == 0) return QCB_OK;
  if (N > 128) return fail(QCB_ERR_UNSUPPORTED, "in-kernel initial states cover at most 128 visible units (one Philox call)");
  philox_init_states_kernel<<<grid_for(C * 4), 256, 0, st>>>(v, C, N, (uint32_t)seed, (uint32_t)(seed >> 32), chain_id0, (uint32_t)sweep0);
  QCB_LAUNCH_CHECK();
  return QCB_OK;
}

}  // namespace qcb
