// Block Gibbs for SMALL layers (N, nh + na <= 16) without tensor cores: one THREAD per chain, pre-activations from float
// nibble tables in shared memory.  BinaryRBM.gibbs_steps (rbm/binary_rbm.py:200-228) / PurificationRBM.gibbs_steps
// (rbm/purification_rbm.py:309-337) at the sizes of the tutorials (Tutorial 4: 10^6 chains, N = nh = 10, k = 100).
//
// The tensor-core filter kernels pad a 10-unit layer to 16 or 32 units and pay a join of several warps per conditional update;
// at this size the contraction is 100 multiply-adds per update, i.e. nothing.  Here
//   x_i = sum_t T[layer][i][t][ nibble t of the input bits ],   T[..][m] = (bias of i if t = 0) + sum of i's weights towards
//   inputs 4 t .. 4 t + 3 selected by the bits of m
// is ceil(n_in / 4) table reads and adds per output unit, the chain's bits never leave the thread's registers, and nothing is
// exchanged between threads.  The tables are replicated (up to 32 times, entry-major) so that the lanes of a warp, which all
// read different entries, hit different banks.
// Same Philox convention (unit u of a layer: call q0 + u / 4, word u % 4, high 16-bit half; include/qcb.h), same decision
// scheme as gibbs_imma.cu / gibbs_umma.cu — float filter against the 16-bit cell with a rigorous margin
// ((n_in + 6) 2^-24 (|bias| + sum |w|) for the float table sums, a quarter of it plus 2^-19 for the filter itself), exact FP64
// comparison with the 53-bit uniform for what the filter leaves open — hence the same samples bit for bit.
#include "common.cuh"

namespace qcb {

constexpr int GT_THREADS = 1024;
constexpr int GT_MAX_UNITS = 16;
constexpr size_t GT_TABLE_BUDGET = 160 * 1024;   // shared memory for the replicated tables

struct GtHeader {
  float margin_h, margin_v;   // filter margins of the two layers (the largest over the layer's units)
  int pad[2];
};
// image: header (16 B) | Th [nhid][QV][16] floats | Tv [N][QH][16] floats      (QV = ceil(N / 4), QH = ceil(nhid / 4))
__host__ __device__ inline int gt_entries(int N, int nhid) { return (nhid * ((N + 3) / 4) + N * ((nhid + 3) / 4)) * 16; }
__host__ __device__ inline size_t gt_image_bytes(int N, int nhid) { return 16 + (size_t)gt_entries(N, nhid) * sizeof(float); }

struct GtArgs {
  const unsigned char* image;
  StackedParams p;
  const double* v_in;
  double* v;
  const uint8_t* bits_in;
  uint8_t* bits_out;
  int in_mode;            // 0: v_in, 1: bits_in, 2: Philox initial state
  int replicas;           // copies of the tables in shared memory (power of two, <= 32)
  int64_t C;
  int k;
  uint32_t key0, key1;
  uint64_t chain_id0;
  uint32_t sweep0;
  uint32_t rk[20];        // Philox round keys
};

// one CTA: tables and margins
__global__ void __launch_bounds__(256) gibbs_pack_table_kernel(StackedParams p, unsigned char* __restrict__ img) {
  __shared__ double red[2][8];
  const int N = p.N, H = p.nhid();
  const int QV = (N + 3) / 4, QH = (H + 3) / 4;
  float* Th = reinterpret_cast<float*>(img + 16);
  float* Tv = Th + (size_t)H * QV * 16;
  for (int e = threadIdx.x; e < H * QV * 16; e += blockDim.x) {
    const int m = e & 15, t = (e >> 4) % QV, i = (e >> 4) / QV;
    float v = t == 0 ? (float)p.hbias(i) : 0.0f;
#pragma unroll
    for (int b = 0; b < 4; ++b)
      if (((m >> b) & 1) && 4 * t + b < N) v += (float)p.w(i, 4 * t + b);
    Th[e] = v;
  }
  for (int e = threadIdx.x; e < N * QH * 16; e += blockDim.x) {
    const int m = e & 15, t = (e >> 4) % QH, j = (e >> 4) / QH;
    float v = t == 0 ? (float)p.b[j] : 0.0f;
#pragma unroll
    for (int b = 0; b < 4; ++b)
      if (((m >> b) & 1) && 4 * t + b < H) v += (float)p.w(4 * t + b, j);
    Tv[e] = v;
  }
  // margins: |x~ - x| <= (n_in + 6) 2^-24 (|bias| + sum |w|) (conversions and every float addition of the table sums);
  // sigmoid is 1/4-Lipschitz; the float evaluation of the test (ex2.approx, two FMAs) is good to 2^-21.5
  double sh = 0.0, sv = 0.0;
  if (threadIdx.x < H) {
    sh = fabs(p.hbias(threadIdx.x));
    for (int j = 0; j < N; ++j) sh += fabs(p.w(threadIdx.x, j));
  } else if (threadIdx.x >= 32 && threadIdx.x < 32 + N) {
    const int j = threadIdx.x - 32;
    sv = fabs(p.b[j]);
    for (int i = 0; i < H; ++i) sv += fabs(p.w(i, j));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sh = fmax(sh, __shfl_xor_sync(0xffffffffu, sh, o));
    sv = fmax(sv, __shfl_xor_sync(0xffffffffu, sv, o));
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sh; red[1][threadIdx.x >> 5] = sv; }
  __syncthreads();
  if (threadIdx.x == 0) {
    sh = sv = 0.0;
    for (int w = 0; w < 8; ++w) { sh = fmax(sh, red[0][w]); sv = fmax(sv, red[1][w]); }
    const double mh = (0.25 * (double)(N + 6) * 0x1.0p-24 * sh + 0x1.0p-19) * 1.001, mv = (0.25 * (double)(H + 6) * 0x1.0p-24 * sv + 0x1.0p-19) * 1.001;
    GtHeader* hd = reinterpret_cast<GtHeader*>(img);
    hd->margin_h = mh < 0.25 ? (float)mh : 4.0f;   // non-finite or huge weights: every unit goes to the exact path
    hd->margin_v = mv < 0.25 ? (float)mv : 4.0f;
    hd->pad[0] = hd->pad[1] = 0;
  }
}

__device__ __forceinline__ void gt_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const uint32_t (&rk)[20], uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ rk[2 * r];
    c1 = lo1;
    c2 = hi0 ^ c3 ^ rk[2 * r + 1];
    c3 = lo0;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// exact decision for unit `unit` of the layer (hidden_out: h | v) given the input bits: FP64 pre-activation from the original
// parameters, 53-bit uniform from the three Philox streams, exact sigmoid
__device__ __noinline__ bool gt_decide_exact(const GtArgs* a, int hidden_out, uint32_t in_bits, int unit, uint32_t cid0, uint32_t cid1,
                                             uint32_t sweep, uint32_t q) {
  const StackedParams& p = a->p;
  const int n_in = hidden_out ? p.N : p.nhid();
  double x = hidden_out ? p.hbias(unit) : p.b[unit];
  for (int j = 0; j < n_in; ++j)
    if ((in_bits >> j) & 1u) x += hidden_out ? p.w(unit, j) : p.w(j, unit);
  const uint32_t w = (uint32_t)(unit & 3);
  uint32_t hi[4], la[4], lb[4];
  philox4x32_10(cid0, cid1, sweep, (0u << 29) | q, a->key0, a->key1, hi);
  philox4x32_10(cid0, cid1, sweep, (1u << 29) | q, a->key0, a->key1, la);   // units below 16: the high half (half = 0)
  philox4x32_10(cid0, cid1, sweep, (2u << 29) | q, a->key0, a->key1, lb);
  const uint32_t hw = w == 0 ? hi[0] : (w == 1 ? hi[1] : (w == 2 ? hi[2] : hi[3]));
  const uint32_t aw = w == 0 ? la[0] : (w == 1 ? la[1] : (w == 2 ? la[2] : la[3]));
  const uint32_t bw = w == 0 ? lb[0] : (w == 1 ? lb[1] : (w == 2 ? lb[2] : lb[3]));
  const uint64_t ubits = ((uint64_t)(hw >> 16) << 37) | ((uint64_t)aw << 5) | (uint64_t)(bw >> 27);
  return (double)ubits * 0x1.0p-53 < sigmoid_exact(x);
}

// QI: nibbles of the input layer, QO: Philox calls (groups of four units) of the output layer, RS: log2 of the table replicas.
// `tab`: the layer's replicated tables, entry (i, t, m) of replica r at float index (((i QI + t) 16 + m) << RS) + r.
template <int QI, int QO, int RS>
__device__ __forceinline__ uint32_t gt_update(const GtArgs& a, const float* __restrict__ tab, int lane_r, int units, int hidden_out,
                                              uint32_t in_bits, float c1, float d01, uint32_t cid0, uint32_t cid1, uint32_t sweep, uint32_t q0) {
  // this thread's read position inside the 16-entry table of nibble t: a register per nibble, everything else is an
  // immediate offset of the load
  const float* pos[QI];
#pragma unroll
  for (int t = 0; t < QI; ++t) pos[t] = tab + ((((in_bits >> (4 * t)) & 15u) << RS) + (uint32_t)lane_r);
  uint32_t acc1 = 0u, acc0 = 0u;   // sign bits of the two tests, units visited from the highest down
#pragma unroll
  for (int c = QO - 1; c >= 0; --c) {
    uint32_t rnd[4];
    gt_philox(cid0, cid1, sweep, q0 + (uint32_t)c, a.rk, rnd);
#pragma unroll
    for (int w = 3; w >= 0; --w) {
      const int i = 4 * c + w;
      if (i < units) {   // only the last call of a layer can run past it: uniform over the grid
        float x = pos[0][(size_t)((i * QI) * 16) << RS];
#pragma unroll
        for (int t = 1; t < QI; ++t) x += pos[t][(size_t)((i * QI + t) * 16) << RS];
        float e;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
        const float q = 1.0f + e;
        const float m = __uint_as_float(__byte_perm(rnd[w], 0x7Fu, 0x5432u) << 7);   // 1 + cell 2^-16
        const float s1 = fmaf(m + c1, q, -1.0f);       // (u_f + 2^-16 + margin) q - 1 < 0  => certainly u < p
        const float s0 = fmaf(d01, q, -s1);            // 1 - (u_f - margin) q < 0          => certainly u >= p
        acc1 = __funnelshift_l(__float_as_uint(s1), acc1, 1);
        acc0 = __funnelshift_l(__float_as_uint(s0), acc0, 1);
      } else {
        acc1 <<= 1;
        acc0 = (acc0 << 1) | 1u;
      }
    }
  }
  // a pre-activation beyond the float exponent range makes q infinite and s0 a NaN: neither sign bit is set and the unit
  // takes the exact path like any other undecided one
  uint32_t out = acc1;
  uint32_t open = ~(acc0 | acc1) & ((units >= 32) ? 0xffffffffu : ((1u << units) - 1u));
  while (open != 0u) {   // rare (about 3e-5 per unit): thread-level exact path
    const int i = __ffs(open) - 1;
    open &= open - 1u;
    if (gt_decide_exact(&a, hidden_out, in_bits, i, cid0, cid1, sweep, q0 + (uint32_t)(i >> 2))) out |= 1u << i;
  }
  return out;
}

template <int QV, int QH, int RS>
__global__ void __launch_bounds__(GT_THREADS, 1) gibbs_table_kernel(const GtArgs a) {
  extern __shared__ __align__(16) float gt_smem[];
  const int N = a.p.N, H = a.p.nhid();
  constexpr int R = 1 << RS;
  const GtHeader* hd = reinterpret_cast<const GtHeader*>(a.image);
  const float* img = reinterpret_cast<const float*>(a.image + 16);
  const int entries_h = H * QV * 16, entries = entries_h + N * QH * 16;
  for (int idx = threadIdx.x; idx < entries * R; idx += GT_THREADS) gt_smem[idx] = img[idx >> RS];
  __syncthreads();
  const float* tab_h = gt_smem;
  const float* tab_v = gt_smem + ((size_t)entries_h << RS);
  const int lane_r = threadIdx.x & (R - 1);
  const float mh = hd->margin_h, mv = hd->margin_v;
  const float c1h = (0x1.0p-16f + mh) - 1.0f, d01h = 0x1.0p-16f + 2.0f * mh;
  const float c1v = (0x1.0p-16f + mv) - 1.0f, d01v = 0x1.0p-16f + 2.0f * mv;
  const uint32_t qv0 = (uint32_t)(4 * ((H + 31) / 32));
  const uint32_t vmask = N >= 32 ? 0xffffffffu : ((1u << N) - 1u);
  const int rowbytes = packed_row_bytes(N);

  for (int64_t ch = (int64_t)blockIdx.x * GT_THREADS + threadIdx.x; ch < a.C; ch += (int64_t)gridDim.x * GT_THREADS) {
    const uint64_t id = a.chain_id0 + (uint64_t)ch;
    const uint32_t cid0 = (uint32_t)id, cid1 = (uint32_t)(id >> 32);
    uint32_t v = 0u;
    if (a.in_mode == 0) {
      const unsigned long long* rp = reinterpret_cast<const unsigned long long*>(a.v_in) + ch * N;
      for (int j = 0; j < N; ++j) v |= ((rp[j] << 1) != 0ull ? 1u : 0u) << j;
    } else if (a.in_mode == 1) {
      const uint8_t* rp = a.bits_in + ch * rowbytes;
      uint32_t x = 0u;
      for (int b = 0; b < rowbytes; ++b) x |= (uint32_t)rp[b] << (8 * b);
      v = __byte_perm(__brev(x), 0u, 0x0123u) & vmask;   // unit 8 b + i = bit 7 - i of byte b  ->  bit 8 b + i
    } else {
      uint32_t rnd[4];
      philox4x32_10(cid0, cid1, a.sweep0, PHILOX_STREAM_V0 << 29, a.key0, a.key1, rnd);
      v = rnd[0] & vmask;
    }
#pragma unroll 1
    for (int s = 0; s < a.k; ++s) {
      const uint32_t sweep = a.sweep0 + (uint32_t)s;
      const uint32_t h = gt_update<QV, QH, RS>(a, tab_h, lane_r, H, 1, v, c1h, d01h, cid0, cid1, sweep, 0u);
      v = gt_update<QH, QV, RS>(a, tab_v, lane_r, N, 0, h, c1v, d01v, cid0, cid1, sweep, qv0);
    }
    if (a.v != nullptr) {
      double* rp = a.v + ch * N;
      for (int j = 0; j < N; ++j) rp[j] = (v >> j) & 1u ? 1.0 : 0.0;
    }
    if (a.bits_out != nullptr) {
      const uint32_t x = __byte_perm(__brev(v), 0u, 0x0123u);
      uint8_t* rp = a.bits_out + ch * rowbytes;
      for (int b = 0; b < rowbytes; ++b) rp[b] = (uint8_t)(x >> (8 * b));
    }
  }
}

size_t gibbs_table_workspace_bytes(int N, int nhid) {
  if (N < 1 || nhid < 1 || N > GT_MAX_UNITS || nhid > GT_MAX_UNITS) return 0;
  return align_up(gt_image_bytes(N, nhid), 256);
}

int64_t gibbs_table_wave_chains() { return (int64_t)sm_count() * GT_THREADS; }

template <int QV, int QH, int RS>
static int launch_gibbs_table_rs(const GtArgs& a, size_t smem, int grid, cudaStream_t st) {
  static bool configured[QCB_MAX_DEVICES] = {false};
  bool& done = configured[device_slot()];
  if (!done) {
    QCB_CUDA((cudaFuncSetAttribute(gibbs_table_kernel<QV, QH, RS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(GT_TABLE_BUDGET + 1024))));
    done = true;
  }
  gibbs_table_kernel<QV, QH, RS><<<grid, GT_THREADS, smem, st>>>(a);
  QCB_LAUNCH_CHECK();
  return QCB_OK;
}
template <int QV, int QH>
static int launch_gibbs_table(const GtArgs& a, size_t smem, int grid, cudaStream_t st) {
  // 32 replicas wherever they fit the budget (up to 12 x 12 units and every shape with one layer of at most 8), else 16
  return a.replicas == 32 ? launch_gibbs_table_rs<QV, QH, 5>(a, smem, grid, st) : launch_gibbs_table_rs<QV, QH, 4>(a, smem, grid, st);
}

// *handled = 1 when this kernel took the call (both layers of at most 16 units, own Philox draws)
int try_gibbs_table(const StackedParams& p, const ChainIO& io, int64_t C, int k, uint64_t seed, uint64_t chain_id0, uint64_t sweep0,
                    void* ws, size_t ws_bytes, int* handled, cudaStream_t st) {
  *handled = 0;
  const int N = p.N, H = p.nhid();
  const size_t need = gibbs_table_workspace_bytes(N, H);
  if (need == 0 || ws == nullptr || ws_bytes < need || k < 1) return QCB_OK;
  unsigned char* img = reinterpret_cast<unsigned char*>(ws);
  gibbs_pack_table_kernel<<<1, 256, 0, st>>>(p, img);
  QCB_LAUNCH_CHECK();
  GtArgs a;
  a.image = img;
  a.p = p;
  a.v_in = io.v_in;
  a.v = io.v_out;
  a.bits_in = io.bits_in;
  a.bits_out = io.bits_out;
  a.in_mode = io.philox_init ? 2 : (io.bits_in != nullptr ? 1 : 0);
  const int R = (size_t)gt_entries(N, H) * 32 * sizeof(float) <= GT_TABLE_BUDGET ? 32 : 16;   // 16 x 16 units: 2 048 entries, 128 KB at 16 replicas
  a.replicas = R;
  a.C = C;
  a.k = k;
  a.key0 = (uint32_t)seed;
  a.key1 = (uint32_t)(seed >> 32);
  a.chain_id0 = chain_id0;
  a.sweep0 = (uint32_t)sweep0;
  for (int r = 0; r < 10; ++r) {
    a.rk[2 * r] = a.key0 + (uint32_t)r * 0x9E3779B9u;
    a.rk[2 * r + 1] = a.key1 + (uint32_t)r * 0xBB67AE85u;
  }
  const size_t smem = (size_t)gt_entries(N, H) * R * sizeof(float);
  const int64_t want = (C + GT_THREADS - 1) / GT_THREADS, cap = (int64_t)sm_count();
  const int grid = (int)(want < cap ? want : cap);
  const int QV = (N + 3) / 4, QH = (H + 3) / 4;
  int rc = QCB_OK;
#define QCB_GT_CASE(V, Hq) \
  if (QV == V && QH == Hq) rc = launch_gibbs_table<V, Hq>(a, smem, grid, st); else
  QCB_GT_CASE(1, 1) QCB_GT_CASE(1, 2) QCB_GT_CASE(1, 3) QCB_GT_CASE(1, 4) QCB_GT_CASE(2, 1) QCB_GT_CASE(2, 2) QCB_GT_CASE(2, 3) QCB_GT_CASE(2, 4)
  QCB_GT_CASE(3, 1) QCB_GT_CASE(3, 2) QCB_GT_CASE(3, 3) QCB_GT_CASE(3, 4) QCB_GT_CASE(4, 1) QCB_GT_CASE(4, 2) QCB_GT_CASE(4, 3) QCB_GT_CASE(4, 4)
  return QCB_OK;
#undef QCB_GT_CASE
  if (rc == QCB_OK) *handled = 1;
  return rc;
}

}  // namespace qcb
