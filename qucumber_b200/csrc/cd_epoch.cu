// One launch for a whole contrastive-divergence epoch of a small BinaryRBM (N, nh <= 64, plain SGD).
//
// Replaces the batch loop of NeuralStateBase.fit (qucumber/nn_states/neural_state.py:611-627): for every
// batch  g = <grad E>_data / B - <grad E>_{v_k} / C  (compute_batch_gradients, :417-447, with
// effective_energy_gradient of rbm/binary_rbm.py:98-120 and gibbs_steps :200-228) followed by
// p -= lr * g  (torch.optim.SGD without momentum).  At the tutorial size (N = nh = 10, 100 + 200 rows per
// batch, k = 10) a step is a few microseconds of arithmetic, so the per-step path is bound by its seven
// launches; here the batch loop runs inside one kernel:
//
//  * one thread-block cluster of 16 CTAs (8 where 16 cannot be co-scheduled); every CTA keeps its own FP64 copy of W, b, c in shared memory;
//  * rows and chains of a batch are dealt to (CTA, warp, slot); a slot is SW = 8/16/32 lanes, lane u of a
//    slot is hidden unit u in the h | v update and visible unit u in the v | h update, states are bit
//    masks exchanged by ballot;
//  * the Gibbs draws follow the library's Philox convention (include/qcb.h) with chain id = row of the
//    negative batch and sweep = sweep0 + step * k + s, i.e. the samples are those of qcb_rbm_cd_step;
//    a decision is taken by a float filter against the 16-bit cell of u (margin 2^-17, |x| <= 64); what the
//    filter leaves open goes to the exact path: FP64 sigmoid against the cell, and the remaining 37 bits of u
//    only when p falls inside the cell — together exactly  u < p;
//  * per step each CTA reduces its gradient share in a fixed order, the shares are exchanged through
//    distributed shared memory (one cluster barrier per step, double-buffered), and every CTA applies the
//    identical update to its copy.
#include <cooperative_groups.h>

#include <cstdlib>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace qcb {

constexpr int CE_CLUSTER_PORTABLE = 8, CE_CLUSTER_MAX = 16;

struct CdEpochArgs {
  double* W;
  double* b;
  double* c;
  int N, nh;
  const double* pos;
  int64_t B_total;
  int pos_batch;
  const double* neg;
  int64_t C_total;
  int neg_batch;
  int num_batches;
  int k;
  double lr;
  uint32_t key0, key1;
  uint32_t sweep0;
  double* g_last;  // optional (nh*N + N + nh): gradient of the last batch
};

__host__ __device__ constexpr int ce_threads(int SW) { return SW == 32 ? 512 : 256; }

// shared memory: W (nh x LD), b, c, the CTA's gradient share (2 x G, read by the other CTAs), per-warp shares (warps x G)
__host__ __device__ inline size_t ce_smem_doubles(int N, int nh, int SW, int warps) {
  const int LD = N | 1, G = nh * N + N + nh;
  return (size_t)SW * LD + N + nh + 2 * (size_t)G + (size_t)warps * (G + nh);  // per-warp shares are padded (GP = G + nh at most)
}

// The 16-bit cell of the unit's uniform for (chain, sweep); q0 = first Philox call of the layer.
__device__ __forceinline__ uint32_t ce_cell(int u, uint32_t q0, uint32_t cid0, uint32_t cid1, uint32_t sweep, uint32_t key0, uint32_t key1) {
  const uint32_t q = q0 + 4u * (uint32_t)(u >> 5) + (uint32_t)((u & 15) >> 2);
  const int word = u & 3;
  uint32_t r[4];
  philox4x32_10(cid0, cid1, sweep, q, key0, key1, r);
  const uint32_t w32 = word == 0 ? r[0] : (word == 1 ? r[1] : (word == 2 ? r[2] : r[3]));
  return ((u >> 4) & 1) ? (w32 & 0xFFFFu) : (w32 >> 16);
}

// Branch-free float filter: 1 / 0 when  u < sigmoid(x)  is certainly true / false, 2 when undecided.
// |p~ - sigmoid(x)| <= 0.25 (3 |x| 2^-24 + 2^-22) + 2^-22 < 2^-18 for |x| <= 64 (conversion of x, the two roundings of
// the exponent argument, ex2.approx, 1 + e, rcp.approx); tested with margin 2^-17.
__device__ __forceinline__ int ce_filter(double x, uint32_t h16) {
  const float xf = (float)x;
  float e, pf;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(xf * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(pf) : "f"(1.0f + e));
  const float lof = (float)h16 * 0x1.0p-16f;
  const bool sure1 = pf - 0x1.0p-17f >= lof + 0x1.0p-16f, sure0 = pf + 0x1.0p-17f <= lof;
  const int verdict = sure1 ? 1 : (sure0 ? 0 : 2);
  return fabsf(xf) <= 64.0f ? verdict : 2;
}

// The same filter for a pre-activation that was itself accumulated in float: `margin` = a quarter of the bound on
// |x~ - x| plus the 2^-17 of the float sigmoid.
__device__ __forceinline__ int ce_filter_f(float xf, uint32_t h16, float margin) {
  float e, pf;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(xf * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(pf) : "f"(1.0f + e));
  const float lof = (float)h16 * 0x1.0p-16f;
  const bool sure1 = pf - margin >= lof + 0x1.0p-16f, sure0 = pf + margin <= lof;
  const int verdict = sure1 ? 1 : (sure0 ? 0 : 2);
  return fabsf(xf) <= 64.0f ? verdict : 2;
}

// Cold path: the exact FP64 comparison  u < sigmoid(x)  with the full 53-bit uniform.
__device__ __noinline__ bool ce_decide_exact(double x, uint32_t h16, int u, uint32_t q0, uint32_t cid0, uint32_t cid1, uint32_t sweep,
                                             uint32_t key0, uint32_t key1) {
  const uint32_t q = q0 + 4u * (uint32_t)(u >> 5) + (uint32_t)((u & 15) >> 2);
  const int word = u & 3;
  const uint32_t half = (uint32_t)(u >> 4) & 1u;
  const double p = sigmoid_exact(x);
  const double lo = (double)h16 * 0x1.0p-16;
  if (p >= lo + 0x1.0p-16) return true;   // u < lo + 2^-16 <= p
  if (p <= lo) return false;              // u >= lo >= p
  uint32_t la[4], lb[4];
  philox4x32_10(cid0, cid1, sweep, ((1u + 2u * half) << 29) | q, key0, key1, la);
  philox4x32_10(cid0, cid1, sweep, ((2u + 2u * half) << 29) | q, key0, key1, lb);
  const uint32_t aw = word == 0 ? la[0] : (word == 1 ? la[1] : (word == 2 ? la[2] : la[3]));
  const uint32_t bw = word == 0 ? lb[0] : (word == 1 ? lb[1] : (word == 2 ? lb[2] : lb[3]));
  const uint64_t bits = ((uint64_t)h16 << 37) | ((uint64_t)aw << 5) | (uint64_t)(bw >> 27);
  return (double)bits * 0x1.0p-53 < p;
}

template <int SW>
__global__ void __launch_bounds__(ce_threads(SW), 1) cd_epoch_kernel(const CdEpochArgs a) {
  extern __shared__ __align__(16) double ce_smem[];
  constexpr int NT = ce_threads(SW), NW = NT / 32, CPW = 32 / SW;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int N = a.N, nh = a.nh, LD = N | 1, G = nh * N + N + nh, GP = nh * LD + N + nh;
  double* Wsm = ce_smem;
  double* bsm = Wsm + (size_t)SW * LD;  // W region padded to SW rows
  double* csm = bsm + N;
  double* partial = csm + nh;
  double* wpart = partial + 2 * (size_t)G;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int u = lane % SW, sub = lane / SW;
  const uint32_t slot_mask = SW == 32 ? 0xffffffffu : ((1u << SW) - 1u);
  const int nranks = (int)cluster.num_blocks();
  const int slot = (rank * NW + warp) * CPW + sub, nslots = nranks * NW * CPW;

  __shared__ double e2tab[EXP2_TABLE];
  fill_exp2_table(e2tab);
  for (int idx = tid; idx < nh * N; idx += NT) Wsm[(idx / N) * LD + idx % N] = a.W[idx];
  for (int idx = tid; idx < N; idx += NT) bsm[idx] = a.b[idx];
  for (int idx = tid; idx < nh; idx += NT) csm[idx] = a.c[idx];
  __syncthreads();

  // theta_u = c_u + sum_j v_j W_uj  for hidden unit u of this lane.  Branch-free: lanes past the layer compute on a
  // clamped row, and columns j >= N are loaded (the W region is padded to SW rows, so this stays in bounds) but never
  // selected because mask bits >= N are zero.
  const int uh = u < nh ? u : nh - 1, uv = u < N ? u : N - 1;
  auto hidden_pre = [&](uint32_t vm) {
    const double* row = Wsm + uh * LD;
    double x[4] = {csm[uh], 0.0, 0.0, 0.0};
#pragma unroll
    for (int j = 0; j < SW; ++j)
      if ((vm >> j) & 1u) x[j & 3] += row[j];
    return (x[0] + x[1]) + (x[2] + x[3]);
  };
  auto visible_pre = [&](uint32_t hm) {
    const double* col = Wsm + uv;
    double x[4] = {bsm[uv], 0.0, 0.0, 0.0};
#pragma unroll
    for (int i = 0; i < SW; ++i)
      if ((hm >> i) & 1u) x[i & 3] += col[i * LD];
    return (x[0] + x[1]) + (x[2] + x[3]);
  };
  auto slot_bits = [&](bool bit) { return (__ballot_sync(0xffffffffu, bit) >> (sub * SW)) & slot_mask; };

  // the slot's first row of each phase is fetched one step ahead (the batches are known in advance)
  auto first_row = [&](const double* rows_all, int64_t total, int batch, int step) {
    const int64_t start = (int64_t)step * batch;
    const int64_t rows = total - start < batch ? total - start : batch;
    return (step < a.num_batches && slot < rows && u < N) ? rows_all[(start + slot) * N + u] : 0.0;
  };
  double next_pos = first_row(a.pos, a.B_total, a.pos_batch, 0), next_neg = first_row(a.neg, a.C_total, a.neg_batch, 0);

  for (int step = 0; step < a.num_batches; ++step) {
    const int64_t p0 = (int64_t)step * a.pos_batch, c0 = (int64_t)step * a.neg_batch;
    const int Bs = (int)(a.B_total - p0 < a.pos_batch ? a.B_total - p0 : a.pos_batch);
    const int Cs = (int)(a.C_total - c0 < a.neg_batch ? a.C_total - c0 : a.neg_batch);
    const double* pos_rows = a.pos + p0 * N;
    const double* neg_rows = a.neg + c0 * N;
    const double first_pos = next_pos, first_neg = next_neg;
    next_pos = first_row(a.pos, a.B_total, a.pos_batch, step + 1);
    next_neg = first_row(a.neg, a.C_total, a.neg_batch, step + 1);
    double accW[SW];
#pragma unroll
    for (int j = 0; j < SW; ++j) accW[j] = 0.0;
    double accb = 0.0, accc = 0.0;
    // g = (1/B) sum_data grad E - (1/C) sum_chains grad E,  grad E(v) = -[p v^T ; v ; p]
    auto accumulate = [&](uint32_t vm, bool valid, double weight) {
      const double sg = sigmoid_tab(hidden_pre(vm), e2tab);
      const double pw = (valid && u < nh) ? weight * sg : 0.0;
      accc += pw;
#pragma unroll
      for (int j = 0; j < SW; ++j)
        if ((vm >> j) & 1u) accW[j] += pw;
      if (valid && ((vm >> u) & 1u)) accb += weight;
    };

    // positive phase
    for (int r0 = 0; r0 < Bs; r0 += nslots) {
      const int r = r0 + slot;
      const bool valid = r < Bs;
      const double raw = r0 == 0 ? first_pos : ((valid && u < N) ? pos_rows[(int64_t)r * N + u] : 0.0);
      const uint32_t vm = slot_bits(raw != 0.0);
      accumulate(vm, valid, -1.0 / (double)Bs);
    }
    // negative phase: k sweeps from the rows of the negative batch, then the gradient at v_k
    for (int ch0 = 0; ch0 < Cs; ch0 += nslots) {
      const int ch = ch0 + slot;
      const bool valid = ch < Cs;
      const double raw = ch0 == 0 ? first_neg : ((valid && u < N) ? neg_rows[(int64_t)ch * N + u] : 0.0);
      uint32_t vm = slot_bits(raw != 0.0);
      const uint32_t cid0 = (uint32_t)ch, cid1 = 0u;
      const bool warp_has_chain = __any_sync(0xffffffffu, valid);
      for (int s = 0; s < (warp_has_chain ? a.k : 0); ++s) {
        const uint32_t sweep = a.sweep0 + (uint32_t)(step * a.k + s);
        const uint32_t qv0 = 4u * (uint32_t)((nh + 31) / 32);
        const uint32_t cell_h = ce_cell(u, 0u, cid0, cid1, sweep, a.key0, a.key1);
        const uint32_t cell_v = ce_cell(u, qv0, cid0, cid1, sweep, a.key0, a.key1);  // independent of the state: overlaps the h | v update
        const double xh = hidden_pre(vm);
        int verdict = ce_filter(xh, cell_h);
        verdict = u < nh ? verdict : 0;
        if (__any_sync(0xffffffffu, verdict == 2)) {
          if (verdict == 2) verdict = ce_decide_exact(xh, cell_h, u, 0u, cid0, cid1, sweep, a.key0, a.key1) ? 1 : 0;
          __syncwarp();
        }
        const uint32_t hm = slot_bits(verdict == 1);
        const double xv = visible_pre(hm);
        verdict = ce_filter(xv, cell_v);
        verdict = u < N ? verdict : 0;
        if (__any_sync(0xffffffffu, verdict == 2)) {
          if (verdict == 2) verdict = ce_decide_exact(xv, cell_v, u, qv0, cid0, cid1, sweep, a.key0, a.key1) ? 1 : 0;
          __syncwarp();
        }
        vm = slot_bits(verdict == 1);
      }
      accumulate(vm, valid, 1.0 / (double)Cs);
    }

    // the slots of a warp, then the warps of the CTA, in a fixed order
#pragma unroll
    for (int off = SW; off < 32; off <<= 1) {
#pragma unroll
      for (int j = 0; j < SW; ++j) accW[j] += __shfl_xor_sync(0xffffffffu, accW[j], off);
      accb += __shfl_xor_sync(0xffffffffu, accb, off);
      accc += __shfl_xor_sync(0xffffffffu, accc, off);
    }
    // per-warp shares: rows of the W part padded to LD (odd) so that the 32 lanes of a warp hit 32 banks
    double* mine = wpart + (size_t)warp * GP;
    if (sub == 0) {
      if (u < nh) {
#pragma unroll
        for (int j = 0; j < SW; ++j)
          if (j < N) mine[u * LD + j] = accW[j];
        mine[nh * LD + N + u] = accc;
      }
      if (u < N) mine[nh * LD + u] = accb;
    }
    __syncthreads();
    double* my_partial = partial;       // this CTA's share of the step's gradient (read by the slice owners)
    double* my_red = partial + G;       // the slice of the cluster sum this CTA owns (read by everybody)
    for (int idx = tid; idx < G; idx += NT) {
      const int off = idx < nh * N ? (idx / N) * LD + idx % N : idx - nh * N + nh * LD;
      double t = 0.0;
      for (int w = 0; w < NW; ++w) t += wpart[(size_t)w * GP + off];
      my_partial[idx] = t;
    }
    cluster.sync();
    // reduce-scatter: CTA `rank` sums slice `rank` over the cluster; all-gather: everybody pulls all slices
    const int slice = (G + nranks - 1) / nranks;
    const int lo = rank * slice, hi = lo + slice < G ? lo + slice : G;
    for (int idx = lo + tid; idx < hi; idx += NT) {
      double g = 0.0;
      for (int rk = 0; rk < nranks; ++rk) g += cluster.map_shared_rank(my_partial, rk)[idx];
      my_red[idx - lo] = g;
      if (a.g_last != nullptr && step == a.num_batches - 1) a.g_last[idx] = g;
    }
    cluster.sync();
    if (a.lr != 0.0) {  // lr = 0: gradient only (qcb_rbm_cd_step semantics in one launch)
      for (int idx = tid; idx < G; idx += NT) {
        const int owner = idx / slice;
        const double g = cluster.map_shared_rank(my_red, owner)[idx - owner * slice];
        double* prm = idx < nh * N ? &Wsm[(idx / N) * LD + idx % N] : (idx < nh * N + N ? &bsm[idx - nh * N] : &csm[idx - nh * N - N]);
        *prm -= a.lr * g;
      }
    }
    __syncthreads();
  }
  cluster.sync();  // nobody leaves while its shares may still be read
  if (rank == 0 && a.lr != 0.0) {
    for (int idx = tid; idx < nh * N; idx += NT) a.W[idx] = Wsm[(idx / N) * LD + idx % N];
    for (int idx = tid; idx < N; idx += NT) a.b[idx] = bsm[idx];
    for (int idx = tid; idx < nh; idx += NT) a.c[idx] = csm[idx];
  }
}

// ---- warp-specialised variant for layers of at most 32 units (round 2) ----------------------------------------------------
// ncu on cd_epoch_kernel: issue slots 27 % busy, a latency chain — a step is 20 dependent conditional updates of 0.36 us, and
// half of the instructions a warp issues per update are the Philox rounds of draws that do not depend on the chain's state.
// Here the CTA has two kinds of warps:
//  * PRODUCER warps (a) draw the 16-bit cells of every (chain slot, sweep, layer, unit) of the next sweeps into a ring of
//    shared-memory buffers (one lane per Philox CALL: four / eight units at once, not one call per unit) and (b) after every
//    parameter update rebuild float NIBBLE TABLES of the pre-activations: T[layer][unit][t][m] = sum of the unit's weights
//    towards input units 4 t .. 4 t + 3 selected by the bits of m (the bias rides in table 0), plus the unit's filter margin;
//  * CONSUMER warps run what depends on the state: the pre-activation is SW / 4 table reads and a short add tree (instead of
//    SW predicated adds), the filter in the division-free form (L q - 1 with q = 1 + 2^(-x log2 e)), two votes.  The float sum
//    carries a rigorous margin ((n_in + 6) 2^-24 (|bias| + sum |w|): conversions and every float addition, a quarter of it on
//    top of the 2^-17 of the float sigmoid); a warp with an undecided lane recomputes that lane in FP64 from the shared-memory
//    weights and takes the exact path.
// Hand-over by named barriers (full[b]: producers arrive, consumers sync; empty[b]: the reverse; tables: producers arrive,
// consumers sync).  Same draws and the same exact decisions as cd_epoch_kernel; up to 16 units also the same reduction order —
// the two kernels return the same parameters bit for bit there (17..32 units: 8 consumer warps of one chain each against the
// 16 warps of cd_epoch_kernel<32>, so the rows are dealt to different slots and the sums agree to rounding).
constexpr int WS_CONS = 256, WS_PROD = 256, WS_THREADS = WS_CONS + WS_PROD;
// ring depth x sweeps per buffer, measured at N = nh = 10, k = 10: 4 x 2: 0.657 ms per epoch, 6 x 2: 0.643, 3 x 5: 0.610, 2 x 10: 0.594
// (every hand-over is a 512-thread named barrier that also couples the consumer warps to each other)
#ifndef WS_NBUF_V
#define WS_NBUF_V 2
#endif
#ifndef WS_KC_V
#define WS_KC_V 10
#endif
constexpr int WS_NBUF = WS_NBUF_V;       // ring depth
constexpr int WS_KC = WS_KC_V;         // sweeps per ring buffer
constexpr int WS_BAR_CONS = 1 + 2 * WS_NBUF, WS_BAR_TABLES = 2 + 2 * WS_NBUF, WS_BAR_STEP = 3 + 2 * WS_NBUF;
__device__ __forceinline__ void ws_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void ws_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// floats behind the FP64 region: cells [WS_NBUF][SPC][2 WS_KC][SW], tables [2 layers][SW units][SW / 4 x 16 + 1 pad], margins [2][SW]
template <int SW>
__host__ __device__ inline size_t ws_smem_bytes(int N, int nh) {
  constexpr int NWC = WS_CONS / 32, SPC = NWC * (32 / SW);
  return ce_smem_doubles(N, nh, SW, NWC) * sizeof(double) +
         ((size_t)WS_NBUF * SPC * 2 * WS_KC * SW + (size_t)2 * SW * ((SW / 4) * 16 + 1) + 2 * SW) * sizeof(float);
}

template <int SW>
__global__ void __launch_bounds__(WS_THREADS, 1) cd_epoch_ws_kernel(const CdEpochArgs a) {
  extern __shared__ __align__(16) double ce_smem[];
  constexpr int NT = WS_CONS, NW = NT / 32, CPW = 32 / SW, SPC = NW * CPW;   // consumer threads / warps, slots per warp / per CTA
  constexpr int CHUNK_FLOATS = SPC * 2 * WS_KC * SW;
  constexpr int NT4 = SW / 4;                                                 // nibble tables per unit
  constexpr int TAB_LD = NT4 * 16 + 1;                                        // a unit's tables, padded: the lanes of a slot read the SAME nibble
  constexpr int TAB_FLOATS = SW * TAB_LD;                                     // per layer
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int N = a.N, nh = a.nh, LD = N | 1, G = nh * N + N + nh, GP = nh * LD + N + nh;
  double* Wsm = ce_smem;
  double* bsm = Wsm + (size_t)SW * LD;  // W region padded to SW rows
  double* csm = bsm + N;
  double* partial = csm + nh;
  double* wpart = partial + 2 * (size_t)G;
  float* cells = reinterpret_cast<float*>(ce_smem + ce_smem_doubles(N, nh, SW, NW));  // cell * 2^-16
  float* tabs = cells + (size_t)WS_NBUF * CHUNK_FLOATS;                               // [layer][unit][t][m]
  float* margins = tabs + 2 * TAB_FLOATS;                                             // [layer][unit]
  const int tid = threadIdx.x, lane = tid & 31;
  const int nranks = (int)cluster.num_blocks();
  const int nslots = nranks * SPC;
  const uint32_t qv0 = 4u * (uint32_t)((nh + 31) / 32);

  __shared__ double e2tab[EXP2_TABLE];
  fill_exp2_table(e2tab);
  for (int idx = tid; idx < nh * N; idx += WS_THREADS) Wsm[(idx / N) * LD + idx % N] = a.W[idx];
  for (int idx = tid; idx < N; idx += WS_THREADS) bsm[idx] = a.b[idx];
  for (int idx = tid; idx < nh; idx += WS_THREADS) csm[idx] = a.c[idx];
  __syncthreads();

  if (tid >= WS_CONS) {
    // ================================= producers =================================
    const int ptid = tid - WS_CONS;
    int ring_next = 0;   // chunks drawn so far = the ring position of the next one
    // weight of hidden unit i / visible unit j as seen from output unit `unit` of `layer` towards its input `in`
    auto weight = [&](int layer, int unit, int in) -> double {
      if (layer == 0) return (unit < nh && in < N) ? Wsm[unit * LD + in] : 0.0;   // h | v: row of W
      return (unit < N && in < nh) ? Wsm[in * LD + unit] : 0.0;                    // v | h: column of W
    };
    auto build_tables = [&]() {
      // one thread per (layer, unit, nibble t): its four weights once, then the 16 subset sums in a fixed order
      for (int e = ptid; e < 2 * SW * NT4; e += WS_PROD) {
        const int t = e % NT4, unit = (e / NT4) % SW, layer = e / (SW * NT4);
        const int units = layer ? N : nh;
        float w4[4] = {0.0f, 0.0f, 0.0f, 0.0f}, base = 0.0f;
        if (unit < units) {
          if (t == 0) base = (float)(layer ? bsm[unit] : csm[unit]);
#pragma unroll
          for (int i = 0; i < 4; ++i) w4[i] = (float)weight(layer, unit, 4 * t + i);
        }
        float* dst = tabs + (size_t)layer * TAB_FLOATS + (size_t)unit * TAB_LD + 16 * t;
#pragma unroll
        for (int m = 0; m < 16; ++m) {
          float v = base;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if ((m >> i) & 1) v += w4[i];
          dst[m] = v;
        }
      }
      // margins: the last warps of the producer group, one unit each
      for (int e = ptid - (WS_PROD - 2 * SW); e >= 0 && e < 2 * SW; e += WS_PROD) {
        const int unit = e % SW, layer = e / SW;
        const int units = layer ? N : nh, n_in = layer ? nh : N;
        float mg = 4.0f;   // padding lanes, non-finite or huge weights: everything goes to the exact path (or is masked)
        if (unit < units) {
          double sabs = fabs(layer ? bsm[unit] : csm[unit]);
          for (int in = 0; in < n_in; ++in) sabs += fabs(weight(layer, unit, in));
          const double md = (0.25 * (double)(n_in + 6) * 0x1.0p-24 * sabs + 0x1.0p-17) * 1.001;
          if (md < 0.25) mg = (float)md;
        }
        margins[e] = mg;
      }
      __threadfence_block();
      __syncwarp();
      ws_bar_arrive(WS_BAR_TABLES, WS_THREADS);
    };
    build_tables();
    // Chunks are numbered through the whole epoch: chunk q belongs to step q / cps (cps chunks per full step; a short last
    // negative batch has fewer rounds, its surplus chunks are never drawn).  Before the cluster barriers of step t the
    // producers have drawn all of step t's chunks and up to WS_NBUF chunks of step t + 1: the ring slots those need were last
    // used by chunks of step <= t, which the consumers release before they reach the same barriers — no producer ever waits for
    // something that happens behind a barrier it has not joined.
    auto draw_chunk = [&](int step, int local) {   // local = round * chunks_per_round + sweep chunk, within the step
      const int cpr = (a.k + WS_KC - 1) / WS_KC;
      const int ch0 = (local / cpr) * nslots, s0 = (local % cpr) * WS_KC;
      const int ns = a.k - s0 < WS_KC ? a.k - s0 : WS_KC;
      for (int item = ptid; item < SPC * 2 * WS_KC * 4; item += WS_PROD) {
        // item = (slot of the CTA, conditional update of the chunk, Philox call c): the call serves units 4 c .. 4 c + 3 (high
        // 16-bit halves) and 16 + 4 c .. 16 + 4 c + 3 (low halves) of the layer
        const int c = item & 3, cond = (item >> 2) % (2 * WS_KC), sl = (item >> 2) / (2 * WS_KC);
        const int layer = cond & 1;                       // 0: h | v, 1: v | h
        const int units = layer ? N : nh;
        if (4 * c >= units || (cond >> 1) >= ns) continue;
        const uint32_t ch = (uint32_t)(ch0 + rank * SPC + sl);
        const uint32_t sweep = a.sweep0 + (uint32_t)(step * a.k + s0 + (cond >> 1));
        uint32_t r[4];
        philox4x32_10(ch, 0u, sweep, (layer ? qv0 : 0u) + (uint32_t)c, a.key0, a.key1, r);
        float* dst = cells + (size_t)(ring_next % WS_NBUF) * CHUNK_FLOATS + ((size_t)sl * 2 * WS_KC + cond) * SW;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          if (4 * c + w < SW) dst[4 * c + w] = (float)(r[w] >> 16) * 0x1.0p-16f;
          if (SW > 16 && 16 + 4 * c + w < SW) dst[16 + 4 * c + w] = (float)(r[w] & 0xFFFFu) * 0x1.0p-16f;
        }
      }
    };
    auto chunks_of_step = [&](int step) {
      if (step >= a.num_batches) return 0;
      const int64_t c0 = (int64_t)step * a.neg_batch;
      const int Cs = (int)(a.C_total - c0 < a.neg_batch ? a.C_total - c0 : a.neg_batch);
      return ((Cs + nslots - 1) / nslots) * ((a.k + WS_KC - 1) / WS_KC);
    };
    int done_in_next = 0;   // chunks of the coming step that were drawn ahead
    for (int step = 0; step < a.num_batches; ++step) {
      const int mine = chunks_of_step(step), next = chunks_of_step(step + 1);
      const int ahead = next < WS_NBUF ? next : WS_NBUF;
      for (int phase = 0; phase < 2; ++phase) {          // 0: the rest of this step, 1: the first chunks of the next one
        const int from = phase == 0 ? done_in_next : 0, to = phase == 0 ? mine : ahead;
        for (int local = from; local < to; ++local, ++ring_next) {
          const int b = ring_next % WS_NBUF;
          if (ring_next >= WS_NBUF) ws_bar_sync(1 + WS_NBUF + b, WS_THREADS);  // the consumers have released this buffer
          draw_chunk(step + phase, local);
          __threadfence_block();
          __syncwarp();
          ws_bar_arrive(1 + b, WS_THREADS);                                    // full
        }
      }
      done_in_next = ahead;
      cluster.sync();   // the cluster barriers of the step involve every thread of the cluster
      cluster.sync();
      ws_bar_sync(WS_BAR_STEP, WS_THREADS);   // the consumers have applied the update
      if (step + 1 < a.num_batches) build_tables();
    }
    cluster.sync();
    return;
  }

  // ================================= consumers =================================
  const int warp = tid >> 5;
  const int u = lane % SW, sub = lane / SW;
  const uint32_t slot_mask = SW == 32 ? 0xffffffffu : ((1u << SW) - 1u);
  const int slot_in_cta = warp * CPW + sub;
  const int slot = rank * SPC + slot_in_cta;
  const int uh = u < nh ? u : nh - 1, uv = u < N ? u : N - 1;
  auto hidden_pre = [&](uint32_t vm) {
    const double* row = Wsm + uh * LD;
    double x[4] = {csm[uh], 0.0, 0.0, 0.0};
#pragma unroll
    for (int j = 0; j < SW; ++j)
      if ((vm >> j) & 1u) x[j & 3] += row[j];
    return (x[0] + x[1]) + (x[2] + x[3]);
  };
  auto visible_pre = [&](uint32_t hm) {
    const double* col = Wsm + uv;
    double x[4] = {bsm[uv], 0.0, 0.0, 0.0};
#pragma unroll
    for (int i = 0; i < SW; ++i)
      if ((hm >> i) & 1u) x[i & 3] += col[i * LD];
    return (x[0] + x[1]) + (x[2] + x[3]);
  };
  auto slot_bits = [&](bool bit) { return (__ballot_sync(0xffffffffu, bit) >> (sub * SW)) & slot_mask; };
  auto first_row = [&](const double* rows_all, int64_t total, int batch, int step) {
    const int64_t start = (int64_t)step * batch;
    const int64_t rows = total - start < batch ? total - start : batch;
    return (step < a.num_batches && slot < rows && u < N) ? rows_all[(start + slot) * N + u] : 0.0;
  };
  double next_pos = first_row(a.pos, a.B_total, a.pos_batch, 0), next_neg = first_row(a.neg, a.C_total, a.neg_batch, 0);

  // float pre-activation of this lane's unit from the nibble tables: SW / 4 independent reads and a short add tree
  const float* tab_h = tabs + (size_t)u * TAB_LD;
  const float* tab_v = tabs + TAB_FLOATS + (size_t)u * TAB_LD;
  auto table_pre = [&](const float* tab, uint32_t mask) {
    float part[NT4];
#pragma unroll
    for (int t = 0; t < NT4; ++t) part[t] = tab[16 * t + ((mask >> (4 * t)) & 15u)];
#pragma unroll
    for (int w = 1; w < NT4; w <<= 1)
#pragma unroll
      for (int t = 0; t + w < NT4; t += 2 * w) part[t] += part[t + w];
    return part[0];
  };
  // bit 0: u < sigmoid(x) certainly true, bit 1: certainly false (neither: undecided).  lof = cell 2^-16;
  // L q - 1 < 0  <=>  L < 1 / q = sigmoid
  auto filter = [&](float xf, float lof, float margin, bool& sure1, bool& sure0) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(xf * -1.4426950408889634f));
    const float q = 1.0f + e;
    const bool in_range = fabsf(xf) <= 64.0f;
    sure1 = in_range && fmaf(lof + (0x1.0p-16f + margin), q, -1.0f) < 0.0f;
    sure0 = in_range && fmaf(lof - margin, q, -1.0f) >= 0.0f;
  };

  int chunk = 0;
  for (int step = 0; step < a.num_batches; ++step) {
    const int64_t p0 = (int64_t)step * a.pos_batch, c0 = (int64_t)step * a.neg_batch;
    const int Bs = (int)(a.B_total - p0 < a.pos_batch ? a.B_total - p0 : a.pos_batch);
    const int Cs = (int)(a.C_total - c0 < a.neg_batch ? a.C_total - c0 : a.neg_batch);
    const double* pos_rows = a.pos + p0 * N;
    const double* neg_rows = a.neg + c0 * N;
    const double first_pos = next_pos, first_neg = next_neg;
    next_pos = first_row(a.pos, a.B_total, a.pos_batch, step + 1);
    next_neg = first_row(a.neg, a.C_total, a.neg_batch, step + 1);
    double accW[SW];
#pragma unroll
    for (int j = 0; j < SW; ++j) accW[j] = 0.0;
    double accb = 0.0, accc = 0.0;
    auto accumulate = [&](uint32_t vm, bool valid, double weight) {
      const double sg = sigmoid_tab(hidden_pre(vm), e2tab);
      const double pw = (valid && u < nh) ? weight * sg : 0.0;
      accc += pw;
#pragma unroll
      for (int j = 0; j < SW; ++j)
        if ((vm >> j) & 1u) accW[j] += pw;
      if (valid && ((vm >> u) & 1u)) accb += weight;
    };

    for (int r0 = 0; r0 < Bs; r0 += nslots) {
      const int r = r0 + slot;
      const bool valid = r < Bs;
      const double raw = r0 == 0 ? first_pos : ((valid && u < N) ? pos_rows[(int64_t)r * N + u] : 0.0);
      const uint32_t vm = slot_bits(raw != 0.0);
      accumulate(vm, valid, -1.0 / (double)Bs);
    }
    ws_bar_sync(WS_BAR_TABLES, WS_THREADS);   // the producers have rebuilt the tables and margins from this step's parameters
    const float mar_h = margins[u], mar_v = margins[SW + u];
    for (int ch0 = 0; ch0 < Cs; ch0 += nslots) {
      const int ch = ch0 + slot;
      const bool valid = ch < Cs;
      const double raw = ch0 == 0 ? first_neg : ((valid && u < N) ? neg_rows[(int64_t)ch * N + u] : 0.0);
      uint32_t vm = slot_bits(raw != 0.0);
      const uint32_t cid0 = (uint32_t)ch, cid1 = 0u;
      const bool warp_has_chain = __any_sync(0xffffffffu, valid);
      for (int s0 = 0; s0 < a.k; s0 += WS_KC, ++chunk) {
        const int b = chunk % WS_NBUF;
        ws_bar_sync(1 + b, WS_THREADS);   // full: the producers have drawn this chunk's cells
        const float* buf = cells + (size_t)b * CHUNK_FLOATS + (size_t)slot_in_cta * 2 * WS_KC * SW + u;
        const int ns = a.k - s0 < WS_KC ? a.k - s0 : WS_KC;
        if (warp_has_chain) {
          for (int ss = 0; ss < ns; ++ss) {
            const uint32_t sweep = a.sweep0 + (uint32_t)(step * a.k + s0 + ss);
            const float lof_h = buf[(2 * ss) * SW], lof_v = buf[(2 * ss + 1) * SW];
            bool s1, s0f;
            filter(table_pre(tab_h, vm), lof_h, mar_h, s1, s0f);
            bool one = s1 && u < nh;
            const bool open_h = !(s1 || s0f) && u < nh;
            uint32_t hm_all = __ballot_sync(0xffffffffu, one);           // both votes are issued back to back
            if (__any_sync(0xffffffffu, open_h)) {
              if (open_h) one = ce_decide_exact(hidden_pre(vm), (uint32_t)(lof_h * 65536.0f), u, 0u, cid0, cid1, sweep, a.key0, a.key1);
              __syncwarp();
              hm_all = __ballot_sync(0xffffffffu, one);
            }
            const uint32_t hm = (hm_all >> (sub * SW)) & slot_mask;
            filter(table_pre(tab_v, hm), lof_v, mar_v, s1, s0f);
            one = s1 && u < N;
            const bool open_v = !(s1 || s0f) && u < N;
            uint32_t vm_all = __ballot_sync(0xffffffffu, one);
            if (__any_sync(0xffffffffu, open_v)) {
              if (open_v) one = ce_decide_exact(visible_pre(hm), (uint32_t)(lof_v * 65536.0f), u, qv0, cid0, cid1, sweep, a.key0, a.key1);
              __syncwarp();
              vm_all = __ballot_sync(0xffffffffu, one);
            }
            vm = (vm_all >> (sub * SW)) & slot_mask;
          }
        }
        ws_bar_arrive(1 + WS_NBUF + b, WS_THREADS);   // empty
      }
      accumulate(vm, valid, 1.0 / (double)Cs);
    }

    // the slots of a warp, then the warps of the CTA, in a fixed order (as in cd_epoch_kernel)
#pragma unroll
    for (int off = SW; off < 32; off <<= 1) {
#pragma unroll
      for (int j = 0; j < SW; ++j) accW[j] += __shfl_xor_sync(0xffffffffu, accW[j], off);
      accb += __shfl_xor_sync(0xffffffffu, accb, off);
      accc += __shfl_xor_sync(0xffffffffu, accc, off);
    }
    double* mine = wpart + (size_t)warp * GP;
    if (sub == 0) {
      if (u < nh) {
#pragma unroll
        for (int j = 0; j < SW; ++j)
          if (j < N) mine[u * LD + j] = accW[j];
        mine[nh * LD + N + u] = accc;
      }
      if (u < N) mine[nh * LD + u] = accb;
    }
    ws_bar_sync(WS_BAR_CONS, WS_CONS);
    double* my_partial = partial;
    double* my_red = partial + G;
    for (int idx = tid; idx < G; idx += NT) {
      const int off = idx < nh * N ? (idx / N) * LD + idx % N : idx - nh * N + nh * LD;
      double t = 0.0;
      for (int w = 0; w < NW; ++w) t += wpart[(size_t)w * GP + off];
      my_partial[idx] = t;
    }
    // the exchange of cd_epoch_kernel (reduce-scatter and all-gather by remote READS, two cluster barriers).  Measured against
    // it in this kernel and not faster: an all-to-all push of every share with one barrier (0.661 vs 0.652 ms per epoch at
    // N = 10, 0.770 vs 0.723 at N = 16), and pushes to the slice owner / from the owner to everybody (0.679 vs 0.658 ms) —
    // the barrier's release waits for the remote stores, which costs what the remote reads' latency costs.
    cluster.sync();
    const int slice = (G + nranks - 1) / nranks;
    const int lo = rank * slice, hi = lo + slice < G ? lo + slice : G;
    for (int idx = lo + tid; idx < hi; idx += NT) {
      double g = 0.0;
      for (int rk = 0; rk < nranks; ++rk) g += cluster.map_shared_rank(my_partial, rk)[idx];
      my_red[idx - lo] = g;
      if (a.g_last != nullptr && step == a.num_batches - 1) a.g_last[idx] = g;
    }
    cluster.sync();
    if (a.lr != 0.0) {
      for (int idx = tid; idx < G; idx += NT) {
        const int owner = idx / slice;
        const double g = cluster.map_shared_rank(my_red, owner)[idx - owner * slice];
        double* prm = idx < nh * N ? &Wsm[(idx / N) * LD + idx % N] : (idx < nh * N + N ? &bsm[idx - nh * N] : &csm[idx - nh * N - N]);
        *prm -= a.lr * g;
      }
    }
    ws_bar_sync(WS_BAR_STEP, WS_THREADS);   // parameters final: the producers rebuild the tables, the consumers go on
  }
  cluster.sync();  // nobody leaves while its shares may still be read
  if (rank == 0 && a.lr != 0.0) {
    for (int idx = tid; idx < nh * N; idx += NT) a.W[idx] = Wsm[(idx / N) * LD + idx % N];
    for (int idx = tid; idx < N; idx += NT) a.b[idx] = bsm[idx];
    for (int idx = tid; idx < nh; idx += NT) a.c[idx] = csm[idx];
  }
}

// ---- layers of 33 .. 64 units ------------------------------------------------------------------------
// Same scheme with one chain per warp and two units per lane (u and u + 32).  A lane cannot keep a row of the
// gradient in registers any more, so after every round the warps park (w p, v, w) of their row in shared memory
// and all threads add the round's outer products to the gradient entries they own (17 per thread).  The step's
// gradient is reduce-scattered and all-gathered across the cluster through distributed shared memory
// (two cluster barriers per step), and every CTA applies the identical update to its copy of W, b, c.
constexpr int CEW_THREADS = 512, CEW_WARPS = CEW_THREADS / 32, CEW_LD = 65, CEW_GMAX = 64 * 64 + 128;
constexpr int CEW_TLD = 16 * 16 + 1;  // floats of one unit's nibble tables (+ 1: bank padding)
constexpr int CEW_EPT = (CEW_GMAX + CEW_THREADS - 1) / CEW_THREADS;  // gradient entries owned by a thread
static_assert(CEW_EPT * CEW_THREADS >= CEW_GMAX, "every gradient entry needs an owner");

__host__ __device__ inline size_t cew_smem_doubles() {
  // W (64 x 65) | b (64) | c (64) | parked rows (warps x 64) | row weights | row masks | own share (G) | reduced slice (<= G / 8)
  // ... | float nibble tables of both layers and the per-unit filter margins (2 x 64 x 257 + 2 x 64 floats)
  return (size_t)64 * CEW_LD + 64 + 64 + CEW_WARPS * 64 + CEW_WARPS + CEW_WARPS + CEW_GMAX + (CEW_GMAX + 7) / 8 + 8 + (2 * 64 * CEW_TLD + 2 * 64 + 1) / 2;
}

__global__ void __launch_bounds__(CEW_THREADS, 1) cd_epoch_wide_kernel(const CdEpochArgs a) {
  extern __shared__ __align__(16) double ce_smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank(), nranks = (int)cluster.num_blocks();
  const int N = a.N, nh = a.nh, G = nh * N + N + nh;
  double* Wsm = ce_smem;
  double* bsm = Wsm + 64 * CEW_LD;
  double* csm = bsm + 64;
  double* Prow = csm + 64;
  double* wts = Prow + CEW_WARPS * 64;
  unsigned long long* masks = reinterpret_cast<unsigned long long*>(wts + CEW_WARPS);
  double* Gpart = reinterpret_cast<double*>(masks + CEW_WARPS);
  double* Gred = Gpart + CEW_GMAX;
  // float NIBBLE TABLES of the pre-activations for the filter: T[unit][t][m] = (bias if t = 0) + the unit's weights towards
  // inputs 4 t .. 4 t + 3 selected by the bits of m; a unit's 16 tables are padded by one float (the 32 lanes of a warp read the
  // same (t, m) of 32 different units)
  float* Th = reinterpret_cast<float*>(Gred + (CEW_GMAX + 7) / 8 + 8);
  float* Tv = Th + 64 * CEW_TLD;
  float* mrg_h = Tv + 64 * CEW_TLD;  // filter margins per hidden / visible unit
  float* mrg_v = mrg_h + 64;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int slot = rank * CEW_WARPS + warp, nslots = nranks * CEW_WARPS;
  const int u0 = lane, u1 = lane + 32;
  const int slice = (G + nranks - 1) / nranks;  // gradient entries reduced by one CTA

  __shared__ double e2tab[EXP2_TABLE];
  fill_exp2_table(e2tab);
  for (int idx = tid; idx < 64 * CEW_LD; idx += CEW_THREADS) Wsm[idx] = 0.0;
  __syncthreads();
  for (int idx = tid; idx < nh * N; idx += CEW_THREADS) Wsm[(idx / N) * CEW_LD + idx % N] = a.W[idx];
  for (int idx = tid; idx < 64; idx += CEW_THREADS) {
    bsm[idx] = idx < N ? a.b[idx] : 0.0;
    csm[idx] = idx < nh ? a.c[idx] : 0.0;
  }
  __syncthreads();

  // pre-activations of this lane's two hidden (visible) units; padded rows / columns of W are zero
  auto hidden_pre2 = [&](unsigned long long vm, double& x0, double& x1) {
    const double* r0 = Wsm + u0 * CEW_LD;
    const double* r1 = Wsm + u1 * CEW_LD;
    const uint32_t lo = (uint32_t)vm, hi = (uint32_t)(vm >> 32);
    double s0[4] = {csm[u0], 0.0, 0.0, 0.0}, s1[4] = {csm[u1], 0.0, 0.0, 0.0};
#pragma unroll
    for (int j = 0; j < 32; ++j) {  // fully unrolled: predicated adds, no branches
      if ((lo >> j) & 1u) { s0[j & 3] += r0[j]; s1[j & 3] += r1[j]; }
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if ((hi >> j) & 1u) { s0[j & 3] += r0[32 + j]; s1[j & 3] += r1[32 + j]; }
    }
    x0 = (s0[0] + s0[1]) + (s0[2] + s0[3]);
    x1 = (s1[0] + s1[1]) + (s1[2] + s1[3]);
  };
  auto visible_pre2 = [&](unsigned long long hm, double& x0, double& x1) {
    const double* c0 = Wsm + u0;
    const double* c1 = Wsm + u1;
    const uint32_t lo = (uint32_t)hm, hi = (uint32_t)(hm >> 32);
    double s0[4] = {bsm[u0], 0.0, 0.0, 0.0}, s1[4] = {bsm[u1], 0.0, 0.0, 0.0};
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      if ((lo >> i) & 1u) { s0[i & 3] += c0[i * CEW_LD]; s1[i & 3] += c1[i * CEW_LD]; }
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      if ((hi >> i) & 1u) { s0[i & 3] += c0[(32 + i) * CEW_LD]; s1[i & 3] += c1[(32 + i) * CEW_LD]; }
    }
    x0 = (s0[0] + s0[1]) + (s0[2] + s0[3]);
    x1 = (s1[0] + s1[1]) + (s1[2] + s1[3]);
  };
  // Float image and margins, rebuilt whenever the parameters change.  x~ (float sum of float-rounded weights) obeys
  // |x~ - x| <= (n_in + 20) 2^-24 (|bias| + sum |w|) =: dx (see below), so |sigmoid(x~) - sigmoid(x)| <= dx / 4; 2^-17 covers the
  // float evaluation of the sigmoid itself (ce_filter).  Units the filter leaves open are redone in FP64.
  auto refresh_float = [&]() {
    // one thread per (layer, unit, nibble): its four weights once, then the 16 subset sums in a fixed order
    for (int e = tid; e < 2 * 64 * 16; e += CEW_THREADS) {
      const int t = e & 15, unit = (e >> 4) & 63, layer = e >> 10;
      float w4[4], base = 0.0f;
#pragma unroll
      for (int i = 0; i < 4; ++i) w4[i] = (float)(layer ? Wsm[(4 * t + i) * CEW_LD + unit] : Wsm[unit * CEW_LD + 4 * t + i]);
      if (t == 0) base = (float)(layer ? bsm[unit] : csm[unit]);
      float* dst = (layer ? Tv : Th) + unit * CEW_TLD + 16 * t;
#pragma unroll
      for (int m = 0; m < 16; ++m) {
        float v = base;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if ((m >> i) & 1) v += w4[i];
        dst[m] = v;
      }
    }
    // |x~ - x| <= (n_in + 20) 2^-24 (|bias| + sum |w|): the conversions, at most four additions per table and the 15 of the tree
    if (tid < 64) {
      double S = fabs(csm[tid]);
      for (int j = 0; j < 64; ++j) S += fabs(Wsm[tid * CEW_LD + j]);
      mrg_h[tid] = (float)(((N + 20) * 0x1.0p-24 * S * 0.25 + 0x1.0p-17) * 1.001);
    } else if (tid < 128) {
      const int j = tid - 64;
      double S = fabs(bsm[j]);
      for (int i = 0; i < 64; ++i) S += fabs(Wsm[i * CEW_LD + j]);
      mrg_v[j] = (float)(((nh + 20) * 0x1.0p-24 * S * 0.25 + 0x1.0p-17) * 1.001);
    }
    __syncthreads();
  };
  // float pre-activations of this lane's two units from the tables: 16 reads and a four-level add tree per unit (the
  // predicated load-add pairs per weight of the first version were 256 instructions per conditional update)
  auto table_pre2 = [&](const float* tab, unsigned long long m, float& x0, float& x1) {
    const float* t0 = tab + u0 * CEW_TLD;
    const float* t1 = tab + u1 * CEW_TLD;
    float p0[16], p1[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) {
      const int idx = 16 * t + (int)((m >> (4 * t)) & 15ull);
      p0[t] = t0[idx];
      p1[t] = t1[idx];
    }
#pragma unroll
    for (int w = 1; w < 16; w <<= 1)
#pragma unroll
      for (int t = 0; t + w < 16; t += 2 * w) { p0[t] += p0[t + w]; p1[t] += p1[t + w]; }
    x0 = p0[0];
    x1 = p1[0];
  };
  auto hidden_pre2_f = [&](unsigned long long vm, float& x0, float& x1) { table_pre2(Th, vm, x0, x1); };
  auto visible_pre2_f = [&](unsigned long long hm, float& x0, float& x1) { table_pre2(Tv, hm, x0, x1); };
  auto mask64 = [&](bool b0, bool b1) {
    return (unsigned long long)__ballot_sync(0xffffffffu, b0) | ((unsigned long long)__ballot_sync(0xffffffffu, b1) << 32);
  };
  auto load_row = [&](const double* rows, int64_t r, bool valid) {
    const double a0 = (valid && u0 < N) ? rows[r * N + u0] : 0.0;
    const double a1 = (valid && u1 < N) ? rows[r * N + u1] : 0.0;
    return mask64(a0 != 0.0, a1 != 0.0);
  };
  // a Bernoulli draw for units u0, u1 of one layer from the state `in`: float pre-activations and filter; if any lane
  // of the warp is left undecided the warp recomputes the FP64 pre-activations and those lanes take the exact path
  auto draw2 = [&](unsigned long long in, bool hidden_layer, uint32_t q0, uint32_t cid0, uint32_t sweep) {
    const int n_units = hidden_layer ? nh : N;
    const float* mrg = hidden_layer ? mrg_h : mrg_v;
    float xf0, xf1;
    if (hidden_layer) hidden_pre2_f(in, xf0, xf1); else visible_pre2_f(in, xf0, xf1);
    const uint32_t cell0 = ce_cell(u0, q0, cid0, 0u, sweep, a.key0, a.key1);
    const uint32_t cell1 = ce_cell(u1, q0, cid0, 0u, sweep, a.key0, a.key1);
    int v0 = ce_filter_f(xf0, cell0, mrg[u0]), v1 = ce_filter_f(xf1, cell1, mrg[u1]);
    v0 = u0 < n_units ? v0 : 0;
    v1 = u1 < n_units ? v1 : 0;
    if (__any_sync(0xffffffffu, v0 == 2 || v1 == 2)) {
      double x0, x1;
      if (hidden_layer) hidden_pre2(in, x0, x1); else visible_pre2(in, x0, x1);
      if (v0 == 2) v0 = ce_decide_exact(x0, cell0, u0, q0, cid0, 0u, sweep, a.key0, a.key1) ? 1 : 0;
      if (v1 == 2) v1 = ce_decide_exact(x1, cell1, u1, q0, cid0, 0u, sweep, a.key0, a.key1) ? 1 : 0;
      __syncwarp();
    }
    return mask64(v0 == 1, v1 == 1);
  };
  refresh_float();

  for (int step = 0; step < a.num_batches; ++step) {
    const int64_t p0 = (int64_t)step * a.pos_batch, c0 = (int64_t)step * a.neg_batch;
    const int Bs = (int)(a.B_total - p0 < a.pos_batch ? a.B_total - p0 : a.pos_batch);
    const int Cs = (int)(a.C_total - c0 < a.neg_batch ? a.C_total - c0 : a.neg_batch);
    const double* pos_rows = a.pos + p0 * N;
    const double* neg_rows = a.neg + c0 * N;
    double gacc[CEW_EPT];
#pragma unroll
    for (int e = 0; e < CEW_EPT; ++e) gacc[e] = 0.0;

    // the warp's row: (w p, v, w) to shared memory; then every thread adds the round to its entries
    auto park = [&](unsigned long long vm, bool valid, double weight) {
      double x0, x1;
      hidden_pre2(vm, x0, x1);
      Prow[warp * 64 + u0] = (valid && u0 < nh) ? weight * sigmoid_tab(x0, e2tab) : 0.0;
      Prow[warp * 64 + u1] = (valid && u1 < nh) ? weight * sigmoid_tab(x1, e2tab) : 0.0;
      if (lane == 0) {
        wts[warp] = valid ? weight : 0.0;
        masks[warp] = valid ? vm : 0ull;
      }
    };
    auto accumulate_round = [&]() {
      __syncthreads();
#pragma unroll
      for (int e = 0; e < CEW_EPT; ++e) {
        const int idx = tid + CEW_THREADS * e;
        if (idx < G) {
          double g = 0.0;
          if (idx < nh * N) {
            const int i = idx / N, j = idx - i * N;
#pragma unroll
            for (int r = 0; r < CEW_WARPS; ++r)
              if ((masks[r] >> j) & 1ull) g += Prow[r * 64 + i];
          } else if (idx < nh * N + N) {
            const int j = idx - nh * N;
#pragma unroll
            for (int r = 0; r < CEW_WARPS; ++r)
              if ((masks[r] >> j) & 1ull) g += wts[r];
          } else {
            const int i = idx - nh * N - N;
#pragma unroll
            for (int r = 0; r < CEW_WARPS; ++r) g += Prow[r * 64 + i];
          }
          gacc[e] += g;
        }
      }
      __syncthreads();
    };

    for (int r0 = 0; r0 < Bs; r0 += nslots) {  // positive phase: grad E(v) = -[p v^T ; v ; p], weight -1/B
      const int r = r0 + slot;
      const bool valid = r < Bs;
      park(load_row(pos_rows, r, valid), valid, -1.0 / (double)Bs);
      accumulate_round();
    }
    for (int ch0 = 0; ch0 < Cs; ch0 += nslots) {  // negative phase: k sweeps, then weight +1/C
      const int ch = ch0 + slot;
      const bool valid = ch < Cs;
      unsigned long long vm = load_row(neg_rows, ch, valid);
      const uint32_t qv0 = 4u * (uint32_t)((nh + 31) / 32);
      for (int s = 0; s < (valid ? a.k : 0); ++s) {  // a warp without a chain in this round sits the sweeps out
        const uint32_t sweep = a.sweep0 + (uint32_t)(step * a.k + s);
        const unsigned long long hm = draw2(vm, true, 0u, (uint32_t)ch, sweep);
        vm = draw2(hm, false, qv0, (uint32_t)ch, sweep);
      }
      park(vm, valid, 1.0 / (double)Cs);
      accumulate_round();
    }

    // own share -> shared memory; slice `rank` of the cluster sum; everybody pulls all slices
#pragma unroll
    for (int e = 0; e < CEW_EPT; ++e) {
      const int idx = tid + CEW_THREADS * e;
      if (idx < G) Gpart[idx] = gacc[e];
    }
    cluster.sync();
    const int lo = rank * slice, hi = lo + slice < G ? lo + slice : G;
    for (int idx = lo + tid; idx < hi; idx += CEW_THREADS) {
      double g = 0.0;
      for (int rk = 0; rk < nranks; ++rk) g += cluster.map_shared_rank(Gpart, rk)[idx];
      Gred[idx - lo] = g;
      if (a.g_last != nullptr && step == a.num_batches - 1) a.g_last[idx] = g;
    }
    cluster.sync();
    if (a.lr != 0.0) {
      for (int idx = tid; idx < G; idx += CEW_THREADS) {
        const int owner = idx / slice;
        const double g = cluster.map_shared_rank(Gred, owner)[idx - owner * slice];
        double* prm = idx < nh * N ? &Wsm[(idx / N) * CEW_LD + idx % N] : (idx < nh * N + N ? &bsm[idx - nh * N] : &csm[idx - nh * N - N]);
        *prm -= a.lr * g;
      }
    }
    __syncthreads();
    if (a.lr != 0.0 && step + 1 < a.num_batches) refresh_float();
  }
  cluster.sync();  // nobody leaves while its shares may still be read
  if (rank == 0 && a.lr != 0.0) {
    for (int idx = tid; idx < nh * N; idx += CEW_THREADS) a.W[idx] = Wsm[(idx / N) * CEW_LD + idx % N];
    for (int idx = tid; idx < N; idx += CEW_THREADS) a.b[idx] = bsm[idx];
    for (int idx = tid; idx < nh; idx += CEW_THREADS) a.c[idx] = csm[idx];
  }
}

// One cluster per launch: 16 CTAs (non-portable size) when the device can co-schedule them, else the portable 8.
template <typename Kernel>
static int launch_one_cluster(Kernel kernel, int threads, size_t smem, const CdEpochArgs& a, cudaStream_t st, int* cluster_cache) {
  constexpr size_t SMEM_MAX = 226 * 1024;  // 227 KB per block on sm_100 includes the kernels' 512 bytes of static shared memory
  if (smem > SMEM_MAX) return fail(QCB_ERR_UNSUPPORTED, "cd_epoch: layer too large for the fused epoch kernel");
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int& cluster_size = cluster_cache[device_slot()];  // cached per device
  if (cluster_size == 0) {
    QCB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX));
    cluster_size = CE_CLUSTER_PORTABLE;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
      int nclusters = 0;
      cfg.gridDim = dim3(CE_CLUSTER_MAX, 1, 1);
      attr[0].val.clusterDim.x = CE_CLUSTER_MAX;
      cfg.dynamicSmemBytes = SMEM_MAX;
      if (cudaOccupancyMaxActiveClusters(&nclusters, kernel, &cfg) == cudaSuccess && nclusters >= 1) cluster_size = CE_CLUSTER_MAX;
      cfg.dynamicSmemBytes = smem;
    }
    (void)cudaGetLastError();
  }
  cfg.gridDim = dim3(cluster_size, 1, 1);
  attr[0].val.clusterDim.x = cluster_size;
  QCB_CUDA(cudaLaunchKernelEx(&cfg, kernel, a));
  QCB_LAUNCH_CHECK();
  return QCB_OK;
}

template <int SW>
static int launch_cd_epoch(const CdEpochArgs& a, cudaStream_t st) {
  constexpr int NT = ce_threads(SW);
  static int cluster_on[QCB_MAX_DEVICES] = {0};
  return launch_one_cluster(cd_epoch_kernel<SW>, NT, ce_smem_doubles(a.N, a.nh, SW, NT / 32) * sizeof(double), a, st, cluster_on);
}

template <int SW>
static int launch_cd_epoch_ws(const CdEpochArgs& a, cudaStream_t st) {
  static int cluster_on[QCB_MAX_DEVICES] = {0};
  return launch_one_cluster(cd_epoch_ws_kernel<SW>, WS_THREADS, ws_smem_bytes<SW>(a.N, a.nh), a, st, cluster_on);
}

// 0: automatic (the warp-specialised kernel for layers of at most 32 units), 1: never the warp-specialised kernel
int g_cd_epoch_mode = 0;

static int launch_cd_epoch_wide(const CdEpochArgs& a, cudaStream_t st) {
  static int cluster_on[QCB_MAX_DEVICES] = {0};
  return launch_one_cluster(cd_epoch_wide_kernel, CEW_THREADS, cew_smem_doubles() * sizeof(double), a, st, cluster_on);
}

int run_cd_epoch(double* W, double* b, double* c, int N, int nh, const double* pos, int64_t B_total, int pos_batch, const double* neg,
                 int64_t C_total, int neg_batch, int num_batches, int k, double lr, uint64_t seed, uint64_t sweep0, double* g_last,
                 cudaStream_t st) {
  if (N < 1 || nh < 1 || N > 64 || nh > 64) return fail(QCB_ERR_UNSUPPORTED, "cd_epoch: needs 1 <= N, nh <= 64");
  if (W == nullptr || b == nullptr || c == nullptr || pos == nullptr || neg == nullptr || pos_batch < 1 || neg_batch < 1 ||
      num_batches < 0 || k < 0)
    return fail(QCB_ERR_BAD_ARG, "cd_epoch: bad arguments");
  if ((int64_t)(num_batches - 1) * pos_batch >= B_total && num_batches > 0)
    return fail(QCB_ERR_BAD_ARG, "cd_epoch: fewer positive rows than batches");
  if ((int64_t)(num_batches - 1) * neg_batch >= C_total && num_batches > 0)
    return fail(QCB_ERR_BAD_ARG, "cd_epoch: fewer negative rows than batches");
  if (sweep0 + (uint64_t)num_batches * (uint64_t)k >= (1ull << 32)) return fail(QCB_ERR_BAD_ARG, "cd_epoch: sweep counter overflow");
  if (num_batches == 0) return QCB_OK;
  CdEpochArgs a;
  a.W = W; a.b = b; a.c = c;
  a.N = N; a.nh = nh;
  a.pos = pos; a.B_total = B_total; a.pos_batch = pos_batch;
  a.neg = neg; a.C_total = C_total; a.neg_batch = neg_batch;
  a.num_batches = num_batches;
  a.k = k;
  a.lr = lr;
  a.key0 = (uint32_t)seed;
  a.key1 = (uint32_t)(seed >> 32);
  a.sweep0 = (uint32_t)sweep0;
  a.g_last = g_last;
  const int m = N > nh ? N : nh;
  if (g_cd_epoch_mode == 0 && k >= 1) {
    if (m <= 8) return launch_cd_epoch_ws<8>(a, st);
    if (m <= 16) return launch_cd_epoch_ws<16>(a, st);
    // 17..32 units: one chain per consumer warp, i.e. 128 slots against the single-role kernel's 256 (rows and chains are dealt
    // differently, so the sums are associated differently: agreement to rounding, not bit for bit); 1.48 against 2.09 ms per epoch at
    // N = nh = 32 all the same
    if (m <= 32) return launch_cd_epoch_ws<32>(a, st);
  }
  if (m <= 8) return launch_cd_epoch<8>(a, st);
  if (m <= 16) return launch_cd_epoch<16>(a, st);
  if (m <= 32) return launch_cd_epoch<32>(a, st);
  return launch_cd_epoch_wide(a, st);
}

}  // namespace qcb
