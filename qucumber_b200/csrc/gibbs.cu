// Fused k-sweep block Gibbs sampler for BinaryRBM / PurificationRBM on sm_100a.
//
// Replaces the reference's per-sweep chain of torch ops (rbm/binary_rbm.py:129-228,
// rbm/purification_rbm.py:196-337: matmul, add_, sigmoid_, clamp_, bernoulli — 10 to 17 launches
// and two HBM round trips of the chain state per sweep) with ONE launch for all k sweeps:
//
//  * a warp owns 8 chains; their visible/hidden states live in registers as bit masks for the
//    whole call (HBM is touched once to read v0 and once to write v_k);
//  * the pre-activations x = W v + c and x' = W^T h + b are FP64 tensor-core tiles
//    (mma.sync m8n8k4 f64, "DMMA"): A = the 8 chains' state bits as 0.0/1.0, B = W fragments
//    read from shared memory, accumulators start at the bias;
//  * W lives in shared memory for the whole call in *fragment order* (one conflict-free
//    LDS.128 per lane feeds two DMMAs).  A tiny pack kernel writes that image once per call into
//    the caller's workspace; every CTA pulls it with one 1-D TMA bulk copy (cp.async.bulk +
//    mbarrier complete_tx);
//  * uniforms are Philox4x32-10 draws made in-kernel (convention in include/qcb.h), or fed by
//    the caller for bit-exact replay of the reference;
//  * the Bernoulli decision u < sigmoid(x) is made with an exactness-preserving filter: a float
//    sigmoid and the top 23 bits of u decide whenever |u - p| > 2^-16; the rare remaining cases
//    (probability ~3e-5) take the exact FP64 path.  Decisions are identical to evaluating
//    u < 1/(1+exp(-x)) in FP64 for every unit.
//
// The PurificationRBM is the same kernel on the stacked matrix [W; U] with biases [c; d]
// (h | v and a | v are both "hidden-like"; v | h,a sums both contributions).
#include "common.cuh"

namespace qcb {

constexpr int GIBBS_THREADS = 256;
constexpr int GIBBS_WARPS = GIBBS_THREADS / 32;
constexpr int CHAINS_PER_WARP = 8;

struct GibbsArgs {
  const double* image;   // packed fragment image in global memory (workspace)
  uint32_t image_bytes;
  const double* v_in;    // (C, N) initial states (may alias v)
  double* v;             // (C, N) final states
  const double* uniforms;  // NULL or (k, C, nhid + N)
  int64_t C;
  int N, nhid;
  int k;
  uint32_t key0, key1;
  uint64_t chain_id0;
  uint32_t sweep0;
  int64_t num_groups;
};

// ---- pack kernel: builds the fragment-ordered shared-memory image ------------------------------
// Layout (doubles): Ih[4*VB][HB][32][2] | Iv[4*HB][VB][32][2] | biasH[16*HB] | biasV[16*VB]
// Ih feeds h|v: element (kk, gq, lane=(g,t), s) = Wst[hid = 16gq + 4(g/2) + 2s + (g%2)][vis = 4kk + t]
// Iv feeds v|h: element (kk, gq, lane, s)       = Wst[hid = 4kk + t][vis = 16gq + 4(g/2) + 2s + (g%2)]
// With this column permutation lane (g,t) ends up owning output units 16gq + 4t + {0,1,2,3} of
// chain g, i.e. exactly the four words of one Philox call.
__global__ void gibbs_pack_kernel(StackedParams p, int HB, int VB, double* __restrict__ img) {
  const int Hpad = 16 * HB, Npad = 16 * VB;
  const int nIh = Npad * Hpad;
  const int total = 2 * nIh + Hpad + Npad;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    double val = 0.0;
    if (idx < 2 * nIh) {
      const bool hside = idx < nIh;
      const int r = hside ? idx : idx - nIh;
      const int OB = hside ? HB : VB;
      const int s = r & 1;
      const int lane = (r >> 1) & 31;
      const int gq = (r >> 6) % OB;
      const int kk = (r >> 6) / OB;
      const int g = lane >> 2, t = lane & 3;
      const int outu = 16 * gq + 4 * (g >> 1) + 2 * s + (g & 1);
      const int inu = 4 * kk + t;
      const int hid = hside ? outu : inu;
      const int vis = hside ? inu : outu;
      if (hid < p.nhid() && vis < p.N) val = p.w(hid, vis);
    } else if (idx < 2 * nIh + Hpad) {
      const int hid = idx - 2 * nIh;
      if (hid < p.nhid()) val = p.hbias(hid);
    } else {
      const int vis = idx - 2 * nIh - Hpad;
      if (vis < p.N) val = p.b[vis];
    }
    img[idx] = val;
  }
}

// ---- Bernoulli decisions for the four units of one Philox call ----------------------------------
// Returns a 4-bit mask, bit w set iff u_w < sigmoid(x_w) evaluated exactly in FP64.
struct RngCtx {
  uint32_t c0, c1;   // chain id
  uint32_t key0, key1;
};

// Uniform convention (include/qcb.h): unit u of a layer whose first Philox call is q0 uses call
// q = q0 + 4 (u / 32) + (u % 16) / 4, output word u % 4 and half (u / 16) % 2 of that word:
//   h16  = 16-bit half of stream 0            (the only bits the fast path needs)
//   lo_a = word of stream 1 + 2 half,  lo_b = word of stream 2 + 2 half   (cold path only)
//   u53  = (h16 * 2^37 + lo_a * 2^5 + (lo_b >> 27)) * 2^-53
// so one Philox call serves EIGHT decisions.  The filter decides from h16 alone unless p lies within
// 2^-17 of the 2^-16-wide cell [h16, h16 + 1) * 2^-16 (probability 3e-5 per decision).
constexpr float FILTER_EPS = 0x1.0p-17f;
constexpr float CELL = 0x1.0p-16f;

__device__ __forceinline__ uint32_t stream_ctr(uint32_t stream, uint32_t q) { return (stream << 29) | q; }

// Cold path, kept out of line (one call site per half sweep) so the hot loop stays small: rebuilds
// the full 53-bit uniforms of the flagged units and compares them with the exact FP64 sigmoid.
// xs holds the lane's pre-activations (4 per 16-unit block), nibw its decisions (4 bits per block),
// slowmask the undecided ones.
__device__ __noinline__ uint32_t fixup_slow(const double* __restrict__ xs, uint32_t nibw, uint32_t slowmask, uint32_t c0,
                                            uint32_t c1, uint32_t key0, uint32_t key1, uint32_t sweep, uint32_t q0, int t) {
  while (slowmask != 0u) {
    const int bitpos = __ffs(slowmask) - 1;
    const int gq = bitpos >> 2;
    const uint32_t sb = (slowmask >> (4 * gq)) & 0xFu;
    slowmask &= ~(0xFu << (4 * gq));
    const uint32_t half = (uint32_t)(gq & 1), q = q0 + 4u * (uint32_t)(gq >> 1) + (uint32_t)t;
    uint32_t hi[4], la[4], lb[4];
    philox4x32_10(c0, c1, sweep, stream_ctr(0u, q), key0, key1, hi);
    philox4x32_10(c0, c1, sweep, stream_ctr(1u + 2u * half, q), key0, key1, la);
    philox4x32_10(c0, c1, sweep, stream_ctr(2u + 2u * half, q), key0, key1, lb);
#pragma unroll 1
    for (int w = 0; w < 4; ++w) {
      if (sb & (1u << w)) {
        const uint64_t h16 = (hi[w] >> (16u * (1u - half))) & 0xFFFFu;
        const uint64_t bits = (h16 << 37) | ((uint64_t)la[w] << 5) | (uint64_t)(lb[w] >> 27);
        const double u = (double)bits * 0x1.0p-53;
        const uint32_t m = 1u << (4 * gq + w);
        if (u < sigmoid_exact(xs[4 * gq + w])) nibw |= m; else nibw &= ~m;
      }
    }
  }
  return nibw;
}

// Fed-uniform mode (bit-exact replay of the reference): every decision is the exact FP64 comparison.
__device__ __noinline__ uint32_t decide_fed(const double* __restrict__ xs, const double* __restrict__ fed, int nblocks,
                                            int n_out_real, int t) {
  uint32_t nibw = 0u;
  for (int gq = 0; gq < nblocks; ++gq) {
    for (int w = 0; w < 4; ++w) {
      const int unit = 16 * gq + 4 * t + w;
      if (unit < n_out_real && fed[unit] < sigmoid_exact(xs[4 * gq + w])) nibw |= 1u << (4 * gq + w);
    }
  }
  return nibw;
}

// Fast decisions for the four units that use half `half` of the words r[0..3]: bit w of the result is
// set when u < p is certain, bit w of `slow` when the 16-bit cell cannot decide.
__device__ __forceinline__ uint32_t decide4_fast(const double (&x)[4], const uint32_t (&r)[4], uint32_t half, uint32_t& slow) {
  uint32_t nib = 0;
  slow = 0;
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    const float xf = (float)x[w];
    const float p = __fdividef(1.0f, 1.0f + __expf(-xf));
    const uint32_t h16 = half ? (r[w] & 0xFFFFu) : (r[w] >> 16);
    const float uf = __uint_as_float(0x3F800000u | (h16 << 7)) - 1.0f;  // u in [uf, uf + 2^-16)
    const bool sure1 = uf + (CELL + FILTER_EPS) <= p;
    const bool sure0 = uf >= p + FILTER_EPS;
    nib |= (sure1 ? 1u : 0u) << w;
    slow |= ((sure1 || sure0) ? 0u : 1u) << w;
  }
  return nib;
}

// ---- tensor-pipe ping-pong ------------------------------------------------------------------------
// Warps w and w+4 of a CTA share an SM sub-partition (and its DMMA pipe).  Left alone, warps that
// contend for the pipe finish their DMMA phases together and then all sit in their epilogues while
// the pipe idles (measured: 74 % DMMA utilisation).  The pair therefore takes turns: one runs its DMMA
// phase while the other runs its Bernoulli epilogue, handing over with two named barriers
// (bar.arrive by the finisher, bar.sync by the waiter; 64 = the two warps).
__device__ __forceinline__ void pipe_acquire(int warp) {
  asm volatile("bar.sync %0, 64;" ::"r"(1 + 2 * (warp & 3) + (warp >> 2)) : "memory");
}
__device__ __forceinline__ void pipe_release(int warp) {
  asm volatile("bar.arrive %0, 64;" ::"r"(1 + 2 * (warp & 3) + ((warp >> 2) ^ 1)) : "memory");
}

// ---- one conditional update (all output units of 8 chains) --------------------------------------
// KB_IN / OB: number of 16-unit blocks on the input / output side.
template <int KB_IN, int OB>
__device__ __forceinline__ void half_sweep(const uint32_t (&min)[(KB_IN + 1) / 2], uint32_t (&mout)[(OB + 1) / 2],
                                           const double2* __restrict__ img, const double* __restrict__ bias,
                                           int n_out_real, const RngCtx& rc, uint32_t sweep, uint32_t q0,
                                           const double* __restrict__ fed /* this chain's block or NULL */,
                                           int lane, int lane_warp) {
  const int t = lane & 3;
  double acc[2 * OB][2];
#pragma unroll
  for (int gq = 0; gq < OB; ++gq) {
    const double2 b01 = *reinterpret_cast<const double2*>(bias + 16 * gq + 4 * t);
    const double2 b23 = *reinterpret_cast<const double2*>(bias + 16 * gq + 4 * t + 2);
    acc[2 * gq][0] = b01.x; acc[2 * gq][1] = b01.y;
    acc[2 * gq + 1][0] = b23.x; acc[2 * gq + 1][1] = b23.y;
  }
  uint32_t mt[(KB_IN + 1) / 2];
#pragma unroll
  for (int w = 0; w < (KB_IN + 1) / 2; ++w) mt[w] = min[w] >> t;

  constexpr int KK = 4 * KB_IN;
  pipe_acquire(lane_warp);
  // B fragments are double-buffered in registers: the loads of k-step kk+1 are issued before the
  // DMMAs of k-step kk, so a load never has to wait for the DMMA that last read its destination
  // registers (the write-after-read wait that otherwise serialises a single warp's DMMA stream).
  double2 bf[2][OB];
#pragma unroll
  for (int gq = 0; gq < OB; ++gq) bf[0][gq] = img[gq * 32 + lane];
#pragma unroll
  for (int kk = 0; kk < KK; ++kk) {
    if (kk + 1 < KK) {
      const double2* row = img + (size_t)(kk + 1) * OB * 32 + lane;
#pragma unroll
      for (int gq = 0; gq < OB; ++gq) bf[(kk + 1) & 1][gq] = row[gq * 32];
    }
    const uint32_t bit = (mt[kk / 8] >> (4 * (kk % 8))) & 1u;
    const double a = __hiloint2double((int)(bit * 0x3FF00000u), 0);
#pragma unroll
    for (int gq = 0; gq < OB; ++gq) {
      dmma884(acc[2 * gq][0], acc[2 * gq][1], a, bf[kk & 1][gq].x);
      dmma884(acc[2 * gq + 1][0], acc[2 * gq + 1][1], a, bf[kk & 1][gq].y);
    }
  }
  pipe_release(lane_warp);

  // Bernoulli decisions: nibw collects 4 bits per 16-unit block, slowmask the undecided ones
  uint32_t nibw = 0u, slowmask = 0u;
  if (fed == nullptr) {
#pragma unroll
    for (int pq = 0; pq < (OB + 1) / 2; ++pq) {
      // one Philox call per pair of 16-unit blocks: its four words carry two 16-bit uniforms each
      uint32_t r[4];
      philox4x32_10(rc.c0, rc.c1, sweep, stream_ctr(0u, q0 + 4 * pq + t), rc.key0, rc.key1, r);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int gq = 2 * pq + half;
        if (gq < OB) {
          int nvalid = n_out_real - (16 * gq + 4 * t);
          nvalid = nvalid < 0 ? 0 : (nvalid > 4 ? 4 : nvalid);
          const uint32_t valid = (1u << nvalid) - 1u;
          const double x[4] = {acc[2 * gq][0], acc[2 * gq][1], acc[2 * gq + 1][0], acc[2 * gq + 1][1]};
          uint32_t slow;
          const uint32_t nib = decide4_fast(x, r, (uint32_t)half, slow);
          nibw |= (nib & valid) << (4 * gq);
          slowmask |= (slow & valid) << (4 * gq);
        }
      }
    }
  }
  if (fed != nullptr || slowmask != 0u) {  // rare (probability ~3e-5 per decision) unless uniforms are fed
    double xs[4 * OB];
#pragma unroll
    for (int gq = 0; gq < OB; ++gq) {
      xs[4 * gq] = acc[2 * gq][0]; xs[4 * gq + 1] = acc[2 * gq][1];
      xs[4 * gq + 2] = acc[2 * gq + 1][0]; xs[4 * gq + 3] = acc[2 * gq + 1][1];
    }
    if (fed != nullptr) nibw = decide_fed(xs, fed, OB, n_out_real, t);
    else nibw = fixup_slow(xs, nibw, slowmask, rc.c0, rc.c1, rc.key0, rc.key1, sweep, q0, t);
  }
#pragma unroll
  for (int w = 0; w < (OB + 1) / 2; ++w) mout[w] = 0u;
#pragma unroll
  for (int gq = 0; gq < OB; ++gq) mout[gq / 2] |= ((nibw >> (4 * gq)) & 0xFu) << (16 * (gq % 2) + 4 * t);
#pragma unroll
  for (int w = 0; w < (OB + 1) / 2; ++w) {
    mout[w] |= __shfl_xor_sync(0xffffffffu, mout[w], 1);
    mout[w] |= __shfl_xor_sync(0xffffffffu, mout[w], 2);
  }
}

template <int HB, int VB>
__global__ void __launch_bounds__(GIBBS_THREADS) gibbs_kernel(const GibbsArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* smem = reinterpret_cast<double*>(smem_raw);
  constexpr int Hpad = 16 * HB, Npad = 16 * VB;
  constexpr int nI = Hpad * Npad;
  const double2* Ih = reinterpret_cast<const double2*>(smem);
  const double2* Iv = reinterpret_cast<const double2*>(smem + nI);
  const double* biasH = smem + 2 * nI;
  const double* biasV = biasH + Hpad;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + 2 * nI + Hpad + Npad);  // [0]: image, [1 + warp]: that warp's state tile
  double* tiles = reinterpret_cast<double*>(smem_raw + align_up((size_t)(2 * nI + Hpad + Npad) * 8 + 8 * (1 + GIBBS_WARPS), 128));

  // Stage the packed parameters with one TMA bulk copy per <=32 KB chunk.
  if (threadIdx.x < 1 + GIBBS_WARPS) mbar_init(mbar + threadIdx.x, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(mbar, a.image_bytes);
    const char* src = reinterpret_cast<const char*>(a.image);
    char* dst = reinterpret_cast<char*>(smem);
    for (uint32_t off = 0; off < a.image_bytes; off += 32768u) {
      const uint32_t n = a.image_bytes - off < 32768u ? a.image_bytes - off : 32768u;
      tma_bulk_g2s(dst + off, src + off, n, mbar);
    }
  }
  mbar_wait(mbar, 0);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2;
  constexpr int MV = (VB + 1) / 2, MH = (HB + 1) / 2;
  const int nunits = a.nhid + a.N;
  const uint32_t qv0 = (uint32_t)(4 * ((a.nhid + 31) / 32));  // first Philox call of the visible layer

  // A group's 8 chain rows are one contiguous span of v_in: the warp's tile is fetched by one TMA bulk copy issued
  // a whole group ahead (while the previous group is being swept), so its latency never shows.
  double* tile = tiles + (size_t)warp * (CHAINS_PER_WARP * Npad);
  uint64_t* tbar = mbar + 1 + warp;
  uint32_t tphase = 0u;
  const bool bulk_base_ok = (reinterpret_cast<uintptr_t>(a.v_in) & 15u) == 0u;
  auto tile_doubles = [&](int64_t gi) -> uint32_t {
    const int64_t left = a.C - gi * CHAINS_PER_WARP;
    return left <= 0 ? 0u : (uint32_t)(left < CHAINS_PER_WARP ? left : CHAINS_PER_WARP) * (uint32_t)a.N;
  };
  auto bulk_fetch = [&](int64_t gi) {  // true when the copy was handed to the TMA unit
    const uint32_t bytes = tile_doubles(gi) * 8u;
    if (bytes == 0u || !bulk_base_ok || (bytes & 15u) != 0u) return false;
    if (lane == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic reads of the tile vs. the async write
      mbar_expect_tx(tbar, bytes);
      tma_bulk_g2s(tile, a.v_in + gi * CHAINS_PER_WARP * (int64_t)a.N, bytes, tbar);
    }
    return true;
  };

  // every warp runs the same number of iterations (groups past the end are dummies) so that the
  // ping-pong partners stay in lock step; role-1 warps let their partners start first
  if (warp >> 2) pipe_release(warp);
  const int64_t stride = (int64_t)gridDim.x * GIBBS_WARPS;
  const int64_t iters = (a.num_groups + stride - 1) / stride;
  bool fetched = iters > 0 && bulk_fetch((int64_t)blockIdx.x * GIBBS_WARPS + warp);
  for (int64_t it = 0; it < iters; ++it) {
    const int64_t gi = it * stride + (int64_t)blockIdx.x * GIBBS_WARPS + warp;
    const int64_t chain_base = gi * CHAINS_PER_WARP;
    const int64_t my_chain = chain_base + g;
    const bool my_valid = my_chain < a.C;

    // v0 tile (shared memory) -> bit masks
    const uint32_t ndbl = tile_doubles(gi);
    const int nch = (int)(ndbl / (uint32_t)a.N);
    if (fetched) {
      mbar_wait(tbar, tphase);
      tphase ^= 1u;
    } else if (ndbl != 0u) {  // unaligned view or odd tile size: plain copy
      for (uint32_t idx = lane; idx < ndbl; idx += 32) tile[idx] = a.v_in[chain_base * a.N + idx];
      __syncwarp();
    }
    uint32_t mv[MV], mh[MH];
#pragma unroll
    for (int w = 0; w < MV; ++w) mv[w] = 0u;
#pragma unroll
    for (int cc = 0; cc < CHAINS_PER_WARP; ++cc) {
#pragma unroll
      for (int w = 0; w < MV; ++w) {
        const int j = w * 32 + lane;
        unsigned long long bits = 0ull;  // integer image of the double: the != 0 test stays off the FP64 pipe
        if (cc < nch && j < a.N) bits = reinterpret_cast<const unsigned long long*>(tile)[cc * a.N + j];
        const uint32_t bal = __ballot_sync(0xffffffffu, (bits << 1) != 0ull);
        if (g == cc) mv[w] = bal;
      }
    }
    __syncwarp();
    fetched = it + 1 < iters && bulk_fetch(gi + stride);

    RngCtx rc;
    const uint64_t cid = a.chain_id0 + (uint64_t)my_chain;
    rc.c0 = (uint32_t)cid; rc.c1 = (uint32_t)(cid >> 32);
    rc.key0 = a.key0; rc.key1 = a.key1;

    for (int s = 0; s < a.k; ++s) {
      const double* fed_h = nullptr;
      const double* fed_v = nullptr;
      if (a.uniforms != nullptr) {
        const int64_t ch = my_valid ? my_chain : 0;
        fed_h = a.uniforms + ((int64_t)s * a.C + ch) * nunits;
        fed_v = fed_h + a.nhid;
      }
      const uint32_t sweep = a.sweep0 + (uint32_t)s;
      half_sweep<VB, HB>(mv, mh, Ih, biasH, a.nhid, rc, sweep, 0u, fed_h, lane, warp);
      half_sweep<HB, VB>(mh, mv, Iv, biasV, a.N, rc, sweep, qv0, fed_v, lane, warp);
    }

    // bit masks -> v_k (coalesced row writes)
#pragma unroll
    for (int cc = 0; cc < CHAINS_PER_WARP; ++cc) {
      const int64_t ch = chain_base + cc;
#pragma unroll
      for (int w = 0; w < MV; ++w) {
        const uint32_t word = __shfl_sync(0xffffffffu, mv[w], cc * 4);
        const int j = w * 32 + lane;
        if (ch < a.C && j < a.N) a.v[ch * a.N + j] = ((word >> lane) & 1u) ? 1.0 : 0.0;
      }
    }
  }
}

static size_t gibbs_image_doubles(int HB, int VB) { return (size_t)2 * 256 * HB * VB + 16 * HB + 16 * VB; }

static int pick_blocks(int units) {
  if (units <= 16) return 1;
  if (units <= 32) return 2;
  if (units <= 64) return 4;
  if (units <= 128) return 8;
  return 0;
}

size_t gibbs_imma_workspace_bytes(int N, int nhid);
size_t gibbs_umma_workspace_bytes(int N, int nhid);
size_t gibbs_umma_wide_workspace_bytes(int N, int nhid);
int64_t gibbs_umma_wide_wave_chains(int N, int nhid);
int try_gibbs_umma_wide(const StackedParams& p, const ChainIO& io, int64_t C, int k, uint64_t seed, uint64_t chain_id0, uint64_t sweep0,
                        void* ws, size_t ws_bytes, int* handled, cudaStream_t st);

size_t gibbs_table_workspace_bytes(int N, int nhid);
int64_t gibbs_table_wave_chains();
int try_gibbs_table(const StackedParams& p, const ChainIO& io, int64_t C, int k, uint64_t seed, uint64_t chain_id0, uint64_t sweep0,
                    void* ws, size_t ws_bytes, int* handled, cudaStream_t st);

size_t gibbs_large_workspace_bytes(int N, int nhid, int64_t C);
int run_gibbs_large(const StackedParams& p, const double* v_in, double* v, int64_t C, int k, uint64_t seed, uint64_t chain_id0,
                    uint64_t sweep0, const double* uniforms, void* ws, size_t ws_bytes, cudaStream_t st);
static bool gibbs_fused_shape(int N, int nhid);
size_t gibbs_workspace_bytes(int N, int nhid);

// rows: the number of chains (only the composed path for large layers needs scratch that grows with it)
size_t gibbs_workspace_bytes_rows(int N, int nhid, int64_t rows) {
  if (!gibbs_fused_shape(N, nhid)) {
    const size_t a = gibbs_large_workspace_bytes(N, nhid, rows), w = gibbs_umma_wide_workspace_bytes(N, nhid);
    return a > w ? a : w;
  }
  return gibbs_workspace_bytes(N, nhid);
}

size_t gibbs_workspace_bytes(int N, int nhid) {
  const int HB = pick_blocks(nhid), VB = pick_blocks(N);
  if (HB == 0 || VB == 0) return 0;
  const size_t a = align_up(gibbs_image_doubles(HB, VB) * sizeof(double), 256), b = gibbs_imma_workspace_bytes(N, nhid);
  size_t c = gibbs_umma_workspace_bytes(N, nhid);
  const size_t w = gibbs_umma_wide_workspace_bytes(N, nhid), tb = gibbs_table_workspace_bytes(N, nhid);
  if (w > c) c = w;
  if (tb > c) c = tb;
  return a > b ? (a > c ? a : c) : (b > c ? b : c);
}

template <int HB, int VB>
static int launch_gibbs(const GibbsArgs& a, cudaStream_t st) {
  // image + mbarriers + one 8-chain state tile per warp
  const size_t smem = align_up(gibbs_image_doubles(HB, VB) * sizeof(double) + 8 * (1 + GIBBS_WARPS), 128) +
                      (size_t)GIBBS_WARPS * CHAINS_PER_WARP * 16 * VB * sizeof(double);
  static int blocks_on[QCB_MAX_DEVICES] = {0};  // 0: this device is not configured yet
  int& blocks_per_sm = blocks_on[device_slot()];
  if (blocks_per_sm == 0) {
    int occ = 0;
    QCB_CUDA(cudaFuncSetAttribute(gibbs_kernel<HB, VB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    QCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gibbs_kernel<HB, VB>, GIBBS_THREADS, smem));
    if (occ < 1) return fail(QCB_ERR_CUDA, "gibbs kernel does not fit on an SM (smem %zu B)", smem);
    blocks_per_sm = occ;
  }
  const int64_t need = (a.num_groups + GIBBS_WARPS - 1) / GIBBS_WARPS;
  const int64_t cap = (int64_t)sm_count() * blocks_per_sm;
  const int grid = (int)(need < cap ? need : cap);
  gibbs_kernel<HB, VB><<<grid, GIBBS_THREADS, smem, st>>>(a);
  QCB_LAUNCH_CHECK();
  return QCB_OK;
}

// 0: automatic (integer-exact filter where it applies: tcgen05 / TMEM kernel for two layers of 33..64 units and at least
// GIBBS_UMMA_MIN_K sweeps, mma.sync kernel otherwise), 1: always the FP64 DMMA kernel, 2: tcgen05 kernel wherever its shape
// fits, 3: never the tcgen05 kernel.  All of them draw the same numbers and make the same decisions.
int g_gibbs_mode = 0;
constexpr int GIBBS_UMMA_MIN_K = 3;  // below that the mma.sync kernel's TMA-prefetched state tiles win (k = 1: 0.29 vs 0.35 ms)
static bool use_umma(int k) { return g_gibbs_mode == 2 || (g_gibbs_mode == 0 && k >= GIBBS_UMMA_MIN_K); }
// The table kernel (gibbs_table.cu: one thread per chain, nibble tables in shared memory) takes layers of at most 16 units in
// the automatic mode; mode 5 sends every shape it fits through it whatever the mode's other choices.
constexpr int GIBBS_TABLE_MIN_K = 1;  // measured faster than the mma.sync kernel from one sweep on (2^20 chains, N = nh = 10: 0.084 vs 0.141 ms)
static bool use_table(int N, int nhid, int k) {
  if (gibbs_table_workspace_bytes(N, nhid) == 0) return false;
  return g_gibbs_mode == 5 || (g_gibbs_mode == 0 && k >= GIBBS_TABLE_MIN_K);
}
// The general tcgen05 kernel (gibbs_umma_wide.cu: any two layers of up to 128 units).  Layers of 65..128 units have no mma.sync
// filter kernel: it takes them in modes 0 and 2; mode 4 sends every shape it fits through it (measurement / test hook).
static bool use_umma_wide(int N, int nhid) {
  if (gibbs_umma_wide_workspace_bytes(N, nhid) == 0) return false;
  if (g_gibbs_mode == 4) return true;
  return (g_gibbs_mode == 0 || g_gibbs_mode == 2) && (N > 64 || nhid > 64);
}

size_t gibbs_imma_workspace_bytes(int N, int nhid);
int try_gibbs_imma(const StackedParams& p, const ChainIO& io, int64_t C, int k, uint64_t seed, uint64_t chain_id0,
                   uint64_t sweep0, void* ws, size_t ws_bytes, unsigned long long* dbg, int* handled, cudaStream_t st);
int run_pack_states(const double* v, int64_t rows, int N, uint8_t* bits, cudaStream_t st);
int run_unpack_states(const uint8_t* bits, int64_t rows, int N, double* v, cudaStream_t st);
int run_philox_init_states(double* v, int64_t C, int N, uint64_t seed, uint64_t chain_id0, uint64_t sweep0, cudaStream_t st);
int64_t gibbs_imma_wave_chains(int N, int nhid);
size_t gibbs_umma_workspace_bytes(int N, int nhid);
int64_t gibbs_umma_wave_chains();
int try_gibbs_umma(const StackedParams& p, const ChainIO& io, int64_t C, int k, uint64_t seed, uint64_t chain_id0, uint64_t sweep0,
                   void* ws, size_t ws_bytes, int* handled, cudaStream_t st);

static bool gibbs_fused_shape(int N, int nhid) {
  const int HB = pick_blocks(nhid), VB = pick_blocks(N);
  return HB != 0 && VB != 0 && !(HB == 8 && VB == 8);
}

static int gibbs_shape_check(const StackedParams& p, int* HBo, int* VBo) {
  const int HB = pick_blocks(p.nhid()), VB = pick_blocks(p.N);
  if (HB == 0 || VB == 0 || (HB == 8 && VB == 8))
    return fail(QCB_ERR_UNSUPPORTED, "gibbs: N=%d, nh+na=%d exceeds the shared-memory resident limit (<=128 each, not both >64)",
                p.N, p.nhid());
  *HBo = HB;
  *VBo = VB;
  return QCB_OK;
}

// Workspace of a call with compact or in-kernel-initialised states: the fragment image plus, for shapes or modes the
// integer-filter kernel does not take, one float64 copy of the chain state.
size_t gibbs_io_workspace_bytes(int N, int nhid, int64_t C) {
  const size_t img = gibbs_workspace_bytes_rows(N, nhid, C);
  const bool imma = gibbs_imma_workspace_bytes(N, nhid) != 0 || use_umma_wide(N, nhid);
  return img + (imma && g_gibbs_mode != 1 ? 0 : align_up((size_t)(C > 0 ? C : 0) * N * sizeof(double), 256));
}

// Chains one full wave of the persistent grid processes (slicing a host pipeline at multiples of it avoids ragged tails)
int64_t gibbs_wave_chains(int N, int nhid) {
  if ((g_gibbs_mode == 0 || g_gibbs_mode == 2) && gibbs_umma_workspace_bytes(N, nhid) != 0) return gibbs_umma_wave_chains();
  if (use_umma_wide(N, nhid)) return gibbs_umma_wide_wave_chains(N, nhid);
  if (use_table(N, nhid, GIBBS_TABLE_MIN_K)) return gibbs_table_wave_chains();
  const int64_t w = gibbs_imma_wave_chains(N, nhid);
  return w > 0 ? w : (int64_t)sm_count() * 2 * GIBBS_WARPS * CHAINS_PER_WARP;
}

static int run_gibbs_rows(const StackedParams& p, const double* v_in, double* v, int64_t C, int k, uint64_t seed, uint64_t chain_id0,
                          uint64_t sweep0, const double* uniforms, void* ws, size_t ws_bytes, int allow_imma, cudaStream_t st);

// General entry: any source (float64 rows, packed rows, in-kernel Bernoulli(1/2)) and any sink (float64 rows, packed rows).
int run_gibbs_io(const StackedParams& p, const ChainIO& io, int64_t C, int k, uint64_t seed, uint64_t chain_id0, uint64_t sweep0,
                 void* ws, size_t ws_bytes, cudaStream_t st) {
  int rc = QCB_OK;
  const size_t img_bytes = gibbs_workspace_bytes_rows(p.N, p.nhid(), C);
  if (ws == nullptr || ws_bytes < img_bytes) return fail(QCB_ERR_BAD_ARG, "gibbs: workspace too small (%zu < %zu)", ws_bytes, img_bytes);
  if (C == 0) return QCB_OK;
  if (sweep0 + (uint64_t)k > (1ull << 32))
    return fail(QCB_ERR_BAD_ARG, "gibbs: sweep0 + k exceeds the 32-bit sweep counter of the Philox convention (use a new seed)");
  if (use_umma(k)) {
    int handled = 0;
    rc = try_gibbs_umma(p, io, C, k, seed, chain_id0, sweep0, ws, ws_bytes, &handled, st);
    if (rc != QCB_OK || handled) return rc;
  }
  if (use_umma_wide(p.N, p.nhid())) {
    int handled = 0;
    rc = try_gibbs_umma_wide(p, io, C, k, seed, chain_id0, sweep0, ws, ws_bytes, &handled, st);
    if (rc != QCB_OK || handled) return rc;
  }
  if (use_table(p.N, p.nhid(), k)) {
    int handled = 0;
    rc = try_gibbs_table(p, io, C, k, seed, chain_id0, sweep0, ws, ws_bytes, &handled, st);
    if (rc != QCB_OK || handled) return rc;
  }
  if (g_gibbs_mode != 1) {
    int handled = 0;
    rc = try_gibbs_imma(p, io, C, k, seed, chain_id0, sweep0, ws, ws_bytes, nullptr, &handled, st);
    if (rc != QCB_OK || handled) return rc;
  }
  // FP64 kernel: float64 rows in and out; compact / generated states are converted around it
  double* rows = io.v_out;
  if (rows == nullptr) {
    const size_t need = img_bytes + align_up((size_t)C * p.N * sizeof(double), 256);
    if (ws_bytes < need) return fail(QCB_ERR_BAD_ARG, "gibbs: workspace too small for the float64 staging rows (%zu < %zu)", ws_bytes, need);
    rows = reinterpret_cast<double*>(static_cast<char*>(ws) + img_bytes);
  }
  const double* src = io.v_in;
  if (io.philox_init) {
    rc = run_philox_init_states(rows, C, p.N, seed, chain_id0, sweep0, st);
    src = rows;
  } else if (io.bits_in != nullptr) {
    rc = run_unpack_states(io.bits_in, C, p.N, rows, st);
    src = rows;
  }
  if (rc != QCB_OK) return rc;
  rc = run_gibbs_rows(p, src, rows, C, k, seed, chain_id0, sweep0, nullptr, ws, img_bytes, 0, st);
  if (rc != QCB_OK) return rc;
  if (io.bits_out != nullptr) rc = run_pack_states(rows, C, p.N, io.bits_out, st);
  return rc;
}

int run_gibbs(const StackedParams& p, const double* v_in, double* v, int64_t C, int k, uint64_t seed, uint64_t chain_id0, uint64_t sweep0,
              const double* uniforms, void* ws, size_t ws_bytes, cudaStream_t st) {
  return run_gibbs_rows(p, v_in, v, C, k, seed, chain_id0, sweep0, uniforms, ws, ws_bytes, 1, st);
}

static int run_gibbs_rows(const StackedParams& p, const double* v_in, double* v, int64_t C, int k, uint64_t seed, uint64_t chain_id0,
                          uint64_t sweep0, const double* uniforms, void* ws, size_t ws_bytes, int allow_imma, cudaStream_t st) {
  const int nhid = p.nhid();
  if (uniforms == nullptr && allow_imma && k >= 1 && C > 0 && use_umma_wide(p.N, nhid) && sweep0 + (uint64_t)k <= (1ull << 32)) {
    int handled = 0;
    ChainIO io;
    io.v_in = v_in;
    io.v_out = v;
    const int rcw = try_gibbs_umma_wide(p, io, C, k, seed, chain_id0, sweep0, ws, ws_bytes, &handled, st);
    if (rcw != QCB_OK || handled) return rcw;
  }
  if (!gibbs_fused_shape(p.N, nhid)) {  // layers beyond the fused kernels: one affine + Bernoulli pass per conditional update
    if (sweep0 + (uint64_t)k > (1ull << 32))
      return fail(QCB_ERR_BAD_ARG, "gibbs: sweep0 + k exceeds the 32-bit sweep counter of the Philox convention (use a new seed)");
    return run_gibbs_large(p, v_in, v, C, k, seed, chain_id0, sweep0, uniforms, ws, ws_bytes, st);
  }
  int HB, VB;
  int rc0 = gibbs_shape_check(p, &HB, &VB);
  if (rc0 != QCB_OK) return rc0;
  const size_t need = gibbs_workspace_bytes(p.N, nhid);
  if (ws == nullptr || ws_bytes < need) return fail(QCB_ERR_BAD_ARG, "gibbs: workspace too small (%zu < %zu)", ws_bytes, need);
  if (C == 0) return QCB_OK;
  if (k == 0) {  // no sweep: v_k is the initial state (out-of-place callers still expect it in v)
    if (v != v_in) QCB_CUDA(cudaMemcpyAsync(v, v_in, (size_t)C * p.N * sizeof(double), cudaMemcpyDeviceToDevice, st));
    return QCB_OK;
  }
  if (sweep0 + (uint64_t)k > (1ull << 32))
    return fail(QCB_ERR_BAD_ARG, "gibbs: sweep0 + k exceeds the 32-bit sweep counter of the Philox convention (use a new seed)");
  if (uniforms == nullptr && g_gibbs_mode != 1 && allow_imma) {
    int handled = 0;
    ChainIO io;
    io.v_in = v_in;
    io.v_out = v;
    if (use_umma(k)) {
      const int rcu = try_gibbs_umma(p, io, C, k, seed, chain_id0, sweep0, ws, ws_bytes, &handled, st);
      if (rcu != QCB_OK || handled) return rcu;
    }
    if (use_table(p.N, nhid, k)) {
      const int rct = try_gibbs_table(p, io, C, k, seed, chain_id0, sweep0, ws, ws_bytes, &handled, st);
      if (rct != QCB_OK || handled) return rct;
    }
    const int rc = try_gibbs_imma(p, io, C, k, seed, chain_id0, sweep0, ws, ws_bytes, nullptr, &handled, st);
    if (rc != QCB_OK || handled) return rc;
  }

  double* img = reinterpret_cast<double*>(ws);
  const int total = (int)gibbs_image_doubles(HB, VB);
  gibbs_pack_kernel<<<(total + 255) / 256, 256, 0, st>>>(p, HB, VB, img);
  QCB_LAUNCH_CHECK();

  GibbsArgs a;
  a.image = img;
  a.image_bytes = (uint32_t)(total * sizeof(double));
  a.v_in = v_in;
  a.v = v;
  a.uniforms = uniforms;
  a.C = C;
  a.N = p.N;
  a.nhid = nhid;
  a.k = k;
  a.key0 = (uint32_t)seed;
  a.key1 = (uint32_t)(seed >> 32);
  a.chain_id0 = chain_id0;
  a.sweep0 = (uint32_t)sweep0;
  a.num_groups = (C + CHAINS_PER_WARP - 1) / CHAINS_PER_WARP;

#define QCB_GIBBS_CASE(H, V) \
  if (HB == H && VB == V) return launch_gibbs<H, V>(a, st);
  QCB_GIBBS_CASE(1, 1) QCB_GIBBS_CASE(1, 2) QCB_GIBBS_CASE(1, 4) QCB_GIBBS_CASE(1, 8)
  QCB_GIBBS_CASE(2, 1) QCB_GIBBS_CASE(2, 2) QCB_GIBBS_CASE(2, 4) QCB_GIBBS_CASE(2, 8)
  QCB_GIBBS_CASE(4, 1) QCB_GIBBS_CASE(4, 2) QCB_GIBBS_CASE(4, 4) QCB_GIBBS_CASE(4, 8)
  QCB_GIBBS_CASE(8, 1) QCB_GIBBS_CASE(8, 2) QCB_GIBBS_CASE(8, 4)
#undef QCB_GIBBS_CASE
  return fail(QCB_ERR_UNSUPPORTED, "gibbs: no kernel for blocks (%d, %d)", HB, VB);
}

}  // namespace qcb
