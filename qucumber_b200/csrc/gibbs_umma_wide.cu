// Block Gibbs with the integer-exact filter on tcgen05 / TMEM for layers of 65..128 units (either side, or both):
// BinaryRBM.gibbs_steps (rbm/binary_rbm.py:200-228) / PurificationRBM.gibbs_steps (rbm/purification_rbm.py:309-337).
// Same quantisation, digits, margins, Philox convention and decisions as gibbs_umma.cu / gibbs_imma.cu — hence the same
// samples bit for bit as the FP64 kernel and the composed path these shapes took before (and as the oracle).
//
// Differences from gibbs_umma.cu (two layers of 33..64 units):
//  * layers are padded to PV, PH in {64, 128}; a conditional update with PI input and PO output units is PI / 16
//    tcgen05.mma (M = 128 chains, N = 2 PO, K = 32): A = [v ; 128 v] (2 PI bytes per chain), B rows 0..PO-1 = [d0 ; 2 d1],
//    rows PO..2 PO - 1 = [d2 ; 0], accumulators acc01 in TMEM columns 0..PO-1 and acc2 in columns PO..2 PO - 1;
//  * a CTA of 1 024 threads runs TWO independent tiles of 128 chains (threads 0..511 and 512..1023) around ONE copy of the
//    weight images (up to 2 x 64 KB) — one tile's MMAs run under the other tile's Bernoulli epilogue; each tile has its own
//    arrival counter, mbarrier, named barrier and 256 TMEM columns;
//  * a thread owns one chain and the 32-unit group g = (warp % 16) / 4 of either layer; groups beyond a layer's padded width
//    (and groups of padding only) skip the decisions of that update but keep the arrive / wait protocol.
// Units are kept in "p-order" inside the kernel (8 units served by one Philox call = 8 consecutive TMEM columns):
// unit u = 32 g + 16 h + 4 c + w  <->  p = 32 g + 8 c + 4 h + w.
#include "common.cuh"
#include <cstdlib>

namespace qcb {

constexpr int GW_TILE = 128;            // chains per tile (M of the MMA)
constexpr int GW_THREADS = 1024;        // per CTA: 8 / GMAX tiles of 128 GMAX threads (GMAX = 32-unit groups of the wider layer)
constexpr int GW_A_LBO = 128 * 16;      // A operand: byte distance of consecutive 16-byte K chunks (128 rows, no swizzle)
constexpr uint32_t GW_TMEM_COLS = 512;  // tiles x 64 GMAX columns
#ifndef GW_PREFETCH
#define GW_PREFETCH 0                   // Philox calls of an update drawn before the MMA-completion wait (measured: no gain here, spills at 64 registers)
#endif

struct GwHeader {
  float inv_scale;   // -log2(e) * 2^-s
  float margin_h, margin_v;
  float dx_h, dx_v;
  int s;
  int degenerate;
  int pad;
};
static_assert(sizeof(GwHeader) == 32, "header layout");

// image: header | biasH[128] | biasV[128] | Bh (4 PV PH bytes) | Bv (4 PV PH bytes)
constexpr size_t GW_BIAS_OFF = 32;
constexpr size_t GW_B_OFF = 2048;
__host__ __device__ inline size_t gw_b_bytes(int PV, int PH) { return (size_t)4 * PV * PH; }
__host__ __device__ inline size_t gw_image_bytes(int PV, int PH) { return GW_B_OFF + 2 * gw_b_bytes(PV, PH); }

struct GwArgs {
  const unsigned char* image;
  StackedParams p;        // original FP64 parameters (cold path)
  const double* v_in;
  double* v;
  const uint8_t* bits_in;
  uint8_t* bits_out;
  int in_mode;            // 0: v_in, 1: bits_in, 2: Philox initial state
  int PV, PH;             // padded layer widths (64 or 128)
  int64_t C;
  int k;
  uint32_t key0, key1;
  uint64_t chain_id0;
  uint32_t sweep0;
  uint32_t rk[20];        // Philox round keys
};

// within a 32-unit group: u5 = 16 h + 4 c + w  <->  p5 = 8 c + 4 h + w
__host__ __device__ inline int gw_u5_of_p5(int p) { return (((p >> 2) & 1) << 4) | (((p >> 3) & 3) << 2) | (p & 3); }
__host__ __device__ inline int gw_p5_of_u5(int u) { return (((u >> 2) & 3) << 3) | (((u >> 4) & 1) << 2) | (u & 3); }
__host__ __device__ inline int gw_unit_of_p(int p) { return (p & ~31) | gw_u5_of_p5(p & 31); }

constexpr long long GW_WQ_MAX = 127ll * 32768 + 63 * 256 + 127;

// ---- pack, step 1 (one CTA): scale, margins, biases -----------------------------------------------------------------------
__global__ void __launch_bounds__(1024) gibbs_pack_wide_header_kernel(StackedParams p, unsigned char* __restrict__ img) {
  __shared__ double red[3][32];
  const int H = p.nhid(), N = p.N;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double mw = 0.0;
  for (int idx = threadIdx.x; idx < H * N; idx += blockDim.x) mw = fmax(mw, fabs(p.w(idx / N, idx % N)));
  double rm = 0.0, cm = 0.0;
  if (threadIdx.x < H) {                      // hidden-like unit: |bias| + sum_j |W_ij|
    const int i = threadIdx.x;
    rm = fabs(p.hbias(i));
    for (int j = 0; j < N; ++j) rm += fabs(p.w(i, j));
  } else if (threadIdx.x >= 512 && threadIdx.x < 512 + N) {
    const int j = threadIdx.x - 512;
    cm = fabs(p.b[j]);
    for (int i = 0; i < H; ++i) cm += fabs(p.w(i, j));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mw = fmax(mw, __shfl_xor_sync(0xffffffffu, mw, o));
    rm = fmax(rm, __shfl_xor_sync(0xffffffffu, rm, o));
    cm = fmax(cm, __shfl_xor_sync(0xffffffffu, cm, o));
  }
  if (lane == 0) { red[0][warp] = mw; red[1][warp] = rm; red[2][warp] = cm; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s_maxw = red[0][0], s_rowmax = red[1][0], s_colmax = red[2][0];
    for (int w = 1; w < 32; ++w) {
      s_maxw = fmax(s_maxw, red[0][w]);
      s_rowmax = fmax(s_rowmax, red[1][w]);
      s_colmax = fmax(s_colmax, red[2][w]);
    }
    int s = 0;
    if (s_maxw > 0.0 && s_maxw < 1e300) s = 21 - ilogb(s_maxw);
    if (s > 60) s = 60;
    if (s_maxw > 0.0 && s_maxw < 1e300 && llrint(s_maxw * ldexp(1.0, s)) > GW_WQ_MAX) s -= 1;
    const bool degenerate = !(s_maxw < 1e300) || s < -60;  // every unit goes to the exact FP64 path
    if (s < -60) s = -60;
    GwHeader* hd = reinterpret_cast<GwHeader*>(img);
    hd->inv_scale = (float)(-1.4426950408889634 * ldexp(1.0, -s));
    hd->s = s;
    hd->degenerate = degenerate ? 1 : 0;
    const double dxh = N * ldexp(1.0, -(s + 1)) + ldexp(1.0, -21) * s_rowmax;
    const double dxv = H * ldexp(1.0, -(s + 1)) + ldexp(1.0, -21) * s_colmax;
    hd->dx_h = (float)dxh;
    hd->dx_v = (float)dxv;
    const double mh = (dxh * 0.25 + ldexp(1.0, -19)) * 1.0001, mv = (dxv * 0.25 + ldexp(1.0, -19)) * 1.0001;
    hd->margin_h = (degenerate || !(mh < 4.0)) ? 4.0f : (float)mh;
    hd->margin_v = (degenerate || !(mv < 4.0)) ? 4.0f : (float)mv;
    hd->pad = 0;
  }
  float* biasH = reinterpret_cast<float*>(img + GW_BIAS_OFF);
  float* biasV = biasH + 128;
  for (int pi = threadIdx.x; pi < 128; pi += blockDim.x) {
    const int u = gw_unit_of_p(pi);
    biasH[pi] = u < H ? (float)(-1.4426950408889634 * p.hbias(u)) : 0.0f;
    biasV[pi] = u < N ? (float)(-1.4426950408889634 * p.b[u]) : 0.0f;
  }
}

// canonical K-major no-swizzle offset of element (row n, byte k) of an operand with `rows` rows
__host__ __device__ inline size_t gw_operand_offset(int rows, int n, int k) { return (size_t)(k >> 4) * rows * 16 + (size_t)n * 16 + (k & 15); }

// ---- pack, step 2: one thread per (hidden p-index, visible p-index) of the padded layer pair writes its digits into both
// B operands.  Bh (h | v): rows = output p-index (hidden), k = input p-index (visible); Bv (v | h): the transpose. ----------
__global__ void __launch_bounds__(256) gibbs_pack_wide_digits_kernel(StackedParams p, int PV, int PH, unsigned char* __restrict__ img) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= PV * PH) return;
  const GwHeader* hd = reinterpret_cast<const GwHeader*>(img);
  const int ph = idx / PV, pv = idx - ph * PV;
  const int hid = gw_unit_of_p(ph), vis = gw_unit_of_p(pv);
  int q0 = 0, q1 = 0, q2 = 0;
  if (!hd->degenerate && hid < p.nhid() && vis < p.N) {
    long long wq = llrint(p.w(hid, vis) * ldexp(1.0, hd->s));
    q0 = (int)(((wq + 128) & 255) - 128);
    wq = (wq - q0) >> 8;
    q1 = (int)(((wq + 64) & 127) - 64);
    q2 = (int)((wq - q1) >> 7);
  }
  unsigned char* Bh = img + GW_B_OFF;
  unsigned char* Bv = Bh + gw_b_bytes(PV, PH);
  const unsigned char b0 = (unsigned char)(q0 & 255), b1 = (unsigned char)((2 * q1) & 255), b2 = (unsigned char)(q2 & 255);
  Bh[gw_operand_offset(2 * PH, ph, pv)] = b0;
  Bh[gw_operand_offset(2 * PH, ph, PV + pv)] = b1;
  Bh[gw_operand_offset(2 * PH, PH + ph, pv)] = b2;
  Bh[gw_operand_offset(2 * PH, PH + ph, PV + pv)] = 0;
  Bv[gw_operand_offset(2 * PV, pv, ph)] = b0;
  Bv[gw_operand_offset(2 * PV, pv, PH + ph)] = b1;
  Bv[gw_operand_offset(2 * PV, PV + pv, ph)] = b2;
  Bv[gw_operand_offset(2 * PV, PV + pv, PH + ph)] = 0;
}

// ---- device helpers (same arithmetic as gibbs_umma.cu) --------------------------------------------------------------------
__device__ __forceinline__ void gw_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const uint32_t (&rk)[20], uint32_t (&out)[4]) {
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
// K-major, no swizzle: start address, LBO (K direction) and SBO (8-row groups, 128 bytes) in 16-byte units, version 1
__device__ __forceinline__ uint64_t gw_desc(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void gw_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void gw_ld8(uint32_t taddr, int (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ uint32_t gw_spread4(uint32_t nib) { return (nib * 0x00204081u) & 0x01010101u; }
__device__ __forceinline__ float gw_ex2(float y) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(y));
  return r;
}
// natural bit order (bit j = unit 32 g + j) <-> p-order (bit 8 c + 4 h + w = unit 32 g + 16 h + 4 c + w)
__device__ __forceinline__ uint32_t gw_natural_to_p(uint32_t x) {
  uint32_t r = 0u;
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int h = 0; h < 2; ++h) r |= ((x >> (16 * h + 4 * c)) & 0xFu) << (8 * c + 4 * h);
  return r;
}
__device__ __forceinline__ uint32_t gw_p_to_natural(uint32_t x) {
  uint32_t r = 0u;
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int h = 0; h < 2; ++h) r |= ((x >> (8 * c + 4 * h)) & 0xFu) << (16 * h + 4 * c);
  return r;
}

// Cold path (all 32 lanes converged): exact FP64 pre-activation from the original parameters, full 53-bit uniform, exact
// sigmoid, for every undecided unit of every lane.  in_words: this conditional's input bits of the tile's 128 chains,
// p-order, [group][row]; g: the 32-unit group of the output layer this warp decides.
__device__ __noinline__ uint32_t fixup_slow_wide(const GwArgs* a, int hidden_out, const uint32_t* in_words, int row, uint32_t nib,
                                                 uint32_t slowmask, uint32_t c0, uint32_t c1, uint32_t sweep, uint32_t q0, int g, int lane) {
  const StackedParams& p = a->p;
  const int n_in = hidden_out ? p.N : p.nhid();
  uint32_t pending = __ballot_sync(0xffffffffu, slowmask != 0u);
  while (pending != 0u) {
    const int src = __ffs(pending) - 1;
    pending &= pending - 1u;
    uint32_t sm = __shfl_sync(0xffffffffu, slowmask, src);
    const int srow = __shfl_sync(0xffffffffu, row, src);
    const uint32_t sc0 = __shfl_sync(0xffffffffu, c0, src), sc1 = __shfl_sync(0xffffffffu, c1, src);
    while (sm != 0u) {
      const int pbit = __ffs(sm) - 1;
      sm &= sm - 1u;
      const int unit = 32 * g + gw_u5_of_p5(pbit);
      double part = 0.0;
      for (int kin = lane; kin < n_in; kin += 32) {   // input unit kin: group kin / 32, p-index of (kin % 32) = that of lane
        const uint32_t word = in_words[(kin >> 5) * GW_TILE + srow];
        if ((word >> gw_p5_of_u5(kin & 31)) & 1u) part += hidden_out ? p.w(unit, kin) : p.w(kin, unit);
      }
      part = warp_sum(part);
      const double x = (hidden_out ? p.hbias(unit) : p.b[unit]) + part;
      const uint32_t c = (uint32_t)(pbit >> 3), half = (uint32_t)((pbit >> 2) & 1), w = (uint32_t)(pbit & 3);
      const uint32_t q = q0 + 4u * (uint32_t)g + c;
      uint32_t hi[4], la[4], lb[4];
      philox4x32_10(sc0, sc1, sweep, (0u << 29) | q, a->key0, a->key1, hi);
      philox4x32_10(sc0, sc1, sweep, ((1u + 2u * half) << 29) | q, a->key0, a->key1, la);
      philox4x32_10(sc0, sc1, sweep, ((2u + 2u * half) << 29) | q, a->key0, a->key1, lb);
      const uint32_t hw = w == 0 ? hi[0] : (w == 1 ? hi[1] : (w == 2 ? hi[2] : hi[3]));
      const uint32_t aw = w == 0 ? la[0] : (w == 1 ? la[1] : (w == 2 ? la[2] : la[3]));
      const uint32_t bw = w == 0 ? lb[0] : (w == 1 ? lb[1] : (w == 2 ? lb[2] : lb[3]));
      const uint64_t h16 = (hw >> (16u * (1u - half))) & 0xFFFFu;
      const uint64_t ubits = (h16 << 37) | ((uint64_t)aw << 5) | (uint64_t)(bw >> 27);
      const double u = (double)ubits * 0x1.0p-53;
      const bool one = u < sigmoid_exact(x);
      if (lane == src) nib = one ? (nib | (1u << pbit)) : (nib & ~(1u << pbit));
    }
  }
  return nib;
}

// GMAX: 32-unit groups of the wider (padded) layer: 4 (65..128 units), 2 (33..64), 1 (<= 32).  A tile of 128 chains is run by
// 128 GMAX threads; a CTA of 1 024 threads runs 8 / GMAX tiles around one copy of the weight images.
// GVT / GHT: the layers' group counts as compile-time constants (0: read them from the arguments).
template <int GMAX, int GVT = 0, int GHT = 0>
__global__ void __launch_bounds__(GW_THREADS, 1) gibbs_umma_wide_kernel(const GwArgs a) {
  constexpr int TT = 128 * GMAX;              // threads per tile
  constexpr int TILES = GW_THREADS / TT;      // tiles per CTA
  constexpr int A_BYTES = 128 * 64 * GMAX;    // 128 chains x 2 x 32 GMAX bytes
  constexpr uint32_t TCOLS = 64u * GMAX;      // TMEM columns per tile: 2 x 32 GMAX
  constexpr int RPW = 32 / GMAX;              // rows per warp in the float64 I/O phases
  extern __shared__ __align__(128) unsigned char smem[];
  const int PV = GVT > 0 ? 32 * GVT : a.PV, PH = GHT > 0 ? 32 * GHT : a.PH;
  const size_t image_bytes = gw_image_bytes(PV, PH), b_bytes = gw_b_bytes(PV, PH);
  const GwHeader* hd = reinterpret_cast<const GwHeader*>(smem);
  const float* biasH = reinterpret_cast<const float*>(smem + GW_BIAS_OFF);
  const float* biasV = biasH + 128;
  const int tid = threadIdx.x;
  const int tile_slot = tid / TT;                          // which of the CTA's tiles
  const int ttid = tid % TT, wl = ttid >> 5, lane = tid & 31;
  const int g = wl >> 2;                                   // 32-unit group of either layer
  const int row = 32 * (wl & 3) + lane;                    // chain row of the tile = TMEM lane
  unsigned char* sA = smem + image_bytes + (size_t)tile_slot * A_BYTES;      // this tile's state bytes (A operand)
  uint32_t* words_all = reinterpret_cast<uint32_t*>(smem + image_bytes + (size_t)TILES * A_BYTES);
  uint32_t* words = words_all + (size_t)tile_slot * 2 * GMAX * GW_TILE;       // [2 buffers][GMAX groups][128 rows]
  uint64_t* mbar = reinterpret_cast<uint64_t*>(words_all + (size_t)TILES * 2 * GMAX * GW_TILE);  // [0]: image, [1 + t]: MMA of tile t
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1 + TILES);
  uint32_t* arrive_cnt = tmem_slot + 1 + tile_slot;        // threads of this tile that have published their bits

  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(GW_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (tid == 0) {
    for (int i = 0; i < 1 + TILES; ++i) mbar_init(mbar + i, 1);
    for (int i = 0; i < TILES; ++i) tmem_slot[1 + i] = 0u;
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(mbar, (uint32_t)image_bytes);
    for (size_t off = 0; off < image_bytes; off += 16384u) {
      const uint32_t n = image_bytes - off < 16384u ? (uint32_t)(image_bytes - off) : 16384u;
      tma_bulk_g2s(smem + off, a.image + off, n, mbar);
    }
  }
  mbar_wait(mbar, 0);
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t tmem = *tmem_slot + TCOLS * (uint32_t)tile_slot;
  const uint32_t taddr = tmem + ((uint32_t)(32 * (wl & 3)) << 16) + 32u * (uint32_t)g;  // this thread's lane, first acc01 column of its group
  const float k1 = hd->inv_scale;
  const int nhid = a.p.nhid(), N = a.p.N;
  const int GV = PV >> 5, GH = PH >> 5;
  const uint32_t qv0 = (uint32_t)(4 * ((nhid + 31) / 32));
  // real (unpadded) units among this thread's 32 of either layer, in p-order
  uint32_t valid_h = 0u, valid_v = 0u;
#pragma unroll
  for (int pb = 0; pb < 32; ++pb) {
    const int u = 32 * g + gw_u5_of_p5(pb);
    valid_h |= (u < nhid ? 1u : 0u) << pb;
    valid_v |= (u < N ? 1u : 0u) << pb;
  }
  const int rowbytes = packed_row_bytes(N);
  const uint64_t adesc = gw_desc(smem_u32(sA), GW_A_LBO);
  const uint64_t bdesc_h = gw_desc(smem_u32(smem + GW_B_OFF), (uint32_t)(2 * PH * 16));
  const uint64_t bdesc_v = gw_desc(smem_u32(smem + GW_B_OFF + b_bytes), (uint32_t)(2 * PV * 16));
  // instruction descriptors: D s32, A u8, B s8, both K-major, M = 128, N = 2 PO
  const uint32_t idesc_h = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)((2 * PH) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t idesc_v = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)((2 * PV) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  uint32_t mma_phase = 0u;
  const int bar_id = 1 + tile_slot;
  auto tile_bar = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(TT) : "memory"); };

  const int64_t num_tiles = (a.C + GW_TILE - 1) / GW_TILE;
  for (int64_t ti = (int64_t)blockIdx.x * TILES + tile_slot; ti < num_tiles; ti += (int64_t)gridDim.x * TILES) {
    const int64_t chain_base = ti * GW_TILE;
    const int64_t ch = chain_base + row;
    const uint64_t id = a.chain_id0 + (uint64_t)ch;
    const uint32_t cid0 = (uint32_t)id, cid1 = (uint32_t)(id >> 32);

    // ---- initial state of this thread's (chain, visible group), natural bit order -> p-order ----
    uint32_t cur = 0u;
    if (a.in_mode == 0) {
      // float64 rows: warp wl fetches rows RPW wl .. RPW wl + RPW - 1 of the tile, 256 contiguous bytes per load, ballots the bits
      // and leaves the (up to GMAX) 32-bit words of each row in buffer 0 of `words`
#pragma unroll 4
      for (int rr = 0; rr < RPW; ++rr) {
        const int r = RPW * wl + rr;
        const int64_t c = chain_base + r;
        unsigned long long d[GMAX];
#pragma unroll
        for (int j = 0; j < GMAX; ++j) d[j] = 0ull;
        if (c < a.C) {
          const unsigned long long* rp = reinterpret_cast<const unsigned long long*>(a.v_in) + c * N;
#pragma unroll
          for (int j = 0; j < GMAX; ++j)
            if (lane + 32 * j < N) d[j] = rp[lane + 32 * j];
        }
#pragma unroll
        for (int j = 0; j < GMAX; ++j) {
          const uint32_t bj = __ballot_sync(0xffffffffu, (d[j] << 1) != 0ull);
          if (lane == 0) words[j * GW_TILE + r] = bj;
        }
      }
      tile_bar();
      cur = words[g * GW_TILE + row];
      // pull this tile slot's NEXT tile (one contiguous span of v_in) into L2 while this one is swept
      const int64_t nbase = (ti + (int64_t)gridDim.x * TILES) * GW_TILE;
      if (nbase < a.C) {
        const int64_t nbytes = (a.C - nbase < GW_TILE ? a.C - nbase : (int64_t)GW_TILE) * N * (int64_t)sizeof(double);
        const char* np = reinterpret_cast<const char*>(a.v_in + nbase * N);
        for (int64_t off = (int64_t)ttid * 128; off < nbytes; off += (int64_t)TT * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(np + off));
      }
    } else if (a.in_mode == 1) {
      if (ch < a.C) {
        const uint8_t* rowp = a.bits_in + ch * rowbytes;
        uint32_t x = 0u;
#pragma unroll
        for (int bb = 0; bb < 4; ++bb)
          if (4 * g + bb < rowbytes) x |= (uint32_t)rowp[4 * g + bb] << (8 * bb);
        cur = __brev(__byte_perm(x, 0u, 0x0123u));
      }
    } else {
      uint32_t rnd[4];
      philox4x32_10(cid0, cid1, a.sweep0, PHILOX_STREAM_V0 << 29, a.key0, a.key1, rnd);
      cur = g == 0 ? rnd[0] : (g == 1 ? rnd[1] : (g == 2 ? rnd[2] : rnd[3]));
    }
    cur = gw_natural_to_p(cur) & valid_v;

#pragma unroll 1
    for (int hs = 0; hs < 2 * a.k; ++hs) {
      const bool back = (hs & 1) != 0;  // false: h | v, true: v | h
      const int GI = back ? GH : GV, GO = back ? GV : GH;   // 32-unit groups of the input / output layer
      // ---- publish this thread's 32 input bits: A operand bytes [v ; 128 v] and the word for the cold path ----
      uint32_t* in_words = words + (hs & 1) * GMAX * GW_TILE;
      if (g < GI) {
        in_words[g * GW_TILE + row] = cur;
        uint4 lo, hi4;
        lo.x = gw_spread4(cur & 0xFu); lo.y = gw_spread4((cur >> 4) & 0xFu); lo.z = gw_spread4((cur >> 8) & 0xFu); lo.w = gw_spread4((cur >> 12) & 0xFu);
        hi4.x = gw_spread4((cur >> 16) & 0xFu); hi4.y = gw_spread4((cur >> 20) & 0xFu); hi4.z = gw_spread4((cur >> 24) & 0xFu); hi4.w = gw_spread4(cur >> 28);
        unsigned char* base = sA + row * 16;
        *reinterpret_cast<uint4*>(base + (2 * g) * GW_A_LBO) = lo;
        *reinterpret_cast<uint4*>(base + (2 * g + 1) * GW_A_LBO) = hi4;
        lo.x <<= 7; lo.y <<= 7; lo.z <<= 7; lo.w <<= 7;
        hi4.x <<= 7; hi4.y <<= 7; hi4.z <<= 7; hi4.w <<= 7;
        *reinterpret_cast<uint4*>(base + (2 * GI + 2 * g) * GW_A_LBO) = lo;       // the 128 v copy starts at K byte 32 GI
        *reinterpret_cast<uint4*>(base + (2 * GI + 2 * g + 1) * GW_A_LBO) = hi4;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core's reads
      asm volatile("tcgen05.fence::before_thread_sync;\n");       // ... and this thread's TMEM reads of the previous accumulators are done
      // every thread of the tile checks in (release); the LAST one (acquire: it sees all publications) issues the MMAs
      uint32_t arrived;
      asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(arrived) : "r"(smem_u32(arrive_cnt)) : "memory");
      if (arrived == TT - 1) {
        *reinterpret_cast<volatile uint32_t*>(arrive_cnt) = 0u;  // nobody checks in again before this MMA has completed
        asm volatile("tcgen05.fence::after_thread_sync;\n");
        const uint64_t bdesc = back ? bdesc_v : bdesc_h;
        const uint32_t idesc = back ? idesc_v : idesc_h;
        const uint64_t a_step = (uint64_t)((2 * GW_A_LBO) >> 4);                   // consecutive K = 32 slices: two K chunks on
        const uint64_t b_step = (uint64_t)((2 * (2 * (back ? PV : PH) * 16)) >> 4);
        const int nslices = 2 * GI;                                                // K = 2 PI = 64 GI bytes
        for (int ks = 0; ks < nslices; ++ks) gw_mma(tmem, adesc + (uint64_t)ks * a_step, bdesc + (uint64_t)ks * b_step, idesc, ks > 0 ? 1u : 0u);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar + 1 + tile_slot)) : "memory");
      }
      const uint32_t valid = g < GO ? (back ? valid_v : valid_h) : 0u;
      const float* bias = (back ? biasV : biasH) + 32 * g;
      const float margin = back ? hd->margin_v : hd->margin_h;
      const uint32_t sweep = a.sweep0 + (uint32_t)(hs >> 1);
      const uint32_t q0 = (back ? qv0 : 0u) + 4u * (uint32_t)g;
      const float c1 = (0x1.0p-16f + margin) - 1.0f;   // s1 = (m + c1) q - 1 < 0             => certainly u < p
      const float d01 = 0x1.0p-16f + 2.0f * margin;    // s0 = d01 q - s1 = 1 - (u_f - margin) q < 0 => certainly u >= p
      const uint32_t acc2_off = 32u * (uint32_t)GO;    // PO columns on
      // the draws of the first GW_PREFETCH calls are made while the tile's MMAs run; the remaining calls overlap the TMEM loads
      uint32_t pre[GW_PREFETCH > 0 ? GW_PREFETCH : 1][4];
      if (valid != 0u) {
#pragma unroll
        for (int j = 0; j < GW_PREFETCH; ++j) gw_philox(cid0, cid1, sweep, q0 + (uint32_t)(3 - j), a.rk, pre[j]);
      }
      mbar_wait(mbar + 1 + tile_slot, mma_phase);
      mma_phase ^= 1u;
      asm volatile("tcgen05.fence::after_thread_sync;\n");

      if (valid != 0u) {   // warp-uniform: a group is a set of whole warps
        uint32_t acc1 = 0u, acc0 = 0u;  // bit p: "certainly 1" / "certainly 0"; units are visited from p = 31 down
#pragma unroll
        for (int c = 3; c >= 0; --c) {
          const uint32_t vb = (valid >> (8 * c)) & 0xFFu;
          if (GMAX < 4 && vb == 0u) {   // a Philox call that serves padding only (narrow layers): warp-uniform skip
            acc1 <<= 8;
            acc0 = (acc0 << 8) | 0xFFu;
            continue;
          }
          int a01[8], a2[8];
          gw_ld8(taddr + 8u * (uint32_t)c, a01);
          gw_ld8(taddr + acc2_off + 8u * (uint32_t)c, a2);
          uint32_t rnd[4];
          if (3 - c < GW_PREFETCH) {
#pragma unroll
            for (int w = 0; w < 4; ++w) rnd[w] = pre[3 - c < GW_PREFETCH ? 3 - c : 0][w];
          } else {
            gw_philox(cid0, cid1, sweep, q0 + (uint32_t)c, a.rk, rnd);  // stream 0; overlaps the TMEM loads
          }
          const float4 b_lo = *reinterpret_cast<const float4*>(bias + 8 * c);
          const float4 b_hi = *reinterpret_cast<const float4*>(bias + 8 * c + 4);
          const float bs[8] = {b_lo.x, b_lo.y, b_lo.z, b_lo.w, b_hi.x, b_hi.y, b_hi.z, b_hi.w};
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (GMAX < 4 && (vb & 0xF0u) == 0u) {   // layers of at most 16 units use the high 16-bit halves only (units 16.. are padding)
            acc1 <<= 4;
            acc0 = (acc0 << 4) | 0xFu;
#pragma unroll
            for (int i = 3; i >= 0; --i) {
              const int t = a2[i] * 32768 + a01[i];
              const float y = fmaf((float)t, k1, bs[i]);
              const float q = 1.0f + gw_ex2(y);
              const float m = __uint_as_float(__byte_perm(rnd[i], 0x7Fu, 0x5432u) << 7);
              const float s1 = fmaf(m + c1, q, -1.0f);
              const float s0 = fmaf(d01, q, -s1);
              acc1 = __funnelshift_l(__float_as_uint(s1), acc1, 1);
              acc0 = __funnelshift_l(__float_as_uint(s0), acc0, 1);
            }
            continue;
          }
#pragma unroll
          for (int i = 7; i >= 0; --i) {   // i = 4 h + w
            const int half = i >> 2, w = i & 3;
            const int t = a2[i] * 32768 + a01[i];
            const float y = fmaf((float)t, k1, bs[i]);  // -log2(e) x~
            const float q = 1.0f + gw_ex2(y);
            const float m = __uint_as_float(__byte_perm(rnd[w], 0x7Fu, half ? 0x5410u : 0x5432u) << 7);
            const float s1 = fmaf(m + c1, q, -1.0f);
            const float s0 = fmaf(d01, q, -s1);
            acc1 = __funnelshift_l(__float_as_uint(s1), acc1, 1);
            acc0 = __funnelshift_l(__float_as_uint(s0), acc0, 1);
          }
        }
        uint32_t nib = acc1 & valid;
        const uint32_t slowmask = ~(acc0 | acc1) & valid;
        if (__any_sync(0xffffffffu, slowmask != 0u))
          nib = fixup_slow_wide(&a, back ? 0 : 1, in_words, row, nib, slowmask, cid0, cid1, sweep, back ? qv0 : 0u, g, lane);
        cur = nib;
      } else if (g < GO) {
        cur = 0u;   // a group of padding only
      }
      asm volatile("tcgen05.fence::before_thread_sync;\n");  // this thread's TMEM reads are ordered before the next MMA (via the arrival)
    }

    // ---- final state: p-order -> natural, float64 rows and / or packed rows ----
    const uint32_t nat = gw_p_to_natural(cur & valid_v);
    if (a.v != nullptr) {
      // float64 rows out: the words go through buffer 1 of `words` (free once every thread of the tile has left the last
      // update's cold path), then warp wl writes its RPW rows, 256 contiguous bytes per store
      tile_bar();
      if (g < GV) words[GMAX * GW_TILE + g * GW_TILE + row] = nat;
      tile_bar();
#pragma unroll 4
      for (int rr = 0; rr < RPW; ++rr) {
        const int r = RPW * wl + rr;
        const int64_t c = chain_base + r;
        if (c < a.C) {
          double* rp = a.v + c * N;
#pragma unroll
          for (int j = 0; j < GMAX; ++j)
            if (lane + 32 * j < N) rp[lane + 32 * j] = (words[GMAX * GW_TILE + j * GW_TILE + r] >> lane) & 1u ? 1.0 : 0.0;
        }
      }
    }
    if (a.bits_out != nullptr && ch < a.C && g < GV) {
      const uint32_t x = __byte_perm(__brev(nat), 0u, 0x0123u);
      uint8_t* rowp = a.bits_out + ch * rowbytes;
#pragma unroll
      for (int bb = 0; bb < 4; ++bb)
        if (4 * g + bb < rowbytes) rowp[4 * g + bb] = (uint8_t)(x >> (8 * bb));
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(*tmem_slot), "r"(GW_TMEM_COLS));
}

static int gw_pad(int units) { return units <= 32 ? 32 : (units <= 64 ? 64 : 128); }
static int gw_gmax(int PV, int PH) { return (PV > PH ? PV : PH) / 32; }
static size_t gw_smem_bytes(int PV, int PH) {
  const int gmax = gw_gmax(PV, PH), tiles = 8 / gmax;
  return gw_image_bytes(PV, PH) + (size_t)tiles * 128 * 64 * gmax + (size_t)tiles * 2 * gmax * GW_TILE * sizeof(uint32_t) + (1 + tiles) * 8 +
         (1 + tiles) * 4 + 16;
}

// shapes: both layers of 1..128 units (the dispatcher in gibbs.cu decides which of them come here)
size_t gibbs_umma_wide_workspace_bytes(int N, int nhid) {
  if (N < 1 || nhid < 1 || N > 128 || nhid > 128) return 0;
  return align_up(gw_image_bytes(gw_pad(N), gw_pad(nhid)), 256);
}

int64_t gibbs_umma_wide_wave_chains(int N, int nhid) { return (int64_t)sm_count() * (8 / gw_gmax(gw_pad(N), gw_pad(nhid))) * GW_TILE; }

// *handled = 1 when this kernel took the call (own Philox draws; layers of up to 128 units)
int try_gibbs_umma_wide(const StackedParams& p, const ChainIO& io, int64_t C, int k, uint64_t seed, uint64_t chain_id0, uint64_t sweep0,
                        void* ws, size_t ws_bytes, int* handled, cudaStream_t st) {
  *handled = 0;
  const size_t need = gibbs_umma_wide_workspace_bytes(p.N, p.nhid());
  if (need == 0 || ws == nullptr || ws_bytes < need || k < 1) return QCB_OK;
  const int PV = gw_pad(p.N), PH = gw_pad(p.nhid());
  unsigned char* img = reinterpret_cast<unsigned char*>(ws);
  gibbs_pack_wide_header_kernel<<<1, 1024, 0, st>>>(p, img);
  QCB_LAUNCH_CHECK();
  gibbs_pack_wide_digits_kernel<<<(PV * PH + 255) / 256, 256, 0, st>>>(p, PV, PH, img);
  QCB_LAUNCH_CHECK();
  GwArgs a;
  a.image = img;
  a.p = p;
  a.v_in = io.v_in;
  a.v = io.v_out;
  a.bits_in = io.bits_in;
  a.bits_out = io.bits_out;
  a.in_mode = io.philox_init ? 2 : (io.bits_in != nullptr ? 1 : 0);
  a.PV = PV;
  a.PH = PH;
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
  const size_t smem = gw_smem_bytes(PV, PH);
  const int gmax = gw_gmax(PV, PH), tiles_per_cta = 8 / gmax;
  static bool configured[QCB_MAX_DEVICES] = {false};
  bool& done = configured[device_slot()];
  if (!done) {
    QCB_CUDA(cudaFuncSetAttribute(gibbs_umma_wide_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gw_smem_bytes(128, 128)));
    QCB_CUDA((cudaFuncSetAttribute(gibbs_umma_wide_kernel<4, 4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gw_smem_bytes(128, 128))));
    QCB_CUDA(cudaFuncSetAttribute(gibbs_umma_wide_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gw_smem_bytes(64, 64)));
    QCB_CUDA(cudaFuncSetAttribute(gibbs_umma_wide_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gw_smem_bytes(32, 32)));
    done = true;
  }
  // one CTA per SM (all 512 TMEM columns), 8 / GMAX tiles per CTA
  const int64_t tiles = (C + GW_TILE - 1) / GW_TILE;
  const int64_t want = (tiles + tiles_per_cta - 1) / tiles_per_cta, cap = (int64_t)sm_count();
  const int grid = (int)(want < cap ? want : cap);
  if (PV == 128 && PH == 128) gibbs_umma_wide_kernel<4, 4, 4><<<grid, GW_THREADS, smem, st>>>(a);
  else if (gmax == 4) gibbs_umma_wide_kernel<4><<<grid, GW_THREADS, smem, st>>>(a);
  else if (gmax == 2) gibbs_umma_wide_kernel<2><<<grid, GW_THREADS, smem, st>>>(a);
  else gibbs_umma_wide_kernel<1><<<grid, GW_THREADS, smem, st>>>(a);
  QCB_LAUNCH_CHECK();
  *handled = 1;
  return QCB_OK;
}

}  // namespace qcb
