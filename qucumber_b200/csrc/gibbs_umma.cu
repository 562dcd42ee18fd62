// Block Gibbs with the integer-exact filter on the 5th-generation tensor cores: tcgen05.mma.kind::i8 with the
// accumulators in tensor memory (TMEM).  Same algorithm, same Philox convention and the same decisions — hence the same
// samples, bit for bit — as gibbs_imma.cu (mma.sync), for layers of 33..64 units on both sides:
// BinaryRBM.gibbs_steps (rbm/binary_rbm.py:200-228) / PurificationRBM.gibbs_steps (rbm/purification_rbm.py:309-337).
//
// What changes is where the contraction lives.  A CTA owns a tile of 128 chains (M = 128):
//  * the chains' state bits sit in shared memory as the A operand, one row of bytes per chain, K-major canonical
//    layout (8 x 16-byte core matrices), 128 bytes per row: [v ; 128 v];
//  * the weights sit in shared memory as the B operand (N = 128 rows x K = 128): rows 0..63 = [d0 ; 2 d1], rows
//    64..127 = [d2 ; 0] of the fixed-point digits Wq = d2 2^15 + d1 2^8 + d0 (8 / 7 / 8 bits), so ONE accumulator tile
//    holds acc01 = sum_k v_k (d0 + 256 d1) in columns 0..63 and acc2 = sum_k v_k d2 in columns 64..127 and
//    t = acc2 2^15 + acc01 is the exact 24-bit fixed-point pre-activation;
//  * one elected thread issues 4 x tcgen05.mma (K = 32 each) per conditional update and commits to an mbarrier;
//  * 8 warps read the accumulators with tcgen05.ld (warp w and w + 4 share TMEM lane quarter w % 4: a thread owns ONE
//    chain and 32 of its 64 output units), draw Philox uniforms, decide, write the new bits straight back as the next
//    A operand (4 x 16-byte stores) — no shuffles, no fragment permutations, no accumulator registers.
// Units are kept in "p-order" inside the kernel so that the 8 units served by one Philox call are 8 consecutive TMEM
// columns: unit u = 32 hh + 16 h + 4 c + w  <->  p = 32 hh + 8 c + 4 h + w  (hh: 32-unit half, c: call, h: 16-bit half).
#include "common.cuh"
#include <cstdlib>

namespace qcb {

constexpr int GU_THREADS = 256;
constexpr int GU_TILE = 128;          // chains per CTA tile (M of the MMA)
constexpr int GU_MIN_CTAS = 4;        // 4 x 128 TMEM columns = the whole tensor memory of an SM
constexpr uint32_t GU_TMEM_COLS = 128;
#ifndef GU_PREFETCH
#define GU_PREFETCH 1                 // Philox calls of an update drawn before the MMA-completion wait (registers: 4 words each)
#endif
constexpr int GU_OPERAND_BYTES = 128 * 128;  // A (128 chains x 128 bytes) and each B image (128 rows x 128 bytes)
constexpr int GU_LBO = (128 / 8) * 128;      // byte distance of consecutive 16-byte K chunks (no swizzle, K-major)

struct UmmaHeader {
  float inv_scale;   // -log2(e) * 2^-s
  float margin_h, margin_v;
  float dx_h, dx_v;
  int s;
  int pad[2];
};

// image: header | biasH[64] | biasV[64] | Bh | Bv   (everything the CTA keeps in shared memory)
constexpr size_t GU_BIAS_OFF = 32;
constexpr size_t GU_BH_OFF = 1024;
constexpr size_t GU_BV_OFF = GU_BH_OFF + GU_OPERAND_BYTES;
constexpr size_t GU_IMAGE_BYTES = GU_BV_OFF + GU_OPERAND_BYTES;
static_assert(sizeof(UmmaHeader) == 32, "header layout");

struct GibbsUmmaArgs {
  const unsigned char* image;
  StackedParams p;        // original FP64 parameters (cold path)
  const double* v_in;
  double* v;
  const uint8_t* bits_in;
  uint8_t* bits_out;
  int in_mode;            // 0: v_in, 1: bits_in, 2: Philox initial state
  int64_t C;
  int k;
  uint32_t key0, key1;
  uint64_t chain_id0;
  uint32_t sweep0;
  uint32_t rk[20];        // Philox round keys (key0 + r W0, key1 + r W1): launch constants, read straight from the constant bank
};

__host__ __device__ inline int gu_unit_of_p(int p) { return (p & 32) | (((p >> 2) & 1) << 4) | (((p >> 3) & 3) << 2) | (p & 3); }
__host__ __device__ inline int gu_p_of_unit(int u) { return (u & 32) | (((u >> 2) & 3) << 3) | (((u >> 4) & 1) << 2) | (u & 3); }
// canonical K-major no-swizzle offset of element (row, k) of a 128-row operand
__host__ __device__ inline int gu_operand_offset(int row, int k) { return (k >> 4) * GU_LBO + row * 16 + (k & 15); }

constexpr long long GU_WQ_MAX = 127ll * 32768 + 63 * 256 + 127;

// ---- pack: quantise (same fixed point, digits and margins as gibbs_imma.cu), write the B operands ---------------
constexpr int GU_PACK_LD = 65;
constexpr int GU_PACK_THREADS = 1024;
__global__ void __launch_bounds__(GU_PACK_THREADS) gibbs_pack_umma_kernel(StackedParams p, unsigned char* __restrict__ img) {
  __shared__ double Ws[64 * GU_PACK_LD];
  __shared__ double red[3][GU_PACK_THREADS / 32];
  const int H = p.nhid(), N = p.N;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double mw = 0.0;
  for (int idx = threadIdx.x; idx < H * N; idx += blockDim.x) {
    const int i = idx / N, j = idx - i * N;
    const double w = p.w(i, j);
    Ws[i * GU_PACK_LD + j] = w;
    mw = fmax(mw, fabs(w));
  }
  __syncthreads();
  // |bias| + sum |w| of every output unit: 8 threads per unit (threads 0..511: hidden-like units, 512..1023: visible units)
  double rm = 0.0, cm = 0.0;
  {
    const int unit = (threadIdx.x & 511) >> 3, part = threadIdx.x & 7;
    double acc = 0.0;
    if (threadIdx.x < 512) {
      if (unit < H) {
        for (int j = part; j < N; j += 8) acc += fabs(Ws[unit * GU_PACK_LD + j]);
        if (part == 0) acc += fabs(p.hbias(unit));
      }
    } else if (unit < N) {
      for (int i = part; i < H; i += 8) acc += fabs(Ws[i * GU_PACK_LD + unit]);
      if (part == 0) acc += fabs(p.b[unit]);
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (threadIdx.x < 512) rm = acc; else cm = acc;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mw = fmax(mw, __shfl_xor_sync(0xffffffffu, mw, o));
    rm = fmax(rm, __shfl_xor_sync(0xffffffffu, rm, o));
    cm = fmax(cm, __shfl_xor_sync(0xffffffffu, cm, o));
  }
  if (lane == 0) { red[0][warp] = mw; red[1][warp] = rm; red[2][warp] = cm; }
  __syncthreads();
  double s_maxw = red[0][0], s_rowmax = red[1][0], s_colmax = red[2][0];
  for (int w = 1; w < GU_PACK_THREADS / 32; ++w) {
    s_maxw = fmax(s_maxw, red[0][w]);
    s_rowmax = fmax(s_rowmax, red[1][w]);
    s_colmax = fmax(s_colmax, red[2][w]);
  }
  int s = 0;
  if (s_maxw > 0.0 && s_maxw < 1e300) s = 21 - ilogb(s_maxw);
  if (s > 60) s = 60;
  if (s_maxw > 0.0 && s_maxw < 1e300 && llrint(s_maxw * ldexp(1.0, s)) > GU_WQ_MAX) s -= 1;
  const bool degenerate = !(s_maxw < 1e300) || s < -60;  // every unit goes to the exact FP64 path
  if (s < -60) s = -60;
  const double scale = degenerate ? 0.0 : ldexp(1.0, s);
  if (threadIdx.x == 0) {
    UmmaHeader* hd = reinterpret_cast<UmmaHeader*>(img);
    hd->inv_scale = (float)(-1.4426950408889634 * ldexp(1.0, -s));
    hd->s = s;
    const double dxh = N * ldexp(1.0, -(s + 1)) + ldexp(1.0, -21) * s_rowmax;
    const double dxv = H * ldexp(1.0, -(s + 1)) + ldexp(1.0, -21) * s_colmax;
    hd->dx_h = (float)dxh;
    hd->dx_v = (float)dxv;
    const double mh = (dxh * 0.25 + ldexp(1.0, -19)) * 1.0001, mv = (dxv * 0.25 + ldexp(1.0, -19)) * 1.0001;
    hd->margin_h = (degenerate || !(mh < 4.0)) ? 4.0f : (float)mh;
    hd->margin_v = (degenerate || !(mv < 4.0)) ? 4.0f : (float)mv;
    hd->pad[0] = hd->pad[1] = 0;
  }
  float* biasH = reinterpret_cast<float*>(img + GU_BIAS_OFF);
  float* biasV = biasH + 64;
  for (int pi = threadIdx.x; pi < 64; pi += blockDim.x) {
    const int u = gu_unit_of_p(pi);
    biasH[pi] = u < H ? (float)(-1.4426950408889634 * p.hbias(u)) : 0.0f;
    biasV[pi] = u < N ? (float)(-1.4426950408889634 * p.b[u]) : 0.0f;
  }
  // B operands, one byte per (row n, k): the thread that owns the pair (hidden p-index ph, visible p-index pv) computes the
  // three digits once and writes them into both images.  Bh (h | v): row ph, k = pv: digit 0; k = 64 + pv: twice digit 1 (it
  // meets the 128 v copy of the state); row 64 + ph: digit 2 for k = pv, zero above.  Bv (v | h): the transpose.
  unsigned char* Bh = img + GU_BH_OFF;
  unsigned char* Bv = img + GU_BV_OFF;
  for (int idx = threadIdx.x; idx < 64 * 64; idx += blockDim.x) {
    const int ph = idx >> 6, pv = idx & 63;
    const int hid = gu_unit_of_p(ph), vis = gu_unit_of_p(pv);
    int q0 = 0, q1 = 0, q2 = 0;
    if (!degenerate && hid < H && vis < N) {
      long long wq = llrint(Ws[hid * GU_PACK_LD + vis] * scale);
      q0 = (int)(((wq + 128) & 255) - 128);
      wq = (wq - q0) >> 8;
      q1 = (int)(((wq + 64) & 127) - 64);
      q2 = (int)((wq - q1) >> 7);
    }
    const unsigned char b0 = (unsigned char)(q0 & 255), b1 = (unsigned char)((2 * q1) & 255), b2 = (unsigned char)(q2 & 255);
    Bh[gu_operand_offset(ph, pv)] = b0;
    Bh[gu_operand_offset(ph, 64 + pv)] = b1;
    Bh[gu_operand_offset(64 + ph, pv)] = b2;
    Bh[gu_operand_offset(64 + ph, 64 + pv)] = 0;
    Bv[gu_operand_offset(pv, ph)] = b0;
    Bv[gu_operand_offset(pv, 64 + ph)] = b1;
    Bv[gu_operand_offset(64 + pv, ph)] = b2;
    Bv[gu_operand_offset(64 + pv, 64 + ph)] = 0;
  }
}

// Philox4x32-10 with the key schedule precomputed on the host (same function as philox4x32_10 of common.cuh)
__device__ __forceinline__ void gu_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const uint32_t (&rk)[20], uint32_t (&out)[4]) {
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

// ---- tcgen05 helpers ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t gu_desc(uint32_t saddr) {
  // K-major, no swizzle: start address, LBO (K direction) and SBO (8-row groups, 128 bytes) in 16-byte units, version 1
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(GU_LBO >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void gu_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void gu_ld8(uint32_t taddr, int (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ uint32_t gu_spread4(uint32_t nib) { return (nib * 0x00204081u) & 0x01010101u; }
__device__ __forceinline__ float gu_ex2(float y) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(y));
  return r;
}
// natural bit order (bit j = unit 32 hh + j) <-> p-order (bit 8 c + 4 h + w = unit 32 hh + 16 h + 4 c + w): a transposition of the
// 2 x 4 matrix of nibbles
__device__ __forceinline__ uint32_t gu_natural_to_p(uint32_t x) {
  uint32_t r = 0u;
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int h = 0; h < 2; ++h) r |= ((x >> (16 * h + 4 * c)) & 0xFu) << (8 * c + 4 * h);
  return r;
}
__device__ __forceinline__ uint32_t gu_p_to_natural(uint32_t x) {
  uint32_t r = 0u;
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int h = 0; h < 2; ++h) r |= ((x >> (8 * c + 4 * h)) & 0xFu) << (16 * h + 4 * c);
  return r;
}

// Cold path (all 32 lanes converged): exact FP64 pre-activation from the original parameters, full 53-bit uniform, exact
// sigmoid, for every undecided unit of every lane.  in_words: this conditional's input bits of all 128 chains, p-order,
// [half][row]; hh: which 32-unit half of the output layer this warp decides.
__device__ __noinline__ uint32_t fixup_slow_umma(const GibbsUmmaArgs* a, int hidden_out, const uint32_t* in_words, int row, uint32_t nib,
                                                 uint32_t slowmask, uint32_t c0, uint32_t c1, uint32_t sweep, uint32_t q0, int hh, int lane) {
  const StackedParams& p = a->p;
  const int n_in = hidden_out ? p.N : p.nhid();
  uint32_t pending = __ballot_sync(0xffffffffu, slowmask != 0u);
  while (pending != 0u) {
    const int src = __ffs(pending) - 1;
    pending &= pending - 1u;
    uint32_t sm = __shfl_sync(0xffffffffu, slowmask, src);
    const int srow = __shfl_sync(0xffffffffu, row, src);
    const uint32_t sc0 = __shfl_sync(0xffffffffu, c0, src), sc1 = __shfl_sync(0xffffffffu, c1, src);
    const uint32_t m0 = in_words[srow], m1 = in_words[GU_TILE + srow];
    while (sm != 0u) {
      const int pbit = __ffs(sm) - 1;
      sm &= sm - 1u;
      const int unit = gu_unit_of_p(32 * hh + pbit);
      double part = 0.0;
      for (int kin = lane; kin < n_in; kin += 32) {
        const int pk = gu_p_of_unit(kin);
        if (((pk < 32 ? m0 : m1) >> (pk & 31)) & 1u) part += hidden_out ? p.w(unit, kin) : p.w(kin, unit);
      }
      part = warp_sum(part);
      const double x = (hidden_out ? p.hbias(unit) : p.b[unit]) + part;
      const uint32_t c = (uint32_t)(pbit >> 3), half = (uint32_t)((pbit >> 2) & 1), w = (uint32_t)(pbit & 3);
      const uint32_t q = q0 + 4u * (uint32_t)hh + c;
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

__global__ void __launch_bounds__(GU_THREADS, GU_MIN_CTAS) gibbs_umma_kernel(const GibbsUmmaArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const UmmaHeader* hd = reinterpret_cast<const UmmaHeader*>(smem);
  const float* biasH = reinterpret_cast<const float*>(smem + GU_BIAS_OFF);
  const float* biasV = biasH + 64;
  unsigned char* sA = smem + GU_IMAGE_BYTES;                                        // 16 KB: the chains' state bytes
  uint32_t* words = reinterpret_cast<uint32_t*>(smem + GU_IMAGE_BYTES + GU_OPERAND_BYTES);  // [2 buffers][2 halves][128 rows]
  uint64_t* mbar = reinterpret_cast<uint64_t*>(words + 2 * 2 * GU_TILE);            // [0]: image copy, [1]: MMA completion
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 2);
  uint32_t* arrive_cnt = tmem_slot + 1;  // threads that have published their bits for the next conditional update

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int hh = warp >> 2;                  // which 32-unit half of the output layer this thread decides
  const int row = 32 * (warp & 3) + lane;    // chain row of the tile = TMEM lane

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(GU_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (tid == 0) {
    mbar_init(mbar, 1);
    mbar_init(mbar + 1, 1);
    *arrive_cnt = 0u;
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(mbar, (uint32_t)GU_IMAGE_BYTES);
    for (uint32_t off = 0; off < GU_IMAGE_BYTES; off += 16384u) {
      const uint32_t n = GU_IMAGE_BYTES - off < 16384u ? (uint32_t)GU_IMAGE_BYTES - off : 16384u;
      tma_bulk_g2s(smem + off, a.image + off, n, mbar);
    }
  }
  mbar_wait(mbar, 0);
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t tmem = *tmem_slot;
  const uint32_t taddr = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + 32u * (uint32_t)hh;  // this thread's lane, first acc01 column
  const float k1 = hd->inv_scale;
  const int nhid = a.p.nhid(), N = a.p.N;
  const uint32_t qv0 = (uint32_t)(4 * ((nhid + 31) / 32));
  // real (unpadded) units among this thread's 32, in p-order
  uint32_t valid_h = 0u, valid_v = 0u;
#pragma unroll
  for (int pb = 0; pb < 32; ++pb) {
    const int u = gu_unit_of_p(32 * hh + pb);
    valid_h |= (u < nhid ? 1u : 0u) << pb;
    valid_v |= (u < N ? 1u : 0u) << pb;
  }
  const int rowbytes = packed_row_bytes(N);
  // instruction descriptor: D s32, A u8, B s8, both K-major, N = 128, M = 128
  const uint32_t idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint64_t adesc = gu_desc(smem_u32(sA)), bdesc_h = gu_desc(smem_u32(smem + GU_BH_OFF)), bdesc_v = gu_desc(smem_u32(smem + GU_BV_OFF));
  uint32_t mma_phase = 0u;

  const int64_t num_tiles = (a.C + GU_TILE - 1) / GU_TILE;
  for (int64_t ti = blockIdx.x; ti < num_tiles; ti += gridDim.x) {
    const int64_t chain_base = ti * GU_TILE;
    const int64_t ch = chain_base + row;
    const uint64_t id = a.chain_id0 + (uint64_t)ch;
    const uint32_t cid0 = (uint32_t)id, cid1 = (uint32_t)(id >> 32);

    // ---- initial state of this thread's (chain, visible half), natural bit order -> p-order ----
    uint32_t cur = 0u;
    if (a.in_mode == 0) {
      // float64 rows: warp w fetches rows 16 w .. 16 w + 15 of the tile, 256 contiguous bytes per load (two per row), ballots
      // the bits and leaves the two 32-bit words of each row in shared memory (buffer 0 of `words`, which nobody reads at
      // this point); after the barrier every thread picks up the word of its own (chain, half)
#pragma unroll 8
      for (int rr = 0; rr < 16; ++rr) {
        const int r = 16 * warp + rr;
        const int64_t c = chain_base + r;
        unsigned long long d0 = 0ull, d1 = 0ull;
        if (c < a.C) {
          const unsigned long long* rp = reinterpret_cast<const unsigned long long*>(a.v_in) + c * N;
          if (lane < N) d0 = rp[lane];
          if (lane + 32 < N) d1 = rp[lane + 32];
        }
        const uint32_t b0 = __ballot_sync(0xffffffffu, (d0 << 1) != 0ull), b1 = __ballot_sync(0xffffffffu, (d1 << 1) != 0ull);
        if (lane == 0) {
          words[r] = b0;
          words[GU_TILE + r] = b1;
        }
      }
      __syncthreads();
      cur = words[hh * GU_TILE + row];
      // pull this CTA's NEXT tile (one contiguous span of v_in) into L2 while this one is swept
      const int64_t nbase = (ti + gridDim.x) * GU_TILE;
      if (nbase < a.C) {
        const int64_t nbytes = (a.C - nbase < GU_TILE ? a.C - nbase : (int64_t)GU_TILE) * N * (int64_t)sizeof(double);
        const char* np = reinterpret_cast<const char*>(a.v_in + nbase * N);
        for (int64_t off = (int64_t)tid * 128; off < nbytes; off += (int64_t)GU_THREADS * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(np + off));
      }
    } else if (a.in_mode == 1) {
      if (ch < a.C) {
        const uint8_t* rowp = a.bits_in + ch * rowbytes;
        uint32_t x = 0u;
#pragma unroll
        for (int bb = 0; bb < 4; ++bb)
          if (4 * hh + bb < rowbytes) x |= (uint32_t)rowp[4 * hh + bb] << (8 * bb);
        cur = __brev(__byte_perm(x, 0u, 0x0123u));
      }
    } else {
      uint32_t rnd[4];
      philox4x32_10(cid0, cid1, a.sweep0, PHILOX_STREAM_V0 << 29, a.key0, a.key1, rnd);
      cur = hh ? rnd[1] : rnd[0];
    }
    cur = gu_natural_to_p(cur) & valid_v;

#pragma unroll 1
    for (int hs = 0; hs < 2 * a.k; ++hs) {
      const bool back = (hs & 1) != 0;  // false: h | v, true: v | h
      // ---- publish this thread's 32 input bits: A operand bytes [v ; 128 v] and the word for the cold path ----
      uint32_t* in_words = words + (hs & 1) * 2 * GU_TILE;
      in_words[hh * GU_TILE + row] = cur;
      {
        uint4 lo, hi4;
        lo.x = gu_spread4(cur & 0xFu); lo.y = gu_spread4((cur >> 4) & 0xFu); lo.z = gu_spread4((cur >> 8) & 0xFu); lo.w = gu_spread4((cur >> 12) & 0xFu);
        hi4.x = gu_spread4((cur >> 16) & 0xFu); hi4.y = gu_spread4((cur >> 20) & 0xFu); hi4.z = gu_spread4((cur >> 24) & 0xFu); hi4.w = gu_spread4(cur >> 28);
        unsigned char* base = sA + row * 16;
        *reinterpret_cast<uint4*>(base + (2 * hh) * GU_LBO) = lo;
        *reinterpret_cast<uint4*>(base + (2 * hh + 1) * GU_LBO) = hi4;
        lo.x <<= 7; lo.y <<= 7; lo.z <<= 7; lo.w <<= 7;
        hi4.x <<= 7; hi4.y <<= 7; hi4.z <<= 7; hi4.w <<= 7;
        *reinterpret_cast<uint4*>(base + (4 + 2 * hh) * GU_LBO) = lo;
        *reinterpret_cast<uint4*>(base + (4 + 2 * hh + 1) * GU_LBO) = hi4;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core's reads
      asm volatile("tcgen05.fence::before_thread_sync;\n");       // ... and this thread's TMEM reads of the previous accumulators are done
      // No CTA barrier: every thread checks in on a shared counter (release) and carries on with its Philox draws; the LAST one
      // to arrive (acquire: it sees all 256 publications) resets the counter and issues the MMAs.  A barrier here made 24 % of all
      // warp samples wait for the slowest warp of the CTA (ncu, profiles/gibbs_umma_r02b_ncu.txt).
      uint32_t arrived;
      asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(arrived) : "r"(smem_u32(arrive_cnt)) : "memory");
      if (arrived == GU_THREADS - 1) {
        *reinterpret_cast<volatile uint32_t*>(arrive_cnt) = 0u;  // nobody checks in again before this MMA has completed
        asm volatile("tcgen05.fence::after_thread_sync;\n");
        const uint64_t bdesc = back ? bdesc_v : bdesc_h;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)  // consecutive K = 32 slices are 2 LBO apart: only the start-address field moves
          gu_mma(tmem, adesc + (uint64_t)(ks * ((2 * GU_LBO) >> 4)), bdesc + (uint64_t)(ks * ((2 * GU_LBO) >> 4)), idesc, ks > 0 ? 1u : 0u);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar + 1)) : "memory");
      }
      const float* bias = (back ? biasV : biasH) + 32 * hh;
      const float margin = back ? hd->margin_v : hd->margin_h;
      const uint32_t valid = back ? valid_v : valid_h;
      const uint32_t sweep = a.sweep0 + (uint32_t)(hs >> 1);
      const uint32_t q0 = (back ? qv0 : 0u) + 4u * (uint32_t)hh;
      const float c1 = (0x1.0p-16f + margin) - 1.0f;   // s1 = (m + c1) q - 1 < 0             => certainly u < p
      const float d01 = 0x1.0p-16f + 2.0f * margin;    // s0 = d01 q - s1 = 1 - (u_f - margin) q < 0 => certainly u >= p
      // the draws of the first GU_PREFETCH calls are made while the tile's MMAs run (the wait below is where 19 % of all warp
      // samples sat: the join of the tile's 8 warps plus the MMA latency); the remaining calls overlap the TMEM loads
      uint32_t pre[GU_PREFETCH > 0 ? GU_PREFETCH : 1][4];
#pragma unroll
      for (int j = 0; j < GU_PREFETCH; ++j) gu_philox(cid0, cid1, sweep, q0 + (uint32_t)(3 - j), a.rk, pre[j]);
      mbar_wait(mbar + 1, mma_phase);
      mma_phase ^= 1u;
      asm volatile("tcgen05.fence::after_thread_sync;\n");

      uint32_t acc1 = 0u, acc0 = 0u;  // bit p: "certainly 1" / "certainly 0"; units are visited from p = 31 down
#pragma unroll
      for (int c = 3; c >= 0; --c) {
        int a01[8], a2[8];
        gu_ld8(taddr + 8u * (uint32_t)c, a01);
        gu_ld8(taddr + 64u + 8u * (uint32_t)c, a2);
        uint32_t rnd[4];
        if (3 - c < GU_PREFETCH) {
#pragma unroll
          for (int w = 0; w < 4; ++w) rnd[w] = pre[3 - c < GU_PREFETCH ? 3 - c : 0][w];
        } else {
          gu_philox(cid0, cid1, sweep, q0 + (uint32_t)c, a.rk, rnd);  // stream 0; overlaps the TMEM loads
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const float4 b_lo = *reinterpret_cast<const float4*>(bias + 8 * c);
        const float4 b_hi = *reinterpret_cast<const float4*>(bias + 8 * c + 4);
        const float bs[8] = {b_lo.x, b_lo.y, b_lo.z, b_lo.w, b_hi.x, b_hi.y, b_hi.z, b_hi.w};
#pragma unroll
        for (int i = 7; i >= 0; --i) {   // i = 4 h + w
          const int half = i >> 2, w = i & 3;
          const int t = a2[i] * 32768 + a01[i];
          const float y = fmaf((float)t, k1, bs[i]);  // -log2(e) x~
          const float q = 1.0f + gu_ex2(y);
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
        nib = fixup_slow_umma(&a, back ? 0 : 1, in_words, row, nib, slowmask, cid0, cid1, sweep, back ? qv0 : 0u, hh, lane);
      cur = nib;
      asm volatile("tcgen05.fence::before_thread_sync;\n");  // this thread's TMEM reads are ordered before the next MMA (via the barrier)
    }

    // ---- final state: p-order -> natural, float64 rows and / or packed rows ----
    const uint32_t nat = gu_p_to_natural(cur & valid_v);
    if (a.v != nullptr) {
      // float64 rows out: the words go through buffer 1 of `words` (free once every thread has left the last update's cold
      // path), then warp w writes rows 16 w .. 16 w + 15, 256 contiguous bytes per store
      __syncthreads();
      words[2 * GU_TILE + hh * GU_TILE + row] = nat;
      __syncthreads();
#pragma unroll 8
      for (int rr = 0; rr < 16; ++rr) {
        const int r = 16 * warp + rr;
        const int64_t c = chain_base + r;
        if (c < a.C) {
          const uint32_t w0 = words[2 * GU_TILE + r], w1 = words[3 * GU_TILE + r];
          double* rp = a.v + c * N;
          if (lane < N) rp[lane] = (w0 >> lane) & 1u ? 1.0 : 0.0;
          if (lane + 32 < N) rp[lane + 32] = (w1 >> lane) & 1u ? 1.0 : 0.0;
        }
      }
    }
    if (a.bits_out != nullptr && ch < a.C) {
      const uint32_t x = __byte_perm(__brev(nat), 0u, 0x0123u);
      uint8_t* rowp = a.bits_out + ch * rowbytes;
#pragma unroll
      for (int bb = 0; bb < 4; ++bb)
        if (4 * hh + bb < rowbytes) rowp[4 * hh + bb] = (uint8_t)(x >> (8 * bb));
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(GU_TMEM_COLS));
}

static size_t umma_smem_bytes() { return GU_IMAGE_BYTES + GU_OPERAND_BYTES + 2 * 2 * GU_TILE * sizeof(uint32_t) + 2 * 8 + 16; }

size_t gibbs_umma_workspace_bytes(int N, int nhid) {
  if (N <= 32 || N > 64 || nhid <= 32 || nhid > 64) return 0;
  return align_up(GU_IMAGE_BYTES, 256);
}

int64_t gibbs_umma_wave_chains() { return (int64_t)sm_count() * GU_MIN_CTAS * GU_TILE; }

// *handled = 1 when this kernel took the call (both layers 33..64 units, own Philox draws)
int try_gibbs_umma(const StackedParams& p, const ChainIO& io, int64_t C, int k, uint64_t seed, uint64_t chain_id0, uint64_t sweep0,
                   void* ws, size_t ws_bytes, int* handled, cudaStream_t st) {
  *handled = 0;
  const size_t need = gibbs_umma_workspace_bytes(p.N, p.nhid());
  if (need == 0 || ws == nullptr || ws_bytes < need) return QCB_OK;
  unsigned char* img = reinterpret_cast<unsigned char*>(ws);
  gibbs_pack_umma_kernel<<<1, GU_PACK_THREADS, 0, st>>>(p, img);
  QCB_LAUNCH_CHECK();
  GibbsUmmaArgs a;
  a.image = img;
  a.p = p;
  a.v_in = io.v_in;
  a.v = io.v_out;
  a.bits_in = io.bits_in;
  a.bits_out = io.bits_out;
  a.in_mode = io.philox_init ? 2 : (io.bits_in != nullptr ? 1 : 0);
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
  const size_t smem = umma_smem_bytes();
  static int blocks_on[QCB_MAX_DEVICES] = {0};
  int& blocks_per_sm = blocks_on[device_slot()];
  if (blocks_per_sm == 0) {
    int occ = 0;
    QCB_CUDA(cudaFuncSetAttribute(gibbs_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    QCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gibbs_umma_kernel, GU_THREADS, smem));
    if (occ < 1) return fail(QCB_ERR_CUDA, "gibbs (umma) kernel does not fit on an SM");
    if (getenv("QCB_DEBUG_OCCUPANCY") != nullptr) fprintf(stderr, "qcb: gibbs_umma occupancy query = %d CTAs / SM (smem %zu B)\n", occ, smem);
    // 4 CTAs per SM by construction: 64 registers x 256 threads, 52 KB of shared memory and 128 of the 512 TMEM columns each
    // (shared memory keeps a fifth CTA out, so tcgen05.alloc never has to wait).  The occupancy query of this toolkit
    // answers 1 for a kernel that allocates tensor memory; CTAs beyond what is resident simply wait their turn.
    blocks_per_sm = GU_MIN_CTAS;
  }
  const int64_t tiles = (C + GU_TILE - 1) / GU_TILE;
  const int64_t cap = (int64_t)sm_count() * blocks_per_sm;
  const int grid = (int)(tiles < cap ? tiles : cap);
  gibbs_umma_kernel<<<grid, GU_THREADS, smem, st>>>(a);
  QCB_LAUNCH_CHECK();
  *handled = 1;
  return QCB_OK;
}

}  // namespace qcb
