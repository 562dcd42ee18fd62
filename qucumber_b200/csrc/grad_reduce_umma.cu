// Reduced energy gradient of large batches of BINARY rows with BOTH contractions on the 5th-generation tensor cores
// (tcgen05.mma.kind::i8, accumulators in tensor memory), for layers of 33..64 units on both sides:
//   g_partial = sum_rows [ p_r (x) v_r ; v_r ; p_r ],  p_r = sigmoid(W v_r + c)               (rbm/binary_rbm.py:113-120)
// The FP64 tensor-core kernel (grad_reduce_dmma.cu) spends 0.75 ms on 2^20 rows at N = nh = 64 against a DMMA floor of
// 0.23 ms.  For 0/1 rows neither contraction needs FP64 products:
//   phase 1  theta = c + W v  exactly as in gibbs_umma.cu, but with TWO 23-bit fixed-point groups of W (46 bits: the
//            quantisation error of theta stays below 2e-12 max|W|), each group as [d0 ; 2 d1 | d2 ; 0] against the chains'
//            byte rows [v ; 128 v]: 8 x tcgen05.mma per tile of 128 rows, four 64-column accumulators in TMEM;
//   FP64     p = sigmoid(theta) (table logistic of common.cuh) — the only FP64 work left — and P = rint(p (2^48 - 1));
//   phase 2  G = P^T V with P split into six unsigned bytes: the byte tiles are written M-major by the thread that owns the
//            row (M = 2 digits x 64 hidden units, K = 128 rows), V is the SAME shared-memory tile that served as A operand
//            of phase 1, now read N-major (N = 64 visible columns + one all-ones column for the hidden-bias sums); three
//            int32 accumulators (one per digit pair) stay in TMEM for the whole kernel and are read out once:
//            G_ij = sum_d 2^(8 d) G_d,ij / (2^48 - 1).  Every partial sum is an exact integer, so the result is the exact sum of
//            the rounded p: absolute error <= rows 2^-49, i.e. far below 1e-10 of any entry that is not itself negligible.
// The visible-bias gradient is a column count of bits.  A row that is not 0/1, or weights the two groups cannot hold
// (non-finite, max|W| > 8 or < 1e-290), raise *flag and the FP64 kernel queued behind redoes the batch (same protocol as the
// integer phase 1 of grad_reduce_dmma.cu: no host synchronisation, capturable).
#include "rows.cuh"

namespace qcb {

#ifndef GM_SPLIT
#define GM_SPLIT 4                              // threads per row: each owns 64 / GM_SPLIT hidden units (2: 8 warps, 4: 16 warps per CTA)
#endif
constexpr int GM_THREADS = 128 * GM_SPLIT;
constexpr int GM_UPT = 64 / GM_SPLIT;           // hidden units per thread
constexpr int GM_RPW = 128 / (4 * GM_SPLIT);    // rows per warp when a tile is loaded
constexpr int GM_TILE = 128;
constexpr int GM_LBO = 2048;                    // 16-byte chunk stride of a 128-row operand (K-major view)
constexpr int GM_A_BYTES = 10 * GM_LBO;         // chunks 0-3: v, 4: ones, 5: zeros, 6-9: 128 v
constexpr int GM_B1_BYTES = 128 * 128;          // one weight group
constexpr int GM_DIG_BYTES = 128 * 128;         // one digit pair: M = 128, K = 128 rows
constexpr uint32_t GM_TMEM_COLS = 512;
constexpr uint32_t GM_COL_G = 256;              // G accumulators: 3 x 80 columns from here
constexpr long long GM_WQ_MAX = 127LL * 32768 + 63 * 256 + 127;

struct GradUmmaHeader {
  double inv_scale;  // 2^-s
  int ok;            // 0: the weights do not fit the two groups
  int pad;
};
constexpr size_t GM_IMG_HB = 64;                       // double hb[64]
constexpr size_t GM_IMG_B1 = 1024;                     // two groups
constexpr size_t GM_IMAGE_BYTES = GM_IMG_B1 + 2 * GM_B1_BYTES;

__host__ __device__ inline int gm_operand_offset(int row, int k) { return (k >> 4) * GM_LBO + row * 16 + (k & 15); }

// ---- pack: scale, two 23-bit groups, digits, canonical K-major layout (same digit convention as gibbs_umma.cu) ----
__global__ void __launch_bounds__(256) grad_pack_umma_kernel(AffineSpec s, unsigned char* __restrict__ img) {
  __shared__ double red[8];
  __shared__ int bad_s;
  const int N = s.n_in1, H = s.n_out;
  double mw = 0.0;
  bool bad = false;
  for (int idx = threadIdx.x; idx < H * N; idx += blockDim.x) {
    const double a = fabs(s.mop(idx / N, idx % N));
    bad = bad || !(a < 1e290);
    mw = fmax(mw, a);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mw = fmax(mw, __shfl_xor_sync(0xffffffffu, mw, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mw;
  if (threadIdx.x == 0) bad_s = 0;
  __syncthreads();
  if (bad) bad_s = 1;
  __syncthreads();
  mw = 0.0;
  for (int w = 0; w < 8; ++w) mw = fmax(mw, red[w]);
  // two groups resolve theta to 64 * 2^-45 * 2^ilogb(max|W|): keep that below ~2e-11
  const bool ok = !bad_s && !(mw > 0.0 && mw < 1e-290) && mw <= 8.0;
  int sc = 0;
  if (ok && mw > 0.0) {
    sc = 21 - ilogb(mw);
    if (llrint(scalbn(mw, sc)) > GM_WQ_MAX) sc -= 1;
  }
  if (threadIdx.x == 0) {
    GradUmmaHeader* hd = reinterpret_cast<GradUmmaHeader*>(img);
    hd->inv_scale = scalbn(1.0, -sc);
    hd->ok = ok ? 1 : 0;
    hd->pad = 0;
  }
  double* hb = reinterpret_cast<double*>(img + GM_IMG_HB);
  // padded hidden units get a bias far below the logistic function's clamp: p = e^-700, which rounds to zero digits
  for (int i = threadIdx.x; i < 64; i += blockDim.x) hb[i] = i < H ? s.bias(i) : -1.0e4;
  // B operands: group g (0 = high 23 bits, 1 = the next 23), row n, column k:
  //   n < 64: hidden unit n, k < 64: digit 0 of visible k, k >= 64: twice digit 1 of visible k - 64;  n >= 64: digit 2 / zero
  for (int idx = threadIdx.x; idx < 128 * 128; idx += blockDim.x) {
    const int n = idx >> 7, k = idx & 127;
    const int hid = n & 63, vis = k & 63;
    int val[2] = {0, 0};
    if (ok && hid < H && vis < N) {
      double x = scalbn(s.mop(hid, vis), sc);
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const double q = rint(x);
        x = (x - q) * 4194304.0;  // remainder, exact; the next group is 2^22 finer
        long long wq = (long long)q;
        const int d0 = (int)(((wq + 128) & 255) - 128);
        wq = (wq - d0) >> 8;
        const int d1 = (int)(((wq + 64) & 127) - 64);
        const int d2 = (int)((wq - d1) >> 7);
        val[g] = n < 64 ? (k < 64 ? d0 : 2 * d1) : (k < 64 ? d2 : 0);
      }
    }
    img[GM_IMG_B1 + gm_operand_offset(n, k)] = (unsigned char)(val[0] & 255);
    img[GM_IMG_B1 + GM_B1_BYTES + gm_operand_offset(n, k)] = (unsigned char)(val[1] & 255);
  }
}

__device__ __forceinline__ uint64_t gm_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void gm_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void gm_ld8(uint32_t taddr, int (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void gm_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t gm_spread4(uint32_t nib) { return (nib * 0x00204081u) & 0x01010101u; }
// exact int32 -> double through the 2^52 bit pattern
__device__ __forceinline__ double gm_i2d(int t) { return __hiloint2double(0x43300000, (int)((uint32_t)t ^ 0x80000000u)) - 4503601774854144.0; }

// bits != NULL: the rows come packed (ceil(N / 8) bytes per row, numpy.packbits order) and v is not read — the negative phase
// of a CD step takes v_k from the Gibbs kernel this way, 8 bytes per chain instead of 512 (N = 64).
__global__ void __launch_bounds__(GM_THREADS, 1) grad_reduce_umma_kernel(const unsigned char* __restrict__ image, const double* __restrict__ v,
                                                                         const unsigned char* __restrict__ bits, int N, int H, int64_t rows,
                                                                         double* __restrict__ partials, int nslots, int* __restrict__ flag) {
  extern __shared__ __align__(128) unsigned char smem[];
  const GradUmmaHeader* hd = reinterpret_cast<const GradUmmaHeader*>(smem);
  const double* hb = reinterpret_cast<const double*>(smem + GM_IMG_HB);
  unsigned char* sA = smem + GM_IMAGE_BYTES;                       // the tile's state bytes
  unsigned char* sD = sA + GM_A_BYTES;                             // three digit-pair tiles
  uint32_t* words = reinterpret_cast<uint32_t*>(sD + 3 * GM_DIG_BYTES);  // [2 halves][128 rows]
  double* e2tab = reinterpret_cast<double*>(words + 2 * GM_TILE);  // [EXP2_TABLE]
  uint64_t* mbar = reinterpret_cast<uint64_t*>(e2tab + EXP2_TABLE);  // [0]: image, [1]: theta MMAs, [2]: outer-product MMAs, [3]: row tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 4);
  int* colcnt = reinterpret_cast<int*>(tmem_slot + 2);             // [64] column counts of the state bits (visible-bias gradient)
  // [128][N]: the NEXT tile's rows, fetched by TMA while this one is processed (bulk copies need a 16-byte aligned destination)
  double* stage = reinterpret_cast<double*>(smem + align_up((size_t)(reinterpret_cast<unsigned char*>(colcnt + 64) - smem), 128));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int hh = warp >> 2;                 // which GM_UPT-unit slice of the hidden layer (and of the state bits) this thread owns
  const int row = 32 * (warp & 3) + lane;
  const size_t E = (size_t)H * N + N + H;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(GM_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (tid == 0) {
    mbar_init(mbar, 1);
    mbar_init(mbar + 1, 1);
    mbar_init(mbar + 2, 1);
    mbar_init(mbar + 3, 1);
  }
  if (tid < 64) colcnt[tid] = 0;
  fill_exp2_table(e2tab);
  // constant chunks of the state tile: chunk 4 = (1, 0, ..., 0) per row (the all-ones column of phase 2), chunk 5 = zeros
  if (tid < 128) reinterpret_cast<uint4*>(sA + 4 * GM_LBO)[tid] = make_uint4(1u, 0u, 0u, 0u);
  else if (tid < 256) reinterpret_cast<uint4*>(sA + 5 * GM_LBO)[tid - 128] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(mbar, (uint32_t)GM_IMAGE_BYTES);
    for (uint32_t off = 0; off < GM_IMAGE_BYTES; off += 16384u) {
      const uint32_t n = GM_IMAGE_BYTES - off < 16384u ? (uint32_t)GM_IMAGE_BYTES - off : 16384u;
      tma_bulk_g2s(smem + off, image + off, n, mbar);
    }
  }
  mbar_wait(mbar, 0);
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t tmem = *tmem_slot;
  const bool weights_ok = hd->ok != 0;
  const double inv_scale = hd->inv_scale;
  const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;

  // descriptors.  phase 1: A = state bytes, K-major (chunk stride LBO, 8-row groups SBO = 128); B = weight groups, same layout.
  // phase 2: A = digit tiles, M-major (16-unit groups SBO = 2048, 8-row K groups LBO = 128); B = state bytes, N-major, same numbers.
  const uint32_t idesc1 = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t idesc2 = (2u << 4) | (0u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(80 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t sA_addr = smem_u32(sA), sB1_addr = smem_u32(smem + GM_IMG_B1), sD_addr = smem_u32(sD);
  uint32_t ph1 = 0u, ph2 = 0u;
  bool g_started = false;
  int cnt0 = 0, cnt1 = 0;  // this lane's column counts (columns lane and lane + 32) over the rows its warp loads

  const int64_t ntiles = (rows + GM_TILE - 1) / GM_TILE;
  uint32_t ph3 = 0u;
  const bool base_aligned = (reinterpret_cast<uintptr_t>(v) & 15u) == 0u;
  auto tile_bytes = [&](int64_t t) -> uint32_t {
    const int64_t left = rows - t * GM_TILE;
    return (uint32_t)((left < GM_TILE ? left : (int64_t)GM_TILE) * N * (int64_t)sizeof(double));
  };
  auto tile_staged = [&](int64_t t) { return bits == nullptr && base_aligned && ((tile_bytes(t) & 15u) == 0u) && (((int64_t)GM_TILE * N) & 1) == 0; };
  const int rowbytes = packed_row_bytes(N);
  const bool bits_wide = rowbytes == 8 && (reinterpret_cast<uintptr_t>(bits) & 7u) == 0u;
  const uint32_t hi_mask = N >= 64 ? 0xffffffffu : ((1u << (N - 32)) - 1u);
  auto stage_fetch = [&](int64_t t) {  // one thread: bulk copies of at most 16 KB each, all completing on mbar[3]
    if (t >= ntiles || !tile_staged(t)) return;
    const uint32_t bytes = tile_bytes(t);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic reads of the staging rows vs. the async writes
    mbar_expect_tx(mbar + 3, bytes);
    const char* src = reinterpret_cast<const char*>(v) + t * (int64_t)GM_TILE * N * (int64_t)sizeof(double);
    for (uint32_t off = 0; off < bytes; off += 16384u)
      tma_bulk_g2s(reinterpret_cast<char*>(stage) + off, src + off, bytes - off < 16384u ? bytes - off : 16384u, mbar + 3);
  };
  if (tid == 0 && weights_ok) stage_fetch(blockIdx.x);
  bool give_up = !weights_ok;
  int why = 2;  // what ends up in *flag when this CTA gives up: 2 = weights outside the two groups, 3 = a row that is not 0/1
  for (int64_t tile = blockIdx.x; tile < ntiles && !give_up; tile += gridDim.x) {
    const int64_t r0 = tile * GM_TILE;
    // ---- rows -> bits.  A tile's rows are one contiguous span: it was handed to the TMA engine a whole tile ago (16-byte
    // aligned spans; anything else is read directly).  Warp w turns rows 16 w .. 16 w + 15 into two ballot words each.
    bool nonbinary = false;
    const bool staged = tile_staged(tile);
    if (staged) {
      mbar_wait(mbar + 3, ph3);
      ph3 ^= 1u;
    }
    if (bits != nullptr) {
      // packed rows: lanes 0..15 fetch one row each (one 8-byte load at N > 56), undo the MSB-first bit order of the bytes and
      // publish the two words; the column counts come from 16 broadcasts
      uint32_t w0 = 0u, w1 = 0u;
      const int64_t gr = r0 + GM_RPW * warp + lane;
      if (lane < GM_RPW && gr < rows) {
        const unsigned char* rp = bits + gr * rowbytes;
        if (bits_wide) {
          const uint2 x = *reinterpret_cast<const uint2*>(rp);
          w0 = x.x;
          w1 = x.y;
        } else {
          for (int b = 0; b < rowbytes; ++b) {
            const uint32_t byte = rp[b];
            if (b < 4) w0 |= byte << (8 * b);
            else w1 |= byte << (8 * (b - 4));
          }
        }
        w0 = __byte_perm(__brev(w0), 0u, 0x0123u);  // unit 8 b + i = bit 7 - i of byte b  ->  bit 8 b + i
        w1 = __byte_perm(__brev(w1), 0u, 0x0123u) & hi_mask;
      }
      if (lane < GM_RPW) {
        words[GM_RPW * warp + lane] = w0;
        words[GM_TILE + GM_RPW * warp + lane] = w1;
      }
#pragma unroll
      for (int rr = 0; rr < GM_RPW; ++rr) {
        cnt0 += (int)((__shfl_sync(0xffffffffu, w0, rr) >> lane) & 1u);
        cnt1 += (int)((__shfl_sync(0xffffffffu, w1, rr) >> lane) & 1u);
      }
    } else {
#pragma unroll 8
    for (int rr = 0; rr < GM_RPW; ++rr) {
      const int r = GM_RPW * warp + rr;
      const int64_t gr = r0 + r;
      unsigned long long d0 = 0ull, d1 = 0ull;
      if (gr < rows) {
        const unsigned long long* rp = staged ? reinterpret_cast<const unsigned long long*>(stage) + r * N
                                              : reinterpret_cast<const unsigned long long*>(v) + gr * N;
        if (lane < N) d0 = rp[lane];
        if (lane + 32 < N) d1 = rp[lane + 32];
      }
      nonbinary |= ((d0 << 1) != 0ull && d0 != 0x3FF0000000000000ull) | ((d1 << 1) != 0ull && d1 != 0x3FF0000000000000ull);
      const uint32_t b0 = __ballot_sync(0xffffffffu, (d0 << 1) != 0ull), b1 = __ballot_sync(0xffffffffu, (d1 << 1) != 0ull);
      cnt0 += (int)((b0 >> lane) & 1u);
      cnt1 += (int)((b1 >> lane) & 1u);
      if (lane == 0) {
        words[r] = b0;
        words[GM_TILE + r] = b1;
      }
    }
    }
    if (tid == 0) nonbinary |= *reinterpret_cast<volatile int*>(flag) != 0;  // another CTA gave up already
    if (__syncthreads_or(nonbinary)) {
      give_up = true;
      why = 3;
      break;
    }
    if (tid == 0) stage_fetch(tile + gridDim.x);  // the staging rows have been consumed (barrier above): fetch the next tile
    if (g_started) {  // the previous tile's outer-product MMAs still read the state bytes and the digit tiles
      mbar_wait(mbar + 2, ph2);
      ph2 ^= 1u;
    }
    if (GM_SPLIT == 2) {  // publish this thread's 32 bits as bytes: [v ; 128 v]
      const uint32_t cur = words[hh * GM_TILE + row];
      uint4 lo, hi4;
      lo.x = gm_spread4(cur & 0xFu); lo.y = gm_spread4((cur >> 4) & 0xFu); lo.z = gm_spread4((cur >> 8) & 0xFu); lo.w = gm_spread4((cur >> 12) & 0xFu);
      hi4.x = gm_spread4((cur >> 16) & 0xFu); hi4.y = gm_spread4((cur >> 20) & 0xFu); hi4.z = gm_spread4((cur >> 24) & 0xFu); hi4.w = gm_spread4(cur >> 28);
      unsigned char* base = sA + row * 16;
      *reinterpret_cast<uint4*>(base + (2 * hh) * GM_LBO) = lo;
      *reinterpret_cast<uint4*>(base + (2 * hh + 1) * GM_LBO) = hi4;
      lo.x <<= 7; lo.y <<= 7; lo.z <<= 7; lo.w <<= 7;
      hi4.x <<= 7; hi4.y <<= 7; hi4.z <<= 7; hi4.w <<= 7;
      *reinterpret_cast<uint4*>(base + (6 + 2 * hh) * GM_LBO) = lo;
      *reinterpret_cast<uint4*>(base + (6 + 2 * hh + 1) * GM_LBO) = hi4;
    } else {              // four threads per row: 16 bits each, one 16-byte chunk of v and one of 128 v
      const uint32_t cur = (words[(hh >> 1) * GM_TILE + row] >> (16 * (hh & 1))) & 0xFFFFu;
      uint4 lo;
      lo.x = gm_spread4(cur & 0xFu); lo.y = gm_spread4((cur >> 4) & 0xFu); lo.z = gm_spread4((cur >> 8) & 0xFu); lo.w = gm_spread4(cur >> 12);
      unsigned char* base = sA + row * 16;
      *reinterpret_cast<uint4*>(base + hh * GM_LBO) = lo;
      lo.x <<= 7; lo.y <<= 7; lo.z <<= 7; lo.w <<= 7;
      *reinterpret_cast<uint4*>(base + (6 + hh) * GM_LBO) = lo;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;\n");
#pragma unroll
      for (int g = 0; g < 2; ++g)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const int chunk = ks < 2 ? 2 * ks : 6 + 2 * (ks - 2);  // K slices {0,1}, {2,3} (v) and {6,7}, {8,9} (128 v)
          gm_mma(tmem + 128u * (uint32_t)g, gm_desc(sA_addr + chunk * GM_LBO, GM_LBO, 128),
                 gm_desc(sB1_addr + g * GM_B1_BYTES + ks * 2 * GM_LBO, GM_LBO, 128), idesc1, ks > 0 ? 1u : 0u);
        }
      gm_commit(mbar + 1);
    }
    mbar_wait(mbar + 1, ph1);
    ph1 ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;\n");

    // ---- theta -> p -> six bytes of rint(p 2^48), written M-major into the digit tiles
    // P = rint(p (2^48 - 1)) fits six bytes for every p in [0, 1] (no clamp); rows beyond the batch get a zero scale
    const double pscale = r0 + row < rows ? 281474976710655.0 : 0.0;
    const uint32_t tcol = tmem + lane_base + (uint32_t)(GM_UPT * hh);
#pragma unroll 1
    for (int c = 0; c < GM_UPT / 8; ++c) {
      int h01[8], h2[8], l01[8], l2[8];
      gm_ld8(tcol + 8u * (uint32_t)c, h01);
      gm_ld8(tcol + 64u + 8u * (uint32_t)c, h2);
      gm_ld8(tcol + 128u + 8u * (uint32_t)c, l01);
      gm_ld8(tcol + 192u + 8u * (uint32_t)c, l2);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      unsigned long long P[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int unit = GM_UPT * hh + 8 * c + i;
        const double thi = gm_i2d(h2[i] * 32768 + h01[i]), tlo = gm_i2d(l2[i] * 32768 + l01[i]);
        const double theta = fma(fma(tlo, 2.384185791015625e-07, thi), inv_scale, hb[unit]);  // 2^-22
        P[i] = __double2ull_rn(sigmoid_tab(theta, e2tab) * pscale);
      }
      // byte d of the eight values -> eight consecutive bytes of digit d's M-major tile
      const int u0 = GM_UPT * hh + 8 * c;  // first unit of the chunk: m = (d & 1) * 64 + u0 + i
#pragma unroll
      for (int d = 0; d < 6; ++d) {
        // byte d of P[0..3] and of P[4..7] -> two words: three byte permutes each (bytes 0..3 live in the low word, 4..5 in the high one)
        uint32_t w[2];
#pragma unroll
        for (int q4 = 0; q4 < 2; ++q4) {
          const uint32_t x0 = d < 4 ? (uint32_t)P[4 * q4] : (uint32_t)(P[4 * q4] >> 32), x1 = d < 4 ? (uint32_t)P[4 * q4 + 1] : (uint32_t)(P[4 * q4 + 1] >> 32);
          const uint32_t x2 = d < 4 ? (uint32_t)P[4 * q4 + 2] : (uint32_t)(P[4 * q4 + 2] >> 32), x3 = d < 4 ? (uint32_t)P[4 * q4 + 3] : (uint32_t)(P[4 * q4 + 3] >> 32);
          const uint32_t sel = (uint32_t)(d & 3) | ((uint32_t)(4 + (d & 3)) << 4);  // byte (d & 3) of the first, then of the second operand
          w[q4] = __byte_perm(__byte_perm(x0, x1, sel), __byte_perm(x2, x3, sel), 0x5410u);
        }
        const int m = (d & 1) * 64 + u0;
        unsigned char* dst = sD + (d >> 1) * GM_DIG_BYTES + (m >> 4) * GM_LBO + row * 16 + (m & 15);
        *reinterpret_cast<uint2*>(dst) = make_uint2(w[0], w[1]);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;\n");
#pragma unroll
      for (int pr = 0; pr < 3; ++pr)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)  // K = rows: 32 rows = 4 groups of 8 = 512 bytes per slice
          gm_mma(tmem + GM_COL_G + 80u * (uint32_t)pr, gm_desc(sD_addr + pr * GM_DIG_BYTES + ks * 512, 128, GM_LBO),
                 gm_desc(sA_addr + ks * 512, 128, GM_LBO), idesc2, (g_started || ks > 0) ? 1u : 0u);
      gm_commit(mbar + 2);
    }
    g_started = true;
  }

  // ---- outcome
  double* out = partials + (size_t)blockIdx.x * E;
  if (give_up) {
    if (tid == 0) *flag = why;
    if (g_started) mbar_wait(mbar + 2, ph2);  // no MMA may be in flight when the tensor memory is released
  } else {
    if (g_started) {
      mbar_wait(mbar + 2, ph2);
      asm volatile("tcgen05.fence::after_thread_sync;\n");
    }
    // G accumulators: TMEM lane m = (d & 1) * 64 + hidden unit, columns [256 + 80 pr, +80): 64 visible columns, the ones column.
    // Warps 0..3 read their lane quarter, scale each digit and add the three pairs; the two digit halves (lanes m and m + 64)
    // meet in shared memory.
    double* gsm = reinterpret_cast<double*>(sA);  // [128][66]: the state and digit tiles (68 KB, contiguous) are free now
    if (warp < 4) {
      const int m = row;
#pragma unroll 1
      for (int c0 = 0; c0 < 80; c0 += 8) {
        double acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.0;
        if (g_started) {
#pragma unroll
          for (int pr = 0; pr < 3; ++pr) {
            int gi[8];
            gm_ld8(tmem + lane_base + GM_COL_G + 80u * (uint32_t)pr + (uint32_t)c0, gi);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const double w = scalbn(1.0, 8 * (2 * pr + (m >> 6))) / 281474976710655.0;  // digit weight 2^(8 d) over the scale 2^48 - 1
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fma(gm_i2d(gi[i]), w, acc[i]);
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (c0 + i < 66) gsm[m * 66 + c0 + i] = acc[i];
      }
    }
    // column counts of the state bits: one shared-memory atomic per lane and column (integers: order does not matter)
    atomicAdd(colcnt + lane, cnt0);
    atomicAdd(colcnt + 32 + lane, cnt1);
    __syncthreads();
    for (int idx = tid; idx < H * N; idx += GM_THREADS) {
      const int i = idx / N, j = idx - i * N;
      out[idx] = gsm[i * 66 + j] + gsm[(64 + i) * 66 + j];
    }
    if (tid < N) out[(size_t)H * N + tid] = (double)colcnt[tid];
    if (tid < H) out[(size_t)H * N + N + tid] = gsm[tid * 66 + 64] + gsm[(64 + tid) * 66 + 64];
    // the FP64 kernel behind this one sizes its grid differently: the slots it would have filled beyond ours count as zero
    for (int slot = gridDim.x + blockIdx.x; slot < nslots; slot += gridDim.x)
      for (size_t idx = tid; idx < E; idx += GM_THREADS) partials[(size_t)slot * E + idx] = 0.0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(GM_TMEM_COLS));
}

static size_t gm_smem_bytes() {
  return GM_IMAGE_BYTES + GM_A_BYTES + 3 * GM_DIG_BYTES + 2 * GM_TILE * sizeof(uint32_t) + EXP2_TABLE * sizeof(double) + 4 * 8 + 8 + 64 * sizeof(int) +
         128 + (size_t)GM_TILE * 64 * sizeof(double);
}

size_t grad_umma_workspace_bytes() { return align_up(GM_IMAGE_BYTES, 256); }

// Launches pack + kernel when the shape fits; *launched = 1 then (the caller queues the FP64 kernel behind it with the same flag).
// img: device scratch of grad_umma_workspace_bytes(); nslots: partial slots the finalize pass will sum.
bool grad_umma_shape(int N, int H, int64_t rows) { return N >= 33 && N <= 64 && H >= 33 && H <= 64 && rows >= 16384; }

int try_grad_reduce_umma(const AffineSpec& s, const double* v, const unsigned char* bits, int H, int64_t rows, double* partials, int nslots,
                         int* flag, unsigned char* img, int* launched, cudaStream_t st) {
  *launched = 0;
  const int N = s.n_in1;
  // both layers 33..64 units: below that the FP64 tensor-core kernel, whose tiles shrink with the layer, is faster (measured)
  if (flag == nullptr || img == nullptr || H != s.n_out || !grad_umma_shape(N, H, rows)) return QCB_OK;
  const int grid = sm_count() < nslots ? sm_count() : nslots;
  if (rows / grid > (int64_t)8000000) return QCB_OK;  // int32 accumulators: 255 * rows per CTA must stay below 2^31
  const size_t smem = gm_smem_bytes();
  static bool configured[QCB_MAX_DEVICES] = {false};
  if (!configured[device_slot()]) {
    QCB_CUDA(cudaFuncSetAttribute(grad_reduce_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[device_slot()] = true;
  }
  grad_pack_umma_kernel<<<1, 256, 0, st>>>(s, img);
  QCB_LAUNCH_CHECK();
  grad_reduce_umma_kernel<<<grid, GM_THREADS, smem, st>>>(img, v, bits, N, H, rows, partials, nslots, flag);
  QCB_LAUNCH_CHECK();
  *launched = 1;
  return QCB_OK;
}

}  // namespace qcb
