// Block Gibbs with an integer-exact tensor-core filter (layers up to 64 x 64, in-kernel Philox draws).
// Like gibbs.cu it replaces the k-sweep loops of BinaryRBM.gibbs_steps (rbm/binary_rbm.py:200-228: h | v then
// v | h through prob_*/sample_*, :129-198) and PurificationRBM.gibbs_steps (rbm/purification_rbm.py:309-337) with
// one launch; the PurificationRBM runs as the stacked hidden layer [W; U], [c; d].
//
// A Bernoulli decision u < sigmoid(x) only needs x in full FP64 when u falls next to p.  This kernel
// therefore evaluates x~ = W v + c with EXACT integer arithmetic on the int8 tensor cores and keeps the
// FP64 evaluation for the rare undecided units, so the samples are those of the all-FP64 kernel
// (gibbs.cu) — same Philox convention, same decisions — at a fraction of the FP64 work:
//
//  * W is quantised once per call to 24-bit fixed point, Wq = rint(W 2^s) = d2 2^15 + d1 2^8 + d0 with signed
//    digits of 8, 7 and 8 bits; a state bit is 0 or 1, so  t = sum_k v_k Wq_ik  is an exact int32 computed by
//    3 x mma.sync.m16n8k32 (u8 x s8, IMMA) passes: d2 first, one shift by 15, then 2 d1 against the state
//    bits scaled to 128 and d0 against the plain bits.  x~ = float(t) 2^-s + c.
//  * |x~ - x| <= n_in 2^-(s+1) + 2^-22 (sum_k |W_ik| + |c_i|) =: dx  is a rigorous bound (quantisation +
//    the two float roundings), hence |sigmoid(x~) - sigmoid(x)| <= dx / 4.  The filter decides from the
//    16-bit cell of u unless p~ lies within margin = dx/4 + 2^-19 of it (the float evaluation of the test is good to 2^-21.5);
//    undecided units (~1e-4 of them) recompute x in FP64 from the original doubles and compare the full
//    53-bit uniform with the exact sigmoid (fixup_slow_imma, out of line).
//  * a warp owns 16 chains (two rows of the M = 16 tile per lane); states are bit masks in registers,
//    the A operand is built from mask nibbles with one multiply-and-mask, the B image (36 KB for
//    64 x 64, both directions) sits in shared memory, staged by one TMA bulk copy.
//
// Fed uniforms and larger layers use the FP64 DMMA kernel.
#include "common.cuh"

namespace qcb {

#ifndef QCB_GI_THREADS  // overridable for occupancy experiments (experiments/gibbs_imma_variants.sh)
#define QCB_GI_THREADS 256
#define QCB_GI_MIN_CTAS 2
#endif
constexpr int GI_THREADS = QCB_GI_THREADS;
constexpr int GI_WARPS = GI_THREADS / 32;
constexpr int GI_CPW = 16;  // chains per warp

struct GibbsImmaArgs {
  const unsigned char* image;
  uint32_t image_bytes;
  StackedParams p;        // original FP64 parameters (cold path)
  const double* v_in;
  double* v;
  const uint8_t* bits_in;  // packed rows in (in_mode 1)
  uint8_t* bits_out;       // packed rows out (optional, next to / instead of v)
  int in_mode;             // 0: v_in (float64 rows), 1: bits_in, 2: v0 ~ Bernoulli(1/2) from Philox stream 5
  int64_t C;
  int k;
  uint32_t key0, key1;
  uint64_t chain_id0;
  uint32_t sweep0;
  unsigned long long* dbg;  // optional: [0] = max over cold-path units of |x~ - x| / dx scaled by 2^20, [1] = cold-path units
};

struct ImmaHeader {
  float inv_scale;   // -log2(e) * 2^-s: the kernel works with y = -log2(e) x so that e^{-x} = 2^y
  float margin_h, margin_v;
  float dx_h, dx_v;
  int s;
  int pad[2];
};

__host__ __device__ inline int imma_kk(int units_pad) { return (units_pad + 31) / 32; }

// Fixed-point weights are split as Wq = d2 2^15 + d1 2^8 + d0 with signed digits d0 (8 bits), d1 (7 bits), d2 (8 bits)
constexpr long long IMMA_WQ_MAX = 127ll * 32768 + 63 * 256 + 127;

// bytes of the image for HB / VB 16-unit blocks
static size_t imma_image_bytes(int HB, int VB) {
  const int Hpad = 16 * HB, Npad = 16 * VB;
  const size_t bh = (size_t)3 * imma_kk(Npad) * (Hpad / 8) * 32 * 8;  // h | v : K over visible, N over hidden
  const size_t bv = (size_t)3 * imma_kk(Hpad) * (Npad / 8) * 32 * 8;
  return align_up(sizeof(ImmaHeader) + bh + bv + (size_t)(Hpad + Npad) * 4, 16);
}

// unit owned by column c (0..7) of n-block nb: lane t ends up with the eight units of one Philox call
__host__ __device__ inline int imma_unit(int nb, int c) { return 32 * (nb >> 2) + 16 * ((nb >> 1) & 1) + 4 * (c >> 1) + 2 * (nb & 1) + (c & 1); }

// ---- pack: quantise, split into int8 digits, write fragment-ordered image ------------------------
// One CTA (the global maximum fixes the scale); the parameters are staged in shared memory once so that
// every later pass (maxima, abs-sums, the three digit images) reads them from there.
constexpr int GI_PACK_LD = 65;  // row stride of the staged weights (odd: column walks are conflict-free)
__global__ void __launch_bounds__(1024) gibbs_pack_imma_kernel(StackedParams p, int HB, int VB, unsigned char* __restrict__ img) {
  __shared__ double Ws[64 * GI_PACK_LD];
  __shared__ double red[3][32];
  const int H = p.nhid(), N = p.N;
  const int Hpad = 16 * HB, Npad = 16 * VB;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double mw = 0.0;
  for (int idx = threadIdx.x; idx < H * N; idx += blockDim.x) {
    const int i = idx / N, j = idx - i * N;
    const double w = p.w(i, j);
    Ws[i * GI_PACK_LD + j] = w;
    mw = fmax(mw, fabs(w));
  }
  __syncthreads();
  double rm = 0.0, cm = 0.0;
  if (threadIdx.x < H) {  // |bias| + abs-sum of one row
    const int i = threadIdx.x;
    rm = fabs(p.hbias(i));
    for (int j = 0; j < N; ++j) rm += fabs(Ws[i * GI_PACK_LD + j]);
  } else if (threadIdx.x >= 64 && threadIdx.x < 64 + N) {  // ... of one column (other warps)
    const int j = threadIdx.x - 64;
    cm = fabs(p.b[j]);
    for (int i = 0; i < H; ++i) cm += fabs(Ws[i * GI_PACK_LD + j]);
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
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
    s_maxw = fmax(s_maxw, red[0][w]);
    s_rowmax = fmax(s_rowmax, red[1][w]);
    s_colmax = fmax(s_colmax, red[2][w]);
  }
  int s = 0;
  if (s_maxw > 0.0 && s_maxw < 1e300) s = 21 - ilogb(s_maxw);  // |Wq| <= 2^22
  if (s > 60) s = 60;
  if (s_maxw > 0.0 && s_maxw < 1e300 && llrint(s_maxw * ldexp(1.0, s)) > IMMA_WQ_MAX) s -= 1;  // keep the top digit inside int8
  // weights beyond 2^81 (or non-finite) do not fit the fixed-point image: the digits are zeroed and the margins
  // set to 4, which makes every unit undecided, i.e. the kernel degenerates to its exact FP64 path
  const bool degenerate = !(s_maxw < 1e300) || s < -60;
  if (s < -60) s = -60;
  const double scale = degenerate ? 0.0 : ldexp(1.0, s);
  ImmaHeader* hd = reinterpret_cast<ImmaHeader*>(img);
  if (threadIdx.x == 0) {
    hd->inv_scale = (float)(-1.4426950408889634 * ldexp(1.0, -s));
    hd->s = s;
    const double dxh = N * ldexp(1.0, -(s + 1)) + ldexp(1.0, -21) * s_rowmax;
    const double dxv = H * ldexp(1.0, -(s + 1)) + ldexp(1.0, -21) * s_colmax;
    hd->dx_h = (float)dxh;
    hd->dx_v = (float)dxv;
    // round the margins up a little: they are compared in float
    const double mh = (dxh * 0.25 + ldexp(1.0, -19)) * 1.0001, mv = (dxv * 0.25 + ldexp(1.0, -19)) * 1.0001;
    hd->margin_h = (degenerate || !(mh < 4.0)) ? 4.0f : (float)mh;  // a margin >= 1 leaves every unit undecided
    hd->margin_v = (degenerate || !(mv < 4.0)) ? 4.0f : (float)mv;
    hd->pad[0] = hd->pad[1] = 0;
  }
  const int KKv = imma_kk(Npad), NBh = Hpad / 8, KKh = imma_kk(Hpad), NBv = Npad / 8;
  uint32_t* Bh = reinterpret_cast<uint32_t*>(img + sizeof(ImmaHeader));
  const size_t nBh = (size_t)3 * KKv * NBh * 32 * 2;
  uint32_t* Bv = Bh + nBh;
  const size_t nBv = (size_t)3 * KKh * NBv * 32 * 2;
  float* biasH = reinterpret_cast<float*>(Bv + nBv);
  float* biasV = biasH + Hpad;
  // one uint32 = 4 consecutive k of one slice; index = ((((slice * KK + kk) * NB/2 + nb/2) * 32 + lane) * 2 + nb%2) * 2 + half
  const size_t perH = nBh / 3, perV = nBv / 3;  // words of one digit slice
  for (size_t idx = threadIdx.x; idx < perH + perV; idx += blockDim.x) {
    const bool hside = idx < perH;
    const size_t within = hside ? idx : idx - perH;
    size_t r = within;
    const int NB = hside ? NBh : NBv;
    const int half = r & 1; r >>= 1;
    const int nblo = r & 1; r >>= 1;  // two n-blocks share one 16-byte fragment load
    const int ln = r & 31; r >>= 5;
    const int nb = 2 * (int)(r % (NB / 2)) + nblo;
    const int kk = (int)(r / (NB / 2));
    const int g = ln >> 2, t = ln & 3;
    const int outu = imma_unit(nb, g);
    uint32_t word[3] = {0u, 0u, 0u};
    for (int i = 0; i < 4; ++i) {
      const int inu = 32 * kk + 16 * half + 4 * t + i;
      const int hid = hside ? outu : inu, vis = hside ? inu : outu;
      if (!degenerate && hid < H && vis < N) {
        long long wq = llrint(Ws[hid * GI_PACK_LD + vis] * scale);
        const int q0 = (int)(((wq + 128) & 255) - 128);
        wq = (wq - q0) >> 8;
        const int q1 = (int)(((wq + 64) & 127) - 64);   // 7 bits: stored doubled, multiplied by the 0 / 128 state operand
        const int q2 = (int)((wq - q1) >> 7);           // |q2| <= 127 because |Wq| <= IMMA_WQ_MAX
        word[0] |= (uint32_t)(q0 & 255) << (8 * i);
        word[1] |= (uint32_t)((2 * q1) & 255) << (8 * i);
        word[2] |= (uint32_t)(q2 & 255) << (8 * i);
      }
    }
    uint32_t* dst = hside ? Bh : Bv;
    const size_t per = hside ? perH : perV;
    for (int slice = 0; slice < 3; ++slice) dst[slice * per + within] = word[slice];
  }
  for (int i = threadIdx.x; i < Hpad; i += blockDim.x) biasH[i] = i < H ? (float)(-1.4426950408889634 * p.hbias(i)) : 0.0f;
  for (int j = threadIdx.x; j < Npad; j += blockDim.x) biasV[j] = j < N ? (float)(-1.4426950408889634 * p.b[j]) : 0.0f;
}

__device__ __forceinline__ void imma16832(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  // A (state bits, 0 / 1 / 128) unsigned, B (weight digits) signed
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// four mask bits -> four int8 (0/1) in one register
__device__ __forceinline__ uint32_t spread4(uint32_t nib) { return (nib * 0x00204081u) & 0x01010101u; }

__device__ __forceinline__ uint32_t stream_ctr_i(uint32_t stream, uint32_t q) { return (stream << 29) | q; }

// Cold path, executed by the whole warp for every lane that has undecided units: exact FP64 pre-activation
// from the original parameters (each lane adds the weights of its two input units, then a warp reduction),
// full 53-bit uniform, exact sigmoid.  in0 / in1: the row's input-layer bits (identical on the four lanes
// of a row); hidden_out: which conditional.  Must be called with all 32 lanes converged.
__device__ __noinline__ uint32_t fixup_slow_imma(const GibbsImmaArgs* a, int hidden_out, uint32_t in0, uint32_t in1, uint32_t nibw,
                                                 uint32_t slowmask, uint32_t c0, uint32_t c1, uint32_t sweep, uint32_t q0, int lane,
                                                 const float* xs, float dx) {
  const StackedParams& p = a->p;
  const int n_in = hidden_out ? p.N : p.nhid();
  uint32_t pending = __ballot_sync(0xffffffffu, slowmask != 0u);
  while (pending != 0u) {
    const int src = __ffs(pending) - 1;
    pending &= pending - 1u;
    uint32_t sm = __shfl_sync(0xffffffffu, slowmask, src);
    const uint32_t m0 = __shfl_sync(0xffffffffu, in0, src), m1 = __shfl_sync(0xffffffffu, in1, src);
    const uint32_t sc0 = __shfl_sync(0xffffffffu, c0, src), sc1 = __shfl_sync(0xffffffffu, c1, src);
    const int ts = src & 3;
    while (sm != 0u) {
      const int bit = __ffs(sm) - 1;
      sm &= sm - 1u;
      const int gq = bit >> 2, w = bit & 3;
      const int unit = 16 * gq + 4 * ts + w;
      double part = 0.0;
      for (int kin = lane; kin < n_in; kin += 32)
        if (((kin < 32 ? m0 : m1) >> (kin & 31)) & 1u) part += hidden_out ? p.w(unit, kin) : p.w(kin, unit);
      part = warp_sum(part);
      const double x = (hidden_out ? p.hbias(unit) : p.b[unit]) + part;
      const uint32_t half = (uint32_t)(gq & 1), q = q0 + 4u * (uint32_t)(gq >> 1) + (uint32_t)ts;
      uint32_t hi[4], la[4], lb[4];
      philox4x32_10(sc0, sc1, sweep, stream_ctr_i(0u, q), a->key0, a->key1, hi);
      philox4x32_10(sc0, sc1, sweep, stream_ctr_i(1u + 2u * half, q), a->key0, a->key1, la);
      philox4x32_10(sc0, sc1, sweep, stream_ctr_i(2u + 2u * half, q), a->key0, a->key1, lb);
      const uint32_t hw = w == 0 ? hi[0] : (w == 1 ? hi[1] : (w == 2 ? hi[2] : hi[3]));
      const uint32_t aw = w == 0 ? la[0] : (w == 1 ? la[1] : (w == 2 ? la[2] : la[3]));
      const uint32_t bw = w == 0 ? lb[0] : (w == 1 ? lb[1] : (w == 2 ? lb[2] : lb[3]));
      const uint64_t h16 = (hw >> (16u * (1u - half))) & 0xFFFFu;
      const uint64_t bits = (h16 << 37) | ((uint64_t)aw << 5) | (uint64_t)(bw >> 27);
      const double u = (double)bits * 0x1.0p-53;
      const bool one = u < sigmoid_exact(x);
      if (lane == src) {
        if (a->dbg != nullptr && xs != nullptr) {
          const double ratio = fabs((double)xs[bit] - x) / (double)dx;
          atomicMax(a->dbg, (unsigned long long)(ratio * 1048576.0));
          atomicAdd(a->dbg + 1, 1ull);
        }
        nibw = one ? (nibw | (1u << bit)) : (nibw & ~(1u << bit));
      }
    }
  }
  return nibw;
}

__device__ __forceinline__ float ex2_approx(float y) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(y));
  return r;
}

// One conditional update of 16 chains.  KB_IN / OB: 16-unit blocks of the input / output layer.
// DBG: keep the filter's pre-activations for the error-bound check of the cold path (test hook only).
template <int KB_IN, int OB, bool DBG>
__device__ __forceinline__ void half_sweep_imma(const uint32_t (&min)[2][(KB_IN + 1) / 2], uint32_t (&mout)[2][(OB + 1) / 2],
                                                const uint2* __restrict__ Bimg, const float* __restrict__ bias, float k1,
                                                float margin, float dx, uint32_t valid_mask, const uint32_t (&cid)[2][2], uint32_t key0,
                                                uint32_t key1, uint32_t sweep, uint32_t q0, const GibbsImmaArgs* args,
                                                int hidden_out, int lane) {
  constexpr int NB = 2 * OB;                 // n-blocks of 8 output units
  constexpr int KK = (KB_IN + 1) / 2;        // k-steps of 32 input units (= mask words)
  const int t = lane & 3;
  // A operand: the two chains' state bits as int8, built once per k-step
  uint32_t afr[KK][4];
#pragma unroll
  for (int kk = 0; kk < KK; ++kk) {
    afr[kk][0] = spread4((min[0][kk] >> (4 * t)) & 0xFu);
    afr[kk][1] = spread4((min[1][kk] >> (4 * t)) & 0xFu);
    afr[kk][2] = spread4((min[0][kk] >> (16 + 4 * t)) & 0xFu);
    afr[kk][3] = spread4((min[1][kk] >> (16 + 4 * t)) & 0xFu);
  }
  // t = ((digit2 * 256 + digit1) * 256 + digit0) . v : the running total re-enters the next pass as C operand
  int tot[NB][4];
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) { tot[nb][0] = 0; tot[nb][1] = 0; tot[nb][2] = 0; tot[nb][3] = 0; }
  // Wq = d2 2^15 + d1 2^8 + d0 (d0, d2: 8 bits, d1: 7 bits).  The top digit goes first and is shifted once; the middle
  // digit needs no shift of the accumulators: it is stored as 2 d1 and multiplied by a state operand scaled to 128.
#pragma unroll
  for (int slice = 2; slice >= 0; --slice) {
    if (slice == 1) {
#pragma unroll
      for (int nb = 0; nb < NB; ++nb)
#pragma unroll
        for (int e = 0; e < 4; ++e) tot[nb][e] <<= 15;
    }
#pragma unroll
    for (int kk = 0; kk < KK; ++kk) {
      const uint4* row = reinterpret_cast<const uint4*>(Bimg) + ((size_t)(slice * KK + kk) * OB) * 32 + lane;
      const int sh = slice == 1 ? 7 : 0;  // state bits as 0 / 128 (u8) for the middle digit, 0 / 1 otherwise
      const uint32_t a0 = afr[kk][0] << sh, a1 = afr[kk][1] << sh, a2 = afr[kk][2] << sh, a3 = afr[kk][3] << sh;
#pragma unroll
      for (int np = 0; np < OB; ++np) {  // one 16-byte load feeds two n-blocks
        const uint4 b = row[np * 32];
        imma16832(tot[2 * np], a0, a1, a2, a3, b.x, b.y);
        imma16832(tot[2 * np + 1], a0, a1, a2, a3, b.z, b.w);
      }
    }
  }

  // Bernoulli decisions; chain row r of this lane owns units 16 gq + 4 t + w, i.e. n-block 2 gq + w/2, column 2t + w%2.
  // With q = 1 + e^{-x}:  u < 1/q  <=>  u q < 1.  u lies in the cell [uf, uf + 2^-16), uf = m - 1.
  // The two tests are sign tests of one FMA each (its sign is exact); the sign bits are shifted straight into
  // the result masks, so units are visited from the highest index down.
  // A verdict is a SET sign bit, so a NaN (q = inf when x~ < -88: inf - inf) or a margin >= 1 reads as "undecided".
  const float c1 = (0x1.0p-16f + margin) - 1.0f;  // s1 = (m + c1) q - 1 < 0            =>  certainly u < p
  const float d01 = 0x1.0p-16f + 2.0f * margin;   // s0 = d01 q - s1 = 1 - (u_f - margin) q < 0  =>  certainly u >= p
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    uint32_t acc1 = 0u, acc0 = 0u;  // bit 4 gq + w: "certainly 1" / "certainly 0"
    float xs[DBG ? 4 * OB : 1];
#pragma unroll
    for (int pq = (OB + 1) / 2 - 1; pq >= 0; --pq) {
      uint32_t rnd[4];
      philox4x32_10(cid[r][0], cid[r][1], sweep, stream_ctr_i(0u, q0 + 4 * pq + t), key0, key1, rnd);
#pragma unroll
      for (int half = 1; half >= 0; --half) {
        const int gq = 2 * pq + half;
        if (gq < OB) {
#pragma unroll
          for (int w = 3; w >= 0; --w) {
            const int nb = 2 * gq + (w >> 1);
            const float y = fmaf((float)tot[nb][2 * r + (w & 1)], k1, bias[16 * gq + 4 * t + w]);  // -log2(e) x~
            if (DBG) xs[4 * gq + w] = y * -0.6931471805599453f;
            const float q = 1.0f + ex2_approx(y);
            // m = 1 + u_f: the 16-bit cell of u in the top mantissa bits, (0x7F0000 | h16) << 7
            const float m = __uint_as_float(__byte_perm(rnd[w], 0x7Fu, half ? 0x5410u : 0x5432u) << 7);
            const float s1 = fmaf(m + c1, q, -1.0f);
            const float s0 = fmaf(d01, q, -s1);  // re-rounding s1 costs 2^-23 in p, inside the 2^-19 allowance
            acc1 = __funnelshift_l(__float_as_uint(s1), acc1, 1);
            acc0 = __funnelshift_l(__float_as_uint(s0), acc0, 1);
          }
        }
      }
    }
    uint32_t nibw = acc1 & valid_mask;
    const uint32_t slowmask = ~(acc0 | acc1) & valid_mask;  // the two verdicts exclude each other
    if (__any_sync(0xffffffffu, slowmask != 0u))
      nibw = fixup_slow_imma(args, hidden_out, min[r][0], KK > 1 ? min[r][KK - 1] : 0u, nibw, slowmask, cid[r][0], cid[r][1], sweep,
                             q0, lane, DBG ? xs : nullptr, dx);
#pragma unroll
    for (int w = 0; w < (OB + 1) / 2; ++w) mout[r][w] = 0u;
#pragma unroll
    for (int gq = 0; gq < OB; ++gq) mout[r][gq / 2] |= ((nibw >> (4 * gq)) & 0xFu) << (16 * (gq % 2) + 4 * t);
#pragma unroll
    for (int w = 0; w < (OB + 1) / 2; ++w) {
      mout[r][w] |= __shfl_xor_sync(0xffffffffu, mout[r][w], 1);
      mout[r][w] |= __shfl_xor_sync(0xffffffffu, mout[r][w], 2);
    }
  }
}

template <int HB, int VB, bool DBG>
__global__ void __launch_bounds__(GI_THREADS, QCB_GI_MIN_CTAS) gibbs_imma_kernel(const GibbsImmaArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int Hpad = 16 * HB, Npad = 16 * VB;
  constexpr int KKv = (VB + 1) / 2, NBh = 2 * HB, KKh = (HB + 1) / 2, NBv = 2 * VB;
  constexpr size_t nBh = (size_t)3 * KKv * NBh * 32, nBv = (size_t)3 * KKh * NBv * 32;  // in uint2
  const ImmaHeader* hd = reinterpret_cast<const ImmaHeader*>(smem_raw);
  const uint2* Bh = reinterpret_cast<const uint2*>(smem_raw + sizeof(ImmaHeader));
  const uint2* Bv = Bh + nBh;
  const float* biasH = reinterpret_cast<const float*>(Bv + nBv);
  const float* biasV = biasH + Hpad;
  constexpr size_t image_end = align_up(sizeof(ImmaHeader) + (nBh + nBv) * 8 + (Hpad + Npad) * 4, 16);
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw + image_end);  // [0]: image, [1 + warp]: that warp's state tile
  double* tiles = reinterpret_cast<double*>(smem_raw + align_up(image_end + 8 * (1 + GI_WARPS), 128));

  if (threadIdx.x < 1 + GI_WARPS) mbar_init(mbar + threadIdx.x, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(mbar, a.image_bytes);
    for (uint32_t off = 0; off < a.image_bytes; off += 32768u) {
      const uint32_t n = a.image_bytes - off < 32768u ? a.image_bytes - off : 32768u;
      tma_bulk_g2s(smem_raw + off, a.image + off, n, mbar);
    }
  }
  mbar_wait(mbar, 0);
  const float inv_scale = hd->inv_scale, margin_h = hd->margin_h, margin_v = hd->margin_v, dx_h = hd->dx_h, dx_v = hd->dx_v;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2;
  constexpr int MV = (VB + 1) / 2, MH = (HB + 1) / 2;
  const int nhid = a.p.nhid(), N = a.p.N;
  const uint32_t qv0 = (uint32_t)(4 * ((nhid + 31) / 32));
  const int64_t num_groups = (a.C + GI_CPW - 1) / GI_CPW;
  uint32_t valid_h = 0u, valid_v = 0u;  // this lane's real (unpadded) units, bit 4 gq + w <-> unit 16 gq + 4 t + w
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int unit = 16 * (i >> 2) + 4 * (lane & 3) + (i & 3);
    valid_h |= (unit < nhid ? 1u : 0u) << i;
    valid_v |= (unit < N ? 1u : 0u) << i;
  }

  // A group's 16 chain rows are one contiguous span of v_in: the warp's tile is fetched by one TMA bulk copy,
  // issued a whole group ahead (while the previous group is being swept), so its latency never shows.
  double* tile = tiles + (size_t)warp * (GI_CPW * Npad);
  uint64_t* tbar = mbar + 1 + warp;
  uint32_t tphase = 0u;
  const bool bulk_base_ok = (reinterpret_cast<uintptr_t>(a.v_in) & 15u) == 0u;
  auto tile_doubles = [&](int64_t gi) -> uint32_t {
    const int64_t left = a.C - gi * GI_CPW;
    return (uint32_t)(left < GI_CPW ? left : GI_CPW) * (uint32_t)N;
  };
  auto bulk_fetch = [&](int64_t gi) {  // true when the copy was handed to the TMA unit
    const uint32_t bytes = tile_doubles(gi) * 8u;
    if (!bulk_base_ok || (bytes & 15u) != 0u) return false;
    if (lane == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic reads of the tile vs. the async write
      mbar_expect_tx(tbar, bytes);
      tma_bulk_g2s(tile, a.v_in + gi * GI_CPW * (int64_t)N, bytes, tbar);
    }
    return true;
  };
  const int64_t gi0 = (int64_t)blockIdx.x * GI_WARPS + warp, gstride = (int64_t)gridDim.x * GI_WARPS;
  const int rowbytes = packed_row_bytes(N);
  uint32_t vword[MV];  // real (unpadded) visible units per mask word
#pragma unroll
  for (int w = 0; w < MV; ++w) vword[w] = N >= 32 * (w + 1) ? 0xffffffffu : (N > 32 * w ? (1u << (N - 32 * w)) - 1u : 0u);
  const bool from_rows = a.in_mode == 0;
  bool fetched = from_rows && gi0 < num_groups && bulk_fetch(gi0);

  for (int64_t gi = gi0; gi < num_groups; gi += gstride) {
    const int64_t chain_base = gi * GI_CPW;
    const int nch = (int)(tile_doubles(gi) / (uint32_t)N);
    uint32_t cid[2][2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const uint64_t id = a.chain_id0 + (uint64_t)(chain_base + 8 * r + g);
      cid[r][0] = (uint32_t)id;
      cid[r][1] = (uint32_t)(id >> 32);
    }
    uint32_t mv[2][MV], mh[2][MH];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int w = 0; w < MV; ++w) mv[r][w] = 0u;
    if (from_rows) {
      if (fetched) {
        mbar_wait(tbar, tphase);
        tphase ^= 1u;
      } else {  // unaligned view or odd tile size: plain copy
        for (uint32_t idx = lane; idx < tile_doubles(gi); idx += 32) tile[idx] = a.v_in[chain_base * N + idx];
        __syncwarp();
      }
#pragma unroll
      for (int cc = 0; cc < GI_CPW; ++cc) {
#pragma unroll
        for (int w = 0; w < MV; ++w) {
          const int j = w * 32 + lane;
          unsigned long long bits = 0ull;
          if (cc < nch && j < N) bits = reinterpret_cast<const unsigned long long*>(tile)[cc * N + j];
          const uint32_t bal = __ballot_sync(0xffffffffu, (bits << 1) != 0ull);
          if (g == (cc & 7)) mv[cc >> 3][w] = bal;
        }
      }
      __syncwarp();
      fetched = gi + gstride < num_groups && bulk_fetch(gi + gstride);
    } else if (a.in_mode == 1) {
      // packed rows: unit j = bit 7 - j % 8 of byte j / 8 (numpy.packbits order); the four lanes of a row read the same bytes
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int64_t ch = chain_base + 8 * r + g;
        if (ch < a.C) {
          const uint8_t* row = a.bits_in + ch * rowbytes;
#pragma unroll
          for (int w = 0; w < MV; ++w) {
            uint32_t x = 0u;
#pragma unroll
            for (int bb = 0; bb < 4; ++bb)
              if (4 * w + bb < rowbytes) x |= (uint32_t)row[4 * w + bb] << (8 * bb);
            mv[r][w] = __brev(__byte_perm(x, 0u, 0x0123u)) & vword[w];
          }
        }
      }
    } else {
      // v0 ~ Bernoulli(1/2): unit j = bit j % 32 of word j / 32 of ONE Philox call (stream 5, call 0) at sweep index sweep0
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        uint32_t rnd[4];
        philox4x32_10(cid[r][0], cid[r][1], a.sweep0, stream_ctr_i(PHILOX_STREAM_V0, 0u), a.key0, a.key1, rnd);
#pragma unroll
        for (int w = 0; w < MV; ++w) mv[r][w] = rnd[w] & vword[w];
      }
    }
    if constexpr (HB == VB) {
      // Equal paddings: both conditional updates run through ONE copy of the half-sweep code (direction
      // selected at run time), which keeps the loop body inside the 32 KB instruction cache.
#pragma unroll 1
      for (int hs = 0; hs < 2 * a.k; ++hs) {
        const bool back = (hs & 1) != 0;  // false: h | v, true: v | h
        half_sweep_imma<VB, HB, DBG>(mv, mh, back ? Bv : Bh, back ? biasV : biasH, inv_scale, back ? margin_v : margin_h,
                                     back ? dx_v : dx_h, back ? valid_v : valid_h, cid, a.key0, a.key1, a.sweep0 + (uint32_t)(hs >> 1),
                                     back ? qv0 : 0u, &a, back ? 0 : 1, lane);
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int w = 0; w < MV; ++w) mv[r][w] = mh[r][w];
      }
    } else {
#pragma unroll 1
      for (int s = 0; s < a.k; ++s) {
        const uint32_t sweep = a.sweep0 + (uint32_t)s;
        half_sweep_imma<VB, HB, DBG>(mv, mh, Bh, biasH, inv_scale, margin_h, dx_h, valid_h, cid, a.key0, a.key1, sweep, 0u, &a, 1, lane);
        half_sweep_imma<HB, VB, DBG>(mh, mv, Bv, biasV, inv_scale, margin_v, dx_v, valid_v, cid, a.key0, a.key1, sweep, qv0, &a, 0, lane);
      }
    }
    if (a.v != nullptr) {
#pragma unroll
      for (int cc = 0; cc < GI_CPW; ++cc) {
        const int64_t ch = chain_base + cc;
#pragma unroll
        for (int w = 0; w < MV; ++w) {
          const uint32_t word = __shfl_sync(0xffffffffu, mv[cc >> 3][w], (cc & 7) * 4);
          const int j = w * 32 + lane;
          if (ch < a.C && j < N) a.v[ch * N + j] = ((word >> lane) & 1u) ? 1.0 : 0.0;
        }
      }
    }
    if (a.bits_out != nullptr && (lane & 3) == 0) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int64_t ch = chain_base + 8 * r + g;
        if (ch < a.C) {
          uint8_t* row = a.bits_out + ch * rowbytes;
#pragma unroll
          for (int w = 0; w < MV; ++w) {
            const uint32_t x = __byte_perm(__brev(mv[r][w] & vword[w]), 0u, 0x0123u);
#pragma unroll
            for (int bb = 0; bb < 4; ++bb)
              if (4 * w + bb < rowbytes) row[4 * w + bb] = (uint8_t)(x >> (8 * bb));
          }
        }
      }
    }
  }
}

// image + mbarriers + one 16-chain state tile per warp
static size_t imma_smem_bytes(int HB, int VB) {
  return align_up(imma_image_bytes(HB, VB) + 8 * (1 + GI_WARPS), 128) + (size_t)GI_WARPS * GI_CPW * 16 * VB * sizeof(double);
}

static int pick_blocks_imma(int units) {
  if (units <= 16) return 1;
  if (units <= 32) return 2;
  if (units <= 64) return 4;
  return 0;
}

template <int HB, int VB, bool DBG>
static int launch_gibbs_imma(const GibbsImmaArgs& a, cudaStream_t st) {
  const size_t smem = imma_smem_bytes(HB, VB);
  static int blocks_on[QCB_MAX_DEVICES] = {0};  // 0: this device is not configured yet
  int& blocks_per_sm = blocks_on[device_slot()];
  if (blocks_per_sm == 0) {
    int occ = 0;
    QCB_CUDA(cudaFuncSetAttribute(gibbs_imma_kernel<HB, VB, DBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    QCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gibbs_imma_kernel<HB, VB, DBG>, GI_THREADS, smem));
    if (occ < 1) return fail(QCB_ERR_CUDA, "gibbs (imma) kernel does not fit on an SM");
    blocks_per_sm = occ;
  }
  const int64_t groups = (a.C + GI_CPW - 1) / GI_CPW;
  const int64_t need = (groups + GI_WARPS - 1) / GI_WARPS;
  const int64_t cap = (int64_t)sm_count() * blocks_per_sm;
  const int grid = (int)(need < cap ? need : cap);
  gibbs_imma_kernel<HB, VB, DBG><<<grid, GI_THREADS, smem, st>>>(a);
  QCB_LAUNCH_CHECK();
  return QCB_OK;
}

// chains in one full wave of the persistent grid (2 CTAs per SM at 64 x 64); 0 when the shape is not handled here
int64_t gibbs_imma_wave_chains(int N, int nhid) {
  const int HB = pick_blocks_imma(nhid), VB = pick_blocks_imma(N);
  if (HB == 0 || VB == 0) return 0;
  return (int64_t)sm_count() * QCB_GI_MIN_CTAS * GI_WARPS * GI_CPW;
}

size_t gibbs_imma_workspace_bytes(int N, int nhid) {
  const int HB = pick_blocks_imma(nhid), VB = pick_blocks_imma(N);
  if (HB == 0 || VB == 0) return 0;
  return align_up(imma_image_bytes(HB, VB), 256);
}

// *handled = 1 when this kernel took the call (layers <= 64 units, own Philox draws)
int try_gibbs_imma(const StackedParams& p, const ChainIO& io, int64_t C, int k, uint64_t seed, uint64_t chain_id0,
                   uint64_t sweep0, void* ws, size_t ws_bytes, unsigned long long* dbg, int* handled, cudaStream_t st) {
  *handled = 0;
  const int HB = pick_blocks_imma(p.nhid()), VB = pick_blocks_imma(p.N);
  if (HB == 0 || VB == 0) return QCB_OK;
  const size_t need = gibbs_imma_workspace_bytes(p.N, p.nhid());
  if (ws == nullptr || ws_bytes < need) return QCB_OK;
  unsigned char* img = reinterpret_cast<unsigned char*>(ws);
  gibbs_pack_imma_kernel<<<1, 1024, 0, st>>>(p, HB, VB, img);  // one CTA; the word loop below is what takes the time
  QCB_LAUNCH_CHECK();
  GibbsImmaArgs a;
  a.image = img;
  a.image_bytes = (uint32_t)imma_image_bytes(HB, VB);
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
  a.dbg = dbg;
  int rc = QCB_OK;
#define QCB_GI_CASE(H, V) \
  if (HB == H && VB == V) rc = (dbg != nullptr ? launch_gibbs_imma<H, V, true>(a, st) : launch_gibbs_imma<H, V, false>(a, st)); else
  QCB_GI_CASE(1, 1) QCB_GI_CASE(1, 2) QCB_GI_CASE(1, 4) QCB_GI_CASE(2, 1) QCB_GI_CASE(2, 2) QCB_GI_CASE(2, 4)
  QCB_GI_CASE(4, 1) QCB_GI_CASE(4, 2) QCB_GI_CASE(4, 4)
  return QCB_OK;
#undef QCB_GI_CASE
  if (rc == QCB_OK) *handled = 1;
  return rc;
}

}  // namespace qcb
