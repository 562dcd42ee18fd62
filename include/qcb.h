/*
 * qcb.h — C ABI of libqucumber_b200.so: B200 (sm_100a) kernels for the RBM hot path of
 * PIQuIL/QuCumber v1.3.3 (block Gibbs sampling, effective energies and their gradients,
 * exact enumeration, rotated-basis amplitudes/gradients).
 *
 * The reference has no FFI of its own: the seam is the Python object protocol of
 * qucumber.rbm.BinaryRBM / PurificationRBM as consumed through nn_state.rbm_am / rbm_ph.
 * Each entry point below names the reference code (file:line, relative to the reference
 * checkout) whose arithmetic it replaces.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *  - Every pointer is a CUDA device pointer to contiguous row-major float64 unless stated.
 *  - W is (nh, N) "weights"; b (N) "visible_bias"; c (nh) "hidden_bias"
 *    (rbm/binary_rbm.py:48-72).  PurificationRBM adds U (na, N) "weights_U" and d (na)
 *    "aux_bias" (rbm/purification_rbm.py:89-128).  na == 0 (U, d NULL) means BinaryRBM.
 *  - Flat gradient order is [W row-major, (U row-major,) b, c, (d)] — the
 *    parameters_to_vector order of rbm/binary_rbm.py:120 and rbm/purification_rbm.py:185.
 *  - Visible/hidden states are float64 arrays holding exactly 0.0 or 1.0 (the reference's
 *    domain).  Energy / gradient / probability entry points accept arbitrary real v.
 *  - Kernels are enqueued on `stream` (a cudaStream_t) and return immediately; no hidden
 *    synchronisation, no allocation.  Scratch memory is caller-owned: ask
 *    qcb_workspace_bytes() and pass a device buffer of at least that size.
 *  - Return value: QCB_OK or an error class; qcb_last_error() gives a thread-local message.
 *  - There is no CPU fallback and no other backend: unsupported shapes return
 *    QCB_ERR_UNSUPPORTED.
 *
 * Random numbers (used when `uniforms` is NULL)
 *  Philox4x32-10.  key = (seed low 32, seed high 32);
 *  counter = (chain id low 32, chain id high 32, sweep index, stream << 29 | q).
 *  The hidden-like units form one layer [h; a] of nh + na units whose first call is q0 = 0; the
 *  visible layer starts at q0 = 4 * ceil((nh + na) / 32).  Unit u of a layer uses call
 *  q = q0 + 4 (u / 32) + (u % 16) / 4, output word u % 4 and 16-bit half (u / 16) % 2 of that word
 *  (half 0 = high bits), so one call serves eight units.  Its 53-bit uniform is
 *      u53 = (h16 * 2^37 + lo_a * 2^5 + (lo_b >> 27)) * 2^-53,
 *  h16 from stream 0, lo_a from stream 1 + 2 half, lo_b from stream 2 + 2 half — the same grid
 *  {m * 2^-53} torch's CPU bernoulli uses.  A unit is set when u53 < p.  The kernel decides from
 *  h16 whenever that is already conclusive and evaluates the remaining bits only otherwise; the
 *  outcome is always that of the full comparison.  Chain ids are chain_id0 + row and sweep
 *  indices sweep0 + s, so a run sharded over ranks or split into several calls draws exactly the
 *  same numbers as one call.  (qcb_bernoulli, the single-step sampler, uses slot s -> call s / 4,
 *  word s % 4, u = (hi * 2^21 + (lo >> 11)) * 2^-53 with lo from counter word q | 2^31.)
 *
 * Fed uniforms (bit-exact replay of the reference under a given torch seed)
 *  uniforms is (k, C, nh + na + N): per sweep and chain the hidden block, the auxiliary
 *  block, then the visible block — the order rbm/binary_rbm.py:224-226 and
 *  rbm/purification_rbm.py:332-335 consume torch.bernoulli draws.
 */
#ifndef QCB_H_
#define QCB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QCB_VERSION 110 /* 0.1.1 */

enum {
  QCB_OK = 0,
  QCB_ERR_BAD_ARG = 1,
  QCB_ERR_UNSUPPORTED = 2,
  QCB_ERR_CUDA = 3
};

/* Operation ids for qcb_workspace_bytes */
enum {
  QCB_OP_GIBBS = 1,
  QCB_OP_ENERGY = 2,
  QCB_OP_GRAD_REDUCE = 3,
  QCB_OP_LOGZ = 4,
  QCB_OP_EXACT_STATS = 5,
  QCB_OP_CWF_ROTATED = 6,
  QCB_OP_DM_ROTATED = 7,
  QCB_OP_CDK_NEG = 8,
  QCB_OP_CD_STEP = 9,
  QCB_OP_KL_TERMS = 10,
  QCB_OP_SAMPLE = 11 /* qcb_rbm_sample / qcb_rbm_gibbs_packed; rows = C */
};

int qcb_version(void);
const char* qcb_last_error(void);

/* Scratch bytes an operation needs for the given sizes (rows = C, B, or 0 when unused). */
size_t qcb_workspace_bytes(int op, int N, int nh, int na, int64_t rows);

/* ---- Block Gibbs sampling ------------------------------------------------------------
 * k sweeps of (h | v), (a | v), (v | h, a) on C chains, in place on v (C, N).
 * Replaces BinaryRBM.gibbs_steps / sample_h_given_v / sample_v_given_h
 * (rbm/binary_rbm.py:170-228) and PurificationRBM.gibbs_steps (rbm/purification_rbm.py:261-337).
 * uniforms == NULL: in-kernel Philox draws (see above). */
int qcb_rbm_gibbs(const double* W, const double* U, const double* b, const double* c, const double* d,
                  int N, int nh, int na, double* v, int64_t C, int k,
                  uint64_t seed, uint64_t chain_id0, uint64_t sweep0, const double* uniforms,
                  void* ws, size_t ws_bytes, void* stream);

/* Kernel selection for qcb_rbm_gibbs / qcb_rbm_sample / the packed and host variants when uniforms == NULL.
 * Layers of up to 128 units use an integer-exact tensor-core filter: W v in exact 24-bit fixed point on the int8 tensor
 * cores, FP64 re-evaluation of the units the rigorous error bound cannot decide.  Three implementations:
 *   csrc/gibbs_umma.cu      — tcgen05.mma.kind::i8, accumulators in tensor memory (both layers 33..64 units);
 *   csrc/gibbs_umma_wide.cu — the same for any two layers of up to 128 units (several 128-chain tiles per CTA around one
 *                             weight image);
 *   csrc/gibbs_imma.cu      — mma.sync.m16n8k32 (any layers up to 64 units).
 * mode 0 (default): automatic — the fastest of them for the shape and k, else the FP64 kernel (fed uniforms), else the
 * composed path (layers of more than 128 units); mode 1: never the integer filter (FP64 tensor-core kernel of csrc/gibbs.cu
 * or the composed path); mode 2: the tcgen05 kernels wherever their shapes fit (any k); mode 3: never a tcgen05 kernel;
 * mode 4: the general tcgen05 kernel for every shape up to 128 units; mode 5: the table kernel (csrc/gibbs_table.cu: one
 * thread per chain, float nibble tables of the pre-activations in shared memory, no tensor cores) for every shape up to 16
 * units — the automatic choice for such layers.  All kernels draw the same numbers and make the same
 * decisions. */
int qcb_set_gibbs_mode(int mode);

/* Test hook: runs the integer-filter kernel and reports in stats2 (device, two uint64, caller-zeroed)
 * [0] = 2^20 * max over re-evaluated units of |x~ - x| / dx (must stay below 2^20: the bound holds),
 * [1] = number of re-evaluated units. */
int qcb_debug_gibbs_filter(const double* W, const double* U, const double* b, const double* c, const double* d,
                           int N, int nh, int na, double* v, int64_t C, int k, uint64_t seed,
                           unsigned long long* stats2, void* ws, size_t ws_bytes, void* stream);

/* ---- Conditionals --------------------------------------------------------------------
 * out (rows, n_out) = clamp(sigmoid(x M^T + bias [+ x2 M2^T]), 0, 1).
 * prob_h_given_v: M = W (nh, N), transpose_m = 0;  prob_v_given_h: M = W used as (N, nh),
 * transpose_m = 1 (rbm/binary_rbm.py:129-168; rbm/purification_rbm.py:196-259, where
 * prob_v_given_ha adds the second term x2 = a, M2 = U).  m_rows/m_cols are the stored shape. */
int qcb_linear_sigmoid(const double* M, int m_rows, int m_cols, int transpose_m, const double* bias,
                       const double* x, const double* M2, int m2_rows, const double* x2,
                       int64_t rows, double* out, void* stream);

/* out (rows, n_out) = x M^T + bias without the sigmoid — F.linear as used by
 * PurificationRBM.mixing_term (rbm/purification_rbm.py:340-350) and DensityMatrix.pi
 * (nn_states/density_matrix.py:130-134).  bias may be NULL. */
int qcb_linear(const double* M, int m_rows, int m_cols, int transpose_m, const double* bias,
               const double* x, int64_t rows, double* out, void* stream);

/* out[i] = (u[i] < p[i]) as 0.0/1.0 — torch.bernoulli(p, out=p) (rbm/binary_rbm.py:182,197)
 * with u either fed (uniforms != NULL) or Philox slot i of chain `chain_id0`, sweep `sweep0`. */
int qcb_bernoulli(const double* p, int64_t n, const double* uniforms, uint64_t seed,
                  uint64_t chain_id0, uint64_t sweep0, double* out, void* stream);

/* ---- Effective energy ----------------------------------------------------------------
 * E_out[r] = -(b.v_r + sum_i softplus(W_i.v_r + c_i) [+ sum_l softplus(U_l.v_r + d_l)]),
 * softplus with torch's threshold 20.  Replaces BinaryRBM.effective_energy
 * (rbm/binary_rbm.py:75-96) and PurificationRBM.effective_energy with a=None
 * (rbm/purification_rbm.py:131-158). */
int qcb_rbm_energy(const double* W, const double* U, const double* b, const double* c, const double* d,
                   int N, int nh, int na, const double* v, int64_t rows, double* E_out,
                   void* ws, size_t ws_bytes, void* stream);

/* ---- Energy gradients ----------------------------------------------------------------
 * g (G) = [accumulate ? g : 0] + scale * sum_r grad E(v_r), grad E(v) = -[p_h (x) v, (p_a (x) v,) v, p_h (, p_a)].
 * Replaces effective_energy_gradient(v, reduce=True) (rbm/binary_rbm.py:98-120;
 * rbm/purification_rbm.py:160-185) and the 1/batch scalings of
 * nn_states/neural_state.py:383,445.  v may hold arbitrary reals like the reference's; batches of 0/1 rows take a
 * faster exact-integer evaluation of the pre-activations that is selected on the device (a flag word in ws), so the
 * call never synchronises and can be captured in a CUDA graph either way. */
int qcb_rbm_grad_reduce(const double* W, const double* U, const double* b, const double* c, const double* d,
                        int N, int nh, int na, const double* v, int64_t rows, double scale,
                        double* g, int accumulate, void* ws, size_t ws_bytes, void* stream);

/* Kernel selection for large batches (>= 16384 rows) of 0/1 rows on 33..64-unit layers.  mode 0 (default): automatic —
 * both contractions on tcgen05.mma.kind::i8 with TMEM accumulators (csrc/grad_reduce_umma.cu: theta from two 23-bit
 * fixed-point groups of W, p digitised to 48 bits for the outer products), falling back on the device to the kernels
 * below when a row is not binary or the weights do not fit; mode 3: never tcgen05 (integer theta on mma.sync + FP64
 * tensor-core outer products, csrc/grad_reduce_dmma.cu); mode 1: FP64 tensor cores only.  Modes 1 and 3 also make
 * qcb_cwf_rotated reduce the expanded rows through those kernels instead of the Gray-walk subset sums of
 * csrc/rotated_gray.cu, and make qcb_cwf_rotated / qcb_dm_rotated send small basis groups (at most 256 expanded rows, layers
 * of at most 16 units) through the general kernels instead of the one-launch kernels (same results to rounding; for tests
 * and measurements).
 * Small layers (N <= 12, nh + na <= 16) with >= 16 * 2^N binary rows are COUNTED in mode 0 (csrc/grad_reduce_small.cu: a
 * histogram of the at most 2^N distinct rows — one coalesced read of the batch — then one weighted pass over the states);
 * a non-binary row hands the batch to the streaming kernel on the device.  Modes 1 and 3 always stream. */
int qcb_set_grad_reduce_mode(int mode);

/* g_rows (rows, G): per-row gradients, effective_energy_gradient(v, reduce=False)
 * (rbm/binary_rbm.py:121-126; rbm/purification_rbm.py:186-193). */
int qcb_rbm_grad_batch(const double* W, const double* U, const double* b, const double* c, const double* d,
                       int N, int nh, int na, const double* v, int64_t rows, double* g_rows, void* stream);

/* out (rows, N): out[r][i] = E(v_r) - E(v_r with visible unit i flipped) = log p(v^(i)) - log p(v), for binary rows.
 * The single-spin-flip ratios psi(v^(i)) / psi(v) of the local estimators (observables/pauli.py:50-83,
 * examples/Tutorial4_DataGeneration_CalculateObservables/quantum_ising_chain.py:45-78) are
 * exp(out / 2) for the amplitude RBM and exp(i out / 2) for the phase RBM. */
int qcb_rbm_flip_logratio(const double* W, const double* U, const double* b, const double* c, const double* d,
                          int N, int nh, int na, const double* v, int64_t rows, double* out, void* stream);

/* Fused CD-k negative phase: k Gibbs sweeps on v (C, N) in place, then
 * g += scale * sum_chains grad E(v_k)  (nn_states/neural_state.py:443-445 with scale = -1/C). */
int qcb_rbm_cdk_neg(const double* W, const double* U, const double* b, const double* c, const double* d,
                    int N, int nh, int na, double* v, int64_t C, int k,
                    uint64_t seed, uint64_t chain_id0, uint64_t sweep0, const double* uniforms,
                    double scale, double* g, void* ws, size_t ws_bytes, void* stream);

/* Whole CD-k gradient of one batch in one call (compute_batch_gradients, nn_states/neural_state.py:417-447):
 * g = pos_scale * sum_rows grad E(pos) + neg_scale * sum_chains grad E(v_k), v_k = k Gibbs sweeps from
 * neg (not modified; v_k lives in the workspace).  The reference uses pos_scale = 1/B, neg_scale = -1/C. */
int qcb_rbm_cd_step(const double* W, const double* U, const double* b, const double* c, const double* d,
                    int N, int nh, int na, const double* pos, int64_t B, const double* neg, int64_t C, int k,
                    uint64_t seed, uint64_t chain_id0, uint64_t sweep0, double pos_scale, double neg_scale,
                    double* g, void* ws, size_t ws_bytes, void* stream);

/* The same step with both batches as PACKED rows (see "Compact states": ceil(N / 8) bytes per row, numpy.packbits order) — the
 * form in which fit() keeps large binary data sets: the positive-phase reduction, the Gibbs kernel and the negative-phase
 * reduction all read and write bits (8 bytes per row at N = 64 instead of 512), nothing is expanded to float64 unless the
 * tcgen05 gradient kernel hands a batch to its FP64 stand-in.  Same arithmetic, same Philox draws and the same result as
 * qcb_rbm_cd_step on the unpacked rows.  Shapes: both layers of 33..64 units and at least 16384 rows per phase, k >= 1;
 * anything else returns QCB_ERR_UNSUPPORTED.  Workspace: qcb_workspace_bytes(QCB_OP_CD_STEP, N, nh, na, max(B, C)). */
int qcb_rbm_cd_step_packed(const double* W, const double* U, const double* b, const double* c, const double* d,
                           int N, int nh, int na, const uint8_t* pos_bits, int64_t B, const uint8_t* neg_bits, int64_t C, int k,
                           uint64_t seed, uint64_t chain_id0, uint64_t sweep0, double pos_scale, double neg_scale,
                           double* g, void* ws, size_t ws_bytes, void* stream);

/* A whole CD-k epoch with plain SGD in ONE launch, for a small BinaryRBM (1 <= N, nh <= 64): the batch loop of
 * NeuralStateBase.fit (nn_states/neural_state.py:611-627).  For step i = 0 .. num_batches-1:
 *   g = qcb_rbm_cd_step(pos[i*pos_batch : (i+1)*pos_batch], neg[i*neg_batch : (i+1)*neg_batch], k, seed,
 *                       chain_id0 = 0, sweep0 + i*k, 1/B_i, -1/C_i);   W, b, c -= lr * g   (in place)
 * pos (B_total, N) and neg (C_total, N) hold the epoch's batches back to back (the last one may be short).
 * g_last (optional, nh*N + N + nh) receives the gradient of the last step (what param.grad would hold).
 * lr = 0 leaves W, b, c untouched (never written): with num_batches = 1 the call is the CD-k gradient of one
 * batch in a single launch, for optimizers and callbacks that act per batch.
 * Same Philox draws as the per-step call, hence the same chains; sums are associated differently
 * (agreement to rounding).  Larger layers return QCB_ERR_UNSUPPORTED: use qcb_rbm_cd_step per batch. */
int qcb_rbm_cd_epoch(double* W, double* b, double* c, int N, int nh,
                     const double* pos, int64_t B_total, int pos_batch,
                     const double* neg, int64_t C_total, int neg_batch, int num_batches,
                     int k, double lr, uint64_t seed, uint64_t sweep0, double* g_last, void* stream);

/* Kernel selection for qcb_rbm_cd_epoch (process-wide; for tests and measurements).  0 (default): layers of at most 32
 * units run the warp-specialised kernel (producer warps draw the Philox cells of the coming sweeps into a shared-memory ring,
 * consumer warps run the state-dependent chain from float weight copies in registers, exact FP64 path for what the float
 * filter leaves open); 1: always the single-role kernels of round 1.  Both make the same decisions; up to 16 units they also
 * add the gradients in the same order (identical parameters bit for bit), from 17 units on the sums agree to rounding. */
int qcb_set_cd_epoch_mode(int mode);

/* ---- Exact enumeration ---------------------------------------------------------------
 * Basis states are generated from their index r in [first_state, first_state + count):
 * visible unit j = (r >> (N-1-j)) & 1 — the row order of generate_hilbert_space
 * (nn_states/neural_state.py:156-176) — so no (2^N, N) tensor exists anywhere.
 * out2[0] = m = max_r(-E(r)), out2[1] = sum_r exp(-E(r) - m): log Z over the range is
 * m + log(out2[1]); ranges combine with a max/sum.  Replaces partition()
 * (rbm/binary_rbm.py:230-241; rbm/purification_rbm.py:453-463).  N <= 62. */
int qcb_rbm_logz_range(const double* W, const double* U, const double* b, const double* c, const double* d,
                       int N, int nh, int na, uint64_t first_state, uint64_t count,
                       double* out2, void* ws, size_t ws_bytes, void* stream);

/* out2 = (m, s) with logsumexp_i(-E[i]) = m + log(s) for an explicit energy vector — partition()
 * called with a user-supplied `space` tensor (rbm/binary_rbm.py:240). */
int qcb_neg_logsumexp(const double* E, int64_t n, double* out2, void* ws, size_t ws_bytes, void* stream);

/* One pass over the same index range with log Z known, for PositiveWaveFunction metrics and
 * the exact negative phase.  target_re/target_im (may be NULL) point at the entries of the
 * target state for `first_state` onwards.  out8:
 *   [0] sum_r t_re(r) * psi(r)        [1] sum_r t_im(r) * psi(r)      psi = exp(-E/2 - logZ/2)
 *   [2] sum_r p_t(r) * log clamp(p_t(r))   [3] sum_r p_t(r) * log clamp(p(r))
 *   [4] sum_r p(r)                          [5..7] reserved (0)
 * with p = exp(-E - logZ), p_t = t_re^2 + t_im^2, clamp to [2^-52, 1 - 2^-52]
 * (utils/training_statistics.py:57-59, 131-135, 179-183).  If g_exact != NULL it receives
 * (accumulated) sum_r p(r) * grad E(r), the exact negative phase of
 * nn_states/neural_state.py:406-413. */
int qcb_rbm_exact_stats_range(const double* W, const double* U, const double* b, const double* c, const double* d,
                              int N, int nh, int na, uint64_t first_state, uint64_t count, double logZ,
                              const double* target_re, const double* target_im,
                              double* out8, double* g_exact, void* ws, size_t ws_bytes, void* stream);

/* ---- Rotated-basis amplitude and gradient: ComplexWaveFunction ------------------------
 * For each sample (row of `samples`, B x N) measured in `basis` (a HOST array of N codes indexing
 * `unitaries`; code 0 must be the identity / "Z"), expand over the S non-identity sites (2^S states,
 * expansion bit order of utils/unitaries.py:154-181), with psi = exp(-E_am/2) * exp(-i E_ph/2):
 *   Upsi[b]  = sum_i Ut[i,b] psi(v_ib)                             -> Upsi (2, B) [re; im]
 *   g_am    += sum_b Re[ sum_i Ut psi grad E_am(v_ib) / Upsi[b] ]
 *   g_ph    += sum_b Re[ i sum_i Ut psi grad E_ph(v_ib) / Upsi[b] ]
 * unitaries is (K, 2, 2, 2) = [code][re/im][row][col] (utils/unitaries.py:36-48).
 * Replaces rotate_psi_inner_prod (utils/unitaries.py:189-233) and
 * ComplexWaveFunction.rotated_gradient (nn_states/complex_wavefunction.py:162-211).
 * g_am / g_ph may be NULL (amplitudes only); both are accumulated into, not overwritten. */
int qcb_cwf_rotated(const double* W_am, const double* b_am, const double* c_am,
                    const double* W_ph, const double* b_ph, const double* c_ph, int N, int nh,
                    const double* samples, int64_t B, const uint8_t* basis, const double* unitaries, int K,
                    double* Upsi, double* g_am, double* g_ph, void* ws, size_t ws_bytes, void* stream);

/* ---- Rotated-basis probability and gradient: DensityMatrix ----------------------------
 * Two PurificationRBMs (amplitude, phase) with the same (N, nh, na).  For each sample:
 *   w_ij = Ut_i conj(Ut_j) rho~(v_i, v_j),  P[b] = sum_ij Re w_ij              -> UrhoU (B)
 *   g_x += - sum_b sum_ij Re[ w_ij dlog rho~(v_i, v_j)/dx ] / (P[b] + 1e-8)
 * Replaces rotate_rho_probs (utils/unitaries.py:236-283), DensityMatrix.pi / rho / pi_grad /
 * rotated_gradient (nn_states/density_matrix.py:113-330) and PurificationRBM.gamma /
 * gamma_grad (rbm/purification_rbm.py:352-451). */
int qcb_dm_rotated(const double* W_am, const double* U_am, const double* b_am, const double* c_am, const double* d_am,
                   const double* W_ph, const double* U_ph, const double* b_ph, const double* c_ph, const double* d_ph,
                   int N, int nh, int na, const double* samples, int64_t B, const uint8_t* basis,
                   const double* unitaries, int K, double* UrhoU, double* g_am, double* g_ph,
                   void* ws, size_t ws_bytes, void* stream);

/* A batch grouped by measurement basis in ONE call (the loop over np.unique(bases) of NeuralStateBase.gradient,
 * nn_states/neural_state.py:326-369): group gi = rows offsets[gi] .. offsets[gi + 1] of `samples`, measured in basis
 * bases[gi * N .. gi * N + N) (codes as above; offsets (ngroups + 1) and bases are HOST arrays).  An all-Z group adds the
 * plain gradient of the amplitude RBM (scale 1), the others add their rotated-basis gradients — the same kernels in the same
 * order as one qcb_rbm_grad_reduce / qcb_cwf_rotated / qcb_dm_rotated call per group.  scratch: device doubles for the
 * per-group amplitudes / probabilities (2 x the largest group for cwf, 1 x for dm).  Workspace: that of the single-basis
 * call for the largest group. */
int qcb_cwf_rotated_groups(const double* W_am, const double* b_am, const double* c_am,
                           const double* W_ph, const double* b_ph, const double* c_ph, int N, int nh,
                           const double* samples, const int64_t* offsets, const uint8_t* bases, int ngroups,
                           const double* unitaries, int K, double* scratch, double* g_am, double* g_ph,
                           void* ws, size_t ws_bytes, void* stream);
int qcb_dm_rotated_groups(const double* W_am, const double* U_am, const double* b_am, const double* c_am, const double* d_am,
                          const double* W_ph, const double* U_ph, const double* b_ph, const double* c_ph, const double* d_ph,
                          int N, int nh, int na, const double* samples, const int64_t* offsets, const uint8_t* bases, int ngroups,
                          const double* unitaries, int K, double* scratch, double* g_am, double* g_ph,
                          void* ws, size_t ws_bytes, void* stream);

/* ---- One training epoch on data with measurement bases ------------------------------------------
 * The batch loop of NeuralStateBase.fit (nn_states/neural_state.py:593-623) for a ComplexWaveFunction (na == 0; U_*, d_*
 * NULL) or a DensityMatrix (na >= 1) with the reference's default optimizer (vanilla SGD), in ONE call.  Batch t holds the
 * rows offsets[t * ngroups] .. offsets[(t + 1) * ngroups] of `pos`, grouped by basis: group gi (basis bases[gi * N ..], codes
 * into `unitaries` as for qcb_cwf_rotated) owns rows offsets[t * ngroups + gi] .. offsets[t * ngroups + gi + 1].  offsets
 * (nbatches * ngroups + 1 entries) and bases are HOST arrays.  Per batch: g = [g_am ; g_ph] <- (1 / B_t) * the grouped
 * positive phase (compute_batch_gradients, neural_state.py:417-447; same kernels and order as qcb_*_rotated_groups), then k
 * sweeps from rows t * neg_batch_size .. of `neg` with Philox key `seed`, chains 0 .. neg_batch_size - 1 and sweeps
 * sweep0 + t * k .., g_am -= (1 / neg_batch_size) * sum grad E(v_k) (as qcb_rbm_cdk_neg), then p <- p - lr * g on the ten
 * (six) parameter tensors IN PLACE; lr == 0 leaves them untouched.  On return g holds the last batch's gradient.  Each RBM's
 * part of g is ordered [W ; U ; b ; c ; d].  scratch: 2 x (rows of the largest batch) doubles; vk: neg_batch_size * N doubles;
 * workspace: max of qcb_workspace_bytes(QCB_OP_CWF_ROTATED / QCB_OP_DM_ROTATED, largest group) and (QCB_OP_CDK_NEG,
 * neg_batch_size).
 * desc_host (PINNED host memory) / desc_dev (device memory), desc_bytes >= qcb_fit_epoch_desc_bytes(ngroups, nbatches) each,
 * or NULL: with them, a batch all of whose basis groups are small (at most 256 expanded rows each, layers of at most 16 units)
 * takes ONE launch for its positive phase — the call writes one descriptor per group into desc_host, copies them to desc_dev
 * on `stream`, and every group's CTA pair writes a partial gradient that is added in group order (the sums of the
 * group-by-group path, bit for bit).  desc_host must stay untouched until the work queued by the call has run. */
int qcb_fit_epoch_bases(double* W_am, double* U_am, double* b_am, double* c_am, double* d_am,
                        double* W_ph, double* U_ph, double* b_ph, double* c_ph, double* d_ph, int N, int nh, int na,
                        const double* pos, const int64_t* offsets, const uint8_t* bases, int ngroups, int nbatches,
                        const double* unitaries, int K, const double* neg, int64_t neg_batch_size, int k, double lr,
                        uint64_t seed, uint64_t sweep0, double* scratch, double* g, double* vk,
                        void* ws, size_t ws_bytes, void* desc_host, void* desc_dev, size_t desc_bytes, void* stream);
size_t qcb_fit_epoch_desc_bytes(int ngroups, int nbatches);

/* ---- Full-space rotated metrics --------------------------------------------------------
 * qcb_kron_rotate: x <- (U_0 (x) U_1 (x) ... (x) U_{n-1}) x on the leading index of a planar complex array
 * x[i * batch + t], i in [0, 2^n), in place (re and im planes).  basis is a HOST array of n codes into
 * unitaries (K, 2, 2, 2) = [code][re/im][row][col]; code 0 is the identity.  Site s acts on bit n-1-s of i
 * (row order of generate_hilbert_space).  Replaces _kron_mult / rotate_psi / rotate_rho
 * (utils/unitaries.py:64-150): rho -> U rho U^dagger is two calls (rotate, conjugate-transpose, rotate). */
int qcb_kron_rotate(double* re, double* im, int n, int64_t batch, const uint8_t* basis,
                    const double* unitaries, int K, void* stream);

/* qcb_kl_terms: out2[0] = sum_i p_t log clamp(p_t), out2[1] = sum_i p_t log clamp(p), clamp to
 * [2^-52, 1 - 2^-52] (_single_basis_KL, utils/training_statistics.py:131-135; KL = out2[0] - out2[1]).
 * p_t = t_re^2 + t_im^2, or t_re itself when t_im is NULL; p = (p_re^2 + p_im^2) exp(-logZ), or
 * p_re exp(-logZ) when p_im is NULL.  Workspace: qcb_workspace_bytes(QCB_OP_KL_TERMS, ...). */
int qcb_kl_terms(const double* t_re, const double* t_im, const double* p_re, const double* p_im, int64_t n,
                 double logZ, double* out2, void* ws, size_t ws_bytes, void* stream);

/* ---- Compact states -------------------------------------------------------------------
 * A chain state is one bit per unit; the reference carries it as a float64 (rbm/binary_rbm.py:220: `.to(self.weights)`).
 * "Packed rows" are ceil(N / 8) bytes per chain, unit j = bit 7 - j % 8 of byte j / 8 — exactly
 * numpy.packbits(v.astype(uint8), axis=1) — so a host program packs / unpacks with numpy and the PCIe payload of a
 * sampling call is N / 8 bytes per chain instead of 8 N. */
int qcb_pack_states(const double* v, int64_t rows, int N, uint8_t* bits, void* stream);   /* any non-zero entry packs as 1 */
int qcb_unpack_states(const uint8_t* bits, int64_t rows, int N, double* v, void* stream); /* writes 0.0 / 1.0 */

/* NeuralStateBase.sample(k, num_samples) without an initial_state (nn_states/neural_state.py:107-131): the initial
 * state v0 ~ Bernoulli(1/2) is drawn IN the kernel (no host RNG, no host-to-device copy) — unit j of chain c is bit
 * j % 32 of word j / 32 of the Philox call with counter (chain id, sweep0, stream 5 << 29 | 0) — then k block-Gibbs
 * sweeps exactly as qcb_rbm_gibbs would run them from that v0 (same draws for the same seed / chain ids / sweep0).
 * v_out (C, N) float64 and / or bits_out (C, ceil(N / 8)) packed rows receive v_k; either may be NULL, not both.
 * Workspace: qcb_workspace_bytes(QCB_OP_SAMPLE, N, nh, na, C).  N <= 128. */
int qcb_rbm_sample(const double* W, const double* U, const double* b, const double* c, const double* d, int N, int nh, int na,
                   int64_t C, int k, uint64_t seed, uint64_t chain_id0, uint64_t sweep0, double* v_out, uint8_t* bits_out,
                   void* ws, size_t ws_bytes, void* stream);

/* qcb_rbm_gibbs on packed rows, in place (device memory).  Workspace as for qcb_rbm_sample. */
int qcb_rbm_gibbs_packed(const double* W, const double* U, const double* b, const double* c, const double* d, int N, int nh,
                         int na, uint8_t* bits, int64_t C, int k, uint64_t seed, uint64_t chain_id0, uint64_t sweep0, void* ws,
                         size_t ws_bytes, void* stream);

/* ---- Host-buffer convenience (the end-to-end path bench.py times) ---------------------
 * Same as qcb_rbm_gibbs but v_host / params are HOST pointers (pinned for full speed); the
 * call copies in, runs, copies out on `stream` using the caller's device scratch
 * dev_scratch (>= qcb_gibbs_host_scratch_bytes) and synchronises the stream before returning.  Float64 rows: the
 * reference's own representation, 8 N bytes per chain each way — PCIe-bound; prefer the packed variant below. */
size_t qcb_gibbs_host_scratch_bytes(int N, int nh, int na, int64_t C);
int qcb_rbm_gibbs_host(const double* W_host, const double* U_host, const double* b_host, const double* c_host,
                       const double* d_host, int N, int nh, int na, double* v_host, int64_t C, int k,
                       uint64_t seed, uint64_t chain_id0, uint64_t sweep0,
                       void* dev_scratch, size_t dev_scratch_bytes, void* stream);

/* The same with packed rows in HOST memory (in place): bits_host (C, ceil(N / 8)).  draw_initial = 0: bits_host holds
 * the initial states; draw_initial = 1: v0 ~ Bernoulli(1/2) is drawn in-kernel as in qcb_rbm_sample and bits_host is
 * output only.  Slices of whole grid waves are copied in, sampled and copied out concurrently.  Calls on one device
 * serialise on an internal lock (they share that device's copy streams); different devices run concurrently. */
size_t qcb_gibbs_host_packed_scratch_bytes(int N, int nh, int na, int64_t C);
int qcb_rbm_gibbs_host_packed(const double* W_host, const double* U_host, const double* b_host, const double* c_host,
                              const double* d_host, int N, int nh, int na, uint8_t* bits_host, int64_t C, int k, uint64_t seed,
                              uint64_t chain_id0, uint64_t sweep0, int draw_initial, void* dev_scratch, size_t dev_scratch_bytes,
                              void* stream);

/* Kernel-launch counter (incremented by every kernel this library launches); bench.py reads it
 * to report gpu_launches. */
uint64_t qcb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* QCB_H_ */
