"""CPU oracle for the QuCumber RBM hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this package. The product (``qucumber_b200``) never does; it fails loudly when
its CUDA library is missing.

This is a plain numpy float64 restatement of the reference's algorithm (PIQuIL/QuCumber v1.3.3,
paths relative to the reference checkout). Each function cites the reference lines it follows.
The arithmetic below the reference lives in PyTorch ATen (torch>=1.2, a third-party dependency
that is not part of the reference tree); the two ATen behaviours that matter for parity are
restated here: ``F.softplus`` has a hard threshold at 20 and ``probs_to_logits`` clamps to
[eps, 1-eps] with eps = 2**-52.

Parity pin: ``tests/golden/*.npz`` were produced by importing the reference itself in the build
container (``tests/golden/make_golden.py``); ``tests/test_oracle_golden.py`` checks every
function here against them, so the oracle is pinned to the reference, not to itself.
"""
from __future__ import annotations

import numpy as np

EPS = float(np.finfo(np.float64).eps)  # torch.finfo(torch.double).eps, training_statistics.py:17


# --------------------------------------------------------------------------------------------
# ATen scalar functions the path relies on
# --------------------------------------------------------------------------------------------
def softplus(x):
    """F.softplus(x) with beta=1, threshold=20 (rbm/binary_rbm.py:94): x if x > 20 else log1p(e^x)."""
    x = np.asarray(x, dtype=np.float64)
    return np.where(x > 20.0, x, np.log1p(np.exp(np.minimum(x, 20.0))))


def sigmoid(x):
    """torch.sigmoid: 1/(1+e^-x) (rbm/binary_rbm.py:145,166)."""
    x = np.asarray(x, dtype=np.float64)
    with np.errstate(over="ignore"):
        return 1.0 / (1.0 + np.exp(-x))


def probs_to_logits(p):
    """torch.distributions.utils.probs_to_logits for non-binary probs: log(clamp(p, eps, 1-eps))
    (utils/training_statistics.py:17,101,131-135)."""
    return np.log(np.clip(np.asarray(p, dtype=np.float64), EPS, 1.0 - EPS))


# --------------------------------------------------------------------------------------------
# Hilbert-space enumeration order
# --------------------------------------------------------------------------------------------
def hilbert_rows(n, first=0, count=None):
    """Rows [first, first+count) of generate_hilbert_space(n): row r, column j = (r >> (n-1-j)) & 1,
    column 0 is the most significant bit (nn_states/neural_state.py:173-176)."""
    count = (1 << n) - first if count is None else count
    idx = np.arange(first, first + count, dtype=np.uint64)
    shifts = np.arange(n - 1, -1, -1, dtype=np.uint64)
    return ((idx[:, None] >> shifts[None, :]) & np.uint64(1)).astype(np.float64)


def basis_index(states):
    """_convert_basis_element_to_index (utils/unitaries.py:184-186): MSB-first integer of each row."""
    states = np.asarray(states)
    n = states.shape[-1]
    powers = (1 << np.arange(n - 1, -1, -1)).astype(np.int64)
    return (np.rint(states).astype(np.int64) * powers).sum(-1)


# --------------------------------------------------------------------------------------------
# Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11) and the
# uniform-draw convention of the CUDA kernels (include/qcb.h, "Random numbers")
# --------------------------------------------------------------------------------------------
_PHILOX_M0 = np.uint64(0xD2511F53)
_PHILOX_M1 = np.uint64(0xCD9E8D57)
_PHILOX_W0 = 0x9E3779B9
_PHILOX_W1 = 0xBB67AE85
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32 with 10 rounds. Counter words are uint32 arrays (broadcastable),
    key words are python ints. Returns four uint32 arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(*(np.asarray(c, dtype=np.uint64) & _MASK32 for c in (c0, c1, c2, c3)))
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _PHILOX_M0 * c0
        p1 = _PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + _PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + _PHILOX_W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def layer_uniforms(seed, chain_ids, sweep, q0, count):
    """53-bit uniforms of the `count` units of one layer for every chain in ``chain_ids`` at global
    sweep index ``sweep`` — the convention of qcb_rbm_gibbs when no uniforms are fed (include/qcb.h):

    unit u uses Philox call q = q0 + 4 (u // 32) + (u % 16) // 4, output word w = u % 4 and
    half = (u // 16) % 2 of that word (half 0 = the high 16 bits). Counter = (chain id low 32,
    chain id high 32, sweep, stream << 29 | q), key = the 64-bit seed split low/high, and
        u53 = (h16 * 2^37 + lo_a * 2^5 + (lo_b >> 27)) * 2^-53
    with h16 from stream 0, lo_a from stream 1 + 2 half and lo_b from stream 2 + 2 half."""
    chain_ids = np.asarray(chain_ids, dtype=np.uint64)
    u = np.arange(count, dtype=np.uint64)
    q = (np.uint64(q0) + np.uint64(4) * (u // np.uint64(32)) + (u % np.uint64(16)) // np.uint64(4))[None, :]
    w = (u % np.uint64(4)).astype(np.int64)
    half = ((u // np.uint64(16)) % np.uint64(2))
    c0 = (chain_ids & _MASK32)[:, None]
    c1 = (chain_ids >> np.uint64(32))[:, None]
    k0, k1 = int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF

    def words(stream):
        ctr = (np.asarray(stream, dtype=np.uint64) << np.uint64(29)) | q
        out4 = np.stack(philox4x32_10(c0, c1, np.uint64(sweep), ctr, k0, k1), axis=-1)
        sel = np.broadcast_to(w[None, :, None], out4.shape[:2] + (1,))
        return np.take_along_axis(out4, sel, axis=-1)[..., 0].astype(np.uint64)

    hi = words(np.zeros_like(half)[None, :])
    la = words((np.uint64(1) + np.uint64(2) * half)[None, :])
    lb = words((np.uint64(2) + np.uint64(2) * half)[None, :])
    h16 = (hi >> (np.uint64(16) * (np.uint64(1) - half))[None, :]) & np.uint64(0xFFFF)
    bits = (h16 << np.uint64(37)) | (la << np.uint64(5)) | (lb >> np.uint64(27))
    return bits.astype(np.float64) * (2.0 ** -53)


def philox_initial_states(seed, chain_ids, sweep0, n):
    """v0 ~ Bernoulli(1/2) of qcb_rbm_sample (the in-kernel stand-in for the host draw of NeuralStateBase.sample,
    nn_states/neural_state.py:124-129): unit j of a chain = bit j % 32 of word j // 32 of the Philox call with
    counter (chain id low, chain id high, sweep0, 5 << 29), key = seed. n <= 128."""
    chain_ids = np.asarray(chain_ids, dtype=np.uint64)
    k0, k1 = int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF
    words = np.stack(philox4x32_10(chain_ids & _MASK32, chain_ids >> np.uint64(32), np.uint64(sweep0), np.uint64(5 << 29), k0, k1),
                     axis=-1).astype(np.uint64)                                   # (C, 4)
    j = np.arange(int(n))
    return ((words[:, j // 32] >> (j % 32).astype(np.uint64)[None, :]) & np.uint64(1)).astype(np.float64)


def pack_states(v):
    """Packed rows of include/qcb.h: numpy.packbits order, ceil(N / 8) bytes per row."""
    return np.packbits(np.asarray(v) != 0, axis=-1)


def unpack_states(bits, n):
    return np.unpackbits(np.asarray(bits, dtype=np.uint8), axis=-1)[..., : int(n)].astype(np.float64)


def visible_q0(nhid):
    """First Philox call of the visible layer: the hidden-like layer occupies 4 calls per 32 units."""
    return 4 * ((int(nhid) + 31) // 32)


def philox_uniforms(seed, chain_ids, sweep, slot0, count):
    """Uniforms of qcb_bernoulli (single conditional steps): slot s uses call s // 4, word s % 4,
    u = (hi * 2^21 + (lo >> 11)) * 2^-53 with hi from counter word q and lo from q | 2^31."""
    chain_ids = np.asarray(chain_ids, dtype=np.uint64)
    slots = np.arange(slot0, slot0 + count, dtype=np.uint64)
    q = (slots >> np.uint64(2))[None, :]
    w = (slots & np.uint64(3)).astype(np.int64)
    c0 = (chain_ids & _MASK32)[:, None]
    c1 = (chain_ids >> np.uint64(32))[:, None]
    k0, k1 = int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF
    hi4 = np.stack(philox4x32_10(c0, c1, np.uint64(sweep), q, k0, k1), axis=-1)
    lo4 = np.stack(philox4x32_10(c0, c1, np.uint64(sweep), q | np.uint64(1 << 31), k0, k1), axis=-1)
    sel = w[None, :, None]
    hi = np.take_along_axis(hi4, np.broadcast_to(sel, hi4.shape[:2] + (1,)), axis=-1)[..., 0].astype(np.uint64)
    lo = np.take_along_axis(lo4, np.broadcast_to(sel, lo4.shape[:2] + (1,)), axis=-1)[..., 0].astype(np.uint64)
    bits = (hi << np.uint64(21)) | (lo >> np.uint64(11))
    return bits.astype(np.float64) * (2.0 ** -53)


# --------------------------------------------------------------------------------------------
# BinaryRBM (rbm/binary_rbm.py)
# --------------------------------------------------------------------------------------------
class BinaryRBMOracle:
    """Restates qucumber.rbm.BinaryRBM: weights (nh, N), visible_bias (N), hidden_bias (nh)
    (rbm/binary_rbm.py:48-72)."""

    def __init__(self, weights, visible_bias, hidden_bias):
        self.W = np.ascontiguousarray(weights, dtype=np.float64)
        self.b = np.ascontiguousarray(visible_bias, dtype=np.float64)
        self.c = np.ascontiguousarray(hidden_bias, dtype=np.float64)
        self.nh, self.n = self.W.shape
        self.num_pars = self.nh * self.n + self.n + self.nh

    # rbm/binary_rbm.py:75-96
    def effective_energy(self, v):
        v = np.asarray(v, dtype=np.float64)
        return -(v @ self.b + softplus(v @ self.W.T + self.c).sum(-1))

    # rbm/binary_rbm.py:150-168 (the clamp is the identity for finite input)
    def prob_h_given_v(self, v):
        return np.clip(sigmoid(np.asarray(v, dtype=np.float64) @ self.W.T + self.c), 0.0, 1.0)

    # rbm/binary_rbm.py:129-147
    def prob_v_given_h(self, h):
        return np.clip(sigmoid(np.asarray(h, dtype=np.float64) @ self.W + self.b), 0.0, 1.0)

    # rbm/binary_rbm.py:98-126; flat order [W row-major, b, c] = parameters_to_vector order
    def effective_energy_gradient(self, v, reduce=True):
        v = np.asarray(v, dtype=np.float64)
        v = v[None, :] if v.ndim < 2 else v
        p = self.prob_h_given_v(v)
        if reduce:
            v2, p2 = v.reshape(-1, self.n), p.reshape(-1, self.nh)
            if v.ndim > 2:  # reference semantics for >2-D input are transpose(0,-1) based; only 2-D is used
                raise ValueError("reduce=True oracle supports 2-D input only")
            return np.concatenate([-(p2.T @ v2).reshape(-1), -v2.sum(0), -p2.sum(0)])
        wg = -(p[..., :, None] * v[..., None, :]).reshape(*v.shape[:-1], -1)
        return np.concatenate([wg, -v, -p], axis=-1)

    # rbm/binary_rbm.py:200-228 with the Bernoulli draws replaced by the thresholding they are
    # equivalent to: per sweep h = [u_h < p(h|v)], then v = [u_v < p(v|h)];
    # uniforms[s, chain, :nh] is the hidden block, uniforms[s, chain, nh:] the visible block.
    def gibbs_steps_fed(self, k, initial_state, uniforms):
        v = np.array(initial_state, dtype=np.float64, copy=True)
        u = np.asarray(uniforms, dtype=np.float64)
        for s in range(k):
            h = (u[s, ..., : self.nh] < self.prob_h_given_v(v)).astype(np.float64)
            v = (u[s, ..., self.nh:] < self.prob_v_given_h(h)).astype(np.float64)
        return v

    def gibbs_steps_philox(self, k, initial_state, seed, chain_id0=0, sweep0=0):
        """Same sweeps with the kernel's own counter-based draws (layer_uniforms)."""
        v = np.array(initial_state, dtype=np.float64, copy=True).reshape(-1, self.n)
        ids = np.arange(chain_id0, chain_id0 + v.shape[0], dtype=np.uint64)
        qv = visible_q0(self.nh)
        for s in range(k):
            uh = layer_uniforms(seed, ids, sweep0 + s, 0, self.nh)
            h = (uh < self.prob_h_given_v(v)).astype(np.float64)
            uv = layer_uniforms(seed, ids, sweep0 + s, qv, self.n)
            v = (uv < self.prob_v_given_h(h)).astype(np.float64)
        return v.reshape(np.shape(initial_state))

    def sample_chains_philox(self, k, num_chains, seed, chain_id0=0, sweep0=0):
        """qcb_rbm_sample: in-kernel Bernoulli(1/2) start, then the sweeps of gibbs_steps_philox."""
        ids = np.arange(chain_id0, chain_id0 + num_chains, dtype=np.uint64)
        return self.gibbs_steps_philox(k, philox_initial_states(seed, ids, sweep0, self.n), seed, chain_id0, sweep0)

    # rbm/binary_rbm.py:230-241 evaluated range-wise: returns (m, s) with
    # logsumexp_{r in range}(-E(row r)) = m + log(s)
    def logz_range(self, first=0, count=None, chunk=1 << 16):
        total = 1 << self.n
        count = total - first if count is None else count
        m, s = -np.inf, 0.0
        pos = first
        while pos < first + count:
            c = min(chunk, first + count - pos)
            x = -self.effective_energy(hilbert_rows(self.n, pos, c))
            mx = float(x.max())
            if mx > m:
                s = s * np.exp(m - mx) if np.isfinite(m) else 0.0
                m = mx
            s += float(np.exp(x - m).sum())
            pos += c
        return m, s

    def partition(self, space=None):
        if space is not None:
            x = -self.effective_energy(space)
            m = x.max()
            return float(np.exp(m + np.log(np.exp(x - m).sum())))
        m, s = self.logz_range()
        return float(np.exp(m + np.log(s)))

    def log_partition(self):
        m, s = self.logz_range()
        return m + np.log(s)

    def log_partition_hidden_side(self):
        """Independent identity used as a size-independent check: tracing the *visible* units
        analytically gives Z = sum_h exp(c.h) prod_j (1 + exp(b_j + sum_i h_i W_ij)); note it
        uses log1p(e^x) without the softplus threshold, so it agrees with `partition` only to
        ~2e-9 * (number of pre-activations above 20)."""
        hs = hilbert_rows(self.nh)
        x = hs @ self.c + np.logaddexp(0.0, hs @ self.W + self.b).sum(-1)
        m = x.max()
        return float(m + np.log(np.exp(x - m).sum()))


# --------------------------------------------------------------------------------------------
# PurificationRBM (rbm/purification_rbm.py)
# --------------------------------------------------------------------------------------------
class PurificationRBMOracle:
    """Restates qucumber.rbm.PurificationRBM: weights_W (nh,N), weights_U (na,N), visible_bias,
    hidden_bias, aux_bias (rbm/purification_rbm.py:89-128); flat order [W, U, b, c, d]."""

    def __init__(self, weights_W, weights_U, visible_bias, hidden_bias, aux_bias):
        self.W = np.ascontiguousarray(weights_W, dtype=np.float64)
        self.U = np.ascontiguousarray(weights_U, dtype=np.float64)
        self.b = np.ascontiguousarray(visible_bias, dtype=np.float64)
        self.c = np.ascontiguousarray(hidden_bias, dtype=np.float64)
        self.d = np.ascontiguousarray(aux_bias, dtype=np.float64)
        self.nh, self.n = self.W.shape
        self.na = self.U.shape[0]
        self.num_pars = self.nh * self.n + self.na * self.n + self.n + self.nh + self.na

    # rbm/purification_rbm.py:131-158
    def effective_energy(self, v, a=None):
        v = np.asarray(v, dtype=np.float64)
        vis = v @ self.b + softplus(v @ self.W.T + self.c).sum(-1)
        if a is None:
            return -(vis + softplus(v @ self.U.T + self.d).sum(-1))
        a = np.asarray(a, dtype=np.float64)
        return -(vis + a @ self.d + np.einsum("...v,av,...a->...", v, self.U, a))

    def prob_h_given_v(self, v):  # :196-215
        return np.clip(sigmoid(np.asarray(v, dtype=np.float64) @ self.W.T + self.c), 0.0, 1.0)

    def prob_a_given_v(self, v):  # :217-236
        return np.clip(sigmoid(np.asarray(v, dtype=np.float64) @ self.U.T + self.d), 0.0, 1.0)

    def prob_v_given_ha(self, h, a):  # :238-259
        h = np.asarray(h, dtype=np.float64)
        a = np.asarray(a, dtype=np.float64)
        return np.clip(sigmoid((h @ self.W + self.b) + a @ self.U), 0.0, 1.0)

    def effective_energy_gradient(self, v, reduce=True):  # :160-193
        v = np.asarray(v, dtype=np.float64)
        v = v[None, :] if v.ndim < 2 else v
        ph, pa = self.prob_h_given_v(v), self.prob_a_given_v(v)
        if reduce:
            if v.ndim > 2:
                raise ValueError("reduce=True oracle supports 2-D input only")
            return np.concatenate([-(ph.T @ v).reshape(-1), -(pa.T @ v).reshape(-1), -v.sum(0), -ph.sum(0), -pa.sum(0)])
        wg = -(ph[..., :, None] * v[..., None, :]).reshape(*v.shape[:-1], -1)
        ug = -(pa[..., :, None] * v[..., None, :]).reshape(*v.shape[:-1], -1)
        return np.concatenate([wg, ug, -v, -ph, -pa], axis=-1)

    # :309-337, draw order h, a, v; uniforms[s, chain] = [h block | a block | v block]
    def gibbs_steps_fed(self, k, initial_state, uniforms):
        v = np.array(initial_state, dtype=np.float64, copy=True)
        u = np.asarray(uniforms, dtype=np.float64)
        for s in range(k):
            h = (u[s, ..., : self.nh] < self.prob_h_given_v(v)).astype(np.float64)
            a = (u[s, ..., self.nh: self.nh + self.na] < self.prob_a_given_v(v)).astype(np.float64)
            v = (u[s, ..., self.nh + self.na:] < self.prob_v_given_ha(h, a)).astype(np.float64)
        return v

    def gibbs_steps_philox(self, k, initial_state, seed, chain_id0=0, sweep0=0):
        """Hidden and auxiliary units form one stacked layer [h; a] of nh + na units (layer_uniforms)."""
        v = np.array(initial_state, dtype=np.float64, copy=True).reshape(-1, self.n)
        ids = np.arange(chain_id0, chain_id0 + v.shape[0], dtype=np.uint64)
        qv = visible_q0(self.nh + self.na)
        for s in range(k):
            uha = layer_uniforms(seed, ids, sweep0 + s, 0, self.nh + self.na)
            h = (uha[:, : self.nh] < self.prob_h_given_v(v)).astype(np.float64)
            a = (uha[:, self.nh:] < self.prob_a_given_v(v)).astype(np.float64)
            v = (layer_uniforms(seed, ids, sweep0 + s, qv, self.n) < self.prob_v_given_ha(h, a)).astype(np.float64)
        return v.reshape(np.shape(initial_state))

    def sample_chains_philox(self, k, num_chains, seed, chain_id0=0, sweep0=0):
        ids = np.arange(chain_id0, chain_id0 + num_chains, dtype=np.uint64)
        return self.gibbs_steps_philox(k, philox_initial_states(seed, ids, sweep0, self.n), seed, chain_id0, sweep0)

    def logz_range(self, first=0, count=None, chunk=1 << 16):  # :453-463
        total = 1 << self.n
        count = total - first if count is None else count
        m, s = -np.inf, 0.0
        pos = first
        while pos < first + count:
            c = min(chunk, first + count - pos)
            x = -self.effective_energy(hilbert_rows(self.n, pos, c))
            mx = float(x.max())
            if mx > m:
                s = s * np.exp(m - mx) if np.isfinite(m) else 0.0
                m = mx
            s += float(np.exp(x - m).sum())
            pos += c
        return m, s

    def partition(self, space=None):
        if space is not None:
            x = -self.effective_energy(space)
            m = x.max()
            return float(np.exp(m + np.log(np.exp(x - m).sum())))
        m, s = self.logz_range()
        return float(np.exp(m + np.log(s)))

    def mixing_term(self, v):  # :340-350
        return np.asarray(v, dtype=np.float64) @ (0.5 * self.U).T + self.d

    def _f(self, v):
        """b.v + sum softplus(Wv + c), the quantity Gamma is built from (:375-396)."""
        v = np.asarray(v, dtype=np.float64)
        return v @ self.b + softplus(v @ self.W.T + self.c).sum(-1)

    def gamma(self, v, vp, eta=1, expand=True):  # :352-396
        sign = float(np.sign(eta))
        t1, t2 = self._f(v), self._f(vp)
        if np.ndim(v) < 2 and np.ndim(vp) < 2:
            return 0.5 * (t1 + sign * t2)
        if expand:
            return 0.5 * (t1[:, None, ...] + sign * t2[None, :, ...])
        return 0.5 * (t1 + sign * t2)

    def gamma_grad(self, v, vp, eta=1, expand=False):  # :398-451 (real part; imaginary part is zero)
        sign = float(np.sign(eta))
        v = np.asarray(v, dtype=np.float64)
        vp = np.asarray(vp, dtype=np.float64)
        v = v[None] if v.ndim < 2 else v
        vp = vp[None] if vp.ndim < 2 else vp
        ph, php = self.prob_h_given_v(v), self.prob_h_given_v(vp)
        wg = ph[..., :, None] * v[..., None, :]
        wgp = php[..., :, None] * vp[..., None, :]
        if expand:
            W_grad = 0.5 * (wg[:, None] + sign * wgp[None, :])
            vb = 0.5 * (v[:, None] + sign * vp[None, :])
            hb = 0.5 * (ph[:, None] + sign * php[None, :])
            bs = (v.shape[0], vp.shape[0], *v.shape[1:-1])
        else:
            W_grad = 0.5 * (wg + sign * wgp)
            vb = 0.5 * (v + sign * vp)
            hb = 0.5 * (ph + sign * php)
            bs = v.shape[:-1]
        return np.concatenate(
            [W_grad.reshape(*bs, -1), np.zeros((*bs, self.na * self.n)), vb, hb, np.zeros((*bs, self.na))], axis=-1
        )


# --------------------------------------------------------------------------------------------
# Wavefunctions and pure-state metrics
# --------------------------------------------------------------------------------------------
def psi_positive(rbm_am, v):
    """PositiveWaveFunction.psi (nn_states/positive_wavefunction.py:95-124, wavefunction.py:33-46):
    amplitude = sqrt(exp(-E)), phase 0. Returns complex128."""
    return np.sqrt(np.exp(-rbm_am.effective_energy(v))).astype(np.complex128)


def psi_complex(rbm_am, rbm_ph, v):
    """ComplexWaveFunction.psi (nn_states/wavefunction.py:62-81, complex_wavefunction.py:129-142):
    amplitude = sqrt(exp(-E_am)), phase = -E_ph / 2."""
    amp = np.sqrt(np.exp(-rbm_am.effective_energy(v)))
    ph = -0.5 * rbm_ph.effective_energy(v)
    return amp * np.cos(ph) + 1j * (amp * np.sin(ph))


def fidelity_pure(psi_unnorm, Z, target):
    """utils/training_statistics.py:57-59: |<target|psi/sqrt(Z)>|^2 (complex128 vectors)."""
    return float(np.abs(np.vdot(target, psi_unnorm / np.sqrt(Z))) ** 2)


def single_basis_KL(target_probs, nn_probs):
    """utils/training_statistics.py:131-135."""
    return float(np.sum(target_probs * probs_to_logits(target_probs)) - np.sum(target_probs * probs_to_logits(nn_probs)))


def KL_positive(rbm_am, target):
    """ts.KL with bases=None (utils/training_statistics.py:164,179-183)."""
    space = hilbert_rows(rbm_am.n)
    Z = rbm_am.partition(space)
    return single_basis_KL(np.abs(target) ** 2, np.exp(-rbm_am.effective_energy(space)) / Z)


def NLL_positive(rbm_am, samples):
    """ts.NLL with sample_bases=None (utils/training_statistics.py:97-102)."""
    Z = rbm_am.partition(hilbert_rows(rbm_am.n))
    return float(-np.mean(probs_to_logits(np.exp(-rbm_am.effective_energy(samples)) / Z)))


def exact_gradient_positive(rbm_am, samples):
    """compute_exact_gradients for a PositiveWaveFunction (nn_states/neural_state.py:386-415)."""
    samples = np.asarray(samples, dtype=np.float64)
    grad = rbm_am.effective_energy_gradient(samples) / float(samples.shape[0])
    space = hilbert_rows(rbm_am.n)
    probs = np.exp(-rbm_am.effective_energy(space))
    probs = probs / probs.sum()
    return grad - rbm_am.effective_energy_gradient(space, reduce=False).T @ probs


# --------------------------------------------------------------------------------------------
# Basis rotations (utils/unitaries.py)
# --------------------------------------------------------------------------------------------
def default_unitaries():
    """create_dict defaults as complex 2x2 matrices (utils/unitaries.py:36-48)."""
    r = 1.0 / np.sqrt(2.0)
    return {
        "X": np.array([[1, 1], [1, -1]], dtype=np.complex128) * r,
        "Y": np.array([[1, -1j], [1, 1j]], dtype=np.complex128) * r,
        "Z": np.array([[1, 0], [0, 1]], dtype=np.complex128),
    }


def rotate_basis_state(basis, states, unitaries=None):
    """_rotate_basis_state (utils/unitaries.py:154-181): expand each sample over its S non-Z
    sites. Returns Ut (2^S, B) complex and v (2^S, B, N); the expansion index i sets rotated
    site number t to bit (i >> (S-1-t)) & 1, and Ut[i,b] = prod_t U_t[sigma_b[site_t], v_i[site_t]]."""
    unitaries = unitaries if unitaries is not None else default_unitaries()
    states = np.asarray(states, dtype=np.float64)
    basis = np.array(list(basis))
    sites = np.where(basis != "Z")[0]
    S = sites.size
    if S == 0:
        return np.ones((1,) + states.shape[:-1], dtype=np.complex128), states[None].copy()
    v = np.repeat(states[None], 1 << S, axis=0)
    sub = hilbert_rows(S)  # (2^S, S)
    v[..., sites] = sub[:, None, :]
    sig = np.rint(states[..., sites]).astype(np.int64)  # (B, S)
    vp = sub.astype(np.int64)  # (2^S, S)
    Ut = np.ones((1 << S,) + states.shape[:-1], dtype=np.complex128)
    for t, site in enumerate(sites):
        U = np.asarray(unitaries[basis[site]], dtype=np.complex128)
        Ut = Ut * U[sig[None, :, t], vp[:, None, t]]
    return Ut, v


def rotate_psi_inner_prod(psi_fn, basis, states, unitaries=None, include_extras=False):
    """utils/unitaries.py:189-233 with psi evaluated by ``psi_fn(v) -> complex``."""
    Ut, v = rotate_basis_state(basis, states, unitaries)
    Upsi_v = Ut * psi_fn(v)
    Upsi = Upsi_v.sum(0)
    return (Upsi, Upsi_v, v) if include_extras else Upsi


def kron_rotate(basis, vec, unitaries=None):
    """rotate_psi via the full Kronecker product (utils/unitaries.py:64-115), site 0 = MSB."""
    unitaries = unitaries if unitaries is not None else default_unitaries()
    out = np.asarray(vec, dtype=np.complex128)
    n = len(basis)
    out = out.reshape((2,) * n)
    for site, b in enumerate(basis):
        out = np.moveaxis(np.tensordot(np.asarray(unitaries[b], dtype=np.complex128), out, axes=([1], [site])), 0, site)
    return out.reshape(-1)


def complex_rotated_gradient(rbm_am, rbm_ph, basis, samples, unitaries=None):
    """ComplexWaveFunction.rotated_gradient (nn_states/complex_wavefunction.py:162-211):
    g_am = sum_b Re[ sum_i T_ib grad E_am(v_ib) / sum_i T_ib ],
    g_ph = sum_b Re[ i sum_i T_ib grad E_ph(v_ib) / sum_i T_ib ],  T = Ut * psi(v)."""
    Upsi, Upsi_v, v = rotate_psi_inner_prod(lambda x: psi_complex(rbm_am, rbm_ph, x), basis, samples, unitaries, True)
    g_am = rbm_am.effective_energy_gradient(v, reduce=False)
    g_ph = 1j * rbm_ph.effective_energy_gradient(v, reduce=False)
    rot_am = np.einsum("ib,ibg->bg", Upsi_v, g_am)
    rot_ph = np.einsum("ib,ibg->bg", Upsi_v, g_ph)
    inv = 1.0 / Upsi
    return np.real(np.einsum("b,bg->g", inv, rot_am)), np.real(np.einsum("b,bg->g", inv, rot_ph))


def group_by_basis(bases):
    """np.unique(bases, axis=0, return_inverse=True) grouping of nn_states/neural_state.py:351."""
    bases = np.asarray(bases)
    uniq, inv = np.unique(bases, axis=0, return_inverse=True)
    return uniq, np.asarray(inv).reshape(-1)


def complex_gradient(rbm_am, rbm_ph, samples, bases, unitaries=None):
    """NeuralStateBase.gradient for a ComplexWaveFunction (nn_states/neural_state.py:326-369)."""
    samples = np.asarray(samples, dtype=np.float64)
    g = [np.zeros(rbm_am.num_pars), np.zeros(rbm_ph.num_pars)]
    uniq, inv = group_by_basis(bases)
    for i in range(uniq.shape[0]):
        sub = samples[inv == i]
        if np.any(uniq[i] != "Z"):
            ga, gp = complex_rotated_gradient(rbm_am, rbm_ph, uniq[i], sub, unitaries)
            g[0] += ga
            g[1] += gp
        else:
            g[0] += rbm_am.effective_energy_gradient(sub)
    return g


# --------------------------------------------------------------------------------------------
# DensityMatrix (nn_states/density_matrix.py)
# --------------------------------------------------------------------------------------------
def dm_pi(rbm_am, rbm_ph, v, vp, expand=True):
    """DensityMatrix.pi (nn_states/density_matrix.py:113-157); phase RBM aux bias is NOT used."""
    v = np.asarray(v, dtype=np.float64)
    vp = np.asarray(vp, dtype=np.float64)
    m_am, mp_am = v @ rbm_am.U.T + rbm_am.d, vp @ rbm_am.U.T + rbm_am.d
    m_ph, mp_ph = v @ rbm_ph.U.T, vp @ rbm_ph.U.T
    if expand and v.ndim >= 2:
        m_am, m_ph = m_am[:, None], m_ph[:, None]
    if expand and vp.ndim >= 2:
        mp_am, mp_ph = mp_am[None, :], mp_ph[None, :]
    x = (m_am + mp_am) / 2
    ph = (m_ph - mp_ph) / 2
    real = np.log(np.sqrt(1 + 2 * np.exp(x) * np.cos(ph) + np.exp(2 * x))).sum(-1)
    imag = np.arctan2(np.exp(x) * np.sin(ph), 1 + np.exp(x) * np.cos(ph)).sum(-1)
    return real + 1j * imag


def dm_rho(rbm_am, rbm_ph, v, vp=None, expand=True):
    """DensityMatrix.rho (nn_states/density_matrix.py:245-273), complex128."""
    if expand is False and vp is None:
        return np.exp(-rbm_am.effective_energy(v)).astype(np.complex128)
    vp = v if vp is None else vp
    pi_ = dm_pi(rbm_am, rbm_ph, v, vp, expand)
    amp = np.exp(rbm_am.gamma(v, vp, +1, expand) + pi_.real)
    ph = rbm_ph.gamma(v, vp, -1, expand) + pi_.imag
    return amp * np.cos(ph) + 1j * (amp * np.sin(ph))


def complex_sigmoid(x, y):
    """cplx.sigmoid (utils/cplx.py:319-334): 1 / (1 + exp(-(x + iy)))."""
    z = np.asarray(x, dtype=np.float64) + 1j * np.asarray(y, dtype=np.float64)
    return 1.0 / (1.0 + np.exp(-z))


def dm_pi_grad(rbm_am, rbm_ph, v, vp, phase=False):
    """DensityMatrix.pi_grad with expand=True (nn_states/density_matrix.py:159-243), complex."""
    v = np.asarray(v, dtype=np.float64)
    vp = np.asarray(vp, dtype=np.float64)
    arg_r = 0.5 * ((v @ rbm_am.U.T + rbm_am.d)[:, None] + (vp @ rbm_am.U.T + rbm_am.d)[None, :])
    arg_i = 0.5 * ((v @ rbm_ph.U.T)[:, None] - (vp @ rbm_ph.U.T)[None, :])
    sig = complex_sigmoid(arg_r, arg_i)  # (i, j, ..., na)
    bs = (v.shape[0], vp.shape[0], *v.shape[1:-1])
    n, nh, na = rbm_am.n, rbm_am.nh, rbm_am.na
    if phase:
        temp = v[:, None] - vp[None, :]
        sig = sig * 1j
        ab = np.zeros((*bs, na), dtype=np.complex128)
    else:
        temp = v[:, None] + vp[None, :]
        ab = sig
    U_grad = 0.5 * sig[..., :, None] * temp[..., None, :]
    return np.concatenate(
        [np.zeros((*bs, nh * n), dtype=np.complex128), U_grad.reshape(*bs, -1),
         np.zeros((*bs, n), dtype=np.complex128), np.zeros((*bs, nh), dtype=np.complex128), ab], axis=-1)


def rotate_rho_probs(rho_fn, basis, states, unitaries=None, include_extras=False):
    """utils/unitaries.py:236-283 with rho evaluated by ``rho_fn(v) -> complex (2^S, 2^S, B)``."""
    Ut, v = rotate_basis_state(basis, states, unitaries)
    Ut2 = np.einsum("ib,jb->ijb", Ut, np.conj(Ut))
    UrhoU_v = Ut2 * rho_fn(v)
    UrhoU = UrhoU_v.real.sum((0, 1))
    return (UrhoU, UrhoU_v, v) if include_extras else UrhoU


def dm_rotated_gradient(rbm_am, rbm_ph, basis, samples, unitaries=None):
    """DensityMatrix.rotated_gradient (nn_states/density_matrix.py:281-330)."""
    UrhoU, UrhoU_v, v = rotate_rho_probs(lambda x: dm_rho(rbm_am, rbm_ph, x), basis, samples, unitaries, True)
    inv = 1.0 / (UrhoU + 1e-8)
    am = rbm_am.gamma_grad(v, v, +1, expand=True) + dm_pi_grad(rbm_am, rbm_ph, v, v, phase=False)
    ph = 1j * rbm_ph.gamma_grad(v, v, -1, expand=True) + dm_pi_grad(rbm_am, rbm_ph, v, v, phase=True)
    out = []
    for g in (am, ph):
        rot = -np.real(np.einsum("ijb,ijbg->bg", UrhoU_v, g))
        out.append(np.einsum("b,bg->g", inv, rot))
    return out[0], out[1]


def dm_gradient(rbm_am, rbm_ph, samples, bases, unitaries=None):
    """NeuralStateBase.gradient for a DensityMatrix (nn_states/neural_state.py:326-369)."""
    samples = np.asarray(samples, dtype=np.float64)
    g = [np.zeros(rbm_am.num_pars), np.zeros(rbm_ph.num_pars)]
    uniq, inv = group_by_basis(bases)
    for i in range(uniq.shape[0]):
        sub = samples[inv == i]
        if np.any(uniq[i] != "Z"):
            ga, gp = dm_rotated_gradient(rbm_am, rbm_ph, uniq[i], sub, unitaries)
            g[0] += ga
            g[1] += gp
        else:
            g[0] += rbm_am.effective_energy_gradient(sub)
    return g


# --------------------------------------------------------------------------------------------
# Metrics with measurement bases (utils/training_statistics.py)
# --------------------------------------------------------------------------------------------
def KL_complex(rbm_am, rbm_ph, target_by_basis, bases, unitaries=None):
    """ts.KL for a WaveFunctionBase with bases and a per-basis target dict
    (utils/training_statistics.py:185-200): rotate the full psi, average the per-basis KL."""
    space = hilbert_rows(rbm_am.n)
    Z = rbm_am.partition(space)
    psi = psi_complex(rbm_am, rbm_ph, space)
    total = 0.0
    for basis in bases:
        psi_r = kron_rotate(basis, psi, unitaries)
        total += single_basis_KL(np.abs(target_by_basis[basis]) ** 2, np.abs(psi_r) ** 2 / Z)
    return total / float(len(bases))


def NLL_complex(rbm_am, rbm_ph, samples, sample_bases, unitaries=None):
    """ts.NLL with sample_bases for a wavefunction (utils/training_statistics.py:103-128)."""
    samples = np.asarray(samples, dtype=np.float64)
    Z = rbm_am.partition(hilbert_rows(rbm_am.n))
    uniq, inv = group_by_basis(sample_bases)
    nll = 0.0
    for i in range(uniq.shape[0]):
        sub = samples[inv == i]
        if np.any(uniq[i] != "Z"):
            Upsi = rotate_psi_inner_prod(lambda x: psi_complex(rbm_am, rbm_ph, x), uniq[i], sub, unitaries)
            probs = np.abs(Upsi) ** 2 / Z
        else:
            probs = np.exp(-rbm_am.effective_energy(sub)) / Z
        nll -= probs_to_logits(probs).sum()
    return float(nll / len(samples))


def KL_dm(rbm_am, rbm_ph, target_rho, bases, unitaries=None):
    """ts.KL for a DensityMatrix with a full target matrix (utils/training_statistics.py:201-216)."""
    space = hilbert_rows(rbm_am.n)
    Z = rbm_am.partition(space)
    idx = basis_index(space)
    total = 0.0
    for basis in bases:
        Ut, v = rotate_basis_state(basis, space, unitaries)
        Ut2 = np.einsum("ib,jb->ijb", Ut, np.conj(Ut))
        vi = basis_index(v)  # (2^S, B)
        t_probs = (Ut2 * target_rho[vi[:, None, :], vi[None, :, :]]).real.sum((0, 1))
        n_probs = rotate_rho_probs(lambda x: dm_rho(rbm_am, rbm_ph, x), basis, space, unitaries) / Z
        total += single_basis_KL(t_probs, n_probs)
    del idx
    return total / float(len(bases))


def NLL_dm(rbm_am, rbm_ph, samples, sample_bases, unitaries=None):
    samples = np.asarray(samples, dtype=np.float64)
    Z = rbm_am.partition(hilbert_rows(rbm_am.n))
    uniq, inv = group_by_basis(sample_bases)
    nll = 0.0
    for i in range(uniq.shape[0]):
        sub = samples[inv == i]
        if np.any(uniq[i] != "Z"):
            probs = rotate_rho_probs(lambda x: dm_rho(rbm_am, rbm_ph, x), uniq[i], sub, unitaries) / Z
        else:
            probs = np.exp(-rbm_am.effective_energy(sub)) / Z
        nll -= probs_to_logits(probs).sum()
    return float(nll / len(samples))


def fidelity_dm(rbm_am, rbm_ph, target_rho):
    """Mixed-state fidelity (utils/training_statistics.py:60-76): (sum sqrt |eig(target rho)|)^2."""
    space = hilbert_rows(rbm_am.n)
    rho = dm_rho(rbm_am, rbm_ph, space, space) / rbm_am.partition(space)
    eig = np.abs(np.linalg.eigvals(target_rho @ rho).real)
    return float(np.sum(np.sqrt(eig)) ** 2)


def exact_gradient_with_bases(kind, rbm_am, rbm_ph, samples, bases, unitaries=None):
    """compute_exact_gradients with bases (nn_states/neural_state.py:386-415)."""
    fn = complex_gradient if kind == "complex" else dm_gradient
    g = fn(rbm_am, rbm_ph, samples, bases, unitaries)
    g = [x / float(len(samples)) for x in g]
    space = hilbert_rows(rbm_am.n)
    probs = np.exp(-rbm_am.effective_energy(space))
    probs = probs / probs.sum()
    g[0] = g[0] - rbm_am.effective_energy_gradient(space, reduce=False).T @ probs
    return g


# ------------------------------------------------------------------------------- local estimators
def pauli_magnetisation(axis, weight_fn, v, absolute=False):
    """SigmaX / SigmaY / SigmaZ .apply (observables/pauli.py:50-83, 100-135, 151-171): per-sample magnetisation
    (1/N) sum_i Re[c_i w(v with unit i flipped, v)], c_i = 1 (X) or i (2 v_i - 1) (Y); Z is the mean of the +-1
    spins. ``weight_fn(vp, v)`` is the importance-sampling weight (neural_state.py:283-300): psi(vp) / psi(v) for
    wavefunctions (wavefunction.py:84-92), rho(vp, v) / rho(v, v) for density matrices (density_matrix.py:275-279)."""
    v = np.asarray(v, dtype=np.float64)
    if axis == "Z":
        res = (2.0 * v - 1.0).mean(1)
        return np.abs(res) if absolute else res
    total = np.zeros(v.shape[0], dtype=np.complex128)
    for i in range(v.shape[1]):
        vf = v.copy()
        vf[:, i] = np.abs(vf[:, i] - 1.0)  # flip_spin, pauli.py:23-33
        coeff = 1.0 if axis == "X" else 1j * (2.0 * v[:, i] - 1.0)
        total = total + coeff * weight_fn(vf, v)
    res = total.real / v.shape[1]
    return np.abs(res) if absolute else res


def wavefunction_weight(psi_fn):
    return lambda vp, v: psi_fn(vp) / psi_fn(v)


def dm_weight(rbm_am, rbm_ph):
    return lambda vp, v: dm_rho(rbm_am, rbm_ph, vp, v, expand=False) / dm_rho(rbm_am, rbm_ph, v, v, expand=False).real


def tfim_chain_energy(rbm_am, v, h):
    """TFIMChainEnergy.apply (examples/Tutorial4_DataGeneration_CalculateObservables/quantum_ising_chain.py:45-78):
    -(sum_i s_i s_{i+1} + h sum_i psi(v^(i)) / psi(v)) / N with open boundaries, positive wavefunction."""
    v = np.asarray(v, dtype=np.float64)
    s = 2.0 * v - 1.0
    log_psi = -0.5 * rbm_am.effective_energy(v)
    flipped = np.empty(v.shape)
    for i in range(v.shape[1]):
        vf = v.copy()
        vf[:, i] = 1.0 - vf[:, i]
        flipped[:, i] = -0.5 * rbm_am.effective_energy(vf)
    m = flipped.max(1, keepdims=True)
    lse = (m + np.log(np.exp(flipped - m).sum(1, keepdims=True)))[:, 0]
    transverse = np.exp(lse - log_psi)
    return -(h * transverse + (s[:, :-1] * s[:, 1:]).sum(1)) / v.shape[1]


def swap_estimator(weight_fn, v, A):
    """SWAP.apply (observables/entanglement.py:58-80): replica 2 = the batch rolled by one row (torch.roll(s, 1, 0)),
    region A exchanged between the replicas, Re[w(s1', s1) w(s2', s2)]."""
    s1 = np.asarray(v, dtype=np.float64)
    s2 = np.roll(s1, 1, axis=0)
    s1_, s2_ = s1.copy(), s2.copy()
    s1_[:, A] = s2[:, A]
    s2_[:, A] = s1[:, A]
    return (weight_fn(s1_, s1) * weight_fn(s2_, s2)).real
