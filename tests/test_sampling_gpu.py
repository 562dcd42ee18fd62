"""GPU parity of the sampling entry points added for the end-to-end path: chains started in-kernel
(NeuralStateBase.sample, reference nn_states/neural_state.py:107-131), packed chain states, the packed host call, and
the single-step samplers (reference rbm/binary_rbm.py:170-198) under fed and Philox uniforms.
Everything here is bit-exact against the numpy oracle (integer / index work)."""
import ctypes

import numpy as np
import pytest
import torch

from helpers import load_binary, load_purification, make_binary_params, make_purification_params, random_states
from oracle import rbm_oracle as orc_mod
from oracle.rbm_oracle import BinaryRBMOracle, PurificationRBMOracle

pytestmark = pytest.mark.gpu

SHAPES = [(10, 10), (64, 64), (4, 9), (20, 17), (33, 70), (100, 40), (1, 1), (57, 64)]


def _pair(n, nh, seed):
    from qucumber_b200 import BinaryRBM

    W, b, c = make_binary_params(n, nh, seed)
    return load_binary(BinaryRBM(n, nh, gpu=True), W, b, c), BinaryRBMOracle(W, b, c)


@pytest.mark.parametrize("n", [1, 7, 8, 9, 10, 31, 32, 33, 63, 64, 65, 100, 128])
def test_pack_unpack_equals_numpy_packbits(lib, cuda, n):
    import qucumber_b200 as qb

    v = random_states(1001, n, n)
    bits = qb.pack_states(torch.tensor(v, device=cuda))
    assert bits.dtype == torch.uint8 and bits.shape == (1001, (n + 7) // 8)
    assert np.array_equal(bits.cpu().numpy(), np.packbits(v.astype(np.uint8), axis=1))
    assert np.array_equal(qb.unpack_states(bits, n).cpu().numpy(), v)
    # leading dimensions and empty input
    v3 = torch.tensor(v[:24].reshape(2, 12, n), device=cuda)
    assert qb.pack_states(v3).shape == (2, 12, (n + 7) // 8)
    assert torch.equal(qb.unpack_states(qb.pack_states(v3), n), v3)
    assert qb.pack_states(torch.zeros(0, n, dtype=torch.double, device=cuda)).shape == (0, (n + 7) // 8)


@pytest.mark.parametrize("n,nh", SHAPES)
@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4, 5])
def test_sample_chains_bit_exact_vs_oracle(lib, cuda, n, nh, mode):
    """v0 drawn in the kernel (Philox stream 5) + k sweeps == the oracle's restatement, for every kernel (mode 0 automatic,
    1 forces the FP64 kernel, which goes through the separate init / pack kernels, 2 the tcgen05 / TMEM kernel where the
    shape fits, 3 the mma.sync kernel)."""
    rbm, orc = _pair(n, nh, 61)
    lib.qcb_set_gibbs_mode(mode)
    try:
        got = rbm.sample_chains(3, 333, seed=0xABCDEF0123, chain_id0=11, sweep0=5)
        bits = rbm.sample_chains(3, 333, seed=0xABCDEF0123, chain_id0=11, sweep0=5, packed=True)
        v00 = rbm.sample_chains(0, 50, seed=77, chain_id0=3, sweep0=9)
    finally:
        lib.qcb_set_gibbs_mode(0)
    want = orc.sample_chains_philox(3, 333, seed=0xABCDEF0123, chain_id0=11, sweep0=5)
    assert got.shape == (333, n) and got.dtype == torch.double
    assert np.array_equal(got.cpu().numpy(), want)
    assert np.array_equal(bits.cpu().numpy(), orc_mod.pack_states(want))
    assert np.array_equal(v00.cpu().numpy(), orc_mod.philox_initial_states(77, np.arange(3, 53), 9, n))


def test_sample_chains_purification(lib, cuda):
    from qucumber_b200 import PurificationRBM

    for (n, nh, na) in [(2, 2, 2), (6, 5, 7), (10, 10, 10), (20, 40, 30), (40, 20, 30), (64, 32, 32)]:  # the last two: tcgen05 shapes
        W, U, b, c, d = make_purification_params(n, nh, na, 5)
        rbm = load_purification(PurificationRBM(n, nh, na, gpu=True), W, U, b, c, d)
        orc = PurificationRBMOracle(W, U, b, c, d)
        got = rbm.sample_chains(4, 129, seed=4242, chain_id0=1, sweep0=2)
        assert np.array_equal(got.cpu().numpy(), orc.sample_chains_philox(4, 129, seed=4242, chain_id0=1, sweep0=2))
        bits = rbm.sample_chains(4, 129, seed=4242, chain_id0=1, sweep0=2, packed=True)
        assert np.array_equal(bits.cpu().numpy(), orc_mod.pack_states(got.cpu().numpy()))


@pytest.mark.parametrize("n,nh", SHAPES)
def test_gibbs_packed_equals_float64_rows(lib, cuda, n, nh):
    import qucumber_b200 as qb

    rbm, orc = _pair(n, nh, 62)
    v0 = random_states(517, n, 63)
    want = rbm.gibbs_steps(5, torch.tensor(v0, device=cuda), seed=31337, chain_id0=2, sweep0=1)
    bits0 = qb.pack_states(torch.tensor(v0, device=cuda))
    keep = bits0.clone()
    got = rbm.gibbs_steps_packed(5, bits0, seed=31337, chain_id0=2, sweep0=1)
    assert torch.equal(bits0, keep) and got.data_ptr() != bits0.data_ptr()
    assert torch.equal(qb.unpack_states(got, n), want)
    inplace = rbm.gibbs_steps_packed(5, bits0, overwrite=True, seed=31337, chain_id0=2, sweep0=1)
    assert inplace.data_ptr() == bits0.data_ptr() and torch.equal(inplace, got)
    assert np.array_equal(want.cpu().numpy(), orc.gibbs_steps_philox(5, v0, seed=31337, chain_id0=2, sweep0=1))


@pytest.mark.parametrize("n,nh,chains", [(64, 64, 700_001), (10, 10, 5000), (33, 70, 3000), (64, 64, 16)])
@pytest.mark.parametrize("draw", [0, 1])
def test_host_packed_call_equals_device_call(lib, cuda, n, nh, chains, draw):
    """qcb_rbm_gibbs_host_packed (pinned host buffers, sliced pipeline) == the device calls, chain for chain."""
    from qucumber_b200 import _lib

    rbm, _ = _pair(n, nh, 64)
    rb = (n + 7) // 8
    rng = np.random.default_rng(65)
    v0 = (rng.random((chains, n)) < 0.5)
    host = torch.from_numpy(np.packbits(v0, axis=1)).pin_memory()
    assert host.shape == (chains, rb)
    if draw:
        want = rbm.sample_chains(3, chains, seed=99, chain_id0=1000, sweep0=4, packed=True)
    else:
        want = rbm.gibbs_steps_packed(3, host.to(cuda), seed=99, chain_id0=1000, sweep0=4)
    nbytes = lib.qcb_gibbs_host_packed_scratch_bytes(n, nh, 0, chains)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=cuda)
    W, b, c = (t.detach().cpu().contiguous() for t in (rbm.weights, rbm.visible_bias, rbm.hidden_bias))
    rc = lib.qcb_rbm_gibbs_host_packed(_lib.ptr(W), None, _lib.ptr(b), _lib.ptr(c), None, n, nh, 0, _lib.ptr(host), chains, 3, 99, 1000,
                                       4, draw, _lib.ptr(scratch), nbytes, _lib.stream_ptr(cuda))
    _lib.check(rc, "qcb_rbm_gibbs_host_packed")
    assert torch.equal(host, want.cpu())
    # too small a scratch buffer is an error, not a crash
    rc = lib.qcb_rbm_gibbs_host_packed(_lib.ptr(W), None, _lib.ptr(b), _lib.ptr(c), None, n, nh, 0, _lib.ptr(host), chains, 3, 99, 1000,
                                       4, draw, _lib.ptr(scratch), 16, _lib.stream_ptr(cuda))
    assert rc != 0


def test_state_sample_starts_in_kernel_and_is_reproducible(lib, cuda):
    """nn_state.sample(k, num_samples): no host draw, torch.manual_seed reproduces it, packed == float64 rows
    (reference contract: tests/test_models_misc.py:44-76)."""
    import qucumber_b200 as qb

    for state in (qb.PositiveWaveFunction(10, gpu=True), qb.ComplexWaveFunction(6, gpu=True), qb.DensityMatrix(4, gpu=True)):
        n = state.num_visible
        torch.manual_seed(1234)
        a = state.sample(5, 1000)
        torch.manual_seed(1234)
        b = state.sample(5, 1000)
        c = state.sample(5, 1000)
        torch.manual_seed(1234)
        p = state.sample(5, 1000, packed=True)
        assert a.shape == (1000, n) and a.dtype == torch.double and a.device.type == "cuda"
        assert set(a.unique().tolist()) <= {0.0, 1.0}
        assert torch.equal(a, b) and not torch.equal(a, c)
        assert torch.equal(qb.unpack_states(p, n), a)
        # an explicit initial state keeps the reference's route (and its overwrite aliasing)
        v0 = torch.zeros(7, n, dtype=torch.double, device=cuda)
        out = state.sample(2, initial_state=v0, overwrite=True)
        assert out.data_ptr() == v0.data_ptr()


def test_sample_statistics_vs_exact_distribution(lib, cuda):
    """Chains started in-kernel follow exp(-E)/Z: chi-square of the 1024-bin histogram (N = 10, 2^21 samples)."""
    rbm, orc = _pair(10, 10, 51)
    chains = 1 << 21
    vk = rbm.sample_chains(50, chains, seed=777)
    idx = (vk.to(torch.int64) * (2 ** torch.arange(9, -1, -1, device=cuda))).sum(1)
    counts = torch.bincount(idx, minlength=1024).cpu().numpy().astype(np.float64)
    p = np.exp(-orc.effective_energy(orc_mod.hilbert_rows(10)))
    p /= p.sum()
    chi2 = float((((counts - chains * p) ** 2) / (chains * p)).sum())
    assert abs(chi2 - 1023) < 5 * np.sqrt(2 * 1023), f"chi2 = {chi2:.1f} for 1023 dof"
    # the initial state itself is fair: every unit is set in half of the chains (5 sigma)
    v0 = rbm.sample_chains(0, chains, seed=778)
    assert torch.all((v0.mean(0) - 0.5).abs() < 5 * 0.5 / np.sqrt(chains))


# ---- single conditional steps ----------------------------------------------------------
@pytest.mark.parametrize("n,nh", [(10, 10), (64, 64), (4, 9), (33, 70)])
def test_single_step_samplers_fed_uniforms(lib, cuda, n, nh):
    """sample_h_given_v / sample_v_given_h == (u < prob) for fed uniforms — torch.bernoulli(p, out=p) under a replayed
    stream (rbm/binary_rbm.py:170-198); in-place on `out` like the reference."""
    rbm, orc = _pair(n, nh, 71)
    v = random_states(300, n, 72)
    h = random_states(300, nh, 73)
    rng = np.random.default_rng(74)
    uh, uv = rng.random((300, nh)), rng.random((300, n))
    got_h = rbm.sample_h_given_v(torch.tensor(v, device=cuda), uniforms=torch.tensor(uh, device=cuda))
    got_v = rbm.sample_v_given_h(torch.tensor(h, device=cuda), uniforms=torch.tensor(uv, device=cuda))
    assert np.array_equal(got_h.cpu().numpy(), (uh < orc.prob_h_given_v(v)).astype(np.float64))
    assert np.array_equal(got_v.cpu().numpy(), (uv < orc.prob_v_given_h(h)).astype(np.float64))
    out = torch.empty(300, nh, dtype=torch.double, device=cuda)
    res = rbm.sample_h_given_v(torch.tensor(v, device=cuda), out=out, uniforms=torch.tensor(uh, device=cuda))
    assert res.data_ptr() == out.data_ptr() and torch.equal(res, got_h)
    # uniforms that sit exactly on p: strict comparison, like torch.bernoulli's u < p
    p = orc.prob_h_given_v(v)
    on = rbm.sample_h_given_v(torch.tensor(v, device=cuda), uniforms=rbm.prob_h_given_v(torch.tensor(v, device=cuda)))
    assert float(on.sum()) == 0.0 and p.min() > 0.0


@pytest.mark.parametrize("count", [1, 3, 4, 5, 1000, 4097])
def test_bernoulli_philox_stream_vs_oracle(lib, cuda, count):
    """qcb_bernoulli's own stream (include/qcb.h: slot s -> call s / 4, word s % 4, 53-bit uniform from two calls)
    against oracle.philox_uniforms, through the C ABI with an explicit key / chain / sweep."""
    from qucumber_b200 import _lib

    p = np.random.default_rng(count).random(count)
    pt = torch.tensor(p, device=cuda)
    out = torch.empty_like(pt)
    seed, chain, sweep = 0x1234567890ABCDEF & ((1 << 62) - 1), (1 << 33) + 17, 12345
    _lib.check(lib.qcb_bernoulli(_lib.ptr(pt), count, None, seed, chain, sweep, _lib.ptr(out), _lib.stream_ptr(cuda)), "qcb_bernoulli")
    u = orc_mod.philox_uniforms(seed, np.array([chain], dtype=np.uint64), sweep, 0, count)[0]
    assert np.array_equal(out.cpu().numpy(), (u < p).astype(np.float64))


def test_single_step_samplers_follow_the_conditional(lib, cuda):
    """Own-RNG single steps: the empirical mean over 2^18 rows matches prob_* within 5 sigma per unit."""
    rbm, orc = _pair(10, 10, 75)
    rows = 1 << 18
    v = torch.tensor(np.tile(random_states(1, 10, 76), (rows, 1)), device=cuda)
    torch.manual_seed(5)
    h = rbm.sample_h_given_v(v)
    p = torch.tensor(orc.prob_h_given_v(v[:1].cpu().numpy())[0], device=cuda)
    sigma = (p * (1 - p) / rows).sqrt()
    assert torch.all((h.mean(0) - p).abs() < 5 * sigma)
    assert set(h.unique().tolist()) <= {0.0, 1.0}


# ---- layers beyond the fused kernels (round-1 review: "Gibbs returns QCB_ERR_UNSUPPORTED above 128 units or with both layers > 64")
@pytest.mark.parametrize("n,nh", [(100, 100), (130, 20), (20, 140), (150, 150), (65, 65)])
def test_gibbs_large_layers_composed_path(lib, cuda, n, nh):
    """Layers the fused kernels cannot hold take one affine + Bernoulli pass per conditional update (gibbs_large.cu): same
    Philox convention, exact FP64 decisions — bit-exact against the oracle under its own draws and under fed uniforms, shard- and
    split-invariant like the fused kernels; packed rows and in-kernel starts (N <= 128) work through the conversions."""
    import qucumber_b200 as qb

    rbm, orc = _pair(n, nh, 401)
    v0 = random_states(257, n, 402)
    got = rbm.gibbs_steps(3, torch.tensor(v0, device=cuda), seed=0xFEEDBEEF, chain_id0=7, sweep0=2)
    assert np.array_equal(got.cpu().numpy(), orc.gibbs_steps_philox(3, v0, seed=0xFEEDBEEF, chain_id0=7, sweep0=2))
    u = np.random.default_rng(403).random((2, 257, nh + n))
    fed = rbm.gibbs_steps(2, torch.tensor(v0, device=cuda), uniforms=torch.tensor(u, device=cuda))
    assert np.array_equal(fed.cpu().numpy(), orc.gibbs_steps_fed(2, v0, u))
    lo = rbm.gibbs_steps(3, torch.tensor(v0[:100], device=cuda), seed=5)
    hi = rbm.gibbs_steps(3, torch.tensor(v0[100:], device=cuda), seed=5, chain_id0=100)
    assert torch.equal(torch.cat([lo, hi]), rbm.gibbs_steps(3, torch.tensor(v0, device=cuda), seed=5))
    bits = rbm.gibbs_steps_packed(3, qb.pack_states(torch.tensor(v0, device=cuda)), seed=5)
    assert torch.equal(qb.unpack_states(bits, n), torch.cat([lo, hi]))
    if n <= 128:
        s = rbm.sample_chains(2, 300, seed=11, chain_id0=4)
        assert np.array_equal(s.cpu().numpy(), orc.sample_chains_philox(2, 300, seed=11, chain_id0=4))


def test_gibbs_large_purification_and_cd_step(lib, cuda):
    """PurificationRBM with nh + na = 150 hidden-like units (the stacked layer continues the unit numbering across h and a),
    and the CD-k gradient call on a 100 x 100 BinaryRBM (Gibbs through the composed path, gradients through the FP64 kernels)."""
    from qucumber_b200 import BinaryRBM, PurificationRBM, _lib

    n, nh, na = 30, 90, 60
    W, U, b, c, d = make_purification_params(n, nh, na, 411)
    rbm = load_purification(PurificationRBM(n, nh, na, gpu=True), W, U, b, c, d)
    orc = PurificationRBMOracle(W, U, b, c, d)
    v0 = random_states(99, n, 412)
    got = rbm.gibbs_steps(3, torch.tensor(v0, device=cuda), seed=77, chain_id0=3, sweep0=1)
    assert np.array_equal(got.cpu().numpy(), orc.gibbs_steps_philox(3, v0, seed=77, chain_id0=3, sweep0=1))
    u = np.random.default_rng(413).random((2, 99, nh + na + n))
    fed = rbm.gibbs_steps(2, torch.tensor(v0, device=cuda), uniforms=torch.tensor(u, device=cuda))
    assert np.array_equal(fed.cpu().numpy(), orc.gibbs_steps_fed(2, v0, u))
    n, nh, B, C, k = 100, 100, 300, 200, 2
    Wb, bb, cb = make_binary_params(n, nh, 414)
    big = load_binary(BinaryRBM(n, nh, gpu=True), Wb, bb, cb)
    o = BinaryRBMOracle(Wb, bb, cb)
    pos, neg = random_states(B, n, 415), random_states(C, n, 416)
    vk = o.gibbs_steps_philox(k, neg, seed=555, chain_id0=9, sweep0=3)
    want = o.effective_energy_gradient(pos) / B - o.effective_energy_gradient(vk) / C
    args, keep = big._param_args()
    tp, tn = torch.tensor(pos, device=cuda), torch.tensor(neg, device=cuda)
    g = torch.empty(big.num_pars, dtype=torch.double, device=cuda)
    ws = _lib.workspace(cuda, lib.qcb_workspace_bytes(_lib.OP_CD_STEP, n, nh, 0, max(B, C)))
    _lib.check(lib.qcb_rbm_cd_step(*args, _lib.ptr(tp), B, _lib.ptr(tn), C, k, 555, 9, 3, 1.0 / B, -1.0 / C, _lib.ptr(g), _lib.ptr(ws),
                                   ws.numel(), _lib.stream_ptr(cuda)), "qcb_rbm_cd_step")
    from helpers import assert_close

    assert_close(g, want, what="CD-k gradient, 100 x 100 layer")
