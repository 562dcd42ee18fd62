"""GPU parity of the tcgen05 Gibbs kernel for layers of 65..128 units (csrc/gibbs_umma_wide.cu) against the numpy oracle:
BinaryRBM.gibbs_steps (reference rbm/binary_rbm.py:200-228) and PurificationRBM.gibbs_steps (rbm/purification_rbm.py:309-337)
under the library's Philox convention — bit-exact (index / byte work), for float64 rows, packed rows and chains started in the
kernel; and sample for sample equal to the all-FP64 / composed path (mode 1) on many chains and extreme parameters."""
import numpy as np
import pytest
import torch

from helpers import load_binary, load_purification, make_binary_params, make_purification_params, random_states
from oracle import rbm_oracle as orc_mod
from oracle.rbm_oracle import BinaryRBMOracle, PurificationRBMOracle

pytestmark = pytest.mark.gpu

WIDE_SHAPES = [(128, 128), (100, 100), (65, 128), (128, 64), (70, 3), (5, 100), (97, 33), (64, 65), (65, 65), (96, 127)]


def _pair(n, nh, seed, **kw):
    from qucumber_b200 import BinaryRBM

    W, b, c = make_binary_params(n, nh, seed, **kw)
    return load_binary(BinaryRBM(n, nh, gpu=True), W, b, c), BinaryRBMOracle(W, b, c)


@pytest.mark.parametrize("n,nh", WIDE_SHAPES)
@pytest.mark.parametrize("mode", [0, 1])
def test_wide_layers_bit_exact_vs_oracle(lib, cuda, n, nh, mode):
    """mode 0: the tcgen05 kernel; mode 1: the FP64 kernel / composed path these shapes took before. 333 chains = two full
    tiles of 128 and a ragged one; chain-id and sweep offsets; split calls and shards."""
    import qucumber_b200 as qb

    rbm, orc = _pair(n, nh, 701)
    v0 = random_states(333, n, 702)
    lib.qcb_set_gibbs_mode(mode)
    try:
        launches0 = lib.qcb_launch_count()
        got = rbm.gibbs_steps(3, torch.tensor(v0, device=cuda), seed=0xC0FFEE1234, chain_id0=9, sweep0=4)
        launches = lib.qcb_launch_count() - launches0
        one = rbm.gibbs_steps(1, torch.tensor(v0, device=cuda), seed=17)
        bits = rbm.gibbs_steps_packed(3, qb.pack_states(torch.tensor(v0, device=cuda)), seed=0xC0FFEE1234, chain_id0=9, sweep0=4)
        started = rbm.sample_chains(2, 300, seed=11, chain_id0=4, sweep0=1)
        started_bits = rbm.sample_chains(2, 300, seed=11, chain_id0=4, sweep0=1, packed=True)
        a = rbm.gibbs_steps(2, torch.tensor(v0, device=cuda), seed=99, sweep0=0)
        a = rbm.gibbs_steps(1, a, seed=99, sweep0=2)
        lo = rbm.gibbs_steps(3, torch.tensor(v0[:130], device=cuda), seed=99)
        hi = rbm.gibbs_steps(3, torch.tensor(v0[130:], device=cuda), seed=99, chain_id0=130)
    finally:
        lib.qcb_set_gibbs_mode(0)
    if mode == 0:
        assert launches == 3, f"expected the two pack kernels and one tcgen05 Gibbs launch, saw {launches}"
    want = orc.gibbs_steps_philox(3, v0, seed=0xC0FFEE1234, chain_id0=9, sweep0=4)
    assert np.array_equal(got.cpu().numpy(), want)
    assert np.array_equal(one.cpu().numpy(), orc.gibbs_steps_philox(1, v0, seed=17))
    assert np.array_equal(bits.cpu().numpy(), orc_mod.pack_states(want))
    want_started = orc.sample_chains_philox(2, 300, seed=11, chain_id0=4, sweep0=1)
    assert np.array_equal(started.cpu().numpy(), want_started)
    assert np.array_equal(started_bits.cpu().numpy(), orc_mod.pack_states(want_started))
    whole = orc.gibbs_steps_philox(3, v0, seed=99)
    assert np.array_equal(a.cpu().numpy(), whole)
    assert np.array_equal(torch.cat([lo, hi]).cpu().numpy(), whole)


@pytest.mark.parametrize("n,nh,na", [(70, 40, 50), (30, 90, 30), (128, 64, 64), (100, 20, 10)])
def test_wide_purification_vs_oracle(lib, cuda, n, nh, na):
    """The stacked hidden-like layer [h ; a] continues the unit numbering across h and a."""
    from qucumber_b200 import PurificationRBM

    W, U, b, c, d = make_purification_params(n, nh, na, 711)
    rbm = load_purification(PurificationRBM(n, nh, na, gpu=True), W, U, b, c, d)
    orc = PurificationRBMOracle(W, U, b, c, d)
    v0 = random_states(200, n, 712)
    got = rbm.gibbs_steps(3, torch.tensor(v0, device=cuda), seed=77, chain_id0=3, sweep0=1)
    assert np.array_equal(got.cpu().numpy(), orc.gibbs_steps_philox(3, v0, seed=77, chain_id0=3, sweep0=1))
    s = rbm.sample_chains(2, 129, seed=4242, chain_id0=1, sweep0=2)
    assert np.array_equal(s.cpu().numpy(), orc.sample_chains_philox(2, 129, seed=4242, chain_id0=1, sweep0=2))


@pytest.mark.parametrize("n,nh", [(128, 128), (100, 40), (40, 100)])
@pytest.mark.parametrize("regime", ["default", "saturated", "tiny", "zero", "mixed", "huge"])
def test_wide_equals_fp64_path_on_many_chains(lib, cuda, n, nh, regime):
    """50 000 chains (several tiles per tile slot of every CTA) x 4 sweeps: the integer filter + exact cold path makes the
    decisions of the all-FP64 evaluation, whatever the scale of the parameters."""
    from qucumber_b200 import BinaryRBM

    rng = np.random.default_rng(721)
    W, b, c = make_binary_params(n, nh, 722)
    if regime == "saturated":
        W, b, c = W * 20, b * 20, c * 20
    elif regime == "tiny":
        W, b, c = W * 1e-6, b * 1e-6, c * 1e-6
    elif regime == "zero":
        W, b, c = W * 0, b * 0, c * 0
    elif regime == "mixed":
        W = W * np.exp(rng.uniform(-12, 2, size=W.shape))
    elif regime == "huge":
        W, b, c = W * 1e6, b * 1e3, c * 1e3
    rbm = load_binary(BinaryRBM(n, nh, gpu=True), W, b, c)
    C = 50_000
    g = torch.Generator(device="cuda").manual_seed(5)
    v0 = (torch.rand(C, n, generator=g, device=cuda) < 0.5).double()
    fast = rbm.gibbs_steps(4, v0, seed=2025, chain_id0=123456789012, sweep0=3)
    lib.qcb_set_gibbs_mode(1)
    try:
        slow = rbm.gibbs_steps(4, v0, seed=2025, chain_id0=123456789012, sweep0=3)
    finally:
        lib.qcb_set_gibbs_mode(0)
    assert torch.equal(fast, slow)
    assert set(fast.unique().tolist()) <= {0.0, 1.0}


def test_wide_statistics_vs_conditional(lib, cuda):
    """One sweep from a fixed visible state: the empirical mean of v' follows sum_h p(v' | h) p(h | v) within 5 sigma
    (N = nh = 100; the hidden marginal is sampled, so compare with a 2-stage Monte-Carlo of the oracle's conditionals)."""
    n = nh = 100
    rbm, orc = _pair(n, nh, 731)
    v0 = np.tile(random_states(1, n, 732), (200_000, 1))
    vk = rbm.gibbs_steps(1, torch.tensor(v0, device=cuda), seed=31).cpu().numpy()
    rng = np.random.default_rng(733)
    ph = orc.prob_h_given_v(v0[:1])
    h = (rng.random((200_000, nh)) < ph).astype(np.float64)
    pv = orc.prob_v_given_h(h).mean(0)
    sigma = np.sqrt(0.25 / 200_000) * 2  # two independent Monte-Carlo estimates
    assert np.all(np.abs(vk.mean(0) - pv) < 5 * sigma)


@pytest.mark.parametrize("n,nh", [(10, 10), (16, 16), (3, 12), (13, 5)])
def test_table_kernel_full_size_properties(lib, cuda, n, nh):
    """Layers of at most 16 units run the thread-per-chain table kernel (csrc/gibbs_table.cu).  At the Tutorial-4 size (2^20
    chains) it makes the decisions of the all-FP64 path, halves sampled separately (by global chain id) equal one call, and
    two calls of k / 2 sweeps equal one of k; packed rows and chains started in the kernel agree with the float64 rows."""
    import qucumber_b200 as qb
    from qucumber_b200 import BinaryRBM

    W, b, c = make_binary_params(n, nh, 741, bias_scale=0.4, weight_scale=1.3)
    rbm = load_binary(BinaryRBM(n, nh, gpu=True), W, b, c)
    C = 1 << 20
    g = torch.Generator(device="cuda").manual_seed(6)
    v0 = (torch.rand(C, n, generator=g, device=cuda) < 0.5).double()
    launches0 = lib.qcb_launch_count()
    whole = rbm.gibbs_steps(6, v0, seed=31415, chain_id0=1 << 33, sweep0=2)
    assert lib.qcb_launch_count() - launches0 == 2  # table build + one kernel
    lib.qcb_set_gibbs_mode(1)
    try:
        ref = rbm.gibbs_steps(6, v0, seed=31415, chain_id0=1 << 33, sweep0=2)
    finally:
        lib.qcb_set_gibbs_mode(0)
    assert torch.equal(whole, ref)
    half = C // 2 + 77
    lo = rbm.gibbs_steps(6, v0[:half], seed=31415, chain_id0=1 << 33, sweep0=2)
    hi = rbm.gibbs_steps(6, v0[half:], seed=31415, chain_id0=(1 << 33) + half, sweep0=2)
    assert torch.equal(torch.cat([lo, hi]), whole)
    two = rbm.gibbs_steps(3, rbm.gibbs_steps(3, v0, seed=31415, chain_id0=1 << 33, sweep0=2), seed=31415, chain_id0=1 << 33, sweep0=5)
    assert torch.equal(two, whole)
    bits = rbm.gibbs_steps_packed(6, qb.pack_states(v0), seed=31415, chain_id0=1 << 33, sweep0=2)
    assert torch.equal(qb.unpack_states(bits, n), whole)
    started = rbm.sample_chains(4, 100_000, seed=99, chain_id0=5)
    lib.qcb_set_gibbs_mode(1)
    try:
        started_ref = rbm.sample_chains(4, 100_000, seed=99, chain_id0=5)
    finally:
        lib.qcb_set_gibbs_mode(0)
    assert torch.equal(started, started_ref)
