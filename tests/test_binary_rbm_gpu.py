"""GPU parity of the CUDA BinaryRBM against the CPU oracle (same seeded inputs).

Mirrors the reference's tests/test_models_misc.py (aliasing, 1-D inputs, seed replay) and the
property style of tests/test_grads.py, with the tolerance the north star states: 1e-10 relative
for energies/gradients/normalisations, bit-exact samples under fed uniforms."""
import os

import numpy as np
import pytest
import torch

from helpers import assert_close, load_binary, make_binary_params, random_states
from oracle.rbm_oracle import BinaryRBMOracle

pytestmark = pytest.mark.gpu

SHAPES = [(10, 10), (64, 64), (4, 9), (20, 17), (33, 70), (100, 40), (1, 1)]


def _pair(n, nh, seed, cuda):
    from qucumber_b200 import BinaryRBM

    W, b, c = make_binary_params(n, nh, seed)
    rbm = load_binary(BinaryRBM(n, nh, gpu=True), W, b, c)
    assert rbm.weights.is_cuda
    return rbm, BinaryRBMOracle(W, b, c)


@pytest.mark.parametrize("n,nh", SHAPES)
def test_effective_energy(lib, cuda, n, nh):
    rbm, orc = _pair(n, nh, 1, cuda)
    v = random_states(257, n, 2)
    assert_close(rbm.effective_energy(torch.tensor(v, device=cuda)), orc.effective_energy(v), what="energy")
    # 1-D in -> 0-D out, arbitrary leading dims, real-valued input
    assert rbm.effective_energy(torch.tensor(v[0], device=cuda)).dim() == 0
    v3 = np.random.default_rng(3).standard_normal((3, 5, n))
    assert_close(rbm.effective_energy(torch.tensor(v3, device=cuda)), orc.effective_energy(v3), what="energy 3-D real")


def test_softplus_threshold(lib, cuda):
    """F.softplus switches to the identity above 20 (a -2e-9 jump the energies must reproduce)."""
    from qucumber_b200 import BinaryRBM

    n, nh = 6, 5
    W, b, c = make_binary_params(n, nh, 5)
    c = c + np.array([19.9999, 20.0001, 25.0, 0.0, -30.0])
    W = W * 1e-3
    rbm = load_binary(BinaryRBM(n, nh, gpu=True), W, b, c)
    v = random_states(64, n, 6)
    assert_close(rbm.effective_energy(torch.tensor(v, device=cuda)), BinaryRBMOracle(W, b, c).effective_energy(v), rtol=1e-13,
                 what="energy near threshold")


@pytest.mark.parametrize("n,nh", SHAPES)
def test_conditionals(lib, cuda, n, nh):
    rbm, orc = _pair(n, nh, 7, cuda)
    v = random_states(100, n, 8)
    h = random_states(100, nh, 9)
    assert_close(rbm.prob_h_given_v(torch.tensor(v, device=cuda)), orc.prob_h_given_v(v), what="prob_h_given_v")
    assert_close(rbm.prob_v_given_h(torch.tensor(h, device=cuda)), orc.prob_v_given_h(h), what="prob_v_given_h")
    out = torch.zeros(100, nh, dtype=torch.double, device=cuda)
    res = rbm.prob_h_given_v(torch.tensor(v, device=cuda), out=out)
    assert res.data_ptr() == out.data_ptr()
    assert rbm.prob_h_given_v(torch.tensor(v[0], device=cuda)).shape == (nh,)


@pytest.mark.parametrize("n,nh", SHAPES)
@pytest.mark.parametrize("rows", [1, 100, 1000])
def test_gradient_reduce(lib, cuda, n, nh, rows):
    rbm, orc = _pair(n, nh, 11, cuda)
    v = random_states(rows, n, 12)
    assert_close(rbm.effective_energy_gradient(torch.tensor(v, device=cuda)), orc.effective_energy_gradient(v),
                 what="reduced gradient")


@pytest.mark.parametrize("n,nh", [(10, 10), (4, 9), (20, 17)])
def test_gradient_batch(lib, cuda, n, nh):
    rbm, orc = _pair(n, nh, 13, cuda)
    v = random_states(37, n, 14)
    assert_close(rbm.effective_energy_gradient(torch.tensor(v, device=cuda), reduce=False),
                 orc.effective_energy_gradient(v, reduce=False), what="batch gradient")
    v3 = random_states(24, n, 15).reshape(4, 6, n)
    got = rbm.effective_energy_gradient(torch.tensor(v3, device=cuda), reduce=False)
    assert got.shape == (4, 6, rbm.num_pars)
    assert_close(got, orc.effective_energy_gradient(v3, reduce=False), what="batch gradient 3-D")
    g1 = rbm.effective_energy_gradient(torch.tensor(v[0], device=cuda), reduce=False)
    assert g1.shape == (1, rbm.num_pars)


@pytest.mark.parametrize("n,nh", SHAPES)
@pytest.mark.parametrize("chains,k", [(1, 1), (200, 10), (1000, 3)])
def test_gibbs_fed_uniforms_bit_exact(lib, cuda, n, nh, chains, k):
    rbm, orc = _pair(n, nh, 21, cuda)
    v0 = random_states(chains, n, 22)
    u = np.random.default_rng(23).random((k, chains, nh + n))
    got = rbm.gibbs_steps(k, torch.tensor(v0, device=cuda), uniforms=torch.tensor(u, device=cuda))
    want = orc.gibbs_steps_fed(k, v0, u)
    assert np.array_equal(got.cpu().numpy(), want)


@pytest.mark.parametrize("n,nh", SHAPES)
def test_gibbs_philox_bit_exact(lib, cuda, n, nh):
    """With its own Philox draws the kernel reproduces the oracle's restatement of the same stream,
    including chain-id / sweep offsets (what makes sharded runs identical to unsharded ones)."""
    rbm, orc = _pair(n, nh, 31, cuda)
    v0 = random_states(333, n, 32)
    got = rbm.gibbs_steps(4, torch.tensor(v0, device=cuda), seed=0xDEADBEEFCAFE, chain_id0=5, sweep0=7)
    want = orc.gibbs_steps_philox(4, v0, seed=0xDEADBEEFCAFE, chain_id0=5, sweep0=7)
    assert np.array_equal(got.cpu().numpy(), want)
    # split in two calls and in two shards == one call
    a = rbm.gibbs_steps(3, torch.tensor(v0, device=cuda), seed=99, sweep0=0)
    a = rbm.gibbs_steps(2, a, seed=99, sweep0=3)
    b = rbm.gibbs_steps(5, torch.tensor(v0, device=cuda), seed=99)
    assert torch.equal(a, b)
    lo = rbm.gibbs_steps(5, torch.tensor(v0[:100], device=cuda), seed=99)
    hi = rbm.gibbs_steps(5, torch.tensor(v0[100:], device=cuda), seed=99, chain_id0=100)
    assert torch.equal(torch.cat([lo, hi]), b)


def test_gibbs_api_contract(lib, cuda):
    """overwrite aliasing, 1-D input, seed replay (reference tests/test_models_misc.py:44-76,237-258)."""
    rbm, _ = _pair(10, 10, 41, cuda)
    v0 = torch.tensor(random_states(50, 10, 42), device=cuda)
    keep = v0.clone()
    out = rbm.gibbs_steps(5, v0, overwrite=False)
    assert torch.equal(v0, keep) and out.data_ptr() != v0.data_ptr()
    out = rbm.gibbs_steps(5, v0, overwrite=True)
    assert out.data_ptr() == v0.data_ptr()
    one = rbm.gibbs_steps(3, keep[0])
    assert one.shape == (10,)
    torch.manual_seed(1234)
    a = rbm.gibbs_steps(10, keep)
    torch.manual_seed(1234)
    b = rbm.gibbs_steps(10, keep)
    c = rbm.gibbs_steps(10, keep)
    assert torch.equal(a, b) and not torch.equal(a, c)
    h = rbm.sample_h_given_v(keep[0])
    assert h.shape == (10,) and set(h.unique().tolist()) <= {0.0, 1.0}


def test_gibbs_statistics_vs_exact_distribution(lib, cuda):
    """Own-RNG samples follow exp(-E)/Z: chi-square of the 2^N-bin histogram (N = 10, 2e6 samples)."""
    n, nh = 10, 10
    rbm, orc = _pair(n, nh, 51, cuda)
    chains = 2_000_000
    v0 = torch.tensor(random_states(chains, n, 52), device=cuda)
    vk = rbm.gibbs_steps(50, v0, seed=2024)
    idx = (vk.to(torch.int64) * (2 ** torch.arange(n - 1, -1, -1, device=cuda))).sum(1)
    counts = torch.bincount(idx, minlength=1 << n).cpu().numpy().astype(np.float64)
    from oracle.rbm_oracle import hilbert_rows

    p = np.exp(-orc.effective_energy(hilbert_rows(n)))
    p /= p.sum()
    chi2 = float((((counts - chains * p) ** 2) / (chains * p)).sum())
    dof = (1 << n) - 1
    # chi2 ~ N(dof, 2 dof): accept within 5 sigma
    assert abs(chi2 - dof) < 5 * np.sqrt(2 * dof), f"chi2 = {chi2:.1f} for {dof} dof"


@pytest.mark.parametrize("n,nh", [(10, 10), (4, 9), (16, 20), (20, 20)])
def test_partition(lib, cuda, n, nh):
    rbm, orc = _pair(n, nh, 61, cuda)
    want = orc.log_partition()
    out2 = rbm.log_partition().cpu().numpy()
    assert abs((out2[0] + np.log(out2[1])) - want) <= 1e-10 * abs(want)
    assert_close(rbm.partition(), np.float64(np.exp(want)), what="Z from index enumeration")
    if n <= 16:
        from qucumber_b200 import HilbertSpace

        space = HilbertSpace(n).tensor(device=cuda)
        assert_close(rbm.partition(space), np.float64(np.exp(want)), what="Z from explicit space")
    # ranges combine: sharded == unsharded, ragged boundaries included
    total = 1 << n
    cuts = [0, 1, total // 3, total // 2 + 5, total]
    parts = [rbm.log_partition(a, b - a).cpu().numpy() for a, b in zip(cuts[:-1], cuts[1:])]
    m = max(p[0] for p in parts)
    s = sum(p[1] * np.exp(p[0] - m) for p in parts)
    assert abs((m + np.log(s)) - want) <= 1e-10 * abs(want)
    empty = rbm.log_partition(5, 0).cpu().numpy()
    assert empty[1] == 0.0


@pytest.mark.parametrize("case", ["plain", "threshold", "fallback", "wide"])
def test_partition_kernel_paths(lib, cuda, case):
    """log Z through the product-form kernel (with and without pre-activations above torch's softplus
    threshold) and through the log-domain kernel it falls back to when the product could overflow."""
    from qucumber_b200 import BinaryRBM

    n, nh = (12, 12) if case != "wide" else (14, 40)
    W, b, c = make_binary_params(n, nh, 71, bias_scale=0.5)
    if case == "threshold":
        c = c + np.where(np.arange(nh) % 3 == 0, 21.0, 0.0)
    if case == "fallback":
        W, b, c = W * 60.0, b * 60.0, c * 60.0
    rbm = load_binary(BinaryRBM(n, nh, gpu=True), W, b, c)
    want = BinaryRBMOracle(W, b, c).log_partition()
    out2 = rbm.log_partition().cpu().numpy()
    got = out2[0] + np.log(out2[1])
    assert abs(got - want) <= 1e-10 * abs(want), (case, got, want)
    lo = rbm.log_partition(0, 1000).cpu().numpy()
    hi = rbm.log_partition(1000, (1 << n) - 1000).cpu().numpy()
    m = max(lo[0], hi[0])
    both = m + np.log(lo[1] * np.exp(lo[0] - m) + hi[1] * np.exp(hi[0] - m))
    assert abs(both - want) <= 1e-10 * abs(want)


def test_gibbs_host_buffers_pipelined(lib, cuda):
    """The host-buffer entry point (sliced H2D / sample / D2H pipeline) returns exactly what the
    device-resident call returns for the same seed."""
    from qucumber_b200 import BinaryRBM, _lib

    n, nh, C, k = 64, 64, 300_001, 3
    W, b, c = make_binary_params(n, nh, 81)
    rbm = load_binary(BinaryRBM(n, nh, gpu=True), W, b, c)
    v0 = torch.tensor(random_states(C, n, 82))
    want = rbm.gibbs_steps(k, v0.to(cuda), seed=4242, chain_id0=7, sweep0=1).cpu()
    vh = v0.clone().pin_memory()
    Wh, bh, ch = (torch.tensor(x).contiguous().pin_memory() for x in (W, b, c))
    nbytes = lib.qcb_gibbs_host_scratch_bytes(n, nh, 0, C)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=cuda)
    _lib.check(lib.qcb_rbm_gibbs_host(_lib.ptr(Wh), None, _lib.ptr(bh), _lib.ptr(ch), None, n, nh, 0, _lib.ptr(vh), C, k, 4242, 7, 1,
                                      _lib.ptr(scratch), nbytes, _lib.stream_ptr(cuda)), "qcb_rbm_gibbs_host")
    assert torch.equal(vh, want)


@pytest.mark.parametrize("n,nh", SHAPES + [(128, 60), (17, 128)])
def test_gradient_reduce_large_batch_tensor_core_path(lib, cuda, n, nh):
    """Batches of >= 2048 rows take the DMMA reduction (grad_reduce_dmma.cu); same 1e-10 gate,
    binary and real-valued rows, ragged last tile."""
    rbm, orc = _pair(n, nh, 91, cuda)
    v = random_states(5003, n, 92)
    assert_close(rbm.effective_energy_gradient(torch.tensor(v, device=cuda)), orc.effective_energy_gradient(v),
                 what="reduced gradient (large batch)")
    vr = np.random.default_rng(93).standard_normal((2048, n))
    assert_close(rbm.effective_energy_gradient(torch.tensor(vr, device=cuda)), orc.effective_energy_gradient(vr),
                 what="reduced gradient (large batch, real rows)")


def test_cd_step_matches_oracle_philox(lib, cuda):
    """qcb_rbm_cd_step = pos_scale * sum grad E(pos) + neg_scale * sum grad E(v_k) with v_k from the
    kernel's own Philox stream; the negative batch is left untouched."""
    from qucumber_b200 import BinaryRBM, _lib

    n, nh, B, C, k = 12, 20, 300, 500, 4
    W, b, c = make_binary_params(n, nh, 101)
    rbm = load_binary(BinaryRBM(n, nh, gpu=True), W, b, c)
    o = BinaryRBMOracle(W, b, c)
    pos, neg = random_states(B, n, 102), random_states(C, n, 103)
    vk = o.gibbs_steps_philox(k, neg, seed=555, chain_id0=9, sweep0=3)
    want = o.effective_energy_gradient(pos) / B - o.effective_energy_gradient(vk) / C
    args, keep = rbm._param_args()
    tp, tn = torch.tensor(pos, device=cuda), torch.tensor(neg, device=cuda)
    g = torch.empty(rbm.num_pars, dtype=torch.double, device=cuda)
    ws = _lib.workspace(cuda, _lib.workspace_bytes(_lib.OP_CD_STEP, n, nh, 0, max(B, C)))
    _lib.check(lib.qcb_rbm_cd_step(*args, _lib.ptr(tp), B, _lib.ptr(tn), C, k, 555, 9, 3, 1.0 / B, -1.0 / C, _lib.ptr(g), _lib.ptr(ws),
                                   ws.numel(), _lib.stream_ptr(cuda)), "qcb_rbm_cd_step")
    assert_close(g, want, what="CD-k gradient")
    assert np.array_equal(tn.cpu().numpy(), neg)


@pytest.mark.parametrize("n,nh,k,wscale", [(64, 64, 3, None), (40, 50, 2, None), (57, 33, 1, None), (64, 64, 2, 12.0)])
def test_cd_step_packed_vk_matches_oracle(lib, cuda, n, nh, k, wscale):
    """Round 2: where the integer-filter Gibbs kernels and the tcgen05 gradient kernel both take the shape (33..64 units,
    >= 16384 chains), v_k travels between them PACKED inside qcb_rbm_cd_step (ceil(N / 8) bytes per chain). Same gate as the
    float64-row path: the oracle's Philox chains and its gradients at 1e-10, elementwise. Row widths that are not whole words
    (N = 40, 57) exercise the byte-wise loader; the last case has weights the two fixed-point groups cannot hold (max|W| > 8):
    the tcgen05 kernel raises its flag, the rows are unpacked on the device and the FP64 kernel redoes the batch."""
    from qucumber_b200 import BinaryRBM, _lib

    B, C = 1000, 16384 + 37  # ragged last tile
    W, b, c = make_binary_params(n, nh, 1300 + n + nh, weight_scale=wscale)
    if wscale is not None:
        assert np.abs(W).max() > 8.0
    rbm = load_binary(BinaryRBM(n, nh, gpu=True), W, b, c)
    o = BinaryRBMOracle(W, b, c)
    pos, neg = random_states(B, n, 1302), random_states(C, n, 1303)
    vk = o.gibbs_steps_philox(k, neg, seed=4242, chain_id0=5, sweep0=2)
    want = o.effective_energy_gradient(pos) / B - o.effective_energy_gradient(vk) / C
    args, keep = rbm._param_args()
    tp, tn = torch.tensor(pos, device=cuda), torch.tensor(neg, device=cuda)
    g = torch.empty(rbm.num_pars, dtype=torch.double, device=cuda)
    ws = _lib.workspace(cuda, _lib.workspace_bytes(_lib.OP_CD_STEP, n, nh, 0, max(B, C)))
    _lib.check(lib.qcb_rbm_cd_step(*args, _lib.ptr(tp), B, _lib.ptr(tn), C, k, 4242, 5, 2, 1.0 / B, -1.0 / C, _lib.ptr(g), _lib.ptr(ws),
                                   ws.numel(), _lib.stream_ptr(cuda)), "qcb_rbm_cd_step")
    assert_close(g, want, what=f"CD-{k} gradient with packed v_k ({n} x {nh})")
    assert np.array_equal(tn.cpu().numpy(), neg)
    # the same call with the float64-row route forced (FP64 gradient kernels only) gives the same vector
    g2 = torch.empty_like(g)
    lib.qcb_set_grad_reduce_mode(1)
    try:
        _lib.check(lib.qcb_rbm_cd_step(*args, _lib.ptr(tp), B, _lib.ptr(tn), C, k, 4242, 5, 2, 1.0 / B, -1.0 / C, _lib.ptr(g2),
                                       _lib.ptr(ws), ws.numel(), _lib.stream_ptr(cuda)), "qcb_rbm_cd_step")
    finally:
        lib.qcb_set_grad_reduce_mode(0)
    assert_close(g, g2.cpu().numpy(), what="packed v_k route == float64-row route")


@pytest.mark.parametrize("n,nh,B,C,pb,nb_,k", [(7, 5, 53, 90, 10, 17, 2), (10, 10, 230, 460, 100, 200, 3), (20, 30, 64, 64, 16, 16, 2),
                                                (32, 32, 40, 40, 40, 40, 1), (48, 40, 300, 520, 100, 200, 2),
                                                (64, 64, 260, 260, 130, 130, 2), (33, 64, 50, 70, 25, 35, 1), (64, 33, 9, 9, 4, 4, 3)])
def test_cd_epoch_matches_oracle(lib, cuda, n, nh, B, C, pb, nb_, k):
    """qcb_rbm_cd_epoch (a whole epoch of CD-k + SGD in one launch) == the oracle's batch loop:
    g_i = <grad E>_pos_i / B_i - <grad E>_{v_k} / C_i with v_k from the Philox stream (chain id = row of
    the negative batch, sweep = sweep0 + i k + s), W, b, c -= lr g_i; short last batches included."""
    from qucumber_b200 import _lib

    W, b, c = make_binary_params(n, nh, 400 + n)
    pos, neg = random_states(B, n, 401), random_states(C, n, 402)
    lr, seed, sweep0 = 0.05, 777, 11
    num_batches = min(-(-B // pb), -(-C // nb_))
    Wo, bo, co = W.copy(), b.copy(), c.copy()
    g_want = None
    for i in range(num_batches):
        o = BinaryRBMOracle(Wo, bo, co)
        p_i, n_i = pos[i * pb:(i + 1) * pb], neg[i * nb_:(i + 1) * nb_]
        vk = o.gibbs_steps_philox(k, n_i, seed=seed, chain_id0=0, sweep0=sweep0 + i * k)
        g_want = o.effective_energy_gradient(p_i) / len(p_i) - o.effective_energy_gradient(vk) / len(n_i)
        Wo = Wo - lr * g_want[:nh * n].reshape(nh, n)
        bo = bo - lr * g_want[nh * n:nh * n + n]
        co = co - lr * g_want[nh * n + n:]
    tW, tb, tc = (torch.tensor(x, device=cuda) for x in (W, b, c))
    tp, tn = torch.tensor(pos, device=cuda), torch.tensor(neg, device=cuda)
    g_last = torch.empty(nh * n + n + nh, dtype=torch.double, device=cuda)
    _lib.check(lib.qcb_rbm_cd_epoch(_lib.ptr(tW), _lib.ptr(tb), _lib.ptr(tc), n, nh, _lib.ptr(tp), B, pb, _lib.ptr(tn), C, nb_,
                                    num_batches, k, lr, seed, sweep0, _lib.ptr(g_last), _lib.stream_ptr(cuda)), "qcb_rbm_cd_epoch")
    assert_close(tW, Wo, what="weights after the fused epoch")
    assert_close(tb, bo, what="visible bias after the fused epoch")
    assert_close(tc, co, what="hidden bias after the fused epoch")
    assert_close(g_last, g_want, what="gradient of the last batch")
    assert np.array_equal(tn.cpu().numpy(), neg) and np.array_equal(tp.cpu().numpy(), pos)


@pytest.mark.parametrize("n,nh,B,C,pb,nb_,k,scale", [(10, 10, 1000, 2000, 100, 200, 10, None), (7, 5, 53, 90, 10, 17, 3, None),
                                                      (16, 16, 700, 1300, 350, 650, 1, None), (3, 12, 64, 64, 16, 16, 5, None),
                                                      (12, 9, 120, 150, 120, 150, 3, 3.0), (16, 10, 90, 90, 30, 30, 4, 40.0),
                                                      (1, 1, 10, 10, 5, 5, 2, None), (8, 8, 2000, 4096, 1000, 2048, 2, None),
                                                      (32, 32, 300, 300, 100, 100, 3, None), (20, 30, 64, 64, 16, 16, 2, 2.0),
                                                      (31, 17, 500, 700, 250, 350, 1, None), (24, 24, 1000, 2000, 100, 200, 10, None)])
def test_cd_epoch_warp_specialised_equals_single_role_kernel(lib, cuda, n, nh, B, C, pb, nb_, k, scale):
    """Layers of at most 32 units run the warp-specialised epoch kernel (producer warps draw the Philox cells into a ring and
    rebuild float nibble tables of the pre-activations after every update, consumer warps run the chain from those tables
    with a rigorous margin and the exact FP64 path behind it): same
    decisions and (up to 16 units) the same summation order as the single-role kernel — identical parameters and gradients, bit for
    bit there, to rounding for 17..32 units; more
    chains than slots (several rounds per step), sweeps that do not fill a ring buffer, short last batches, strong couplings."""
    from qucumber_b200 import _lib

    W, b, c = make_binary_params(n, nh, 900 + n, bias_scale=0.1 if scale is None else 1.0, weight_scale=scale)
    pos, neg = random_states(B, n, 901), random_states(C, n, 902)
    num_batches = min(-(-B // pb), -(-C // nb_))
    outs = []
    for mode in (0, 1):
        tW, tb, tc = (torch.tensor(x, device=cuda) for x in (W, b, c))
        tp, tn = torch.tensor(pos, device=cuda), torch.tensor(neg, device=cuda)
        g_last = torch.empty(nh * n + n + nh, dtype=torch.double, device=cuda)
        _lib.check(lib.qcb_set_cd_epoch_mode(mode), "mode")
        try:
            _lib.check(lib.qcb_rbm_cd_epoch(_lib.ptr(tW), _lib.ptr(tb), _lib.ptr(tc), n, nh, _lib.ptr(tp), B, pb, _lib.ptr(tn), C, nb_,
                                            num_batches, k, 0.05, 4711, 9, _lib.ptr(g_last), _lib.stream_ptr(cuda)), "qcb_rbm_cd_epoch")
            torch.cuda.synchronize()
        finally:
            _lib.check(lib.qcb_set_cd_epoch_mode(0), "mode")
        outs.append([t.cpu() for t in (tW, tb, tc, g_last)])
    for x, y in zip(*outs):
        if max(n, nh) <= 16:
            assert torch.equal(x, y)
        else:  # 17..32 units: the rows are dealt to different slots, the sums agree to rounding
            assert_close(x, y.numpy(), rtol=1e-12, what="warp-specialised vs single-role epoch kernel")
    assert lib.qcb_set_cd_epoch_mode(2) != 0


def test_cd_epoch_rejects_large_layers(lib, cuda):
    from qucumber_b200 import _lib

    z = torch.zeros(65 * 65, dtype=torch.double, device=cuda)
    rc = lib.qcb_rbm_cd_epoch(_lib.ptr(z), _lib.ptr(z), _lib.ptr(z), 65, 65, _lib.ptr(z), 4, 2, _lib.ptr(z), 4, 2, 2, 1, 0.1, 1, 0, None,
                              _lib.stream_ptr(cuda))
    assert rc != 0 and b"cd_epoch" in lib.qcb_last_error()


def test_empty_and_degenerate_inputs(lib, cuda):
    """Empty batches, zero sweeps and error reporting through the C ABI."""
    from qucumber_b200 import BinaryRBM, _lib

    rbm, orc = _pair(10, 7, 111, cuda)
    empty = torch.zeros(0, 10, dtype=torch.double, device=cuda)
    assert rbm.effective_energy(empty).shape == (0,)
    assert rbm.gibbs_steps(3, empty).shape == (0, 10)
    assert torch.equal(rbm.effective_energy_gradient(empty), torch.zeros(rbm.num_pars, dtype=torch.double, device=cuda))
    v = torch.tensor(random_states(9, 10, 112), device=cuda)
    assert torch.equal(rbm.gibbs_steps(0, v), v)
    with pytest.raises(ValueError):
        rbm.effective_energy(torch.zeros(3, 11, dtype=torch.double, device=cuda))
    big = BinaryRBM(800, 800, gpu=True)
    # 800 x 800: the weights do not fit in shared memory and a row of 800 inputs is beyond the row-tiled kernels (767 inputs
    # is what one CTA stages; tests/test_large_gpu.py covers everything below that)
    with pytest.raises(RuntimeError, match="input units per row tile"):
        big.gibbs_steps(1, torch.zeros(4, 800, dtype=torch.double, device=cuda))
    rc = lib.qcb_rbm_energy(None, None, None, None, None, 4, 4, 0, None, 1, None, None, 0, None)
    assert rc == 1 and b"must not be NULL" in lib.qcb_last_error()


def test_full_size_properties(lib, cuda):
    """Size-independent properties at the BASELINE sizes the oracle cannot reach directly:
    (1) 2^20 chains of N = nh = 64: halves sampled separately (by global chain id) == one call, and two
        calls of k/2 sweeps == one call of k; (2) exact log Z over 2^30 states equals the closed form
        obtained by tracing the *visible* units (2^nh terms) and is additive over index ranges."""
    from qucumber_b200 import BinaryRBM

    n, nh, C = 64, 64, 1 << 20
    W, b, c = make_binary_params(n, nh, 121)
    rbm = load_binary(BinaryRBM(n, nh, gpu=True), W, b, c)
    g = torch.Generator(device="cuda").manual_seed(5)
    v0 = (torch.rand(C, n, generator=g, device=cuda) < 0.5).double()
    whole = rbm.gibbs_steps(4, v0, seed=31337)
    half = C // 2 + 13
    lo = rbm.gibbs_steps(4, v0[:half].clone(), seed=31337)
    hi = rbm.gibbs_steps(4, v0[half:].clone(), seed=31337, chain_id0=half)
    assert torch.equal(torch.cat([lo, hi]), whole)
    two = rbm.gibbs_steps(2, rbm.gibbs_steps(2, v0, seed=31337), seed=31337, sweep0=2)
    assert torch.equal(two, whole)
    assert set(whole.unique().tolist()) == {0.0, 1.0}
    frac = float(whole.mean())
    assert 0.3 < frac < 0.7
    # exact Z at N = 30
    n2, nh2 = 30, 10
    W2, b2, c2 = make_binary_params(n2, nh2, 122, bias_scale=0.3)
    rbm2 = load_binary(BinaryRBM(n2, nh2, gpu=True), W2, b2, c2)
    want = BinaryRBMOracle(W2, b2, c2).log_partition_hidden_side()
    out2 = rbm2.log_partition().cpu().numpy()
    got = out2[0] + np.log(out2[1])
    assert abs(got - want) <= 1e-10 * abs(want), (got, want)
    cut = (1 << n2) // 3 + 7
    a, bb = rbm2.log_partition(0, cut).cpu().numpy(), rbm2.log_partition(cut, (1 << n2) - cut).cpu().numpy()
    m = max(a[0], bb[0])
    both = m + np.log(a[1] * np.exp(a[0] - m) + bb[1] * np.exp(bb[0] - m))
    assert abs(both - want) <= 1e-10 * abs(want)


@pytest.mark.parametrize("n,nh", [(10, 10), (64, 64), (33, 60), (100, 40), (4, 9)])
def test_effective_energy_large_batch_tensor_core_path(lib, cuda, n, nh):
    """Batches >= 2048 rows: theta on the FP64 tensor cores, softplus sum in threshold-exact product form."""
    from qucumber_b200 import BinaryRBM

    rbm, orc = _pair(n, nh, 141, cuda)
    v = random_states(4099, n, 142)
    assert_close(rbm.effective_energy(torch.tensor(v, device=cuda)), orc.effective_energy(v), what="energy (large batch)")
    vr = np.random.default_rng(143).standard_normal((2048, n)) * 3.0
    assert_close(rbm.effective_energy(torch.tensor(vr, device=cuda)), orc.effective_energy(vr), what="energy (large batch, real rows)")
    # pre-activations straddling torch's softplus threshold, and very negative ones
    W, b, c = make_binary_params(n, nh, 144)
    c = c + np.where(np.arange(nh) % 4 == 0, 20.00005, np.where(np.arange(nh) % 4 == 1, -45.0, 0.0))
    W = W * 1e-4
    r2 = load_binary(BinaryRBM(n, nh, gpu=True), W, b, c)
    assert_close(r2.effective_energy(torch.tensor(v, device=cuda)), BinaryRBMOracle(W, b, c).effective_energy(v), rtol=1e-13,
                 what="energy near threshold (large batch)")
    # rows starting 8 bytes off a 16-byte boundary (no TMA row copies) and a NaN entry (stays a NaN, only in its row)
    backing = torch.empty(v.size + 1, dtype=torch.double, device=cuda)
    vt = backing[1:].view(*v.shape)
    vt.copy_(torch.tensor(v))
    assert_close(rbm.effective_energy(vt), orc.effective_energy(v), what="energy (large batch, misaligned rows)")
    vn = torch.tensor(v, device=cuda)
    vn[7, 0] = float("nan")
    en = rbm.effective_energy(vn).cpu().numpy()
    assert np.isnan(en[7]) and np.isfinite(np.delete(en, 7)).all()


def test_integer_filter_kernel_equals_fp64_kernel(lib, cuda):
    """The int8-tensor-core filter kernel (default for layers <= 64 units) and the FP64 DMMA kernel make
    identical decisions on 2^20 chains; the rigorous error bound of the filter is never exceeded."""
    from qucumber_b200 import BinaryRBM, _lib

    for (n, nh, scale) in [(64, 64, None), (10, 10, None), (33, 60, 1.5), (64, 20, 0.01), (48, 64, 0.3), (64, 33, None)]:
        W, b, c = make_binary_params(n, nh, 161, bias_scale=0.5, weight_scale=scale)
        rbm = load_binary(BinaryRBM(n, nh, gpu=True), W, b, c)
        C = 1 << 20 if n == 64 and nh == 64 else 1 << 16
        g = torch.Generator(device="cuda").manual_seed(7)
        v0 = (torch.rand(C, n, generator=g, device=cuda) < 0.5).double()
        try:
            _lib.check(lib.qcb_set_gibbs_mode(1), "mode")
            ref = rbm.gibbs_steps(3, v0, seed=2718, chain_id0=11, sweep0=5)
        finally:
            _lib.check(lib.qcb_set_gibbs_mode(0), "mode")
        for mode in (0, 2, 3, 4, 5):  # automatic, tcgen05 / TMEM kernel (where the shape fits), mma.sync kernel, general tcgen05 kernel
            try:
                _lib.check(lib.qcb_set_gibbs_mode(mode), "mode")
                got = rbm.gibbs_steps(3, v0, seed=2718, chain_id0=11, sweep0=5)
            finally:
                _lib.check(lib.qcb_set_gibbs_mode(0), "mode")
            assert torch.equal(got, ref), (n, nh, mode)
        stats = torch.zeros(2, dtype=torch.int64, device=cuda)
        args, keep = rbm._param_args()
        vv = v0.clone()
        ws = _lib.workspace(cuda, _lib.workspace_bytes(_lib.OP_GIBBS, n, nh, 0, C))
        _lib.check(lib.qcb_debug_gibbs_filter(*args, _lib.ptr(vv), C, 3, 2718, _lib.ptr(stats), _lib.ptr(ws), ws.numel(),
                                              _lib.stream_ptr(cuda)), "debug filter")
        ratio, count = stats[0].item() / 2 ** 20, stats[1].item()
        assert count > 0 and ratio < 0.5, (n, nh, ratio, count)          # observed error well inside the bound
        assert count < 2e-3 * C * 3 * (n + nh), (n, nh, count)            # and the cold path stays rare


@pytest.mark.parametrize("case", ["saturated", "huge", "tiny", "mixed_scales", "zero"])
def test_integer_filter_kernel_extreme_parameters(lib, cuda, case):
    """Regimes where the fixed-point filter cannot decide (pre-activations beyond +-88, weights beyond the
    24-bit window, margins >= 1) must fall through to the exact path: same samples as the FP64 kernel."""
    from qucumber_b200 import BinaryRBM, _lib

    _run_extreme_case(lib, cuda, case, 24, 40)
    _run_extreme_case(lib, cuda, case, 40, 64)   # both layers in 33..64: the tcgen05 kernel's shapes


def _run_extreme_case(lib, cuda, case, n, nh):
    from qucumber_b200 import BinaryRBM, _lib

    C = 4096
    W, b, c = make_binary_params(n, nh, 171, bias_scale=0.5)
    if case == "saturated":
        W *= 40.0            # |x| up to several hundred: e^{-x} overflows float
        c[::3] -= 150.0
        b[::4] += 120.0
    elif case == "huge":
        W *= 1e30            # does not fit the fixed-point image at all
    elif case == "tiny":
        W *= 1e-30
    elif case == "mixed_scales":
        W[0, :] *= 1e6       # one row dominates the scale: the others quantise to (almost) nothing
    elif case == "zero":
        W[:] = 0.0
    rbm = load_binary(BinaryRBM(n, nh, gpu=True), W, b, c)
    g = torch.Generator(device="cuda").manual_seed(9)
    v0 = (torch.rand(C, n, generator=g, device=cuda) < 0.5).double()
    try:
        _lib.check(lib.qcb_set_gibbs_mode(1), "mode")
        ref = rbm.gibbs_steps(4, v0, seed=99, chain_id0=3, sweep0=1)
    finally:
        _lib.check(lib.qcb_set_gibbs_mode(0), "mode")
    for mode in (0, 2, 3, 4, 5):
        try:
            _lib.check(lib.qcb_set_gibbs_mode(mode), "mode")
            got = rbm.gibbs_steps(4, v0, seed=99, chain_id0=3, sweep0=1)
        finally:
            _lib.check(lib.qcb_set_gibbs_mode(0), "mode")
        assert torch.equal(got, ref), (case, n, nh, mode)


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("n,nh,C", [(7, 5, 37), (33, 20, 50), (64, 64, 19), (10, 10, 1), (63, 41, 300)])
def test_gibbs_ragged_and_unaligned_state_tiles(lib, cuda, mode, n, nh, C):
    """State tiles reach the kernels by TMA bulk copies when the span is 16-byte aligned and a multiple of 16 bytes;
    odd N, a ragged last group and an 8-byte-misaligned view must take the plain-copy path and give the same
    samples as the oracle, in place and out of place, for both Gibbs kernels."""
    from qucumber_b200 import BinaryRBM, _lib

    W, b, c = make_binary_params(n, nh, 181 + n)
    rbm = load_binary(BinaryRBM(n, nh, gpu=True), W, b, c)
    o = BinaryRBMOracle(W, b, c)
    v0 = random_states(C, n, 182)
    want = o.gibbs_steps_philox(3, v0, seed=4321, chain_id0=5, sweep0=2)
    backing = torch.zeros(C * n + 1, dtype=torch.double, device=cuda)
    view = backing[1:].view(C, n)          # contiguous, data pointer 8 bytes off a 16-byte boundary
    view.copy_(torch.tensor(v0, device=cuda))
    assert view.data_ptr() % 16 == 8
    try:
        _lib.check(lib.qcb_set_gibbs_mode(mode), "mode")
        out = rbm.gibbs_steps(3, view, seed=4321, chain_id0=5, sweep0=2)
        assert np.array_equal(out.cpu().numpy(), want) and np.array_equal(view.cpu().numpy(), v0)
        rbm.gibbs_steps(3, view, overwrite=True, seed=4321, chain_id0=5, sweep0=2)
        assert np.array_equal(view.cpu().numpy(), want)
        aligned = torch.tensor(v0, device=cuda)
        assert np.array_equal(rbm.gibbs_steps(3, aligned, seed=4321, chain_id0=5, sweep0=2).cpu().numpy(), want)
    finally:
        _lib.check(lib.qcb_set_gibbs_mode(0), "mode")
    assert backing[0].item() == 0.0


def test_reduced_gradient_saturated_units_large_batch(lib, cuda):
    """The tensor-core gradient path evaluates the logistic function with a table + Newton scheme
    (common.cuh: sigmoid_tab); saturated and far-out pre-activations (|theta| up to ~1000) must still match
    the oracle's 1 / (1 + exp(-x)) to 1e-10 in every gradient entry."""
    from qucumber_b200 import BinaryRBM

    n, nh, B = 24, 40, 4096
    W, b, c = make_binary_params(n, nh, 191, bias_scale=0.5)
    W[::2] *= 60.0
    c[1::4] -= 400.0
    c[2::4] += 300.0
    rbm = load_binary(BinaryRBM(n, nh, gpu=True), W, b, c)
    o = BinaryRBMOracle(W, b, c)
    v = random_states(B, n, 192)
    with np.errstate(over="ignore"):
        want = o.effective_energy_gradient(v)
    assert_close(rbm.effective_energy_gradient(torch.tensor(v, device=cuda)), want, what="reduced gradient, saturated units")


@pytest.mark.parametrize("n,nh,case", [(64, 64, "plain"), (32, 64, "plain"), (64, 25, "plain"), (40, 33, "plain"), (64, 64, "mixed_scale"),
                                       (64, 64, "large"), (64, 64, "tiny"), (64, 64, "zero"), (48, 64, "nonbinary_tail"),
                                       (64, 64, "nonbinary"), (64, 64, "nonfinite"), (64, 64, "subnormal_scale"), (49, 64, "plain"),
                                       (64, 64, "unaligned"), (56, 48, "unaligned_real"), (64, 64, "saturated"), (33, 33, "plain"),
                                       (63, 41, "plain"), (64, 64, "ragged_tail")])
@pytest.mark.parametrize("mode", [0, 3])
def test_reduced_gradient_integer_theta_large_batch(lib, cuda, n, nh, case, mode):
    """Batches of >= 16384 binary rows. mode 0 (automatic): on 33..64-unit layers BOTH contractions run on tcgen05 /
    TMEM (grad_reduce_umma.cu: theta from two 23-bit fixed-point groups, p digitised to 48 bits); mode 3: theta from 3 x 23-bit
    fixed-point images on mma.sync + FP64 tensor-core outer products (grad_reduce_dmma.cu, also mode 0's choice for 25..32-unit
    layers). Non-binary rows anywhere in the batch, non-finite weights or weights outside the images' range hand the whole
    batch to the FP64 kernel on the device. Aligned row tiles arrive by TMA, odd widths and misaligned views element by
    element. Same 1e-10 gates (max-norm and elementwise) in every case."""
    from qucumber_b200 import BinaryRBM, _lib

    _lib.check(lib.qcb_set_grad_reduce_mode(mode), "mode")
    try:
        _integer_theta_case(lib, cuda, n, nh, case)
    finally:
        _lib.check(lib.qcb_set_grad_reduce_mode(0), "mode")


def _integer_theta_case(lib, cuda, n, nh, case):
    from qucumber_b200 import BinaryRBM

    rows = 16384 + (1 if case == "ragged_tail" else 37)
    W, b, c = make_binary_params(n, nh, 301, bias_scale=0.3)
    v = random_states(rows, n, 302)
    rng = np.random.default_rng(303)
    if case == "saturated":   # hidden units pinned on and off: p = 1 and p ~ 1e-40 must digitise to all-ones / all-zero bytes
        c[::3] += 90.0
        c[1::3] -= 90.0
    if case == "mixed_scale":
        W = W * 10.0 ** rng.uniform(-9, 0.5, size=W.shape)
    elif case == "large":
        W = W * 37.0
    elif case == "tiny":
        W = W * 1e-12
    elif case == "zero":
        W = np.zeros_like(W)
    elif case == "nonbinary_tail":
        v[-1, 3] = 0.5
    elif case == "nonbinary":
        v = rng.standard_normal((rows, n))
    elif case == "nonfinite":
        W[5, 7] = np.inf
    elif case == "subnormal_scale":
        W = W * 1e-300
    rbm = load_binary(BinaryRBM(n, nh, gpu=True), W, b, c)
    o = BinaryRBMOracle(W, b, c)
    with np.errstate(all="ignore"):
        want = o.effective_energy_gradient(v)
    if case == "unaligned_real":
        v = rng.standard_normal((rows, n))
        with np.errstate(all="ignore"):
            want = o.effective_energy_gradient(v)
    if case.startswith("unaligned"):  # rows that start 8 bytes off a 16-byte boundary: no TMA row copies
        backing = torch.empty(rows * n + 1, dtype=torch.double, device=cuda)
        vt = backing[1:].view(rows, n)
        vt.copy_(torch.tensor(v))
        assert vt.data_ptr() % 16 == 8
    else:
        vt = torch.tensor(v, device=cuda)
    got = rbm.effective_energy_gradient(vt).cpu().numpy()
    if case == "nonfinite":
        ok = np.isfinite(want)  # everything outside hidden unit 5
        assert_close(got[ok], want[ok], what="finite entries next to a non-finite weight")
    else:
        assert_close(got, want, what=f"reduced gradient, integer theta ({case})")
    # per-entry check on the weight block as well (the bias sums dominate the global scale)
    if case in ("plain", "mixed_scale", "tiny"):
        gw, ww = got[: nh * n], want[: nh * n]
        assert float(np.max(np.abs(gw - ww) / (np.abs(ww) + 1e-6 * np.abs(ww).max()))) < 1e-9


def test_calls_are_capturable_in_a_cuda_graph(lib, cuda):
    """The C ABI only enqueues work on the caller's stream (no allocation, no synchronisation, no stream query that
    would invalidate a capture): a Gibbs call and a gradient call captured in a CUDA graph replay to the eager result."""
    from qucumber_b200 import BinaryRBM

    n, nh, C = 16, 12, 300
    W, b, c = make_binary_params(n, nh, 201)
    rbm = load_binary(BinaryRBM(n, nh, gpu=True), W, b, c)
    v0 = torch.tensor(random_states(C, n, 202), device=cuda)
    want_v = rbm.gibbs_steps(4, v0, seed=77)
    want_g = rbm.effective_energy_gradient(want_v)
    static_v = v0.clone()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        rbm.gibbs_steps(4, static_v.clone(), overwrite=True, seed=77)  # warm-up: workspaces, function attributes
        rbm.effective_energy_gradient(static_v)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            rbm.gibbs_steps(4, static_v, overwrite=True, seed=77)
            g_out = rbm.effective_energy_gradient(static_v)
    torch.cuda.current_stream().wait_stream(s)
    for _ in range(2):
        static_v.copy_(v0)
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(static_v, want_v)
        assert_close(g_out, want_g.cpu().numpy(), what="gradient from a replayed graph")


@pytest.mark.parametrize("binary", [True, False])
def test_large_batch_gradient_is_capturable(lib, cuda, binary):
    """The large-batch gradient is a memset of the flag word, the integer-theta kernel and the FP64 kernel behind it,
    selected on the device: the sequence captures into a CUDA graph and replays correctly for 0/1 rows (integer path)
    and for real rows (flag raised, FP64 kernel does the batch) with the SAME captured graph shape."""
    from qucumber_b200 import BinaryRBM

    n, nh, rows = 48, 40, 16384 + 5
    W, b, c = make_binary_params(n, nh, 211)
    rbm = load_binary(BinaryRBM(n, nh, gpu=True), W, b, c)
    o = BinaryRBMOracle(W, b, c)
    rng = np.random.default_rng(212)
    first = random_states(rows, n, 213) if binary else rng.standard_normal((rows, n))
    second = rng.standard_normal((rows, n)) if binary else random_states(rows, n, 214)  # replay with the other kind of rows
    static_v = torch.tensor(first, device=cuda)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        rbm.effective_energy_gradient(static_v)  # warm-up
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            g_out = rbm.effective_energy_gradient(static_v)
    torch.cuda.current_stream().wait_stream(s)
    for rows_np in (first, second, first):
        static_v.copy_(torch.tensor(rows_np))
        graph.replay()
        torch.cuda.synchronize()
        assert_close(g_out, o.effective_energy_gradient(rows_np), what="large-batch gradient from a replayed graph")


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_one_process_two_devices(lib, cuda):
    """A model living on cuda:1 while the current device stays cuda:0 runs on cuda:1 (the library switches to the
    device of the caller's stream or, for the legacy default stream, of the data) and gives cuda:0's results."""
    from qucumber_b200 import BinaryRBM

    torch.cuda.set_device(0)
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        W, b, c = make_binary_params(24, 24, 211)
        rbm = load_binary(BinaryRBM(24, 24, gpu=True), W, b, c).to(dev)
        rbm.device = torch.device(dev)
        v = torch.tensor(random_states(500, 24, 212), device=dev)
        outs.append([rbm.gibbs_steps(3, v, seed=5).cpu(), rbm.effective_energy(v).cpu(), rbm.effective_energy_gradient(v).cpu(),
                     rbm.log_partition().cpu()])
        assert torch.cuda.current_device() == 0
    for a, b_ in zip(*outs):
        assert torch.equal(a, b_)


@pytest.mark.parametrize("n,nh,scale", [(12, 9, 3.0), (40, 36, 1.5), (64, 64, 0.6), (30, 30, 40.0)])
def test_cd_epoch_large_weights(lib, cuda, n, nh, scale):
    """Strong couplings (pre-activations of tens to thousands): the epoch kernels' float filters must hand the
    undecided and the out-of-range units to the exact path, i.e. the chains stay those of the oracle."""
    from qucumber_b200 import _lib

    W, b, c = make_binary_params(n, nh, 500 + n, bias_scale=1.0, weight_scale=scale)
    pos, neg = random_states(120, n, 501), random_states(150, n, 502)
    k, seed = 3, 31
    o = BinaryRBMOracle(W, b, c)
    vk = o.gibbs_steps_philox(k, neg, seed=seed, chain_id0=0, sweep0=4)
    with np.errstate(over="ignore"):
        want = o.effective_energy_gradient(pos) / len(pos) - o.effective_energy_gradient(vk) / len(neg)
    tW, tb, tc = (torch.tensor(x, device=cuda) for x in (W, b, c))
    tp, tn = torch.tensor(pos, device=cuda), torch.tensor(neg, device=cuda)  # named: they must outlive the asynchronous call
    g = torch.empty(nh * n + n + nh, dtype=torch.double, device=cuda)
    _lib.check(lib.qcb_rbm_cd_epoch(_lib.ptr(tW), _lib.ptr(tb), _lib.ptr(tc), n, nh, _lib.ptr(tp), 120, 120, _lib.ptr(tn), 150, 150, 1, k,
                                    0.0, seed, 4, _lib.ptr(g), _lib.stream_ptr(cuda)), "qcb_rbm_cd_epoch")
    assert_close(g, want, what="CD-k gradient from the fused kernel (lr = 0)")
    assert np.array_equal(tW.cpu().numpy(), W) and np.array_equal(tb.cpu().numpy(), b) and np.array_equal(tc.cpu().numpy(), c)


def test_cd_step_with_zero_sweeps(lib, cuda):
    """k = 0: the negative phase is evaluated at the negative batch itself (v_k = v_0), also through the out-of-place
    path of qcb_rbm_cd_step / qcb_rbm_cdk_neg."""
    from qucumber_b200 import BinaryRBM, _lib

    n, nh, B, C = 11, 6, 40, 70
    W, b, c = make_binary_params(n, nh, 601)
    rbm = load_binary(BinaryRBM(n, nh, gpu=True), W, b, c)
    o = BinaryRBMOracle(W, b, c)
    pos, neg = random_states(B, n, 602), random_states(C, n, 603)
    want = o.effective_energy_gradient(pos) / B - o.effective_energy_gradient(neg) / C
    args, keep = rbm._param_args()
    tp, tn = torch.tensor(pos, device=cuda), torch.tensor(neg, device=cuda)
    g = torch.empty(rbm.num_pars, dtype=torch.double, device=cuda)
    ws = _lib.workspace(cuda, _lib.workspace_bytes(_lib.OP_CD_STEP, n, nh, 0, max(B, C)))
    ws.fill_(255)  # stale workspace contents must not leak into the result
    _lib.check(lib.qcb_rbm_cd_step(*args, _lib.ptr(tp), B, _lib.ptr(tn), C, 0, 1, 0, 0, 1.0 / B, -1.0 / C, _lib.ptr(g), _lib.ptr(ws),
                                   ws.numel(), _lib.stream_ptr(cuda)), "qcb_rbm_cd_step")
    assert_close(g, want, what="CD-0 gradient")


def test_bench_json_contract(lib, cuda):
    """`python bench.py` prints ONE JSON line on stdout (everything else goes to stderr) with the driver's keys: the metric
    and its config, device-timed value, roofline, CPU baseline, end-to-end figure with its copy sizes, launch count, clocks."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE")}
    run = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "3", "--warmup", "3", "--chains", "131072",
                          "--ref-chains", "16384", "--sustain-s", "0.2"], capture_output=True, text=True, timeout=900, env=env, cwd=root)
    assert run.returncode == 0, run.stderr[-3000:]
    lines = [ln for ln in run.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines[:3]
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in d, key
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3 and d["value"] > 1e8 and d["gpu_launches"] > 0
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and "workload" in d["config"] and d["vs_baseline"] is None
    r = d["roofline"]
    # frac = the issue-slot fraction (the real limiter); the int8 tensor-pipe view and the FP64-equivalent restatement sit beside it
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s", "TOP/s", "Gwarp-instr/s") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert 0 < r["frac"] < 1 and 0 < r["tensor_pipe"]["frac"] < 1 and r["fp64_equivalent"]["ratio"] > 0 and 0 < r["fp64_path"]["frac"] < 1
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] > 0 and d["cpu_baseline"]["cores"] >= 1
    e = d["e2e"]
    # packed rows: 8 bytes per chain each way (+ the parameters on the way in), copies inside the timed region
    assert e["value"] > 0 and e["h2d_bytes_per_step"] >= 131072 * 8 and e["d2h_bytes_per_step"] == 131072 * 8 and e["gpu_launches"] > 0
    assert e["value"] < d["value"]
    for sub in ("sample_api", "sample_api_float64_rows", "float64_rows_host_call"):
        assert e[sub]["value"] > 0 and e[sub]["d2h_bytes_per_step"] > 0, sub
    assert e["float64_rows_host_call"]["h2d_bytes_per_step"] >= 131072 * 64 * 8
    assert d["sustained"]["seconds"] >= 0.2 and d["sustained"]["value"] > 1e8
    t = d["torch_cuda_baseline"]
    assert t["kind"] == "reference" and t["gibbs_s64"]["chain_sweeps_per_s"] > 0 and t["tfim_epoch"]["ms_per_epoch"] > 0, t
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}


# ---- round 2: parity gaps named by the review ---------------------------------------------------------------------
@pytest.mark.parametrize("n,nh,k", [(64, 64, 2), (10, 10, 10)])
def test_gibbs_fed_uniforms_bit_exact_131072_chains(lib, cuda, n, nh, k):
    """fed-uniform Gibbs at C = 2^17 (every chain group of a full persistent grid and more)."""
    rbm, orc = _pair(n, nh, 221, cuda)
    chains = 1 << 17
    v0 = random_states(chains, n, 222)
    u = np.random.default_rng(223).random((k, chains, nh + n))
    got = rbm.gibbs_steps(k, torch.tensor(v0, device=cuda), uniforms=torch.tensor(u, device=cuda))
    assert np.array_equal(got.cpu().numpy(), orc.gibbs_steps_fed(k, v0, u))


def test_log_partition_n24_vs_chunked_oracle(lib, cuda):
    """N = nh = 24 (beyond the reference's max_size = 20, nn_states/neural_state.py:52-55): the oracle evaluates the
    reference's effective_energy chunk-wise with a streaming log-sum-exp (rbm/binary_rbm.py:92-96, 240-241)."""
    n = nh = 24
    rbm, orc = _pair(n, nh, 231, cuda)
    m, s = orc.logz_range(chunk=1 << 18)
    want = m + np.log(s)
    out2 = rbm.log_partition().cpu().numpy()
    got = out2[0] + np.log(out2[1])
    assert abs(got - want) <= 1e-10 * abs(want), (got, want)
    # a sub-range against the same oracle, and additivity of two ranges
    first, count = 123_457, 3_000_001
    ms, ss = orc.logz_range(first, count, chunk=1 << 18)
    o = rbm.log_partition(first, count).cpu().numpy()
    assert abs((o[0] + np.log(o[1])) - (ms + np.log(ss))) <= 1e-10 * abs(ms + np.log(ss))


@pytest.mark.parametrize("nh", [12, 20])
def test_log_partition_n32_vs_hidden_side_trace(lib, cuda, nh):
    """N = 32 (2^32 states, the benchmark's size) against the closed form that traces the VISIBLE units analytically and
    sums over the 2^nh hidden states — independent of the enumeration kernel. Weights small enough that no
    pre-activation reaches the softplus threshold (the closed form has none)."""
    n = 32
    W, b, c = make_binary_params(n, nh, 241, bias_scale=0.2)
    from qucumber_b200 import BinaryRBM

    rbm = load_binary(BinaryRBM(n, nh, gpu=True), W, b, c)
    assert np.abs(W).sum(1).max() + np.abs(c).max() < 20.0
    want = BinaryRBMOracle(W, b, c).log_partition_hidden_side()
    out2 = rbm.log_partition().cpu().numpy()
    got = out2[0] + np.log(out2[1])
    assert abs(got - want) <= 1e-10 * abs(want), (got, want)


@pytest.mark.parametrize("n,nh", [(10, 10), (4, 9), (16, 16), (1, 1), (13, 3), (3, 16), (16, 15)])
def test_gradient_reduce_small_layers_large_batch(lib, cuda, n, nh):
    """Layers of at most 16 x 16 units with >= 4096 explicit rows take the streaming kernel of grad_reduce_small.cu (the T4
    shape of SURVEY 8(d)): 1e-10 against the oracle for binary rows, real-valued rows, ragged last tiles (rows not a
    multiple of 512), an 8-byte-misaligned view, saturated units, and scale / accumulate through the C ABI."""
    from qucumber_b200 import BinaryRBM, _lib

    rbm, orc = _pair(n, nh, 301, cuda)
    for rows in (4096, 5003, 70_001):
        v = random_states(rows, n, 302 + rows)
        assert_close(rbm.effective_energy_gradient(torch.tensor(v, device=cuda)), orc.effective_energy_gradient(v),
                     what=f"small-layer gradient, {rows} binary rows")
    vr = np.random.default_rng(303).standard_normal((4999, n)) * 2.0
    assert_close(rbm.effective_energy_gradient(torch.tensor(vr, device=cuda)), orc.effective_energy_gradient(vr),
                 what="small-layer gradient, real rows")
    backing = torch.zeros(6001 * n + 1, dtype=torch.double, device=cuda)
    view = backing[1:].view(6001, n)
    v = random_states(6001, n, 304)
    view.copy_(torch.tensor(v, device=cuda))
    assert view.data_ptr() % 16 == 8
    assert_close(rbm.effective_energy_gradient(view), orc.effective_energy_gradient(v), what="small-layer gradient, misaligned rows")
    # saturated hidden units (|theta| in the hundreds) and the C ABI's scale / accumulate
    W, b, c = make_binary_params(n, nh, 305, bias_scale=0.5)
    W *= 60.0
    c[::2] += 300.0
    sat = load_binary(BinaryRBM(n, nh, gpu=True), W, b, c)
    v = random_states(8192, n, 306)
    want = BinaryRBMOracle(W, b, c).effective_energy_gradient(v)
    assert_close(sat.effective_energy_gradient(torch.tensor(v, device=cuda)), want, what="small-layer gradient, saturated units")
    args, keep = sat._param_args()
    g = torch.full((sat.num_pars,), 2.0, dtype=torch.double, device=cuda)
    tv = torch.tensor(v, device=cuda)
    ws = _lib.workspace(cuda, lib.qcb_workspace_bytes(_lib.OP_GRAD_REDUCE, n, nh, 0, 8192))
    _lib.check(lib.qcb_rbm_grad_reduce(*args, _lib.ptr(tv), 8192, -0.25, _lib.ptr(g), 1, _lib.ptr(ws), ws.numel(), _lib.stream_ptr(cuda)),
               "qcb_rbm_grad_reduce")
    assert_close(g, 2.0 - 0.25 * want, what="small-layer gradient, scale and accumulate")


@pytest.mark.parametrize("n,nh", [(10, 10), (12, 16), (5, 3), (8, 13)])
def test_gradient_reduce_histogram_path(lib, cuda, n, nh):
    """Round 2: small visible layers (N <= 12) with >= 16 * 2^N binary rows are counted (state histogram + one weighted pass
    over the 2^N states, grad_reduce_small.cu) instead of streamed. Same 1e-10 gate against the oracle on uniform rows, on
    concentrated distributions (every row equal; 90 % one state — the warp-aggregated counters), on a ragged row count, with
    one non-binary entry in the last row (device-side hand-over to the streaming kernel), and equal to the streaming kernel."""
    rbm, orc = _pair(n, nh, 1401, cuda)
    rows = (16 << n) + 4099
    rng = np.random.default_rng(1402)
    v = random_states(rows, n, 1403)
    assert_close(rbm.effective_energy_gradient(torch.tensor(v, device=cuda)), orc.effective_energy_gradient(v), what="histogram path, uniform rows")
    same = np.repeat(random_states(1, n, 1404), rows, axis=0)
    assert_close(rbm.effective_energy_gradient(torch.tensor(same, device=cuda)), orc.effective_energy_gradient(same), what="histogram path, one state")
    skew = np.where(rng.random((rows, 1)) < 0.9, same, v)
    want = orc.effective_energy_gradient(skew)
    got = rbm.effective_energy_gradient(torch.tensor(skew, device=cuda))
    assert_close(got, want, what="histogram path, 90 % one state")
    lib.qcb_set_grad_reduce_mode(3)
    try:
        streamed = rbm.effective_energy_gradient(torch.tensor(skew, device=cuda))
    finally:
        lib.qcb_set_grad_reduce_mode(0)
    assert_close(got, streamed.cpu().numpy(), what="histogram path == streaming kernel")
    odd = skew.copy()
    odd[-1, n - 1] = 0.5
    assert_close(rbm.effective_energy_gradient(torch.tensor(odd, device=cuda)), orc.effective_energy_gradient(odd),
                 what="histogram path, non-binary row handed to the streaming kernel")


def test_gradient_reduce_histogram_path_purification(lib, cuda):
    """The same for a PurificationRBM (stacked [W; U], gradient order [W, U, b, c, d])."""
    from helpers import load_purification, make_purification_params
    from oracle.rbm_oracle import PurificationRBMOracle
    from qucumber_b200 import PurificationRBM

    n, nh, na = 6, 5, 4
    W, U, b, c, d = make_purification_params(n, nh, na, 1410)
    rbm = load_purification(PurificationRBM(n, nh, na, gpu=True), W, U, b, c, d)
    v = random_states(5000, n, 1411)
    assert_close(rbm.effective_energy_gradient(torch.tensor(v, device=cuda)), PurificationRBMOracle(W, U, b, c, d).effective_energy_gradient(v),
                 what="histogram path, purification RBM")
