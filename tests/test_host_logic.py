"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/qcb.h
declares, the Python mirror's host logic, and the multi-rank combine logic on gloo (world_size 2)."""
import os
import re
import socket
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol(lib):
    from qucumber_b200 import _lib

    header = open(os.path.join(ROOT, "include", "qcb.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(qcb_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in qcb.h but not exported"
    assert lib.qcb_version() == 110
    assert lib.qcb_workspace_bytes(_lib.OP_GIBBS, 64, 64, 0, 1) > 2 * 64 * 64 * 8
    # layers beyond the fused kernels take the composed path: scratch for the hidden states and the probabilities of all chains
    assert lib.qcb_workspace_bytes(_lib.OP_GIBBS, 150, 150, 0, 1000) >= 2 * 1000 * 150 * 8


def test_no_cpu_fallback():
    from qucumber_b200 import BinaryRBM, PositiveWaveFunction, PurificationRBM

    rbm = BinaryRBM(4, 3, gpu=False)
    assert set(rbm.state_dict()) == {"weights", "visible_bias", "hidden_bias"} and rbm.num_pars == 4 * 3 + 4 + 3
    with pytest.raises(RuntimeError, match="no CPU path"):
        rbm.effective_energy(torch.zeros(2, 4, dtype=torch.double))
    with pytest.raises(RuntimeError, match="no CPU path"):
        rbm.gibbs_steps(1, torch.zeros(2, 4, dtype=torch.double))
    p = PurificationRBM(4, 3, 2, gpu=False)
    assert set(p.state_dict()) == {"weights_W", "weights_U", "visible_bias", "hidden_bias", "aux_bias"}
    assert p.num_pars == 4 * 3 + 4 * 2 + 4 + 3 + 2
    with pytest.raises(RuntimeError, match="no CPU path"):
        PositiveWaveFunction(4, gpu=False).sample(k=1, num_samples=2)


def test_epoch_fusion_sees_instance_level_batch_hooks():
    """a callback whose per-batch hooks are INSTANCE attributes (the reference's LambdaCallback,
    callbacks/lambda_callback.py:84-97) must keep the per-batch path, exactly like class-level hooks; hooks that do
    nothing (pass / docstring only / absent) must not."""
    from qucumber_b200 import PositiveWaveFunction
    from qucumber_b200.callbacks import CallbackBase, Timer
    from qucumber_b200.nn_states.neural_state import _Callbacks, _is_noop_hook

    nn = PositiveWaveFunction(4, 4, gpu=False)
    opt = torch.optim.SGD(list(nn.rbm_am.parameters()), lr=0.1)

    class LambdaLike(CallbackBase):  # hooks installed on the instance, class-level ones are the base no-ops
        def __init__(self, **hooks):
            for name, fn in hooks.items():
                setattr(self, name, fn)

    class Documented:
        def on_batch_end(self, nn_state, epoch, batch):
            """only a docstring"""

    class ClassLevel(CallbackBase):
        def on_batch_start(self, nn_state, epoch, batch):
            nn_state.seen = batch

    class CallableObject:
        def __call__(self, *a):
            return None

    calls = []
    fusable = lambda *cbs: nn._epoch_fusable(opt, _Callbacks(list(cbs)), None)  # noqa: E731
    assert fusable() and fusable(CallbackBase()) and fusable(Timer(verbose=False)) and fusable(Documented())
    assert fusable(LambdaLike(on_epoch_end=lambda s, e: calls.append(e)))
    assert not fusable(LambdaLike(on_batch_end=lambda s, e, b: calls.append(b)))
    assert not fusable(LambdaLike(on_batch_start=lambda s, e, b: setattr(s, "stop_training", True)))
    assert not fusable(ClassLevel())
    assert not fusable(LambdaLike(on_batch_end=CallableObject()))
    assert not fusable(LambdaLike(on_batch_end=print))
    assert _is_noop_hook(None) and _is_noop_hook(lambda *a: None) and not _is_noop_hook(lambda *a: calls.append(1))
    # the reference's own LambdaCallback, when the vendored copy is present
    sys.path.insert(0, ROOT)
    from oracle import make_ref

    if make_ref.available():
        make_ref.import_reference()
        from qucumber.callbacks import LambdaCallback

        assert not fusable(LambdaCallback(on_batch_end=lambda s, e, b: calls.append(b)))
        assert fusable(LambdaCallback(on_epoch_end=lambda s, e: calls.append(e)))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "qucumber_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f"{f} imports the oracle"
                assert "/root/reference" not in text


def test_host_helpers(tmp_path):
    from qucumber_b200 import HilbertSpace
    from qucumber_b200.utils import auto_unsqueeze_args, cplx, unitaries
    from qucumber_b200.utils.data import extract_refbasis_samples, load_data
    from qucumber_b200.utils.gradients_utils import vector_to_grads
    from oracle.rbm_oracle import default_unitaries, hilbert_rows

    assert np.array_equal(HilbertSpace(5).tensor().numpy(), hilbert_rows(5))
    assert np.array_equal(HilbertSpace(6).tensor(9, 11).numpy(), hilbert_rows(6, 9, 11))

    class M:
        @auto_unsqueeze_args()
        def f(self, v):
            assert v.dim() == 2
            return v * 2

    assert M().f(torch.ones(3)).shape == (3,) and M().f(torch.ones(2, 3)).shape == (2, 3)
    lin = torch.nn.Linear(3, 2)
    vec = torch.arange(8.0)
    vector_to_grads(vec, lin.parameters())
    assert torch.equal(lin.weight.grad, vec[:6].view(2, 3)) and torch.equal(lin.bias.grad, vec[6:])
    d = unitaries.create_dict()
    for name, u in default_unitaries().items():
        assert np.allclose(cplx.numpy(d[name]), u, atol=0, rtol=1e-15)
    codes, names = unitaries.encode_basis(list("XZYZ"), d)
    assert names[0] == "Z" and [names[c] for c in codes] == list("XZYZ")
    a = cplx.make_complex(torch.tensor([1.0, 2.0]), torch.tensor([3.0, -1.0]))
    b = cplx.make_complex(torch.tensor([0.5, -2.0]), torch.tensor([1.0, 4.0]))
    assert np.allclose(cplx.numpy(cplx.scalar_mult(a, b)), cplx.numpy(a) * cplx.numpy(b))
    assert np.allclose(cplx.numpy(cplx.inner_prod(a, b)), np.vdot(cplx.numpy(a), cplx.numpy(b)))
    assert np.allclose(cplx.numpy(cplx.inverse(a)), 1 / cplx.numpy(a))
    assert np.allclose(cplx.numpy(cplx.einsum("i,i->", a, b)), np.sum(cplx.numpy(a) * cplx.numpy(b)))
    # the full-space rotation is a CUDA kernel (qcb_kron_rotate; parity in tests/test_states_gpu.py): no CPU path
    rng = np.random.default_rng(0)
    psi = rng.standard_normal(8) + 1j * rng.standard_normal(8)

    class S:
        unitary_dict = d
        device = torch.device("cpu")

    with pytest.raises(RuntimeError, match="no CPU path"):
        unitaries.rotate_psi(S(), "XZY", None, psi=cplx.make_complex(psi))
    samples = tmp_path / "s.txt"
    samples.write_text("1 0\n0 1\n1 1\n")
    bases = tmp_path / "b.txt"
    bases.write_text("Z Z\nX Z\nZ Z\n")
    s, bs = load_data(str(samples), tr_bases_path=str(bases))
    assert s.dtype == torch.double and s.shape == (3, 2) and bs.shape == (3, 2)
    assert extract_refbasis_samples(s, bs).tolist() == [[1.0, 0.0], [1.0, 1.0]]


def test_shard_range_covers_everything():
    from qucumber_b200.distributed import shard_range

    for total in (0, 1, 7, 1 << 20, (1 << 32) + 5):
        for ws in (1, 2, 3, 8):
            pieces = [shard_range(total, r, ws) for r in range(ws)]
            assert pieces[0][0] == 0 and sum(c for _, c in pieces) == total
            for (f0, c0), (f1, _) in zip(pieces[:-1], pieces[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in pieces) - min(c for _, c in pieces) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world_size, port, out_dir):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import make_binary_params, random_states
    from oracle.rbm_oracle import BinaryRBMOracle
    from qucumber_b200 import distributed as qd

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world_size)
    n, nh = 9, 6
    W, b, c = make_binary_params(n, nh, 3)
    orc = BinaryRBMOracle(W, b, c)
    # exact enumeration: each rank's (m, s) over its index range, merged by all-gather
    first, count = qd.shard_range(1 << n)
    m, s = orc.logz_range(first, count)
    logz = qd.combine_logz(torch.tensor([m, s], dtype=torch.double))
    assert abs(float(logz) - orc.log_partition()) <= 1e-12 * abs(orc.log_partition())
    # CD gradient: partial sums scaled by global sizes + one allreduce == full-batch gradient
    pos = random_states(50, n, 4)
    f, cnt = qd.shard_range(50)
    part = torch.tensor(orc.effective_energy_gradient(pos[f:f + cnt]) / 50.0)
    extra = torch.tensor([float(rank + 1)], dtype=torch.double)
    qd.allreduce_sum_([part, extra])
    assert np.allclose(part.numpy(), orc.effective_energy_gradient(pos) / 50.0, rtol=1e-13, atol=0)
    assert float(extra) == sum(range(1, world_size + 1))
    # Philox draws keyed by global chain id: shards concatenate to the single-process run
    v0 = random_states(20, n, 5)
    f, cnt = qd.shard_range(20)
    mine = orc.gibbs_steps_philox(3, v0[f:f + cnt], seed=11, chain_id0=f)
    gathered = [None] * world_size
    dist.all_gather_object(gathered, mine)
    assert np.array_equal(np.concatenate(gathered), orc.gibbs_steps_philox(3, v0, seed=11))
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


def _fit_worker(rank, world_size, port, out_dir):
    """fit() with the batch sharded over two gloo ranks. The per-rank numbers come from the oracle (no GPU here): the
    state class below overrides the ONE numeric hook of the sharded path, so what is exercised is the host logic that
    ships — row / chain shards, global scales, global chain ids, seed broadcast, identical shuffles, all-reduce, update."""
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import make_binary_params, random_states
    from oracle.rbm_oracle import BinaryRBMOracle
    from qucumber_b200 import PositiveWaveFunction
    from qucumber_b200 import distributed as qd

    calls = []

    class OracleBacked(PositiveWaveFunction):
        def _cd_gradients_local(self, k, pos, neg, bases, pos_scale, neg_scale, chain_id0, seed, sweep0):
            rbm = self.rbm_am
            orc = BinaryRBMOracle(rbm.weights.numpy(), rbm.visible_bias.numpy(), rbm.hidden_bias.numpy())
            g = np.zeros(rbm.num_pars)
            if pos.shape[0]:
                g += pos_scale * orc.effective_energy_gradient(pos.numpy())
            if neg.shape[0]:
                vk = orc.gibbs_steps_philox(k, neg.numpy(), seed=seed, chain_id0=chain_id0, sweep0=sweep0)
                g += neg_scale * orc.effective_energy_gradient(vk)
            calls.append((pos.shape[0], neg.shape[0], chain_id0))
            return [torch.tensor(g)]

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world_size)
    n, nh = 6, 5
    W, b, c = make_binary_params(n, nh, 3)
    data = torch.tensor(random_states(53, n, 4))

    def run(sharded):
        torch.manual_seed(100 + rank)  # ranks seeded DIFFERENTLY on purpose: fit() must broadcast rank 0's key
        st = OracleBacked(n, nh, gpu=False)
        with torch.no_grad():
            st.rbm_am.weights.copy_(torch.tensor(W)); st.rbm_am.visible_bias.copy_(torch.tensor(b)); st.rbm_am.hidden_bias.copy_(torch.tensor(c))
        st.distributed = sharded
        st.fuse_epochs = False  # the one-launch-per-epoch kernel needs the GPU; the replica run goes batch by batch too
        del calls[:]
        torch.manual_seed(7 if rank == 0 else 8)
        st.fit(data, epochs=2, pos_batch_size=10, neg_batch_size=15, k=2, lr=0.1)
        return torch.cat([p.detach().reshape(-1) for p in st.rbm_am.parameters()]), list(calls)

    sharded, calls_s = run(True)
    # every rank took its shard of each batch (last batch: 3 positive rows -> 2 + 1), chains numbered globally
    assert [c[0] for c in calls_s[:6]] == ([5] * 5 + [2 if rank == 0 else 1])
    assert all(c[1] in (7, 8) for c in calls_s) and all(c[2] == (0 if rank == 0 else 8) for c in calls_s)
    gathered = [torch.empty_like(sharded) for _ in range(world_size)]
    dist.all_gather(gathered, sharded)
    assert torch.equal(gathered[0], gathered[1]), "ranks ended with different parameters"
    # the unsharded run: same key (run() reseeds torch's generator, rank 0's key is broadcast either way) and therefore the
    # same shuffles and draws, full batches on every rank -> the same parameters up to the order of the sums
    replica, calls_r = run(False)
    assert [c[0] for c in calls_r[:6]] == [10] * 5 + [3] and all(c[2] == 0 for c in calls_r)
    assert torch.allclose(sharded, replica, rtol=1e-12, atol=1e-14), float((sharded - replica).abs().max())
    assert not torch.equal(sharded, torch.cat([torch.tensor(x).reshape(-1) for x in (W, b, c)])), "fit() did not move the parameters"
    open(os.path.join(out_dir, f"fit_ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


def test_two_rank_sharded_fit_gloo(tmp_path):
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_fit_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "fit_ok0").exists() and (tmp_path / "fit_ok1").exists()


def test_two_rank_combine_logic_gloo(tmp_path):
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_complex_helper_algebra():
    """utils/cplx.py against numpy complex arithmetic on random inputs (what the reference's tests/test_cplx.py pins with
    literal tensors): construction, products, conjugation, division, norms, Kronecker products, the complex sigmoid,
    and the error cases for mismatched shapes."""
    from qucumber_b200.utils import cplx

    rng = np.random.default_rng(5)

    def rand(*shape):
        z = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
        return z, cplx.make_complex(torch.tensor(z.real), torch.tensor(z.imag))

    def same(t, z, tol=1e-13):
        got = cplx.numpy(t) if t.shape[0] == 2 and t.dim() == np.ndim(z) + 1 else t.numpy()
        assert got.shape == np.shape(z), (got.shape, np.shape(z))
        assert np.allclose(got, z, rtol=tol, atol=tol)

    s, S = rand()            # scalar
    v, V = rand(5)           # vector
    w, W = rand(5)
    m, M = rand(4, 5)        # matrices
    k, K = rand(5, 3)
    q, Q = rand(4, 5)
    assert torch.equal(cplx.make_complex(torch.tensor([1.0, 2.0])), torch.tensor([[1.0, 2.0], [0.0, 0.0]]))
    assert torch.equal(cplx.real(V), torch.tensor(v.real)) and torch.equal(cplx.imag(V), torch.tensor(v.imag))
    same(cplx.scalar_mult(S, V), s * v)
    same(cplx.scalar_mult(V, W), v * w)
    same(cplx.elementwise_mult(M, Q), m * q)
    same(cplx.elementwise_division(M, Q), m / q)
    same(cplx.scalar_divide(V, S), v / s)
    same(cplx.matmul(M, K), m @ k)
    same(cplx.matmul(M, V), m @ v)
    same(cplx.inner_prod(V, W), np.vdot(v, w))
    same(cplx.inner_prod(S, cplx.make_complex(torch.tensor(2.0), torch.tensor(-1.0))), np.conj(s) * (2 - 1j))
    same(cplx.outer_prod(V, W), np.outer(v, w))
    same(cplx.einsum("ij,jk->ik", M, K), m @ k)
    assert np.allclose(cplx.einsum("ij,jk->ik", M, K, imag_part=False).numpy(), (m @ k).real)
    assert np.allclose(cplx.einsum("ij,jk->ik", M, K, real_part=False).numpy(), (m @ k).imag)
    same(cplx.conjugate(M), m.conj().T)
    same(cplx.conjugate(V), v.conj())
    same(cplx.conj(M), m.conj())
    assert np.allclose(cplx.absolute_value(M).numpy(), np.abs(m))
    same(cplx.kronecker_prod(M, K), np.kron(m, k))
    same(cplx.sigmoid(torch.tensor(m.real), torch.tensor(m.imag)), 1.0 / (1.0 + np.exp(-m)))
    same(cplx.inverse(M), 1.0 / m)
    assert np.isclose(float(cplx.norm_sqr(V)), np.vdot(v, v).real) and np.isclose(float(cplx.norm(V)), np.linalg.norm(v))
    for bad in (lambda: cplx.elementwise_division(M, K), lambda: cplx.outer_prod(M, V), lambda: cplx.kronecker_prod(M, V),
                lambda: cplx.inner_prod(M, V), lambda: cplx.einsum("i,i->", V, W, real_part=False, imag_part=False)):
        with pytest.raises(ValueError):
            bad()


def test_every_golden_fixture_has_a_committed_generator():
    """Each file under tests/golden/ is written by one of the committed make_golden*.py scripts (which import the
    reference read-only in the build container) — no fixture of unknown origin."""
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    scripts = sorted(f for f in os.listdir(gold) if f.startswith("make_golden") and f.endswith(".py"))
    assert scripts, "generator scripts missing"
    source = "".join(open(os.path.join(gold, f)).read() for f in scripts)
    for name in sorted(os.listdir(gold)):
        if name in scripts or name.startswith("__"):
            continue
        stem, ext = os.path.splitext(name)
        assert ext in (".npz", ".pt"), name
        prefix = stem.rsplit("_", 1)[0] + "_"  # e.g. binary_rbm_{tag}.npz, ref_checkpoint_{tag}.pt
        assert name in source or (prefix in source and ext in source), f"{name}: no generator writes it"


def test_reference_arm_json_contract():
    """`bench.py --impl reference` runs without a GPU (it times the CPU restatement of the reference's operator sequence)
    and prints ONE JSON line with the contract's keys; ranks other than 0 exit 0 without output."""
    import json
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE")}
    run = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, env=env, cwd=root)
    assert run.returncode == 0, run.stderr[-2000:]
    lines = [ln for ln in run.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["steps"] == 1 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["value"] > 0
    assert "workload" in d["config"] and d["vs_baseline"] is None and d["higher_is_better"] is True
    from oracle import make_ref

    want_kind = "reference" if (make_ref.available() or os.path.isdir(make_ref.REF_SRC)) else "port"
    assert d["cpu_baseline"]["kind"] == want_kind and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    env.update(RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29871")
    other = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                           capture_output=True, text=True, timeout=120, env=env, cwd=root)
    assert other.returncode == 0 and other.stdout.strip() == ""


def test_rotated_shape_limits_raise_before_training():
    """shapes the fused rotated-basis kernels cannot take are refused with a clear ValueError at the start of fit(),
    not with QCB_ERR_UNSUPPORTED on the first rotated batch."""
    from qucumber_b200 import ComplexWaveFunction, DensityMatrix

    big = ComplexWaveFunction(600, 4, gpu=False)  # the site tables of the rotated-basis kernels hold 512 visible units
    z_only = np.array([["Z"] * 600] * 3)
    big._validate_rotated_shapes(z_only)  # reference-basis data needs no rotated kernel
    rotated = z_only.copy()
    rotated[1, 5] = "X"
    with pytest.raises(ValueError, match="num_visible <= 512"):
        big._validate_rotated_shapes(rotated)
    ok = ComplexWaveFunction(40, gpu=False)  # default num_hidden = 40 > 32: the row-tiled kernels take it
    ok._validate_rotated_shapes(rotated[:, :40])
    many = np.array([["X"] * 22 + ["Z"] * 18])
    with pytest.raises(ValueError, match="at most 21 non-Z sites"):
        ok._validate_rotated_shapes(many)
    dm = DensityMatrix(20, gpu=False)  # default num_aux = 20 > 16: the warp-per-row pair kernel takes it
    dm._validate_rotated_shapes(np.array([["X"] + ["Z"] * 19]))
    dm = DensityMatrix(4, 4, 300, gpu=False)
    with pytest.raises(ValueError, match="num_aux <= 256"):
        dm._validate_rotated_shapes(np.array([["X"] + ["Z"] * 3]))
    dm = DensityMatrix(12, 12, 12, gpu=False)
    with pytest.raises(ValueError, match="at most 10 non-Z sites"):
        dm._validate_rotated_shapes(np.array([["Y"] * 11 + ["Z"]]))


def test_full_space_fast_path_only_for_the_canonical_enumeration():
    """an explicit space tensor takes the index-generated path only when it IS the canonical enumeration."""
    from types import SimpleNamespace

    from qucumber_b200.hilbert import HilbertSpace
    from qucumber_b200.utils.training_statistics import _is_full_space

    st = SimpleNamespace(num_visible=4)
    canonical = HilbertSpace(4, "cpu").tensor()
    assert _is_full_space(st, None) and _is_full_space(st, HilbertSpace(4, "cpu")) and _is_full_space(st, canonical)
    assert not _is_full_space(st, canonical.flip(0))  # another row order
    dup = canonical.clone()
    dup[3] = dup[2]
    assert not _is_full_space(st, dup)  # repeated rows
    assert not _is_full_space(st, canonical[:8])


def test_checkpoint_loader_prefers_tensors_only(tmp_path):
    """checkpoints are read with the tensors-only unpickler; arbitrary-object metadata falls back with a warning."""
    import warnings

    from qucumber_b200.nn_states.neural_state import load_checkpoint

    with warnings.catch_warnings():
        warnings.simplefilter("error")
        for kind in ("positive", "complex", "density"):
            assert "rbm_am" in load_checkpoint(os.path.join(os.path.dirname(__file__), "golden", f"ref_checkpoint_{kind}.pt"))
    path = tmp_path / "odd.pt"
    torch.save({"rbm_am": {"weights": torch.zeros(2, 2)}, "note": SimpleNamespaceForPickle()}, path)
    with pytest.warns(RuntimeWarning, match="trusted files only"):
        assert "note" in load_checkpoint(path)


class SimpleNamespaceForPickle:
    """A module-level class: picklable, and not on torch's tensors-only allowlist."""


def test_mixed_state_fidelity_formula():
    """The Hermitian two-eigensolve form of the mixed-state fidelity against the reference's eigvals(sigma rho) recipe
    (training_statistics.py:63-76) on full-rank matrices (1e-12), and against <psi|rho|psi> for a pure target, where the
    reference's own round-off is ~1e-8."""
    from qucumber_b200.utils.training_statistics import _mixed_state_fidelity

    rng = np.random.default_rng(11)

    def density(dim, rank):
        a = rng.normal(size=(dim, rank)) + 1j * rng.normal(size=(dim, rank))
        m = a @ a.conj().T
        return m / np.trace(m).real

    def planes(m):
        return torch.tensor(np.stack([m.real, m.imag]))

    for dim in (2, 4, 16, 32):
        sigma, rho = density(dim, dim), density(dim, dim)
        ref = np.sum(np.sqrt(np.abs(np.linalg.eigvals(sigma @ rho).real))) ** 2
        assert abs(_mixed_state_fidelity(planes(sigma), planes(rho)) - ref) <= 1e-12 * ref
    psi = rng.normal(size=(16, 1)) + 1j * rng.normal(size=(16, 1))
    psi /= np.linalg.norm(psi)
    rho = density(16, 16)
    truth = (psi.conj().T @ rho @ psi).real.item()
    assert abs(_mixed_state_fidelity(planes(psi @ psi.conj().T), planes(rho)) - truth) <= 1e-13 * truth
    assert abs(_mixed_state_fidelity(planes(rho), planes(rho)) - 1.0) <= 1e-12


@pytest.mark.parametrize("n,groups,batch", [(500, 5, 100), (900, 9, 100), (230, 6, 50), (64, 3, 64), (7, 4, 3), (1, 1, 1)])
def test_epoch_grouping_matches_per_batch_grouping(n, groups, batch):
    """fit() over data with bases sorts the whole epoch by (batch, basis) once (qcb_fit_epoch_bases takes the offsets table): per
    batch that must be the stable grouping NeuralStateBase.gradient applies to a single batch (neural_state.py:326-369), group
    after group, including empty groups and a short last batch."""
    from qucumber_b200.nn_states.neural_state import _group_epoch_by_basis

    rng = np.random.default_rng(n + groups)
    ids = rng.integers(0, groups, size=n)
    if n > 10:
        ids[ids == groups - 1] = 0  # one basis never measured: an empty group in every batch
    nb = -(-n // batch)
    order, offsets, counts = _group_epoch_by_basis(ids, groups, batch, nb)
    assert sorted(order.tolist()) == list(range(n))
    assert offsets[0] == 0 and offsets[-1] == n and np.all(np.diff(offsets) == counts)
    for t in range(nb):
        rows = np.arange(t * batch, min((t + 1) * batch, n))
        want = rows[np.argsort(ids[rows], kind="stable")]
        first, last = offsets[t * groups], offsets[(t + 1) * groups]
        assert np.array_equal(order[first:last], want)
        assert np.array_equal(counts[t * groups:(t + 1) * groups], np.bincount(ids[rows], minlength=groups))
        for g in range(groups):
            seg = order[offsets[t * groups + g]:offsets[t * groups + g + 1]]
            assert np.all(ids[seg] == g) and np.all(np.diff(seg) > 0)


def test_ctypes_signatures_match_the_header():
    """Every prototype of include/qcb.h against the ctypes table of qucumber_b200/_lib.py: same number of parameters, and each
    parameter's C type maps to the ctypes type listed for it (pointers -> c_void_p, int -> c_int, int64_t -> c_int64, uint64_t ->
    c_uint64, size_t -> c_size_t, double -> c_double) — an argument added on one side only would shift every later one."""
    import ctypes

    from qucumber_b200 import _lib

    header = open(os.path.join(ROOT, "include", "qcb.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    header = re.sub(r"//[^\n]*", "", header)
    protos = re.findall(r"\b([A-Za-z_][A-Za-z0-9_ ]*?[\s\*]+)(qcb_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", header)
    assert len(protos) == len(_lib.SIGNATURES), (len(protos), len(_lib.SIGNATURES))

    def ctype_of(decl):
        decl = decl.strip()
        if "*" in decl:
            return ctypes.c_void_p
        words = [w for w in re.split(r"\s+", decl) if w not in ("const", "unsigned", "long")]
        base = decl if decl.startswith("unsigned long long") else words[0]
        table = {"int": ctypes.c_int, "int64_t": ctypes.c_int64, "uint64_t": ctypes.c_uint64, "size_t": ctypes.c_size_t,
                 "double": ctypes.c_double, "unsigned long long": ctypes.c_uint64}
        for key, val in table.items():
            if base.startswith(key) and (len(base) == len(key) or not (base[len(key)].isalnum() or base[len(key)] == "_")):
                return val
        raise AssertionError(f"unmapped C type in {decl!r}")

    for ret, name, params in protos:
        restype, argtypes = _lib.SIGNATURES[name]
        params = params.strip()
        plist = [] if params in ("", "void") else [p for p in params.split(",")]
        assert len(plist) == len(argtypes), f"{name}: {len(plist)} parameters in qcb.h, {len(argtypes)} in _lib.SIGNATURES"
        for i, (p, a) in enumerate(zip(plist, argtypes)):
            want = ctype_of(p)
            same = want is a or (ctypes.sizeof(want) == ctypes.sizeof(a) and {want, a} <= {ctypes.c_void_p, ctypes.c_char_p})
            assert same, f"{name}: parameter {i} is {p.strip()!r} in qcb.h but {a.__name__} in _lib.SIGNATURES"
