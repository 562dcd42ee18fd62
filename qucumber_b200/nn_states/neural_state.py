"""Host-side mirror of ``qucumber.nn_states.NeuralStateBase`` (reference:
qucumber/nn_states/neural_state.py:30-636) driving the CUDA RBMs.

The class keeps the reference's public surface (fit / sample / gradient / compute_batch_gradients /
compute_exact_gradients / normalization / save / load ...) and control flow (shuffling, optimizer,
callbacks), and routes every numeric step to libqucumber_b200.so:

* negative phase of CD-k: one fused call (k Gibbs sweeps + gradient reduction, qcb_rbm_cdk_neg);
* positive phase: qcb_rbm_grad_reduce (scaled by 1/batch in-kernel) or the fused rotated-basis
  kernels, accumulated straight into the gradient vectors;
* normalisation / exact gradients: index-generated enumeration (no Hilbert-space tensor).
"""
import abc
from itertools import chain
from math import ceil

import numpy as np
import torch

from .. import _lib
from ..utils import unitaries
from ..hilbert import HilbertSpace
from ..rng import next_seed
from ..utils import cplx
from ..utils.data import extract_refbasis_samples
from ..utils.gradients_utils import vector_to_grads


def load_checkpoint(location, map_location="cpu"):
    """torch.load of a QuCumber checkpoint. Tensors-only unpickling is tried first; a file whose metadata holds other Python
    objects (the reference's save() accepts any picklable metadata, neural_state.py:199-226) needs the full unpickler, which can
    run code from the file: that fallback warns, so only load such checkpoints from sources you trust."""
    import pickle
    import warnings

    try:
        return torch.load(location, map_location=map_location, weights_only=True)
    except pickle.UnpicklingError:
        warnings.warn(f"{location}: metadata is not tensors-only; unpickling arbitrary objects (trusted files only)",
                      RuntimeWarning, stacklevel=3)
        return torch.load(location, map_location=map_location, weights_only=False)


class _Callbacks:
    """Minimal dispatcher for the reference's callback protocol (on_train_start, on_epoch_start,
    on_batch_start, on_batch_end, on_epoch_end, on_train_end); any object implementing a subset of
    those hooks — e.g. the reference's own CallbackBase subclasses — can be passed to fit()."""

    HOOKS = ("on_train_start", "on_train_end", "on_epoch_start", "on_epoch_end", "on_batch_start", "on_batch_end")

    def __init__(self, callbacks):
        self.callbacks = list(callbacks or [])

    def fire(self, hook, *args):
        for cb in self.callbacks:
            fn = getattr(cb, hook, None)
            if fn is not None:
                fn(*args)


def _noop_plain(*args, **kwargs):
    pass


def _noop_documented(*args, **kwargs):
    """A hook that only carries a docstring."""


_NOOP_BYTECODE = {_noop_plain.__code__.co_code, _noop_documented.__code__.co_code, (lambda *a, **k: None).__code__.co_code}


def _is_noop_hook(fn):
    """True for a hook that is absent or a plain Python function whose body is `pass` / a docstring / `return None`
    (compared with the bytecode this interpreter emits for such bodies). Anything else — including any callable that is
    not a Python function — counts as doing work."""
    if fn is None:
        return True
    fn = getattr(fn, "__func__", fn)
    code = getattr(fn, "__code__", None)
    return code is not None and code.co_code in _NOOP_BYTECODE and not code.co_names


def stored_attribute(slot):
    """A settable property over a plain instance slot. The reference exposes ``rbm_am`` / ``rbm_ph`` / ``device`` as
    settable properties (e.g. complex_wavefunction.py:74-103) so that a module can be swapped in from outside —
    INTEGRATION.md section 1 relies on exactly that; the subclasses here declare them with this helper."""
    return property(lambda self: getattr(self, slot), lambda self, value: setattr(self, slot, value))


def _group_epoch_by_basis(ids, num_groups, batch_size, num_batches):
    """Rows of an epoch (``ids[i]``: basis group of the i-th row of the shuffled data set; batch t = rows t * batch_size ..)
    sorted by (batch, group), stable within a group — per batch the order ``np.argsort(ids_of_batch, kind="stable")`` gives, which
    is what NeuralStateBase.gradient does with one batch. Returns (order, offsets, counts): row ``order[p]`` goes to position
    p; group g of batch t owns positions offsets[t * num_groups + g] .. offsets[t * num_groups + g + 1]."""
    ids = np.asarray(ids, dtype=np.int64).reshape(-1)
    key = (np.arange(ids.shape[0], dtype=np.int64) // batch_size) * num_groups + ids
    order = np.argsort(key, kind="stable")
    counts = np.bincount(key, minlength=num_batches * num_groups)
    offsets = np.zeros(num_batches * num_groups + 1, dtype=np.int64)
    np.cumsum(counts, out=offsets[1:])
    return order, offsets, counts


class _IndexedRows:
    """Rows ``source[index]`` that have not been gathered yet: fit() hands the batches of a large data set to
    compute_batch_gradients in this form, so that a rank gathers only the rows of its own shard (under torch.distributed
    every rank would otherwise copy the whole shuffled data set every epoch) and no data-set sized temporary exists."""

    def __init__(self, source, index):
        self.source, self.index = source, index
        self.shape = (index.shape[0],) + tuple(source.shape[1:])

    def rows(self, first=0, count=None):
        idx = self.index if count is None else self.index[first:first + count]
        return self.source.index_select(0, idx)


class _GroupedBases:
    """The bases of one batch together with the group every row belongs to: fit() groups the whole data set by basis ONCE
    (np.unique over string rows costs more than the kernels of a two-qubit batch) and hands each batch its slice."""

    def __init__(self, strings, ids, unique):
        self.strings, self.ids, self.unique = strings, ids, unique

    def __len__(self):
        return len(self.strings)

    def __getitem__(self, item):  # row slices (the shards of compute_batch_gradients)
        return _GroupedBases(self.strings[item], self.ids[item], self.unique)

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.strings, dtype=dtype)


_SHUFFLE_STREAMS = {}


def _shuffle_stream(device):
    """The side stream fit() draws the next epoch's shuffle on (created once per device)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _SHUFFLE_STREAMS:
        _SHUFFLE_STREAMS[idx] = torch.cuda.Stream(device=device)
    return _SHUFFLE_STREAMS[idx]


class NeuralStateBase(abc.ABC):
    """Abstract base class for neural-network quantum states backed by CUDA RBMs."""

    _stop_training = False

    @property
    def stop_training(self):
        """If True, training stops at the end of the current batch/epoch (neural_state.py:33-50)."""
        return self._stop_training

    @stop_training.setter
    def stop_training(self, new_val):
        if isinstance(new_val, bool):
            self._stop_training = new_val
        else:
            raise ValueError("stop_training must be a boolean value!")

    @property
    def max_size(self):
        """Largest Hilbert space that generate_hilbert_space will materialise as a tensor."""
        return 20

    @property
    @abc.abstractmethod
    def networks(self):
        """Names of the internal RBMs."""

    @property
    @abc.abstractmethod
    def rbm_am(self):
        """The RBM that models the amplitude."""

    @property
    @abc.abstractmethod
    def device(self):
        """The device the model lives on."""

    def reinitialize_parameters(self):
        for net in self.networks:
            getattr(self, net).initialize_parameters()

    def __getattr__(self, attr):
        return getattr(self.rbm_am, attr)

    # ---- probabilities, sampling, enumeration -------------------------------------------------
    def probability(self, v, Z=1.0):
        """exp(-E(v)) / Z in the computational basis (neural_state.py:90-105)."""
        v = v.to(device=self.device, dtype=torch.double)
        return (-self.rbm_am.effective_energy(v)).exp() / Z

    def sample(self, k, num_samples=1, initial_state=None, overwrite=False, packed=False):
        """k block-Gibbs sweeps from a random (or given) state (neural_state.py:107-131).

        Without ``initial_state`` the reference draws ``Bernoulli(0.5)`` on the host and copies it over
        (neural_state.py:124-129); here the chains start from Bernoulli(1/2) bits drawn INSIDE the sampling kernel
        (Philox keyed from torch's generator, so ``torch.manual_seed`` still reproduces a run): one launch, nothing
        crosses PCIe. ``packed=True`` (not in the reference) returns ``(num_samples, ceil(N / 8))`` uint8 rows in
        numpy.packbits order instead of float64 rows — 64 x fewer bytes to bring to the host."""
        if initial_state is None:
            if hasattr(self.rbm_am, "sample_chains"):
                return self.rbm_am.sample_chains(k, num_samples, packed=packed)
            dist = torch.distributions.Bernoulli(probs=0.5)  # a foreign RBM module behind rbm_am: the reference's own route
            initial_state = dist.sample(torch.Size((num_samples, self.num_visible))).to(device=self.device, dtype=torch.double)
        out = self.rbm_am.gibbs_steps(k, initial_state, overwrite=overwrite)
        if packed:
            from ..rbm._sampling import pack_states

            return pack_states(out)
        return out

    def subspace_vector(self, num, size=None, device=None):
        """Basis state number ``num`` (MSB first) as a vector (neural_state.py:133-154)."""
        device = device if device is not None else self.device
        size = size if size else self.num_visible
        bits = [(num >> (size - 1 - j)) & 1 for j in range(size)]
        return torch.tensor(bits, dtype=torch.double, device=device)

    def generate_hilbert_space(self, size=None, device=None):
        """All 2^size basis states, row r = binary digits of r, column 0 most significant
        (neural_state.py:156-176). Raises ValueError above ``max_size`` like the reference; use
        :meth:`hilbert_space` for a lazy, size-unlimited handle the CUDA kernels understand."""
        device = device if device is not None else self.device
        size = size if size else self.rbm_am.num_visible
        if size > self.max_size:
            raise ValueError("Size of the Hilbert space is too large!")
        return HilbertSpace(size, device).tensor()

    def hilbert_space(self):
        """Lazy handle on the full visible space: accepted wherever a ``space`` tensor is."""
        return HilbertSpace(self.rbm_am.num_visible, self.device)

    #: "auto" splits a full-space enumeration over the ranks from this many visible units on (smaller spaces are launch-bound)
    shard_min_visible = 24

    def _shards_space(self):
        """(rank, world_size) to split a full-space enumeration over, or (0, 1)."""
        from .. import distributed as qd

        rank, world_size = qd.world()
        if world_size == 1 or self.distributed is False:
            return 0, 1
        if self.distributed is True or self.rbm_am.num_visible >= self.shard_min_visible:
            return rank, world_size
        return 0, 1

    def normalization(self, space=None):
        """Partition function / trace of the unnormalised state (neural_state.py:178-197). Over the full space
        (``space=None`` or a HilbertSpace handle) and under torch.distributed, each rank enumerates its index range and
        the (max, sum) pairs are merged after one all-gather."""
        rank, world_size = self._shards_space()
        if world_size > 1 and (space is None or isinstance(space, HilbertSpace)):
            from .. import distributed as qd

            first, count = qd.shard_range(1 << self.rbm_am.num_visible, rank, world_size)
            return qd.combine_logz(self.rbm_am.log_partition(first, count)).exp()
        return self.rbm_am.partition(space)

    def compute_normalization(self, space=None):
        return self.normalization(space)

    # ---- checkpointing (same file format as the reference) -------------------------------------
    def save(self, location, metadata=None):
        """torch.save of {net: state_dict, **metadata} (neural_state.py:199-226)."""
        metadata = dict(metadata) if metadata else {}
        if hasattr(self, "unitary_dict"):
            if "unitary_dict" in metadata:
                raise ValueError("Invalid key in metadata; unitary_dict cannot be a key!")
            metadata["unitary_dict"] = self.unitary_dict
        for net in self.networks:
            if net in metadata:
                raise ValueError(f"Invalid key in metadata; '{net}' cannot be a key!")
        data = {net: getattr(self, net).state_dict() for net in self.networks}
        data.update(**metadata)
        torch.save(data, location)

    def load(self, location):
        """Load parameters saved by this class or by stock QuCumber (neural_state.py:228-246)."""
        state_dict = load_checkpoint(location, map_location=self.device)
        for net in self.networks:
            getattr(self, net).load_state_dict(state_dict[net])
        if hasattr(self, "unitary_dict") and "unitary_dict" in state_dict:
            self.unitary_dict = state_dict["unitary_dict"]

    @abc.abstractmethod
    def importance_sampling_numerator(self, vp, v):
        """Numerator of the importance-sampling weight of vp w.r.t. v."""

    @abc.abstractmethod
    def importance_sampling_denominator(self, v):
        """Denominator of the importance-sampling weight."""

    def importance_sampling_weight(self, vp, v):
        return cplx.elementwise_division(self.importance_sampling_numerator(vp, v), self.importance_sampling_denominator(v))

    # ---- gradients ------------------------------------------------------------------------------
    def _accumulate_plain_gradient(self, rbm, samples, scale, g):
        """g += scale * sum_rows grad E(row) with one reduction kernel."""
        lib = _lib.load()
        args, keep = rbm._param_args()
        rows = samples.contiguous()
        na = getattr(rbm, "num_aux", 0)
        ws = _lib.workspace(rows.device, lib.qcb_workspace_bytes(_lib.OP_GRAD_REDUCE, rbm.num_visible, rbm.num_hidden, na, rows.shape[0]))
        _lib.check(lib.qcb_rbm_grad_reduce(*args, _lib.ptr(rows), rows.shape[0], float(scale), _lib.ptr(g), 1, _lib.ptr(ws),
                                           ws.numel(), _lib.stream_ptr(rows.device)), "qcb_rbm_grad_reduce")

    def _accumulate_rotated_gradient(self, basis, samples, g_am, g_ph):
        """Subclasses with measurement bases accumulate the rotated-basis gradient of ``samples``."""
        raise NotImplementedError

    #: "cwf" / "dm": the grouped C entry point that serves a whole batch of this state type (None: one call per basis)
    _grouped_call = None

    #: shape limits of the fused rotated-basis kernels: (attribute of the state, largest supported value, what it is)
    _rotated_limits = ()
    _max_rotated_sites = 21

    def _validate_rotated_shapes(self, bases):
        """Raises ValueError when rotated-basis data cannot be handled by the fused kernels of this state type (the reference
        has no such limits; they are listed in README.md)."""
        rotated = int(np.max(np.sum(np.asarray(bases) != "Z", axis=-1))) if len(bases) else 0
        if rotated == 0:
            return
        for attr, limit, what in self._rotated_limits:
            if getattr(self, attr) > limit:
                raise ValueError(f"{type(self).__name__}: rotated-basis data needs {what} <= {limit} in qucumber_b200 "
                                 f"(this state has {getattr(self, attr)})")
        if rotated > self._max_rotated_sites:
            raise ValueError(f"{type(self).__name__}: at most {self._max_rotated_sites} non-Z sites per basis are supported "
                             f"(the data has a basis with {rotated})")

    def gradient(self, samples, bases=None):
        """Summed gradients of a batch measured in the given bases (neural_state.py:326-369)."""
        samples = samples.to(device=self.device, dtype=torch.double)
        grad = [torch.zeros(getattr(self, net).num_pars, dtype=torch.double, device=self.device) for net in self.networks]
        if bases is None:
            s2 = samples.unsqueeze(0) if samples.dim() < 2 else samples
            self._accumulate_plain_gradient(self.rbm_am, s2, 1.0, grad[0])
            return grad
        if samples.dim() < 2:
            samples = samples.unsqueeze(0)
            bases = np.array(list(bases)).reshape(1, -1)
        # group the rows by basis on the host, gather them ONCE on the device (a boolean-mask index per basis would
        # synchronise the stream every time): after the gather each basis owns a contiguous slice
        if isinstance(bases, _GroupedBases):
            unique_bases, indices = bases.unique, bases.ids
        else:
            unique_bases, indices = np.unique(bases, axis=0, return_inverse=True)
        indices = np.asarray(indices).reshape(-1)
        order = np.argsort(indices, kind="stable")
        counts = np.bincount(indices, minlength=unique_bases.shape[0])
        grouped = samples.index_select(0, torch.as_tensor(order, device=samples.device))
        if self._grouped_call is not None and grouped.is_cuda:
            # every basis of the batch in ONE C call (the per-basis loop, plain all-Z groups included, runs in C)
            unitaries._call_groups(self._grouped_call, self.rbm_am, self.rbm_ph, unique_bases, counts, grouped, self.unitary_dict,
                                   grad[0], grad[1])
            return grad
        start = 0
        for i in range(unique_bases.shape[0]):
            basis = unique_bases[i, :]
            sub = grouped[start:start + int(counts[i]), :]
            start += int(counts[i])
            if np.any(basis != "Z"):
                self._accumulate_rotated_gradient(basis, sub, grad[0], grad[1])
            else:
                self._accumulate_plain_gradient(self.rbm_am, sub, 1.0, grad[0])
        return grad

    def positive_phase_gradients(self, samples_batch, bases_batch=None):
        """Data-driven part of the gradient, averaged over the batch (neural_state.py:371-384)."""
        grad = self.gradient(samples_batch, bases=bases_batch)
        return [gr / float(samples_batch.shape[0]) for gr in grad]

    def compute_exact_gradients(self, samples_batch, space=None, bases_batch=None):
        """Positive phase minus the exact negative phase sum_sigma p(sigma) grad E(sigma)
        (neural_state.py:386-415); the sum runs over index-generated states on the GPU."""
        grad = self.positive_phase_gradients(samples_batch, bases_batch=bases_batch)
        rbm = self.rbm_am
        if space is not None and not isinstance(space, HilbertSpace) and space.shape[0] != (1 << rbm.num_visible):
            probs = self.probability(space, Z=1.0)
            probs = probs / probs.sum()
            all_grads = rbm.effective_energy_gradient(space, reduce=False)
            grad[0] -= torch.mv(all_grads.t(), probs)
            return grad
        from .. import distributed as qd

        lib = _lib.load()
        args, keep = rbm._param_args()
        # the index range splits over the ranks like the partition function; with log Z known the partial
        # sums of p(r) grad E(r) are plain sums (one all-reduce of num_pars doubles)
        rank, world_size = self._shards_space()
        first, count = qd.shard_range(1 << rbm.num_visible, rank, world_size)
        out2 = rbm.log_partition(first, count)
        logZ = float(qd.combine_logz(out2) if world_size > 1 else out2[0] + out2[1].log())
        na = getattr(rbm, "num_aux", 0)
        neg = torch.zeros(rbm.num_pars, dtype=torch.double, device=self.device)
        out8 = torch.empty(8, dtype=torch.double, device=self.device)
        ws = _lib.workspace(self.device, lib.qcb_workspace_bytes(_lib.OP_EXACT_STATS, rbm.num_visible, rbm.num_hidden, na, count))
        _lib.check(lib.qcb_rbm_exact_stats_range(*args, first, count, logZ, None, None, _lib.ptr(out8), _lib.ptr(neg), _lib.ptr(ws),
                                                 ws.numel(), _lib.stream_ptr(self.device)), "qcb_rbm_exact_stats_range")
        if world_size > 1:
            qd.allreduce_sum_(neg)
        grad[0] -= neg
        return grad

    _cd_seed = None   # Philox key of the current fit() call (None: draw a fresh key per call)
    _cd_sweep = 0     # sweeps consumed under that key

    def _next_cd_stream(self, k):
        """(seed, sweep0) for the next k Gibbs sweeps: inside fit() one key serves the whole run and the
        sweep counter advances, so no per-step host RNG call is needed; outside, a fresh key per call."""
        if self._cd_seed is None:
            return next_seed(), 0
        sweep0 = self._cd_sweep
        self._cd_sweep += int(k)
        if self._cd_sweep >= (1 << 31):
            self._cd_seed, self._cd_sweep = next_seed(), 0
        return self._cd_seed, sweep0

    #: "auto": under torch.distributed (world_size > 1) large batches are sharded over the ranks and the partial gradient
    #: sums all-reduced; small ones run as replicas. True / False force one or the other.
    distributed = "auto"
    #: "auto" shards a batch of at least this many rows per rank (rotated-basis batches: 1/64 of it, a row there costs 2^S rows)
    shard_min_rows_per_rank = 4096

    def _shards_batches(self, pos_rows, neg_rows, with_bases):
        """(rank, world_size) to shard this batch over, or (0, 1) to run it on every rank as a replica."""
        from .. import distributed as qd

        rank, world_size = qd.world()
        if world_size == 1 or self.distributed is False:
            return 0, 1
        if self.distributed is True:
            return rank, world_size
        floor = self.shard_min_rows_per_rank // (64 if with_bases else 1)
        return (rank, world_size) if max(pos_rows, neg_rows) >= floor * world_size else (0, 1)

    def _cd_gradients_local(self, k, pos, neg, bases, pos_scale, neg_scale, chain_id0, seed, sweep0):
        """pos_scale * sum_rows grad E(pos rows [in their bases]) + neg_scale * sum_chains grad E(v_k) for THESE rows and
        chains — a partial sum when the batch is sharded. Chains are numbered chain_id0, chain_id0 + 1, ... (the Philox
        key of a chain is its global number, so shards draw what the unsharded batch would). Without bases the whole
        thing is ONE call (qcb_rbm_cd_step: gradient reduction, k fused sweeps, gradient reduction); v_k never leaves
        the device workspace and ``neg`` is not modified."""
        rbm = self.rbm_am
        lib = _lib.load()
        na = getattr(rbm, "num_aux", 0)
        C = neg.shape[0]
        args, keep = rbm._param_args()
        if bases is None:
            B = pos.shape[0]
            g = torch.empty(rbm.num_pars, dtype=torch.double, device=self.device)
            ws = _lib.workspace(self.device, _lib.workspace_bytes(_lib.OP_CD_STEP, rbm.num_visible, rbm.num_hidden, na, max(B, C, 1)))
            _lib.check(lib.qcb_rbm_cd_step(*args, _lib.ptr(pos), B, _lib.ptr(neg), C, int(k), seed, int(chain_id0), sweep0,
                                           float(pos_scale), float(neg_scale), _lib.ptr(g), _lib.ptr(ws), ws.numel(),
                                           _lib.stream_ptr(self.device)), "qcb_rbm_cd_step")
            grad = [g]
            for net in self.networks[1:]:
                grad.append(torch.zeros(getattr(self, net).num_pars, dtype=torch.double, device=self.device))
            return grad
        grad = self.gradient(pos, bases=bases) if pos.shape[0] else [
            torch.zeros(getattr(self, net).num_pars, dtype=torch.double, device=self.device) for net in self.networks]
        for gr in grad:
            gr.mul_(float(pos_scale))
        if C:
            vk = neg.clone()
            ws = _lib.workspace(self.device, _lib.workspace_bytes(_lib.OP_CDK_NEG, rbm.num_visible, rbm.num_hidden, na, C))
            _lib.check(lib.qcb_rbm_cdk_neg(*args, _lib.ptr(vk), C, int(k), seed, int(chain_id0), sweep0, None, float(neg_scale),
                                           _lib.ptr(grad[0]), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(self.device)), "qcb_rbm_cdk_neg")
        return grad

    #: rows per phase from which qcb_rbm_cd_step_packed applies (the tcgen05 gradient kernel's threshold)
    packed_min_rows = 16384

    def _packed_training_shape(self):
        """Whether fit() may keep the data as packed rows: a BinaryRBM with both layers of 33..64 units."""
        rbm = self.rbm_am
        return (len(self.networks) == 1 and hasattr(rbm, "weights") and getattr(rbm, "num_aux", 0) == 0
                and 33 <= rbm.num_visible <= 64 and 33 <= rbm.num_hidden <= 64)

    def _cd_gradients_packed(self, k, pos_bits, neg_bits, bases, pos_scale, neg_scale, chain_id0, seed, sweep0):
        """_cd_gradients_local for batches that arrive as PACKED rows (uint8, ceil(N / 8) bytes per row): ONE call,
        qcb_rbm_cd_step_packed — both reductions and the sweeps read and write bits. Phases of fewer than ``packed_min_rows``
        rows (a short last batch) are unpacked and take the float64 call."""
        from ..rbm._sampling import unpack_states

        rbm = self.rbm_am
        B, C = pos_bits.shape[0], neg_bits.shape[0]
        if min(B, C) < self.packed_min_rows or int(k) < 1:
            pos = unpack_states(pos_bits, rbm.num_visible)
            neg = pos if neg_bits is pos_bits else unpack_states(neg_bits, rbm.num_visible)
            return self._cd_gradients_local(k, pos, neg, None, pos_scale, neg_scale, chain_id0, seed, sweep0)
        lib = _lib.load()
        args, keep = rbm._param_args()
        g = torch.empty(rbm.num_pars, dtype=torch.double, device=self.device)
        ws = _lib.workspace(self.device, _lib.workspace_bytes(_lib.OP_CD_STEP, rbm.num_visible, rbm.num_hidden, 0, max(B, C)))
        _lib.check(lib.qcb_rbm_cd_step_packed(*args, _lib.ptr(pos_bits), B, _lib.ptr(neg_bits), C, int(k), seed, int(chain_id0), sweep0,
                                              float(pos_scale), float(neg_scale), _lib.ptr(g), _lib.ptr(ws), ws.numel(),
                                              _lib.stream_ptr(self.device)), "qcb_rbm_cd_step_packed")
        return [g]

    def compute_batch_gradients(self, k, samples_batch, neg_batch, bases_batch=None):
        """CD-k gradient of one batch (neural_state.py:417-447): positive phase from the data, negative phase from k
        Gibbs sweeps started at ``neg_batch``. Under torch.distributed a large batch is sharded: each rank reduces its
        rows and chains (scaled by the GLOBAL batch sizes), one sum all-reduce of the gradient vectors follows, issued
        asynchronously and waited on right before the vectors are handed back for the update; every rank returns the
        same gradients (the same as a single process up to summation order)."""
        from .. import distributed as qd

        rbm = self.rbm_am
        lib = _lib.load()
        na = getattr(rbm, "num_aux", 0)
        seed, sweep0 = self._next_cd_stream(k)
        if isinstance(samples_batch, _IndexedRows):  # batches of fit() that are still (data, index) pairs: shard first, gather after
            B, C = samples_batch.shape[0], neg_batch.shape[0]
            rank, world_size = self._shards_batches(B, C, False)
            # equal batch sizes reuse the positive permutation for the chains' starting rows (neural_state.py:449-498): one gather
            # serves both (the CD step does not modify its negative batch)
            same = (B == C and neg_batch.source is samples_batch.source and neg_batch.index.data_ptr() == samples_batch.index.data_ptr())
            packed = samples_batch.source.dtype == torch.uint8   # fit() keeps large binary data sets as packed rows
            if world_size > 1 or packed:
                pf, pc = qd.shard_range(B, rank, world_size) if world_size > 1 else (0, B)
                nf, nc = qd.shard_range(C, rank, world_size) if world_size > 1 else (0, C)
                pos_rows = samples_batch.rows(pf, pc)
                neg_rows = pos_rows if same else neg_batch.rows(nf, nc)
                local = self._cd_gradients_packed if packed else self._cd_gradients_local
                grad = local(k, pos_rows, neg_rows, None, 1.0 / float(B), -1.0 / float(C), nf, seed, sweep0)
                if world_size > 1:
                    qd.allreduce_sum_(grad)
                return grad
            samples_batch = samples_batch.rows()
            neg_batch = samples_batch if same else neg_batch.rows()
        neg = neg_batch.to(device=self.device, dtype=torch.double).contiguous()
        pos = samples_batch.to(device=self.device, dtype=torch.double).contiguous()
        B, C = pos.shape[0], neg.shape[0]
        rank, world_size = self._shards_batches(B, C, bases_batch is not None)
        if world_size > 1:
            pf, pc = qd.shard_range(B, rank, world_size)
            nf, nc = qd.shard_range(C, rank, world_size)
            bases = None if bases_batch is None else (bases_batch[pf:pf + pc] if isinstance(bases_batch, _GroupedBases)
                                                      else np.asarray(bases_batch)[pf:pf + pc])
            grad = self._cd_gradients_local(k, pos[pf:pf + pc], neg[nf:nf + nc], bases, 1.0 / float(B), -1.0 / float(C), nf, seed, sweep0)
            qd.allreduce_sum_(grad)
            return grad
        if (bases_batch is None and self.fuse_epochs and na == 0 and rbm.num_visible <= 64 and rbm.num_hidden <= 64 and 0 < B <= 4096
                and 0 < C <= 4096 and hasattr(rbm, "weights")):
            # small RBM, small batch: positive phase, k sweeps and negative phase in ONE cluster launch
            # (qcb_rbm_cd_epoch with a single batch and lr = 0: parameters untouched, gradient out)
            g = torch.empty(rbm.num_pars, dtype=torch.double, device=self.device)
            _lib.check(lib.qcb_rbm_cd_epoch(_lib.ptr(rbm.weights.data), _lib.ptr(rbm.visible_bias.data),
                                            _lib.ptr(rbm.hidden_bias.data), rbm.num_visible, rbm.num_hidden, _lib.ptr(pos), B, B,
                                            _lib.ptr(neg), C, C, 1, int(k), 0.0, seed, sweep0, _lib.ptr(g),
                                            _lib.stream_ptr(self.device)), "qcb_rbm_cd_epoch")
            grad = [g]
            for net in self.networks[1:]:
                grad.append(torch.zeros(getattr(self, net).num_pars, dtype=torch.double, device=self.device))
            return grad
        return self._cd_gradients_local(k, pos, neg, bases_batch, 1.0 / float(B), -1.0 / float(C), 0, seed, sweep0)

    # ---- training loop --------------------------------------------------------------------------
    def _shuffle_data(self, pos_batch_size, neg_batch_size, num_batches, train_samples, input_bases, z_samples):
        """Per-epoch batching (neural_state.py:449-498): positive batches are a permutation of the
        data; negative batches reuse it when the sizes match, otherwise they are drawn with
        replacement (from the all-Z samples when bases are given)."""
        gen = self._shuffle_generator
        if (input_bases is None and gen is not None and gen.device.type == "cuda" and train_samples.is_cuda
                and train_samples.shape[0] >= self.lazy_batches_min_rows):
            # large data: only the permutations are drawn now, the rows are gathered batch by batch (and shard by shard)
            n = train_samples.shape[0]
            pos_perm = torch.randperm(n, generator=gen, device=train_samples.device)
            neg_perm = pos_perm if neg_batch_size == pos_batch_size else torch.randint(
                n, size=(num_batches * neg_batch_size,), dtype=torch.long, generator=gen, device=train_samples.device)
            rows = self._train_bits if self._train_bits is not None else train_samples
            pos_b = [_IndexedRows(rows, pos_perm[s:s + pos_batch_size]) for s in range(0, n, pos_batch_size)]
            neg_b = [_IndexedRows(rows, neg_perm[s:s + neg_batch_size]) for s in range(0, neg_perm.shape[0], neg_batch_size)]
            return zip(pos_b, neg_b)
        pos, neg, pos_perm = self._shuffle_tensors(pos_batch_size, neg_batch_size, num_batches, train_samples, input_bases, z_samples,
                                                   generator=self._shuffle_generator)
        pos_batches = [pos[s:s + pos_batch_size] for s in range(0, len(pos), pos_batch_size)]
        neg_batches = [neg[s:s + neg_batch_size] for s in range(0, len(neg), neg_batch_size)]
        if input_bases is not None:
            pos_bases = input_bases[pos_perm]
            if self._basis_groups is not None:
                uniq, ids_all = self._basis_groups
                pos_ids = ids_all[np.asarray(pos_perm)]
                bases_batches = [_GroupedBases(pos_bases[s:s + pos_batch_size], pos_ids[s:s + pos_batch_size], uniq)
                                 for s in range(0, len(train_samples), pos_batch_size)]
            else:
                bases_batches = [pos_bases[s:s + pos_batch_size] for s in range(0, len(train_samples), pos_batch_size)]
            return zip(pos_batches, neg_batches, bases_batches)
        return zip(pos_batches, neg_batches)

    @staticmethod
    def _shuffle_tensors(pos_batch_size, neg_batch_size, num_batches, train_samples, input_bases, z_samples, generator=None):
        """The epoch's positive rows, negative rows (batches back to back) and the permutation used. ``generator``: the
        host generator to draw from (None: torch's default; under torch.distributed one seeded identically on all ranks)."""
        where = generator.device if generator is not None else torch.device("cpu")  # a device generator shuffles on the device
        pos_perm = torch.randperm(train_samples.shape[0], generator=generator, device=where)
        pos = train_samples[pos_perm]
        if input_bases is None:
            if neg_batch_size == pos_batch_size:
                neg_perm = pos_perm
            else:
                neg_perm = torch.randint(train_samples.shape[0], size=(num_batches * neg_batch_size,), dtype=torch.long,
                                         generator=generator, device=where)
            neg = train_samples[neg_perm]
        else:
            neg_perm = torch.randint(z_samples.shape[0], size=(num_batches * neg_batch_size,), dtype=torch.long, generator=generator,
                                     device=where)
            neg = z_samples[neg_perm]
        return pos, neg, pos_perm

    _shuffle_generator = None  # set by fit() under torch.distributed: every rank must draw the same batches
    _train_bits = None         # set by fit(): the packed copy of a large binary data set (see _cd_gradients_packed)
    _basis_groups = None       # set by fit(): (unique bases, group id of every sample) of the data set
    _basis_codes = None        # set by fit(): (codes of the unique bases, names of the unitaries they index)
    _index_staging = None      # set by _fit_epoch_bases(): two pinned staging slots (indices, event of the last use, descriptors)
    _desc_dev = None           # set by _fit_epoch_bases(): device copy of the per-group descriptors
    #: set to False to keep fit() on float64 rows whatever the data
    pack_training_data = True
    # data without bases is shuffled on the device: the permutation, the gathers and the epoch's kernels are then all queued
    # asynchronously (an index tensor drawn on the host has to be copied over first, which makes the host wait for the previous
    # epoch's kernel: 1.30 -> 1.1 ms per TFIM epoch); set it to a row count to keep host shuffles below that size
    device_shuffle_min_rows = 0
    #: from this many samples on, fit() hands batches to compute_batch_gradients as (data, index) pairs (see _IndexedRows)
    lazy_batches_min_rows = 1 << 16

    #: set to False to force one C call per batch even where a whole epoch fits one launch
    fuse_epochs = True

    def _epoch_fusable(self, optimizer, cbs, input_bases):
        """True when the batch loop of an epoch can run as ONE kernel (qcb_rbm_cd_epoch): a single small
        BinaryRBM, no bases, vanilla SGD, and no callback that acts per batch."""
        from ..rbm import BinaryRBM

        pg = optimizer.param_groups
        rbm = getattr(self, "rbm_am", None)
        if not (self.fuse_epochs and input_bases is None and list(self.networks) == ["rbm_am"] and type(rbm) is BinaryRBM
                and rbm.num_visible <= 64 and rbm.num_hidden <= 64):
            return False
        if not (type(optimizer) is torch.optim.SGD and len(pg) == 1 and pg[0].get("momentum", 0) == 0
                and pg[0].get("weight_decay", 0) == 0 and pg[0].get("dampening", 0) == 0 and not pg[0].get("nesterov", False)
                and not pg[0].get("maximize", False)):
            return False

        # hooks are resolved exactly as fire() resolves them (instance attributes included: the reference's
        # LambdaCallback installs its hooks on the instance, callbacks/lambda_callback.py:84-97)
        return all(_is_noop_hook(getattr(cb, hook, None)) for cb in cbs.callbacks for hook in ("on_batch_start", "on_batch_end"))

    def _bases_epoch_fusable(self, plain_sgd, cbs):
        """True when the batch loop of an epoch over data WITH bases can run as one C call (qcb_fit_epoch_bases): a state type
        with a grouped entry point on a CUDA device, vanilla SGD, and no callback that acts per batch."""
        if not (self.fuse_epochs and plain_sgd and self._grouped_call is not None and self.device.type == "cuda"
                and list(self.networks) == ["rbm_am", "rbm_ph"]):
            return False
        if not all(prm.is_contiguous() and prm.is_cuda for net in self.networks for prm in getattr(self, net).parameters()):
            return False
        return all(_is_noop_hook(getattr(cb, hook, None)) for cb in cbs.callbacks for hook in ("on_batch_start", "on_batch_end"))

    def _fit_epoch_bases(self, train_samples, z_samples, pos_batch_size, neg_batch_size, num_batches, k, lr, want_grads):
        """All batches of one epoch over data with bases in one call; parameters are updated in place. Draws the permutations
        exactly as _shuffle_tensors does and hands every batch over grouped by basis (stable within a group), so the
        batches, their row order and the chains' Philox streams are those of the per-batch loop."""
        lib = _lib.load()
        dev = self.device
        am, ph = self.rbm_am, self.rbm_ph
        uniq, ids_all = self._basis_groups
        n, N = train_samples.shape
        G = int(uniq.shape[0])
        gen = self._shuffle_generator
        pos_perm = torch.randperm(n, generator=gen).numpy()
        neg_perm = torch.randint(z_samples.shape[0], size=(num_batches * neg_batch_size,), dtype=torch.long, generator=gen)
        order, offsets, counts = _group_epoch_by_basis(ids_all[pos_perm], G, int(pos_batch_size), int(num_batches))
        # both index vectors go over in ONE asynchronous copy from a pinned staging buffer (a pageable copy would make the host
        # wait for the previous epoch's kernels before it could prepare this one); two buffers alternate, each guarded by an event
        # (the same pair of staging slots carries the per-group descriptors qcb_fit_epoch_bases writes for its one-launch batches)
        stage = self._index_staging
        desc_bytes = int(lib.qcb_fit_epoch_desc_bytes(G, int(num_batches)))
        if stage is None or stage[0][0].numel() < n + neg_perm.numel() or stage[0][2].numel() < desc_bytes:
            stage = self._index_staging = [[torch.empty(n + neg_perm.numel(), dtype=torch.long).pin_memory(), None,
                                            torch.empty(max(desc_bytes, 16), dtype=torch.uint8).pin_memory()] for _ in range(2)]
            self._desc_dev = torch.empty(max(desc_bytes, 16), dtype=torch.uint8, device=dev)
        stage.reverse()
        buf, busy, desc_host = stage[0]
        if busy is not None:
            busy.synchronize()
        host = buf[:n + neg_perm.numel()]
        host[:n] = torch.from_numpy(pos_perm[order])
        host[n:] = neg_perm
        index = host.to(dev, non_blocking=True)
        pos = train_samples.index_select(0, index[:n])
        neg = z_samples.index_select(0, index[n:])
        codes, names = self._basis_codes
        table = unitaries.unitary_table(self.unitary_dict, names, dev)
        na = int(getattr(am, "num_aux", 0))
        bmax = int(counts.max()) if counts.size else 0
        C = int(neg_batch_size)
        op = _lib.OP_DM_ROTATED if na else _lib.OP_CWF_ROTATED
        ws = _lib.workspace(dev, max(lib.qcb_workspace_bytes(op, N, am.num_hidden, na, bmax),
                                     lib.qcb_workspace_bytes(_lib.OP_CDK_NEG, N, am.num_hidden, na, C), 256))
        scratch = torch.empty(2 * max(int(pos_batch_size), 1), dtype=torch.double, device=dev)
        g = torch.empty(am.num_pars + ph.num_pars, dtype=torch.double, device=dev)
        vk = torch.empty(max(C, 1) * N, dtype=torch.double, device=dev)
        if self._cd_sweep + num_batches * int(k) >= (1 << 31):
            self._cd_seed, self._cd_sweep = next_seed(), 0
        # parameters straight from the modules' dictionaries (nn.Module.__getattr__ is the slow path, and this runs every epoch)
        names_p = ("weights_W", "weights_U", "visible_bias", "hidden_bias", "aux_bias") if na else ("weights", None, "visible_bias",
                                                                                                   "hidden_bias", None)
        ptrs = [r._parameters[name].data_ptr() if name is not None else None for r in (am, ph) for name in names_p]
        _lib.check(lib.qcb_fit_epoch_bases(*ptrs, N, am.num_hidden, na, _lib.ptr(pos), offsets.ctypes.data, codes.ctypes.data, G,
                                           int(num_batches), _lib.ptr(table), len(names), _lib.ptr(neg), C, int(k), float(lr),
                                           self._cd_seed, self._cd_sweep, _lib.ptr(scratch), _lib.ptr(g), _lib.ptr(vk), _lib.ptr(ws),
                                           ws.numel(), desc_host.data_ptr(), self._desc_dev.data_ptr(), desc_bytes,
                                           _lib.stream_ptr(dev)), "qcb_fit_epoch_bases")
        stage[0][1] = torch.cuda.Event()   # this slot's index vector and descriptors have crossed once this has happened
        stage[0][1].record(torch.cuda.current_stream(dev))
        self._cd_sweep += num_batches * int(k)
        if want_grads:
            vector_to_grads(g[:am.num_pars], am.parameters())
            vector_to_grads(g[am.num_pars:], ph.parameters())

    def _fit_epoch_fused(self, pos, neg, num_batches, pos_batch_size, neg_batch_size, k, lr, want_grads):
        """All batches of one epoch in one launch; parameters are updated in place."""
        rbm = self.rbm_am
        lib = _lib.load()
        if self._cd_sweep + num_batches * int(k) >= (1 << 31):
            self._cd_seed, self._cd_sweep = next_seed(), 0
        pos = pos.contiguous()
        neg = neg.contiguous()
        g_last = torch.empty(rbm.num_pars, dtype=torch.double, device=self.device) if want_grads else None
        _lib.check(lib.qcb_rbm_cd_epoch(_lib.ptr(rbm.weights.data), _lib.ptr(rbm.visible_bias.data), _lib.ptr(rbm.hidden_bias.data),
                                        rbm.num_visible, rbm.num_hidden, _lib.ptr(pos), pos.shape[0], int(pos_batch_size),
                                        _lib.ptr(neg), neg.shape[0], int(neg_batch_size), int(num_batches), int(k), float(lr),
                                        self._cd_seed, self._cd_sweep, _lib.ptr(g_last), _lib.stream_ptr(self.device)),
                   "qcb_rbm_cd_epoch")
        self._cd_sweep += num_batches * int(k)
        if want_grads:
            vector_to_grads(g_last, rbm.parameters())

    def fit(self, data, epochs=100, pos_batch_size=100, neg_batch_size=None, k=1, lr=1e-3, input_bases=None, progbar=False,
            starting_epoch=1, time=False, callbacks=None, optimizer=torch.optim.SGD, optimizer_args=None, scheduler=None,
            scheduler_args=None, **kwargs):
        """Train with contrastive divergence (neural_state.py:500-636). Same arguments and hook
        order as the reference; ``progbar`` accepts True/"notebook" when tqdm is installed."""
        if self.stop_training:
            return
        cbs = _Callbacks(callbacks)
        if time:
            from ..callbacks import Timer

            cbs.callbacks.append(Timer())
        neg_batch_size = neg_batch_size if neg_batch_size else pos_batch_size
        if isinstance(data, torch.Tensor):
            train_samples = data.clone().detach().to(device=self.device, dtype=torch.double)
        else:
            train_samples = torch.tensor(data, device=self.device, dtype=torch.double)
        all_params = list(chain(*[getattr(self, net).parameters() for net in self.networks]))
        optimizer = optimizer(all_params, lr=lr, **(optimizer_args or {}))
        if scheduler is not None:
            scheduler = scheduler(optimizer, **(scheduler_args or {}))
        # plain SGD (no momentum / weight decay / hooks): fused multi-tensor update, .grad set only when callbacks can look
        pg = optimizer.param_groups
        plain_sgd = (type(optimizer) is torch.optim.SGD and len(pg) == 1 and scheduler is None
                     and pg[0].get("momentum", 0) == 0 and pg[0].get("weight_decay", 0) == 0
                     and pg[0].get("dampening", 0) == 0 and not pg[0].get("nesterov", False)
                     and not pg[0].get("maximize", False))
        net_params = [list(getattr(self, net).parameters()) for net in self.networks]
        z_samples = None
        if input_bases is not None:
            self._validate_rotated_shapes(input_bases)  # a clear error now instead of QCB_ERR_UNSUPPORTED on the first rotated batch
            z_samples = extract_refbasis_samples(train_samples, input_bases).to(device=self.device)
            input_bases = np.asarray(input_bases)
            uniq, ids_all = np.unique(input_bases, axis=0, return_inverse=True)
            self._basis_groups = (uniq, np.asarray(ids_all).reshape(-1))
            rows, names = [], None
            for basis in uniq:
                c_, names = unitaries.encode_basis(basis, self.unitary_dict)
                rows.append(c_)
            self._basis_codes = (np.ascontiguousarray(np.stack(rows).astype(np.uint8)), names)

        from .. import distributed as qd

        # one Philox key per fit() call; under torch.distributed rank 0's key is broadcast and also seeds the host generator of
        # the shuffles, so that every rank walks through the same batches whatever its own seeding was
        self._cd_seed, self._cd_sweep = qd.broadcast_seed(next_seed(), self.device), 0
        # data sets without bases are shuffled on the device (nothing for the host to wait on; a host randperm of 8 M rows
        # also takes longer than the epoch's kernels); the generator is seeded from the fit key, so torch.manual_seed still
        # reproduces the run
        if input_bases is None and train_samples.is_cuda and train_samples.shape[0] >= self.device_shuffle_min_rows:
            self._shuffle_generator = torch.Generator(device=train_samples.device).manual_seed(self._cd_seed)
        else:
            self._shuffle_generator = torch.Generator().manual_seed(self._cd_seed) if qd.world()[1] > 1 else None
        # large binary data for a 33..64-unit BinaryRBM: keep a packed copy (1 bit per unit) and hand the batches out as
        # packed rows — the shuffle gathers 8 bytes per sample instead of 512 and the whole CD step runs on bits
        self._train_bits = None
        if (self.pack_training_data and input_bases is None and self._shuffle_generator is not None
                and self._shuffle_generator.device.type == "cuda" and train_samples.shape[0] >= self.lazy_batches_min_rows
                and self._packed_training_shape() and int(k) >= 1
                and min(pos_batch_size, neg_batch_size) // max(1, qd.world()[1]) >= self.packed_min_rows
                and bool(((train_samples == 0) | (train_samples == 1)).all())):
            from ..rbm._sampling import pack_states

            self._train_bits = pack_states(train_samples)
        cbs.fire("on_train_start", self)
        num_batches = ceil(train_samples.shape[0] / pos_batch_size)
        epoch_iter = range(starting_epoch, epochs + 1)
        if progbar:
            from tqdm import tqdm, tqdm_notebook

            epoch_iter = (tqdm_notebook if progbar == "notebook" else tqdm)(epoch_iter, desc="Epochs ")
        # one launch per epoch: small RBM, small batches (a single thread-block cluster walks the whole epoch); schedulers count
        # optimizer.step() calls, so they keep the per-batch loop
        fused_epoch = (scheduler is None and max(pos_batch_size, neg_batch_size) <= 4096
                       and self._shards_batches(pos_batch_size, neg_batch_size, input_bases is not None)[1] == 1
                       and self._epoch_fusable(optimizer, cbs, input_bases))
        # data with bases: the whole batch loop of an epoch in one C call (the kernels of the per-batch path in the same order,
        # the SGD step included; at two qubits the per-batch loop was host time)
        fused_bases = (input_bases is not None and scheduler is None and self._bases_epoch_fusable(plain_sgd, cbs)
                       and self._shards_batches(pos_batch_size, neg_batch_size, True)[1] == 1 and train_samples.shape[0] > 0)
        # fused epochs with device-side shuffles: the NEXT epoch's permutation and gathers run on a side stream while this
        # epoch's kernel (one 16-CTA cluster, the rest of the GPU is idle) runs on the caller's stream
        side = None
        if fused_epoch and self._shuffle_generator is not None and self._shuffle_generator.device.type == "cuda":
            side = _shuffle_stream(self.device)  # one per device for the process: the caching allocator pools blocks per stream
            side.wait_stream(torch.cuda.current_stream(self.device))

        def shuffled_epoch():
            if side is None:
                return self._shuffle_tensors(pos_batch_size, neg_batch_size, num_batches, train_samples, None, None,
                                             generator=self._shuffle_generator)[:2] + (None,)
            with torch.cuda.stream(side):
                p_, n_, _ = self._shuffle_tensors(pos_batch_size, neg_batch_size, num_batches, train_samples, None, None,
                                                  generator=self._shuffle_generator)
                ready = torch.cuda.Event()
                ready.record(side)
            return p_, n_, ready

        ahead = shuffled_epoch() if fused_epoch else None
        in_flight = []
        for ep in epoch_iter:
            if fused_epoch:
                pos, neg, ready = ahead
                if ready is not None:
                    main = torch.cuda.current_stream(self.device)
                    main.wait_event(ready)
                    pos.record_stream(main)
                    neg.record_stream(main)
                cbs.fire("on_epoch_start", self, ep)
                nb = min(num_batches, ceil(neg.shape[0] / neg_batch_size))
                self._fit_epoch_fused(pos, neg, nb, pos_batch_size, neg_batch_size, k, optimizer.param_groups[0]["lr"],
                                      want_grads=bool(cbs.callbacks))
                batches = ()
                if ep != epochs:
                    ahead = shuffled_epoch()  # queued behind nothing but the previous shuffle: overlaps the kernel just launched
                if side is not None:
                    # nothing above makes the host wait, so bound its lead: with hundreds of epochs queued, every epoch's
                    # shuffled copy of the data would be alive at once (and the caching allocator would have to grow for it)
                    done = torch.cuda.Event()
                    done.record(torch.cuda.current_stream(self.device))
                    in_flight.append(done)
                    if len(in_flight) > 3:
                        in_flight.pop(0).synchronize()
            elif fused_bases:
                cbs.fire("on_epoch_start", self, ep)
                self._fit_epoch_bases(train_samples, z_samples, pos_batch_size, neg_batch_size, num_batches, k,
                                      optimizer.param_groups[0]["lr"], want_grads=bool(cbs.callbacks))
                batches = ()
            else:
                batches = self._shuffle_data(pos_batch_size, neg_batch_size, num_batches, train_samples, input_bases, z_samples)
                cbs.fire("on_epoch_start", self, ep)
            for b, batch in enumerate(batches):
                cbs.fire("on_batch_start", self, ep, b)
                all_grads = self.compute_batch_gradients(k, *batch)
                if plain_sgd:
                    # vanilla SGD (the reference's default optimizer): p -= lr * g as one fused multi-tensor
                    # update on views of the flat gradient vectors; same arithmetic as optimizer.step()
                    views = []
                    for i, net in enumerate(self.networks):
                        offset = 0
                        for prm in net_params[i]:
                            views.append(all_grads[i][offset:offset + prm.numel()].view(prm.size()))
                            offset += prm.numel()
                    if cbs.callbacks:  # observers find this batch's gradient in .grad, as after optimizer.step()
                        for prm, view in zip(all_params, views):
                            prm.grad = view
                    torch._foreach_add_(all_params, views, alpha=-optimizer.param_groups[0]["lr"])
                else:
                    optimizer.zero_grad()
                    for i, net in enumerate(self.networks):
                        vector_to_grads(all_grads[i], getattr(self, net).parameters())
                    optimizer.step()
                cbs.fire("on_batch_end", self, ep, b)
                if self.stop_training:
                    break
            if scheduler is not None:
                scheduler.step()
            cbs.fire("on_epoch_end", self, ep)
            if self.stop_training:
                break
        self._cd_seed = None
        self._shuffle_generator = None
        self._train_bits = None
        self._basis_groups = None
        self._basis_codes = None
        self._index_staging = None
        self._desc_dev = None
        cbs.fire("on_train_end", self)
