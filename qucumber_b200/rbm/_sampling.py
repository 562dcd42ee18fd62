"""Sampling entry points shared by BinaryRBM and PurificationRBM that have no counterpart in the reference's RBM
classes: chains started IN the kernel (NeuralStateBase.sample without an initial state,
qucumber/nn_states/neural_state.py:124-131) and compact ("packed") chain states — ceil(N / 8) bytes per chain,
numpy.packbits order — next to the reference's float64 rows (qucumber/rbm/binary_rbm.py:220).
"""
import torch

from .. import _lib
from ..rng import next_seed


def packed_row_bytes(num_visible):
    return (int(num_visible) + 7) // 8


def pack_states(v):
    """(…, N) tensor of 0/1 values on a CUDA device -> (…, ceil(N / 8)) uint8, numpy.packbits(axis=-1) order."""
    _lib.require_cuda(v, "pack_states")
    n = v.shape[-1]
    rows = v.to(torch.double).reshape(-1, n).contiguous()
    bits = torch.empty(rows.shape[0], packed_row_bytes(n), dtype=torch.uint8, device=v.device)
    _lib.check(_lib.load().qcb_pack_states(_lib.ptr(rows), rows.shape[0], n, _lib.ptr(bits), _lib.stream_ptr(v.device)),
               "qcb_pack_states")
    return bits.reshape(*v.shape[:-1], packed_row_bytes(n))


def unpack_states(bits, num_visible):
    """Inverse of :func:`pack_states`: (…, ceil(N / 8)) uint8 on a CUDA device -> (…, N) float64 of 0.0 / 1.0."""
    _lib.require_cuda(bits, "unpack_states")
    rb = packed_row_bytes(num_visible)
    if bits.dtype != torch.uint8 or bits.shape[-1] != rb:
        raise ValueError(f"unpack_states: expected uint8 rows of {rb} bytes, got {bits.dtype} {tuple(bits.shape)}")
    flat = bits.reshape(-1, rb).contiguous()
    v = torch.empty(flat.shape[0], int(num_visible), dtype=torch.double, device=bits.device)
    _lib.check(_lib.load().qcb_unpack_states(_lib.ptr(flat), flat.shape[0], int(num_visible), _lib.ptr(v),
                                             _lib.stream_ptr(bits.device)), "qcb_unpack_states")
    return v.reshape(*bits.shape[:-1], int(num_visible))


class ChainSamplingMixin:
    """Needs ``_param_args()`` -> ((W, U, b, c, d, N, nh, na) ctypes arguments, tensors kept alive)."""

    def _sampling_device(self):
        return next(self.parameters()).device

    def sample_chains(self, k, num_chains, seed=None, chain_id0=0, sweep0=0, packed=False):
        """``num_chains`` chains started from v0 ~ Bernoulli(1/2) drawn in the kernel, k block-Gibbs sweeps, ONE launch,
        no host random numbers and no host-to-device copy. Returns (num_chains, N) float64 like the reference's
        ``sample``, or (num_chains, ceil(N / 8)) uint8 packed rows with ``packed=True``."""
        args, keep = self._param_args()
        dev = self._sampling_device()
        lib = _lib.load()
        C, n = int(num_chains), self.num_visible
        na = getattr(self, "num_aux", 0)
        v = None if packed else torch.empty(C, n, dtype=torch.double, device=dev)
        bits = torch.empty(C, packed_row_bytes(n), dtype=torch.uint8, device=dev) if packed else None
        ws = _lib.workspace(dev, lib.qcb_workspace_bytes(_lib.OP_SAMPLE, n, self.num_hidden, na, C if packed else 0))
        _lib.check(lib.qcb_rbm_sample(*args, C, int(k), next_seed() if seed is None else int(seed), int(chain_id0), int(sweep0),
                                      _lib.ptr(v), _lib.ptr(bits), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)),
                   "qcb_rbm_sample")
        return bits if packed else v

    def gibbs_steps_packed(self, k, bits, overwrite=False, seed=None, chain_id0=0, sweep0=0):
        """``gibbs_steps`` on packed rows (C, ceil(N / 8)) uint8: same chains as the float64 call for the same key."""
        args, keep = self._param_args()
        n, rb = self.num_visible, packed_row_bytes(self.num_visible)
        if bits.dtype != torch.uint8 or bits.shape[-1] != rb:
            raise ValueError(f"gibbs_steps_packed: expected uint8 rows of {rb} bytes, got {bits.dtype} {tuple(bits.shape)}")
        _lib.require_cuda(bits, "gibbs_steps_packed")
        out = bits if overwrite else bits.clone()
        oc = out if out.is_contiguous() else out.contiguous()
        C = oc.numel() // rb
        lib = _lib.load()
        na = getattr(self, "num_aux", 0)
        ws = _lib.workspace(oc.device, lib.qcb_workspace_bytes(_lib.OP_SAMPLE, n, self.num_hidden, na, C))
        _lib.check(lib.qcb_rbm_gibbs_packed(*args, _lib.ptr(oc), C, int(k), next_seed() if seed is None else int(seed),
                                            int(chain_id0), int(sweep0), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(oc.device)),
                   "qcb_rbm_gibbs_packed")
        if oc is not out:
            out.copy_(oc)
        return out
