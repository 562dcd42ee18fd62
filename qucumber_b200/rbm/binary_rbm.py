"""CUDA-backed drop-in for ``qucumber.rbm.BinaryRBM`` (reference: qucumber/rbm/binary_rbm.py:26-241).

Same constructor, parameter names (``weights``, ``visible_bias``, ``hidden_bias`` — so
``state_dict`` checkpoints interchange), methods and return conventions as the reference
class, but every numeric method is one call into libqucumber_b200.so (sm_100a kernels).
There is no CPU path: calling a numeric method on a model that is not on a CUDA device raises.
"""
import ctypes

import numpy as np
import torch
from torch import nn

from .. import _lib
from ..rng import next_seed
from ..utils import auto_unsqueeze_args
from ._sampling import ChainSamplingMixin


def _as_rows(x, like, width, what):
    """(…, width) tensor -> contiguous float64 (rows, width) on the parameter device."""
    x = x.to(like)
    if x.shape[-1] != width:
        raise ValueError(f"{what}: last dimension is {x.shape[-1]}, expected {width}")
    return x.reshape(-1, width).contiguous()


class BinaryRBM(ChainSamplingMixin, nn.Module):
    def __init__(self, num_visible, num_hidden=None, zero_weights=False, gpu=True):
        super().__init__()
        self.num_visible = int(num_visible)
        self.num_hidden = int(num_hidden) if num_hidden else self.num_visible
        self.num_pars = (self.num_visible * self.num_hidden) + self.num_visible + self.num_hidden
        self.gpu = gpu and torch.cuda.is_available()
        # an INDEXED device: streams and workspaces are looked up per device, and "cuda" alone would follow whatever device is
        # current at call time rather than the one that owns the parameters
        self.device = torch.device("cuda", torch.cuda.current_device()) if self.gpu else torch.device("cpu")
        self.initialize_parameters(zero_weights=zero_weights)

    def __repr__(self):
        return f"BinaryRBM(num_visible={self.num_visible}, num_hidden={self.num_hidden}, gpu={self.gpu})"

    def initialize_parameters(self, zero_weights=False):
        """Randomize the parameters of the RBM (binary_rbm.py:48-72: W ~ randn / sqrt(N), zero biases)."""
        gen_tensor = torch.zeros if zero_weights else torch.randn
        self.weights = nn.Parameter(
            gen_tensor(self.num_hidden, self.num_visible, device=self.device, dtype=torch.double)
            / np.sqrt(self.num_visible),
            requires_grad=False,
        )
        self.visible_bias = nn.Parameter(
            torch.zeros(self.num_visible, device=self.device, dtype=torch.double), requires_grad=False
        )
        self.hidden_bias = nn.Parameter(
            torch.zeros(self.num_hidden, device=self.device, dtype=torch.double), requires_grad=False
        )

    # ------------------------------------------------------------------------------------
    def _params(self):
        w = self.weights.data
        _lib.require_cuda(w, "BinaryRBM")
        return (w.contiguous(), self.visible_bias.data.contiguous(), self.hidden_bias.data.contiguous())

    def _param_args(self):
        w, b, c = self._params()
        return (_lib.ptr(w), None, _lib.ptr(b), _lib.ptr(c), None, self.num_visible, self.num_hidden, 0), (w, b, c)

    @auto_unsqueeze_args()
    def effective_energy(self, v):
        """E(v) = -(b.v + sum_i softplus(c_i + W_i.v)) for every row of v (binary_rbm.py:75-96)."""
        (args, keep) = self._param_args()
        vr = _as_rows(v, self.weights, self.num_visible, "effective_energy")
        out = torch.empty(vr.shape[0], dtype=torch.double, device=vr.device)
        lib = _lib.load()
        _lib.check(
            lib.qcb_rbm_energy(*args, _lib.ptr(vr), vr.shape[0], _lib.ptr(out), None, 0, _lib.stream_ptr(vr.device)),
            "qcb_rbm_energy",
        )
        return out.reshape(v.shape[:-1])

    def effective_energy_gradient(self, v, reduce=True):
        """Gradients of the effective energy (binary_rbm.py:98-126), flat order [W, b, c]."""
        v = (v.unsqueeze(0) if v.dim() < 2 else v).to(self.weights)
        (args, keep) = self._param_args()
        lib = _lib.load()
        if reduce:
            if v.dim() != 2:
                raise ValueError("effective_energy_gradient(reduce=True) expects a 2-D batch")
            vr = _as_rows(v, self.weights, self.num_visible, "effective_energy_gradient")
            g = torch.empty(self.num_pars, dtype=torch.double, device=vr.device)
            nbytes = lib.qcb_workspace_bytes(_lib.OP_GRAD_REDUCE, self.num_visible, self.num_hidden, 0, vr.shape[0])
            ws = _lib.workspace(vr.device, nbytes)
            _lib.check(
                lib.qcb_rbm_grad_reduce(*args, _lib.ptr(vr), vr.shape[0], 1.0, _lib.ptr(g), 0, _lib.ptr(ws), ws.numel(),
                                        _lib.stream_ptr(vr.device)),
                "qcb_rbm_grad_reduce",
            )
            return g
        vr = _as_rows(v, self.weights, self.num_visible, "effective_energy_gradient")
        out = torch.empty(vr.shape[0], self.num_pars, dtype=torch.double, device=vr.device)
        _lib.check(
            lib.qcb_rbm_grad_batch(*args, _lib.ptr(vr), vr.shape[0], _lib.ptr(out), _lib.stream_ptr(vr.device)),
            "qcb_rbm_grad_batch",
        )
        return out.reshape(*v.shape[:-1], self.num_pars)

    def _linear_sigmoid(self, x, transpose, n_in, n_out, bias, out, what):
        w = self.weights.data.contiguous()
        _lib.require_cuda(w, what)
        xr = _as_rows(x, self.weights, n_in, what)
        res = out if (out is not None and out.is_contiguous() and out.dtype == torch.double and out.device == xr.device
                      and out.numel() == xr.shape[0] * n_out) else torch.empty(
            *x.shape[:-1], n_out, dtype=torch.double, device=xr.device)
        lib = _lib.load()
        _lib.check(
            lib.qcb_linear_sigmoid(_lib.ptr(w), self.num_hidden, self.num_visible, int(transpose), _lib.ptr(bias), _lib.ptr(xr),
                                   None, 0, None, xr.shape[0], _lib.ptr(res), _lib.stream_ptr(xr.device)),
            "qcb_linear_sigmoid",
        )
        if out is not None and res is not out:
            out.copy_(res.reshape(out.shape))
            return out
        return res

    @auto_unsqueeze_args()
    def prob_v_given_h(self, h, out=None):
        """p(v_j = 1 | h) = clamp(sigmoid(h W + b), 0, 1) (binary_rbm.py:129-147)."""
        return self._linear_sigmoid(h, True, self.num_hidden, self.num_visible, self.visible_bias.data.contiguous(), out,
                                    "prob_v_given_h")

    @auto_unsqueeze_args()
    def prob_h_given_v(self, v, out=None):
        """p(h_i = 1 | v) = clamp(sigmoid(v W^T + c), 0, 1) (binary_rbm.py:150-168)."""
        return self._linear_sigmoid(v, False, self.num_visible, self.num_hidden, self.hidden_bias.data.contiguous(), out,
                                    "prob_h_given_v")

    def _bernoulli(self, p, uniforms):
        pc = p if p.is_contiguous() else p.contiguous()
        lib = _lib.load()
        u = None if uniforms is None else uniforms.to(pc).contiguous()
        _lib.check(
            lib.qcb_bernoulli(_lib.ptr(pc), pc.numel(), _lib.ptr(u), next_seed(), 0, 0, _lib.ptr(pc),
                              _lib.stream_ptr(pc.device)),
            "qcb_bernoulli",
        )
        if pc is not p:
            p.copy_(pc)
        return p

    def sample_v_given_h(self, h, out=None, uniforms=None):
        """Sample a visible state given a hidden state (binary_rbm.py:170-183)."""
        v = self.prob_v_given_h(h, out=out)
        return self._bernoulli(v, uniforms)

    def sample_h_given_v(self, v, out=None, uniforms=None):
        """Sample a hidden state given a visible state (binary_rbm.py:185-198)."""
        h = self.prob_h_given_v(v, out=out)
        return self._bernoulli(h, uniforms)

    def gibbs_steps(self, k, initial_state, overwrite=False, uniforms=None, seed=None, chain_id0=0, sweep0=0):
        """k steps of block Gibbs sampling in ONE kernel launch (binary_rbm.py:200-228).

        ``overwrite`` keeps the reference's aliasing contract: the returned tensor IS
        ``initial_state`` when it already lives on the model's device as float64.
        Extras (not in the reference): ``uniforms`` of shape (k, *batch, nh + N) replays given
        draws bit-exactly; ``seed`` / ``chain_id0`` / ``sweep0`` pin the Philox stream
        (default: a fresh key from torch's default generator, so ``torch.manual_seed`` makes
        runs reproducible)."""
        v = (initial_state if overwrite else initial_state.clone()).to(self.weights)
        (args, keep) = self._param_args()
        vc = v if v.is_contiguous() else v.contiguous()
        C = vc.numel() // self.num_visible if self.num_visible else 0
        if vc.shape[-1] != self.num_visible:
            raise ValueError(f"gibbs_steps: last dimension is {vc.shape[-1]}, expected {self.num_visible}")
        lib = _lib.load()
        u = None
        if uniforms is not None:
            u = uniforms.to(self.weights).reshape(int(k), C, self.num_hidden + self.num_visible).contiguous()
        ws = _lib.workspace(vc.device, lib.qcb_workspace_bytes(_lib.OP_GIBBS, self.num_visible, self.num_hidden, 0, C))
        _lib.check(
            lib.qcb_rbm_gibbs(*args, _lib.ptr(vc), C, int(k), next_seed() if seed is None else int(seed), int(chain_id0),
                              int(sweep0), _lib.ptr(u), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(vc.device)),
            "qcb_rbm_gibbs",
        )
        if vc is not v:
            v.copy_(vc)
        return v

    def log_partition(self, first_state=0, count=None):
        """(m, s) device tensor with log sum_{r in range} exp(-E(r)) = m + log(s); basis states are
        generated from their index in-kernel (no Hilbert-space tensor)."""
        (args, keep) = self._param_args()
        lib = _lib.load()
        total = 1 << self.num_visible
        count = total - first_state if count is None else count
        out2 = torch.empty(2, dtype=torch.double, device=self.weights.device)
        ws = _lib.workspace(out2.device, lib.qcb_workspace_bytes(_lib.OP_LOGZ, self.num_visible, self.num_hidden, 0, 0))
        _lib.check(
            lib.qcb_rbm_logz_range(*args, int(first_state), int(count), _lib.ptr(out2), _lib.ptr(ws), ws.numel(),
                                   _lib.stream_ptr(out2.device)),
            "qcb_rbm_logz_range",
        )
        return out2

    def partition(self, space=None):
        """Partition function Z = exp(logsumexp(-E(space))) (binary_rbm.py:230-241).

        ``space=None`` or a :class:`qucumber_b200.HilbertSpace` marker enumerates all 2^N states
        from their indices; an explicit tensor is evaluated row by row like the reference."""
        from ..hilbert import HilbertSpace

        if space is None or isinstance(space, HilbertSpace):
            out2 = self.log_partition()
        else:
            E = self.effective_energy(space.to(self.weights)).reshape(-1).contiguous()
            lib = _lib.load()
            out2 = torch.empty(2, dtype=torch.double, device=E.device)
            ws = _lib.workspace(E.device, lib.qcb_workspace_bytes(_lib.OP_LOGZ, self.num_visible, self.num_hidden, 0, 0))
            _lib.check(
                lib.qcb_neg_logsumexp(_lib.ptr(E), E.numel(), _lib.ptr(out2), _lib.ptr(ws), ws.numel(),
                                      _lib.stream_ptr(E.device)),
                "qcb_neg_logsumexp",
            )
        return (out2[0] + out2[1].log()).exp()
