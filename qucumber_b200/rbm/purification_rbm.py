"""CUDA-backed drop-in for ``qucumber.rbm.PurificationRBM`` (reference:
qucumber/rbm/purification_rbm.py:26-463): an RBM with a hidden and an auxiliary ("purification")
layer, both connected to the visible units, used by DensityMatrix.

Same parameter names (``weights_W``, ``weights_U``, ``visible_bias``, ``hidden_bias``,
``aux_bias``), methods and shapes as the reference; the arithmetic runs in libqucumber_b200.so.
Internally the kernels treat [W; U] as one stacked weight matrix with biases [c; d].
"""
import numpy as np
import torch
from torch import nn

from .. import _lib
from ..rng import next_seed
from ..utils import auto_unsqueeze_args
from ..utils import cplx
from ._sampling import ChainSamplingMixin
from .binary_rbm import _as_rows


class PurificationRBM(ChainSamplingMixin, nn.Module):
    def __init__(self, num_visible, num_hidden=None, num_aux=None, zero_weights=False, gpu=False):
        super().__init__()
        self.num_visible = int(num_visible)
        self.num_hidden = int(num_hidden) if num_hidden is not None else self.num_visible
        self.num_aux = int(num_aux) if num_aux is not None else self.num_visible
        # parameter order: W (visible-hidden), U (visible-auxiliary), b (visible), c (hidden), d (auxiliary)
        self.num_pars = (
            self.num_visible * self.num_hidden + self.num_visible * self.num_aux
            + self.num_visible + self.num_hidden + self.num_aux
        )
        self.gpu = gpu and torch.cuda.is_available()
        # an INDEXED device: streams and workspaces are looked up per device, and "cuda" alone would follow whatever device is
        # current at call time rather than the one that owns the parameters
        self.device = torch.device("cuda", torch.cuda.current_device()) if self.gpu else torch.device("cpu")
        self.initialize_parameters(zero_weights=zero_weights)

    def __repr__(self):
        return (f"PurificationBinaryRBM(num_visible={self.num_visible}, num_hidden={self.num_hidden}, "
                f"num_aux={self.num_aux}, gpu={self.gpu})")

    def initialize_parameters(self, zero_weights=False):
        """purification_rbm.py:89-128: W, U ~ randn / sqrt(N) (or zeros), all biases zero."""
        gen = torch.zeros if zero_weights else torch.randn
        kw = dict(dtype=torch.double, device=self.device)
        scale = np.sqrt(self.num_visible)
        self.weights_W = nn.Parameter(gen(self.num_hidden, self.num_visible, **kw) / scale, requires_grad=False)
        self.weights_U = nn.Parameter(gen(self.num_aux, self.num_visible, **kw) / scale, requires_grad=False)
        self.visible_bias = nn.Parameter(torch.zeros(self.num_visible, **kw), requires_grad=False)
        self.hidden_bias = nn.Parameter(torch.zeros(self.num_hidden, **kw), requires_grad=False)
        self.aux_bias = nn.Parameter(torch.zeros(self.num_aux, **kw), requires_grad=False)

    # ------------------------------------------------------------------------------------
    def _tensors(self):
        t = tuple(p.data.contiguous() for p in (self.weights_W, self.weights_U, self.visible_bias, self.hidden_bias,
                                                self.aux_bias))
        _lib.require_cuda(t[0], "PurificationRBM")
        return t

    def _param_args(self, with_aux=True):
        W, U, b, c, d = self._tensors()
        if with_aux:
            return (_lib.ptr(W), _lib.ptr(U), _lib.ptr(b), _lib.ptr(c), _lib.ptr(d), self.num_visible, self.num_hidden,
                    self.num_aux), (W, U, b, c, d)
        return (_lib.ptr(W), None, _lib.ptr(b), _lib.ptr(c), None, self.num_visible, self.num_hidden, 0), (W, U, b, c, d)

    def _energy(self, v, with_aux):
        args, keep = self._param_args(with_aux)
        vr = _as_rows(v, self.weights_W, self.num_visible, "effective_energy")
        out = torch.empty(vr.shape[0], dtype=torch.double, device=vr.device)
        _lib.check(_lib.load().qcb_rbm_energy(*args, _lib.ptr(vr), vr.shape[0], _lib.ptr(out), None, 0,
                                              _lib.stream_ptr(vr.device)), "qcb_rbm_energy")
        return out.reshape(v.shape[:-1])

    def _linear(self, M, x, bias, transpose=False, sigmoid=False, M2=None, x2=None):
        """out = x M^T + bias (or x M for transpose) [+ x2 M2], optionally through clamp(sigmoid)."""
        lib = _lib.load()
        n_in = M.shape[0] if transpose else M.shape[1]
        n_out = M.shape[1] if transpose else M.shape[0]
        xr = _as_rows(x, self.weights_W, n_in, "linear")
        out = torch.empty(*x.shape[:-1], n_out, dtype=torch.double, device=xr.device)
        if sigmoid:
            x2r = None if x2 is None else _as_rows(x2, self.weights_W, M2.shape[0], "linear")
            rc = lib.qcb_linear_sigmoid(_lib.ptr(M), M.shape[0], M.shape[1], int(transpose), _lib.ptr(bias), _lib.ptr(xr),
                                        _lib.ptr(M2), 0 if M2 is None else M2.shape[0], _lib.ptr(x2r), xr.shape[0],
                                        _lib.ptr(out), _lib.stream_ptr(xr.device))
            _lib.check(rc, "qcb_linear_sigmoid")
        else:
            rc = lib.qcb_linear(_lib.ptr(M), M.shape[0], M.shape[1], int(transpose), _lib.ptr(bias), _lib.ptr(xr), xr.shape[0],
                                _lib.ptr(out), _lib.stream_ptr(xr.device))
            _lib.check(rc, "qcb_linear")
        return out

    @staticmethod
    def _into(res, out):
        if out is None:
            return res
        out.copy_(res.reshape(out.shape))
        return out

    @auto_unsqueeze_args()
    def effective_energy(self, v, a=None):
        """purification_rbm.py:131-158. With ``a=None`` the auxiliary units are traced out
        analytically (one fused kernel); with explicit ``a`` the bilinear form is evaluated."""
        v = v.to(self.weights_W)
        if a is None:
            return self._energy(v, with_aux=True)
        a = (a.unsqueeze(0) if a.dim() < 2 else a).to(self.weights_W)
        W, U, b, c, d = self._tensors()
        vis_term = -self._energy(v, with_aux=False)              # b.v + sum softplus(W v + c)
        ua = self._linear(U, v, d)                                # (…, na): U v + d
        return -(vis_term + (ua * a).sum(-1))                     # + d.a + a^T U v

    def effective_energy_gradient(self, v, reduce=True):
        """purification_rbm.py:160-193, flat order [W, U, b, c, d]."""
        v = (v.unsqueeze(0) if v.dim() < 2 else v).to(self.weights_W)
        args, keep = self._param_args()
        lib = _lib.load()
        vr = _as_rows(v, self.weights_W, self.num_visible, "effective_energy_gradient")
        if reduce:
            if v.dim() != 2:
                raise ValueError("effective_energy_gradient(reduce=True) expects a 2-D batch")
            g = torch.empty(self.num_pars, dtype=torch.double, device=vr.device)
            ws = _lib.workspace(vr.device, lib.qcb_workspace_bytes(_lib.OP_GRAD_REDUCE, self.num_visible, self.num_hidden,
                                                                   self.num_aux, vr.shape[0]))
            _lib.check(lib.qcb_rbm_grad_reduce(*args, _lib.ptr(vr), vr.shape[0], 1.0, _lib.ptr(g), 0, _lib.ptr(ws), ws.numel(),
                                               _lib.stream_ptr(vr.device)), "qcb_rbm_grad_reduce")
            return g
        out = torch.empty(vr.shape[0], self.num_pars, dtype=torch.double, device=vr.device)
        _lib.check(lib.qcb_rbm_grad_batch(*args, _lib.ptr(vr), vr.shape[0], _lib.ptr(out), _lib.stream_ptr(vr.device)),
                   "qcb_rbm_grad_batch")
        return out.reshape(*v.shape[:-1], self.num_pars)

    @auto_unsqueeze_args()
    def prob_h_given_v(self, v, out=None):
        """purification_rbm.py:196-215."""
        W, U, b, c, d = self._tensors()
        return self._into(self._linear(W, v, c, sigmoid=True), out)

    @auto_unsqueeze_args()
    def prob_a_given_v(self, v, out=None):
        """purification_rbm.py:217-236."""
        W, U, b, c, d = self._tensors()
        return self._into(self._linear(U, v, d, sigmoid=True), out)

    @auto_unsqueeze_args(1, 2)
    def prob_v_given_ha(self, h, a, out=None):
        """purification_rbm.py:238-259: clamp(sigmoid(h W + b + a U), 0, 1)."""
        W, U, b, c, d = self._tensors()
        return self._into(self._linear(W, h, b, transpose=True, sigmoid=True, M2=U, x2=a), out)

    def _bernoulli(self, p, uniforms):
        pc = p if p.is_contiguous() else p.contiguous()
        u = None if uniforms is None else uniforms.to(pc).contiguous()
        _lib.check(_lib.load().qcb_bernoulli(_lib.ptr(pc), pc.numel(), _lib.ptr(u), next_seed(), 0, 0, _lib.ptr(pc),
                                             _lib.stream_ptr(pc.device)), "qcb_bernoulli")
        if pc is not p:
            p.copy_(pc)
        return p

    def sample_a_given_v(self, v, out=None, uniforms=None):
        return self._bernoulli(self.prob_a_given_v(v, out=out), uniforms)

    def sample_h_given_v(self, v, out=None, uniforms=None):
        return self._bernoulli(self.prob_h_given_v(v, out=out), uniforms)

    def sample_v_given_ha(self, h, a, out=None, uniforms=None):
        return self._bernoulli(self.prob_v_given_ha(h, a, out=out), uniforms)

    def gibbs_steps(self, k, initial_state, overwrite=False, uniforms=None, seed=None, chain_id0=0, sweep0=0):
        """k sweeps of (h | v), (a | v), (v | h, a) in one kernel launch (purification_rbm.py:309-337).
        ``uniforms`` (k, *batch, nh + na + N) replays given draws in the reference's order h, a, v."""
        v = (initial_state if overwrite else initial_state.clone()).to(self.weights_W)
        args, keep = self._param_args()
        vc = v if v.is_contiguous() else v.contiguous()
        if vc.shape[-1] != self.num_visible:
            raise ValueError(f"gibbs_steps: last dimension is {vc.shape[-1]}, expected {self.num_visible}")
        C = vc.numel() // self.num_visible
        lib = _lib.load()
        u = None
        if uniforms is not None:
            u = uniforms.to(self.weights_W).reshape(int(k), C, self.num_hidden + self.num_aux + self.num_visible).contiguous()
        ws = _lib.workspace(vc.device, lib.qcb_workspace_bytes(_lib.OP_GIBBS, self.num_visible, self.num_hidden, self.num_aux, C))
        _lib.check(lib.qcb_rbm_gibbs(*args, _lib.ptr(vc), C, int(k), next_seed() if seed is None else int(seed), int(chain_id0),
                                     int(sweep0), _lib.ptr(u), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(vc.device)),
                   "qcb_rbm_gibbs")
        if vc is not v:
            v.copy_(vc)
        return v

    @auto_unsqueeze_args()
    def mixing_term(self, v):
        """V = (1/2) U v + d (purification_rbm.py:340-350)."""
        W, U, b, c, d = self._tensors()
        return self._linear((0.5 * U).contiguous(), v.to(self.weights_W), d)

    def _f(self, v):
        """b.v + sum_i softplus(W_i.v + c_i): the quantity both Gamma matrices are built from."""
        return -self._energy(v.to(self.weights_W), with_aux=False)

    def gamma(self, v, vp, eta=1, expand=True):
        """Elements of Gamma^(eta) = (1/2)[f(v) + eta f(v')] (purification_rbm.py:352-396)."""
        sign = float(np.sign(eta))
        if v.dim() < 2 and vp.dim() < 2:
            return 0.5 * (self._f(v.unsqueeze(0))[0] + sign * self._f(vp.unsqueeze(0))[0])
        t1, t2 = self._f(v), self._f(vp)
        if expand:
            return 0.5 * (t1.unsqueeze(1) + sign * t2.unsqueeze(0))
        return 0.5 * (t1 + sign * t2)

    def gamma_grad(self, v, vp, eta=1, expand=False):
        """Gradient of Gamma^(eta) w.r.t. this RBM's parameters as a complex tensor (2, …, G)
        (purification_rbm.py:398-451): only the W, b, c blocks are non-zero."""
        sign = float(np.sign(eta))
        unsqueezed = v.dim() < 2 or vp.dim() < 2
        v = (v.unsqueeze(0) if v.dim() < 2 else v).to(self.weights_W)
        vp = (vp.unsqueeze(0) if vp.dim() < 2 else vp).to(self.weights_W)
        ph, php = self.prob_h_given_v(v), self.prob_h_given_v(vp)
        wg = ph.unsqueeze(-1) * v.unsqueeze(-2)
        wgp = php.unsqueeze(-1) * vp.unsqueeze(-2)
        if expand:
            W_grad = 0.5 * (wg.unsqueeze(1) + sign * wgp.unsqueeze(0))
            vb_grad = 0.5 * (v.unsqueeze(1) + sign * vp.unsqueeze(0))
            hb_grad = 0.5 * (ph.unsqueeze(1) + sign * php.unsqueeze(0))
            batch = (v.shape[0], vp.shape[0], *v.shape[1:-1])
        else:
            W_grad = 0.5 * (wg + sign * wgp)
            vb_grad = 0.5 * (v + sign * vp)
            hb_grad = 0.5 * (ph + sign * php)
            batch = tuple(v.shape[:-1])
        zeros_U = torch.zeros(*batch, self.num_aux * self.num_visible, dtype=torch.double, device=v.device)
        zeros_d = torch.zeros(*batch, self.num_aux, dtype=torch.double, device=v.device)
        parts = [W_grad.reshape(*batch, -1), zeros_U, vb_grad, hb_grad, zeros_d]
        if unsqueezed and not expand:
            parts = [p.squeeze(0) for p in parts]
        return cplx.make_complex(torch.cat(parts, dim=-1))

    def log_partition(self, first_state=0, count=None):
        args, keep = self._param_args()
        lib = _lib.load()
        count = (1 << self.num_visible) - first_state if count is None else count
        out2 = torch.empty(2, dtype=torch.double, device=self.weights_W.device)
        ws = _lib.workspace(out2.device, lib.qcb_workspace_bytes(_lib.OP_LOGZ, self.num_visible, self.num_hidden, self.num_aux, 0))
        _lib.check(lib.qcb_rbm_logz_range(*args, int(first_state), int(count), _lib.ptr(out2), _lib.ptr(ws), ws.numel(),
                                          _lib.stream_ptr(out2.device)), "qcb_rbm_logz_range")
        return out2

    def partition(self, space=None):
        """purification_rbm.py:453-463; ``space=None`` / HilbertSpace marker enumerates from indices."""
        from ..hilbert import HilbertSpace

        if space is None or isinstance(space, HilbertSpace):
            out2 = self.log_partition()
        else:
            E = self.effective_energy(space.to(self.weights_W)).reshape(-1).contiguous()
            lib = _lib.load()
            out2 = torch.empty(2, dtype=torch.double, device=E.device)
            ws = _lib.workspace(E.device, lib.qcb_workspace_bytes(_lib.OP_LOGZ, self.num_visible, self.num_hidden, self.num_aux, 0))
            _lib.check(lib.qcb_neg_logsumexp(_lib.ptr(E), E.numel(), _lib.ptr(out2), _lib.ptr(ws), ws.numel(),
                                             _lib.stream_ptr(E.device)), "qcb_neg_logsumexp")
        return (out2[0] + out2[1].log()).exp()
