"""Mirror of ``qucumber.nn_states.DensityMatrix`` (qucumber/nn_states/density_matrix.py:27-384):
a mixed state parametrised by two PurificationRBMs (amplitude and phase)."""
import torch

from ..rbm import PurificationRBM
from ..utils import cplx, unitaries
from .neural_state import NeuralStateBase, load_checkpoint, stored_attribute


class DensityMatrix(NeuralStateBase):
    _rbm_am = None
    _rbm_ph = None
    _device = None
    rbm_am = stored_attribute("_rbm_am")
    rbm_ph = stored_attribute("_rbm_ph")
    device = stored_attribute("_device")
    _grouped_call = "dm"
    _rotated_limits = (("num_aux", 256, "num_aux"), ("num_visible", 512, "num_visible"))
    _max_rotated_sites = 10  # 4^S pairs per sample

    def __init__(self, num_visible, num_hidden=None, num_aux=None, unitary_dict=None, gpu=True, module=None):
        if module is not None:
            raise ValueError("module= injection is not supported for two-RBM states (it is broken in the reference too)")
        self.rbm_am = PurificationRBM(num_visible, num_hidden, num_aux, gpu=gpu)
        self.rbm_ph = PurificationRBM(num_visible, num_hidden, num_aux, gpu=gpu)
        self.num_visible = self.rbm_am.num_visible
        self.num_hidden = self.rbm_am.num_hidden
        self.num_aux = self.rbm_am.num_aux
        self.device = self.rbm_am.device
        self.unitary_dict = unitary_dict if unitary_dict else unitaries.create_dict()
        self.unitary_dict = {k: v.to(device=self.device) for k, v in self.unitary_dict.items()}

    @property
    def networks(self):
        return ["rbm_am", "rbm_ph"]

    def _mix_terms(self, v, vp, expand):
        """Arguments of the Pi matrix: (1/2)[U_am(v + v') + 2 d_am] and (1/2) U_ph (v - v')
        (density_matrix.py:130-141) — the affine parts come from qcb_linear."""
        am, ph = self.rbm_am, self.rbm_ph
        Ua, Up = am.weights_U.data.contiguous(), ph.weights_U.data.contiguous()
        m_am, mp_am = am._linear(Ua, v, am.aux_bias.data.contiguous()), am._linear(Ua, vp, am.aux_bias.data.contiguous())
        m_ph, mp_ph = ph._linear(Up, v, None), ph._linear(Up, vp, None)
        if expand and v.dim() >= 2:
            m_am, m_ph = m_am.unsqueeze(1), m_ph.unsqueeze(1)
        if expand and vp.dim() >= 2:
            mp_am, mp_ph = mp_am.unsqueeze(0), mp_ph.unsqueeze(0)
        return (m_am + mp_am) / 2, (m_ph - mp_ph) / 2

    def pi(self, v, vp, expand=True):
        """Elements of the Pi matrix as (log-modulus, argument) (density_matrix.py:113-157)."""
        v = v.to(device=self.device, dtype=torch.double)
        vp = vp.to(device=self.device, dtype=torch.double)
        exp_arg, phase = self._mix_terms(v, vp, expand)
        real = (1 + 2 * exp_arg.exp() * phase.cos() + (2 * exp_arg).exp()).sqrt().log().sum(-1)
        imag = torch.atan2(exp_arg.exp() * phase.sin(), 1 + exp_arg.exp() * phase.cos()).sum(-1)
        return cplx.make_complex(real, imag)

    def pi_grad(self, v, vp, phase=False, expand=False):
        """Gradient of Pi w.r.t. the amplitude (phase=False) or phase RBM parameters
        (density_matrix.py:159-243); only the U (and, for the amplitude, d) blocks are non-zero."""
        unsqueezed = v.dim() < 2 or vp.dim() < 2
        v = (v.unsqueeze(0) if v.dim() < 2 else v).to(device=self.device, dtype=torch.double)
        vp = (vp.unsqueeze(0) if vp.dim() < 2 else vp).to(device=self.device, dtype=torch.double)
        if expand:
            arg_real, arg_imag = self._mix_terms(v, vp, True)
        else:
            arg_real = self.rbm_am.mixing_term(v + vp)
            arg_imag = self.rbm_ph.mixing_term(v - vp)
        sig = cplx.sigmoid(arg_real, arg_imag)
        batch = (v.shape[0], vp.shape[0], *v.shape[1:-1]) if expand else tuple(v.shape[:-1])
        am = self.rbm_am
        kw = dict(dtype=torch.double, device=self.device)
        W_zero = torch.zeros(*batch, am.num_hidden * am.num_visible, **kw)
        vb_zero = torch.zeros(*batch, am.num_visible, **kw)
        hb_zero = torch.zeros(*batch, am.num_hidden, **kw)
        if phase:
            temp = (v.unsqueeze(1) - vp.unsqueeze(0)) if expand else (v - vp)
            sig = cplx.scalar_mult(sig, cplx.I)
            ab_real = torch.zeros(*batch, am.num_aux, **kw)
            ab_imag = ab_real.clone()
        else:
            temp = (v.unsqueeze(1) + vp.unsqueeze(0)) if expand else (v + vp)
            ab_real, ab_imag = cplx.real(sig), cplx.imag(sig)
        U_grad = 0.5 * torch.einsum("c...j,...k->c...jk", sig, temp)
        parts_r = [W_zero, cplx.real(U_grad).reshape(*batch, -1), vb_zero, hb_zero, ab_real]
        parts_i = [W_zero.clone(), cplx.imag(U_grad).reshape(*batch, -1), vb_zero.clone(), hb_zero.clone(), ab_imag]
        if unsqueezed and not expand:
            parts_r = [p.squeeze(0) for p in parts_r]
            parts_i = [p.squeeze(0) for p in parts_i]
        return cplx.make_complex(torch.cat(parts_r, dim=-1), torch.cat(parts_i, dim=-1))

    def rho(self, v, vp=None, expand=True):
        """Unnormalised density-matrix elements rho~(v, v') (density_matrix.py:245-273)."""
        if expand is False and vp is None:
            return cplx.make_complex(self.probability(v))
        if vp is None:
            vp = v
        v = v.to(device=self.device, dtype=torch.double)
        vp = vp.to(device=self.device, dtype=torch.double)
        pi_ = self.pi(v, vp, expand=expand)
        amp = (self.rbm_am.gamma(v, vp, eta=+1, expand=expand) + cplx.real(pi_)).exp()
        phase = self.rbm_ph.gamma(v, vp, eta=-1, expand=expand) + cplx.imag(pi_)
        return cplx.make_complex(amp * phase.cos(), amp * phase.sin())

    def importance_sampling_numerator(self, vp, v):
        return self.rho(vp, v, expand=False)

    def importance_sampling_denominator(self, v):
        return cplx.make_complex(self.probability(v))

    # fused rotated-basis path ------------------------------------------------------------------
    def _rotated_probs(self, basis, states, unitaries_=None):
        return unitaries._call_dm(self.rbm_am, self.rbm_ph, basis, states, unitaries_ if unitaries_ else self.unitary_dict)

    def _accumulate_rotated_gradient(self, basis, samples, g_am, g_ph):
        unitaries._call_dm(self.rbm_am, self.rbm_ph, basis, samples, self.unitary_dict, g_am, g_ph)

    def rotated_gradient(self, basis, sample):
        """[g_am, g_ph] of samples measured in ``basis`` (density_matrix.py:281-304) in one fused call."""
        sample = sample.to(device=self.device, dtype=torch.double)
        if sample.dim() < 2:
            sample = sample.unsqueeze(0)
        g = [torch.zeros(self.rbm_am.num_pars, dtype=torch.double, device=self.device),
             torch.zeros(self.rbm_ph.num_pars, dtype=torch.double, device=self.device)]
        self._accumulate_rotated_gradient(basis, sample, g[0], g[1])
        return g

    def am_grads(self, v):
        """density_matrix.py:306-317."""
        return self.rbm_am.gamma_grad(v, v, eta=+1, expand=True) + self.pi_grad(v, v, phase=False, expand=True)

    def ph_grads(self, v):
        """density_matrix.py:319-330."""
        return cplx.scalar_mult(self.rbm_ph.gamma_grad(v, v, eta=-1, expand=True), cplx.I) + self.pi_grad(v, v, phase=True, expand=True)

    def fit(self, data, epochs=100, pos_batch_size=100, neg_batch_size=None, k=1, lr=1, input_bases=None, progbar=False,
            starting_epoch=1, time=False, callbacks=None, optimizer=torch.optim.SGD, optimizer_args=None, scheduler=None,
            scheduler_args=None, **kwargs):
        if input_bases is None:
            raise ValueError("input_bases must be provided to train a DensityMatrix!")
        return super().fit(data=data, epochs=epochs, pos_batch_size=pos_batch_size, neg_batch_size=neg_batch_size, k=k, lr=lr,
                           input_bases=input_bases, progbar=progbar, starting_epoch=starting_epoch, time=time, callbacks=callbacks,
                           optimizer=optimizer, optimizer_args=optimizer_args, scheduler=scheduler, scheduler_args=scheduler_args,
                           **kwargs)

    @staticmethod
    def autoload(location, gpu=True):
        state_dict = load_checkpoint(location)
        nn_state = DensityMatrix(unitary_dict=state_dict["unitary_dict"],
                                 num_visible=len(state_dict["rbm_am"]["visible_bias"]),
                                 num_hidden=len(state_dict["rbm_am"]["hidden_bias"]),
                                 num_aux=len(state_dict["rbm_am"]["aux_bias"]), gpu=gpu)
        nn_state.load(location)
        return nn_state
