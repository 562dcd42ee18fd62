"""Mirror of ``qucumber.nn_states.ComplexWaveFunction`` (qucumber/nn_states/complex_wavefunction.py:26-266)."""
import torch

from ..rbm import BinaryRBM
from ..utils import unitaries
from .neural_state import load_checkpoint, stored_attribute
from .wavefunction import WaveFunctionBase


class ComplexWaveFunction(WaveFunctionBase):
    """A wavefunction with a phase: one BinaryRBM for the amplitude, one for the phase. Unlike the
    reference (which warns that its GPU path is slower than its CPU path because the rotated-basis
    code round-trips through NumPy), the rotated-basis amplitude and gradient are fused CUDA kernels."""

    _rbm_am = None
    _rbm_ph = None
    _device = None
    rbm_am = stored_attribute("_rbm_am")
    rbm_ph = stored_attribute("_rbm_ph")
    device = stored_attribute("_device")
    _grouped_call = "cwf"
    _rotated_limits = (("num_visible", 512, "num_visible"),)  # larger hidden layers take the row-tiled kernels (csrc/large.cu)
    _max_rotated_sites = 21

    def __init__(self, num_visible, num_hidden=None, unitary_dict=None, gpu=True, module=None):
        if module is not None:
            raise ValueError("module= injection is not supported for two-RBM states (it is broken in the reference too)")
        self.rbm_am = BinaryRBM(num_visible, num_hidden, gpu=gpu)
        self.rbm_ph = BinaryRBM(num_visible, num_hidden, gpu=gpu)
        self.num_visible = self.rbm_am.num_visible
        self.num_hidden = self.rbm_am.num_hidden
        self.device = self.rbm_am.device
        self.unitary_dict = unitary_dict if unitary_dict else unitaries.create_dict()
        self.unitary_dict = {k: v.to(device=self.device) for k, v in self.unitary_dict.items()}

    @property
    def networks(self):
        return ["rbm_am", "rbm_ph"]

    def phase(self, v):
        """-E_ph(v) / 2 (complex_wavefunction.py:129-142)."""
        return -0.5 * self.rbm_ph.effective_energy(v)

    # fused rotated-basis path ------------------------------------------------------------------
    def _rotated_amplitudes(self, basis, states, unitaries_=None):
        return unitaries._call_cwf(self.rbm_am, self.rbm_ph, basis, states, unitaries_ if unitaries_ else self.unitary_dict)

    def _accumulate_rotated_gradient(self, basis, samples, g_am, g_ph):
        unitaries._call_cwf(self.rbm_am, self.rbm_ph, basis, samples, self.unitary_dict, g_am, g_ph)

    def rotated_gradient(self, basis, sample):
        """[g_am, g_ph] of the samples measured in ``basis`` (complex_wavefunction.py:162-186), one
        fused CUDA call instead of materialising (2, 2^S, B, G) gradient tensors."""
        sample = sample.to(device=self.device, dtype=torch.double)
        if sample.dim() < 2:
            sample = sample.unsqueeze(0)
        g = [torch.zeros(self.rbm_am.num_pars, dtype=torch.double, device=self.device),
             torch.zeros(self.rbm_ph.num_pars, dtype=torch.double, device=self.device)]
        self._accumulate_rotated_gradient(basis, sample, g[0], g[1])
        return g

    def am_grads(self, v):
        """Per-state amplitude-RBM gradients as a complex tensor (complex_wavefunction.py:188-197)."""
        from ..utils import cplx

        return cplx.make_complex(self.rbm_am.effective_energy_gradient(v, reduce=False))

    def ph_grads(self, v):
        """Per-state phase-RBM gradients times i (complex_wavefunction.py:199-211)."""
        from ..utils import cplx

        return cplx.scalar_mult(cplx.make_complex(self.rbm_ph.effective_energy_gradient(v, reduce=False)), cplx.I)

    def fit(self, data, epochs=100, pos_batch_size=100, neg_batch_size=None, k=1, lr=1e-3, input_bases=None, progbar=False,
            starting_epoch=1, time=False, callbacks=None, optimizer=torch.optim.SGD, optimizer_args=None, scheduler=None,
            scheduler_args=None, **kwargs):
        if input_bases is None:
            raise ValueError("input_bases must be provided to train a ComplexWaveFunction!")
        return super().fit(data=data, epochs=epochs, pos_batch_size=pos_batch_size, neg_batch_size=neg_batch_size, k=k, lr=lr,
                           input_bases=input_bases, progbar=progbar, starting_epoch=starting_epoch, time=time, callbacks=callbacks,
                           optimizer=optimizer, optimizer_args=optimizer_args, scheduler=scheduler, scheduler_args=scheduler_args,
                           **kwargs)

    @staticmethod
    def autoload(location, gpu=True):
        state_dict = load_checkpoint(location)
        wvfn = ComplexWaveFunction(unitary_dict=state_dict["unitary_dict"],
                                   num_visible=len(state_dict["rbm_am"]["visible_bias"]),
                                   num_hidden=len(state_dict["rbm_am"]["hidden_bias"]), gpu=gpu)
        wvfn.load(location)
        return wvfn
