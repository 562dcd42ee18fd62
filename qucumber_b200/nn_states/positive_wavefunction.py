"""Mirror of ``qucumber.nn_states.PositiveWaveFunction`` (qucumber/nn_states/positive_wavefunction.py:24-240)."""
import torch

from ..rbm import BinaryRBM
from .neural_state import load_checkpoint, stored_attribute
from .wavefunction import WaveFunctionBase


class PositiveWaveFunction(WaveFunctionBase):
    """A wavefunction with real, positive coefficients: one BinaryRBM for the amplitude."""

    _rbm_am = None
    _device = None
    rbm_am = stored_attribute("_rbm_am")
    device = stored_attribute("_device")

    def __init__(self, num_visible, num_hidden=None, gpu=True, module=None):
        if module is None:
            self.rbm_am = BinaryRBM(num_visible, num_hidden, gpu=gpu)
        else:
            self.rbm_am = module.to(torch.device("cuda", torch.cuda.current_device()) if (gpu and torch.cuda.is_available()) else torch.device("cpu"))
            self.rbm_am.device = next(self.rbm_am.parameters()).device
        self.num_visible = self.rbm_am.num_visible
        self.num_hidden = self.rbm_am.num_hidden
        self.device = self.rbm_am.device

    @property
    def networks(self):
        return ["rbm_am"]

    def phase(self, v):
        """Identically zero (positive_wavefunction.py:95-106)."""
        if v.dim() == 1:
            v = v.unsqueeze(0)
            return torch.zeros(v.shape[:-1], dtype=torch.double, device=self.device).squeeze(0)
        return torch.zeros(v.shape[:-1], dtype=torch.double, device=self.device)

    def gradient(self, v, *args, **kwargs):
        """Gradient of the amplitude RBM only; bases are ignored (positive_wavefunction.py:126-146)."""
        return super().gradient(v, bases=None)

    def positive_phase_gradients(self, samples_batch, *args, **kwargs):
        """Positive phase of the gradients; extra arguments are ignored (positive_wavefunction.py:142-154)."""
        return super().positive_phase_gradients(samples_batch, bases_batch=None)

    def compute_batch_gradients(self, k, samples_batch, neg_batch, *args, **kwargs):
        return super().compute_batch_gradients(k, samples_batch, neg_batch, bases_batch=None)

    def compute_exact_gradients(self, samples_batch, space=None, *args, **kwargs):
        return super().compute_exact_gradients(samples_batch, space, bases_batch=None)

    def compute_exact_grads(self, samples_batch, space=None, *args, **kwargs):
        """Name used by the reference's subclass (positive_wavefunction.py:156-171); same as compute_exact_gradients."""
        return self.compute_exact_gradients(samples_batch, space)

    def fit(self, data, epochs=100, pos_batch_size=100, neg_batch_size=None, k=1, lr=1e-3, progbar=False, starting_epoch=1,
            time=False, callbacks=None, optimizer=torch.optim.SGD, optimizer_args=None, scheduler=None, scheduler_args=None,
            **kwargs):
        """Bases are not used for a positive wavefunction (positive_wavefunction.py:194-229)."""
        kwargs["input_bases"] = None
        return super().fit(data=data, epochs=epochs, pos_batch_size=pos_batch_size, neg_batch_size=neg_batch_size, k=k, lr=lr,
                           progbar=progbar, starting_epoch=starting_epoch, time=time, callbacks=callbacks, optimizer=optimizer,
                           optimizer_args=optimizer_args, scheduler=scheduler, scheduler_args=scheduler_args, **kwargs)

    @staticmethod
    def autoload(location, gpu=True):
        state_dict = load_checkpoint(location)
        wvfn = PositiveWaveFunction(num_visible=len(state_dict["rbm_am"]["visible_bias"]),
                                    num_hidden=len(state_dict["rbm_am"]["hidden_bias"]), gpu=gpu)
        wvfn.load(location)
        return wvfn
