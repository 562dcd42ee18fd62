"""Mirror of ``qucumber.nn_states.WaveFunctionBase`` (qucumber/nn_states/wavefunction.py:22-87)."""
import abc

from ..utils import cplx
from .neural_state import NeuralStateBase


class WaveFunctionBase(NeuralStateBase):
    def amplitude(self, v):
        """|psi(v)| = sqrt(exp(-E_am(v)))  (wavefunction.py:33-46)."""
        return (-self.rbm_am.effective_energy(v)).exp().sqrt()

    @abc.abstractmethod
    def phase(self, v):
        """Phase of psi(v)."""

    def psi(self, v):
        """Unnormalised psi(v) = |psi| (cos(phase) + i sin(phase)) as (2, …) (wavefunction.py:62-81)."""
        amplitude, phase = self.amplitude(v), self.phase(v)
        return cplx.make_complex(amplitude * phase.cos(), amplitude * phase.sin())

    def importance_sampling_numerator(self, vp, v):
        return self.psi(vp)

    def importance_sampling_denominator(self, v):
        return self.psi(v)
