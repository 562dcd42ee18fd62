"""Host-side mirror of ``qucumber.nn_states`` (control flow only: batching, grouping by basis, optimizer, callbacks);
the arithmetic underneath runs in the CUDA library. Class names and signatures are the reference's."""
from . import complex_wavefunction as _cwf
from . import density_matrix as _dm
from . import neural_state as _base
from . import positive_wavefunction as _pwf
from . import wavefunction as _wf

NeuralStateBase = _base.NeuralStateBase
WaveFunctionBase = _wf.WaveFunctionBase
PositiveWaveFunction = _pwf.PositiveWaveFunction
ComplexWaveFunction = _cwf.ComplexWaveFunction
DensityMatrix = _dm.DensityMatrix

__all__ = ["NeuralStateBase", "WaveFunctionBase", "PositiveWaveFunction", "ComplexWaveFunction", "DensityMatrix"]
