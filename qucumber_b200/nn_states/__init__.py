from .neural_state import NeuralStateBase  # noqa: F401
from .wavefunction import WaveFunctionBase  # noqa: F401
from .positive_wavefunction import PositiveWaveFunction  # noqa: F401
from .complex_wavefunction import ComplexWaveFunction  # noqa: F401
from .density_matrix import DensityMatrix  # noqa: F401
