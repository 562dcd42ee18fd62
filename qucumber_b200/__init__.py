"""qucumber_b200 — B200 (sm_100a) backend for QuCumber's RBM training and sampling hot path.

Host code is PyTorch (device memory, streams, torch.distributed); the arithmetic is
hand-written CUDA in libqucumber_b200.so behind the C ABI of include/qcb.h.
"""
from .hilbert import HilbertSpace  # noqa: F401
from .rbm import BinaryRBM, PurificationRBM  # noqa: F401
from .rbm._sampling import pack_states, unpack_states  # noqa: F401
from .rng import set_random_seed  # noqa: F401
from . import nn_states, utils  # noqa: F401
from .nn_states import ComplexWaveFunction, DensityMatrix, PositiveWaveFunction  # noqa: F401

__version__ = "0.1.0"
