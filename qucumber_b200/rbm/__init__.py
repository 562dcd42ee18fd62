from .binary_rbm import BinaryRBM  # noqa: F401
from .purification_rbm import PurificationRBM  # noqa: F401
