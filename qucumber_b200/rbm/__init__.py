from .binary_rbm import BinaryRBM  # noqa: F401
