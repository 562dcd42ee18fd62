"""The two RBM modules of the drop-in boundary: same class names, constructor arguments, method names and
``state_dict`` keys as ``qucumber.rbm`` (binary_rbm.py, purification_rbm.py); every method body is a call into
``libqucumber_b200.so`` (include/qcb.h). CUDA only."""
from . import binary_rbm as _binary
from . import purification_rbm as _purification

BinaryRBM = _binary.BinaryRBM
PurificationRBM = _purification.PurificationRBM

__all__ = ["BinaryRBM", "PurificationRBM"]
