"""Flat gradient vector -> ``param.grad`` views (reference: qucumber/utils/gradients_utils.py:21-52)."""
import torch


def vector_to_grads(vec, parameters):
    """Slice ``vec`` in parameter order and install the pieces as the parameters' ``.grad``."""
    if not isinstance(vec, torch.Tensor):
        raise TypeError(f"expected torch.Tensor, but got: {torch.typename(vec)}")
    offset = 0
    device = None
    for param in parameters:
        if device is None:
            device = param.device
        elif param.device != device:
            raise TypeError("Found two parameters on different devices, this is currently not supported.")
        n = param.numel()
        param.grad = vec[offset:offset + n].view(param.size()).data
        offset += n
