"""Lazy stand-in for the (2^N, N) tensor of ``generate_hilbert_space`` (reference:
qucumber/nn_states/neural_state.py:156-176). The CUDA enumeration kernels generate basis states
from their index, so the full space is only a size; ``tensor()`` materialises rows on demand for
callers that really need them (same row order: column 0 is the most significant bit)."""
import torch


class HilbertSpace:
    def __init__(self, num_visible, device=None):
        self.num_visible = int(num_visible)
        self.device = device

    @property
    def shape(self):
        return (1 << self.num_visible, self.num_visible)

    def __len__(self):
        return 1 << self.num_visible

    def tensor(self, first=0, count=None, device=None):
        n = self.num_visible
        count = (1 << n) - first if count is None else count
        dev = device if device is not None else self.device
        idx = torch.arange(first, first + count, dtype=torch.int64, device=dev)
        shifts = torch.arange(n - 1, -1, -1, dtype=torch.int64, device=dev)
        return ((idx[:, None] >> shifts[None, :]) & 1).to(torch.double)
