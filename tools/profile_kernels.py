"""Runs the non-Gibbs hot kernels once each at benchmark shapes (target of an ncu --set full capture)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qucumber_b200 import BinaryRBM  # noqa: E402

dev = torch.device("cuda")
g = torch.Generator().manual_seed(1234)
n = 28
rbm = BinaryRBM(n, n, gpu=True)
with torch.no_grad():
    rbm.weights.copy_(torch.randn(n, n, dtype=torch.double, generator=g) / np.sqrt(n))
for _ in range(2):
    rbm.log_partition()
big = BinaryRBM(64, 64, gpu=True)
v = (torch.rand(1 << 20, 64, device=dev) < 0.5).double()
for _ in range(2):
    big.effective_energy_gradient(v)
    big.effective_energy(v)
torch.cuda.synchronize()
print("done")
