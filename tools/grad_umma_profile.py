"""A few launches of the tcgen05 reduced gradient at N = nh = 64, 2^20 binary rows: ncu capture target."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qucumber_b200 import BinaryRBM  # noqa: E402

n = 64
rb = BinaryRBM(n, n, gpu=True)
g = torch.Generator().manual_seed(99)
with torch.no_grad():
    rb.weights.copy_(torch.randn(n, n, dtype=torch.double, generator=g) / np.sqrt(n))
v = (torch.rand(1 << 20, n, generator=g) < 0.5).double().cuda()
for _ in range(3):
    rb.effective_energy_gradient(v)
torch.cuda.synchronize()
print("done")
