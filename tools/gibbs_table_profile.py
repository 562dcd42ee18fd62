"""Two launches of the table Gibbs kernel at the Tutorial-4 shape (2^20 chains, N = nh = 10, k = 100): ncu --set full target."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qucumber_b200 import BinaryRBM  # noqa: E402

rbm = BinaryRBM(10, 10, gpu=True)
v = (torch.rand(1 << 20, 10, device="cuda") < 0.5).double()
for _ in range(2):
    rbm.gibbs_steps(100, v, overwrite=True, seed=7)
torch.cuda.synchronize()
print("done")
