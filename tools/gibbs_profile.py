"""Launches of a Gibbs kernel at benchmark shape (target of an ncu --set full capture): gibbs_profile.py [k] [mode]
(mode 0: mma.sync integer filter, 1: FP64 DMMA, 2: tcgen05 / TMEM integer filter)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qucumber_b200 import BinaryRBM, _lib  # noqa: E402

dev = torch.device("cuda")
k = int(sys.argv[1]) if len(sys.argv) > 1 else 10
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
_lib.load().qcb_set_gibbs_mode(mode)
n = 64
g = torch.Generator().manual_seed(7)
rbm = BinaryRBM(n, n, gpu=True)
with torch.no_grad():
    rbm.weights.copy_(torch.randn(n, n, dtype=torch.double, generator=g) / np.sqrt(n))
v = (torch.rand(1 << 20, n, device=dev) < 0.5).double()
for _ in range(2):
    rbm.gibbs_steps(k, v, overwrite=True, seed=1)
torch.cuda.synchronize()
print("done")
