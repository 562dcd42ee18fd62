"""Launch the single-spin-flip kernel a few times on 2^20 samples (for ncu). Usage: python tools/flip_profile.py [N]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qucumber_b200 import BinaryRBM  # noqa: E402
from qucumber_b200.observables import flip_logratio  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
rbm = BinaryRBM(n, n, gpu=True)
v = (torch.rand(1 << 20, n, device="cuda:0") < 0.5).double()
for _ in range(3):
    out = flip_logratio(rbm, v)
torch.cuda.synchronize()
print(float(out.sum()))
