"""One rotated-basis gradient call at the benchmark shape (for an ncu launch list)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qucumber_b200 import ComplexWaveFunction, DensityMatrix  # noqa: E402

dev = torch.device("cuda")
rng = np.random.default_rng(7)
n, S, B = 20, 10, 4096
cw = ComplexWaveFunction(n, n, gpu=True)
sites = np.sort(rng.choice(n, S, replace=False))
basis = np.array(["Z"] * n)
basis[sites] = rng.choice(["X", "Y"], S)
samples = torch.tensor((rng.random((B, n)) < 0.5).astype(np.float64), device=dev)
for _ in range(2):
    cw.rotated_gradient(basis, samples)
torch.cuda.synchronize()
dm = DensityMatrix(6, 6, 6, gpu=True)
basis6 = np.array(list("XYZXZY"))
s6 = torch.tensor((rng.random((4096, 6)) < 0.5).astype(np.float64), device=dev)
for _ in range(2):
    dm.rotated_gradient(basis6, s6)
torch.cuda.synchronize()
print("done")
