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

# ---- the kernels added later in round 1: fused CD epoch, product-form fidelity/KL sums, full-space rotation, rotated gradient
from qucumber_b200 import ComplexWaveFunction, PositiveWaveFunction, _lib  # noqa: E402
from qucumber_b200.utils import training_statistics as ts  # noqa: E402
from qucumber_b200.utils.unitaries import rotate_psi  # noqa: E402

torch.manual_seed(0)
small = PositiveWaveFunction(10, 10, gpu=True)
data = (torch.rand(10000, 10) < 0.5).double()
small.fit(data, epochs=2, pos_batch_size=100, neg_batch_size=200, k=10, lr=0.01)
n = 26
st = PositiveWaveFunction(n, n, gpu=True)
t = torch.zeros(2, 1 << n, dtype=torch.double, device=dev)
t[0] = torch.rand(1 << n, dtype=torch.double, device=dev)
t /= t[0].norm()
for _ in range(2):
    ts._positive_stats(st, t)
cw = ComplexWaveFunction(20, 20, gpu=True)
psi = torch.randn(2, 1 << 20, dtype=torch.double, device=dev)
for _ in range(2):
    rotate_psi(cw, "XYZXYZXYZXYZXYZXYZXY", None, psi=psi)
rng = np.random.default_rng(7)
basis = np.array(["Z"] * 20)
basis[np.sort(rng.choice(20, 10, replace=False))] = "X"
samples = torch.tensor((rng.random((4096, 20)) < 0.5).astype(np.float64), device=dev)
for _ in range(2):
    cw.rotated_gradient(basis, samples)
torch.cuda.synchronize()
print("done (late kernels)")
