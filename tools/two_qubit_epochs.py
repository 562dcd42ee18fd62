"""Epoch times of the Tutorial-2 / Tutorial-3 style trainings (ComplexWaveFunction and DensityMatrix on 2 qubits)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qucumber_b200 import ComplexWaveFunction, DensityMatrix  # noqa: E402

rng = np.random.default_rng(3)
for name, make, n_samples, lr in (("complex", lambda: ComplexWaveFunction(2, 2, gpu=True), 500, 0.1),
                                  ("density", lambda: DensityMatrix(2, 2, 2, gpu=True), 900, 0.1)):
    bases_all = ["XZ", "ZX", "YZ", "ZY", "ZZ"] if name == "complex" else ["XX", "XY", "XZ", "YX", "YY", "YZ", "ZX", "ZY", "ZZ"]
    bases = np.array([list(bases_all[i % len(bases_all)]) for i in range(n_samples)])
    data = torch.tensor((rng.random((n_samples, 2)) < 0.5).astype(np.float64))
    torch.manual_seed(1)
    nn = make()
    nn.fit(data, epochs=2, pos_batch_size=100, neg_batch_size=100, k=10, lr=lr, input_bases=bases)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    nn.fit(data, epochs=10, pos_batch_size=100, neg_batch_size=100, k=10, lr=lr, input_bases=bases)
    torch.cuda.synchronize()
    print(f"{name}: {(time.perf_counter() - t0) / 10 * 1e3:.2f} ms/epoch ({n_samples} samples, {len(bases_all)} bases)")
