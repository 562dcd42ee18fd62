"""Two epochs of the Tutorial-2 / Tutorial-3 style trainings, for a launch list:
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/two_qubit_launches.py density"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qucumber_b200 import ComplexWaveFunction, DensityMatrix  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "complex"
rng = np.random.default_rng(3)
bases_all = ["XZ", "ZX", "YZ", "ZY", "ZZ"] if kind == "complex" else ["XX", "XY", "XZ", "YX", "YY", "YZ", "ZX", "ZY", "ZZ"]
n_samples = 500 if kind == "complex" else 900
bases = np.array([list(bases_all[i % len(bases_all)]) for i in range(n_samples)])
data = torch.tensor((rng.random((n_samples, 2)) < 0.5).astype(np.float64))
torch.manual_seed(1)
nn = ComplexWaveFunction(2, 2, gpu=True) if kind == "complex" else DensityMatrix(2, 2, 2, gpu=True)
nn.fit(data, epochs=2, pos_batch_size=100, neg_batch_size=100, k=10, lr=0.1, input_bases=bases)
torch.cuda.synchronize()
