"""Host-side cProfile of the two-qubit trainings (Tutorial-2 / Tutorial-3 style): where the per-batch Python time goes."""
import cProfile
import os
import pstats
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qucumber_b200 import ComplexWaveFunction, DensityMatrix
rng = np.random.default_rng(3)
bases_all = ["XX", "XY", "XZ", "YX", "YY", "YZ", "ZX", "ZY", "ZZ"]
bases = np.array([list(bases_all[i % 9]) for i in range(900)])
data = torch.tensor((rng.random((900, 2)) < 0.5).astype(np.float64))
torch.manual_seed(1)
nn = DensityMatrix(2, 2, 2, gpu=True)
nn.fit(data, epochs=2, pos_batch_size=100, neg_batch_size=100, k=10, lr=0.1, input_bases=bases)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
nn.fit(data, epochs=10, pos_batch_size=100, neg_batch_size=100, k=10, lr=0.1, input_bases=bases)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
