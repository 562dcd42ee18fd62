"""Host-side profile of the TFIM N=10 training loop (where do the microseconds per CD step go?)."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qucumber_b200 import PositiveWaveFunction  # noqa: E402

gold = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "binary_rbm_tfim10.npz"))
train = torch.tensor(np.unpackbits(gold["train_packed"], axis=1)[:, :10].astype(np.float64))
nn = PositiveWaveFunction(10, 10, gpu=True)
nn.fit(train, epochs=2, pos_batch_size=100, neg_batch_size=200, k=10, lr=0.01)
torch.cuda.synchronize()
t0 = time.perf_counter()
nn.fit(train, epochs=10, pos_batch_size=100, neg_batch_size=200, k=10, lr=0.01)
torch.cuda.synchronize()
print("ms/epoch", (time.perf_counter() - t0) / 10 * 1e3)
pr = cProfile.Profile()
pr.enable()
nn.fit(train, epochs=3, pos_batch_size=100, neg_batch_size=200, k=10, lr=0.01)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
