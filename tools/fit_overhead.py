"""Where the wall clock of a fused-epoch fit() goes (host-side shuffle vs kernel)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qucumber_b200 import PositiveWaveFunction  # noqa: E402

torch.manual_seed(0)
nn = PositiveWaveFunction(10, 10, gpu=True)
data = (torch.rand(10000, 10) < 0.5).double()
nn.fit(data, epochs=5, pos_batch_size=100, neg_batch_size=200, k=10, lr=0.01)
torch.cuda.synchronize()
for epochs in (50, 200):
    t0 = time.perf_counter()
    nn.fit(data, epochs=epochs, pos_batch_size=100, neg_batch_size=200, k=10, lr=0.01)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"epochs={epochs}: host returned after {(t1 - t0) / epochs * 1e3:.3f} ms/epoch, GPU done after {(t2 - t0) / epochs * 1e3:.3f} ms/epoch")
train = data.cuda()
t0 = time.perf_counter()
for _ in range(200):
    pos, neg, _ = nn._shuffle_tensors(100, 200, 100, train, None, None)
torch.cuda.synchronize()
print(f"shuffle alone: {(time.perf_counter() - t0) / 200 * 1e3:.3f} ms/epoch")
