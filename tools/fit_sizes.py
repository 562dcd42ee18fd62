"""Epoch time of fit() (10k samples, pos 100 / neg 200, CD-10, SGD) against the layer size, fused epoch vs per-batch path."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qucumber_b200 import PositiveWaveFunction  # noqa: E402

for n in (10, 16, 24, 32, 48, 64):
    data = (torch.rand(10000, n) < 0.5).double()
    row = []
    class BatchHook:  # any per-batch callback switches the epoch fusion off
        n = 0

        def on_batch_end(self, s, ep, b):
            self.n += 1

    for fuse, cbs in ((True, None), (True, [BatchHook()]), (False, [BatchHook()]), (False, None)):
        torch.manual_seed(0)
        nn = PositiveWaveFunction(n, n, gpu=True)
        nn.fuse_epochs = fuse
        nn.fit(data, epochs=2, pos_batch_size=100, neg_batch_size=200, k=10, lr=0.01, callbacks=cbs)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        nn.fit(data, epochs=10, pos_batch_size=100, neg_batch_size=200, k=10, lr=0.01, callbacks=cbs)
        torch.cuda.synchronize()
        row.append((time.perf_counter() - t0) / 10 * 1e3)
    print(f"N=nh={n}: one launch per epoch {row[0]:.2f} ms/epoch, with a per-batch callback (optimizer.step path): one launch per batch {row[1]:.2f} vs separate kernels {row[2]:.2f}; separate kernels, fused SGD {row[3]:.2f}")
