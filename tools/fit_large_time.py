"""Epoch time of fit() on a large binary data set (2^21 samples x 64 units, two batches of 2^20 rows and chains, CD-10, SGD):
packed training data (qcb_rbm_cd_step_packed, 8-byte rows) against float64 rows (qcb_rbm_cd_step, 512-byte rows)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qucumber_b200 import PositiveWaveFunction  # noqa: E402

dev = torch.device("cuda")
gen = torch.Generator(device="cuda").manual_seed(3)
data = (torch.rand(1 << 21, 64, device=dev, generator=gen) < 0.5).double()
for packed in (True, False, True, False):
    torch.manual_seed(0)
    nn = PositiveWaveFunction(64, 64, gpu=True)
    nn.pack_training_data = packed
    nn.fit(data, epochs=1, pos_batch_size=1 << 20, neg_batch_size=1 << 20, k=10, lr=1e-3)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    nn.fit(data, epochs=5, pos_batch_size=1 << 20, neg_batch_size=1 << 20, k=10, lr=1e-3)
    torch.cuda.synchronize()
    print(f"packed={packed}: {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms per epoch (2 batches of 2^20 rows + 2^20 chains, CD-10), incl. the per-call set-up of fit()")
