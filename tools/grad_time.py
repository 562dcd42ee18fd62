"""Times the reduced-gradient path (effective_energy_gradient) at the benchmark shapes and the energy kernel."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qucumber_b200 import BinaryRBM  # noqa: E402

dev = torch.device("cuda")
rows = 1 << 20
for nv, nh in ((64, 64), (32, 32), (20, 20), (10, 10)):
    rb = BinaryRBM(nv, nh, gpu=True)
    g = torch.Generator().manual_seed(99)
    with torch.no_grad():
        rb.weights.copy_(torch.randn(nh, nv, dtype=torch.double, generator=g) / np.sqrt(nv))
    v = (torch.rand(rows, nv, generator=g) < 0.5).double().to(dev)
    for name, fn in (("grad", lambda: rb.effective_energy_gradient(v)), ("energy", lambda: rb.effective_energy(v))):
        fn()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(f"N={nv} nh={nh} {name}: {best:.3f} ms")
