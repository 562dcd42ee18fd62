"""Times log Z and the fidelity / KL sums (qcb_rbm_exact_stats_range) over the full space."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qucumber_b200 import PositiveWaveFunction  # noqa: E402
from qucumber_b200.utils import training_statistics as ts  # noqa: E402

dev = torch.device("cuda")
for n in [int(x) for x in sys.argv[1:]] or (20, 24, 26, 28):
    torch.manual_seed(1234)
    nn = PositiveWaveFunction(n, n, gpu=True)
    g = torch.Generator(device="cuda").manual_seed(5)
    t = torch.zeros(2, 1 << n, dtype=torch.double, device=dev)
    t[0].uniform_(generator=g)
    t /= t[0].norm()

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    ms_z = timed(lambda: nn.rbm_am.log_partition())
    ms_s = timed(lambda: ts._positive_stats(nn, t))
    ms_g = timed(lambda: nn.compute_exact_gradients(None, None) if False else None) if False else 0.0
    print(f"N={n}: logZ {ms_z:.3f} ms ({(1 << n) / ms_z / 1e6:.1f} G states/s), logZ + fidelity/KL sums {ms_s:.3f} ms "
          f"({(1 << n) / max(ms_s - ms_z, 1e-9) / 1e6:.1f} G states/s for the stats pass)")
