"""Time the single-spin-flip kernel and the SigmaZ / SigmaX / TFIM-energy estimators on 2^20 samples
(config 3: N = nh = 10; also 64 x 64). Usage: python tools/observables_time.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qucumber_b200 import PositiveWaveFunction  # noqa: E402
from qucumber_b200.observables import SigmaX, SigmaZ, TFIMChainEnergy, flip_logratio  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


dev = torch.device("cuda:0")
for n in (10, 64):
    nn = PositiveWaveFunction(n, n, gpu=True)
    if n == 10:
        g = np.load(os.path.join(ROOT, "tests", "golden", "binary_rbm_tfim10.npz"))
        with torch.no_grad():
            nn.rbm_am.weights.copy_(torch.tensor(g["weights"])); nn.rbm_am.visible_bias.copy_(torch.tensor(g["visible_bias"]))
            nn.rbm_am.hidden_bias.copy_(torch.tensor(g["hidden_bias"]))
    v = (torch.rand(1 << 20, n, device=dev) < 0.5).double()
    s = nn.rbm_am.gibbs_steps(20, v, seed=3)
    print(f"N={n}: flip kernel {timed(lambda: flip_logratio(nn.rbm_am, s)):.3f} ms; "
          f"SigmaZ {timed(lambda: SigmaZ(absolute=True).statistics_from_samples(nn, s)):.3f} ms; "
          f"SigmaX {timed(lambda: SigmaX().statistics_from_samples(nn, s)):.3f} ms; "
          f"TFIM energy {timed(lambda: TFIMChainEnergy(1.0).statistics_from_samples(nn, s)):.3f} ms")
