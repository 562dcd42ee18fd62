"""Splits the rotated-basis gradient time (ComplexWaveFunction N = nh = 20, B = 4096) into the amplitude pass and the
two weighted reductions over the expansion."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qucumber_b200 import BinaryRBM, ComplexWaveFunction  # noqa: E402
from qucumber_b200.utils import unitaries as un  # noqa: E402

dev = torch.device("cuda:0")
n, B = 20, 4096
nn = ComplexWaveFunction(n, n, gpu=True)
samples = (torch.rand(B, n, device=dev) < 0.5).double()


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


for S in (4, 8, 10):
    basis = np.array(["X"] * S + ["Z"] * (n - S))
    t_amp = timed(lambda: un.rotate_psi_inner_prod(nn, basis, samples))
    t_grad = timed(lambda: nn.rotated_gradient(basis, samples))
    rows = B << S
    v = (torch.rand(rows, n, device=dev) < 0.5).double()
    rb = BinaryRBM(n, n, gpu=True)
    t_red = timed(lambda: rb.effective_energy_gradient(v))
    print(f"S={S}: amplitudes {t_amp:.3f} ms, amplitudes + gradient {t_grad:.3f} ms, one explicit-row reduction of {rows} rows {t_red:.3f} ms")
