"""Times qcb_rbm_gibbs against the number of sweeps per launch (fixed cost of a launch vs cost per sweep).

usage: python tools/gibbs_k_sweep.py [chains_log2]   (prints one JSON line per (mode, k))"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qucumber_b200 import BinaryRBM, _lib  # noqa: E402

dev = torch.device("cuda")
C = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 20)
n = 64
g = torch.Generator().manual_seed(7)
rbm = BinaryRBM(n, n, gpu=True)
with torch.no_grad():
    rbm.weights.copy_(torch.randn(n, n, dtype=torch.double, generator=g) / np.sqrt(n))
v = (torch.rand(C, n, device=dev) < 0.5).double()
lib = _lib.load()
for mode in ((2, 3) if "--filters" in sys.argv else (0, 1)):
    lib.qcb_set_gibbs_mode(mode)
    for k in ((1, 2, 3, 4) if "--filters" in sys.argv else (1, 2, 5, 10, 20, 50)):
        for _ in range(3):
            rbm.gibbs_steps(k, v, overwrite=True, seed=1)
        reps = 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            rbm.gibbs_steps(k, v, overwrite=True, seed=1)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(json.dumps({"mode": mode, "k": k, "ms": round(ms, 4), "sweeps_per_s": C * k / ms * 1e3}))
lib.qcb_set_gibbs_mode(0)
