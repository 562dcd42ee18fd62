"""tcgen05 / TMEM Gibbs kernel (mode 2) against the mma.sync integer-filter kernel (mode 0): identical samples for every
I/O mode, then timings at N = nh = 64.  Run on the GPU box under a short `timeout` (a tcgen05 mistake can hang a kernel)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qucumber_b200 as qb  # noqa: E402
from qucumber_b200 import BinaryRBM, _lib  # noqa: E402

dev = torch.device("cuda")
lib = _lib.load()
out = {"equal": {}, "timing": []}


def make(n, nh, seed):
    g = torch.Generator().manual_seed(seed)
    rbm = BinaryRBM(n, nh, gpu=True)
    with torch.no_grad():
        rbm.weights.copy_(torch.randn(nh, n, dtype=torch.double, generator=g) / np.sqrt(n))
        rbm.visible_bias.copy_(0.3 * torch.randn(n, dtype=torch.double, generator=g))
        rbm.hidden_bias.copy_(0.3 * torch.randn(nh, dtype=torch.double, generator=g))
    return rbm


def both(fn):
    res = []
    for mode in (0, 2):
        lib.qcb_set_gibbs_mode(mode)
        res.append(fn())
    lib.qcb_set_gibbs_mode(0)
    torch.cuda.synchronize()
    return res


ok = True
for (n, nh) in [(64, 64), (40, 64), (64, 33), (50, 50), (33, 33)]:
    rbm = make(n, nh, 3)
    for C in (1, 100, 128, 129, 5000, 200_001):
        v0 = (torch.rand(C, n, device=dev) < 0.5).double()
        a, b = both(lambda: rbm.gibbs_steps(3, v0, seed=77, chain_id0=5, sweep0=2))
        e1 = torch.equal(a, b)
        a, b = both(lambda: rbm.sample_chains(2, C, seed=78, chain_id0=9, sweep0=1, packed=True))
        e2 = torch.equal(a, b)
        bits0 = qb.pack_states(v0)
        a, b = both(lambda: rbm.gibbs_steps_packed(2, bits0, seed=79))
        e3 = torch.equal(a, b)
        a, b = both(lambda: rbm.sample_chains(0, C, seed=80))
        e4 = torch.equal(a, b)
        out["equal"][f"{n}x{nh} C={C}"] = [e1, e2, e3, e4]
        ok = ok and e1 and e2 and e3 and e4
print(json.dumps({"all_equal": ok, "cases": out["equal"]}))
if not ok:
    sys.exit(2)

n = 64
rbm = make(n, n, 7)
C = 1 << 20
v = (torch.rand(C, n, device=dev) < 0.5).double()
for mode in (0, 2):
    lib.qcb_set_gibbs_mode(mode)
    for k in (1, 10, 100):
        for _ in range(3):
            rbm.gibbs_steps(k, v, overwrite=True, seed=1)
        reps = 10 if k < 100 else 3
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            rbm.gibbs_steps(k, v, overwrite=True, seed=1)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(json.dumps({"mode": mode, "k": k, "ms": round(ms, 4), "chain_sweeps_per_s": C * k / ms * 1e3}))
lib.qcb_set_gibbs_mode(0)
