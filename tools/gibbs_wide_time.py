"""Block Gibbs on layers of 65..128 units: the tcgen05 kernel (gibbs_umma_wide.cu, mode 0) against the path these shapes took
before (mode 1: FP64 DMMA kernel where one layer is <= 64 units, the composed affine + Bernoulli passes otherwise)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qucumber_b200 import BinaryRBM, _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda")
C = 1 << 20
import sys as _sys
NARROW = "--narrow" in _sys.argv
TABLE = "--table" in _sys.argv
SHAPES = [(10, 10, 100), (10, 10, 10), (10, 10, 2), (10, 10, 1), (16, 16, 10), (4, 4, 10), (12, 7, 10), (2, 2, 100)] if TABLE else \
    [(10, 10, 100), (16, 16, 10), (20, 20, 10), (32, 32, 10), (10, 64, 10), (48, 48, 10), (64, 64, 10), (64, 64, 2), (10, 10, 2)] if NARROW else \
    [(128, 128, 10), (100, 100, 10), (128, 64, 10), (64, 128, 10), (100, 40, 10), (64, 64, 10)]
for (n, nh, k) in SHAPES:
    torch.manual_seed(1)
    rbm = BinaryRBM(n, nh, gpu=True)
    v = (torch.rand(C, n, device=dev) < 0.5).double()
    out = {}
    for mode in ((5, 3) if TABLE else (4, 0) if NARROW else (0, 1)):
        lib.qcb_set_gibbs_mode(mode)
        reps = 2 if mode == 1 else 5
        for _ in range(2):
            rbm.gibbs_steps(k, v, overwrite=True, seed=7)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            rbm.gibbs_steps(k, v, overwrite=True, seed=7)
        e1.record()
        torch.cuda.synchronize()
        out[mode] = e0.elapsed_time(e1) / reps
    lib.qcb_set_gibbs_mode(0)
    a, b = ((5, 3) if TABLE else (4, 0) if NARROW else (0, 1))
    print(f"N={n:4d} nh={nh:4d}  C=2^20 k={k}: mode {a} {out[a]:8.3f} ms = {C * k / out[a] / 1e6:7.2f}e9 chain-sweeps/s "
          f"({C * k * (n + nh) / out[a] / 1e6 / 1e3:6.2f}e12 unit updates/s) | mode {b} {out[b]:9.3f} ms  ({out[b] / out[a]:5.2f} x)", flush=True)
    del v
    torch.cuda.empty_cache()
