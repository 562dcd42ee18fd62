"""GPU time of one fused CD-10 epoch at the Tutorial-1 shape (N = nh = 10, 10k rows, pos 100 / neg 200)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qucumber_b200 import _lib  # noqa: E402

dev = torch.device("cuda")
lib = _lib.load()
k = int(sys.argv[1]) if len(sys.argv) > 1 else 10
n = nh = int(sys.argv[2]) if len(sys.argv) > 2 else 10
torch.manual_seed(0)
W = (torch.randn(nh, n, dtype=torch.double) / n ** 0.5).to(dev)
b = torch.zeros(n, dtype=torch.double, device=dev)
c = torch.zeros(nh, dtype=torch.double, device=dev)
pos = (torch.rand(10000, n, device=dev) < 0.5).double()
neg = (torch.rand(20000, n, device=dev) < 0.5).double()


def run():
    _lib.check(lib.qcb_rbm_cd_epoch(_lib.ptr(W), _lib.ptr(b), _lib.ptr(c), n, nh, _lib.ptr(pos), 10000, 100, _lib.ptr(neg), 20000, 200,
                                    100, k, 0.01, 1, 0, None, _lib.stream_ptr(dev)), "cd_epoch")


for mode in (0, 1):  # 0: warp-specialised kernel where it applies (layers <= 16 units), 1: single-role kernel
    lib.qcb_set_cd_epoch_mode(mode)
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    print(f"fused epoch kernel N = nh = {n}, mode {mode}: {e0.elapsed_time(e1) / 10:.3f} ms per epoch of 100 steps (k = {k})")
lib.qcb_set_cd_epoch_mode(0)
