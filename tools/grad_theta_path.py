"""Which phase-1 path did the reduced gradient take? Reads the flag word the integer-theta kernel leaves in the
workspace (0 = integer images did the batch, 1 = handed to the FP64 kernel) and times both kinds of input."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qucumber_b200 import BinaryRBM, _lib  # noqa: E402

dev = torch.device("cuda:0")
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
lib = _lib.load()
for nv, nh in ((64, 64), (32, 64), (48, 40)):
    rb = BinaryRBM(nv, nh, gpu=True)
    g = torch.Generator().manual_seed(99)
    with torch.no_grad():
        rb.weights.copy_(torch.randn(nh, nv, dtype=torch.double, generator=g) / np.sqrt(nv))
    vb = (torch.rand(rows, nv, generator=g) < 0.5).double().to(dev)
    vr = torch.rand(rows, nv, generator=g, dtype=torch.double).to(dev)
    need = lib.qcb_workspace_bytes(_lib.OP_GRAD_REDUCE, nv, nh, 0, rows)
    for name, v in (("binary", vb), ("real", vr)):
        rb.effective_energy_gradient(v)
        torch.cuda.synchronize()
        ws = _lib.workspace(dev, need)
        flag = int(ws[need - 256: need - 252].view(torch.int32).item())
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rb.effective_energy_gradient(v)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(f"N={nv} nh={nh} {name}: flag={flag} {best:.3f} ms")
