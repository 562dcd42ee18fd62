"""tcgen05 reduced-gradient kernel (mode 0) against the mma.sync / DMMA kernels (modes 3 and 1) and the oracle, then timings
at N = nh = 64 and N = nh = 32, 2^20 rows.  Run on the GPU box under a short `timeout`."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import load_binary, make_binary_params, random_states  # noqa: E402
from oracle.rbm_oracle import BinaryRBMOracle  # noqa: E402
from qucumber_b200 import BinaryRBM, _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda")
ok = True
for (n, nh, rows) in [(64, 64, 16384), (64, 64, 16421), (40, 33, 20000), (33, 64, 16500), (64, 35, 16500), (50, 50, 100_003)]:
    W, b, c = make_binary_params(n, nh, 5, bias_scale=0.3)
    rbm = load_binary(BinaryRBM(n, nh, gpu=True), W, b, c)
    v = random_states(rows, n, 6)
    want = BinaryRBMOracle(W, b, c).effective_energy_gradient(v)
    tv = torch.tensor(v, device=dev)
    res = {}
    for mode in (0, 3, 1):
        lib.qcb_set_grad_reduce_mode(mode)
        g = rbm.effective_energy_gradient(tv).cpu().numpy()
        res[mode] = float(np.max(np.abs(g - want) / (np.abs(want) + 1e-13 * np.abs(want).max())))
    lib.qcb_set_grad_reduce_mode(0)
    print(json.dumps({"shape": [n, nh, rows], "max_elementwise_rel_err": res}))
    ok = ok and all(e < 1e-10 for e in res.values())
if not ok:
    sys.exit(2)
for n in (64, 32):
    W, b, c = make_binary_params(n, n, 7)
    rbm = load_binary(BinaryRBM(n, n, gpu=True), W, b, c)
    v = (torch.rand(1 << 20, n, device=dev) < 0.5).double()
    for mode in (0, 3, 1):
        lib.qcb_set_grad_reduce_mode(mode)
        for _ in range(3):
            rbm.effective_energy_gradient(v)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            rbm.effective_energy_gradient(v)
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({"N": n, "mode": mode, "ms": round(e0.elapsed_time(e1) / 10, 4)}))
    lib.qcb_set_grad_reduce_mode(0)
