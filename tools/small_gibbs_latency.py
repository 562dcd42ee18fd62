"""GPU time of one small Gibbs call (N = nh = 10, 200 chains, k = 10): the per-step cost inside fit()."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qucumber_b200 import BinaryRBM  # noqa: E402

rbm = BinaryRBM(10, 10, gpu=True)
v = (torch.rand(200, 10, device="cuda") < 0.5).double()
for _ in range(50):
    rbm.gibbs_steps(10, v, overwrite=True, seed=1)
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    rbm.gibbs_steps(10, v, overwrite=True, seed=1)
    with torch.cuda.graph(g, stream=s):
        for _ in range(100):
            rbm.gibbs_steps(10, v, overwrite=True, seed=1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    g.replay()
e1.record()
torch.cuda.synchronize()
print(f"graph-replayed small gibbs call: {e0.elapsed_time(e1) / 1000 * 1e3:.2f} us per call (pack + sweep kernels)")
