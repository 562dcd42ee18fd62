"""Two fused epochs of a 64 x 64 PositiveWaveFunction fit (target of an ncu capture of cd_epoch_wide_kernel)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qucumber_b200 import PositiveWaveFunction
torch.manual_seed(0)
n = 64
data = (torch.rand(2000, n) < 0.5).double()
nn = PositiveWaveFunction(n, n, gpu=True)
nn.fit(data, epochs=2, pos_batch_size=100, neg_batch_size=200, k=10, lr=0.01)
torch.cuda.synchronize()
print("done")
