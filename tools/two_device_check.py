"""One process, two GPUs: the same calls on cuda:0 and cuda:1 (current device left at 0) must agree bit for bit."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qucumber_b200 import BinaryRBM, PositiveWaveFunction  # noqa: E402

assert torch.cuda.device_count() >= 2
torch.cuda.set_device(0)
res = []
for dev in ("cuda:0", "cuda:1"):
    torch.manual_seed(3)
    rbm = BinaryRBM(64, 64, gpu=True).to(dev) if hasattr(BinaryRBM, "to") else BinaryRBM(64, 64, gpu=True)
    rbm = rbm.to(dev)
    rbm.device = torch.device(dev)
    g = torch.Generator().manual_seed(5)
    v = (torch.rand(1 << 14, 64, generator=g) < 0.5).double().to(dev)
    out = [rbm.gibbs_steps(5, v, seed=11).cpu(), rbm.effective_energy(v).cpu(), rbm.effective_energy_gradient(v).cpu()]
    small = BinaryRBM(20, 20, gpu=True).to(dev)
    small.device = torch.device(dev)
    o2 = small.log_partition().cpu()
    out.append(o2)
    torch.manual_seed(9)
    nn = PositiveWaveFunction(10, 10, gpu=True)
    nn.rbm_am = nn.rbm_am.to(dev)
    nn.rbm_am.device = torch.device(dev)
    nn.device = torch.device(dev)
    data = (torch.rand(1000, 10, generator=g) < 0.5).double()
    nn.fit(data, epochs=2, pos_batch_size=100, neg_batch_size=200, k=3, lr=0.05)
    out.append(torch.cat([p.detach().flatten().cpu() for p in nn.rbm_am.parameters()]))
    assert torch.cuda.current_device() == 0
    res.append(out)
names = ["gibbs", "energy", "gradient", "log_partition", "fit parameters"]
bad = [n for n, a, b in zip(names, *res) if not torch.equal(a, b)]
for n, a, b in zip(names, *res):
    print(n, "equal" if torch.equal(a, b) else f"DIFFERENT max |d| = {(a - b).abs().max().item():.3e}")
assert not bad, bad
print("two-device check ok:", len(res[0]), "results identical on cuda:0 and cuda:1")
