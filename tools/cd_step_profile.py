"""One CD-10 step at the S64 benchmark shape (positive reduction from float64 rows, tcgen05 Gibbs with packed v_k, negative
reduction from the packed rows) and the T4 reduced gradient (state histogram + weighted pass): target of an
`ncu --set full` capture; without ncu it prints the step times."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qucumber_b200 import BinaryRBM, PositiveWaveFunction  # noqa: E402

dev = torch.device("cuda")
torch.manual_seed(1234)
C = 1 << 20
st = PositiveWaveFunction(64, 64, gpu=True)
with torch.no_grad():
    st.rbm_am.weights.copy_(torch.randn(64, 64, dtype=torch.double) / 8)
    st.rbm_am.visible_bias.copy_(0.1 * torch.randn(64, dtype=torch.double))
    st.rbm_am.hidden_bias.copy_(0.1 * torch.randn(64, dtype=torch.double))
big = (torch.rand(C, 64, device=dev) < 0.5).double()
small = BinaryRBM(10, 10, gpu=True)
v10 = (torch.rand(C, 10, device=dev) < 0.5).double()


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print("cd10 step S64: %.3f ms" % timed(lambda: st.compute_batch_gradients(10, big, big)))
print("reduced gradient T4: %.4f ms" % timed(lambda: small.effective_energy_gradient(v10), 20))
print("reduced gradient S64 (float64 rows): %.3f ms" % timed(lambda: st.rbm_am.effective_energy_gradient(big)))
