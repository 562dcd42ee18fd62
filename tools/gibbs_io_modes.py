import os, sys, torch
sys.path.insert(0, os.getcwd())
from qucumber_b200 import BinaryRBM
import qucumber_b200 as qb
dev = torch.device("cuda"); C = 1 << 20
torch.manual_seed(1)
rbm = BinaryRBM(64, 64, gpu=True)
v = (torch.rand(C, 64, device=dev) < 0.5).double()
bits = qb.pack_states(v)
def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
for k in (10, 100):
    print(k, "float64 rows in/out      %.3f ms" % timed(lambda: rbm.gibbs_steps(k, v, overwrite=True, seed=3)))
    print(k, "packed rows in/out       %.3f ms" % timed(lambda: rbm.gibbs_steps_packed(k, bits, overwrite=True, seed=3)))
    print(k, "in-kernel start, packed  %.3f ms" % timed(lambda: rbm.sample_chains(k, C, seed=3, packed=True)))
