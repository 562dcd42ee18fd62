import os, sys, torch
sys.path.insert(0, os.getcwd())
from qucumber_b200 import BinaryRBM, _lib
lib=_lib.load(); lib.qcb_set_gibbs_mode(5)
rbm=BinaryRBM(10,10,gpu=True)
v=(torch.rand(1<<20,10,device='cuda')<0.5).double()
for _ in range(2): rbm.gibbs_steps(100, v, overwrite=True, seed=7)
torch.cuda.synchronize(); print("done")
