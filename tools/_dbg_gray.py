import sys, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from helpers import make_binary_params, load_binary, random_states
from oracle import rbm_oracle as orc
from qucumber_b200 import ComplexWaveFunction, _lib
lib=_lib.load()
n,nh,S,B=24,17,12,3
rng=np.random.default_rng(1000+S)
nn=ComplexWaveFunction(n,nh,gpu=True)
pa=make_binary_params(n,nh,41,bias_scale=0.3); pp=make_binary_params(n,nh,42,bias_scale=0.3)
load_binary(nn.rbm_am,*pa); load_binary(nn.rbm_ph,*pp)
sites=np.sort(rng.choice(n,S,replace=False)); basis=np.array(["Z"]*n); basis[sites]=rng.choice(["X","Y"],S)
samples=random_states(B,n,43); ts=torch.tensor(samples,device='cuda')
new=[g.cpu().numpy() for g in nn.rotated_gradient(basis,ts)]
lib.qcb_set_grad_reduce_mode(3)
old=[g.cpu().numpy() for g in nn.rotated_gradient(basis,ts)]
lib.qcb_set_grad_reduce_mode(0)
ga,gp=orc.complex_rotated_gradient(orc.BinaryRBMOracle(*pa),orc.BinaryRBMOracle(*pp),list(basis),samples)
for name,a,b in (("new-old am",new[0],old[0]),("new-orc am",new[0],ga),("old-orc am",old[0],ga),("new-old ph",new[1],old[1]),("new-orc ph",new[1],gp),("old-orc ph",old[1],gp)):
    d=np.abs(a-b); i=np.argmax(d); print(name,'max abs',d.max(),'at',i,'val',b[i],'max|e|',np.abs(b).max())
print('sites',sites, 'G', len(ga))
