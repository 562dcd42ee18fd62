"""2-GPU check of qucumber_b200.distributed on real hardware (run under torchrun)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qucumber_b200 import BinaryRBM, PositiveWaveFunction  # noqa: E402
from qucumber_b200 import distributed as qd  # noqa: E402

local = int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
rank, world = qd.world()
torch.manual_seed(3)
n = 24
rbm = BinaryRBM(n, n, gpu=True)
with torch.no_grad():
    g = torch.Generator().manual_seed(1)
    rbm.weights.copy_(torch.randn(n, n, dtype=torch.double, generator=g) / np.sqrt(n))
full = rbm.log_partition()
full = float(full[0] + full[1].log())
sharded = float(qd.sharded_log_partition(rbm))
assert abs(full - sharded) <= 1e-12 * abs(full), (full, sharded)
# chains sharded by global id == unsharded
C = 4096
g = torch.Generator().manual_seed(2)
v0 = (torch.rand(C, n, generator=g) < 0.5).double().to(dev)
ref = rbm.gibbs_steps(5, v0, seed=99)
first, count = qd.shard_range(C)
mine = qd.sharded_gibbs(rbm, 5, v0[first:first + count].clone(), seed=99, total_chains=C)
assert torch.equal(mine, ref[first:first + count])
# CD gradient: sharded batch + allreduce == single-GPU gradient with the same draws
nn_state = PositiveWaveFunction(n, n, gpu=True)
nn_state.rbm_am.load_state_dict(rbm.state_dict())
pos = (torch.rand(2048, n, generator=g) < 0.5).double().to(dev)
pf, pc = qd.shard_range(2048)
gs = qd.sharded_cd_gradients(nn_state, 3, pos[pf:pf + pc], v0[first:first + count], 2048, C, seed=7, chain_id0=first)
from qucumber_b200 import _lib  # noqa: E402
lib = _lib.load()
args, keep = rbm._param_args()
gref = torch.empty(rbm.num_pars, dtype=torch.double, device=dev)
ws = _lib.workspace(dev, _lib.workspace_bytes(_lib.OP_CD_STEP, n, n, 0, C))
_lib.check(lib.qcb_rbm_cd_step(*args, _lib.ptr(pos), 2048, _lib.ptr(v0), C, 3, 7, 0, 0, 1.0 / 2048, -1.0 / C, _lib.ptr(gref), _lib.ptr(ws),
                               ws.numel(), _lib.stream_ptr(dev)), "cd_step")
err = float((gs - gref).abs().max() / gref.abs().max())
assert err <= 1e-12, err
if rank == 0:
    print(f"dist_check ok on {world} GPUs: sharded logZ {sharded:.12f}, gibbs shards identical, CD gradient rel err {err:.2e}")
dist.destroy_process_group()
