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
# ---- the same things through the public API (round 2): fit(), normalization(), fidelity / KL, exact gradients ----
from qucumber_b200 import ComplexWaveFunction  # noqa: E402
from qucumber_b200.utils import training_statistics as ts  # noqa: E402


def flat(state):
    return torch.cat([p.detach().reshape(-1) for net in state.networks for p in getattr(state, net).parameters()])


def fitted(cls, distributed, data, bases=None, **kw):
    torch.manual_seed(11)
    st = cls(data.shape[1], gpu=True)
    st.distributed = distributed
    torch.manual_seed(12)
    st.fit(data, input_bases=bases, **kw)
    return flat(st)


gd = torch.Generator().manual_seed(5)
data = (torch.rand(40000, 20, generator=gd) < 0.3).double()
a = fitted(PositiveWaveFunction, True, data, epochs=2, pos_batch_size=10000, neg_batch_size=12000, k=3, lr=0.05)
b = fitted(PositiveWaveFunction, False, data, epochs=2, pos_batch_size=10000, neg_batch_size=12000, k=3, lr=0.05)
fit_err = float((a - b).abs().max() / b.abs().max())
assert fit_err <= 1e-10, fit_err
gathered = [torch.empty_like(a) for _ in range(world)]
dist.all_gather(gathered, a)
assert all(torch.equal(gathered[0], x) for x in gathered), "ranks ended fit() with different parameters"
# rotated-basis training: the batch rows (with their bases) are sharded, both gradient vectors all-reduced
nq = 6
dq = (torch.rand(512, nq, generator=gd) < 0.5).double()
bq = np.array([list("XZYZZZ") if i % 3 == 0 else (list("ZZZZXY") if i % 3 == 1 else list("ZZZZZZ")) for i in range(512)])
ac = fitted(ComplexWaveFunction, True, dq, bq, epochs=2, pos_batch_size=128, neg_batch_size=128, k=2, lr=0.05)
bc = fitted(ComplexWaveFunction, False, dq, bq, epochs=2, pos_batch_size=128, neg_batch_size=128, k=2, lr=0.05)
cfit_err = float((ac - bc).abs().max() / bc.abs().max())
assert cfit_err <= 1e-10, cfit_err
# normalization / fidelity / KL / exact gradients with the index range split over the ranks
st = PositiveWaveFunction(n, n, gpu=True)
st.rbm_am.load_state_dict(rbm.state_dict())
tg = torch.Generator(device="cuda").manual_seed(9)
target = torch.zeros(2, 1 << n, dtype=torch.double, device=dev)
target[0].copy_(torch.rand(1 << n, dtype=torch.double, device=dev, generator=tg))
target /= target[0].norm()
vals = {}
for mode in (True, False):
    st.distributed = mode
    vals[mode] = (float(st.normalization()), ts.fidelity(st, target), ts.KL(st, target),
                  st.compute_exact_gradients(pos[:64])[0])
for i in range(3):
    assert abs(vals[True][i] - vals[False][i]) <= 1e-11 * abs(vals[False][i]), (i, vals[True][i], vals[False][i])
eg_err = float((vals[True][3] - vals[False][3]).abs().max() / vals[False][3].abs().max())
assert eg_err <= 1e-10, eg_err
if rank == 0:
    print(f"dist_check ok on {world} GPUs: sharded logZ {sharded:.12f}, gibbs shards identical, CD gradient rel err {err:.2e}, "
          f"fit() sharded vs replica {fit_err:.1e} (positive) {cfit_err:.1e} (complex, bases), Z / fidelity / KL equal, "
          f"exact gradient {eg_err:.1e}")
dist.destroy_process_group()
