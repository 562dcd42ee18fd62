"""Multi-GPU plumbing: one process per GPU under ``torchrun``, NCCL through ``torch.distributed``.

The hot path shards over independent units, so almost nothing is exchanged:

* Gibbs sampling / generation — chains are independent; rank r owns a contiguous block of chains
  and keys its Philox draws by the *global* chain id, so 1/2/4/8-GPU runs produce the same
  samples. No collective.
* CD-k training — each rank reduces the gradient of its share of the positive batch and of its
  chains (already scaled by the *global* batch sizes); one sum-allreduce of ``num_pars`` doubles
  per step, then every rank applies the identical optimizer step to its replicated parameters.
* Exact enumeration — rank r owns the index range [r 2^N / R, (r+1) 2^N / R); the per-rank
  (max, sum exp) pairs are all-gathered (2 doubles per rank) and merged.

Everything here is host logic over ``torch.distributed`` and also runs on the gloo backend (CPU)
for tests; the per-rank numbers come from the CUDA kernels.
"""
import math

import torch
import torch.distributed as dist


def is_distributed():
    return dist.is_available() and dist.is_initialized()


def world():
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if is_distributed():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(total, rank=None, world_size=None):
    """Contiguous, balanced shard [first, first + count) of ``total`` units for ``rank``."""
    if rank is None or world_size is None:
        rank, world_size = world()
    base, rem = divmod(int(total), int(world_size))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def allreduce_sum_(tensors):
    """In-place sum over ranks of a tensor or list of tensors (no-op for a single process)."""
    if not is_distributed() or dist.get_world_size() == 1:
        return tensors
    seq = [tensors] if isinstance(tensors, torch.Tensor) else list(tensors)
    if len(seq) == 1:
        dist.all_reduce(seq[0], op=dist.ReduceOp.SUM)
    else:  # one collective for all of them (latency-bound: 33 KB at N = nh = 64)
        flat = torch.cat([t.reshape(-1) for t in seq])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        torch._foreach_copy_(seq, [flat[o:o + t.numel()].view_as(t) for o, t in zip(_offsets(seq), seq)])
    return tensors


def _offsets(seq):
    out, pos = [], 0
    for t in seq:
        out.append(pos)
        pos += t.numel()
    return out


def merge_lse_pairs(pairs):
    """Merge (m_r, s_r) pairs (rows of a (R, 2) tensor) with sum_r s_r e^{m_r} = s e^{m}; returns log of it as a 0-d
    tensor on the pairs' device. Pure tensor arithmetic: nothing here reads a value back to the host (an empty range,
    s_r = 0, drops out through the masks; all ranges empty gives -inf)."""
    m = pairs[:, 0]
    s = pairs[:, 1]
    live = s > 0
    neg_inf = torch.full_like(m, -math.inf)
    M = torch.where(live, m, neg_inf).max()
    total = torch.where(live, s * torch.exp(torch.where(live, m - M, torch.zeros_like(m))), torch.zeros_like(s)).sum()
    return torch.where(total > 0, M + total.log(), torch.full_like(total, -math.inf))


def combine_logz(out2):
    """log Z from this rank's (max, sum exp(x - max)) pair: ONE all-gather of 2 doubles per rank into a (R, 2) tensor,
    merged on the device."""
    if not is_distributed() or dist.get_world_size() == 1:
        return merge_lse_pairs(out2.reshape(1, 2))
    gathered = torch.empty(2 * dist.get_world_size(), dtype=out2.dtype, device=out2.device)
    dist.all_gather_into_tensor(gathered, out2.contiguous().reshape(2))
    return merge_lse_pairs(gathered.view(-1, 2))


def broadcast_seed(seed, device):
    """Rank 0's 62-bit seed on every rank (identity for a single process)."""
    if not is_distributed() or dist.get_world_size() == 1:
        return int(seed)
    on_gpu = dist.get_backend() == "nccl"
    t = torch.tensor([int(seed)], dtype=torch.int64, device=device if on_gpu else "cpu")
    dist.broadcast(t, src=0)
    return int(t.item())


def sharded_log_partition(rbm):
    """log Z with the Hilbert space split across ranks by index range."""
    first, count = shard_range(1 << rbm.num_visible)
    return combine_logz(rbm.log_partition(first, count))


def sharded_gibbs(rbm, k, local_chains, seed, total_chains=None, sweep0=0, overwrite=True):
    """k sweeps on this rank's block of chains; draws are keyed by global chain id, so the
    concatenation over ranks equals a single-GPU run with the same seed."""
    rank, world_size = world()
    if total_chains is None:
        first = rank * local_chains.shape[0]
    else:
        first, count = shard_range(total_chains, rank, world_size)
        if count != local_chains.shape[0]:
            raise ValueError(f"rank {rank} holds {local_chains.shape[0]} chains, expected {count}")
    return rbm.gibbs_steps(k, local_chains, overwrite=overwrite, seed=seed, chain_id0=first, sweep0=sweep0)


def sharded_cd_gradients(nn_state, k, pos_local, neg_local, global_pos, global_neg, seed, chain_id0, sweep0=0):
    """CD-k gradient of a PositiveWaveFunction with the batch sharded over ranks: local partial sums
    scaled by the global batch sizes, then one allreduce. Every rank returns the same vector."""
    from . import _lib

    rbm = nn_state.rbm_am
    lib = _lib.load()
    dev = nn_state.device
    g = torch.zeros(rbm.num_pars, dtype=torch.double, device=dev)
    args, keep = rbm._param_args()
    na = getattr(rbm, "num_aux", 0)
    pos = pos_local.to(device=dev, dtype=torch.double).contiguous()
    ws = _lib.workspace(dev, max(lib.qcb_workspace_bytes(_lib.OP_GRAD_REDUCE, rbm.num_visible, rbm.num_hidden, na, pos.shape[0]),
                                 lib.qcb_workspace_bytes(_lib.OP_CDK_NEG, rbm.num_visible, rbm.num_hidden, na, neg_local.shape[0])))
    st = _lib.stream_ptr(dev)
    _lib.check(lib.qcb_rbm_grad_reduce(*args, _lib.ptr(pos), pos.shape[0], 1.0 / float(global_pos), _lib.ptr(g), 1, _lib.ptr(ws),
                                       ws.numel(), st), "qcb_rbm_grad_reduce")
    vk = neg_local.to(device=dev, dtype=torch.double).clone().contiguous()
    _lib.check(lib.qcb_rbm_cdk_neg(*args, _lib.ptr(vk), vk.shape[0], int(k), int(seed), int(chain_id0), int(sweep0), None,
                                   -1.0 / float(global_neg), _lib.ptr(g), _lib.ptr(ws), ws.numel(), st), "qcb_rbm_cdk_neg")
    allreduce_sum_(g)
    return g
