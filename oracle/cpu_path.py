"""The reference's CPU execution path for the benchmark baseline — TEST/BENCH INFRASTRUCTURE ONLY.

The reference (PIQuIL/QuCumber v1.3.3) is pure Python over PyTorch and cannot travel to the GPU
box, so `bench.py --impl reference` and the `cpu_baseline` leg time this port instead: the same
ATen CPU operators in the same order as the reference issues them (float64, in-place, using
every host thread torch gives them), i.e. what `gpu=False` QuCumber executes.
Only bench.py and tests/ may import this module.
"""
import torch


def gibbs_steps_cpu(W, b, c, k, v):
    """k block-Gibbs sweeps, in place on v (chains, N) — op sequence of rbm/binary_rbm.py:142-147,
    163-168, 181-182, 196-197, 220-228: matmul(out=), add_, sigmoid_, clamp_, bernoulli(out=)."""
    h = torch.zeros(*v.shape[:-1], W.shape[0], dtype=v.dtype)
    Wt = W.t()
    for _ in range(k):
        torch.matmul(v, Wt, out=h).add_(c).sigmoid_().clamp_(min=0, max=1)
        torch.bernoulli(h, out=h)
        torch.matmul(h, W, out=v).add_(b).sigmoid_().clamp_(min=0, max=1)
        torch.bernoulli(v, out=v)
    return v


def effective_energy_cpu(W, b, c, v):
    """rbm/binary_rbm.py:92-96."""
    return -(torch.matmul(v, b) + torch.nn.functional.softplus(torch.nn.functional.linear(v, W, c)).sum(-1))


def partition_cpu(W, b, c, n):
    """generate_hilbert_space + partition (nn_states/neural_state.py:173-176, rbm/binary_rbm.py:240-241)."""
    import numpy as np

    dim = np.arange(2 ** n)
    space = ((dim[:, None] & (1 << np.arange(n))) > 0)[:, ::-1].astype(int)
    space = torch.tensor(space, dtype=torch.double)
    return (-effective_energy_cpu(W, b, c, space)).logsumexp(0).exp()


def grad_reduce_cpu(W, b, c, v):
    """effective_energy_gradient(reduce=True), rbm/binary_rbm.py:113-120."""
    p = torch.matmul(v, W.t()).add_(c).sigmoid_().clamp_(min=0, max=1)
    return torch.cat([(-torch.matmul(p.t(), v)).reshape(-1), -v.sum(0), -p.sum(0)])
