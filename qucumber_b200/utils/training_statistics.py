"""Exact-enumeration metrics (reference: qucumber/utils/training_statistics.py): fidelity, KL and NLL.

For a PositiveWaveFunction measured in the computational basis, fidelity and KL are ONE
enumeration pass on the GPU (qcb_rbm_exact_stats_range): basis states are generated from their
index, exp(-E) is normalised by the log-partition function from a first pass, and the overlap
and both KL sums come back as five numbers. The other cases (bases, density matrices) follow the
reference's structure on top of the CUDA-backed state classes.
"""
import numpy as np
import torch
from torch.distributions.utils import probs_to_logits

from .. import _lib
from ..hilbert import HilbertSpace
from . import cplx
from .unitaries import rotate_psi, rotate_psi_inner_prod, rotate_rho_probs


def _is_full_space(nn_state, space):
    """True when ``space`` is the canonical enumeration (row r = binary digits of r, MSB first): None, a HilbertSpace handle,
    or an explicit tensor that EQUALS it — a caller-supplied space in another row order (or with repeated rows) must be
    evaluated row by row like the reference does, because the target is indexed in space order."""
    if space is None or isinstance(space, HilbertSpace):
        return True
    n = nn_state.num_visible
    if space.dim() != 2 or space.shape[0] != (1 << n) or space.shape[1] != n or n > 24:
        return False
    canonical = HilbertSpace(n, space.device).tensor()
    return bool(torch.equal(space.to(torch.double), canonical))


def _positive_stats(nn_state, target):
    """(logZ, out8) for a PositiveWaveFunction over the full space; target is (2, 2^N) or None."""
    from ..nn_states import PositiveWaveFunction

    from .. import distributed as qd

    assert isinstance(nn_state, PositiveWaveFunction)
    rbm = nn_state.rbm_am
    lib = _lib.load()
    # Under torch.distributed the index range is split over the ranks: per-rank (max, sum) pairs are
    # all-gathered and merged into log Z, the five partial sums of the second pass are plain sums -> one all-reduce.
    rank, world_size = nn_state._shards_space()
    first, count = qd.shard_range(1 << rbm.num_visible, rank, world_size)
    out2 = rbm.log_partition(first, count)
    logZ = float(qd.combine_logz(out2) if world_size > 1 else out2[0] + out2[1].log())
    args, keep = rbm._param_args()
    t_re = t_im = None
    if target is not None:
        target = target.to(device=nn_state.device, dtype=torch.double)
        t_re, t_im = target[0][first:first + count].contiguous(), target[1][first:first + count].contiguous()
    out8 = torch.empty(8, dtype=torch.double, device=nn_state.device)
    ws = _lib.workspace(nn_state.device, lib.qcb_workspace_bytes(_lib.OP_EXACT_STATS, rbm.num_visible, rbm.num_hidden, 0, count))
    _lib.check(lib.qcb_rbm_exact_stats_range(*args, first, count, logZ, _lib.ptr(t_re), _lib.ptr(t_im), _lib.ptr(out8), None,
                                             _lib.ptr(ws), ws.numel(), _lib.stream_ptr(nn_state.device)),
               "qcb_rbm_exact_stats_range")
    if world_size > 1:
        qd.allreduce_sum_(out8)
    return logZ, out8


def fidelity(nn_state, target, space=None, **kwargs):
    """|<target|psi>|^2 for pure states; (tr sqrt(sqrt(rho) sigma sqrt(rho)))^2 for mixed states
    (training_statistics.py:27-76)."""
    from ..nn_states import PositiveWaveFunction, WaveFunctionBase

    if "target_psi" in kwargs or "target_rho" in kwargs:
        target = kwargs.get("target_psi", kwargs.get("target_rho"))
    target = target.to(nn_state.device)
    if isinstance(nn_state, PositiveWaveFunction) and _is_full_space(nn_state, space):
        assert target.dim() == 2, "target must be a complex vector!"
        _, out8 = _positive_stats(nn_state, target)
        o = out8.cpu()
        return float(o[0] ** 2 + o[1] ** 2)
    space = space if space is not None else nn_state.generate_hilbert_space()
    if isinstance(space, HilbertSpace):
        space = space.tensor(device=nn_state.device)
    Z = nn_state.normalization(space)
    if isinstance(nn_state, WaveFunctionBase):
        assert target.dim() == 2, "target must be a complex vector!"
        psi = nn_state.psi(space) / Z.sqrt()
        F = cplx.inner_prod(target, psi)
        return cplx.absolute_value(F).pow_(2).item()
    assert target.dim() == 3, "target must be a complex matrix!"
    rho = nn_state.rho(space, space) / Z
    return _mixed_state_fidelity(target.to(torch.double), rho)


def _mixed_state_fidelity(target, rho):
    """(tr sqrt(sqrt(sigma) rho sqrt(sigma)))^2 of two density matrices given as (2, D, D) re/im planes, on the device that
    holds them. The reference takes the eigenvalues of the NON-Hermitian product sigma rho on the host
    (training_statistics.py:63-76); the same spectrum belongs to the Hermitian matrix sqrt(sigma) rho sqrt(sigma), which two
    symmetric eigensolves (cuSOLVER through torch.linalg.eigh) give without leaving the GPU and without complex round-off in
    the eigenvalues. Negative eigenvalues are round-off and enter by absolute value, as in the reference; eigenvalues of
    round-off size are dropped before the square root (the reference keeps them, which costs it ~1e-8 absolute on
    rank-deficient targets: sqrt(1e-17) per null direction)."""
    sigma = torch.complex(target[0], target[1])
    r = torch.complex(rho[0], rho[1])
    dim = sigma.shape[-1]

    def spectrum(w):  # |w|, with round-off-sized entries (below dim * eps * largest) set to zero before the square root
        w = w.abs()
        return torch.where(w < w.max() * (dim * torch.finfo(torch.double).eps), torch.zeros_like(w), w)

    w, vec = torch.linalg.eigh(0.5 * (sigma + sigma.mH))
    root = (vec * spectrum(w).sqrt().to(vec.dtype)) @ vec.mH
    m = root @ r @ root
    ev = spectrum(torch.linalg.eigvalsh(0.5 * (m + m.mH)))
    return float(ev.sqrt().sum() ** 2)


def NLL(nn_state, samples, space=None, sample_bases=None, **kwargs):
    """Negative log-likelihood of the samples (training_statistics.py:79-128)."""
    from ..nn_states import WaveFunctionBase

    samples = samples.to(device=nn_state.device, dtype=torch.double)
    Z = nn_state.normalization(space if not isinstance(space, HilbertSpace) else None)
    if sample_bases is None:
        return -torch.mean(probs_to_logits(nn_state.probability(samples, Z))).item()
    total = 0.0
    unique_bases, indices = np.unique(sample_bases, axis=0, return_inverse=True)
    indices = torch.as_tensor(np.asarray(indices).reshape(-1), device=samples.device)
    for i in range(unique_bases.shape[0]):
        basis = unique_bases[i, :]
        sub = samples[indices == i, :]
        if np.any(basis != "Z"):
            if isinstance(nn_state, WaveFunctionBase):
                nn_probs = (cplx.absolute_value(rotate_psi_inner_prod(nn_state, basis, sub)) ** 2) / Z
            else:
                nn_probs = rotate_rho_probs(nn_state, basis, sub) / Z
        else:
            nn_probs = nn_state.probability(sub, Z)
        total -= torch.sum(probs_to_logits(nn_probs))
    return (total / float(len(samples))).item()


def _single_basis_KL(target_probs, nn_probs):
    if target_probs.is_cuda and nn_probs.is_cuda:
        return _kl_terms(target_probs, None, nn_probs, None, 0.0)
    return torch.sum(target_probs * probs_to_logits(target_probs)) - torch.sum(target_probs * probs_to_logits(nn_probs))


def _kl_terms(t_re, t_im, p_re, p_im, logZ):
    """sum p_t log clamp(p_t) - sum p_t log clamp(p) in one reduction kernel (qcb_kl_terms); p_t and p are
    given as amplitudes (re, im) or, with im = None, as probabilities; p is scaled by exp(-logZ)."""
    lib = _lib.load()
    dev = t_re.device
    planes = [None if x is None else x.to(device=dev, dtype=torch.double).contiguous() for x in (t_re, t_im, p_re, p_im)]
    n = planes[0].numel()
    if planes[2].numel() != n:
        raise ValueError("target and model distributions differ in size")
    out2 = torch.empty(2, dtype=torch.double, device=dev)
    ws = _lib.workspace(dev, lib.qcb_workspace_bytes(_lib.OP_KL_TERMS, 0, 0, 0, n))
    _lib.check(lib.qcb_kl_terms(*[_lib.ptr(x) for x in planes], n, float(logZ), _lib.ptr(out2), _lib.ptr(ws), ws.numel(),
                                _lib.stream_ptr(dev)), "qcb_kl_terms")
    return out2[0] - out2[1]


def KL(nn_state, target, space=None, bases=None, **kwargs):
    """KL divergence between target and model distributions, averaged over the given bases
    (training_statistics.py:138-218)."""
    from ..nn_states import PositiveWaveFunction, WaveFunctionBase

    if "target_psi" in kwargs or "target_rho" in kwargs:
        target = kwargs.get("target_psi", kwargs.get("target_rho"))
    if bases is None and not isinstance(target, dict) and isinstance(nn_state, PositiveWaveFunction) and _is_full_space(nn_state, space):
        _, out8 = _positive_stats(nn_state, target)
        o = out8.cpu()
        return float(o[2] - o[3])
    space = space if space is not None else nn_state.generate_hilbert_space()
    if isinstance(space, HilbertSpace):
        space = space.tensor(device=nn_state.device)
    Z = nn_state.normalization(space)
    if isinstance(target, dict):
        target = {k: v.to(nn_state.device) for k, v in target.items()}
        if bases is None:
            bases = list(target.keys())
        else:
            assert set(bases) == set(target.keys()), "Given bases must match the keys of the target_psi dictionary."
    else:
        target = target.to(nn_state.device)
    kl = 0.0
    if bases is None:
        kl += _single_basis_KL(cplx.absolute_value(target) ** 2, nn_state.probability(space, Z))
    elif isinstance(nn_state, WaveFunctionBase):
        for basis in bases:
            if isinstance(target, dict):
                target_psi_r = target[basis]
            else:
                target_psi_r = rotate_psi(nn_state, basis, space, psi=target)
            assert target_psi_r.dim() == 2, "target must be a complex vector!"
            psi_r = rotate_psi(nn_state, basis, space)
            kl += _kl_terms(target_psi_r[0], target_psi_r[1], psi_r[0], psi_r[1], float(torch.log(torch.as_tensor(Z))))
        kl /= float(len(bases))
    else:
        for basis in bases:
            if isinstance(target, dict):
                target_rho_r = target[basis]
                assert target_rho_r.dim() == 3, "target must be a complex matrix!"
                target_probs_r = torch.diagonal(cplx.real(target_rho_r))
            else:
                assert target.dim() == 3, "target must be a complex matrix!"
                target_probs_r = rotate_rho_probs(nn_state, basis, space, rho=target)
            kl += _single_basis_KL(target_probs_r, rotate_rho_probs(nn_state, basis, space) / Z)
        kl /= float(len(bases))
    return kl.item()
